"""CPU-side checks of the drop-in boundary: the library builds, loads, exports every symbol include/*.h declares, and
the Python host mirrors the reference's argument validation.  No device work."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(gpsig_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_builds_and_exports_every_declared_symbol():
    from gpsig_b200 import _build, _lib
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), "include/gpsig_b200.h declares %s but the library does not export it" % name
    # the ctypes prototype table covers exactly the header
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    lib.gpsig_version.restype = ctypes.c_int
    assert lib.gpsig_version() == _lib.ABI_VERSION
    lib.gpsig_error_string.restype = ctypes.c_char_p
    assert lib.gpsig_error_string(-1) == b"invalid argument"


def test_library_contains_sm100a_tma_code():
    """The recursion kernel must carry TMA tensor loads (UTMALDG) compiled for sm_100a."""
    import shutil
    import subprocess
    from gpsig_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UTMALDG" in out


def test_argument_errors_are_reported_not_crashed():
    from gpsig_b200 import _lib
    lib = _lib.load()
    rc = lib.gpsig_sigkern_levels(0, 1, 1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, None)
    assert rc == -1
    with pytest.raises(_lib.GPSigError):
        _lib.check(rc, "gpsig_sigkern_levels")
    assert lib.gpsig_seq_kern_workspace_bytes(0, 1, 1, 1, 1, 0) == 0
    assert lib.gpsig_seq_kern_workspace_bytes(8, 16, 8, 16, 3, 1 << 20) > 0


def test_kernel_constructor_validation_mirrors_reference():
    from gpsig_b200 import kernels
    with pytest.raises(ValueError):                      # kernels.py:98-101
        kernels.SignatureLinear(10, 3, 2)
    with pytest.raises(NotImplementedError):             # kernels.py:59-60
        kernels.SignatureLinear(12, 3, 4, order=2, low_rank=True)
    with pytest.raises(ValueError):                      # kernels.py:74-75
        kernels.SignatureLinear(12, 3, 2, num_lags=-1)
    with pytest.raises(ValueError):                      # kernels.py:128-133
        kernels.SignatureLinear(12, 3, 2, variances=[1.0, 2.0])
    k = kernels.SignatureRBF(12, 3, 4, order=-1)
    assert k.order == 4 and k.len_examples == 4          # kernels.py:56-57
    assert kernels.SignatureLinear(12, 3, 4, order=7).order == 4
    assert kernels.SignatureGauss is kernels.SignatureRBF
    assert np.allclose(k.variances, np.ones(5)) and k.lengthscales.shape == (3,)


def test_inducing_tensor_shape_asserts():
    from gpsig_b200 import inducing_variables as iv
    Z = np.zeros((6, 5, 3))
    f = iv.InducingTensors(Z, num_levels=3)
    assert len(f) == 5 and f.len_tensors == 6
    with pytest.raises(AssertionError):                  # inducing_variables.py:40
        iv.InducingTensors(Z, num_levels=4)
    with pytest.raises(AssertionError):                  # inducing_variables.py:42-43
        iv.InducingTensors(Z, num_levels=3, increments=True)
    fw = iv.InducingTensors(np.zeros((6, 5, 2, 3)), num_levels=3, increments=True, learn_weights=True)
    assert fw.W.shape == (3, 5, 5)


def test_no_cpu_fallback():
    """Without a CUDA device the product raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from gpsig_b200 import kernels, _lib
    k = kernels.SignatureLinear(12, 3, 2)
    with pytest.raises(_lib.GPSigError):
        k.K(np.zeros((2, 12)))


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "gpsig_b200", "*.py")):
        src = open(path).read()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), path


def test_low_rank_projection_csc_matches_the_dense_definition():
    """Host logic of the low-rank mode (no device): the CSC form handed to gpsig_lr_hadamard_csc encodes exactly
    C = scale * (A (x) B) R with row q of R pairing A[q % k1] and B[q // k1] (low_rank_calculations.py:165-170, :182-193)."""
    from gpsig_b200 import low_rank_calculations as L
    from oracle import gpsig_oracle as O
    rng = np.random.default_rng(0)
    k1, k2, r = 5, 7, 6
    proj = L.draw_projection(k1, k2, r, "sqrt", seed=[3, 1], device="cpu")
    again = L.draw_projection(k1, k2, r, "sqrt", seed=[3, 1], device="cpu")
    assert np.array_equal(proj.dense, again.dense)                       # same seed, same projection
    colptr, ia, ib, val = (t.numpy() for t in (proj.colptr, proj.ia, proj.ib, proj.val))
    assert colptr[0] == 0 and colptr[-1] == len(val) == np.count_nonzero(proj.dense)
    A, B = rng.standard_normal((4, k1)), rng.standard_normal((4, k2))
    C = np.zeros((4, r))
    for c in range(r):
        for e in range(colptr[c], colptr[c + 1]):
            C[:, c] += A[:, ia[e]] * B[:, ib[e]] * val[e]
    C *= proj.scale
    ref = O.lr_hadamard_prod_sparse(A, B, proj.dense, O.sparse_scale(k1 * k2, "sqrt"))
    np.testing.assert_allclose(C, ref, rtol=1e-6, atol=1e-7)
    sub = L.draw_projection(k1, k2, r, "lin", seed=5, device="cpu")       # subsampling: one signed entry per column
    assert np.array_equal(sub.colptr.numpy(), np.arange(r + 1)) and set(np.abs(sub.val.numpy())) == {1.0}
