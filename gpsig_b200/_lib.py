"""
ctypes binding of libgpsig_b200.so -- the only way the Python host reaches the device code.

There is NO CPU fallback: if the library is missing it is built with nvcc (gpsig_b200/_build.py); if that fails, or a
call is made without a CUDA device, an exception is raised.  The oracle under oracle/ is never imported from here.
"""
import ctypes
import os

from . import _build

_c_float_p = ctypes.c_void_p  # device pointers travel as integers (torch .data_ptr())

_PROTOTYPES = {
    "gpsig_version": (ctypes.c_int, []),
    "gpsig_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "gpsig_last_error_detail": (ctypes.c_char_p, []),
    "gpsig_launch_count": (ctypes.c_longlong, []),
    "gpsig_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "gpsig_profile_reset": (ctypes.c_int, []),
    "gpsig_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong),
                                          ctypes.POINTER(ctypes.c_double)]),
    "gpsig_scale_features": (ctypes.c_int, [_c_float_p, ctypes.c_long, ctypes.c_int, _c_float_p, ctypes.c_int, _c_float_p,
                                            ctypes.c_void_p]),
    "gpsig_add_lags": (ctypes.c_int, [_c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_int, _c_float_p,
                                      ctypes.c_void_p]),
    "gpsig_gram": (ctypes.c_int, [ctypes.c_int, _c_float_p, ctypes.c_long, _c_float_p, ctypes.c_long, ctypes.c_int,
                                  ctypes.c_void_p, _c_float_p, ctypes.c_long, ctypes.c_void_p]),
    "gpsig_sigkern_levels": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                            ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "gpsig_seq_kern_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_size_t]),
    "gpsig_seq_kern_levels": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_int, _c_float_p,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_long, ctypes.c_long,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "gpsig_seq_kern_levels_blocks": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                    _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                    ctypes.c_int, _c_float_p, ctypes.c_long, ctypes.c_void_p,
                                                    ctypes.c_size_t, ctypes.c_void_p]),
    "gpsig_seq_kern_diag_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t]),
    "gpsig_set_knob": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "gpsig_mirror_upper": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "gpsig_assemble_symmetric": (ctypes.c_int, [_c_float_p, ctypes.c_void_p, ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "gpsig_seq_kern_diag_levels": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                  ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "gpsig_normalize_weight_sum": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, _c_float_p, _c_float_p,
                                                  ctypes.c_void_p, ctypes.c_float, ctypes.c_int, _c_float_p, _c_float_p,
                                                  _c_float_p, ctypes.c_void_p]),
    "gpsig_tensor_kern_levels": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, _c_float_p,
                                                ctypes.c_void_p]),
    "gpsig_tens_vs_seq_levels": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "gpsig_tens_seq_kern_levels": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, _c_float_p, ctypes.c_long, ctypes.c_int,
                                                  _c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, _c_float_p,
                                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_float_p, ctypes.c_void_p]),
    "gpsig_sigkern_levels_vjp": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long,
                                                ctypes.c_long, ctypes.c_long, ctypes.c_int, _c_float_p, _c_float_p,
                                                ctypes.c_void_p]),
    "gpsig_tens_vs_seq_levels_vjp": (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                                    _c_float_p, _c_float_p, ctypes.c_void_p]),
    "gpsig_lr_hadamard_csc": (ctypes.c_int, [_c_float_p, ctypes.c_long, ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, _c_float_p, ctypes.c_int, ctypes.c_float,
                                             _c_float_p, ctypes.c_void_p]),
    "gpsig_lr_seq_level": (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_float_p, ctypes.c_int,
                                          ctypes.c_float, _c_float_p, _c_float_p, ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)
ABI_VERSION = 200  # include/gpsig_b200.h GPSIG_B200_VERSION

_lib = None


class GPSigError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def library_path():
    return _build.LIBPATH


def load():
    """Load (building first if needed) the shared library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    # build() is digest-cached (stamp file over sources + header + flags): a stale library is rebuilt, a current one costs
    # one hash of the sources.  Without nvcc (a box that only received the prebuilt .so) the existing library is used.
    path = _build.LIBPATH
    try:
        path = _build.build(force=bool(os.environ.get("GPSIG_B200_REBUILD")))
    except Exception:
        if not os.path.exists(path):
            raise
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.gpsig_version() != ABI_VERSION:
        raise GPSigError("libgpsig_b200.so reports ABI version %d, the Python host expects %d: rebuild (GPSIG_B200_REBUILD=1)"
                         % (lib.gpsig_version(), ABI_VERSION))
    _lib = lib
    return lib


def set_knob(name, value):
    """gpsig_set_knob: runtime switch of a tuning knob (bench.py's `pipeline` pass, experiments)."""
    check(load().gpsig_set_knob(name.encode(), int(value)), "gpsig_set_knob")


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.gpsig_error_string(rc).decode()
        detail = lib.gpsig_last_error_detail().decode() if rc < 0 else ""
        raise GPSigError("%s failed: %s (%d)%s" % (what, msg, rc, (": " + detail) if detail else ""))
