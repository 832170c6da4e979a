"""GPU: the row-sharded multi-GPU path.  World size 1 in-process, and world size 2 as two processes (NCCL when two
devices are visible, otherwise both ranks share cuda:0 and exchange through gloo) -- the sharded result must be
bit-identical to the single-GPU symmetric result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import random_walks

pytestmark = pytest.mark.gpu


def _kernel():
    from gpsig_b200 import kernels
    return kernels.SignatureRBF(64 * 4, 4, 4, lengthscales=1.3)


def test_sharded_world_size_1_matches_single():
    from gpsig_b200 import parallel
    X = random_walks(150, 64, 4, 11).reshape(150, -1)
    k = _kernel()
    assert torch.equal(parallel.sharded_K_symm(k, X), k.K(X))
    k.normalization = False
    assert torch.equal(parallel.sharded_K_symm(k, X), k.K(X))
    from gpsig_b200 import kernels
    kl = kernels.SignatureLinear(64 * 4, 4, 4, lengthscales=1.3)
    assert torch.equal(parallel.sharded_K_symm(kl, X), kl.K(X))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    ndev = torch.cuda.device_count()
    backend = "nccl" if ndev >= ws else "gloo"
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group(backend, rank=rank, world_size=ws)
    from gpsig_b200 import parallel
    X = random_walks(150, 64, 4, 11).reshape(150, -1)
    k = _kernel()
    K = parallel.sharded_K_symm(k, X)
    ref = k.K(X)
    ok = bool(torch.equal(K, ref))
    from gpsig_b200 import kernels
    kl = kernels.SignatureLinear(64 * 4, 4, 4, lengthscales=1.3)      # Linear takes the warp-fused kernel (row ranges per rank)
    ok = ok and bool(torch.equal(parallel.sharded_K_symm(kl, X), kl.K(X)))
    # rectangular block, Kuf column shards and the data-parallel ELBO
    Y2 = random_walks(37, 64, 4, 12).reshape(37, -1)
    # no call-level centre in the RBF arithmetic: shards are bit-equal to the single call
    close = lambda a, b: bool(torch.equal(a, b))  # noqa: E731
    ok = ok and close(parallel.sharded_K(k, X, Y2), k.K(X, Y2))
    rng = np.random.default_rng(5)
    Z = 0.4 * rng.standard_normal((10, 9, 2, 4))
    ok = ok and close(parallel.sharded_K_tens_vs_seq(k, Z, X, increments=True), k.K_tens_vs_seq(Z, X, increments=True))
    from gpsig_b200 import models, inducing_variables as iv
    Yl = (rng.standard_normal((150, 1)) > 0).astype(np.float64)
    m = models.SVGP(X, Yl, k, models.Bernoulli(), iv.InducingTensors(Z, 4, increments=True), num_latent=1,
                    q_mu=0.2 * rng.standard_normal((9, 1)))
    e1, e2 = float(parallel.sharded_elbo(m).item()), m.compute_log_likelihood()
    ok = ok and abs(e1 - e2) < 1e-4 * abs(e2)
    q.put((rank, ok, backend))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_world_size_2_matches_single():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res
