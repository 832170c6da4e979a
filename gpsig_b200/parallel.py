"""
Multi-GPU covariance: the N x N sequence-pair batch is sharded by ROW BLOCKS over ranks (one process per GPU), X is
replicated, and the assembled covariance rows are exchanged with a single all-gather over NCCL/NVLink.
The reference has no distributed code; this module is new.

Symmetric K(X, X): a rank computes, for each row block it owns, only the tiles j >= i (C ABI: row_begin/row_end of
gpsig_seq_kern_levels), normalises / weights / sums its rows locally (the level diagonals of ALL sequences are
recomputed per rank -- N tiles, <1 % of the work), all-gathers the (rows, N) shards of the final matrix, and mirrors
the lower triangle.  Row blocks are dealt in boustrophedon order (0..W-1, W-1..0, ...) so that every rank gets the same
share of long and short rows of the triangle.  Results are bit-identical to the single-GPU symmetric path.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def row_blocks(n, world_size, blocks_per_rank=8, multiple=8):
    """Block boundaries [(b, e), ...] covering [0, n): about blocks_per_rank blocks per rank, sizes multiple of 8."""
    size = -(-n // (world_size * blocks_per_rank))
    size = max(multiple, -(-size // multiple) * multiple)
    return [(b, min(n, b + size)) for b in range(0, n, size)]


def owner_of_block(k, world_size):
    """Boustrophedon dealing: block k of round r = k // W goes to rank k % W (even rounds) or W-1 - k % W (odd rounds)."""
    r, c = divmod(k, world_size)
    return c if r % 2 == 0 else world_size - 1 - c


def partition(n, world_size, blocks_per_rank=8):
    """[(rank's list of (begin, end))] for every rank."""
    blocks = row_blocks(n, world_size, blocks_per_rank)
    parts = [[] for _ in range(world_size)]
    for k, be in enumerate(blocks):
        parts[owner_of_block(k, world_size)].append(be)
    return parts


def rows_of(blocks):
    if not blocks:
        return np.zeros((0,), dtype=np.int64)
    return np.concatenate([np.arange(b, e, dtype=np.int64) for b, e in blocks])


_INDEX_CACHE = {}


def _shard_index(n, parts, rank, dev):
    """Index tensors of a partition (cached per device): rows of `rank`, their int32 copy (diagonal columns), and for
    every global row its position in the flattened (world * maxr) all-gather result."""
    key = (n, tuple(tuple(p) for p in parts), rank, str(dev))
    hit = _INDEX_CACHE.get(key)
    if hit is None:
        counts = [sum(e - b for b, e in p) for p in parts]
        maxr = max(counts) if counts else 0
        src = np.zeros((n,), dtype=np.int64)
        for r, p in enumerate(parts):
            rr = rows_of(p)
            src[rr] = r * maxr + np.arange(len(rr))
        rows = rows_of(parts[rank])
        hit = (rows, torch.as_tensor(rows, device=dev), torch.as_tensor(rows.astype(np.int32), device=dev),
               torch.as_tensor(src, device=dev), torch.as_tensor(src.astype(np.int32), device=dev), maxr)
        if len(_INDEX_CACHE) > 64:
            _INDEX_CACHE.clear()
        _INDEX_CACHE[key] = hit
    return hit


def _all_gather_padded(padded, world_size, group):
    """ONE collective on equally sized shards: all_gather_into_tensor (NCCL); the gloo backend used by the CPU tests (or
    two ranks sharing one GPU) stages through host memory."""
    if world_size == 1:
        return padded[None]
    if dist.get_backend(group) == "nccl":
        gathered = torch.empty((world_size,) + tuple(padded.shape), device=padded.device, dtype=padded.dtype)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        return gathered
    host = padded.cpu()
    chunks = [torch.empty_like(host) for _ in range(world_size)]
    dist.all_gather(chunks, host, group=group)
    return torch.stack(chunks).to(padded.device)


def gather_rows(local_rows, n, parts, group=None):
    """All-gather row shards (rank r holds the rows rows_of(parts[r]), shape (len, n2)) into the full (n, n2) matrix:
    one collective on shards padded to a common row count, one gather kernel that puts the rows in order."""
    world_size = len(parts)
    rank = dist.get_rank(group) if (dist.is_initialized() and world_size > 1) else 0
    dev, dt, n2 = local_rows.device, local_rows.dtype, local_rows.shape[1]
    _, _, _, src, _, maxr = _shard_index(n, parts, rank, dev)
    if local_rows.shape[0] == maxr:
        padded = local_rows.contiguous()
    else:
        padded = torch.zeros((maxr, n2), device=dev, dtype=dt)
        padded[:local_rows.shape[0]] = local_rows
    gathered = _all_gather_padded(padded, world_size, group)
    return gathered.reshape(world_size * maxr, n2).index_select(0, src)


def sharded_K_symm(kern, X, group=None, blocks_per_rank=8, gather=True):
    """K(X, X) (kernels.py:400-476, X2 is None) with the pair batch sharded over the process group.

    Per rank: ONE fused launch over its list of row blocks (entries j >= i only), one launch for the level diagonals of
    all sequences, one normalise / weight / sum launch that writes straight into the all-gather input, ONE all-gather,
    and one kernel that orders the gathered rows and fills the lower triangle from the transpose.

    Returns the full (N, N) matrix on every rank; gather=False returns (local rows (len, N) with only j >= i valid,
    row indices) without communicating."""
    kern._check_supported()
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rk = dist.get_rank(group) if dist.is_initialized() else 0
    Xs = kern._seqs(X)
    n = Xs.shape[0]
    parts = partition(n, ws, blocks_per_rank)
    mine = parts[rk]
    dev = Xs.device
    rows, ridx, cols, _, src32, maxr = _shard_index(n, parts, rk, dev)
    padded = torch.empty((maxr, n), device=dev, dtype=torch.float32)
    if len(rows) < maxr:
        padded[len(rows):].zero_()
    if len(rows):
        lv = kern._K_seq(Xs, row_blocks=mine)                                   # (M+1, len(rows), N), j >= i only
        if kern.normalization:
            dg = kern._K_seq_diag(Xs)                                           # (M+1, N)
            d1 = dg.index_select(1, ridx)
            kern._finish(lv, d1, dg, normalize=True, diag_cols=cols, out=padded[:len(rows)])
        else:
            kern._finish(lv, normalize=False, out=padded[:len(rows)])
    if not gather:
        return padded[:len(rows)], rows
    gathered = _all_gather_padded(padded, ws, group)
    K = torch.empty((n, n), device=dev, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.gpsig_assemble_symmetric(gathered.data_ptr(), src32.data_ptr(), n, K.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "gpsig_assemble_symmetric")
    return K


def sharded_K(kern, X, X2=None, group=None, blocks_per_rank=8):
    """K(X, X2): symmetric case -> sharded_K_symm; rectangular case shards the rows of X (contiguous blocks dealt the same
    way, no triangle to balance) and all-gathers."""
    if X2 is None:
        return sharded_K_symm(kern, X, group, blocks_per_rank)
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rk = dist.get_rank(group) if dist.is_initialized() else 0
    n = X.shape[0]
    parts = partition(n, ws, blocks_per_rank)
    rows = rows_of(parts[rk])
    Xr = X[torch.as_tensor(rows, device=X.device)] if isinstance(X, torch.Tensor) else X[rows]
    local = kern.K(Xr, X2) if len(rows) else torch.zeros((0, X2.shape[0]), device=kern._dev(), dtype=torch.float32)
    return gather_rows(local, n, parts, group)


# ----------------------------------------------------------------------------------------------------------------------
# Inducing-tensor covariances and the SVGP bound: shard the SEQUENCE axis.  Kzz is tiny and replicated;
# every (z, n) entry of Kzx depends on one inducing tensor and one sequence only, so the column shards are independent
# and one all-gather assembles Kzx; the ELBO needs one all-reduce of a scalar.
# ----------------------------------------------------------------------------------------------------------------------
def column_shards(n, world_size):
    """contiguous [begin, end) of the sequence axis per rank (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    out, b = [], 0
    for r in range(world_size):
        e = b + base + (1 if r < rem else 0)
        out.append((b, e))
        b = e
    return out


def _world(group):
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_columns(local, shards, group=None):
    """All-gather column shards (rank r holds (rows, e_r - b_r)) into (rows, n) with ONE collective."""
    ws = len(shards)
    if ws == 1:
        return local
    width = max(e - b for b, e in shards)
    rows = local.shape[0]
    padded = torch.zeros((width, rows), device=local.device, dtype=local.dtype)
    padded[:local.shape[1]] = local.transpose(0, 1)
    if dist.get_backend(group) == "nccl":
        gathered = torch.empty((ws, width, rows), device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(gathered, padded, group=group)
    else:
        host = padded.cpu()
        chunks = [torch.empty_like(host) for _ in range(ws)]
        dist.all_gather(chunks, host, group=group)
        gathered = torch.stack(chunks).to(local.device)
    return torch.cat([gathered[r, :e - b].transpose(0, 1) for r, (b, e) in enumerate(shards)], dim=1).contiguous()


def sharded_K_tens_vs_seq(kern, Z, X, increments=False, group=None):
    """Kuf = kern.K_tens_vs_seq(Z, X) (kernels.py:538-588) with the sequences sharded over the ranks; full (nz, N) on every
    rank; equal to the single-GPU result up to rounding (the RBF path centres the points on the first sequence of the call;
    exact mode only -- the low-rank mode draws per call and is not sharded)."""
    ws, rk = _world(group)
    shards = column_shards(X.shape[0], ws)
    b, e = shards[rk]
    nz = Z.shape[1]
    if e > b:
        local = kern.K_tens_vs_seq(Z, X[b:e], increments=increments)
    else:
        local = torch.zeros((nz, 0), device=kern._dev(), dtype=torch.float32)
    return gather_columns(local, shards, group)


def sharded_elbo(model, X=None, Y=None, group=None):
    """SVGP bound (models.py:39-59) data-parallel over the sequences: every rank evaluates Kuu_Kuf_Kff, the conditional and
    the variational expectations of ITS shard, the partial sums are all-reduced (one scalar), the KL term is replicated."""
    from . import models as _models
    ws, rk = _world(group)
    X = model.X if X is None else X
    Y = model.Y if Y is None else Y
    n = X.shape[0]
    b, e = column_shards(n, ws)[rk]
    f_mean, f_var, Kzz, q_mu, q_sqrt = model._build_predict(X[b:e], return_Kzz=not model.whiten, return_q=True)
    dev = f_mean.device
    KL = _models.gauss_kl(q_mu, q_sqrt, K=Kzz)
    part = torch.sum(model.likelihood.variational_expectations(f_mean, f_var, model._dev(Y[b:e], dev))).reshape(1).double()
    if ws > 1:
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(part, group=group)
        else:
            host = part.cpu()
            dist.all_reduce(host, group=group)
            part = host.to(dev)
    scale = float(model.num_data) / float(n)
    return part[0].to(torch.float32) * scale - KL
