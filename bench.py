#!/usr/bin/env python
"""
bench.py -- throughput of the signature-kernel covariance path on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg3|cfg5|cfg1] [--kernel linear|rbf]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own algorithm (fp64 NumPy oracle) on the host cores

Workloads (a "step" is one evaluation of the workload's covariance through the public API of gpsig_b200):
  cfg4 (default)  full K(X, X), N=4096 L=128 d=8 M=5, SignatureRBF      -- BASELINE.json configs[3], the metric's shape
  cfg2            full K(X, X), N=1024 L=64 d=6 M=4, SignatureLinear    -- configs[1]
  cfg1            full K(X, X), N=32 L=20 d=3 M=3, SignatureRBF          -- configs[0] (the reference's CPU-sized case)
  cfg3            Kuf, InducingTensors Z=256 (increments) vs N=4096 L=128 d=8 M=5 -- configs[2] (--low-rank: low-rank mode)
  cfg5            SVGP ELBO step (Kuu_Kuf_Kff + Cholesky + conditional + KL), N=8192 Z=512 L=100 d=10 M=6 -- configs[4]
With N > 1 ranks the pair batch is sharded (gpsig_b200.parallel): K(X, X) by row blocks + ONE all-gather, Kuf by sequence
shards + one all-gather, the ELBO data-parallel + one scalar all-reduce.  The problem size is fixed ("scaling": "strong").

  value     work units / step time with the inputs resident in HBM (CUDA events, max over ranks, L2 flushed between
            steps).  Units: output pairs N^2 (K(X, X)), tensor-sequence pairs Z N (Kuf, ELBO).
  e2e       the same through the public API with HOST buffers: inputs start in pinned host memory, the result ends in
            pinned host memory, both copies inside the timed region.  With N > 1 every rank uploads the (replicated)
            inputs and downloads ITS slab of the result into one host buffer shared by the ranks.
  roofline  the dominant kernel of the step, from CUDA events recorded around every launch inside the library
            (gpsig_profile_*).  The default K(X, X) path is the warp-fused kernel (Gram + recursion in one launch, no HBM
            intermediate), which the FP32 pipe bounds: `achieved` = algorithmic FP32 lane-operations per second (per Gram
            entry: d + 2 dot-product / norm terms, 2 for the 2-D increment, 2M - 1 for the recursion -- RBF; d + 2M - 1
            Linear; each FADD / FFMA / FMUL lane-op counts once) against 148 SMs x 128 lanes x the SM clock measured during
            the run.  `hbm_equivalent` keeps the figure the north star is phrased in (Gram bytes the kernel stands for per
            second against the measured copy bandwidth), and `pipeline` carries the same K steps through the HBM-staged
            two-kernel path (knob warpfused = 0) with the roofline of its recursion kernel sigkern_fo_stream_kernel -- the
            kernel the HBM roofline really bounds.  Kuf / ELBO: the tensor-vs-sequence kernel, FP32 lane-ops likewise.
  cpu_baseline  the fp64 NumPy oracle (op-for-op restatement of the reference, oracle/gpsig_oracle.py) on a bounded
            sample of the same workload over all host cores; a reported baseline, not the target.
  parity    after the timed region: entries of the result the timed steps produced against the fp64 oracle (K(X, X): a
            12 x 12 corner AND 512 entries drawn uniformly from the whole N x N matrix; Kuf / ELBO: a sample block).

Nothing here reads /root/reference.  oracle/ is executed only as the CPU baseline (cpu_baseline / --impl reference) and,
after the timed region, as the checker.
"""
import argparse
import datetime
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "cfg4": dict(type="ksymm", N=4096, L=128, d=8, M=5, kernel="rbf", desc="Full K(X,X) N=4096 L=128 d=8 M=5"),
    "cfg2": dict(type="ksymm", N=1024, L=64, d=6, M=4, kernel="linear", desc="SignatureLinear K(X,X) N=1024 L=64 d=6 M=4"),
    "cfg1": dict(type="ksymm", N=32, L=20, d=3, M=3, kernel="rbf", desc="SignatureRBF K(X,X) N=32 L=20 d=3 M=3"),
    "cfg3": dict(type="kuf", N=4096, Z=256, L=128, d=8, M=5, kernel="rbf",
                 desc="Kuf with InducingTensors Z=256, N=4096 L=128 d=8 M=5"),
    "cfg5": dict(type="elbo", N=8192, Z=512, L=100, d=10, M=6, kernel="rbf",
                 desc="SVGP ELBO step (Kuu Cholesky + Kuf matmul + KL) SignatureRBF N=8192 Z=512 L=100 d=10 M=6"),
}
METRICS = {"ksymm": "sequence-pairs/sec for full K(X,X)", "kuf": "tensor-sequence pairs/sec for Kuf",
           "elbo": "tensor-sequence pairs/sec through one SVGP ELBO step"}
UNIT = "pairs/s"


def synth_X(N, L, d, seed=0):
    """unit-scale random walks."""
    rng = np.random.default_rng(seed)
    return (np.cumsum(rng.standard_normal((N, L, d)), axis=1) / np.sqrt(L)).reshape(N, L * d)


def synth_Z(X, L, d, M, nz, seed=2):
    """SURVEY.md 8d / gpsig/utils.py:25-63: pairs of consecutive observations sampled from X plus 0.4 randn."""
    rng = np.random.default_rng(seed)
    T = M * (M + 1) // 2
    Xr = X.reshape(X.shape[0], L, d)
    seq = rng.integers(0, X.shape[0], size=(T, nz))
    t = rng.integers(0, L - 1, size=(T, nz))
    return np.stack([Xr[seq, t], Xr[seq, t + 1]], axis=2) + 0.4 * rng.standard_normal((T, nz, 2, d))


def lengthscales_for(kind, d):
    # RBF: sqrt(d)-scale heuristic (gpsig/utils.py:88-97 gives sqrt(E|x-x'|^2 d) ~ O(sqrt(d)) for unit-scale walks)
    return float(np.sqrt(d)) if kind == "rbf" else 1.0


def units_per_step(wl):
    return wl["N"] * wl["N"] if wl["type"] == "ksymm" else wl["Z"] * wl["N"]


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle over a process pool
# ----------------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(kind, L, d, M, ls, payload):
    try:
        from threadpoolctl import threadpool_limits
        _W["lim"] = threadpool_limits(1)
    except Exception:
        pass
    from oracle import gpsig_oracle as O
    _W["ko"] = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=ls)
    _W.update(payload)


def _cpu_rows(be):
    b, e = be
    ko, Xs = _W["ko"], _W["Xs"]
    return b, ko._K_seq(Xs[b:e], Xs)  # (M+1, e-b, n): Gram + recursion exactly as kernels.py:226 + signature_algs.py:8-35


def _cpu_kuf_cols(be):
    b, e = be
    ko = _W["ko"]
    return b, ko.K_tens_vs_seq(_W["Z"], _W["X"][b:e], increments=True)  # kernels.py:538-588 on a block of sequences


class CpuReference:
    """The workload's covariance on a bounded sample by the oracle, blocks spread over `cores` worker processes."""

    def __init__(self, wl, kind, n_s, cores):
        import multiprocessing as mp
        from oracle import gpsig_oracle as O
        self.wl, self.kind, self.n_s, self.cores = wl, kind, n_s, cores
        L, d, M = wl["L"], wl["d"], wl["M"]
        self.ls = lengthscales_for(kind, d)
        self.ko = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=self.ls)
        self.X = synth_X(n_s, L, d, seed=0)
        if wl["type"] == "ksymm":
            self.Xs = self.ko._scale_seq(self.ko._seqs(self.X))
            self.blocks = [(b, min(n_s, b + 4)) for b in range(0, n_s, 4)]
            payload = {"Xs": self.Xs}
            self.units = n_s * n_s
            self.what = "K(X,X) of %d x %d pairs" % (n_s, n_s)
        else:
            self.nz = min(wl["Z"], 64)
            self.Z = synth_Z(self.X, L, d, M, self.nz)
            step = max(1, n_s // (4 * cores))
            self.blocks = [(b, min(n_s, b + step)) for b in range(0, n_s, step)]
            payload = {"Z": self.Z, "X": self.X}
            self.units = self.nz * n_s
            self.what = "%s of %d tensors x %d sequences" % (
                "Kuf" if wl["type"] == "kuf" else "ELBO step (Kuf blocks + dense algebra)", self.nz, n_s)
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(kind, L, d, M, self.ls, payload))

    def step(self):
        from oracle import gpsig_oracle as O
        if self.wl["type"] == "ksymm":
            lv = np.empty((self.wl["M"] + 1, self.n_s, self.n_s))
            for b, part in self.pool.imap_unordered(_cpu_rows, self.blocks):
                lv[:, b:b + part.shape[1]] = part
            # kernels.py:430-433, :471-476
            lv = lv + self.ko.jitter * np.eye(self.n_s)[None]
            dsq = np.sqrt(np.diagonal(lv, axis1=-2, axis2=-1))
            lv = lv / (dsq[:, :, None] * dsq[:, None, :])
            return (lv * self.ko._weights()[:, None, None]).sum(axis=0)
        Kzx = np.empty((self.nz, self.n_s))
        for b, part in self.pool.imap_unordered(_cpu_kuf_cols, self.blocks):
            Kzx[:, b:b + part.shape[1]] = part
        if self.wl["type"] == "kuf":
            return Kzx
        # ELBO (models.py:39-73): Kzz, conditional, KL, Bernoulli expectations on top of the Kuf blocks
        Kzz = self.ko.K_tens(self.Z, increments=True) + self.ko.jitter * np.eye(self.nz)
        Kxx = np.full((self.n_s,), float(np.sum(self.ko._weights()))) + self.ko.jitter
        q_mu, q_sqrt = np.zeros((self.nz, 1)), np.eye(self.nz)[None]
        fm, fv = O.base_conditional(Kzx, Kzz, Kxx, q_mu, full_cov=False, q_sqrt=q_sqrt, white=True)
        Y = (np.arange(self.n_s)[:, None] % 2).astype(np.float64)
        return np.array([np.sum(O.bernoulli_variational_expectations(fm, fv, Y)) - O.gauss_kl(q_mu, q_sqrt)])

    def close(self):
        self.pool.close()
        self.pool.join()


def time_cpu_reference(wl, kind, n_s, steps, warmup):
    cores = os.cpu_count() or 1
    ref = CpuReference(wl, kind, n_s, cores)
    try:
        for _ in range(warmup):
            ref.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            K = ref.step()
        dt = (time.perf_counter() - t0) / steps
    finally:
        ref.close()
    assert np.isfinite(K).all()
    return dict(value=ref.units / dt, unit=UNIT, cores=cores, kind="port",
                sample="fp64 NumPy oracle (oracle/gpsig_oracle.py), %s at L=%d d=%d M=%d %s, %d worker processes, %.2f s/step"
                       % (ref.what, wl["L"], wl["d"], wl["M"], kind, cores, dt)), dt


def cpu_sample_size(args, wl):
    return min(args.cpu_sample_n, wl["N"]) if wl["type"] == "ksymm" else min(4 * args.cpu_sample_n, wl["N"])


def run_reference(args, wl):
    """--impl reference: only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kind = args.kernel or wl["kernel"]
    cb, dt = time_cpu_reference(wl, kind, cpu_sample_size(args, wl), max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRICS[wl["type"]], "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "static_kernel": kind, "N": wl["N"], "L": wl["L"], "d": wl["d"], "M": wl["M"],
                   "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 100` next to the run.  It is started BEFORE the warm-up steps -- its start-up (NVML initialisation,
    device enumeration) holds driver locks for a few hundred milliseconds and would otherwise stretch the first timed
    steps -- and only the samples stamped inside [mark_begin(), mark_end()] (the timed region) are reported."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def mark_begin(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def __init__(self, uuid):
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", uuid, "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

        def collect(windowed):
            sm, smax, pw, reasons = [], [], [], set()
            for ln in out.strip().splitlines():
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                if windowed:
                    try:
                        ts = datetime.datetime.strptime(f[7], "%Y/%m/%d %H:%M:%S.%f")
                    except (ValueError, IndexError):
                        return None
                    if ts < self.t0 or ts > self.t1 + datetime.timedelta(milliseconds=100):
                        continue
                try:
                    sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return (sm, smax, pw, reasons) if sm else None

        got, window = None, "timed region"
        if self.t0 is not None and self.t1 is not None:
            got = collect(True)
        if got is None:  # timed region shorter than the sampling period (or no timestamps): every sample of the run
            got, window = collect(False), "warm-up + timed region (timed region shorter than the 100 ms sampling period)"
        if got is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm, smax, pw, reasons = got
        # "under load": samples drawing more than half of the peak power seen
        lim = 0.5 * max(pw)
        load = [s for s, p in zip(sm, pw) if p >= lim] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "power_w_max": max(pw), "samples": len(sm),
                "window": window, "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def load_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu captures (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


def fp32_ops_per_entry(kind, d, M):
    """algorithmic FP32 lane-operations per Gram entry of the fused K(X, X) kernel (see the module docstring)"""
    return (d + 2) + 2 + (2 * M - 1) if kind == "rbf" else d + (2 * M - 1)


def fp32_ops_per_kuf_entry(kind, d):
    """per (component, tensor, sequence, time step): two static-kernel evaluations (RBF: d + 2 each; Linear: one d-term dot
    product of increments), the tensor increment, the time increment, 2 for the recursion"""
    return 2 * (d + 2) + 1 + 1 + 2 if kind == "rbf" else d + 2


def tcgen05_takes_kuf(kind, Xnp, Znp, ls, d, M, low_rank):
    """Host-side restatement of the device gate of tens_tc.cu (tens_tc_supported + the radius flag of tc_prep_x_kernel): does
    the tcgen05 kernel take this Kuf call, or does the CUDA-core kernel of tens.cu?  Only names the roofline; no effect on
    what runs."""
    if kind != "rbf" or low_rank or os.environ.get("GPSIG_TENS_TC", "1") == "0" or 3 * d + 6 > 64 or M > 6:
        return False
    rs = np.sqrt(np.log2(np.e) / 2.0)
    c = (Znp.reshape(-1, d).astype(np.float64) / ls).mean(axis=0)
    r2 = (((Xnp.reshape(-1, d).astype(np.float64) / ls - c) * rs) ** 2).sum(axis=1).max()
    return bool(r2 <= 16.0)


class SharedHostMatrix:
    """One pinned host matrix visible to every rank of the node (file in /dev/shm mapped shared, registered with CUDA):
    each rank copies ITS slab of the result there, so the e2e path needs no second collective."""

    def __init__(self, shape, rank, world, tag):
        import torch
        import torch.distributed as dist
        self.path = "/dev/shm/gpsig_b200_%s_%s" % (tag, os.environ.get("MASTER_PORT", str(os.getpid())))
        n = int(np.prod(shape))
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(n * 4)
        if world > 1:
            dist.barrier()
        self.t = torch.from_file(self.path, shared=True, size=n, dtype=torch.float32).reshape(shape)
        self.registered = False
        try:
            rc = torch.cuda.cudart().cudaHostRegister(self.t.data_ptr(), n * 4, 0)
            self.registered = (int(rc) == 0)
        except Exception:
            self.registered = False
        if world > 1:
            dist.barrier()
        self.rank = rank

    def close(self):
        import torch
        try:
            if self.registered:
                torch.cuda.cudart().cudaHostUnregister(self.t.data_ptr())
        except Exception:
            pass
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def read_profile(lib, _lib, classes):
    out = {}
    for name, c in classes:
        ms, n, un = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        _lib.check(lib.gpsig_profile_read(c, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(un)), "gpsig_profile_read")
        out[name] = (ms.value, n.value, un.value)
    return out


PROF_CLASSES = (("prep", 0), ("producer", 1), ("recursion", 2), ("recursion_other", 3), ("epilogue", 4), ("tens", 5), ("fused", 6))


def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gpsig_b200 import kernels, parallel, models, inducing_variables as iv, _lib, settings
    lib = _lib.load()
    if args.workspace_gb:
        settings.workspace_budget_bytes = int(args.workspace_gb * (1 << 30))

    wtype = wl["type"]
    N, L, d, M = wl["N"], wl["L"], wl["d"], wl["M"]
    kind = args.kernel or wl["kernel"]
    cls = dict(linear=kernels.SignatureLinear, rbf=kernels.SignatureRBF)[kind]
    kw = dict(low_rank=True) if (args.low_rank and wtype == "kuf") else {}
    kern = cls(L * d, d, M, lengthscales=lengthscales_for(kind, d), **kw)
    Xnp = synth_X(N, L, d, seed=0).astype(np.float32)
    Xh = torch.from_numpy(Xnp).pin_memory()
    Xd = Xh.to(dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    units = units_per_step(wl)
    Zd = Zh = model = None
    tc_path = False
    if wtype in ("kuf", "elbo"):
        nz = wl["Z"]
        Znp = synth_Z(Xnp.astype(np.float64), L, d, M, nz).astype(np.float32)
        Zh = torch.from_numpy(Znp).pin_memory()
        Zd = Zh.to(dev)
        tc_path = tcgen05_takes_kuf(kind, Xnp, Znp, lengthscales_for(kind, d), d, M, args.low_rank)
    if wtype == "elbo":
        Ynp = (np.arange(N)[:, None] % 2).astype(np.float64)
        rngq = np.random.default_rng(7)
        model = models.SVGP(Xd, torch.from_numpy(Ynp).to(dev), kern, models.Bernoulli(), iv.InducingTensors(Zd, M, increments=True),
                            num_latent=1, q_mu=0.1 * rngq.standard_normal((nz, 1)))

    # result buffers on the host: one matrix shared by the ranks (every rank writes its slab)
    out_shape = {"ksymm": (N, N), "kuf": (wl.get("Z", 1), N), "elbo": (1, 1)}[wtype]
    shared = SharedHostMatrix(out_shape, rank, world, args.workload) if wtype != "elbo" else None
    Hh = shared.t if shared is not None else torch.empty(out_shape, dtype=torch.float32).pin_memory()
    slab = parallel.column_shards(out_shape[0] if wtype == "ksymm" else out_shape[1], world)[rank]

    def step_dev(X=None, Z=None):
        X = Xd if X is None else X
        if wtype == "ksymm":
            return parallel.sharded_K_symm(kern, X, blocks_per_rank=args.blocks_per_rank) if world > 1 else kern.K(X)
        if wtype == "kuf":
            Zz = Zd if Z is None else Z
            if world > 1:
                return parallel.sharded_K_tens_vs_seq(kern, Zz, X, increments=True)
            return kern.K_tens_vs_seq(Zz, X, increments=True)
        with torch.no_grad():
            return parallel.sharded_elbo(model, X=X) if world > 1 else model._build_likelihood(X, model.Y)

    def step_e2e():
        X = Xh.to(dev, non_blocking=True)
        if wtype == "ksymm":
            K = step_dev(X)
            Hh[slab[0]:slab[1]].copy_(K[slab[0]:slab[1]], non_blocking=True)
        elif wtype == "kuf":
            K = step_dev(X, Zh.to(dev, non_blocking=True))
            Hh[:, slab[0]:slab[1]].copy_(K[:, slab[0]:slab[1]], non_blocking=True)
        else:
            K = step_dev(X)
            Hh.copy_(K.reshape(1, 1).to(torch.float32), non_blocking=True)
        return K

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        if sampler:
            sampler.mark_begin()
        for e0, e1 in ev:
            flush.zero_()          # L2 flush, outside the event pair
            e0.record()
            fn()
            e1.record()
        barrier()
        if sampler:
            torch.cuda.synchronize()
            sampler.mark_end()
        clocks = sampler.stop() if sampler else None
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", "") if rank == 0 else None
    sampler = ClockSampler(uuid) if uuid else None
    for _ in range(args.warmup):
        step_dev()
    barrier()

    lib.gpsig_profile_reset()
    lib.gpsig_profile_enable(1)
    l0 = lib.gpsig_launch_count()
    ms_total, clocks = timed(step_dev, args.steps, sampler)
    l1 = lib.gpsig_launch_count()
    lib.gpsig_profile_enable(0)
    prof = read_profile(lib, _lib, PROF_CLASSES)
    lib.gpsig_profile_reset()
    launches = torch.tensor([l1 - l0], device=dev, dtype=torch.int64)
    per_rank = None
    if world > 1:
        dist.all_reduce(launches)
        # the dominant kernel's time on EVERY rank (the step time is the max over ranks: a slow GPU or an unbalanced shard
        # shows here, a slow collective does not)
        mine = torch.tensor([(prof["fused"][0] + prof["tens"][0]) / args.steps], device=dev, dtype=torch.float64)
        allk = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allk, mine)
        per_rank = [round(float(t.item()), 3) for t in allk]

    # e2e: host buffers in, host buffer out
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    barrier()

    # K(X, X): the same steps through the two-kernel pipeline (increment-Gram producer -> HBM -> stream recursion), the
    # design the HBM roofline describes; the default path above is the warp-fused kernel, which never writes the Gram tensor
    prof_pipe, ms_pipe = None, None
    if wtype == "ksymm" and prof["fused"][0] > 0 and not args.no_pipeline:
        _lib.set_knob("warpfused", 0)
        try:
            step_dev()
            lib.gpsig_profile_reset()
            lib.gpsig_profile_enable(1)
            ms_pipe, _ = timed(step_dev, args.steps)
            lib.gpsig_profile_enable(0)
            prof_pipe = read_profile(lib, _lib, (("prep", 0), ("producer", 1), ("recursion", 2), ("epilogue", 4)))
            lib.gpsig_profile_reset()
        finally:
            _lib.set_knob("warpfused", 1)

    barrier()  # every rank's slab of the last e2e step has landed in the shared host buffer
    if rank != 0:
        if shared is not None:
            barrier()  # rank 0 reads the shared buffer for the parity check before anyone unmaps it
            shared.close()
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = units / (ms_step * 1e-3)
    peak_hbm, peak_src = load_peaks()
    sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
    peak_fp32 = 148 * 128 * sm_hz / 1e12  # T lane-ops / s

    def roof_hbm(kname, cls_prof, total_ms, traffic):
        b_pair = 4 * L * L + 4 * (M + 1)
        r_ms, r_n, r_units = cls_prof
        achieved = (r_units * b_pair) / (r_ms * 1e-3) / 1e9 if r_ms > 0 else 0.0
        return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak_hbm, "unit": "GB/s", "frac": achieved / peak_hbm,
                "traffic": traffic, "peak_source": peak_src, "bytes_per_pair": b_pair, "pairs_per_launch": r_units / max(r_n, 1),
                "launches": r_n, "avg_launch_ms": r_ms / max(r_n, 1), "kernel_share_of_step": r_ms / total_ms}

    def roof_fp32(kname, cls_prof, total_ms, ops_per_unit, traffic, how):
        r_ms, r_n, r_units = cls_prof
        achieved = r_units * ops_per_unit / (r_ms * 1e-3) / 1e12 if r_ms > 0 else 0.0
        return {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": peak_fp32, "unit": "Tlaneop/s",
                "frac": achieved / peak_fp32, "traffic": traffic,
                "peak_source": "148 SMs x 128 FP32 lanes x %.0f MHz (SM clock sampled under load during this run)" % (sm_hz / 1e6),
                "lane_ops_per_unit": ops_per_unit, "unit_of_work": how, "units_per_launch": r_units / max(r_n, 1), "launches": r_n,
                "avg_launch_ms": r_ms / max(r_n, 1), "kernel_share_of_step": r_ms / total_ms}

    pipeline = None
    if wtype == "ksymm":
        if prof["fused"][0] > 0:
            per_pair = fp32_ops_per_entry(kind, d, M) * (L - 1) * (L - 1)
            traffic = load_traffic("%s_warpfused_%s_bytes_per_launch" % (args.workload, kind)) if world == 1 else None
            roofline = roof_fp32("sigkern_warpfused_kernel", prof["fused"], ms_total, per_pair, traffic,
                                 "sequence pair = (L-1)^2 Gram entries x %d lane-ops" % fp32_ops_per_entry(kind, d, M))
            roofline["hbm_equivalent"] = roof_hbm("sigkern_warpfused_kernel", prof["fused"], ms_total, traffic)
            roofline["hbm_equivalent"]["note"] = ("Gram bytes the fused kernel stands for per second; it never reads them "
                                                  "(see `pipeline.roofline` for the HBM-staged recursion kernel)")
            if prof_pipe is not None:
                pipeline = {"ms_per_step": ms_pipe / args.steps, "value": units / (ms_pipe / args.steps * 1e-3), "unit": UNIT,
                            "how": "same steps with knob warpfused=0: increment-Gram producer -> HBM chunk -> stream recursion",
                            "stages": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                                       for k, v in prof_pipe.items()},
                            "roofline": roof_hbm("sigkern_fo_stream_kernel", prof_pipe["recursion"], ms_pipe,
                                                 load_traffic("%s_recursion_bytes_per_launch" % args.workload) if world == 1 else None)}
        else:
            roofline = roof_hbm("sigkern_fo_stream_kernel", prof["recursion"], ms_total,
                                load_traffic("%s_recursion_bytes_per_launch" % args.workload) if world == 1 else None)
    else:
        T = M * (M + 1) // 2
        per_pair = fp32_ops_per_kuf_entry(kind, d) * T * L
        traffic = load_traffic("%s_tens_bytes_per_launch" % args.workload) if world == 1 else None
        fp32_roof = roof_fp32("tens_seq_fast_kernel", prof["tens"], ms_total, per_pair, traffic,
                              "(tensor, sequence) pair = T L = %d x %d component-time entries x %d lane-ops"
                              % (T, L, fp32_ops_per_kuf_entry(kind, d)))
        if tc_path:
            # the tcgen05 kernel moves the Gram onto the tensor cores; what is left per Gram entry is ONE MUFU.EX2, and the
            # SFU (16 lanes / clk / SM, tools/ubench/pipes.cu) is the pipe that bounds it: 2 T L exponentials per pair
            r_ms, r_n, r_units = prof["tens"]
            r_n = r_n / 2  # every call launches the tcgen05 kernel and, behind it, the CUDA-core kernel, which returns at once
            ex_per_pair = 2 * T * L
            peak_mufu = 148 * 16 * sm_hz / 1e12
            achieved = r_units * ex_per_pair / (r_ms * 1e-3) / 1e12 if r_ms > 0 else 0.0
            roofline = {"bound": "mufu", "kernel": "tens_seq_tc_kernel", "achieved": achieved, "peak": peak_mufu, "unit": "Tex2/s",
                        "frac": achieved / peak_mufu, "traffic": traffic,
                        "peak_source": "148 SMs x 16 SFU lanes x %.0f MHz (MUFU.EX2 measured at 0.5 warp-instr/clk/SM, "
                                       "profiles/r2a_pipes.log)" % (sm_hz / 1e6),
                        "ex2_per_unit": ex_per_pair, "unit_of_work": "(tensor, sequence) pair = 2 T L = 2 x %d x %d static-kernel "
                        "evaluations; their d-term dot products run on tcgen05 (kind::tf32, split operands)" % (T, L),
                        "units_per_launch": r_units / max(r_n, 1), "launches": r_n, "avg_launch_ms": r_ms / max(r_n, 1),
                        "kernel_share_of_step": r_ms / ms_total,
                        "timing_note": "avg_launch_ms includes the few microseconds of the early-exit CUDA-core launch behind it",
                        "cuda_core_equivalent": dict(fp32_roof, note="the same work counted as the CUDA-core kernel's FP32 lane-ops "
                                                     "(a fraction above 1 is what moving the Gram to the tensor cores bought)")}
        else:
            roofline = fp32_roof
    stages = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}

    # parity (not timed): entries of what the timed steps produced against the fp64 oracle
    from oracle import gpsig_oracle as O
    ko = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=lengthscales_for(kind, d))
    X64 = Xnp.astype(np.float64)
    if wtype == "ksymm":
        Kn = Hh.numpy()
        nc = min(12, N)
        ref = ko.K(X64[:nc])
        parity = {"corner": "%dx%d" % (nc, nc),
                  "max_abs_err_over_max_abs_ref": float(np.max(np.abs(Kn[:nc, :nc] - ref)) / np.max(np.abs(ref))), "tolerance": 1e-4}
        rngp = np.random.default_rng(123)
        ns = min(512, N * N)
        ii, jj = rngp.integers(0, N, size=ns), rngp.integers(0, N, size=ns)
        scaled = lambda s: ko._scale_seq(ko._seqs(X64[s:s + 1]))  # noqa: E731
        dg = {int(s): ko._K_seq_diag(scaled(s))[:, 0] for s in np.unique(np.concatenate([ii, jj]))}  # (M+1,) per sequence
        w = ko._weights()
        worst, big = 0.0, 0.0
        for a, b in zip(ii, jj):
            lv = ko._K_seq(scaled(a), scaled(b))[:, 0, 0]
            lv = lv + (ko.jitter if a == b else 0.0)                                              # kernels.py:431
            val = float(np.sum(w * lv / (np.sqrt(dg[int(a)] + ko.jitter) * np.sqrt(dg[int(b)] + ko.jitter))))
            worst, big = max(worst, abs(float(Kn[a, b]) - val)), max(big, abs(val))
        parity["random_entries"] = {"count": int(ns), "max_abs_err_over_max_abs_ref": worst / big, "tolerance": 1e-4,
                                    "how": "entries (i, j) drawn uniformly from the N x N result of the timed e2e steps"}
    elif wtype == "kuf" and not args.low_rank:
        zs, ns_ = min(8, wl["Z"]), min(24, N)
        Z64 = Zh.numpy().astype(np.float64)
        ref = ko.K_tens_vs_seq(Z64[:, :zs], X64[:ns_], increments=True)
        parity = {"block": "%dx%d" % (zs, ns_),
                  "max_abs_err_over_max_abs_ref": float(np.max(np.abs(Hh.numpy()[:zs, :ns_] - ref)) / np.max(np.abs(ref))),
                  "tolerance": 1e-4}
    elif wtype == "elbo":
        # the bound itself on a sub-problem small enough for the oracle
        zs, ns_ = 16, 48
        Z64 = Zh.numpy().astype(np.float64)[:, :zs]
        q_mu_s = np.asarray(model.q_mu)[:zs]
        msub = models.SVGP(Xd[:ns_], model.Y[:ns_], kern, models.Bernoulli(), iv.InducingTensors(Zd[:, :zs], M, increments=True),
                           num_latent=1, q_mu=q_mu_s)
        got = msub.compute_log_likelihood()
        ref = O.svgp_elbo(ko, Z64, X64[:ns_], model.Y[:ns_].cpu().numpy(), q_mu_s, np.eye(zs)[None],
                          likelihood="bernoulli", increments=True)[0]
        parity = {"sub_problem": "Z=%d N=%d" % (zs, ns_), "elbo": got, "oracle_elbo": float(ref),
                  "rel_err": abs(got - ref) / abs(ref), "tolerance": 1e-4}
    else:
        parity = {"note": "low-rank mode is randomised: parity is defined on injected draws (tests/test_gpu_lowrank.py)"}
    if shared is not None and world > 1:
        barrier()  # the other ranks may unmap the shared buffer now

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = time_cpu_reference(wl, kind, cpu_sample_size(args, wl), 1, 0)
    par = ({"ksymm": "row-sharded x%d + one all-gather", "kuf": "sequence-sharded x%d + one all-gather",
            "elbo": "data-parallel x%d + one scalar all-reduce"}[wtype] % world) if world > 1 else "single GPU"
    h2d = int(Xh.numel() * 4 + (Zh.numel() * 4 if (Zh is not None and wtype == "kuf") else 0))
    d2h_rank = {"ksymm": (slab[1] - slab[0]) * N * 4, "kuf": out_shape[0] * (slab[1] - slab[0]) * 4, "elbo": 4}[wtype]
    line = {
        "metric": METRICS[wtype], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "static_kernel": kind, "N": N, "L": L, "d": d, "M": M, "order": 1,
                   "normalization": True, "parallelism": par, "l2": "flushed between steps (512 MiB memset)",
                   "workspace_budget_GiB": settings.workspace_budget_bytes / (1 << 30)},
        "e2e": {"value": units / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": int(np.prod(out_shape)) * 4, "ms_per_step": ms_e2e / args.steps,
                "per_rank": {"h2d_bytes": h2d, "d2h_bytes": int(d2h_rank)},
                "host_result": ("one pinned host buffer shared by the ranks; every rank writes its slab" if world > 1
                                else "pinned host buffer")},
        "gpu_launches": int(launches.item()),
        "per_rank_kernel_ms": per_rank,
        "clocks": clocks,
        "roofline": roofline,
        "pipeline": pipeline,
        "stages": stages,
        "parity": parity,
        "cpu_baseline": cpu_baseline,
    }
    if wtype == "ksymm":
        line["config"]["symmetric_half_computed"] = True
    else:
        line["config"]["Z"] = wl["Z"]
        line["config"]["low_rank"] = bool(args.low_rank)
    print(json.dumps(line), flush=True)
    if shared is not None:
        shared.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default=None, choices=["linear", "rbf"])
    ap.add_argument("--low-rank", action="store_true", help="cfg3: low-rank mode (kernels.py low_rank=True)")
    ap.add_argument("--cpu-sample-n", type=int, default=128,
                    help="CPU arm: n_s x n_s pairs (K) / 4 n_s sequences (Kuf, ELBO) per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the second pass through the two-kernel pipeline")
    ap.add_argument("--workspace-gb", type=float, default=None)
    ap.add_argument("--blocks-per-rank", type=int, default=8)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: spawn the ranks ourselves (the driver launches torchrun directly)
        import random
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", str(random.randint(20000, 40000)), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, wl)


if __name__ == "__main__":
    main()
