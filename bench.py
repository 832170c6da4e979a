#!/usr/bin/env python
"""
bench.py -- sequence-pairs/sec of the full signature-kernel covariance K(X, X) (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2] [--kernel linear|rbf]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own algorithm (fp64 NumPy oracle) on the host cores

A "step" is one evaluation of K(X, X) for the whole workload: point prep -> warp-fused increment-Gram + level recursion
kernel (or, as the `pipeline` pass: chunked increment-Gram producer -> bulk-copy-staged stream recursion) -> normalise /
weight / sum -> mirror (gpsig_b200.kernels.SignatureKernel.K; with N > 1 gpsig_b200.parallel.sharded_K_symm: row blocks
dealt over the ranks, ONE all-gather of the assembled rows).  The problem size is fixed as N grows ("scaling": "strong").

  value     N^2 output pairs / step time, X already resident in HBM (CUDA events, max over ranks, L2 flushed between
            steps).
  e2e       the same through the public API with HOST buffers: X starts in pinned host memory, K ends in pinned
            host memory, both copies inside the timed region.
  roofline  the dominant kernel of the step: algorithmic bytes per pair (4 L1 L2 + 4 (M+1)) x pairs processed
            / its own CUDA-event duration (events recorded around every launch inside the library: gpsig_profile_*),
            against MEASURED_PEAKS.json's hbm_gbs.  The default path is the warp-fused kernel (Gram + recursion in one
            launch, no HBM intermediate: FP32-bound, the figure is the Gram bytes it stands for); `pipeline` carries the
            same K steps through the HBM-staged two-kernel path (GPSIG_WARPFUSED=0) with the roofline of its recursion
            kernel sigkern_fo_stream_kernel -- the kernel the HBM roofline really bounds.
  cpu_baseline  the fp64 NumPy oracle (op-for-op restatement of the reference, oracle/gpsig_oracle.py) on a bounded
            sample (n_s x n_s pairs of the same L/d/M) over all host cores; a reported baseline, not the target.

Nothing here reads /root/reference.  oracle/ is executed only as the CPU baseline (cpu_baseline / --impl reference) and,
after the timed region, as the checker of a 12 x 12 corner of the result ("parity").
"""
import argparse
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
    # BASELINE.json configs[3] / the north star's target shape; fits one GPU through the chunked pipeline
    "cfg4": dict(N=4096, L=128, d=8, M=5, kernel="rbf", desc="Full K(X,X) N=4096 L=128 d=8 M=5"),
    # BASELINE.json configs[1]
    "cfg2": dict(N=1024, L=64, d=6, M=4, kernel="linear", desc="SignatureLinear K(X,X) N=1024 L=64 d=6 M=4"),
}
METRIC = "sequence-pairs/sec for full K(X,X)"
UNIT = "pairs/s"


def synth_X(N, L, d, seed=0):
    """unit-scale random walks."""
    rng = np.random.default_rng(seed)
    return (np.cumsum(rng.standard_normal((N, L, d)), axis=1) / np.sqrt(L)).reshape(N, L * d)


def lengthscales_for(kind, d):
    # RBF: sqrt(d)-scale heuristic (gpsig/utils.py:88-97 gives sqrt(E|x-x'|^2 d) ~ O(sqrt(d)) for unit-scale walks)
    return float(np.sqrt(d)) if kind == "rbf" else 1.0


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle over a process pool (one row block per task)
# ----------------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(kind, L, d, M, ls, Xs):
    try:
        from threadpoolctl import threadpool_limits
        _W["lim"] = threadpool_limits(1)
    except Exception:
        pass
    from oracle import gpsig_oracle as O
    _W["ko"] = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=ls)
    _W["Xs"] = Xs


def _cpu_rows(be):
    b, e = be
    ko, Xs = _W["ko"], _W["Xs"]
    return b, ko._K_seq(Xs[b:e], Xs)  # (M+1, e-b, n): Gram + recursion exactly as kernels.py:226 + signature_algs.py:8-35


class CpuReference:
    """K(X, X) of n_s sequences by the oracle, rows blocks spread over `cores` worker processes."""

    def __init__(self, kind, L, d, M, n_s, cores, row_block=4):
        import multiprocessing as mp
        from oracle import gpsig_oracle as O
        self.kind, self.L, self.d, self.M, self.n_s, self.cores = kind, L, d, M, n_s, cores
        self.ls = lengthscales_for(kind, d)
        self.ko = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=self.ls)
        self.X = synth_X(n_s, L, d, seed=0)
        self.Xs = self.ko._scale_seq(self.ko._seqs(self.X))
        self.blocks = [(b, min(n_s, b + row_block)) for b in range(0, n_s, row_block)]
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init,
                                                 initargs=(kind, L, d, M, self.ls, self.Xs))

    def step(self):
        lv = np.empty((self.M + 1, self.n_s, self.n_s))
        for b, part in self.pool.imap_unordered(_cpu_rows, self.blocks):
            lv[:, b:b + part.shape[1]] = part
        # kernels.py:430-433, :471-476
        lv = lv + self.ko.jitter * np.eye(self.n_s)[None]
        dsq = np.sqrt(np.diagonal(lv, axis1=-2, axis2=-1))
        lv = lv / (dsq[:, :, None] * dsq[:, None, :])
        return (lv * self.ko._weights()[:, None, None]).sum(axis=0)

    def close(self):
        self.pool.close()
        self.pool.join()


def time_cpu_reference(kind, L, d, M, n_s, steps, warmup):
    cores = os.cpu_count() or 1
    ref = CpuReference(kind, L, d, M, n_s, cores)
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
    return dict(value=n_s * n_s / dt, unit=UNIT, cores=cores, kind="port",
                sample="fp64 NumPy oracle (oracle/gpsig_oracle.py), K(X,X) of %d x %d pairs at L=%d d=%d M=%d %s, %d worker "
                       "processes, %.2f s/step" % (n_s, n_s, L, d, M, kind, cores, dt)), dt


def run_reference(args, wl):
    """--impl reference: only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    kind = args.kernel or wl["kernel"]
    cb, dt = time_cpu_reference(kind, wl["L"], wl["d"], wl["M"], args.cpu_sample_n, max(1, args.steps), min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "static_kernel": kind, "N": wl["N"], "L": wl["L"], "d": wl["d"], "M": wl["M"],
                   "sample": "%d x %d pairs per step" % (args.cpu_sample_n, args.cpu_sample_n)},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.proc = None
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
        sm, smax, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load": samples drawing more than half of the peak power seen
        lim = 0.5 * max(pw)
        load = [s for s, p in zip(sm, pw) if p >= lim] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def load_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the recursion kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(key)
    except Exception:
        return None


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
    from gpsig_b200 import kernels, parallel, _lib, settings
    lib = _lib.load()
    if args.workspace_gb:
        settings.workspace_budget_bytes = int(args.workspace_gb * (1 << 30))

    N, L, d, M = wl["N"], wl["L"], wl["d"], wl["M"]
    kind = args.kernel or wl["kernel"]
    cls = dict(linear=kernels.SignatureLinear, rbf=kernels.SignatureRBF)[kind]
    kern = cls(L * d, d, M, lengthscales=lengthscales_for(kind, d))
    Xnp = synth_X(N, L, d, seed=0).astype(np.float32)
    Xh = torch.from_numpy(Xnp).pin_memory()
    Xd = Xh.to(dev)
    Kh = torch.empty((N, N), dtype=torch.float32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_dev():
        if world > 1:
            return parallel.sharded_K_symm(kern, Xd, blocks_per_rank=args.blocks_per_rank)
        return kern.K(Xd)

    def step_e2e():
        K = parallel.sharded_K_symm(kern, Xh, blocks_per_rank=args.blocks_per_rank) if world > 1 else kern.K(Xh)
        Kh.copy_(K, non_blocking=True)
        return K

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler_uuid=None):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        sampler = ClockSampler(sampler_uuid) if sampler_uuid else None
        for e0, e1 in ev:
            flush.zero_()          # L2 flush, outside the event pair
            e0.record()
            fn()
            e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    for _ in range(args.warmup):
        K = step_dev()
    barrier()

    lib.gpsig_profile_reset()
    lib.gpsig_profile_enable(1)
    l0 = lib.gpsig_launch_count()
    uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", "") if rank == 0 else None
    ms_total, clocks = timed(step_dev, args.steps, uuid)
    l1 = lib.gpsig_launch_count()
    lib.gpsig_profile_enable(0)
    import ctypes
    prof = {}
    for name, c in (("prep", 0), ("producer", 1), ("recursion", 2), ("recursion_other", 3), ("epilogue", 4), ("fused", 6)):
        ms, n, un = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        _lib.check(lib.gpsig_profile_read(c, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(un)), "gpsig_profile_read")
        prof[name] = (ms.value, n.value, un.value)
    lib.gpsig_profile_reset()
    launches = torch.tensor([l1 - l0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)

    # e2e: host buffers in, host buffer out
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    # the same steps through the two-kernel pipeline (increment-Gram producer -> HBM -> stream recursion), the design the
    # HBM roofline describes; the default path above is the warp-fused kernel, which never writes the Gram tensor
    prof_pipe, ms_pipe = None, None
    if prof["fused"][0] > 0:
        _lib.set_knob("warpfused", 0)
        try:
            step_dev()
            lib.gpsig_profile_reset()
            lib.gpsig_profile_enable(1)
            ms_pipe, _ = timed(step_dev, args.steps)
            lib.gpsig_profile_enable(0)
            prof_pipe = {}
            for name, c in (("prep", 0), ("producer", 1), ("recursion", 2), ("epilogue", 4)):
                ms, n, un = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
                _lib.check(lib.gpsig_profile_read(c, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(un)), "gpsig_profile_read")
                prof_pipe[name] = (ms.value, n.value, un.value)
            lib.gpsig_profile_reset()
        finally:
            _lib.set_knob("warpfused", 1)
    if rank == 0:
        assert np.isfinite(Kh.numpy()[:8]).all()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = N * N / (ms_step * 1e-3)
    # roofline of the recursion kernel (rank 0's launches)
    peak, peak_src = load_peaks()
    b_pair = 4 * L * L + 4 * (M + 1)
    def roof(kname, cls_prof, total_ms, traffic, note=None):
        r_ms, r_n, r_units = cls_prof
        achieved = (r_units * b_pair) / (r_ms * 1e-3) / 1e9 if r_ms > 0 else 0.0
        out = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
               "traffic": traffic, "peak_source": peak_src, "bytes_per_pair": b_pair, "pairs_per_launch": r_units / max(r_n, 1),
               "launches": r_n, "avg_launch_ms": r_ms / max(r_n, 1), "kernel_share_of_step": r_ms / total_ms}
        if note:
            out["note"] = note
        return out

    pipeline = None
    if prof["fused"][0] > 0:
        # default path: Gram + recursion in ONE kernel.  `achieved` is the ALGORITHMIC Gram bytes (same per-pair figure) the
        # kernel stands for per second; its real DRAM traffic (`traffic`) is the inputs and outputs only -- the kernel is
        # FP32-issue bound, the HBM roofline is what it removes.  The `pipeline` object below carries the HBM-staged
        # two-kernel path measured in this same run.
        roofline = roof("sigkern_warpfused_kernel", prof["fused"], ms_total,
                        load_traffic("%s_warpfused_bytes_per_launch" % args.workload) if world == 1 else None,
                        "fused Gram + recursion: no HBM intermediate (FP32-issue bound); see `pipeline.roofline` for the "
                        "HBM-staged recursion kernel")
        if prof_pipe is not None:
            pipeline = {"ms_per_step": ms_pipe / args.steps, "value": N * N / (ms_pipe / args.steps * 1e-3), "unit": UNIT,
                        "how": "same steps with GPSIG_WARPFUSED=0: increment-Gram producer -> HBM chunk -> stream recursion",
                        "stages": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                                   for k, v in prof_pipe.items()},
                        "roofline": roof("sigkern_fo_stream_kernel", prof_pipe["recursion"], ms_pipe,
                                         load_traffic("%s_recursion_bytes_per_launch" % args.workload) if world == 1 else None)}
    else:
        roofline = roof("sigkern_fo_stream_kernel", prof["recursion"], ms_total,
                        load_traffic("%s_recursion_bytes_per_launch" % args.workload) if world == 1 else None)
    stages = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
    p_ms, p_n, p_units = prof["producer"]
    if p_ms > 0:
        stages["producer"]["store_GBps"] = p_units * 4 * (L - 1) * kernels_pitch(L) / (p_ms * 1e-3) / 1e9

    # parity spot check (not timed): a corner of the matrix the timed steps produced against the fp64 oracle
    from oracle import gpsig_oracle as O
    nc = 12
    ko = O.SignatureKernelOracle(kind, L * d, d, M, lengthscales=lengthscales_for(kind, d))
    ref = ko.K(Xnp[:nc].astype(np.float64))
    parity = {"corner": "%dx%d" % (nc, nc), "max_abs_err_over_max_abs_ref": float(np.max(np.abs(Kh.numpy()[:nc, :nc] - ref)) / np.max(np.abs(ref))),
              "tolerance": 1e-4}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = time_cpu_reference(kind, L, d, M, args.cpu_sample_n, 1, 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "static_kernel": kind, "N": N, "L": L, "d": d, "M": M, "order": 1,
                   "normalization": True, "symmetric_half_computed": True,
                   "parallelism": "row-sharded x%d + one all-gather" % world if world > 1 else "single GPU",
                   "l2": "flushed between steps (512 MiB memset)",
                   "workspace_budget_GiB": settings.workspace_budget_bytes / (1 << 30)},
        "e2e": {"value": N * N / (ms_e2e / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(Xh.numel() * 4),
                "d2h_bytes_per_step": int(Kh.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches.item()),
        "clocks": clocks,
        "roofline": roofline,
        "pipeline": pipeline,
        "stages": stages,
        "parity": parity,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernels_pitch(L):
    """column pitch of the increment-Gram chunk buffer (16 columns per lane, power-of-two lanes per pair)."""
    need = max(2, -(-(L - 1) // 16))
    lp = 1
    while lp < need:
        lp *= 2
    return 16 * lp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default=None, choices=["linear", "rbf"])
    ap.add_argument("--cpu-sample-n", type=int, default=128, help="CPU arm: n_s x n_s pairs per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
