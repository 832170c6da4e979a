#!/usr/bin/env python
"""
Turn ncu artefacts from gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv            > profiles/x_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/x_prof.ncu-rep [launch#]  > profiles/x_prof.md

`launches`: per-kernel count / total / share of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised:
compare SHARES).  `kernel`: the headline counters of one `--set full` capture, the warp-stall breakdown and the hottest
SASS instructions (needs -lineinfo builds; runs `ncu -i` here, no GPU needed).
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]


def launches(path):
    rows = list(csv.reader(open(path)))
    i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, body = rows[i0], rows[i0 + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in body:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split("(")[0].replace("void ", ""), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("source: %s  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)" % path)
    print("total %.3f ms over %d launches\n" % (tot, sum(a[0] for a in agg.values())))
    print("| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% | %.4f |" % (k[:80], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


def _ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def kernel(path, which=0):
    raw = list(csv.reader(io.StringIO(_ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    r = rows[which]
    idx = {k: i for i, k in enumerate(hdr)}
    print("source: %s, launch %d of %d captured (ncu --set full --clock-control none --import-source on)" % (path, which, len(rows)))
    print("kernel: `%s`\n" % r[idx["Kernel Name"]])
    print("| counter | value |\n|---|---:|")
    for k, label in KEYS:
        if k in idx:
            print("| %s (`%s`) | %s %s |" % (label, k, r[idx[k]], units[idx[k]]))
    src = list(csv.reader(io.StringIO(_ncu(["-i", path, "--page", "source", "--csv", "--print-source", "sass"]))))
    starts = [i for i, rr in enumerate(src) if rr and rr[0] == "Address"]
    if not starts:
        return
    s0 = starts[min(which, len(starts) - 1)]
    s1 = [s for s in starts if s > s0]
    shdr = src[s0]
    body = [rr for rr in src[s0 + 1:(s1[0] - 1 if s1 else len(src))] if len(rr) >= len(shdr)]
    sidx = {k: i for i, k in enumerate(shdr)}
    stalls = [k for k in shdr if k.startswith("stall_") and "Not Issued" not in k]
    ns = sum(int(rr[sidx["# Samples"]] or 0) for rr in body)
    tot = {k: sum(int(rr[sidx[k]] or 0) for rr in body) for k in stalls}
    print("\nwarp-stall sampling (%d samples):\n\n| reason | share |\n|---|---:|" % ns)
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        if v:
            print("| %s | %.1f%% |" % (k, 100.0 * v / max(ns, 1)))
    ops = collections.Counter()
    nfull = none = 0
    for rr in body:
        n = int(rr[sidx["Instructions Executed"]] or 0)
        t = float(rr[sidx["Avg. Threads Executed"]] or 0)
        if t <= 1.5:
            none += n
        else:
            nfull += n
        tok = rr[sidx["Source"]].strip().split()
        if tok:
            o = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
            ops[o.split(".")[0]] += n
    print("\nwarp instructions executed: %d (of which %d with a single active thread)\n" % (nfull + none, none))
    print("| opcode | share of executed warp instructions |\n|---|---:|")
    for k, v in ops.most_common(14):
        print("| %s | %.1f%% |" % (k, 100.0 * v / max(nfull + none, 1)))
    print("\nhottest SASS instructions by stall samples:\n\n| samples | executed | avg threads | instruction | top stall |\n|---:|---:|---:|---|---|")
    for rr in sorted(body, key=lambda x: -int(x[sidx["# Samples"]] or 0))[:14]:
        st = sorted(((k, int(rr[sidx[k]] or 0)) for k in stalls), key=lambda x: -x[1])[0]
        print("| %s | %s | %s | `%s` | %s |" % (rr[sidx["# Samples"]], rr[sidx["Instructions Executed"]], rr[sidx["Avg. Threads Executed"]],
                                                rr[sidx["Source"]].strip()[:60], st[0]))


if __name__ == "__main__":
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
