"""
Host-side pieces of bench.py that need no GPU: the clock-sample parser (nvidia-smi lines -> the `clocks` object, filtered
to the timed window), the host restatement of the tcgen05 gate that only NAMES the Kuf roofline, the algorithmic
operation counts the rooflines are quoted in, and the reference arm's JSON contract on a tiny sample.
"""
import datetime
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _FakeProc:
    def __init__(self, text):
        self.text = text

    def terminate(self):
        pass

    def communicate(self, timeout=None):
        return self.text, ""


LINES = ("1965, 1965, 500.1, Not Active, Not Active, Not Active, Not Active, 2026/10/17 15:31:02.123\n"
         "1900, 1965, 510.1, Not Active, Not Active, Not Active, Active, 2026/10/17 15:31:02.223\n"
         "1200, 1965, 100.1, Active, Not Active, Not Active, Not Active, 2026/10/17 15:31:05.223\n")


def _sampler(t0, t1):
    cs = bench.ClockSampler.__new__(bench.ClockSampler)
    cs.proc, cs.t0, cs.t1 = _FakeProc(LINES), t0, t1
    return cs


def test_clock_samples_are_filtered_to_the_timed_window():
    t0 = datetime.datetime(2026, 10, 17, 15, 31, 2, 100000)
    out = _sampler(t0, t0 + datetime.timedelta(milliseconds=200)).stop()
    assert out["samples"] == 2 and out["window"] == "timed region"
    assert out["sm_mhz"] == 1932.5 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]          # the hw_slowdown sample lies outside the window


def test_clock_samples_fall_back_to_the_whole_run_for_short_regions():
    t0 = datetime.datetime(2026, 10, 17, 15, 31, 3)
    out = _sampler(t0, t0 + datetime.timedelta(milliseconds=1)).stop()
    assert out["samples"] == 3 and out["window"].startswith("warm-up + timed region")
    assert "hw_slowdown" in out["reasons"]


def test_tcgen05_gate_restatement():
    rng = np.random.default_rng(0)
    L, d, M, nz = 16, 8, 3, 5
    X = (np.cumsum(rng.standard_normal((6, L, d)), axis=1) / np.sqrt(L)).reshape(6, -1)
    Z = bench.synth_Z(X, L, d, M, nz)
    ls = bench.lengthscales_for("rbf", d)
    assert bench.tcgen05_takes_kuf("rbf", X, Z, ls, d, M, False)
    assert not bench.tcgen05_takes_kuf("linear", X, Z, ls, d, M, False)          # no exponent: CUDA-core kernel
    assert not bench.tcgen05_takes_kuf("rbf", X, Z, ls, d, M, True)              # low-rank mode
    assert not bench.tcgen05_takes_kuf("rbf", X + 100.0, Z, ls, d, M, False)     # data far from the tensors: gate closes
    assert not bench.tcgen05_takes_kuf("rbf", X, Z, ls, 20, M, False)            # 3 d + 6 > 64 operand slots


def test_algorithmic_operation_counts():
    assert bench.fp32_ops_per_entry("rbf", 8, 5) == 21 and bench.fp32_ops_per_entry("linear", 8, 5) == 17
    assert bench.units_per_step(bench.WORKLOADS["cfg4"]) == 4096 * 4096
    assert bench.units_per_step(bench.WORKLOADS["cfg3"]) == 256 * 4096


def test_reference_arm_contract_on_a_tiny_sample():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-n", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == bench.UNIT
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
