"""bench.py's contract, as far as a box without a GPU can check it: the reference arm's JSON
line (it runs the oracle, the one CPU leg bench.py has), the `config` both arms share, the
workload each GPU count selects, and the clock sampler's "rows inside the timed region" rule."""
import json
import os
import subprocess
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def args(**kw):
    base = dict(gpus=1, steps=1, warmup=0, impl="ours", workload=None, reads=0, genome=0, ksize=0, cpu_sample_reads=2000,
                parity_reads=1000, no_cpu_baseline=True, no_e2e=True, no_parity=True, round_mw=256, table_hint=-1)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_workload_per_gpu_count():
    c2 = bench.workload(args(), 1)
    assert c2["name"] == "C2" and c2["ksize"] == 31 and c2["reads_total"] == 10_000_000 and c2["genome_total"] == 5_000_000
    c3 = bench.workload(args(gpus=8), 8)
    # BASELINE.json configs[2] at eight GPUs: 100 M reads from a 100 Mbp genome, k=21
    assert c3["name"] == "C3" and c3["ksize"] == 21 and c3["reads_total"] == 100_000_000 and c3["genome_total"] == 100_000_000
    assert bench.workload(args(gpus=2), 2)["reads"] == c3["reads"]  # weak scaling: per-GPU work fixed
    cfg = bench.config_of(c3, 8)
    assert cfg["workload"].startswith("C3") and cfg["reads"] == 100_000_000 and "l2_policy" in cfg


def test_algorithmic_bytes():
    # L/(L-k+1) bytes of sequence + 16-B slot read + 8-B count write per k-mer
    assert abs(bench.alg_bytes_per_kmer(150, 31) - (150 / 120 + 24)) < 1e-12
    assert bench.alg_bytes_per_kmer(150, 21, 8) > bench.alg_bytes_per_kmer(150, 21, 1)  # + the routed hash


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-reads", "2000", "--workload", "c3", "--reads", "4000", "--genome", "100000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # same `config` as our arm would print for this command line
    a = args(impl="reference", workload="c3", reads=4000, genome=100000)
    assert d["config"] == bench.config_of(bench.workload(a, 1), 1)


def test_reference_arm_other_ranks_do_nothing(capsys):
    bench.run_reference(args(impl="reference"), rank=1, world=2)
    assert capsys.readouterr().out == ""


def test_clock_rows_outside_the_timed_region_do_not_count():
    c = bench.ClockSampler.__new__(bench.ClockSampler)  # no nvidia-smi here: feed the rows by hand
    c.proc, c.rows = None, []
    row = lambda mhz, cap: [str(mhz), "1965", "700", "0x0", "Not Active", "Not Active", "Not Active", cap]
    now = time.perf_counter()
    c.rows.append((now - 5.0, row(1200, "Active")))      # during warm-up
    c.t_in, c.t_out = now - 1.0, now
    c.rows.append((now - 0.5, row(1965, "Not Active")))  # inside
    c.rows.append((now - 0.2, row(1950, "Not Active")))  # inside
    c.rows.append((now + 3.0, row(600, "Active")))       # long after
    s = c.summary()
    assert s["samples"] == 2 and s["sm_mhz"] == 1957.5 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == []
    c.rows.append((now - 0.1, row(1800, "Active")))
    assert c.summary()["reasons"] == ["sw_power_cap"]
