"""The CPU-only parts of bench.py (no GPU): the reference arm end to end on a small sample — one JSON line with the keys the
driver reads — and the `cpu_baseline` leg of the GPU arm as a function."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--workload", "fib",
                        "--log-rows", "12", "--ref-log-rows", "10", "--ref-extra-log-rows", "8,12"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "rows/s" and d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "2^10" in d["cpu_baseline"]["sample"]
    assert set(d["sizes"]) == {"2^10", "2^8", "2^12"} and all("rows_per_s" in v for v in d["sizes"].values())
    assert d["config"]["full_workload_measured_once"] is True and d["config"]["same_config"] is False
    assert d["host"]["usable"] >= 1 and d["host"]["affinity"] >= 1
    assert d["config1_prove_fibonacci_n25"]["ms_per_proof_1_threads"] > 0           # BASELINE.json configs[0]
    assert d["config2_ntt"]["roundtrip_bit_exact"] is True and d["config2_ntt"]["achieved"] > 0      # configs[1] on the CPU
    # the product library is not loaded in the arm's own process (its traces arrive as files from a child process)
    assert "libvalida_b200" not in r.stderr


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_baseline_leg():
    # in a child process: the leg changes process-wide allocator settings (it is the last thing bench.py does)
    code = ("import sys, json; sys.path.insert(0, %r); import bench; bench.host_cpus(); print(json.dumps(bench.cpu_baseline_leg('fib', 9)))" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["kind"] == "port" and d["unit"] == "rows/s" and d["value"] > 0 and d["cores"] >= 1
    assert any(len(v) == 2 for v in d["thread_sweep_s"].values())          # the fastest thread count ran twice
