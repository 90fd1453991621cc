"""The host witness generator reproduces, bit for bit, the traces recorded in tests/golden/trace_hashes.json (written by
tests/golden/make_trace_hashes.py with the original single-threaded generator): guards the parallel radix sort of the
memory log, the flat cell map and the multi-threaded row loops."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_traces_match_recorded_digests(built):
    import make_trace_hashes

    want = json.load(open(os.path.join(HERE, "golden", "trace_hashes.json")))
    got = make_trace_hashes.compute()
    assert sorted(got) == sorted(want)
    for name in sorted(want):
        assert got[name] == want[name], name


def test_memory_log_sort_is_stable_for_any_thread_count(built):
    """Same digests with 1, 3 and 7 OpenMP threads (chunk boundaries of the radix passes move with the thread count)."""
    import subprocess

    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import make_trace_hashes as m; "
            "import valida_b200 as vb; d = m.digest(vb.run_program(vb.fib_program(9360), initial_fp=0x1000)); print(json.dumps(d, sort_keys=True))"
            % (os.path.dirname(HERE), os.path.join(HERE, "golden")))
    want = json.load(open(os.path.join(HERE, "golden", "trace_hashes.json")))["fib_9360"]
    for threads in ("1", "3", "7"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1]
        assert json.loads(out) == want, threads
