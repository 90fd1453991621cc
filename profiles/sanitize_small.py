"""A small end-to-end run for compute-sanitizer (memcheck / racecheck): single-GPU prove, device witness, two in-process ranks.
    compute-sanitizer --tool memcheck python profiles/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import valida_b200 as vb
import oracle_binding

orc = oracle_binding.Oracle()
ctx = vb.Context(0)
cfg = vb.StarkConfig(ctx, orc.rc480)
t = vb.run_program(vb.fib_program(25), initial_fp=0x1000)
p = vb.prove_machine(cfg, t)
assert p == orc.prove(t.main, t.preprocessed, debug_checks=False).cbor()
vb.verify_machine(cfg, p, t.preprocessed)
n = ((1 << 13) - 17) // 7
log = vb.run_program_log(vb.fib_program(n))
host = log.traces()
dm, dp = log.witness_device(ctx)
assert all(np.array_equal(d.download(), m) for d, m in zip(dm, host.main))
single = vb.prove_machine(cfg, host, device_resident=(dm, dp))
ranks = [vb.Context(0), vb.Context(0)]
vb.comm_init_local(ranks)
cfgs = [vb.StarkConfig(c, orc.rc480) for c in ranks]
assert all(q == single for q in vb.run_ranks(lambda r, c: vb.prove_machine(cfgs[r], host), ranks))
for c in ranks:
    c.close()
print("sanitize_small ok")
