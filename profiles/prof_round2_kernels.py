"""Launches of the kernels added in round 2, for ncu: device witness (row fill + memory-log sort), PoW search, the two peer-store
exchanges of a split commit (two in-process ranks on device 0).
    ncu --set full -k regex:'cpu_rows|mem_rows|sort_scatter|pow_grind|rows_to_cols|cols_to_rows' -c 12 -o out python profiles/prof_round2_kernels.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import valida_b200 as vb

rc = np.random.default_rng(7).integers(0, vb.BABYBEAR_P, 480, dtype=np.uint32)
ctx = vb.Context(0)
cfg = vb.StarkConfig(ctx, rc)
log = vb.run_program_log(vb.fib_program(((1 << 18) - 17) // 7))
dm, dp = log.witness_device(ctx)
host = log.traces()
single = vb.prove_machine(cfg, host, device_resident=(dm, dp))
ranks = [vb.Context(0), vb.Context(0)]
vb.comm_init_local(ranks)
cfgs = [vb.StarkConfig(c, rc) for c in ranks]
assert all(p == single for p in vb.run_ranks(lambda r, c: vb.prove_machine(cfgs[r], host), ranks))
for c in ranks:
    c.close()
print("ok")
