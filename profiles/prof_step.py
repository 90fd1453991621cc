"""One prove step for profiling (ncu wraps this): python profiles/prof_step.py <log_rows> [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import valida_b200 as vb

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = vb.Context(0)
cfg = vb.StarkConfig(ctx, np.random.default_rng(7).integers(0, vb.BABYBEAR_P, 480, dtype=np.uint32))
n = ((1 << log_rows) - 17) // 7
t = vb.run_program(vb.fib_program(n), initial_fp=0x1000)
dm = [ctx.upload(m) for m in t.main]; dp = [ctx.upload(m) for m in t.preprocessed]
for _ in range(steps):
    p = vb.prove_machine(cfg, t, device_resident=(dm, dp))
print("proof bytes", len(p), "launches", ctx.launch_count)
