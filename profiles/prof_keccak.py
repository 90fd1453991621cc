"""Keccak tree of a memory-chip-sized matrix for ncu: python profiles/prof_keccak.py (commit of a 2^24 x 14 matrix, twice)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import valida_b200 as vb

ctx = vb.Context(0)
m = ctx.upload(np.random.default_rng(9).integers(0, vb.BABYBEAR_P, (1 << 24, 14), dtype=np.uint32))
pcs = vb.TwoAdicFriPcs(ctx)
for _ in range(2):
    root, pd = pcs.commit_batches([m])
    pd.free()
print("launches", ctx.launch_count)
