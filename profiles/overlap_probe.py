"""Does a Keccak tree (INT-ALU bound) overlap with coset LDEs (FMA + ALU) when they run on two streams of one GPU?
python profiles/overlap_probe.py   — prints sequential vs concurrent wall time (development probe)."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import valida_b200 as vb

rng = np.random.default_rng(5)
a = vb.Context(0); b = vb.Context(0)
mem = a.upload(rng.integers(0, vb.BABYBEAR_P, (1 << 24, 14), dtype=np.uint32))     # memory-chip shaped
cpu = b.upload(rng.integers(0, vb.BABYBEAR_P, (1 << 22, 67), dtype=np.uint32))     # cpu + add columns
pa, db = vb.TwoAdicFriPcs(a), vb.Radix2Dft(b)

def tree():
    r, pd = pa.commit_batches([mem]); pd.free()
def ldes():
    for _ in range(2):
        l = db.coset_lde_batch(cpu, 1, 31, bit_reversed=True); b.synchronize(); l.free()

for _ in range(2): tree(); ldes()
def timed(f):
    a.synchronize(); b.synchronize(); t = time.perf_counter(); f(); a.synchronize(); b.synchronize(); return (time.perf_counter() - t) * 1e3
t_tree = min(timed(tree) for _ in range(3)); t_lde = min(timed(ldes) for _ in range(3))
def both():
    th = threading.Thread(target=tree); th.start(); ldes(); th.join()
t_both = min(timed(both) for _ in range(3))
print("tree %.1f ms, ldes %.1f ms, sequential %.1f ms, concurrent %.1f ms" % (t_tree, t_lde, t_tree + t_lde, t_both))
