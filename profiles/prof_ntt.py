"""NTT / coset-LDE launches for ncu: python profiles/prof_ntt.py
(a) coset LDE of a 2^22 x 16 matrix (the ADD chip of the 2^22 workload), (b) 2^20 x 64 NTT + iNTT, natural order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import valida_b200 as vb

ctx = vb.Context(0)
rng = np.random.default_rng(3)
a = ctx.upload(rng.integers(0, vb.BABYBEAR_P, (1 << 22, 16), dtype=np.uint32))
dft = vb.Radix2Dft(ctx)
for _ in range(2):
    lde = dft.coset_lde_batch(a, 1, 31, bit_reversed=True)
    del lde
b = ctx.upload(rng.integers(0, vb.BABYBEAR_P, (1 << 20, 64), dtype=np.uint32))
for _ in range(2):
    f = dft.dft_batch(b)
    g = dft.idft_batch(f)
ctx.synchronize()
print("launches", ctx.launch_count)
