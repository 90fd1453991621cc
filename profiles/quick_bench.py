"""Uninstrumented + instrumented timing of one workload (development aid): python profiles/quick_bench.py <log_rows> <steps>"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import valida_b200 as vb
log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 22
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = vb.Context(0, stream=stream.cuda_stream)
cfg = vb.StarkConfig(ctx, np.random.default_rng(7).integers(0, vb.BABYBEAR_P, 480, dtype=np.uint32))
n = ((1 << log_rows) - 17) // 7
t = vb.run_program(vb.fib_program(n), initial_fp=0x1000)
dm = [ctx.upload(m) for m in t.main]; dp = [ctx.upload(m) for m in t.preprocessed]
for _ in range(2): vb.prove_machine(cfg, t, device_resident=(dm, dp))
torch.cuda.synchronize()
per = []
for _ in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    vb.prove_machine(cfg, t, device_resident=(dm, dp))
    e1.record(stream); torch.cuda.synchronize()
    per.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
print("per-step (event ms, wall ms):", ["%.1f/%.1f" % p for p in per])
ms = sum(p[0] for p in per) / steps
print("clean: %.1f ms/step  %.2f Mrows/s" % (ms, (1 << log_rows) / ms / 1e3))
print("phases:", ["%s %.1f" % p for p in vb.last_prove_phases(ctx)])
ctx.set_kernel_timing(True)
vb.prove_machine(cfg, t, device_resident=(dm, dp)); ctx.kernel_stats()
vb.prove_machine(cfg, t, device_resident=(dm, dp))
ks = sorted(ctx.kernel_stats(), key=lambda k: -k[2])
print("kernel sum %.1f ms" % sum(k[2] for k in ks))
for k in ks: print("  %-28s n=%5d %8.2f ms  %8.1f GB/s" % (k[0], k[1], k[2], k[3] / 1e9 / (k[2] / 1e3) if k[2] else 0))
