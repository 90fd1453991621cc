#!/usr/bin/env python
"""bench.py — trace rows/sec proven (Fibonacci, BASELINE.json metric) on N B200s of one node.

A "step" is one full Machine::prove() of the workload (LDE + Keccak Merkle commits + LogUp perm trace
+ quotient + FRI opening), from traces to CBOR proof bytes.
  value  : whole-job rows/s with the traces already resident in HBM (vgpu_prove_device)
  e2e    : the same metric through the reference-facing C-ABI call with HOST buffers
           (vgpu_prove: H2D of the pinned traces + D2H of the proof inside the timed region)
  N > 1  : the path shards by independent proofs (no data-path collective): every rank proves its own
           trace; value = rows of all ranks / max-over-ranks time ("scaling": "weak").
  --impl reference : the CPU restatement of the reference prover (oracle/, all host threads) on a
           bounded sample of the same workload; rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FIB_N = {22: 599183, 20: 149794, 18: 37447, 17: 9360 * 2, 16: 9360, 15: 2339, 12: 582, 8: 25}   # log2(CPU rows) -> n (cycles = 17 + 7n)


def fib_n_for_log_rows(log_rows):
    # largest n with 17 + 7n <= 2^log_rows
    return ((1 << log_rows) - 17) // 7


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_ratio(kernel):
    """DRAM bytes moved / algorithmic bytes for one kernel class, from the committed `ncu --set full` raw pages:
    ntt_pass_kernel: profiles/r01_ntt_v5_raw.csv (8 launches over a 2^22 x 16 matrix, 8 B per element per launch);
    compress_layer_kernel: profiles/r01_keccak_big_raw.csv (tree layers of 2^24 and 2^23 nodes, 96 B per node)."""
    import csv

    spec = {"ntt_pass_kernel": ("r01_ntt_v5_raw.csv", "ntt_pass", lambda gx, gy: 8.0 * gx * gy * (1 << 14)),      # a 2^14-element tile per CTA
            "compress_layer_kernel": ("r01_keccak_big_raw.csv", "compress_layer", lambda gx, gy: 96.0 * gx * 128)}  # a node per thread
    if kernel not in spec:
        return None, None
    fname, tag, alg_bytes = spec[kernel]
    path = os.path.join(ROOT, "profiles", fname)
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw, ig, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Grid Size"), hdr.index("Kernel Name")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        tot, alg = 0.0, 0.0
        for r in rows[2:]:
            if tag not in r[ik]:
                continue
            tot += float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            gx, gy = [int(v) for v in r[ig].strip("()").split(",")[:2]]
            alg += alg_bytes(gx, gy)
        return tot / alg, os.path.relpath(path, ROOT)
    except Exception:
        return None, None


def aggregate_throughput(dist, rows_local, ms_local, device=None):
    """Whole-job rows/s over all ranks: sum of rows / max-over-ranks time (each rank proves its own trace)."""
    import torch

    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    r = torch.tensor([float(rows_local)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    return float(r.item()) / (float(t.item()) / 1000.0), float(t.item())


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.15)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if r[1].isdigit()] or [0])
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(self.rows)}


def tune_oracle_threads(orc, vb):
    """Pick the OpenMP thread count that proves a small sample fastest (large core counts oversubscribe the
    oracle's many short parallel regions); returns the chosen count."""
    t = vb.run_program(vb.fib_program(fib_n_for_log_rows(14)), initial_fp=0x1000)
    best, best_n = None, None
    cores = os.cpu_count() or 1
    for n in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        orc.set_threads(n)
        t0 = time.perf_counter()
        pr = orc.prove(t.main, t.preprocessed, debug_checks=False)
        dt = time.perf_counter() - t0
        del pr
        if best is None or dt < best:
            best, best_n = dt, n
    orc.set_threads(best_n)
    return best_n


def run_reference(args, rank):
    """Reference arm: the oracle prover (CPU restatement of the reference) on a bounded sample."""
    if rank != 0:
        return
    import valida_b200 as vb
    from valida_b200 import build as vbuild
    import oracle_binding

    vbuild.build_oracle()
    orc = oracle_binding.Oracle()
    threads = tune_oracle_threads(orc, vb)
    log_rows = args.ref_log_rows
    n = fib_n_for_log_rows(log_rows)
    t = vb.run_program(vb.fib_program(n), initial_fp=0x1000)
    rows = t.main[0].shape[0]
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        pr = orc.prove(t.main, t.preprocessed, debug_checks=False)
        dt = time.perf_counter() - t0
        del pr
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = rows * len(times) / total
    cores = threads
    sample = "Fibonacci n=%d: 2^%d CPU rows (mem 2^%d), full prove per step; %d OpenMP threads (best of a sweep) on %d host cores" % (
        n, log_rows, (t.main[2].shape[0]).bit_length() - 1, threads, os.cpu_count())
    line = {
        "impl": "reference", "metric": "trace rows/sec proven (Fibonacci)", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (BabyBear, 31-bit modular) + ext5", "data": "synthetic",
        "config": {"workload": workload_name(args.log_rows), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the real reference (Rust + un-vendored Plonky3) cannot be built here; this is oracle/, the C++ restatement, OpenMP on all host threads",
    }
    emit(line)


def workload_name(log_rows):
    return "Fibonacci 2^%d-row full prove (LDE+perm+quotient+FRI+Keccak Merkle), BasicMachine 14 chips, blowup 2, 40 queries" % log_rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-rows", type=int, default=22, help="log2 of the CPU-chip trace height (BASELINE config: 22)")
    ap.add_argument("--ref-log-rows", type=int, default=18, help="bounded sample size of the CPU reference arm")
    ap.add_argument("--cpu-baseline-log-rows", type=int, default=18)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded-timeout", type=int, default=240, help="N > 1: seconds the optional split-proof section may take")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line.  Libraries print there too (NCCL's "NCCL version ..." banner comes out of
    # a C printf on rank 0), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to
    # a private duplicate of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import valida_b200 as vb
    from valida_b200 import build as vbuild

    if not os.path.exists(vb.lib_path):
        vbuild.build()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # an explicit (non-default) torch stream: its handle is non-null, so the library enqueues on it and
    # torch.cuda.Event timings on this stream see the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = vb.Context(local_rank, stream=stream.cuda_stream)
    assert stream.cuda_stream != 0
    rc = np.zeros(480, dtype=np.uint32)
    # documented stand-in for the caller's Poseidon RNG: SplitMix64("valida"), 31-bit rejection sampling
    state, k, M = 0x76616C696461, 0, (1 << 64) - 1
    while k < 480:
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        c = z >> 33
        if c < vb.BABYBEAR_P:
            rc[k] = c
            k += 1
    cfg = vb.StarkConfig(ctx, rc)

    n = fib_n_for_log_rows(args.log_rows)
    t0 = time.perf_counter()
    traces = vb.run_program(vb.fib_program(n), initial_fp=0x1000)
    tracegen_s = time.perf_counter() - t0
    rows = traces.main[0].shape[0]
    assert rows == 1 << args.log_rows
    trace_bytes = sum(m.nbytes for m in traces.main) + sum(m.nbytes for m in traces.preprocessed)

    # pinned host copies for the e2e path; device-resident copies for `value`
    pinned = []
    for m in list(traces.main) + list(traces.preprocessed):
        tt = torch.empty(m.shape, dtype=torch.int32, pin_memory=True)
        tt.numpy().view(np.uint32)[...] = m
        pinned.append(tt)

    class PinnedTraces:
        main = [p.numpy().view(np.uint32) for p in pinned[:14]]
        preprocessed = [p.numpy().view(np.uint32) for p in pinned[14:]]

    dm = [ctx.upload(m) for m in traces.main]
    dp = [ctx.upload(m) for m in traces.preprocessed]
    ctx.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the first warm-up step to the end of the timed region (the same load
    # throughout; one nvidia-smi query takes ~0.3 s, the timed region alone would see one or two samples)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        vb.prove_machine(cfg, traces, device_resident=(dm, dp))

    # ---- timed: device-resident, no instrumentation ----
    launches0 = ctx.launch_count
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        proof = vb.prove_machine(cfg, traces, device_resident=(dm, dp))
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    phases = vb.last_prove_phases(ctx)
    value, ms_total_max = aggregate_throughput(dist, rows * args.steps, ms_total, device="cuda")

    # ---- the same K steps again with a CUDA-event pair around every kernel launch (per-kernel roofline) ----
    ctx.set_kernel_timing(True)
    vb.prove_machine(cfg, traces, device_resident=(dm, dp))   # populates the event pool
    ctx.kernel_stats()
    barrier()
    ei0, ei1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ei0.record(stream)
    for _ in range(args.steps):
        vb.prove_machine(cfg, traces, device_resident=(dm, dp))
    ei1.record(stream)
    barrier()
    ms_instr = ei0.elapsed_time(ei1)
    kstats = ctx.kernel_stats()
    ctx.set_kernel_timing(False)

    # ---- timed: end to end through the host-buffer C-ABI call ----
    vb.prove_machine(cfg, PinnedTraces)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        proof_e2e = vb.prove_machine(cfg, PinnedTraces)
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    e2e_value, _ = aggregate_throughput(dist, rows * args.steps, ms_e2e, device="cuda")
    assert proof_e2e == proof

    def make_line(sharded):

        peak, peak_src = peaks()
        kstats_sorted = sorted(kstats, key=lambda k: -k[2])
        kernels = [{"kernel": k[0], "launches_per_step": k[1] / args.steps, "ms_per_step": k[2] / args.steps,
                    "algorithmic_gb_per_step": k[3] / args.steps / 1e9, "achieved_gbs": (k[3] / 1e9) / (k[2] / 1e3) if k[2] > 0 else None} for k in kstats_sorted]
        top = kstats_sorted[0]
        achieved = (top[3] / 1e9) / (top[2] / 1e3)
        roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src, "share_of_step": top[2] / ms_instr, "ms_per_step_instrumented": ms_instr / args.steps}
        # The Keccak kernels are bound by the INT ALU pipe, not by HBM (profiles/r01_summary.md section 4: 122 LOP3 + 58 SHF per
        # round at 63 lanes/clk/SM = 4.32 G Keccak-f/s on this part, 4.30 measured stand-alone): report that ceiling beside the HBM one.
        KECCAK_PEAK_GPERM = 4.32
        keccak = {}
        for name, bytes_per_perm in (("compress_layer_kernel", 96.0), ("fri_leaf_hash_kernel", 72.0)):
            kk = [k for k in kstats if k[0] == name]
            if kk and kk[0][2] > 0:
                g = kk[0][3] / bytes_per_perm / (kk[0][2] / 1e3) / 1e9     # >= 1 permutation per `bytes_per_perm` algorithmic bytes
                keccak[name] = {"achieved_gperm_s": g, "frac_of_alu_ceiling": g / KECCAK_PEAK_GPERM}
        roofline["int_alu_ceiling"] = {"unit": "G Keccak-f/s", "peak": KECCAK_PEAK_GPERM, "kernels": keccak,
                                       "note": "lower bounds: injected layers and multi-block leaves run more permutations than counted",
                                       "ncu": "profiles/r01_keccak_big_raw.csv: sm__inst_executed_pipe_alu 99.8 % (compress_layer_kernel, 2^24 nodes in 3.96 ms = 4.24 G/s), 92.8 % (leaf_hash_kernel)"}
        ratio, ratio_src = ncu_traffic_ratio(top[0])
        if ratio is not None:
            # GB per launch, like `achieved`: the measured DRAM/algorithmic ratio of the committed capture applied to this
            # run's average launch (ncu cannot run inside the timed region)
            roofline["traffic"] = ratio * (top[3] / top[1]) / 1e9
            roofline["algorithmic_gb_per_launch"] = (top[3] / top[1]) / 1e9
            roofline["traffic_source"] = "%s: DRAM read+write = %.3f x algorithmic bytes" % (ratio_src, ratio)
        ntt = [k for k in kstats if k[0] == "ntt_pass_kernel"]
        if ntt:
            a = (ntt[0][3] / 1e9) / (ntt[0][2] / 1e3)
            roofline["ntt_pass"] = {"achieved": a, "frac": a / peak, "unit": "GB/s", "bytes": "8 B per element per pass (read+write)"}

        # ---- second headline figure: BASELINE config 2 — 2^20 x 64 BabyBear NTT + inverse, device resident ----
        ntt_line = None
        if world == 1:
            hh, ww = 1 << 20, 64
            rr = np.arange(hh, dtype=np.uint64)[:, None]
            cc = np.arange(ww, dtype=np.uint64)[None, :]
            x = ((rr * 64 + cc) * 0x9E3779B1 % vb.BABYBEAR_P).astype(np.uint32)     # SURVEY 8(d) config 2 input
            dft = vb.Radix2Dft(ctx)
            dx = ctx.upload(x)
            for _ in range(3):
                dft.dft_batch(dx); dft.idft_batch(dx)
            torch.cuda.synchronize()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            n0.record(stream)
            for _ in range(reps):
                dft.dft_batch(dx); dft.idft_batch(dx)
            n1.record(stream)
            torch.cuda.synchronize()
            ms_pair = n0.elapsed_time(n1) / reps
            roundtrip_ok = bool(np.array_equal(dx.download(), x))
            gbs = 2 * 8.0 * hh * ww / (ms_pair / 1e3) / 1e9      # two transforms, 8 B per element each (read once + write once)
            ntt_line = {"workload": "2^20 x 64 NTT + iNTT (natural order in/out), 256 MiB working set > L2", "ms_forward_plus_inverse": ms_pair,
                        "achieved": gbs, "unit": "GB/s", "frac": gbs / peak, "bytes": "8*h*w per transform", "roundtrip_bit_exact": roundtrip_ok}
            dx.free()
            # SURVEY 8(d) config 2 also asks for the one-column and the CPU-chip-width shapes; they are extras to the line:
            # any failure is recorded here and changes nothing above
            try:
                others = {}
                for w2 in (51, 1):
                    x2 = np.ascontiguousarray(x[:, :w2])
                    d2 = ctx.upload(x2)
                    for _ in range(3):
                        dft.dft_batch(d2); dft.idft_batch(d2)
                    torch.cuda.synchronize()
                    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    m0.record(stream)
                    for _ in range(reps):
                        dft.dft_batch(d2); dft.idft_batch(d2)
                    m1.record(stream)
                    torch.cuda.synchronize()
                    ms2 = m0.elapsed_time(m1) / reps
                    ok2 = bool(np.array_equal(d2.download(), x2))
                    d2.free()
                    g2 = 2 * 8.0 * hh * w2 / (ms2 / 1e3) / 1e9
                    others["2^20 x %d" % w2] = {"ms_forward_plus_inverse": ms2, "achieved": g2, "unit": "GB/s", "frac": g2 / peak, "roundtrip_bit_exact": ok2,
                                               "l2": "working set %d MiB %s L2" % (hh * w2 * 4 >> 20, ">" if hh * w2 * 4 > 126 << 20 else "fits in")}
                ntt_line["other_widths"] = others
            except Exception as exc:   # noqa: BLE001
                ntt_line["other_widths"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            import oracle_binding

            vbuild.build_oracle()
            orc = oracle_binding.Oracle()
            threads = tune_oracle_threads(orc, vb)
            nb = fib_n_for_log_rows(args.cpu_baseline_log_rows)
            tb = vb.run_program(vb.fib_program(nb), initial_fp=0x1000)
            t0 = time.perf_counter()
            ref = orc.prove(tb.main, tb.preprocessed, debug_checks=False)
            dt = time.perf_counter() - t0
            del ref
            cpu_baseline = {"value": tb.main[0].shape[0] / dt, "unit": "rows/s", "cores": threads, "kind": "port",
                            "sample": "Fibonacci n=%d (2^%d CPU rows), one full oracle prove, %.1f s, %d OpenMP threads (best of a sweep) on %d host cores"
                                      % (nb, args.cpu_baseline_log_rows, dt, threads, os.cpu_count())}

        line = {
            "metric": "trace rows/sec proven (Fibonacci)", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (BabyBear, 31-bit modular) + ext5", "data": "synthetic",
            "config": {"workload": workload_name(args.log_rows), "fib_n": n, "trace_bytes": trace_bytes, "proof_bytes": len(proof),
                       "l2": "inputs (%.2f GB of traces, %.1f GB of LDEs) exceed L2" % (trace_bytes / 1e9, 4.5 * trace_bytes / 1e9),
                       "parallelism": "independent proofs per GPU (no data-path collective)" if world > 1 else "single GPU",
                       "host_tracegen_s": tracegen_s},
            "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": trace_bytes, "d2h_bytes_per_step": len(proof)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "ntt": ntt_line,
            "cpu_baseline": cpu_baseline,
            "phases_ms": {p[0]: p[1] for p in phases},
            "kernels": kernels,
            "sharded": sharded,
        }
        return line

    # ---- N > 1: the SAME single proof split across the ranks (column shares for the LDE, leaf/layer shares for the
    # Keccak trees, NCCL exchange over NVLink) — latency of one proof on N GPUs, beside the replica throughput.
    # The figure is extra to the contract line: a failure is reported inside `sharded`, and a watchdog gives the line out
    # without it if the section does not finish (a hung collective must not lose the replica measurement above). ----
    sharded = None
    if dist is not None:
        done = threading.Event()

        def watchdog():
            if not done.wait(args.sharded_timeout):
                if rank == 0:
                    emit(make_line({"error": "split-proof section did not finish within %d s" % args.sharded_timeout}))
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        try:
            ctx.comm_init_from_torch()
            for _ in range(2):
                proof_sh = vb.prove_machine(cfg, traces, device_resident=(dm, dp))
            identical = proof_sh == proof
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            for _ in range(args.steps):
                vb.prove_machine(cfg, traces, device_resident=(dm, dp))
            s1.record(stream)
            barrier()
            ms_sh = s0.elapsed_time(s1)
            sh_phases = vb.last_prove_phases(ctx)
            _, ms_sh_max = aggregate_throughput(dist, rows * args.steps, ms_sh, device="cuda")
            ctx.set_sharding(False)
            sharded = {"what": "ONE proof per step split across %d GPUs (strong scaling of a single proof)" % world,
                       "ms_per_proof": ms_sh_max / args.steps, "rows_per_s": rows * args.steps / (ms_sh_max / 1e3),
                       "proof_bytes_identical": bool(identical), "phases_ms": {k: v for k, v in sh_phases}}
        except Exception as exc:   # noqa: BLE001
            sharded = {"error": "%s: %s" % (type(exc).__name__, exc)}
        finally:
            done.set()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    emit(make_line(sharded))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
