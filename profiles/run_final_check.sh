mkdir -p gpurun_out
( timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/final_gpu_tests.log
lscpu | grep -i "model name\|socket\|numa\|thread\|^CPU(s)" > gpurun_out/host_cpu.txt 2>&1
timeout 150 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_reference_v2.json 2> gpurun_out/bench_reference_v2.err
OMP_PROC_BIND=close OMP_PLACES=cores timeout 90 python bench.py --impl reference --steps 2 --warmup 3 --ref-extra-log-rows 0 > gpurun_out/bench_reference_bind.json 2> gpurun_out/bench_reference_bind.err
cat gpurun_out/final_gpu_tests.log; cat gpurun_out/host_cpu.txt; python - <<'PY'
import json
for f in ("bench_reference_v2", "bench_reference_bind"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["sizes"], d["thread_sweep_s"])
    except Exception as e:
        print(f, "failed", e)
PY
