#!/bin/bash
# One multi-GPU session (gpurun --gpus N): the process-per-GPU test at 4 ranks, then the split-proof bench on the three
# workloads of BASELINE.json configs 3-5.  Usage: profiles/run_multi_gpu.sh <N> [steps] [warmup]
N=${1:-8}; STEPS=${2:-3}; WARM=${3:-2}
export OMP_NUM_THREADS=$(( $(nproc) / N ))       # torchrun would pin it to 1: the host witness generator uses OpenMP
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
run() {   # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps $STEPS --warmup $WARM $2 > gpurun_out/bench_${N}gpu_$1.json 2> gpurun_out/bench_${N}gpu_$1.err
  echo "$1 rc=$? $(cut -c1-240 gpurun_out/bench_${N}gpu_$1.json)"
  grep -i "error\|Traceback" gpurun_out/bench_${N}gpu_$1.err | grep -v "elastic/errors\|error_file" | head -5
}
run fib22 ""
run fib24 "--workload fib24 --no-replicas"
run config5 "--workload config5 --no-replicas"
