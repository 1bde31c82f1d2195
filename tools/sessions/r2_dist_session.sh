#!/usr/bin/env bash
# Multi-GPU GPU call (gpurun --gpus N): every multi-GPU test on all GPUs of the box, then the bench line at N
# (plus N/2, N/4 ... when asked: SCALE="1 2 4 8"), with phase stamps of a short partitioned solve.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=$(python -c "import torch; print(torch.cuda.device_count())")
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_cli.py -q -m gpu -k "dist or multi_gpu" > gpurun_out/r2_dist_pytest_n$N.txt 2>&1; el "multi-GPU tests on $N GPUs rc=$?"; tail -8 gpurun_out/r2_dist_pytest_n$N.txt
for n in ${SCALE:-$N}; do
  [ "$n" -le "$N" ] || continue
  if [ "$n" -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps ${STEPS:-100} --warmup 5 --all-modes 0 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/r2_dist_bench_n$n.json 2> gpurun_out/r2_dist_bench_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus $n --steps ${STEPS:-100} --warmup 5 $( [ "$n" -eq "$N" ] && echo "${BENCH_EXTRA_TOP:-}" || echo "${BENCH_EXTRA:-}" ) > gpurun_out/r2_dist_bench_n$n.json 2> gpurun_out/r2_dist_bench_n$n.err
  fi
  el "bench N=$n rc=$?"; tail -1 gpurun_out/r2_dist_bench_n$n.json | cut -c1-900; tail -2 gpurun_out/r2_dist_bench_n$n.err | cut -c1-300
done
if [ "$N" -ge 2 ]; then
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 \
      bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_dist_bench_ref_torchrun.json 2> gpurun_out/r2_dist_bench_ref_torchrun.err
  el "reference arm under torchrun rc=$?"; cut -c1-700 gpurun_out/r2_dist_bench_ref_torchrun.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
      tools/dist_phases.py --iters 60 --out gpurun_out/r2_dist_phases_n$N.json > gpurun_out/r2_dist_phases_n$N.log 2>&1
  el "phase stamps N=$N rc=$?"; tail -12 gpurun_out/r2_dist_phases_n$N.log | cut -c1-200
fi
