#!/usr/bin/env bash
# 2-GPU call: whole GPU suite (multi-GPU tests included), then N=1,2 with and without PDL, then phases.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2d2_pytest_n$N.txt 2>&1; el "pytest -m gpu on $N GPUs rc=$?"; tail -5 gpurun_out/r2d2_pytest_n$N.txt
run() { # tag n env...
  local tag=$1 n=$2; shift 2
  if [ "$n" -eq 1 ]; then
    env "$@" timeout 400 python bench.py --gpus 1 --steps 100 --warmup 5 --all-modes 0 --no-cpu-baseline --workloads 0 > gpurun_out/r2d2_bench_${tag}_n$n.json 2> gpurun_out/r2d2_bench_${tag}_n$n.err
  else
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus $n --steps 100 --warmup 5 --workloads 0 > gpurun_out/r2d2_bench_${tag}_n$n.json 2> gpurun_out/r2d2_bench_${tag}_n$n.err
  fi
  el "bench $tag N=$n rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2d2_bench_${tag}_n$n.json").read().strip().splitlines()[-1])
print("   value %.1f (min %.1f max %.1f) dev %.1f e2e %.1f spmv_frac %.3f iter_frac %.3f" % (d["value"], d["blocks"]["it_s_min"], d["blocks"]["it_s_max"], d["device_resident_loop"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_cg_iteration"]["frac"]))
PY
}
run pdl 1; run nopdl 1 ABFT_CUDA_PDL=0; run pdl 2; run nopdl 2 ABFT_CUDA_PDL=0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 \
    tools/dist_phases.py --iters 60 --out gpurun_out/r2d2_phases_n$N.json > gpurun_out/r2d2_phases_n$N.log 2>&1
el "phase stamps N=$N rc=$?"; grep -A3 "gap\|iteration_us" gpurun_out/r2d2_phases_n$N.log | grep -v "^--" | tr -d '\n' | sed 's/},/}\n/g' | cut -c1-220
