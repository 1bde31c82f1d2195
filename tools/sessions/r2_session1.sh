#!/usr/bin/env bash
# Round-2 first GPU call (1 GPU): everything round 1 left unseen on hardware, then the kernel A/B.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/r2_pytest_gpu.txt 2>&1; el "pytest -m gpu rc=$?"; tail -25 gpurun_out/r2_pytest_gpu.txt
ABFT_CUDA_LIB=build/variants/libabftcuda_rowsecc.so timeout 300 python -m pytest tests/test_gpu_rowptr.py -q -m gpu > gpurun_out/r2_pytest_rowptr.txt 2>&1; el "rowptr (rowsecc build) rc=$?"; tail -5 gpurun_out/r2_pytest_rowptr.txt
: > gpurun_out/r2_ab.jsonl
ab() { # name modes [env...]
  local name=$1 modes=$2; shift 2
  env "$@" ABFT_CUDA_LIB=build/variants/libabftcuda_$name.so timeout 240 python tools/ab_time.py --tag $name --modes $modes >> gpurun_out/r2_ab.jsonl 2>> gpurun_out/r2_ab.err
  el "ab $name rc=$?"
}
ab base none,sec7,secded,crc32c
for v in uni t1792 uni_t1792 uni_t1536 lastrefill uni_lastrefill uni_minb3; do ab $v none,sec7,secded; done
for v in crcwalk crc6 crcwalk6 crc5 crcwalk5_t1536; do ab $v crc32c; done
ab rowsecc none,secded
ab rowsecc none,secded ABFT_CUDA_PROTECT_ROWS=secded
cat gpurun_out/r2_ab.jsonl | cut -c1-260
timeout 600 tools/gpu_session.sh (round 1, removed) coo > gpurun_out/r2_coo.txt 2>&1; el "coo rc=$?"; tail -30 gpurun_out/r2_coo.txt | cut -c1-400
timeout 600 tools/gpu_session.sh (round 1, removed) c5 > gpurun_out/r2_c5.txt 2>&1; el "c5 rc=$?"; cut -c1-1200 gpurun_out/r2_c5.txt
timeout 400 python bench.py > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; el "bench rc=$?"; cut -c1-1500 gpurun_out/r2_bench0.json
