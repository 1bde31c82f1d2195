#!/usr/bin/env bash
# Round-2 fifth GPU call (1 GPU): everything green again after the revert; gathers-first / local-OR isolated.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2s5_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2s5_pytest.txt
: > gpurun_out/r2s5_ab.jsonl
ab() { local name=$1 tag=$2 modes=$3; shift 3
  env "$@" ABFT_CUDA_LIB=build/variants/libabftcuda_$name.so timeout 240 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s5_ab.jsonl 2>> gpurun_out/r2s5_ab.err; el "ab $tag rc=$?"; }
ab base base none,sec7,secded,crc32c
ab nogf nogf none,sec7,secded
ab localor localor sec7,secded
ab nogf_localor nogf_localor sec7,secded
ab base base_rows none,secded ABFT_CUDA_PROTECT_ROWS=secded
cut -c1-200 gpurun_out/r2s5_ab.jsonl
timeout 500 python bench.py > gpurun_out/r2s5_bench.json 2> gpurun_out/r2s5_bench.err; el "bench rc=$?"; cut -c1-600 gpurun_out/r2s5_bench.json; tail -3 gpurun_out/r2s5_bench.err
