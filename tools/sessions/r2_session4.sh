#!/usr/bin/env bash
# Round-2 fourth GPU call (1 GPU): split tile refill + gathers first, host polling, compiled host loop.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2s4_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2s4_pytest.txt
: > gpurun_out/r2s4_ab.jsonl
ab() { local name=$1 tag=$2 modes=$3; shift 3
  env "$@" ABFT_CUDA_LIB=build/variants/libabftcuda_$name.so timeout 240 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s4_ab.jsonl 2>> gpurun_out/r2s4_ab.err; el "ab $tag rc=$?"; }
ab prev prev none,sec7,secded,crc32c ABFT_CUDA_SPMV_KERNEL=cta
ab base new none,sec7,secded,crc32c
ab base new_rows none,secded ABFT_CUDA_PROTECT_ROWS=secded
cut -c1-200 gpurun_out/r2s4_ab.jsonl
python tools/prof_spmv.py --mode secded --reps 3 > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
      -o gpurun_out/r2s4_ncu_spmv_secded python tools/prof_spmv.py --mode secded --reps 3 > gpurun_out/r2s4_ncu_spmv_secded.log 2>&1
el "ncu spmv secded rc=$?"
timeout 500 python bench.py --workloads 0 > gpurun_out/r2s4_bench.json 2> gpurun_out/r2s4_bench.err; el "bench rc=$?"; cut -c1-1500 gpurun_out/r2s4_bench.json; tail -3 gpurun_out/r2s4_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 --all-modes 0 --no-cpu-baseline > gpurun_out/r2s4_bench20.json 2> gpurun_out/r2s4_bench20.err; el "bench20 rc=$?"; cut -c1-1500 gpurun_out/r2s4_bench20.json; tail -3 gpurun_out/r2s4_bench20.err
