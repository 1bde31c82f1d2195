#!/usr/bin/env bash
# Round-2 third GPU call (1 GPU): the warp-pipeline spmv kernel against the CTA-tiled one.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2s3_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2s3_pytest.txt
ABFT_CUDA_SPMV_KERNEL=cta timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowptr.py -x -q -m gpu > gpurun_out/r2s3_pytest_cta.txt 2>&1; el "pytest (cta kernel) rc=$?"; tail -3 gpurun_out/r2s3_pytest_cta.txt
: > gpurun_out/r2s3_ab.jsonl
ab() { local name=$1 tag=$2 modes=$3; shift 3
  env "$@" ABFT_CUDA_LIB=build/variants/libabftcuda_$name.so timeout 240 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s3_ab.jsonl 2>> gpurun_out/r2s3_ab.err; el "ab $tag rc=$?"; }
ab base cta none,sec7,secded,crc32c ABFT_CUDA_SPMV_KERNEL=cta
ab base warp none,sec7,secded,crc32c
ab blk512 cta512 none,sec7,secded,crc32c ABFT_CUDA_SPMV_KERNEL=cta
ab base cta_rows none,secded ABFT_CUDA_SPMV_KERNEL=cta ABFT_CUDA_PROTECT_ROWS=secded
ab base warp_rows none,secded ABFT_CUDA_PROTECT_ROWS=secded
cut -c1-200 gpurun_out/r2s3_ab.jsonl
for m in secded; do
  python tools/prof_spmv.py --mode $m --reps 3 > /dev/null 2>&1 || { echo "prof_spmv $m failed without ncu: skipped"; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
      -o gpurun_out/r2s3_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/r2s3_ncu_spmv_$m.log 2>&1
  el "ncu spmv $m rc=$?"
done
timeout 500 python bench.py --workloads 0 > gpurun_out/r2s3_bench.json 2> gpurun_out/r2s3_bench.err; el "bench rc=$?"; cut -c1-1200 gpurun_out/r2s3_bench.json; tail -3 gpurun_out/r2s3_bench.err
