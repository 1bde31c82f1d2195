#!/usr/bin/env bash
# Round-2 second GPU call (1 GPU): the rewritten spmv (block-size knob, uniform walk adopted, new CRC32C check, overlapped
# row-pointer pre-pass) -- parity, A/B, ncu --set full of the two kernels that are below the bar, the new bench.py.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowptr.py tests/test_gpu_fuzz.py tests/test_gpu_cli_coo.py -x -q -m gpu > gpurun_out/r2s2_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2s2_pytest.txt
: > gpurun_out/r2s2_ab.jsonl
ab() { local name=$1 modes=$2; shift 2
  env "$@" ABFT_CUDA_LIB=build/variants/libabftcuda_$name.so timeout 240 python tools/ab_time.py --tag "$name${TAGX:-}" --modes $modes >> gpurun_out/r2s2_ab.jsonl 2>> gpurun_out/r2s2_ab.err; el "ab $name rc=$?"; }
ab base none,sec7,secded,crc32c
ab blk128 none,sec7,secded,crc32c
ab blk64 none,sec7,secded,crc32c
ab crcmulhi crc32c
TAGX=_rows ab base none,secded ABFT_CUDA_PROTECT_ROWS=secded
TAGX=_rows ab blk64 none,secded ABFT_CUDA_PROTECT_ROWS=secded
cut -c1-230 gpurun_out/r2s2_ab.jsonl
for m in secded crc32c; do
  python tools/prof_spmv.py --mode $m --reps 3 > /dev/null 2>&1 || { echo "prof_spmv $m failed without ncu: skipped"; continue; }
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 1 -c 1 -f \
      -o gpurun_out/r2s2_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/r2s2_ncu_spmv_$m.log 2>&1
  el "ncu spmv $m rc=$?"
done
timeout 500 python bench.py > gpurun_out/r2s2_bench.json 2> gpurun_out/r2s2_bench.err; el "bench rc=$?"; cut -c1-3000 gpurun_out/r2s2_bench.json; tail -5 gpurun_out/r2s2_bench.err
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2s2_bench_ref.json 2> gpurun_out/r2s2_bench_ref.err; el "bench ref rc=$?"; cut -c1-1500 gpurun_out/r2s2_bench_ref.json
