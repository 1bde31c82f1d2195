#!/usr/bin/env bash
# Round-2 seventh 1-GPU call (short): stencil instantiations vs the general kernel, row-pointer pre-pass on the upper
# half of the CTA; ncu captures for the record (secded, crc32c on 256^3; sec7 on the power-law matrix at its named size).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > gpurun_out/r2s7_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 240 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s7_ab.jsonl 2>> gpurun_out/r2s7_ab.err; el "ab $tag rc=$?"; }
ab generic none,sec7,secded,crc32c,secdedv ABFT_CUDA_SPMV_KERNEL=generic
ab stencil sec7,secded,crc32c,secdedv
ab stencil_rows none,secded ABFT_CUDA_PROTECT_ROWS=secded
cut -c1-200 gpurun_out/r2s7_ab.jsonl
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_vector_protection.py tests/test_gpu_rowptr.py -x -q -m gpu > gpurun_out/r2s7_pytest.txt 2>&1; el "pytest (parity, vectors, rowptr) rc=$?"; tail -3 gpurun_out/r2s7_pytest.txt
for m in secded crc32c; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
      -o gpurun_out/r2s7_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/r2s7_ncu_spmv_$m.log 2>&1
  el "ncu spmv $m rc=$?"
done
timeout 400 ncu --set full --clock-control none -k regex:spmv -s 1 -c 1 -f \
    -o gpurun_out/r2s7_ncu_spmv_powerlaw_sec7 python tools/prof_spmv.py --workload powerlaw --rows 8388608 --nnz 200000000 --mode sec7 --reps 3 > gpurun_out/r2s7_ncu_spmv_powerlaw.log 2>&1
el "ncu spmv powerlaw rc=$?"
