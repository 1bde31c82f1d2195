#!/usr/bin/env bash
# Round-2 eleventh 1-GPU call (short): the CRC32C stencil kernel with its tables in registers, 18 chunks cut from the
# whole 88-bit element (the tree's library) against session 10's 19-table form and the LDS kernel of commit bbebe08;
# the crc32c GPU tests on it; one ncu capture.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > gpurun_out/r2s11_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 200 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s11_ab.jsonl 2>> gpurun_out/r2s11_ab.err; el "ab $tag rc=$?"; }
ab lds_bbebe08 crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_head.so
ab shfl19_c3 crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_shfl_all_c3.so
ab shfl18_c3 crc32c
cut -c1-230 gpurun_out/r2s11_ab.jsonl
tail -5 gpurun_out/r2s11_ab.err
timeout 400 python -m pytest tests -x -q -m gpu -k "crc32c or crc or fuzz" > gpurun_out/r2s11_pytest_crc.txt 2>&1; el "pytest -k crc rc=$?"; tail -3 gpurun_out/r2s11_pytest_crc.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
    -o gpurun_out/r2s11_ncu_spmv_crc32c python tools/prof_spmv.py --mode crc32c --reps 3 > gpurun_out/r2s11_ncu_spmv_crc32c.log 2>&1
el "ncu spmv crc32c rc=$?"
