#!/usr/bin/env bash
# Round-2 tenth 1-GPU call (short): SHFL.IDX issue rate, alone and next to LDS / SHF; the CRC32C spmv with its chunk
# tables in registers (one entry per lane, looked up by SHFL.IDX: no mask, no shift for chunk 0) -- all 19 tables at
# 3 CTAs per SM, 14 or 7 of them at 4.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 120 tools/microbench > gpurun_out/r2s10_microbench.txt 2>&1; el "microbench rc=$?"; grep -E "SHF|LDS|LOP3 " gpurun_out/r2s10_microbench.txt
: > gpurun_out/r2s10_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 200 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s10_ab.jsonl 2>> gpurun_out/r2s10_ab.err; el "ab $tag rc=$?"; }
ab shipped crc32c
for v in shfl_all_c3 shfl_all_c4 shfl_m3m2_c4 shfl_m3m2_c3 shfl_m3_c4; do ab $v crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_$v.so; done
ab shipped_again crc32c
cut -c1-230 gpurun_out/r2s10_ab.jsonl
tail -5 gpurun_out/r2s10_ab.err
