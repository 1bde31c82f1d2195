#!/usr/bin/env bash
# Round-2 twelfth 1-GPU call (short): CRC32C stencil kernel, the three-bit last chunk computed on the ALU/FMA pipes
# (multiply + PRMT sign masks + LOP3) instead of an 18th SHFL, against the 18-lookup kernel.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > gpurun_out/r2s12_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 200 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s12_ab.jsonl 2>> gpurun_out/r2s12_ab.err; el "ab $tag rc=$?"; }
ab shfl18_c3 crc32c
ab shfl17_alu3 crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_alu3.so
ab shfl18_c3_again crc32c
ab shfl17_alu3_again crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_alu3.so
cut -c1-230 gpurun_out/r2s12_ab.jsonl
tail -5 gpurun_out/r2s12_ab.err
