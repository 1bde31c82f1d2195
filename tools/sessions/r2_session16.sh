#!/usr/bin/env bash
# Round-2 sixteenth 1-GPU call (short): CRC32C stencil kernel with a separate copy of the row walk for full warps
# (no per-element selects) against the shipped one.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > gpurun_out/r2s16_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 200 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s16_ab.jsonl 2>> gpurun_out/r2s16_ab.err; el "ab $tag rc=$?"; }
ab shipped crc32c
ab fullt crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_crcfullt.so
ab shipped_again crc32c
ab fullt_again crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_crcfullt.so
cut -c1-230 gpurun_out/r2s16_ab.jsonl
tail -5 gpurun_out/r2s16_ab.err
