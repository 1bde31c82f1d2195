#!/usr/bin/env bash
# Round-2 ninth 1-GPU call (short): IMAD.HI / IMAD.WIDE issue rates (can the idle FMA pipe take the CRC check's chunk
# shifts?), then the CRC32C spmv with all / half / two of its six shifts per word as multiply-high.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 120 tools/microbench > gpurun_out/r2s9_microbench.txt 2>&1; el "microbench rc=$?"; cat gpurun_out/r2s9_microbench.txt
: > gpurun_out/r2s9_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 200 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s9_ab.jsonl 2>> gpurun_out/r2s9_ab.err; el "ab $tag rc=$?"; }
ab shipped crc32c
for v in crcmulhi_all crcmulhi_odd crcmulhi_two; do ab $v crc32c ABFT_CUDA_LIB=build/variants/libabftcuda_$v.so; done
ab shipped_again crc32c
cut -c1-230 gpurun_out/r2s9_ab.jsonl
tail -5 gpurun_out/r2s9_ab.err
