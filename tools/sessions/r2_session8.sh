#!/usr/bin/env bash
# Round-2 eighth 1-GPU call (short): spmv_stream_kernel (producer warp, stage ring, no CTA barrier) against the stencil
# instantiations of spmv_kernel on 256^3; stage count / CTAs per SM variants of it.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > gpurun_out/r2s8_ab.jsonl
ab() { local tag=$1 modes=$2; shift 2
  env "$@" timeout 200 python tools/ab_time.py --tag "$tag" --modes $modes >> gpurun_out/r2s8_ab.jsonl 2>> gpurun_out/r2s8_ab.err; el "ab $tag rc=$?"; }
M=sec7,secded,crc32c
ab barrier $M ABFT_CUDA_SPMV_KERNEL=stencil
ab stream_s3c3 $M
ab stream_s2c4 $M ABFT_CUDA_LIB=build/variants/libabftcuda_s2c4.so
ab stream_s2c3 $M ABFT_CUDA_LIB=build/variants/libabftcuda_s2c3.so
ab stream_s4c2 $M ABFT_CUDA_LIB=build/variants/libabftcuda_s4c2.so
ab barrier_again $M ABFT_CUDA_SPMV_KERNEL=stencil
cut -c1-260 gpurun_out/r2s8_ab.jsonl
tail -5 gpurun_out/r2s8_ab.err
