#!/usr/bin/env bash
# 2-GPU call (short): the unmodified cg.cpp driving several GPUs from one host thread (ABFT_CUDA_NGPUS) after
# map_vector / unmap_vector moved to one pinned host image with concurrent per-GPU slices.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py -q -m gpu -k "multi_gpu or vector_transfers" > gpurun_out/r2d4_pytest_cli_multi_n$N.txt 2>&1; el "multi-GPU CLI tests on $N GPUs rc=$?"; tail -4 gpurun_out/r2d4_pytest_cli_multi_n$N.txt
