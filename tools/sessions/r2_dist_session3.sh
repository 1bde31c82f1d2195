#!/usr/bin/env bash
# 2-GPU call (short): the partitioned-CG tests after the CRC32C kernel change (register tables, 3 CTAs per SM).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/r2d3_pytest_dist_n$N.txt 2>&1; el "test_gpu_dist on $N GPUs rc=$?"; tail -4 gpurun_out/r2d3_pytest_dist_n$N.txt
