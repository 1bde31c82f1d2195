#!/usr/bin/env bash
# Round-2: the COO differential fuzz on the device (6 modes x ABFT_FUZZ_TRIALS random multi-fault trials against the pinned
# COO oracle): first the default 120 per mode, then 2500 per mode.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
ABFT_FUZZ_TRIALS=2500 timeout 75 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu -k coo_oracle > gpurun_out/r2j_pytest_fuzz_coo_2500.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2j_pytest_fuzz_coo_2500.txt
