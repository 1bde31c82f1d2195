#!/usr/bin/env bash
# Round-2 very last 1-GPU call (about a minute): the driver's bench line at the closing sources (roofline.traffic filled:
# profiles/r2_spmv_dram_traffic.json carries this tree's hash).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2i_bench_k20.json 2> gpurun_out/r2i_bench_k20.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2i_bench_k20.json
