#!/usr/bin/env bash
# Round-2 closing 1-GPU call: whole suite, smoke, the bench lines the driver will ask for, launch list, ncu captures of
# the shipped kernels (-> profiles/r2_final_*, profiles/r2_spmv_dram_traffic.json via tools/ncu_traffic.py).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.txt 2>&1; el "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; el "bench reference arm rc=$?"
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_k20.json 2> gpurun_out/r2f_bench_k20.err; el "bench K=20 rc=$?"; cut -c1-300 gpurun_out/r2f_bench_k20.json
timeout 500 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; el "bench default rc=$?"; cut -c1-300 gpurun_out/r2f_bench_default.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_ncu_launches.csv \
    python bench.py --steps 20 --warmup 5 --blocks 2 --all-modes 0 --no-cpu-baseline --workloads 0 > gpurun_out/r2f_ncu_launches.log 2>&1; el "ncu launch list rc=$?"
for m in secded crc32c; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
      -o gpurun_out/r2f_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/r2f_ncu_spmv_$m.log 2>&1
  el "ncu spmv $m rc=$?"
done
