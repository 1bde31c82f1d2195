#!/usr/bin/env bash
# Round-2 closing 1-GPU call, second edition (after the CRC32C register tables and the streamed COO kernel): whole
# suite, smoke, the bench lines the driver will ask for, launch list, ncu captures of the shipped kernels
# (-> profiles/r2_final2_*, profiles/r2_spmv_dram_traffic.json via tools/ncu_traffic.py).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2g_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2g_pytest.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.txt 2>&1; el "smoke rc=$?"; tail -2 gpurun_out/r2g_smoke.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; el "bench reference arm rc=$?"
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_k20.json 2> gpurun_out/r2g_bench_k20.err; el "bench K=20 rc=$?"; cut -c1-300 gpurun_out/r2g_bench_k20.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_ncu_launches.csv \
    python bench.py --steps 20 --warmup 5 --blocks 2 --all-modes 0 --no-cpu-baseline --workloads 0 > gpurun_out/r2g_ncu_launches.log 2>&1; el "ncu launch list rc=$?"
for m in secded crc32c; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
      -o gpurun_out/r2g_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/r2g_ncu_spmv_$m.log 2>&1
  el "ncu spmv $m rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:coo_spmv_stream -s 1 -c 1 -f \
    -o gpurun_out/r2g_ncu_coo_stream_secded python tools/bench_coo.py --mode secded --steps 3 --warmup 1 > gpurun_out/r2g_ncu_coo_stream_secded.log 2>&1
el "ncu coo stream secded rc=$?"
timeout 200 python tools/bench_coo.py --mode secded --kernel direct --steps 50 > gpurun_out/r2g_coo_direct.json 2> gpurun_out/r2g_coo_direct.err; el "bench_coo direct rc=$?"
