#!/usr/bin/env bash
# Round-2 last 1-GPU call: whole GPU suite at the closing sources, ncu captures of the secded / crc32c spmv for
# profiles/r2_spmv_dram_traffic.json (tagged with the closing hash of csrc/).
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
for m in secded crc32c; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv -s 1 -c 1 -f \
      -o gpurun_out/r2h_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/r2h_ncu_spmv_$m.log 2>&1
  el "ncu spmv $m rc=$?"
done
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/r2h_pytest.txt 2>&1; el "pytest rc=$?"; tail -6 gpurun_out/r2h_pytest.txt
