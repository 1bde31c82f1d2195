#!/usr/bin/env bash
# Round-2 sixth 1-GPU call: mode secdedv on the device, the whole suite after the vector-kernel signature change, bench.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2s6_pytest.txt 2>&1; el "pytest rc=$?"; tail -8 gpurun_out/r2s6_pytest.txt
timeout 500 python bench.py > gpurun_out/r2s6_bench.json 2> gpurun_out/r2s6_bench.err; el "bench rc=$?"; cut -c1-400 gpurun_out/r2s6_bench.json; tail -3 gpurun_out/r2s6_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s6_bench.json").read().strip().splitlines()[-1])
for m,v in d.get("modes",{}).items(): print(m, {k:round(x,3) for k,x in v.items()})
for k,v in d.get("workloads",{}).items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("it_s","iter_frac","spmv_frac","spmv_ms","error","skipped")}, (v.get("gather_rate") or {}).get("frac"))
PY
