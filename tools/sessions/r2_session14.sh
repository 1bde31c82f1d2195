#!/usr/bin/env bash
# Round-2 fourteenth 1-GPU call (short): COO streamed kernel with the call-free fast walk (checks only OR a verdict;
# a group with anything unusual is walked again one by one) against the direct kernel, four modes; COO GPU tests.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
: > gpurun_out/r2s14b_coo.jsonl
for m in secded sec7 sed none; do for k in stream; do
  timeout 200 python tools/bench_coo.py --mode $m --kernel $k --steps 50 >> gpurun_out/r2s14b_coo.jsonl 2>> gpurun_out/r2s14b_coo.err; el "bench_coo $m $k rc=$?"
done; done
python - <<'PY'
import json
for l in open("gpurun_out/r2s14b_coo.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"][-30:], d["config"]["kernel"], "it/s %.0f" % d["value"], "spmv ms %.4f" % d["roofline"]["ms_per_launch"], "frac %.3f" % d["roofline"]["frac"], "faults", d["faults"])
PY
tail -3 gpurun_out/r2s14b_coo.err
timeout 300 python -m pytest tests/test_gpu_coo.py -x -q -m gpu > gpurun_out/r2s14b_pytest_coo.txt 2>&1; el "pytest coo rc=$?"; tail -3 gpurun_out/r2s14b_pytest_coo.txt
