#!/usr/bin/env bash
# Round-2 fifteenth 1-GPU call (short): do the vector kernels' launches flip the SM's L1 / shared-memory split at every
# kernel boundary?  bench.py's plugin path and device loop with the vector kernels asked to keep the spmv kernel's split.
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"; cd "$ROOT"; mkdir -p gpurun_out
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
run() { local tag=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --all-modes 0 --no-cpu-baseline --workloads 0 > gpurun_out/r2s15_bench_$tag.json 2> gpurun_out/r2s15_bench_$tag.err
  el "bench $tag rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/r2s15_bench_$tag.json").read().strip().splitlines()[-1])
print("   value %.1f (min %.1f max %.1f) dev %.1f e2e %.1f spmv_frac %.3f iter_frac %.3f reasons %s" % (d["value"], d["blocks"]["it_s_min"], d["blocks"]["it_s_max"], d["device_resident_loop"]["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_cg_iteration"]["frac"], d["clocks"]["reasons"]))
PY
}
run default; run carve ABFT_CUDA_VEC_CARVEOUT=1; run default2; run carve2 ABFT_CUDA_VEC_CARVEOUT=1
