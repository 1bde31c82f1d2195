#!/usr/bin/env bash
# What the next session with GPU minutes should run first (one `gpurun --timeout 1500 -- tools/gpu_session.sh` call, 1 GPU):
# everything that was written without a GPU gets seen on hardware, then the A/B kit picks kernel variants.
# Each stage writes under gpurun_out/ and never aborts the following ones.
#   1 tests    pytest -m gpu, plus the opt-in suites (GPU fuzz, cg-coo-cuda CLI)
#   2 ab       tools/ab_variants.sh run  (parity per variant, then spmv timings: none / sec7 / secded / crc32c on 256^3)
#   3 coo      COO spmv timings (tools/time_spmv.py --format coo)
#   4 c5       bench.py --workload powerlaw at the size BASELINE names (N = 2^23, ~1.9e8 non-zeros), sec7 and constraints
#   5 bench    the default bench line
#   7 multi    (not in the default list; needs `gpurun --gpus N`) BASELINE configs[3] and [4] row-partitioned over every
#              GPU count the box offers: poisson2d_blocks sec8, powerlaw sec7 / constraints at the named size
#   6 ncu      (not in the default list) `ncu --set full` of one spmv launch per mode named in $NCU_MODES (default
#              "secded crc32c") with the library named by $ABFT_CUDA_LIB, reports under gpurun_out/ for tools/ncu_summary.py;
#              plus the launch list of a short bench run.  Only after the same commands have run clean without ncu.
# Usage: tools/gpu_session.sh [stage ...]   (default: all)
set -uo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
cd "$ROOT"
mkdir -p gpurun_out
stages=("$@"); [ ${#stages[@]} -eq 0 ] && stages=(tests ab coo c5 bench)
for st in "${stages[@]}"; do
  case "$st" in
    tests)
      python -m pytest tests -q -m gpu > gpurun_out/session_pytest_gpu.txt 2>&1; echo "pytest -m gpu: rc=$?"
      ABFT_GPU_FUZZ=1 ABFT_COO_CLI_TESTS=1 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_cli_coo.py -q -m gpu \
          > gpurun_out/session_pytest_optin.txt 2>&1; echo "opt-in suites: rc=$?"
      tail -3 gpurun_out/session_pytest_gpu.txt gpurun_out/session_pytest_optin.txt ;;
    ab)
      ls build/variants/libabftcuda_base.so > /dev/null 2>&1 || echo "build/variants is empty: run tools/ab_variants.sh build BEFORE gpurun (nvcc cross-compiles in the container)"
      tools/ab_variants.sh run 50 > gpurun_out/session_ab.txt 2>&1; echo "ab kit: rc=$?"; tail -40 gpurun_out/session_ab.txt ;;
    coo)
      # the opt-in grouped element layout (coo.cuh): parity first, then both layouts timed
      ABFT_CUDA_COO_LAYOUT=grouped python -m pytest tests/test_gpu_coo.py -q -m gpu > gpurun_out/session_pytest_coo_grouped.txt 2>&1
      echo "COO tests, grouped layout: rc=$?"; tail -2 gpurun_out/session_pytest_coo_grouped.txt
      : > gpurun_out/session_coo.jsonl
      for layout in caller grouped; do
        ABFT_CUDA_COO_LAYOUT=$layout python tools/time_spmv.py --workload poisson3d --format coo --modes none,sed,sec7,secded \
            --reps 30 --warm 50 --tag "coo_$layout" >> gpurun_out/session_coo.jsonl 2>> gpurun_out/session_coo.err
        echo "coo timings ($layout): rc=$?"
      done
      cat gpurun_out/session_coo.jsonl
      for layout in caller grouped; do
        python tools/bench_coo.py --mode secded --layout $layout --steps 100 --warmup 5 \
            > gpurun_out/session_bench_coo_$layout.json 2> gpurun_out/session_bench_coo_$layout.err
        echo "bench_coo ($layout): rc=$?"; cat gpurun_out/session_bench_coo_$layout.json
      done ;;
    c5)
      for m in sec7 constraints; do
        python bench.py --workload powerlaw --mode $m --steps 100 --warmup 10 --all-modes 0 \
            > gpurun_out/session_c5_$m.json 2> gpurun_out/session_c5_$m.err; echo "C5 $m: rc=$?"; cat gpurun_out/session_c5_$m.json
      done ;;
    bench)
      python bench.py > gpurun_out/session_bench.json 2> gpurun_out/session_bench.err; echo "bench: rc=$?"; cat gpurun_out/session_bench.json ;;
    multi)
      ngpu=$(python -c "import torch; print(torch.cuda.device_count())")
      for n in 2 4 8; do
        [ "$n" -le "$ngpu" ] || continue
        for wl in "poisson2d_blocks sec8" "powerlaw sec7" "powerlaw constraints"; do
          set -- $wl
          python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
              bench.py --gpus $n --workload $1 --mode $2 --steps 100 --warmup 10 --all-modes 0 \
              > gpurun_out/session_multi_${1}_${2}_n$n.json 2> gpurun_out/session_multi_${1}_${2}_n$n.err
          echo "$1 $2 x$n: rc=$?"; tail -1 gpurun_out/session_multi_${1}_${2}_n$n.json | cut -c1-400
        done
      done ;;
    ncu)
      for m in ${NCU_MODES:-secded crc32c}; do
        python tools/prof_spmv.py --mode $m --reps 3 > /dev/null 2>&1 || { echo "prof_spmv $m failed without ncu: skipped"; continue; }
        ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 1 -c 1 -f \
            -o gpurun_out/session_ncu_spmv_$m python tools/prof_spmv.py --mode $m --reps 3 > gpurun_out/session_ncu_spmv_$m.log 2>&1
        echo "ncu spmv $m: rc=$?"
      done
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/session_ncu_launches.csv \
          python bench.py --steps 20 --warmup 3 --all-modes 0 --no-cpu-baseline > gpurun_out/session_ncu_launches.log 2>&1
      echo "ncu launch list: rc=$?" ;;
    *) echo "unknown stage $st" ;;
  esac
done
