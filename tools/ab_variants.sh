#!/usr/bin/env bash
# Kernel A/B kit: build libabftcuda.so variants with compile-time knobs and time spmv with each (one GPU).
#
#   tools/ab_variants.sh build            # here (no GPU): nvcc cross-compiles into build/variants/ (travels with gpurun)
#   tools/ab_variants.sh run [reps]       # on the GPU box: parity tests + spmv timings per variant -> gpurun_out/ab_variants.jsonl
#                                         # (AB_ONLY="base uni ..." restricts the list; ~1.5 min of GPU per variant)
#
# Variants (name:flags).  `base` is the shipped configuration.  Knobs: spmv.cuh (ABFT_WALK_UNIFORM, ABFT_TILE_STATIC,
# ABFT_TILE_DYN, ABFT_SPMV_MINB, ABFT_ELEM_BATCH, ABFT_CRC_CHUNK, ABFT_CRC_WALK, ABFT_SPMV_LASTREFILL; ABFT_ROWS_ECC in rowsecc.cuh).  A variant is only ever adopted after `run` shows it parity-green
# and faster; numbers taken here are kernel timings (CUDA events), not bench values.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build/variants"
VARIANTS=(
  "base:"
  "uni:-DABFT_WALK_UNIFORM=1"
  "uni_t1536:-DABFT_WALK_UNIFORM=1 -DABFT_TILE_STATIC=1536"
  "t1536:-DABFT_TILE_STATIC=1536"
  "t1792:-DABFT_TILE_STATIC=1792"
  "uni_t1792:-DABFT_WALK_UNIFORM=1 -DABFT_TILE_STATIC=1792"
  "uni_minb3:-DABFT_WALK_UNIFORM=1 -DABFT_SPMV_MINB=3"
  "crc5:-DABFT_CRC_CHUNK=5"
  "crc5_t1536:-DABFT_CRC_CHUNK=5 -DABFT_TILE_STATIC=1536"
  "crc6:-DABFT_CRC_CHUNK=6"
  "crcwalk:-DABFT_CRC_WALK=1"
  "crcwalk6:-DABFT_CRC_WALK=1 -DABFT_CRC_CHUNK=6"
  "crcwalk5_t1536:-DABFT_CRC_WALK=1 -DABFT_CRC_CHUNK=5 -DABFT_TILE_STATIC=1536"
  "rowsecc:-DABFT_ROWS_ECC=1"
  "lastrefill:-DABFT_SPMV_LASTREFILL=1"
  "uni_lastrefill:-DABFT_WALK_UNIFORM=1 -DABFT_SPMV_LASTREFILL=1"
)
case "${1:-}" in
  build)
    mkdir -p "$OUT"
    for v in "${VARIANTS[@]}"; do
      name="${v%%:*}"; flags="${v#*:}"
      ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
          -ccbin /usr/bin/g++ $flags -Xptxas -v -shared -o "$OUT/libabftcuda_$name.so" \
          "$ROOT/abft-sparse-cg_b200/csrc/abft_cuda.cu" -lcudart > "$OUT/$name.ptxas.log" 2>&1 \
        && echo "built $name" ) &
    done
    wait
    ;;
  run)
    reps="${2:-50}"
    mkdir -p "$ROOT/gpurun_out"
    : > "$ROOT/gpurun_out/ab_variants.jsonl"
    for v in "${VARIANTS[@]}"; do
      name="${v%%:*}"
      # AB_ONLY="base uni t1792": only these (default: all)
      if [ -n "${AB_ONLY:-}" ] && ! [[ " $AB_ONLY " == *" $name "* ]]; then continue; fi
      lib="$OUT/libabftcuda_$name.so"
      [ -f "$lib" ] || { echo "missing $lib (run: tools/ab_variants.sh build)"; continue; }
      # parity first: a variant that is not bit-exact is not timed
      if ABFT_CUDA_LIB="$lib" python -m pytest "$ROOT/tests/test_gpu_parity.py" "$ROOT/tests/test_gpu_rowptr.py" -x -q -m gpu > "$ROOT/gpurun_out/ab_parity_$name.txt" 2>&1; then
        ABFT_CUDA_LIB="$lib" python "$ROOT/tools/time_spmv.py" --workload poisson3d --modes none,sec7,secded,crc32c \
            --reps "$reps" --tag "$name" >> "$ROOT/gpurun_out/ab_variants.jsonl"
      else
        echo "{\"tag\": \"$name\", \"parity\": \"FAILED\"}" >> "$ROOT/gpurun_out/ab_variants.jsonl"
      fi
    done
    cat "$ROOT/gpurun_out/ab_variants.jsonl"
    ;;
  *) echo "usage: $0 build | run [reps]"; exit 2 ;;
esac
