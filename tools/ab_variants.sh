#!/usr/bin/env bash
# Kernel A/B kit: build libabftcuda.so variants with compile-time knobs (here, no GPU: nvcc cross-compiles into
# build/variants/, which travels with gpurun); tools/ab_time.py then times each on the GPU box and holds its results
# to the first library's, bit for bit.  A variant is adopted only when it is parity-green and faster; the knob is then
# removed (none is left in the source at the end of round 2).
#   tools/ab_variants.sh build ["name:flags" ...]
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build/variants"
VARIANTS=("base:")
case "${1:-}" in
  build)
    shift
    [ $# -gt 0 ] && VARIANTS=("$@")
    mkdir -p "$OUT"
    for v in "${VARIANTS[@]}"; do
      name="${v%%:*}"; flags="${v#*:}"
      ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
          -ccbin /usr/bin/g++ $flags -Xptxas -v -shared -o "$OUT/libabftcuda_$name.so" \
          "$ROOT/abft-sparse-cg_b200/csrc/abft_cuda.cu" -lcudart > "$OUT/$name.ptxas.log" 2>&1 \
        && echo "built $name" || { echo "FAILED $name"; tail -5 "$OUT/$name.ptxas.log"; } ) &
    done
    wait
    ;;
  *) echo "usage: $0 build [name:flags ...]"; exit 2 ;;
esac
