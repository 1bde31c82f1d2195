#!/usr/bin/env python
"""A few spmv launches of one workload/mode for ncu.
    python tools/prof_spmv.py --workload powerlaw --mode none --reps 3
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkg_loader  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="poisson3d")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--mode", default="secded")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--rows", type=int, default=1 << 21, help="powerlaw: N")
    ap.add_argument("--nnz", type=int, default=50_000_000, help="powerlaw: target non-zeros")
    a = ap.parse_args()
    pkg = pkg_loader.load()
    mg = pkg.matgen
    if a.workload == "poisson3d":
        t = bench.slab_triplets(mg, a.grid, 0, a.grid ** 3)
    elif a.workload == "poisson2d_blocks":
        t = mg.poisson2d(512, 64)
    else:
        t = mg.powerlaw(a.rows, a.nnz, seed=1)
    ctx = pkg.CUDAContext(a.mode, 0)
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    p, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
    ctx.upload(p, np.random.default_rng(1).random(t.N))
    for _ in range(a.reps):
        ctx.spmv(A, p, w)
    print("pw", ctx.dot(p, w), "tiles", ctx.L.abft_cuda_matrix_tiles(A.handle))
    ctx.close()


if __name__ == "__main__":
    main()
