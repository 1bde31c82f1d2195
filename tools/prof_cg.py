#!/usr/bin/env python
"""Short CG run for ncu: builds the grid^3 Poisson matrix and runs a few iterations of the hot path.
    python tools/prof_cg.py --grid 256 --mode secded --iters 4
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
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--mode", default="secded")
    ap.add_argument("--iters", type=int, default=4)
    a = ap.parse_args()
    pkg = pkg_loader.load()
    N = a.grid ** 3
    t = bench.slab_triplets(pkg.matgen, a.grid, 0, N)
    ctx = pkg.CUDAContext(a.mode, 0)
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    b, x, r, p, w = (ctx.create_vector(N) for _ in range(5))
    ctx.upload(b, np.random.default_rng(1).random(N))
    it, hist = ctx.cg(A, b, x, r, p, w, a.iters, 0.0)
    print("iterations", it, "rr", hist[-1], "launches", ctx.launch_count())
    ctx.close()


if __name__ == "__main__":
    main()
