#!/usr/bin/env python
"""Kernel experiments: time spmv launches (CUDA events on the context's stream) of one workload
for a list of modes, with the library named by ABFT_CUDA_LIB (default: the in-tree build), and
check result[] against scipy's CSR product (relative, reordered sums).

    python tools/time_spmv.py --workload powerlaw --modes none,sec7 --reps 20
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkg_loader  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="powerlaw")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--modes", default="none")
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--rows", type=int, default=1 << 21)
    ap.add_argument("--nnz", type=int, default=50_000_000)
    ap.add_argument("--tag", default="")
    ap.add_argument("--format", default="csr", choices=["csr", "coo"], help="coo: abft_cuda_coo_spmv (DESIGN.md 9)")
    ap.add_argument("--warm", type=int, default=300)
    a = ap.parse_args()
    import torch
    import scipy.sparse as sp
    pkg = pkg_loader.load()
    mg = pkg.matgen
    if a.workload == "poisson3d":
        t = bench.slab_triplets(mg, a.grid, 0, a.grid ** 3)
    elif a.workload == "poisson2d_blocks":
        t = mg.poisson2d(512, 64)
    else:
        t = mg.powerlaw(a.rows, a.nnz, seed=1)
    nnz = int(t.cols.shape[0])
    vec = np.random.default_rng(1).random(t.N)
    want = sp.csr_matrix((t.vals, (t.rows, t.cols)), shape=(t.N, t.N)) @ vec
    _, b_spmv = bench.alg_bytes(t.N, nnz)
    coo = a.format == "coo"
    if coo:
        b_spmv = 16 * nnz + 16 * t.N   # elements once, vec read once, result written once
    for mode in a.modes.split(","):
        ctx = pkg.CUDAContext(mode, 0)
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N) if coo else ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
        spmv = ctx.coo_spmv if coo else ctx.spmv
        p, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
        ctx.upload(p, vec)
        stream = torch.cuda.ExternalStream(ctx.L.abft_cuda_ctx_stream(ctx.ctx))
        for _ in range(a.warm):          # long enough for the clocks to settle
            spmv(A, p, w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(a.reps):
            spmv(A, p, w)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        got = ctx.download(w)
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        print(json.dumps({"tag": a.tag, "lib": os.path.basename(os.environ.get("ABFT_CUDA_LIB", "libabftcuda.so")),
                          "workload": a.workload, "format": a.format, "mode": mode, "N": t.N, "nnz": nnz, "ms": round(ms, 4),
                          "GBs": round(b_spmv / ms / 1e6, 1), "frac": round(b_spmv / ms / 1e6 / 6537.0, 3),
                          "tiles": (0 if coo else int(ctx.L.abft_cuda_matrix_tiles(A.handle))), "max_rel_err": err,
                          "events": len(ctx.messages)}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
