#!/usr/bin/env python
"""A/B of libabftcuda.so builds on the 256^3 workload, cheap enough to run a dozen variants in one GPU call.

One process per library (ABFT_CUDA_LIB).  The triplets, and the FIRST library's result[] / fault events per mode, are
cached under --cache; every later library must reproduce them bit for bit (result[] of the clean matrix, and the
event list + result[] after --flips seeded single-bit flips) before its timing counts.  Timing = CUDA events on the
context's stream around --reps back-to-back spmv launches after --warm warm-up launches.

    ABFT_CUDA_LIB=build/variants/libabftcuda_uni.so python tools/ab_time.py --tag uni --modes none,sec7,secded
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
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--modes", default="none,sec7,secded,crc32c")
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--warm", type=int, default=200)
    ap.add_argument("--flips", type=int, default=64)
    ap.add_argument("--tag", default="")
    ap.add_argument("--cache", default="/tmp/abft_ab_cache")
    a = ap.parse_args()
    import torch
    pkg = pkg_loader.load()
    mg = pkg.matgen
    os.makedirs(a.cache, exist_ok=True)
    tfile = os.path.join(a.cache, f"p3d_{a.grid}.npz")
    if os.path.exists(tfile):
        z = np.load(tfile)
        cols, rows, vals, N = z["cols"], z["rows"], z["vals"], int(z["N"])
    else:
        t = bench.slab_triplets(mg, a.grid, 0, a.grid ** 3)
        cols, rows, vals, N = t.cols, t.rows, t.vals, t.N
        np.savez(tfile, cols=cols, rows=rows, vals=vals, N=N)
    nnz = int(cols.shape[0])
    vec = np.random.default_rng(1).random(N)
    _, b_spmv = bench.alg_bytes(N, nnz)
    rng = np.random.default_rng(5)
    flip_idx = rng.integers(0, nnz, a.flips)
    flip_bit = rng.integers(0, 96, a.flips)
    for mode in a.modes.split(","):
        ctx = pkg.CUDAContext(mode, 0)
        A = ctx.create_matrix(cols, rows, vals, N)
        p, w = ctx.create_vector(N), ctx.create_vector(N)
        ctx.upload(p, vec)
        stream = torch.cuda.ExternalStream(ctx.L.abft_cuda_ctx_stream(ctx.ctx))
        for _ in range(a.warm):
            ctx.spmv(A, p, w)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(a.reps):
            ctx.spmv(A, p, w)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        clean = ctx.download(w)
        ev_clean, _ = ctx.poll_events(4096)
        for i, b in zip(flip_idx, flip_bit):
            ctx.flip_bit(A, int(i), int(b))
        ctx.spmv(A, p, w)
        faulty = ctx.download(w)
        ev, fatal = ctx.poll_events(4096)
        ev = sorted(ev)
        ref = os.path.join(a.cache, f"ref_{a.grid}_{mode}.npz")
        if os.path.exists(ref):
            z = np.load(ref)
            parity = {"clean": bool(np.array_equal(z["clean"].view(np.uint64), clean.view(np.uint64))),
                      "faulty": bool(np.array_equal(z["faulty"].view(np.uint64), faulty.view(np.uint64), )),
                      "events": bool(np.array_equal(z["ev"], np.asarray(ev, dtype=np.int64).reshape(-1, 3))),
                      "fatal": bool(int(z["fatal"]) == int(fatal)), "against": str(z["tag"])}
        else:
            np.savez(ref, clean=clean, faulty=faulty, ev=np.asarray(ev, dtype=np.int64).reshape(-1, 3), fatal=int(fatal), tag=a.tag)
            parity = {"reference_run": True}
        print(json.dumps({"tag": a.tag, "lib": os.path.basename(os.environ.get("ABFT_CUDA_LIB", "libabftcuda.so")),
                          "mode": mode, "ms": round(ms, 4), "GBs": round(b_spmv / ms / 1e6, 1),
                          "frac": round(b_spmv / ms / 1e6 / 6537.0, 3), "clean_events": len(ev_clean),
                          "fault_events": len(ev), "fatal": bool(fatal), "parity": parity}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
