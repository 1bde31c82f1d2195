#!/usr/bin/env python
"""bench.py's measurement for the COO target (cg-coo): one JSON line in the same vocabulary.

    python tools/bench_coo.py [--mode secded] [--grid 256] [--steps 100] [--warmup 5] [--layout caller|grouped]

A step = one CG iteration through the C ABI as cg.cpp drives it (abft_cuda_coo_spmv, dot, calc_xr, calc_p; two
host-visible scalars; there is no device-resident loop for COO; with the streamed kernel the dot is served by the spmv's
own reduction).  `value` = steps / device time (CUDA events on the
context's stream) with operands resident; `e2e` = a whole solve from pinned host buffers (b, x uploaded, x downloaded,
all timed); `roofline` = the COO spmv kernel pair (group walk + stray fix-up) launched back to back, against its
algorithmic bytes: 16 B per element (the 128-bit element, read once) + 8 B per entry of vec (read once) + 8 B per
entry of result (written once).  `perm` / `colptr` (4 B per element / per column) are the implementation's, and are
reported as overhead, not counted.  Results: profiles/r2_coo_stream_kernel.jsonl.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pkg_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="secded", choices=["none", "constraints", "sed", "sec7", "sec8", "secded"])
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--layout", default="grouped", choices=["caller", "grouped"])
    ap.add_argument("--kernel", default="stream", choices=["stream", "direct"],
                    help="grouped layout: the TMA-staged kernel (coo_stream.cuh) or coo_spmv_kernel")
    a = ap.parse_args()
    os.environ.pop("ABFT_CUDA_COO_KERNEL", None)
    if a.kernel == "direct":
        os.environ["ABFT_CUDA_COO_KERNEL"] = "direct"
    if a.layout == "caller":
        os.environ["ABFT_CUDA_COO_LAYOUT"] = "caller"
    else:
        os.environ.pop("ABFT_CUDA_COO_LAYOUT", None)
    import torch
    pkg = pkg_loader.load()
    n = a.grid
    N = n ** 3
    t = bench.slab_triplets(pkg.matgen, n, 0, N)
    nnz = int(t.cols.shape[0])
    peak, peak_src = bench.peaks()
    b_spmv = 16 * nnz + 16 * N
    b_iter = b_spmv + 72 * N                         # + calc_xr 48 N + calc_p 24 N (the separate dot re-reads 16 N: not counted)
    ctx = pkg.CUDAContext(a.mode, 0)
    t0 = time.time()
    A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
    ctx.sync()
    t_create = time.time() - t0
    stream = torch.cuda.ExternalStream(ctx.L.abft_cuda_ctx_stream(ctx.ctx))
    vb, x, r, p, w = (ctx.create_vector(N) for _ in range(5))
    rhs = np.random.default_rng(1).random(N)
    ctx.upload(vb, rhs)

    def iterations(k, rr):
        for _ in range(k):
            ctx.coo_spmv(A, p, w)
            pw = ctx.dot(p, w)
            rr_new = ctx.calc_xr(x, r, p, w, rr / pw)
            ctx.calc_p(p, r, rr_new / rr)
            rr = rr_new
        return rr

    def start():
        ctx.upload(x, np.zeros(N))
        ctx.copy_vector(r, vb)
        ctx.copy_vector(p, r)
        return ctx.dot(r, r)

    sampler = bench.ClockSampler(0)
    rr = iterations(a.warmup, start())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    t_begin = time.time()
    e0.record(stream)
    iterations(a.steps, rr)
    e1.record(stream)
    torch.cuda.synchronize()
    t_end = time.time()
    launches = ctx.launch_count() - launches0
    ms_step = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop(t_begin, t_end)

    # the spmv kernel pair alone, back to back (elements 1.9 GB >> L2)
    for _ in range(5):
        ctx.coo_spmv(A, p, w)
    torch.cuda.synchronize()
    e0.record(stream)
    reps = 30
    for _ in range(reps):
        ctx.coo_spmv(A, p, w)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_spmv = e0.elapsed_time(e1) / reps

    # e2e: host buffers in, host buffer out
    hb = ctx.map_vector(vb)
    hx = ctx.map_vector(x)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    hx[:] = 0.0
    ctx.unmap_vector(vb, hb)
    ctx.unmap_vector(x, hx)
    ctx.copy_vector(r, vb)
    ctx.copy_vector(p, r)
    iterations(a.steps, ctx.dot(r, r))
    out = ctx.map_vector(x)
    checksum = float(out[0])
    e2e_s = time.perf_counter() - w0
    events, fatal = ctx.poll_events()

    line = {"metric": "cg_iterations_per_sec", "value": 1e3 / ms_step, "unit": "it/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"poisson3d_{n} as COO (cg-coo: 128-bit elements; N={N}, nnz={nnz}), mode {a.mode}, -c 0 -i K",
                       "layout": a.layout, "kernel": a.kernel if a.layout == "grouped" else "direct", "l2": f"inputs larger than L2 (element stream {16 * nnz / 1e6:.0f} MB per iteration)",
                       "abi": "C ABI of include/abft_cuda.h via ctypes; host-driven loop (coo_spmv, dot, calc_xr, calc_p)"},
            "e2e": {"value": a.steps / e2e_s, "unit": "it/s", "h2d_bytes_per_step": 16.0 * N / a.steps,
                    "d2h_bytes_per_step": 8.0 * N / a.steps, "what": "b,x uploaded from pinned host buffers + K iterations + x downloaded"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": ("coo_spmv_stream_kernel" if a.layout == "grouped" and a.kernel == "stream" else "coo_spmv_kernel") + f"<{a.mode}> (+ stray fix-up)", "achieved": b_spmv / ms_spmv / 1e6,
                         "peak": peak, "unit": "GB/s", "frac": b_spmv / ms_spmv / 1e6 / peak, "traffic": None,
                         "peak_source": peak_src, "bytes_per_launch": b_spmv, "ms_per_launch": ms_spmv,
                         "implementation_overhead_bytes": 4 * nnz * (a.layout == "caller") + 4 * N},
            "roofline_cg_iteration": {"bound": "hbm", "bytes_per_iteration": b_iter, "achieved": b_iter / ms_step / 1e6,
                                      "peak": peak, "unit": "GB/s", "frac": b_iter / ms_step / 1e6 / peak},
            "clocks": clocks, "setup": {"create_matrix_s": t_create}, "faults": len(events), "x0": checksum}
    print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
