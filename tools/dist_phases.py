#!/usr/bin/env python
"""Where a row-partitioned CG iteration spends its time besides the streams (256^3 secded, one rank per GPU, under
torch.distributed.run): abft_cuda_phases_enable on every rank, `--iters` iterations of the plugin path
(abft_cuda_dist_cg_host_loop) and of the device-resident loop, then the nanosecond globaltimer stamps of all ranks laid
side by side -> per-phase medians, written as JSON (profiles/r2_dist_phases_n<N>.json)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkg_loader  # noqa: E402
import bench  # noqa: E402

SLOTS = 12
NAMES = ["spmv_begin", "spmv_wait_end", "spmv_reduce_enter", "spmv_end", "xr_begin", "xr_reduce_enter", "xr_end", "p_begin", "p_end"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--mode", default="secded")
    ap.add_argument("--out", default="gpurun_out/dist_phases.json")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    pkg = pkg_loader.load()
    args = argparse.Namespace(grid=a.grid, rows=0, nnz=0)
    env = dict(pkg=pkg, torch=torch, dist=dist, rank=rank, world=world, local_rank=lr)
    case = bench.Case(env, args, "poisson3d", a.mode, None)
    L, ctx = case.ctx.L, case.ctx.ctx
    out = {}
    for path in ("plugin", "device_loop"):
        case.start_solve()
        case.iterate(5)
        case.barrier()
        assert L.abft_cuda_phases_enable(ctx, a.iters) == 0
        if path == "plugin":
            case.iterate(a.iters)
        else:
            case.dev_loop(a.iters)
        buf = np.zeros((a.iters, SLOTS), dtype=np.uint64)
        n = L.abft_cuda_phases_read(ctx, buf.ctypes.data, a.iters)
        L.abft_cuda_phases_enable(ctx, 0)
        rows = buf[:n].astype(np.int64)
        allrows = [None] * world
        dist.all_gather_object(allrows, rows)
        if rank == 0:
            R = np.stack(allrows)            # [rank, iteration, slot]
            R = R[:, 2:-1, :]                # first rows: the device loop's set-up reductions / warm-up
            t0 = R[:, :, 0].min(axis=0)      # iteration start = the earliest spmv begin over the ranks
            res = {}

            def stat(x):                     # x: [rank, iteration] in ns -> median over iterations of mean / max over ranks
                return {"mean_over_ranks_us": float(np.median(x.mean(axis=0)) / 1e3), "max_over_ranks_us": float(np.median(x.max(axis=0)) / 1e3)}
            res["spmv_begin_skew"] = stat(R[:, :, 0] - t0)
            res["spmv_wait_for_pushes"] = stat(R[:, :, 1] - R[:, :, 0])
            res["spmv_own_rows"] = stat(R[:, :, 2] - R[:, :, 1])
            res["spmv_allreduce_wait"] = stat(R[:, :, 3] - R[:, :, 2])
            res["gap_spmv_to_calc_xr (host round trip on the plugin path)"] = stat(R[:, :, 4] - R[:, :, 3])
            res["calc_xr_own_rows"] = stat(R[:, :, 5] - R[:, :, 4])
            res["calc_xr_allreduce_wait"] = stat(R[:, :, 6] - R[:, :, 5])
            res["gap_calc_xr_to_calc_p (host round trip on the plugin path)"] = stat(R[:, :, 7] - R[:, :, 6])
            res["calc_p_and_push"] = stat(R[:, :, 8] - R[:, :, 7])
            nxt = R[:, 1:, 0] - R[:, :-1, 8]
            res["gap_calc_p_to_next_spmv"] = stat(nxt)
            per_iter = np.diff(R[:, :, 0].min(axis=0))
            res["iteration_us_median"] = float(np.median(per_iter) / 1e3)
            out[path] = res
            print(path, json.dumps(res, indent=1))
    if rank == 0:
        out["config"] = {"n_gpus": world, "workload": case.workload_str, "iterations": a.iters,
                         "clock": "globaltimer (ns; one clock for all GPUs of a box)"}
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    case.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
