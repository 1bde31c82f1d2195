"""Worker for tests/test_dist_gloo.py and tests/test_gpu_dist.py (run under torch.distributed.run).

mode `cpu`: gloo backend, no GPU.  Exercises the host-side logic of the multi-GPU path -- row
partition, needed-interval computation, exchange plan -- by running the row-partitioned CG with numpy
kernels and point-to-point transfers that follow the plan exactly; anything a rank reads that the plan
did not deliver is NaN, so a wrong plan cannot pass.
mode `gpu`: nccl backend for setup only; the data path is the C ABI's dist_* entry points (peer stores
over NVLink, flags, in-kernel all-reduce).  Compares against the single-rank CPU oracle.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def build(mg, name):
    if name == "p2d":
        return mg.poisson2d(24)
    if name == "p3d":
        return mg.poisson3d(10)
    if name == "blocks":
        return mg.poisson2d(8, 6)
    if name == "powerlaw":
        return mg.powerlaw(800, 12000, seed=3)
    raise ValueError(name)


def main():
    mode, shape, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    import torch
    import torch.distributed as dist
    import pkg_loader
    pkg = pkg_loader.load()
    mg, D = pkg.matgen, pkg.dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "gpu":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    else:
        dist.init_process_group("gloo")

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    t = build(mg, shape)
    b = mg.rhs(t.N)
    rowptr = t.row_pointers().astype(np.int64)
    cuts = D.balanced_row_cuts(rowptr, world)
    r0, r1 = cuts[rank], cuts[rank + 1]
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    tl = mg.Triplets(t.cols[e0:e1], t.rows[e0:e1], t.vals[e0:e1], t.N)
    result = {}

    if mode == "cpu":
        import scipy.sparse as sp
        need = D.needed_intervals(tl.cols, cuts)
        all_need = all_gather(need)
        plan = D.exchange_plan(all_need, rank)
        A = sp.csr_matrix((tl.vals, (tl.rows.astype(np.int64) - r0, tl.cols.astype(np.int64))), shape=(r1 - r0, t.N))
        p_full = np.full(t.N, np.nan)

        def push():
            reqs = []
            for s in range(world):
                lo, hi = plan["push_lo"][s], plan["push_hi"][s]
                if s != rank and hi > lo:
                    assert r0 <= lo and hi <= r1
                    reqs.append(dist.isend(torch.from_numpy(p_full[lo:hi].copy()), s))
            for s in range(world):
                if (plan["wait_mask"] >> s) & 1:
                    lo, hi = all_need[rank][s]
                    buf = torch.empty(hi - lo, dtype=torch.float64)
                    dist.recv(buf, s)
                    p_full[lo:hi] = buf.numpy()
            for q in reqs:
                q.wait()

        def allsum(v):
            tt = torch.tensor([v], dtype=torch.float64)
            dist.all_reduce(tt)
            return float(tt.item())

        x = np.zeros(r1 - r0); r = b[r0:r1].copy()
        p_full[r0:r1] = r
        push()
        rr = allsum(float(r @ r))
        it = 0
        max_it = int(sys.argv[4])          # fixed iteration count: reductions are reordered, so a residual
        while it < max_it:                 # sitting on the threshold could stop one iteration apart
            w = A @ p_full
            assert np.all(np.isfinite(w)), "spmv read an entry the exchange plan did not deliver"
            pw = allsum(float(p_full[r0:r1] @ w))
            alpha = rr / pw
            x += alpha * p_full[r0:r1]; r -= alpha * w
            rr_new = allsum(float(r @ r))
            p_full[r0:r1] = r + (rr_new / rr) * p_full[r0:r1]
            push()
            rr = rr_new
            it += 1
        xs = all_gather(x)
        result = dict(iterations=it, rr=rr, x=np.concatenate(xs).tolist(), cuts=cuts, plan=plan,
                      bytes=D.exchanged_bytes(plan), nnz_local=int(e1 - e0))
    else:
        import oracle_py as O
        gpu_mode = sys.argv[4] if len(sys.argv) > 4 else "secded"
        ctx = pkg.CUDAContext(gpu_mode, int(os.environ["LOCAL_RANK"]))
        cg = D.DistCG(ctx, rank, world, all_gather, tl, cuts, t.N, e0)
        ctx.upload(cg.b, b[r0:r1])
        if len(sys.argv) > 5 and sys.argv[5] == "flip" and rank == world - 1:
            ctx.flip_bit(cg.A, e0 + 7, 33)       # a fault in the last rank's block: global index in the report
        cg.start()
        hist = []
        it = cg.iterate(2000, 1e-10, hist)
        lines = ctx.report_faults()
        x_host_loop = ctx.download(cg.x)
        rr_host_loop = cg.rr
        xs = all_gather(x_host_loop)
        all_lines = all_gather(lines)
        # the device-resident loop over the same partition: same kernels, same scalars -> identical bits
        dev_same = []
        for every in (1, 16):
            ctx.upload(cg.x, np.zeros(r1 - r0))
            it2, rr2 = cg.solve_device(2000, 1e-10, every)
            dev_same.append(bool(it2 == it and rr2 == rr_host_loop and np.array_equal(ctx.download(cg.x), x_host_loop)))
        # and a host-driven solve still works afterwards (sequence numbers were handed back correctly)
        ctx.upload(cg.x, np.zeros(r1 - r0))
        cg.start()
        it3 = cg.iterate(2000, 1e-10)
        dev_same.append(bool(it3 == it and np.array_equal(ctx.download(cg.x), x_host_loop)))
        # the compiled form of that loop (abft_cuda_dist_cg_host_loop: what bench.py times): the same four calls
        ctx.upload(cg.x, np.zeros(r1 - r0))
        cg.start()
        it4 = cg.host_loop(2000, 1e-10)
        dev_same.append(bool(it4 == it and cg.rr == rr_host_loop and np.array_equal(ctx.download(cg.x), x_host_loop)))
        ctx.report_faults()
        result = dict(iterations=it, rr=rr_host_loop, x=np.concatenate(xs).tolist(), hist=hist, lines=all_lines, cuts=cuts,
                      plan=cg.plan, e0_last=e0, dev_same=all_gather(dev_same))
        cg.close()
        ctx.close()
    results = all_gather(dict(iterations=result["iterations"], rr=result["rr"]))
    if rank == 0:
        result["all"] = results
        with open(out_path, "w") as f:
            json.dump(result, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
