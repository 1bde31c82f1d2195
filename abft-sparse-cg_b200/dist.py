"""Multi-GPU CG: 1-D row partition + the exchange plan.

Host-side logic only (numpy); the data path is in csrc/ (calc_p_push_kernel, the
peer all-reduce in the reduction finalizer, the flag wait in spmv_kernel).  The
reference has no distributed code at all; the contract is: same results as the
single-GPU / CPU target on the same input, for any rank count.

    cuts            balanced_row_cuts(rowptr, world)       contiguous row blocks balanced by nnz
    need            needed_intervals(cols_local, cuts)     rows of every peer my block's columns touch
    plan            exchange_plan(all_need, rank)          what I push to whom, whom I wait for
    DistCG          the CG loop of cg.cpp:83-118 over the dist_* entry points of the C ABI
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def balanced_row_cuts(rowptr: np.ndarray, world: int) -> list:
    """Row boundaries [c_0=0, ..., c_world=N] so that every block holds ~nnz/world elements."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    N = rowptr.shape[0] - 1
    nnz = int(rowptr[-1])
    cuts = [0]
    for k in range(1, world):
        target = (nnz * k) // world
        r = int(np.searchsorted(rowptr, target, side="left"))
        cuts.append(min(max(r, cuts[-1]), N))
    cuts.append(N)
    return cuts


def even_row_cuts(N: int, world: int) -> list:
    return [(N * k) // world for k in range(world + 1)]


def needed_intervals(cols_local: np.ndarray, cuts: list, mask24: bool = False) -> list:
    """For every rank s: the interval [lo, hi) of s's rows that the columns of my block reference
    ((0, 0) if none).  Columns are global indices (check bits, if any, masked off)."""
    cols = np.asarray(cols_local, dtype=np.int64)
    if mask24:
        cols = cols & 0x00FFFFFF
    out = []
    for s in range(len(cuts) - 1):
        lo, hi = cuts[s], cuts[s + 1]
        if cols.size == 0 or hi <= lo:
            out.append((0, 0))
            continue
        sel = cols[(cols >= lo) & (cols < hi)]
        out.append((int(sel.min()), int(sel.max()) + 1) if sel.size else (0, 0))
    return out


def exchange_plan(all_need: list, rank: int) -> dict:
    """all_need[r][s] = interval of s's rows that rank r reads.  ->  what `rank` pushes and waits for."""
    world = len(all_need)
    push_lo = [0] * world
    push_hi = [0] * world
    wait_mask = 0
    for s in range(world):
        if s == rank:
            continue
        lo, hi = all_need[s][rank]          # rows of mine that s needs
        push_lo[s], push_hi[s] = int(lo), int(hi)
        lo2, hi2 = all_need[rank][s]        # rows of s that I need => s pushes to me
        if hi2 > lo2:
            wait_mask |= 1 << s
    return dict(push_lo=push_lo, push_hi=push_hi, wait_mask=wait_mask)


def exchanged_bytes(plan: dict) -> int:
    return 8 * sum(max(0, h - l) for l, h in zip(plan["push_lo"], plan["push_hi"]))


class DistCG:
    """One rank of the row-partitioned solver.  `all_gather(obj) -> [obj of rank 0, ...]` is the only
    transport this class needs (torch.distributed.all_gather_object in bench.py): it carries the IPC
    handles and the plan at setup, never the data path."""

    def __init__(self, ctx, rank: int, world: int, all_gather, triplets_local, cuts, N_global: int, nnz_offset: int):
        self.ctx, self.L = ctx, ctx.L
        self.rank, self.world, self.N = rank, world, N_global
        self.r0, self.r1 = cuts[rank], cuts[rank + 1]
        self.nloc = self.r1 - self.r0
        t = triplets_local
        self.A = ctx.create_matrix_part(t.cols, t.rows, t.vals, self.nloc, N_global, self.r0, nnz_offset)
        # vectors: full-length p on every rank, local slices of the rest
        self.p_full = ctx.create_vector(N_global)
        self.p = ctx.view_vector(self.p_full, self.r0, self.nloc)
        self.b, self.x, self.r, self.w = (ctx.create_vector(self.nloc) for _ in range(4))
        # exchange window + handles
        self.comm = self.L.abft_cuda_comm_create(ctx.ctx, rank, world)
        if not self.comm:
            raise RuntimeError("comm_create failed")
        hw = C.create_string_buffer(64)
        hv = C.create_string_buffer(64)
        ctx._ck(self.L.abft_cuda_comm_window_handle(ctx.ctx, self.comm, hw), "comm_window_handle")
        ctx._ck(self.L.abft_cuda_vector_ipc_handle(ctx.ctx, self.p_full.handle, hv), "vector_ipc_handle")
        need = needed_intervals(t.cols, cuts)
        everyone = all_gather(dict(win=hw.raw, vec=hv.raw, need=need))
        ctx._ck(self.L.abft_cuda_comm_attach_self(ctx.ctx, self.comm, self.p_full.handle), "attach_self")
        for s, info in enumerate(everyone):
            if s != rank:
                ctx._ck(self.L.abft_cuda_comm_attach_peer(ctx.ctx, self.comm, s, info["win"], info["vec"]), "attach_peer")
        self.plan = exchange_plan([e["need"] for e in everyone], rank)
        lo = (C.c_int * world)(*self.plan["push_lo"])
        hi = (C.c_int * world)(*self.plan["push_hi"])
        ctx._ck(self.L.abft_cuda_comm_set_plan(ctx.ctx, self.comm, lo, hi, self.plan["wait_mask"], self.r0), "set_plan")
        all_gather(0)    # everyone has opened everyone's handles before the first push

    # ---- the dist_* hot path ----
    def spmv(self):
        self.ctx._ck(self.L.abft_cuda_dist_spmv(self.ctx.ctx, self.comm, self.A.handle, self.p_full.handle, self.w.handle), "dist_spmv")

    def dot(self, a, b) -> float:
        out = C.c_double(0.0)
        self.ctx._ck(self.L.abft_cuda_dist_dot(self.ctx.ctx, self.comm, a.handle, b.handle, C.byref(out)), "dist_dot")
        return out.value

    def calc_xr(self, alpha) -> float:
        out = C.c_double(0.0)
        self.ctx._ck(self.L.abft_cuda_dist_calc_xr(self.ctx.ctx, self.comm, self.x.handle, self.r.handle, self.p.handle,
                                                   self.w.handle, float(alpha), C.byref(out)), "dist_calc_xr")
        return out.value

    def calc_p(self, beta):
        self.ctx._ck(self.L.abft_cuda_dist_calc_p(self.ctx.ctx, self.comm, self.p.handle, self.r.handle, float(beta)), "dist_calc_p")

    def push_p(self):
        self.ctx._ck(self.L.abft_cuda_dist_push(self.ctx.ctx, self.comm, self.p.handle), "dist_push")

    # ---- cg.cpp:83-118 ----
    def start(self):
        self.ctx.copy_vector(self.r, self.b)
        self.ctx.copy_vector(self.p, self.r)
        self.push_p()
        self.rr = self.dot(self.r, self.r)
        return self.rr

    def iterate(self, k: int, conv: float = 0.0, hist=None) -> int:
        rr = self.rr
        done = 0
        while done < k and rr > conv:
            self.spmv()
            pw = self.dot(self.p, self.w)
            alpha = rr / pw if pw != 0.0 else 0.0
            rr_new = self.calc_xr(alpha)
            beta = rr_new / rr if rr != 0.0 else 0.0
            self.calc_p(beta)
            rr = rr_new
            if hist is not None:
                hist.append(rr)
            done += 1
        self.rr = rr
        return done

    def host_loop(self, k: int, conv: float = 0.0) -> int:
        """iterate(), compiled (abft_cuda_dist_cg_host_loop: the same four dist_* calls per iteration)."""
        rr = C.c_double(self.rr)
        n = self.L.abft_cuda_dist_cg_host_loop(self.ctx.ctx, self.comm, self.A.handle, self.x.handle, self.r.handle,
                                               self.p.handle, self.p_full.handle, self.w.handle, int(k), float(conv),
                                               C.byref(rr), None)
        if n < 0:
            from . import capi
            raise RuntimeError("cuda target: dist_cg_host_loop failed: " + capi.last_error())
        self.rr = rr.value
        return n

    def solve_device(self, max_itrs: int, conv: float, check_every: int = 32):
        """Device-resident loop (abft_cuda_dist_cg_solve): r = b, p = r, then iterate; -> (iterations, rr)."""
        rr = C.c_double(0.0)
        it = self.L.abft_cuda_dist_cg_solve(self.ctx.ctx, self.comm, self.A.handle, self.b.handle, self.x.handle,
                                            self.r.handle, self.p.handle, self.p_full.handle, self.w.handle, int(max_itrs),
                                            float(conv), int(check_every), C.byref(rr))
        if it == -2:
            self.ctx.report_faults()     # a fatal fault on this rank: raises FatalFault with the lines the reference prints
        if it < 0:
            from . import capi
            raise RuntimeError("cuda target: dist_cg_solve failed: " + capi.last_error())
        self.rr = rr.value
        return it, rr.value

    def close(self):
        if self.comm:
            self.L.abft_cuda_comm_destroy(self.ctx.ctx, self.comm)
            self.comm = None
