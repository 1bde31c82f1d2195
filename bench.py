#!/usr/bin/env python
"""bench.py -- CG iterations/s of the CUDA target on the BASELINE.json roofline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode secded] [--grid 256] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one CG iteration (cg.cpp:94-118): spmv (+ fused dot), calc_xr, calc_p, driven through the
C ABI exactly as the CGContext plugin drives it (two host-visible scalars per iteration).

Workload (config.workload): 3D 7-point Poisson on a grid^3 mesh (default 256^3: N = 2^24, nnz = 117 047 296,
BASELINE.json configs[2] / the north_star's ">= 100M non-zeros" roofline case), ABFT mode `secded` unless
--mode says otherwise, right-hand side uniform in [0,1) like cg.cpp:70.  Synthetic, generated in memory in the
triplet order cg.cpp hands to create_matrix.  At N > 1 the rows are partitioned across ranks (strong scaling:
the matrix is fixed), the search direction p is all-gathered over NCCL and the two scalars are all-reduced.

Printed JSON keys follow the driver's contract; see DESIGN.md "Measurement" for the definitions of
value / e2e / roofline / cpu_baseline used here.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--mode", default="secded")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--workload", default="poisson3d", choices=["poisson3d", "poisson2d_blocks", "powerlaw"],
                    help="poisson3d: grid^3 7-point (default, BASELINE configs[2]); poisson2d_blocks: 512^2 5-point x 64 "
                         "diagonal blocks (largest 5-point case the 24-bit column allows, configs[3]); powerlaw: irregular "
                         "rows, --rows x --nnz (default N=2^23, ~2e8 nnz: configs[4] as named; round 1 measured 2^21 / 5e7)")
    ap.add_argument("--rows", type=int, default=1 << 23, help="powerlaw: N (<= 2^24 for the ECC modes)")
    ap.add_argument("--nnz", type=int, default=200_000_000, help="powerlaw: target non-zeros after mirroring")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--all-modes", type=int, default=1, help="also time every ABFT mode briefly (N=1 only)")
    ap.add_argument("--cpu-iters", type=int, default=5, help="iterations of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = stores over NVLink fused into the kernels (product); nccl = all-gather/all-reduce calls (baseline)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, t_begin=None, t_end=None):
        """Summary over the samples taken in [t_begin, t_end] (time.time() values; None = all)."""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        total = 0
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            total += 1
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.05):
                    continue
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples_in_timed_region": len(sm),
                "samples_total": total, "reasons": sorted(reasons)}


def ncu_traffic(mode, workload, grid, world):
    """DRAM bytes (read + write) of one spmv launch from the committed ncu capture of this workload/mode
    (profiles/r1_spmv_dram_traffic.json), or None when there is no capture for it."""
    path = os.path.join(ROOT, "profiles", "r1_spmv_dram_traffic.json")
    if workload != "poisson3d" or grid != 256 or world != 1 or not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    return d.get(mode, {}).get("traffic")


def slab_triplets(mg, grid, r0, r1):
    """Rows [r0, r1) of the grid^3 7-point Poisson matrix, global columns, cg.cpp order."""
    n = grid
    idx = np.arange(r0, r1, dtype=np.int64)
    z, y, x = idx // (n * n), (idx // n) % n, idx % n
    offs = np.array([-n * n, -n, -1, 0, 1, n, n * n], dtype=np.int64)
    valid = np.stack([z > 0, y > 0, x > 0, np.ones_like(x, dtype=bool), x < n - 1, y < n - 1, z < n - 1], axis=1)
    vals = np.array([-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0])
    cols = (idx[:, None] + offs[None, :])[valid]
    rows = np.broadcast_to(idx[:, None], valid.shape)[valid]
    v = np.broadcast_to(vals[None, :], valid.shape)[valid]
    return mg.Triplets(cols, rows, v, n ** 3)


def alg_bytes(N, nnz):
    """bytes that must move per CG iteration / per spmv."""
    return 12 * nnz + 92 * N + 4, 12 * nnz + 20 * N + 4


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU target (oracle/_ref, built from its unmodified sources)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(args, t, b, iters, warm):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    if not os.path.exists(O.REF_SO):
        # oracle/_ref (the reference built from its own sources) did not travel: fall back to the oracle's C port
        # of the same algorithm -- serial, so it understates what the reference's OpenMP spmv would do
        mode = args.mode
        t0 = time.time()
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        t_create = time.time() - t0
        L = O.lib()
        N = t.N
        x = np.zeros(N); r = b.copy(); p = b.copy(); w = np.zeros(N)
        rr = L.orc_dot(r, r, N)

        def run(k):
            nonlocal rr
            for _ in range(k):
                y, fatal, ev = om.spmv(p)
                w[:] = y
                pw = L.orc_dot(p, w, N)
                rr_new = L.orc_calc_xr(x, r, p, w, rr / pw, N)
                L.orc_calc_p(p, r, rr_new / rr, N)
                rr = rr_new
        run(min(warm, 1))
        t0 = time.time()
        run(iters)
        dt = time.time() - t0
        return iters / dt, dt, t_create, "port"
    kind = "reference"
    rc = O.RefContext(args.mode if args.mode != "crc32c" else "secded")
    t0 = time.time()
    rc.create_matrix(t.cols, t.rows, t.vals, t.N)
    t_create = time.time() - t0
    R = rc.R
    bv = rc.vector(b); x = rc.vector(); r = rc.vector(); p = rc.vector(); w = rc.vector()
    R.ref_copy_vector(rc.ctx, r, bv)
    R.ref_copy_vector(rc.ctx, p, r)
    rr = ctypes.c_double(R.ref_dot(rc.ctx, r, r))
    if warm:
        R.ref_cg_iterations(rc.ctx, rc.A, x, r, p, w, warm, 0.0, ctypes.byref(rr), None)
    t0 = time.time()
    done = R.ref_cg_iterations(rc.ctx, rc.A, x, r, p, w, iters, 0.0, ctypes.byref(rr), None)
    dt = time.time() - t0
    return done / dt, dt, t_create, kind


def emit(line):
    """The one JSON line goes to the real stdout; everything else written to fd 1 meanwhile (NCCL's version
    banner, library chatter) was diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    import pkg_loader
    pkg = pkg_loader.load()
    mg = pkg.matgen
    n = args.grid
    full = None
    if args.workload == "poisson3d":
        N = n ** 3
        nnz = 7 * N - 6 * n * n
        workload = f"poisson3d_{n} (3D 7-point Poisson {n}^3: N={N}, nnz={nnz}), mode {args.mode}, -c 0 -i K"
    else:
        full = mg.poisson2d(512, 64) if args.workload == "poisson2d_blocks" else mg.powerlaw(args.rows, args.nnz, seed=1)
        N, nnz = full.N, full.nnz
        what = ("2D 5-point Poisson 512^2 replicated -b 64" if args.workload == "poisson2d_blocks"
                else "power-law row lengths (Zipf 2.0, <= 4096 per row), SPD")
        workload = f"{args.workload} ({what}: N={N}, nnz={nnz}), mode {args.mode}, -c 0 -i K"
    B_alg, B_spmv = alg_bytes(N, nnz)
    peak, peak_src = peaks()

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        t = slab_triplets(mg, n, 0, N) if full is None else full
        b = np.random.default_rng(1).random(N)
        its, dt, t_create, kind = cpu_reference_run(args, t, b, args.steps, args.warmup)
        cores = os.cpu_count() if kind == "reference" else 1
        line = {"impl": "reference", "metric": "cg_iterations_per_sec", "value": its, "unit": "it/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / its, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "threads": cores,
                           "note": "reference cg-csr CPU target (CSR/CPUContext.cpp) through its CGContext API, "
                                   "OpenMP over all host cores (only spmv is parallel in the reference)"},
                "cpu_baseline": {"value": its, "unit": "it/s", "cores": cores, "kind": kind,
                                 "sample": f"{args.steps} CG iterations on the full matrix after {args.warmup} warm-up"},
                "e2e": {"value": its, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    os.environ["NCCL_DEBUG"] = "WARN"      # the version banner goes to stdout and would precede the JSON line
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi takes a while to spin up
    # row partition (contiguous, balanced by rows == by nnz up to the boundary planes)
    t_gen = time.time()
    if full is None:
        cuts = [(N * k) // world for k in range(world + 1)]
        r0, r1 = cuts[rank], cuts[rank + 1]
        t = slab_triplets(mg, n, r0, r1)
    else:
        rowptr = full.row_pointers().astype(np.int64)
        cuts = pkg.dist.balanced_row_cuts(rowptr, world)
        r0, r1 = cuts[rank], cuts[rank + 1]
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        t = mg.Triplets(full.cols[e0:e1], full.rows[e0:e1], full.vals[e0:e1], N)
    t_gen = time.time() - t_gen
    nnz_local = t.nnz
    if world > 1:
        counts = torch.tensor([nnz_local], dtype=torch.int64, device="cuda")
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        nnz_offset = int(sum(int(c.item()) for c in allc[:rank]))
    else:
        nnz_offset = 0
    b_full = np.random.default_rng(1).random(N)
    nloc = r1 - r0

    def setup(mode):
        ctx = pkg.CUDAContext(mode, local_rank)
        tu = time.time()
        A = ctx.create_matrix_part(t.cols, t.rows, t.vals, nloc, N, r0, nnz_offset)
        ctx.sync()
        return ctx, A, time.time() - tu

    use_p2p = world > 1 and args.exchange == "p2p"
    if use_p2p:
        def all_gather_obj(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        ctx = pkg.CUDAContext(args.mode, local_rank)
        tu = time.time()
        dcg = pkg.dist.DistCG(ctx, rank, world, all_gather_obj, t, cuts, N, nnz_offset)
        ctx.sync()
        A, t_upload = dcg.A, time.time() - tu
    else:
        ctx, A, t_upload = setup(args.mode)
    stream = torch.cuda.ExternalStream(ctx.L.abft_cuda_ctx_stream(ctx.ctx), device=torch.device("cuda", local_rank))

    # vectors: p is full length on every rank (spmv gathers from it); the rest are the local slices
    if use_p2p:
        p_full, p_loc, vb, x, r, w = dcg.p_full, dcg.p, dcg.b, dcg.x, dcg.r, dcg.w
        p_full_t = None
    else:
        p_full_t = torch.zeros(N, dtype=torch.float64, device="cuda")
        p_full = ctx.wrap_vector(p_full_t.data_ptr(), N)
        p_loc = ctx.view_vector(p_full, r0, nloc)
        vb, x, r, w = (ctx.create_vector(nloc) for _ in range(4))
    scal_t = torch.zeros(1, dtype=torch.float64, device="cuda")

    def allreduce(v):
        if world == 1:
            return v
        scal_t.fill_(v)
        dist.all_reduce(scal_t)
        return float(scal_t.item())

    def exchange_p():
        if world > 1:
            ctx.sync()
            dist.all_gather_into_tensor(p_full_t, p_full_t[r0:r1])
            torch.cuda.current_stream().synchronize()

    hb = ctx.map_vector(vb)
    hb[:] = b_full[r0:r1]
    hx = ctx.map_vector(x)
    hx[:] = 0.0

    state = {"rr": 0.0}

    def start_solve():
        ctx.unmap_vector(vb, hb)          # H2D from the pinned map buffers
        ctx.unmap_vector(x, hx)
        if use_p2p:
            dcg.start()
            state["rr"] = dcg.rr
            return
        ctx.copy_vector(r, vb)
        ctx.copy_vector(p_loc, r)
        exchange_p()
        state["rr"] = allreduce(ctx.dot(r, r))

    def iterate(k):
        if use_p2p:
            dcg.rr = state["rr"]
            dcg.iterate(k, 0.0)
            state["rr"] = dcg.rr
            return
        rr = state["rr"]
        for _ in range(k):
            ctx.spmv(A, p_full, w)
            pw = allreduce(ctx.dot(p_loc, w))
            alpha = rr / pw if pw != 0.0 else 0.0
            rr_new = allreduce(ctx.calc_xr(x, r, p_loc, w, alpha))
            beta = rr_new / rr if rr != 0.0 else 0.0
            ctx.calc_p(p_loc, r, beta)
            exchange_p()
            rr = rr_new
        state["rr"] = rr

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """device time of fn() on the context's stream, max over ranks (ms)"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- host-driven loop (what the CGContext plugin does: two scalars to the host per iteration) ----
    start_solve()
    iterate(max(args.warmup, 3))
    ms_host = timed(lambda: iterate(args.steps))
    its_host = args.steps / (ms_host * 1e-3)
    ctx.report_faults()

    # ---- value: K iterations of the device-resident loop (abft_cuda_[dist_]cg_solve: same kernels, the scalars
    # stay on the device, one readback at the end), operands resident in HBM ----
    if use_p2p:
        def dev_loop(k):
            return dcg.solve_device(k, 0.0, k)[0]
    else:
        if world > 1:
            dev_loop = None          # the NCCL baseline has no device-resident loop
        else:
            def dev_loop(k):
                return ctx.cg_solve_device(A, vb, x, r, p_loc, w, k, 0.0, k)[0]
    if dev_loop is not None:
        dev_loop(max(args.warmup, 3))
        box = {}
        l0 = ctx.launch_count()
        t_begin = time.time()
        ms_total = timed(lambda: box.update(it=dev_loop(args.steps)))
        t_end = time.time()
        launches = ctx.launch_count() - l0
        assert box["it"] == args.steps, box
    else:
        l0 = ctx.launch_count()
        t_begin = time.time()
        ms_total = timed(lambda: iterate(args.steps))
        t_end = time.time()
        launches = ctx.launch_count() - l0
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    its = args.steps / (ms_total * 1e-3)
    ctx.report_faults()

    # ---- e2e: a whole solve from host buffers through the C ABI (b in, x out), K iterations ----
    def e2e():
        start_solve()
        iterate(args.steps)
        ctx.map_vector(x)                 # D2H of the solution
    ms_e2e = timed(e2e)
    its_e2e = args.steps / (ms_e2e * 1e-3)
    h2d = (16 * nloc) / args.steps + 16
    d2h = (8 * nloc) / args.steps + 16

    # ---- roofline of the dominant kernel: spmv alone, back to back on the stream ----
    reps = max(20, min(args.steps, 100))
    # (with the fused exchange the spmv alone would wait for a push that never comes: time the plain kernel)
    for _ in range(3):
        ctx.spmv(A, p_full, w)
    ms_spmv = timed(lambda: [ctx.spmv(A, p_full, w) for _ in range(reps)]) / reps
    spmv_gbs = (B_spmv / world) / (ms_spmv * 1e-3) / 1e9
    iter_gbs = (B_alg / world) / (ms_total / args.steps * 1e-3) / 1e9
    ctx.report_faults()

    line = {"metric": "cg_iterations_per_sec", "value": its, "unit": "it/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "rows_per_gpu": nloc, "nnz_per_gpu": nnz_local,
                       "partition": "1-D contiguous row blocks" if world > 1 else "single GPU",
                       "exchange": ("none" if world == 1 else
                                    "fused: halo of p stored into the peers' copies over NVLink inside calc_p, "
                                    "scalars all-reduced through peer memory inside the reductions (%d B pushed per rank per iteration)"
                                    % pkg.dist.exchanged_bytes(dcg.plan) if use_p2p else
                                    "NCCL all-gather of p + all-reduce of 2 scalars (baseline)"),
                       "l2": "inputs larger than L2 (matrix stream %.0f MB per GPU per iteration)" % (12 * nnz_local / 1e6),
                       "abi": "C ABI of include/abft_cuda.h via ctypes; value = device-resident loop "
                              "(abft_cuda_[dist_]cg_solve), e2e / host_driven_loop = the four plugin calls per iteration"},
            "host_driven_loop": {"value": its_host, "unit": "it/s", "ms_per_step": ms_host / args.steps,
                                 "what": "spmv, dot, calc_xr, calc_p through the C ABI as cg.cpp drives them, operands resident"},
            "e2e": {"value": its_e2e, "unit": "it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "b,x uploaded from pinned host buffers + K iterations + x downloaded, all timed"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": f"spmv_kernel<{args.mode}>", "achieved": spmv_gbs, "peak": peak,
                         "unit": "GB/s", "frac": spmv_gbs / peak,
                         "traffic": ncu_traffic(args.mode, args.workload, args.grid, world), "peak_source": peak_src,
                         "bytes_per_launch": B_spmv // world, "ms_per_launch": ms_spmv},
            "roofline_cg_iteration": {"bound": "hbm", "bytes_per_iteration": B_alg // world, "achieved": iter_gbs,
                                      "peak": peak, "unit": "GB/s", "frac": iter_gbs / peak},
            "clocks": clocks,
            "setup": {"generate_s": t_gen, "create_matrix_s": t_upload}}

    if args.workload == "powerlaw":
        # random columns: every non-zero is one 32-byte sector request for 8 useful bytes, and the B200 serves one such
        # request per clock per SM (tools/gather_bench.cu, profiles/r1_gather_bench.txt: 288 G/s from an L2-resident
        # vector) -- that, not HBM, is this workload's ceiling
        gathers = nnz_local / (ms_spmv * 1e-3) / 1e9
        line["gather_rate"] = {"achieved": gathers, "peak": 288.0, "unit": "G gathers/s", "frac": gathers / 288.0,
                               "peak_source": "measured: profiles/r1_gather_bench.txt (1 sector per clock per SM)"}

    if world == 1 and rank == 0:
        # ---- every ABFT mode, briefly ----
        if args.all_modes:
            modes = {}
            for mode in pkg.MODES:
                if mode == args.mode:
                    modes[mode] = {"it_s": its, "iter_frac": iter_gbs / peak, "spmv_frac": spmv_gbs / peak}
                    continue
                c2, A2, _ = setup(mode)
                s2 = torch.cuda.ExternalStream(c2.L.abft_cuda_ctx_stream(c2.ctx))
                bv, xv, rv, pv, wv = (c2.create_vector(N) for _ in range(5))
                c2.upload(bv, b_full)
                c2.cg_solve_device(A2, bv, xv, rv, pv, wv, 5, 0.0, 5)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(s2)
                c2.cg_solve_device(A2, bv, xv, rv, pv, wv, 50, 0.0, 50)
                e1.record(s2); e1.synchronize()
                ms_it = e0.elapsed_time(e1) / 50
                e0.record(s2)
                for _ in range(30):
                    c2.spmv(A2, pv, wv)
                e1.record(s2); e1.synchronize()
                ms_sp = e0.elapsed_time(e1) / 30
                modes[mode] = {"it_s": 1e3 / ms_it, "iter_frac": B_alg / (ms_it * 1e-3) / 1e9 / peak,
                               "spmv_frac": B_spmv / (ms_sp * 1e-3) / 1e9 / peak}
                c2.report_faults()
                c2.destroy_matrix(A2)
                c2.close()
            line["modes"] = modes

        # ---- CPU baseline: the reference's CPU target on this box's host cores, bounded sample ----
        if not args.no_cpu_baseline:
            try:
                v, dt, t_create, kind = cpu_reference_run(args, t, b_full, args.cpu_iters, 1)
                line["cpu_baseline"] = {"value": v, "unit": "it/s", "cores": os.cpu_count() if kind == "reference" else 1,
                                        "kind": kind,
                                        "sample": f"{args.cpu_iters} CG iterations on the full matrix "
                                                  f"({dt:.1f} s; create_matrix {t_create:.1f} s untimed), OpenMP all cores"}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "it/s", "cores": os.cpu_count(), "kind": "unavailable",
                                        "sample": str(e)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
