#!/usr/bin/env python
"""bench.py -- CG iterations/s of the CUDA target on the BASELINE.json roofline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode secded] [--grid 256] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one CG iteration (cg.cpp:94-118): spmv (+ fused dot), calc_xr, calc_p, driven through the C ABI
exactly as the CGContext plugin drives it -- four calls, two host-visible scalars per iteration.

Workload (config.workload): 3D 7-point Poisson on a grid^3 mesh (default 256^3: N = 2^24, nnz = 117 047 296,
BASELINE.json configs[2] / the north_star's ">= 100M non-zeros" roofline case), ABFT mode `secded` unless --mode
says otherwise, right-hand side uniform in [0,1) like cg.cpp:70.  Synthetic, generated in memory in the triplet
order cg.cpp hands to create_matrix.  At N > 1 the rows are partitioned across ranks (strong scaling: the matrix is
fixed); the halo of the search direction p is stored into the peers' copies over NVLink by calc_p itself and the
two scalars are all-reduced through peer memory inside the reductions (CUDA IPC; torch.distributed / NCCL carries
only the setup and the timing barrier, nothing on the data path).

What the JSON line holds:
  value                  K / median over --blocks timed blocks of K iterations each, every block bracketed by
                         barrier + synchronize, CUDA events on the context's stream, max over ranks; the PLUGIN path
                         (host-driven four-call loop), operands resident in HBM.  `blocks` has min / max / all.
  device_resident_loop   the same kernels with the scalars kept on the device (abft_cuda_[dist_]cg_solve): a path
                         cg.cpp cannot reach, reported beside the value, not as it.
  e2e                    a whole solve from host buffers: b, x uploaded, K iterations, x downloaded, all timed.
  roofline               spmv_kernel alone, back to back (matrix stream >> L2), algorithmic bytes / time / peak.
  check, rr_after        correctness beside the number: rr and sum(x) after `check.iterations` iterations from
                         x = 0 on the same b, and rr after warm-up + K iterations; the reference arm prints the
                         same keys (relative difference of rr <= 1e-10 is the north_star's bar).
  workloads              BASELINE configs[3] (512^2 -b 64, sec8) and configs[4] (power-law, sec7 / constraints) at
                         their named sizes on the same N GPUs, time-boxed.
  formats                (N = 1) the reference's second format, COO (cg-coo), on the primary matrix in the primary mode:
                         it/s through the C ABI, spmv kernel pair against the format's own roofline.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CHECK_ITERS = 3     # iterations of the untimed correctness solve both arms print (check.rr, check.x_sum)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--blocks", type=int, default=10, help="timed blocks of --steps iterations each (median reported)")
    ap.add_argument("--mode", default="secded")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--workload", default="poisson3d", choices=["poisson3d", "poisson2d_blocks", "powerlaw"],
                    help="poisson3d: grid^3 7-point (default, BASELINE configs[2]); poisson2d_blocks: 512^2 5-point x 64 "
                         "diagonal blocks (largest 5-point case the 24-bit column allows, configs[3]); powerlaw: irregular "
                         "rows, --rows x --nnz (default N=2^23, ~1.9e8 nnz: configs[4] as named)")
    ap.add_argument("--rows", type=int, default=1 << 23, help="powerlaw: N (<= 2^24 for the ECC modes)")
    ap.add_argument("--nnz", type=int, default=200_000_000, help="powerlaw: target non-zeros after mirroring")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--all-modes", type=int, default=1, help="also time every ABFT mode briefly (N=1 only)")
    ap.add_argument("--workloads", type=int, default=1, help="also run BASELINE configs[3] and [4] briefly (any N)")
    ap.add_argument("--budget-s", type=float, default=240.0, help="stop adding optional sections after this much wall time")
    ap.add_argument("--cpu-iters", type=int, default=5, help="iterations of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def source_hash():
    """Identifies the kernel sources the loaded library was built from (the ncu traffic capture names the same)."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "abft-sparse-cg_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode()); h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(mode, workload, grid, world):
    """DRAM bytes (read + write) of one spmv launch from the committed `ncu --set full` capture of this workload and
    mode -- only if that capture was taken on a library built from the very sources in this tree; else None."""
    path = os.path.join(ROOT, "profiles", "r2_spmv_dram_traffic.json")
    if workload != "poisson3d" or grid != 256 or world != 1 or not os.path.exists(path):
        return None
    with open(path) as f:
        d = json.load(f)
    if d.get("source_hash") != source_hash():
        return None
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


def describe(args, workload, mode, mg=None, full=None):
    """-> (N, nnz, workload string).  `full` (the generated matrix) is needed for everything but poisson3d."""
    if workload == "poisson3d":
        n = args.grid
        N = n ** 3
        nnz = 7 * N - 6 * n * n
        return N, nnz, f"poisson3d_{n} (3D 7-point Poisson {n}^3: N={N}, nnz={nnz}), mode {mode}, -c 0 -i K"
    what = ("2D 5-point Poisson 512^2 replicated -b 64" if workload == "poisson2d_blocks"
            else "power-law row lengths (Zipf 2.0, <= 4096 per row), SPD")
    return full.N, full.nnz, f"{workload} ({what}: N={full.N}, nnz={full.nnz}), mode {mode}, -c 0 -i K"


def generate(args, mg, workload):
    if workload == "poisson2d_blocks":
        return mg.poisson2d(512, 64)
    if workload == "powerlaw":
        return mg.powerlaw(args.rows, args.nnz, seed=1)
    return None


def config_of(workload_str, N, nnz, mode):
    """Identical in both arms (the driver compares it)."""
    return {"workload": workload_str, "mode": mode, "N": int(N), "nnz": int(nnz), "precision": "f64",
            "l2": ("inputs larger than L2" if 12 * nnz > 2 * 126e6 else "inputs NOT larger than L2")
                  + " (matrix stream %.0f MB per iteration)" % (12 * nnz / 1e6)}


def rhs(N):
    return np.random.default_rng(1).random(N)


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU target (oracle/_ref, built from its unmodified sources)
# ---------------------------------------------------------------------------------------------------
def omp_threads(want=None):
    """Set (optionally) and read back the OpenMP thread count the reference library will use."""
    try:
        g = ctypes.CDLL("libgomp.so.1")
        if want:
            g.omp_set_num_threads(int(want))
        return int(g.omp_get_max_threads())
    except OSError:
        return None


def cpu_reference_run(args, mode, t, b, iters, warm, check_iters=0):
    """-> dict(it_s, seconds, create_s, kind, threads, check, rr_after).  The reference's CPU target driven through
    its CGContext API from x = 0: `warm` untimed iterations, then `iters` timed ones."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the reference arm is to use the host's cores
    ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    import oracle_py as O
    N = t.N
    if not os.path.exists(O.REF_SO):
        # oracle/_ref (the reference built from its own sources) did not travel: the oracle's C port of the same
        # algorithm -- serial, so it understates what the reference's OpenMP spmv would do
        t0 = time.time()
        om = O.OracleMatrix(mode if mode != "secdedv" else "secded", t.cols, t.rows, t.vals, N)
        t_create = time.time() - t0
        L = O.lib()
        x = np.zeros(N); r = b.copy(); p = b.copy(); w = np.zeros(N)
        st = {"rr": L.orc_dot(r, r, N)}

        def run(k):
            for _ in range(k):
                y, fatal, ev = om.spmv(p)
                w[:] = y
                pw = L.orc_dot(p, w, N)
                rr_new = L.orc_calc_xr(x, r, p, w, st["rr"] / pw, N)
                L.orc_calc_p(p, r, rr_new / st["rr"], N)
                st["rr"] = rr_new
        check = None
        if check_iters:
            run(check_iters)
            check = {"iterations": check_iters, "rr": st["rr"], "x_sum": float(x.sum()), "x_abs_sum": float(np.abs(x).sum())}
            x[:] = 0.0; r[:] = b; p[:] = b
            st["rr"] = L.orc_dot(r, r, N)
        run(warm)
        t0 = time.time()
        run(iters)
        dt = time.time() - t0
        return dict(it_s=iters / dt, seconds=dt, create_s=t_create, kind="port", threads=1, check=check,
                    rr_after={"iterations": warm + iters, "rr": st["rr"]})
    ref_mode = mode if mode in ("none", "constraints", "sed", "sec7", "sec8", "secded") else "secded"
    rc = O.RefContext(ref_mode)   # (the reference has no crc32c / secdedv: its costliest mode stands in)
    threads = omp_threads(ncpu)
    t0 = time.time()
    rc.create_matrix(t.cols, t.rows, t.vals, N)
    t_create = time.time() - t0
    R = rc.R
    bv = rc.vector(b); x = rc.vector(); r = rc.vector(); p = rc.vector(); w = rc.vector()
    R.ref_copy_vector(rc.ctx, r, bv)
    R.ref_copy_vector(rc.ctx, p, r)
    rr = ctypes.c_double(R.ref_dot(rc.ctx, r, r))
    check = None

    def run(k):
        return R.ref_cg_iterations(rc.ctx, rc.A, x, r, p, w, k, 0.0, ctypes.byref(rr), None) if k > 0 else 0
    if check_iters:
        # correctness beside the number: a fresh solve of check_iters iterations from x = 0 (our arm does the same)
        run(check_iters)
        xs = rc.view(x)
        check = {"iterations": check_iters, "rr": rr.value, "x_sum": float(xs.sum()), "x_abs_sum": float(np.abs(xs).sum())}
        rc.view(x)[:] = 0.0
        R.ref_copy_vector(rc.ctx, r, bv)
        R.ref_copy_vector(rc.ctx, p, r)
        rr = ctypes.c_double(R.ref_dot(rc.ctx, r, r))
    run(warm)
    t0 = time.time()
    got = run(iters)
    dt = time.time() - t0
    return dict(it_s=got / dt, seconds=dt, create_s=t_create, kind="reference", threads=threads, check=check,
                rr_after={"iterations": warm + got, "rr": rr.value})


def emit(line):
    """The one JSON line goes to the real stdout; everything else written to fd 1 meanwhile (NCCL's version
    banner, library chatter) was diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


# ---------------------------------------------------------------------------------------------------
# our arm: one workload on this rank's GPU (a row block of it when world > 1)
# ---------------------------------------------------------------------------------------------------
class Case:
    """One (workload, mode) on `world` GPUs, driven through the C ABI the way the CGContext plugin drives it."""

    def __init__(self, env, args, workload, mode, full=None):
        self.env, self.args, self.workload, self.mode = env, args, workload, mode
        pkg, torch, dist = env["pkg"], env["torch"], env["dist"]
        rank, world, local_rank = env["rank"], env["world"], env["local_rank"]
        mg = pkg.matgen
        self.N, self.nnz, self.workload_str = describe(args, workload, mode, mg, full)
        N = self.N
        t_gen = time.time()
        if full is None:
            cuts = [(N * k) // world for k in range(world + 1)]
            r0, r1 = cuts[rank], cuts[rank + 1]
            t = slab_triplets(mg, args.grid, r0, r1)
        else:
            rowptr = full.row_pointers().astype(np.int64)
            cuts = pkg.dist.balanced_row_cuts(rowptr, world)
            r0, r1 = cuts[rank], cuts[rank + 1]
            e0, e1 = int(rowptr[r0]), int(rowptr[r1])
            t = mg.Triplets(full.cols[e0:e1], full.rows[e0:e1], full.vals[e0:e1], N)
        self.t_gen = time.time() - t_gen
        self.t, self.r0, self.r1, self.nloc, self.nnz_local = t, r0, r1, r1 - r0, t.nnz
        if world > 1:
            counts = torch.tensor([t.nnz], dtype=torch.int64, device="cuda")
            allc = [torch.zeros_like(counts) for _ in range(world)]
            dist.all_gather(allc, counts)
            nnz_offset = int(sum(int(c.item()) for c in allc[:rank]))
        else:
            nnz_offset = 0
        self.ctx = ctx = pkg.CUDAContext(mode, local_rank)
        tu = time.time()
        self.dcg = None
        if world > 1:
            def all_gather_obj(obj):
                out = [None] * world
                dist.all_gather_object(out, obj)
                return out
            self.dcg = dcg = pkg.dist.DistCG(ctx, rank, world, all_gather_obj, t, cuts, N, nnz_offset)
            self.A = dcg.A
            self.p_full, self.p_loc, self.vb, self.x, self.r, self.w = dcg.p_full, dcg.p, dcg.b, dcg.x, dcg.r, dcg.w
        else:
            self.A = ctx.create_matrix_part(t.cols, t.rows, t.vals, self.nloc, N, r0, nnz_offset)
            self.p_full = ctx.create_vector(N)
            self.p_loc = self.p_full
            self.vb, self.x, self.r, self.w = (ctx.create_vector(self.nloc) for _ in range(4))
        ctx.sync()
        self.t_upload = time.time() - tu
        self.stream = torch.cuda.ExternalStream(ctx.L.abft_cuda_ctx_stream(ctx.ctx), device=torch.device("cuda", local_rank))
        self.hb = ctx.map_vector(self.vb)
        self.hb[:] = rhs(N)[r0:r1]
        self.hx = ctx.map_vector(self.x)
        self.hx[:] = 0.0
        self.rr = 0.0
        self.iterations = 0     # since the last start_solve

    # ---- cg.cpp:45-118 through the plugin calls ----
    def start_solve(self):
        ctx = self.ctx
        ctx.unmap_vector(self.vb, self.hb)          # H2D from the pinned map buffers
        ctx.unmap_vector(self.x, self.hx)
        if self.dcg:
            self.dcg.start()
            self.rr = self.dcg.rr
        else:
            ctx.copy_vector(self.r, self.vb)
            ctx.copy_vector(self.p_loc, self.r)
            self.rr = ctx.dot(self.r, self.r)
        self.iterations = 0

    def iterate(self, k):
        """k iterations of the plugin path: abft_cuda_[dist_]cg_host_loop = cg.cpp's loop body over the four public
        entry points (spmv, dot, calc_xr, calc_p; two host-visible scalars per iteration), compiled like cg.cpp is."""
        if self.dcg:
            self.dcg.rr = self.rr
            self.dcg.host_loop(k, 0.0)
            self.rr = self.dcg.rr
        else:
            _, self.rr = self.ctx.cg_host_loop(self.A, self.x, self.r, self.p_loc, self.w, k, 0.0, self.rr)
        self.iterations += k

    def dev_loop(self, k):
        if self.dcg:
            return self.dcg.solve_device(k, 0.0, k)[0]
        return self.ctx.cg_solve_device(self.A, self.vb, self.x, self.r, self.p_loc, self.w, k, 0.0, k)[0]

    def barrier(self):
        torch, dist = self.env["torch"], self.env["dist"]
        torch.cuda.synchronize()
        if self.env["world"] > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn):
        """device time of fn() on the context's stream, max over ranks (ms)"""
        torch, dist = self.env["torch"], self.env["dist"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        fn()
        e1.record(self.stream)
        e1.synchronize()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if self.env["world"] > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def x_sums(self):
        """(sum x, sum |x|) over all ranks, added in rank order."""
        xs = self.ctx.download(self.x)
        mine = (float(xs.sum()), float(np.abs(xs).sum()))
        if self.env["world"] == 1:
            return mine
        out = [None] * self.env["world"]
        self.env["dist"].all_gather_object(out, mine)
        return float(sum(o[0] for o in out)), float(sum(o[1] for o in out))

    def spmv_ms(self, reps):
        ctx = self.ctx
        # (with the fused exchange the spmv alone would wait for a push that never comes: the plain kernel is timed)
        for _ in range(3):
            ctx.spmv(self.A, self.p_full, self.w)
        return self.timed(lambda: [ctx.spmv(self.A, self.p_full, self.w) for _ in range(reps)]) / reps

    def close(self):
        if self.dcg:
            self.dcg.close()
        self.ctx.destroy_matrix(self.A)
        self.ctx.close()


def brief(case, steps, peak):
    """it/s of the plugin path, the device loop and the spmv kernel of one case (one block each)."""
    world = case.env["world"]
    B_alg, B_spmv = alg_bytes(case.N, case.nnz)
    case.start_solve()
    case.iterate(5)
    ms_host = case.timed(lambda: case.iterate(steps)) / steps
    rr = case.rr
    case.ctx.report_faults()
    case.dev_loop(5)
    box = {}
    ms_dev = case.timed(lambda: box.update(it=case.dev_loop(steps))) / steps
    ms_sp = case.spmv_ms(30)
    case.ctx.report_faults()
    out = {"workload": case.workload_str, "it_s": 1e3 / ms_host, "ms_per_step": ms_host,
           "iter_frac": (B_alg / world) / (ms_host * 1e-3) / 1e9 / peak,
           "device_resident_loop_it_s": 1e3 / ms_dev,
           "spmv_ms": ms_sp, "spmv_frac": (B_spmv / world) / (ms_sp * 1e-3) / 1e9 / peak,
           "rr_after": {"iterations": 5 + steps, "rr": rr},
           "nnz_per_gpu": case.nnz_local, "rows_per_gpu": case.nloc,
           "setup": {"generate_s": case.t_gen, "create_matrix_s": case.t_upload}}
    if case.dcg:
        out["pushed_bytes_per_rank_per_iteration"] = case.env["pkg"].dist.exchanged_bytes(case.dcg.plan)
    if case.workload == "powerlaw":
        # random columns: every non-zero is one 32-byte sector request for 8 useful bytes, and the B200 serves one such
        # request per clock per SM (tools/gather_bench.cu, profiles/r1_gather_bench.txt: 288 G/s from an L2-resident
        # vector) -- that, not HBM, is this workload's ceiling
        g = case.nnz_local / (ms_sp * 1e-3) / 1e9
        out["gather_rate"] = {"achieved": g, "peak": 288.0, "unit": "G gathers/s", "frac": g / 288.0,
                              "peak_source": "measured: profiles/r1_gather_bench.txt (1 sector per clock per SM)"}
    return out


def coo_brief(pkg, t, mode, steps, peak):
    """The reference's second format (cg-coo) on the same matrix: CG through the C ABI the way cg.cpp drives
    it (abft_cuda_coo_spmv, dot -- served by the spmv's own reduction --, calc_xr, calc_p) and the spmv kernel pair back
    to back, against the format's own algorithmic bytes (16 B per element + the two vectors); tools/bench_coo.py is
    the long form."""
    import torch
    N, nnz = t.N, int(t.cols.shape[0])
    ctx = pkg.CUDAContext(mode, 0)
    t0 = time.time()
    A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, N)
    ctx.sync()
    t_create = time.time() - t0
    stream = torch.cuda.ExternalStream(ctx.L.abft_cuda_ctx_stream(ctx.ctx))
    vb, x, r, p, w = (ctx.create_vector(N) for _ in range(5))
    ctx.upload(vb, rhs(N))
    ctx.upload(x, np.zeros(N))
    ctx.copy_vector(r, vb)
    ctx.copy_vector(p, r)
    rr = ctx.dot(r, r)

    def iterations(k, rr):
        for _ in range(k):
            ctx.coo_spmv(A, p, w)
            pw = ctx.dot(p, w)
            rr_new = ctx.calc_xr(x, r, p, w, rr / pw)
            ctx.calc_p(p, r, rr_new / rr)
            rr = rr_new
        return rr

    rr = iterations(5, rr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    n0 = ctx.launch_count()
    e0.record(stream)
    rr = iterations(steps, rr)
    e1.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launch_count() - n0
    ms_step = e0.elapsed_time(e1) / steps
    for _ in range(5):
        ctx.coo_spmv(A, p, w)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(30):
        ctx.coo_spmv(A, p, w)
    e1.record(stream)
    torch.cuda.synchronize()
    ms_spmv = e0.elapsed_time(e1) / 30
    events, fatal = ctx.poll_events()
    ctx.close()
    b_spmv = 16 * nnz + 16 * N
    b_iter = b_spmv + 72 * N
    return {"workload": f"the same matrix as COO (128-bit elements, N={N}, nnz={nnz}), mode {mode}", "it_s": 1e3 / ms_step,
            "ms_per_step": ms_step, "iter_frac": b_iter / ms_step / 1e6 / peak, "spmv_ms": ms_spmv,
            "spmv_frac": b_spmv / ms_spmv / 1e6 / peak, "bytes_per_spmv": b_spmv, "gpu_launches": int(launches),
            "rr_after": {"iterations": 5 + steps, "rr": rr}, "faults": len(events), "setup": {"create_matrix_s": t_create}}


def main():
    global _REAL_STDOUT
    t_start = time.time()
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    warm = args.warmup

    import pkg_loader
    pkg = pkg_loader.load()
    mg = pkg.matgen
    peak, peak_src = peaks()

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        full = generate(args, mg, args.workload)
        N, nnz, workload = describe(args, args.workload, args.mode, mg, full)
        t = slab_triplets(mg, args.grid, 0, N) if full is None else full
        res = cpu_reference_run(args, args.mode, t, rhs(N), args.steps, warm, CHECK_ITERS)
        its = res["it_s"]
        line = {"impl": "reference", "metric": "cg_iterations_per_sec", "value": its, "unit": "it/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 / its, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_of(workload, N, nnz, args.mode),
                "cpu_baseline": {"value": its, "unit": "it/s", "cores": res["threads"], "kind": res["kind"],
                                 "sample": f"{args.steps} CG iterations on the full matrix after {warm} warm-up "
                                           f"({res['seconds']:.1f} s; create_matrix {res['create_s']:.1f} s untimed)"},
                "e2e": {"value": its, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "check": res["check"], "rr_after": res["rr_after"],
                "details": {"omp_threads": res["threads"], "host_cpus": os.cpu_count(),
                            "note": "reference cg-csr CPU target (CSR/CPUContext.cpp) through its CGContext API, OpenMP over "
                                    "all host cores (only spmv is parallel in the reference)"
                                    + ("; the reference has no %s mode: secded stands in" % args.mode if args.mode in ("crc32c", "secdedv") else "")}}
        emit(line)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    env = dict(pkg=pkg, torch=torch, dist=dist, rank=rank, world=world, local_rank=local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi takes a while to spin up

    full = generate(args, mg, args.workload)
    case = Case(env, args, args.workload, args.mode, full)
    ctx = case.ctx
    N, nnz, workload = case.N, case.nnz, case.workload_str
    B_alg, B_spmv = alg_bytes(N, nnz)
    K = args.steps

    # ---- correctness beside the number: CHECK_ITERS iterations from x = 0 (the reference arm prints the same) ----
    case.start_solve()
    case.iterate(CHECK_ITERS)
    xs, xa = case.x_sums()
    check = {"iterations": CHECK_ITERS, "rr": case.rr, "x_sum": xs, "x_abs_sum": xa}
    ctx.report_faults()

    # ---- value: the plugin path (host-driven four-call loop), --blocks blocks of K iterations ----
    case.start_solve()
    case.iterate(warm)
    block_ms, rr_after = [], None
    launches = 0
    t_begin = time.time()
    for blk in range(max(args.blocks, 1)):
        if case.iterations + K > 600:       # keep rr well away from underflow: start the solve over (untimed)
            case.start_solve()
            case.iterate(3)
        l0 = ctx.launch_count()
        block_ms.append(case.timed(lambda: case.iterate(K)))
        launches = ctx.launch_count() - l0
        if blk == 0:
            rr_after = {"iterations": case.iterations, "rr": case.rr}
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    ms_total = float(np.median(block_ms))
    its = K / (ms_total * 1e-3)
    ctx.report_faults()

    # ---- the device-resident loop (same kernels, scalars stay on the device): beside the value, not it ----
    case.dev_loop(warm)
    dev_ms = []
    for _ in range(3):
        box = {}
        dev_ms.append(case.timed(lambda: box.update(it=case.dev_loop(K))))
        assert box["it"] == K, box
    ms_dev = float(np.median(dev_ms))
    ctx.report_faults()

    # ---- e2e: a whole solve from host buffers through the C ABI (b, x in; x out), K iterations ----
    def e2e():
        case.start_solve()
        case.iterate(K)
        ctx.map_vector(case.x)                 # D2H of the solution
    e2e_ms = [case.timed(e2e) for _ in range(3)]
    ms_e2e = float(np.median(e2e_ms))
    its_e2e = K / (ms_e2e * 1e-3)
    h2d = (16 * case.nloc) / K + 16
    d2h = (8 * case.nloc) / K + 16

    # ---- roofline of the dominant kernel: spmv alone, back to back on the stream ----
    reps = max(20, min(K, 100))
    ms_spmv = case.spmv_ms(reps)
    spmv_gbs = (B_spmv / world) / (ms_spmv * 1e-3) / 1e9
    iter_gbs = (B_alg / world) / (ms_total / K * 1e-3) / 1e9
    ctx.report_faults()

    line = {"metric": "cg_iterations_per_sec", "value": its, "unit": "it/s", "n_gpus": world, "steps": K,
            "warmup": warm, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(workload, N, nnz, args.mode),
            "blocks": {"n": len(block_ms), "steps_each": K, "statistic": "median",
                       "it_s_min": K / (max(block_ms) * 1e-3), "it_s_max": K / (min(block_ms) * 1e-3),
                       "ms_per_block": [round(m, 4) for m in block_ms]},
            "path": "plugin: spmv, dot, calc_xr, calc_p through the C ABI of include/abft_cuda.h as cg.cpp drives them "
                    "(two host-visible scalars per iteration), operands resident in HBM",
            "device_resident_loop": {"value": K / (ms_dev * 1e-3), "unit": "it/s", "ms_per_step": ms_dev / K,
                                     "what": "abft_cuda_[dist_]cg_solve: same kernels, scalars stay on the device "
                                             "(a path cg.cpp cannot reach)"},
            "e2e": {"value": its_e2e, "unit": "it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "b,x uploaded from pinned host buffers + K iterations + x downloaded, all timed "
                            "(median of 3); the copies are %.0f MB per solve whatever K is" % (24 * case.nloc / 1e6)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": f"spmv_kernel<{args.mode}>", "achieved": spmv_gbs, "peak": peak,
                         "unit": "GB/s", "frac": spmv_gbs / peak,
                         "traffic": ncu_traffic(args.mode, args.workload, args.grid, world), "peak_source": peak_src,
                         "bytes_per_launch": B_spmv // world, "ms_per_launch": ms_spmv},
            "roofline_cg_iteration": {"bound": "hbm", "bytes_per_iteration": B_alg // world, "achieved": iter_gbs,
                                      "peak": peak, "unit": "GB/s", "frac": iter_gbs / peak},
            "check": check, "rr_after": rr_after,
            "clocks": clocks,
            "details": {"rows_per_gpu": case.nloc, "nnz_per_gpu": case.nnz_local,
                        "partition": "1-D contiguous row blocks" if world > 1 else "single GPU",
                        "collective": ("none (single GPU)" if world == 1 else
                                       "peer-memory stores over NVLink (CUDA IPC), no NCCL on the data path: halo of p stored "
                                       "into the peers' copies inside calc_p, scalars all-reduced through peer memory inside "
                                       "the reductions (%d B pushed per rank per iteration); NCCL only carries setup and the "
                                       "timing barrier" % pkg.dist.exchanged_bytes(case.dcg.plan)),
                        "source_hash": source_hash()},
            "setup": {"generate_s": case.t_gen, "create_matrix_s": case.t_upload}}
    if args.workload == "powerlaw":
        g = case.nnz_local / (ms_spmv * 1e-3) / 1e9
        line["gather_rate"] = {"achieved": g, "peak": 288.0, "unit": "G gathers/s", "frac": g / 288.0,
                               "peak_source": "measured: profiles/r1_gather_bench.txt (1 sector per clock per SM)"}
    b_full = rhs(N)
    t_primary = case.t
    case.close()
    del case

    def over_budget():
        flag = torch.tensor([1 if time.time() - t_start > args.budget_s else 0], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return bool(flag.item())

    # ---- every ABFT mode, briefly (N = 1) ----
    if world == 1 and args.all_modes:
        modes = {args.mode: {"it_s": its, "iter_frac": iter_gbs / peak, "spmv_frac": spmv_gbs / peak,
                             "device_resident_loop_it_s": K / (ms_dev * 1e-3)}}
        for mode in pkg.MODES:
            if mode == args.mode or over_budget():
                continue
            c2 = Case(env, args, args.workload, mode, full)
            r = brief(c2, 50, peak)
            modes[mode] = {k: r[k] for k in ("it_s", "iter_frac", "spmv_frac", "device_resident_loop_it_s")}
            c2.close()
            del c2
        line["modes"] = modes

    # ---- BASELINE configs[3] and configs[4] at their named sizes, on the same GPUs ----
    if args.workloads and args.workload == "poisson3d":
        wl = {}
        full = None
        for name, mode in (("poisson2d_blocks", "sec8"), ("powerlaw", "sec7"), ("powerlaw", "constraints")):
            if over_budget():
                wl[f"{name}_{mode}"] = {"skipped": "time budget (--budget-s)"}
                continue
            try:
                if full is None or full_name != name:
                    full, full_name = generate(args, mg, name), name
                c2 = Case(env, args, name, mode, full)
                wl[f"{name}_{mode}"] = brief(c2, min(K, 100), peak)
                c2.close()
                del c2
            except Exception as e:      # an optional section must not take the headline down with it
                wl[f"{name}_{mode}"] = {"error": repr(e)[:300]}
        line["workloads"] = wl

    # ---- the COO format on the primary matrix, N = 1 ----
    if world == 1 and args.all_modes and args.workload == "poisson3d":
        if over_budget():
            line["formats"] = {"coo_" + args.mode: {"skipped": "time budget (--budget-s)"}}
        elif args.mode in ("crc32c", "secdedv"):
            line["formats"] = {"coo_" + args.mode: {"skipped": "CSR-only mode"}}
        else:
            try:
                line["formats"] = {"coo_" + args.mode: coo_brief(pkg, t_primary, args.mode, min(K, 100), peak)}
            except Exception as e:      # an optional section must not take the headline down with it
                line["formats"] = {"coo_" + args.mode: {"error": repr(e)[:300]}}

    # ---- CPU baseline: the reference's CPU target on this box's host cores, bounded sample ----
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        try:
            t_all = t_primary if args.workload == "poisson3d" else generate(args, mg, args.workload)
            res = cpu_reference_run(args, args.mode, t_all, b_full, args.cpu_iters, 1)
            line["cpu_baseline"] = {"value": res["it_s"], "unit": "it/s", "cores": res["threads"], "kind": res["kind"],
                                    "sample": f"{args.cpu_iters} CG iterations on the full matrix "
                                              f"({res['seconds']:.1f} s; create_matrix {res['create_s']:.1f} s untimed), "
                                              f"OpenMP threads = {res['threads']} of {os.cpu_count()} host CPUs"}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "it/s", "cores": os.cpu_count(), "kind": "unavailable",
                                    "sample": str(e)}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
