"""The LOGIC of the CSR kernels (abft-sparse-cg_b200/csrc/spmv.cuh, encode.cuh, vecops.cuh), executed on the CPU.

There is no GPU in the build container, so the CPU suite checks the flagship kernels this way too: the product's own
kernel source is compiled by g++ against a multi-threaded shim of the CUDA built-ins (tests/helpers/cuda_shim_mt/:
a CTA runs as 256 real threads, __syncthreads and the warp collectives are barriers, the mbarrier / bulk-copy PTX
helpers are a small state machine + memcpy) and compared with the oracle, which is pinned to the reference.  The work
decomposition comes from the real library's abft_cuda_plan_tiles (host logic).  Same bars as tests/test_gpu_parity.py.
What this cannot see: memory ordering on the device, TMA alignment rules, launch configuration, performance.

"""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"]   # (secdedv = index 7: tests/test_vector_protection.py)
FATAL = {1, 4, 5, 6, 7, 8, 9, 11, 12, 13}   # every kind but CORRECTED_BIT (2), CORRECTED_PARITY (3), ROWPTR_CORRECTED (10)
ROWPTR_EVENTS = {10, 11}               # reported ahead of everything else, by pair (csrc/rowsecc.cuh)
ROW_EVENTS = {5, 6, 9}                 # ROW_SIZE, ROW_ORDER, CRC_ROW
REGULAR = 1 << 16
SLOW = os.environ.get("ABFT_SLOW_TESTS") == "1"   # more modes / cases / variant combinations than the default run


class Event(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_uint32), ("bit", C.c_uint32)]


def _build(tmp, name, flags):
    so = str(tmp / f"libcsr_emul_{name}.so")
    h = os.path.join(ROOT, "tests", "helpers")
    # -fno-gnu-unique: the kernels' "shared memory" is function-local statics of templates; as GNU-unique symbols they
    # would be merged across the differently configured builds loaded into this one process
    subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-fno-gnu-unique", "-w", "-pthread", *flags,
                    "-I" + os.path.join(h, "cuda_shim_mt"), "-o", so, os.path.join(h, "csr_emul.cpp")], check=True)
    L = C.CDLL(so)
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.emu_create.restype = C.c_void_p
    L.emu_create.argtypes = [C.c_int, u32p, u32p, f64p, C.c_int, C.c_int]
    L.emu_create_part.restype = C.c_void_p
    L.emu_create_part.argtypes = [C.c_int, u32p, u32p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.emu_destroy.argtypes = [C.c_void_p]
    L.emu_set_tiles.argtypes = [C.c_void_p, u32p, C.c_int, C.c_int, C.c_int]
    for f, t in ((L.emu_cols, C.c_uint32), (L.emu_rowptr, C.c_uint32), (L.emu_values, C.c_double)):
        f.restype = C.POINTER(t)
        f.argtypes = [C.c_void_p]
    L.emu_flip.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.emu_spmv.restype = C.c_int
    L.emu_spmv.argtypes = [C.c_void_p, f64p, f64p, C.POINTER(Event), C.c_int, C.POINTER(C.c_double)]
    L.emu_dot.restype = C.c_double
    L.emu_dot.argtypes = [C.c_void_p, f64p, f64p, C.c_int, C.c_int]
    L.emu_calc_xr.restype = C.c_double
    L.emu_calc_xr.argtypes = [C.c_void_p, f64p, f64p, f64p, f64p, C.c_double, C.c_int, C.c_int]
    L.emu_calc_p.argtypes = [C.c_void_p, f64p, f64p, C.c_double, C.c_int, C.c_int]
    L.emu_vec_encode.argtypes = [f64p, C.c_int]
    L.emu_vec_events.restype = C.c_int
    L.emu_vec_events.argtypes = [C.c_void_p, C.POINTER(Event), C.c_int]
    return L


@pytest.fixture(scope="module")
def EMU(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("csr_emu"), "shipped", [])


class EmuMatrix:
    """create_matrix + the tile plan, as abft_cuda_matrix_create_part does it."""

    def __init__(self, L, pkg, mode, t, grid=3, rows=None):
        """rows=(r0, r1): only that block of rows, as one rank of a row partition holds it (global columns and indices)"""
        if rows is None:
            cols, coo_rows, vals, n_rows, row0, nnz0 = t.cols, t.rows, t.vals, t.N, 0, 0
        else:
            r0, r1 = rows
            sel = (t.rows >= r0) & (t.rows < r1)
            cols, coo_rows, vals = t.cols[sel], t.rows[sel], t.vals[sel]
            n_rows, row0, nnz0 = r1 - r0, r0, int(np.count_nonzero(t.rows < r0))
        self.L, self.mode, self.N, self.nnz = L, mode, n_rows, int(cols.shape[0])
        self.h = L.emu_create_part(MODES.index(mode), np.ascontiguousarray(cols, dtype=np.uint32),
                                   np.ascontiguousarray(coo_rows, dtype=np.uint32), np.ascontiguousarray(vals, dtype=np.float64),
                                   n_rows, self.nnz, t.N, row0, nnz0)
        assert self.h
        lib = pkg.capi.load()
        rowptr = self.rowptr.copy()
        t = type("Part", (), {"N": n_rows})           # (below: the local row count)

        # the decomposition and the kernel (static round-robin or dynamic schedule), as create_matrix picks them
        kind = C.c_int(0)
        n = lib.abft_cuda_plan_matrix(rowptr.ctypes.data, t.N, MODES.index(mode), None, 0, C.byref(kind))
        tiles = np.zeros((max(n, 1), 4), dtype=np.uint32)
        assert lib.abft_cuda_plan_matrix(rowptr.ctypes.data, t.N, MODES.index(mode), tiles.ctypes.data, n, C.byref(kind)) == n
        tiles = tiles[:n]
        self.kind = kind.value
        self.dyn = self.kind == 1
        self.tiles = tiles
        L.emu_set_tiles(self.h, np.ascontiguousarray(tiles.reshape(-1)) if len(tiles) else np.zeros(4, np.uint32),
                        len(tiles), self.kind, grid)

    @property
    def cols(self):
        return np.ctypeslib.as_array(self.L.emu_cols(self.h), shape=(max(self.nnz, 1),))[:self.nnz]

    @property
    def rowptr(self):
        return np.ctypeslib.as_array(self.L.emu_rowptr(self.h), shape=(self.N + 1,))

    @property
    def values(self):
        return np.ctypeslib.as_array(self.L.emu_values(self.h), shape=(max(self.nnz, 1),))[:self.nnz]

    def flip(self, index, bit):
        self.L.emu_flip(self.h, index, bit)

    def spmv(self, vec):
        """-> (result, fatal, events, dot) with the events ordered and cut the way abft_cuda_poll_events does"""
        out = np.full(self.N, np.nan)
        log = (Event * 4096)()
        dot = C.c_double(0.0)
        n = self.L.emu_spmv(self.h, np.ascontiguousarray(vec, dtype=np.float64), out, log, 4096, C.byref(dot))
        assert n >= 0, n
        cons = self.mode == "constraints"

        def key(e):
            kind, index, bit = e
            rowev = kind in ROW_EVENTS
            if kind in ROWPTR_EVENTS:
                return (0, index, 0, 0, kind)
            return (1,) + (((index if rowev else bit), (0 if rowev else 1), index, kind) if cons else (index, 0, bit, kind))
        raw = sorted(((log[i].kind, log[i].index, log[i].bit) for i in range(n)), key=key)
        events, fatal = [], False
        for kind, index, bit in raw:
            if cons and kind in (7, 8):
                bit = 0                       # COL_* events carry their row in `bit` for ordering only
            if events and events[-1] == (kind, index, bit):
                continue
            events.append((kind, index, bit))
            if kind in FATAL:
                fatal = True
                break
        return out, fatal, events, dot.value

    def close(self):
        self.L.emu_destroy(self.h)


def shapes(mg):
    rng = np.random.default_rng(11)

    def ragged():
        N = 700
        lens = rng.integers(0, 10, N)
        lens[[0, 5, 6, 300]] = 0
        lens[-3:] = 0
        rows = np.repeat(np.arange(N), lens)
        cols = np.concatenate([np.sort(rng.choice(N, l, replace=False)) for l in lens if l > 0])
        return mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)

    def short_rows():
        N = 900
        lens = rng.integers(1, 4, N)
        rows = np.repeat(np.arange(N), lens)
        cols = np.concatenate([np.sort(rng.choice(N, l, replace=False)) for l in lens])
        return mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)

    def long_rows():
        N = 6000
        lens = np.full(N, 3)
        lens[[10, 3000, 5999]] = [4000, 2049, 2500]
        rows = np.repeat(np.arange(N), lens)
        cols = np.concatenate([np.sort(rng.choice(N, l, replace=False)) for l in lens])
        return mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)

    return {
        "p2d_40": lambda: mg.poisson2d(40),                 # 5-point: walks of 5, several tiles per CTA
        "p2d_12_b3": lambda: mg.poisson2d(12, 3),
        "p3d_12": lambda: mg.poisson3d(12),                 # 7-point: walks of 7
        "powerlaw": lambda: mg.powerlaw(3000, 60000, seed=2),                                   # dynamic schedule
        "powerlaw_dense": lambda: mg.powerlaw(1500, 100000, alpha=1.6, max_deg=800, seed=9),    # rows of 100s
        "ragged": ragged,
        "short_rows": short_rows,
        "long_rows": long_rows,                             # rows longer than a tile
        "one_by_one": lambda: mg.Triplets([0], [0], [2.5], 1),
    }


@pytest.mark.parametrize("mode", MODES)
def test_create_matrix_and_clean_spmv(EMU, pkg, O, mode):
    rng = np.random.default_rng(3)
    for name, build in shapes(pkg.matgen).items():
        t = build()
        em = EmuMatrix(EMU, pkg, mode, t)
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        assert np.array_equal(em.rowptr, om.rows), name
        assert np.array_equal(em.cols, om.cols), name
        assert np.array_equal(em.values.view(np.uint64), om.values.view(np.uint64)), name
        vec = rng.standard_normal(t.N)
        y, fatal, events, dot = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert not fatal and not fatal0 and events == [] and ev0 == [], (name, events)
        lens = np.bincount(t.rows, minlength=t.N)
        seq = lens <= 32          # sequential order kept: regular tiles, and rows of up to 32 elements anywhere
        assert np.array_equal(y[seq].view(np.uint64), y0[seq].view(np.uint64)), name
        scale = np.abs(t.vals).max() * np.abs(vec).max() * max(int(lens.max()), 1)
        assert np.max(np.abs(y - y0)) <= 1e-13 * scale, name
        exact = float(np.dot(vec, y0))
        assert abs(dot - exact) <= 1e-10 * max(np.abs(vec * y0).sum(), 1e-300), name
        # the schedule must not show: same bits with one CTA taking every tile
        em1 = EmuMatrix(EMU, pkg, mode, t, grid=1)
        y1, _, _, dot1 = em1.spmv(vec)
        assert np.array_equal(y1.view(np.uint64), y.view(np.uint64)), name
        if em.dyn:
            assert np.float64(dot1).view(np.uint64) == np.float64(dot).view(np.uint64), name   # per-tile partials
        em.close()
        em1.close()


@pytest.mark.parametrize("mode", ["sed", "sec7", "sec8", "secded", "crc32c"] if SLOW else ["sed", "secded", "crc32c"])
@pytest.mark.parametrize("shape", ["p3d_12", "powerlaw", "long_rows"])
def test_single_and_double_flips(EMU, pkg, O, mode, shape):
    t = shapes(pkg.matgen)[shape]()
    rng = np.random.default_rng(5)
    vec = rng.standard_normal(t.N)
    lens = np.bincount(t.rows, minlength=t.N)
    scale = np.abs(t.vals).max() * np.abs(vec).max() * max(int(lens.max()), 1)
    em = EmuMatrix(EMU, pkg, mode, t)
    om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    clean_c, clean_v = em.cols.copy(), em.values.copy()
    for trial in range(8 if SLOW else 6):
        index = int(rng.integers(0, t.nnz))
        bits = [int(b) for b in rng.choice(96, 1 + trial % 2, replace=False)]
        for b in bits:
            em.flip(index, b)
            om.flip(index, b)
        y, fatal, events, _ = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (shape, index, bits, events, ev0)
        if not fatal:
            assert np.array_equal(em.cols, om.cols) and np.array_equal(em.values.view(np.uint64), om.values.view(np.uint64))
            assert np.max(np.abs(y - y0)) <= 1e-13 * scale or not np.all(np.isfinite(y0)), (shape, index, bits)
        for arr, clean in ((em.cols, clean_c), (om.cols, clean_c), (em.values, clean_v), (om.values, clean_v)):
            arr[index] = clean[index]
    em.close()


def test_constraints_violations(EMU, pkg, O):
    """The four CONSTRAINTS checks, on regular and irregular tiles, incl. row pointers that leave the staged window."""
    rng = np.random.default_rng(6)
    for shape in ("p3d_12", "powerlaw", "long_rows"):
        t = shapes(pkg.matgen)[shape]()
        vec = rng.standard_normal(t.N)
        rowptr = np.concatenate([[0], np.cumsum(np.bincount(t.rows, minlength=t.N))])
        r = int(rng.integers(1, t.N - 2))
        cases = [("col", int(rowptr[r]), 1 << 30), ("col", int(rowptr[r]) + 1, 0), ("col", int(rowptr[r + 1]) - 1, 0),
                 ("row", r, t.nnz + 5), ("row", r + 1, int(rowptr[r]) - 1 if rowptr[r] > 0 else 0), ("row", r, 0),
                 ("row", t.N, t.nnz + 1)]
        if shape == "long_rows":
            # rows longer than a tile are walked from the tile's own range: their row pointers must still be looked at
            for lr in ((10, 3000) if SLOW else (3000,)):
                cases += [("row", lr + 1, t.nnz + 7), ("row", lr, int(rowptr[lr]) - 1), ("row", lr, int(rowptr[lr]) + 2),
                          ("row", lr + 1, int(rowptr[lr + 1]) - 3), ("row", lr + 1, int(rowptr[lr]) - 1),
                          ("col", int(rowptr[lr]) + 2100, 1 << 28), ("col", int(rowptr[lr]) + 17, 0)]
        for what, where, value in cases:
            em = EmuMatrix(EMU, pkg, "constraints", t)
            om = O.OracleMatrix("constraints", t.cols, t.rows, t.vals, t.N)
            if what == "col":
                em.cols[where] = value
                om.cols[where] = value
            else:
                em.rowptr[where] = value
                om.rows[where] = value
            y, fatal, events, _ = em.spmv(vec)
            y0, fatal0, ev0 = om.spmv(vec)
            assert (fatal, events) == (fatal0, ev0), (shape, what, where, value, events, ev0)
            if not fatal:
                lens = np.bincount(t.rows, minlength=t.N)
                scale = np.abs(t.vals).max() * np.abs(vec).max() * int(lens.max())
                assert np.max(np.abs(y - y0)) <= 1e-13 * scale, (shape, what, where, value)
            em.close()


@pytest.mark.parametrize("mode", ["none", "constraints", "sed", "sec7", "sec8", "secded"])
def test_fuzz_distribution(EMU, pkg, O, mode):
    """The differential-fuzz distribution of tools/fuzz_oracle_vs_ref.py (the oracle passes it against the live
    reference): multi-bit, multi-element and row-pointer faults on small random matrices."""
    spec = importlib.util.spec_from_file_location("fuzz_oracle_vs_ref", os.path.join(ROOT, "tools", "fuzz_oracle_vs_ref.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(77)
    pool = fz.matrices(pkg.matgen, rng)
    for trial in range(30):
        t = pool[int(rng.integers(0, len(pool)))]
        vec = rng.standard_normal(t.N)
        flips, pokes = fz.draw_faults(rng, mode, t)
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        fz.corrupt_arrays(om.cols, om.rows, om.values, flips, pokes)
        em = EmuMatrix(EMU, pkg, mode, t)
        fz.corrupt_arrays(em.cols, em.rowptr, em.values, flips, pokes)
        y0, fatal0, ev0 = om.spmv(vec)
        y, fatal, events, _ = em.spmv(vec)
        what = (mode, trial, flips, pokes)
        assert (fatal, events) == (fatal0, ev0), (what, events, ev0)
        if not fatal:
            assert np.array_equal(em.cols, om.cols) and np.array_equal(em.values.view(np.uint64), om.values.view(np.uint64)), what
            assert fz.same_result(y, y0, t, vec, om.cols, om.values, mode), what
        em.close()


@pytest.mark.parametrize("N", [1, 2, 7, 1000, 4099])
def test_vector_kernels(EMU, pkg, O, N):
    L = O.lib()
    rng = np.random.default_rng(N)
    x, r, p, w = (rng.standard_normal(N) for _ in range(4))
    t = pkg.matgen.Triplets([0], [0], [1.0], 1)
    em = EmuMatrix(EMU, pkg, "none", t)
    grid = 3
    assert abs(EMU.emu_dot(em.h, p, w, N, grid) - L.orc_dot(p, w, N)) <= 1e-12 * np.abs(p * w).sum()
    x2, r2, p2 = x.copy(), r.copy(), p.copy()
    xe, re_, pe = x.copy(), r.copy(), p.copy()
    rr = EMU.emu_calc_xr(em.h, xe, re_, p, w, 0.37, N, grid)
    rr0 = L.orc_calc_xr(x2, r2, p, w, 0.37, N)
    assert np.array_equal(xe, x2) and np.array_equal(re_, r2) and abs(rr - rr0) <= 1e-12 * rr0
    EMU.emu_calc_p(em.h, pe, re_, -1.25, N, grid)
    L.orc_calc_p(p2, r2, -1.25, N)
    assert np.array_equal(pe, p2)
    em.close()


@pytest.mark.parametrize("mode", ["none", "constraints", "sec7", "secded"])
def test_uniform_walk_of_full_warps(EMU, pkg, O, mode):
    """The unrolled walk of regular tiles (rows of up to 7 elements) drops the predicate on its adds when every row of the
    warp fills it: clean results bit for bit on stencils (full warps take the unpredicated adds) and on ragged / short
    rows (mixed warps must not), and the same events and repairs under faults."""
    rng = np.random.default_rng(21)
    for name in ("p3d_12", "p2d_40", "p2d_12_b3", "ragged", "short_rows", "one_by_one"):
        t = shapes(pkg.matgen)[name]()
        em = EmuMatrix(EMU, pkg, mode, t)
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        vec = rng.standard_normal(t.N)
        y, fatal, events, _ = em.spmv(vec)
        y0, _, _ = om.spmv(vec)
        assert not fatal and events == [] and np.array_equal(y.view(np.uint64), y0.view(np.uint64)), name
        if mode in ("sec7", "secded") and t.nnz > 10:
            for trial in range(4):
                index, bit = int(rng.integers(0, t.nnz)), int(rng.integers(0, 96))
                em.flip(index, bit)
                om.flip(index, bit)
                y, fatal, events, _ = em.spmv(vec)
                y0, fatal0, ev0 = om.spmv(vec)
                assert (fatal, events) == (fatal0, ev0), (name, index, bit)
                assert np.array_equal(em.cols, om.cols) and np.array_equal(y.view(np.uint64), y0.view(np.uint64)), (name, index, bit)
        em.close()


def test_crc_row_walk(EMU, pkg, O):
    """The spmv's CRC32C check (5-bit chunk tables of the three linear maps F, F.F, F.F.F, one entry per bank; unrolled row
    walk on regular tiles, the one-by-one loop elsewhere) against the oracle's byte-wise CRC: clean matrices pass (rows of
    >= 4 elements carry the CRC, shorter ones the SECDED byte), every single and double flip in a CRC-protected row is
    caught, in regular tiles and in rows longer than a tile."""
    L, chunk = EMU, 0
    rng = np.random.default_rng(31 + chunk)
    for name in ("p3d_12", "p2d_40", "ragged", "short_rows", "long_rows", "powerlaw"):
        t = shapes(pkg.matgen)[name]()
        em = EmuMatrix(L, pkg, "crc32c", t)
        om = O.OracleMatrix("crc32c", t.cols, t.rows, t.vals, t.N)
        assert np.array_equal(em.cols, om.cols), name            # (encode uses the byte tables in every build)
        vec = rng.standard_normal(t.N)
        y, fatal, events, _ = em.spmv(vec)
        y0, _, _ = om.spmv(vec)
        lens = np.bincount(t.rows, minlength=t.N)
        scale = np.abs(t.vals).max() * np.abs(vec).max() * int(lens.max())
        assert not fatal and events == [] and np.max(np.abs(y - y0)) <= 1e-13 * scale, (name, events)
        assert np.array_equal(y[lens <= 32].view(np.uint64), y0[lens <= 32].view(np.uint64)), name
        clean_c, clean_v = em.cols.copy(), em.values.copy()
        for trial in range(6):
            index = int(rng.integers(0, t.nnz))
            for b in rng.choice(96, 1 + trial % 2, replace=False):
                em.flip(index, int(b))
                om.flip(index, int(b))
            _, fatal, events, _ = em.spmv(vec)
            _, fatal0, ev0 = om.spmv(vec)
            assert (fatal, events) == (fatal0, ev0), (name, index, events, ev0)
            for arr, clean in ((em.cols, clean_c), (om.cols, clean_c), (em.values, clean_v), (om.values, clean_v)):
                arr[index] = clean[index]
        em.close()


def test_cg_trajectory_with_the_emulated_kernels(EMU, pkg, O):
    """cg.cpp:83-118 driven over the emulated spmv (with its fused dot), calc_xr and calc_p -- tree reductions and all --
    against the oracle's sequential loop: one injected flip is repaired on the way, the iteration count is equal and
    the residual history agrees to 1e-10 relative (the tolerance BASELINE.json states for reordered reductions)."""
    t = pkg.matgen.poisson2d(24)
    b = pkg.matgen.rhs(t.N)
    N = t.N
    em = EmuMatrix(EMU, pkg, "secded", t)
    om = O.OracleMatrix("secded", t.cols, t.rows, t.vals, t.N)
    em.flip(1234, 37)
    om.flip(1234, 37)
    ref = om.cg(b, 500, 1e-10)
    x, r, p = np.zeros(N), b.copy(), b.copy()
    rr = EMU.emu_dot(em.h, r, r, N, 3)
    hist, lines = [], []
    for itr in range(500):
        if rr <= 1e-10:
            break
        w, fatal, events, pw = em.spmv(p)
        assert not fatal
        lines += [O.format_event(*e) for e in events]
        rr_new = EMU.emu_calc_xr(em.h, x, r, p, w, rr / pw, N, 3)
        EMU.emu_calc_p(em.h, p, r, rr_new / rr, N, 3)
        rr = rr_new
        hist.append(rr)
    assert lines == [O.format_event(*e) for e in ref["events"]] == ["[ECC] corrected bit 37 at index 1234"]
    assert len(hist) == ref["iterations"]
    assert np.max(np.abs(np.array(hist) - ref["rr_hist"]) / ref["rr_hist"]) <= 1e-10
    assert np.max(np.abs(x - ref["x"])) <= 1e-9 * np.max(np.abs(ref["x"]))
    em.close()


@pytest.mark.parametrize("mode", ["constraints", "sec7", "secded", "crc32c"])
def test_row_partitions_report_global_indices(EMU, pkg, O, mode):
    """A matrix cut into row blocks (what each rank of a multi-GPU run holds: global columns, global element indices):
    every block's result slice equals the whole matrix's, and faults carry GLOBAL row / element numbers, so that the
    per-rank lists put end to end are the single-GPU list."""
    rng = np.random.default_rng(51)
    for name in ("p3d_12", "powerlaw", "long_rows"):
        t = shapes(pkg.matgen)[name]()
        vec = rng.standard_normal(t.N)
        cuts = [0] + sorted(int(c) for c in rng.choice(np.arange(1, t.N), 2, replace=False)) + [t.N]
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        parts = [EmuMatrix(EMU, pkg, mode, t, rows=(cuts[k], cuts[k + 1])) for k in range(3)]
        nnz0 = [int(np.count_nonzero(t.rows < cuts[k])) for k in range(4)]
        faults = []
        for trial in range(3):
            index = int(rng.integers(0, t.nnz))
            bit = int(rng.integers(0, 88)) if mode != "constraints" else int(rng.integers(64, 88))
            k = max(j for j in range(3) if nnz0[j] <= index)
            parts[k].flip(index - nnz0[k], bit)
            om.flip(index, bit)
            faults.append((index, bit))
        y0, fatal0, ev0 = om.spmv(vec)
        got_events, got_fatal, ys = [], False, []
        for k, em in enumerate(parts):
            y, fatal, events, _ = em.spmv(vec)
            ys.append(y)
            if not got_fatal:                      # (the host stops at the first fatal event in row order)
                got_events += events
                got_fatal = fatal
        assert (got_fatal, got_events) == (fatal0, ev0), (name, faults, got_events, ev0)
        if not fatal0:
            y = np.concatenate(ys)
            lens = np.bincount(t.rows, minlength=t.N)
            assert np.array_equal(y[lens <= 32].view(np.uint64), y0[lens <= 32].view(np.uint64)), name
        for em in parts:
            em.close()


def _from_lens(mg, lens, seed):
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, dtype=np.int64)
    N = int(max(len(lens), lens.max() if len(lens) else 1, 1))
    lens = np.concatenate([lens, np.zeros(N - len(lens), dtype=np.int64)])
    rows = np.repeat(np.arange(N), lens)
    cols = np.concatenate([np.sort(rng.choice(N, int(l), replace=False)) for l in lens if l > 0]) if lens.sum() else np.zeros(0)
    return mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)


EDGE_LENS = {
    "exactly_one_static_tile": [2048],            # one row as long as a static tile
    "dyn_tile_boundary": [1536, 1537, 3],         # at and just past the dynamic tile size
    "seq_threshold": [32, 33, 31, 0, 64],         # rows of 32 are summed sequentially, 33 by the warp
    "mostly_empty": [0, 0, 5, 0, 0, 0, 0, 7, 0],
    "trailing_empty": [3, 4] + [0] * 40,
    "row_limit": [1] * 257 + [6] * 300,           # tiles capped by the 256-row limit, then by elements
    "chunked_walk": [8, 9, 10, 11, 12] * 60,      # regular tiles whose rows are longer than the unrolled walks
    "zigzag": [1, 63] * 50,                       # irregular
    "no_elements": [0, 0, 0],
    "odd_sizes": [5] * 259,                       # row / element counts that are not multiples of the 16-byte windows
}


@pytest.mark.parametrize("mode", ["none", "constraints", "secded", "crc32c"])
def test_edge_shapes(EMU, pkg, O, mode):
    rng = np.random.default_rng(61)
    for name, lens in EDGE_LENS.items():
        t = _from_lens(pkg.matgen, lens, 7)
        em = EmuMatrix(EMU, pkg, mode, t)
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        assert np.array_equal(em.rowptr, om.rows) and np.array_equal(em.cols, om.cols), name
        vec = rng.standard_normal(t.N)
        y, fatal, events, dot = em.spmv(vec)
        y0, _, _ = om.spmv(vec)
        rl = np.bincount(t.rows, minlength=t.N)
        scale = max(np.abs(t.vals).max() if t.nnz else 0.0, 1e-300) * np.abs(vec).max() * max(int(rl.max()), 1)
        assert not fatal and events == [], (name, events)
        assert np.array_equal(y[rl <= 32].view(np.uint64), y0[rl <= 32].view(np.uint64)), name
        assert np.max(np.abs(y - y0)) <= 1e-13 * scale, name
        assert abs(dot - float(np.dot(vec, y0))) <= 1e-10 * max(np.abs(vec * y0).sum(), 1e-300), name
        if t.nnz and mode in ("secded", "crc32c"):
            for index in {0, t.nnz - 1, t.nnz // 2}:
                em.flip(index, 17)
                om.flip(index, 17)
                _, fatal, events, _ = em.spmv(vec)
                _, fatal0, ev0 = om.spmv(vec)
                assert (fatal, events) == (fatal0, ev0), (name, index)
                em.cols[index], em.values[index] = om.cols[index], om.values[index]
                if fatal0:          # (crc32c is detect-only: put the bit back on both sides)
                    em.flip(index, 17)
                    om.flip(index, 17)
        em.close()


def test_stencil_matrices_get_the_fixed_length_instantiation(EMU, pkg):
    """5- and 7-point stencil matrices run spmv_kernel<MODE, false, FLEN> (row walk length fixed at compile time, no
    long-row / irregular-tile paths); anything else keeps the general instantiation.  Every test of this file that uses
    p3d_12 / p2d_40 in a compute-bound mode therefore exercises the stencil kernels, ragged / short_rows the general one."""
    EMU.emu_flen.restype = C.c_int
    EMU.emu_flen.argtypes = [C.c_void_p]
    for name, want in (("p3d_12", 7), ("p2d_40", 5), ("p2d_12_b3", 5), ("ragged", 0), ("short_rows", 0), ("powerlaw", 0)):
        em = EmuMatrix(EMU, pkg, "secded", shapes(pkg.matgen)[name]())
        assert EMU.emu_flen(em.h) == want, name
        em.close()


def test_crc32c_register_tables_with_columns_up_to_2_24(EMU, pkg, O):
    """The stencil instantiation's CRC check cuts 18 five-bit chunks from the whole 88-bit element; the last ones hold
    column bits 16..23, which the small test matrices never set.  The last 160 rows of a 2^24-column matrix (one rank's
    block of a row partition), 7 elements each with columns all over [0, 2^24): clean, and with every bit of one
    element's column field flipped in turn, against the oracle."""
    N = 1 << 24
    r0 = N - 160
    rng = np.random.default_rng(24)
    rows = np.repeat(np.arange(r0, N), 7)
    cols = np.concatenate([np.sort(rng.choice(N, 7, replace=False)) for _ in range(160)])
    cols[-1] = N - 1                                                  # every column bit set
    t = pkg.matgen.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)
    vec = rng.standard_normal(N)
    em = EmuMatrix(EMU, pkg, "crc32c", t, rows=(r0, N))
    assert EMU.emu_flen(em.h) == 7 and int((em.cols & 0x00FFFFFF).max()) == N - 1
    om = O.OracleMatrix("crc32c", t.cols, t.rows, t.vals, N)
    assert np.array_equal(em.cols, om.cols)                           # same stored CRC bytes
    y, fatal, events, _ = em.spmv(vec)
    y0, fatal0, ev0 = om.spmv(vec)
    assert not fatal and events == [] == ev0
    assert np.array_equal(y.view(np.uint64), y0[r0:].view(np.uint64))
    index = 7 * 77 + 3
    for bit in range(64, 88):
        em.flip(index, bit); om.flip(index, bit)
        y, fatal, events, _ = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0) and fatal, bit        # [CRC32C] error detected in row r0 + 77
        em.flip(index, bit); om.flip(index, bit)
    em.close()
