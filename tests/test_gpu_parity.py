"""Parity of the CUDA target (through the C ABI) against the CPU oracle on the same
seeded inputs, and against the committed reference fixtures.

Bars (north_star): bit-exact for the integer work -- encoded column words, ECC
detection / correction events and flagged indices, corrected matrices; spmv results
are bit-exact as well (the kernel keeps the reference's per-row operation order and
uses no FMA); reductions (dot, r.r) are reordered, so iteration counts must be equal
and residuals agree to RTOL = 1e-10 relative.
"""
import ctypes
import re

import numpy as np
import pytest

from conftest import unhex, unhex1

pytestmark = pytest.mark.gpu

RTOL = 1e-10
ALL_MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"]
ECC_MODES = ["sed", "sec7", "sec8", "secded"]


def shapes(mg):
    rng = np.random.default_rng(11)

    def ragged():
        # rows of length 0..9, some empty, trailing empty rows, random columns
        N = 700
        lens = rng.integers(0, 10, N)
        lens[[0, 5, 6, 300]] = 0
        lens[-3:] = 0
        rows = np.repeat(np.arange(N), lens)
        cols = np.concatenate([np.sort(rng.choice(N, l, replace=False)) for l in lens if l > 0])
        return mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)

    def short_rows():
        # 1..3 entries per row: exercises the short-row fallback of crc32c and tiles capped by row count
        N = 1500
        lens = rng.integers(1, 4, N)
        rows = np.repeat(np.arange(N), lens)
        cols = np.concatenate([np.sort(rng.choice(N, l, replace=False)) for l in lens])
        return mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)

    return {
        "p2d_40": lambda: mg.poisson2d(40),
        "p2d_16_b5": lambda: mg.poisson2d(16, 5),
        "p3d_12": lambda: mg.poisson3d(12),
        "powerlaw": lambda: mg.powerlaw(3000, 60000, seed=2),
        "powerlaw_dense": lambda: mg.powerlaw(6000, 700000, alpha=1.6, max_deg=1500, seed=9),   # full tiles, rows of 100s
        "ragged": ragged,
        "short_rows": short_rows,
        "one_by_one": lambda: mg.Triplets([0], [0], [2.5], 1),
    }


def make(pkg, O, mode, t):
    ctx = pkg.CUDAContext(mode)
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    return ctx, A, om


def gpu_spmv(ctx, A, vec, n_out=None):
    n_out = A.N if n_out is None else n_out
    v = ctx.create_vector(len(vec))
    w = ctx.create_vector(n_out)
    ctx.upload(v, vec)
    ctx.spmv(A, v, w)
    events, fatal = ctx.poll_events()
    return ctx.download(w), fatal, events, (v, w)


# ---------------------------------------------------------------- create_matrix
@pytest.mark.parametrize("mode", ALL_MODES)
def test_create_matrix_bit_exact(pkg, O, mode):
    for name, build in shapes(pkg.matgen).items():
        t = build()
        ctx, A, om = make(pkg, O, mode, t)
        cols, rowptr, vals = ctx.read_matrix(A)
        assert np.array_equal(rowptr, om.rows), name
        assert np.array_equal(cols, om.cols), name
        assert np.array_equal(vals.view(np.uint64), om.values.view(np.uint64)), name
        ctx.close()


def test_ecc_modes_reject_wide_matrices(pkg):
    # column index is 24 bits in every ECC mode (CSR/CPUContext.cpp:238,282,338,404)
    ctx = pkg.CUDAContext("secded")
    with pytest.raises(RuntimeError, match="2\\^24"):
        ctx.create_matrix([0], [0], [1.0], (1 << 24) + 1)
    ctx.close()
    ctx = pkg.CUDAContext("none")
    A = ctx.create_matrix([0], [0], [1.0], (1 << 24) + 1)   # fine without check bits
    assert A.nnz == 1
    ctx.close()


def test_empty_matrix(pkg):
    ctx = pkg.CUDAContext("secded")
    A = ctx.create_matrix(np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0), 5, 0)
    v, w = ctx.create_vector(5), ctx.create_vector(5)
    ctx.upload(v, np.arange(5.0)); ctx.upload(w, np.ones(5))
    ctx.spmv(A, v, w)
    assert ctx.dot(v, w) == 0.0
    assert np.array_equal(ctx.download(w), np.zeros(5))
    ctx.close()


# ---------------------------------------------------------------- spmv
@pytest.mark.parametrize("mode", ALL_MODES)
def test_spmv_bit_exact(pkg, O, mode):
    rng = np.random.default_rng(3)
    for name, build in shapes(pkg.matgen).items():
        t = build()
        ctx, A, om = make(pkg, O, mode, t)
        vec = rng.standard_normal(t.N)
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert not fatal and not fatal0 and events == [] and ev0 == [], name
        # bit-exact wherever the kernel keeps the reference's sequential order: every row of a regular tile and
        # rows of up to 32 elements anywhere; longer rows of irregular (power-law) tiles are reduced by a fixed
        # tree -- deterministic, but a different order
        lens = np.bincount(t.rows, minlength=t.N)
        seq = lens <= 32
        assert np.array_equal(y[seq].view(np.uint64), y0[seq].view(np.uint64)), name
        scale = np.abs(t.vals).max() * np.abs(vec).max() * max(int(lens.max()), 1)
        assert np.max(np.abs(y - y0)) <= 1e-13 * scale, name
        ctx.close()


def test_spmv_long_rows(pkg, O):
    """Rows longer than a tile are streamed and tree-reduced: not bit-exact, within RTOL."""
    mg = pkg.matgen
    rng = np.random.default_rng(8)
    N = 9000
    lens = np.full(N, 3)
    lens[[10, 4000, 8999]] = [5000, 2049, 8000]
    rows = np.repeat(np.arange(N), lens)
    cols = np.concatenate([np.sort(rng.choice(N, l, replace=False)) for l in lens])
    t = mg.Triplets(cols, rows, rng.standard_normal(rows.shape[0]), N)
    vec = rng.standard_normal(N)
    for mode in ALL_MODES:
        ctx, A, om = make(pkg, O, mode, t)
        cols_d, rowptr_d, _ = ctx.read_matrix(A)
        assert np.array_equal(cols_d, om.cols)
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, _, _ = om.spmv(vec)
        assert not fatal and events == []
        short = lens < 2049
        assert np.array_equal(y[short], y0[short])
        scale = np.abs(t.vals).max() * np.abs(vec).max() * lens.max()
        assert np.max(np.abs(y - y0)) <= 1e-13 * scale
        ctx.close()


def _irregular_big(mg):
    # ~2000 tiles of every kind (regular, irregular, rows longer than a tile): several tiles per CTA,
    # drawn from the tile counter of the dynamic schedule
    return mg.powerlaw(200_000, 4_000_000, max_deg=3000, seed=6)


@pytest.mark.parametrize("mode", ["none", "constraints", "sec7", "secded"])
def test_spmv_dynamic_schedule(pkg, O, mode):
    """Matrices with irregular tiles are scheduled dynamically (CTAs draw tiles from a counter): result[] and the
    fused dot(vec, result) must not depend on which CTA ran which tile -- bitwise equal from launch to launch --
    and the counter must be ready again for the next launch."""
    t = _irregular_big(pkg.matgen)
    vec = np.random.default_rng(5).standard_normal(t.N)
    ctx, A, om = make(pkg, O, mode, t)
    assert ctx.L.abft_cuda_matrix_tiles(A.handle) > 3 * 592
    v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
    ctx.upload(v, vec)
    runs = []
    for _ in range(3):
        ctx.upload(w, np.zeros(t.N))
        ctx.spmv(A, v, w)
        runs.append((ctx.download(w).view(np.uint64).copy(), np.float64(ctx.dot(v, w)).view(np.uint64)))
    events, fatal = ctx.poll_events()
    assert not fatal and events == []
    for y_bits, dot_bits in runs[1:]:
        assert np.array_equal(y_bits, runs[0][0]) and dot_bits == runs[0][1]
    y = runs[0][0].view(np.float64)
    y0, _, _ = om.spmv(vec)
    lens = np.bincount(t.rows, minlength=t.N)
    seq = lens <= 32
    assert np.array_equal(y[seq].view(np.uint64), y0[seq].view(np.uint64))
    scale = np.abs(t.vals).max() * np.abs(vec).max() * int(lens.max())
    assert np.max(np.abs(y - y0)) <= 1e-13 * scale
    exact = O.lib().orc_dot(vec, y0, t.N)
    assert abs(runs[0][1].view(np.float64) - exact) <= RTOL * abs(exact)
    ctx.close()


@pytest.mark.parametrize("mode", ["sed", "sec7", "sec8", "secded"])
def test_faults_in_irregular_tiles(pkg, O, mode):
    """Single flips in short rows, in rows longer than 32 and in rows longer than a tile of a power-law matrix:
    same events, same repaired matrix as the oracle."""
    t = _irregular_big(pkg.matgen)
    vec = np.random.default_rng(6).standard_normal(t.N)
    lens = np.bincount(t.rows, minlength=t.N)
    rowptr = np.concatenate([[0], np.cumsum(lens)])
    short = int(np.flatnonzero((lens > 3) & (lens <= 32))[100])
    mid = int(np.flatnonzero((lens > 100) & (lens <= 2048))[10])
    longr = int(np.flatnonzero(lens > 2048)[0])
    targets = [(int(rowptr[short]) + 2, 5), (int(rowptr[mid]) + 77, 40), (int(rowptr[longr]) + 2100, 70)]
    for index, bit in targets if mode != "sed" else targets[:1]:
        ctx, A, om = make(pkg, O, mode, t)
        ctx.flip_bit(A, index, bit); om.flip(index, bit)
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0) and len(events) == 1, (index, bit)
        cols, _, vals = ctx.read_matrix(A)
        assert np.array_equal(cols, om.cols) and np.array_equal(vals.view(np.uint64), om.values.view(np.uint64))
        if not fatal:
            scale = np.abs(t.vals).max() * np.abs(vec).max() * int(lens.max())
            assert np.max(np.abs(y - y0)) <= 1e-13 * scale
        ctx.close()


def test_constraints_in_irregular_tiles(pkg, O):
    """Column faults CONSTRAINTS must catch in the rows of a power-law matrix (short, > 32 elements, longer than
    a tile), and row-pointer faults that send a row outside its tile."""
    t = _irregular_big(pkg.matgen)
    vec = np.random.default_rng(7).standard_normal(t.N)
    lens = np.bincount(t.rows, minlength=t.N)
    rowptr = np.concatenate([[0], np.cumsum(lens)])
    short = int(np.flatnonzero((lens > 3) & (lens <= 32))[200])
    mid = int(np.flatnonzero((lens > 100) & (lens <= 2048))[20])
    longr = int(np.flatnonzero(lens > 2048)[1])
    col_pokes = [(int(rowptr[short]) + 1, 1 << 30), (int(rowptr[short]) + 2, 0),        # too large / out of order
                 (int(rowptr[mid]) + 50, 1 << 29), (int(rowptr[mid]) + 99, 0),
                 (int(rowptr[longr]) + 2500, 1 << 28), (int(rowptr[longr]) + 10, 0)]
    for index, value in col_pokes:
        ctx, A, om = make(pkg, O, "constraints", t)
        ctx._ck(ctx.L.abft_cuda_matrix_poke_col(ctx.ctx, A.handle, index, value), "poke")
        om.cols[index] = value
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0) and fatal, (index, value)
        ctx.close()
    # (the last two: the pointers of a row longer than a tile, which is walked from the tile's own range --
    # found by running the kernel's logic on the CPU, tests/test_csr_kernel_logic.py)
    for row, value in [(short, t.nnz + 5), (mid, 3), (mid + 1, int(rowptr[mid]) - 1),
                       (longr + 1, t.nnz + 5), (longr + 1, int(rowptr[longr]) - 1)]:
        ctx, A, om = make(pkg, O, "constraints", t)
        ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, row, value), "poke")
        om.rows[row] = value
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0) and fatal, (row, value)
        ctx.close()


def test_fused_dot_equals_standalone_dot(pkg, O):
    t = pkg.matgen.poisson3d(20)
    rng = np.random.default_rng(1)
    vec = rng.standard_normal(t.N)
    ctx, A, om = make(pkg, O, "secded", t)
    y, _, _, (v, w) = gpu_spmv(ctx, A, vec)
    fused = ctx.dot(v, w)                       # served by the spmv's epilogue
    n0 = ctx.launch_count()
    fused_again = ctx.dot(w, v)                 # operand order does not matter
    assert ctx.launch_count() == n0 and fused_again == fused
    ctx.copy_vector(w, w)                       # any write invalidates the cached value
    standalone = ctx.dot(v, w)
    assert ctx.launch_count() == n0 + 1
    exact = O.lib().orc_dot(vec, y, t.N)
    assert abs(fused - exact) <= RTOL * abs(exact) and abs(standalone - exact) <= RTOL * abs(exact)
    ctx.close()


# ---------------------------------------------------------------- vector kernels
def test_vector_transfers_into_caller_buffers(pkg):
    """abft_cuda_vector_download_async / upload_async with a buffer from abft_cuda_host_alloc (what the plugin's
    map_vector / unmap_vector do with several GPUs under one host thread: one host image, one slice per GPU) move
    the same bytes as map / unmap, and an upload invalidates a cached fused dot like any other write."""
    import ctypes as C
    N = 70001
    ctx = pkg.CUDAContext("secded")
    L = ctx.L
    L.abft_cuda_host_alloc.restype = C.c_void_p
    buf = L.abft_cuda_host_alloc(C.c_size_t(8 * 2 * N))
    assert buf
    host = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(2 * N,))
    data = np.random.default_rng(3).standard_normal(N)
    host[:N] = data
    host[N:] = 0.0
    a, b = ctx.create_vector(N), ctx.create_vector(N)
    assert L.abft_cuda_vector_upload_async(ctx.ctx, a.handle, C.c_void_p(buf)) == 0
    ctx.copy_vector(b, a)
    assert L.abft_cuda_vector_download_async(ctx.ctx, b.handle, C.c_void_p(buf + 8 * N)) == 0
    ctx.sync()
    assert np.array_equal(host[N:].view(np.uint64), data.view(np.uint64))
    assert np.array_equal(ctx.download(a).view(np.uint64), data.view(np.uint64))
    d0 = ctx.dot(a, b)
    host[:N] = 2.0 * data
    assert L.abft_cuda_vector_upload_async(ctx.ctx, a.handle, C.c_void_p(buf)) == 0
    ctx.sync()
    d1 = ctx.dot(a, b)
    assert abs(d1 - 2.0 * d0) <= 1e-12 * abs(d0)
    L.abft_cuda_host_free(C.c_void_p(buf))
    ctx.close()


@pytest.mark.parametrize("N", [1, 2, 7, 1000, 65537, 1 << 20])
def test_vector_kernels(pkg, O, N):
    L = O.lib()
    rng = np.random.default_rng(N)
    x, r, p, w = (rng.standard_normal(N) for _ in range(4))
    ctx = pkg.CUDAContext("none")
    vx, vr, vp, vw = (ctx.create_vector(N) for _ in range(4))
    for v, d in zip((vx, vr, vp, vw), (x, r, p, w)):
        ctx.upload(v, d)
    d_gpu, d_cpu = ctx.dot(vp, vw), L.orc_dot(p, w, N)
    assert abs(d_gpu - d_cpu) <= 1e-12 * np.abs(p * w).sum()
    x2, r2, p2 = x.copy(), r.copy(), p.copy()
    rr_gpu = ctx.calc_xr(vx, vr, vp, vw, 0.37)
    rr_cpu = L.orc_calc_xr(x2, r2, p, w, 0.37, N)
    assert np.array_equal(ctx.download(vx), x2) and np.array_equal(ctx.download(vr), r2)   # elementwise: bit-exact
    assert abs(rr_gpu - rr_cpu) <= 1e-12 * rr_cpu
    ctx.calc_p(vp, vr, -1.25)
    L.orc_calc_p(p2, r2, -1.25, N)
    assert np.array_equal(ctx.download(vp), p2)
    # copy_vector + map/unmap round trip
    ctx.copy_vector(vw, vx)
    assert np.array_equal(ctx.download(vw), x2)
    h = ctx.map_vector(vw); h[:] = 3.0; ctx.unmap_vector(vw, h)
    assert np.array_equal(ctx.download(vw), np.full(N, 3.0))
    ctx.close()


def test_unaligned_views(pkg, O):
    """Odd offsets (multi-GPU row partitions) take the scalar path: same results."""
    L = O.lib()
    rng = np.random.default_rng(2)
    N = 4099
    big = [rng.standard_normal(N + 3) for _ in range(4)]
    ctx = pkg.CUDAContext("none")
    parents = [ctx.create_vector(N + 3) for _ in range(4)]
    for v, d in zip(parents, big):
        ctx.upload(v, d)
    vx, vr, vp, vw = (ctx.view_vector(v, 1, N) for v in parents)
    x, r, p, w = (d[1:N + 1].copy() for d in big)
    assert abs(ctx.dot(vp, vw) - L.orc_dot(p, w, N)) <= 1e-12 * np.abs(p * w).sum()
    rr = ctx.calc_xr(vx, vr, vp, vw, -0.5)
    assert abs(rr - L.orc_calc_xr(x, r, p, w, -0.5, N)) <= 1e-12 * rr
    assert np.array_equal(ctx.download(parents[0])[1:N + 1], x) and np.array_equal(ctx.download(parents[1])[1:N + 1], r)
    assert ctx.download(parents[0])[0] == big[0][0] and ctx.download(parents[0])[N + 1] == big[0][N + 1]
    ctx.close()


# ---------------------------------------------------------------- CG
def run_cg(ctx, A, b, max_itrs, conv):
    N = len(b)
    vb, x, r, p, w = (ctx.create_vector(N) for _ in range(5))
    ctx.upload(vb, b)
    it, hist = ctx.cg(A, vb, x, r, p, w, max_itrs, conv)
    return it, hist, ctx.download(x), (vb, x, r, p, w)


@pytest.mark.parametrize("mode", ALL_MODES)
def test_cg_poisson2d_256(pkg, O, mode):
    """BASELINE configs 1/2: 256x256 Poisson, -c 1e-10 -> 683 iterations in every mode."""
    t = pkg.matgen.poisson2d(256)
    b = pkg.matgen.rhs(t.N)
    ctx, A, om = make(pkg, O, mode, t)
    it, hist, x, _ = run_cg(ctx, A, b, 1000, 1e-10)
    ref = om.cg(b, 1000, 1e-10)
    assert ref["iterations"] == 683
    assert it == ref["iterations"]
    assert np.max(np.abs(hist - ref["rr_hist"]) / ref["rr_hist"]) <= RTOL
    assert np.max(np.abs(x - ref["x"])) <= RTOL * np.max(np.abs(ref["x"]))
    assert ctx.messages == []
    ctx.close()


def test_cg_matches_reference_fixture(pkg, golden_cg):
    """Against the committed reference trajectory (no oracle involved)."""
    t = pkg.matgen.poisson2d(golden_cg["grid"])
    b = unhex(golden_cg["b"])
    for mode, g in golden_cg["modes"].items():
        ctx = pkg.CUDAContext(mode)
        A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
        cols, _, _ = ctx.read_matrix(A)
        assert [int(z) for z in cols[:16]] == g["cols_head"] and int(cols.astype(np.uint64).sum()) == g["cols_sum"]
        vb, vw = ctx.create_vector(t.N), ctx.create_vector(t.N)
        ctx.upload(vb, b)
        ctx.spmv(A, vb, vw)
        assert np.array_equal(ctx.download(vw), unhex(g["spmv_b"]))
        it, hist, x, _ = run_cg(ctx, A, b, 1000, 1e-10)
        want = unhex(g["rr_hist"])
        assert it == g["iterations"]
        assert np.max(np.abs(hist - want) / want) <= RTOL
        ctx.close()


def test_device_resident_loop_equals_host_loop(pkg, O):
    t = pkg.matgen.poisson2d(64, 2)
    b = pkg.matgen.rhs(t.N)
    for mode in ["none", "secded", "crc32c"]:
        ctx, A, om = make(pkg, O, mode, t)
        it, hist, x, vecs = run_cg(ctx, A, b, 1000, 1e-10)
        vb2, x2, r2, p2, w2 = (ctx.create_vector(t.N) for _ in range(5))
        ctx.upload(vb2, b)
        for every in (1, 7, 64):
            ctx.upload(x2, np.zeros(t.N))
            it2, rr2 = ctx.cg_solve_device(A, vb2, x2, r2, p2, w2, 1000, 1e-10, check_every=every)
            assert it2 == it and rr2 == hist[-1]
            assert np.array_equal(ctx.download(x2), x)          # same kernels, same scalars: bit-identical
        # iteration cap
        ctx.upload(x2, np.zeros(t.N))
        it3, rr3 = ctx.cg_solve_device(A, vb2, x2, r2, p2, w2, 10, 0.0, check_every=4)
        assert it3 == 10 and rr3 == hist[9]
        ctx.close()


# ---------------------------------------------------------------- injected faults
@pytest.mark.parametrize("mode", ECC_MODES + ["crc32c", "constraints", "none"])
def test_all_96_single_flips(pkg, O, mode):
    """Every bit of one element: identical events, identical corrected matrix, identical result."""
    t = pkg.matgen.poisson2d(12)
    vec = pkg.matgen.rhs(t.N)
    for index in (77, 0, t.nnz - 1):
        for bit in range(96):
            ctx, A, om = make(pkg, O, mode, t)
            ctx.flip_bit(A, index, bit)
            om.flip(index, bit)
            y, fatal, events, _ = gpu_spmv(ctx, A, vec)
            y0, fatal0, ev0 = om.spmv(vec)
            assert (fatal, events) == (fatal0, ev0), (index, bit)
            cols, _, vals = ctx.read_matrix(A)
            assert np.array_equal(cols, om.cols) and np.array_equal(vals.view(np.uint64), om.values.view(np.uint64))
            if not fatal:
                # (an undetected index flip can point outside vec[]: the reference reads out of bounds there,
                # oracle and CUDA target both clamp the column to the last entry)
                assert np.array_equal(y, y0, equal_nan=True), (index, bit)
            # a second spmv finds nothing new in the correcting modes
            if mode in ("sec7", "sec8", "secded"):
                v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
                ctx.upload(v, vec); ctx.spmv(A, v, w)
                assert ctx.poll_events() == ([], False)
            ctx.close()


@pytest.mark.parametrize("mode", ["sec7", "sec8", "secded", "sed"])
def test_double_flips(pkg, O, mode):
    """Pairs of flips in one element.  SECDED: always 'double-bit error detected'.
    SEC7/SEC8 mis-correct; events must still equal the oracle's."""
    t = pkg.matgen.poisson2d(10)
    vec = pkg.matgen.rhs(t.N)
    rng = np.random.default_rng(5)
    pairs = [(43, 44), (9, 18), (0, 95), (88, 3), (88, 95), (63, 64)] + [tuple(rng.choice(96, 2, replace=False)) for _ in range(60)]
    for a, b in pairs:
        ctx, A, om = make(pkg, O, mode, t)
        for bit in (int(a), int(b)):
            ctx.flip_bit(A, 123, bit); om.flip(123, bit)
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (a, b)
        if mode == "secded":
            assert fatal and events[-1][0] == 4
        cols, _, vals = ctx.read_matrix(A)
        assert np.array_equal(cols, om.cols) and np.array_equal(vals.view(np.uint64), om.values.view(np.uint64))
        ctx.close()


def test_faults_match_reference_fixture(pkg, golden_cg):
    """The committed single-flip cases of the reference: same lines, same iteration count, matrix restored."""
    t = pkg.matgen.poisson2d(golden_cg["grid"])
    b = unhex(golden_cg["b"])
    for f in golden_cg["faults"]:
        ctx = pkg.CUDAContext(f["mode"])
        A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
        clean_cols, _, clean_vals = ctx.read_matrix(A)
        ctx.flip_bit(A, f["index"], f["bit"])
        it, hist, x, _ = run_cg(ctx, A, b, 1000, 1e-10)
        assert ctx.messages == f["lines"], f
        assert it == f["iterations"]
        want = unhex1(f["rr_last"])
        assert abs(hist[-1] - want) <= RTOL * want
        cols, _, vals = ctx.read_matrix(A)
        assert np.array_equal(cols, clean_cols) and np.array_equal(vals, clean_vals)
        ctx.close()


def test_cli_fixture_replay(pkg, golden_cli):
    """The reference CLI's pinned-seed runs (-x with srand(time) pinned): the CUDA target, drawing rand()
    in the reference's order, flips the same bits, prints the same lines, stops or converges the same."""
    libc = ctypes.CDLL(None)
    t = pkg.matgen.poisson2d(golden_cli["grid"], golden_cli["blocks"])
    b = pkg.matgen.rhs(t.N)
    for case in golden_cli["cases"]:
        ctx = pkg.CUDAContext(case["mode"])
        A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
        if "-x" in case["extra"]:
            n, kind = 1, 0
            for a in case["extra"][1:]:
                if a == "VALUE": kind = 1
                elif a == "INDEX": kind = 2
                else: n = int(a)
            libc.srand(case["seed"])
            ctx.inject_bitflip(A, kind, n)
        try:
            it, hist, x, _ = run_cg(ctx, A, b, 300, 1e-10)
            code = 0
        except pkg.FatalFault:
            it, code = None, 1
        assert ctx.messages == case["faults"], case
        assert code == case["exit"], case
        if case["ran"]:
            assert case["ran"] == f"ran for {it} iterations"
        ctx.close()


def test_constraints_row_pointer_faults(pkg, O):
    """rows[] corruption is only checked by CONSTRAINTS (CSR/CPUContext.cpp:173-182)."""
    t = pkg.matgen.poisson2d(30)
    vec = pkg.matgen.rhs(t.N)
    pokes = [(17, 1 << 30), (400, 0), (400, 3), (899, t.nnz + 1), (5, 21), (200, 980)]
    for row, value in pokes:
        ctx, A, om = make(pkg, O, "constraints", t)
        ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, row, value), "poke")
        om.rows[row] = value
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (row, value)
        assert fatal
        ctx.close()


@pytest.mark.parametrize("mode", ["none", "constraints", "secded", "crc32c"])
def test_row_pointer_fault_that_lengthens_a_row_inside_its_tile(pkg, O, mode):
    """A corrupted rows[] entry that makes a row LONGER than the tile's build-time longest row while the columns stay
    increasing across the old row boundary: the reference walks the whole [start, end) range
    (CSR/CPUContext.cpp:125,184), so must the unrolled fast walk (it used to stop after max_len elements)."""
    N = 64
    idx = np.arange(N, dtype=np.uint32)
    for t, row, value in [(pkg.matgen.Triplets(idx, idx, np.arange(1.0, N + 1), N), 11, 17),       # diagonal: 1 -> 7 long
                          (pkg.matgen.poisson2d(12), 31, 31 * 5 - 2 + 9)]:
        vec = pkg.matgen.rhs(t.N)
        ctx, A, om = make(pkg, O, mode, t)
        if mode == "crc32c" and t.N == N:
            ctx.close()
            continue   # 1-element rows carry per-element SECDED bytes: lengthening one is a CRC/ECC fault of its own
        ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, row, value), "poke")
        om.rows[row] = value
        y, fatal, events, _ = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (mode, row, value)
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), (mode, row, value)
        ctx.close()


def test_multiple_faults_are_reported_in_serial_order(pkg, O):
    """Several corrupted elements: events come back sorted like the serial CPU loop visits them,
    and everything after the first fatal one is dropped (the reference has exit(1)ed)."""
    t = pkg.matgen.poisson2d(20)
    vec = pkg.matgen.rhs(t.N)
    ctx, A, om = make(pkg, O, "secded", t)
    for index, bits in [(1500, [3]), (20, [70]), (900, [1, 2]), (1700, [5]), (300, [88])]:
        for bit in bits:
            ctx.flip_bit(A, index, bit); om.flip(index, bit)
    y, fatal, events, _ = gpu_spmv(ctx, A, vec)
    y0, fatal0, ev0 = om.spmv(vec)
    assert fatal and fatal0 and events == ev0
    assert [e[1] for e in events] == [20, 300, 900]
    ctx.close()


# ---------------------------------------------------------------- row partitions (multi-GPU building block)
@pytest.mark.parametrize("mode", ["none", "secded", "crc32c", "constraints"])
def test_row_partitions_reproduce_the_whole(pkg, O, mode):
    """A matrix split into row blocks (global columns, global indices in events) gives the same
    result slices, the same encoded words and the same events as the unsplit matrix."""
    t = pkg.matgen.poisson3d(14)
    vec = pkg.matgen.rhs(t.N)
    om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    y0, _, _ = om.spmv(vec)
    rowptr = t.row_pointers().astype(np.int64)
    cuts = [0, 1000, 1001, 2000, t.N]
    ctx = pkg.CUDAContext(mode)
    v = ctx.create_vector(t.N)
    ctx.upload(v, vec)
    total_dot = 0.0
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        e0, e1 = int(rowptr[r0]), int(rowptr[r1])
        A = ctx.create_matrix_part(t.cols[e0:e1], t.rows[e0:e1], t.vals[e0:e1], r1 - r0, t.N, r0, e0)
        cols, rp, _ = ctx.read_matrix(A)
        assert np.array_equal(cols, om.cols[e0:e1]) and np.array_equal(rp, om.rows[r0:r1 + 1] - e0)
        w = ctx.create_vector(r1 - r0)
        ctx.spmv(A, v, w)
        part = ctx.view_vector(v, r0, r1 - r0)
        total_dot += ctx.dot(part, w)
        assert ctx.poll_events() == ([], False)
        assert np.array_equal(ctx.download(w), y0[r0:r1])
        if mode == "secded" and r1 - r0 > 1:
            gi = e0 + 5
            ctx.flip_bit(A, gi, 17)
            ctx.spmv(A, v, w)
            assert ctx.poll_events() == ([(2, gi, 17)], False)      # global element index
    exact = O.lib().orc_dot(vec, y0, t.N)
    assert abs(total_dot - exact) <= RTOL * abs(exact)
    ctx.close()


# ---------------------------------------------------------------- exhaustive flip sweeps
def test_secded_detects_all_4560_double_flips(pkg):
    """Every pair of bits of one element: parity even, syndrome non-zero -> 'double-bit error detected', nothing
    corrected, nothing written back."""
    t = pkg.matgen.poisson2d(3)
    ctx = pkg.CUDAContext("secded")
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    clean_cols, _, clean_vals = ctx.read_matrix(A)
    v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
    ctx.upload(v, np.ones(t.N))
    index = 7
    n = 0
    for a in range(96):
        for b in range(a + 1, 96):
            ctx.flip_bit(A, index, a); ctx.flip_bit(A, index, b)
            ctx.spmv(A, v, w)
            events, fatal = ctx.poll_events()
            assert fatal and events == [(4, index, 0)], (a, b, events)
            ctx.flip_bit(A, index, a); ctx.flip_bit(A, index, b)        # undo
            n += 1
    assert n == 4560
    cols, _, vals = ctx.read_matrix(A)
    assert np.array_equal(cols, clean_cols) and np.array_equal(vals, clean_vals)
    ctx.close()


@pytest.mark.parametrize("mode", ["sec7", "sec8"])
def test_all_double_flips_match_oracle(pkg, O, mode):
    """SEC7/SEC8 mis-correct double flips (768 of the 4560 pairs even point at 'bit >= 96', where the reference
    writes out of bounds): the CUDA target must report exactly what the oracle reports and leave the same matrix.
    One element per pair in a diagonal matrix, one spmv per batch (the device log holds 2048 records)."""
    pairs = [(a, b) for a in range(96) for b in range(a + 1, 96)]
    rng = np.random.default_rng(12)
    for lo in range(0, len(pairs), 1500):
        batch = pairs[lo:lo + 1500]
        n = len(batch)
        t = pkg.matgen.Triplets(np.arange(n), np.arange(n), rng.standard_normal(n) * 3.0, n)
        ctx, A, om = make(pkg, O, mode, t)
        for i, (a, b) in enumerate(batch):
            for bit in (a, b):
                ctx.flip_bit(A, i, bit); om.flip(i, bit)
        v, w = ctx.create_vector(n), ctx.create_vector(n)
        ctx.upload(v, np.ones(n))
        ctx.spmv(A, v, w)
        events, fatal = ctx.poll_events(cap=2048)
        y = ctx.download(w)
        y0, fatal0, ev0 = om.spmv(np.ones(n))
        assert fatal == fatal0 and len(events) == len(ev0)
        assert events == ev0
        cols, _, vals = ctx.read_matrix(A)
        assert np.array_equal(cols, om.cols) and np.array_equal(vals.view(np.uint64), om.values.view(np.uint64))
        assert np.array_equal(y, y0, equal_nan=True)      # (a flipped exponent can make a NaN: payloads may differ)
        ctx.close()


def test_fault_log_overflow_is_reported_not_dropped(pkg):
    """More faults in one spmv than the device log holds (16384 records): every element is still repaired, the records
    that fit come back in order, and the last one says how many were dropped -- fatal, because the printed lines can no
    longer equal the reference's (which has no such limit).  Polling hands the events out in pages."""
    n, faults = 20000, 17000
    rng = np.random.default_rng(5)
    t = pkg.matgen.Triplets(np.arange(n), np.arange(n), rng.standard_normal(n), n)
    ctx = pkg.CUDAContext("sec7")
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    clean_cols, _, clean_vals = ctx.read_matrix(A)
    bits = rng.integers(0, 64, faults)
    for i in range(faults):
        ctx.flip_bit(A, i, int(bits[i]))
    v, w = ctx.create_vector(n), ctx.create_vector(n)
    ctx.upload(v, np.ones(n))
    ctx.spmv(A, v, w)
    events, fatal = ctx.poll_events(cap=1000)
    assert fatal and len(events) == 16384 + 1
    assert events[-1] == (12, faults - 16384, 0)
    assert ctx.format_event(events[-1]) == f"fault log overflow: {faults - 16384} events dropped"
    idx = [e[1] for e in events[:-1]]
    assert idx == sorted(set(idx)) and all(e[0] == 2 and e[2] == bits[e[1]] for e in events[:-1])
    cols, _, vals = ctx.read_matrix(A)
    assert np.array_equal(cols, clean_cols) and np.array_equal(vals.view(np.uint64), clean_vals.view(np.uint64))
    ctx.spmv(A, v, w)
    assert ctx.poll_events() == ([], False)
    ctx.close()


def test_crc32c_row_known_answer_on_device(pkg):
    """RFC 3720 B.4's 48-byte PDU as a one-row crc32c matrix (tests/golden/crc32c_rfc3720.json): the device-side encode
    writes the PUBLISHED CRC-32C into the top bytes of the four column words, and the spmv check (different tables:
    5-bit chunks of the three linear maps) accepts the row; one flipped bit anywhere in the 48 bytes is caught."""
    from conftest import load_golden
    K = load_golden("crc32c_rfc3720.json")["row_known_answer"]
    vals = np.array([np.frombuffer(bytes.fromhex(h), dtype="<f8")[0] for h in K["values_hex"]])
    cols = np.array(K["columns"], dtype=np.uint32)
    N = int(cols.max()) + 1
    ctx = pkg.CUDAContext("crc32c")
    A = ctx.create_matrix(cols, np.zeros(4, np.uint32), vals, N)
    got, _, _ = ctx.read_matrix(A)
    assert [int(c >> 24) for c in got] == K["crc_bytes_in_column_top_bytes"]
    v, w = ctx.create_vector(N), ctx.create_vector(N)
    ctx.upload(v, np.ones(N))
    ctx.spmv(A, v, w)
    assert ctx.poll_events() == ([], False)
    for index, bit in [(0, 0), (1, 37), (2, 63), (3, 64), (2, 87)]:
        ctx.flip_bit(A, index, bit)
        ctx.spmv(A, v, w)
        events, fatal = ctx.poll_events()
        assert fatal and events == [(9, 0, 0)], (index, bit, events)
        ctx.flip_bit(A, index, bit)
    ctx.close()


@pytest.mark.parametrize("mode", ["none", "secded"])
def test_compiled_host_loop_equals_the_four_calls(pkg, O, mode):
    """abft_cuda_cg_host_loop (what bench.py times as the plugin path) is the loop of cg.cpp:94-118 over the same four
    entry points as CUDAContext.cg(): same iteration count, same residual history and solution, bit for bit -- also
    with a fault repaired on the way (the loop stops before using that spmv's result, the fault is reported, and the
    iteration is redone)."""
    t = pkg.matgen.poisson2d(48)
    b = pkg.matgen.rhs(t.N)
    runs = []
    for compiled in (False, True):
        ctx = pkg.CUDAContext(mode)
        A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
        if mode == "secded":
            ctx.flip_bit(A, 777, 41)
        vb, x, r, p, w = (ctx.create_vector(t.N) for _ in range(5))
        ctx.upload(vb, b)
        if compiled:
            ctx.copy_vector(r, vb)
            ctx.copy_vector(p, r)
            hist = []
            it, rr = ctx.cg_host_loop(A, x, r, p, w, 1000, 1e-10, ctx.dot(r, r), hist)
            hist = np.asarray(hist)
        else:
            it, hist = ctx.cg(A, vb, x, r, p, w, 1000, 1e-10)
        runs.append((it, hist, ctx.download(x), list(ctx.messages)))
        ctx.close()
    assert runs[0][0] == runs[1][0] and np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    assert runs[0][3] == runs[1][3] == (["[ECC] corrected bit 41 at index 777"] if mode == "secded" else [])
    ref = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    if mode == "secded":
        ref.flip(777, 41)
    want = ref.cg(b, 1000, 1e-10)
    assert runs[1][0] == want["iterations"] and np.max(np.abs(runs[1][1] - want["rr_hist"]) / want["rr_hist"]) <= RTOL
