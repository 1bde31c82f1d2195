"""The differential fuzz of tools/fuzz_oracle_vs_ref.py, GPU side: the same random cases (random matrices incl. one with
arbitrary double values, 1..3 flips in each of 1..2 elements, row-pointer faults under `constraints`) through the C ABI
on the device and through the CPU oracle -- which the CPU suite runs against the live reference on this very
distribution (tests/test_oracle_ref.py).  Same events in the same order, same fatal verdict, same repaired matrix,
same result[] (bit for bit where the kernels keep the sequential order: rows of up to 32 elements; the oracle and the
kernels clamp out-of-range columns alike, so nothing is skipped here).
"""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz_module():
    spec = importlib.util.spec_from_file_location("fuzz_oracle_vs_ref", os.path.join(ROOT, "tools", "fuzz_oracle_vs_ref.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("mode", ["none", "constraints", "sed", "sec7", "sec8", "secded"])
def test_random_faults_gpu_equals_oracle(pkg, O, mode):
    fz = _fuzz_module()
    rng = np.random.default_rng(77)
    pool = fz.matrices(pkg.matgen, rng)
    ctx = pkg.CUDAContext(mode)
    fatal_seen = repaired_seen = 0
    for trial in range(250):
        t = pool[int(rng.integers(0, len(pool)))]
        vec = rng.standard_normal(t.N)
        flips, pokes = fz.draw_faults(rng, mode, t)
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        fz.corrupt_arrays(om.cols, om.rows, om.values, flips, pokes)
        y0, fatal0, ev0 = om.spmv(vec)
        A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
        for index, bit in flips:
            ctx.flip_bit(A, index, bit)
        for row, value in pokes:
            ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, row, value), "poke_rowptr")
        v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
        ctx.upload(v, vec)
        ctx.spmv(A, v, w)
        events, fatal = ctx.poll_events()
        what = (mode, trial, flips, pokes)
        assert (fatal, events) == (fatal0, ev0), what
        if not fatal:
            # (after a fatal fault the reference has exited; rows behind it were never visited there, while the
            # device has walked -- and possibly repaired -- all of them: nothing to compare)
            cols, rows, vals = ctx.read_matrix(A)
            assert np.array_equal(cols, om.cols) and np.array_equal(vals.view(np.uint64), om.values.view(np.uint64)), what
            assert np.array_equal(rows, om.rows), what
            assert fz.same_result(ctx.download(w), y0, t, vec, om.cols, om.values, mode), what
        fatal_seen += fatal
        repaired_seen += bool(events) and not fatal
        for x in (v, w):
            ctx.destroy_vector(x)
        ctx.destroy_matrix(A)
    ctx.close()
    if mode in ("sed", "secded", "constraints"):
        assert fatal_seen > 10
    if mode in ("sec7", "sec8", "secded"):
        assert repaired_seen > 10


@pytest.mark.parametrize("mode", ["none", "constraints", "sed", "sec7", "sec8", "secded"])
def test_random_faults_gpu_equals_coo_oracle(pkg, O, mode):
    """The COO target (abft_cuda_coo_*, default layout and kernel: grouped elements streamed through shared memory) on
    the distribution tools/fuzz_oracle_vs_ref.py --format coo holds the COO oracle to the live reference on: 1..3 flips
    (any of the 128 bits) in each of 1..2 elements.  Same events in the same order, same fatal verdict, same repaired
    elements, same result[] bit for bit (the kernel keeps the reference's order of additions per output entry; elements
    whose index leaves [0, N) are dropped by kernel and oracle alike)."""
    import coo_oracle_py as CO
    fz = _fuzz_module()
    mg = pkg.matgen
    rng = np.random.default_rng(78)
    # the small pool of the CSR fuzz, plus matrices of several tiles (256 output entries each) and one with columns
    # too dense to stage
    pool = fz.matrices(mg, rng) + [mg.poisson2d(24), mg.poisson3d(8), mg.powerlaw(700, 4000, seed=5)]
    trials = int(os.environ.get("ABFT_FUZZ_TRIALS", "120"))
    ctx = pkg.CUDAContext(mode)
    fatal_seen = repaired_seen = 0
    for trial in range(trials):
        t = pool[int(rng.integers(0, len(pool)))]
        vec = rng.standard_normal(t.N)
        flips = []
        for _ in range(int(rng.integers(1, 3))):
            index = int(rng.integers(0, t.nnz))
            flips += [(index, b) for b in fz.draw_bits(rng, 128, fz.COO_TAME)]
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        for index, bit in flips:
            om.flip(index, bit)
            ctx.coo_flip_bit(A, index, bit)
        y0, fatal0, ev0 = om.spmv(vec)
        v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
        ctx.upload(v, vec)
        ctx.coo_spmv(A, v, w)
        events, fatal = ctx.poll_events()
        what = (mode, trial, t.N, t.nnz, flips)
        assert (fatal, events) == (fatal0, ev0), what
        if not fatal:
            assert np.array_equal(ctx.read_coo_matrix(A).view(np.uint32), om.elements.view(np.uint32)), what
            assert np.array_equal(ctx.download(w).view(np.uint64), y0.view(np.uint64)), what
            # the dot the spmv leaves behind equals the dot of what was stored (strays included)
            got, y = ctx.dot(v, w), ctx.download(w)
            with np.errstate(all="ignore"):
                exact, scale = float(np.dot(vec, y)), float(np.dot(np.abs(vec), np.abs(y)))
            if np.isfinite(scale) and scale < 1e290:      # (one matrix of the pool spans the whole exponent range)
                assert abs(got - exact) <= 1e-12 * scale + 1e-300, what
        fatal_seen += fatal
        repaired_seen += bool(events) and not fatal
        for x in (v, w):
            ctx.destroy_vector(x)
        ctx.destroy_coo_matrix(A)
    ctx.close()
    if trials >= 100:
        if mode in ("sed", "secded", "constraints"):
            assert fatal_seen > 5
        if mode in ("sec7", "sec8", "secded"):
            assert repaired_seen > 5
