"""COO format on the GPU (abft_cuda_coo_*) against the pinned COO oracle (oracle/coo_oracle.c).

Every test runs in both element layouts (coo.cuh): the default grouped one -- with the streamed kernel of
coo_stream.cuh and with coo_spmv_kernel (ABFT_CUDA_COO_KERNEL=direct) -- and the caller-order one
(ABFT_CUDA_COO_LAYOUT=caller).  The same cases also run on the CPU against the kernels' own source through a shim of
the CUDA built-ins (tests/test_coo_kernel_logic.py).

Bars: bit-exact encoded elements, events (kind, index, bit), repaired elements and result[] (the kernel keeps the
reference's accumulation order per output entry and uses no FMA); CG iteration counts equal, rr within 1e-10
relative (reductions are reordered).
"""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10
MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]


@pytest.fixture(params=["grouped", "caller", "grouped_direct"], autouse=True)
def layout(request, monkeypatch):
    """grouped: the default (elements stored in group order, streamed through shared memory by coo_stream.cuh's kernel;
    tiles too large to stage -- the power-law shape -- read global memory); caller: the caller's order, read through
    `perm`; grouped_direct: the grouped layout read straight from global memory (coo_spmv_kernel)."""
    monkeypatch.delenv("ABFT_CUDA_COO_LAYOUT", raising=False)
    monkeypatch.delenv("ABFT_CUDA_COO_KERNEL", raising=False)
    if request.param == "caller":
        monkeypatch.setenv("ABFT_CUDA_COO_LAYOUT", "caller")
    elif request.param == "grouped_direct":
        monkeypatch.setenv("ABFT_CUDA_COO_KERNEL", "direct")
    return request.param


@pytest.fixture(scope="module")
def CO(O):
    import coo_oracle_py
    coo_oracle_py.lib()
    return coo_oracle_py


def shapes(mg):
    return {"p2d_24": lambda: mg.poisson2d(24), "p2d_9_b3": lambda: mg.poisson2d(9, 3), "p3d_8": lambda: mg.poisson3d(8),
            "powerlaw": lambda: mg.powerlaw(2000, 40000, seed=3), "one": lambda: mg.Triplets([0], [0], [2.5], 1)}


def gpu_spmv(ctx, A, vec):
    v, w = ctx.create_vector(len(vec)), ctx.create_vector(A.N)
    ctx.upload(v, vec)
    ctx.coo_spmv(A, v, w)
    events, fatal = ctx.poll_events()
    return ctx.download(w), fatal, events


@pytest.mark.parametrize("mode", MODES)
def test_create_and_spmv_bit_exact(pkg, CO, mode):
    rng = np.random.default_rng(2)
    for name, build in shapes(pkg.matgen).items():
        t = build()
        ctx = pkg.CUDAContext(mode)
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        assert np.array_equal(ctx.read_coo_matrix(A).view(np.uint32), om.elements.view(np.uint32)), name
        vec = rng.standard_normal(t.N)
        y, fatal, events = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert not fatal and events == [] == ev0, name
        assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), name
        ctx.close()


@pytest.mark.parametrize("mode", ["sed", "sec7", "sec8", "secded"])
def test_all_128_single_flips(pkg, CO, mode):
    t = pkg.matgen.poisson2d(12, 2)
    vec = pkg.matgen.rhs(t.N)
    for bit in range(128):
        ctx = pkg.CUDAContext(mode)
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        ctx.coo_flip_bit(A, 77, bit); om.flip(77, bit)
        y, fatal, events = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), bit
        assert [ctx.coo_format_event(e) for e in events] == [CO.format_event(*e) for e in ev0]
        assert np.array_equal(ctx.read_coo_matrix(A).view(np.uint32), om.elements.view(np.uint32)), bit
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), bit
        ctx.close()


@pytest.mark.parametrize("mode", ["none", "sed", "sec7", "sec8"])
def test_silent_index_faults_are_rerouted_like_the_reference(pkg, CO, mode):
    """Index faults no check repairs -- any flip in mode none, double flips under sed / sec8 (parity stays even),
    double flips under sec7 (mis-corrected) -- move an element's product to another output entry, or drop it."""
    t = pkg.matgen.poisson2d(16)
    vec = pkg.matgen.rhs(t.N)
    rng = np.random.default_rng(12)
    for trial in range(40):
        ctx = pkg.CUDAContext(mode)
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        index = int(rng.integers(0, t.nnz))
        bits = [int(rng.integers(0, 9))] if mode == "none" else [int(b) for b in rng.choice(64, 2, replace=False)]
        for b in bits:
            ctx.coo_flip_bit(A, index, b); om.flip(index, b)
        y, fatal, events = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (index, bits)
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), (index, bits)
            assert np.array_equal(ctx.read_coo_matrix(A).view(np.uint32), om.elements.view(np.uint32))
        ctx.close()


def test_constraints_violations(pkg, CO):
    t = pkg.matgen.poisson2d(16)
    vec = pkg.matgen.rhs(t.N)
    # (element, bit): row too large, column too large, row order, column order
    for index, bit in [(300, 45), (300, 23), (301, 38), (302, 1), (0, 32), (t.nnz - 1, 40), (t.nnz - 1, 3)]:
        ctx = pkg.CUDAContext("constraints")
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        om = CO.CooOracleMatrix("constraints", t.cols, t.rows, t.vals, t.N)
        ctx.coo_flip_bit(A, index, bit); om.flip(index, bit)
        y, fatal, events = gpu_spmv(ctx, A, vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (index, bit)
        assert [ctx.coo_format_event(e) for e in events] == [CO.format_event(*e) for e in ev0]
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64))
        ctx.close()


def test_dot_after_coo_spmv(pkg, CO, layout):
    """abft_cuda_dot(vec, result) right after abft_cuda_coo_spmv(A, vec, result) (cg.cpp:97-100): served by the streamed
    kernel's own reduction (no launch) -- also when the stray fix-up kernel rewrites entries of result[] afterwards (an
    index fault in mode none moves an element to another output entry) -- and equal to the dot of what was stored."""
    t = pkg.matgen.poisson2d(40)
    vec = np.random.default_rng(4).standard_normal(t.N)
    for flips in ([], [(77, 3)], [(77, 3), (901, 6), (902, 1)]):
        ctx = pkg.CUDAContext("none")
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        for index, bit in flips:
            ctx.coo_flip_bit(A, index, bit)
        v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
        ctx.upload(v, vec)
        ctx.coo_spmv(A, v, w)
        n0 = ctx.launch_count()
        got = ctx.dot(v, w)
        if layout == "grouped":
            assert ctx.launch_count() == n0, "the streamed kernel reduces dot(vec, result) itself"
            assert ctx.dot(w, v) == got and ctx.launch_count() == n0
        y = ctx.download(w)
        exact = float(np.dot(vec, y))
        scale = float(np.dot(np.abs(vec), np.abs(y)))
        assert abs(got - exact) <= 1e-13 * scale, flips
        ctx.copy_vector(w, w)                       # any write invalidates the cached value
        again = ctx.dot(v, w)
        assert ctx.launch_count() > n0 and abs(again - exact) <= 1e-13 * scale
        ctx.close()


@pytest.mark.parametrize("mode", MODES)
def test_cg_matches_oracle(pkg, CO, mode):
    t = pkg.matgen.poisson2d(48)
    b = pkg.matgen.rhs(t.N)
    ctx = pkg.CUDAContext(mode)
    A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
    vb, x, r, p, w = (ctx.create_vector(t.N) for _ in range(5))
    ctx.upload(vb, b)
    it, hist, events, fatal = ctx.coo_cg(A, vb, x, r, p, w, 1000, 1e-10)
    ref = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N).cg(b, 1000, 1e-10)
    assert not fatal and events == [] and it == ref["iterations"]
    assert float(np.max(np.abs(hist - ref["rr_hist"]) / ref["rr_hist"])) <= RTOL
    ctx.close()


def test_cli_fixture_replay(pkg, CO):
    """The reference cg-coo CLI cases of tests/golden/coo_golden.json, replayed through the C ABI."""
    from conftest import load_golden
    G = load_golden("coo_golden.json")["cli"]
    t = pkg.matgen.poisson2d(G["grid"], G["blocks"])
    b = pkg.matgen.rhs(t.N)
    for case in G["cases"]:
        ctx = pkg.CUDAContext(case["mode"])
        A = ctx.create_coo_matrix(t.cols, t.rows, t.vals, t.N)
        for ln in case["faults"]:
            mm = re.match(r"\*\*\* flipping bit (\d+) at index (\d+) \*\*\*", ln)
            if mm:
                ctx.coo_flip_bit(A, int(mm.group(2)), int(mm.group(1)))
        vb, x, r, p, w = (ctx.create_vector(t.N) for _ in range(5))
        ctx.upload(vb, b)
        it, hist, events, fatal = ctx.coo_cg(A, vb, x, r, p, w, 300, 1e-10)
        assert [ctx.coo_format_event(e) for e in events] == [ln for ln in case["faults"] if not ln.startswith("***")], case
        assert fatal == (case["exit"] != 0), case
        if case["ran"]:
            assert case["ran"] == f"ran for {it} iterations", case
        ctx.close()
