"""Mode `secdedv` (repository-defined, PARITY UNPINNED: the reference protects no vector, CSR/CPUContext.cpp:54-113):
secded matrix elements + SECDED row pointers + a parity bit in the lowest mantissa bit of every vector element.

CPU side (this file's non-gpu tests): the kernels' own source on the CPU emulator against oracle/vecprot_oracle.py --
stored values bit for bit (spmv result, x, r, p given the same scalars), every single flipped bit of a vector element
detected by the kernel that reads it next.  GPU side: the same through the C ABI, plus the CG trajectory.

Stated tolerance: against the secdedv oracle the residual history agrees to 1e-9 relative (the device reorders the
reductions, and every stored value is perturbed by at most one ulp on either side); against the UNPROTECTED reference
arithmetic the iteration count of the 256^2 case stays within +-2 and the solution within 1e-9 relative.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
SECDEDV = 7
VEC_SED = 13


@pytest.fixture(scope="module")
def V(O):
    import vecprot_oracle
    return vecprot_oracle


@pytest.fixture(scope="module")
def EMU(tmp_path_factory):
    from test_csr_kernel_logic import _build
    return _build(tmp_path_factory.mktemp("csr_emu_vp"), "vecprot", [])


def emu_matrix(EMU, pkg, t, protect_rows=True):
    from test_csr_kernel_logic import EmuMatrix, MODES
    MODES.append("secdedv") if "secdedv" not in MODES else None
    em = EmuMatrix(EMU, pkg, "secdedv", t)
    if protect_rows:
        EMU.emu_protect_rows.restype = C.c_int
        EMU.emu_protect_rows.argtypes = [C.c_void_p]
        assert EMU.emu_protect_rows(em.h) == 1
    return em


def vec_events(EMU, em):
    from test_csr_kernel_logic import Event
    buf = (Event * 64)()
    n = EMU.emu_vec_events(em.h, buf, 64)
    return sorted((buf[i].kind, buf[i].index, buf[i].bit) for i in range(n))


def test_encode_gives_even_parity_and_costs_one_ulp(V):
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.standard_normal(1000), [0.0, -0.0, 1.0, 1e-310, 1e308, np.inf]])
    e = V.encode(v)
    assert not V.odd(e).any()
    assert np.array_equal(V.encode(e).view(np.uint64), e.view(np.uint64))          # idempotent
    assert np.all((e.view(np.uint64) ^ v.view(np.uint64)) <= 1)                    # only bit 0 may differ
    for bit in (0, 1, 31, 32, 52, 63):                                             # any single flip shows
        f = (e.view(np.uint64) ^ np.uint64(1 << bit)).view(np.float64)
        assert V.odd(f).all()


def test_vector_kernels_store_encoded_values_and_catch_flips(EMU, pkg, O, V):
    """calc_xr, calc_p, dot on the emulator: stored x, r, p equal the oracle's encode(...) bit for bit; a flipped bit in
    any operand is reported by the kernel that reads it, naming the vector and the element."""
    t = pkg.matgen.poisson2d(12)
    em = emu_matrix(EMU, pkg, t)
    N = 1001
    rng = np.random.default_rng(2)
    x, r, p, w = (V.encode(rng.standard_normal(N)) for _ in range(4))
    x0, r0 = x.copy(), r.copy()
    rr = EMU.emu_calc_xr(em.h, x, r, p, w, 0.37, N, 3)
    xe, re_ = V.encode(x0 + 0.37 * p), V.encode(r0 - 0.37 * w)
    assert np.array_equal(x.view(np.uint64), xe.view(np.uint64)) and np.array_equal(r.view(np.uint64), re_.view(np.uint64))
    assert abs(rr - float(np.dot(re_, re_))) <= 1e-12 * rr and vec_events(EMU, em) == []
    p0 = p.copy()
    EMU.emu_calc_p(em.h, p, r, -1.25, N, 3)
    assert np.array_equal(p.view(np.uint64), V.encode(r + (-1.25) * p0).view(np.uint64)) and vec_events(EMU, em) == []
    assert abs(EMU.emu_dot(em.h, p, r, N, 3) - float(np.dot(p, r))) <= 1e-12 * abs(float(np.dot(p, r))) and vec_events(EMU, em) == []
    # one flipped bit per vector, each found by the next reader, with the vector's name (2 x, 3 r, 4 p, 5 w)
    for which, (arr, name) in enumerate([(x, 2), (r, 3), (p, 4), (w, 5)]):
        i, bit = 17 + 100 * which, [0, 13, 52, 63][which]
        arr.view(np.uint64)[i] ^= np.uint64(1 << bit)
        EMU.emu_calc_xr(em.h, x.copy(), r.copy(), p, w, 0.5, N, 2)
        assert vec_events(EMU, em) == [(VEC_SED, i, name)], (which, i, bit)
        arr.view(np.uint64)[i] ^= np.uint64(1 << bit)
    p.view(np.uint64)[999] ^= np.uint64(1 << 40)
    EMU.emu_calc_p(em.h, p.copy(), r, 0.5, N, 2)
    assert vec_events(EMU, em) == [(VEC_SED, 999, 4)]
    EMU.emu_dot(em.h, p, r, N, 2)
    assert vec_events(EMU, em) == [(VEC_SED, 999, 6)]
    em.close()


def test_spmv_checks_its_operand_and_encodes_its_result(EMU, pkg, O, V):
    """spmv in mode secdedv on the emulator: result = encode(secded row sums), the fused dot uses the stored values;
    a flipped bit in the operand is reported once per row that gathers it (column index, vector 1) and nothing else."""
    for t in (pkg.matgen.poisson3d(8), pkg.matgen.poisson2d(20)):
        em = emu_matrix(EMU, pkg, t)
        vp = V.VecProtCG(t.cols, t.rows, t.vals, t.N)
        vec = V.encode(np.random.default_rng(3).standard_normal(t.N))
        y, fatal, events, dot = em.spmv(vec)
        y0, _, _ = vp.spmv(vec)
        assert not fatal and events == [] and np.array_equal(y.view(np.uint64), y0.view(np.uint64))
        assert abs(dot - float(np.dot(vec, y0))) <= 1e-10 * np.abs(vec * y0).sum()
        bad = vec.copy()
        j = t.N // 2 + 3
        bad.view(np.uint64)[j] ^= np.uint64(1 << 22)
        _, fatal, events, _ = em.spmv(bad)
        assert fatal and events and all(e == (VEC_SED, j, 1) for e in events)
        em.close()


def test_cg_trajectory_secdedv_on_the_emulator(EMU, pkg, O, V):
    t = pkg.matgen.poisson2d(20)
    b = pkg.matgen.rhs(t.N)
    N = t.N
    em = emu_matrix(EMU, pkg, t)
    ref = V.VecProtCG(t.cols, t.rows, t.vals, t.N).cg(b, 400, 1e-10)
    plain = O.OracleMatrix("secded", t.cols, t.rows, t.vals, t.N).cg(b, 400, 1e-10)
    x = np.zeros(N)
    r = b.copy()
    EMU.emu_vec_encode(r, N)
    p = r.copy()
    rr = EMU.emu_dot(em.h, r, r, N, 3)
    hist = []
    for _ in range(400):
        if rr <= 1e-10:
            break
        w, fatal, events, pw = em.spmv(p)
        assert not fatal and events == []
        rr_new = EMU.emu_calc_xr(em.h, x, r, p, w, rr / pw, N, 3)
        EMU.emu_calc_p(em.h, p, r, rr_new / rr, N, 3)
        rr = rr_new
        hist.append(rr)
    assert vec_events(EMU, em) == []
    assert len(hist) == ref["iterations"] and np.max(np.abs(np.array(hist) - ref["rr_hist"]) / ref["rr_hist"]) <= 1e-9
    assert abs(len(hist) - plain["iterations"]) <= 2
    assert np.max(np.abs(x - plain["x"])) <= 1e-9 * np.max(np.abs(plain["x"]))
    em.close()


# ------------------------------------------------------------------------------------------------ on the device
@pytest.mark.gpu
def test_gpu_secdedv_cg_and_faults(pkg, O, V):
    t = pkg.matgen.poisson2d(64)
    b = pkg.matgen.rhs(t.N)
    ctx = pkg.CUDAContext("secdedv")
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    cols, rowptr, _ = ctx.read_matrix(A)
    om = O.OracleMatrix("secded", t.cols, t.rows, t.vals, t.N)
    assert np.array_equal(cols, om.cols)                       # the element code is mode secded's
    assert np.any(rowptr >> 28) and np.array_equal(rowptr & 0x0FFFFFFF, om.rows)    # row pointers carry check bits
    vb, x, r, p, w = (ctx.create_vector(t.N) for _ in range(5))
    ctx.upload(vb, b)
    assert np.array_equal(ctx.download(vb).view(np.uint64), V.encode(b).view(np.uint64))     # encoded on arrival
    it, hist = ctx.cg(A, vb, x, r, p, w, 1000, 1e-10)
    ref = V.VecProtCG(t.cols, t.rows, t.vals, t.N).cg(b, 1000, 1e-10)
    plain = om.cg(b, 1000, 1e-10)
    assert ctx.messages == []
    assert it == ref["iterations"] and np.max(np.abs(hist - ref["rr_hist"]) / ref["rr_hist"]) <= 1e-9
    assert abs(it - plain["iterations"]) <= 2
    xs = ctx.download(x)
    assert not V.odd(xs).any() and np.max(np.abs(xs - plain["x"])) <= 1e-9 * np.max(np.abs(plain["x"]))
    # the compiled host loop and the device-resident loop take the same path
    ctx.upload(x, np.zeros(t.N))
    ctx.copy_vector(r, vb); ctx.copy_vector(p, r)
    it2, _ = ctx.cg_host_loop(A, x, r, p, w, 1000, 1e-10, ctx.dot(r, r))
    assert it2 == it and np.array_equal(ctx.download(x).view(np.uint64), xs.view(np.uint64))
    ctx.upload(x, np.zeros(t.N))
    it3, _ = ctx.cg_solve_device(A, vb, x, r, p, w, 1000, 1e-10, 16)
    assert it3 == it and np.array_equal(ctx.download(x).view(np.uint64), xs.view(np.uint64))
    # a flipped bit in each vector is found by the next kernel that reads it
    ctx.copy_vector(r, vb); ctx.copy_vector(p, r)
    for vec, name, reader in [(p, "spmv operand", lambda: ctx.spmv(A, p, w)), (r, "r", lambda: ctx.calc_p(p, r, 0.5)),
                              (x, "x", lambda: ctx.calc_xr(x, r, p, w, 0.1)), (w, "mapped", lambda: ctx.map_vector(w))]:
        idx, bit = 1234, 17
        ctx.flip_vector_bit(vec, idx, bit)
        reader()
        events, fatal = ctx.poll_events()
        assert fatal and events[0][0] == VEC_SED and events[0][1] == idx, (name, events)
        assert ctx.format_event(events[0]) == f"[ECC] error detected in vector ({name}) at index {idx}"
        ctx.flip_vector_bit(vec, idx, bit)
    # matrix and row-pointer faults: as in secded / with ABFT_CUDA_PROTECT_ROWS
    ctx.flip_bit(A, 777, 41)
    ctx.spmv(A, p, w)
    assert ctx.poll_events() == ([(2, 777, 41)], False)
    ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, 100, int(rowptr[100]) ^ (1 << 5)), "poke")
    ctx.spmv(A, p, w)
    events, fatal = ctx.poll_events()
    assert not fatal and events == [(10, 50, 5)]
    ctx.close()


@pytest.mark.gpu
def test_gpu_secdedv_through_the_plugin(pkg, O, tmp_path):
    """`cg-csr -t cuda -m secdedv` with the unmodified driver: runs, converges like the unprotected targets (+-2)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "cg-csr-cuda")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cg-csr-cuda not built")
    path = str(tmp_path / "p64.mtx")
    pkg.matgen.write_mtx(path, pkg.matgen.poisson2d(64))
    code, text = O.run_cg_csr(exe, ["-f", path, "-b", "1", "-c", "1e-10", "-t", "cuda", "-m", "secdedv"])
    code0, text0 = O.run_cg_csr(exe, ["-f", path, "-b", "1", "-c", "1e-10", "-t", "cpu", "-m", "secded"])
    assert code == 0 == code0, text[-1500:]
    ran = [int(ln.split()[2]) for ln in (text, text0) for ln in ln.splitlines() if ln.startswith("ran for")]
    assert len(ran) == 2 and abs(ran[0] - ran[1]) <= 2
    assert "implementation        = cuda-secdedv" in text
