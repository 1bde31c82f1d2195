"""The CPU oracle against the live, unmodified reference (oracle/_ref, built from the
reference's own sources by oracle/Makefile; shipped prebuilt to the GPU box).
Skipped only where neither the reference tree nor a prebuilt oracle/_ref exists."""
import numpy as np
import pytest

REF_MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]


@pytest.fixture(scope="module")
def ref(O):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (no reference tree here)")
    return O


def test_ecc_random_vectors(ref):
    L, R = ref.lib(), ref.ref()
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2**63, 5000, dtype=np.int64).view(np.float64)
    cols = rng.integers(0, 2**32, 5000, dtype=np.uint64)
    for v, c in zip(bits, cols):
        v, c = float(v), int(c)
        assert L.orc_ecc_compute_col8(v, c) == R.ref_ecc_compute_col8(v, c)
        assert L.orc_ecc_overall_parity(v, c) == R.ref_ecc_overall_parity(v, c)


@pytest.mark.parametrize("mode", REF_MODES)
@pytest.mark.parametrize("shape", ["p2d_24_b1", "p2d_12_b3", "p3d_7", "powerlaw"])
def test_matrix_and_cg_bit_identical(ref, pkg, mode, shape):
    mg = pkg.matgen
    t = {"p2d_24_b1": lambda: mg.poisson2d(24), "p2d_12_b3": lambda: mg.poisson2d(12, 3),
         "p3d_7": lambda: mg.poisson3d(7), "powerlaw": lambda: mg.powerlaw(300, 3000, seed=4)}[shape]()
    b = mg.rhs(t.N)
    om = ref.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    rc = ref.RefContext(mode)
    rc.create_matrix(t.cols, t.rows, t.vals, t.N)
    c, r, v = rc.matrix_arrays()
    assert np.array_equal(c, om.cols) and np.array_equal(r, om.rows) and np.array_equal(v, om.values)
    a, g = om.cg(b, 400, 1e-12), rc.cg(b, 400, 1e-12)
    assert a["iterations"] == g["iterations"] and not g["text"]
    assert np.array_equal(a["rr_hist"], g["rr_hist"]) and np.array_equal(a["x"], g["x"])


@pytest.mark.parametrize("mode", ["sec7", "sec8", "secded"])
def test_all_96_single_flips(ref, pkg, mode):
    """Every bit of one element: same message, same corrected matrix, same spmv result."""
    t = pkg.matgen.poisson2d(6)
    b = pkg.matgen.rhs(t.N)
    index = 77
    for bit in range(96):
        om = ref.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        om.flip(index, bit)
        rc = ref.RefContext(mode)
        rc.create_matrix(t.cols, t.rows, t.vals, t.N)
        c, r, v = rc.matrix_arrays()
        if bit < 64:
            v.view(np.uint64)[index] ^= np.uint64(1) << np.uint64(bit)
        else:
            c[index] ^= np.uint32(1 << (bit - 64))
        vb, vw = rc.vector(b), rc.vector()
        text = rc.spmv(vb, vw)
        y, fatal, ev = om.spmv(b)
        assert not fatal
        assert [ref.format_event(*e) for e in ev] == [ln for ln in text.splitlines() if ln]
        assert np.array_equal(y, rc.view(vw))
        assert np.array_equal(c, om.cols) and np.array_equal(v, om.values)


def test_vector_kernels_bit_identical(ref, pkg):
    rng = np.random.default_rng(5)
    N = 1001
    L = ref.lib()
    rc = ref.RefContext("none")
    rc.N = N
    x, r, p, w = (rng.standard_normal(N) for _ in range(4))
    vx, vr, vp, vw = rc.vector(x), rc.vector(r), rc.vector(p), rc.vector(w)
    assert rc.dot(vp, vw) == L.orc_dot(p, w, N)
    x2, r2, p2 = x.copy(), r.copy(), p.copy()
    assert rc.calc_xr(vx, vr, vp, vw, 0.37) == L.orc_calc_xr(x2, r2, p, w, 0.37, N)
    assert np.array_equal(rc.view(vx), x2) and np.array_equal(rc.view(vr), r2)
    rc.calc_p(vp, vr, -1.25)
    L.orc_calc_p(p2, r2, -1.25, N)
    assert np.array_equal(rc.view(vp), p2)


def test_differential_fuzz_against_the_reference(ref):
    """Random matrices x modes x multi-bit / multi-element faults (and row-pointer faults under `constraints`), one spmv
    each, the reference's side in a forked child because it exit(1)s on fatal faults: same lines, same exit status,
    same repaired matrix, same result[] wherever the reference's behaviour is defined (tools/fuzz_oracle_vs_ref.py).
    Run as its own process: it pins OpenMP to one thread before libgomp loads (serial message order; and a child
    forked from a process whose OpenMP pool already exists would hang in the reference's parallel loop)."""
    import json
    import os
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_oracle_vs_ref.py")
    out = subprocess.run([sys.executable, tool, "--trials", "400", "--seed", "20261018"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    stats = json.loads(out.stdout.strip().splitlines()[-1])
    assert stats["trials"] == 400 and stats["fatal"] > 50 and stats["corrected"] > 50 and stats["result_compared"] > 100
