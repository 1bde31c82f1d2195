"""The COO CPU oracle (oracle/coo_oracle.c) against fixtures generated from the UNMODIFIED reference COO target
(tests/golden/coo_golden.json, made by tests/golden/make_golden_coo.py) -- runs anywhere -- and, where the
reference build exists (oracle/_ref/libref_coo.so), against the live reference on fresh random inputs.

The COO backend is the next row of the scope table (DESIGN.md 0, row f); this is the parity gate the CUDA
side will be held to.  Bars: bit-exact everywhere (masks, encoded words, events, repaired elements, result[],
CG trajectories: the oracle uses the reference's operation order).
"""
import ctypes
import re
import struct

import numpy as np
import pytest

from conftest import load_golden, unhex, unhex1


@pytest.fixture(scope="module")
def CO(O):
    import coo_oracle_py
    coo_oracle_py.lib()
    return coo_oracle_py


@pytest.fixture(scope="module")
def G():
    return load_golden("coo_golden.json")


# ---------------------------------------------------------------- ECC arithmetic
def test_masks_equal_the_reference_constants(CO, G):
    L = CO.lib()
    assert [[L.coo_hamming_mask(p, w) for w in range(4)] for p in range(1, 8)] == G["ecc"]["masks"]
    assert L.coo_hamming_mask(1, 0) == 0x80AAAD5B and L.coo_hamming_mask(7, 3) == 0xFFFFFFFF   # COO/ecc.h:18,51


def test_encode_known_answers(CO, G):
    L = CO.lib()
    for k in G["ecc"]["kats"]:
        v, c, r = unhex1(k["value"]), k["col"], k["row"]
        assert L.coo_ecc_compute_col8(c, r, v) == k["syndrome"]
        assert L.coo_ecc_overall_parity(c, r, v) == k["overall"]
        assert L.coo_encode_column(2, c, r, v) == k["sed"]
        assert L.coo_encode_column(3, c, r, v) == k["sec7"]
        assert L.coo_encode_column(4, c, r, v) == k["sec8"] == L.coo_encode_column(5, c, r, v)
        # an encoded element is a code word
        assert L.coo_ecc_compute_col8(k["sec8"], r, v) == 0 and L.coo_ecc_overall_parity(k["sec8"], r, v) == 0


def test_syndrome_decode_table(CO, G):
    L = CO.lib()
    for d in G["ecc"]["decode"]:
        assert L.coo_ecc_get_flipped_bit_col8(d["syndrome"]) == d["bit"]


def test_all_128_single_flips_of_one_element(CO, G):
    L = CO.lib()
    s = G["ecc"]["single_flips"]
    v = unhex1(s["value"])
    for f in s["flips"]:
        w = [s["encoded"], s["row"]] + list(struct.unpack("<II", struct.pack("<d", v)))
        w[f["bit"] // 32] ^= 1 << (f["bit"] % 32)
        fv = struct.unpack("<d", struct.pack("<II", w[2], w[3]))[0]
        syn = L.coo_ecc_compute_col8(w[0], w[1], fv)
        assert (syn, L.coo_ecc_overall_parity(w[0], w[1], fv)) == (f["syndrome"], f["overall"])
        # every single flip trips the overall parity; the syndrome points back at the flipped bit, except for
        # bit 24 (the overall parity bit itself, covered by no Hamming mask)
        assert f["overall"] == 1
        if f["bit"] == 24:
            assert syn == 0
        else:
            assert L.coo_ecc_get_flipped_bit_col8(syn) == f["bit"]


# ---------------------------------------------------------------- spmv / CG
def _matrix(pkg, G):
    return pkg.matgen.poisson2d(G["cg"]["grid"], G["cg"]["blocks"]), unhex(G["cg"]["b"])


@pytest.mark.parametrize("mode", ["none", "constraints", "sed", "sec7", "sec8", "secded"])
def test_clean_matrix_spmv_and_cg_bit_exact(CO, pkg, G, mode):
    t, b = _matrix(pkg, G)
    g = G["cg"]["modes"][mode]
    m = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    words = m.elements.view(np.uint32).reshape(-1, 4)
    assert int(words[:, 0].astype(np.uint64).sum()) == g["col_words_sum"]
    assert [int(z) for z in words[:16, 0]] == g["col_words_head"]
    assert np.array_equal(words[:, 1], t.rows.astype(np.uint32)) and np.array_equal(m.elements["value"], t.vals)
    y, fatal, ev = m.spmv(b)
    assert not fatal and not ev and np.array_equal(y, unhex(g["spmv_b"]))
    res = m.cg(b, 500, 1e-10)
    assert res["iterations"] == g["iterations"]
    assert np.array_equal(res["rr_hist"], unhex(g["rr_hist"])) and np.array_equal(res["x"], unhex(g["x"]))


def test_every_single_flip_of_an_element_is_corrected_like_the_reference(CO, pkg, G):
    t, b = _matrix(pkg, G)
    index = G["cg"]["single_flips"]["index"]
    for f in G["cg"]["single_flips"]["cases"]:
        m = CO.CooOracleMatrix(f["mode"], t.cols, t.rows, t.vals, t.N)
        clean = m.elements.view(np.uint32).copy()
        m.flip(index, f["bit"])
        y, fatal, ev = m.spmv(b)
        assert [CO.format_event(*e) for e in ev] == f["lines"], f
        assert not fatal
        assert f["restored"] == bool(np.array_equal(m.elements.view(np.uint32), clean)), f
        assert float(np.sum(y)) == unhex1(f["y_sum"]), f


def _replay(CO, pkg, case, grid, blocks):
    """Replay a cg-coo CLI fixture with the oracle: same triplets, same flips (parsed from the reference's own
    '*** flipping' lines), compare messages / exit status / iteration count."""
    t = pkg.matgen.poisson2d(grid, blocks)
    b = pkg.matgen.rhs(t.N)
    m = CO.CooOracleMatrix(case["mode"], t.cols, t.rows, t.vals, t.N)
    expect = [ln for ln in case["faults"] if not ln.startswith("***")]
    for ln in case["faults"]:
        mm = re.match(r"\*\*\* flipping bit (\d+) at index (\d+) \*\*\*", ln)
        if mm:
            m.flip(int(mm.group(2)), int(mm.group(1)))
    res = m.cg(b, 300, 1e-10)
    assert [CO.format_event(*e) for e in res["events"]] == expect, case
    assert res["fatal"] == (case["exit"] != 0), case
    if case["ran"]:
        assert case["ran"] == f"ran for {res['iterations']} iterations", case


def test_cli_fixture_replay(CO, pkg, G):
    for case in G["cli"]["cases"]:
        _replay(CO, pkg, case, G["cli"]["grid"], G["cli"]["blocks"])
    # the fixture covers all four CONSTRAINTS messages and both fatal ECC ones
    text = " ".join(ln for c in G["cli"]["cases"] for ln in c["faults"])
    for needle in ["row size constraint violated for index", "column size constraint violated for index",
                   "row index order violated at index", "column index order violated at index",
                   "double-bit error detected", "error detected at index"]:
        assert needle in text


def test_rand_driven_injection_order(CO, pkg, G):
    """inject_bitflip draws rand() as COO/CPUContext.cpp:125,136 does: same seed, same flips."""
    libc = ctypes.CDLL(None)
    t = pkg.matgen.poisson2d(G["cli"]["grid"], G["cli"]["blocks"])
    for case in G["cli"]["cases"]:
        if case["seed"] is None or "-x" not in case["extra"]:
            continue
        n, kind = 1, 0
        for a in case["extra"][1:]:
            if a == "VALUE": kind = 1
            elif a == "INDEX": kind = 2
            else: n = int(a)
        m = CO.CooOracleMatrix(case["mode"], t.cols, t.rows, t.vals, t.N)
        libc.srand(case["seed"])
        picked = m.inject(kind, n)
        want = [re.match(r"\*\*\* flipping bit (\d+) at index (\d+)", ln) for ln in case["faults"] if ln.startswith("***")]
        assert [int(w.group(2)) for w in want] == [picked[0]] * n
        assert [int(w.group(1)) for w in want] == picked[1:]


def test_coo_and_csr_oracles_agree_on_symmetric_matrices(CO, O, pkg):
    """result[col] += value * vec[row] over a symmetric matrix is the CSR product, and each output entry is
    accumulated in the same order (ascending row == ascending column of the mirrored entry): same bits."""
    rng = np.random.default_rng(4)
    for t in [pkg.matgen.poisson2d(9, 3), pkg.matgen.poisson3d(5), pkg.matgen.powerlaw(300, 3000, seed=4)]:
        vec = rng.standard_normal(t.N)
        y_coo, _, _ = CO.CooOracleMatrix("none", t.cols, t.rows, t.vals, t.N).spmv(vec)
        y_csr, _, _ = O.OracleMatrix("none", t.cols, t.rows, t.vals, t.N).spmv(vec)
        assert np.array_equal(y_coo.view(np.uint64), y_csr.view(np.uint64))


# ---------------------------------------------------------------- live reference (build container only)
def test_live_reference_random_matrices_and_flips(CO, pkg):
    if not CO.have_ref():
        pytest.skip("oracle/_ref/libref_coo.so not built (no reference tree here)")
    rng = np.random.default_rng(99)
    shapes = [pkg.matgen.poisson2d(7, 2), pkg.matgen.powerlaw(200, 2500, seed=8)]
    for t in shapes:
        vec = rng.standard_normal(t.N)
        for mode in CO.MODES:
            m = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
            rc = CO.RefCooContext(mode)
            rc.create_matrix(t.cols, t.rows, t.vals, t.N)
            assert np.array_equal(m.elements.view(np.uint32), rc.elements.view(np.uint32))
            if mode in ("sec7", "sec8", "secded"):   # correctable: a single flip somewhere
                i, bit = int(rng.integers(0, t.nnz)), int(rng.integers(0, 128))
                m.flip(i, bit); rc.flip(i, bit)
            y, fatal, ev = m.spmv(vec)
            y_ref, text = rc.spmv(vec)
            assert not fatal and [CO.format_event(*e) for e in ev] == text.splitlines()
            assert np.array_equal(y.view(np.uint64), y_ref.view(np.uint64))
            assert np.array_equal(m.elements.view(np.uint32), rc.elements.view(np.uint32))


def test_differential_fuzz_against_the_reference_coo(CO):
    """tools/fuzz_oracle_vs_ref.py --format coo: random matrices x modes x 1..3 flips (any of the 128 bits) in each of
    1..2 elements, the reference's cg-coo target in a forked child (it exit(1)s on fatal faults): same lines, same exit
    status, same repaired elements, same result[] wherever the reference's behaviour is defined."""
    import json
    import os
    import subprocess
    import sys
    if not CO.have_ref():
        pytest.skip("oracle/_ref/libref_coo.so not built (no reference tree here)")
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_oracle_vs_ref.py")
    out = subprocess.run([sys.executable, tool, "--format", "coo", "--trials", "400", "--seed", "20261018"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    stats = json.loads(out.stdout.strip().splitlines()[-1])
    assert stats["trials"] == 400 and stats["fatal"] > 20 and stats["corrected"] > 40 and stats["result_compared"] > 100
