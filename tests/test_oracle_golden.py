"""The CPU oracle against the fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  Integer work bit-exact; the CG trajectories too,
because the oracle keeps the reference's operation order and has no FMA."""
import numpy as np
import pytest

from conftest import unhex, unhex1

REF_MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]


def test_hamming_masks_match_reference_constants(O, golden_ecc):
    L = O.lib()
    for p in range(1, 8):
        for w in range(3):
            assert L.orc_hamming_mask(p, w) == golden_ecc["masks"][p - 1][w]


def test_encode_known_answers(O, golden_ecc):
    L = O.lib()
    for k in golden_ecc["kats"]:
        v, c = unhex1(k["value"]), k["column"]
        assert L.orc_ecc_compute_col8(v, c) == k["syndrome"]
        assert L.orc_ecc_overall_parity(v, c) == k["overall"]
        assert L.orc_encode_column(O.MODE_ID["sed"], v, c) == k["sed"]
        assert L.orc_encode_column(O.MODE_ID["sec7"], v, c) == k["sec7"]
        assert L.orc_encode_column(O.MODE_ID["sec8"], v, c) == k["sec8"]
        assert L.orc_encode_column(O.MODE_ID["secded"], v, c) == k["sec8"]
        assert L.orc_encode_column(O.MODE_ID["none"], v, c) == c


def test_survey_known_answers(O):
    # (value, column) -> SED, SEC7, SEC8/SECDED column words
    L = O.lib()
    kats = [(4.0, 0, 0x00000000, 0xBE000000, 0xBE000000), (-1.0, 1, 0x00000001, 0xAE000001, 0xAF000001),
            (-1.0, 255, 0x800000FF, 0xBC0000FF, 0xBC0000FF), (4.0, 65535, 0x0000FFFF, 0xBE00FFFF, 0xBE00FFFF),
            (-1.0, 16777215, 0x80FFFFFF, 0xBCFFFFFF, 0xBCFFFFFF), (0.5, 123456, 0x8001E240, 0x4A01E240, 0x4A01E240),
            (3.141592653589793, 1000000, 0x000F4240, 0xA80F4240, 0xA90F4240)]
    for v, c, sed, s7, s8 in kats:
        assert L.orc_encode_column(2, v, c) == sed
        assert L.orc_encode_column(3, v, c) == s7
        assert L.orc_encode_column(4, v, c) == s8
        assert L.orc_encode_column(5, v, c) == s8


def test_syndrome_decode_table(O, golden_ecc):
    L = O.lib()
    for d in golden_ecc["decode"]:
        assert L.orc_ecc_get_flipped_bit_col8(d["syndrome"]) == d["bit"]


def test_single_flip_syndromes(O, golden_ecc):
    import struct
    L = O.lib()
    sf = golden_ecc["single_flips"]
    v = unhex1(sf["value"])
    for f in sf["flips"]:
        w = list(struct.unpack("<II", struct.pack("<d", v))) + [sf["encoded"]]
        w[f["bit"] // 32] ^= 1 << (f["bit"] % 32)
        fv = struct.unpack("<d", struct.pack("<II", w[0], w[1]))[0]
        syn = L.orc_ecc_compute_col8(fv, w[2])
        assert syn == f["syndrome"] and L.orc_ecc_overall_parity(fv, w[2]) == f["overall"]
        if f["bit"] == 88:
            assert syn == 0          # column bit 24: only the overall parity sees it
        else:
            assert L.orc_ecc_get_flipped_bit_col8(syn) == f["bit"]


def test_double_flips_are_detected_by_secded(O):
    # all 4560 pairs: overall parity even, syndrome non-zero
    import struct
    L = O.lib()
    v, c = 3.141592653589793, 1000000
    enc = L.orc_encode_column(5, v, c)
    base = list(struct.unpack("<II", struct.pack("<d", v))) + [enc]
    for a in range(96):
        for b in range(a + 1, 96):
            w = list(base)
            w[a // 32] ^= 1 << (a % 32)
            w[b // 32] ^= 1 << (b % 32)
            fv = struct.unpack("<d", struct.pack("<II", w[0], w[1]))[0]
            assert L.orc_ecc_overall_parity(fv, w[2]) == 0
            assert L.orc_ecc_compute_col8(fv, w[2]) != 0


def test_crc32c_check_value(O):
    assert O.lib().orc_crc32c(b"123456789", 9) == 0xE3069283
    assert O.lib().orc_crc32c(b"", 0) == 0


@pytest.mark.parametrize("mode", REF_MODES)
def test_cg_trajectory_matches_reference(O, pkg, golden_cg, mode):
    t = pkg.matgen.poisson2d(golden_cg["grid"])
    assert (t.N, t.nnz) == (golden_cg["N"], golden_cg["nnz"])
    b = unhex(golden_cg["b"])
    assert np.array_equal(b, pkg.matgen.rhs(t.N))
    g = golden_cg["modes"][mode]
    m = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    assert int(m.cols.astype(np.uint64).sum()) == g["cols_sum"]
    assert [int(z) for z in m.cols[:16]] == g["cols_head"]
    y, fatal, ev = m.spmv(b)
    assert not fatal and not ev and np.array_equal(y, unhex(g["spmv_b"]))
    res = m.cg(b, 1000, 1e-10)
    assert res["iterations"] == g["iterations"]
    assert np.array_equal(res["rr_hist"], unhex(g["rr_hist"]))
    assert np.array_equal(res["x"], unhex(g["x"]))


def test_injected_single_flips_match_reference(O, pkg, golden_cg):
    t = pkg.matgen.poisson2d(golden_cg["grid"])
    b = unhex(golden_cg["b"])
    for f in golden_cg["faults"]:
        m = O.OracleMatrix(f["mode"], t.cols, t.rows, t.vals, t.N)
        clean = m.cols.copy()
        m.flip(f["index"], f["bit"])
        res = m.cg(b, 1000, 1e-10)
        lines = [O.format_event(*e) for e in res["events"]]
        assert lines == f["lines"], f
        assert res["iterations"] == f["iterations"]
        assert res["rr_hist"][-1] == unhex1(f["rr_last"])
        assert f["restored"] and np.array_equal(m.cols, clean) and np.array_equal(m.values, t.vals)


def _replay(O, pkg, case, grid, blocks):
    """Replay a CLI fixture with the oracle: same triplets, same flips (parsed from the
    reference's own '*** flipping' lines), compare messages / exit / iteration count."""
    import re
    t = pkg.matgen.poisson2d(grid, blocks)
    b = pkg.matgen.rhs(t.N)
    m = O.OracleMatrix(case["mode"], t.cols, t.rows, t.vals, t.N)
    expect = [ln for ln in case["faults"] if not ln.startswith("***")]
    for ln in case["faults"]:
        mm = re.match(r"\*\*\* flipping bit (\d+) at index (\d+) \*\*\*", ln)
        if mm:
            m.flip(int(mm.group(2)), int(mm.group(1)))
    res = m.cg(b, 300, 1e-10)
    lines = [O.format_event(*e) for e in res["events"]]
    assert lines == expect, case
    assert res["fatal"] == (case["exit"] != 0)
    if case["ran"]:
        assert case["ran"] == f"ran for {res['iterations']} iterations"


def test_cli_fixture_replay(O, pkg, golden_cli):
    for case in golden_cli["cases"]:
        _replay(O, pkg, case, golden_cli["grid"], golden_cli["blocks"])


def test_rand_driven_injection_order(O, pkg, golden_cli):
    """inject_bitflip draws rand() as CSR/CPUContext.cpp:137,148 does: same seed, same flips."""
    import ctypes, re
    libc = ctypes.CDLL(None)
    t = pkg.matgen.poisson2d(golden_cli["grid"], golden_cli["blocks"])
    for case in golden_cli["cases"]:
        if case["seed"] is None or "-x" not in case["extra"]:
            continue
        n = 1
        kind = 0
        for a in case["extra"][1:]:
            if a == "VALUE": kind = 1
            elif a == "INDEX": kind = 2
            else: n = int(a)
        m = O.OracleMatrix(case["mode"], t.cols, t.rows, t.vals, t.N)
        libc.srand(case["seed"])
        picked = m.inject(kind, n)
        want = [re.match(r"\*\*\* flipping bit (\d+) at index (\d+)", ln) for ln in case["faults"] if ln.startswith("***")]
        assert [int(w.group(2)) for w in want] == [picked[0]] * n
        assert [int(w.group(1)) for w in want] == picked[1:]


# ---- crc32c: repo-defined, no reference to pin against; properties only ----
def test_crc32c_mode_roundtrip_and_detection(O, pkg):
    t = pkg.matgen.poisson2d(8)
    b = pkg.matgen.rhs(t.N)
    m = O.OracleMatrix("crc32c", t.cols, t.rows, t.vals, t.N)
    plain = O.OracleMatrix("none", t.cols, t.rows, t.vals, t.N)
    assert np.array_equal(m.cols & 0x00FFFFFF, plain.cols)
    y, fatal, ev = m.spmv(b)
    y0, _, _ = plain.spmv(b)
    assert not fatal and not ev and np.array_equal(y, y0)
    rows = m.rows
    rng = np.random.default_rng(3)
    for _ in range(40):
        mm = O.OracleMatrix("crc32c", t.cols, t.rows, t.vals, t.N)
        i = int(rng.integers(0, t.nnz)); bit = int(rng.integers(0, 96))
        row = int(np.searchsorted(rows, i, side="right") - 1)
        k = i - int(rows[row])
        mm.flip(i, bit)
        y, fatal, ev = mm.spmv(b)
        harmless = bit >= 88 and k >= 4 and (rows[row + 1] - rows[row]) >= 4   # unused top byte, masked off
        if harmless:
            assert not fatal and np.array_equal(y, y0)
        else:
            assert fatal and ev == [(9, row, 0)]
