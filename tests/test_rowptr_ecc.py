"""Row-pointer protection (scope row f.2; repository-defined, PARITY UNPINNED -- the reference trusts rows[]):
the device code of csrc/rowsecc.cuh, compiled for the host and run through the CSR kernel emulator
(tests/helpers/cuda_shim_mt/), against the independent bit-by-bit restatement in
oracle/rowptr_ecc.c.  Opt-in at compile time until it has run on a GPU (tools/ab_variants.sh carries `rowsecc`)."""
import ctypes as C
import os

import numpy as np
import pytest

from test_csr_kernel_logic import EmuMatrix, _build, shapes  # noqa: F401  (the emulator harness)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def RO(O):
    O.build()
    L = C.CDLL(os.path.join(ROOT, "oracle", "librowptr_ecc.so"))
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    L.rpo_encode.argtypes = [u32p, C.c_int]
    L.rpo_check_all.restype = C.c_int
    L.rpo_check_all.argtypes = [u32p, C.c_int, u32p, C.c_int, C.POINTER(C.c_int)]
    return L


@pytest.fixture(scope="module")
def EMU_ROWS(tmp_path_factory):
    L = _build(tmp_path_factory.mktemp("csr_emu_rows"), "rowsecc", [])
    L.emu_protect_rows.restype = C.c_int
    L.emu_protect_rows.argtypes = [C.c_void_p]
    return L


def oracle_prepass(RO, words):
    """-> (events [(kind, pair, bit)], fatal); repairs `words` in place"""
    ev = np.zeros(3 * 256, dtype=np.uint32)
    fatal = C.c_int(0)
    n = RO.rpo_check_all(words, len(words), ev, 256, C.byref(fatal))
    return [tuple(int(x) for x in ev[3 * i:3 * i + 3]) for i in range(min(n, 256))], bool(fatal.value)


def protected(EMU_ROWS, pkg, RO, mode, t):
    em = EmuMatrix(EMU_ROWS, pkg, mode, t)
    assert EMU_ROWS.emu_protect_rows(em.h) == 1
    n_words = (t.N + 2) & ~1
    enc = np.zeros(n_words, dtype=np.uint32)
    enc[:t.N + 1] = t.row_pointers()
    RO.rpo_encode(enc, n_words)
    got = np.ctypeslib.as_array(EMU_ROWS.emu_rowptr(em.h), shape=(n_words,))
    return em, enc, got


@pytest.mark.parametrize("mode", ["none", "secded", "constraints", "crc32c"])
def test_encode_and_clean_spmv(EMU_ROWS, pkg, O, RO, mode):
    rng = np.random.default_rng(1)
    for name in ("p3d_12", "powerlaw", "long_rows", "ragged", "one_by_one"):
        t = shapes(pkg.matgen)[name]()
        em, enc, got = protected(EMU_ROWS, pkg, RO, mode, t)
        assert np.array_equal(got, enc), name                         # same check bits as the oracle
        assert np.array_equal(got[:t.N + 1] & 0x0FFFFFFF, t.row_pointers()), name
        vec = rng.standard_normal(t.N)
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        y, fatal, events, _ = em.spmv(vec)
        y0, _, _ = om.spmv(vec)
        lens = np.bincount(t.rows, minlength=t.N)
        scale = np.abs(t.vals).max() * np.abs(vec).max() * max(int(lens.max()), 1)
        assert not fatal and events == [] and np.max(np.abs(y - y0)) <= 1e-13 * scale, (name, events)
        assert np.array_equal(y[lens <= 32].view(np.uint64), y0[lens <= 32].view(np.uint64)), name
        em.close()


def test_every_single_flip_of_a_pair_is_repaired(EMU_ROWS, pkg, O, RO):
    t = shapes(pkg.matgen)["p3d_12"]()
    vec = np.random.default_rng(2).standard_normal(t.N)
    om = O.OracleMatrix("secded", t.cols, t.rows, t.vals, t.N)
    y0, _, _ = om.spmv(vec)
    for pair in (0, 128, (t.N + 1) // 2):                   # first, a tile boundary (row 256), last
        for bit in range(64):
            em, enc, got = protected(EMU_ROWS, pkg, RO, "secded", t)
            word = 2 * pair + bit // 32
            got[word] ^= np.uint32(1 << (bit % 32))
            bad = enc.copy()
            bad[word] ^= np.uint32(1 << (bit % 32))
            ev0, fatal0 = oracle_prepass(RO, bad)
            y, fatal, events, _ = em.spmv(vec)
            assert (events, fatal) == (ev0, fatal0) == ([(10, pair, bit)], False), (pair, bit, events, ev0)
            assert np.array_equal(got, enc) and np.array_equal(bad, enc), (pair, bit)      # repaired in memory
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), (pair, bit)
            em.close()


def test_double_flips_are_detected_and_mixed_faults_keep_their_order(EMU_ROWS, pkg, O, RO):
    t = shapes(pkg.matgen)["p3d_12"]()
    vec = np.random.default_rng(3).standard_normal(t.N)
    rng = np.random.default_rng(4)
    for trial in range(25):
        em, enc, got = protected(EMU_ROWS, pkg, RO, "secded", t)
        om = O.OracleMatrix("secded", t.cols, t.rows, t.vals, t.N)
        bad = enc.copy()
        n_pairs = len(enc) // 2
        for pair in rng.choice(n_pairs, 1 + trial % 3, replace=False):       # 1..3 faulty pairs
            for bit in rng.choice(64, 1 + int(rng.integers(0, 2)), replace=False):   # 1 or 2 flips each
                word = 2 * int(pair) + int(bit) // 32
                got[word] ^= np.uint32(1 << (int(bit) % 32))
                bad[word] ^= np.uint32(1 << (int(bit) % 32))
        index, ebit = int(rng.integers(0, t.nnz)), int(rng.integers(0, 96))   # and one element fault
        em.flip(index, ebit)
        om.flip(index, ebit)
        ev0, fatal0 = oracle_prepass(RO, bad)
        if not fatal0:
            _, fatal0, ev_el = om.spmv(vec)
            ev0 = ev0 + ev_el
        y, fatal, events, _ = em.spmv(vec)
        assert (events, fatal) == (ev0, fatal0), (trial, events, ev0)
        em.close()
