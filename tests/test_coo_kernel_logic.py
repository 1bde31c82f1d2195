"""The LOGIC of the COO kernels (abft-sparse-cg_b200/csrc/coo.cuh), executed on the CPU.

There is no GPU in the build container, so the CPU suite checks the COO kernels this way: the product's own kernel
source, unmodified, is compiled by g++ against a shim of the CUDA built-ins (tests/helpers/cuda_shim/) and run one
emulated thread after the other (tests/helpers/coo_emul.cpp), then compared with the COO oracle -- which is pinned
against the reference (tests/test_oracle_coo.py).  Bit-exact bars, as for the GPU tests (tests/test_gpu_coo.py, the
same cases through the C ABI on a B200).  What this cannot see: races, memory ordering, launch configuration,
performance.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]
FATAL = {1, 4, 5, 6, 7, 8}
ELEM = np.dtype([("col", np.uint32), ("row", np.uint32), ("value", np.float64)])


class Event(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_uint32), ("bit", C.c_uint32)]


@pytest.fixture(scope="module", params=["caller_order", "grouped"])
def EMU(request, tmp_path_factory):
    """Both element layouts (coo.cuh): the shipped one and the opt-in grouped one (ABFT_CUDA_COO_LAYOUT=grouped)."""
    so = str(tmp_path_factory.mktemp("emu") / "libcoo_emul.so")
    h = os.path.join(ROOT, "tests", "helpers")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I" + os.path.join(h, "cuda_shim"),
                    "-o", so, os.path.join(h, "coo_emul.cpp")], check=True)
    L = C.CDLL(so)
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.emu_create.restype = C.c_void_p
    L.emu_create.argtypes = [C.c_int, u32p, u32p, f64p, C.c_int, C.c_int, C.c_int]
    L.emu_destroy.argtypes = [C.c_void_p]
    L.emu_elements.restype = C.c_void_p
    L.emu_elements.argtypes = [C.c_void_p]
    L.emu_flip.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.emu_spmv.restype = C.c_int
    L.emu_spmv.argtypes = [C.c_void_p, f64p, f64p, C.POINTER(Event), C.c_int]
    L.grouped = int(request.param == "grouped")
    return L


@pytest.fixture(scope="module")
def CO(O):
    import coo_oracle_py
    coo_oracle_py.lib()
    return coo_oracle_py


class EmuMatrix:
    def __init__(self, L, mode, t):
        self.L, self.mode, self.N, self.nnz = L, mode, t.N, int(t.cols.shape[0])
        self.h = L.emu_create(MODES.index(mode), np.ascontiguousarray(t.cols, dtype=np.uint32),
                              np.ascontiguousarray(t.rows, dtype=np.uint32), np.ascontiguousarray(t.vals, dtype=np.float64),
                              t.N, self.nnz, getattr(L, "grouped", 0))
        assert self.h

    @property
    def elements(self):
        buf = (C.c_char * (16 * max(self.nnz, 1))).from_address(self.L.emu_elements(self.h))
        return np.frombuffer(buf, dtype=ELEM)[:self.nnz]

    def flip(self, index, bit):
        self.L.emu_flip(self.h, index, bit)

    def spmv(self, vec):
        """-> (result, fatal, events) with the events ordered and cut the way abft_cuda_poll_events does"""
        out = np.full(self.N, np.nan)
        log = (Event * 4096)()
        n = self.L.emu_spmv(self.h, np.ascontiguousarray(vec, dtype=np.float64), out, log, 4096)
        assert n >= 0
        raw = sorted(((log[i].index, log[i].bit, log[i].kind) for i in range(n)))
        events, fatal = [], False
        for index, bit, kind in raw:
            if self.mode == "constraints" and kind in (7, 8):
                bit = 0      # COL_* events carry the index in `bit` for ordering only
            events.append((kind, index, bit))
            if kind in FATAL:
                fatal = True
                break
        return out, fatal, events

    def close(self):
        self.L.emu_destroy(self.h)


def shapes(mg):
    return {"p2d_24": lambda: mg.poisson2d(24), "p2d_9_b3": lambda: mg.poisson2d(9, 3), "p3d_8": lambda: mg.poisson3d(8),
            "powerlaw": lambda: mg.powerlaw(2000, 40000, seed=3), "one": lambda: mg.Triplets([0], [0], [2.5], 1)}


@pytest.mark.parametrize("mode", MODES)
def test_encode_and_spmv_bit_exact(EMU, CO, pkg, mode):
    rng = np.random.default_rng(2)
    for name, build in shapes(pkg.matgen).items():
        t = build()
        em = EmuMatrix(EMU, mode, t)
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        assert np.array_equal(em.elements.view(np.uint32), om.elements.view(np.uint32)), name
        vec = rng.standard_normal(t.N)
        y, fatal, events = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert not fatal and events == [] == ev0, name
        assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), name
        em.close()


@pytest.mark.parametrize("mode", ["sed", "sec7", "sec8", "secded"])
def test_all_128_single_flips(EMU, CO, pkg, mode):
    t = pkg.matgen.poisson2d(12, 2)
    vec = pkg.matgen.rhs(t.N)
    for bit in range(128):
        em = EmuMatrix(EMU, mode, t)
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        em.flip(77, bit); om.flip(77, bit)
        y, fatal, events = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), bit
        assert np.array_equal(em.elements.view(np.uint32), om.elements.view(np.uint32)), bit
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), bit
        em.close()


@pytest.mark.parametrize("mode", ["none", "sed", "sec7", "sec8", "secded"])
def test_double_flips_and_silent_index_faults(EMU, CO, pkg, mode):
    """Faults no check repairs -- any flip in mode none; double flips: undetected by sed / sec8 (parity stays even),
    mis-corrected by sec7, detected by secded -- including the ones that move a product to another output entry
    (the stray list and its fix-up kernel) or out of the matrix."""
    t = pkg.matgen.poisson2d(16)
    vec = pkg.matgen.rhs(t.N)
    rng = np.random.default_rng(12)
    rerouted = 0
    for trial in range(150):
        em = EmuMatrix(EMU, mode, t)
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        n_el = 1 + int(rng.integers(0, 3))              # 1..3 corrupted elements per trial
        for index in rng.choice(t.nnz, n_el, replace=False):
            bits = [int(rng.integers(0, 9))] if mode == "none" else [int(b) for b in rng.choice(40, 2, replace=False)]
            for b in bits:
                em.flip(int(index), b); om.flip(int(index), b)
        y, fatal, events = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), trial
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), trial
            assert np.array_equal(em.elements.view(np.uint32), om.elements.view(np.uint32)), trial
            clean, _, _ = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N).spmv(vec)
            rerouted += int(np.count_nonzero(y0 != clean) >= 2)
        em.close()
    if mode != "secded":
        assert rerouted > 0     # the stray path was exercised


def test_two_strays_into_the_same_output_entry(EMU, CO, pkg):
    """Two elements of different groups whose corrupted columns coincide, plus one that leaves and one that joins a
    group whose own member strays away: the fix-up merges by element index."""
    t = pkg.matgen.poisson2d(8)
    vec = pkg.matgen.rhs(t.N)
    cols = t.cols.astype(np.int64)
    em = EmuMatrix(EMU, "none", t)
    om = CO.CooOracleMatrix("none", t.cols, t.rows, t.vals, t.N)
    # elements 5, 40, 90 all get column 17; element with column 17 (first one) gets column 3
    def set_col(index, newcol):
        diff = int(cols[index]) ^ int(newcol)
        for b in range(24):
            if diff >> b & 1:
                em.flip(index, b); om.flip(index, b)
    for idx in (90, 5, 40):
        set_col(idx, 17)
    member = int(np.flatnonzero(cols == 17)[0])
    set_col(member, 3)
    y, fatal, events = em.spmv(vec)
    y0, _, _ = om.spmv(vec)
    assert not fatal and events == []
    assert np.array_equal(y.view(np.uint64), y0.view(np.uint64))
    em.close()


def test_constraints_violations(EMU, CO, pkg):
    t = pkg.matgen.poisson2d(16)
    vec = pkg.matgen.rhs(t.N)
    kinds = set()
    rng = np.random.default_rng(5)
    cases = [(300, 45), (300, 23), (301, 38), (302, 1), (0, 32), (t.nnz - 1, 40), (t.nnz - 1, 3)]
    cases += [(int(rng.integers(0, t.nnz)), int(rng.integers(0, 64))) for _ in range(120)]
    for index, bit in cases:
        em = EmuMatrix(EMU, "constraints", t)
        om = CO.CooOracleMatrix("constraints", t.cols, t.rows, t.vals, t.N)
        em.flip(index, bit); om.flip(index, bit)
        y, fatal, events = em.spmv(vec)
        y0, fatal0, ev0 = om.spmv(vec)
        assert (fatal, events) == (fatal0, ev0), (index, bit)
        kinds |= {e[0] for e in events}
        if not fatal:
            assert np.array_equal(y.view(np.uint64), y0.view(np.uint64)), (index, bit)
        em.close()
    assert kinds == {5, 6, 7, 8}


def test_cli_fixture_replay(EMU, CO, O, pkg):
    """The reference cg-coo CLI cases (tests/golden/coo_golden.json) through the emulated kernels, with the CG loop
    of cg.cpp:83-118 around them (sequential reductions, as on the CPU: iteration counts must be equal)."""
    G = load_golden("coo_golden.json")["cli"]
    t = pkg.matgen.poisson2d(G["grid"], G["blocks"])
    b = pkg.matgen.rhs(t.N)
    V = O.lib()
    for case in G["cases"]:
        em = EmuMatrix(EMU, case["mode"], t)
        for ln in case["faults"]:
            mm = re.match(r"\*\*\* flipping bit (\d+) at index (\d+) \*\*\*", ln)
            if mm:
                em.flip(int(mm.group(2)), int(mm.group(1)))
        N = t.N
        x = np.zeros(N); r = b.copy(); p = b.copy()
        rr = V.orc_dot(r, r, N)
        lines, fatal, itr = [], False, 0
        while itr < 300 and rr > 1e-10:
            w, fatal, events = em.spmv(p)
            lines += [CO.format_event(*e) for e in events]
            if fatal:
                break
            pw = V.orc_dot(p, w, N)
            rr_new = V.orc_calc_xr(x, r, p, w, rr / pw, N)
            V.orc_calc_p(p, r, rr_new / rr, N)
            rr = rr_new
            itr += 1
        assert lines == [ln for ln in case["faults"] if not ln.startswith("***")], case
        assert fatal == (case["exit"] != 0), case
        if case["ran"]:
            assert case["ran"] == f"ran for {itr} iterations", case
        em.close()
