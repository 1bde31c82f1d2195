"""ctypes bindings of the COO CPU oracle (oracle/libcoo_oracle.so) and of the window onto the
unmodified reference COO target (oracle/_ref/libref_coo.so, oracle/_ref/cg-coo).

TEST INFRASTRUCTURE (see coo_oracle.h): importable from tests/ and tests/golden/ only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import oracle_py as O   # build(), capture_stdout, run_cg_csr (runs any of the reference CLIs), FAKETIME_SO

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]
MODE_ID = {m: i for i, m in enumerate(MODES)}
FATAL_EVENTS = {1, 4, 5, 6, 7, 8}

# the reference's struct coo_element {uint32_t col; uint32_t row; double value} (COO/ecc.h:11-16)
ELEM = np.dtype([("col", np.uint32), ("row", np.uint32), ("value", np.float64)])


class Event(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_uint32), ("bit", C.c_uint32)]


class CooMatrix(C.Structure):
    _fields_ = [("N", C.c_uint32), ("nnz", C.c_uint32), ("mode", C.c_int32), ("elements", C.c_void_p)]


_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libcoo_oracle.so")
        if not os.path.exists(so):
            O.build(force=True)
        L = C.CDLL(so)
        L.coo_hamming_mask.restype = C.c_uint32
        L.coo_hamming_mask.argtypes = [C.c_int, C.c_int]
        for f in (L.coo_ecc_compute_col8, L.coo_ecc_overall_parity):
            f.restype = C.c_uint32
            f.argtypes = [C.c_uint32, C.c_uint32, C.c_double]
        L.coo_ecc_get_flipped_bit_col8.restype = C.c_uint32
        L.coo_ecc_get_flipped_bit_col8.argtypes = [C.c_uint32]
        L.coo_encode_column.restype = C.c_uint32
        L.coo_encode_column.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_double]
        L.coo_create_matrix.restype = C.POINTER(CooMatrix)
        L.coo_create_matrix.argtypes = [C.c_int, _u32p, _u32p, _f64p, C.c_int, C.c_int]
        L.coo_destroy_matrix.argtypes = [C.POINTER(CooMatrix)]
        L.coo_spmv.restype = C.c_int
        L.coo_spmv.argtypes = [C.POINTER(CooMatrix), _f64p, _f64p, C.POINTER(Event), C.c_int, C.POINTER(C.c_int)]
        L.coo_flip_bit.argtypes = [C.POINTER(CooMatrix), C.c_int, C.c_int]
        L.coo_inject_bitflip.argtypes = [C.POINTER(CooMatrix), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.coo_cg_solve.restype = C.c_int
        L.coo_cg_solve.argtypes = [C.POINTER(CooMatrix), _f64p, _f64p, _f64p, _f64p, _f64p, C.c_int, C.c_double,
                                   C.c_void_p, C.POINTER(Event), C.c_int, C.POINTER(C.c_int)]
        L.coo_format_event.restype = C.c_int
        L.coo_format_event.argtypes = [C.POINTER(Event), C.c_char_p, C.c_int]
        _lib = L
    return _lib


def format_event(kind: int, index: int, bit: int) -> str:
    e = Event(kind, index, bit)
    buf = C.create_string_buffer(128)
    lib().coo_format_event(C.byref(e), buf, 128)
    return buf.value.decode()


class CooOracleMatrix:
    """Owns a coo_matrix; `elements` is a numpy structured view of its storage."""

    def __init__(self, mode: str, cols, rows, vals, N: int):
        self.mode = mode
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.N, self.nnz = int(N), int(cols.shape[0])
        self._m = lib().coo_create_matrix(MODE_ID[mode], cols, rows, vals, self.N, self.nnz)

    @property
    def elements(self):
        buf = (C.c_char * (16 * max(self.nnz, 1))).from_address(self._m.contents.elements)
        return np.frombuffer(buf, dtype=ELEM)[:self.nnz]

    def flip(self, index: int, bit: int) -> None:
        lib().coo_flip_bit(self._m, index, bit)

    def inject(self, kind: int, num_flips: int):
        picked = (C.c_int * (1 + num_flips))()
        lib().coo_inject_bitflip(self._m, kind, num_flips, picked)
        return list(picked)

    def spmv(self, vec: np.ndarray, cap: int = 4096):
        """-> (result, fatal, [(kind, index, bit), ...])"""
        out = np.zeros(self.N, dtype=np.float64)
        log = (Event * cap)()
        n = C.c_int(0)
        fatal = lib().coo_spmv(self._m, np.ascontiguousarray(vec, dtype=np.float64), out, log, cap, C.byref(n))
        return out, bool(fatal), [(log[i].kind, log[i].index, log[i].bit) for i in range(n.value)]

    def cg(self, b: np.ndarray, max_itrs: int, conv: float, cap: int = 4096):
        N = self.N
        x = np.zeros(N); r = np.zeros(N); p = np.zeros(N); w = np.zeros(N)
        hist = np.zeros(max(max_itrs, 1))
        log = (Event * cap)()
        n = C.c_int(0)
        it = lib().coo_cg_solve(self._m, np.ascontiguousarray(b, dtype=np.float64), x, r, p, w, max_itrs, conv,
                                hist.ctypes.data, log, cap, C.byref(n))
        fatal = it < 0
        iters = (-1 - it) if fatal else it
        return dict(iterations=iters, rr_hist=hist[:iters].copy(), x=x, fatal=fatal,
                    events=[(log[i].kind, log[i].index, log[i].bit) for i in range(n.value)])

    def __del__(self):
        try:
            lib().coo_destroy_matrix(self._m)
        except Exception:
            pass


# ---------------------------------------------------------------------------
# the unmodified reference COO target (only where oracle/_ref was built / shipped)
# ---------------------------------------------------------------------------
REF_SO = os.path.join(HERE, "_ref", "libref_coo.so")
REF_EXE = os.path.join(HERE, "_ref", "cg-coo")


def have_ref() -> bool:
    O.build()
    if os.path.exists("/root/reference/cg.cpp") and not os.path.exists(REF_SO):
        O.build(force=True)
    return os.path.exists(REF_SO) and os.path.exists(REF_EXE)


_ref = None


def ref():
    global _ref
    if _ref is None:
        R = C.CDLL(REF_SO)
        R.refcoo_hamming_mask.restype = C.c_uint32
        R.refcoo_hamming_mask.argtypes = [C.c_int, C.c_int]
        for f in (R.refcoo_ecc_compute_col8, R.refcoo_ecc_overall_parity):
            f.restype = C.c_uint32
            f.argtypes = [C.c_uint32, C.c_uint32, C.c_double]
        R.refcoo_ecc_get_flipped_bit_col8.restype = C.c_uint32
        R.refcoo_ecc_get_flipped_bit_col8.argtypes = [C.c_uint32]
        R.refcoo_context_create.restype = C.c_void_p
        R.refcoo_context_create.argtypes = [C.c_char_p]
        R.refcoo_context_destroy.argtypes = [C.c_void_p]
        R.refcoo_create_matrix.restype = C.c_void_p
        R.refcoo_create_matrix.argtypes = [C.c_void_p, _u32p, _u32p, _f64p, C.c_int, C.c_int]
        R.refcoo_destroy_matrix.argtypes = [C.c_void_p, C.c_void_p]
        R.refcoo_matrix_elements.restype = C.c_void_p
        R.refcoo_matrix_elements.argtypes = [C.c_void_p]
        R.refcoo_create_vector.restype = C.c_void_p
        R.refcoo_create_vector.argtypes = [C.c_void_p, C.c_int]
        R.refcoo_destroy_vector.argtypes = [C.c_void_p, C.c_void_p]
        R.refcoo_map_vector.restype = C.POINTER(C.c_double)
        R.refcoo_map_vector.argtypes = [C.c_void_p, C.c_void_p]
        R.refcoo_copy_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        R.refcoo_dot.restype = C.c_double
        R.refcoo_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        R.refcoo_spmv.argtypes = [C.c_void_p] * 4
        R.refcoo_inject_bitflip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        R.refcoo_cg_iterations.restype = C.c_int
        R.refcoo_cg_iterations.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p]
        _ref = R
    return _ref


class RefCooContext:
    """The reference's COO CPUContext for one mode, driven through CGContext's API."""

    def __init__(self, mode: str):
        self.R = ref()
        self.mode = mode
        self.ctx = self.R.refcoo_context_create(mode.encode())
        self.N = 0

    def create_matrix(self, cols, rows, vals, N):
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.N, self.nnz = int(N), int(cols.shape[0])
        self.A = self.R.refcoo_create_matrix(self.ctx, cols, rows, vals, self.N, self.nnz)
        return self.A

    @property
    def elements(self):
        buf = (C.c_char * (16 * max(self.nnz, 1))).from_address(self.R.refcoo_matrix_elements(self.A))
        return np.frombuffer(buf, dtype=ELEM)[:self.nnz]

    def flip(self, index: int, bit: int) -> None:
        """COO/CPUContext.cpp:138 on a chosen element."""
        words = self.elements.view(np.uint32).reshape(-1, 4)
        words[index, bit // 32] ^= np.uint32(1 << (bit % 32))

    def vector(self, data=None):
        v = self.R.refcoo_create_vector(self.ctx, self.N)
        h = self.view(v)
        h[:] = 0.0 if data is None else data
        return v

    def view(self, v):
        return np.ctypeslib.as_array(self.R.refcoo_map_vector(self.ctx, v), shape=(self.N,))

    def spmv(self, vec):
        """-> (result, text the reference printed); non-fatal cases only (a fatal one exit(1)s the process)"""
        v, out = self.vector(vec), self.vector()
        with O.capture_stdout() as cap:
            self.R.refcoo_spmv(self.ctx, self.A, v, out)
        res = self.view(out).copy()
        for t in (v, out):
            self.R.refcoo_destroy_vector(self.ctx, t)
        return res, cap.text

    def cg(self, b, max_itrs, conv):
        bv = self.vector(b); x = self.vector(); r = self.vector(); p = self.vector(); w = self.vector()
        self.R.refcoo_copy_vector(self.ctx, r, bv)
        self.R.refcoo_copy_vector(self.ctx, p, r)
        rr = C.c_double(self.R.refcoo_dot(self.ctx, r, r))
        hist = np.zeros(max(max_itrs, 1))
        with O.capture_stdout() as cap:
            it = self.R.refcoo_cg_iterations(self.ctx, self.A, x, r, p, w, max_itrs, conv, C.byref(rr), hist.ctypes.data)
        out = dict(iterations=it, rr_hist=hist[:it].copy(), x=self.view(x).copy(), text=cap.text)
        for v in (bv, x, r, p, w):
            self.R.refcoo_destroy_vector(self.ctx, v)
        return out
