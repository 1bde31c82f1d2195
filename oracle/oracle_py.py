"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and of the window
onto the unmodified reference (oracle/_ref/libref_csr.so).

TEST INFRASTRUCTURE: importable from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"]
MODE_ID = {m: i for i, m in enumerate(MODES)}

EV_NAMES = {1: "sed_detected", 2: "corrected_bit", 3: "corrected_parity", 4: "double_bit",
            5: "row_size", 6: "row_order", 7: "col_size", 8: "col_order", 9: "crc_row"}
FATAL_EVENTS = {1, 4, 5, 6, 7, 8, 9}


class Event(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_uint32), ("bit", C.c_uint32)]


class OrcMatrix(C.Structure):
    _fields_ = [("N", C.c_uint32), ("nnz", C.c_uint32), ("mode", C.c_int32),
                ("cols", C.POINTER(C.c_uint32)), ("rows", C.POINTER(C.c_uint32)),
                ("values", C.POINTER(C.c_double))]


def build(force: bool = False) -> None:
    """make -C oracle (liboracle.so, libfaketime.so, and _ref/ when the reference tree exists)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or not os.path.exists(
            os.path.join(HERE, "libfaketime.so")) or (
            os.path.exists("/root/reference/cg.cpp") and not os.path.exists(os.path.join(HERE, "_ref", "libref_csr.so"))):
        subprocess.run(["make", "-C", HERE, "-s"], check=True, stdout=subprocess.DEVNULL)


_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.orc_hamming_mask.restype = C.c_uint32
        L.orc_hamming_mask.argtypes = [C.c_int, C.c_int]
        for f in (L.orc_ecc_compute_col8, L.orc_ecc_overall_parity):
            f.restype = C.c_uint32
            f.argtypes = [C.c_double, C.c_uint32]
        L.orc_ecc_get_flipped_bit_col8.restype = C.c_uint32
        L.orc_ecc_get_flipped_bit_col8.argtypes = [C.c_uint32]
        L.orc_encode_column.restype = C.c_uint32
        L.orc_encode_column.argtypes = [C.c_int, C.c_double, C.c_uint32]
        L.orc_crc32c.restype = C.c_uint32
        L.orc_crc32c.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_crc32c_row.restype = C.c_uint32
        L.orc_crc32c_row.argtypes = [_u32p, _f64p, C.c_uint32]
        L.orc_create_matrix.restype = C.POINTER(OrcMatrix)
        L.orc_create_matrix.argtypes = [C.c_int, _u32p, _u32p, _f64p, C.c_int, C.c_int]
        L.orc_destroy_matrix.argtypes = [C.POINTER(OrcMatrix)]
        L.orc_spmv.restype = C.c_int
        L.orc_spmv.argtypes = [C.POINTER(OrcMatrix), _f64p, _f64p, C.POINTER(Event), C.c_int, C.POINTER(C.c_int)]
        L.orc_dot.restype = C.c_double
        L.orc_dot.argtypes = [_f64p, _f64p, C.c_int]
        L.orc_calc_xr.restype = C.c_double
        L.orc_calc_xr.argtypes = [_f64p, _f64p, _f64p, _f64p, C.c_double, C.c_int]
        L.orc_calc_p.argtypes = [_f64p, _f64p, C.c_double, C.c_int]
        L.orc_flip_bit.argtypes = [C.POINTER(OrcMatrix), C.c_int, C.c_int]
        L.orc_inject_bitflip.argtypes = [C.POINTER(OrcMatrix), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_cg_solve.restype = C.c_int
        L.orc_cg_solve.argtypes = [C.POINTER(OrcMatrix), _f64p, _f64p, _f64p, _f64p, _f64p, C.c_int, C.c_double,
                                   C.c_void_p, C.POINTER(Event), C.c_int, C.POINTER(C.c_int)]
        L.orc_format_event.restype = C.c_int
        L.orc_format_event.argtypes = [C.POINTER(Event), C.c_char_p, C.c_int]
        _lib = L
    return _lib


def format_event(kind: int, index: int, bit: int) -> str:
    e = Event(kind, index, bit)
    buf = C.create_string_buffer(128)
    lib().orc_format_event(C.byref(e), buf, 128)
    return buf.value.decode()


class OracleMatrix:
    """Owns an orc_matrix; numpy views of its storage."""

    def __init__(self, mode: str, cols, rows, vals, N: int):
        self.mode = mode
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.N, self.nnz = int(N), int(cols.shape[0])
        self._m = lib().orc_create_matrix(MODE_ID[mode], cols, rows, vals, self.N, self.nnz)

    @property
    def cols(self):
        return np.ctypeslib.as_array(self._m.contents.cols, shape=(max(self.nnz, 1),))[:self.nnz]

    @property
    def rows(self):
        return np.ctypeslib.as_array(self._m.contents.rows, shape=(self.N + 1,))

    @property
    def values(self):
        return np.ctypeslib.as_array(self._m.contents.values, shape=(max(self.nnz, 1),))[:self.nnz]

    def flip(self, index: int, bit: int) -> None:
        lib().orc_flip_bit(self._m, index, bit)

    def inject(self, kind: int, num_flips: int):
        picked = (C.c_int * (1 + num_flips))()
        lib().orc_inject_bitflip(self._m, kind, num_flips, picked)
        return list(picked)

    def spmv(self, vec: np.ndarray, cap: int = 4096):
        """-> (result, fatal, [(kind, index, bit), ...])"""
        out = np.zeros(self.N, dtype=np.float64)
        log = (Event * cap)()
        n = C.c_int(0)
        fatal = lib().orc_spmv(self._m, np.ascontiguousarray(vec, dtype=np.float64), out, log, cap, C.byref(n))
        return out, bool(fatal), [(log[i].kind, log[i].index, log[i].bit) for i in range(n.value)]

    def cg(self, b: np.ndarray, max_itrs: int, conv: float, cap: int = 4096):
        """cg.cpp:83-118 -> dict(iterations, rr_hist, x, fatal, events)"""
        N = self.N
        x = np.zeros(N); r = np.zeros(N); p = np.zeros(N); w = np.zeros(N)
        hist = np.zeros(max(max_itrs, 1))
        log = (Event * cap)()
        n = C.c_int(0)
        it = lib().orc_cg_solve(self._m, np.ascontiguousarray(b, dtype=np.float64), x, r, p, w, max_itrs, conv,
                                hist.ctypes.data, log, cap, C.byref(n))
        fatal = it < 0
        iters = (-1 - it) if fatal else it
        return dict(iterations=iters, rr_hist=hist[:iters].copy(), x=x, r=r, p=p, fatal=fatal,
                    events=[(log[i].kind, log[i].index, log[i].bit) for i in range(n.value)])

    def __del__(self):
        try:
            lib().orc_destroy_matrix(self._m)
        except Exception:
            pass


# ---------------------------------------------------------------------------
# the unmodified reference (only where oracle/_ref was built / shipped)
# ---------------------------------------------------------------------------
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libref_csr.so")
REF_EXE = os.path.join(REF_DIR, "cg-csr")
FAKETIME_SO = os.path.join(HERE, "libfaketime.so")


def have_ref() -> bool:
    build()
    return os.path.exists(REF_SO) and os.path.exists(REF_EXE)


_ref = None


def ref():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_SO)
        R.ref_hamming_mask.restype = C.c_uint32
        R.ref_hamming_mask.argtypes = [C.c_int, C.c_int]
        for f in (R.ref_ecc_compute_col8, R.ref_ecc_overall_parity):
            f.restype = C.c_uint32
            f.argtypes = [C.c_double, C.c_uint32]
        R.ref_ecc_get_flipped_bit_col8.restype = C.c_uint32
        R.ref_ecc_get_flipped_bit_col8.argtypes = [C.c_uint32]
        R.ref_context_create.restype = C.c_void_p
        R.ref_context_create.argtypes = [C.c_char_p]
        R.ref_context_destroy.argtypes = [C.c_void_p]
        R.ref_create_matrix.restype = C.c_void_p
        R.ref_create_matrix.argtypes = [C.c_void_p, _u32p, _u32p, _f64p, C.c_int, C.c_int]
        R.ref_destroy_matrix.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_matrix_cols.restype = C.POINTER(C.c_uint32)
        R.ref_matrix_rows.restype = C.POINTER(C.c_uint32)
        R.ref_matrix_values.restype = C.POINTER(C.c_double)
        for f in (R.ref_matrix_cols, R.ref_matrix_rows, R.ref_matrix_values):
            f.argtypes = [C.c_void_p]
        R.ref_create_vector.restype = C.c_void_p
        R.ref_create_vector.argtypes = [C.c_void_p, C.c_int]
        R.ref_destroy_vector.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_map_vector.restype = C.POINTER(C.c_double)
        R.ref_map_vector.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_unmap_vector.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        R.ref_copy_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_dot.restype = C.c_double
        R.ref_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_calc_xr.restype = C.c_double
        R.ref_calc_xr.argtypes = [C.c_void_p] * 5 + [C.c_double]
        R.ref_calc_p.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        R.ref_spmv.argtypes = [C.c_void_p] * 4
        R.ref_inject_bitflip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        R.ref_cg_iterations.restype = C.c_int
        R.ref_cg_iterations.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_double, C.POINTER(C.c_double), C.c_void_p]
        _ref = R
    return _ref


class capture_stdout:
    """Capture what C code printf()s to fd 1 (the reference reports faults that way)."""

    def __enter__(self):
        sys.stdout.flush()
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        self._saved = os.dup(1)
        os.dup2(self._tmp.fileno(), 1)
        self.text = ""
        return self

    def __exit__(self, *exc):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode()
        self._tmp.close()
        return False


class RefContext:
    """The reference's CPUContext for one mode, driven through CGContext's API."""

    def __init__(self, mode: str):
        self.R = ref()
        self.mode = mode
        self.ctx = self.R.ref_context_create(mode.encode())
        self.N = 0

    def create_matrix(self, cols, rows, vals, N):
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.N, self.nnz = int(N), int(cols.shape[0])
        self.A = self.R.ref_create_matrix(self.ctx, cols, rows, vals, self.N, self.nnz)
        return self.A

    def matrix_arrays(self):
        c = np.ctypeslib.as_array(self.R.ref_matrix_cols(self.A), shape=(self.nnz,))
        r = np.ctypeslib.as_array(self.R.ref_matrix_rows(self.A), shape=(self.N + 1,))
        v = np.ctypeslib.as_array(self.R.ref_matrix_values(self.A), shape=(self.nnz,))
        return c, r, v

    def vector(self, data=None):
        v = self.R.ref_create_vector(self.ctx, self.N)
        h = self.view(v)
        h[:] = 0.0 if data is None else data
        return v

    def view(self, v):
        return np.ctypeslib.as_array(self.R.ref_map_vector(self.ctx, v), shape=(self.N,))

    def spmv(self, vec, out):
        """-> text the reference printed"""
        with capture_stdout() as cap:
            self.R.ref_spmv(self.ctx, self.A, vec, out)
        return cap.text

    def dot(self, a, b):
        return self.R.ref_dot(self.ctx, a, b)

    def calc_xr(self, x, r, p, w, alpha):
        return self.R.ref_calc_xr(self.ctx, x, r, p, w, alpha)

    def calc_p(self, p, r, beta):
        self.R.ref_calc_p(self.ctx, p, r, beta)

    def cg(self, b, max_itrs, conv):
        """cg.cpp:83-118 through the plugin API -> dict like OracleMatrix.cg (non-fatal cases only)."""
        bv = self.vector(b); x = self.vector(); r = self.vector(); p = self.vector(); w = self.vector()
        self.R.ref_copy_vector(self.ctx, r, bv)
        self.R.ref_copy_vector(self.ctx, p, r)
        rr = C.c_double(self.R.ref_dot(self.ctx, r, r))
        hist = np.zeros(max(max_itrs, 1))
        with capture_stdout() as cap:
            it = self.R.ref_cg_iterations(self.ctx, self.A, x, r, p, w, max_itrs, conv, C.byref(rr), hist.ctypes.data)
        out = dict(iterations=it, rr_hist=hist[:it].copy(), x=self.view(x).copy(), r=self.view(r).copy(),
                   p=self.view(p).copy(), text=cap.text)
        for v in (bv, x, r, p, w):
            self.R.ref_destroy_vector(self.ctx, v)
        return out


def run_cg_csr(exe: str, args, fake_time=None, env_extra=None, timeout=600):
    """Run a cg-csr binary (the reference CLI) -> (exit code, stdout)."""
    env = dict(os.environ)
    if fake_time is not None:
        env["ABFT_FAKE_TIME"] = str(fake_time)
        env["LD_PRELOAD"] = FAKETIME_SO
    if env_extra:
        env.update(env_extra)
    p = subprocess.run([exe] + [str(a) for a in args], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=timeout)
    return p.returncode, p.stdout.decode(errors="replace")
