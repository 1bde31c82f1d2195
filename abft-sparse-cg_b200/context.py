"""Host-side mirror of the reference's plugin interface for this path.

`CUDAContext` has the method names, argument order and error behaviour of
CGContext (CGContext.h:15-36) so the parity tests read like a run of cg.cpp;
every method is one call into the C ABI (capi).  The one difference forced by
Python: where the reference printf()s a fault message and exit(1)s from inside
spmv (CSR/CPUContext.cpp:233-234 etc.), `report_faults()` returns the message
lines and raises `FatalFault` -- the C++ plugin (host/CUDAContext.cpp) does the
real print + exit(1).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

MODES = capi.MODES
ANY, VALUE, INDEX = 0, 1, 2  # CGContext::BitFlipKind (CGContext.h:11)


class FatalFault(RuntimeError):
    """A fault the reference would have exit(1)ed on; .lines holds the messages printed so far."""

    def __init__(self, lines, events):
        super().__init__(lines[-1] if lines else "fatal fault")
        self.lines = lines
        self.events = events


class Matrix:
    def __init__(self, handle, N, nnz, nnz_offset=0):
        self.handle, self.N, self.nnz, self.nnz_offset = handle, N, nnz, nnz_offset


class Vector:
    def __init__(self, handle, N):
        self.handle, self.N = handle, N


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class CUDAContext:
    def __init__(self, mode: str = "none", device: int = 0):
        self.L = capi.load()
        mid = self.L.abft_cuda_mode_from_name(mode.encode())
        if mid < 0:
            # CGContext::create's behaviour for an unknown (target, mode): CGContext.cpp:18-23
            raise ValueError(f"No implementation found for cuda-{mode}")
        self.mode = mode
        self.ctx = self.L.abft_cuda_ctx_create(mid, device)
        if not self.ctx:
            raise RuntimeError("cuda target: context creation failed: " + capi.last_error())
        self.messages = []   # every fault line reported so far (what the reference prints)
        self.events = []

    def _ck(self, rc, what):
        if rc:
            raise RuntimeError(f"cuda target: {what} failed: {capi.last_error()}")

    # ---- matrix ----
    def create_matrix(self, columns, rows, values, N, nnz=None):
        columns = np.ascontiguousarray(columns, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        nnz = int(columns.shape[0]) if nnz is None else int(nnz)
        h = self.L.abft_cuda_matrix_create(self.ctx, _ptr(columns), _ptr(rows), _ptr(values), int(N), nnz)
        if not h:
            raise RuntimeError("cuda target: create_matrix failed: " + capi.last_error())
        return Matrix(h, int(N), nnz)

    def create_matrix_part(self, columns, rows, values, n_rows, N_global, row_offset, nnz_offset):
        columns = np.ascontiguousarray(columns, dtype=np.uint32)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        nnz = int(columns.shape[0])
        h = self.L.abft_cuda_matrix_create_part(self.ctx, _ptr(columns), _ptr(rows), _ptr(values), int(n_rows), nnz,
                                                int(N_global), int(row_offset), int(nnz_offset))
        if not h:
            raise RuntimeError("cuda target: create_matrix_part failed: " + capi.last_error())
        return Matrix(h, int(n_rows), nnz, int(nnz_offset))

    def destroy_matrix(self, mat):
        self.L.abft_cuda_matrix_destroy(self.ctx, mat.handle)
        mat.handle = None

    def read_matrix(self, mat):
        cols = np.empty(mat.nnz, dtype=np.uint32)
        rowptr = np.empty(mat.N + 1, dtype=np.uint32)
        vals = np.empty(mat.nnz, dtype=np.float64)
        self._ck(self.L.abft_cuda_matrix_read(self.ctx, mat.handle, _ptr(cols), _ptr(rowptr), _ptr(vals)), "matrix_read")
        return cols, rowptr, vals

    # ---- vectors ----
    def create_vector(self, N):
        h = self.L.abft_cuda_vector_create(self.ctx, int(N))
        if not h:
            raise RuntimeError("cuda target: create_vector failed: " + capi.last_error())
        return Vector(h, int(N))

    def wrap_vector(self, dev_ptr, N):
        h = self.L.abft_cuda_vector_wrap(self.ctx, C.c_void_p(int(dev_ptr)), int(N))
        if not h:
            raise RuntimeError("cuda target: wrap_vector failed: " + capi.last_error())
        return Vector(h, int(N))

    def view_vector(self, parent, offset, N):
        h = self.L.abft_cuda_vector_view(self.ctx, parent.handle, int(offset), int(N))
        if not h:
            raise RuntimeError("cuda target: view_vector failed: " + capi.last_error())
        return Vector(h, int(N))

    def destroy_vector(self, vec):
        self.L.abft_cuda_vector_destroy(self.ctx, vec.handle)
        vec.handle = None

    def map_vector(self, v) -> np.ndarray:
        p = self.L.abft_cuda_vector_map(self.ctx, v.handle)
        if not p:
            raise RuntimeError("cuda target: map_vector failed: " + capi.last_error())
        return np.ctypeslib.as_array(p, shape=(max(v.N, 1),))[:v.N]

    def unmap_vector(self, v, h: np.ndarray):
        self._ck(self.L.abft_cuda_vector_unmap(self.ctx, v.handle, _ptr(h)), "unmap_vector")

    def copy_vector(self, dst, src):
        self._ck(self.L.abft_cuda_vector_copy(self.ctx, dst.handle, src.handle), "copy_vector")

    # convenience on top of map/unmap
    def upload(self, v, data):
        data = np.ascontiguousarray(data, dtype=np.float64)
        assert data.shape[0] == v.N
        self.unmap_vector(v, data)

    def download(self, v) -> np.ndarray:
        return self.map_vector(v).copy()

    def flip_vector_bit(self, v, index, bit):
        """mode secdedv's injection aid: flip one bit of a vector element on the device"""
        self._ck(self.L.abft_cuda_vector_flip_bit(self.ctx, v.handle, int(index), int(bit)), "vector_flip_bit")

    # ---- hot path ----
    def dot(self, a, b) -> float:
        out = C.c_double(0.0)
        self._ck(self.L.abft_cuda_dot(self.ctx, a.handle, b.handle, C.byref(out)), "dot")
        return out.value

    def calc_xr(self, x, r, p, w, alpha) -> float:
        out = C.c_double(0.0)
        self._ck(self.L.abft_cuda_calc_xr(self.ctx, x.handle, r.handle, p.handle, w.handle, float(alpha),
                                          C.byref(out)), "calc_xr")
        return out.value

    def calc_p(self, p, r, beta):
        self._ck(self.L.abft_cuda_calc_p(self.ctx, p.handle, r.handle, float(beta)), "calc_p")

    def spmv(self, mat, vec, result):
        self._ck(self.L.abft_cuda_spmv(self.ctx, mat.handle, vec.handle, result.handle), "spmv")

    def sync(self):
        self._ck(self.L.abft_cuda_ctx_sync(self.ctx), "sync")

    # ---- injection (CSR/CPUContext.cpp:135-159) ----
    def flip_bit(self, mat, index, bit):
        self._ck(self.L.abft_cuda_flip_bit(self.ctx, mat.handle, int(index), int(bit)), "flip_bit")

    def inject_bitflip(self, mat, kind=ANY, num_flips=1, rand=None):
        """Draws rand() in the reference's order; `rand` defaults to libc's rand()."""
        if rand is None:
            libc = C.CDLL(None)
            libc.rand.restype = C.c_int
            rand = libc.rand
        index = rand() % mat.nnz
        start, end = 0, 96
        if kind == VALUE:
            end = 64
        elif kind == INDEX:
            start = 64
        picked = [index]
        for _ in range(num_flips):
            bit = (rand() % (end - start)) + start
            self.messages.append(f"*** flipping bit {bit} at index {index} ***")
            self.flip_bit(mat, index, bit)
            picked.append(bit)
        return picked

    # ---- fault reporting ----
    def poll_events(self, cap=256):
        """Every queued event in the reference's serial order (the library hands them out `cap` at a time), and
        whether the last one is fatal."""
        buf = (capi.Event * cap)()
        fatal = C.c_int(0)
        out = []
        while True:
            n = self.L.abft_cuda_poll_events(self.ctx, buf, cap, C.byref(fatal))
            if n < 0:
                raise RuntimeError("cuda target: poll_events failed: " + capi.last_error())
            out += [(buf[i].kind, buf[i].index, buf[i].bit) for i in range(n)]
            if n == 0 or fatal.value:
                return out, bool(fatal.value)

    def format_event(self, ev) -> str:
        e = capi.Event(*ev)
        buf = C.create_string_buffer(128)
        self.L.abft_cuda_format_event(C.byref(e), buf, 128)
        return buf.value.decode()

    def report_faults(self):
        """What host/CUDAContext.cpp::report_faults does, minus the exit(1)."""
        events, fatal = self.poll_events()
        lines = [self.format_event(e) for e in events]
        self.events += events
        self.messages += lines
        if fatal:
            raise FatalFault(list(self.messages), list(self.events))
        return lines

    # ---- the CG loop of cg.cpp:83-118, driven through the plugin API ----
    def cg(self, A, b, x, r, p, w, max_itrs, conv_threshold):
        """-> (iterations, rr history).  Raises FatalFault where the reference would exit(1)."""
        self.copy_vector(r, b)
        self.copy_vector(p, r)
        rr = self.dot(r, r)
        hist = []
        itr = 0
        while itr < max_itrs and rr > conv_threshold:
            self.spmv(A, p, w)
            pw = self.dot(p, w)
            self.report_faults()
            alpha = rr / pw
            rr_new = self.calc_xr(x, r, p, w, alpha)
            beta = rr_new / rr
            self.calc_p(p, r, beta)
            rr = rr_new
            hist.append(rr)
            itr += 1
        return itr, np.asarray(hist)

    def cg_host_loop(self, A, x, r, p, w, max_itrs, conv_threshold, rr, hist=None):
        """abft_cuda_cg_host_loop: the loop body of cg() above, compiled (same four entry points, same order) ->
        (iterations, rr).  Raises FatalFault where the reference would exit(1); a repaired fault is reported and the
        loop goes on."""
        rr_io = C.c_double(rr)
        done = 0
        while done < max_itrs and rr_io.value > conv_threshold:
            buf = np.zeros(max(max_itrs - done, 1)) if hist is not None else None
            n = self.L.abft_cuda_cg_host_loop(self.ctx, A.handle, x.handle, r.handle, p.handle, w.handle, int(max_itrs - done),
                                              float(conv_threshold), C.byref(rr_io), _ptr(buf) if buf is not None else None)
            if n < 0:
                raise RuntimeError("cuda target: cg_host_loop failed: " + capi.last_error())
            if hist is not None:
                hist.extend(buf[:n].tolist())
            done += n
            if self.L.abft_cuda_events_pending(self.ctx):
                self.report_faults()
            elif n == 0:
                break
        return done, rr_io.value

    # ---- COO format (the reference's cg-coo; COO/CPUContext.cpp) ----
    COO_ELEM = np.dtype([("col", np.uint32), ("row", np.uint32), ("value", np.float64)])

    def create_coo_matrix(self, columns, rows, values, N):
        c = np.ascontiguousarray(columns, dtype=np.uint32)
        r = np.ascontiguousarray(rows, dtype=np.uint32)
        v = np.ascontiguousarray(values, dtype=np.float64)
        h = self.L.abft_cuda_coo_matrix_create(self.ctx, _ptr(c), _ptr(r), _ptr(v), int(N), int(c.shape[0]))
        if not h:
            raise RuntimeError("cuda target: coo_matrix_create failed: " + capi.last_error())
        m = Matrix(h, int(N), int(c.shape[0]))
        m.coo = True
        return m

    def destroy_coo_matrix(self, mat):
        self.L.abft_cuda_coo_matrix_destroy(self.ctx, mat.handle)
        mat.handle = None

    def read_coo_matrix(self, mat) -> np.ndarray:
        out = np.zeros(max(mat.nnz, 1), dtype=self.COO_ELEM)
        self._ck(self.L.abft_cuda_coo_matrix_read(self.ctx, mat.handle, _ptr(out)), "coo_matrix_read")
        return out[:mat.nnz]

    def coo_spmv(self, mat, vec, result):
        self._ck(self.L.abft_cuda_coo_spmv(self.ctx, mat.handle, vec.handle, result.handle), "coo_spmv")

    def coo_flip_bit(self, mat, index, bit):
        self._ck(self.L.abft_cuda_coo_flip_bit(self.ctx, mat.handle, int(index), int(bit)), "coo_flip_bit")

    def coo_format_event(self, ev) -> str:
        e = capi.Event(*ev)
        buf = C.create_string_buffer(128)
        self.L.abft_cuda_coo_format_event(C.byref(e), buf, 128)
        return buf.value.decode()

    def coo_cg(self, A, b, x, r, p, w, max_itrs, conv_threshold):
        """cg.cpp:83-118 over a COO matrix -> (iterations, rr history, events, fatal)."""
        self.copy_vector(r, b)
        self.copy_vector(p, r)
        rr = self.dot(r, r)
        hist, events = [], []
        itr = 0
        while itr < max_itrs and rr > conv_threshold:
            self.coo_spmv(A, p, w)
            pw = self.dot(p, w)
            ev, fatal = self.poll_events()
            events += ev
            if fatal:
                return itr, np.asarray(hist), events, True
            alpha = rr / pw
            rr_new = self.calc_xr(x, r, p, w, alpha)
            beta = rr_new / rr
            self.calc_p(p, r, beta)
            rr = rr_new
            hist.append(rr)
            itr += 1
        return itr, np.asarray(hist), events, False

    def cg_solve_device(self, A, b, x, r, p, w, max_itrs, conv_threshold, check_every=16):
        """Device-resident loop (abft_cuda_cg_solve) -> (iterations, final rr)."""
        rr = C.c_double(0.0)
        it = self.L.abft_cuda_cg_solve(self.ctx, A.handle, b.handle, x.handle, r.handle, p.handle, w.handle,
                                       int(max_itrs), float(conv_threshold), int(check_every), C.byref(rr))
        if it == -2:
            self.report_faults()
        if it < 0:
            raise RuntimeError("cuda target: cg_solve failed: " + capi.last_error())
        return it, rr.value

    def launch_count(self) -> int:
        return int(self.L.abft_cuda_launch_count(self.ctx))

    def close(self):
        if self.ctx:
            self.L.abft_cuda_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
