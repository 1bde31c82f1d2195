"""Row-pointer protection on the device (csrc/rowsecc.cuh; repository-defined, parity unpinned): the C ABI with
ABFT_CUDA_PROTECT_ROWS=secded against oracle/rowptr_ecc.c.  Needs a library built with -DABFT_ROWS_ECC=1
(ABFT_CUDA_LIB=build/variants/libabftcuda_rowsecc.so; `tools/ab_variants.sh build`): skipped with the shipped build,
which refuses the option.  The same cases run on the CPU emulator in tests/test_rowptr_ecc.py."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
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


@pytest.fixture()
def protect(monkeypatch):
    monkeypatch.setenv("ABFT_CUDA_PROTECT_ROWS", "secded")


def make(pkg, mode, t):
    ctx = pkg.CUDAContext(mode)
    try:
        A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    except RuntimeError as e:
        ctx.close()
        if "built without row-pointer protection" in str(e):
            pytest.skip("library built without -DABFT_ROWS_ECC=1")
        raise
    return ctx, A


def spmv(ctx, A, vec):
    v, w = ctx.create_vector(len(vec)), ctx.create_vector(A.N)
    ctx.upload(v, vec)
    ctx.spmv(A, v, w)
    events, fatal = ctx.poll_events()
    return ctx.download(w), fatal, events


@pytest.mark.parametrize("mode", ["none", "secded", "constraints", "crc32c"])
def test_protected_rows_clean_and_single_flips(pkg, O, RO, protect, mode):
    t = pkg.matgen.poisson3d(12)
    vec = np.random.default_rng(1).standard_normal(t.N)
    ctx, A = make(pkg, mode, t)
    om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
    enc = np.zeros((t.N + 2) & ~1, dtype=np.uint32)
    enc[:t.N + 1] = t.row_pointers()
    RO.rpo_encode(enc, len(enc))
    _, rowptr, _ = ctx.read_matrix(A)
    assert np.array_equal(rowptr, enc[:t.N + 1])
    y0, _, _ = om.spmv(vec)
    y, fatal, events = spmv(ctx, A, vec)
    assert not fatal and events == [] and np.array_equal(y.view(np.uint64), y0.view(np.uint64))
    rng = np.random.default_rng(2)
    for trial in range(40):
        row, bit = int(rng.integers(0, t.N + 1)), int(rng.integers(0, 32))
        ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, row, int(enc[row]) ^ (1 << bit)), "poke_rowptr")
        y, fatal, events = spmv(ctx, A, vec)
        assert (fatal, events) == (False, [(10, row // 2, 32 * (row % 2) + bit)]), (row, bit, events)
        assert np.array_equal(y.view(np.uint64), y0.view(np.uint64))
        _, rowptr, _ = ctx.read_matrix(A)
        assert np.array_equal(rowptr, enc[:t.N + 1])            # repaired in memory
    # two flips in one pair: detected, fatal
    ctx._ck(ctx.L.abft_cuda_matrix_poke_rowptr(ctx.ctx, A.handle, 300, int(enc[300]) ^ 0b101), "poke_rowptr")
    _, fatal, events = spmv(ctx, A, vec)
    assert fatal and events == [(11, 150, 0)]
    ctx.close()
