"""Full BASELINE.json sizes (too big for the CPU oracle in seconds): size-independent properties.

  * 3D 7-point Poisson 256^3 (N = 2^24, nnz = 117 047 296), secded          -- configs[2]
  * 2D 5-point Poisson 512^2 replicated -b 64 (N = 2^24, nnz = 83 755 008), sec8 -- configs[3] as re-dimensioned
    in SURVEY.md 8(d) (the 24-bit column of the ECC modes caps N at 2^24)

Properties: A*1 equals the analytic row sums bit for bit (and the fused dot is their exact integer sum); a checksum of
all encoded column words equals an independent numpy re-encode from the reference's mask constants (golden file);
encode -> flip 100 random bits -> spmv reports exactly those 100 faults at the right indices, repairs the matrix back to
the clean checksum and returns the clean result; every ABFT mode gives bit-identical CG iterates on a clean matrix.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def poisson3d_256(pkg):
    import bench
    n = 256
    return bench.slab_triplets(pkg.matgen, n, 0, n ** 3), n


def numpy_encode(golden_ecc, mode, cols, vals):
    """Vectorised restatement of generate_ecc_bits from the reference's 21 mask constants (golden file)."""
    bits = vals.view(np.uint64)
    w0 = (bits & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    w1 = (bits >> np.uint64(32)).astype(np.uint32)
    w2 = cols.astype(np.uint32).copy()

    def parity(x):
        x = x ^ (x >> np.uint32(16)); x = x ^ (x >> np.uint32(8)); x = x ^ (x >> np.uint32(4))
        x = x ^ (x >> np.uint32(2)); x = x ^ (x >> np.uint32(1))
        return x & np.uint32(1)

    if mode in ("sec7", "sec8", "secded"):
        syn = np.zeros_like(w2)
        for p in range(1, 8):
            m = [np.uint32(v) for v in golden_ecc["masks"][p - 1]]
            syn |= parity((w0 & m[0]) ^ (w1 & m[1]) ^ (w2 & m[2])) << np.uint32(32 - p)
        w2 |= syn
        if mode != "sec7":
            w2 |= parity(w0 ^ w1 ^ w2) << np.uint32(24)
    elif mode == "sed":
        w2 |= parity(w0 ^ w1 ^ w2) << np.uint32(31)
    return w2


def checksum(a):
    a = np.ascontiguousarray(a)
    return int(np.bitwise_xor.reduce(a.view(np.uint32))), int(a.view(np.uint32).astype(np.uint64).sum())


def test_poisson3d_256_secded_properties(pkg, golden_ecc):
    t, n = poisson3d_256(pkg)
    N = t.N
    assert (N, t.nnz) == (1 << 24, 117047296)
    ctx = pkg.CUDAContext("secded")
    A = ctx.create_matrix(t.cols, t.rows, t.vals, N)
    # (1) checksum of checksums: every encoded word, against an independent re-encode
    cols, rowptr, vals = ctx.read_matrix(A)
    want = numpy_encode(golden_ecc, "secded", t.cols, t.vals)
    assert checksum(cols) == checksum(want)
    assert np.array_equal(cols[::9973], want[::9973]) and np.array_equal(vals, t.vals)
    assert np.array_equal(rowptr, t.row_pointers())
    clean = (checksum(cols), checksum(vals))
    # (2) A * 1 == analytic row sums, bit for bit; the fused dot is their exact integer sum
    idx = np.arange(N, dtype=np.int64)
    z, y, x = idx // (n * n), (idx // n) % n, idx % n
    nbrs = (z > 0).astype(int) + (z < n - 1) + (y > 0) + (y < n - 1) + (x > 0) + (x < n - 1)
    rowsum = 6.0 - nbrs
    v, w = ctx.create_vector(N), ctx.create_vector(N)
    ctx.upload(v, np.ones(N))
    ctx.spmv(A, v, w)
    assert ctx.dot(v, w) == float(rowsum.sum())
    assert ctx.poll_events() == ([], False)
    w_clean = ctx.download(w)
    assert np.array_equal(w_clean, rowsum)
    # (3) encode -> corrupt -> spmv detects, repairs and still returns the clean result
    rng = np.random.default_rng(2026)
    where = np.sort(rng.choice(t.nnz, 100, replace=False))
    bits = rng.integers(0, 96, 100)
    for i, b in zip(where, bits):
        ctx.flip_bit(A, int(i), int(b))
    cols2, _, vals2 = ctx.read_matrix(A)
    assert (checksum(cols2), checksum(vals2)) != clean
    ctx.spmv(A, v, w)
    events, fatal = ctx.poll_events()
    assert not fatal
    assert events == [((3, int(i), 0) if b == 88 else (2, int(i), int(b))) for i, b in zip(where, bits)]
    assert np.array_equal(ctx.download(w), w_clean)
    cols3, _, vals3 = ctx.read_matrix(A)
    assert (checksum(cols3), checksum(vals3)) == clean
    ctx.spmv(A, v, w)
    assert ctx.poll_events() == ([], False)          # idempotent: nothing left to correct
    # (4) a double flip is detected, never silently accepted
    ctx.flip_bit(A, 77777777, 5); ctx.flip_bit(A, 77777777, 70)
    ctx.spmv(A, v, w)
    events, fatal = ctx.poll_events()
    assert fatal and events == [(4, 77777777, 0)]
    ctx.close()


def test_poisson3d_256_all_modes_give_identical_iterates(pkg):
    """On a clean matrix the checks never change the arithmetic: bit-identical CG iterates in every reference mode,
    and run to run.  (The fused dot's tree has one partial per CTA: this holds because these modes' spmv kernels run
    the same persistent grid, 4 CTAs per SM.  crc32c keeps its tables in registers and runs 3 CTAs per SM: the same
    result[] bit for bit, another tree under p.w, so its residual history agrees to rounding, not to the bit.)"""
    t, n = poisson3d_256(pkg)
    N = t.N
    b = np.random.default_rng(1).random(N)
    ref = ref_hist = None
    for mode in ["none", "secded", "secded", "sec7", "crc32c", "constraints"]:
        ctx = pkg.CUDAContext(mode)
        A = ctx.create_matrix(t.cols, t.rows, t.vals, N)
        vb, x, r, p, w = (ctx.create_vector(N) for _ in range(5))
        ctx.upload(vb, b)
        it, hist = ctx.cg(A, vb, x, r, p, w, 25, 0.0)
        got = (hist.tobytes(), checksum(ctx.download(x)))
        assert it == 25 and ctx.messages == []
        if ref is None:
            ref, ref_hist = got, hist.copy()
        if mode == "crc32c":
            assert np.allclose(hist, ref_hist, rtol=1e-12, atol=0.0), mode
        else:
            assert got == ref, mode
        ctx.close()


def test_poisson2d_512_b64_sec8_properties(pkg, golden_ecc):
    t = pkg.matgen.poisson2d(512, 64)
    N = t.N
    assert (N, t.nnz) == (1 << 24, 83755008)
    ctx = pkg.CUDAContext("sec8")
    A = ctx.create_matrix(t.cols, t.rows, t.vals, N)
    cols, rowptr, vals = ctx.read_matrix(A)
    assert checksum(cols) == checksum(numpy_encode(golden_ecc, "sec8", t.cols, t.vals))
    clean = (checksum(cols), checksum(vals))
    v, w = ctx.create_vector(N), ctx.create_vector(N)
    ctx.upload(v, np.ones(N))
    ctx.spmv(A, v, w)
    rowsum = 4.0 - (np.bincount(t.rows, minlength=N) - 1)
    assert np.array_equal(ctx.download(w), rowsum) and ctx.dot(v, w) == float(rowsum.sum())
    rng = np.random.default_rng(7)
    where = np.sort(rng.choice(t.nnz, 50, replace=False))
    bits = rng.integers(0, 96, 50)
    for i, b in zip(where, bits):
        ctx.flip_bit(A, int(i), int(b))
    ctx.spmv(A, v, w)
    events, fatal = ctx.poll_events()
    assert not fatal and [e[1] for e in events] == [int(i) for i in where]
    assert np.array_equal(ctx.download(w), rowsum)
    cols3, _, vals3 = ctx.read_matrix(A)
    assert (checksum(cols3), checksum(vals3)) == clean
    ctx.close()
