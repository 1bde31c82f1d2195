"""Synthetic SPD matrices of the shapes BASELINE.json names (no network: the
reference's default input is a SuiteSparse download, matrices/Makefile:7-11).

Every generator returns the triplets exactly as the reference's loader hands
them to ``create_matrix`` (cg.cpp:394-418): 0-based, both triangles present,
sorted by (row, col), block-replicated ``blocks`` times on the diagonal.
``write_mtx`` emits the Matrix-Market file the unmodified ``cg-csr`` binary
reads (lower triangle + diagonal, 1-based; cg.cpp:365-392 mirrors it).
"""
from __future__ import annotations

import numpy as np


class Triplets:
    """Row-sorted COO triplets + shape, the argument list of create_matrix."""

    __slots__ = ("cols", "rows", "vals", "N", "nnz")

    def __init__(self, cols, rows, vals, N):
        self.cols = np.ascontiguousarray(cols, dtype=np.uint32)
        self.rows = np.ascontiguousarray(rows, dtype=np.uint32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.N = int(N)
        self.nnz = int(self.cols.shape[0])

    def row_pointers(self) -> np.ndarray:
        """CSR row pointers as CSR/CPUContext.cpp:36-41 builds them."""
        counts = np.bincount(self.rows, minlength=self.N).astype(np.int64)
        ptr = np.zeros(self.N + 1, dtype=np.int64)
        np.cumsum(counts, out=ptr[1:])
        return ptr.astype(np.uint32)


def _stencil(shape, diag):
    """(2*dim+1)-point Poisson stencil on a grid of the given shape."""
    dims = len(shape)
    N = int(np.prod(shape))
    idx = np.arange(N, dtype=np.int64)
    strides = [int(np.prod(shape[d + 1:])) for d in range(dims)]
    coords = [(idx // strides[d]) % shape[d] for d in range(dims)]
    # neighbour offsets in increasing column order: -s0, -s1, ..., 0, ..., +s1, +s0
    offs, valid, vals = [], [], []
    for d in range(dims):
        offs.append(-strides[d]); valid.append(coords[d] > 0); vals.append(-1.0)
    offs.append(0); valid.append(np.ones(N, dtype=bool)); vals.append(float(diag))
    for d in reversed(range(dims)):
        offs.append(strides[d]); valid.append(coords[d] < shape[d] - 1); vals.append(-1.0)
    k = len(offs)
    mask = np.stack(valid, axis=1)                       # (N, k) row-major => (row, col) order
    cols = (idx[:, None] + np.asarray(offs, dtype=np.int64)[None, :])[mask]
    rows = np.broadcast_to(idx[:, None], (N, k))[mask]
    v = np.broadcast_to(np.asarray(vals)[None, :], (N, k))[mask]
    return cols, rows, v, N


def _replicate(cols, rows, vals, width, blocks):
    """cg.cpp:401-416: copy the block `blocks` times down the diagonal."""
    if blocks == 1:
        return cols, rows, vals, width
    shift = (np.arange(blocks, dtype=np.int64) * width)[:, None]
    return ((cols[None, :] + shift).ravel(), (rows[None, :] + shift).ravel(),
            np.tile(vals, blocks), width * blocks)


def poisson2d(n: int, blocks: int = 1) -> Triplets:
    """2D 5-point Poisson on an n x n grid (diag 4, off-diag -1)."""
    c, r, v, N = _stencil((n, n), 4.0)
    c, r, v, N = _replicate(c, r, v, N, blocks)
    return Triplets(c, r, v, N)


def poisson3d(n: int, blocks: int = 1) -> Triplets:
    """3D 7-point Poisson on an n^3 grid (diag 6, off-diag -1)."""
    c, r, v, N = _stencil((n, n, n), 6.0)
    c, r, v, N = _replicate(c, r, v, N, blocks)
    return Triplets(c, r, v, N)


def powerlaw(N: int, target_nnz: int, alpha: float = 2.0, max_deg: int = 4096, seed: int = 1) -> Triplets:
    """Irregular SPD matrix with power-law row lengths (BASELINE config 5).

    Lower-triangle degree d_i ~ Zipf(alpha) clipped to [1, min(i, max_deg)],
    scaled so that the mirrored matrix has about ``target_nnz`` entries;
    distinct random columns j < i, off-diagonal -1, diagonal = row degree + 1
    (strictly diagonally dominant, hence SPD).
    """
    rng = np.random.default_rng(seed)
    deg = rng.zipf(alpha, size=N).astype(np.int64)
    want_lower = max((target_nnz - N) // 2, 1)
    scale = want_lower / max(int(np.minimum(deg, max_deg).sum()), 1)
    if scale > 1.0:
        deg = np.ceil(deg * scale).astype(np.int64)
    deg = np.minimum(np.minimum(deg, max_deg), np.arange(N, dtype=np.int64))
    rows_l = np.repeat(np.arange(N, dtype=np.int64), deg)
    cols_l = (rng.random(rows_l.shape[0]) * rows_l).astype(np.int64)          # j in [0, i)
    key = rows_l * N + cols_l
    key.sort()                                                                # (np.unique is 70x slower here)
    if key.shape[0]:
        keep = np.empty(key.shape, dtype=bool)
        keep[0] = True
        np.not_equal(key[1:], key[:-1], out=keep[1:])
        key = key[keep]                                                       # distinct, sorted
    rows_l, cols_l = key // N, key % N
    del key
    return _mirror_lower(rows_l, cols_l, N)


def _mirror_lower(rows_l, cols_l, N) -> Triplets:
    """Strict lower triangle (sorted by (row, col)) -> the full symmetric matrix, sorted by (row, col), without a
    sort of the mirrored entries: row i of the result is [row i of L | diagonal | row i of L^T], and the transpose
    is a counting sort (a 2e8-entry argsort takes minutes; this takes seconds)."""
    n_l = rows_l.shape[0]
    cnt_l = np.bincount(rows_l, minlength=N)            # entries left of the diagonal
    cnt_u = np.bincount(cols_l, minlength=N)            # entries right of it
    # transpose of L by a stable counting sort on the column: within a column the rows stay ascending
    t_order = np.argsort(cols_l, kind="stable") if n_l < (1 << 16) else _stable_counting_order(cols_l, cnt_u, N)
    ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(cnt_l + cnt_u + 1, out=ptr[1:])
    start_l = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(cnt_l, out=start_l[1:])
    start_u = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(cnt_u, out=start_u[1:])
    nnz = int(ptr[-1])
    c = np.empty(nnz, dtype=np.uint32)
    r = np.empty(nnz, dtype=np.uint32)
    v = np.empty(nnz, dtype=np.float64)
    r[:] = np.repeat(np.arange(N, dtype=np.uint32), cnt_l + cnt_u + 1)
    v[:] = -1.0
    # lower part of row i sits at ptr[i] + (k - start_l[i]) for the row's k-th entry of L
    pos = np.arange(n_l, dtype=np.int64) + np.repeat(ptr[:-1] - start_l[:-1], cnt_l)
    c[pos] = cols_l
    # diagonal
    dpos = ptr[:-1] + cnt_l
    c[dpos] = np.arange(N, dtype=np.uint32)
    v[dpos] = (cnt_l + cnt_u).astype(np.float64) + 1.0
    # upper part: row i = column i of L, entries (their rows) ascending
    pos = np.arange(n_l, dtype=np.int64) + np.repeat(dpos + 1 - start_u[:-1], cnt_u)
    c[pos] = rows_l[t_order]
    return Triplets(c, r, v, N)


def _stable_counting_order(keys, counts, N):
    """argsort(keys, kind='stable') for integer keys in [0, N) with known counts, through scipy's O(nnz) CSR->CSC
    conversion (a counting sort in C)."""
    import scipy.sparse as sp
    n = keys.shape[0]
    # one entry per element: row = its position's "chunk" does not matter; use a 1 x n trick: matrix with
    # rows = key, cols = position is already "sorted by position within a key" after tocsr()
    m = sp.coo_matrix((np.ones(n, dtype=np.int8), (keys, np.arange(n, dtype=np.int64))), shape=(N, n)).tocsr()
    return m.indices.astype(np.int64)


def write_mtx(path: str, t: Triplets) -> None:
    """Matrix-Market `coordinate real symmetric`, lower triangle + diagonal."""
    keep = t.cols <= t.rows
    r = t.rows[keep].astype(np.int64) + 1
    c = t.cols[keep].astype(np.int64) + 1
    v = t.vals[keep]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real symmetric\n")
        f.write(f"{t.N} {t.N} {r.shape[0]}\n")
        np.savetxt(f, np.column_stack([r, c, v]), fmt=["%d", "%d", "%.17g"])


def rhs(N: int) -> np.ndarray:
    """The right-hand side cg.cpp:68-72 builds: glibc rand() from its default
    seed (srand is only called later, cg.cpp:79), divided by RAND_MAX."""
    import ctypes
    libc = ctypes.CDLL(None)
    libc.rand.restype = ctypes.c_int
    libc.srand(1)                      # glibc's default seed
    rand = libc.rand
    out = np.empty(N, dtype=np.float64)
    for i in range(N):
        out[i] = rand()
    return out / 2147483647.0
