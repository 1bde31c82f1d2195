"""Synthetic matrices: shapes BASELINE.json names, in the order cg.cpp hands them over."""
import os

import numpy as np
import pytest


def test_poisson2d_counts(pkg):
    t = pkg.matgen.poisson2d(256)
    assert (t.N, t.nnz) == (65536, 326656)          # C1/C2
    key = t.rows.astype(np.int64) * t.N + t.cols
    assert np.all(np.diff(key) > 0)                  # sorted by (row, col), no duplicates (cg.cpp:394)
    assert t.vals[t.rows == t.cols].min() == 4.0 and set(np.unique(t.vals)) == {-1.0, 4.0}


def test_poisson3d_counts(pkg):
    n = 16
    t = pkg.matgen.poisson3d(n)
    assert t.N == n ** 3 and t.nnz == 7 * n ** 3 - 6 * n * n
    # symmetric
    a = set(zip(t.rows.tolist(), t.cols.tolist()))
    assert all((c, r) in a for r, c in list(a)[:2000])


def test_blocks_replicate_on_the_diagonal(pkg):
    one, four = pkg.matgen.poisson2d(8), pkg.matgen.poisson2d(8, 4)
    assert four.N == 4 * one.N and four.nnz == 4 * one.nnz
    for j in range(4):
        s = slice(j * one.nnz, (j + 1) * one.nnz)
        assert np.array_equal(four.cols[s], one.cols + j * one.N) and np.array_equal(four.rows[s], one.rows + j * one.N)


def test_powerlaw_is_sorted_symmetric_dominant(pkg):
    t = pkg.matgen.powerlaw(2000, 40000, seed=1)
    key = t.rows.astype(np.int64) * t.N + t.cols
    assert np.all(np.diff(key) > 0)
    diag = t.vals[t.rows == t.cols]
    offsum = np.bincount(t.rows[t.rows != t.cols], minlength=t.N)
    assert np.all(diag == offsum + 1.0)
    lens = np.bincount(t.rows, minlength=t.N)
    assert lens.max() > 8 * np.median(lens)          # heavy tail


def test_mtx_roundtrip_through_reference_loader(pkg, O, tmp_path):
    """write_mtx -> the unmodified cg-csr loader reproduces N, nnz and converges like the in-memory triplets."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    t = pkg.matgen.poisson2d(20)
    path = str(tmp_path / "p20.mtx")
    pkg.matgen.write_mtx(path, t)
    code, text = O.run_cg_csr(O.REF_EXE, ["-f", path, "-b", "3", "-c", "1e-10", "-t", "cpu", "-m", "secded"])
    assert code == 0
    t3 = pkg.matgen.poisson2d(20, 3)
    assert f"matrix size           = {t3.N} x {t3.N}" in text
    assert f"number of non-zeros   = {t3.nnz} " in text
    res = O.OracleMatrix("secded", t3.cols, t3.rows, t3.vals, t3.N).cg(pkg.matgen.rhs(t3.N), 1000, 1e-10)
    assert f"ran for {res['iterations']} iterations" in text


@pytest.mark.parametrize("N,nnz", [(1, 1), (2, 4), (50, 300), (2000, 40000), (40000, 600000)])
def test_powerlaw_mirror_equals_a_plain_sort(pkg, N, nnz):
    """The sort-free assembly (row = [L row | diagonal | L^T row], counting-sort transpose; the larger case takes the
    scipy path) gives exactly what sorting the mirrored triplets gives."""
    t = pkg.matgen.powerlaw(N, nnz, seed=4)
    low = t.cols < t.rows
    rl, cl = t.rows[low].astype(np.int64), t.cols[low].astype(np.int64)
    deg = np.bincount(rl, minlength=N) + np.bincount(cl, minlength=N)
    r = np.concatenate([rl, cl, np.arange(N)])
    c = np.concatenate([cl, rl, np.arange(N)])
    v = np.concatenate([np.full(2 * rl.shape[0], -1.0), deg + 1.0])
    order = np.argsort(r * N + c, kind="stable")
    assert np.array_equal(t.rows, r[order]) and np.array_equal(t.cols, c[order]) and np.array_equal(t.vals, v[order])
    assert t.nnz == 2 * rl.shape[0] + N
