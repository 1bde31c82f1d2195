"""Host logic of create_matrix: the spmv work decomposition (abft_cuda_plan_tiles, no device needed)."""
import numpy as np
import pytest

REGULAR, LONG = 1 << 16, 1 << 17


def plan(pkg, lens, tile_nnz=0):
    L = pkg.capi.load()
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    n = L.abft_cuda_plan_tiles(rowptr.ctypes.data, len(lens), tile_nnz, None, 0)
    out = np.zeros((max(n, 1), 4), dtype=np.uint32)
    assert L.abft_cuda_plan_tiles(rowptr.ctypes.data, len(lens), tile_nnz, out.ctypes.data, n) == n
    return rowptr, out[:n]


def check_partition(rowptr, tiles, n_rows, tile_nnz):
    row = 0
    for rb, nb, ne, info in tiles:
        nrows = int(info) & 0xFFFF
        assert rb == row and nb == rowptr[rb] and ne == rowptr[rb + nrows]     # whole rows, contiguous, in order
        assert 1 <= nrows <= 256
        if info & LONG:
            assert nrows == 1 and ne - nb > tile_nnz and not (info & REGULAR)
        else:
            assert ne - nb <= tile_nnz
            lens = np.diff(rowptr[rb:rb + nrows + 1].astype(np.int64))
            assert ((int(info) >> 18) & 0xFFF) == min(int(lens.max()), 0xFFF)
        row += nrows
    assert row == n_rows


def test_stencil_matrices_are_all_regular_2048_tiles(pkg):
    for t in [pkg.matgen.poisson2d(64), pkg.matgen.poisson3d(20), pkg.matgen.poisson2d(16, 5)]:
        lens = np.bincount(t.rows, minlength=t.N)
        rowptr, tiles = plan(pkg, lens)
        check_partition(rowptr, tiles, t.N, 2048)
        assert all(info & REGULAR for *_, info in tiles)
        # full tiles hold 256 rows (row limit) of 5..7 elements
        assert all((int(info) & 0xFFFF) == 256 for *_, info in tiles[:-1])


def test_power_law_matrices_get_the_smaller_tiles(pkg):
    t = pkg.matgen.powerlaw(50_000, 1_000_000, max_deg=3000, seed=6)
    lens = np.bincount(t.rows, minlength=t.N)
    rowptr, tiles = plan(pkg, lens)
    check_partition(rowptr, tiles, t.N, 1536)
    assert any(not (info & REGULAR) for *_, info in tiles) and any(info & LONG for *_, info in tiles) == bool((lens > 1536).any())
    # explicit sizes are honoured
    for size in (512, 2048):
        rowptr, tiles = plan(pkg, lens, size)
        check_partition(rowptr, tiles, t.N, size)


def test_a_handful_of_very_long_rows_is_not_a_thread_per_row_tile(pkg):
    # three rows of 680 elements: similar lengths, but one thread per row would walk 680 elements each
    rowptr, tiles = plan(pkg, [680, 680, 680, 3, 3, 3], 2048)
    assert not (tiles[0][3] & REGULAR)
    # rows of similar, short length are regular; so are tiny matrices
    for lens in ([7] * 300, [1], [5, 5], [0, 0, 4, 0], [40] * 25):
        rowptr, tiles = plan(pkg, lens, 2048)
        check_partition(rowptr, tiles, len(lens), 2048)
        assert all(info & REGULAR for *_, info in tiles), lens
    # one long row among short ones: irregular (element-parallel), not regular
    rowptr, tiles = plan(pkg, [3] * 100 + [900] + [3] * 100, 2048)
    assert not (tiles[0][3] & REGULAR)


def test_empty_and_ragged(pkg):
    rowptr, tiles = plan(pkg, [])
    assert len(tiles) == 0
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 10, 5000)
    lens[[0, 17, 4999]] = 0
    lens[100] = 5000
    rowptr, tiles = plan(pkg, lens, 2048)
    check_partition(rowptr, tiles, 5000, 2048)
    assert sum(1 for *_, info in tiles if info & LONG) == 1
