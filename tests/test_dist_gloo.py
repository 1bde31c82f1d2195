"""Multi-GPU host logic on CPU: world_size 2 and 3 over gloo (no GPU).  The partition, the needed-interval
computation and the exchange plan are the real ones (abft-sparse-cg_b200/dist.py); only the kernels are
numpy stand-ins.  Entries of p the plan does not deliver are NaN, so an incomplete plan fails."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_worker(world, mode, shape, tmp_path, extra=()):
    out = str(tmp_path / f"out_{mode}_{shape}_{world}.json")
    port = 29600 + (os.getpid() + world * 7 + len(shape)) % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_dist_worker.py"),
           mode, shape, out] + list(extra)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert p.returncode == 0, p.stdout.decode()[-3000:]
    with open(out) as f:
        return json.load(f)


def test_cuts_balance_nnz(pkg):
    t = pkg.matgen.powerlaw(4000, 100000, seed=5)
    rowptr = t.row_pointers().astype(np.int64)
    for world in (2, 3, 8):
        cuts = pkg.dist.balanced_row_cuts(rowptr, world)
        assert cuts[0] == 0 and cuts[-1] == t.N and all(a <= b for a, b in zip(cuts, cuts[1:]))
        per = [int(rowptr[b] - rowptr[a]) for a, b in zip(cuts, cuts[1:])]
        assert max(per) - min(per) <= 2 * int(np.diff(rowptr).max())      # within one (longest) row of even


def test_plan_is_a_halo_for_banded_and_empty_for_blocks(pkg):
    D, mg = pkg.dist, pkg.matgen
    # 3D Poisson cut in z-slabs: each neighbour needs exactly one plane
    n = 8
    t = mg.poisson3d(n)
    cuts = D.even_row_cuts(t.N, 4)
    rowptr = t.row_pointers().astype(np.int64)
    need = []
    for r in range(4):
        e0, e1 = int(rowptr[cuts[r]]), int(rowptr[cuts[r + 1]])
        need.append(D.needed_intervals(t.cols[e0:e1], cuts))
    for r in range(4):
        plan = D.exchange_plan(need, r)
        for s in range(4):
            width = plan["push_hi"][s] - plan["push_lo"][s]
            assert width == (n * n if abs(s - r) == 1 else 0)
        assert plan["wait_mask"] == sum(1 << s for s in range(4) if abs(s - r) == 1)
    # block-diagonal matrix cut on block boundaries: nothing to exchange
    t = mg.poisson2d(6, 4)
    cuts = D.even_row_cuts(t.N, 4)
    rowptr = t.row_pointers().astype(np.int64)
    need = [D.needed_intervals(t.cols[int(rowptr[cuts[r]]):int(rowptr[cuts[r + 1]])], cuts) for r in range(4)]
    for r in range(4):
        plan = D.exchange_plan(need, r)
        assert D.exchanged_bytes(plan) == 0 and plan["wait_mask"] == 0


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("shape", ["p2d", "powerlaw", "blocks"])
def test_partitioned_cg_matches_single_rank(pkg, O, tmp_path, world, shape):
    import importlib.util
    spec = importlib.util.spec_from_file_location("_dist_worker", os.path.join(ROOT, "tests", "_dist_worker.py"))
    wk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wk)
    t = wk.build(pkg.matgen, shape)
    ref = O.OracleMatrix("none", t.cols, t.rows, t.vals, t.N).cg(pkg.matgen.rhs(t.N), 2000, 1e-10)
    res = run_worker(world, "cpu", shape, tmp_path, extra=[str(ref["iterations"])])
    assert all(r["iterations"] == ref["iterations"] for r in res["all"])
    assert all(r["rr"] == res["rr"] for r in res["all"])                       # every rank holds the same scalar
    x = np.asarray(res["x"])
    if shape == "powerlaw":
        # widely spread spectrum: in finite precision the CG residual history is sensitive to the order of the
        # reductions (numpy's pairwise sums vs the sequential reference) long before convergence, so only the
        # solution is comparable
        assert np.max(np.abs(x - ref["x"])) <= 1e-6 * np.max(np.abs(ref["x"]))
    else:
        assert abs(res["rr"] - ref["rr_hist"][-1]) <= 1e-6 * ref["rr_hist"][-1]
        assert np.max(np.abs(x - ref["x"])) <= 1e-9 * np.max(np.abs(ref["x"]))
