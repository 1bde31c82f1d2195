"""Row-partitioned CG over the dist_* entry points of the C ABI on 2+ GPUs of one box: peer stores over
NVLink for the p exchange, in-kernel all-reduce of the two scalars, flag waits in spmv.  One process per
GPU (torch.distributed.run); compared with the single-rank CPU oracle.  Skipped on a 1-GPU box (ranks that
wait on each other must not share a GPU)."""
import numpy as np
import pytest

from test_dist_gloo import run_worker, ROOT  # noqa: F401

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("shape,mode", [("p2d", "secded"), ("p3d", "none"), ("blocks", "sec7"), ("powerlaw", "constraints"),
                                        ("p3d", "crc32c")])
def test_dist_cg_matches_oracle(pkg, O, tmp_path, shape, mode):
    world = min(n_gpus(), 8)       # every GPU of the box (the exchange windows hold up to 8 ranks)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("_dist_worker", os.path.join(ROOT, "tests", "_dist_worker.py"))
    wk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wk)
    t = wk.build(pkg.matgen, shape)
    ref = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N).cg(pkg.matgen.rhs(t.N), 2000, 1e-10)
    res = run_worker(world, "gpu", shape, tmp_path, extra=[mode])
    x = np.asarray(res["x"])
    assert all(r["rr"] == res["rr"] for r in res["all"])                  # all-reduce gives every rank the same bits
    if shape == "powerlaw":
        assert np.max(np.abs(x - ref["x"])) <= 1e-6 * np.max(np.abs(ref["x"]))
    else:
        assert all(r["iterations"] == ref["iterations"] for r in res["all"])
        hist = np.asarray(res["hist"])
        assert np.max(np.abs(hist - ref["rr_hist"]) / ref["rr_hist"]) <= 1e-10
        assert np.max(np.abs(x - ref["x"])) <= 1e-10 * np.max(np.abs(ref["x"]))
    assert all(lines == [] for lines in res["lines"])
    # device-resident partitioned loop == host-driven partitioned loop, bit for bit, on every rank
    assert all(all(flags) for flags in res["dev_same"]), res["dev_same"]


def test_dist_fault_reports_global_index(pkg, O, tmp_path):
    world = min(n_gpus(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    res = run_worker(world, "gpu", "p2d", tmp_path, extra=["secded", "flip"])
    t = pkg.matgen.poisson2d(24)
    gi = int(t.row_pointers()[res["cuts"][-2]]) + 7      # element 7 of the LAST rank's block, as a global index
    assert res["lines"][-1] == [f"[ECC] corrected bit 33 at index {gi}"]
    assert all(lines == [] for lines in res["lines"][:-1])
    ref = O.OracleMatrix("secded", t.cols, t.rows, t.vals, t.N).cg(pkg.matgen.rhs(t.N), 2000, 1e-10)
    assert all(r["iterations"] == ref["iterations"] for r in res["all"])
