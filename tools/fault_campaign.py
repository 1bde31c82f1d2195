#!/usr/bin/env python
"""Fault-injection campaign: the experiment abft-sparse-cg exists for, at GPU speed.

For every ABFT mode and every flip count k, inject k random bit flips into one random element of the matrix
(CSR/CPUContext.cpp:135-159: all flips hit the same element), run one spmv, and classify what happened:

    corrected   the mode repaired the element and the result equals the clean result
    miscorrected the mode "repaired" something but the result (or the matrix) is wrong
    detected    a fatal event (the reference would exit(1))
    masked      nothing reported, result equal to the clean one (the fault was harmless)
    silent      nothing reported, result WRONG  (silent data corruption)

    python tools/fault_campaign.py --trials 2000 --flips 1,2,3 [--check-oracle]

--check-oracle runs every trial through the CPU oracle as well and insists on identical events and identical
classification (test infrastructure; used by tests/test_gpu_campaign.py).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkg_loader  # noqa: E402

CLASSES = ["corrected", "miscorrected", "detected", "masked", "silent"]


def classify(fatal, events, same_result, same_matrix):
    if fatal:
        return "detected"
    if events:
        return "corrected" if (same_result and same_matrix) else "miscorrected"
    return "masked" if same_result else "silent"


def campaign(pkg, mode, t, vec, trials, flips, seed, oracle=None):
    rng = np.random.default_rng(seed)
    ctx = pkg.CUDAContext(mode)
    A = ctx.create_matrix(t.cols, t.rows, t.vals, t.N)
    clean_cols, _, clean_vals = ctx.read_matrix(A)
    v, w = ctx.create_vector(t.N), ctx.create_vector(t.N)
    ctx.upload(v, vec)
    ctx.spmv(A, v, w)
    assert ctx.poll_events() == ([], False)
    w_clean = ctx.download(w)
    om = oracle.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N) if oracle else None
    counts = {c: 0 for c in CLASSES}
    for _ in range(trials):
        index = int(rng.integers(0, t.nnz))
        bits = [int(b) for b in rng.choice(96, flips, replace=False)]
        for b in bits:
            ctx.flip_bit(A, index, b)
        ctx.spmv(A, v, w)
        events, fatal = ctx.poll_events()
        got = ctx.download(w)
        same_result = bool(np.array_equal(got, w_clean, equal_nan=True))
        # the element after the spmv (corrections are written back)
        cols_now, _, vals_now = ctx.read_matrix(A) if t.nnz <= 200000 else (None, None, None)
        same_matrix = cols_now is None or bool(cols_now[index] == clean_cols[index] and
                                               vals_now.view(np.uint64)[index] == clean_vals.view(np.uint64)[index])
        cls = classify(fatal, events, same_result, same_matrix)
        counts[cls] += 1
        if om is not None:
            for b in bits:
                om.flip(index, b)
            y0, fatal0, ev0 = om.spmv(vec)
            assert (fatal, events) == (fatal0, ev0), (mode, index, bits, events, ev0)
            assert np.array_equal(cols_now[index], om.cols[index]) and vals_now.view(np.uint64)[index] == om.values.view(np.uint64)[index]
            if not fatal:
                assert np.array_equal(got, y0, equal_nan=True), (mode, index, bits)
            # restore the oracle's element
            om.cols[index] = clean_cols[index]
            om.values[index] = clean_vals[index]
        # restore the device element
        ctx._ck(ctx.L.abft_cuda_matrix_poke_col(ctx.ctx, A.handle, index, int(clean_cols[index])), "poke_col")
        ctx._ck(ctx.L.abft_cuda_matrix_poke_value(ctx.ctx, A.handle, index, float(clean_vals[index])), "poke_value")
    ctx.close()
    return counts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=1000)
    ap.add_argument("--flips", default="1,2,3")
    ap.add_argument("--modes", default="none,constraints,sed,sec7,sec8,secded,crc32c")
    ap.add_argument("--grid", type=int, default=24, help="2D 5-point Poisson grid (x4 blocks)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--check-oracle", action="store_true")
    a = ap.parse_args()
    pkg = pkg_loader.load()
    oracle = None
    if a.check_oracle:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as oracle
    t = pkg.matgen.poisson2d(a.grid, 4)
    vec = pkg.matgen.rhs(t.N) + 0.5
    out = {"matrix": f"2D 5-point Poisson {a.grid}^2 -b 4 (N={t.N}, nnz={t.nnz})", "trials": a.trials, "results": {}}
    print(f"{'mode':12s} {'flips':>5s} " + " ".join(f"{c:>12s}" for c in CLASSES))
    for mode in a.modes.split(","):
        for k in [int(x) for x in a.flips.split(",")]:
            c = campaign(pkg, mode, t, vec, a.trials, k, a.seed + k, oracle)
            out["results"][f"{mode}/{k}"] = c
            print(f"{mode:12s} {k:5d} " + " ".join(f"{c[x]:12d}" for x in CLASSES))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
