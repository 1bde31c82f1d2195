#!/usr/bin/env python
"""Differential fuzzing of the CPU oracle against the live, unmodified reference (test infrastructure; CPU only).

Every trial builds one small matrix in a random reference mode, corrupts it the same way on both sides -- 1..3 bit
flips in each of 1..2 elements (any of the 96 bits), and for `constraints` sometimes a row pointer -- runs ONE spmv
on both and insists on: the same printed lines, the same exit status (the reference exit(1)s from inside spmv on a
fatal fault, so its side runs in a forked child), the same matrix arrays afterwards (corrections are written back),
and, when the reference's behaviour is defined, the same result[] bit for bit.

    python tools/fuzz_oracle_vs_ref.py --trials 5000 --seed 1

Undefined reference behaviour is skipped, not compared: a column >= N after the spmv (the reference has read vec[]
out of bounds, CSR/CPUContext.cpp:127-128), a row pointer that sends a non-`constraints` mode outside the arrays.
Needs oracle/_ref/libref_csr.so (built from /root/reference by oracle/Makefile, or shipped prebuilt).
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

REF_MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]


def matrices(mg, rng):
    """A small pool: stencils, blocks, power-law rows, and one with arbitrary double values."""
    pool = [mg.poisson2d(6), mg.poisson2d(5, 3), mg.poisson3d(4), mg.powerlaw(120, 1500, seed=7)]
    t = mg.powerlaw(80, 600, seed=11)
    vals = rng.standard_normal(t.nnz) * np.exp(rng.uniform(-200, 200, t.nnz))     # whole exponent range
    vals[rng.integers(0, t.nnz, 8)] = [0.0, -0.0, 5e-324, -2.2e-308, 1.7e308, 1.0, -1.0, 3.5]
    pool.append(mg.Triplets(t.cols, t.rows, vals, t.N))
    return pool


# bits whose flip usually keeps the indices inside [0, N) for the small matrices used here (value bits, the low
# index bits, the check byte): half of the elements draw from these, so that most trials stay in defined territory
CSR_TAME = np.array(list(range(64)) + list(range(64, 69)) + list(range(88, 96)))
COO_TAME = np.array(list(range(5)) + list(range(24, 32)) + list(range(32, 37)) + list(range(64, 128)))


def draw_bits(rng, width, tame):
    k = int(rng.integers(1, 4))
    return [int(b) for b in (rng.choice(tame, k, replace=False) if rng.random() < 0.5 else rng.choice(width, k, replace=False))]


def draw_faults(rng, mode, t):
    """-> ([(index, bit), ...], [(row, new_value), ...])"""
    flips = []
    for _ in range(int(rng.integers(1, 3))):
        index = int(rng.integers(0, t.nnz))
        flips += [(index, b) for b in draw_bits(rng, 96, CSR_TAME)]
    pokes = []
    if mode == "constraints" and rng.random() < 0.3:
        row = int(rng.integers(0, t.N + 1))
        pokes.append((row, int(rng.integers(0, t.nnz + 40))))
        if rng.random() < 0.3:
            flips = []
    return flips, pokes


def corrupt_arrays(cols, rows, vals, flips, pokes):
    v64 = vals.view(np.uint64)
    for index, bit in flips:
        if bit < 64:
            v64[index] ^= np.uint64(1) << np.uint64(bit)
        else:
            cols[index] ^= np.uint32(1 << (bit - 64))
    for row, value in pokes:
        rows[row] = value


def same_result(y, y0, t, vec, cols, vals, mode):
    """The bar of the device-vs-oracle comparisons: rows of up to 32 elements keep the reference's sequential order on
    the device (bit-exact); longer rows of irregular tiles are tree-reduced there (1e-13 of the row's |a||x| sum).
    `cols`, `vals`: the matrix after the spmv (repairs included)."""
    lens = np.bincount(t.rows, minlength=t.N)
    seq = lens <= 32
    if not np.array_equal(y[seq].view(np.uint64), y0[seq].view(np.uint64)):
        return False
    if np.all(seq):
        return True
    mask = 0x00FFFFFF if mode not in ("none", "constraints") else 0xFFFFFFFF
    c = np.minimum(cols & np.uint32(mask), t.N - 1).astype(np.int64)      # (columns are clamped, as in the kernels)
    with np.errstate(all="ignore"):
        bound = np.bincount(t.rows, weights=np.abs(vals * vec[c]), minlength=t.N)
        ok = (np.abs(y - y0) <= 1e-13 * bound) | ~np.isfinite(bound) | (np.isnan(y) & np.isnan(y0)) | (y == y0)
    return bool(np.all(ok[~seq]))


def reference_side(O, mode, t, vec, flips, pokes, path):
    """Child process: the reference's create_matrix + corruption + spmv; stdout goes to `path`.txt, arrays to
    `path`.npz when spmv returned (exit status 0), nothing more when it exit(1)ed."""
    pid = os.fork()
    if pid:
        _, status = os.waitpid(pid, 0)
        return os.waitstatus_to_exitcode(status)
    try:
        fd = os.open(path + ".txt", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
        os.dup2(fd, 1)
        os.dup2(os.open(os.devnull, os.O_WRONLY), 2)    # (glibc's "stack smashing detected" when the reference's UB bites)
        rc = O.RefContext(mode)
        rc.create_matrix(t.cols, t.rows, t.vals, t.N)
        c, r, v = rc.matrix_arrays()
        corrupt_arrays(c, r, v, flips, pokes)
        vb, vw = rc.vector(vec), rc.vector()
        O.ref().ref_spmv(rc.ctx, rc.A, vb, vw)          # may exit(1) (flushes stdio on the way out)
        O.C.CDLL(None).fflush(None)
        np.savez(path + ".npz", y=rc.view(vw), cols=c, rows=r, vals=v)
        os._exit(0)
    except BaseException:
        os._exit(99)


def coo_reference_side(CO, mode, t, vec, flips, path):
    """Child process, COO target: as reference_side; elements are the 128-bit {col, row, value} records."""
    pid = os.fork()
    if pid:
        _, status = os.waitpid(pid, 0)
        return os.waitstatus_to_exitcode(status)
    try:
        fd = os.open(path + ".txt", os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600)
        os.dup2(fd, 1)
        os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
        rc = CO.RefCooContext(mode)
        rc.create_matrix(t.cols, t.rows, t.vals, t.N)
        for index, bit in flips:
            rc.flip(index, bit)
        v, out = rc.vector(vec), rc.vector()
        rc.R.refcoo_spmv(rc.ctx, rc.A, v, out)            # may exit(1)
        CO.C.CDLL(None).fflush(None)
        np.savez(path + ".npz", y=rc.view(out), elements=rc.elements.view(np.uint32))
        os._exit(0)
    except BaseException:
        os._exit(99)


def run_coo(trials, seed, verbose=False):
    """The COO target (cg-coo): 1..3 flips (any of the 128 bits) in each of 1..2 elements."""
    import oracle_py as O  # noqa: F401  (builds the oracles)
    import coo_oracle_py as CO
    import pkg_loader
    if not CO.have_ref():
        raise SystemExit("oracle/_ref/libref_coo.so not built")
    mg = pkg_loader.load().matgen
    rng = np.random.default_rng(seed)
    pool = matrices(mg, rng)
    stats = {"trials": 0, "fatal": 0, "corrected": 0, "clean_or_silent": 0, "result_compared": 0, "skipped_undefined": 0}
    path = os.path.join(tempfile.mkdtemp(prefix="fuzz_refcoo_"), "t")
    for trial in range(trials):
        mode = REF_MODES[int(rng.integers(0, len(REF_MODES)))]
        t = pool[int(rng.integers(0, len(pool)))]
        vec = rng.standard_normal(t.N)
        flips = []
        for _ in range(int(rng.integers(1, 3))):
            index = int(rng.integers(0, t.nnz))
            flips += [(index, b) for b in draw_bits(rng, 128, COO_TAME)]
        om = CO.CooOracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        for index, bit in flips:
            om.flip(index, bit)
        y0, fatal0, ev0 = om.spmv(vec)
        lines0 = [CO.format_event(*e) for e in ev0]
        for ext in (".txt", ".npz"):
            if os.path.exists(path + ext):
                os.remove(path + ext)
        code = coo_reference_side(CO, mode, t, vec, flips, path)
        ctx = dict(trial=trial, mode=mode, N=t.N, nnz=t.nnz, flips=flips)
        stats["trials"] += 1
        if code not in (0, 1):
            stats["skipped_undefined"] += 1     # the reference crashed: it indexes result[] / vec[] with whatever it finds
            continue
        with open(path + ".txt") as f:
            lines1 = [ln for ln in f.read().splitlines() if ln]
        got = np.load(path + ".npz") if code == 0 else None
        words = om.elements.view(np.uint32).reshape(-1, 4)
        mask = 0x00FFFFFF if mode not in ("none", "constraints") else 0xFFFFFFFF
        # an index outside [0, N) makes the reference read vec[] / write result[] out of bounds: undefined from the
        # first such element on (it may even have clobbered its own matrix), nothing is compared
        if np.any((words[:, 0] & np.uint32(mask)) >= t.N) or np.any(words[:, 1] >= t.N):
            stats["skipped_undefined"] += 1
            continue
        assert lines0 == lines1, (ctx, lines0, lines1)
        assert fatal0 == (code == 1), (ctx, fatal0, code)
        if fatal0:
            stats["fatal"] += 1
            continue
        assert np.array_equal(got["elements"].reshape(-1, 4), words), ctx
        stats["corrected" if lines0 else "clean_or_silent"] += 1
        assert np.array_equal(got["y"].view(np.uint64), y0.view(np.uint64)), ctx
        stats["result_compared"] += 1
        if verbose and trial % 500 == 0:
            print(trial, stats, flush=True)
    return stats


def run(trials, seed, verbose=False):
    import oracle_py as O
    import pkg_loader
    if not O.have_ref():
        raise SystemExit("oracle/_ref not built")
    mg = pkg_loader.load().matgen
    rng = np.random.default_rng(seed)
    pool = matrices(mg, rng)
    stats = {"trials": 0, "fatal": 0, "corrected": 0, "clean_or_silent": 0, "result_compared": 0, "skipped_undefined": 0}
    tmp = tempfile.mkdtemp(prefix="fuzz_ref_")
    path = os.path.join(tmp, "t")
    for trial in range(trials):
        mode = REF_MODES[int(rng.integers(0, len(REF_MODES)))]
        t = pool[int(rng.integers(0, len(pool)))]
        vec = rng.standard_normal(t.N)
        flips, pokes = draw_faults(rng, mode, t)
        # oracle
        om = O.OracleMatrix(mode, t.cols, t.rows, t.vals, t.N)
        corrupt_arrays(om.cols, om.rows, om.values, flips, pokes)
        y0, fatal0, ev0 = om.spmv(vec)
        lines0 = [O.format_event(*e) for e in ev0]
        # reference
        for ext in (".txt", ".npz"):
            if os.path.exists(path + ext):
                os.remove(path + ext)
        code = reference_side(O, mode, t, vec, flips, pokes, path)
        ctx = dict(trial=trial, mode=mode, N=t.N, nnz=t.nnz, flips=flips, pokes=pokes)
        if code not in (0, 1):
            # the reference crashed (e.g. SIGSEGV on an out-of-range column in mode none): undefined, nothing to compare
            stats["skipped_undefined"] += 1
            stats["trials"] += 1
            continue
        with open(path + ".txt") as f:
            lines1 = [ln for ln in f.read().splitlines() if ln]
        assert lines0 == lines1, (ctx, lines0, lines1)
        assert fatal0 == (code == 1), (ctx, fatal0, code)
        stats["trials"] += 1
        if fatal0:
            stats["fatal"] += 1
            continue
        got = np.load(path + ".npz")
        assert np.array_equal(got["cols"], om.cols), ctx
        assert np.array_equal(got["vals"].view(np.uint64), om.values.view(np.uint64)), ctx
        assert np.array_equal(got["rows"], om.rows), ctx
        stats["corrected" if lines0 else "clean_or_silent"] += 1
        mask = 0x00FFFFFF if mode not in ("none", "constraints") else 0xFFFFFFFF
        if np.any((om.cols & np.uint32(mask)) >= t.N):
            stats["skipped_undefined"] += 1          # the reference read vec[] out of bounds
            continue
        assert np.array_equal(got["y"].view(np.uint64), y0.view(np.uint64)), ctx
        stats["result_compared"] += 1
        if verbose and trial % 500 == 0:
            print(trial, stats, flush=True)
    return stats


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--format", default="csr", choices=["csr", "coo"])
    a = ap.parse_args()
    # the reference's spmv is `omp parallel for` over rows: one thread gives the serial order of the messages
    # (set before libref_csr.so, and with it libgomp, is loaded)
    os.environ["OMP_NUM_THREADS"] = "1"
    print(json.dumps((run_coo if a.format == "coo" else run)(a.trials, a.seed, verbose=True)))
