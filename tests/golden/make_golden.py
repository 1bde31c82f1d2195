#!/usr/bin/env python
"""Generate the fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference; `make -C oracle` builds
oracle/_ref/ from the reference's own sources):

    python tests/golden/make_golden.py

Every number below comes out of the reference's code: its CSR/ecc.h inline
functions and its CPUContext classes through oracle/_ref/libref_csr.so, and its
cg-csr CLI (oracle/_ref/cg-csr) for the cases that exit(1).  The oracle
(oracle/abft_oracle.c) is NOT used here -- the fixtures pin it.
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import pkg_loader  # noqa: E402

matgen = pkg_loader.load().matgen
OUT = os.path.dirname(os.path.abspath(__file__))
REF_MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]


def hexf(x):
    return struct.pack("<d", float(x)).hex()


def hexarr(a):
    return np.ascontiguousarray(a, dtype=np.float64).tobytes().hex()


def ecc_fixture():
    R = O.ref()
    rng = np.random.default_rng(20261018)
    vals = [4.0, -1.0, -1.0, 4.0, -1.0, 0.5, 3.141592653589793, 6.0, 0.0, -0.0, 1e-300, -1e300]
    cols = [0, 1, 255, 65535, 16777215, 123456, 1000000, 7, 0, 1 << 23, 4242, 99]
    vals += list(rng.standard_normal(52) * 10.0 ** rng.integers(-8, 8, 52))
    cols += list(rng.integers(0, 1 << 24, 52))
    kats = []
    for v, c in zip(vals, cols):
        c = int(c)
        syn = R.ref_ecc_compute_col8(v, c)
        sec7 = c | syn
        sec8 = sec7 | (R.ref_ecc_overall_parity(v, sec7) << 24)
        sed = c | (R.ref_ecc_overall_parity(v, c) << 31)
        kats.append(dict(value=hexf(v), column=c, syndrome=syn, overall=R.ref_ecc_overall_parity(v, c),
                         sed=sed, sec7=sec7, sec8=sec8))
    masks = [[R.ref_hamming_mask(p, w) for w in range(3)] for p in range(1, 8)]
    decode = []
    for pos in range(1, 128):
        syn = 0
        for p in range(1, 8):
            if (pos >> (p - 1)) & 1:
                syn |= 1 << (32 - p)
        decode.append(dict(syndrome=syn, bit=R.ref_ecc_get_flipped_bit_col8(syn)))
    # single flips of one encoded element: syndrome/parity the reference computes for each of the 96 bits
    v, c = 3.141592653589793, 1000000
    syn = R.ref_ecc_compute_col8(v, c)
    enc = c | syn
    enc |= R.ref_ecc_overall_parity(v, enc) << 24
    flips = []
    for bit in range(96):
        w = list(struct.unpack("<II", struct.pack("<d", v))) + [enc]
        w[bit // 32] ^= 1 << (bit % 32)
        fv = struct.unpack("<d", struct.pack("<II", w[0], w[1]))[0]
        flips.append(dict(bit=bit, syndrome=R.ref_ecc_compute_col8(fv, w[2]), overall=R.ref_ecc_overall_parity(fv, w[2])))
    json.dump(dict(source="reference CSR/ecc.h via oracle/_ref/libref_csr.so", masks=masks, kats=kats, decode=decode,
                   single_flips=dict(value=hexf(v), column=c, encoded=enc, flips=flips)),
              open(os.path.join(OUT, "ecc_kats.json"), "w"), indent=1)


def parse_faults(text):
    keep = ("[ECC]", "constraint violated", "*** flipping")
    return [ln.strip() for ln in text.splitlines() if any(k in ln for k in keep)]


def cg_fixture():
    """Clean CG trajectories + injected single flips (non-fatal modes) on a 16x16 Poisson matrix."""
    t = matgen.poisson2d(16)
    b = matgen.rhs(t.N)
    out = dict(source="reference CPUContext via oracle/_ref/libref_csr.so", grid=16, N=t.N, nnz=t.nnz, b=hexarr(b),
               modes={})
    clean_cols = {}
    for mode in REF_MODES:
        rc = O.RefContext(mode)
        rc.create_matrix(t.cols, t.rows, t.vals, t.N)
        c, r, v = rc.matrix_arrays()
        enc_cols = c.copy()
        res = rc.cg(b, 1000, 1e-10)
        entry = dict(iterations=res["iterations"], rr_hist=hexarr(res["rr_hist"]), x=hexarr(res["x"]),
                     cols_sum=int(enc_cols.astype(np.uint64).sum()), cols_head=[int(z) for z in enc_cols[:16]])
        clean_cols[mode] = enc_cols
        # spmv of the clean matrix on b
        rc2 = O.RefContext(mode)
        rc2.create_matrix(t.cols, t.rows, t.vals, t.N)
        vb, vw = rc2.vector(b), rc2.vector()
        rc2.spmv(vb, vw)
        entry["spmv_b"] = hexarr(rc2.view(vw))
        out["modes"][mode] = entry
    # injected flips, correcting modes only (the fatal ones go through the CLI below)
    faults = []
    rng = np.random.default_rng(7)
    for mode in ["sec7", "sec8", "secded"]:
        for trial in range(12):
            index = int(rng.integers(0, t.nnz))
            bit = int(rng.integers(0, 96))
            rc = O.RefContext(mode)
            rc.create_matrix(t.cols, t.rows, t.vals, t.N)
            c, r, v = rc.matrix_arrays()
            if bit < 64:
                v.view(np.uint64)[index] ^= np.uint64(1) << np.uint64(bit)
            else:
                c[index] ^= np.uint32(1 << (bit - 64))
            res = rc.cg(b, 1000, 1e-10)
            faults.append(dict(mode=mode, index=index, bit=bit, lines=parse_faults(res["text"]),
                               iterations=res["iterations"], rr_last=hexf(res["rr_hist"][-1]),
                               restored=bool((c == clean_cols[mode]).all() and (v == t.vals).all())))
    out["faults"] = faults
    json.dump(out, open(os.path.join(OUT, "cg_poisson2d_16.json"), "w"), indent=1)


def cli_fixture():
    """cg-csr CLI runs with time() pinned: the rand()-driven injection and the exit(1) cases."""
    t = matgen.poisson2d(32)
    mtx = "/tmp/abft_golden_p32.mtx"
    matgen.write_mtx(mtx, t)
    cases = []
    runs = [("none", [], None), ("secded", ["-x", "1", "VALUE"], 12345), ("secded", ["-x", "1", "INDEX"], 4242),
            ("secded", ["-x", "2"], 777), ("secded", ["-x", "2"], 778), ("sed", ["-x"], 31337), ("sed", ["-x", "2"], 5),
            ("sec7", ["-x", "1"], 99), ("sec8", ["-x", "1"], 100), ("sec8", ["-x", "1", "INDEX"], 101),
            ("constraints", ["-x", "1", "INDEX"], 55), ("constraints", ["-x", "1", "INDEX"], 56),
            ("constraints", ["-x", "1", "INDEX"], 57), ("constraints", ["-x", "1", "INDEX"], 58)]
    for mode, extra, seed in runs:
        args = ["-f", mtx, "-b", "2", "-c", "1e-10", "-i", "300", "-t", "cpu", "-m", mode] + extra
        code, text = O.run_cg_csr(O.REF_EXE, args, fake_time=seed)
        iters = [ln for ln in text.splitlines() if ln.startswith("ran for")]
        cases.append(dict(mode=mode, extra=extra, seed=seed, exit=code, faults=parse_faults(text),
                          ran=iters[0] if iters else None,
                          total_error=[ln for ln in text.splitlines() if ln.startswith("total error")]))
    json.dump(dict(source="reference cg-csr CLI (oracle/_ref/cg-csr), LD_PRELOAD time() pinned", grid=32, blocks=2,
                   cases=cases), open(os.path.join(OUT, "cli_poisson2d_32.json"), "w"), indent=1)


if __name__ == "__main__":
    assert O.have_ref(), "needs the reference build (oracle/_ref)"
    ecc_fixture()
    cg_fixture()
    cli_fixture()
    print("golden fixtures written to", OUT)
