#!/usr/bin/env python
"""Generate tests/golden/coo_*.json from the UNMODIFIED reference COO target.

Run in the build container (needs /root/reference; `make -C oracle` builds oracle/_ref/libref_coo.so and
oracle/_ref/cg-coo from the reference's own sources):

    python tests/golden/make_golden_coo.py

Every number below comes out of the reference's code: its COO/ecc.h inline functions and its COO CPUContext
classes through oracle/_ref/libref_coo.so, and its cg-coo CLI for the cases that exit(1).  The oracle
(oracle/coo_oracle.c) is NOT used here -- the fixtures pin it.
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
import coo_oracle_py as CO  # noqa: E402
import pkg_loader  # noqa: E402

matgen = pkg_loader.load().matgen
OUT = os.path.dirname(os.path.abspath(__file__))


def hexf(x):
    return struct.pack("<d", float(x)).hex()


def hexarr(a):
    return np.ascontiguousarray(a, dtype=np.float64).tobytes().hex()


def parse_faults(text):
    keep = ("[ECC]", "constraint violated", "order violated", "*** flipping")
    return [ln.strip() for ln in text.splitlines() if any(k in ln for k in keep)]


def ecc_fixture():
    R = CO.ref()
    rng = np.random.default_rng(20261018)
    vals = [4.0, -1.0, -1.0, 4.0, -1.0, 0.5, 3.141592653589793, 6.0, 0.0, -0.0, 1e-300, -1e300]
    cols = [0, 1, 255, 65535, 16777215, 123456, 1000000, 7, 0, 1 << 23, 4242, 99]
    rows = [0, 0, 254, 65534, 16777215, 654321, 999999, 8, 4294967295, 1, 4243, 100]
    vals += list(rng.standard_normal(36) * 10.0 ** rng.integers(-8, 8, 36))
    cols += list(rng.integers(0, 1 << 24, 36))
    rows += list(rng.integers(0, 1 << 32, 36))
    kats = []
    for v, c, r in zip(vals, cols, rows):
        c, r = int(c), int(r)
        syn = R.refcoo_ecc_compute_col8(c, r, v)
        sec7 = c | syn
        sec8 = sec7 | (R.refcoo_ecc_overall_parity(sec7, r, v) << 24)
        sed = c | (R.refcoo_ecc_overall_parity(c, r, v) << 31)
        kats.append(dict(value=hexf(v), col=c, row=r, syndrome=syn, overall=R.refcoo_ecc_overall_parity(c, r, v),
                         sed=sed, sec7=sec7, sec8=sec8))
    masks = [[R.refcoo_hamming_mask(p, w) for w in range(4)] for p in range(1, 8)]
    decode = []
    for pos in range(1, 128):
        syn = 0
        for p in range(1, 8):
            if (pos >> (p - 1)) & 1:
                syn |= 1 << (32 - p)
        decode.append(dict(syndrome=syn, bit=R.refcoo_ecc_get_flipped_bit_col8(syn)))
    # every single flip of one SECDED-encoded element: what the reference's check computes for it
    v, c, r = 3.141592653589793, 1000000, 999999
    enc = c | R.refcoo_ecc_compute_col8(c, r, v)
    enc |= R.refcoo_ecc_overall_parity(enc, r, v) << 24
    flips = []
    for bit in range(128):
        w = [enc, r] + list(struct.unpack("<II", struct.pack("<d", v)))
        w[bit // 32] ^= 1 << (bit % 32)
        fv = struct.unpack("<d", struct.pack("<II", w[2], w[3]))[0]
        flips.append(dict(bit=bit, syndrome=R.refcoo_ecc_compute_col8(w[0], w[1], fv),
                          overall=R.refcoo_ecc_overall_parity(w[0], w[1], fv)))
    return dict(masks=masks, kats=kats, decode=decode, single_flips=dict(value=hexf(v), col=c, row=r, encoded=enc, flips=flips))


def cg_fixture():
    """Clean CG trajectories, and every single flip of one element in the correcting modes, on a 12x12 Poisson
    matrix replicated -b 2."""
    t = matgen.poisson2d(12, 2)
    b = matgen.rhs(t.N)
    out = dict(grid=12, blocks=2, N=t.N, nnz=t.nnz, b=hexarr(b), modes={})
    for mode in CO.MODES:
        rc = CO.RefCooContext(mode)
        rc.create_matrix(t.cols, t.rows, t.vals, t.N)
        words = rc.elements.view(np.uint32).reshape(-1, 4).copy()
        y, text = rc.spmv(b)
        res = rc.cg(b, 500, 1e-10)
        assert text == "" and res["text"] == ""
        out["modes"][mode] = dict(iterations=res["iterations"], rr_hist=hexarr(res["rr_hist"]), x=hexarr(res["x"]),
                                  spmv_b=hexarr(y), col_words_sum=int(words[:, 0].astype(np.uint64).sum()),
                                  col_words_head=[int(z) for z in words[:16, 0]])
    flips = []
    index = 77
    for mode in ["sec7", "sec8", "secded"]:
        for bit in range(128):
            rc = CO.RefCooContext(mode)
            rc.create_matrix(t.cols, t.rows, t.vals, t.N)
            clean = rc.elements.view(np.uint32).copy()
            rc.flip(index, bit)
            y, text = rc.spmv(b)
            flips.append(dict(mode=mode, bit=bit, lines=parse_faults(text),
                              restored=bool((rc.elements.view(np.uint32) == clean).all()),
                              y_sum=hexf(float(np.sum(y)))))
    out["single_flips"] = dict(index=index, cases=flips)
    return out


def cli_fixture():
    """cg-coo CLI runs with time() pinned: the rand()-driven injection and the exit(1) cases."""
    t = matgen.poisson2d(32)
    mtx = "/tmp/abft_golden_p32.mtx"
    matgen.write_mtx(mtx, t)
    cases = []
    runs = [("none", [], None), ("secded", ["-x", "1", "VALUE"], 12345), ("secded", ["-x", "1", "INDEX"], 4242),
            ("secded", ["-x", "2"], 777), ("secded", ["-x", "2"], 778), ("sed", ["-x"], 31337), ("sed", ["-x", "2"], 5),
            ("sec7", ["-x", "1"], 99), ("sec8", ["-x", "1"], 100), ("sec8", ["-x", "1", "INDEX"], 101),
            ("constraints", ["-x", "1", "INDEX"], 55), ("constraints", ["-x", "1", "INDEX"], 56),
            ("constraints", ["-x", "1", "INDEX"], 57), ("constraints", ["-x", "1", "INDEX"], 58),
            ("constraints", ["-x", "1", "INDEX"], 59), ("constraints", ["-x", "1", "INDEX"], 60)]
    for mode, extra, seed in runs:
        args = ["-f", mtx, "-b", "2", "-c", "1e-10", "-i", "300", "-t", "cpu", "-m", mode] + extra
        code, text = O.run_cg_csr(CO.REF_EXE, args, fake_time=seed)
        iters = [ln for ln in text.splitlines() if ln.startswith("ran for")]
        cases.append(dict(mode=mode, extra=extra, seed=seed, exit=code, faults=parse_faults(text),
                          ran=iters[0] if iters else None,
                          total_error=[ln for ln in text.splitlines() if ln.startswith("total error")]))
    return dict(grid=32, blocks=2, cases=cases)


if __name__ == "__main__":
    assert CO.have_ref(), "needs the reference build (oracle/_ref)"
    json.dump(dict(source="reference COO/ecc.h + COO/CPUContext.cpp via oracle/_ref/libref_coo.so; CLI cases: "
                          "oracle/_ref/cg-coo with LD_PRELOAD time() pinned",
                   ecc=ecc_fixture(), cg=cg_fixture(), cli=cli_fixture()),
              open(os.path.join(OUT, "coo_golden.json"), "w"), indent=1)
    print("COO golden fixtures written to", OUT)
