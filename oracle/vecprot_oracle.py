"""ORACLE SUPPORT (test infrastructure; see abft_oracle.h) -- mode `secdedv`, PARITY UNPINNED.

The reference protects the matrix elements only: its vectors are plain arrays of doubles (CSR/CPUContext.cpp:54-113)
and nothing in /root/reference defines a vector code.  `secdedv` is this repository's own definition:

    matrix elements   SECDED, exactly mode secded (pinned to the reference)
    row pointers      SECDED per pair (oracle/rowptr_ecc.c, csrc/rowsecc.cuh)
    vectors           SED: the mantissa's least significant bit of every stored double holds the parity of the other
                      63 bits, so every stored element has even parity; an element with odd parity is reported when
                      it is next read ("[ECC] error detected in vector (...) at index i", fatal).

This file restates the vector arithmetic in numpy -- every value a kernel stores goes through `encode`, every value it
reads is the stored one (parity bit and all) -- and the CG loop of cg.cpp:83-118 over it, with the matrix product taken
from the pinned secded oracle.  spmv row sums, x, r and p are bit-identical to the kernels' given the same scalars;
the reductions are reordered on the device, so trajectories agree to the tolerance the tests state.
"""
import numpy as np

import oracle_py as O


def encode(v):
    """the stored form of v: bit 0 := parity of bits 1..63"""
    bits = np.ascontiguousarray(v, dtype=np.float64).view(np.uint64) & ~np.uint64(1)
    return (bits | (np.bitwise_count(bits).astype(np.uint64) & np.uint64(1))).view(np.float64)


def odd(v):
    """True where a stored element's parity is odd (corrupted)"""
    return (np.bitwise_count(np.ascontiguousarray(v, dtype=np.float64).view(np.uint64)) & 1).astype(bool)


class VecProtCG:
    def __init__(self, cols, rows, vals, N):
        self.N = int(N)
        self.A = O.OracleMatrix("secded", cols, rows, vals, N)

    def spmv(self, p):
        y, fatal, events = self.A.spmv(p)
        return encode(y), fatal, events

    def cg(self, b, max_itrs, conv):
        """-> dict(iterations, rr_hist, x): unmap encodes b and x; r = b; p = r; then cg.cpp's loop"""
        N = self.N
        x = encode(np.zeros(N))
        r = encode(b).copy()
        p = r.copy()
        rr = float(np.dot(r, r))
        hist = []
        it = 0
        while it < max_itrs and rr > conv:
            w, fatal, events = self.spmv(p)
            assert not fatal and not events
            pw = float(np.dot(p, w))
            alpha = rr / pw
            x = encode(x + alpha * p)
            r = encode(r - alpha * w)
            rr_new = float(np.dot(r, r))
            p = encode(r + (rr_new / rr) * p)
            rr = rr_new
            hist.append(rr)
            it += 1
        return dict(iterations=it, rr_hist=np.asarray(hist), x=x, r=r, p=p)
