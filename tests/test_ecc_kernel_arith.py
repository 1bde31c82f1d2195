"""The ECC arithmetic of the kernels (abft-sparse-cg_b200/csrc/ecc.cuh: masks, syndrome, overall parity, syndrome -> bit,
encode), compiled for the host through the CUDA shim (tests/helpers/) and held against the oracle, which is pinned to
the reference.  Runs without a GPU; the GPU tests check the same arithmetic where it executes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def E(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ecc") / "libecc_emul.so")
    h = os.path.join(ROOT, "tests", "helpers")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I" + os.path.join(h, "cuda_shim"), "-o", so,
                    os.path.join(h, "ecc_emul.cpp")], check=True)
    L = C.CDLL(so)
    for name in ("ecc_parities_masks", "ecc_syndrome7", "ecc_syndrome7_any"):
        getattr(L, name).restype = C.c_uint32
        getattr(L, name).argtypes = [C.c_uint32] * 3
    L.ecc_flipped_bit.restype = C.c_uint32
    L.ecc_flipped_bit.argtypes = [C.c_uint32]
    L.ecc_encode_column.restype = C.c_uint32
    L.ecc_encode_column.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    L.ecc_hamming_mask.restype = C.c_uint32
    L.ecc_hamming_mask.argtypes = [C.c_int, C.c_int]
    return L


def test_kernel_ecc_arithmetic_equals_the_oracle(E, O):
    """masks, syndrome, overall parity, syndrome -> bit, encode: kernel source vs oracle (itself pinned to the reference)"""
    L = O.lib()
    assert [[E.ecc_hamming_mask(p, w) for w in range(3)] for p in range(1, 8)] == \
           [[L.orc_hamming_mask(p, w) for w in range(3)] for p in range(1, 8)]
    rng = np.random.default_rng(4)
    vals = rng.integers(0, 2**63, 4000, dtype=np.int64).view(np.float64)
    cols = rng.integers(0, 2**32, 4000, dtype=np.uint64)
    for v, c in zip(vals, cols):
        bits = int(np.float64(v).view(np.uint64))
        w0, w1, c = bits & 0xFFFFFFFF, bits >> 32, int(c)
        s = E.ecc_syndrome7(w0, w1, c)
        assert s == L.orc_ecc_compute_col8(float(v), c)
        assert (E.ecc_parities_masks(w0, w1, c) >> 7) == L.orc_ecc_overall_parity(float(v), c)
        assert E.ecc_syndrome7_any(w0, w1, c) == (1 if s else 0)
        for mode in (2, 3, 4, 5):
            assert E.ecc_encode_column(mode, w0, w1, c & 0x00FFFFFF) == L.orc_encode_column(mode, float(v), c & 0x00FFFFFF)
    for pos in range(1, 128):
        s = sum(((pos >> (p - 1)) & 1) << (32 - p) for p in range(1, 8))
        assert E.ecc_flipped_bit(s) == L.orc_ecc_get_flipped_bit_col8(s), pos
