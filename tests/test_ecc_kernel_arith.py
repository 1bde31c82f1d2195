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


def test_crc32c_published_vectors(E, O):
    """RFC 3720 appendix B.4 (tests/golden/crc32c_rfc3720.json): the oracle's CRC and the three table forms the device
    tables are built from (byte table, slicing-by-4, the 5-bit chunk tables of F / F.F / F.F.F the spmv check uses)
    all reproduce the published CRC-32C values."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "crc32c_rfc3720.json")) as f:
        G = json.load(f)
    u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    E.ecc_crc32c_bytes.restype = C.c_uint32
    E.ecc_crc32c_bytes.argtypes = [u8p, C.c_int]
    for f_ in (E.ecc_crc32c_words_by4, E.ecc_crc32c_words_by_maps, E.ecc_crc32c_words_by_elem_chunks):
        f_.restype = C.c_uint32
        f_.argtypes = [u32p, C.c_int]
    seen_maps = seen_elem = 0
    for name, v in G["vectors"].items():
        data = bytes.fromhex(v["bytes"])
        want = int(v["crc32c"], 16)
        assert O.lib().orc_crc32c(data, len(data)) == want, name
        arr = np.frombuffer(data, dtype=np.uint8).copy()
        assert E.ecc_crc32c_bytes(arr, len(data)) == want, name
        if len(data) % 4 == 0:
            words = arr.view("<u4").copy()
            assert E.ecc_crc32c_words_by4(words, len(words)) == want, name
            if len(words) % 3 == 0:
                assert E.ecc_crc32c_words_by_maps(words, len(words)) == want, name
                seen_maps += 1
                if np.all(words[2::3] < 2**24):
                    assert E.ecc_crc32c_words_by_elem_chunks(words, len(words)) == want, name
                    seen_elem += 1
    assert seen_maps >= 1 and seen_elem >= 1     # the 48-byte PDU: four 96-bit elements
    # random messages: the three table forms agree with the oracle
    rng = np.random.default_rng(9)
    for n_words in (3, 6, 21, 300):
        words = rng.integers(0, 2**32, n_words, dtype=np.uint64).astype(np.uint32)
        want = O.lib().orc_crc32c(words.tobytes(), 4 * n_words)
        assert E.ecc_crc32c_words_by4(words, n_words) == want == E.ecc_crc32c_words_by_maps(words, n_words)
        assert E.ecc_crc32c_bytes(words.view(np.uint8).copy(), 4 * n_words) == want
        words[2::3] &= 0x00FFFFFF    # the 88-bit element form: columns of 24 bits
        want = O.lib().orc_crc32c(words.tobytes(), 4 * n_words)
        assert E.ecc_crc32c_words_by_elem_chunks(words, n_words) == want


def test_crc32c_row_known_answer_oracle(O, pkg):
    """The RFC's 48-byte PDU is four 12-byte elements: a one-row crc32c matrix built from it carries the published CRC
    in the top bytes of its four column words (the row format of DESIGN.md 5, pinned to a published vector)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "crc32c_rfc3720.json")) as f:
        K = json.load(f)["row_known_answer"]
    vals = np.array([np.frombuffer(bytes.fromhex(h), dtype="<f8")[0] for h in K["values_hex"]])
    cols = np.array(K["columns"], dtype=np.uint32)
    N = int(cols.max()) + 1
    om = O.OracleMatrix("crc32c", cols, np.zeros(4, np.uint32), vals, N)
    assert [int(c >> 24) for c in om.cols] == K["crc_bytes_in_column_top_bytes"]
    assert [int(c & 0xFFFFFF) for c in om.cols] == K["columns"]
