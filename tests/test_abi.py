"""The C-ABI shared library: builds for sm_100a, loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) when no CUDA device is usable.
No compute calls here -- those are the -m gpu tests."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "abft_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abft_cuda_\w+)\s*\(", text)))


def test_library_builds_and_loads(pkg):
    pkg.capi.build()
    L = pkg.capi.load()
    assert L.abft_cuda_abi_version() == 1


def test_exports_every_declared_symbol(pkg):
    L = pkg.capi.load()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/abft_cuda.h but not exported"
    assert sorted(pkg.capi.SYMBOLS) == syms


def test_mode_names_are_the_reference_registrations(pkg):
    # lower-case, exact strcmp (CGContext.cpp:14; CSR/CPUContext.cpp:415-420); no '-' (run_tests:12-13)
    L = pkg.capi.load()
    names = [L.abft_cuda_mode_name(i).decode() for i in range(8)]
    # (the first six are the reference's; crc32c and secdedv are this repository's own modes)
    assert names == ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c", "secdedv"]
    assert all("-" not in n and n == n.lower() for n in names)
    assert L.abft_cuda_mode_name(8) is None and L.abft_cuda_mode_from_name(b"SECDED") == -1
    for i, n in enumerate(names):
        assert L.abft_cuda_mode_from_name(n.encode()) == i


def test_event_messages_are_the_reference_strings(pkg, O):
    L = pkg.capi.load()
    want = {1: "[ECC] error detected at index 42", 2: "[ECC] corrected bit 7 at index 42",
            3: "[ECC] corrected overall parity bit at index 42", 4: "[ECC] double-bit error detected",
            5: "row size constraint violated for row 42", 6: "row order constraint violated for row42",
            7: "column size constraint violated at index 42", 8: "column order constraint violated at index 42",
            9: "[CRC32C] error detected in row 42"}
    for kind, text in want.items():
        e = pkg.capi.Event(kind, 42, 7)
        buf = ctypes.create_string_buffer(128)
        L.abft_cuda_format_event(ctypes.byref(e), buf, 128)
        assert buf.value.decode() == text
        assert O.format_event(kind, 42, 7) == text            # oracle agrees
        assert bool(L.abft_cuda_event_is_fatal(kind)) == (kind in pkg.capi.FATAL_KINDS)


def test_no_silent_cpu_fallback(pkg):
    L = pkg.capi.load()
    if L.abft_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pkg.CUDAContext("none")
    with pytest.raises(ValueError, match="No implementation found for cuda-bogus"):
        pkg.CUDAContext("bogus")


def test_plugin_binary_registers_cuda_targets(O):
    exe = os.path.join(ROOT, "oracle", "_ref", "cg-csr-cuda")
    if not os.path.exists(exe):
        pytest.skip("cg-csr-cuda not built (needs the reference tree)")
    code, text = O.run_cg_csr(exe, ["--list"])
    assert code == 0
    listed = [ln.strip() for ln in text.splitlines() if "-" in ln]
    for m in ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"]:
        assert f"cuda-{m}" in listed
    for m in ["none", "constraints", "sed", "sec7", "sec8", "secded"]:
        assert f"cpu-{m}" in listed      # the reference's own targets are untouched


def test_device_side_masks_equal_the_reference_constants(tmp_path):
    """The kernels derive their Hamming masks at compile time (csrc/ecc.cuh, csrc/coo.cuh): all 21 CSR and all 28
    COO masks must be the reference's constants (fixtures generated from CSR/ecc.h and COO/ecc.h)."""
    import shutil
    import subprocess
    from conftest import load_golden
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "print_masks")
    src = os.path.join(ROOT, "tests", "helpers", "print_masks.cu")
    subprocess.run([nvcc, "-std=c++17", "-O0", "-o", exe, src], check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    out = subprocess.run([exe], check=True, stdout=subprocess.PIPE).stdout.decode().split("\n")
    csr = [[int(x) for x in ln.split()[2:]] for ln in out if ln.startswith("csr")]
    coo = [[int(x) for x in ln.split()[2:]] for ln in out if ln.startswith("coo")]
    assert csr == load_golden("ecc_kats.json")["masks"]
    assert coo == load_golden("coo_golden.json")["ecc"]["masks"]
