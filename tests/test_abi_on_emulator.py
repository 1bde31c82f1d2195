"""The C ABI end to end on a machine without a GPU: csrc/abft_cuda.cu -- host orchestration AND kernels -- compiled by
g++ against the multi-threaded CUDA shim and a fake runtime (tests/helpers/build_emu_lib.py: the only source
transformation is `kernel<<<...>>>(args)` -> EMU_LAUNCH), loaded in place of libabftcuda.so (ABFT_CUDA_LIB), and
driven by the GPU tests' own code: smoke() and a few tests of tests/test_gpu_parity.py and tests/test_gpu_coo.py small
enough for the CPU suite (ABFT_SLOW_TESTS=1: more of them).
Test infrastructure; it sees logic (tile plans, fault log ordering, fused-dot bookkeeping, the device-resident loop's
flag ring, row partitions ...), not the device.  Whole GPU test files can be run the same way by hand:

    ABFT_CUDA_LIB=$(python tests/helpers/build_emu_lib.py) python -m pytest tests/test_gpu_coo.py -m gpu
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))

# the default run keeps what takes seconds; ABFT_SLOW_TESTS=1 adds the device-resident loop, CONSTRAINTS row-pointer faults
# and the serial event order (about ten minutes: every CTA barrier is a 256-thread rendezvous)
SUBSET = "empty_matrix or ecc_modes_reject or fused_dot or unaligned or vector_kernels or vector_transfers"
if os.environ.get("ABFT_SLOW_TESTS") == "1":
    SUBSET += " or device_resident_loop or constraints_row_pointer_faults or multiple_faults_are_reported"


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    import build_emu_lib
    return build_emu_lib.build(tmp_path_factory.mktemp("abft_emu"))


def test_smoke_on_the_emulated_library(emu_lib):
    out = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, capture_output=True,
                         text=True, timeout=900, env={**os.environ, "ABFT_CUDA_LIB": emu_lib})
    assert out.returncode == 0 and "smoke ok" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_gpu_parity_subset_on_the_emulated_library(emu_lib):
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-x", "-q", "-k", SUBSET],
                         cwd=ROOT, capture_output=True, text=True, timeout=1500, env={**os.environ, "ABFT_CUDA_LIB": emu_lib})
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-1500:]


def test_gpu_coo_subset_on_the_emulated_library(emu_lib):
    """The streamed COO kernel (csrc/coo_stream.cuh: bulk-copy ring, warp-voted walks, fused dot, stray fix-up) in the
    default CPU suite: encode + spmv bit for bit in every mode on all shapes (staged and unstaged tiles), the dot
    after the spmv with and without strays, index faults in mode none."""
    sel = "grouped and not direct and (dot_after or create_and_spmv or (silent_index and none))"
    if os.environ.get("ABFT_SLOW_TESTS") == "1":
        sel = "grouped and not direct and not all_128"
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_coo.py", "-m", "gpu", "-x", "-q", "-k", sel],
                         cwd=ROOT, capture_output=True, text=True, timeout=1500, env={**os.environ, "ABFT_CUDA_LIB": emu_lib})
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-1500:]
    # ... and the COO differential fuzz of tests/test_gpu_fuzz.py, a few trials per mode (ABFT_SLOW_TESTS=1: 300)
    trials = "300" if os.environ.get("ABFT_SLOW_TESTS") == "1" else "15"
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_fuzz.py", "-m", "gpu", "-x", "-q", "-k", "coo_oracle"],
                         cwd=ROOT, capture_output=True, text=True, timeout=1500,
                         env={**os.environ, "ABFT_CUDA_LIB": emu_lib, "ABFT_FUZZ_TRIALS": trials})
    assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-1500:]
