"""The drop-in itself: the reference's UNMODIFIED cg.cpp driver (matrix-market loader, -b blocking,
-x injection, -m modes) linked with host/CUDAContext.cpp, run as `cg-csr -t cuda`, against the same
binary's `-t cpu` and against the committed reference CLI fixtures.  The binary is built in the build
container (it contains reference objects, so it lives in the git-ignored oracle/_ref/) and travels to
the GPU box with the snapshot."""
import os
import re

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "cg-csr-cuda")
CUDA_MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"]
KEEP = ("[ECC]", "[CRC32C]", "constraint violated", "*** flipping")


def faults(text):
    return [ln.strip() for ln in text.splitlines() if any(k in ln for k in KEEP)]


def field(text, prefix):
    for ln in text.splitlines():
        if ln.startswith(prefix):
            return ln.strip()
    return None


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/cg-csr-cuda not built (needs the reference tree at build time)")
    return EXE


@pytest.fixture(scope="module")
def mtx32(pkg, tmp_path_factory):
    path = str(tmp_path_factory.mktemp("mtx") / "p32.mtx")
    pkg.matgen.write_mtx(path, pkg.matgen.poisson2d(32))
    return path


@pytest.fixture(scope="module")
def mtx256(pkg, tmp_path_factory):
    path = str(tmp_path_factory.mktemp("mtx") / "p256.mtx")
    pkg.matgen.write_mtx(path, pkg.matgen.poisson2d(256))
    return path


def test_list_shows_cuda_targets(O, exe):
    code, text = O.run_cg_csr(exe, ["--list"])
    assert code == 0
    for m in CUDA_MODES:
        assert f"cuda-{m}" in text


@pytest.mark.parametrize("mode", CUDA_MODES)
def test_clean_run_matches_cpu_target(O, exe, mtx256, mode):
    """BASELINE configs 1/2 through the real CLI: 256x256 Poisson, -b 1 -c 1e-10 -> 683 iterations."""
    args = ["-f", mtx256, "-b", "1", "-c", "1e-10", "-m"]
    code, text = O.run_cg_csr(exe, args + [mode, "-t", "cuda"])
    assert code == 0, text[-2000:]
    cpu_mode = mode if mode != "crc32c" else "none"          # crc32c has no cpu counterpart in the reference
    code0, text0 = O.run_cg_csr(exe, args + [cpu_mode, "-t", "cpu"])
    assert code0 == 0
    assert field(text, "ran for") == field(text0, "ran for") == "ran for 683 iterations"
    assert field(text, "total error") == field(text0, "total error")
    assert field(text, "max error") == field(text0, "max error")
    assert field(text, "implementation") == f"implementation        = cuda-{mode}"
    # the per-iteration residual lines agree to the 4 decimals cg.cpp prints (cg.cpp:117)
    it = [float(ln.split("=")[1]) for ln in text.splitlines() if ln.startswith("iteration")]
    it0 = [float(ln.split("=")[1]) for ln in text0.splitlines() if ln.startswith("iteration")]
    assert len(it) == len(it0) == 683
    assert all(abs(a - b) <= 1.5e-4 + 1e-9 * abs(b) for a, b in zip(it, it0))


def test_cli_fixture_replay(O, exe, mtx32, golden_cli):
    """Pinned-seed -x runs of the reference CLI (time() pinned by LD_PRELOAD): `-t cuda` flips the same
    bits, prints the same messages, exits with the same status and converges in the same iterations."""
    for case in golden_cli["cases"]:
        args = ["-f", mtx32, "-b", str(golden_cli["blocks"]), "-c", "1e-10", "-i", "300", "-t", "cuda", "-m", case["mode"]]
        code, text = O.run_cg_csr(exe, args + case["extra"], fake_time=case["seed"])
        assert faults(text) == case["faults"], (case, text[-1500:])
        assert code == case["exit"], case
        assert field(text, "ran for") == case["ran"], case
        if case["total_error"]:
            assert field(text, "total error") == case["total_error"][0]


def test_cpu_and_cuda_agree_under_random_seeds(O, exe, mtx32):
    """More seeds than the fixtures hold: run both targets of the same binary side by side."""
    for mode, extra in [("secded", ["-x", "1"]), ("secded", ["-x", "2"]), ("sec7", ["-x", "1", "VALUE"]),
                        ("sec8", ["-x", "1", "INDEX"]), ("sed", ["-x", "1"]), ("constraints", ["-x", "1", "INDEX"])]:
        for seed in range(2000, 2004):
            args = ["-f", mtx32, "-b", "4", "-c", "1e-10", "-i", "400", "-m", mode] + extra
            c1, t1 = O.run_cg_csr(exe, args + ["-t", "cuda"], fake_time=seed)
            c0, t0 = O.run_cg_csr(exe, args + ["-t", "cpu"], fake_time=seed)
            assert (c1, faults(t1), field(t1, "ran for")) == (c0, faults(t0), field(t0, "ran for")), (mode, extra, seed)


def test_reference_run_tests_protocol(O, exe, mtx32):
    """The reference's own smoke test (run_tests:1-82), restated: every --list pair runs clean; `sed` modes
    report 'error detected at index' under -x; `sec` modes exit 0 under -x; `secded` modes report
    'double-bit error detected' under -x 2 (flaky by design when both flips hit the same bit, 1/96)."""
    code, text = O.run_cg_csr(exe, ["--list"])
    impls = [ln.strip() for ln in text.splitlines() if "-" in ln and ln.strip().startswith("cuda")]
    assert len(impls) == 8      # the reference's six modes + crc32c + secdedv
    base = ["-f", mtx32, "-b", "4"]
    for impl in impls:
        target, mode = impl.split("-")
        code, _ = O.run_cg_csr(exe, base + ["-t", target, "-m", mode])
        assert code == 0, impl
        if "sed" in impl:
            _, t = O.run_cg_csr(exe, base + ["-t", target, "-m", mode, "-x"], fake_time=11)
            assert "error detected at index" in t, impl
        if "sec" in impl:
            code, _ = O.run_cg_csr(exe, base + ["-t", target, "-m", mode, "-x"], fake_time=12)
            assert code == 0, impl
        if "secded" in impl:
            _, t = O.run_cg_csr(exe, base + ["-t", target, "-m", mode, "-x", "2"], fake_time=777)
            assert "double-bit error detected" in t, impl


def test_unknown_mode_fails_like_the_reference(O, exe, mtx32):
    # CGContext.cpp:18-23
    code, text = O.run_cg_csr(exe, ["-f", mtx32, "-t", "cuda", "-m", "bogus"])
    assert code == 1 and "No implementation found for cuda-bogus" in text


# ---------------------------------------------------------------- multi-GPU through the same CLI
def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode", ["none", "secded", "constraints", "crc32c"])
def test_multi_gpu_cli_matches_cpu_target(O, exe, mtx256, mode):
    """ABFT_CUDA_NGPUS=g: the plugin row-partitions the matrix over g GPUs of the box; same output as -t cpu."""
    g = min(_ngpus(), 8)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    args = ["-f", mtx256, "-b", "1", "-c", "1e-10", "-m"]
    code, text = O.run_cg_csr(exe, args + [mode, "-t", "cuda"], env_extra={"ABFT_CUDA_NGPUS": str(g)})
    assert code == 0, text[-2000:]
    code0, text0 = O.run_cg_csr(exe, args + [mode if mode != "crc32c" else "none", "-t", "cpu"])
    assert field(text, "ran for") == field(text0, "ran for") == "ran for 683 iterations"
    assert field(text, "total error") == field(text0, "total error")
    assert field(text, "max error") == field(text0, "max error")


def test_multi_gpu_cli_fixture_replay(O, exe, mtx32, golden_cli):
    """The pinned-seed injection fixtures with the matrix split over the GPUs: same flips (routed to the owning
    GPU), same messages with global indices, same exit status."""
    g = min(_ngpus(), 8)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    for case in golden_cli["cases"]:
        args = ["-f", mtx32, "-b", str(golden_cli["blocks"]), "-c", "1e-10", "-i", "300", "-t", "cuda", "-m", case["mode"]]
        code, text = O.run_cg_csr(exe, args + case["extra"], fake_time=case["seed"], env_extra={"ABFT_CUDA_NGPUS": str(g)})
        assert faults(text) == case["faults"], (case, text[-1500:])
        assert code == case["exit"], case
        assert field(text, "ran for") == case["ran"], case
