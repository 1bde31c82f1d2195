"""The COO drop-in: the reference's UNMODIFIED cg.cpp driver linked with host/COOCUDAContext.cpp, run as
`cg-coo -t cuda`, against the same binary's `-t cpu` and against the reference cg-coo CLI fixtures.
"""
import os

import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "cg-coo-cuda")
MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded"]
KEEP = ("[ECC]", "constraint violated", "order violated", "*** flipping")


def faults(text):
    return [ln.strip() for ln in text.splitlines() if any(k in ln for k in KEEP)]


def field(text, prefix):
    for ln in text.splitlines():
        if ln.startswith(prefix):
            return ln.strip()
    return None


def rr_lines(text):
    """rr of every `iteration %5d :  rr = %12.4lf` line (cg.cpp:116)."""
    return [float(ln.split("=")[1]) for ln in text.splitlines() if ln.startswith("iteration") and "rr =" in ln]


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/cg-coo-cuda not built (needs the reference tree at build time)")
    return EXE


@pytest.fixture(scope="module")
def mtx32(pkg, tmp_path_factory):
    path = str(tmp_path_factory.mktemp("mtx") / "p32.mtx")
    pkg.matgen.write_mtx(path, pkg.matgen.poisson2d(32))
    return path


@pytest.mark.parametrize("mode", MODES)
def test_clean_run_matches_cpu_target(O, exe, mtx32, mode):
    args = ["-f", mtx32, "-b", "2", "-c", "1e-10", "-m", mode]
    code, text = O.run_cg_csr(exe, args + ["-t", "cuda"])
    code0, text0 = O.run_cg_csr(exe, args + ["-t", "cpu"])
    assert code == 0 == code0, text[-2000:]
    assert field(text, "ran for") == field(text0, "ran for")
    assert field(text, "total error") == field(text0, "total error")
    assert field(text, "implementation") == f"implementation        = cuda-{mode}"


def test_cli_fixture_replay(O, exe, mtx32):
    """Pinned-seed -x runs of the reference cg-coo (time() pinned by LD_PRELOAD): `-t cuda` flips the same bits,
    prints the same messages, exits with the same status and converges in the same iterations."""
    G = load_golden("coo_golden.json")["cli"]
    for case in G["cases"]:
        args = ["-f", mtx32, "-b", str(G["blocks"]), "-c", "1e-10", "-i", "300", "-t", "cuda", "-m", case["mode"]]
        code, text = O.run_cg_csr(exe, args + case["extra"], fake_time=case["seed"])
        assert faults(text) == case["faults"], (case, text[-1500:])
        assert code == case["exit"], case
        assert field(text, "ran for") == case["ran"], case


def test_cpu_and_cuda_agree_under_random_seeds(O, exe, mtx32):
    for mode, extra in [("secded", ["-x", "1"]), ("secded", ["-x", "2"]), ("sec7", ["-x", "1", "VALUE"]),
                        ("sec8", ["-x", "1", "INDEX"]), ("sed", ["-x", "1"]), ("constraints", ["-x", "1", "INDEX"]),
                        ("none", ["-x", "1", "VALUE"])]:
        for seed in range(3000, 3004):
            args = ["-f", mtx32, "-b", "4", "-c", "1e-10", "-i", "400", "-m", mode] + extra
            c1, t1 = O.run_cg_csr(exe, args + ["-t", "cuda"], fake_time=seed)
            c0, t0 = O.run_cg_csr(exe, args + ["-t", "cpu"], fake_time=seed)
            assert (c1, faults(t1)) == (c0, faults(t0)), (mode, extra, seed)
            if mode == "none":
                # nothing detects the flip: the solve goes on with a corrupted (possibly indefinite) matrix.  Seed 3001
                # flips exponent bit 123 of a diagonal entry (4.0 -> 5e39): after one iteration r_j = r_j - rr/p_j, and
                # the next search direction p_j = r_j + beta p_j cancels to ~1e-12 of its terms -- the 1e-16 difference of
                # a reordered reduction decides whether rr is 1e36 or 1e6 from then on (-t cpu runs 400 iterations,
                # -t cuda "converges" after 388).  Only the first residual is comparable.
                r1, r0 = rr_lines(t1)[:1], rr_lines(t0)[:1]
                assert len(r1) == len(r0) == 1 and abs(r1[0] - r0[0]) <= 1e-9 * abs(r0[0]), (seed, r1, r0)
            else:
                assert field(t1, "ran for") == field(t0, "ran for"), (mode, extra, seed)
