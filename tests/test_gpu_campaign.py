"""Randomised fault campaign (tools/fault_campaign.py): hundreds of random 1-, 2- and 3-bit faults per mode, every one
run through the CUDA target and the CPU oracle; events, repaired elements and results must agree trial by trial, and
the classification table must show what each code promises (SEC7/SEC8/SECDED repair every single flip, SECDED detects
every double flip, SED detects every odd number of flips)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("mode", ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"])
def test_campaign_matches_oracle(pkg, O, mode):
    import fault_campaign as fc
    t = pkg.matgen.poisson2d(16, 2)
    vec = pkg.matgen.rhs(t.N) + 0.5
    res = {k: fc.campaign(pkg, mode, t, vec, 150, k, 100 + k, O) for k in (1, 2, 3)}
    total = {k: sum(v.values()) for k, v in res.items()}
    assert all(n == 150 for n in total.values())
    if mode in ("sec7", "sec8", "secded"):
        # a single flip is either repaired or harmless (SEC7: column bit 24 is unprotected and masked off)
        assert res[1]["corrected"] + res[1]["masked"] == 150 and res[1]["silent"] == 0 and res[1]["miscorrected"] == 0
    if mode == "secded":
        assert res[2]["detected"] == 150
    if mode == "sed":
        assert res[1]["detected"] == 150 and res[3]["detected"] == 150 and res[2]["detected"] == 0
    if mode == "crc32c":
        assert res[1]["silent"] == 0 and res[2]["silent"] == 0 and res[3]["silent"] == 0
