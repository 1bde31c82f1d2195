#!/usr/bin/env python
"""run_benchmark-style timing of the unmodified cg.cpp CLI (reference protocol: run_benchmark:3-26, `time taken`
line of cg.cpp:125) for -t cpu and -t cuda on a synthetic matrix, 2D 5-point Poisson n^2 replicated -b B.
    python tools/cli_bench.py --n 512 --blocks 64 --iters 200 --modes secded,none --ngpus 1,2
"""
import argparse
import json
import os
import re
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402
import pkg_loader  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--blocks", type=int, default=64)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--cpu-iters", type=int, default=10)
    ap.add_argument("--modes", default="none,secded")
    ap.add_argument("--ngpus", default="1")
    ap.add_argument("--runs", type=int, default=3)
    a = ap.parse_args()
    pkg = pkg_loader.load()
    exe = os.path.join(ROOT, "oracle", "_ref", "cg-csr-cuda")
    mtx = os.path.join(tempfile.gettempdir(), f"poisson2d_{a.n}.mtx")
    pkg.matgen.write_mtx(mtx, pkg.matgen.poisson2d(a.n))
    out = {"matrix": f"2D 5-point Poisson {a.n}^2, -b {a.blocks}", "cores": os.cpu_count(), "rows": []}

    def run(target, mode, iters, env=None):
        best = None
        for _ in range(a.runs):
            code, text = O.run_cg_csr(exe, ["-f", mtx, "-b", a.blocks, "-c", "0", "-i", iters, "-t", target, "-m", mode],
                                      env_extra=env, timeout=1800)
            m = re.search(r"time taken =\s*([0-9.]+) ms", text)
            assert code == 0 and m, text[-500:]
            ms = float(m.group(1))
            best = ms if best is None else min(best, ms)
        return best

    for mode in a.modes.split(","):
        ms = run("cpu", mode, a.cpu_iters)
        out["rows"].append(dict(target="cpu", mode=mode, iters=a.cpu_iters, ms=ms, it_s=a.cpu_iters / ms * 1e3))
        for g in a.ngpus.split(","):
            ms = run("cuda", mode, a.iters, {"ABFT_CUDA_NGPUS": g})
            out["rows"].append(dict(target=f"cuda x{g}", mode=mode, iters=a.iters, ms=ms, it_s=a.iters / ms * 1e3))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
