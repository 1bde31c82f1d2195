#!/usr/bin/env python
"""Static instruction budgets of the spmv row walks, read from the SASS of a built library (no GPU needed).

    python tools/sass_report.py abft-sparse-cg_b200/libabftcuda.so build/variants/libabftcuda_uni.so ...

For every spmv_kernel<MODE, static> it finds the straight-line blocks that hold a whole unrolled row walk (many POPC for
the ECC modes, many LDS for the CRC mode) and prints their length and opcode mix: the per-row instruction count that,
with the pipe rates of profiles/r1_microbench_issue_rates.txt, bounds the check-heavy modes.
"""
import re
import subprocess
import sys
from collections import Counter

MODES = ["none", "constraints", "sed", "sec7", "sec8", "secded", "crc32c"]


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, d = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if cur and m:
            d[cur].append(m.group(1))
    return d


def blocks(ins):
    cur = []
    for t in ins:
        cur.append(t)
        if re.search(r"\b(BRA|BRX|EXIT|BSYNC|BAR|RET|CALL)\b", t):
            yield cur
            cur = []
    if cur:
        yield cur


def op(t):
    return re.sub(r"^@!?U?P\d+ ", "", t).split()[0].split(".")[0]


def main():
    for lib in sys.argv[1:]:
        print(f"== {lib}")
        ks = kernels(lib)
        for mode in (3, 5, 6):
            name = next((k for k in ks if f"abft11spmv_kernelILi{mode}ELb0" in k), None)
            if not name:
                continue
            ins = ks[name]
            print(f"  spmv_kernel<{MODES[mode]}, static>: {len(ins)} instructions")
            for b in blocks(ins):
                c = Counter(op(t) for t in b)
                heavy = c["POPC"] >= 28 if mode != 6 else c["LDS"] >= 40
                if heavy:
                    per = ", ".join(f"{k} {v}" for k, v in c.most_common(9))
                    print(f"    block of {len(b):4d}: {per}")


if __name__ == "__main__":
    main()
