"""Build libabftcuda_emu.so: the product's csrc/abft_cuda.cu -- host orchestration and kernels -- compiled by g++
against the multi-threaded CUDA shim and the fake runtime of tests/helpers/cuda_shim_mt/ (test infrastructure).

The only transformation of the source is mechanical: every `kernel<<<grid, block, smem, stream>>>(args);` launch of a
temporary copy becomes `EMU_LAUNCH(grid, block, smem, kernel(args));`.  The result exports the same C ABI, so
ABFT_CUDA_LIB=<it> lets the GPU tests' own code run (small cases) on a machine without a GPU.
"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _balanced(text, i, open_ch, close_ch):
    """text[i] == open_ch -> index just past the matching close_ch"""
    depth = 0
    while True:
        ch = text[i]
        if ch == open_ch:
            depth += 1
        elif ch == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def _split_top_level(text):
    """split at commas that are not inside () or <>"""
    parts, depth, cur = [], 0, []
    for i, ch in enumerate(text):
        if ch in "(<":
            depth += 1
        elif ch == ")" or (ch == ">" and text[i - 1:i] != "-"):      # (not the arrow of `c->stream`)
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


def rewrite_launches(src):
    out, pos, n = [], 0, 0
    for m in re.finditer(r"<<<", src):
        a = m.start()
        if a < pos:
            continue
        # kernel expression: back to the start of the statement
        k = a
        while src[k - 1] not in ";{}\n":
            k -= 1
        kernel = src[k:a].strip()
        b = src.index(">>>", a)
        config = src[a + 3:b]
        assert src[b + 3] == "(", src[b:b + 40]
        e = _balanced(src, b + 3, "(", ")")
        assert src[e] == ";", src[e - 20:e + 5]
        args = src[b + 4:e - 1]
        parts = _split_top_level(config)
        assert len(parts) == 4, (kernel, parts)
        indent = src[k:a][:len(src[k:a]) - len(src[k:a].lstrip())]
        out.append(src[pos:k])
        out.append(f"{indent}EMU_LAUNCH({parts[0]}, {parts[1]}, {parts[2]}, {kernel}({args}));")
        pos = e + 1
        n += 1
    out.append(src[pos:])
    return "".join(out), n


def build(out_dir, flags=()):
    csrc = os.path.join(ROOT, "abft-sparse-cg_b200", "csrc")
    with open(os.path.join(csrc, "abft_cuda.cu")) as f:
        text, n = rewrite_launches(f.read())
    assert n >= 20, n
    text = '#include "cuda_runtime.h"\n#include "emu_runtime.h"\n' + text
    tmp = os.path.join(str(out_dir), "abft_cuda_emu.cpp")
    with open(tmp, "w") as f:
        f.write(text)
    so = os.path.join(str(out_dir), "libabftcuda_emu.so")
    h = os.path.join(ROOT, "tests", "helpers", "cuda_shim_mt")
    subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-fno-gnu-unique", "-w", "-pthread",
                    *flags, "-I" + h, "-I" + csrc, "-o", so, tmp], check=True)
    return so


if __name__ == "__main__":
    import sys
    import tempfile
    print(build(tempfile.mkdtemp(prefix="abft_emu_"), sys.argv[1:]))
