#!/usr/bin/env python
"""The CSR kernels on the CPU emulator under AddressSanitizer or ThreadSanitizer (test infrastructure; no GPU).

    python tools/emu_sanitize.py --san address            # accesses outside the (padded) arrays
    python tools/emu_sanitize.py --san thread [-DABFT_SPMV_LASTREFILL=1 ...]   # data races between the CTA's threads

Builds tests/helpers/csr_emul.cpp (the product's spmv.cuh / encode.cuh / vecops.cuh against the multi-threaded CUDA shim)
with the sanitizer, re-executes itself with the sanitizer runtime preloaded, and runs clean and faulty spmv launches over
stencil, ragged, short-row, long-row and power-law matrices in four modes.  Exit status 0 and "0 reports" = clean.
"""
import argparse
import ctypes as C
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--san", default="address", choices=["address", "thread"])
    a, flags = ap.parse_known_args()
    rt = subprocess.run(["g++", f"-print-file-name=lib{'asan' if a.san == 'address' else 'tsan'}.so"], capture_output=True,
                        text=True).stdout.strip()
    if os.environ.get("ABFT_SANITIZE_CHILD") != "1":
        so = os.path.join(tempfile.mkdtemp(prefix="emu_san_"), "libcsr_emul_san.so")
        h = os.path.join(ROOT, "tests", "helpers")
        subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-gnu-unique",
                        f"-fsanitize={a.san}", "-fno-omit-frame-pointer", "-w", "-pthread", *flags,
                        "-I" + os.path.join(h, "cuda_shim_mt"), "-o", so, os.path.join(h, "csr_emul.cpp")], check=True)
        env = {**os.environ, "ABFT_SANITIZE_CHILD": "1", "ABFT_SANITIZE_LIB": so, "LD_PRELOAD": rt,
               "ASAN_OPTIONS": "detect_leaks=0", "TSAN_OPTIONS": "halt_on_error=0 report_signal_unsafe=0 history_size=4"}
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--san", a.san, *flags], env=env, capture_output=True, text=True)
        text = out.stdout + out.stderr
        reports = text.count("ERROR: AddressSanitizer") + text.count("WARNING: ThreadSanitizer")
        print(out.stdout.strip())
        print(f"{reports} reports (-fsanitize={a.san}{' ' + ' '.join(flags) if flags else ''})")
        if reports:
            print(text[-4000:])
        sys.exit(1 if (reports or out.returncode) else 0)

    # ---- child: sanitizer runtime is loaded
    import numpy as np
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import pkg_loader
    pkg = pkg_loader.load()
    spec = importlib.util.spec_from_file_location("tk", os.path.join(ROOT, "tests", "test_csr_kernel_logic.py"))
    tk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tk)
    L = C.CDLL(os.environ["ABFT_SANITIZE_LIB"])
    u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    L.emu_create_part.restype = C.c_void_p
    L.emu_create_part.argtypes = [C.c_int, u32p, u32p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.emu_destroy.argtypes = [C.c_void_p]
    L.emu_set_tiles.argtypes = [C.c_void_p, u32p, C.c_int, C.c_int, C.c_int]
    for f, t in ((L.emu_cols, C.c_uint32), (L.emu_rowptr, C.c_uint32), (L.emu_values, C.c_double)):
        f.restype = C.POINTER(t)
        f.argtypes = [C.c_void_p]
    L.emu_flip.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.emu_spmv.restype = C.c_int
    L.emu_spmv.argtypes = [C.c_void_p, f64p, f64p, C.POINTER(tk.Event), C.c_int, C.POINTER(C.c_double)]
    rng = np.random.default_rng(1)
    shapes = tk.shapes(pkg.matgen)
    names = ("p3d_12", "ragged", "short_rows", "long_rows", "powerlaw", "one_by_one") if a.san == "address" else ("p2d_40", "powerlaw", "long_rows")
    for name in names:
        t = shapes[name]()
        for mode in ("none", "constraints", "secded", "crc32c"):
            em = tk.EmuMatrix(L, pkg, mode, t)
            vec = rng.standard_normal(t.N)
            em.spmv(vec)
            if t.nnz > 5:
                em.flip(int(rng.integers(0, t.nnz)), int(rng.integers(0, 96)))
                em.spmv(vec)
            em.close()
        print(name, "done", flush=True)


if __name__ == "__main__":
    main()
