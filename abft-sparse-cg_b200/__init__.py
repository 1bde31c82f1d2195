"""abft-sparse-cg_b200 -- B200-native (sm_100a) CUDA target for the CSR CG hot
path of abft-sparse-cg.

The product is csrc/ (kernels + the C ABI of include/abft_cuda.h, built into
libabftcuda.so) and host/CUDAContext.cpp (the reference-side CGContext plugin).
The Python here is only a ctypes mirror of that ABI for tests and bench.py:

    capi      raw ctypes binding of every symbol in include/abft_cuda.h
    context   CUDAContext: same method names / argument meaning as CGContext
    matgen    synthetic Poisson / power-law matrices in cg.cpp's triplet order
    dist      row partition + exchange plan + the multi-GPU CG loop over the dist_* entry points

The directory name is not an importable identifier; load it with
``pkg_loader.load()`` at the repository root.
"""
from . import capi, context, dist, matgen  # noqa: F401
from .context import CUDAContext, FatalFault, MODES  # noqa: F401
