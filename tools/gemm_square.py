"""Main-loop ceiling probe: large square fp16 GEMMs (epilogue and tile-quantisation effects negligible) through the
pxr_test_gemm hook, single-CTA tiles vs CTA pairs, next to torch.matmul (cuBLAS) on the same shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gemm_util import run_gemm  # noqa: E402


def timeit(fn, n=10):
    """fn(repeat) enqueues `repeat` launches back to back (no host sync in between)."""
    fn(2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


for (M, N, K) in [(8192, 8192, 8192), (12608, 3072, 768), (12608, 3072, 6144), (12608, 768, 3072)]:
    A = (torch.randn(M, K, device="cuda") * 0.05).half()
    B = (torch.randn(N, K, device="cuda") * 0.05).half()
    out = torch.zeros(M, N, device="cuda", dtype=torch.half)
    fl = 2.0 * M * N * K
    line = f"M={M} N={N} K={K}:"
    for bn in (256, 192):
        for cg in (1, 2):
            t = timeit(lambda r: run_gemm(A, B, M, N, K, lda=K, ldb=K, block_n=bn, out_f16=out, ldc=N, cta_group=cg, repeat=r))
            line += f"  bn{bn}/cg{cg} {fl / t / 1e12:6.0f}"
    def mm(r):
        for _ in range(r):
            torch.matmul(A, B.t())
    t = timeit(mm)
    line += f"  cuBLAS {fl / t / 1e12:6.0f} TFLOP/s"
    print(line, flush=True)
