"""Vendor-library yardstick for the four hot GEMM shapes (B=256): torch.matmul (cuBLAS, bf16, plain epilogue) against
the engine's tcgen05 kernels with their fused epilogues, both as SUSTAINED loops (thousands of back-to-back launches,
i.e. under the power cap) so the numbers are comparable with MEASURED_PEAKS.json `bf16_tflops_sustained`."""
import os
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_kernels_gpu import run_gemm  # noqa: E402

SHAPES = [("qkv", 3584, 3072, 1024, 0), ("c_proj", 3584, 1024, 1024, 1), ("expert up (SwiGLU)", 7168, 8192, 1024, 2),
          ("expert down", 7168, 1024, 4096, 3), ("square 8192", 8192, 8192, 8192, 3)]


def cublas(M, N, K, reps):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(reps // 4):
        torch.matmul(a, w.t(), out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, w.t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12


for name, M, N, K, epi in SHAPES:
    flops = 2.0 * M * N * K
    reps = max(200, int(0.6 / (flops / 1.2e15)))  # ~0.6 s of sustained work per measurement
    us, tf = cublas(M, N, K, reps)
    print(f"{name}: M={M} N={N} K={K}  cuBLAS sustained {us:.2f} us  {tf:.1f} TFLOP/s  ({reps} launches)", flush=True)
    os.environ["MODE_GEMM_BENCH_REPS"] = str(reps)
    run_gemm(M, N, K, epi, pair=True)
    run_gemm(M, N, K, epi, pair=False)
