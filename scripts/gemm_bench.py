"""Times the tcgen05 GEMM on the hot-path shapes through mode_debug_gemm (MODE_GEMM_BENCH_REPS)."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
os.environ.setdefault("MODE_GEMM_BENCH_REPS", "20")
from test_kernels_gpu import run_gemm  # noqa: E402

for (M, N, K, epi) in [(3584, 3072, 1024, 0), (3584, 1024, 1024, 1), (3584, 16384, 1024, 2), (7168, 1024, 4096, 3),
                       (1792, 3072, 1024, 0), (1792, 16384, 1024, 2), (8192, 8192, 8192, 3)]:
    run_gemm(M, N, K, epi)
    run_gemm(M, N, K, epi, pair=True)
    run_gemm(M, N, K, epi, pair=True, stream_k=True)

# tile widths chosen by engine.cu choose_bn (wave filling on 74 CTA pairs) against the 256-wide tiling
for (M, N, K, epi) in [(3584, 3072, 1024, 0), (3584, 1024, 1024, 1), (7168, 1024, 4096, 3), (5120, 1024, 4096, 3),
                       (1792, 3072, 1024, 0), (1792, 1024, 1024, 1), (3584, 1024, 4096, 3)]:
    for bn in (256, 240, 224, 208, 192, 176, 160, 128):
        run_gemm(M, N, K, epi, pair=True, bn=bn)

# Is the pair kernel bound by L2->SM delivery? Same MMA work with 25 % fewer weight-tile loads (results are garbage).
for (M, N, K, epi) in [(3584, 16384, 1024, 2), (7168, 1024, 4096, 3), (8192, 8192, 8192, 3)]:
    run_gemm(M, N, K, epi, pair=True)
    print("  ^ normal   v every other weight-tile load skipped")
    run_gemm(M, N, K, epi, pair=True, skip_b=True)

# few M-tiles (B = 32 per GPU: 448 token rows, 896 routed rows): narrow tiles put all 74 CTA pairs to work
for (M, N, K, epi) in [(448, 3072, 1024, 0), (448, 1024, 1024, 1), (896, 1024, 4096, 3), (1792, 1024, 4096, 3)]:
    for bn in (256, 192, 128, 96, 64):
        run_gemm(M, N, K, epi, pair=True, bn=bn)
