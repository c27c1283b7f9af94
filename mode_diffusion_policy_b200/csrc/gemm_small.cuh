// Small-M GEMM for rollout-sized batches (M = B*T <= 16 token rows per group: the reference's MoDEAgent.step runs the
// sampler at B = 1, mode_agent.py:584-636).
//
//   D[rows <= 16, N] = A[rows, K] (bf16)  x  W[N, K]^T (bf16, nn.Linear layout)     fp32 accumulate
//
// At this size the tensor-memory kernel (gemm.cuh) is all fixed cost: 256-row tiles that are 95 % padding, TMEM
// allocation, cluster barriers, and a K loop that a handful of CTAs walk alone (the K = 4096 down projection streams its
// weights through 8 CTA pairs). The work is a weight-streaming problem: every weight is read once and used for <= 16
// rows. Here every CTA owns 8 output columns (8 weight rows) and its 8 warps split K; each lane streams its weight row
// with 16-byte loads straight from global memory into mma.sync.m16n8k16 B fragments — a 32-wide K chunk is consumed by
// two MMAs with the k index permuted identically for A and B, so a lane's 16 contiguous bytes ARE its fragments — the A
// rows (<= 16 x K, L1/L2 resident) are loaded the same way, the 8 partial 16 x 8 tiles are summed through shared memory
// in a fixed order, and the epilogue is the same as the big kernel's (bias -> bf16, residual += in fp32, SwiGLU,
// plain bf16 / fp32). Work description is the same device-side M-tile table, so grouped expert GEMMs need no special
// case: one table entry per routed expert, rows_valid <= 16.
#pragma once
#include "gemm.cuh"

namespace mode {

constexpr int SMALL_M_MAX_ROWS = 16;
constexpr int SMALL_M_WARPS = 8;

struct SmallGemmParams {
  const __nv_bfloat16* A;      // [rows, K]
  const __nv_bfloat16* W;      // [weight rows, K]
  const GemmMTile* m_tiles;    // same table as the tensor-memory kernel (a_row0, out_row0, rows_valid <= 16, w_row_base)
  const int* num_m_tiles;      // device scalar
  const float* bias;           // indexed by weight row (packed like the weights), may be null
  void* out;                   // bf16 or fp32 [rows, ldo]
  int K, ldo;
  int w_row_off;
};

__device__ __forceinline__ void mma_bf16_16816_acc(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                   uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// grid (N / 8 column slabs [hidden units / 8 for SwiGLU], max M-tiles); 256 threads.
template <int EPI>
__global__ void __launch_bounds__(SMALL_M_WARPS * 32) gemm_small_m_kernel(const SmallGemmParams p) {
  pdl_trigger();
  constexpr bool GLU = (EPI == EPI_SWIGLU_BF16);
  constexpr int NACC = GLU ? 2 : 1;
  __shared__ float red[SMALL_M_WARPS][NACC][16 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  pdl_wait();
  if (static_cast<int>(blockIdx.y) >= *p.num_m_tiles) return;
  const GemmMTile tile = p.m_tiles[blockIdx.y];
  const int slab = blockIdx.x;  // 8 output columns (or hidden units)
  // weight rows streamed by this lane (row g of the slab; SwiGLU: the projected row and its gate row 128 further)
  int w_row = p.w_row_off + tile.w_row_base;
  if (GLU)
    w_row += (slab * 8 / 128) * 256 + (slab * 8) % 128 + g;
  else
    w_row += slab * 8 + g;
  const __nv_bfloat16* w0 = p.W + static_cast<size_t>(w_row) * p.K;
  const __nv_bfloat16* w1 = w0 + static_cast<size_t>(128) * p.K;  // gate row (GLU only)
  const __nv_bfloat16* a_lo = p.A + static_cast<size_t>(tile.a_row0 + g) * p.K;
  const __nv_bfloat16* a_hi = a_lo + static_cast<size_t>(8) * p.K;
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int k_per_warp = p.K / SMALL_M_WARPS;  // K is a multiple of 256: whole 32-wide chunks
  const int k0 = warp * k_per_warp + tq * 8;
#pragma unroll 4
  for (int kc = 0; kc < k_per_warp; kc += 32) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w0 + k0 + kc));
    const uint4 al = __ldg(reinterpret_cast<const uint4*>(a_lo + k0 + kc));
    const uint4 ah = __ldg(reinterpret_cast<const uint4*>(a_hi + k0 + kc));
    mma_bf16_16816_acc(acc[0], al.x, ah.x, al.y, ah.y, wv.x, wv.y);
    mma_bf16_16816_acc(acc[0], al.z, ah.z, al.w, ah.w, wv.z, wv.w);
    if (GLU) {
      const uint4 gv = __ldg(reinterpret_cast<const uint4*>(w1 + k0 + kc));
      mma_bf16_16816_acc(acc[NACC - 1], al.x, ah.x, al.y, ah.y, gv.x, gv.y);
      mma_bf16_16816_acc(acc[NACC - 1], al.z, ah.z, al.w, ah.w, gv.z, gv.w);
    }
  }
  // C fragment: c0,c1 = (row g, cols 2tq, 2tq+1), c2,c3 = (row g+8, same cols)
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    red[warp][i][g * 8 + 2 * tq] = acc[i][0];
    red[warp][i][g * 8 + 2 * tq + 1] = acc[i][1];
    red[warp][i][(g + 8) * 8 + 2 * tq] = acc[i][2];
    red[warp][i][(g + 8) * 8 + 2 * tq + 1] = acc[i][3];
  }
  __syncthreads();
  if (threadIdx.x >= 128) return;
  const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
  float v[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < SMALL_M_WARPS; ++w) s += red[w][i][r * 8 + c];  // fixed order: deterministic
    v[i] = s;
  }
  if (r >= tile.rows_valid) return;
  const size_t o = static_cast<size_t>(tile.out_row0 + r) * p.ldo + slab * 8 + c;
  const int wr = p.w_row_off + tile.w_row_base;
  if constexpr (EPI == EPI_BIAS_BF16) {
    reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v[0] + p.bias[wr + slab * 8 + c]);
  } else if constexpr (EPI == EPI_RESID_F32) {
    reinterpret_cast<float*>(p.out)[o] += v[0];
  } else if constexpr (EPI == EPI_SWIGLU_BF16) {
    const int prow = wr + (slab * 8 / 128) * 256 + (slab * 8) % 128 + c;
    const float proj = v[0] + p.bias[prow], gate = v[1] + p.bias[prow + 128];
    reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(proj * silu_f(gate));
  } else if constexpr (EPI == EPI_PLAIN_BF16) {
    reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v[0]);
  } else {
    reinterpret_cast<float*>(p.out)[o] = v[0];
  }
}

}  // namespace mode
