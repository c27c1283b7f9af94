// Small-M GEMM for rollout-sized batches (M = B*T <= 32 token rows per group, i.e. B <= 2: the reference's
// MoDEAgent.step runs the sampler at B = 1, mode_agent.py:584-636).
//
//   D[rows <= 32, N] = A[rows, K] (bf16)  x  W[N, K]^T (bf16, nn.Linear layout)     fp32 accumulate
//
// At this size the tensor-memory kernel (gemm.cuh) is all fixed cost: 256-row tiles that are 95 % padding, TMEM
// allocation, cluster barriers, and a K loop that a handful of CTAs walk alone (the K = 4096 down projection streams its
// weights through 8 CTA pairs). The work is a weight-streaming problem: every weight is read once and used for <= 32
// rows (1 or 2 row tiles of 16 per CTA, all fed by the same weight fragments). Here every CTA owns 8 output columns (8 weight rows) and its 8 warps split K; each lane streams its weight row
// with 16-byte loads straight from global memory into mma.sync.m16n8k16 B fragments — a 32-wide K chunk is consumed by
// two MMAs with the k index permuted identically for A and B, so a lane's 16 contiguous bytes ARE its fragments — the A
// rows (<= 16 x K, L1/L2 resident) are loaded the same way, the 8 partial 16 x 8 tiles are summed through shared memory
// in a fixed order, and the epilogue is the same as the big kernel's (bias -> bf16, residual += in fp32, SwiGLU,
// plain bf16 / fp32). Work description is the same device-side M-tile table, so grouped expert GEMMs need no special
// case: one table entry per routed expert, rows_valid <= 32.
#pragma once
#include "gemm.cuh"

namespace mode {

constexpr int SMALL_M_MAX_ROWS = 32;  // up to 2 row tiles of 16 per group (B <= 2 at T = 14); measured: at 64 rows the
                                      // per-CTA re-reads of the A rows cost more than the tensor-memory kernel's fixed overhead
constexpr int SMALL_M_WARPS = 8;

struct SmallGemmParams {
  const __nv_bfloat16* A;      // [rows, K]
  const __nv_bfloat16* W;      // [weight rows, K]
  const GemmMTile* m_tiles;    // same table as the tensor-memory kernel (a_row0, out_row0, rows_valid <= 32, w_row_base)
  const int* num_m_tiles;      // device scalar
  const float* bias;           // indexed by weight row (packed like the weights), may be null
  void* out;                   // bf16 or fp32 [rows, ldo]
  int K, ldo;
  int w_row_off;
  // 1: every CTA asks L2 for its own weight rows BEFORE waiting for the previous kernel (programmatic dependent launch:
  // the CTAs are resident while a 14-row norm / attention kernel runs on a handful of SMs, and the weights do not depend
  // on it), so the HBM stream of this GEMM overlaps its predecessor and the K loop reads L2. Only when the M-tile table
  // is final before the launch chain starts (dense projections; expert GEMMs inside a pre-routed sampler graph).
  int prefetch = 0;
};

__device__ __forceinline__ void mma_bf16_16816_acc(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                   uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Body for one (slab, group) task. `red_raw`: SMALL_M_WARPS * NACC * MT*128 floats of shared memory. COHERENT_A: the A rows
// were written earlier in the SAME kernel launch (persistent small-batch kernel, small_eval.cuh) and must not be read
// through the non-coherent path; weights always are. All threads of the CTA call this together (it synchronises).
template <int EPI, int MT, bool COHERENT_A>
__device__ __forceinline__ void gemm_small_body(const SmallGemmParams& p, int slab, int group, float* red_raw) {
  constexpr bool GLU = (EPI == EPI_SWIGLU_BF16);
  constexpr int NACC = GLU ? 2 : 1;
  float (*red)[NACC][MT * 16 * 8] = reinterpret_cast<float (*)[NACC][MT * 16 * 8]>(red_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  if (group >= *p.num_m_tiles) return;
  const GemmMTile tile = p.m_tiles[group];
  // weight rows streamed by this lane (row g of the slab; SwiGLU: the projected row and its gate row 128 further)
  int w_row = p.w_row_off + tile.w_row_base;
  if (GLU)
    w_row += (slab * 8 / 128) * 256 + (slab * 8) % 128 + g;
  else
    w_row += slab * 8 + g;
  const __nv_bfloat16* w0 = p.W + static_cast<size_t>(w_row) * p.K;
  const __nv_bfloat16* w1 = w0 + static_cast<size_t>(128) * p.K;  // gate row (GLU only)
  const __nv_bfloat16* a_lo = p.A + static_cast<size_t>(tile.a_row0 + g) * p.K;  // row tile m: + m * 16 rows
  float acc[MT][NACC][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[m][i][0] = acc[m][i][1] = acc[m][i][2] = acc[m][i][3] = 0.f;
  const int k_per_warp = p.K / SMALL_M_WARPS;  // K is a multiple of 256: whole 32-wide chunks
  const int k0 = warp * k_per_warp + tq * 8;
#pragma unroll(MT == 1 ? 4 : 2)
  for (int kc = 0; kc < k_per_warp; kc += 32) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w0 + k0 + kc));
    uint4 gv = make_uint4(0, 0, 0, 0);
    if (GLU) gv = __ldg(reinterpret_cast<const uint4*>(w1 + k0 + kc));
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const uint4* pl = reinterpret_cast<const uint4*>(a_lo + static_cast<size_t>(m * 16) * p.K + k0 + kc);
      const uint4* ph = reinterpret_cast<const uint4*>(a_lo + static_cast<size_t>(m * 16 + 8) * p.K + k0 + kc);
      const uint4 al = COHERENT_A ? *pl : __ldg(pl);
      const uint4 ah = COHERENT_A ? *ph : __ldg(ph);
      mma_bf16_16816_acc(acc[m][0], al.x, ah.x, al.y, ah.y, wv.x, wv.y);
      mma_bf16_16816_acc(acc[m][0], al.z, ah.z, al.w, ah.w, wv.z, wv.w);
      if (GLU) {
        mma_bf16_16816_acc(acc[m][NACC - 1], al.x, ah.x, al.y, ah.y, gv.x, gv.y);
        mma_bf16_16816_acc(acc[m][NACC - 1], al.z, ah.z, al.w, ah.w, gv.z, gv.w);
      }
    }
  }
  // C fragment: c0,c1 = (row g, cols 2tq, 2tq+1), c2,c3 = (row g+8, same cols)
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float* dst = red[warp][i] + m * 128;
      dst[g * 8 + 2 * tq] = acc[m][i][0];
      dst[g * 8 + 2 * tq + 1] = acc[m][i][1];
      dst[(g + 8) * 8 + 2 * tq] = acc[m][i][2];
      dst[(g + 8) * 8 + 2 * tq + 1] = acc[m][i][3];
    }
  __syncthreads();
  const int wr = p.w_row_off + tile.w_row_base;
  // MT * 128 outputs, 256 threads: thread t finishes outputs t, t + 256, ...
  for (int idx = threadIdx.x; idx < MT * 128; idx += SMALL_M_WARPS * 32) {
  const int r = idx >> 3, c = idx & 7;
  float v[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < SMALL_M_WARPS; ++w) s += red[w][i][idx];  // fixed order: deterministic
    v[i] = s;
  }
  if (r >= tile.rows_valid) continue;
  const size_t o = static_cast<size_t>(tile.out_row0 + r) * p.ldo + slab * 8 + c;
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
  __syncthreads();  // `red` is reused by the caller's next task
}

// S slabs per task (persistent small-batch kernel): the same arithmetic as gemm_small_body for slabs
// [slab0, slab0 + S) of one group — same K split over the warps, same chunk order, same fixed-order reduction, hence
// bit-identical results — but every lane keeps S (x2 for SwiGLU) independent 16-byte weight loads in flight per K chunk.
// With ONE resident CTA per SM (the persistent kernel's register budget) that is what brings the streamed bytes in
// flight per SM to 48-64 KB, i.e. to what HBM latency x bandwidth / 148 SMs asks for. A rows are always read coherently.
template <int EPI, int MT, int S>
__device__ __forceinline__ void gemm_small_multi_body(const SmallGemmParams& p, int slab0, int n_slabs, int group, float* red_raw) {
  constexpr bool GLU = (EPI == EPI_SWIGLU_BF16);
  constexpr int NACC = GLU ? 2 : 1;
  float (*red)[S][NACC][MT * 16 * 8] = reinterpret_cast<float (*)[S][NACC][MT * 16 * 8]>(red_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  if (group >= *p.num_m_tiles) return;
  const GemmMTile tile = p.m_tiles[group];
  const int wr = p.w_row_off + tile.w_row_base;
  const __nv_bfloat16* w0[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int slab = min(slab0 + s, n_slabs - 1);  // a task past the last slab re-reads it and stores nothing
    const int row = GLU ? (slab * 8 / 128) * 256 + (slab * 8) % 128 + g : slab * 8 + g;
    w0[s] = p.W + static_cast<size_t>(wr + row) * p.K;
  }
  const __nv_bfloat16* a_lo = p.A + static_cast<size_t>(tile.a_row0 + g) * p.K;
  float acc[S][MT][NACC][4];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < NACC; ++i) acc[s][m][i][0] = acc[s][m][i][1] = acc[s][m][i][2] = acc[s][m][i][3] = 0.f;
  const int k_per_warp = p.K / SMALL_M_WARPS;
  const int k0 = warp * k_per_warp + tq * 8;
#pragma unroll(MT == 1 ? 4 : 2)
  for (int kc = 0; kc < k_per_warp; kc += 32) {
    uint4 wv[S], gv[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      wv[s] = __ldg(reinterpret_cast<const uint4*>(w0[s] + k0 + kc));
      if (GLU) gv[s] = __ldg(reinterpret_cast<const uint4*>(w0[s] + static_cast<size_t>(128) * p.K + k0 + kc));
    }
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const uint4 al = *reinterpret_cast<const uint4*>(a_lo + static_cast<size_t>(m * 16) * p.K + k0 + kc);
      const uint4 ah = *reinterpret_cast<const uint4*>(a_lo + static_cast<size_t>(m * 16 + 8) * p.K + k0 + kc);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        mma_bf16_16816_acc(acc[s][m][0], al.x, ah.x, al.y, ah.y, wv[s].x, wv[s].y);
        mma_bf16_16816_acc(acc[s][m][0], al.z, ah.z, al.w, ah.w, wv[s].z, wv[s].w);
        if (GLU) {
          mma_bf16_16816_acc(acc[s][m][NACC - 1], al.x, ah.x, al.y, ah.y, gv[s].x, gv[s].y);
          mma_bf16_16816_acc(acc[s][m][NACC - 1], al.z, ah.z, al.w, ah.w, gv[s].z, gv[s].w);
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        float* dst = red[warp][s][i] + m * 128;
        dst[g * 8 + 2 * tq] = acc[s][m][i][0];
        dst[g * 8 + 2 * tq + 1] = acc[s][m][i][1];
        dst[(g + 8) * 8 + 2 * tq] = acc[s][m][i][2];
        dst[(g + 8) * 8 + 2 * tq + 1] = acc[s][m][i][3];
      }
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * MT * 128; idx += SMALL_M_WARPS * 32) {
    const int s = idx / (MT * 128), e = idx % (MT * 128);
    const int slab = slab0 + s;
    const int r = e >> 3, c = e & 7;
    float v[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < SMALL_M_WARPS; ++w) sum += red[w][s][i][e];  // fixed order: deterministic
      v[i] = sum;
    }
    if (r >= tile.rows_valid || slab >= n_slabs) continue;
    const size_t o = static_cast<size_t>(tile.out_row0 + r) * p.ldo + slab * 8 + c;
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
  __syncthreads();  // `red` is reused by the caller's next task
}

// grid (N / 8 column slabs [hidden units / 8 for SwiGLU], max M-tiles); 256 threads. MT = 16-row tiles per group
// (1 or 2): every weight fragment a lane loads is used for all MT row tiles.
template <int EPI, int MT>
__global__ void __launch_bounds__(SMALL_M_WARPS * 32) gemm_small_m_kernel(const SmallGemmParams p) {
  pdl_trigger();
  __shared__ float red[SMALL_M_WARPS * (EPI == EPI_SWIGLU_BF16 ? 2 : 1) * MT * 16 * 8];
  if (p.prefetch && static_cast<int>(blockIdx.y) < *p.num_m_tiles) {
    constexpr bool GLU = (EPI == EPI_SWIGLU_BF16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slab = blockIdx.x;
    const int row0 = p.w_row_off + p.m_tiles[blockIdx.y].w_row_base + (GLU ? (slab * 8 / 128) * 256 + (slab * 8) % 128 : slab * 8);
    const int k_per_warp = p.K / SMALL_M_WARPS;
    const int lines = k_per_warp / 64;  // 128-byte lines of one weight row inside this warp's K slice
    for (int i = lane; i < 8 * lines * (GLU ? 2 : 1); i += 32) {
      const int row = row0 + (i / lines) % 8 + (i / (8 * lines)) * 128;
      const __nv_bfloat16* src = p.W + static_cast<size_t>(row) * p.K + warp * k_per_warp + (i % lines) * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
    }
  }
  pdl_wait();
  gemm_small_body<EPI, MT, false>(p, blockIdx.x, blockIdx.y, red);
}

}  // namespace mode
