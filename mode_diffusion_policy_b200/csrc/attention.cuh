// Fused causal self-attention for the MoDE token sequence (T <= 64 tokens: sigma | goal | images | actions).
// Reference: Attention.forward, mode/models/networks/modedit.py:133-167 — per-head RMSNorm on q and k (:145-146),
// F.scaled_dot_product_attention(is_causal=True) (:149), heads re-interleaved (:165).
//
// One warp per (sample, head). q/k/v rows are read once from the packed QKV activation (bf16, 16-byte vectors),
// q and k are RMS-normalised in fp32 and rounded to bf16 in shared memory, QK^T and PV run on mma.sync m16n8k16
// (0.05% of the step's FLOPs: the tensor-memory path would be all overhead for 14x14 problems), softmax in fp32.
// Rounding points follow the flash SDPA kernel the reference dispatches to under bf16 autocast: bf16 q/k/v, fp32
// scores, P rounded to bf16 before PV, fp32 row sum, bf16 output.
#pragma once
#include "ptx.cuh"
#include "rng.cuh"

namespace mode {

constexpr int ATTN_WARPS = 4;
constexpr int ATTN_MAX_TPAD = 64;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct alignas(64) AttnParams {
  // TMA view of qkv for attention_tma_kernel: [rows, 3*d] bf16, box {64 columns, 16*MT rows}, SWIZZLE_128B
  CUtensorMap tmap_qkv;
  const __nv_bfloat16* qkv;  // [B*T, 3*d]: q | k | v, each d = H*Dh wide
  __nv_bfloat16* out;        // [B*T, d]
  const float* q_gain;       // [Dh]
  const float* k_gain;       // [Dh]
  int B, T, H;
  float eps;                 // RMSNorm eps (1e-6)
  float inv_sqrt_dh;         // float(Dh ** -0.5): RMSNorm scale (modedit.py:75) and the SDPA softmax scale
  // training only: dropout on the attention probabilities (SDPA dropout_p, modedit.py:149); thr == 0 disables it.
  // Element (b, h, row, col) uses half (col & 1) of word ((b*H + h)*T + row) * ceil(T/2) + col/2 of stream RNG_ATTN.
  DropoutSpec drop = DropoutSpec{0u, 0u, 1.0f};
};

// DH: head dim (32/64/128). MT: number of 16-row tiles covering T (T_pad = 16*MT).
// Body of attention_kernel for virtual block `vblock` with `n_warps` warps per CTA and `attn_smem` holding
// n_warps * 3 * 16*MT * (DH + 8) bf16 (also a phase of the persistent small-batch kernel, small_eval.cuh).
template <int DH, int MT>
__device__ __forceinline__ void attention_body(const AttnParams& p, int vblock, int n_warps, uint8_t* attn_smem) {
  constexpr int TPAD = 16 * MT;
  constexpr int LDS = DH + 8;            // padded row (bf16 elements): conflict-free ldmatrix
  constexpr int VPR = DH / 8;            // 16-byte vectors per row
  constexpr int ROWS_PER_IT = 32 / VPR;  // rows covered by one warp-wide vector load
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = vblock * n_warps + warp;  // (b, h)
  if (item >= p.B * p.H) return;
  const int b = item / p.H, h = item % p.H;
  const int T = p.T, d = p.H * DH;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(attn_smem) + static_cast<size_t>(warp) * 3 * TPAD * LDS;
  __nv_bfloat16* sK = sQ + TPAD * LDS;
  __nv_bfloat16* sV = sK + TPAD * LDS;

  // ---- stage q, k, v into shared memory: all rows are requested up front with cp.async (no register staging, the
  //      whole 3 x T x Dh tile is in flight at once), then q and k are RMS-normalised in place
  const int sub = lane % VPR;  // vector index inside the row
  const float inv_sqrt_dh = p.inv_sqrt_dh;
  for (int r0 = 0; r0 < TPAD; r0 += ROWS_PER_IT) {
    const int row = r0 + lane / VPR;
    const __nv_bfloat16* src = p.qkv + (static_cast<size_t>(b) * T + row) * 3 * d + h * DH + sub * 8;
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      __nv_bfloat16* dst = (which == 0 ? sQ : which == 1 ? sK : sV) + row * LDS + sub * 8;
      if (row < T)
        cp_async_16(smem_u32(dst), src + which * d);
      else
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_all();
  __syncwarp();
  float gq[8], gk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gq[j] = p.q_gain[sub * 8 + j];
    gk[j] = p.k_gain[sub * 8 + j];
  }
  for (int r0 = 0; r0 < TPAD; r0 += ROWS_PER_IT) {
    const int row = r0 + lane / VPR;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* ptr = (which == 0 ? sQ : sK) + row * LDS + sub * 8;
      const uint4 raw = *reinterpret_cast<const uint4*>(ptr);
      float v[8];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
        ss += f.x * f.x + f.y * f.y;
      }
#pragma unroll
      for (int o = VPR / 2; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      // RMSNorm.forward (modedit.py:78-80): x / clamp(||x|| * dim^-0.5, eps) * g   (one division per row)
      const float rn = 1.0f / fmaxf(sqrtf(ss) * inv_sqrt_dh, p.eps);
      const float* g = which == 0 ? gq : gk;
      uint4 o4;
      o4.x = pack_bf16x2((v[0] * rn) * g[0], (v[1] * rn) * g[1]);
      o4.y = pack_bf16x2((v[2] * rn) * g[2], (v[3] * rn) * g[3]);
      o4.z = pack_bf16x2((v[4] * rn) * g[4], (v[5] * rn) * g[5]);
      o4.w = pack_bf16x2((v[6] * rn) * g[6], (v[7] * rn) * g[7]);
      *reinterpret_cast<uint4*>(ptr) = o4;
    }
  }
  __syncwarp();

  const int g = lane >> 2, tq = lane & 3;  // mma fragment coordinates
  const float scale = inv_sqrt_dh;         // SDPA default scale 1/sqrt(Dh)
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);

#pragma unroll 1
  for (int mi = 0; mi < MT; ++mi) {
    if (mi * 16 >= T) break;
    // ---- S = Q K^T for this 16-row tile; only key tiles nj <= 2*mi+1 can be unmasked
    float s[2 * MT][4];
#pragma unroll
    for (int nj = 0; nj < 2 * MT; ++nj) s[nj][0] = s[nj][1] = s[nj][2] = s[nj][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldmatrix_x4(sQ_a + ((mi * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8) * 2, a0, a1, a2, a3);
#pragma unroll
      for (int nj = 0; nj < 2 * MT; ++nj) {
        if (nj <= 2 * mi + 1) {
          uint32_t b0, b1;
          ldmatrix_x2(sK_a + ((nj * 8 + (lane & 7)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8) * 2, b0, b1);
          mma_bf16_16816(s[nj], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    // ---- causal softmax (fp32); rows g and g+8 of the tile live in this quad
    const int row_lo = mi * 16 + g, row_hi = row_lo + 8;
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int nj = 0; nj < 2 * MT; ++nj) {
      if (nj <= 2 * mi + 1) {
        const int c0 = nj * 8 + 2 * tq;
        s[nj][0] = (c0 <= row_lo) ? s[nj][0] * scale : -INFINITY;
        s[nj][1] = (c0 + 1 <= row_lo) ? s[nj][1] * scale : -INFINITY;
        s[nj][2] = (c0 <= row_hi) ? s[nj][2] * scale : -INFINITY;
        s[nj][3] = (c0 + 1 <= row_hi) ? s[nj][3] * scale : -INFINITY;
        m_lo = fmaxf(m_lo, fmaxf(s[nj][0], s[nj][1]));
        m_hi = fmaxf(m_hi, fmaxf(s[nj][2], s[nj][3]));
      }
    }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int nj = 0; nj < 2 * MT; ++nj) {
      if (nj <= 2 * mi + 1) {
        s[nj][0] = __expf(s[nj][0] - m_lo);
        s[nj][1] = __expf(s[nj][1] - m_lo);
        s[nj][2] = __expf(s[nj][2] - m_hi);
        s[nj][3] = __expf(s[nj][3] - m_hi);
        sum_lo += s[nj][0] + s[nj][1];
        sum_hi += s[nj][2] + s[nj][3];
      }
    }
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    if (p.drop.thr) {  // zero the dropped probabilities; the row sums above stay those of the full softmax
      const uint32_t thalf = static_cast<uint32_t>(T + 1) >> 1;
      const uint32_t w_lo = (static_cast<uint32_t>(item) * T + row_lo) * thalf, w_hi = (static_cast<uint32_t>(item) * T + row_hi) * thalf;
#pragma unroll
      for (int nj = 0; nj < 2 * MT; ++nj) {
        if (nj <= 2 * mi + 1) {
          const uint32_t cw = static_cast<uint32_t>(nj * 4 + tq);  // (nj*8 + 2*tq) / 2
          const uint32_t b_lo = rng_bits(p.drop.key, w_lo + cw), b_hi = rng_bits(p.drop.key, w_hi + cw);
          if ((b_lo & 0xffffu) < p.drop.thr) s[nj][0] = 0.f;
          if ((b_lo >> 16) < p.drop.thr) s[nj][1] = 0.f;
          if ((b_hi & 0xffffu) < p.drop.thr) s[nj][2] = 0.f;
          if ((b_hi >> 16) < p.drop.thr) s[nj][3] = 0.f;
        }
      }
    }

    // ---- O = P V (P rounded to bf16, unnormalised; divide by the fp32 row sum at the end)
    float o[DH / 8][4];
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kj = 0; kj < MT; ++kj) {
      if (kj <= mi) {
        const uint32_t a0 = pack_bf16x2(s[2 * kj][0], s[2 * kj][1]);
        const uint32_t a1 = pack_bf16x2(s[2 * kj][2], s[2 * kj][3]);
        const uint32_t a2 = pack_bf16x2(s[2 * kj + 1][0], s[2 * kj + 1][1]);
        const uint32_t a3 = pack_bf16x2(s[2 * kj + 1][2], s[2 * kj + 1][3]);
#pragma unroll
        for (int dn = 0; dn < DH / 8; ++dn) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(sV_a + ((kj * 16 + (lane & 15)) * LDS + dn * 8) * 2, b0, b1);
          mma_bf16_16816(o[dn], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    const float keep_scale = p.drop.thr ? p.drop.scale : 1.0f;
    const float inv_lo = keep_scale / sum_lo, inv_hi = keep_scale / sum_hi;
    // ---- stage O through this tile's (now dead) Q rows so the global store is 16-byte coalesced
    __syncwarp();
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) {
      *reinterpret_cast<uint32_t*>(sQ + row_lo * LDS + dn * 8 + 2 * tq) =
          pack_bf16x2(o[dn][0] * inv_lo, o[dn][1] * inv_lo);
      *reinterpret_cast<uint32_t*>(sQ + row_hi * LDS + dn * 8 + 2 * tq) =
          pack_bf16x2(o[dn][2] * inv_hi, o[dn][3] * inv_hi);
    }
    __syncwarp();
  }
  // ---- write out [T, Dh] for this head
  for (int r0 = 0; r0 < TPAD; r0 += ROWS_PER_IT) {
    const int row = r0 + lane / VPR;
    if (row < T) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + row * LDS + sub * 8);
      *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * T + row) * d + h * DH + sub * 8) = v;
    }
  }
}

template <int DH, int MT>
__global__ void __launch_bounds__(ATTN_WARPS * 32) attention_kernel(const AttnParams p) {
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t attn_smem[];
  pdl_wait();
  attention_body<DH, MT>(p, blockIdx.x, ATTN_WARPS, attn_smem);
}

// ------------------------------------------------------------------------------------------------------------------
// TMA-staged variant (head dims 64 and 128; the default). Same math and rounding points as attention_kernel above; what
// changes is how the (sample, head) tile reaches shared memory and how it is laid out there:
//   * one elected lane per warp issues 3 * DH/64 bulk tensor loads (cp.async.bulk.tensor, SASS UTMALDG) of
//     [16*MT rows x 64 columns] boxes of q, k and v against the warp's own mbarrier (expect_tx = the tile's bytes); no
//     thread moves data through registers and nothing is padded: rows past T are the next sample's rows (finite, masked
//     by causality) or zero fill past the end of the tensor;
//   * the boxes land in the 128-byte-swizzle layout (16-byte chunk c of row r at chunk c ^ (r & 7)), which makes the
//     ldmatrix reads of 8 consecutive rows conflict-free without the 8-element row padding of the cp.async version.
// smem address of element (row, col) of an operand whose first half-tile starts at `base` (1024-byte aligned):
template <int TPAD>
__device__ __forceinline__ uint32_t attn_sw128(uint32_t base, int row, int col) {
  return base + static_cast<uint32_t>(col >> 6) * (TPAD * 128) + static_cast<uint32_t>(row) * 128u +
         (static_cast<uint32_t>(((col & 63) >> 3) ^ (row & 7)) << 4) + static_cast<uint32_t>(col & 7) * 2u;
}

template <int DH, int MT>
__global__ void __launch_bounds__(ATTN_WARPS * 32) attention_tma_kernel(const __grid_constant__ AttnParams p) {
  pdl_trigger();
  constexpr int TPAD = 16 * MT;
  constexpr int VPR = DH / 8;            // 16-byte vectors per row
  constexpr int ROWS_PER_IT = 32 / VPR;  // rows covered by one warp-wide vector access
  constexpr int NH = DH / 64;            // 64-column boxes per operand
  constexpr uint32_t OP_BYTES = TPAD * DH * 2;
  extern __shared__ __align__(1024) uint8_t attn_tma_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * ATTN_WARPS + warp;  // (b, h)
  // [ATTN_WARPS][3 operands][OP_BYTES] then one mbarrier per warp
  const uint32_t smem_base = smem_u32(attn_tma_smem);
  if ((smem_base & 1023u) != 0) __trap();  // the swizzle pattern is a function of the address bits
  const uint32_t sQ_a = smem_base + static_cast<uint32_t>(warp) * 3u * OP_BYTES, sK_a = sQ_a + OP_BYTES, sV_a = sK_a + OP_BYTES;
  const uint32_t bar = smem_base + ATTN_WARPS * 3u * OP_BYTES + 8u * warp;
  if (item >= p.B * p.H) return;
  const int b = item / p.H, h = item % p.H;
  const int T = p.T, d = p.H * DH;
  if (lane == 0) {
    tma_prefetch_desc(&p.tmap_qkv);
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  pdl_wait();
  if (lane == 0) {
    mbar_arrive_expect_tx(bar, 3u * OP_BYTES);
#pragma unroll
    for (int which = 0; which < 3; ++which)
#pragma unroll
      for (int hf = 0; hf < NH; ++hf)
        tma_load_2d(sQ_a + which * OP_BYTES + hf * (TPAD * 128), &p.tmap_qkv, bar, which * d + h * DH + hf * 64, b * T);
  }
  const int sub = lane % VPR;
  const float inv_sqrt_dh = p.inv_sqrt_dh;
  float gq[8], gk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // the gain loads overlap the tensor loads
    gq[j] = p.q_gain[sub * 8 + j];
    gk[j] = p.k_gain[sub * 8 + j];
  }
  mbar_wait(bar, 0);
  // ---- per-head RMSNorm of q and k in place (fp32 math, bf16 storage)
  for (int r0 = 0; r0 < TPAD; r0 += ROWS_PER_IT) {
    const int row = r0 + lane / VPR;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t a = attn_sw128<TPAD>(which == 0 ? sQ_a : sK_a, row, sub * 8);
      uint4 raw;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(a));
      float v[8];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
        ss += f.x * f.x + f.y * f.y;
      }
#pragma unroll
      for (int o = VPR / 2; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rn = 1.0f / fmaxf(sqrtf(ss) * inv_sqrt_dh, p.eps);  // RMSNorm.forward, modedit.py:78-80
      const float* g = which == 0 ? gq : gk;
      st_shared_v4(a, pack_bf16x2((v[0] * rn) * g[0], (v[1] * rn) * g[1]), pack_bf16x2((v[2] * rn) * g[2], (v[3] * rn) * g[3]),
                   pack_bf16x2((v[4] * rn) * g[4], (v[5] * rn) * g[5]), pack_bf16x2((v[6] * rn) * g[6], (v[7] * rn) * g[7]));
    }
  }
  __syncwarp();

  const int g = lane >> 2, tq = lane & 3;
  const float scale = inv_sqrt_dh;
#pragma unroll 1
  for (int mi = 0; mi < MT; ++mi) {
    if (mi * 16 >= T) break;
    float s[2 * MT][4];
#pragma unroll
    for (int nj = 0; nj < 2 * MT; ++nj) s[nj][0] = s[nj][1] = s[nj][2] = s[nj][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t a0, a1, a2, a3;
      ldmatrix_x4(attn_sw128<TPAD>(sQ_a, mi * 16 + (lane & 15), kk * 16 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
      for (int nj = 0; nj < 2 * MT; ++nj) {
        if (nj <= 2 * mi + 1) {
          uint32_t b0, b1;
          ldmatrix_x2(attn_sw128<TPAD>(sK_a, nj * 8 + (lane & 7), kk * 16 + ((lane >> 3) & 1) * 8), b0, b1);
          mma_bf16_16816(s[nj], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    const int row_lo = mi * 16 + g, row_hi = row_lo + 8;
    float m_lo = -INFINITY, m_hi = -INFINITY;
#pragma unroll
    for (int nj = 0; nj < 2 * MT; ++nj) {
      if (nj <= 2 * mi + 1) {
        const int c0 = nj * 8 + 2 * tq;
        s[nj][0] = (c0 <= row_lo) ? s[nj][0] * scale : -INFINITY;
        s[nj][1] = (c0 + 1 <= row_lo) ? s[nj][1] * scale : -INFINITY;
        s[nj][2] = (c0 <= row_hi) ? s[nj][2] * scale : -INFINITY;
        s[nj][3] = (c0 + 1 <= row_hi) ? s[nj][3] * scale : -INFINITY;
        m_lo = fmaxf(m_lo, fmaxf(s[nj][0], s[nj][1]));
        m_hi = fmaxf(m_hi, fmaxf(s[nj][2], s[nj][3]));
      }
    }
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 1));
    m_lo = fmaxf(m_lo, __shfl_xor_sync(0xffffffffu, m_lo, 2));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 1));
    m_hi = fmaxf(m_hi, __shfl_xor_sync(0xffffffffu, m_hi, 2));
    float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
    for (int nj = 0; nj < 2 * MT; ++nj) {
      if (nj <= 2 * mi + 1) {
        s[nj][0] = __expf(s[nj][0] - m_lo);
        s[nj][1] = __expf(s[nj][1] - m_lo);
        s[nj][2] = __expf(s[nj][2] - m_hi);
        s[nj][3] = __expf(s[nj][3] - m_hi);
        sum_lo += s[nj][0] + s[nj][1];
        sum_hi += s[nj][2] + s[nj][3];
      }
    }
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
    sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
    sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
    if (p.drop.thr) {
      const uint32_t thalf = static_cast<uint32_t>(T + 1) >> 1;
      const uint32_t w_lo = (static_cast<uint32_t>(item) * T + row_lo) * thalf, w_hi = (static_cast<uint32_t>(item) * T + row_hi) * thalf;
#pragma unroll
      for (int nj = 0; nj < 2 * MT; ++nj) {
        if (nj <= 2 * mi + 1) {
          const uint32_t cw = static_cast<uint32_t>(nj * 4 + tq);
          const uint32_t b_lo = rng_bits(p.drop.key, w_lo + cw), b_hi = rng_bits(p.drop.key, w_hi + cw);
          if ((b_lo & 0xffffu) < p.drop.thr) s[nj][0] = 0.f;
          if ((b_lo >> 16) < p.drop.thr) s[nj][1] = 0.f;
          if ((b_hi & 0xffffu) < p.drop.thr) s[nj][2] = 0.f;
          if ((b_hi >> 16) < p.drop.thr) s[nj][3] = 0.f;
        }
      }
    }
    float o[DH / 8][4];
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kj = 0; kj < MT; ++kj) {
      if (kj <= mi) {
        const uint32_t a0 = pack_bf16x2(s[2 * kj][0], s[2 * kj][1]);
        const uint32_t a1 = pack_bf16x2(s[2 * kj][2], s[2 * kj][3]);
        const uint32_t a2 = pack_bf16x2(s[2 * kj + 1][0], s[2 * kj + 1][1]);
        const uint32_t a3 = pack_bf16x2(s[2 * kj + 1][2], s[2 * kj + 1][3]);
#pragma unroll
        for (int dn = 0; dn < DH / 8; ++dn) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(attn_sw128<TPAD>(sV_a, kj * 16 + (lane & 15), dn * 8), b0, b1);
          mma_bf16_16816(o[dn], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    const float keep_scale = p.drop.thr ? p.drop.scale : 1.0f;
    const float inv_lo = keep_scale / sum_lo, inv_hi = keep_scale / sum_hi;
    // stage O through this tile's (now dead) Q rows so the global store is 16-byte coalesced
    __syncwarp();
#pragma unroll
    for (int dn = 0; dn < DH / 8; ++dn) {
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(attn_sw128<TPAD>(sQ_a, row_lo, dn * 8 + 2 * tq)),
                   "r"(pack_bf16x2(o[dn][0] * inv_lo, o[dn][1] * inv_lo)) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(attn_sw128<TPAD>(sQ_a, row_hi, dn * 8 + 2 * tq)),
                   "r"(pack_bf16x2(o[dn][2] * inv_hi, o[dn][3] * inv_hi)) : "memory");
    }
    __syncwarp();
  }
  for (int r0 = 0; r0 < TPAD; r0 += ROWS_PER_IT) {
    const int row = r0 + lane / VPR;
    if (row < T) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(attn_sw128<TPAD>(sQ_a, row, sub * 8)));
      *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * T + row) * d + h * DH + sub * 8) = v;
    }
  }
}

}  // namespace mode
