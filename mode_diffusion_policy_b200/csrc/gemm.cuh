// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[rows, N] = A[rows, K] (bf16, K-major)  x  W[N, K]^T (bf16, K-major, nn.Linear layout)   fp32 accumulate in TMEM
//
// One kernel serves every contraction on the MoDE hot path (SURVEY.md §2.2 rows for modedit.py:141-143, :166,
// :561-566, :760-766): dense projections and the *grouped* expert GEMMs share the same code because work is described
// by a device-side table of M-tiles {A row, output row, valid rows, weight row base}. A grouped GEMM is just a table
// whose M-tiles point at different experts' weight rows.
//
// Roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (+ TMEM allocator warp),
// warps 2..5 = epilogue (TMEM -> registers -> fused epilogue -> global). Pipelines: smem ring full/empty mbarriers
// (TMA <-> MMA), two TMEM accumulators of 128x256 fp32 with full/empty mbarriers (MMA <-> epilogue).
#pragma once
#include "ptx.cuh"

namespace mode {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_N = 256;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B atom
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BLOCK_N * GEMM_BLOCK_K * 2;  // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_TMEM_COLS = 512;  // two 256-column accumulators
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,    // out_bf16 = acc + bias[w_row]                       (packed QKV projection)
  EPI_RESID_F32 = 1,    // out_f32  = resid_f32 + acc   (in-place allowed)    (attention c_proj + residual)
  EPI_SWIGLU_BF16 = 2,  // out_bf16[., nb*128+j] = (acc[j]+b[j]) * silu(acc[128+j]+b[128+j])   (expert up-projection)
  EPI_PLAIN_BF16 = 3,   // out_bf16 = acc                                      (expert down-projection)
  EPI_PLAIN_F32 = 4,    // out_f32  = acc                                      (obs/goal token embeddings)
};

// One M-tile of work. For grouped GEMMs consecutive tiles may belong to different experts.
struct GemmMTile {
  int a_row0;      // first row of the A operand (TMA coordinate)
  int out_row0;    // first row of the output
  int rows_valid;  // rows of this tile that are real (<= 128); the rest are never stored
  int w_row_base;  // first weight row of this tile's problem (expert / layer offset), bias uses the same index
};

struct alignas(64) GemmParams {
  CUtensorMap tmap_a;          // [rows, K] bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmap_w;          // [weight rows, K] bf16, box {64, 256}, SWIZZLE_128B
  const GemmMTile* m_tiles;    // device table
  const int* num_m_tiles;      // device scalar (grouped GEMMs build it on the GPU)
  int n_blocks;                // N / 256
  int k_blocks;                // K / 64
  void* out;                   // bf16 or f32, row-major
  int ld_out;                  // elements
  const float* bias;           // indexed by weight row (packed like the weights), may be null
  const float* resid;          // EPI_RESID_F32 only; may alias out
  int w_row_off;               // added to every tile's w_row_base (layer offset of dense projections)
};

__device__ __forceinline__ float silu_f(float g) { return __fdividef(g, 1.0f + __expf(-g)); }

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  // barrier slots: full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + 2 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_w);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);  // every epilogue thread arrives
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr_smem), GEMM_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_m = *p.num_m_tiles;
  const int total = n_m * p.n_blocks;
  const int k_blocks = p.k_blocks;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int mt = t % n_m, nb = t / n_m;
        const GemmMTile tile = p.m_tiles[mt];
        const int w_row = p.w_row_off + tile.w_row_base + nb * GEMM_BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * GEMM_STAGE_BYTES;
          const uint32_t b_dst = a_dst + GEMM_A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), GEMM_STAGE_BYTES);
          tma_load_2d(a_dst, &p.tmap_a, full_bar(stage), kb * GEMM_BLOCK_K, tile.a_row0);
          tma_load_2d(b_dst, &p.tmap_w, full_bar(stage), kb * GEMM_BLOCK_K, w_row);
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, GEMM_BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * GEMM_BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);  // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * GEMM_STAGE_BYTES;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tfull_bar(as));  // accumulator complete
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
      const int mt = t % n_m, nb = t / n_m;
      const GemmMTile tile = p.m_tiles[mt];
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const bool row_ok = row_in_tile < tile.rows_valid;
      const size_t out_row = static_cast<size_t>(tile.out_row0 + row_in_tile);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GEMM_BLOCK_N;
      const int w_row = p.w_row_off + tile.w_row_base + nb * GEMM_BLOCK_N;

      if constexpr (EPI == EPI_SWIGLU_BF16) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ld_out + nb * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t rp[32], rg[32];
          tmem_ld_32x32(taddr + c * 32, rp);
          tmem_ld_32x32(taddr + 128 + c * 32, rg);
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(p.bias + w_row + c * 32);
          const float4* bg = reinterpret_cast<const float4*>(p.bias + w_row + 128 + c * 32);
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b0 = __ldg(bp + j), b1 = __ldg(bg + j);
            const float h0 = (__uint_as_float(rp[4 * j + 0]) + b0.x) * silu_f(__uint_as_float(rg[4 * j + 0]) + b1.x);
            const float h1 = (__uint_as_float(rp[4 * j + 1]) + b0.y) * silu_f(__uint_as_float(rg[4 * j + 1]) + b1.y);
            const float h2 = (__uint_as_float(rp[4 * j + 2]) + b0.z) * silu_f(__uint_as_float(rg[4 * j + 2]) + b1.z);
            const float h3 = (__uint_as_float(rp[4 * j + 3]) + b0.w) * silu_f(__uint_as_float(rg[4 * j + 3]) + b1.w);
            packed[2 * j] = pack_bf16x2(h0, h1);
            packed[2 * j + 1] = pack_bf16x2(h2, h3);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < GEMM_BLOCK_N / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col = nb * GEMM_BLOCK_N + c * 32;
          if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_PLAIN_BF16) {
            uint32_t packed[16];
            if constexpr (EPI == EPI_BIAS_BF16) {
              const float4* b = reinterpret_cast<const float4*>(p.bias + w_row + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = __ldg(b + j);
                packed[2 * j] = pack_bf16x2(__uint_as_float(r[4 * j]) + bb.x, __uint_as_float(r[4 * j + 1]) + bb.y);
                packed[2 * j + 1] =
                    pack_bf16x2(__uint_as_float(r[4 * j + 2]) + bb.z, __uint_as_float(r[4 * j + 3]) + bb.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                packed[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ld_out + col);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
          } else {  // fp32 outputs
            if (row_ok) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ld_out + col);
              if constexpr (EPI == EPI_RESID_F32) {
                const float4* src = reinterpret_cast<const float4*>(p.resid + out_row * p.ld_out + col);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 x = src[j];
                  dst[j] = make_float4(x.x + __uint_as_float(r[4 * j]), x.y + __uint_as_float(r[4 * j + 1]),
                                       x.z + __uint_as_float(r[4 * j + 2]), x.w + __uint_as_float(r[4 * j + 3]));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace mode
