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
#include "rng.cuh"

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
constexpr int GEMM_EPI_BUF_BYTES = 32 * 128;  // one warp's staging tile: 32 rows x 128 B (SWIZZLE_128B box of the TMA store)
constexpr int GEMM_EPI_BYTES = 4 * 2 * GEMM_EPI_BUF_BYTES;  // 4 epilogue warps, double-buffered
constexpr int GEMM_BIAS_BYTES = 2 * GEMM_BLOCK_N * 4;  // per-tile bias slice, double-buffered with the accumulators
constexpr int GEMM_TAIL_BYTES = GEMM_EPI_BYTES + GEMM_BIAS_BYTES + 256 /*barriers*/;
// shared memory layout (1024-byte aligned base): [stage ring][epilogue staging 32 KB][bias 2 KB][mbarriers + tmem ptr]
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_TAIL_BYTES;

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,    // out_bf16 = acc + bias[w_row]                       (packed QKV projection)
  EPI_RESID_F32 = 1,    // out_f32 += acc  (TMA reduce-add into the residual stream) (attention c_proj + residual)
  EPI_SWIGLU_BF16 = 2,  // out_bf16[., nb*128+j] = (acc[j]+b[j]) * silu(acc[128+j]+b[128+j])   (expert up-projection)
  EPI_PLAIN_BF16 = 3,   // out_bf16 = acc                                      (expert down-projection)
  EPI_PLAIN_F32 = 4,    // out_f32  = acc                                      (obs/goal token embeddings)
  EPI_SWIGLU_SAVE = 5,  // EPI_SWIGLU_BF16 + the pre-activations z = acc + bias -> tmap_out2 (training forward)
  EPI_CONV_BF16 = 6,    // out_bf16 = film(relu(acc + bias + residual))  (FiLM-ResNet convolutions as GEMMs, resnet.inc)
};
// Extras of EPI_CONV_BF16: folded-BatchNorm bias comes through GemmParams::bias; optional residual (identity / projected
// shortcut of a bottleneck), ReLU, and FiLM (1 + gamma[n, c]) * v + beta[n, c] with n = row / film_hw (reference
// pretrained_resnets.py:19-23, applied to the output of each ResNet stage).
struct ConvEpilogue {
  const __nv_bfloat16* res;  // [rows, ld_res] or nullptr
  int ld_res;
  int relu;
  const float* film_g;       // [images, film_c] or nullptr
  const float* film_b;
  int film_hw, film_c;       // rows per image (film_hw) for the FiLM / border lookups
  // zero-bordered activation layout (resnet.inc, implicit 3x3 convolutions): every image is stored as (pad_h + 2) x pad_w2
  // positions with a one-pixel zero border; rows that are border positions are written as zeros. pad_w2 = 0: dense layout.
  int pad_w2, pad_h;
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
  CUtensorMap tmap_out;        // output [rows, N_out] bf16 (box {64, 32}) or f32 (box {32, 32}), SWIZZLE_128B
  CUtensorMap tmap_out2;       // EPI_SWIGLU_SAVE: pre-activation output [rows, N] bf16
  const float* bias;           // indexed by weight row (packed like the weights), may be null
  int w_row_off;               // added to every tile's w_row_base (layer offset of dense projections)
  // stream-K over the last, partial wave of tiles (CTA-pair kernel only; see SkIter)
  int sk_enable;
  float4* sk_partials;         // [grid CTAs][8 chunks][8 vec][128 rows] fp32 partial accumulators (128 KB per CTA)
  int* sk_flags;               // [grid CTAs][4 epilogue warps], 0 between kernels
  // EPI_SWIGLU_SAVE only (training): nn.Dropout between SwishGLU and the down projection (modedit.py:254) applied to h
  // in the epilogue. Hidden unit j of token m in expert e uses half (j & 1) of word (m*E + e)*(F/2) + j/2 of stream
  // RNG_MLP; row_token maps a permuted row to its token (written by ln2_permute_kernel). drop.thr == 0: no dropout.
  DropoutSpec drop;
  const int* row_token;
  int drop_rows_per_expert;    // weight rows per expert (8d): expert = (w_row_base / this) % drop_E
  int drop_E;
  int drop_half_F;             // F / 2 = 2d
  // Tile width of the CTA-pair kernel (multiple of 16, <= 256; 256 everywhere else). The host picks the width that
  // fills whole waves of the 74 CTA pairs (e.g. 208 for N = 1024 / 3072 at B = 256: 5 resp. 15 column tiles instead of
  // 4 / 12), which changes which tile an output column belongs to but not the order in which its K terms are summed.
  // n_blocks = ceil(n_total / bn); tmap_w has a bn/2-row box; columns [bn/64*64, bn) of a tile (bf16 outputs; bn/32*32
  // for fp32) are stored straight from registers through out_ptr, everything else through tmap_out as before.
  ConvEpilogue conv;           // EPI_CONV_BF16 only
  // Tile order of the CTA-pair kernel. band = 0: tile t -> (mt = t % n_m, nb = t / n_m), i.e. all M-tiles of one column
  // block before the next. band = G > 0: bands of G M-tiles; inside a band all column blocks of those rows are walked
  // (M fastest), so one wave of tiles re-uses a band of A rows for every column block while they are still in L2. Used
  // for the expert down-projection, whose A operand (h, 58.7 MB at B = 256) does not survive in L2 across the 4-5 passes of
  // the column-block-major order (ncu: 102.7 MB DRAM reads for 75.5 MB of operands). Values do not depend on the order.
  int band;
  // Implicit convolution (resnet.inc): n_taps > 0 splits the K loop into n_taps groups of kb_per_tap k-blocks; group t reads
  // the A rows shifted by tap_off[t] (a 3x3 tap of a zero-bordered NHWC activation is a row offset) at columns
  // [0, 64 * kb_per_tap); the weight columns advance through all n_taps * kb_per_tap k-blocks as usual. Rows that fall
  // outside the tensor are zero-filled by TMA.
  int n_taps, kb_per_tap;
  int tap_off[9];
  int bn;
  int n_total;                 // N of the whole GEMM (output columns / weight rows per problem)
  void* out_ptr;               // raw output base, row stride ldo elements, rows >= out_rows are never written
  int ldo, out_rows;
};
// Column geometry handed to the epilogue (defaults = the fixed 256-wide tiles of the other kernels).
struct EpiGeom {
  int bn = GEMM_BLOCK_N;
  int n_total = 0x7fffffff;
  void* out = nullptr;
  int ldo = 0, out_rows = 0;
  ConvEpilogue conv = ConvEpilogue{nullptr, 0, 0, nullptr, nullptr, 1, 0, 0, 0};
};

// Work decomposition of the CTA-pair kernel. Tiles of full waves are data-parallel (tile = worker + i * P). The last,
// partial wave (R = T mod P tiles) is processed stream-K style: its R * KB k-block units are divided evenly over the
// workers, so a tile may be produced by several workers. The worker that owns a tile's first k-block finishes the
// tile (it adds the other workers' fp32 partial accumulators before the normal epilogue); it handles that segment LAST
// while the contributors handle theirs FIRST in the stream-K phase, so waits are short and cannot deadlock.
struct SkIter {
  int w, P, T, KB, T_dp, P_sk, units;  // worker, workers, tiles, k-blocks/tile, data-parallel tiles, sk workers, sk units
  int i_dp, u, u_end;
  __device__ SkIter(int worker, int workers, int tiles, int kblocks, int sk_enable) {
    w = worker; P = workers; T = tiles; KB = kblocks;
    const int R = T % P;
    // stream-K only when there is at least one full wave before the remainder (otherwise plain data-parallel)
    const bool sk = sk_enable && R > 0 && T > P;
    T_dp = sk ? T - R : T;
    P_sk = sk ? min(P, R * 4) : 0;  // at most 4 contributors per tile
    units = sk ? R * KB : 0;
    i_dp = 0;
    u = (sk && w < P_sk) ? static_cast<int>(static_cast<long long>(w) * units / P_sk) : 0;
    u_end = (sk && w < P_sk) ? static_cast<int>(static_cast<long long>(w + 1) * units / P_sk) : 0;
  }
  __device__ int sk_begin(int worker) const { return static_cast<int>(static_cast<long long>(worker) * units / P_sk); }
  // next segment: tile index, k-block range [kb0, kb1)
  __device__ bool next(int& tile, int& kb0, int& kb1) {
    const int t = w + i_dp * P;
    if (t < T_dp) {
      ++i_dp;
      tile = t; kb0 = 0; kb1 = KB;
      return true;
    }
    if (u < u_end) {
      const int tt = u / KB;
      kb0 = u - tt * KB;
      kb1 = min(KB, kb0 + (u_end - u));
      tile = T_dp + tt;
      u += kb1 - kb0;
      return true;
    }
    return false;
  }
  // workers after `w` that hold a part of stream-K tile `tile` (only meaningful for a segment with kb0 == 0 < kb1 < KB)
  __device__ int contributors(int tile) const {
    const int tile_end = (tile - T_dp + 1) * KB;
    int n = 0;
    while (w + 1 + n < P_sk && sk_begin(w + 1 + n) < tile_end) ++n;
    return n;
  }
};

__device__ __forceinline__ void decode_tile(int t, int n_m, int n_blocks, int band, int& mt, int& nb) {
  if (band <= 0) {
    mt = t % n_m;
    nb = t / n_m;
    return;
  }
  const int per_band = band * n_blocks;
  const int b = t / per_band, idx = t - b * per_band;
  const int rows = min(band, n_m - b * band);  // the last band may be short
  mt = b * band + idx % rows;
  nb = idx / rows;
}

// silu(g) = g / (1 + 2^(-g*log2 e)) with the SFU's ex2/rcp approximations (~2^-22 relative error each; the result is
// rounded to bf16 right after). 5 instructions per element: the SwiGLU epilogue must stay below the MMA time of a tile.
__device__ __forceinline__ float silu_f(float g) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(g * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return g * r;
}

// Fused epilogue of one warp's 32 accumulator rows x 256 columns. Each lane owns one row (TMEM lane); values go
// TMEM -> registers -> epilogue math -> a 32 x 128 B staging tile in shared memory laid out in the TMA SWIZZLE_128B
// pattern (16-byte chunk j of row r lives at chunk j ^ (r & 7): conflict-free v4 stores) -> ONE bulk tensor store
// (or reduce-add) per chunk issued by lane 0, double-buffered with bulk async-groups. Global writes are therefore
// full 128-byte rows instead of 32 scattered 16-byte pieces per instruction.
// Partial accumulators in global memory: chunk c (32 columns), vector j (4 columns), row r -> float4 index
// (c*8 + j)*128 + r: consecutive lanes (rows) touch consecutive 16-byte words, i.e. fully coalesced both ways.
struct SkParts {
  const float4* base;  // first contributor's slot, this CTA rank (contributors are 2 CTA slots apart)
  int n;               // number of contributors (0: plain tile)
  int row;             // row of this lane inside the CTA's 128-row half
};
__device__ __forceinline__ void load_acc32(uint32_t taddr_col, uint32_t (&r)[32], const SkParts& sk, int chunk) {
  tmem_ld_32x32(taddr_col, r);
  tmem_ld_wait();
  for (int i = 0; i < sk.n; ++i) {
    const float4* src = sk.base + static_cast<size_t>(i) * 2 * (8 * 8 * 128) + (chunk * 8) * 128 + sk.row;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 v = __ldcg(src + j * 128);
      r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + v.x);
      r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
      r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
      r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
    }
  }
}
// A contributor's segment: dump this warp's 32 x 256 fp32 accumulator rows to its slot and publish them.
__device__ __forceinline__ void store_partial_warp(uint32_t taddr, float4* slot, int* flag, int row, int lane) {
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      __stcg(slot + (c * 8 + j) * 128 + row,
             make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                         __uint_as_float(r[4 * j + 3])));
  }
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(flag, 1);
}
__device__ __forceinline__ void wait_flag(const int* flag) {
  uint32_t spins = 0;
  int v;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (!v && ++spins > MODE_SPIN_LIMIT) {
      printf("mode: stream-K partial wait timed out (block %d)\n", blockIdx.x);
      __trap();
    }
  } while (!v);
}

// per-lane dropout state of the SwiGLU epilogue: word index of this row's hidden unit 0
struct EpiDrop {
  uint32_t key, thr, base;
  float scale;
};
template <int EPI, bool DROP = false>
__device__ __forceinline__ void gemm_epilogue_warp(const CUtensorMap* tmap_out, uint32_t taddr, uint32_t stage_smem, int lane,
                                                   int out_row0, const float* sbias, int nb, uint32_t& n_stores,
                                                   const SkParts sk = SkParts{nullptr, 0, 0},
                                                   const EpiDrop dr = EpiDrop{0, 0, 0, 1.0f},
                                                   const EpiGeom g = EpiGeom{}) {
  const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  auto chunk_addr = [&](uint32_t buf, uint32_t j) { return buf + row_off + ((j ^ sw) << 4); };
  auto begin_chunk = [&]() -> uint32_t {
    // the store issued two chunks ago read this buffer: wait until at most one store is still reading smem
    if (lane == 0) bulk_wait_group_read<1>();
    __syncwarp();
    return stage_smem + (n_stores & 1u) * GEMM_EPI_BUF_BYTES;
  };
  auto end_chunk = [&](uint32_t buf, int col) {
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
      if constexpr (EPI == EPI_RESID_F32)
        tma_reduce_add_2d(tmap_out, buf, col, out_row0);
      else
        tma_store_2d(tmap_out, buf, col, out_row0);
      bulk_commit_group();
    }
    ++n_stores;
  };

  if constexpr (EPI == EPI_SWIGLU_BF16) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {  // 64 hidden units per store
      const uint32_t buf = begin_chunk();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c32 = c * 64 + hh * 32;
        uint32_t rp[32], rg[32];
        load_acc32(taddr + c32, rp, sk, c32 / 32);
        load_acc32(taddr + 128 + c32, rg, sk, (128 + c32) / 32);
        const float4* bp = reinterpret_cast<const float4*>(sbias + c32);
        const float4* bg = reinterpret_cast<const float4*>(sbias + 128 + c32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t pk[4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 b0 = bp[2 * j + u], b1 = bg[2 * j + u];
            const int o = 8 * j + 4 * u;
            float h0 = (__uint_as_float(rp[o + 0]) + b0.x) * silu_f(__uint_as_float(rg[o + 0]) + b1.x);
            float h1 = (__uint_as_float(rp[o + 1]) + b0.y) * silu_f(__uint_as_float(rg[o + 1]) + b1.y);
            float h2 = (__uint_as_float(rp[o + 2]) + b0.z) * silu_f(__uint_as_float(rg[o + 2]) + b1.z);
            float h3 = (__uint_as_float(rp[o + 3]) + b0.w) * silu_f(__uint_as_float(rg[o + 3]) + b1.w);
            if constexpr (DROP) {
              const uint32_t w = dr.base + static_cast<uint32_t>(nb * 128 + c32 + o) / 2u;
              const uint32_t r0 = rng_bits(dr.key, w), r1 = rng_bits(dr.key, w + 1u);
              h0 = (r0 & 0xffffu) < dr.thr ? 0.f : h0 * dr.scale;
              h1 = (r0 >> 16) < dr.thr ? 0.f : h1 * dr.scale;
              h2 = (r1 & 0xffffu) < dr.thr ? 0.f : h2 * dr.scale;
              h3 = (r1 >> 16) < dr.thr ? 0.f : h3 * dr.scale;
            }
            pk[2 * u] = pack_bf16x2(h0, h1);
            pk[2 * u + 1] = pack_bf16x2(h2, h3);
          }
          st_shared_v4(chunk_addr(buf, hh * 4 + j), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      end_chunk(buf, nb * 128 + c * 64);
    }
  } else if constexpr (EPI == EPI_CONV_BF16) {
    const int col0 = nb * g.bn;
    const int nfull = g.bn >> 6;  // convolution widths are multiples of 64
    const int row = out_row0 + lane;
    const bool row_ok = row < g.out_rows;
    const int img = row_ok ? row / g.conv.film_hw : 0;
    bool border = false;
    if (g.conv.pad_w2 > 0 && row_ok) {
      const int q = row - img * g.conv.film_hw, y = q / g.conv.pad_w2, x = q - y * g.conv.pad_w2;
      border = y == 0 || y > g.conv.pad_h || x == 0 || x == g.conv.pad_w2 - 1;
    }
#pragma unroll 1
    for (int c = 0; c < nfull; ++c) {
      if (col0 + c * 64 >= g.n_total) break;
      const uint32_t buf = begin_chunk();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c32 = c * 64 + hh * 32;
        uint32_t r[32];
        load_acc32(taddr + c32, r, sk, c32 / 32);
        const float4* b = reinterpret_cast<const float4*>(sbias + c32);
        const uint4* rs = (g.conv.res && row_ok)
                              ? reinterpret_cast<const uint4*>(g.conv.res + static_cast<size_t>(row) * g.conv.ld_res + col0 + c32)
                              : nullptr;
        const float4* fg = (g.conv.film_g && row_ok)
                               ? reinterpret_cast<const float4*>(g.conv.film_g + static_cast<size_t>(img) * g.conv.film_c + col0 + c32)
                               : nullptr;
        const float4* fb = fg ? reinterpret_cast<const float4*>(g.conv.film_b + static_cast<size_t>(img) * g.conv.film_c + col0 + c32)
                              : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8];
          const float4 b0 = b[2 * j], b1 = b[2 * j + 1];
          v[0] = __uint_as_float(r[8 * j + 0]) + b0.x; v[1] = __uint_as_float(r[8 * j + 1]) + b0.y;
          v[2] = __uint_as_float(r[8 * j + 2]) + b0.z; v[3] = __uint_as_float(r[8 * j + 3]) + b0.w;
          v[4] = __uint_as_float(r[8 * j + 4]) + b1.x; v[5] = __uint_as_float(r[8 * j + 5]) + b1.y;
          v[6] = __uint_as_float(r[8 * j + 6]) + b1.z; v[7] = __uint_as_float(r[8 * j + 7]) + b1.w;
          if (rs) {
            const uint4 q = __ldg(rs + j);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 f = __bfloat1622float2(h2[u]);
              v[2 * u] += f.x;
              v[2 * u + 1] += f.y;
            }
          }
          if (g.conv.relu) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.f);
          }
          if (fg) {
            const float4 g0 = __ldg(fg + 2 * j), g1 = __ldg(fg + 2 * j + 1), e0 = __ldg(fb + 2 * j), e1 = __ldg(fb + 2 * j + 1);
            v[0] = fmaf(1.f + g0.x, v[0], e0.x); v[1] = fmaf(1.f + g0.y, v[1], e0.y);
            v[2] = fmaf(1.f + g0.z, v[2], e0.z); v[3] = fmaf(1.f + g0.w, v[3], e0.w);
            v[4] = fmaf(1.f + g1.x, v[4], e1.x); v[5] = fmaf(1.f + g1.y, v[5], e1.y);
            v[6] = fmaf(1.f + g1.z, v[6], e1.z); v[7] = fmaf(1.f + g1.w, v[7], e1.w);
          }
          if (border) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = 0.f;
          }
          st_shared_v4(chunk_addr(buf, hh * 4 + j), pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                       pack_bf16x2(v[6], v[7]));
        }
      }
      end_chunk(buf, col0 + c * 64);
    }
  } else if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_PLAIN_BF16) {
    const int col0 = nb * g.bn;
    const int nfull = g.bn >> 6;
#pragma unroll 1
    for (int c = 0; c < nfull; ++c) {  // 64 output columns per store
      if (col0 + c * 64 >= g.n_total) break;  // tile overhangs the matrix (its weight rows were out-of-range filler)
      const uint32_t buf = begin_chunk();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c32 = c * 64 + hh * 32;
        uint32_t r[32];
        load_acc32(taddr + c32, r, sk, c32 / 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t pk[4];
          if constexpr (EPI == EPI_BIAS_BF16) {
            const float4* b = reinterpret_cast<const float4*>(sbias + c32);
            const float4 b0 = b[2 * j], b1 = b[2 * j + 1];
            pk[0] = pack_bf16x2(__uint_as_float(r[8 * j + 0]) + b0.x, __uint_as_float(r[8 * j + 1]) + b0.y);
            pk[1] = pack_bf16x2(__uint_as_float(r[8 * j + 2]) + b0.z, __uint_as_float(r[8 * j + 3]) + b0.w);
            pk[2] = pack_bf16x2(__uint_as_float(r[8 * j + 4]) + b1.x, __uint_as_float(r[8 * j + 5]) + b1.y);
            pk[3] = pack_bf16x2(__uint_as_float(r[8 * j + 6]) + b1.z, __uint_as_float(r[8 * j + 7]) + b1.w);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              pk[u] = pack_bf16x2(__uint_as_float(r[8 * j + 2 * u]), __uint_as_float(r[8 * j + 2 * u + 1]));
          }
          st_shared_v4(chunk_addr(buf, hh * 4 + j), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      end_chunk(buf, col0 + c * 64);
    }
    // columns past the last full 64-column chunk (tile widths like 208 = 3 x 64 + 16): 16 at a time, one 32-byte
    // sector per row, stored straight from registers
#pragma unroll 1
    for (int t = nfull * 64; t < g.bn; t += 16) {
      if (col0 + t >= g.n_total) break;
      uint32_t r[16];
      tmem_ld_32x16(taddr + t, r);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float v0 = __uint_as_float(r[2 * u]), v1 = __uint_as_float(r[2 * u + 1]);
        if constexpr (EPI == EPI_BIAS_BF16) {
          v0 += sbias[t + 2 * u];
          v1 += sbias[t + 2 * u + 1];
        }
        pk[u] = pack_bf16x2(v0, v1);
      }
      const int row = out_row0 + lane;
      if (row < g.out_rows) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(g.out) + static_cast<size_t>(row) * g.ldo + col0 + t);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
  } else {  // fp32 outputs: 32 columns (128 B) per store
    const int col0 = nb * g.bn;
    const int nfull = g.bn >> 5;
#pragma unroll 1
    for (int c = 0; c < nfull; ++c) {
      if (col0 + c * 32 >= g.n_total) break;
      const uint32_t buf = begin_chunk();
      uint32_t r[32];
      load_acc32(taddr + c * 32, r, sk, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) st_shared_v4(chunk_addr(buf, j), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      end_chunk(buf, col0 + c * 32);
    }
    if ((g.bn & 16) && col0 + nfull * 32 < g.n_total) {  // the last 16 columns of a tile whose width is 16 mod 32
      const int t = nfull * 32;
      uint32_t r[16];
      tmem_ld_32x16(taddr + t, r);
      tmem_ld_wait();
      const int row = out_row0 + lane;
      if (row < g.out_rows) {
        float* dst = reinterpret_cast<float*>(g.out) + static_cast<size_t>(row) * g.ldo + col0 + t;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if constexpr (EPI == EPI_RESID_F32)
            red_add_v4_f32(dst + 4 * j, __uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                           __uint_as_float(r[4 * j + 3]));
          else
            *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
      }
    }
  }
}

// The 128 epilogue threads copy the tile's 256 bias values into shared memory BEFORE they wait for the accumulator, so
// the global-load latency hides behind the tile's MMAs; named barrier 1 (epilogue warps only) publishes them.
template <int EPI>
__device__ __forceinline__ void stage_bias(const GemmParams& p, float* sbias, int w_row, int epi_tid, int col0 = 0) {
  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_SWIGLU_BF16 || EPI == EPI_SWIGLU_SAVE || EPI == EPI_CONV_BF16) {
    // entries past the tile width or past the matrix edge (an overhanging last tile) are never used: read nothing there
    const int j0 = epi_tid, j1 = 128 + epi_tid;
    sbias[j0] = (j0 < p.bn && col0 + j0 < p.n_total) ? __ldg(p.bias + w_row + j0) : 0.f;
    sbias[j1] = (j1 < p.bn && col0 + j1 < p.n_total) ? __ldg(p.bias + w_row + j1) : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
  }
}

// EPI_SWIGLU_SAVE = the bias->bf16 epilogue into tmap_out2 followed by the SwiGLU epilogue into tmap_out
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_dispatch(const GemmParams& p, uint32_t taddr, uint32_t stage_smem, int lane,
                                                       int out_row0, const float* sbias, int nb, uint32_t& n_stores,
                                                       int w_row_base, const SkParts sk = SkParts{nullptr, 0, 0}) {
  if constexpr (EPI == EPI_SWIGLU_SAVE) {
    gemm_epilogue_warp<EPI_BIAS_BF16>(&p.tmap_out2, taddr, stage_smem, lane, out_row0, sbias, nb, n_stores, sk);
    if (p.drop.thr) {
      const uint32_t token = static_cast<uint32_t>(p.row_token[out_row0 + lane]);
      const uint32_t expert = static_cast<uint32_t>((w_row_base / p.drop_rows_per_expert) % p.drop_E);
      const EpiDrop dr{p.drop.key, p.drop.thr, (token * p.drop_E + expert) * static_cast<uint32_t>(p.drop_half_F), p.drop.scale};
      gemm_epilogue_warp<EPI_SWIGLU_BF16, true>(&p.tmap_out, taddr, stage_smem, lane, out_row0, sbias, nb, n_stores, sk, dr);
    } else {
      gemm_epilogue_warp<EPI_SWIGLU_BF16>(&p.tmap_out, taddr, stage_smem, lane, out_row0, sbias, nb, n_stores, sk);
    }
  } else {
    gemm_epilogue_warp<EPI>(&p.tmap_out, taddr, stage_smem, lane, out_row0, sbias, nb, n_stores, sk, EpiDrop{0, 0, 0, 1.0f},
                            EpiGeom{p.bn, p.n_total, p.out_ptr, p.ldo, p.out_rows, p.conv});
  }
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();
  float* sbias_all = reinterpret_cast<float*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_EPI_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_EPI_BYTES + GEMM_BIAS_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  // barrier slots: full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + 2 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);
  const uint32_t epi_smem = smem_base + GEMM_STAGES * GEMM_STAGE_BYTES;  // 1024-aligned: stages are 48 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
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

  pdl_wait();  // setup above overlapped the previous kernel's tail; its outputs are read from here on
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
    const uint32_t stage_smem = epi_smem + static_cast<uint32_t>(q) * 2 * GEMM_EPI_BUF_BYTES;
    uint32_t n_stores = 0;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++iter) {
      const int mt = t % n_m, nb = t / n_m;
      const GemmMTile tile = p.m_tiles[mt];
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      float* sbias = sbias_all + as * GEMM_BLOCK_N;
      stage_bias<EPI>(p, sbias, p.w_row_off + tile.w_row_base + nb * GEMM_BLOCK_N, q * 32 + lane, nb * GEMM_BLOCK_N);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      if (q * 32 < tile.rows_valid) {  // warp-uniform: this warp's 32 rows hold at least one real row
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GEMM_BLOCK_N;
        gemm_epilogue_dispatch<EPI>(p, taddr, stage_smem, lane, tile.out_row0 + q * 32, sbias, nb, n_stores, tile.w_row_base);
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld): hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
    }
    if (lane == 0) bulk_wait_group<0>();  // outstanding stores must land before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster compute one 256 x 256 tile. Each CTA loads its own
// 128 A rows and HALF of the weight tile (128 of the 256 N rows); the tensor core reads both halves, so the L2->SM
// traffic per MMA drops from 48 KB to 32 KB per k-block and six stages fit in shared memory. The leader CTA (rank 0)
// issues the MMAs for the pair; both CTAs run a TMA producer and an epilogue over their own 128 accumulator rows.
// M-tiles in the table are 256 rows here.
constexpr int G2_STAGES = 6;
constexpr int G2_HALF_N = GEMM_BLOCK_N / 2;
constexpr int G2_B_BYTES = G2_HALF_N * GEMM_BLOCK_K * 2;       // 16 KB
constexpr int G2_STAGE_BYTES = GEMM_A_BYTES + G2_B_BYTES;      // 32 KB
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + GEMM_TAIL_BYTES;

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_2cta_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();
  float* sbias_all = reinterpret_cast<float*>(smem + G2_STAGES * G2_STAGE_BYTES + GEMM_EPI_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES + GEMM_EPI_BYTES + GEMM_BIAS_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (G2_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * G2_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * G2_STAGES + 2 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);
  const uint32_t epi_smem = smem_base + G2_STAGES * G2_STAGE_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_a);
    tma_prefetch_desc(&p.tmap_w);
    tma_prefetch_desc(&p.tmap_out);
    if constexpr (EPI == EPI_SWIGLU_SAVE) tma_prefetch_desc(&p.tmap_out2);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(full_bar(s), 2);   // leader's: one arrival per CTA's producer (+ both CTAs' TMA bytes)
      mbar_init(empty_bar(s), 1);  // each CTA's own: the pair's MMA commit is multicast to both
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);     // each CTA's own (multicast commit)
      mbar_init(tempty_bar(s), 256);  // leader's: all 2 x 128 epilogue threads of the pair
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(smem_u32(tmem_ptr_smem), GEMM_TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  pdl_wait();  // setup above overlapped the previous kernel's tail; its outputs are read from here on
  const int n_m = *p.num_m_tiles;
  const int total = n_m * p.n_blocks;
  const int k_blocks = p.k_blocks;
  const int bn = p.bn, half_n = p.bn >> 1;
  const uint32_t stage_tx = GEMM_A_BYTES + static_cast<uint32_t>(half_n) * (GEMM_BLOCK_K * 2);  // bytes per CTA per k-block

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      SkIter it(pair, n_pairs, total, k_blocks, p.sk_enable & 1);
      int t, kb0, kb1;
      while (it.next(t, kb0, kb1)) {
        int mt, nb;
        decode_tile(t, n_m, p.n_blocks, p.band, mt, nb);
        const GemmMTile tile = p.m_tiles[mt];
        const int a_row = tile.a_row0 + static_cast<int>(rank) * GEMM_BLOCK_M;
        const int w_row = p.w_row_off + tile.w_row_base + nb * bn + static_cast<int>(rank) * half_n;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * G2_STAGE_BYTES;
          const uint32_t b_dst = a_dst + GEMM_A_BYTES;
          const uint32_t leader_full = mapa_cluster(full_bar(stage), 0);
          // measurement aid (mode_debug_gemm bit 10): odd k-blocks reuse the stale weight tile of their stage, i.e. the
          // same MMA work with 25 % less L2->SM traffic; results are garbage
          const bool skip_b = (p.sk_enable & 2) && (kb & 1);
          if (rank == 0)
            mbar_arrive_expect_tx(full_bar(stage), skip_b ? 2 * GEMM_A_BYTES : 2 * stage_tx);
          else
            mbar_arrive_cluster(leader_full);
          if (p.n_taps > 0) {
            const int tap = kb / p.kb_per_tap;
            tma_load_2d_2sm(a_dst, &p.tmap_a, leader_full, (kb - tap * p.kb_per_tap) * GEMM_BLOCK_K, a_row + p.tap_off[tap]);
          } else {
            tma_load_2d_2sm(a_dst, &p.tmap_a, leader_full, kb * GEMM_BLOCK_K, a_row);
          }
          if (!skip_b) tma_load_2d_2sm(b_dst, &p.tmap_w, leader_full, kb * GEMM_BLOCK_K, w_row);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, static_cast<uint32_t>(bn));
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      SkIter it(pair, n_pairs, total, k_blocks, p.sk_enable & 1);
      int t, kb0, kb1;
      for (; it.next(t, kb0, kb1); ++iter) {
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * GEMM_BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * G2_STAGE_BYTES;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
            umma_bf16_2cta(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_2cta(empty_bar(stage), 0b11);  // frees the stage in BOTH CTAs
          if (kb == kb1 - 1) umma_commit_2cta(tfull_bar(as), 0b11);
          if (++stage == G2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const uint32_t stage_smem = epi_smem + static_cast<uint32_t>(q) * 2 * GEMM_EPI_BUF_BYTES;
    uint32_t n_stores = 0;
    int iter = 0;
    SkIter it(pair, n_pairs, total, k_blocks, p.sk_enable & 1);
    int t, kb0, kb1;
    constexpr size_t kSlot = 8 * 8 * 128;  // float4 per CTA slot
    for (; it.next(t, kb0, kb1); ++iter) {
      int mt, nb;
      decode_tile(t, n_m, p.n_blocks, p.band, mt, nb);
      const GemmMTile tile = p.m_tiles[mt];
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      float* sbias = sbias_all + as * GEMM_BLOCK_N;
      stage_bias<EPI>(p, sbias, p.w_row_off + tile.w_row_base + nb * bn, q * 32 + lane, nb * bn);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int row0 = static_cast<int>(rank) * GEMM_BLOCK_M + q * 32;  // first row of this warp inside the 256-row tile
      if (row0 < tile.rows_valid) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GEMM_BLOCK_N;
        if (kb0 != 0) {
          // contributor: publish the partial accumulator of this segment in this CTA's slot
          store_partial_warp(taddr, p.sk_partials + static_cast<size_t>(blockIdx.x) * kSlot, p.sk_flags + blockIdx.x * 4 + q,
                             q * 32 + lane, lane);
        } else {
          SkParts sk{nullptr, 0, q * 32 + lane};
          if (kb1 != k_blocks) {  // finisher of a split tile: collect the contributors' partials first
            sk.n = it.contributors(t);
            const int first = 2 * (pair + 1) + static_cast<int>(rank);
            sk.base = p.sk_partials + static_cast<size_t>(first) * kSlot;
            for (int i = 0; i < sk.n; ++i) wait_flag(p.sk_flags + (first + 2 * i) * 4 + q);
          }
          gemm_epilogue_dispatch<EPI>(p, taddr, stage_smem, lane, tile.out_row0 + row0, sbias, nb, n_stores, tile.w_row_base, sk);
          if (sk.n > 0) {
            __syncwarp();
            if (lane == 0)
              for (int i = 0; i < sk.n; ++i) p.sk_flags[(2 * (pair + 1 + i) + static_cast<int>(rank)) * 4 + q] = 0;
          }
        }
      }
      tc_fence_before();
      if (rank == 0)
        mbar_arrive(tempty_bar(as));
      else
        mbar_arrive_cluster(mapa_cluster(tempty_bar(as), 0));
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be arriving on the leader's barriers / reading its smem via the MMA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace mode
