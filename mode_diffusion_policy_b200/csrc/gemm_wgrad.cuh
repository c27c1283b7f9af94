// Weight-gradient GEMM for sm_100a:  dW[N_out, K_out] = dY[rows, N_out]^T  x  X[rows, K_out]   (fp32 out)
//
// The contraction runs over token ROWS, so both operands are consumed "MN-major": a TMA box of {64 columns, 64 rows}
// of a row-major activation lands in shared memory as 64 contraction rows x 128 bytes of the MN index, which is exactly
// the canonical SWIZZLE_128B MN-major tcgen05 operand layout — no transposed copies of activations are ever made.
// One descriptor covers all 64-column atoms of an operand (LBO = 8 KB between atoms, SBO = 1 KB between 8-row groups);
// advancing 16 contraction rows inside the stage is +2 KB. Reference being replaced: torch.autograd's Linear backward
// (grad_weight = grad_output.t() @ input) for every nn.Linear on the path (modedit.py:108-111, :87, :255).
//
// Work = a device-side table of problems {first row, k-blocks, output row base}: a grouped (per-expert) weight
// gradient is a table with one problem per expert whose row range is that expert's token group. Problems with zero rows
// are skipped (the gradient buffer is zero-filled per step, un-routed experts keep exact zeros like the reference).
#pragma once
#include "gemm.cuh"

namespace mode {

struct WgradProblem {
  int row0;          // first contraction row (same in both operand buffers)
  int k_blocks;      // contraction rows / 64 (rows beyond the real count must be zero in the dY operand)
  int out_row_base;  // first output row of this problem in the gradient matrix (packed weight-row index space)
  int pad;
};

struct alignas(64) WgradParams {
  CUtensorMap tmap_dy;   // [rows, N_out] bf16, box {64, 64}, SWIZZLE_128B
  CUtensorMap tmap_x;    // [rows, K_out] bf16, box {64, 64}, SWIZZLE_128B
  CUtensorMap tmap_out;  // [grad rows, K_out] fp32, box {32, 32}, SWIZZLE_128B
  const WgradProblem* problems;
  int n_problems;
  int m_tiles;           // N_out / 128
  int n_blocks;          // K_out / 256
  int swiglu_half;       // > 0: output rows are in the interleaved SwiGLU packing; un-interleave on store (half = 4d)
};

// MN-major SWIZZLE_128B operand descriptor (see header comment).
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(8192u >> 4) << 16;  // LBO: next 64-column atom (64 rows x 128 B)
  d |= static_cast<uint64_t>(1024u >> 4) << 32;  // SBO: next group of 8 contraction rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_EPI_BYTES + GEMM_BIAS_BYTES);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + 2 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);
  const uint32_t epi_smem = smem_base + GEMM_STAGES * GEMM_STAGE_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_dy);
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_out);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
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
  pdl_wait();

  const int tiles_per_problem = p.m_tiles * p.n_blocks;
  const int total = p.n_problems * tiles_per_problem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const WgradProblem pr = p.problems[t / tiles_per_problem];
        if (pr.k_blocks == 0) continue;
        const int r = t % tiles_per_problem, mt = r % p.m_tiles, nb = r / p.m_tiles;
        for (int kb = 0; kb < pr.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * GEMM_STAGE_BYTES;
          const uint32_t b_dst = a_dst + GEMM_A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), GEMM_STAGE_BYTES);
          const int row = pr.row0 + kb * 64;
#pragma unroll
          for (int j = 0; j < 2; ++j) tma_load_2d(a_dst + j * 8192, &p.tmap_dy, full_bar(stage), mt * 128 + j * 64, row);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_2d(b_dst + j * 8192, &p.tmap_x, full_bar(stage), nb * 256 + j * 64, row);
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major: instruction-descriptor bits 15 (A) and 16 (B)
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, GEMM_BLOCK_N) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const WgradProblem pr = p.problems[t / tiles_per_problem];
        if (pr.k_blocks == 0) continue;
        const int as = iter & 1;
        const uint32_t aphase = (iter >> 1) & 1;
        ++iter;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * GEMM_BLOCK_N;
        for (int kb = 0; kb < pr.k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * GEMM_STAGE_BYTES;
          const uint64_t a_desc = make_smem_desc_mn_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_mn_sw128(a_addr + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)  // +16 contraction rows = +2 KB = +128 in the >>4 field
            umma_bf16(tmem_d, a_desc + 128u * k, b_desc + 128u * k, idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));
          if (kb == pr.k_blocks - 1) umma_commit(tfull_bar(as));
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const uint32_t stage_smem = epi_smem + static_cast<uint32_t>(q) * 2 * GEMM_EPI_BUF_BYTES;
    uint32_t n_stores = 0;
    int iter = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const WgradProblem pr = p.problems[t / tiles_per_problem];
      const int r = t % tiles_per_problem, mt = r % p.m_tiles, nb = r / p.m_tiles;
      int out_row = mt * 128;  // row inside the problem, packed index space
      if (p.swiglu_half > 0) {
        // packed blocks of 256 rows = 128 projected rows then their 128 gate rows -> reference row order
        const int blk = out_row / 256, in_blk = out_row % 256;
        out_row = (in_blk < 128 ? 0 : p.swiglu_half) + blk * 128;
      }
      out_row += pr.out_row_base + q * 32;
      if (pr.k_blocks == 0) {
        // no token was routed to this expert: its gradient is exactly zero (the reference leaves .grad = None)
        for (int c = 0; c < 8; ++c) {
          if (lane == 0) bulk_wait_group_read<1>();
          __syncwarp();
          const uint32_t buf = stage_smem + (n_stores & 1u) * GEMM_EPI_BUF_BYTES;
#pragma unroll
          for (int j = 0; j < 8; ++j) st_shared_v4(buf + lane * 128 + j * 16, 0u, 0u, 0u, 0u);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmap_out, buf, nb * GEMM_BLOCK_N + c * 32, out_row);
            bulk_commit_group();
          }
          ++n_stores;
        }
        continue;
      }
      const int as = iter & 1;
      const uint32_t aphase = (iter >> 1) & 1;
      ++iter;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GEMM_BLOCK_N;
      gemm_epilogue_warp<EPI_PLAIN_F32>(&p.tmap_out, taddr, stage_smem, lane, out_row, nullptr, nb, n_stores);
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace mode
