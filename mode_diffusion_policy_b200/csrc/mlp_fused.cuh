// Expert MLP as ONE persistent CTA-pair kernel: the grouped up-projection (+bias, SwiGLU) and the grouped
// down-projection of a NoiseBlockMoE (reference modedit.py:555-595, Mlp :220-265) share a single launch.
//
// EXPERIMENTAL, off by default (MODE_MLP_FUSED=1). Why it exists: at B=256 the up GEMM is 896 CTA-pair tiles (12.1 waves
// on 74 pairs) and the down GEMM 112 tiles of 4x the K length (1.5 waves); as two static launches they cost
// 13 + 2*4 = 21 tile-units where 18.2 are needed. Here every pair pulls its next tile from a global in-order queue (one
// atomicAdd per tile, prefetched a tile ahead), and the down tiles of an M-tile are queued LAG M-tiles after its up
// tiles so that their dependency (all 32 h column blocks of those rows written) is already satisfied when fetched.
// Measured on B200 (profiles/r01_mlp_fused_variants.log): bit-identical output, but the true dependency makes the queue
// end in a tail of 4-unit down tiles: LAG 3/5/8/18 -> 297/329/357/371 denoising-steps/s against 363-369 for the
// two-launch path, and 376 with the dependency tracking removed (an upper bound, not a valid configuration).
//
// Queue order (n_m M-tiles, NU up column blocks, ND down column blocks):
//   up(0,*) .. up(LAG-1,*) | up(g,*), down(g-LAG,*) for g = LAG..n_m-1 | down(n_m-LAG,*) .. down(n_m-1,*)
// Each tile is still computed by exactly one pair with a fixed k order: results are bit-identical to the two-launch path.
//
// Cross-CTA plumbing: the leader CTA's producer thread is the scheduler. It publishes each fetched item in a 4-entry
// ring that lives in BOTH CTAs' shared memory (remote st.shared::cluster + release.cluster arrive on the peer's `sfull`
// barrier); the ten consumers of an entry (MMA thread, 2 x 4 epilogue warps, the peer's producer) arrive on the
// leader's `sempty` barrier once they have copied it to registers. h tiles are published to the other pairs through a
// per-M-tile counter in global memory: an epilogue warp increments it (release, gpu scope) after its TMA stores of an
// up tile have completed; a producer acquires it (and fences the async proxy) before its TMA loads of h.
#pragma once
#include "gemm.cuh"

namespace mode {

constexpr int MLP_RING = 4;
constexpr int MLP_LAG = 3;           // M-tiles between an up tile row and its down tiles in the queue
constexpr int MLP_ITEM_END = -1;
constexpr int MLP_ARRIVALS_PER_TILE = 8;  // 2 CTAs x 4 epilogue warps signal each up tile

struct alignas(64) MlpParams {
  GemmParams up;    // A = permuted tokens, W = packed up weights, out = h, bias = up bias, m_tiles = up table
  GemmParams down;  // A = h, W = down weights, out = y, m_tiles = down table (same M-tiles, other weight rows)
  int* sync;        // [0] queue head, [1 + m] arrivals of M-tile m's up tiles; zeroed by the kernel before (ln2_permute)
  int flags;        // experiments (MODE_MLP_FLAGS): 1 deferred signalling, 2 no dependency tracking (timing only),
                    // 4 static round-robin instead of the atomic queue, 8 m-fastest up order with all down tiles last
};

__device__ __forceinline__ int mlp_decode(int s, int n_m, int NU, int ND, int flags) {
  // -> kind << 30 | m << 8 | nb, or MLP_ITEM_END
  if (flags & 8) {
    if (s < n_m * NU) return ((s % n_m) << 8) | (s / n_m);
    s -= n_m * NU;
    if (s < n_m * ND) return (1 << 30) | ((s % n_m) << 8) | (s / n_m);
    return MLP_ITEM_END;
  }
  const int lag = min((flags >> 8) ? (flags >> 8) : MLP_LAG, n_m);
  const int head = lag * NU;
  if (s < head) return ((s / NU) << 8) | (s % NU);
  s -= head;
  const int per = NU + ND, mid = (n_m - lag) * per;
  if (s < mid) {
    const int g = lag + s / per, r = s % per;
    return r < NU ? ((g << 8) | r) : ((1 << 30) | ((g - lag) << 8) | (r - NU));
  }
  s -= mid;
  if (s < lag * ND) return (1 << 30) | ((n_m - lag + s / ND) << 8) | (s % ND);
  return MLP_ITEM_END;
}

__device__ __forceinline__ void mlp_wait_rows_ready(const int* counter, int target) {
  uint32_t spins = 0;
  int v;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    if (v >= target) break;
    if (++spins > MODE_SPIN_LIMIT) {
      printf("mode: expert MLP dependency wait timed out (block %d, have %d of %d)\n", blockIdx.x, v, target);
      __trap();
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");  // the acquired h rows are read through the async proxy (TMA)
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    mlp_fused_2cta_kernel(const __grid_constant__ MlpParams p) {
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
  constexpr int kSchedBar0 = 2 * G2_STAGES + 5;
  auto sfull_bar = [&](int s) { return bar_base + 8u * (kSchedBar0 + s); };
  auto sempty_bar = [&](int s) { return bar_base + 8u * (kSchedBar0 + MLP_RING + s); };
  volatile int* ring = reinterpret_cast<volatile int*>(bars + kSchedBar0 + 2 * MLP_RING);
  const uint32_t ring_base = smem_u32(const_cast<int*>(ring));
  static_assert((kSchedBar0 + 2 * MLP_RING) * 8 + MLP_RING * 4 <= 256, "barrier block overflows its 256 bytes");
  const uint32_t epi_smem = smem_base + G2_STAGES * G2_STAGE_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.up.tmap_a);
    tma_prefetch_desc(&p.up.tmap_w);
    tma_prefetch_desc(&p.up.tmap_out);
    tma_prefetch_desc(&p.down.tmap_a);
    tma_prefetch_desc(&p.down.tmap_w);
    tma_prefetch_desc(&p.down.tmap_out);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(full_bar(s), 2);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 256);
    }
    for (int s = 0; s < MLP_RING; ++s) {
      mbar_init(sfull_bar(s), 1);    // each CTA's own: the scheduler's arrival
      mbar_init(sempty_bar(s), 10);  // leader's: MMA thread + 4 epilogue warps (leader), producer + 4 epilogue warps (peer)
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

  pdl_wait();  // the routing tables, permuted tokens and the zeroed queue come from the kernels before
  const int n_m = *p.up.num_m_tiles;
  const int NU = p.up.n_blocks, ND = p.down.n_blocks;
  const int ready_target = MLP_ARRIVALS_PER_TILE * NU;

  // consumer side of the scheduler ring: wait for entry `it`, copy it, release the slot (one arrival per caller)
  auto take_item = [&](int it) -> int {
    const int slot = it % MLP_RING;
    mbar_wait_cluster(sfull_bar(slot), (it / MLP_RING) & 1);  // the peer's entry was written from the leader CTA
    const int item = ring[slot];
    return item;
  };
  auto release_item = [&](int it) {
    const int slot = it % MLP_RING;
    if (rank == 0)
      mbar_arrive(sempty_bar(slot));
    else
      mbar_arrive_cluster(mapa_cluster(sempty_bar(slot), 0));
  };
  // loads of one tile's k-blocks (this CTA's 128 A rows and its half of the weight tile)
  int stage = 0;
  uint32_t phase = 0;
  auto produce_tile = [&](int item) {
    const int kind = item >> 30, m = (item >> 8) & 0x3fffff, nb = item & 0xff;
    const GemmParams& g = kind ? p.down : p.up;
    const GemmMTile tile = g.m_tiles[m];
    const int a_row = tile.a_row0 + static_cast<int>(rank) * GEMM_BLOCK_M;
    const int w_row = g.w_row_off + tile.w_row_base + nb * GEMM_BLOCK_N + static_cast<int>(rank) * G2_HALF_N;
    if (kind && !(p.flags & 2)) mlp_wait_rows_ready(p.sync + 1 + m, ready_target);
    const int k_blocks = g.k_blocks;
    for (int kb = 0; kb < k_blocks; ++kb) {
      mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t a_dst = smem_base + stage * G2_STAGE_BYTES;
      const uint32_t b_dst = a_dst + GEMM_A_BYTES;
      const uint32_t leader_full = mapa_cluster(full_bar(stage), 0);
      if (rank == 0)
        mbar_arrive_expect_tx(full_bar(stage), 2 * G2_STAGE_BYTES);
      else
        mbar_arrive_cluster(leader_full);
      tma_load_2d_2sm(a_dst, &g.tmap_a, leader_full, kb * GEMM_BLOCK_K, a_row);
      tma_load_2d_2sm(b_dst, &g.tmap_w, leader_full, kb * GEMM_BLOCK_K, w_row);
      if (++stage == G2_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  };

  if (warp == 0) {
    if (lane == 0 && rank == 0) {
      // ===================== scheduler + TMA producer (leader CTA) =====================
      const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
      const bool stat = p.flags & 4;
      int next_s = stat ? pair : atomicAdd(p.sync, 1);
      for (int it = 0;; ++it) {
        const int item = mlp_decode(next_s, n_m, NU, ND, p.flags);
        const int slot = it % MLP_RING;
        mbar_wait(sempty_bar(slot), ((it / MLP_RING) & 1) ^ 1);
        ring[slot] = item;
        st_shared_cluster_u32(mapa_cluster(ring_base + 4u * slot, 1), static_cast<uint32_t>(item));
        mbar_arrive(sfull_bar(slot));
        mbar_arrive_release_cluster(mapa_cluster(sfull_bar(slot), 1));
        if (item == MLP_ITEM_END) break;
        next_s = stat ? next_s + n_pairs : atomicAdd(p.sync, 1);  // in flight while this tile's loads are issued
        produce_tile(item);
      }
    } else if (lane == 0) {
      // ===================== TMA producer (peer CTA) =====================
      for (int it = 0;; ++it) {
        const int item = take_item(it);
        release_item(it);
        if (item == MLP_ITEM_END) break;
        produce_tile(item);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===================== MMA issuer (leader CTA only) =====================
      constexpr uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, GEMM_BLOCK_N);
      int mstage = 0;
      uint32_t mphase = 0;
      for (int it = 0;; ++it) {
        const int item = take_item(it);
        release_item(it);
        if (item == MLP_ITEM_END) break;
        const int k_blocks = (item >> 30) ? p.down.k_blocks : p.up.k_blocks;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * GEMM_BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(full_bar(mstage), mphase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + mstage * G2_STAGE_BYTES;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + GEMM_A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
            umma_bf16_2cta(tmem_d, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2cta(empty_bar(mstage), 0b11);
          if (kb == k_blocks - 1) umma_commit_2cta(tfull_bar(as), 0b11);
          if (++mstage == G2_STAGES) {
            mstage = 0;
            mphase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const uint32_t stage_smem = epi_smem + static_cast<uint32_t>(q) * 2 * GEMM_EPI_BUF_BYTES;
    uint32_t n_stores = 0;
    int pending_m = -1;  // up tile whose completion this warp still has to signal (deferred mode)
    // publish this warp's h rows of M-tile mm: its TMA stores have completed -> visible at gpu scope -> count the arrival
    auto signal = [&](int mm) {
      if (!(p.flags & 16)) asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p.sync + 1 + mm), "r"(1) : "memory");
    };
    for (int it = 0;; ++it) {
      const int item = take_item(it);
      __syncwarp();
      if (lane == 0) release_item(it);
      if (item == MLP_ITEM_END) break;
      const int kind = item >> 30, m = (item >> 8) & 0x3fffff, nb = item & 0xff;
      // A warp may only hold an unsignalled up tile while it works on another up tile (those never wait on anything):
      // before a down tile, whose producer may be waiting for exactly this signal, flush it.
      if (kind && lane == 0 && pending_m >= 0) {
        bulk_wait_group<0>();
        signal(pending_m);
        pending_m = -1;
      }
      const GemmParams& g = kind ? p.down : p.up;
      const GemmMTile tile = g.m_tiles[m];
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      float* sbias = sbias_all + as * GEMM_BLOCK_N;
      if (!kind)
        stage_bias<EPI_SWIGLU_BF16>(p.up, sbias, g.w_row_off + tile.w_row_base + nb * GEMM_BLOCK_N, q * 32 + lane);
      else  // keep the four warps within one item of each other: the bias buffers are reused every second item
        asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const int row0 = static_cast<int>(rank) * GEMM_BLOCK_M + q * 32;
      if (row0 < tile.rows_valid) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * GEMM_BLOCK_N;
        if (kind)
          gemm_epilogue_warp<EPI_PLAIN_BF16>(&p.down.tmap_out, taddr, stage_smem, lane, tile.out_row0 + row0, sbias, nb, n_stores);
        else
          gemm_epilogue_warp<EPI_SWIGLU_BF16>(&p.up.tmap_out, taddr, stage_smem, lane, tile.out_row0 + row0, sbias, nb, n_stores);
      }
      tc_fence_before();
      if (rank == 0)
        mbar_arrive(tempty_bar(as));
      else
        mbar_arrive_cluster(mapa_cluster(tempty_bar(as), 0));
      if (lane == 0 && !(p.flags & 2)) {
        if (p.flags & 1) {
          // deferred: the previous up tile's stores are older than the two groups this up tile committed
          if (!kind) {
            if (pending_m >= 0) {
              if (row0 < tile.rows_valid)
                bulk_wait_group<2>();
              else
                bulk_wait_group<0>();
              signal(pending_m);
            }
            pending_m = m;
          }
        } else if (!kind) {
          bulk_wait_group<0>();
          signal(m);
        }
      }
    }
    if (lane == 0) {
      bulk_wait_group<0>();
      if (pending_m >= 0) signal(pending_m);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace mode
