// Persistent kernel for rollout-sized batches: a whole sampler loop (n steps x (embed + L blocks + head)) in ONE
// cooperative launch.
//
// Why: the reference's inference mode is B = 1 (MoDEAgent.step, mode_agent.py:584-637). There the step is a chain of
// 860 dependent kernels of 14 token rows each; every one of them is latency, not work (7.2 ms per 10-step sample against
// 1.1 ms of weight streaming, profiles/r02_small_batch_profile.log). Here the same device code (the bodies of the row
// kernels, the cp.async attention and the weight-streaming GEMM) runs as PHASES of one resident grid: each CTA loops over
// the phase's tasks (task = blockIdx.x, += gridDim.x), then all CTAs meet at a grid barrier (one atomic per CTA, ~1 us)
// instead of a kernel boundary. The phase list is a device table built once per (batch, schedule length, sampler) by the
// host; per-step scalars (sigma, update coefficients) are read from device tables at run time, so every schedule of a
// sampler replays the same table.
//
// STATUS: correct and deterministic (tests/test_engine_gpu.py), but on B200 it is SLOWER than the CUDA graph of per-phase
// kernels it was meant to replace: 8.6 ms vs 7.2 ms per 10-step sample at B = 1 (profiles/r02_small_fused.log). The
// union of all phases needs ~196 registers, i.e. one CTA of 8 warps per SM, and a phase that is a dependent chain of
// 3-4 L2 round trips (every row phase) does not get shorter by removing its launch: the graph path overlaps those
// chains with programmatic dependent launch, the barrier serialises them. Opt-in (MODE_SMALL_FUSED=1).
//
// Memory ordering: a phase reads what other CTAs wrote in the phase before. Writers: plain stores, then
// __syncthreads(), then thread 0 fences (gpu scope) and arrives on the barrier counter. Readers: thread 0 spins with
// ld.acquire.gpu, fences, __syncthreads(). Activations are therefore never read through the non-coherent path inside this
// kernel (gemm_small_body<.., COHERENT_A = true>; the row bodies use plain loads; attention stages with cp.async.cg, which
// reads L2); weights and other launch-invariant tables may be.
#pragma once
#include "attention.cuh"
#include "gemm_small.cuh"
#include "rowwise.cuh"

namespace mode {

enum SmallPhaseKind : int { SP_EMBED = 0, SP_GEMM = 1, SP_ATTN = 2, SP_LN2 = 3, SP_COMBINE = 4, SP_HEAD = 5 };
constexpr int SMALL_PHASE_RAW = 288;  // bytes for the largest parameter struct (AttnParams: 128-byte tensor map + fields)

struct alignas(16) SmallPhase {
  int kind;     // SmallPhaseKind
  int ntasks;   // virtual blocks of the phase
  int epi;      // SP_GEMM: GemmEpilogue
  int slabs;    // SP_GEMM: column slabs (8 output columns each) per group; a task covers small_slabs_per_task(epi) of them
  alignas(16) unsigned char raw[SMALL_PHASE_RAW];
};
static_assert(sizeof(EmbedParams) <= SMALL_PHASE_RAW && sizeof(SmallGemmParams) <= SMALL_PHASE_RAW &&
                  sizeof(AttnParams) <= SMALL_PHASE_RAW && sizeof(Ln2Params) <= SMALL_PHASE_RAW &&
                  sizeof(CombineParams) <= SMALL_PHASE_RAW && sizeof(HeadParams) <= SMALL_PHASE_RAW,
              "SMALL_PHASE_RAW too small");

// slabs per GEMM task, by epilogue (= by projection): chosen so that a lane has 12-16 weight loads in flight and the
// phase is a small number of rounds over one CTA per SM (d = 1024: QKV 384 slabs -> 128 tasks, c_proj 128 -> 64,
// expert up 2 x 512 -> 512, expert down 2 x 128 -> 128)
__host__ __device__ constexpr int small_slabs_per_task(int epi) { return epi == EPI_BIAS_BF16 ? 3 : 2; }

constexpr int SMALL_EVAL_THREADS = SMALL_M_WARPS * 32;  // 256: 8 warps, as the row kernels and the small GEMM expect

__device__ __forceinline__ void small_grid_barrier(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v < epoch && ++spins > MODE_SPIN_LIMIT) {
        printf("mode: small-batch grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, v, epoch);
        __trap();
      }
    } while (v < epoch);
    __threadfence();
  }
  __syncthreads();
}

// shared memory: max(attention staging for 8 warps, GEMM partial sums) — sized by the host (small_eval_smem_bytes)
template <int DH, int MT>
constexpr int small_eval_smem_bytes() {
  constexpr int attn = SMALL_M_WARPS * 3 * 16 * (DH + 8) * 2;
  constexpr int red = SMALL_M_WARPS * 4 * MT * 128 * 4;  // S * NACC <= 4 partial tiles per warp
  return (attn > red ? attn : red) + static_cast<int>(sizeof(SmallPhase));
}

// NVEC = d / 128, DH = head dim, MT = 16-row tiles per GEMM group (token rows of the batch <= 16 * MT, T <= 16).
template <int NVEC, int DH, int MT>
__global__ void __launch_bounds__(SMALL_EVAL_THREADS) small_eval_kernel(const SmallPhase* __restrict__ phases, int n_phases,
                                                                        unsigned* barrier_counter) {
  extern __shared__ __align__(16) uint8_t small_smem[];
  SmallPhase* ph = reinterpret_cast<SmallPhase*>(small_smem);
  uint8_t* work = small_smem + sizeof(SmallPhase);
  unsigned epoch = 0;
  for (int pi = 0; pi < n_phases; ++pi) {
    // every CTA keeps its own copy of the phase descriptor in shared memory (the table itself never changes)
    {
      const uint4* src = reinterpret_cast<const uint4*>(phases + pi);
      uint4* dst = reinterpret_cast<uint4*>(ph);
      for (int i = threadIdx.x; i < static_cast<int>(sizeof(SmallPhase) / 16); i += SMALL_EVAL_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int kind = ph->kind, ntasks = ph->ntasks;
    for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
      switch (kind) {
        case SP_EMBED: embed_body<NVEC>(*reinterpret_cast<const EmbedParams*>(ph->raw), task); break;
        case SP_LN2: ln2_permute_body<NVEC>(*reinterpret_cast<const Ln2Params*>(ph->raw), task); break;
        case SP_COMBINE: combine_body<NVEC>(*reinterpret_cast<const CombineParams*>(ph->raw), task); break;
        case SP_HEAD: head_body<NVEC>(*reinterpret_cast<const HeadParams*>(ph->raw), task); break;
        case SP_ATTN:
          attention_body<DH, 1>(*reinterpret_cast<const AttnParams*>(ph->raw), task, SMALL_M_WARPS, work);
          __syncthreads();  // the staging buffers are reused by this CTA's next task / phase
          break;
        default: {
          const SmallGemmParams& g = *reinterpret_cast<const SmallGemmParams*>(ph->raw);
          const int per = small_slabs_per_task(ph->epi);
          const int tasks_per_group = (ph->slabs + per - 1) / per;
          const int slab0 = (task % tasks_per_group) * per, group = task / tasks_per_group;
          float* red = reinterpret_cast<float*>(work);
          switch (ph->epi) {
            case EPI_BIAS_BF16: gemm_small_multi_body<EPI_BIAS_BF16, MT, 3>(g, slab0, ph->slabs, group, red); break;
            case EPI_RESID_F32: gemm_small_multi_body<EPI_RESID_F32, MT, 2>(g, slab0, ph->slabs, group, red); break;
            case EPI_SWIGLU_BF16: gemm_small_multi_body<EPI_SWIGLU_BF16, MT, 2>(g, slab0, ph->slabs, group, red); break;
            default: gemm_small_multi_body<EPI_PLAIN_BF16, MT, 2>(g, slab0, ph->slabs, group, red); break;
          }
        }
      }
    }
    small_grid_barrier(barrier_counter, epoch);
  }
}

}  // namespace mode
