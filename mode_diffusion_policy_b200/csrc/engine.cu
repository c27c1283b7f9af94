// Host side of libmode_engine.so: configuration, weight packing, workspace, routing tables, kernel sequencing,
// CUDA-graph capture of the DDIM loop, and the C ABI declared in include/mode_engine.h.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <utility>
#include <string>
#include <vector>

#include "../../include/mode_engine.h"
#include "attention.cuh"
#include "gemm.cuh"
#include "gemm_wgrad.cuh"
#include "gemm_small.cuh"
#include "mlp_fused.cuh"
#include "rowwise.cuh"
#include "resnet.cuh"
#include "small_eval.cuh"
#include "backward.cuh"
#include "optimizer.cuh"

using namespace mode;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU_OK(expr)                                                                                          \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess)                                                                                   \
      return fail(MODE_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)
#define RET_IF(expr)         \
  do {                       \
    int _r = (expr);         \
    if (_r != MODE_OK) return _r; \
  } while (0)

extern "C" const char* mode_last_error(void) { return g_err; }

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ small kernels
namespace {

// fp32 rows -> packed destination (bf16 or fp32), optionally with the SwiGLU interleave:
// source rows [0, half) are `projected`, [half, 2*half) are `gate` (SwishGLU.forward, modedit.py:88-90); destination
// blocks of 256 rows hold 128 projected rows followed by their 128 gate rows so that one 128x256 GEMM tile contains
// both halves of 128 hidden units and the SwiGLU product can be formed in the epilogue.
__global__ void pack_rows_kernel(const float* __restrict__ src, void* __restrict__ dst, int rows, int cols,
                                 size_t dst_row0, int swiglu_half, int to_bf16, int transpose) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(rows) * cols) return;
  const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
  if (transpose) {
    reinterpret_cast<float*>(dst)[static_cast<size_t>(c) * rows + r] = src[i];
    return;
  }
  int dr = r;
  if (swiglu_half > 0) {
    const int is_gate = r >= swiglu_half;
    const int rr = is_gate ? r - swiglu_half : r;
    dr = (rr / 128) * 256 + (is_gate ? 128 : 0) + (rr % 128);
  }
  const size_t o = (dst_row0 + dr) * static_cast<size_t>(cols) + c;
  if (to_bf16)
    reinterpret_cast<__nv_bfloat16*>(dst)[o] = __float2bfloat16_rn(src[i]);
  else
    reinterpret_cast<float*>(dst)[o] = src[i];
}

// Same mapping, four columns per thread (cols % 4 == 0, no transpose): 16-byte loads, 8- or 16-byte stores. A training loop
// re-packs every parameter after each optimiser step, so this pass runs at HBM speed rather than one element per thread.
__global__ void pack_rows_vec4_kernel(const float4* __restrict__ src, void* __restrict__ dst, int rows, int cols4,
                                      size_t dst_row0, int swiglu_half, int to_bf16) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(rows) * cols4) return;
  const int r = static_cast<int>(i / cols4), c4 = static_cast<int>(i % cols4);
  int dr = r;
  if (swiglu_half > 0) {
    const int is_gate = r >= swiglu_half;
    const int rr = is_gate ? r - swiglu_half : r;
    dr = (rr / 128) * 256 + (is_gate ? 128 : 0) + (rr % 128);
  }
  const float4 v = src[i];
  const size_t o = (dst_row0 + dr) * static_cast<size_t>(cols4) + c4;
  if (to_bf16) {
    uint2 pk;
    pk.x = mode::pack_bf16x2(v.x, v.y);
    pk.y = mode::pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[o] = pk;
  } else {
    reinterpret_cast<float4*>(dst)[o] = v;
  }
}

// out[r] = sum_k W[r,k]*vec[k] (+ add[r]) accumulated in fp64; one warp per row. Used once at weight-load time.
__global__ void matvec_f64_kernel(const float* __restrict__ W, const float* __restrict__ vec,
                                  const float* __restrict__ add, float* __restrict__ out, int rows, int cols) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  double acc = 0.0;
  for (int k = lane; k < cols; k += 32) acc += static_cast<double>(W[static_cast<size_t>(r) * cols + k]) * vec[k];
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[r] = static_cast<float>(acc + (add ? static_cast<double>(add[r]) : 0.0));
}

// Generic first router Linear for the block-level entry (arbitrary c): z[b, j] = W1[j,:] . c[b,:] + b1[j]  (fp32).
__global__ void router_hidden_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                     const float* __restrict__ c, float* __restrict__ z, int B, int Hd, int d) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= B * Hd) return;
  const int b = item / Hd, j = item % Hd;
  float acc = 0.f;
  for (int k = lane; k < d; k += 32) acc = fmaf(W1[static_cast<size_t>(j) * d + k], c[static_cast<size_t>(b) * d + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) z[item] = acc + b1[j];
}

constexpr int SCHED_COEFS = 4;  // update coefficients per sampler step (HeadParams::coefs)
struct ScheduleArg {
  float v[(1 + SCHED_COEFS) * 64];
  int n;
};
// Writes the per-step {sigma} and update-coefficient tables the captured graph reads.
__global__ void set_schedule_kernel(const ScheduleArg a, float* __restrict__ sig, float* __restrict__ coefs) {
  const int i = threadIdx.x;
  if (i < a.n) {
    sig[i] = a.v[(1 + SCHED_COEFS) * i];
    for (int c = 0; c < SCHED_COEFS; ++c) coefs[SCHED_COEFS * i + c] = a.v[(1 + SCHED_COEFS) * i + 1 + c];
  }
}

// Training-mode goal masking (MoDeDiT.mask_cond, modedit.py:882-893): every goal feature is zeroed independently with
// probability cond_mask_prob, no rescaling. Element i uses half (i & 1) of word i/2 of stream RNG_GOAL. With `grad` set
// the same mask is applied in place to the gradient w.r.t. the goal instead.
__global__ void goal_mask_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, uint32_t key, uint32_t thr) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t bits = rng_bits(key, static_cast<uint32_t>(i >> 1));
  const bool dropped = ((i & 1) ? (bits >> 16) : (bits & 0xffffu)) < thr;
  out[i] = dropped ? 0.f : in[i];
}

// noised = action + noise * sigma  (GCDenoiser.loss, score_wrappers.py:59)
__global__ void noise_actions_kernel(const float* __restrict__ action, const float* __restrict__ noise,
                                     const float* __restrict__ sigma, float* __restrict__ out, int per_sample, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fadd_rn(action[i], __fmul_rn(noise[i], sigma[i / per_sample]));
}

}  // namespace

// ------------------------------------------------------------------------------------------------ engine
constexpr int ROUTE_SLOTS = 65;   // routing-table slots: 0..63 = sampler steps of a schedule, 64 = plain evaluations
constexpr int ROUTE_SLOT_EVAL = 64;

struct WeightSpec {
  void* dst;
  size_t dst_row0;
  int rows, cols;
  int swiglu_half;
  int to_bf16;
  bool ignore;
  bool transpose;  // store [cols, rows] (fp32 only)
  bool provided;
};

// Buffers and tensor maps of one NoiseBlockMoE evaluation. Inference uses ONE set for every layer with the residual
// stream updated in place (x_in = x1 = xn = x_out); training uses one set per layer so that the backward pass finds every
// activation it needs.
struct LayerIO {
  float *x_in, *x1, *xn, *x_out, *x_out_copy;
  __nv_bfloat16 *hA, *qkv, *attn, *perm, *h, *y, *z, *hA_next;
  CUtensorMap tm_hA, tm_attn, tm_perm, tm_h;          // GEMM A-operand loads
  CUtensorMap tm_qkv_attn;                             // q|k|v boxes of the TMA-staged attention kernel
  CUtensorMap to_qkv, to_x1, to_h, to_y, to_z;         // GEMM output stores
};
struct TrainState;

struct mode_engine {
  mode_config_t cfg;
  int d, L, H, Dh, E, K, T, S, A, adim, maxB, maxM, Hd, F, obs, gdim;
  int num_sms;
  int perm_rows, max_tiles;
  float inv_sqrt_d, inv_sqrt_dh;
  bool pair;   // CTA-pair GEMM kernel (MODE_GEMM_CTA_PAIR, default on)
  bool mlp_fused;  // expert up+down projections as one dynamically scheduled launch for every batch (MODE_MLP_FUSED=1; needs pair)
  int mlp_fused_max_rows;  // default: fused up to this many token rows (B <= 128: one launch less per block is worth 1-3 %
                           // there, profiles/r02_batch_sweep.log; at B = 256 two launches with their own tile widths win)
  int* mlp_sync;   // its tile queue head + per-M-tile dependency counters
  int tile_m;  // rows per M-tile: 256 with CTA pairs, 128 otherwise
  bool small_m = true;  // rollout-sized batches (B*T <= 16 rows) use the weight-streaming GEMM (MODE_SMALL_M=0 disables)
  int trim_rows = 0;  // inference: the last block's experts only run on the action rows (MODE_TRIM_LAST=0 disables)
  bool finalized = false;
  std::map<std::string, WeightSpec> specs;
  std::vector<void*> allocs;

  // packed weights
  __nv_bfloat16 *w_qkv, *w_proj, *w_up, *w_down, *w_tok, *w_goal;
  float *b_qkv, *b_up, *ln1_g, *ln2_g, *qn_g, *kn_g, *lnf_g, *pos;
  float *sig_w1, *sig_b1, *sig_w2, *sig_u, *sig_v;
  float *r_w1, *r_b1, *r_w2, *r_b2, *r_a, *r_b;
  float *w_act, *w_out, *b_out;
  float* stage;
  size_t stage_elems;

  // workspace
  float *x, *cvec, *xnorm, *state_tok, *goal_tok, *x_work, *sig_dev, *coefs_dev, *tok_sqerr, *zbuf;
  float* d_prev;  // previous step's denoised actions (multistep samplers)
  float *x_probe = nullptr, *hist = nullptr, *prog_dev = nullptr, *noise_buf = nullptr;  // sampler programs (mode_sample_program)
  // persistent small-batch kernel (small_eval.cuh): phase tables per sampler loop, recorded by the enqueue functions
  std::vector<SmallPhase>* rec = nullptr;                 // non-null while a table is being recorded (nothing is launched)
  struct SmallProgram { SmallPhase* dev; int n; };
  std::map<std::string, SmallProgram> small_programs;
  unsigned* small_barrier = nullptr;
  int small_prefetch_mask = -1;  // bit EPI_*: that weight-streaming GEMM prefetches its weights ahead of the dependency wait
                                 // (MODE_SMALL_PREFETCH=mask; -1 = by batch size, see small_gemm)
  bool band_down = true;                                  // MODE_GEMM_BAND=0: column-block-major tile order for the down GEMM
  bool small_fused = false;                               // MODE_SMALL_FUSED=1 opts in (measured slower, see mode_create)
  std::map<std::string, cudaGraphExec_t> prog_graphs;
  std::map<std::string, int64_t> prog_graph_launches;
  float *in_state, *in_goal, *in_x;  // device staging of the *_host entry points
  __nv_bfloat16 *hA, *qkv, *attn, *perm, *hbuf, *ybuf, *st_bf16, *goal_bf16;
  int *topk_idx, *sel_idx, *pos_tab, *num_tiles, *dense_counts;
  float *topk_w, *sel_w, *probs, *logits;
  GemmMTile *up_tiles, *down_tiles, *downT_tiles, *dense_tiles;
  WgradProblem *wg_up = nullptr, *wg_down = nullptr;  // training only
  unsigned long long *usage, *tokens;
  int dense_cap;  // tiles per dense table
  int cur_B = -1;
  // dense M-tile tables per batch size seen so far: switching between batch sizes (a rollout server alternating B) only
  // swaps pointers and rebuilds tensor maps on the host — no device synchronisation, and the CUDA graphs / phase tables
  // captured for the other batch sizes stay valid (they reference their own tables)
  struct BatchCtx { GemmMTile* tiles; int* counts; };
  std::map<int, BatchCtx> batch_ctx;
  int last_slot = ROUTE_SLOT_EVAL;  // slot holding the routing of the most recent evaluation (mode_get_routing)

  CUtensorMap tm_hA, tm_attn, tm_perm, tm_h, tm_st, tm_goal;
  CUtensorMap tm_wqkv, tm_wproj, tm_wup, tm_wdown, tm_wtok, tm_wgoal;
  float4* sk_partials;  // stream-K workspace: 128 KB per CTA of the GEMM grid
  int* sk_flags;
  CUtensorMap to_h, to_y;                          // grouped outputs (padded rows, fixed extent)
  CUtensorMap to_qkv, to_x, to_state, to_goal;     // dense outputs: extent = exact rows of the current batch

  cudaStream_t cap_stream = nullptr;
  std::map<std::pair<int, int>, CUtensorMap> w_maps;  // (weight matrix id, tile width) -> weight map with a width/2-row box
  std::map<std::pair<int, int>, cudaGraphExec_t> graphs;
  std::map<std::pair<int, int>, int64_t> graph_launches;
  int64_t launch_count = 0;
  LayerIO io;                   // inference buffer set (rebuilt by ensure_batch)
  bool train_weights_dirty = true;  // transposed weight copies of the training path are stale
  // Per-block "packed weights are current" events (mode_weights_record_ready): a sharded data-parallel optimizer finishes
  // block l's weights on a side stream while the next step's forward is already running; the first launch that reads
  // block l waits for its event (training forward: block by block; every other entry: all of them up front).
  std::vector<cudaEvent_t> block_w_ready;
  std::vector<char> block_w_pending;
  bool defer_block_waits = false;   // true while mode_train_step enqueues
  cudaEvent_t weights_ready = nullptr;  // recorded on the default stream after the last (re)pack
  bool weights_wait_pending = false;
  TrainState* train = nullptr;  // lazily created by the first training call

  // stochastic training mode (mode_train_set_stochastic): dropout probabilities, multinomial routing, RNG position
  struct {
    float p_attn = 0.f, p_mlp = 0.f, p_goal = 0.f, p_embed = 0.f;
    int multinomial = 0;
    unsigned long long seed = 0;
    uint32_t step = 0;
  } stoch;
  bool stoch_active = false;  // true only while mode_train_step enqueues its forward/backward
  bool train_forward = false; // true while mode_train_step enqueues (its saved activations need the tensor-memory path)
  // token-level routing tables [L][maxB*T][K] (multinomial routing draws per token), permuted-row -> token map
  int *tok_topk_idx = nullptr, *tok_sel_idx = nullptr, *tok_pos = nullptr, *row_token = nullptr;
  float *tok_topk_w = nullptr, *tok_sel_w = nullptr, *goal_masked = nullptr;

  // optional per-kernel-class timing (mode_profile_eval): event pairs around every launch of one evaluation
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_cls;
};

enum ProfClass { PC_ROUTE = 0, PC_EMBED, PC_QKV, PC_ATTN, PC_PROJ, PC_LN2, PC_UP, PC_DOWN, PC_COMBINE, PC_HEAD, PC_COND, PC_COUNT };

struct ProfScope {
  mode_engine* e;
  cudaStream_t st;
  ProfScope(mode_engine* e_, cudaStream_t st_, int cls) : e(e_), st(st_) {
    if (!e->prof_on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    e->prof_ev.push_back(a);
    e->prof_ev.push_back(b);
    e->prof_cls.push_back(cls);
  }
  ~ProfScope() {
    if (e->prof_on) cudaEventRecord(e->prof_ev.back(), st);
  }
};

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int get_encode_fn() {
  if (g_encode) return MODE_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CU_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) return fail(MODE_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return MODE_OK;
}

// Row-major matrix [rows, cols] of bf16 (elem_bytes 2) or fp32 (4); tiles of {128 bytes of columns, box_rows rows},
// 128-byte swizzle (matches make_smem_desc_sw128 for operands and the epilogue staging layout for outputs).
static int make_tmap_ex(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, int elem_bytes) {
  RET_IF(get_encode_fn());
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {128u / (uint32_t)elem_bytes, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MODE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u elem=%d", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, box_rows, elem_bytes);
  return MODE_OK;
}
// GEMM operand map: K-major bf16, box {64 columns, box_rows}
static int make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_ex(m, base, rows, cols, box_rows, 2);
}
// GEMM output map: one epilogue warp stores 32 rows x 128 bytes per bulk tensor store
static int make_out_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, int elem_bytes) {
  return make_tmap_ex(m, base, rows, cols, 32, elem_bytes);
}

template <typename T>
static int dev_alloc(mode_engine* e, T** p, size_t n, bool zero = true) {
  void* q = nullptr;
  const size_t bytes = (n ? n : 1) * sizeof(T);
  cudaError_t err = cudaMalloc(&q, bytes);
  if (err != cudaSuccess) return fail(MODE_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
  if (zero) {
    err = cudaMemset(q, 0, bytes);
    if (err != cudaSuccess) return fail(MODE_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(err));
  }
  if (e) e->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return MODE_OK;
}

// ------------------------------------------------------------------------------------------------ launches
// All hot-path kernels go through cudaLaunchKernelEx with programmatic stream serialization (PDL): see ptx.cuh.
static bool pdl_enabled() {
  static const bool on = !(getenv("MODE_PDL") && atoi(getenv("MODE_PDL")) == 0);
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
static int g_attr_done = 0;
template <int EPI>
static int gemm_set_attr() {
  CU_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  return MODE_OK;
}
template <int EPI>
static int gemm2_set_attr() {
  CU_OK(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
  return MODE_OK;
}
static int set_kernel_attrs() {
  if (g_attr_done) return MODE_OK;
  RET_IF(gemm2_set_attr<EPI_BIAS_BF16>());
  RET_IF(gemm2_set_attr<EPI_RESID_F32>());
  RET_IF(gemm2_set_attr<EPI_SWIGLU_BF16>());
  RET_IF(gemm2_set_attr<EPI_PLAIN_BF16>());
  RET_IF(gemm2_set_attr<EPI_PLAIN_F32>());
  RET_IF(gemm2_set_attr<EPI_SWIGLU_SAVE>());
  RET_IF(gemm2_set_attr<EPI_CONV_BF16>());
  RET_IF(gemm_set_attr<EPI_BIAS_BF16>());
  RET_IF(gemm_set_attr<EPI_RESID_F32>());
  RET_IF(gemm_set_attr<EPI_SWIGLU_BF16>());
  RET_IF(gemm_set_attr<EPI_PLAIN_BF16>());
  RET_IF(gemm_set_attr<EPI_PLAIN_F32>());
  RET_IF(gemm_set_attr<EPI_SWIGLU_SAVE>());
  CU_OK(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CU_OK(cudaFuncSetAttribute(mlp_fused_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
  g_attr_done = 1;
  return MODE_OK;
}

// pair = true: CTA-pair kernel (cta_group::2, 256-row M-tiles, cluster of 2); false: single-CTA kernel (128-row tiles)
static int launch_gemm(int epi, bool pair, int num_sms, cudaStream_t st, const GemmParams& p) {
  dim3 grid(pair ? (num_sms & ~1) : num_sms), block(GEMM_THREADS);
  if (pair) {
    switch (epi) {
      case EPI_BIAS_BF16: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_BIAS_BF16>, grid, block, G2_SMEM_BYTES, st, p)); break;
      case EPI_RESID_F32: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_RESID_F32>, grid, block, G2_SMEM_BYTES, st, p)); break;
      case EPI_SWIGLU_BF16: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_SWIGLU_BF16>, grid, block, G2_SMEM_BYTES, st, p)); break;
      case EPI_PLAIN_BF16: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_PLAIN_BF16>, grid, block, G2_SMEM_BYTES, st, p)); break;
      case EPI_PLAIN_F32: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_PLAIN_F32>, grid, block, G2_SMEM_BYTES, st, p)); break;
      case EPI_SWIGLU_SAVE: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_SWIGLU_SAVE>, grid, block, G2_SMEM_BYTES, st, p)); break;
      case EPI_CONV_BF16: CU_OK(launch_k(gemm_tcgen05_2cta_kernel<EPI_CONV_BF16>, grid, block, G2_SMEM_BYTES, st, p)); break;
      default: return fail(MODE_ERR_INVALID, "unknown GEMM epilogue %d", epi);
    }
  } else {
    switch (epi) {
      case EPI_BIAS_BF16: CU_OK(launch_k(gemm_tcgen05_kernel<EPI_BIAS_BF16>, grid, block, GEMM_SMEM_BYTES, st, p)); break;
      case EPI_RESID_F32: CU_OK(launch_k(gemm_tcgen05_kernel<EPI_RESID_F32>, grid, block, GEMM_SMEM_BYTES, st, p)); break;
      case EPI_SWIGLU_BF16: CU_OK(launch_k(gemm_tcgen05_kernel<EPI_SWIGLU_BF16>, grid, block, GEMM_SMEM_BYTES, st, p)); break;
      case EPI_PLAIN_BF16: CU_OK(launch_k(gemm_tcgen05_kernel<EPI_PLAIN_BF16>, grid, block, GEMM_SMEM_BYTES, st, p)); break;
      case EPI_PLAIN_F32: CU_OK(launch_k(gemm_tcgen05_kernel<EPI_PLAIN_F32>, grid, block, GEMM_SMEM_BYTES, st, p)); break;
      case EPI_SWIGLU_SAVE: CU_OK(launch_k(gemm_tcgen05_kernel<EPI_SWIGLU_SAVE>, grid, block, GEMM_SMEM_BYTES, st, p)); break;
      default: return fail(MODE_ERR_INVALID, "unknown GEMM epilogue %d", epi);
    }
  }
  CU_OK(cudaGetLastError());
  return MODE_OK;
}

// p.B == 0 only configures the kernel (opt-in shared memory); done at create time so that nothing but launches happens
// while the DDIM loop is being captured into a CUDA graph.
// Weight-streaming GEMM for <= 32 rows per group (gemm_small.cuh). n_cols: output columns (hidden units for SwiGLU);
// max_rows: upper bound on the rows of any group (selects the number of 16-row tiles per CTA).
template <int MT>
static int launch_gemm_small_mt(int epi, cudaStream_t st, const SmallGemmParams& p, int n_cols, int max_tiles) {
  const dim3 grid(n_cols / 8, max_tiles), block(SMALL_M_WARPS * 32);
  switch (epi) {
    case EPI_BIAS_BF16: CU_OK(launch_k(gemm_small_m_kernel<EPI_BIAS_BF16, MT>, grid, block, 0, st, p)); break;
    case EPI_RESID_F32: CU_OK(launch_k(gemm_small_m_kernel<EPI_RESID_F32, MT>, grid, block, 0, st, p)); break;
    case EPI_SWIGLU_BF16: CU_OK(launch_k(gemm_small_m_kernel<EPI_SWIGLU_BF16, MT>, grid, block, 0, st, p)); break;
    case EPI_PLAIN_BF16: CU_OK(launch_k(gemm_small_m_kernel<EPI_PLAIN_BF16, MT>, grid, block, 0, st, p)); break;
    case EPI_PLAIN_F32: CU_OK(launch_k(gemm_small_m_kernel<EPI_PLAIN_F32, MT>, grid, block, 0, st, p)); break;
    default: return fail(MODE_ERR_INVALID, "small-M GEMM: unsupported epilogue %d", epi);
  }
  CU_OK(cudaGetLastError());
  return MODE_OK;
}
static int launch_gemm_small(int epi, cudaStream_t st, const SmallGemmParams& p, int n_cols, int max_tiles, int max_rows) {
  if (max_rows <= 16) return launch_gemm_small_mt<1>(epi, st, p, n_cols, max_tiles);
  if (max_rows <= SMALL_M_MAX_ROWS) return launch_gemm_small_mt<2>(epi, st, p, n_cols, max_tiles);
  return fail(MODE_ERR_INVALID, "small-M GEMM: %d rows per group exceed %d", max_rows, SMALL_M_MAX_ROWS);
}
static SmallGemmParams small_params(const void* A, const void* W, int K, const GemmMTile* tiles, const int* ntiles,
                                    const float* bias, void* out, int ldo, int w_row_off) {
  SmallGemmParams p;
  p.A = reinterpret_cast<const __nv_bfloat16*>(A); p.W = reinterpret_cast<const __nv_bfloat16*>(W);
  p.m_tiles = tiles; p.num_m_tiles = ntiles; p.bias = bias; p.out = out; p.K = K; p.ldo = ldo; p.w_row_off = w_row_off;
  return p;
}

template <int DH, int MT>
static int launch_attn_t(cudaStream_t st, const AttnParams& p) {
  constexpr int smem = ATTN_WARPS * 3 * (16 * MT) * (DH + 8) * 2;
  static bool attr = false;
  if (!attr) {
    CU_OK(cudaFuncSetAttribute(attention_kernel<DH, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  if (p.B == 0) return MODE_OK;
  const int items = p.B * p.H;
  CU_OK(launch_k(attention_kernel<DH, MT>, dim3((items + ATTN_WARPS - 1) / ATTN_WARPS), dim3(ATTN_WARPS * 32), smem, st, p));
  CU_OK(cudaGetLastError());
  return MODE_OK;
}
// TMA-staged attention (attention_tma_kernel, the default for head dims 64 / 128; MODE_ATTN_TMA=0 selects the cp.async
// staging of attention_kernel). p.tmap_qkv must describe p.qkv with a {64, 16*MT}-box (make_attn_tmap).
static bool attn_tma_enabled() {
  static const bool on = !(getenv("MODE_ATTN_TMA") && atoi(getenv("MODE_ATTN_TMA")) == 0);
  return on;
}
template <int DH, int MT>
static int launch_attn_tma_t(cudaStream_t st, const AttnParams& p) {
  constexpr int smem = ATTN_WARPS * 3 * (16 * MT) * DH * 2 + 8 * ATTN_WARPS;
  static bool attr = false;
  if (!attr) {
    CU_OK(cudaFuncSetAttribute(attention_tma_kernel<DH, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  if (p.B == 0) return MODE_OK;
  const int items = p.B * p.H;
  CU_OK(launch_k(attention_tma_kernel<DH, MT>, dim3((items + ATTN_WARPS - 1) / ATTN_WARPS), dim3(ATTN_WARPS * 32), smem, st, p));
  CU_OK(cudaGetLastError());
  return MODE_OK;
}
template <int DH>
static int launch_attn_dh(cudaStream_t st, const AttnParams& p) {
  const int mt = (p.T + 15) / 16;
  if constexpr (DH >= 64) {
    if (attn_tma_enabled()) {
      switch (mt) {
        case 1: RET_IF((launch_attn_tma_t<DH, 1>(st, p))); break;
        case 2: RET_IF((launch_attn_tma_t<DH, 2>(st, p))); break;
        case 3: RET_IF((launch_attn_tma_t<DH, 3>(st, p))); break;
        case 4: RET_IF((launch_attn_tma_t<DH, 4>(st, p))); break;
        default: return fail(MODE_ERR_INVALID, "attention supports T <= 64 tokens (got %d)", p.T);
      }
      return MODE_OK;
    }
  }
  switch (mt) {
    case 1: return launch_attn_t<DH, 1>(st, p);
    case 2: return launch_attn_t<DH, 2>(st, p);
    case 3: return launch_attn_t<DH, 3>(st, p);
    case 4: return launch_attn_t<DH, 4>(st, p);
    default: return fail(MODE_ERR_INVALID, "attention supports T <= 64 tokens (got %d)", p.T);
  }
}
static int launch_attn(cudaStream_t st, const AttnParams& p, int Dh) {
  switch (Dh) {
    case 32: return launch_attn_dh<32>(st, p);
    case 64: return launch_attn_dh<64>(st, p);
    case 128: return launch_attn_dh<128>(st, p);
    default: return fail(MODE_ERR_INVALID, "head dim must be 32, 64 or 128 (got %d)", Dh);
  }
}

// ------------------------------------------------------------------------------------------------ create / destroy
static void add_spec(mode_engine* e, const std::string& name, void* dst, size_t row0, int rows, int cols, int half,
                     int bf16, bool ignore = false, bool transpose = false) {
  WeightSpec s;
  s.dst = dst;
  s.dst_row0 = row0;
  s.rows = rows;
  s.cols = cols;
  s.swiglu_half = half;
  s.to_bf16 = bf16;
  s.ignore = ignore;
  s.transpose = transpose;
  s.provided = false;
  e->specs[name] = s;
}

static void destroy_train(TrainState* t);

extern "C" void mode_destroy(mode_engine_t* e) {
  if (!e) return;
  destroy_train(e->train);
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.second);
  for (auto& g : e->prog_graphs) cudaGraphExecDestroy(g.second);
  for (auto& sp : e->small_programs) cudaFree(sp.second.dev);
  for (auto& c : e->batch_ctx) {
    cudaFree(c.second.tiles);
    cudaFree(c.second.counts);
  }
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->weights_ready) cudaEventDestroy(e->weights_ready);
  for (cudaEvent_t ev : e->block_w_ready)
    if (ev) cudaEventDestroy(ev);
  for (void* p : e->allocs) cudaFree(p);
  delete e;
}

extern "C" int mode_create(const mode_config_t* c, mode_engine_t** out) {
  if (!c || !out) return fail(MODE_ERR_INVALID, "null argument");
  *out = nullptr;
  int dev = 0;
  CU_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(MODE_ERR_CUDA, "this library only runs on sm_100 (B200); device %d is sm_%d%d — there is no fallback path",
                dev, prop.major, prop.minor);
  const int d = c->embed_dim;
  if (d <= 0 || d % 256 != 0 || d > MAX_D) return fail(MODE_ERR_INVALID, "embed_dim must be a multiple of 256 and <= %d", MAX_D);
  if (c->n_heads <= 0 || d % c->n_heads != 0) return fail(MODE_ERR_INVALID, "embed_dim %% n_heads != 0");
  const int Dh = d / c->n_heads;
  if (Dh != 32 && Dh != 64 && Dh != 128) return fail(MODE_ERR_INVALID, "head dim %d unsupported (32/64/128)", Dh);
  if (c->num_experts < 1 || c->num_experts > MAX_EXPERTS) return fail(MODE_ERR_INVALID, "num_experts must be in [1, %d]", MAX_EXPERTS);
  if (c->top_k < 1 || c->top_k > MAX_TOPK || c->top_k > c->num_experts) return fail(MODE_ERR_INVALID, "top_k must be in [1, min(%d, num_experts)]", MAX_TOPK);
  if (c->action_dim < 1 || c->action_dim > 8) return fail(MODE_ERR_INVALID, "action_dim must be in [1, 8]");
  if (c->obs_dim % 64 || c->goal_dim % 64 || c->obs_dim <= 0 || c->goal_dim <= 0)
    return fail(MODE_ERR_INVALID, "obs_dim and goal_dim must be positive multiples of 64");
  if (c->n_state_tokens < 1 || c->action_seq_len < 1) return fail(MODE_ERR_INVALID, "n_state_tokens / action_seq_len must be >= 1");
  const int T = 2 + c->n_state_tokens + c->action_seq_len;
  if (T > ATTN_MAX_TPAD) return fail(MODE_ERR_INVALID, "token sequence %d exceeds %d", T, ATTN_MAX_TPAD);
  if (c->max_batch < 1 || c->n_layers < 1) return fail(MODE_ERR_INVALID, "max_batch / n_layers must be >= 1");
  RET_IF(set_kernel_attrs());
  {
    AttnParams cfg_only{};
    cfg_only.T = T;
    RET_IF(launch_attn(nullptr, cfg_only, Dh));
  }

  mode_engine* e = new mode_engine();
  e->cfg = *c;
  e->d = d; e->L = c->n_layers; e->H = c->n_heads; e->Dh = Dh; e->E = c->num_experts; e->K = c->top_k;
  e->T = T; e->S = c->n_state_tokens; e->A = c->action_seq_len; e->adim = c->action_dim; e->maxB = c->max_batch;
  e->maxM = e->maxB * T; e->Hd = 2 * d; e->F = 4 * d; e->obs = c->obs_dim; e->gdim = c->goal_dim;
  e->num_sms = prop.multiProcessorCount;
  e->inv_sqrt_d = static_cast<float>(pow(static_cast<double>(d), -0.5));
  e->inv_sqrt_dh = static_cast<float>(pow(static_cast<double>(Dh), -0.5));
  const int L = e->L, E = e->E, K = e->K, Hd = e->Hd, F = e->F;
  const int maxM_pad = round_up(e->maxM, 256);
  {
    const char* env = getenv("MODE_GEMM_CTA_PAIR");
    e->pair = env ? atoi(env) != 0 : true;
    e->tile_m = e->pair ? 256 : 128;
    const char* small_env = getenv("MODE_SMALL_M");
    e->small_m = !(small_env && atoi(small_env) == 0);
    const char* trim_env = getenv("MODE_TRIM_LAST");
    e->trim_rows = (trim_env && atoi(trim_env) == 0) ? 0 : e->A;
    // persistent one-launch sampler for B <= 2 (small_eval.cuh): correct (tested) but measured SLOWER than the CUDA graph of
    // per-phase kernels on B200 (8.6 vs 7.2 ms per 10-step sample at B = 1, profiles/r02_small_fused.log) -> opt-in
    const char* pf_env = getenv("MODE_SMALL_PREFETCH");
    e->small_prefetch_mask = pf_env ? atoi(pf_env) : -1;
    const char* band_env = getenv("MODE_GEMM_BAND");
    e->band_down = !(band_env && atoi(band_env) == 0);
    const char* sf_env = getenv("MODE_SMALL_FUSED");
    e->small_fused = sf_env && atoi(sf_env) != 0;
    env = getenv("MODE_MLP_FUSED");
    e->mlp_fused = e->pair && (env ? atoi(env) != 0 : false);
    e->mlp_fused_max_rows = (e->pair && !env) ? 1792 : 0;
  }
  e->max_tiles = (K * e->maxM + e->tile_m - 1) / e->tile_m + E;
  e->perm_rows = e->max_tiles * e->tile_m;
  int rc = MODE_OK;
#define A_(call) do { if (rc == MODE_OK) rc = (call); } while (0)
  A_(dev_alloc(e, &e->w_qkv, (size_t)L * 3 * d * d));
  A_(dev_alloc(e, &e->w_proj, (size_t)L * d * d));
  A_(dev_alloc(e, &e->w_up, (size_t)L * E * 8 * d * d));
  A_(dev_alloc(e, &e->w_down, (size_t)L * E * d * F));
  A_(dev_alloc(e, &e->w_tok, (size_t)d * e->obs));
  A_(dev_alloc(e, &e->w_goal, (size_t)d * e->gdim));
  A_(dev_alloc(e, &e->b_qkv, (size_t)L * 3 * d));
  A_(dev_alloc(e, &e->b_up, (size_t)L * E * 8 * d));
  A_(dev_alloc(e, &e->ln1_g, (size_t)L * d));
  A_(dev_alloc(e, &e->ln2_g, (size_t)L * d));
  A_(dev_alloc(e, &e->qn_g, (size_t)L * Dh));
  A_(dev_alloc(e, &e->kn_g, (size_t)L * Dh));
  A_(dev_alloc(e, &e->lnf_g, (size_t)d));
  A_(dev_alloc(e, &e->pos, (size_t)(1 + e->A) * d));
  A_(dev_alloc(e, &e->sig_w1, (size_t)d));
  A_(dev_alloc(e, &e->sig_b1, (size_t)d));
  A_(dev_alloc(e, &e->sig_w2, (size_t)d * d));
  A_(dev_alloc(e, &e->sig_u, (size_t)d));
  A_(dev_alloc(e, &e->sig_v, (size_t)d));
  A_(dev_alloc(e, &e->r_w1, (size_t)L * Hd * d));
  A_(dev_alloc(e, &e->r_b1, (size_t)L * Hd));
  A_(dev_alloc(e, &e->r_w2, (size_t)L * E * Hd));
  A_(dev_alloc(e, &e->r_b2, (size_t)L * E));
  A_(dev_alloc(e, &e->r_a, (size_t)L * Hd));
  A_(dev_alloc(e, &e->r_b, (size_t)L * Hd));
  A_(dev_alloc(e, &e->w_act, (size_t)d * e->adim));
  A_(dev_alloc(e, &e->w_out, (size_t)e->adim * d));
  A_(dev_alloc(e, &e->b_out, (size_t)e->adim));
  e->stage_elems = (size_t)8 * d * d;
  if ((size_t)d * e->obs > e->stage_elems) e->stage_elems = (size_t)d * e->obs;
  if ((size_t)d * e->gdim > e->stage_elems) e->stage_elems = (size_t)d * e->gdim;
  A_(dev_alloc(e, &e->stage, e->stage_elems, false));
  // workspace
  const int st_rows = round_up(e->maxB * e->S, 256), goal_rows = round_up(e->maxB, 256);
  A_(dev_alloc(e, &e->x, (size_t)maxM_pad * d));
  A_(dev_alloc(e, &e->cvec, (size_t)e->maxB * d));
  A_(dev_alloc(e, &e->xnorm, (size_t)maxM_pad * d));
  A_(dev_alloc(e, &e->state_tok, (size_t)st_rows * d));
  A_(dev_alloc(e, &e->goal_tok, (size_t)goal_rows * d));
  A_(dev_alloc(e, &e->x_work, (size_t)e->maxB * e->A * e->adim));
  A_(dev_alloc(e, &e->sig_dev, 64 + (size_t)e->maxB));
  A_(dev_alloc(e, &e->coefs_dev, SCHED_COEFS * 64));
  A_(dev_alloc(e, &e->d_prev, (size_t)e->maxB * e->A * e->adim));
  A_(dev_alloc(e, &e->tok_sqerr, (size_t)e->maxB * e->A));
  A_(dev_alloc(e, &e->zbuf, (size_t)e->maxB * Hd));
  A_(dev_alloc(e, &e->in_state, (size_t)e->maxB * e->S * e->obs));
  A_(dev_alloc(e, &e->in_goal, (size_t)e->maxB * e->gdim));
  A_(dev_alloc(e, &e->in_x, (size_t)e->maxB * e->A * e->adim));
  A_(dev_alloc(e, &e->hA, (size_t)maxM_pad * d));
  A_(dev_alloc(e, &e->qkv, (size_t)maxM_pad * 3 * d));
  A_(dev_alloc(e, &e->attn, (size_t)maxM_pad * d));
  A_(dev_alloc(e, &e->perm, (size_t)e->perm_rows * d));
  A_(dev_alloc(e, &e->hbuf, (size_t)e->perm_rows * F));
  A_(dev_alloc(e, &e->ybuf, (size_t)e->perm_rows * d));
  A_(dev_alloc(e, &e->st_bf16, (size_t)st_rows * e->obs));
  A_(dev_alloc(e, &e->goal_bf16, (size_t)goal_rows * e->gdim));
  A_(dev_alloc(e, &e->topk_idx, ROUTE_SLOTS * (size_t)L * e->maxB * K));
  A_(dev_alloc(e, &e->sel_idx, ROUTE_SLOTS * (size_t)L * e->maxB * K));
  A_(dev_alloc(e, &e->pos_tab, ROUTE_SLOTS * (size_t)L * e->maxB * K));
  A_(dev_alloc(e, &e->topk_w, ROUTE_SLOTS * (size_t)L * e->maxB * K));
  A_(dev_alloc(e, &e->sel_w, ROUTE_SLOTS * (size_t)L * e->maxB * K));
  A_(dev_alloc(e, &e->probs, ROUTE_SLOTS * (size_t)L * e->maxB * E));
  A_(dev_alloc(e, &e->logits, ROUTE_SLOTS * (size_t)L * e->maxB * E));
  A_(dev_alloc(e, &e->up_tiles, ROUTE_SLOTS * (size_t)L * e->max_tiles));
  A_(dev_alloc(e, &e->down_tiles, ROUTE_SLOTS * (size_t)L * e->max_tiles));
  A_(dev_alloc(e, &e->downT_tiles, ROUTE_SLOTS * (size_t)L * e->max_tiles));
  A_(dev_alloc(e, &e->num_tiles, ROUTE_SLOTS * (size_t)L));
  e->dense_cap = maxM_pad / 128;  // enough for either tile size
  e->dense_tiles = nullptr;  // per batch size, see ensure_batch
  e->dense_counts = nullptr;
  A_(dev_alloc(e, &e->sk_partials, (size_t)e->num_sms * 8 * 8 * 128, false));
  A_(dev_alloc(e, &e->sk_flags, (size_t)e->num_sms * 4));
  A_(dev_alloc(e, &e->mlp_sync, (size_t)1 + e->max_tiles));
  A_(dev_alloc(e, &e->usage, (size_t)L * E));
  A_(dev_alloc(e, &e->tokens, (size_t)L));
  // tensor maps
  A_(make_tmap(&e->tm_hA, e->hA, maxM_pad, d, 128));
  A_(make_tmap(&e->tm_attn, e->attn, maxM_pad, d, 128));
  A_(make_tmap(&e->tm_perm, e->perm, e->perm_rows, d, 128));
  A_(make_tmap(&e->tm_h, e->hbuf, e->perm_rows, F, 128));
  A_(make_tmap(&e->tm_st, e->st_bf16, st_rows, e->obs, 128));
  A_(make_tmap(&e->tm_goal, e->goal_bf16, goal_rows, e->gdim, 128));
  A_(make_tmap(&e->tm_wqkv, e->w_qkv, (uint64_t)L * 3 * d, d, e->pair ? 128 : 256));
  A_(make_tmap(&e->tm_wproj, e->w_proj, (uint64_t)L * d, d, e->pair ? 128 : 256));
  A_(make_tmap(&e->tm_wup, e->w_up, (uint64_t)L * E * 8 * d, d, e->pair ? 128 : 256));
  A_(make_tmap(&e->tm_wdown, e->w_down, (uint64_t)L * E * d, F, e->pair ? 128 : 256));
  A_(make_tmap(&e->tm_wtok, e->w_tok, d, e->obs, e->pair ? 128 : 256));
  A_(make_tmap(&e->tm_wgoal, e->w_goal, d, e->gdim, e->pair ? 128 : 256));
  A_(make_out_tmap(&e->to_h, e->hbuf, e->perm_rows, F, 2));
  A_(make_out_tmap(&e->to_y, e->ybuf, e->perm_rows, d, 2));
  if (rc == MODE_OK && cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking) != cudaSuccess)
    rc = fail(MODE_ERR_CUDA, "cudaStreamCreate failed");
#undef A_
  if (rc != MODE_OK) {
    mode_destroy(e);
    return rc;
  }
  // expected state_dict (reference names and shapes; SURVEY.md §8b "Weights contract")
  add_spec(e, "pos_emb", e->pos, 0, 1 + e->A, d, 0, 0);
  add_spec(e, "sigma_emb.weight", e->sig_w1, 0, d, 1, 0, 0);
  add_spec(e, "sigma_emb.bias", e->sig_b1, 0, 1, d, 0, 0);
  add_spec(e, "sigma_linear.weight", e->sig_w2, 0, d, d, 0, 0);
  add_spec(e, "tok_emb.weight", e->w_tok, 0, d, e->obs, 0, 1);
  add_spec(e, "gripper_embed.weight", nullptr, 0, d, e->obs, 0, 0, true);
  add_spec(e, "goal_emb.weight", e->w_goal, 0, d, e->gdim, 0, 1);
  add_spec(e, "action_emb.weight", e->w_act, 0, d, e->adim, 0, 0, false, true);  // stored transposed [adim, d]
  add_spec(e, "ln.g", e->lnf_g, 0, 1, d, 0, 0);
  add_spec(e, "out.weight", e->w_out, 0, e->adim, d, 0, 0);
  add_spec(e, "out.bias", e->b_out, 0, 1, e->adim, 0, 0);
  for (int l = 0; l < L; ++l) {
    const std::string b = "blocks." + std::to_string(l) + ".";
    add_spec(e, b + "ln_1.g", e->ln1_g + (size_t)l * d, 0, 1, d, 0, 0);
    add_spec(e, b + "ln_2.g", e->ln2_g + (size_t)l * d, 0, 1, d, 0, 0);
    add_spec(e, b + "attn.query.weight", e->w_qkv, (size_t)l * 3 * d, d, d, 0, 1);
    add_spec(e, b + "attn.key.weight", e->w_qkv, (size_t)l * 3 * d + d, d, d, 0, 1);
    add_spec(e, b + "attn.value.weight", e->w_qkv, (size_t)l * 3 * d + 2 * d, d, d, 0, 1);
    add_spec(e, b + "attn.query.bias", e->b_qkv + (size_t)l * 3 * d, 0, 1, d, 0, 0);
    add_spec(e, b + "attn.key.bias", e->b_qkv + (size_t)l * 3 * d + d, 0, 1, d, 0, 0);
    add_spec(e, b + "attn.value.bias", e->b_qkv + (size_t)l * 3 * d + 2 * d, 0, 1, d, 0, 0);
    add_spec(e, b + "attn.c_proj.weight", e->w_proj, (size_t)l * d, d, d, 0, 1);
    add_spec(e, b + "attn.q_norm.g", e->qn_g + (size_t)l * Dh, 0, 1, Dh, 0, 0);
    add_spec(e, b + "attn.k_norm.g", e->kn_g + (size_t)l * Dh, 0, 1, Dh, 0, 0);
    add_spec(e, b + "router.router.mlp.0.weight", e->r_w1 + (size_t)l * Hd * d, 0, Hd, d, 0, 0);
    add_spec(e, b + "router.router.mlp.0.bias", e->r_b1 + (size_t)l * Hd, 0, 1, Hd, 0, 0);
    add_spec(e, b + "router.router.mlp.3.weight", e->r_w2 + (size_t)l * E * Hd, 0, E, Hd, 0, 0);
    add_spec(e, b + "router.router.mlp.3.bias", e->r_b2 + (size_t)l * E, 0, 1, E, 0, 0);
    for (int x = 0; x < E; ++x) {
      const std::string eb = b + "experts.expert_" + std::to_string(x) + ".mlp.";
      add_spec(e, eb + "0.project.weight", e->w_up, ((size_t)l * E + x) * 8 * d, 8 * d, d, F, 1);
      // bias: a column vector packed with the same row interleave
      add_spec(e, eb + "0.project.bias", e->b_up, ((size_t)l * E + x) * 8 * d, 8 * d, 1, F, 0);
      add_spec(e, eb + "2.weight", e->w_down, ((size_t)l * E + x) * d, d, F, 0, 1);
    }
  }
  *out = e;
  return MODE_OK;
}

extern "C" int mode_set_weight_on_stream(mode_engine_t* e, const char* name, const void* data, int is_device,
                                         const int64_t* shape, int ndim, void* stream) {
  if (!e || !name || !data) return fail(MODE_ERR_INVALID, "null argument");
  cudaStream_t pst = reinterpret_cast<cudaStream_t>(stream);
  auto it = e->specs.find(name);
  if (it == e->specs.end()) return fail(MODE_ERR_UNKNOWN_NAME, "'%s' is not a MoDeDiT state_dict key for this configuration", name);
  WeightSpec& s = it->second;
  size_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= (size_t)shape[i];
  if (numel != (size_t)s.rows * s.cols)
    return fail(MODE_ERR_INVALID, "'%s': expected %zu elements, got %zu", name, (size_t)s.rows * s.cols, numel);
  s.provided = true;
  e->finalized = false;
  e->train_weights_dirty = true;
  if (s.ignore) return MODE_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)((numel + threads - 1) / threads);
  if (is_device) {
    // device source: pack straight from the caller's tensor, stream-ordered on the CALLER's stream (the one its
    // parameter updates were enqueued on), no synchronisation (a training loop with a torch optimizer re-packs all
    // 686 M parameters after every step)
    if (!s.transpose && s.cols % 4 == 0 && (reinterpret_cast<uintptr_t>(data) & 15) == 0) {
      const size_t n4 = numel / 4;
      pack_rows_vec4_kernel<<<(unsigned)((n4 + threads - 1) / threads), threads, 0, pst>>>(
          reinterpret_cast<const float4*>(data), s.dst, s.rows, s.cols / 4, s.dst_row0, s.swiglu_half, s.to_bf16);
    } else {
      pack_rows_kernel<<<blocks, threads, 0, pst>>>(reinterpret_cast<const float*>(data), s.dst, s.rows, s.cols, s.dst_row0,
                                                    s.swiglu_half, s.to_bf16, s.transpose ? 1 : 0);
    }
    CU_OK(cudaGetLastError());
    return MODE_OK;
  }
  if (numel > e->stage_elems) return fail(MODE_ERR_INVALID, "'%s' larger than the staging buffer", name);
  CU_OK(cudaMemcpyAsync(e->stage, data, numel * sizeof(float), cudaMemcpyHostToDevice, pst));
  pack_rows_kernel<<<blocks, threads, 0, pst>>>(e->stage, s.dst, s.rows, s.cols, s.dst_row0, s.swiglu_half, s.to_bf16,
                                                s.transpose ? 1 : 0);
  CU_OK(cudaGetLastError());
  CU_OK(cudaStreamSynchronize(pst));  // the staging buffer and the caller's host array are reused right away
  return MODE_OK;
}

extern "C" int mode_set_weight(mode_engine_t* e, const char* name, const void* data, int is_device,
                               const int64_t* shape, int ndim) {
  return mode_set_weight_on_stream(e, name, data, is_device, shape, ndim, nullptr);
}

// Quantities derived from the packed weights (the sigma-affine collapse of DESIGN.md §5), recomputed whenever they change.
static int refresh_derived(mode_engine* e, cudaStream_t st) {
  const int d = e->d, Hd = e->Hd;
  // emb_t(s) = sigma_linear(sigma_emb(s)) = s * (W2 w1) + (W2 b1)            (modedit.py:823-832)
  matvec_f64_kernel<<<(d + 7) / 8, 256, 0, st>>>(e->sig_w2, e->sig_w1, nullptr, e->sig_u, d, d);
  matvec_f64_kernel<<<(d + 7) / 8, 256, 0, st>>>(e->sig_w2, e->sig_b1, nullptr, e->sig_v, d, d);
  // router first Linear on c = emb_t(s): s * (W1 u) + (W1 v + b1)            (modedit.py:304-310, :336)
  // all blocks in one launch each: r_w1 [L, Hd, d], r_a / r_b1 / r_b [L, Hd] are contiguous over the blocks (this runs
  // after every optimizer step of the training path, not only at load time)
  const int rows = e->L * Hd;
  matvec_f64_kernel<<<(rows + 7) / 8, 256, 0, st>>>(e->r_w1, e->sig_u, nullptr, e->r_a, rows, d);
  matvec_f64_kernel<<<(rows + 7) / 8, 256, 0, st>>>(e->r_w1, e->sig_v, e->r_b1, e->r_b, rows, d);
  CU_OK(cudaGetLastError());
  return MODE_OK;
}

extern "C" int mode_finalize_weights_on_stream(mode_engine_t* e, void* stream) {
  if (!e) return fail(MODE_ERR_INVALID, "null engine");
  cudaStream_t pst = reinterpret_cast<cudaStream_t>(stream);
  for (auto& kv : e->specs)
    if (!kv.second.provided && !kv.second.ignore) return fail(MODE_ERR_STATE, "weight '%s' was never set", kv.first.c_str());
  RET_IF(refresh_derived(e, pst));
  // no host synchronisation: host-source weights were already synchronised in mode_set_weight, device-source packing is
  // stream-ordered on `stream`; the first call on any stream waits for this event (ensure_batch)
  if (!e->weights_ready) CU_OK(cudaEventCreateWithFlags(&e->weights_ready, cudaEventDisableTiming));
  CU_OK(cudaEventRecord(e->weights_ready, pst));
  e->weights_wait_pending = true;
  e->finalized = true;
  return MODE_OK;
}

extern "C" int mode_finalize_weights(mode_engine_t* e) { return mode_finalize_weights_on_stream(e, nullptr); }

// ------------------------------------------------------------------------------------------------ batch tables
// layer < 0: every block with a pending event
static int wait_block_weights(mode_engine* e, cudaStream_t st, int layer) {
  if (e->block_w_pending.empty()) return MODE_OK;
  const int l0 = layer < 0 ? 0 : layer, l1 = layer < 0 ? e->L : layer + 1;
  for (int l = l0; l < l1; ++l)
    if (e->block_w_pending[l]) {
      CU_OK(cudaStreamWaitEvent(st, e->block_w_ready[l], 0));
      e->block_w_pending[l] = 0;
    }
  return MODE_OK;
}

static int ensure_batch(mode_engine* e, int B, cudaStream_t st) {
  if (B < 1 || B > e->maxB) return fail(MODE_ERR_INVALID, "batch %d outside [1, max_batch=%d]", B, e->maxB);
  if (!e->finalized) return fail(MODE_ERR_STATE, "mode_finalize_weights has not been called");
  if (!e->defer_block_waits) RET_IF(wait_block_weights(e, st, -1));
  if (e->weights_wait_pending) {
    // weights were (re)packed on the default stream without host synchronisation: order this stream after them
    CU_OK(cudaStreamWaitEvent(st, e->weights_ready, 0));
    e->weights_wait_pending = false;
  }
  if (B == e->cur_B) return MODE_OK;
  auto ctx = e->batch_ctx.find(B);
  if (ctx == e->batch_ctx.end()) {
    if (e->batch_ctx.size() >= 32) {  // bounded cache: drop everything captured for the old batch sizes
      CU_OK(cudaDeviceSynchronize());
      for (auto& g : e->graphs) cudaGraphExecDestroy(g.second);
      e->graphs.clear();
      e->graph_launches.clear();
      for (auto& g : e->prog_graphs) cudaGraphExecDestroy(g.second);
      e->prog_graphs.clear();
      e->prog_graph_launches.clear();
      for (auto& sp : e->small_programs) cudaFree(sp.second.dev);
      e->small_programs.clear();
      for (auto& c : e->batch_ctx) {
        cudaFree(c.second.tiles);
        cudaFree(c.second.counts);
      }
      e->batch_ctx.clear();
    }
    std::vector<GemmMTile> tiles(3 * (size_t)e->dense_cap);
    int counts[3];
    const int rows_[3] = {B * e->T, B * e->S, B};
    for (int k = 0; k < 3; ++k) {
      const int tm = e->tile_m;
      const int n = (rows_[k] + tm - 1) / tm;
      counts[k] = n;
      for (int i = 0; i < n; ++i) {
        GemmMTile t;
        t.a_row0 = i * tm;
        t.out_row0 = i * tm;
        t.rows_valid = rows_[k] - i * tm < tm ? rows_[k] - i * tm : tm;
        t.w_row_base = 0;
        tiles[(size_t)k * e->dense_cap + i] = t;
      }
    }
    mode_engine::BatchCtx c{nullptr, nullptr};
    RET_IF(dev_alloc<GemmMTile>(nullptr, &c.tiles, tiles.size(), false));
    RET_IF(dev_alloc<int>(nullptr, &c.counts, 3, false));
    // fresh allocations nothing else reads: plain synchronous copies, no device-wide synchronisation
    CU_OK(cudaMemcpy(c.tiles, tiles.data(), tiles.size() * sizeof(GemmMTile), cudaMemcpyHostToDevice));
    CU_OK(cudaMemcpy(c.counts, counts, sizeof(counts), cudaMemcpyHostToDevice));
    ctx = e->batch_ctx.emplace(B, c).first;
  }
  e->dense_tiles = ctx->second.tiles;
  e->dense_counts = ctx->second.counts;
  const int rows[3] = {B * e->T, B * e->S, B};
  // output maps clip at the exact row count, so partially filled 32-row store boxes never touch rows >= M
  RET_IF(make_out_tmap(&e->to_qkv, e->qkv, rows[0], 3 * e->d, 2));
  RET_IF(make_out_tmap(&e->to_x, e->x, rows[0], e->d, 4));
  RET_IF(make_out_tmap(&e->to_state, e->state_tok, rows[1], e->d, 4));
  RET_IF(make_out_tmap(&e->to_goal, e->goal_tok, rows[2], e->d, 4));
  {
    LayerIO& io = e->io;
    io.x_in = io.x1 = io.xn = io.x_out = e->x;
    io.x_out_copy = nullptr;
    io.hA = io.hA_next = e->hA; io.qkv = e->qkv; io.attn = e->attn; io.perm = e->perm; io.h = e->hbuf; io.y = e->ybuf;
    io.z = nullptr;
    io.tm_hA = e->tm_hA; io.tm_attn = e->tm_attn; io.tm_perm = e->tm_perm; io.tm_h = e->tm_h;
    io.to_qkv = e->to_qkv; io.to_x1 = e->to_x; io.to_h = e->to_h; io.to_y = e->to_y;
    RET_IF(make_tmap(&io.tm_qkv_attn, e->qkv, rows[0], 3 * e->d, 16 * ((e->T + 15) / 16)));
  }
  e->cur_B = B;
  return MODE_OK;
}

static GemmParams gemm_params(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tout,
                              const GemmMTile* tiles, const int* ntiles, int N, int Kdim, const float* bias) {
  GemmParams p;
  p.tmap_a = ta;
  p.tmap_w = tw;
  p.tmap_out = tout;
  p.m_tiles = tiles;
  p.num_m_tiles = ntiles;
  p.n_blocks = N / GEMM_BLOCK_N;
  p.bn = GEMM_BLOCK_N;
  p.n_total = N;
  p.out_ptr = nullptr;
  p.ldo = 0;
  p.out_rows = 0;
  p.conv = ConvEpilogue{nullptr, 0, 0, nullptr, nullptr, 1, 0, 0, 0};
  p.band = 0;
  p.n_taps = 0;
  p.kb_per_tap = 1;
  for (int i = 0; i < 9; ++i) p.tap_off[i] = 0;
  p.k_blocks = Kdim / GEMM_BLOCK_K;
  p.bias = bias;
  p.w_row_off = 0;
  p.sk_enable = 0;
  p.sk_partials = nullptr;
  p.sk_flags = nullptr;
  p.drop = DropoutSpec{0u, 0u, 1.0f};
  p.row_token = nullptr;
  p.drop_rows_per_expert = 1;
  p.drop_E = 1;
  p.drop_half_F = 0;
  return p;
}
// Tile width of a CTA-pair GEMM with `n_m` M-tiles, N output columns and k_blocks 64-wide K steps: the multiple of 16 that
// minimises  waves x (k_blocks x max(width, 240) + 12 x width + 1024).  The three terms are the measured structure of a
// tile's time (profiles/r02_gemm_tile_widths.log, r02_gemm_small_m.log): a k-block costs the same below ~240 columns (the
// 6-stage TMA pipeline is latency-bound there: K = 4096 at M = 896 takes 24.4 us at every width from 64 to 256), the
// epilogue is proportional to the width (K = 1024 at M = 448: 18.2 -> 13.4 us from 256 to 128), and every tile pays a fixed
// ramp. So narrow tiles pay off when they fill idle CTA pairs or whole waves, never by shortening the K loop; ties go to
// the wider tile. At B = 256 (14 M-tiles, 74 pairs): QKV N = 3072 -> 208 (210 tiles = 3 waves of 0.81 instead of 168 =
// 3 waves of 1.0), c_proj -> 208 (70 tiles, one wave), expert down (28 M-tiles) -> 208 (140 tiles = 2 waves).
static int choose_bn(int n_m, int N, int pairs, int k_blocks) {
  static const int forced = getenv("MODE_GEMM_BN") ? atoi(getenv("MODE_GEMM_BN")) : 0;
  if (forced >= 64 && forced <= 256 && forced % 16 == 0) return forced;
  int best = GEMM_BLOCK_N;
  long best_cost = -1;
  for (int bn = GEMM_BLOCK_N; bn >= 64; bn -= 16) {
    const long tiles = (long)n_m * ((N + bn - 1) / bn);
    const long per_tile = (long)k_blocks * (bn > 240 ? bn : 240) + 12L * bn + 1024;
    const long cost = ((tiles + pairs - 1) / pairs) * per_tile;
    if (best_cost < 0 || cost < best_cost) {
      best = bn;
      best_cost = cost;
    }
  }
  return best;
}

// Switches a pair-kernel GEMM to `bn`-wide tiles: weight map with a bn/2-row box (cached per weight matrix and width),
// raw output pointer for the columns the TMA store chunks do not cover.
static int narrow_gemm(mode_engine* e, GemmParams& p, int bn, int which_w, const void* w_base, uint64_t w_rows, int Kdim,
                       void* out_ptr, int ldo, int out_rows) {
  if (!e->pair || bn == GEMM_BLOCK_N) return MODE_OK;
  const auto key = std::make_pair(which_w, bn);
  auto it = e->w_maps.find(key);
  if (it == e->w_maps.end()) {
    CUtensorMap m;
    RET_IF(make_tmap(&m, w_base, w_rows, (uint64_t)Kdim, (uint32_t)(bn / 2)));
    it = e->w_maps.emplace(key, m).first;
  }
  p.tmap_w = it->second;
  p.bn = bn;
  p.n_blocks = (p.n_total + bn - 1) / bn;
  p.out_ptr = out_ptr;
  p.ldo = ldo;
  p.out_rows = out_rows;
  p.sk_enable = 0;  // stream-K partials are laid out for 256-wide tiles
  return MODE_OK;
}

// Stream-K over the last partial wave (QKV, expert up/down). Correct (tests/test_kernels_gpu.py) but OFF by default:
// measured on B200 the fp32 partial round trip costs as much as the removed tail (profiles/r01_gemm_variants.log), and
// splitting K makes the accumulation order depend on the batch size, which would break the bit-exact batch-split
// invariance of the sampler. MODE_GEMM_STREAM_K=1 turns it on for experiments.
static void enable_stream_k(mode_engine* e, GemmParams& p) {
  static const bool on = getenv("MODE_GEMM_STREAM_K") && atoi(getenv("MODE_GEMM_STREAM_K")) != 0;
  if (!e->pair || !on) return;
  p.sk_enable = 1;
  p.sk_partials = e->sk_partials;
  p.sk_flags = e->sk_flags;
}

// embed_dim is a multiple of 256 -> d/128 float4 vectors per lane, compile-time for register-resident rows
#define LAUNCH_ROW_KERNEL(KERNEL, d, grid, st, params)                                        \
  do {                                                                                        \
    switch ((d) / 128) {                                                                      \
      case 2: CU_OK(launch_k(KERNEL<2>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                  \
      case 4: CU_OK(launch_k(KERNEL<4>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                  \
      case 6: CU_OK(launch_k(KERNEL<6>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                  \
      case 8: CU_OK(launch_k(KERNEL<8>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                  \
      case 10: CU_OK(launch_k(KERNEL<10>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                \
      case 12: CU_OK(launch_k(KERNEL<12>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                \
      case 14: CU_OK(launch_k(KERNEL<14>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                \
      case 16: CU_OK(launch_k(KERNEL<16>, dim3(grid), dim3(ROW_WARPS * 32), 0, (st), params)); break;                \
      default: return fail(MODE_ERR_INVALID, "embed_dim %d unsupported by the row kernels", (d)); \
    }                                                                                         \
  } while (0)

static inline unsigned row_blocks(int rows) { return (unsigned)((rows + ROW_WARPS - 1) / ROW_WARPS); }

// ---- recording of the persistent small-batch kernel's phase table (small_eval.cuh)
template <typename P>
static void rec_phase(mode_engine* e, int kind, const P& params, int ntasks, int epi = 0, int slabs = 1) {
  static_assert(sizeof(P) <= SMALL_PHASE_RAW, "parameter struct does not fit a phase descriptor");
  SmallPhase ph;
  memset(&ph, 0, sizeof(ph));
  ph.kind = kind; ph.ntasks = ntasks; ph.epi = epi; ph.slabs = slabs;
  memcpy(ph.raw, &params, sizeof(P));
  e->rec->push_back(ph);
}
// LAUNCH_ROW_KERNEL or, while recording, one phase of `kind`
#define ROW_PHASE(KIND, KERNEL, d, grid, st, params)                  \
  do {                                                                \
    if (e->rec)                                                       \
      rec_phase(e, (KIND), (params), (int)(grid));                    \
    else                                                              \
      LAUNCH_ROW_KERNEL(KERNEL, d, grid, st, params);                 \
  } while (0)
static int small_gemm(mode_engine* e, int epi, cudaStream_t st, const SmallGemmParams& p_in, int n_cols, int max_tiles, int max_rows,
                      bool tables_final = true) {
  SmallGemmParams p = p_in;
  // measured (profiles/r02_small_fused.log): with one 16-row tile per group (B = 1) prefetching the expert weights ahead of
  // the wait costs more than it hides (5.5 -> 6.8 ms per sample), QKV + c_proj are a small win; with two tiles (B = 2) all four help
  const int mask = e->small_prefetch_mask >= 0 ? e->small_prefetch_mask : (max_rows <= 16 ? 3 : 15);
  p.prefetch = (tables_final && (mask >> epi & 1)) ? 1 : 0;
  if (e->rec) {
    const int per = small_slabs_per_task(epi), slabs = n_cols / 8;
    rec_phase(e, SP_GEMM, p, ((slabs + per - 1) / per) * max_tiles, epi, slabs);
    return MODE_OK;
  }
  return launch_gemm_small(epi, st, p, n_cols, max_tiles, max_rows);
}

// obs / goal token embeddings: computed once per trajectory, not per denoising step.
static int enqueue_cond(mode_engine* e, cudaStream_t st, int B, const float* state_dev, const float* goal_dev) {
  ProfScope ps(e, st, PC_COND);
  const size_t n_st = (size_t)B * e->S * e->obs / 4, n_g = (size_t)B * e->gdim / 4;
  cast_bf16_kernel<<<(unsigned)((n_st + 255) / 256), 256, 0, st>>>(state_dev, e->st_bf16, n_st);
  cast_bf16_kernel<<<(unsigned)((n_g + 255) / 256), 256, 0, st>>>(goal_dev, e->goal_bf16, n_g);
  CU_OK(cudaGetLastError());
  const bool small = e->small_m && !e->train_forward && B * e->S <= SMALL_M_MAX_ROWS && e->obs % 256 == 0 && e->gdim % 256 == 0;
  if (small) {
    RET_IF(launch_gemm_small(EPI_PLAIN_F32, st, small_params(e->st_bf16, e->w_tok, e->obs, e->dense_tiles + e->dense_cap,
                                                             e->dense_counts + 1, nullptr, e->state_tok, e->d, 0), e->d, 1, B * e->S));
    RET_IF(launch_gemm_small(EPI_PLAIN_F32, st, small_params(e->goal_bf16, e->w_goal, e->gdim, e->dense_tiles + 2 * e->dense_cap,
                                                             e->dense_counts + 2, nullptr, e->goal_tok, e->d, 0), e->d, 1, B));
    e->launch_count += 4;
    return MODE_OK;
  }
  GemmParams p = gemm_params(e->tm_st, e->tm_wtok, e->to_state, e->dense_tiles + e->dense_cap, e->dense_counts + 1,
                             e->d, e->obs, nullptr);
  RET_IF(launch_gemm(EPI_PLAIN_F32, e->pair, e->num_sms, st, p));
  p = gemm_params(e->tm_goal, e->tm_wgoal, e->to_goal, e->dense_tiles + 2 * e->dense_cap, e->dense_counts + 2, e->d,
                  e->gdim, nullptr);
  RET_IF(launch_gemm(EPI_PLAIN_F32, e->pair, e->num_sms, st, p));
  e->launch_count += 4;
  return MODE_OK;
}

// Routing tables of layer l as the routed kernels see them: `units` routing units of `rt` token rows each.
struct RouteView {
  const int* sel_idx;
  const float* sel_w;
  int* pos;
  int units, rt;
  bool per_token;
};
static inline bool token_routing(const mode_engine* e) { return e->stoch_active && e->stoch.multinomial; }
static RouteView route_view(mode_engine* e, int B, size_t lt, int l) {
  if (token_routing(e)) {
    const size_t o = (size_t)l * B * e->T * e->K;
    return RouteView{e->tok_sel_idx + o, e->tok_sel_w + o, e->tok_pos + o, B * e->T, 1, true};
  }
  const size_t o = lt * B * e->K;
  return RouteView{e->sel_idx + o, e->sel_w + o, e->pos_tab + o, B, e->T, false};
}
static DropoutSpec dropout_spec(const mode_engine* e, float p, uint32_t stream, int layer) {
  if (!e->stoch_active || p <= 0.f) return DropoutSpec{0u, 0u, 1.0f};
  return DropoutSpec{rng_key(e->stoch.seed, e->stoch.step, stream, (uint32_t)layer), drop_threshold(p), 1.0f / (1.0f - p)};
}

// Routes `n_slots` evaluations at once: slot s uses sigma + s * sigma_slot_stride (a whole sampler schedule in one
// launch), tables [slot][L][B][..].
static int enqueue_routing(mode_engine* e, cudaStream_t st, int B, const float* sigma, int stride, const float* z_explicit,
                           int layer0, int n_layers, int slot0 = ROUTE_SLOT_EVAL, int n_slots = 1,
                           int sigma_slot_stride = 0, int trim_rows = 0) {
  ProfScope ps(e, st, PC_ROUTE);
  RouterParams r;
  r.sc = StepScalars{sigma, stride, e->cfg.sigma_data};
  r.ra = e->r_a; r.rb = e->r_b; r.w2 = e->r_w2; r.b2 = e->r_b2;
  r.z_explicit = z_explicit;
  r.topk_idx = e->topk_idx; r.topk_w = e->topk_w; r.sel_idx = e->sel_idx; r.sel_w = e->sel_w;
  r.probs = e->probs; r.logits = e->logits;
  r.L = n_layers; r.layer0 = layer0; r.B = B; r.E = e->E; r.K = e->K; r.Hd = e->Hd; r.normalize = e->cfg.router_normalize;
  r.slot0 = slot0; r.Ltot = e->L; r.sigma_slot_stride = sigma_slot_stride;
  const bool tok = token_routing(e);
  if (tok && (stride != 1 || n_slots != 1 || layer0 != 0 || n_layers != e->L))
    return fail(MODE_ERR_STATE, "per-token multinomial routing needs per-sample sigma and a full-network evaluation");
  r.multinomial = tok ? 1 : 0; r.T = e->T; r.seed = e->stoch.seed; r.step = e->stoch.step;
  r.tok_topk_idx = e->tok_topk_idx; r.tok_topk_w = e->tok_topk_w; r.tok_sel_idx = e->tok_sel_idx; r.tok_sel_w = e->tok_sel_w;
  const int distinct_rows = (stride == 0 && !z_explicit) ? 1 : B;
  if (distinct_rows > 1)  // many rows: a warp per row, 8 rows of a layer per CTA
    CU_OK(launch_k(router_kernel<true>, dim3(n_slots * n_layers * ((distinct_rows + ROW_WARPS - 1) / ROW_WARPS)), dim3(ROW_WARPS * 32), 0, st, r));
  else  // one row per (slot, layer): a CTA per row
    CU_OK(launch_k(router_kernel<false>, dim3(n_slots * n_layers), dim3(ROW_WARPS * 32), 0, st, r));
  PlanParams pl;
  pl.sel_idx = e->sel_idx; pl.pos = e->pos_tab; pl.up_tiles = e->up_tiles; pl.down_tiles = e->down_tiles;
  pl.downT_tiles = e->downT_tiles; pl.wg_up = e->wg_up; pl.wg_down = e->wg_down;
  pl.num_tiles = e->num_tiles; pl.usage = e->usage; pl.tokens = e->tokens;
  pl.L = e->L; pl.B = B; pl.K = e->K; pl.E = e->E; pl.T = e->T; pl.max_tiles = e->max_tiles;
  pl.up_rows_per_expert = 8 * e->d; pl.down_rows_per_expert = e->d; pl.layer0 = layer0; pl.tile_m = e->tile_m; pl.slot0 = slot0; pl.n_layers = n_layers;
  pl.route_lt_sub = 0;
  pl.trim_rows = tok ? 0 : trim_rows;
  if (tok) {  // token-level tables have no slot dimension; the tile tables stay in the evaluation slot
    pl.sel_idx = e->tok_sel_idx; pl.pos = e->tok_pos; pl.B = B * e->T; pl.T = 1; pl.route_lt_sub = slot0 * e->L;
  }
  CU_OK(launch_k(plan_kernel, dim3(n_slots * n_layers), dim3(256), 0, st, pl));
  e->last_slot = slot0 + n_slots - 1;
  CU_OK(cudaGetLastError());
  e->launch_count += 2;
  return MODE_OK;
}

// One NoiseBlockMoE (modedit.py:530-595) given hA = bf16(ln_1(x)+c) and routing tables for layer l.
static int enqueue_block(mode_engine* e, cudaStream_t st, int B, int l, int combine_mode, int slot, const LayerIO& io,
                         int trim_rows = 0) {
  const int d = e->d, M = B * e->T;
  const size_t lt = (size_t)slot * e->L + l;  // layer index inside the routing tables
  // measurement aid (scripts/skip_diag.py): bit PC_x set = do not launch that kernel class; outputs are then garbage
  static const unsigned skip = getenv("MODE_DEBUG_SKIP") ? (unsigned)strtoul(getenv("MODE_DEBUG_SKIP"), nullptr, 0) : 0u;
  GemmParams p = gemm_params(io.tm_hA, e->tm_wqkv, io.to_qkv, e->dense_tiles, e->dense_counts, 3 * d, d, e->b_qkv);
  p.w_row_off = l * 3 * d;
  enable_stream_k(e, p);
  const int pairs = e->num_sms >> 1;
  const int n_m_dense = (M + e->tile_m - 1) / e->tile_m;
  if (!e->train_forward)
    RET_IF(narrow_gemm(e, p, choose_bn(n_m_dense, 3 * d, pairs, d / GEMM_BLOCK_K), 0, e->w_qkv, (uint64_t)e->L * 3 * d, d, io.qkv, 3 * d, M));
  // rollout-sized batch: every group has <= 16 rows -> weight-streaming kernels (gemm_small.cuh)
  const bool small = e->small_m && !io.z && !e->train_forward && M <= SMALL_M_MAX_ROWS && d % 256 == 0;
  const int small_groups = e->E < B * e->K ? e->E : B * e->K;  // upper bound on routed groups
  {
    ProfScope ps(e, st, PC_QKV);
    if (small)
      RET_IF(small_gemm(e, EPI_BIAS_BF16, st, small_params(io.hA, e->w_qkv, d, e->dense_tiles, e->dense_counts, e->b_qkv,
                                                               io.qkv, 3 * d, l * 3 * d), 3 * d, 1, M));
    else if (!(skip >> PC_QKV & 1)) RET_IF(launch_gemm(EPI_BIAS_BF16, e->pair, e->num_sms, st, p));
  }
  AttnParams a;
  a.tmap_qkv = io.tm_qkv_attn;
  a.qkv = io.qkv; a.out = io.attn; a.q_gain = e->qn_g + (size_t)l * e->Dh; a.k_gain = e->kn_g + (size_t)l * e->Dh;
  a.B = B; a.T = e->T; a.H = e->H; a.eps = e->cfg.rms_eps; a.inv_sqrt_dh = e->inv_sqrt_dh;
  a.drop = dropout_spec(e, e->stoch.p_attn, RNG_ATTN, l);
  {
    ProfScope ps(e, st, PC_ATTN);
    if (e->rec)
      rec_phase(e, SP_ATTN, a, (B * e->H + SMALL_M_WARPS - 1) / SMALL_M_WARPS);
    else if (!(skip >> PC_ATTN & 1)) RET_IF(launch_attn(st, a, e->Dh));
  }
  p = gemm_params(io.tm_attn, e->tm_wproj, io.to_x1, e->dense_tiles, e->dense_counts, d, d, nullptr);  // x1 += acc
  p.w_row_off = l * d;
  if (!e->train_forward)
    RET_IF(narrow_gemm(e, p, choose_bn(n_m_dense, d, pairs, d / GEMM_BLOCK_K), 1, e->w_proj, (uint64_t)e->L * d, d, io.x1, d, M));
  {
    ProfScope ps(e, st, PC_PROJ);
    if (small)
      RET_IF(small_gemm(e, EPI_RESID_F32, st, small_params(io.attn, e->w_proj, d, e->dense_tiles, e->dense_counts, nullptr,
                                                               io.x1, d, l * d), d, 1, M));
    else if (!(skip >> PC_PROJ & 1)) RET_IF(launch_gemm(EPI_RESID_F32, e->pair, e->num_sms, st, p));
  }
  Ln2Params n2;
  const RouteView rv = route_view(e, B, lt, l);
  const DropoutSpec mlp_drop = dropout_spec(e, e->stoch.p_mlp, RNG_MLP, l);
  n2.x = io.x1; n2.x_out = io.xn; n2.g = e->ln2_g + (size_t)l * d; n2.pos = rv.pos; n2.perm = io.perm;
  int* row_token = mlp_drop.thr ? e->row_token + (size_t)l * e->perm_rows : nullptr;  // [L][perm_rows], training only
  n2.row_token = row_token;
  const int t_skip = (trim_rows > 0 && !rv.per_token) ? e->T - trim_rows : 0;  // must match the plan of this layer
  n2.t_skip = t_skip;
  n2.B = rv.units; n2.T = rv.rt; n2.K = e->K; n2.d = d; n2.eps = e->cfg.rms_eps; n2.inv_sqrt_d = e->inv_sqrt_d;
  // the training forward keeps the two-launch path (it saves z)
  const bool fused_mlp = (e->mlp_fused || M <= e->mlp_fused_max_rows) && !io.z && !e->train_forward && !small;
  n2.zero = fused_mlp ? e->mlp_sync : nullptr;
  n2.n_zero = 1 + e->max_tiles;
  {
    ProfScope ps(e, st, PC_LN2);
    if (!(skip >> PC_LN2 & 1)) ROW_PHASE(SP_LN2, ln2_permute_kernel, d, row_blocks(M), st, n2);
  }
  CU_OK(cudaGetLastError());
  p = gemm_params(io.tm_perm, e->tm_wup, io.to_h, e->up_tiles + lt * e->max_tiles, e->num_tiles + lt, 8 * d, d,
                  e->b_up);
  GemmParams pd = gemm_params(io.tm_h, e->tm_wdown, io.to_y, e->down_tiles + lt * e->max_tiles, e->num_tiles + lt, d, e->F,
                              nullptr);
  if (fused_mlp) {
    // up-projection + SwiGLU and down-projection from one dynamic tile queue (mlp_fused.cuh); reported as PC_UP
    ProfScope ps(e, st, PC_UP);
    MlpParams mp;
    mp.up = p;
    mp.down = pd;
    mp.sync = e->mlp_sync;
    static const int mlp_flags = getenv("MODE_MLP_FLAGS") ? atoi(getenv("MODE_MLP_FLAGS")) : (1 | (18 << 8));
    mp.flags = mlp_flags;  // deferred signalling, down tiles queued 18 M-tiles behind
    CU_OK(launch_k(mlp_fused_2cta_kernel, dim3(e->num_sms & ~1), dim3(GEMM_THREADS), G2_SMEM_BYTES, st, mp));
  } else {
    enable_stream_k(e, p);
    {
      ProfScope ps(e, st, PC_UP);
      if (io.z) {  // training: also keep the pre-activations for the SwiGLU backward
        p.tmap_out2 = io.to_z;
        p.drop = mlp_drop; p.row_token = row_token; p.drop_rows_per_expert = 8 * d; p.drop_E = e->E; p.drop_half_F = e->F / 2;
        RET_IF(launch_gemm(EPI_SWIGLU_SAVE, e->pair, e->num_sms, st, p));
      } else if (small) {
        RET_IF(small_gemm(e, EPI_SWIGLU_BF16, st, small_params(io.perm, e->w_up, d, e->up_tiles + lt * e->max_tiles,
                                                                   e->num_tiles + lt, e->b_up, io.h, e->F, 0), e->F, small_groups, M,
                          /*tables_final=*/slot != ROUTE_SLOT_EVAL));  // a pre-routed schedule was planned before the launch chain
      } else if (!(skip >> PC_UP & 1)) {
        RET_IF(launch_gemm(EPI_SWIGLU_BF16, e->pair, e->num_sms, st, p));
      }
    }
    enable_stream_k(e, pd);
    if (!e->train_forward) {
      // routed rows of this block (the last block only keeps its action rows), before group padding
      // M-tiles of the grouped GEMM: every routed expert's group is padded to whole tiles; with one sigma for the batch
      // (samplers) all rows go to the same K experts, which is also the estimate used for heterogeneous batches
      const int group_rows = rv.units * (rv.rt - t_skip);
      RET_IF(narrow_gemm(e, pd, choose_bn(e->K * ((group_rows + e->tile_m - 1) / e->tile_m), d, pairs, e->F / GEMM_BLOCK_K), 2, e->w_down,
                         (uint64_t)e->L * e->E * d, e->F, io.y, d, e->perm_rows));
    }
    if (e->pair && e->band_down) pd.band = std::max(1, (e->num_sms >> 1) / pd.n_blocks);  // one wave = one band x all column blocks
    {
      ProfScope ps(e, st, PC_DOWN);
      if (small)
        RET_IF(small_gemm(e, EPI_PLAIN_BF16, st, small_params(io.h, e->w_down, e->F, e->down_tiles + lt * e->max_tiles,
                                                                  e->num_tiles + lt, nullptr, io.y, d, 0), d, small_groups, M,
                          /*tables_final=*/slot != ROUTE_SLOT_EVAL));
      else if (!(skip >> PC_DOWN & 1)) RET_IF(launch_gemm(EPI_PLAIN_BF16, e->pair, e->num_sms, st, pd));
    }
  }
  CombineParams c;
  c.x = io.xn; c.x_out = io.x_out; c.x_copy = io.x_out_copy; c.y = io.y;
  c.pos = rv.pos; c.w = rv.sel_w;
  c.g_next = (combine_mode == 0) ? e->ln1_g + (size_t)(l + 1) * d : e->lnf_g;
  c.cvec = e->cvec; c.hA = io.hA_next; c.xnorm = e->xnorm;
  c.t_skip = t_skip;
  c.B = rv.units; c.T = rv.rt; c.Tc = e->T; c.K = e->K; c.d = d; c.mode = combine_mode; c.eps = e->cfg.rms_eps; c.inv_sqrt_d = e->inv_sqrt_d;
  {
    ProfScope ps(e, st, PC_COMBINE);
    if (!(skip >> PC_COMBINE & 1)) ROW_PHASE(SP_COMBINE, combine_kernel, d, row_blocks(M), st, c);
  }
  CU_OK(cudaGetLastError());
  e->launch_count += fused_mlp ? 6 : 7;
  return MODE_OK;
}

// One network evaluation: router + plan + embed + L blocks + head. head_mode as HeadParams.mode.
// `prerouted_slot` >= 0: the routing tables of that slot were filled ahead of time (sampler schedule); otherwise the
// evaluation routes itself into ROUTE_SLOT_EVAL.
static int enqueue_eval(mode_engine* e, cudaStream_t st, int B, const float* sigma, int stride, const float* actions,
                        int apply_c_in, int head_mode, float* out, const float* coefs, const float* clean,
                        int prerouted_slot = -1, const LayerIO* layer_io = nullptr, const float* prog = nullptr,
                        const float* noise = nullptr) {
  // layer_io: per-layer buffer sets (training); nullptr = the shared in-place inference set
  const LayerIO& io0 = layer_io ? layer_io[0] : e->io;
  const int slot = prerouted_slot >= 0 ? prerouted_slot : ROUTE_SLOT_EVAL;
  // the last block's dead rows are only dropped on the inference path (training keeps every row: its saved activations
  // and tile tables are shared with the backward pass); pre-routed schedules were planned with the same setting
  const int trim = layer_io ? 0 : e->trim_rows;
  if (prerouted_slot < 0) RET_IF(enqueue_routing(e, st, B, sigma, stride, nullptr, 0, e->L, ROUTE_SLOT_EVAL, 1, 0, trim));
  EmbedParams em;
  em.sc = StepScalars{sigma, stride, e->cfg.sigma_data};
  em.sig_u = e->sig_u; em.sig_v = e->sig_v; em.goal_tok = e->goal_tok; em.state_tok = e->state_tok; em.pos = e->pos;
  em.w_act_t = e->w_act; em.actions = actions; em.ln1_g = e->ln1_g; em.x = io0.x_in; em.x_copy = layer_io ? io0.x1 : nullptr; em.cvec = e->cvec; em.hA = io0.hA;
  em.B = B; em.T = e->T; em.S = e->S; em.A = e->A; em.action_dim = e->adim; em.d = e->d; em.apply_c_in = apply_c_in;
  em.eps = e->cfg.rms_eps; em.inv_sqrt_d = e->inv_sqrt_d;
  em.drop = dropout_spec(e, e->stoch.p_embed, RNG_EMBED, 0);
  {
    ProfScope ps(e, st, PC_EMBED);
    ROW_PHASE(SP_EMBED, embed_kernel, e->d, row_blocks(B * e->T), st, em);
  }
  CU_OK(cudaGetLastError());
  for (int l = 0; l < e->L; ++l) {
    RET_IF(wait_block_weights(e, st, l));
    RET_IF(enqueue_block(e, st, B, l, l + 1 < e->L ? 0 : 1, slot, layer_io ? layer_io[l] : e->io, l + 1 < e->L ? 0 : trim));
  }
  if (head_mode < 0) {  // training forward: the loss/head backward kernel consumes the final-ln output directly
    e->launch_count += 1;
    return MODE_OK;
  }
  HeadParams h;
  h.sc = StepScalars{sigma, stride, e->cfg.sigma_data};
  h.xnorm = e->xnorm; h.w_out = e->w_out; h.b_out = e->b_out; h.x_act = actions; h.out = out; h.clean = clean;
  h.tok_sqerr = e->tok_sqerr; h.coefs = coefs; h.d_prev = e->d_prev;
  h.prog = prog; h.xbase = e->x_work; h.xprobe = e->x_probe; h.hist = e->hist; h.noise = noise;
  h.B = B; h.T = e->T; h.A = e->A; h.action_dim = e->adim; h.d = e->d; h.mode = head_mode;
  {
    ProfScope ps(e, st, PC_HEAD);
    ROW_PHASE(SP_HEAD, head_kernel, e->d, row_blocks(B * e->A), st, h);
  }
  CU_OK(cudaGetLastError());
  e->launch_count += 2;
  return MODE_OK;
}

extern "C" int mode_weights_record_ready(mode_engine_t* e, int layer, void* stream) {
  if (!e) return fail(MODE_ERR_INVALID, "null engine");
  if (layer < 0 || layer >= e->L) return fail(MODE_ERR_INVALID, "block %d out of range", layer);
  if (e->block_w_ready.empty()) {
    e->block_w_ready.assign(e->L, nullptr);
    e->block_w_pending.assign(e->L, 0);
  }
  if (!e->block_w_ready[layer]) CU_OK(cudaEventCreateWithFlags(&e->block_w_ready[layer], cudaEventDisableTiming));
  CU_OK(cudaEventRecord(e->block_w_ready[layer], reinterpret_cast<cudaStream_t>(stream)));
  e->block_w_pending[layer] = 1;
  return MODE_OK;
}

// ------------------------------------------------------------------------------------------------ public entry points
static int eval_common(mode_engine* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                       const float* sigma_dev, int sigma_stride, float* out_dev, int B, void* stream, int apply_c_in,
                       int head_mode) {
  if (!e || !state_dev || !goal_dev || !actions_dev || !sigma_dev || !out_dev) return fail(MODE_ERR_INVALID, "null argument");
  if (sigma_stride != 0 && sigma_stride != 1) return fail(MODE_ERR_INVALID, "sigma_stride must be 0 or 1");
  RET_IF(ensure_batch(e, B, reinterpret_cast<cudaStream_t>(stream)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  e->launch_count = 0;
  RET_IF(enqueue_cond(e, st, B, state_dev, goal_dev));
  return enqueue_eval(e, st, B, sigma_dev, sigma_stride, actions_dev, apply_c_in, head_mode, out_dev, nullptr, nullptr);
}

extern "C" int mode_forward(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                            const float* sigma_dev, int sigma_stride, float* out_dev, int B, void* stream) {
  return eval_common(e, state_dev, goal_dev, actions_dev, sigma_dev, sigma_stride, out_dev, B, stream, 0, 0);
}

extern "C" int mode_denoise(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                            const float* sigma_dev, int sigma_stride, float* out_dev, int B, void* stream) {
  return eval_common(e, state_dev, goal_dev, actions_dev, sigma_dev, sigma_stride, out_dev, B, stream, 1, 1);
}

extern "C" int mode_profile_eval(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                                 const float* sigma_dev, int sigma_stride, float* out_dev, int B, int reps, void* stream,
                                 float* ms_host, int32_t* launches_host) {
  if (!e || !ms_host || !launches_host || reps < 1) return fail(MODE_ERR_INVALID, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < PC_COUNT; ++i) {
    ms_host[i] = 0.f;
    launches_host[i] = 0;
  }
  // one untimed pass (also validates arguments), then `reps` instrumented passes
  RET_IF(mode_denoise(e, state_dev, goal_dev, actions_dev, sigma_dev, sigma_stride, out_dev, B, stream));
  for (int r = 0; r < reps; ++r) {
    e->prof_on = true;
    int rc = mode_denoise(e, state_dev, goal_dev, actions_dev, sigma_dev, sigma_stride, out_dev, B, stream);
    e->prof_on = false;
    cudaError_t ce = cudaStreamSynchronize(st);
    for (size_t i = 0; i < e->prof_cls.size(); ++i) {
      float ms = 0.f;
      if (rc == MODE_OK && ce == cudaSuccess) cudaEventElapsedTime(&ms, e->prof_ev[2 * i], e->prof_ev[2 * i + 1]);
      ms_host[e->prof_cls[i]] += ms / reps;
      if (r == 0) launches_host[e->prof_cls[i]] += 1;
      cudaEventDestroy(e->prof_ev[2 * i]);
      cudaEventDestroy(e->prof_ev[2 * i + 1]);
    }
    e->prof_ev.clear();
    e->prof_cls.clear();
    if (rc != MODE_OK) return rc;
    if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "profile pass failed: %s", cudaGetErrorString(ce));
  }
  return MODE_OK;
}

extern "C" int mode_loss(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* action_dev,
                         const float* noise_dev, const float* sigma_dev, float* loss_dev, float* out_dev, int B,
                         void* stream) {
  if (!e || !state_dev || !goal_dev || !action_dev || !noise_dev || !sigma_dev || !loss_dev)
    return fail(MODE_ERR_INVALID, "null argument");
  RET_IF(ensure_batch(e, B, reinterpret_cast<cudaStream_t>(stream)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  e->launch_count = 0;
  const int per = e->A * e->adim, n = B * per;
  noise_actions_kernel<<<(n + 255) / 256, 256, 0, st>>>(action_dev, noise_dev, sigma_dev, e->x_work, per, n);
  CU_OK(cudaGetLastError());
  RET_IF(enqueue_cond(e, st, B, state_dev, goal_dev));
  RET_IF(enqueue_eval(e, st, B, sigma_dev, 1, e->x_work, 1, 3, out_dev, nullptr, action_dev));
  loss_reduce_kernel<<<1, 256, 0, st>>>(e->tok_sqerr, B * e->A, (float)n, loss_dev);
  CU_OK(cudaGetLastError());
  e->launch_count += 2;
  return MODE_OK;
}


// ------------------------------------------------------------------------------------------------ persistent small-batch kernel
static bool small_fused_ok(const mode_engine* e, int B) {
  const int M = B * e->T, nvec = e->d / 128;
  return e->small_fused && e->small_m && M <= SMALL_M_MAX_ROWS && e->T <= 16 && e->d % 256 == 0 &&
         ((nvec == 2 && e->Dh == 64) || (nvec == 4 && e->Dh == 128) || (nvec == 8 && e->Dh == 128));
}
template <int NVEC, int DH, int MT>
static int launch_small_eval_t(mode_engine* e, cudaStream_t st, const mode_engine::SmallProgram& prog) {
  static int grid = 0;
  constexpr int smem = small_eval_smem_bytes<DH, MT>();
  if (!grid) {
    CU_OK(cudaFuncSetAttribute(small_eval_kernel<NVEC, DH, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CU_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, small_eval_kernel<NVEC, DH, MT>, SMALL_EVAL_THREADS, smem));
    if (occ < 1) return fail(MODE_ERR_CUDA, "small-batch kernel does not fit on an SM");
    grid = e->num_sms * occ;  // all CTAs resident at once: required by the grid barrier (cooperative launch checks it)
  }
  CU_OK(cudaMemsetAsync(e->small_barrier, 0, sizeof(unsigned), st));
  const SmallPhase* dev = prog.dev;
  int n = prog.n;
  unsigned* bar = e->small_barrier;
  void* args[] = {(void*)&dev, (void*)&n, (void*)&bar};
  CU_OK(cudaLaunchCooperativeKernel((const void*)small_eval_kernel<NVEC, DH, MT>, dim3(grid), dim3(SMALL_EVAL_THREADS), args, smem, st));
  return MODE_OK;
}
static int launch_small_eval(mode_engine* e, cudaStream_t st, const mode_engine::SmallProgram& prog, int B) {
  const int nvec = e->d / 128;
  const bool two = B * e->T > 16;
  if (nvec == 2) return two ? launch_small_eval_t<2, 64, 2>(e, st, prog) : launch_small_eval_t<2, 64, 1>(e, st, prog);
  if (nvec == 4) return two ? launch_small_eval_t<4, 128, 2>(e, st, prog) : launch_small_eval_t<4, 128, 1>(e, st, prog);
  return two ? launch_small_eval_t<8, 128, 2>(e, st, prog) : launch_small_eval_t<8, 128, 1>(e, st, prog);
}
// Records (once per key) the phase table of `n_evals` network evaluations; eval_fn(i) enqueues evaluation i.
template <typename F>
static int get_small_program(mode_engine* e, const std::string& key, int n_evals, F eval_fn, mode_engine::SmallProgram* out) {
  auto it = e->small_programs.find(key);
  if (it != e->small_programs.end()) {
    *out = it->second;
    return MODE_OK;
  }
  if (!e->small_barrier) RET_IF(dev_alloc(e, &e->small_barrier, 1));
  std::vector<SmallPhase> v;
  e->rec = &v;
  int rc = MODE_OK;
  for (int i = 0; i < n_evals && rc == MODE_OK; ++i) rc = eval_fn(i);
  e->rec = nullptr;
  RET_IF(rc);
  mode_engine::SmallProgram pr;
  pr.n = (int)v.size();
  void* q = nullptr;
  CU_OK(cudaMalloc(&q, v.size() * sizeof(SmallPhase)));
  pr.dev = reinterpret_cast<SmallPhase*>(q);
  CU_OK(cudaMemcpy(pr.dev, v.data(), v.size() * sizeof(SmallPhase), cudaMemcpyHostToDevice));
  e->small_programs[key] = pr;
  *out = pr;
  return MODE_OK;
}

// head_mode: 2 DDIM, 4 Euler, 5 DPM-Solver++(2M) (HeadParams::mode)
static int get_sampler_graph(mode_engine* e, int B, int n, int head_mode, cudaGraphExec_t* exec) {
  const auto key = std::make_pair(B, n + 128 * head_mode);
  auto it = e->graphs.find(key);
  if (it != e->graphs.end()) {
    *exec = it->second;
    e->launch_count += e->graph_launches[key];
    return MODE_OK;
  }
  const int64_t before = e->launch_count;
  cudaGraph_t graph = nullptr;
  CU_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  int rc = MODE_OK;
  for (int i = 0; i < n && rc == MODE_OK; ++i)
    rc = enqueue_eval(e, e->cap_stream, B, e->sig_dev + i, 0, e->x_work, 1, head_mode, e->x_work, e->coefs_dev + SCHED_COEFS * i, nullptr, i);
  cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
  if (rc != MODE_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
  cudaGraphExec_t ex = nullptr;
  ce = cudaGraphInstantiate(&ex, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
  e->graphs[key] = ex;
  e->graph_launches[key] = e->launch_count - before;
  *exec = ex;
  return MODE_OK;
}

// Per-step update coefficients in fp32, following the op order of the reference samplers (0-dim fp32 tensor arithmetic):
//   DDIM / DPM-Solver-1 (gc_sampling.py:936-950)  {sigma_fn(t')/sigma_fn(t), expm1(-h)}
//   Euler, s_churn = 0 (:164-211)                 {sigma' - sigma}
//   DPM-Solver++(2M) (:699-734)                   {ratio, expm1(-h), 1 + 1/(2r), 1/(2r)}; the last two are {1, 0} on the
//                                                 first step and on a step that ends at sigma = 0 (first-order update)
static void sampler_schedule(int sampler, const float* sigmas, int n, ScheduleArg* a) {
  a->n = n;
  for (int i = 0; i < n; ++i) {
    float* v = a->v + (1 + SCHED_COEFS) * i;
    const float s = sigmas[i], sn = sigmas[i + 1];
    const float t = -logf(s), tn = -logf(sn);  // t_fn = sigma.log().neg(); log(0) = -inf -> tn = +inf
    const float h = tn - t;
    const float ratio = expf(-tn) / expf(-t);  // sigma_fn(t_next) / sigma_fn(t)
    const float em1 = expm1f(-h);              // (-h).expm1()
    v[0] = s;
    v[1] = v[2] = v[3] = v[4] = 0.f;
    if (sampler == MODE_SAMPLER_EULER) {
      v[1] = sn - s;
    } else {
      v[1] = ratio;
      v[2] = em1;
      v[3] = 1.f;
      if (sampler == MODE_SAMPLER_DPMPP_2M && i > 0 && sn != 0.f) {
        const float h_last = t - (-logf(sigmas[i - 1]));
        const float r = h_last / h;
        v[3] = 1.f + 1.f / (2.f * r);
        v[4] = 1.f / (2.f * r);
      }
    }
  }
}

extern "C" int mode_sample(mode_engine_t* e, int sampler, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                           const float* sigmas_host, int n_plus_1, int B, void* stream) {
  if (!e || !state_dev || !goal_dev || !x_inout_dev || !sigmas_host) return fail(MODE_ERR_INVALID, "null argument");
  if (sampler != MODE_SAMPLER_DDIM && sampler != MODE_SAMPLER_EULER && sampler != MODE_SAMPLER_DPMPP_2M)
    return fail(MODE_ERR_INVALID, "unknown fused sampler %d", sampler);
  const int n = n_plus_1 - 1;
  if (n < 1 || n > 64) return fail(MODE_ERR_INVALID, "number of sampling steps must be in [1, 64]");
  for (int i = 0; i < n; ++i)
    if (!(sigmas_host[i] > 0.f)) return fail(MODE_ERR_INVALID, "sigmas[%d] must be > 0", i);
  RET_IF(ensure_batch(e, B, reinterpret_cast<cudaStream_t>(stream)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  e->launch_count = 0;
  cudaGraphExec_t exec = nullptr;
  const int head_mode = sampler == MODE_SAMPLER_DDIM ? 2 : (sampler == MODE_SAMPLER_EULER ? 4 : 5);
  const bool fused_small = small_fused_ok(e, B);  // rollout-sized batch: the whole loop is one persistent kernel
  mode_engine::SmallProgram sp{nullptr, 0};
  if (fused_small) {
    const std::string key = "s:" + std::to_string(B) + ":" + std::to_string(n) + ":" + std::to_string(head_mode);
    RET_IF(get_small_program(e, key, n, [&](int i) {
      return enqueue_eval(e, nullptr, B, e->sig_dev + i, 0, e->x_work, 1, head_mode, e->x_work, e->coefs_dev + SCHED_COEFS * i, nullptr, i);
    }, &sp));
    e->launch_count = 0;
  } else {
    RET_IF(get_sampler_graph(e, B, n, head_mode, &exec));
  }
  ScheduleArg sa;
  sampler_schedule(sampler, sigmas_host, n, &sa);
  set_schedule_kernel<<<1, 64, 0, st>>>(sa, e->sig_dev, e->coefs_dev);
  CU_OK(cudaGetLastError());
  const size_t xbytes = (size_t)B * e->A * e->adim * sizeof(float);
  CU_OK(cudaMemcpyAsync(e->x_work, x_inout_dev, xbytes, cudaMemcpyDeviceToDevice, st));
  // route the whole sigma schedule (n steps x L layers) in one router + one plan launch, ahead of the captured loop
  RET_IF(enqueue_routing(e, st, B, e->sig_dev, 0, nullptr, 0, e->L, 0, n, 1, e->trim_rows));
  RET_IF(enqueue_cond(e, st, B, state_dev, goal_dev));
  if (fused_small)
    RET_IF(launch_small_eval(e, st, sp, B));
  else
    CU_OK(cudaGraphLaunch(exec, st));
  CU_OK(cudaMemcpyAsync(x_inout_dev, e->x_work, xbytes, cudaMemcpyDeviceToDevice, st));
  e->launch_count += 1;
  return MODE_OK;
}

extern "C" int mode_sample_ddim(mode_engine_t* e, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                                const float* sigmas_host, int n_plus_1, int B, void* stream) {
  return mode_sample(e, MODE_SAMPLER_DDIM, state_dev, goal_dev, x_inout_dev, sigmas_host, n_plus_1, B, stream);
}

extern "C" int mode_sample_ddim_host(mode_engine_t* e, const float* state_host, const float* goal_host,
                                     float* x_inout_host, const float* sigmas_host, int n_plus_1, int B, void* stream) {
  if (!e || !state_host || !goal_host || !x_inout_host) return fail(MODE_ERR_INVALID, "null argument");
  RET_IF(ensure_batch(e, B, reinterpret_cast<cudaStream_t>(stream)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n_state = (size_t)B * e->S * e->obs, n_goal = (size_t)B * e->gdim, n_x = (size_t)B * e->A * e->adim;
  CU_OK(cudaMemcpyAsync(e->in_state, state_host, n_state * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_OK(cudaMemcpyAsync(e->in_goal, goal_host, n_goal * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_OK(cudaMemcpyAsync(e->in_x, x_inout_host, n_x * sizeof(float), cudaMemcpyHostToDevice, st));
  RET_IF(mode_sample_ddim(e, e->in_state, e->in_goal, e->in_x, sigmas_host, n_plus_1, B, stream));
  CU_OK(cudaMemcpyAsync(x_inout_host, e->in_x, n_x * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_OK(cudaStreamSynchronize(st));
  return MODE_OK;
}

// Sampler programs: any k-diffusion sampler whose update is linear in {X, P, D, history, noise} as one CUDA graph
// (gc_sampling.py: sample_heun :257, sample_dpm_2 :315, sample_lms :430, sample_dpmpp_2s :956, the ancestral variants
// :214 / :376 / :874). The host (mode_diffusion_policy_b200/gc_sampling.py) turns the sigma schedule into one row of
// coefficients per network evaluation; the graph only depends on the evaluation count and on which evaluations read
// the probe, so every schedule of the same sampler replays the same graph.
extern "C" int mode_sample_program(mode_engine_t* e, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                                   const float* sigma_eval_host, const int32_t* reads_probe_host, const float* prog_host,
                                   const float* noise_dev, int n_evals, int B, void* stream) {
  if (!e || !state_dev || !goal_dev || !x_inout_dev || !sigma_eval_host || !reads_probe_host || !prog_host)
    return fail(MODE_ERR_INVALID, "null argument");
  if (n_evals < 1 || n_evals > 64) return fail(MODE_ERR_INVALID, "a sampler program has 1..64 network evaluations (got %d)", n_evals);
  for (int i = 0; i < n_evals; ++i)
    if (!(sigma_eval_host[i] > 0.f)) return fail(MODE_ERR_INVALID, "evaluation %d: sigma must be > 0", i);
  RET_IF(ensure_batch(e, B, reinterpret_cast<cudaStream_t>(stream)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n_el = (size_t)e->maxB * e->A * e->adim;
  if (!e->x_probe) {
    RET_IF(dev_alloc(e, &e->x_probe, n_el));
    RET_IF(dev_alloc(e, &e->hist, 4 * n_el));
    RET_IF(dev_alloc(e, &e->prog_dev, (size_t)64 * HEAD_PROG_FLOATS));
    RET_IF(dev_alloc(e, &e->noise_buf, 64 * n_el));
  }
  e->launch_count = 0;
  std::string key = std::to_string(B) + (noise_dev ? ":n:" : ":-:");
  for (int i = 0; i < n_evals; ++i) key.push_back(reads_probe_host[i] ? 'P' : 'X');
  const size_t cur_el = (size_t)B * e->A * e->adim;
  const bool fused_small = small_fused_ok(e, B);
  mode_engine::SmallProgram sp{nullptr, 0};
  if (fused_small)
    RET_IF(get_small_program(e, "p:" + key, n_evals, [&](int i) {
      return enqueue_eval(e, nullptr, B, e->sig_dev + i, 0, reads_probe_host[i] ? e->x_probe : e->x_work, 1, 6, nullptr, nullptr,
                          nullptr, i, nullptr, e->prog_dev + (size_t)i * HEAD_PROG_FLOATS,
                          noise_dev ? e->noise_buf + (size_t)i * cur_el : nullptr);
    }, &sp));
  auto it = e->prog_graphs.find(key);
  if (fused_small) {
    e->launch_count = 0;
  } else if (it == e->prog_graphs.end()) {
    const int64_t before = e->launch_count;
    cudaGraph_t graph = nullptr;
    CU_OK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = MODE_OK;
    for (int i = 0; i < n_evals && rc == MODE_OK; ++i)
      rc = enqueue_eval(e, e->cap_stream, B, e->sig_dev + i, 0, reads_probe_host[i] ? e->x_probe : e->x_work, 1, 6, nullptr,
                        nullptr, nullptr, i, nullptr, e->prog_dev + (size_t)i * HEAD_PROG_FLOATS,
                        noise_dev ? e->noise_buf + (size_t)i * cur_el : nullptr);
    cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc != MODE_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
    cudaGraphExec_t ex = nullptr;
    ce = cudaGraphInstantiate(&ex, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    it = e->prog_graphs.emplace(key, ex).first;
    e->prog_graph_launches[key] = e->launch_count - before;
  } else {
    e->launch_count += e->prog_graph_launches[key];
  }
  ScheduleArg sa;
  sa.n = n_evals;
  for (int i = 0; i < n_evals; ++i) {
    float* v = sa.v + (1 + SCHED_COEFS) * i;
    v[0] = sigma_eval_host[i];
    v[1] = v[2] = v[3] = v[4] = 0.f;
  }
  set_schedule_kernel<<<1, 64, 0, st>>>(sa, e->sig_dev, e->coefs_dev);
  CU_OK(cudaGetLastError());
  CU_OK(cudaMemcpyAsync(e->prog_dev, prog_host, (size_t)n_evals * HEAD_PROG_FLOATS * sizeof(float), cudaMemcpyHostToDevice, st));
  if (noise_dev)
    CU_OK(cudaMemcpyAsync(e->noise_buf, noise_dev, (size_t)n_evals * cur_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CU_OK(cudaMemsetAsync(e->hist, 0, 4 * cur_el * sizeof(float), st));
  CU_OK(cudaMemcpyAsync(e->x_work, x_inout_dev, cur_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CU_OK(cudaMemcpyAsync(e->x_probe, x_inout_dev, cur_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
  RET_IF(enqueue_routing(e, st, B, e->sig_dev, 0, nullptr, 0, e->L, 0, n_evals, 1, e->trim_rows));
  RET_IF(enqueue_cond(e, st, B, state_dev, goal_dev));
  if (fused_small)
    RET_IF(launch_small_eval(e, st, sp, B));
  else
    CU_OK(cudaGraphLaunch(it->second, st));
  CU_OK(cudaMemcpyAsync(x_inout_dev, e->x_work, cur_el * sizeof(float), cudaMemcpyDeviceToDevice, st));
  e->launch_count += 1;
  return MODE_OK;
}

extern "C" int mode_block_forward(mode_engine_t* e, int layer, const float* x_dev, const float* c_dev, float* out_dev,
                                  int B, void* stream) {
  if (!e || !x_dev || !c_dev || !out_dev) return fail(MODE_ERR_INVALID, "null argument");
  if (layer < 0 || layer >= e->L) return fail(MODE_ERR_INVALID, "layer %d out of range", layer);
  RET_IF(ensure_batch(e, B, reinterpret_cast<cudaStream_t>(stream)));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  e->launch_count = 0;
  const int d = e->d, M = B * e->T;
  CU_OK(cudaMemcpyAsync(e->x, x_dev, (size_t)M * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CU_OK(cudaMemcpyAsync(e->cvec, c_dev, (size_t)B * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // generic router input: z = W1 c + b1 for an arbitrary conditioning vector
  router_hidden_kernel<<<(unsigned)((B * e->Hd + 7) / 8), 256, 0, st>>>(e->r_w1 + (size_t)layer * e->Hd * d,
                                                                        e->r_b1 + (size_t)layer * e->Hd, e->cvec, e->zbuf,
                                                                        B, e->Hd, d);
  CU_OK(cudaGetLastError());
  RET_IF(enqueue_routing(e, st, B, e->sig_dev, 0, e->zbuf, layer, 1));
  Ln1Params l1;
  l1.x = e->x; l1.cvec = e->cvec; l1.g = e->ln1_g + (size_t)layer * d; l1.hA = e->hA; l1.rows = M; l1.T = e->T; l1.d = d;
  l1.eps = e->cfg.rms_eps; l1.inv_sqrt_d = e->inv_sqrt_d;
  LAUNCH_ROW_KERNEL(ln1_kernel, d, row_blocks(M), st, l1);
  CU_OK(cudaGetLastError());
  RET_IF(enqueue_block(e, st, B, layer, 2, ROUTE_SLOT_EVAL, e->io));
  CU_OK(cudaMemcpyAsync(out_dev, e->x, (size_t)M * d * sizeof(float), cudaMemcpyDeviceToDevice, st));
  e->launch_count += 2;
  return MODE_OK;
}

extern "C" int mode_get_routing_at(mode_engine_t* e, int step, int layer, int B, int32_t* idx_host, float* w_host,
                                   float* probs_host) {
  if (!e) return fail(MODE_ERR_INVALID, "null engine");
  if (layer < 0 || layer >= e->L || B < 1 || B > e->maxB) return fail(MODE_ERR_INVALID, "layer/B out of range");
  if (step < -1 || step >= ROUTE_SLOT_EVAL) return fail(MODE_ERR_INVALID, "step %d out of range", step);
  CU_OK(cudaDeviceSynchronize());
  const size_t lt = (size_t)(step < 0 ? e->last_slot : step) * e->L + layer;
  const size_t o = lt * B * e->K;
  if (idx_host) CU_OK(cudaMemcpy(idx_host, e->topk_idx + o, (size_t)B * e->K * sizeof(int), cudaMemcpyDeviceToHost));
  if (w_host) CU_OK(cudaMemcpy(w_host, e->topk_w + o, (size_t)B * e->K * sizeof(float), cudaMemcpyDeviceToHost));
  if (probs_host)
    CU_OK(cudaMemcpy(probs_host, e->probs + lt * B * e->E, (size_t)B * e->E * sizeof(float), cudaMemcpyDeviceToHost));
  return MODE_OK;
}

extern "C" int mode_get_routing(mode_engine_t* e, int layer, int B, int32_t* idx_host, float* w_host, float* probs_host) {
  return mode_get_routing_at(e, -1, layer, B, idx_host, w_host, probs_host);
}

extern "C" int mode_get_expert_usage(mode_engine_t* e, int layer, int64_t* usage_host, int64_t* total_tokens_host) {
  if (!e || layer < 0 || layer >= e->L) return fail(MODE_ERR_INVALID, "bad engine/layer");
  CU_OK(cudaDeviceSynchronize());
  if (usage_host) CU_OK(cudaMemcpy(usage_host, e->usage + (size_t)layer * e->E, e->E * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (total_tokens_host) CU_OK(cudaMemcpy(total_tokens_host, e->tokens + layer, sizeof(int64_t), cudaMemcpyDeviceToHost));
  return MODE_OK;
}

extern "C" int mode_reset_expert_usage(mode_engine_t* e) {
  if (!e) return fail(MODE_ERR_INVALID, "null engine");
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemset(e->usage, 0, (size_t)e->L * e->E * sizeof(unsigned long long)));
  CU_OK(cudaMemset(e->tokens, 0, (size_t)e->L * sizeof(unsigned long long)));
  return MODE_OK;
}

extern "C" int64_t mode_last_launch_count(const mode_engine_t* e) { return e ? e->launch_count : 0; }

// ------------------------------------------------------------------------------------------------ unit-test entries
extern "C" int mode_debug_gemm(const void* a_dev, const void* w_dev, const float* bias_dev, const float* resid_dev,
                               void* out_dev, int M, int N, int Kdim, int epilogue, void* stream) {
  if (!a_dev || !w_dev || !out_dev) return fail(MODE_ERR_INVALID, "null argument");
  const bool pair = (epilogue & 0x100) != 0;      // bit 8 selects the CTA-pair kernel
  const bool stream_k = (epilogue & 0x200) != 0;  // bit 9 enables stream-K on the last partial wave (pair kernel)
  const bool skip_b_debug = (epilogue & 0x400) != 0;  // bit 10: measurement aid, see gemm.cuh (garbage results)
  const int bn = ((epilogue >> 16) & 0xff) ? ((epilogue >> 16) & 0xff) * 16 : GEMM_BLOCK_N;  // bits 16-23: tile width / 16
  epilogue &= 0xff;
  if (bn != GEMM_BLOCK_N && (!pair || stream_k || bn < 64 || bn > 256 || epilogue == EPI_SWIGLU_BF16 || epilogue == EPI_SWIGLU_SAVE))
    return fail(MODE_ERR_INVALID, "tile width %d needs the plain CTA-pair kernel and a non-SwiGLU epilogue", bn);
  const int tm = pair ? 256 : 128;
  if (M < 1 || N % 256 || Kdim % 64 || N < 256 || Kdim < 64) return fail(MODE_ERR_INVALID, "need N %% 256 == 0 and K %% 64 == 0");
  RET_IF(set_kernel_attrs());
  int dev = 0;
  CU_OK(cudaGetDevice(&dev));
  int sms = 0;
  CU_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n = (M + tm - 1) / tm;
  std::vector<GemmMTile> tiles(n);
  for (int i = 0; i < n; ++i) tiles[i] = GemmMTile{i * tm, i * tm, M - i * tm < tm ? M - i * tm : tm, 0};
  GemmMTile* d_tiles = nullptr;
  int* d_n = nullptr;
  RET_IF(dev_alloc<GemmMTile>(nullptr, &d_tiles, n, false));
  RET_IF(dev_alloc<int>(nullptr, &d_n, 1, false));
  CU_OK(cudaMemcpy(d_tiles, tiles.data(), n * sizeof(GemmMTile), cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(d_n, &n, sizeof(int), cudaMemcpyHostToDevice));
  CUtensorMap ta, tw;
  // The caller allocates A with round_up(M, 256) rows so the TMA box never exceeds the tensor extent.
  RET_IF(make_tmap(&ta, a_dev, round_up(M, 256), Kdim, 128));
  RET_IF(make_tmap(&tw, w_dev, N, Kdim, pair ? bn / 2 : 256));
  const int n_out = (epilogue == EPI_SWIGLU_BF16) ? N / 2 : N;
  const int out_bytes = (epilogue == EPI_RESID_F32 || epilogue == EPI_PLAIN_F32) ? 4 : 2;
  CUtensorMap tout;
  RET_IF(make_out_tmap(&tout, out_dev, M, n_out, out_bytes));
  if (epilogue == EPI_RESID_F32) {  // the kernel accumulates into the output: seed it with the residual
    if (!resid_dev) return fail(MODE_ERR_INVALID, "resid epilogue needs resid_dev");
    CU_OK(cudaMemcpyAsync(out_dev, resid_dev, (size_t)M * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  GemmParams p = gemm_params(ta, tw, tout, d_tiles, d_n, N, Kdim, bias_dev);
  if (bn != GEMM_BLOCK_N) {
    p.bn = bn;
    p.n_blocks = (N + bn - 1) / bn;
    p.out_ptr = out_dev;
    p.ldo = N;
    p.out_rows = M;
  }
  static float4* dbg_parts = nullptr;
  static int* dbg_flags = nullptr;
  if (pair && stream_k) {
    if (!dbg_parts) {
      RET_IF(dev_alloc<float4>(nullptr, &dbg_parts, (size_t)sms * 8 * 8 * 128, false));
      RET_IF(dev_alloc<int>(nullptr, &dbg_flags, (size_t)sms * 4, true));
    }
    p.sk_enable = 1;
    p.sk_partials = dbg_parts;
    p.sk_flags = dbg_flags;
  }
  if (pair && skip_b_debug) p.sk_enable |= 2;
  int rc = launch_gemm(epilogue, pair, sms, st, p);
  // MODE_GEMM_BENCH_REPS=n: time n further back-to-back launches with CUDA events and print the average
  if (rc == MODE_OK && getenv("MODE_GEMM_BENCH_REPS")) {
    const int reps = atoi(getenv("MODE_GEMM_BENCH_REPS"));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps && rc == MODE_OK; ++i) rc = launch_gemm(epilogue, pair, sms, st, p);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * M * (double)N * Kdim;
    printf("mode_debug_gemm%s%s M=%d N=%d K=%d epi=%d bn=%d: %.3f us/launch, %.1f TFLOP/s\n", pair ? "[pair]" : "", (pair && stream_k) ? "[sk]" : "", M, N, Kdim, epilogue,
           bn, 1e3 * ms / reps, flops * reps / (ms * 1e-3) / 1e12);
    fflush(stdout);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(d_tiles);
  cudaFree(d_n);
  if (rc != MODE_OK) return rc;
  if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "GEMM kernel failed: %s", cudaGetErrorString(ce));
  return MODE_OK;
}

// Activation operand of the weight-gradient GEMM: row-major bf16 [rows, cols], box {64 columns, 64 rows}.
static int make_mn_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols) {
  return make_tmap_ex(m, base, rows, cols, 64, 2);
}

static int launch_wgrad(int num_sms, cudaStream_t st, const WgradParams& p) {
  CU_OK(launch_k(gemm_wgrad_kernel, dim3(num_sms), dim3(GEMM_THREADS), GEMM_SMEM_BYTES, st, p));
  return MODE_OK;
}

extern "C" int mode_debug_wgrad(const void* dy_dev, const void* x_dev, float* out_dev, int rows, int n_out, int k_out,
                                int swiglu_half, void* stream) {
  if (!dy_dev || !x_dev || !out_dev) return fail(MODE_ERR_INVALID, "null argument");
  if (rows < 64 || rows % 64 || n_out % 128 || k_out % 256) return fail(MODE_ERR_INVALID, "need rows %% 64 == 0, N_out %% 128 == 0, K_out %% 256 == 0");
  RET_IF(set_kernel_attrs());
  int dev = 0, sms = 0;
  CU_OK(cudaGetDevice(&dev));
  CU_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  WgradParams p;
  RET_IF(make_mn_tmap(&p.tmap_dy, dy_dev, rows, n_out));
  RET_IF(make_mn_tmap(&p.tmap_x, x_dev, rows, k_out));
  RET_IF(make_out_tmap(&p.tmap_out, out_dev, n_out, k_out, 4));
  WgradProblem pr{0, rows / 64, 0, 0};
  WgradProblem* d_pr = nullptr;
  RET_IF(dev_alloc<WgradProblem>(nullptr, &d_pr, 1, false));
  CU_OK(cudaMemcpy(d_pr, &pr, sizeof(pr), cudaMemcpyHostToDevice));
  p.problems = d_pr;
  p.n_problems = 1;
  p.m_tiles = n_out / 128;
  p.n_blocks = k_out / 256;
  p.swiglu_half = swiglu_half;
  int rc = launch_wgrad(sms, st, p);
  if (rc == MODE_OK && getenv("MODE_GEMM_BENCH_REPS")) {
    const int reps = atoi(getenv("MODE_GEMM_BENCH_REPS"));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps && rc == MODE_OK; ++i) rc = launch_wgrad(sms, st, p);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("mode_debug_wgrad rows=%d N_out=%d K_out=%d: %.3f us/launch, %.1f TFLOP/s\n", rows, n_out, k_out,
           1e3 * ms / reps, 2.0 * rows * (double)n_out * k_out * reps / (ms * 1e-3) / 1e12);
    fflush(stdout);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  cudaError_t ce = cudaStreamSynchronize(st);
  cudaFree(d_pr);
  if (rc != MODE_OK) return rc;
  if (ce != cudaSuccess) return fail(MODE_ERR_CUDA, "wgrad kernel failed: %s", cudaGetErrorString(ce));
  return MODE_OK;
}

extern "C" int mode_debug_attention(const void* qkv_dev, const float* q_gain_dev, const float* k_gain_dev, void* out_dev,
                                    int B, int T, int H, int Dh, float eps, void* stream) {
  if (!qkv_dev || !q_gain_dev || !k_gain_dev || !out_dev) return fail(MODE_ERR_INVALID, "null argument");
  AttnParams a;
  a.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv_dev);
  a.out = reinterpret_cast<__nv_bfloat16*>(out_dev);
  a.q_gain = q_gain_dev; a.k_gain = k_gain_dev; a.B = B; a.T = T; a.H = H; a.eps = eps;
  a.inv_sqrt_dh = static_cast<float>(pow(static_cast<double>(Dh), -0.5));
  if (Dh >= 64 && T <= 64) RET_IF(make_tmap(&a.tmap_qkv, qkv_dev, (uint64_t)B * T, (uint64_t)3 * H * Dh, 16 * ((T + 15) / 16)));
  return launch_attn(reinterpret_cast<cudaStream_t>(stream), a, Dh);
}

#include "train.inc"
#include "resnet.inc"
