// HBM-bound row kernels of the MoDE denoising step: token embedding, RMSNorm (+sigma-conditioning add, + bf16 cast
// for the next GEMM's A operand), router softmax/top-k, routing plan (counting sort by expert), token permute,
// expert combine, output head + EDM preconditioning + DDIM update.
// One warp per token row (d floats), 16-byte vector accesses, fp32 math, warp-shuffle reductions.
#pragma once
#include "gemm_wgrad.cuh"
#include "ptx.cuh"
#include "rng.cuh"

namespace mode {

constexpr int ROW_WARPS = 8;        // warps per CTA for row kernels
constexpr int MAX_D = 2048;         // embed_dim upper bound (register-resident row: MAX_D/128 float4 per lane)
constexpr int MAX_VEC = MAX_D / 128;
constexpr int MAX_EXPERTS = 32;     // one lane per expert in the router
constexpr int MAX_TOPK = 8;

// Launch-invariant vectors (norm gains, embedding tables, head weights) are read through the non-coherent path: besides
// the cache hint this tells the compiler that no store of the kernel can alias them, so their loads are hoisted above
// the stores of earlier loop iterations instead of waiting for them (activations keep plain loads).
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// RMSNorm.forward, modedit.py:72-80: x / clamp(||x||_2 * dim^-0.5, min=eps) * g. Returns 1/denominator: one IEEE
// division per row, then a multiply per element (differs from the per-element division by <= 1 fp32 ulp).
__device__ __forceinline__ float rms_inv_denominator(float sumsq, float inv_sqrt_dim, float eps) {
  return 1.0f / fmaxf(sqrtf(sumsq) * inv_sqrt_dim, eps);
}

// ------------------------------------------------------------------------------------------------------------
// Cast fp32 -> bf16 (A operand of the obs/goal embedding GEMMs).
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n4) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n4) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ------------------------------------------------------------------------------------------------------------
// Per-sample scalars of one network evaluation. sigma_stride = 0 broadcasts one sigma to the whole batch
// (samplers: sigmas[i] * s_in, gc_sampling.py:945); 1 = per-sample sigma (training loss, mode_agent.py:669).
struct StepScalars {
  const float* sigma;
  int sigma_stride;
  float sigma_data;
};
__device__ __forceinline__ float load_sigma(const StepScalars& s, int b) { return s.sigma[b * s.sigma_stride]; }

// ------------------------------------------------------------------------------------------------------------
// Token embedding + layer-0 ln_1 (+c): builds the input sequence (MoDeDiT.forward, modedit.py:754-790, :847-860)
//   row 0          : emb_t = sigma_linear(sigma_emb(ln(sigma)/4))  = s*u + v   (both layers are affine in s)
//   row 1          : goal_emb(goal) + pos[0]
//   rows 2..2+S-1  : tok_emb(state_images) + pos[1]       (every image token shares pos row 1)
//   rows 2+S..T-1  : action_emb(action * c_in) + pos[1 + j]
// and writes x (fp32 residual stream) and hA = bf16(rms(x)*g_ln1 + c) for the first QKV GEMM.
struct EmbedParams {
  StepScalars sc;
  const float* sig_u;      // [d]  sigma_linear.weight @ sigma_emb.weight[:,0]
  const float* sig_v;      // [d]  sigma_linear.weight @ sigma_emb.bias
  const float* goal_tok;   // [B, d]   fp32 goal_emb(goal)       (no pos yet; computed once per trajectory)
  const float* state_tok;  // [B*S, d] fp32 tok_emb(state_images)
  const float* pos;        // [1+A, d]
  const float* w_act_t;    // [action_dim, d]: action_emb.weight transposed at load time (coalesced float4 loads)
  const float* actions;    // [B, A, action_dim]
  const float* ln1_g;      // [d] layer-0 ln_1 gain
  float* x;                // [B*T, d]
  float* x_copy;           // optional second copy (training: seeds the c_proj accumulation buffer x1)
  float* cvec;             // [B, d] conditioning vector c = emb_t (consumed by later kernels)
  __nv_bfloat16* hA;       // [B*T, d]
  int B, T, S, A, action_dim, d;
  int apply_c_in;          // 1: GCDenoiser.forward scales the action input by c_in (score_wrappers.py:79-80)
  float eps;
  float inv_sqrt_d;  // float(d ** -0.5), computed on the host exactly as RMSNorm.__init__ does (modedit.py:75)
  // training only: nn.Dropout(embed_pdrob) on the goal / image / action token embeddings (+pos), not on the sigma token
  // (modedit.py:779-784). Element (row, col) uses half (col & 1) of word (row*d + col)/2 of stream RNG_EMBED.
  DropoutSpec drop = DropoutSpec{0u, 0u, 1.0f};
};

// keep/scale factors of 4 consecutive elements starting at flat element index e0 (a multiple of 4)
__device__ __forceinline__ float4 dropout_mask4(const DropoutSpec& d, uint32_t e0) {
  const uint32_t b0 = rng_bits(d.key, e0 >> 1), b1 = rng_bits(d.key, (e0 >> 1) + 1u);
  return make_float4((b0 & 0xffffu) < d.thr ? 0.f : d.scale, (b0 >> 16) < d.thr ? 0.f : d.scale,
                     (b1 & 0xffffu) < d.thr ? 0.f : d.scale, (b1 >> 16) < d.thr ? 0.f : d.scale);
}

// body of embed_kernel for virtual block `vblock` (also a phase of the persistent small-batch kernel, small_eval.cuh)
template <int NVEC>
__device__ __forceinline__ void embed_body(const EmbedParams& p, int vblock) {
  const int row = vblock * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.T) return;
  const int b = row / p.T, t = row % p.T;
  const float sigma = load_sigma(p.sc, b);
  const float s = logf(sigma) / 4.0f;  // process_sigma_embeddings, modedit.py:824
  float4 xv[NVEC];
  float ss = 0.f;
  float act[8];
  const int n_act = p.action_dim;
  if (t >= 2 + p.S) {
    float c_in = 1.0f;
    if (p.apply_c_in) c_in = 1.0f / sqrtf(sigma * sigma + p.sc.sigma_data * p.sc.sigma_data);
    const float* a = p.actions + (static_cast<size_t>(b) * p.A + (t - 2 - p.S)) * n_act;
#pragma unroll
    for (int j = 0; j < 8; ++j) act[j] = j < n_act ? a[j] * c_in : 0.f;  // fixed trip count: registers, loads issued together
  }
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    {
      const int col = (i * 32 + lane) * 4;
      const float4 u = ldg4(p.sig_u + col);
      const float4 v = ldg4(p.sig_v + col);
      const float4 c = make_float4(fmaf(s, u.x, v.x), fmaf(s, u.y, v.y), fmaf(s, u.z, v.z), fmaf(s, u.w, v.w));
      float4 x;
      if (t == 0) {
        x = c;
      } else if (t < 2 + p.S) {
        const float* src = (t == 1) ? p.goal_tok + static_cast<size_t>(b) * p.d
                                    : p.state_tok + (static_cast<size_t>(b) * p.S + (t - 2)) * p.d;
        const float4 e = *reinterpret_cast<const float4*>(src + col);
        const float4 pe = ldg4(p.pos + (t == 1 ? 0 : 1) * p.d + col);
        x = make_float4(e.x + pe.x, e.y + pe.y, e.z + pe.z, e.w + pe.w);
      } else {
        const int j = t - 2 - p.S;
        const float4 pe = ldg4(p.pos + (1 + j) * p.d + col);
        // the action_dim (<= 8) weight rows are requested back to back (a loop with a run-time trip count issued one
        // load per iteration and waited for it: 56 serialised round trips per row, 29 us at B = 1), summed in the same order
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        float4 wk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          wk[k] = k < n_act ? ldg4(p.w_act_t + static_cast<size_t>(k) * p.d + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < n_act) {
            e[0] = fmaf(act[k], wk[k].x, e[0]);
            e[1] = fmaf(act[k], wk[k].y, e[1]);
            e[2] = fmaf(act[k], wk[k].z, e[2]);
            e[3] = fmaf(act[k], wk[k].w, e[3]);
          }
        x = make_float4(e[0] + pe.x, e[1] + pe.y, e[2] + pe.z, e[3] + pe.w);
      }
      if (p.drop.thr && t > 0) {
        const float4 m = dropout_mask4(p.drop, static_cast<uint32_t>(row) * p.d + col);
        x = make_float4(x.x * m.x, x.y * m.y, x.z * m.z, x.w * m.w);
      }
      xv[i] = x;
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  }
  ss = warp_sum(ss);
  const float rn = rms_inv_denominator(ss, p.inv_sqrt_d, p.eps);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    {
      const int col = (i * 32 + lane) * 4;
      const float4 x = xv[i];
      const float4 u = ldg4(p.sig_u + col);
      const float4 v = ldg4(p.sig_v + col);
      const float4 c = make_float4(fmaf(s, u.x, v.x), fmaf(s, u.y, v.y), fmaf(s, u.z, v.z), fmaf(s, u.w, v.w));
      const float4 g = ldg4(p.ln1_g + col);
      *reinterpret_cast<float4*>(p.x + static_cast<size_t>(row) * p.d + col) = x;
      if (p.x_copy) *reinterpret_cast<float4*>(p.x_copy + static_cast<size_t>(row) * p.d + col) = x;
      if (t == 0) *reinterpret_cast<float4*>(p.cvec + static_cast<size_t>(b) * p.d + col) = c;
      const float h0 = (x.x * rn) * g.x + c.x, h1 = (x.y * rn) * g.y + c.y;
      const float h2 = (x.z * rn) * g.z + c.z, h3 = (x.w * rn) * g.w + c.w;
      *reinterpret_cast<uint2*>(p.hA + static_cast<size_t>(row) * p.d + col) =
          make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
    }
  }
}
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32, 4) embed_kernel(const EmbedParams p) {
  pdl_trigger();
  pdl_wait();
  embed_body<NVEC>(p, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------------------
// Block-level entry (NoiseBlockMoE.forward, modedit.py:530-532): hA = bf16(rms(x)*g + c) for caller-supplied x, c.
struct Ln1Params {
  const float* x;      // [rows, d]
  const float* cvec;   // [B, d]
  const float* g;      // [d]
  __nv_bfloat16* hA;   // [rows, d]
  int rows, T, d;
  float eps;
  float inv_sqrt_d;  // float(d ** -0.5), computed on the host exactly as RMSNorm.__init__ does (modedit.py:75)
};
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32, 4) ln1_kernel(const Ln1Params p) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.rows) return;
  const int b = row / p.T;
  float4 xv[NVEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * p.d + (i * 32 + lane) * 4);
      xv[i] = x;
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  ss = warp_sum(ss);
  const float rn = rms_inv_denominator(ss, p.inv_sqrt_d, p.eps);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 x = xv[i];
      const float4 g = *reinterpret_cast<const float4*>(p.g + col);
      const float4 c = *reinterpret_cast<const float4*>(p.cvec + static_cast<size_t>(b) * p.d + col);
      *reinterpret_cast<uint2*>(p.hA + static_cast<size_t>(row) * p.d + col) =
          make_uint2(pack_bf16x2((x.x * rn) * g.x + c.x, (x.y * rn) * g.y + c.y),
                     pack_bf16x2((x.z * rn) * g.z + c.z, (x.w * rn) * g.w + c.w));
    }
}

// ------------------------------------------------------------------------------------------------------------
// Router: noise-conditioned top-k (RouterCond.forward, modedit.py:312-421; CondRouterMLP :170-217).
// The router only sees c = emb_t = s*u + v (router_context_cond_only, modedit.py:328-332), so its first Linear is
// affine in the scalar s = ln(sigma)/4:  W1 c + b1 = s*(W1 u) + (W1 v + b1) =: s*ra + rb  (ra, rb precomputed in fp64
// at weight-load time). One warp per (layer, sample): lanes stride the 2d hidden units, GELU(erf), dot with W2 rows,
// warp-shuffle reduce, softmax, clamp(1e-9, 1-1e-9), k rounds of arg-max (lowest index wins ties), renormalise.
struct RouterParams {
  StepScalars sc;
  const float* ra;   // [L, Hd]
  const float* rb;   // [L, Hd]
  const float* w2;   // [L, E, Hd]
  const float* b2;   // [L, E]
  const float* z_explicit;  // optional [B, Hd]: pre-activation of the first Linear for an arbitrary c (block entry)
  int* topk_idx;     // [L, B, K]  descending probability (torch.topk order)
  float* topk_w;     // [L, B, K]  renormalised probabilities, same order
  int* sel_idx;      // [L, B, K]  the same experts sorted ascending (the reference's accumulation order, :561-566)
  float* sel_w;      // [L, B, K]
  float* probs;      // [L, B, E]  clamped softmax (true_probs)
  float* logits;     // [L, B, E]  logits - max (RouterCond.logits, :345-346)
  int L, B, E, K, Hd;
  int layer0;        // first layer handled (block-level entry routes a single layer)
  int normalize;
  // Tables have a leading `slot` dimension [slot][L][B][...] so that a whole sigma schedule can be routed by ONE launch
  // (slot = sampler step) ahead of the captured denoising loop; plain evaluations use a single slot.
  int slot0, Ltot;   // first slot written; layers per slot in the table layout
  int sigma_slot_stride;  // sigma of slot s starts at sc.sigma + s * sigma_slot_stride
  // Training with use_argmax=False (modedit.py:389-390): every TOKEN draws its K experts with torch.multinomial(probs, K,
  // replacement=False) semantics, i.e. K sequential draws proportional to the remaining clamped probabilities. Draw k
  // of token m inverts the CDF over the remaining experts (ascending index) at u * S, u = rng_uniform(bits(m*K + k)) of
  // stream RNG_ROUTE, S = fp32 sum of the remaining probabilities in ascending order. Tables [L][B*T][K], layer-major.
  int multinomial = 0, T = 1;
  unsigned long long seed = 0;
  uint32_t step = 0;
  int* tok_topk_idx = nullptr;   // draw order (what torch.multinomial returns)
  float* tok_topk_w = nullptr;
  int* tok_sel_idx = nullptr;    // ascending expert order
  float* tok_sel_w = nullptr;
};

// Two work layouts, same arithmetic per row:
//   WARP_ROWS = false  one CTA (8 warps) per (slot, layer, row): each warp covers Hd/8 hidden units, partial logits are
//                      reduced through shared memory, warp 0 finishes. Lowest latency; used when there are few rows —
//                      sampler schedules (sigma_stride == 0: ONE distinct row per layer, broadcast to all B samples).
//   WARP_ROWS = true   one WARP per row, the 8 warps of a CTA take 8 consecutive rows of the same layer, so the layer's
//                      router weights (ra, rb, E rows of W2: 8 KB each at d=1024) are fetched once per CTA through L1
//                      instead of once per row. Used for per-sample sigma (training, loss): 2.3x faster at B=256.
// The two layouts add the same products in different orders: logits agree to fp32 rounding, not bit for bit.
template <bool WARP_ROWS>
__global__ void __launch_bounds__(ROW_WARPS * 32) router_kernel(const RouterParams p) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part_all[ROW_WARPS][2][MAX_EXPERTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float (*part)[MAX_EXPERTS] = part_all[WARP_ROWS ? warp : 0];  // [0]: clamped probabilities, [1]: shifted logits
  const int R = p.sc.sigma_stride == 0 && !p.z_explicit ? 1 : p.B;  // distinct rows
  const int groups = WARP_ROWS ? (R + ROW_WARPS - 1) / ROW_WARPS : R;
  const int per_slot = p.L * groups;
  const int slot_i = blockIdx.x / per_slot, rem = blockIdx.x % per_slot;
  const int l = p.layer0 + rem / groups, b = WARP_ROWS ? (rem % groups) * ROW_WARPS + warp : rem % groups;
  if (b >= R) return;
  const int lt = (p.slot0 + slot_i) * p.Ltot + l;  // layer index inside the routing tables
  const float s = p.z_explicit ? 0.f : logf(p.sc.sigma[slot_i * p.sigma_slot_stride + b * p.sc.sigma_stride]) / 4.0f;
  const float* ra = p.ra + static_cast<size_t>(l) * p.Hd;
  const float* rb = p.rb + static_cast<size_t>(l) * p.Hd;
  const float* w2 = p.w2 + static_cast<size_t>(l) * p.E * p.Hd;
  float acc[MAX_EXPERTS];
#pragma unroll
  for (int e = 0; e < MAX_EXPERTS; ++e) acc[e] = 0.f;
  // each thread owns 4 consecutive hidden units per pass (Hd = 2d is a multiple of 512): 16-byte loads, all of a
  // pass's loads (ra, rb, E rows of W2) are independent and issued together
  const int j0 = (WARP_ROWS ? lane : static_cast<int>(threadIdx.x)) * 4, jstep = (WARP_ROWS ? 32 : ROW_WARPS * 32) * 4;
  for (int j = j0; j < p.Hd; j += jstep) {
    float4 z4;
    if (p.z_explicit) {
      z4 = *reinterpret_cast<const float4*>(p.z_explicit + static_cast<size_t>(b) * p.Hd + j);
    } else {
      const float4 a4 = *reinterpret_cast<const float4*>(ra + j), b4 = *reinterpret_cast<const float4*>(rb + j);
      z4 = make_float4(fmaf(s, a4.x, b4.x), fmaf(s, a4.y, b4.y), fmaf(s, a4.z, b4.z), fmaf(s, a4.w, b4.w));
    }
    constexpr float kInvSqrt2 = 0.70710678118654752440f;
    float4 h4;  // nn.GELU() (erf form)
    h4.x = 0.5f * z4.x * (1.0f + erff(z4.x * kInvSqrt2));
    h4.y = 0.5f * z4.y * (1.0f + erff(z4.y * kInvSqrt2));
    h4.z = 0.5f * z4.z * (1.0f + erff(z4.z * kInvSqrt2));
    h4.w = 0.5f * z4.w * (1.0f + erff(z4.w * kInvSqrt2));
#pragma unroll
    for (int e = 0; e < MAX_EXPERTS; ++e)
      if (e < p.E) {
        const float4 w = *reinterpret_cast<const float4*>(w2 + static_cast<size_t>(e) * p.Hd + j);
        acc[e] = fmaf(h4.x, w.x, fmaf(h4.y, w.y, fmaf(h4.z, w.z, fmaf(h4.w, w.w, acc[e]))));
      }
  }
  // lane e ends up owning expert e's logit
  float logit = -INFINITY;
  if constexpr (WARP_ROWS) {
#pragma unroll
    for (int e = 0; e < MAX_EXPERTS; ++e) {
      if (e < p.E) {
        const float v = warp_sum(acc[e]);
        if (lane == e) logit = v + p.b2[l * p.E + e];
      }
    }
  } else {
    __shared__ float wpart[ROW_WARPS][MAX_EXPERTS];
#pragma unroll
    for (int e = 0; e < MAX_EXPERTS; ++e) {
      if (e < p.E) {
        const float v = warp_sum(acc[e]);
        if (lane == 0) wpart[warp][e] = v;
      }
    }
    __syncthreads();
    if (warp != 0) return;
    if (lane < p.E) {  // fixed summation order over the 8 warps
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) v += wpart[w][lane];
      logit = v + p.b2[l * p.E + lane];
    }
  }
  float mx = logit;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float zl = logit - mx;  // (logits - max) / temperature(=1)
  const float ex = lane < p.E ? expf(zl) : 0.f;
  const float den = warp_sum(ex);
  float prob = ex / den;
  prob = fminf(fmaxf(prob, 1e-9f), 1.0f - 1e-9f);
  // top-k: k rounds of warp arg-max, ties -> lowest expert index
  float cand = lane < p.E ? prob : -1.f;
  int sel[MAX_TOPK], srt[MAX_TOPK];
  float selp[MAX_TOPK], srtp[MAX_TOPK];
  float psum = 0.f;
#pragma unroll
  for (int k = 0; k < MAX_TOPK; ++k) {
    sel[k] = 0;
    selp[k] = 0.f;
    if (k < p.K) {
      float bv = cand;
      int bi = lane;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      sel[k] = bi;
      selp[k] = bv;
      psum += bv;
      if (lane == bi) cand = -1.f;
    }
  }
  // renormalise (modedit.py:418-419) and build the ascending-expert order by rank counting (K <= 8, registers only)
#pragma unroll
  for (int k = 0; k < MAX_TOPK; ++k)
    if (k < p.K && p.normalize) selp[k] = selp[k] / psum;
#pragma unroll
  for (int k = 0; k < MAX_TOPK; ++k) {
    srt[k] = 0;
    srtp[k] = 0.f;
  }
#pragma unroll
  for (int k = 0; k < MAX_TOPK; ++k) {
    if (k < p.K) {
      int rank = 0;
#pragma unroll
      for (int j = 0; j < MAX_TOPK; ++j)
        if (j < p.K && sel[j] < sel[k]) ++rank;
#pragma unroll
      for (int j = 0; j < MAX_TOPK; ++j)
        if (j == rank) {
          srt[j] = sel[k];
          srtp[j] = selp[k];
        }
    }
  }
  // write this row's result to its own slot, or to every sample's slot when one sigma serves the whole batch
  const int n_dst = (R == 1) ? p.B : 1;
  if (lane < p.E) {  // stage this row's probabilities / logits in the warp's shared slot
    part[0][lane] = prob;
    part[1][lane] = zl;
  }
  __syncwarp();
  for (int i = lane; i < n_dst * p.E; i += 32) {
    const int bb = (R == 1) ? i / p.E : b, e = i % p.E;
    const size_t o = (static_cast<size_t>(lt) * p.B + bb) * p.E + e;
    p.probs[o] = part[0][e];
    p.logits[o] = part[1][e];
  }
  for (int bb = (R == 1 ? lane : (lane == 0 ? b : p.B)); bb < (R == 1 ? p.B : b + 1); bb += 32) {
    const size_t pk = (static_cast<size_t>(lt) * p.B + bb) * p.K;
#pragma unroll
    for (int k = 0; k < MAX_TOPK; ++k)
      if (k < p.K) {
        p.topk_idx[pk + k] = sel[k];
        p.topk_w[pk + k] = selp[k];
        p.sel_idx[pk + k] = srt[k];
        p.sel_w[pk + k] = srtp[k];
      }
  }
  if (!p.multinomial) return;
  // ---- per-token multinomial draws (lane t handles token t of this sample); part[0][e] holds the clamped probabilities
  const uint32_t key = rng_key(p.seed, p.step, RNG_ROUTE, static_cast<uint32_t>(l));
  for (int t = lane; t < p.T; t += 32) {
    const uint32_t m = static_cast<uint32_t>(b) * p.T + t;
    uint32_t avail = p.E >= 32 ? 0xffffffffu : ((1u << p.E) - 1u);
    int dr[MAX_TOPK];
    float dw[MAX_TOPK];
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < MAX_TOPK; ++k) {
      dr[k] = 0;
      dw[k] = 0.f;
      if (k < p.K) {
        float S = 0.f;
        for (int e = 0; e < p.E; ++e)
          if (avail >> e & 1u) S = __fadd_rn(S, part[0][e]);
        const float target = __fmul_rn(rng_uniform(rng_bits(key, m * p.K + k)), S);
        float cum = 0.f;
        int chosen = -1, last = 0;
        for (int e = 0; e < p.E; ++e)
          if (avail >> e & 1u) {
            last = e;
            cum = __fadd_rn(cum, part[0][e]);
            if (chosen < 0 && target < cum) chosen = e;
          }
        if (chosen < 0) chosen = last;
        avail &= ~(1u << chosen);
        dr[k] = chosen;
        dw[k] = part[0][chosen];
        tot = __fadd_rn(tot, dw[k]);
      }
    }
    const size_t o = (static_cast<size_t>(l) * p.B * p.T + m) * p.K;
#pragma unroll
    for (int k = 0; k < MAX_TOPK; ++k) {
      if (k < p.K) {
        const float w = p.normalize ? dw[k] / tot : dw[k];
        int rank = 0;
#pragma unroll
        for (int j = 0; j < MAX_TOPK; ++j)
          if (j < p.K && dr[j] < dr[k]) ++rank;
        p.tok_topk_idx[o + k] = dr[k];
        p.tok_topk_w[o + k] = w;
        p.tok_sel_idx[o + rank] = dr[k];
        p.tok_sel_w[o + rank] = w;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Routing plan: per layer, a stable counting sort of samples by expert. Produces
//   pos[l, b, k]  first row of sample b's T tokens inside expert sel_idx[l,b,k]'s group (groups padded to 128 rows)
//   up/down M-tile tables for the grouped GEMMs and their tile count
//   expert-usage counters (NoiseBlockMoE.inference_expert_usage / total_tokens_processed, modedit.py:568-572, :594)
// One CTA per layer; rank of a sample inside its group = exclusive block scan of "selects expert e" flags.
struct PlanParams {
  const int* sel_idx;            // [L, B, K]
  int* pos;                      // [L, B, K]
  GemmMTile* up_tiles;           // [L, max_tiles]
  GemmMTile* down_tiles;         // [L, max_tiles]
  GemmMTile* downT_tiles;        // [L, max_tiles] transposed down weights (4d rows per expert): data gradient of `down`
  WgradProblem* wg_up;           // [L, E] weight-gradient problems of the up projection (training; may be null)
  WgradProblem* wg_down;         // [L, E]
  int* num_tiles;                // [L]
  unsigned long long* usage;     // [L, E]
  unsigned long long* tokens;    // [L]
  int L, B, K, E, T, max_tiles;
  int up_rows_per_expert;        // 8d  (packed SwiGLU rows)
  int down_rows_per_expert;      // d
  int layer0;                    // first layer handled by blockIdx.x == 0 (block-level entry plans a single layer)
  int tile_m;                    // rows per GEMM M-tile (128 single-CTA, 256 CTA-pair): groups are padded to it
  int slot0, n_layers;           // grid = n_slots * n_layers CTAs: slot = slot0 + blockIdx.x / n_layers
  int trim_rows = 0;             // > 0: the LAST layer only routes the final trim_rows token rows of every unit (see below)
  int route_lt_sub = 0;              // subtracted from the table-layer index for sel_idx / pos (token-level tables: no slots)
};

__global__ void __launch_bounds__(256) plan_kernel(const PlanParams p) {
  pdl_trigger();
  pdl_wait();
  const int l = p.layer0 + blockIdx.x % p.n_layers;
  const int lt = (p.slot0 + blockIdx.x / p.n_layers) * p.L + l;  // layer index inside the routing tables
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int warp_tot[8];
  __shared__ int cnt[MAX_EXPERTS];
  __shared__ int grp_row0[MAX_EXPERTS];
  __shared__ int grp_tile0[MAX_EXPERTS + 1];
  // Dead-row elimination: only the last action_seq_len tokens of the last block reach the output head
  // (modedit.py:806-808) and nothing downstream attends to the others, so the last block's experts run on those rows
  // only. The usage counters keep counting every routed token like the reference does.
  const int Tl = (p.trim_rows > 0 && l == p.L - 1) ? p.trim_rows : p.T;
  const int* sel = p.sel_idx + static_cast<size_t>(lt - p.route_lt_sub) * p.B * p.K;
  int* pos = p.pos + static_cast<size_t>(lt - p.route_lt_sub) * p.B * p.K;
  // pass 1: rank of every (sample, slot) inside its expert group; stored temporarily in pos
  for (int e = 0; e < p.E; ++e) {
    int running = 0;
    for (int base = 0; base < p.B; base += 256) {
      const int b = base + tid;
      int slot = -1;
      if (b < p.B)
        for (int k = 0; k < p.K; ++k)
          if (sel[b * p.K + k] == e) slot = k;
      const unsigned bal = __ballot_sync(0xffffffffu, slot >= 0);
      const int pre = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) warp_tot[warp] = __popc(bal);
      __syncthreads();
      int woff = 0, tot = 0;
      for (int w = 0; w < 8; ++w) {
        if (w < warp) woff += warp_tot[w];
        tot += warp_tot[w];
      }
      if (slot >= 0) pos[b * p.K + slot] = running + woff + pre;
      running += tot;
      __syncthreads();
    }
    if (tid == 0) cnt[e] = running;
  }
  __syncthreads();
  if (tid == 0) {
    int row = 0, tile = 0;
    for (int e = 0; e < p.E; ++e) {
      grp_row0[e] = row;
      grp_tile0[e] = tile;
      const int nt = (cnt[e] * Tl + p.tile_m - 1) / p.tile_m;
      row += nt * p.tile_m;
      tile += nt;
    }
    grp_tile0[p.E] = tile;
    p.num_tiles[lt] = tile;
    atomicAdd(p.tokens + l, static_cast<unsigned long long>(p.B) * p.T);
  }
  __syncthreads();
  if (tid < p.E) atomicAdd(p.usage + static_cast<size_t>(l) * p.E + tid, static_cast<unsigned long long>(cnt[tid]) * p.T);
  // pass 2: ranks -> row positions
  for (int i = tid; i < p.B * p.K; i += 256) pos[i] = grp_row0[sel[i]] + pos[i] * Tl;
  if (p.wg_up && tid < p.E) {
    const int nt = grp_tile0[tid + 1] - grp_tile0[tid];
    p.wg_up[static_cast<size_t>(lt) * p.E + tid] = WgradProblem{grp_row0[tid], nt * p.tile_m / 64, (l * p.E + tid) * p.up_rows_per_expert, 0};
    p.wg_down[static_cast<size_t>(lt) * p.E + tid] = WgradProblem{grp_row0[tid], nt * p.tile_m / 64, (l * p.E + tid) * p.down_rows_per_expert, 0};
  }
  // pass 3: tile tables
  for (int e = 0; e < p.E; ++e) {
    const int nt = grp_tile0[e + 1] - grp_tile0[e];
    const int rows = cnt[e] * Tl;
    for (int i = tid; i < nt; i += 256) {
      GemmMTile t;
      t.a_row0 = grp_row0[e] + i * p.tile_m;
      t.out_row0 = t.a_row0;
      t.rows_valid = min(p.tile_m, rows - i * p.tile_m);
      t.w_row_base = (l * p.E + e) * p.up_rows_per_expert;
      p.up_tiles[static_cast<size_t>(lt) * p.max_tiles + grp_tile0[e] + i] = t;
      t.w_row_base = (l * p.E + e) * p.down_rows_per_expert;
      p.down_tiles[static_cast<size_t>(lt) * p.max_tiles + grp_tile0[e] + i] = t;
      t.w_row_base = (l * p.E + e) * 4 * p.down_rows_per_expert;
      p.downT_tiles[static_cast<size_t>(lt) * p.max_tiles + grp_tile0[e] + i] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// ln_2 + permute: x <- rms(x)*g (the block REPLACES the residual stream with its norm, modedit.py:539), and the
// bf16 copy of each token row is written straight to its K expert groups (the gather of :563-566 done as a scatter
// with 16-byte coalesced stores; no separate permute pass).
struct Ln2Params {
  const float* x;         // [B*T, d] in (x1)
  float* x_out;           // [B*T, d] out (xn); may alias x (inference: in place)
  const float* g;         // [d]
  const int* pos;         // [B, K] (this layer)
  __nv_bfloat16* perm;    // [rows_perm, d]
  int B, T, K, d;
  float eps;
  float inv_sqrt_d;  // float(d ** -0.5), computed on the host exactly as RMSNorm.__init__ does (modedit.py:75)
  int* zero;         // tile queue + dependency counters of the fused expert-MLP kernel that follows (or null)
  int n_zero;
  int t_skip = 0;    // rows t < t_skip of every unit are dead (last block, see plan_kernel): neither normalised nor routed
  int* row_token = nullptr;  // optional [rows_perm]: token row of every permuted row (MLP dropout of the training path)
};
// (B, T) of the routed kernels are ROUTING units x rows per unit: (samples, tokens per sample) when all tokens of a
// sample share their experts (eval / arg-max routing), (B*T tokens, 1) under per-token multinomial routing.
// body of ln2_permute_kernel for virtual block `vblock` (also a phase of the persistent small-batch kernel, small_eval.cuh)
// KFIX = 2: the top-2 routing of every shipped configuration, fully unrolled — the generic variant (KFIX = 0) predicates its
// loops over MAX_TOPK = 8 slots, i.e. issues four times the instructions of the two live ones (ncu: 1 765 instructions
// per row in combine_kernel, IPC-bound at 19 % of DRAM bandwidth).
template <int NVEC, int KFIX>
__device__ __forceinline__ void ln2_permute_impl(const Ln2Params& p, int vblock) {
  constexpr int KMAX = KFIX ? KFIX : MAX_TOPK;
  if (vblock == 0 && p.zero)
    for (int i = threadIdx.x; i < p.n_zero; i += ROW_WARPS * 32) p.zero[i] = 0;
  const int row = vblock * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.T) return;
  const int b = row / p.T, t = row % p.T - p.t_skip;
  if (t < 0) return;
  float4 xv[NVEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * p.d + (i * 32 + lane) * 4);
      xv[i] = x;
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
  ss = warp_sum(ss);
  const float rn = rms_inv_denominator(ss, p.inv_sqrt_d, p.eps);
  int dst_row[MAX_TOPK];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (KFIX || k < p.K) dst_row[k] = p.pos[b * p.K + k] + t;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 x = xv[i];
      const float4 g = ldg4(p.g + col);
      const float4 y = make_float4((x.x * rn) * g.x, (x.y * rn) * g.y, (x.z * rn) * g.z,
                                   (x.w * rn) * g.w);
      *reinterpret_cast<float4*>(p.x_out + static_cast<size_t>(row) * p.d + col) = y;
      const uint2 pk = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (KFIX || k < p.K) *reinterpret_cast<uint2*>(p.perm + static_cast<size_t>(dst_row[k]) * p.d + col) = pk;
    }
  if (p.row_token && lane == 0) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (KFIX || k < p.K) p.row_token[dst_row[k]] = row;
  }
}
template <int NVEC>
__device__ __forceinline__ void ln2_permute_body(const Ln2Params& p, int vblock) {
  if (p.K == 2)
    ln2_permute_impl<NVEC, 2>(p, vblock);
  else
    ln2_permute_impl<NVEC, 0>(p, vblock);
}
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32, 4) ln2_permute_kernel(const Ln2Params p) {
  pdl_trigger();
  pdl_wait();
  ln2_permute_body<NVEC>(p, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------------------
// Expert combine (+ next layer's ln_1 + c): x <- xn + sum_k w_k * y[pos_k]  accumulated in ascending expert order
// with separately rounded products, exactly as `next_states[idx] += probs * expert(x)` does (modedit.py:561-566, :595);
// then hA = bf16(rms(x)*g_next + c) feeds the next block's QKV GEMM. For the last block g_next is the final `ln`
// gain (modedit.py:818) and the normalised row is written back as fp32 for the head instead.
struct CombineParams {
  const float* x;              // [B*T, d] in: xn
  float* x_out;                // [B*T, d] block output; may alias x (inference: in place)
  float* x_copy;               // optional second copy of the block output (training: next layer's x1 seed)
  const __nv_bfloat16* y;      // [rows_perm, d] expert outputs (bf16)
  const int* pos;              // [B, K]
  const float* w;              // [B, K] (ascending expert order)
  const float* g_next;         // [d]
  const float* cvec;           // [B, d]
  __nv_bfloat16* hA;           // [B*T, d] (mode 0)
  float* xnorm;                // [B*T, d] (mode 1: final ln output, fp32)
  int B, T, K, d;
  int mode;                    // 0: next block's ln_1 + c -> hA ; 1: final ln -> xnorm ; 2: none (block-level entry)
  int t_skip = 0;              // rows t < t_skip of every unit are dead (last block): skipped
  int Tc;                      // token rows per cvec row (tokens per sample; differs from T under per-token routing)
  float eps;
  float inv_sqrt_d;  // float(d ** -0.5), computed on the host exactly as RMSNorm.__init__ does (modedit.py:75)
};
// body of combine_kernel for virtual block `vblock` (also a phase of the persistent small-batch kernel, small_eval.cuh)
// KFIX = 2: the top-2 routing of every shipped configuration, fully unrolled — the generic variant (KFIX = 0) predicates its
// loops over MAX_TOPK = 8 slots, i.e. issues four times the instructions of the two live ones (ncu: 1 765 instructions
// per row in combine_kernel, IPC-bound at 19 % of DRAM bandwidth).
template <int NVEC, int KFIX>
__device__ __forceinline__ void combine_impl(const CombineParams& p, int vblock) {
  constexpr int KMAX = KFIX ? KFIX : MAX_TOPK;
  const int row = vblock * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.T) return;
  const int b = row / p.T, t = row % p.T - p.t_skip;
  if (t < 0) return;
  int src_row[MAX_TOPK];
  float wk[MAX_TOPK];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
    if (KFIX || k < p.K) {
      src_row[k] = p.pos[b * p.K + k] + t;
      wk[k] = p.w[b * p.K + k];
    }
  // Three passes over the row's NVEC vectors, so that no store sits between loads: x_out may alias x (inference updates
  // the residual stream in place) and a store inside the load loop made every later load wait for it — eight serialised
  // round trips per row (ncu: 20.9 us at 17 % of DRAM bandwidth for 50 MB). (1) all x loads, (2) all expert-row loads and
  // the weighted sum in the reference's order, (3) all stores.
  float4 xv[NVEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
    xv[i] = *reinterpret_cast<const float4*>(p.x + static_cast<size_t>(row) * p.d + (i * 32 + lane) * 4);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (KFIX || k < p.K) {
          const uint2 raw = *reinterpret_cast<const uint2*>(p.y + static_cast<size_t>(src_row[k]) * p.d + col);
          const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
          const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
          acc.x = __fadd_rn(acc.x, __fmul_rn(wk[k], y01.x));
          acc.y = __fadd_rn(acc.y, __fmul_rn(wk[k], y01.y));
          acc.z = __fadd_rn(acc.z, __fmul_rn(wk[k], y23.x));
          acc.w = __fadd_rn(acc.w, __fmul_rn(wk[k], y23.y));
        }
      float4 x = xv[i];
      x = make_float4(x.x + acc.x, x.y + acc.y, x.z + acc.z, x.w + acc.w);
      xv[i] = x;
      ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(p.x_out + static_cast<size_t>(row) * p.d + col) = xv[i];
    if (p.x_copy) *reinterpret_cast<float4*>(p.x_copy + static_cast<size_t>(row) * p.d + col) = xv[i];
  }
  if (p.mode == 2) return;
  ss = warp_sum(ss);
  const float rn = rms_inv_denominator(ss, p.inv_sqrt_d, p.eps);
  // normalise in registers first (loads only), store afterwards
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 x = xv[i];
      const float4 g = ldg4(p.g_next + col);
      float4 y = make_float4((x.x * rn) * g.x, (x.y * rn) * g.y, (x.z * rn) * g.z, (x.w * rn) * g.w);
      if (p.mode == 0) {
        const float4 c = *reinterpret_cast<const float4*>(p.cvec + static_cast<size_t>(row / p.Tc) * p.d + col);
        y = make_float4(y.x + c.x, y.y + c.y, y.z + c.z, y.w + c.w);
      }
      xv[i] = y;
    }
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 y = xv[i];
      if (p.mode == 0)
        *reinterpret_cast<uint2*>(p.hA + static_cast<size_t>(row) * p.d + col) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      else
        *reinterpret_cast<float4*>(p.xnorm + static_cast<size_t>(row) * p.d + col) = y;
    }
}
template <int NVEC>
__device__ __forceinline__ void combine_body(const CombineParams& p, int vblock) {
  if (p.K == 2)
    combine_impl<NVEC, 2>(p, vblock);
  else
    combine_impl<NVEC, 0>(p, vblock);
}
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32, 4) combine_kernel(const CombineParams p) {
  pdl_trigger();
  pdl_wait();
  combine_body<NVEC>(p, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------------------
// Output head + EDM preconditioning + sampler update, one warp per action token:
//   F   = out(ln(x)[:, -A:])                                   modedit.py:806-808 (Linear(d, action_dim) + bias)
//   D   = c_out * F + c_skip * x_act                           score_wrappers.py:79-80 (mode >= 1)
//   x'  = ratio * x_act + coef * D                             gc_sampling.py:948-950 DDIM / DPM-Solver-1 (mode 2)
//   err = sum (F - target)^2, target = (a - c_skip*noised)/c_out   score_wrappers.py:58-62 (mode 3, loss)
struct HeadParams {
  StepScalars sc;
  const float* xnorm;      // [B*T, d] final-ln output
  const float* w_out;      // [action_dim, d]
  const float* b_out;      // [action_dim]
  const float* x_act;      // [B, A, action_dim] the (unscaled) noisy actions fed to this evaluation
  float* out;              // [B, A, action_dim]
  const float* clean;      // mode 3: clean actions
  float* tok_sqerr;        // mode 3: [B*A] per-token sum of squared errors (reduced deterministically afterwards)
  const float* coefs;      // modes 2, 4, 5: this step's update coefficients (host fp32, reference op order; engine.cu sampler_schedule)
  float* d_prev;           // mode 5: the previous step's denoised actions [B, A, action_dim] (read, then overwritten)
  int B, T, A, action_dim, d;
  int mode;                // 0 raw F, 1 denoised D, 2 DDIM update, 3 loss (out <- F), 4 Euler update, 5 DPM-Solver++(2M) update,
                           // 6 sampler program (below)
  // mode 6 — one row of a "sampler program" (mode_sample_program): every k-diffusion update of gc_sampling.py is a linear
  // combination of the step's base sample X, the probe P fed to a second evaluation, this evaluation's denoised D, up to
  // four history tensors H (Heun's first slope, the LMS derivative history) and a caller-drawn noise tensor:
  //   dst <- cX*X + cP*P + cD*D + sum_j cH[j]*H[j] + cN*noise,  dst = X or P;   H[slot] <- hX*x_in + hD*D  (optional)
  // prog = {cX, cP, cD, cH0..cH3, cN, hX, hD, slot (-1: none), dst (0: X, 1: P)} as 12 floats, written by the host.
  const float* prog;
  float* xbase;            // X  [B, A, action_dim]
  float* xprobe;           // P
  float* hist;             // H  [4][B*A*action_dim]
  const float* noise;      // this evaluation's noise tensor or nullptr
};
constexpr int HEAD_PROG_FLOATS = 16;
// body of head_kernel for virtual block `vblock` (also a phase of the persistent small-batch kernel, small_eval.cuh)
template <int NVEC>
__device__ __forceinline__ void head_body(const HeadParams& p, int vblock) {
  const int item = vblock * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= p.B * p.A) return;
  const int b = item / p.A, j = item % p.A;
  const int row = b * p.T + (p.T - p.A) + j;
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const float4 x = *reinterpret_cast<const float4*>(p.xnorm + static_cast<size_t>(row) * p.d + col);
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (a < p.action_dim) {
          const float4 w = ldg4(p.w_out + static_cast<size_t>(a) * p.d + col);
          acc[a] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[a]))));
        }
    }
  float mine = 0.f;
#pragma unroll
  for (int a = 0; a < 8; ++a)
    if (a < p.action_dim) {
      const float v = warp_sum(acc[a]);
      if (lane == a) mine = v + p.b_out[a];
    }
  float sq = 0.f;
  if (lane < p.action_dim) {
    const size_t o = (static_cast<size_t>(b) * p.A + j) * p.action_dim + lane;
    float result = mine;
    if (p.mode >= 1) {
      const float sigma = load_sigma(p.sc, b);
      const float sd = p.sc.sigma_data;
      const float s2 = sigma * sigma + sd * sd;
      const float c_skip = (sd * sd) / s2;
      const float c_out = sigma * sd / sqrtf(s2);
      const float xa = p.x_act[o];
      if (p.mode == 3) {
        const float target = (p.clean[o] - c_skip * xa) / c_out;
        const float e = mine - target;
        sq = e * e;
      } else {
        // inner * c_out + action * c_skip, each product rounded separately as ATen does
        const float den = __fadd_rn(__fmul_rn(mine, c_out), __fmul_rn(xa, c_skip));
        if (p.mode == 2) {
          // (sigma_next / sigma) * action - expm1(-h) * denoised            gc_sampling.py:948-950
          result = __fsub_rn(__fmul_rn(p.coefs[0], xa), __fmul_rn(p.coefs[1], den));
        } else if (p.mode == 4) {
          // d = (action - denoised) / sigma;  action + d * (sigma_next - sigma)   gc_sampling.py:73-75, :207-209
          result = __fadd_rn(xa, __fmul_rn(__fdiv_rn(__fsub_rn(xa, den), sigma), p.coefs[0]));
        } else if (p.mode == 5) {
          // denoised_d = (1 + 1/(2r)) * denoised - (1/(2r)) * old_denoised (first-order on the first / last step)   :722-731
          float target = den;
          if (p.coefs[3] != 0.f) target = __fsub_rn(__fmul_rn(p.coefs[2], den), __fmul_rn(p.coefs[3], p.d_prev[o]));
          p.d_prev[o] = den;
          result = __fsub_rn(__fmul_rn(p.coefs[0], xa), __fmul_rn(p.coefs[1], target));
        } else if (p.mode == 6) {
          const float* c = p.prog;
          const size_t n_el = static_cast<size_t>(p.B) * p.A * p.action_dim;
          float v = c[2] * den;
          if (c[0] != 0.f) v = fmaf(c[0], p.xbase[o], v);
          if (c[1] != 0.f) v = fmaf(c[1], p.xprobe[o], v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c[3 + j] != 0.f) v = fmaf(c[3 + j], p.hist[j * n_el + o], v);
          if (c[7] != 0.f) v = fmaf(c[7], p.noise[o], v);
          const int slot = static_cast<int>(c[10]);
          if (slot >= 0) p.hist[slot * n_el + o] = fmaf(c[8], xa, c[9] * den);  // every source above was read first
          (c[11] != 0.f ? p.xprobe : p.xbase)[o] = v;
          result = v;
        } else {
          result = den;
        }
      }
    }
    if (p.out) p.out[o] = result;
  }
  if (p.mode == 3) {
    sq = warp_sum(sq);
    if (lane == 0) p.tok_sqerr[item] = sq;
  }
}
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32) head_kernel(const HeadParams p) {
  pdl_trigger();
  pdl_wait();
  head_body<NVEC>(p, blockIdx.x);
}

// Deterministic mean of the per-token squared errors: loss = sum / (B*A*action_dim)  (.pow(2).flatten(1).mean()).
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float* __restrict__ tok_sqerr, int n, float denom,
                                                          float* __restrict__ loss) {
  __shared__ float part[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += tok_sqerr[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    loss[0] = t / denom;
  }
}

}  // namespace mode
