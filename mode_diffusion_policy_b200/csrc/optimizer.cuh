// AdamW over the engine's flat gradient buffer, fused with the weight re-pack.
//
// The reference trains with torch.optim.AdamW on parameter groups split by name (mode_agent.py:362-384: names containing
// 'bias', 'LayerNorm' or 'embedding' get weight_decay 0). A step there costs three passes over 686 M parameters: the
// optimizer (read p, g, m, v; write p, m, v), and for this engine a re-pack of the fp32 masters into the bf16 layouts the
// GEMMs read, plus autograd handing out gradient copies. Here ONE launch walks every bound tensor: it reads the gradient
// section, updates the moments (engine-owned flat buffers with the gradient layout) and the caller-owned fp32 master in
// place, and writes the packed copy (bf16 or fp32, SwiGLU row interleave, transposed action embedding) in the same pass.
// Arithmetic follows torch's single-tensor AdamW: p *= 1 - lr*wd; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
// p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
#pragma once
#include "ptx.cuh"

namespace mode {

constexpr int OPT_BLOCK_ELEMS = 1024;  // 256 threads x 4 elements

struct OptTensor {
  float* p;                      // fp32 master (caller-owned, reference layout [rows, cols])
  void* dst;                     // packed engine copy
  unsigned long long g_off;      // element offset of this tensor in the flat gradient / moment buffers
  unsigned long long dst_row0;
  int rows, cols, swiglu_half, to_bf16, transpose, decay;
  unsigned block0;               // first block of the launch that belongs to this tensor
  unsigned pad;
};

struct AdamWParams {
  const OptTensor* tab;
  int n;
  const float* grads;
  float *m, *v;
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
  const float* grad_scale;  // device scalar multiplied into every gradient (the loss's incoming gradient), or null
  unsigned block_base;      // OptTensor::block0 of the first tensor of this launch (group launches start mid-table)
  // Exponential moving average of the updated weights in the same pass (reference mode/callbacks/ema.py:119-126:
  // diff = ema - w; diff *= 1 - decay; ema -= diff). Engine-owned flat buffer with the gradient layout; null = off.
  // ema_init: the first step seeds the average with the weights BEFORE the update (the callback clones them at
  // on_train_start, ema.py:96).
  float* ema;
  float ema_one_minus_decay;
  int ema_init;
  // Sharded update (data-parallel ranks each own 1/shard_world of every tensor of the launch, mode_optimizer_set_sharding):
  // the launch has 1/shard_world of the blocks; block j of a tensor's shard is block shard_rank * (blocks / shard_world) + j
  // of the tensor. The updated weights go to `staging` (bf16, gradient layout) instead of the packed copy: the ranks
  // all-gather the staging spans and re-pack them locally (pack_from_staging_kernel). shard_world <= 1: off.
  int shard_rank, shard_world;
  __nv_bfloat16* staging;
};

// block -> (tensor, first element). Shards: every tensor of the launch has a block count divisible by shard_world (host
// checked), so block0 / shard_world indexes the launch's blocks.
__device__ __forceinline__ OptTensor opt_find_tensor(const AdamWParams& a, size_t& base) {
  const unsigned W = a.shard_world > 1 ? a.shard_world : 1;
  const unsigned q = blockIdx.x + a.block_base / W;
  int lo = 0, hi = a.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (a.tab[mid].block0 / W <= q)
      lo = mid;
    else
      hi = mid - 1;
  }
  const OptTensor t = a.tab[lo];
  unsigned blk = q - t.block0 / W;
  if (W > 1) {
    const size_t numel = static_cast<size_t>(t.rows) * t.cols;
    blk += a.shard_rank * static_cast<unsigned>(numel / OPT_BLOCK_ELEMS / W);
  }
  base = static_cast<size_t>(blk) * OPT_BLOCK_ELEMS;
  return t;
}

__device__ __forceinline__ float ema_update(float ema, float p_old, float p_new, const AdamWParams& a) {
  const float e = a.ema_init ? p_old : ema;
  return __fsub_rn(e, __fmul_rn(__fsub_rn(e, p_new), a.ema_one_minus_decay));  // no FMA contraction: torch does sub, mul, sub
}

__device__ __forceinline__ int opt_dst_row(int r, int swiglu_half) {
  if (swiglu_half <= 0) return r;
  const int is_gate = r >= swiglu_half;
  const int rr = is_gate ? r - swiglu_half : r;
  return (rr / 128) * 256 + (is_gate ? 128 : 0) + (rr % 128);
}

__device__ __forceinline__ float adamw_update(float p, float g, float& m, float& v, const AdamWParams& a, float decay_mul) {
  p *= decay_mul;
  m = m + (g - m) * (1.0f - a.beta1);
  v = v * a.beta2 + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  return p - (a.lr / a.bc1) * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_pack_kernel(const AdamWParams a) {
  // which tensor does this block belong to? (binary search over the first-block table; a few hundred entries)
  size_t base;
  const OptTensor t = opt_find_tensor(a, base);
  const size_t numel = static_cast<size_t>(t.rows) * t.cols;
  const float gs = a.grad_scale ? *a.grad_scale : 1.0f;
  const float decay_mul = t.decay ? 1.0f - a.lr * a.wd : 1.0f;
  if (!t.transpose && (t.cols & 3) == 0 && (t.g_off & 3) == 0 && (reinterpret_cast<uintptr_t>(t.p) & 15) == 0) {
    const size_t i = base + threadIdx.x * 4;
    if (i >= numel) return;
    const size_t gi = t.g_off + i;  // sections are 128-byte aligned
    float4 p = *reinterpret_cast<const float4*>(t.p + i);
    const float4 p_old = p;
    const float4 g = *reinterpret_cast<const float4*>(a.grads + gi);
    float4 m = *reinterpret_cast<const float4*>(a.m + gi);
    float4 v = *reinterpret_cast<const float4*>(a.v + gi);
    p.x = adamw_update(p.x, g.x * gs, m.x, v.x, a, decay_mul);
    p.y = adamw_update(p.y, g.y * gs, m.y, v.y, a, decay_mul);
    p.z = adamw_update(p.z, g.z * gs, m.z, v.z, a, decay_mul);
    p.w = adamw_update(p.w, g.w * gs, m.w, v.w, a, decay_mul);
    *reinterpret_cast<float4*>(t.p + i) = p;
    *reinterpret_cast<float4*>(a.m + gi) = m;
    *reinterpret_cast<float4*>(a.v + gi) = v;
    if (a.ema) {
      float4 em = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!a.ema_init) em = *reinterpret_cast<const float4*>(a.ema + gi);
      em.x = ema_update(em.x, p_old.x, p.x, a);
      em.y = ema_update(em.y, p_old.y, p.y, a);
      em.z = ema_update(em.z, p_old.z, p.z, a);
      em.w = ema_update(em.w, p_old.w, p.w, a);
      *reinterpret_cast<float4*>(a.ema + gi) = em;
    }
    if (a.staging) {  // sharded: bf16 in gradient layout, packed after the all-gather (host: to_bf16 tensors only)
      uint2 pk;
      pk.x = pack_bf16x2(p.x, p.y);
      pk.y = pack_bf16x2(p.z, p.w);
      *reinterpret_cast<uint2*>(a.staging + gi) = pk;
      return;
    }
    const int r = static_cast<int>(i / t.cols), c = static_cast<int>(i % t.cols);
    const size_t o = (t.dst_row0 + opt_dst_row(r, t.swiglu_half)) * static_cast<size_t>(t.cols) + c;
    if (t.to_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(p.x, p.y);
      pk.y = pack_bf16x2(p.z, p.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(t.dst) + o) = pk;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(t.dst) + o) = p;
    }
    return;
  }
  for (int k = 0; k < 4; ++k) {  // small / oddly shaped tensors: one element at a time, coalesced across the block
    const size_t i = base + k * 256 + threadIdx.x;
    if (i >= numel) return;
    const size_t gi = t.g_off + i;
    float m = a.m[gi], v = a.v[gi];
    const float p_old = t.p[i];
    const float p = adamw_update(p_old, a.grads[gi] * gs, m, v, a, decay_mul);
    t.p[i] = p;
    a.m[gi] = m;
    a.v[gi] = v;
    if (a.ema) a.ema[gi] = ema_update(a.ema_init ? 0.f : a.ema[gi], p_old, p, a);
    const int r = static_cast<int>(i / t.cols), c = static_cast<int>(i % t.cols);
    if (t.transpose) {
      reinterpret_cast<float*>(t.dst)[static_cast<size_t>(c) * t.rows + r] = p;
    } else {
      const size_t o = (t.dst_row0 + opt_dst_row(r, t.swiglu_half)) * static_cast<size_t>(t.cols) + c;
      if (t.to_bf16)
        reinterpret_cast<__nv_bfloat16*>(t.dst)[o] = __float2bfloat16_rn(p);
      else
        reinterpret_cast<float*>(t.dst)[o] = p;
    }
  }
}

// After the all-gather of a sharded group's staging spans: every rank writes the packed bf16 copies of ALL of the group's
// tensors (SwiGLU row interleave, row offsets) from the gathered bf16 values. Same block -> tensor table, whole tensors.
__global__ void __launch_bounds__(256) pack_from_staging_kernel(AdamWParams a) {
  a.shard_world = 1;
  size_t base;
  const OptTensor t = opt_find_tensor(a, base);
  const size_t i = base + threadIdx.x * 4;
  if (i >= static_cast<size_t>(t.rows) * t.cols) return;
  const uint2 pk = *reinterpret_cast<const uint2*>(a.staging + t.g_off + i);
  const int r = static_cast<int>(i / t.cols), c = static_cast<int>(i % t.cols);
  const size_t o = (t.dst_row0 + opt_dst_row(r, t.swiglu_half)) * static_cast<size_t>(t.cols) + c;
  *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(t.dst) + o) = pk;
}

// ------------------------------------------------------------------------------------------------------------------
// Per-tensor sums of squares of the flat gradient buffer in two launches and ONE device->host copy: the reference's
// gradient monitoring (MoDEAgent.on_before_zero_grad, mode_agent.py:304-359) calls `.item()` several times per
// parameter (~1400 host synchronisations per step for the 12-layer model). Deterministic: fixed chunking, fixed order.
constexpr int SUMSQ_CHUNK = 8192;  // elements per block of the first pass

struct SumsqSegment {
  unsigned long long off, numel;  // span of the flat buffer
  unsigned block0;                // first block of the first pass that belongs to this segment
  unsigned pad;
};

__global__ void __launch_bounds__(256) grad_sumsq_partial_kernel(const float* __restrict__ grads, const SumsqSegment* __restrict__ seg,
                                                                 int n_seg, float* __restrict__ partial) {
  int lo = 0, hi = n_seg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg[mid].block0 <= blockIdx.x)
      lo = mid;
    else
      hi = mid - 1;
  }
  const SumsqSegment s = seg[lo];
  const size_t begin = static_cast<size_t>(blockIdx.x - s.block0) * SUMSQ_CHUNK;
  const size_t end = begin + SUMSQ_CHUNK < s.numel ? begin + SUMSQ_CHUNK : s.numel;
  float acc = 0.f;
  for (size_t i = begin + threadIdx.x; i < end; i += 256) {
    const float g = grads[s.off + i];
    acc = fmaf(g, g, acc);
  }
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w >= 1; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// one block per segment: its partials summed in double, fixed order
__global__ void __launch_bounds__(256) grad_sumsq_finish_kernel(const float* __restrict__ partial, const SumsqSegment* __restrict__ seg,
                                                                int n_blocks_total, int n_seg, float* __restrict__ out) {
  const int sI = blockIdx.x;
  const unsigned b0 = seg[sI].block0, b1 = sI + 1 < n_seg ? seg[sI + 1].block0 : static_cast<unsigned>(n_blocks_total);
  double acc = 0.0;
  for (unsigned b = b0 + threadIdx.x; b < b1; b += 256) acc += static_cast<double>(partial[b]);
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w >= 1; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[sI] = static_cast<float>(red[0]);
}

}  // namespace mode
