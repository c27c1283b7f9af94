// Backward (gradient) row kernels of the MoDE block for the training path (GCDenoiser.loss -> backward,
// reference score_wrappers.py:45-63 + torch.autograd over modedit.py). Deterministic mode of SURVEY.md A.5: dropout 0,
// top-k routing. GEMM-shaped gradients run on the tcgen05 kernels (dgrad: gemm.cuh with transposed weight copies,
// wgrad: gemm_wgrad.cuh); everything here is the HBM-bound glue between them. All reductions over rows have a fixed
// order (per-CTA partials + a column-sum kernel) so gradients are bit-reproducible.
#pragma once
#include "rowwise.cuh"

namespace mode {

// dx_i = r * (g_i dy_i - xh_i * mean_j(g_j dy_j xh_j)),  xh = x * r,  r = 1 / max(||x|| d^-1/2, eps)
// (RMSNorm.forward, modedit.py:72-80; when the clamp is active r is a constant and the second term vanishes)
struct RmsBwd {
  float r;        // 1 / denominator
  float coef;     // mean_j(g_j dy_j xh_j) (0 when clamped)
};

constexpr int COLSUM_CHUNKS = 32;  // row chunks of the two-level deterministic column sums

// ------------------------------------------------------------------------------------------------------------
// Loss + head backward: F = out(ln(x)[-A:]); loss = mean((F - target)^2) (score_wrappers.py:58-62).
// One warp per action token: recomputes F from the saved final-ln output, writes dF (for the head weight gradient),
// the loss partial, and the gradient w.r.t. the final-ln INPUT row (x_last) into dX; non-action rows get zeros.
struct HeadBwdParams {
  StepScalars sc;
  const float* xnorm;     // [B*T, d] final ln output (saved by the forward)
  const float* x_last;    // [B*T, d] final ln input
  const float* lnf_g;     // [d]
  const float* w_out;     // [adim, d]
  const float* b_out;     // [adim]
  const float* noised;    // [B, A, adim] noised actions fed to the network (before c_in)
  const float* clean;     // [B, A, adim]
  float* dF;              // [B*A, 8] (adim <= 8, padded)
  float* tok_sqerr;       // [B*A]
  float* dX;              // [B*T, d] out: gradient w.r.t. x_last
  float* g_part;          // [B*A, d] per-row contribution to d ln.g
  float* F_out;           // optional [B, A, adim]
  int B, T, A, action_dim, d;
  float inv_count;        // 1 / (B*A*adim)
  float eps, inv_sqrt_d;
};

template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32) head_bwd_kernel(const HeadBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int item = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= p.B * p.T) return;
  const int b = item / p.T, t = item % p.T;
  const int row = item;
  if (t < p.T - p.A) {  // not an action token: no gradient enters here from the head
#pragma unroll
    for (int i = 0; i < NVEC; ++i)
      *reinterpret_cast<float4*>(p.dX + static_cast<size_t>(row) * p.d + (i * 32 + lane) * 4) = make_float4(0, 0, 0, 0);
    return;
  }
  const int j = t - (p.T - p.A);
  const int tok = b * p.A + j;
  float4 xn[NVEC];
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    xn[i] = *reinterpret_cast<const float4*>(p.xnorm + static_cast<size_t>(row) * p.d + col);
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a < p.action_dim) {
        const float4 w = *reinterpret_cast<const float4*>(p.w_out + static_cast<size_t>(a) * p.d + col);
        acc[a] = fmaf(xn[i].x, w.x, fmaf(xn[i].y, w.y, fmaf(xn[i].z, w.z, fmaf(xn[i].w, w.w, acc[a]))));
      }
  }
  const float sigma = load_sigma(p.sc, b);
  const float sd = p.sc.sigma_data, s2 = sigma * sigma + sd * sd;
  const float c_skip = (sd * sd) / s2, c_out = sigma * sd / sqrtf(s2);
  float dF[8];
  float sq = 0.f;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    dF[a] = 0.f;
    if (a < p.action_dim) {
      const float F = warp_sum(acc[a]) + p.b_out[a];
      const size_t o = static_cast<size_t>(tok) * p.action_dim + a;
      const float target = (p.clean[o] - c_skip * p.noised[o]) / c_out;
      const float e = F - target;
      sq += e * e;
      dF[a] = 2.0f * e * p.inv_count;
      if (lane == 0 && p.F_out) p.F_out[o] = F;
    }
  }
  if (lane == 0) {
    p.tok_sqerr[tok] = sq;
#pragma unroll
    for (int a = 0; a < 8; ++a) p.dF[tok * 8 + a] = dF[a];
  }
  // d xnorm = dF . W_out ; then the final RMSNorm backward
  float4 dy[NVEC], xl[NVEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    float4 v = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a < p.action_dim) {
        const float4 w = *reinterpret_cast<const float4*>(p.w_out + static_cast<size_t>(a) * p.d + col);
        v.x = fmaf(dF[a], w.x, v.x); v.y = fmaf(dF[a], w.y, v.y); v.z = fmaf(dF[a], w.z, v.z); v.w = fmaf(dF[a], w.w, v.w);
      }
    dy[i] = v;
    xl[i] = *reinterpret_cast<const float4*>(p.x_last + static_cast<size_t>(row) * p.d + col);
    ss += xl[i].x * xl[i].x + xl[i].y * xl[i].y + xl[i].z * xl[i].z + xl[i].w * xl[i].w;
  }
  ss = warp_sum(ss);
  const float den = sqrtf(ss) * p.inv_sqrt_d;
  const bool clamped = den < p.eps;
  const float r = 1.0f / fmaxf(den, p.eps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float4 g = *reinterpret_cast<const float4*>(p.lnf_g + (i * 32 + lane) * 4);
    dot += g.x * dy[i].x * xl[i].x + g.y * dy[i].y * xl[i].y + g.z * dy[i].z * xl[i].z + g.w * dy[i].w * xl[i].w;
  }
  dot = clamped ? 0.f : warp_sum(dot) * r * r / static_cast<float>(p.d);  // mean_j(g dy xh) * r  (xh = x r)
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(p.lnf_g + col);
    float4 dx;
    dx.x = r * (g.x * dy[i].x - xl[i].x * dot);
    dx.y = r * (g.y * dy[i].y - xl[i].y * dot);
    dx.z = r * (g.z * dy[i].z - xl[i].z * dot);
    dx.w = r * (g.w * dy[i].w - xl[i].w * dot);
    *reinterpret_cast<float4*>(p.dX + static_cast<size_t>(row) * p.d + col) = dx;
    *reinterpret_cast<float4*>(p.g_part + static_cast<size_t>(tok) * p.d + col) =
        make_float4(dy[i].x * xl[i].x * r, dy[i].y * xl[i].y * r, dy[i].z * xl[i].z * r, dy[i].w * xl[i].w * r);
  }
}

// partial[chunk, a, :] = sum over the chunk's action tokens of dF[tok, a] * xnorm[row(tok), :]  (+ column d: dF sums)
// grid (ceil((d+1)/256), adim, COLSUM_CHUNKS); finished by colsum_f32_kernel over the chunks.
__global__ void __launch_bounds__(256) head_wgrad_partial_kernel(const float* __restrict__ dF, const float* __restrict__ xnorm,
                                                                 float* __restrict__ partial, int B, int T, int A, int adim,
                                                                 int d) {
  pdl_trigger();
  pdl_wait();
  const int col = blockIdx.x * 256 + threadIdx.x;  // col == d is the bias column
  const int a = blockIdx.y;
  if (col > d) return;
  const int n = B * A, per = (n + COLSUM_CHUNKS - 1) / COLSUM_CHUNKS;
  const int t0 = blockIdx.z * per, t1 = min(n, t0 + per);
  float acc = 0.f;
  for (int tok = t0; tok < t1; ++tok) {
    const int row = (tok / A) * T + (T - A) + tok % A;
    const float g = dF[tok * 8 + a];
    acc = col < d ? fmaf(g, xnorm[static_cast<size_t>(row) * d + col], acc) : acc + g;
  }
  partial[(static_cast<size_t>(blockIdx.z) * adim + a) * (d + 1) + col] = acc;
}
// scatter the finished [adim, d+1] sums into out.weight [adim, d] and out.bias [adim]
__global__ void __launch_bounds__(256) head_wgrad_finish_kernel(const float* __restrict__ partial, float* __restrict__ g_wout,
                                                                float* __restrict__ g_bout, int adim, int d) {
  pdl_trigger();
  pdl_wait();
  const int col = blockIdx.x * 256 + threadIdx.x;
  const int a = blockIdx.y;
  if (col > d) return;
  float acc = 0.f;
  for (int i = 0; i < COLSUM_CHUNKS; ++i) acc += partial[(static_cast<size_t>(i) * adim + a) * (d + 1) + col];
  if (col < d)
    g_wout[static_cast<size_t>(a) * d + col] = acc;
  else
    g_bout[a] = acc;
}

// ------------------------------------------------------------------------------------------------------------
// Combine backward: out = xn + sum_k w_k y_k  (modedit.py:561-566, :595)
//   dY[pos_k + t] = bf16(w_k * dOut)            (permuted rows; pad rows were zero-filled)
//   dw_row[row, k] = <dOut[row], y_k[row]>       (summed over the sample's tokens by the router backward)
struct CombineBwdParams {
  const float* dX;           // [B*T, d] gradient w.r.t. the block output (left untouched: it is also d xn)
  const __nv_bfloat16* y;    // [P, d] saved expert outputs
  const int* pos;            // [B, K]
  const float* w;            // [B, K]
  __nv_bfloat16* dY;         // [P, d]
  float* dw_row;             // [B*T, K]
  int B, T, K, d;
};
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32, 4) combine_bwd_kernel(const CombineBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.T) return;
  const int b = row / p.T, t = row % p.T;
  float4 g[NVEC];
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
    g[i] = *reinterpret_cast<const float4*>(p.dX + static_cast<size_t>(row) * p.d + (i * 32 + lane) * 4);
  for (int k = 0; k < p.K; ++k) {
    const int prow = p.pos[b * p.K + k] + t;
    const float wk = p.w[b * p.K + k];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int col = (i * 32 + lane) * 4;
      const uint2 raw = *reinterpret_cast<const uint2*>(p.y + static_cast<size_t>(prow) * p.d + col);
      const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      dot += g[i].x * y01.x + g[i].y * y01.y + g[i].z * y23.x + g[i].w * y23.y;
      *reinterpret_cast<uint2*>(p.dY + static_cast<size_t>(prow) * p.d + col) =
          make_uint2(pack_bf16x2(wk * g[i].x, wk * g[i].y), pack_bf16x2(wk * g[i].z, wk * g[i].w));
    }
    dot = warp_sum(dot);
    if (lane == 0) p.dw_row[row * p.K + k] = dot;
  }
}

// ------------------------------------------------------------------------------------------------------------
// SwiGLU backward on the packed pre-activations z = [128 projected | 128 gate] per 256 columns:
//   h = zp * silu(zg);  d zp = dh * silu(zg);  d zg = dh * zp * sig(zg) * (1 + zg * (1 - sig(zg)))
// Rows beyond a tile's valid count are written as zeros so that the weight-gradient GEMMs may contract over the padded
// row range. One thread per 8 hidden units (16-byte accesses).
struct SwigluBwdParams {
  const __nv_bfloat16* z;    // [P, 8d]
  const __nv_bfloat16* dH;   // [P, 4d]
  __nv_bfloat16* dZ;         // [P, 8d]
  const GemmMTile* tiles;    // this layer's up-projection tile table (tile i covers rows [i*tile_m, (i+1)*tile_m))
  const int* num_tiles;
  int tile_m, F;             // F = 4d
  DropoutSpec drop = DropoutSpec{0u, 0u, 1.0f};  // the forward's dropout on h (gemm.cuh SwiGLU epilogue): d swiglu = dH o keep / (1 - p)
  const int* row_token = nullptr;  // [P] token of every permuted row
  int E = 1, rows_per_expert = 1;    // expert of a tile = (w_row_base / rows_per_expert) % E
};
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const SwigluBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int n_rows = *p.num_tiles * p.tile_m;
  const int vec_per_row = p.F / 8;
  const size_t idx = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  const size_t row = idx / vec_per_row;
  if (row >= static_cast<size_t>(n_rows)) return;
  const int c = static_cast<int>(idx % vec_per_row) * 8;  // hidden-unit index of the first of 8
  const int blk = c / 128, in_blk = c % 128;
  const size_t zp_off = row * (2 * p.F) + blk * 256 + in_blk;
  const size_t zg_off = zp_off + 128;
  const GemmMTile tile = p.tiles[row / p.tile_m];
  uint4 ozp = make_uint4(0, 0, 0, 0), ozg = make_uint4(0, 0, 0, 0);
  if (static_cast<int>(row % p.tile_m) < tile.rows_valid) {
    const uint4 rzp = *reinterpret_cast<const uint4*>(p.z + zp_off);
    const uint4 rzg = *reinterpret_cast<const uint4*>(p.z + zg_off);
    const uint4 rdh = *reinterpret_cast<const uint4*>(p.dH + row * p.F + c);
    const __nv_bfloat162* zp2 = reinterpret_cast<const __nv_bfloat162*>(&rzp);
    const __nv_bfloat162* zg2 = reinterpret_cast<const __nv_bfloat162*>(&rzg);
    const __nv_bfloat162* dh2 = reinterpret_cast<const __nv_bfloat162*>(&rdh);
    uint32_t* o1 = reinterpret_cast<uint32_t*>(&ozp);
    uint32_t* o2 = reinterpret_cast<uint32_t*>(&ozg);
    uint32_t word0 = 0;
    if (p.drop.thr) {
      const uint32_t token = static_cast<uint32_t>(p.row_token[row]);
      const uint32_t expert = static_cast<uint32_t>((tile.w_row_base / p.rows_per_expert) % p.E);
      word0 = (token * p.E + expert) * static_cast<uint32_t>(p.F / 2) + static_cast<uint32_t>(c) / 2u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 zp = __bfloat1622float2(zp2[j]), zg = __bfloat1622float2(zg2[j]);
      float2 dh = __bfloat1622float2(dh2[j]);
      if (p.drop.thr) {
        const uint32_t bits = rng_bits(p.drop.key, word0 + j);
        dh.x = (bits & 0xffffu) < p.drop.thr ? 0.f : dh.x * p.drop.scale;
        dh.y = (bits >> 16) < p.drop.thr ? 0.f : dh.y * p.drop.scale;
      }
      const float s0 = 1.0f / (1.0f + __expf(-zg.x)), s1 = 1.0f / (1.0f + __expf(-zg.y));
      o1[j] = pack_bf16x2(dh.x * zg.x * s0, dh.y * zg.y * s1);
      o2[j] = pack_bf16x2(dh.x * zp.x * s0 * (1.0f + zg.x * (1.0f - s0)), dh.y * zp.y * s1 * (1.0f + zg.y * (1.0f - s1)));
    }
  }
  *reinterpret_cast<uint4*>(p.dZ + zp_off) = ozp;
  *reinterpret_cast<uint4*>(p.dZ + zg_off) = ozg;
}

// ------------------------------------------------------------------------------------------------------------
// ln_2 backward (+ un-permute of the expert-input gradient): the block replaces its residual by xn = rms(x1) g2
// (modedit.py:539), and xn feeds both the output (identity) and the k expert groups:
//   d xn = dOut + sum_k dXp[pos_k + t];   d x1 = rms_bwd(x1, g2, d xn)
// Writes d x1 to dX (fp32, the running residual-stream gradient) and as bf16 (operand of the c_proj gradients).
struct Ln2BwdParams {
  float* dX;                   // [B*T, d] in: dOut, out: d x1
  const __nv_bfloat16* dXp;    // [P, d] gradient w.r.t. the permuted expert inputs
  const int* pos;              // [B, K]
  const float* x1;             // [B*T, d] saved ln_2 input
  const float* g;              // [d]
  __nv_bfloat16* dx1_bf16;     // [B*T (padded), d]
  float* g_part;               // [B*T, d] per-row contribution to d ln_2.g
  int B, T, K, d;
  float eps, inv_sqrt_d;
};
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32) ln2_bwd_kernel(const Ln2BwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.T) return;
  const int b = row / p.T, t = row % p.T;
  float4 dy[NVEC], x[NVEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    dy[i] = *reinterpret_cast<const float4*>(p.dX + static_cast<size_t>(row) * p.d + col);
    x[i] = *reinterpret_cast<const float4*>(p.x1 + static_cast<size_t>(row) * p.d + col);
    ss += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
  }
  for (int k = 0; k < p.K; ++k) {
    const int prow = p.pos[b * p.K + k] + t;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const uint2 raw = *reinterpret_cast<const uint2*>(p.dXp + static_cast<size_t>(prow) * p.d + (i * 32 + lane) * 4);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      dy[i].x += a.x; dy[i].y += a.y; dy[i].z += c.x; dy[i].w += c.y;
    }
  }
  ss = warp_sum(ss);
  const float den = sqrtf(ss) * p.inv_sqrt_d;
  const bool clamped = den < p.eps;
  const float r = 1.0f / fmaxf(den, p.eps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float4 g = *reinterpret_cast<const float4*>(p.g + (i * 32 + lane) * 4);
    dot += g.x * dy[i].x * x[i].x + g.y * dy[i].y * x[i].y + g.z * dy[i].z * x[i].z + g.w * dy[i].w * x[i].w;
  }
  dot = clamped ? 0.f : warp_sum(dot) * r * r / static_cast<float>(p.d);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(p.g + col);
    float4 dx;
    dx.x = r * (g.x * dy[i].x - x[i].x * dot);
    dx.y = r * (g.y * dy[i].y - x[i].y * dot);
    dx.z = r * (g.z * dy[i].z - x[i].z * dot);
    dx.w = r * (g.w * dy[i].w - x[i].w * dot);
    *reinterpret_cast<float4*>(p.dX + static_cast<size_t>(row) * p.d + col) = dx;
    *reinterpret_cast<uint2*>(p.dx1_bf16 + static_cast<size_t>(row) * p.d + col) =
        make_uint2(pack_bf16x2(dx.x, dx.y), pack_bf16x2(dx.z, dx.w));
    *reinterpret_cast<float4*>(p.g_part + static_cast<size_t>(row) * p.d + col) =
        make_float4(dy[i].x * x[i].x * r, dy[i].y * x[i].y * r, dy[i].z * x[i].z * r, dy[i].w * x[i].w * r);
  }
}

// ------------------------------------------------------------------------------------------------------------
// ln_1 backward: hA = rms(x_in) g1 + c  (modedit.py:532).  d x_in = d x1 (residual) + rms_bwd(x_in, g1, d hA);
// d c[b] += sum_t d hA[b, t].  One warp per row; the per-sample sum of d hA is taken afterwards (dc_reduce_kernel).
struct Ln1BwdParams {
  float* dX;                    // [B*T, d] in: d x1, out: d x_in
  const __nv_bfloat16* dhA;     // [B*T, d]
  const float* x_in;            // [B*T, d]
  const float* g;               // [d]
  float* g_part;                // [B*T, d]
  int B, T, d;
  float eps, inv_sqrt_d;
};
template <int NVEC>
__global__ void __launch_bounds__(ROW_WARPS * 32) ln1_bwd_kernel(const Ln1BwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= p.B * p.T) return;
  float4 dy[NVEC], x[NVEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const uint2 raw = *reinterpret_cast<const uint2*>(p.dhA + static_cast<size_t>(row) * p.d + col);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    dy[i] = make_float4(a.x, a.y, c.x, c.y);
    x[i] = *reinterpret_cast<const float4*>(p.x_in + static_cast<size_t>(row) * p.d + col);
    ss += x[i].x * x[i].x + x[i].y * x[i].y + x[i].z * x[i].z + x[i].w * x[i].w;
  }
  ss = warp_sum(ss);
  const float den = sqrtf(ss) * p.inv_sqrt_d;
  const bool clamped = den < p.eps;
  const float r = 1.0f / fmaxf(den, p.eps);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float4 g = *reinterpret_cast<const float4*>(p.g + (i * 32 + lane) * 4);
    dot += g.x * dy[i].x * x[i].x + g.y * dy[i].y * x[i].y + g.z * dy[i].z * x[i].z + g.w * dy[i].w * x[i].w;
  }
  dot = clamped ? 0.f : warp_sum(dot) * r * r / static_cast<float>(p.d);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(p.g + col);
    float4 dx = *reinterpret_cast<const float4*>(p.dX + static_cast<size_t>(row) * p.d + col);
    dx.x += r * (g.x * dy[i].x - x[i].x * dot);
    dx.y += r * (g.y * dy[i].y - x[i].y * dot);
    dx.z += r * (g.z * dy[i].z - x[i].z * dot);
    dx.w += r * (g.w * dy[i].w - x[i].w * dot);
    *reinterpret_cast<float4*>(p.dX + static_cast<size_t>(row) * p.d + col) = dx;
    *reinterpret_cast<float4*>(p.g_part + static_cast<size_t>(row) * p.d + col) =
        make_float4(dy[i].x * x[i].x * r, dy[i].y * x[i].y * r, dy[i].z * x[i].z * r, dy[i].w * x[i].w * r);
  }
}

// dc[b, :] += sum_t src[b*T + t, :]  (src: bf16 d hA rows; optionally also the fp32 row `extra_t` of dX, used once for
// the sigma token whose embedding IS c). One thread per (b, 4 columns); fixed order over t.
__global__ void __launch_bounds__(256) dc_reduce_kernel(const __nv_bfloat16* __restrict__ dhA, const float* __restrict__ dX,
                                                        int extra_t, float* __restrict__ dc, int B, int T, int d) {
  pdl_trigger();
  pdl_wait();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int vec_per_row = d / 4;
  if (idx >= B * vec_per_row) return;
  const int b = idx / vec_per_row, col = (idx % vec_per_row) * 4;
  float4 acc = *reinterpret_cast<const float4*>(dc + static_cast<size_t>(b) * d + col);
  if (dhA)
    for (int t = 0; t < T; ++t) {
      const uint2 raw = *reinterpret_cast<const uint2*>(dhA + (static_cast<size_t>(b) * T + t) * d + col);
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      acc.x += a.x; acc.y += a.y; acc.z += c.x; acc.w += c.y;
    }
  if (dX && extra_t >= 0) {
    const float4 v = *reinterpret_cast<const float4*>(dX + (static_cast<size_t>(b) * T + extra_t) * d + col);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(dc + static_cast<size_t>(b) * d + col) = acc;
}

// Deterministic column sums over many rows, two levels: COLSUM_CHUNKS partial sums over contiguous row chunks (one CTA
// per (256 columns, chunk): coalesced reads, fixed order inside the chunk), then a fixed-order sum over the chunks.

// partial[chunk, c] = sum over the chunk's rows of part[row, c]      (fp32 per-row contributions, e.g. norm gains)
__global__ void __launch_bounds__(256) colsum_f32_partial_kernel(const float* __restrict__ part, int n, int cols,
                                                                 float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int per = (n + COLSUM_CHUNKS - 1) / COLSUM_CHUNKS;
  const int r0 = blockIdx.y * per, r1 = min(n, r0 + per);
  float acc = 0.f;
#pragma unroll 8  // the loads are independent: keep 8 in flight per thread (the adds stay in row order)
  for (int i = r0; i < r1; ++i) acc += part[static_cast<size_t>(i) * cols + c];
  partial[static_cast<size_t>(blockIdx.y) * cols + c] = acc;
}
// out[c] = sum_chunk partial[chunk, c]
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ partial, int n, int cols,
                                                         float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += partial[static_cast<size_t>(i) * cols + c];
  out[c] = acc;
}

// Bias gradients from bf16 activations-gradients: grid (col blocks, chunks, problems); a problem is a row range
// (an expert's token group, or all rows).  partial[(problem, chunk), c]
__global__ void __launch_bounds__(256) colsum_bf16_partial_kernel(const __nv_bfloat16* __restrict__ src, int ld,
                                                                  const WgradProblem* __restrict__ problems,
                                                                  int cols_per_problem, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols_per_problem) return;
  const WgradProblem pr = problems[blockIdx.z];
  const int n = pr.k_blocks * 64;
  const int per = (n + COLSUM_CHUNKS - 1) / COLSUM_CHUNKS;
  const int r0 = blockIdx.y * per, r1 = min(n, r0 + per);
  float acc = 0.f;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(src[static_cast<size_t>(pr.row0 + r) * ld + c]);
  partial[(static_cast<size_t>(blockIdx.z) * COLSUM_CHUNKS + blockIdx.y) * cols_per_problem + c] = acc;
}
// out[out_row_base + remap(c)] = sum_chunk partial[(problem, chunk), c]; swiglu_half > 0 un-interleaves packed columns
__global__ void __launch_bounds__(256) colsum_bf16_finish_kernel(const float* __restrict__ partial,
                                                                 const WgradProblem* __restrict__ problems,
                                                                 int cols_per_problem, int swiglu_half,
                                                                 float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols_per_problem) return;
  const WgradProblem pr = problems[blockIdx.y];
  float acc = 0.f;
  for (int i = 0; i < COLSUM_CHUNKS; ++i)
    acc += partial[(static_cast<size_t>(blockIdx.y) * COLSUM_CHUNKS + i) * cols_per_problem + c];
  int oc = c;
  if (swiglu_half > 0) {
    const int blk = c / 256, in_blk = c % 256;
    oc = (in_blk < 128 ? 0 : swiglu_half) + blk * 128 + (in_blk % 128);
  }
  out[pr.out_row_base + oc] = acc;  // out_row_base counts weight rows of the problem == bias entries of the problem
}

// ------------------------------------------------------------------------------------------------------------
// Attention backward for one (sample, head) per warp (T <= 32): recomputes the normalised q, k and the softmax from the
// saved qkv, then  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - rowsum(dP o P)) / sqrt(Dh),  dq^ = dS k^,  dk^ = dS^T q^,
// and the per-head RMSNorm backward (Attention.forward, modedit.py:141-149). fp32 SIMT: 0.05 % of the step's FLOPs.
struct AttnBwdParams {
  const __nv_bfloat16* qkv;    // [B*T, 3d] saved
  const __nv_bfloat16* dO;     // [B*T, d]
  __nv_bfloat16* dqkv;         // [B*T, 3d]
  const float* q_gain;         // [Dh]
  const float* k_gain;
  float* gq_part;              // [B*H, Dh] per-(sample, head) contribution to d q_norm.g
  float* gk_part;
  int B, T, H;
  float eps, inv_sqrt_dh;
  DropoutSpec drop = DropoutSpec{0u, 0u, 1.0f};  // the forward's attention-probability dropout (attention.cuh); thr == 0: none
};
constexpr int ATTN_BWD_MAX_T = 32;
// per-warp shared memory: q^, k^, v, dO as bf16 [T][DH] (all four ARE bf16 values: rounded operands / bf16 tensors),
// P and dS fp32 [T][32], the two reciprocal-norm vectors. 18 KB at T=14, Dh=128 -> 12 warps per SM.
__host__ __device__ constexpr int attn_bwd_bytes_per_warp(int T, int DH) {
  return 4 * T * DH * 2 + (2 * T * ATTN_BWD_MAX_T + 2 * ATTN_BWD_MAX_T) * 4;
}
template <int DH>
__global__ void __launch_bounds__(ATTN_WARPS * 32) attention_bwd_kernel(const AttnBwdParams p) {
  constexpr int DPL = DH / 32;  // head dims per lane
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t attn_bwd_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * ATTN_WARPS + warp;
  pdl_wait();
  if (item >= p.B * p.H) return;
  const int b = item / p.H, h = item % p.H;
  const int T = p.T, d = p.H * DH;
  // per-warp shared memory: qh, kh (bf16-rounded normalised q, k with gain), v, dO : [T][DH] bf16;
  // P, dS : [T][32] fp32; rq, rk : [T]. The gain-free normalised q, k of the RMSNorm backward are recomputed from the
  // (L2-resident) qkv rows at the end instead of being kept.
  uint8_t* wbase = attn_bwd_smem + static_cast<size_t>(warp) * attn_bwd_bytes_per_warp(T, DH);
  __nv_bfloat16* qh = reinterpret_cast<__nv_bfloat16*>(wbase);
  __nv_bfloat16* kh = qh + T * DH;
  __nv_bfloat16* vv = kh + T * DH;
  __nv_bfloat16* dO = vv + T * DH;
  float* P = reinterpret_cast<float*>(dO + T * DH);
  float* dS = P + T * ATTN_BWD_MAX_T;
  float* rq = dS + T * ATTN_BWD_MAX_T;
  float* rk = rq + ATTN_BWD_MAX_T;
  float gq[DPL], gk[DPL];
#pragma unroll
  for (int u = 0; u < DPL; ++u) {
    gq[u] = p.q_gain[lane + 32 * u];
    gk[u] = p.k_gain[lane + 32 * u];
  }
  for (int t = 0; t < T; ++t) {
    const __nv_bfloat16* src = p.qkv + (static_cast<size_t>(b) * T + t) * 3 * d + h * DH;
    float q[DPL], k[DPL];
    float sq = 0.f, sk = 0.f;
#pragma unroll
    for (int u = 0; u < DPL; ++u) {
      const int c = lane + 32 * u;
      q[u] = __bfloat162float(src[c]);
      k[u] = __bfloat162float(src[d + c]);
      vv[t * DH + c] = src[2 * d + c];
      dO[t * DH + c] = p.dO[(static_cast<size_t>(b) * T + t) * d + h * DH + c];
      sq += q[u] * q[u];
      sk += k[u] * k[u];
    }
    sq = warp_sum(sq);
    sk = warp_sum(sk);
    const float r_q = 1.0f / fmaxf(sqrtf(sq) * p.inv_sqrt_dh, p.eps), r_k = 1.0f / fmaxf(sqrtf(sk) * p.inv_sqrt_dh, p.eps);
    if (lane == 0) {
      rq[t] = (sqrtf(sq) * p.inv_sqrt_dh < p.eps) ? -r_q : r_q;  // sign bit marks an active clamp
      rk[t] = (sqrtf(sk) * p.inv_sqrt_dh < p.eps) ? -r_k : r_k;
    }
#pragma unroll
    for (int u = 0; u < DPL; ++u) {
      const int c = lane + 32 * u;
      qh[t * DH + c] = __float2bfloat16_rn(q[u] * r_q * gq[u]);
      kh[t * DH + c] = __float2bfloat16_rn(k[u] * r_k * gk[u]);
    }
  }
  __syncwarp();
  // scores, softmax, dP
  for (int i = 0; i < T; ++i) {
    float mx = -INFINITY;
    for (int j = 0; j <= i; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int u = 0; u < DPL; ++u) {
        const int c = lane + 32 * u;
        s = fmaf(__bfloat162float(qh[i * DH + c]), __bfloat162float(kh[j * DH + c]), s);
        dp = fmaf(__bfloat162float(dO[i * DH + c]), __bfloat162float(vv[j * DH + c]), dp);
      }
      s = warp_sum(s) * p.inv_sqrt_dh;
      dp = warp_sum(dp);
      if (lane == 0) {
        P[i * ATTN_BWD_MAX_T + j] = s;
        dS[i * ATTN_BWD_MAX_T + j] = dp;
      }
      mx = fmaxf(mx, s);
    }
    __syncwarp();
    // lane j handles column j of row i
    const bool on = lane <= i;
    const float e = on ? __expf(P[i * ATTN_BWD_MAX_T + lane] - mx) : 0.f;
    const float den = warp_sum(e);
    const float pij = e / den;
    // dropout: O = (P o m) V with m = keep / (1 - p)  ->  dP = (dO V^T) o m, and dV sees P o m
    float m = 1.0f;
    if (p.drop.thr) {
      const uint32_t word = (static_cast<uint32_t>(item) * T + i) * (static_cast<uint32_t>(T + 1) >> 1) + (lane >> 1);
      const uint32_t bits = rng_bits(p.drop.key, word);
      m = (((lane & 1) ? (bits >> 16) : (bits & 0xffffu)) < p.drop.thr) ? 0.f : p.drop.scale;
    }
    const float dpij = on ? dS[i * ATTN_BWD_MAX_T + lane] * m : 0.f;
    const float rowdot = warp_sum(pij * dpij);
    if (lane < T) {
      P[i * ATTN_BWD_MAX_T + lane] = pij * m;
      dS[i * ATTN_BWD_MAX_T + lane] = pij * (dpij - rowdot) * p.inv_sqrt_dh;
    }
    __syncwarp();
  }
  // dV, dq^, dk^ : each lane owns its head dims
  float gq_acc[DPL], gk_acc[DPL];
#pragma unroll
  for (int u = 0; u < DPL; ++u) gq_acc[u] = gk_acc[u] = 0.f;
  for (int t = 0; t < T; ++t) {
    float dv[DPL], dq[DPL], dk[DPL];
#pragma unroll
    for (int u = 0; u < DPL; ++u) dv[u] = dq[u] = dk[u] = 0.f;
    for (int j = 0; j <= t; ++j) {  // row t as a query: keys j <= t
      const float ds = dS[t * ATTN_BWD_MAX_T + j];
#pragma unroll
      for (int u = 0; u < DPL; ++u) dq[u] = fmaf(ds, __bfloat162float(kh[j * DH + lane + 32 * u]), dq[u]);
    }
    for (int i = t; i < T; ++i) {  // row t as a key/value: queries i >= t
      const float pp = P[i * ATTN_BWD_MAX_T + t], ds = dS[i * ATTN_BWD_MAX_T + t];
#pragma unroll
      for (int u = 0; u < DPL; ++u) {
        dv[u] = fmaf(pp, __bfloat162float(dO[i * DH + lane + 32 * u]), dv[u]);
        dk[u] = fmaf(ds, __bfloat162float(qh[i * DH + lane + 32 * u]), dk[u]);
      }
    }
    // per-head RMSNorm backward: q^ = qn * g, qn = q * r (recomputed from the saved row)
    const float r_q = fabsf(rq[t]), r_k = fabsf(rk[t]);
    const __nv_bfloat16* srow = p.qkv + (static_cast<size_t>(b) * T + t) * 3 * d + h * DH;
    float qn[DPL], kn[DPL];
    float dotq = 0.f, dotk = 0.f;
#pragma unroll
    for (int u = 0; u < DPL; ++u) {
      const int c = lane + 32 * u;
      qn[u] = __bfloat162float(srow[c]) * r_q;
      kn[u] = __bfloat162float(srow[d + c]) * r_k;
      dotq += gq[u] * dq[u] * qn[u];
      dotk += gk[u] * dk[u] * kn[u];
      gq_acc[u] += dq[u] * qn[u];
      gk_acc[u] += dk[u] * kn[u];
    }
    dotq = rq[t] < 0.f ? 0.f : warp_sum(dotq) / static_cast<float>(DH);
    dotk = rk[t] < 0.f ? 0.f : warp_sum(dotk) / static_cast<float>(DH);
    __nv_bfloat16* dst = p.dqkv + (static_cast<size_t>(b) * T + t) * 3 * d + h * DH;
#pragma unroll
    for (int u = 0; u < DPL; ++u) {
      const int c = lane + 32 * u;
      dst[c] = __float2bfloat16_rn(r_q * (gq[u] * dq[u] - qn[u] * dotq));
      dst[d + c] = __float2bfloat16_rn(r_k * (gk[u] * dk[u] - kn[u] * dotk));
      dst[2 * d + c] = __float2bfloat16_rn(dv[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < DPL; ++u) {
    p.gq_part[static_cast<size_t>(item) * DH + lane + 32 * u] = gq_acc[u];
    p.gk_part[static_cast<size_t>(item) * DH + lane + 32 * u] = gk_acc[u];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Router backward for one (layer, sample) per warp. Forward (RouterCond, modedit.py:343-352, :418-419):
//   z = W1 c + b1; hid = gelu(z); logits = W2 hid + b2; p = clamp(softmax(logits)); w_k = p_sel_k / sum_sel p
// Input: d w_k (sum over the sample's tokens of <dOut, y_k>). Outputs: d z (bf16 [B, 2d], feeds the W1 / c gradients
// through the tensor-core GEMMs), hid (bf16, operand of the W2 gradient is tiny: done here), d logits.
struct RouterBwdParams {
  StepScalars sc;
  const float* ra;        // [Hd] this layer
  const float* rb;
  const float* w2;        // [E, Hd]
  const float* probs;     // [B, E] clamped softmax (saved)
  const int* sel_idx;     // [B, K]
  const float* dw_row;    // [B*T, K]
  __nv_bfloat16* dz;      // [B (padded to 64), Hd]
  float* dlogit;          // [B, E]
  float* hid;             // [B, Hd] fp32 (for the W2 gradient)
  int B, T, E, K, Hd;
  int normalize;
  int per_token = 0;      // 1: sel_idx is [B*T, K] (multinomial routing draws per token); 0: [B, K] shared by the sample
};
// One CTA per sample: every warp recomputes the (tiny) per-expert prelude, the 2d hidden units are split over the CTA.
__global__ void __launch_bounds__(ROW_WARPS * 32) router_bwd_kernel(const RouterBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31;
  // lane e owns expert e
  const float pe = lane < p.E ? p.probs[b * p.E + lane] : 0.f;
  float dp = 0.f;
  // per_token: every token has its own selection; d p accumulates over the sample's tokens (the probabilities are shared)
  for (int tt = 0; tt < (p.per_token ? p.T : 1); ++tt) {
    float dwk = 0.f;   // d w for the slot whose expert is this lane
    bool selected = false;
    for (int k = 0; k < p.K; ++k) {
      const int e = p.per_token ? p.sel_idx[(b * p.T + tt) * p.K + k] : p.sel_idx[b * p.K + k];
      float s = 0.f;
      if (p.per_token) {
        s = p.dw_row[(b * p.T + tt) * p.K + k];
      } else {
        for (int t = 0; t < p.T; ++t) s += p.dw_row[(b * p.T + t) * p.K + k];
      }
      if (lane == e) {
        dwk = s;
        selected = true;
      }
    }
    // w_e = p_e / S over selected (normalize) -> d p_e = (dw_e - sum_sel(dw w)) / S
    if (p.normalize) {
      const float S = warp_sum(selected ? pe : 0.f);
      const float dot = warp_sum(selected ? dwk * pe / S : 0.f);
      dp += selected ? (dwk - dot) / S : 0.f;
    } else {
      dp += selected ? dwk : 0.f;
    }
  }
  // clamp(p, 1e-9, 1-1e-9): gradient passes only strictly inside the interval
  if (!(pe > 1e-9f && pe < 1.0f - 1e-9f)) dp = 0.f;
  // softmax backward: d logit_e = p_e (dp_e - sum_j dp_j p_j)
  const float sdp = warp_sum(lane < p.E ? dp * pe : 0.f);
  const float dlog = lane < p.E ? pe * (dp - sdp) : 0.f;
  if (lane < p.E && threadIdx.x < 32) p.dlogit[b * p.E + lane] = dlog;
  // d hid = W2^T dlogit ; d z = d hid * gelu'(z)
  const float s = logf(load_sigma(p.sc, b)) / 4.0f;
  for (int j = threadIdx.x; j < p.Hd; j += ROW_WARPS * 32) {  // Hd is a multiple of 256: uniform trip count
    const float z = fmaf(s, p.ra[j], p.rb[j]);
    const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * z * z);
    float dh = 0.f;
    for (int e = 0; e < p.E; ++e) dh = fmaf(__shfl_sync(0xffffffffu, dlog, e), p.w2[static_cast<size_t>(e) * p.Hd + j], dh);
    p.hid[static_cast<size_t>(b) * p.Hd + j] = z * cdf;
    p.dz[static_cast<size_t>(b) * p.Hd + j] = __float2bfloat16_rn(dh * (cdf + z * pdf));
  }
}

// g_w2[e, j] = sum_b dlogit[b, e] hid[b, j];  g_b2[e] = sum_b dlogit[b, e];  g_b1[j] = sum_b dz[b, j]  (fixed order).
// grid (Hd/256, E + 1): blockIdx.y < E handles expert y's row of W2 (and its bias), blockIdx.y == E the b1 sums, so the
// E + 1 serial loops over the batch run side by side instead of one after the other in each thread.
__global__ void __launch_bounds__(256) router_wgrad_small_kernel(const float* __restrict__ dlogit, const float* __restrict__ hid,
                                                                 const __nv_bfloat16* __restrict__ dz, float* __restrict__ g_w2,
                                                                 float* __restrict__ g_b2, float* __restrict__ g_b1, int B,
                                                                 int E, int Hd) {
  pdl_trigger();
  pdl_wait();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= Hd) return;
  const int e = blockIdx.y;
  if (e == E) {
    float b1 = 0.f;
#pragma unroll 4
    for (int b = 0; b < B; ++b) b1 += __bfloat162float(dz[static_cast<size_t>(b) * Hd + j]);
    g_b1[j] = b1;
    return;
  }
  float acc = 0.f, accb = 0.f;
#pragma unroll 4
  for (int b = 0; b < B; ++b) {
    const float g = dlogit[b * E + e];
    acc = fmaf(g, hid[static_cast<size_t>(b) * Hd + j], acc);
    accb += g;
  }
  g_w2[static_cast<size_t>(e) * Hd + j] = acc;
  if (j == 0) g_b2[e] = accb;
}

// ------------------------------------------------------------------------------------------------------------
// Embedding backward (MoDeDiT.forward, modedit.py:754-790): from the gradient of the input sequence dX [B, T, d]
//   goal row  -> dgoal (bf16, operand of the goal_emb weight gradient), pos[0]
//   image rows -> dstate (bf16, operand of the tok_emb weight gradient), pos[1]
//   action rows -> action_emb.weight gradient, pos[1 + j]
// (the sigma-token row is added to dc by dc_reduce_kernel).
// Embedding-dropout backward: dX[row, :] *= keep/(1-p) for every non-sigma token row, in place, before embed_bwd_kernel.
__global__ void __launch_bounds__(256) embed_dropout_bwd_kernel(float* __restrict__ dX, int rows, int T, int d, DropoutSpec drop) {
  pdl_trigger();
  pdl_wait();
  const size_t i4 = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;  // float4 index
  if (i4 >= static_cast<size_t>(rows) * d / 4) return;
  const size_t e0 = i4 * 4;
  if ((e0 / d) % T == 0) return;  // sigma token: not dropped
  const float4 m = dropout_mask4(drop, static_cast<uint32_t>(e0));
  float4 g = reinterpret_cast<float4*>(dX)[i4];
  g = make_float4(g.x * m.x, g.y * m.y, g.z * m.z, g.w * m.w);
  reinterpret_cast<float4*>(dX)[i4] = g;
}

struct EmbedBwdParams {
  StepScalars sc;
  const float* dX;           // [B*T, d]
  const float* noised;       // [B, A, adim]
  __nv_bfloat16* dgoal;      // [B (padded), d]
  __nv_bfloat16* dstate;     // [B*S (padded), d]
  float* g_pos;              // [1 + A, d]
  float* g_wact;             // [d, adim]
  int B, T, S, A, adim, d;
};
// grid (d / 256, A + 2, EMBED_BWD_CHUNKS): blockIdx.y <= A -> position row y; A + 1 -> action_emb weight (8 values per
// column); blockIdx.z = chunk of samples. Partial sums go to `partial[chunk][y][col(*8)]`, finished by
// embed_bwd_finish_kernel in a fixed order.
constexpr int EMBED_BWD_CHUNKS = 16;
__global__ void __launch_bounds__(256) embed_bwd_kernel(const EmbedBwdParams p, float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= p.d) return;
  const int y = blockIdx.y;
  const int per = (p.B + EMBED_BWD_CHUNKS - 1) / EMBED_BWD_CHUNKS;
  const int b0 = blockIdx.z * per, b1 = min(p.B, b0 + per);
  // partial layout: [chunk][(A + 1) rows of d | d * 8 action-weight values]
  float* mine = partial + static_cast<size_t>(blockIdx.z) * ((p.A + 1) * p.d + p.d * 8);
  if (y == 0) {  // goal token (t = 1): pos row 0
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) {
      const float g = p.dX[(static_cast<size_t>(b) * p.T + 1) * p.d + col];
      acc += g;
      p.dgoal[static_cast<size_t>(b) * p.d + col] = __float2bfloat16_rn(g);
    }
    mine[col] = acc;
  } else if (y <= p.A) {  // pos row y: action j = y - 1, plus the image tokens for y == 1
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) {
      acc += p.dX[(static_cast<size_t>(b) * p.T + 2 + p.S + (y - 1)) * p.d + col];
      if (y == 1)
        for (int s = 0; s < p.S; ++s) {
          const float g = p.dX[(static_cast<size_t>(b) * p.T + 2 + s) * p.d + col];
          acc += g;
          p.dstate[(static_cast<size_t>(b) * p.S + s) * p.d + col] = __float2bfloat16_rn(g);
        }
    }
    mine[static_cast<size_t>(y) * p.d + col] = acc;
  } else {  // action_emb.weight[col, a] = sum_{b, j} dX[b, 2+S+j, col] * (c_in * noised[b, j, a])
    float acc[8];
    for (int a = 0; a < 8; ++a) acc[a] = 0.f;
    for (int b = b0; b < b1; ++b) {
      const float sigma = load_sigma(p.sc, b);
      const float c_in = 1.0f / sqrtf(sigma * sigma + p.sc.sigma_data * p.sc.sigma_data);
      for (int j = 0; j < p.A; ++j) {
        const float g = p.dX[(static_cast<size_t>(b) * p.T + 2 + p.S + j) * p.d + col] * c_in;
        const float* a_in = p.noised + (static_cast<size_t>(b) * p.A + j) * p.adim;
        for (int a = 0; a < p.adim; ++a) acc[a] = fmaf(g, a_in[a], acc[a]);
      }
    }
    for (int a = 0; a < 8; ++a) mine[static_cast<size_t>(p.A + 1) * p.d + static_cast<size_t>(col) * 8 + a] = acc[a];
  }
}
// grid (d / 256, A + 2): sums the chunks and writes pos_emb / action_emb.weight gradients
__global__ void __launch_bounds__(256) embed_bwd_finish_kernel(const float* __restrict__ partial, float* __restrict__ g_pos,
                                                               float* __restrict__ g_wact, int A, int adim, int d) {
  pdl_trigger();
  pdl_wait();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= d) return;
  const int y = blockIdx.y;
  const size_t stride = static_cast<size_t>(A + 1) * d + static_cast<size_t>(d) * 8;
  if (y <= A) {
    float acc = 0.f;
    for (int c = 0; c < EMBED_BWD_CHUNKS; ++c) acc += partial[c * stride + static_cast<size_t>(y) * d + col];
    g_pos[static_cast<size_t>(y) * d + col] = acc;
  } else {
    for (int a = 0; a < adim; ++a) {
      float acc = 0.f;
      for (int c = 0; c < EMBED_BWD_CHUNKS; ++c)
        acc += partial[c * stride + static_cast<size_t>(A + 1) * d + static_cast<size_t>(col) * 8 + a];
      g_wact[static_cast<size_t>(col) * adim + a] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Sigma-embedding backward: c = W2 e1, e1 = s w1 + b1 (modedit.py:823-832).  de1 = dc W2 (fp32 SIMT, [B,d]x[d,d]);
// g_w1 = sum_b de1 s_b; g_b1 = sum_b de1. Also writes e1 and dc as bf16 for the W2 weight gradient (dc^T e1).
struct SigmaBwdParams {
  StepScalars sc;
  const float* dc;          // [B, d]
  const float* w1;          // [d]
  const float* b1;          // [d]
  const float* w2;          // [d, d] (out, in)
  float* g_w1;              // [d]
  float* g_b1;              // [d]
  __nv_bfloat16* dc_bf16;   // [B (padded), d]
  __nv_bfloat16* e1_bf16;   // [B (padded), d]
  int B, d;
};
// grid (d / 256, B): de1[b, i] = sum_o dc[b, o] * W2[o, i]  (+ the bf16 operands of the W2 weight gradient)
__global__ void __launch_bounds__(256) sigma_bwd_kernel(const SigmaBwdParams p, float* __restrict__ de1) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;  // input feature of sigma_linear
  const int b = blockIdx.y;
  if (i >= p.d) return;
  const float s = logf(load_sigma(p.sc, b)) / 4.0f;
  float acc = 0.f;
  for (int o = 0; o < p.d; ++o) acc = fmaf(p.dc[static_cast<size_t>(b) * p.d + o], p.w2[static_cast<size_t>(o) * p.d + i], acc);
  de1[static_cast<size_t>(b) * p.d + i] = acc;
  p.dc_bf16[static_cast<size_t>(b) * p.d + i] = __float2bfloat16_rn(p.dc[static_cast<size_t>(b) * p.d + i]);
  p.e1_bf16[static_cast<size_t>(b) * p.d + i] = __float2bfloat16_rn(fmaf(s, p.w1[i], p.b1[i]));
}
// g_w1[i] = sum_b de1[b, i] * s_b ; g_b1[i] = sum_b de1[b, i]   (fixed order over b)
__global__ void __launch_bounds__(256) sigma_bwd_finish_kernel(const SigmaBwdParams p, const float* __restrict__ de1) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= p.d) return;
  float gw = 0.f, gb = 0.f;
  for (int b = 0; b < p.B; ++b) {
    const float v = de1[static_cast<size_t>(b) * p.d + i];
    gw = fmaf(v, logf(load_sigma(p.sc, b)) / 4.0f, gw);
    gb += v;
  }
  p.g_w1[i] = gw;
  p.g_b1[i] = gb;
}

// [rows, cols] (bf16 or fp32) -> bf16 [cols, rows]: weight transposes for the data-gradient GEMMs, once per weight update
template <typename SrcT>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const SrcT* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const size_t mat = static_cast<size_t>(blockIdx.z) * rows * cols;
  for (int r = ty; r < 32; r += 8)
    if (by + r < rows && bx + tx < cols) tile[r][tx] = __nv_bfloat16(src[mat + static_cast<size_t>(by + r) * cols + bx + tx]);
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bx + r < cols && by + tx < rows) dst[mat + static_cast<size_t>(bx + r) * rows + by + tx] = tile[tx][r];
}

// bf16 [rows, cols] -> bf16 [cols, rows] for rows, cols multiples of 64: 64 x 64 tiles, 16-byte global accesses on both
// sides (a source row segment and a destination row segment are each 128 contiguous bytes = 8 lanes), staged through a
// 33-word-stride shared tile (conflict-free 4-byte writes, <= 2-way conflicts on the transposed reads).
__global__ void __launch_bounds__(256) transpose_bf16_64_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                int rows, int cols) {
  __shared__ uint32_t tile[64][33];  // tile[r][w]: source row r, bf16 columns 2w and 2w+1
  const int bx = blockIdx.x * 64, by = blockIdx.y * 64;
  const size_t mat = static_cast<size_t>(blockIdx.z) * rows * cols;
  const int t = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = pass * 32 + (t >> 3), c8 = t & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(src + mat + static_cast<size_t>(by + r) * cols + bx + c8 * 8);
    tile[r][c8 * 4 + 0] = v.x;
    tile[r][c8 * 4 + 1] = v.y;
    tile[r][c8 * 4 + 2] = v.z;
    tile[r][c8 * 4 + 3] = v.w;
  }
  __syncthreads();
  const __nv_bfloat16* th = reinterpret_cast<const __nv_bfloat16*>(&tile[0][0]);
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int c = pass * 32 + (t >> 3), seg = t & 7;  // destination row bx + c, its source rows seg*8 .. seg*8+7
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint16_t lo = reinterpret_cast<const uint16_t*>(th)[(seg * 8 + 2 * k) * 66 + c];
      const uint16_t hi = reinterpret_cast<const uint16_t*>(th)[(seg * 8 + 2 * k + 1) * 66 + c];
      o[k] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
    *reinterpret_cast<uint4*>(dst + mat + static_cast<size_t>(bx + c) * rows + by + seg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Generic weight gradient out[n, k] = sum_r dy[r, n] * x[r, k] for shapes the tensor-core kernel does not tile
// (embedding widths that are not multiples of 256, used by small test configurations).
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                         float* __restrict__ out, int rows, int n_out, int k_out) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (idx >= static_cast<size_t>(n_out) * k_out) return;
  const int n = static_cast<int>(idx / k_out), k = static_cast<int>(idx % k_out);
  float acc = 0.f;
  for (int r = 0; r < rows; ++r)
    acc = fmaf(__bfloat162float(dy[static_cast<size_t>(r) * n_out + n]), __bfloat162float(x[static_cast<size_t>(r) * k_out + k]), acc);
  out[idx] = acc;
}

// Generic data gradient dx[r, i] = sum_o dy[r, o] * w[o, i] (w = nn.Linear weight [d_out, n_in], bf16) for input widths
// the tensor-core kernel does not tile (small test configurations).
__global__ void __launch_bounds__(256) dgrad_simt_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ w,
                                                         float* __restrict__ dx, int rows, int d_out, int n_in) {
  pdl_trigger();
  pdl_wait();
  const size_t idx = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (idx >= static_cast<size_t>(rows) * n_in) return;
  const int r = static_cast<int>(idx / n_in), i = static_cast<int>(idx % n_in);
  float acc = 0.f;
  for (int o = 0; o < d_out; ++o)
    acc = fmaf(__bfloat162float(dy[static_cast<size_t>(r) * d_out + o]), __bfloat162float(w[static_cast<size_t>(o) * n_in + i]), acc);
  dx[idx] = acc;
}

}  // namespace mode
