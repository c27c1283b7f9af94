// Data-movement kernels of the FiLM-ResNet-50 token producer (reference mode/models/perceptual_encoders/
// pretrained_resnets.py:25-60, called from MoDEAgent.embed_visual_obs, mode_agent.py:548-567).
//
// Every convolution runs as a GEMM on the tcgen05 CTA-pair kernel (gemm.cuh, EPI_CONV_BF16: folded-BatchNorm bias,
// shortcut add, ReLU and FiLM in the epilogue). Activations are NHWC bf16, i.e. row-major [images*H*W, C] matrices, so a
// 1x1 stride-1 convolution needs no data movement at all; the kernels here build the A operand of the others
// (im2col rows in (kh, kw, c) order, 16 bytes = 8 channels per thread), fold BatchNorm into the packed weights, and do
// the two pooling layers and the FiLM coefficient Linear.
#pragma once
#include "ptx.cuh"

namespace mode {

// conv1: fp32 NCHW images -> im2col rows [images*Ho*Wo, Kpad] bf16, k = (kh*KW + kw)*C + c, zero beyond KH*KW*C.
__global__ void im2col_nchw_f32_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ col, int n_img, int C, int H,
                                       int W, int Ho, int Wo, int KH, int KW, int stride, int pad, int Kpad) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const size_t total = static_cast<size_t>(n_img) * Ho * Wo * Kpad;
  if (i >= total) return;
  const int k = static_cast<int>(i % Kpad);
  const size_t row = i / Kpad;
  float v = 0.f;
  if (k < KH * KW * C) {
    const int c = k % C, tap = k / C, kw = tap % KW, kh = tap / KW;
    const int wo = static_cast<int>(row % Wo), ho = static_cast<int>((row / Wo) % Ho), n = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    const int h = ho * stride - pad + kh, w = wo * stride - pad + kw;
    if (h >= 0 && h < H && w >= 0 && w < W) v = img[((static_cast<size_t>(n) * C + c) * H + h) * W + w];
  }
  col[i] = __float2bfloat16_rn(v);
}

// conv1 specialisation (7x7, stride 2, pad 3, 3 channels): one CTA per output row (image n, row ho). The 7 input rows of
// the 3 channels are staged in shared memory as bf16 with coalesced reads, then the Wo x Kpad im2col rows are written as
// whole 16-byte vectors (8 consecutive k = (kh*7 + kw)*3 + c): the generic kernel above spends 4.6 ms on 256 images of
// 224x224 on per-element index arithmetic and strided reads, this one is bound by its 1.2 GB of stores.
__global__ void __launch_bounds__(256) im2col_conv1_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ col, int H,
                                                           int W, int Ho, int Wo, int Kpad) {
  extern __shared__ __nv_bfloat16 rows_s[];  // [3][7][W + 6]
  const int n = blockIdx.x / Ho, ho = blockIdx.x % Ho;
  const int Wp = W + 6;
  for (int i = threadIdx.x; i < 21 * Wp; i += 256) {
    const int x = i % Wp - 3, ck = i / Wp, kh = ck % 7, c = ck / 7;
    const int h = ho * 2 - 3 + kh;
    float v = 0.f;
    if (x >= 0 && x < W && h >= 0 && h < H) v = img[((static_cast<size_t>(n) * 3 + c) * H + h) * W + x];
    rows_s[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int vec_per_row = Kpad >> 3;
  uint4* dst = reinterpret_cast<uint4*>(col + (static_cast<size_t>(n) * Ho + ho) * Wo * Kpad);
  for (int i = threadIdx.x; i < Wo * vec_per_row; i += 256) {
    const int wo = i / vec_per_row, k0 = (i % vec_per_row) * 8;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u;
      __nv_bfloat16 e = __float2bfloat16_rn(0.f);
      if (k < 147) {
        const int c = k % 3, tap = k / 3, kw = tap % 7, kh = tap / 7;
        e = rows_s[(c * 7 + kh) * Wp + wo * 2 + kw];
      }
      v[u] = e;
    }
    dst[i] = *reinterpret_cast<const uint4*>(v);
  }
}

// NHWC bf16 activation -> im2col rows [images*Ho*Wo, KH*KW*C] (C % 8 == 0): one 16-byte vector (8 channels) per thread.
// Also used with KH = KW = 1, stride 2 for the strided 1x1 projection shortcuts.
__global__ void im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ col, int n_img, int C,
                                   int H, int W, int Ho, int Wo, int KH, int KW, int stride, int pad) {
  // 32-bit index arithmetic: the host checks that the vector count fits (it is 58 M for 256 images of 224x224)
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c8n = C >> 3, taps = KH * KW;
  const unsigned total = static_cast<unsigned>(n_img) * Ho * Wo * taps * c8n;
  if (i >= total) return;
  const unsigned c8 = i % c8n, t = i / c8n, tap = t % taps, row = t / taps;
  const int kw = tap % KW, kh = tap / KW;
  const int wo = row % Wo, ho = (row / Wo) % Ho, n = row / (static_cast<unsigned>(Wo) * Ho);
  const int h = ho * stride - pad + kh, w = wo * stride - pad + kw;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (h >= 0 && h < H && w >= 0 && w < W)
    v = __ldg(reinterpret_cast<const uint4*>(act + ((static_cast<size_t>(n) * H + h) * W + w) * C) + c8);
  reinterpret_cast<uint4*>(col)[i] = v;
}

// MaxPool2d(3, stride 2, padding 1) on NHWC bf16 (timm / torchvision resnet stem), 8 channels per thread.
__global__ void maxpool3x3s2_nhwc_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n_img, int C,
                                         int H, int W, int Ho, int Wo) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int c8n = C >> 3;
  const size_t total = static_cast<size_t>(n_img) * Ho * Wo * c8n;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % c8n);
  const size_t row = i / c8n;
  const int wo = static_cast<int>(row % Wo), ho = static_cast<int>((row / Wo) % Ho), n = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
  float m[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) m[u] = -INFINITY;
  for (int dh = 0; dh < 3; ++dh) {
    const int h = ho * 2 - 1 + dh;
    if (h < 0 || h >= H) continue;
    for (int dw = 0; dw < 3; ++dw) {
      const int w = wo * 2 - 1 + dw;
      if (w < 0 || w >= W) continue;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(n) * H + h) * W + w) * C) + c8);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __bfloat1622float2(h2[u]);
        m[2 * u] = fmaxf(m[2 * u], f.x);
        m[2 * u + 1] = fmaxf(m[2 * u + 1], f.y);
      }
    }
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]),
                                                pack_bf16x2(m[6], m[7]));
}

// ---- zero-bordered layout (implicit 3x3 convolutions): an image of H x W positions is stored as (H + 2) x (W + 2)
// positions whose one-pixel border is zero, so the tap (kh, kw) of a stride-1 3x3 convolution is the row offset
// (kh - 1) * (W + 2) + (kw - 1) of the [positions, C] matrix and the GEMM reads it straight from the activation.

// MaxPool2d(3, 2, 1): dense NHWC in -> zero-bordered NHWC out.
__global__ void maxpool3x3s2_to_padded_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n_img,
                                              int C, int H, int W, int Ho, int Wo) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c8n = C >> 3, Wp = Wo + 2, Hp = Ho + 2;
  if (i >= static_cast<unsigned>(n_img) * Hp * Wp * c8n) return;
  const unsigned c8 = i % c8n, row = i / c8n;
  const int xo = row % Wp, yo = (row / Wp) % Hp, n = row / (Wp * Hp);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (xo >= 1 && xo <= Wo && yo >= 1 && yo <= Ho) {
    const int wo = xo - 1, ho = yo - 1;
    float m[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = -INFINITY;
    for (int dh = 0; dh < 3; ++dh) {
      const int h = ho * 2 - 1 + dh;
      if (h < 0 || h >= H) continue;
      for (int dw = 0; dw < 3; ++dw) {
        const int w = wo * 2 - 1 + dw;
        if (w < 0 || w >= W) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(n) * H + h) * W + w) * C) + c8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = __bfloat1622float2(h2[u]);
          m[2 * u] = fmaxf(m[2 * u], f.x);
          m[2 * u + 1] = fmaxf(m[2 * u + 1], f.y);
        }
      }
    }
    r = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
  }
  reinterpret_cast<uint4*>(out)[i] = r;
}

// im2col of a STRIDED convolution (the 3x3/2 and the 1x1/2 projection of a stage's first block) from a zero-bordered
// input to rows in zero-bordered OUTPUT order: row (n, yo, xo) of the (Ho + 2) x (Wo + 2) grid; border rows are zero. The
// input border supplies the convolution's padding, so no bounds checks are needed (pad <= 1).
__global__ void im2col_padded_kernel(const __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ col, int n_img, int C,
                                     int H, int W, int Ho, int Wo, int KH, int KW, int stride, int pad) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned c8n = C >> 3, taps = KH * KW, Wpo = Wo + 2, Hpo = Ho + 2;
  if (i >= static_cast<unsigned>(n_img) * Hpo * Wpo * taps * c8n) return;
  const unsigned c8 = i % c8n, t = i / c8n, tap = t % taps, row = t / taps;
  const int kw = tap % KW, kh = tap / KW;
  const int xo = row % Wpo, yo = (row / Wpo) % Hpo, n = row / (Wpo * Hpo);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (xo >= 1 && xo <= Wo && yo >= 1 && yo <= Ho) {
    const int h = (yo - 1) * stride - pad + kh + 1, w = (xo - 1) * stride - pad + kw + 1;  // +1: position in the bordered input
    v = __ldg(reinterpret_cast<const uint4*>(act + ((static_cast<size_t>(n) * (H + 2) + h) * (W + 2) + w) * C) + c8);
  }
  reinterpret_cast<uint4*>(col)[i] = v;
}

// Global average pool over the interior of a zero-bordered map (the border is zero: summing everything is the same).
__global__ void avgpool_padded_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int n_img, int H, int W, int C) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int c2n = C >> 1;
  if (i >= static_cast<size_t>(n_img) * c2n) return;
  const int c2 = static_cast<int>(i % c2n), n = static_cast<int>(i / c2n);
  const int rows = (H + 2) * (W + 2);
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(in + static_cast<size_t>(n) * rows * C) + c2;
  float a = 0.f, b = 0.f;
  for (int y = 1; y <= H; ++y)
    for (int x = 1; x <= W; ++x) {
      const float2 f = __bfloat1622float2(p[static_cast<size_t>(y * (W + 2) + x) * c2n]);
      a += f.x;
      b += f.y;
    }
  const float inv = 1.0f / static_cast<float>(H * W);
  out[static_cast<size_t>(n) * C + 2 * c2] = a * inv;
  out[static_cast<size_t>(n) * C + 2 * c2 + 1] = b * inv;
}

// Global average pool: NHWC bf16 [images, hw, C] -> fp32 [images, C] (timm global_pool, pretrained_resnets.py:57-58).
// One thread per (image, channel pair); consecutive threads read consecutive channels: coalesced.
__global__ void avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int n_img, int hw, int C) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int c2n = C >> 1;
  if (i >= static_cast<size_t>(n_img) * c2n) return;
  const int c2 = static_cast<int>(i % c2n), n = static_cast<int>(i / c2n);
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(in + static_cast<size_t>(n) * hw * C) + c2;
  float a = 0.f, b = 0.f;
  for (int s = 0; s < hw; ++s) {
    const float2 f = __bfloat1622float2(p[static_cast<size_t>(s) * c2n]);
    a += f.x;
    b += f.y;
  }
  const float inv = 1.0f / static_cast<float>(hw);
  out[static_cast<size_t>(n) * C + 2 * c2] = a * inv;
  out[static_cast<size_t>(n) * C + 2 * c2 + 1] = b * inv;
}

// FiLM coefficients: gamma[n, c] = Wg[c, :] . cond[n, :] + bg[c], beta likewise (FiLMLayer.forward,
// pretrained_resnets.py:19-21), fp32, one warp per (n, c).
__global__ void film_linear_kernel(const float* __restrict__ cond, const float* __restrict__ wg, const float* __restrict__ bg,
                                   const float* __restrict__ wb, const float* __restrict__ bb, float* __restrict__ gamma,
                                   float* __restrict__ beta, int n_img, int C, int cond_dim) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= n_img * C) return;
  const int n = item / C, c = item % C;
  float a = 0.f, b = 0.f;
  for (int k = lane; k < cond_dim; k += 32) {
    const float x = cond[static_cast<size_t>(n) * cond_dim + k];
    a = fmaf(wg[static_cast<size_t>(c) * cond_dim + k], x, a);
    b = fmaf(wb[static_cast<size_t>(c) * cond_dim + k], x, b);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    gamma[item] = a + bg[c];
    beta[item] = b + bb[c];
  }
}

// Inference BatchNorm folded into the convolution that precedes it, and the weight repacked for the GEMM:
//   w'[co, (kh, kw, ci)] = w[co, ci, kh, kw] * g[co] / sqrt(var[co] + eps)   (bf16, K padded with zeros to Kpad)
//   b'[co] = beta[co] - mean[co] * g[co] / sqrt(var[co] + eps)              (fp32)
__global__ void fold_bn_pack_kernel(const float* __restrict__ w, const float* __restrict__ g, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                    __nv_bfloat16* __restrict__ wp, float* __restrict__ bp, int Cout, int Cin, int KH, int KW,
                                    int Kpad) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<size_t>(Cout) * Kpad) return;
  const int k = static_cast<int>(i % Kpad), co = static_cast<int>(i / Kpad);
  const float s = g[co] / sqrtf(var[co] + eps);
  float v = 0.f;
  if (k < KH * KW * Cin) {
    const int ci = k % Cin, tap = k / Cin, kw = tap % KW, kh = tap / KW;
    v = w[((static_cast<size_t>(co) * Cin + ci) * KH + kh) * KW + kw] * s;
  }
  wp[i] = __float2bfloat16_rn(v);
  if (k == 0) bp[co] = beta[co] - mean[co] * s;
}

// Dense M-tile table of a [rows, .] GEMM: tile i covers rows [i*tile_m, min(rows, (i+1)*tile_m)).
__global__ void fill_dense_tiles_kernel(GemmMTile* __restrict__ tiles, int* __restrict__ count, int rows, int tile_m) {
  const int n = (rows + tile_m - 1) / tile_m;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *count = n;
  if (i < n) tiles[i] = GemmMTile{i * tile_m, i * tile_m, min(tile_m, rows - i * tile_m), 0};
}

}  // namespace mode
