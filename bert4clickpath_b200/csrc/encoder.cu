// Encoder-layer kernels that are not GEMMs (the Dense layers go through gemm_umma.cu):
//   - fused short-sequence masked self-attention, forward and backward
//     (transformer.py:64-97 scaled_dot_product_attention, :130-156 split/merge heads);
//   - residual + dropout + LayerNormalization(eps=1e-6), forward and backward
//     (transformer.py:202-213, LayerNormalization non-fused path: biased variance);
//   - column sums (bias gradients) and partial-sum reductions.
// Sequences are short (S <= 256): one CTA owns one (sequence, head); Q/K/V/dO live in shared
// memory as bf16 with an odd word stride, scores never touch HBM.
#include <algorithm>

#include <cstdlib>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

static constexpr int MAX_S = 256;

// tensor-core path for S <= 128 and head depth 32 / 64 (attention_mma.cu)
bool attention_mma_supported(int S, int dh);
int attention_mma_fwd(const void* qkv, const int32_t* ids, int B, int S, int H, int dh, void* out,
                      float* lse, cudaStream_t st);
int attention_mma_bwd(const void* qkv, const void* fwd_out, const void* dout, const float* lse,
                      const int32_t* ids, int B,
                      int S, int H, int dh, void* dqkv, cudaStream_t st);

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

// cooperative load of one head's [S x dh] slice (row stride ld in global) into smem [S][dh+2]
__device__ __forceinline__ void load_head_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int S,
                                               int dh, long ld) {
  const int words = dh / 2;
  for (int i = threadIdx.x; i < S * words; i += blockDim.x) {
    const int r = i / words, w = i - r * words;
    reinterpret_cast<uint32_t*>(dst + (size_t)r * (dh + 2))[w] =
        reinterpret_cast<const uint32_t*>(src + (size_t)r * ld)[w];
  }
}

// ------------------------------------------------------------------------- attention forward
// grid = B*H, block = 128 or 256.  qkv: [T][3d] bf16 (q | k | v), out: [T][d] bf16, lse: [B][H][S].
__global__ void __launch_bounds__(256)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ ids,
                     int S, int H, int dh, __nv_bfloat16* __restrict__ out,
                     float* __restrict__ lse_out) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int d = H * dh;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int st = dh + 2;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sm);
  __nv_bfloat16* sK = sQ + (size_t)S * st;
  __nv_bfloat16* sV = sK + (size_t)S * st;
  float* sPad = reinterpret_cast<float*>(sV + (size_t)S * st);
  const int nwarps = blockDim.x >> 5;
  float* sP = sPad + S;  // [nwarps][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * d + h * dh;
  load_head_tile(sQ, base, S, dh, 3L * d);
  load_head_tile(sK, base + d, S, dh, 3L * d);
  load_head_tile(sV, base + 2 * d, S, dh, 3L * d);
  for (int j = threadIdx.x; j < S; j += blockDim.x)
    sPad[j] = (ids[(size_t)b * S + j] == 0) ? -1e9f : 0.f;  // create_padding_mask * -1e9
  __syncthreads();
  const float sqrt_dh = sqrtf((float)dh);
  float* myP = sP + (size_t)warp * S;
  for (int i = warp; i < S; i += nwarps) {
    float z[MAX_S / 32];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      z[t] = -INFINITY;
      if (j < S) {
        float acc = 0.f;
        const __nv_bfloat162* qr = reinterpret_cast<const __nv_bfloat162*>(sQ + (size_t)i * st);
        const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(sK + (size_t)j * st);
        for (int c = 0; c < dh / 2; ++c) {
          const float2 qv = __bfloat1622float2(qr[c]);
          const float2 kv = __bfloat1622float2(kr[c]);
          acc = fmaf(qv.x, kv.x, acc);
          acc = fmaf(qv.y, kv.y, acc);
        }
        z[t] = __fdiv_rn(acc, sqrt_dh) + sPad[j];
        m = fmaxf(m, z[t]);
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < S) {
        z[t] = expf(z[t] - m);
        sum += z[t];
      }
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int t = 0; t < MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < S) myP[j] = z[t] * inv;
    }
    if (lane == 0 && lse_out) lse_out[((size_t)b * H + h) * S + i] = m + logf(sum);
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc = fmaf(myP[j], bf2f(sV[(size_t)j * st + c]), acc);
      out[((size_t)b * S + i) * d + h * dh + c] = __float2bfloat16_rn(acc);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------- attention backward
// dqkv: [T][3d] bf16.  Phase 1 (warp per query row): delta_i and dQ_i.  Phase 2 (warp per key
// row): dK_j, dV_j.  Probabilities are recomputed from the saved log-sum-exp.
__global__ void __launch_bounds__(256)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ lse_in, const int32_t* __restrict__ ids, int S,
                     int H, int dh, __nv_bfloat16* __restrict__ dqkv) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int d = H * dh;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int st = dh + 2;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sm);
  __nv_bfloat16* sK = sQ + (size_t)S * st;
  __nv_bfloat16* sV = sK + (size_t)S * st;
  __nv_bfloat16* sDO = sV + (size_t)S * st;
  float* sPad = reinterpret_cast<float*>(sDO + (size_t)S * st);
  float* sLse = sPad + S;
  float* sDelta = sLse + S;
  const int nwarps = blockDim.x >> 5;
  float* sP = sDelta + S;             // [nwarps][S]
  float* sDZ = sP + (size_t)nwarps * S;  // [nwarps][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * d + h * dh;
  load_head_tile(sQ, base, S, dh, 3L * d);
  load_head_tile(sK, base + d, S, dh, 3L * d);
  load_head_tile(sV, base + 2 * d, S, dh, 3L * d);
  load_head_tile(sDO, dout + (size_t)b * S * d + h * dh, S, dh, (long)d);
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    sPad[j] = (ids[(size_t)b * S + j] == 0) ? -1e9f : 0.f;
    sLse[j] = lse_in[((size_t)b * H + h) * S + j];
  }
  __syncthreads();
  const float sqrt_dh = sqrtf((float)dh);
  const float inv_sqrt = 1.f / sqrt_dh;
  float* myP = sP + (size_t)warp * S;
  float* myDZ = sDZ + (size_t)warp * S;
  __nv_bfloat16* dq_out = dqkv + (size_t)b * S * 3 * d + h * dh;

  // ---- phase 1: rows of the score matrix
  for (int i = warp; i < S; i += nwarps) {
    float p[MAX_S / 32], da[MAX_S / 32];
    float delta = 0.f;
    const float lse_i = sLse[i];
#pragma unroll
    for (int t = 0; t < MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      p[t] = 0.f;
      da[t] = 0.f;
      if (j < S) {
        float acc = 0.f, acc2 = 0.f;
        const __nv_bfloat162* qr = reinterpret_cast<const __nv_bfloat162*>(sQ + (size_t)i * st);
        const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(sK + (size_t)j * st);
        const __nv_bfloat162* gr = reinterpret_cast<const __nv_bfloat162*>(sDO + (size_t)i * st);
        const __nv_bfloat162* vr = reinterpret_cast<const __nv_bfloat162*>(sV + (size_t)j * st);
        for (int c = 0; c < dh / 2; ++c) {
          const float2 qv = __bfloat1622float2(qr[c]), kv = __bfloat1622float2(kr[c]);
          const float2 gv = __bfloat1622float2(gr[c]), vv = __bfloat1622float2(vr[c]);
          acc = fmaf(qv.x, kv.x, acc);
          acc = fmaf(qv.y, kv.y, acc);
          acc2 = fmaf(gv.x, vv.x, acc2);
          acc2 = fmaf(gv.y, vv.y, acc2);
        }
        p[t] = expf(__fdiv_rn(acc, sqrt_dh) + sPad[j] - lse_i);
        da[t] = acc2;
        delta = fmaf(p[t], da[t], delta);
      }
    }
    delta = warp_sum(delta);
    if (lane == 0) sDelta[i] = delta;
#pragma unroll
    for (int t = 0; t < MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < S) myDZ[j] = p[t] * (da[t] - delta);
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc = fmaf(myDZ[j], bf2f(sK[(size_t)j * st + c]), acc);
      dq_out[(size_t)i * 3 * d + c] = __float2bfloat16_rn(acc * inv_sqrt);
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- phase 2: columns of the score matrix
  for (int j = warp; j < S; j += nwarps) {
    const bool key_is_pad = sPad[j] != 0.f;  // uniform per warp
    if (!key_is_pad) {
#pragma unroll
      for (int t = 0; t < MAX_S / 32; ++t) {
        const int i = lane + 32 * t;
        if (i < S) {
          float acc = 0.f, acc2 = 0.f;
          const __nv_bfloat162* qr = reinterpret_cast<const __nv_bfloat162*>(sQ + (size_t)i * st);
          const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(sK + (size_t)j * st);
          const __nv_bfloat162* gr = reinterpret_cast<const __nv_bfloat162*>(sDO + (size_t)i * st);
          const __nv_bfloat162* vr = reinterpret_cast<const __nv_bfloat162*>(sV + (size_t)j * st);
          for (int c = 0; c < dh / 2; ++c) {
            const float2 qv = __bfloat1622float2(qr[c]), kv = __bfloat1622float2(kr[c]);
            const float2 gv = __bfloat1622float2(gr[c]), vv = __bfloat1622float2(vr[c]);
            acc = fmaf(qv.x, kv.x, acc);
            acc = fmaf(qv.y, kv.y, acc);
            acc2 = fmaf(gv.x, vv.x, acc2);
            acc2 = fmaf(gv.y, vv.y, acc2);
          }
          const float pij = expf(__fdiv_rn(acc, sqrt_dh) - sLse[i]);
          myP[i] = pij;
          myDZ[i] = pij * (acc2 - sDelta[i]);
        }
      }
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float dk = 0.f, dv = 0.f;
      if (!key_is_pad) {
        for (int i = 0; i < S; ++i) {
          dv = fmaf(myP[i], bf2f(sDO[(size_t)i * st + c]), dv);
          dk = fmaf(myDZ[i], bf2f(sQ[(size_t)i * st + c]), dk);
        }
      }
      dq_out[(size_t)j * 3 * d + d + c] = __float2bfloat16_rn(dk * inv_sqrt);
      dq_out[(size_t)j * 3 * d + 2 * d + c] = __float2bfloat16_rn(dv);
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------- residual + dropout + LayerNorm
// y = LN(x + dropout(r)) * gamma + beta ; one warp per row, d <= 256.
static constexpr int LN_MAX_PER_LANE = 8;  // d <= 256; kernels are templated on ceil(d/32)

struct DropSpec {
  float inv_keep;
  uint32_t thresh24;
  uint64_t seed;
  uint32_t site;
};

template <int NPL>
__global__ void __launch_bounds__(256)
residual_ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r, long T, int d,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       DropSpec dp, float eps, float* __restrict__ y_f32,
                       __nv_bfloat16* __restrict__ y_bf16, long ld_bf16) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T) return;
  if (dp.thresh24) dp.seed = resolve_seed(dp.seed);
  float v[NPL];
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < NPL; ++t) {
    const int c = lane + 32 * t;
    v[t] = 0.f;
    if (c < d) {
      float rv = r[row * d + c];
      if (dp.thresh24)
        rv = dropout_keep(dp.seed, dp.site, (uint64_t)(row * d + c), dp.thresh24) ? rv * dp.inv_keep : 0.f;
      v[t] = x[row * d + c] + rv;
      sum += v[t];
    }
  }
  const float mean = warp_sum(sum) / (float)d;
  float var = 0.f;
#pragma unroll
  for (int t = 0; t < NPL; ++t) {
    const int c = lane + 32 * t;
    if (c < d) {
      const float dv = v[t] - mean;
      var = fmaf(dv, dv, var);
    }
  }
  var = warp_sum(var) / (float)d;
  const float rstd = rsqrtf(var + eps);
#pragma unroll
  for (int t = 0; t < NPL; ++t) {
    const int c = lane + 32 * t;
    if (c < d) {
      const float o = (v[t] - mean) * rstd * gamma[c] + beta[c];
      if (y_f32) y_f32[row * d + c] = o;
      if (y_bf16) y_bf16[row * ld_bf16 + c] = __float2bfloat16_rn(o);
    }
  }
}

// Backward.  Outputs: dx (fp32, the residual branch), dr (bf16 and/or fp32: gradient w.r.t. the
// un-dropped Dense output r), and per-block partial column sums [gridDim.x][3][d] of
// (dy*xhat, dy, dr) = (dgamma, dbeta, bias gradient of the Dense that produced r).
template <int NPL>
__global__ void __launch_bounds__(256)
residual_ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                       const float* __restrict__ r, long T, int d,
                       const float* __restrict__ gamma, DropSpec dp, float eps,
                       float* __restrict__ dx, __nv_bfloat16* __restrict__ dr_bf16, long ld_bf16,
                       float* __restrict__ dr_f32, float* __restrict__ partial) {
  __shared__ float red[8][3][NPL * 32];
  if (dp.thresh24) dp.seed = resolve_seed(dp.seed);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  float pg[NPL], pb[NPL], pr[NPL];
#pragma unroll
  for (int t = 0; t < NPL; ++t) pg[t] = pb[t] = pr[t] = 0.f;
  for (long row = (long)blockIdx.x * nwarps + warp; row < T; row += (long)gridDim.x * nwarps) {
    float v[NPL], g[NPL], keepf[NPL];
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int c = lane + 32 * t;
      v[t] = 0.f;
      keepf[t] = 0.f;
      if (c < d) {
        float rv = r[row * d + c];
        keepf[t] = 1.f;
        if (dp.thresh24) {
          const bool keep = dropout_keep(dp.seed, dp.site, (uint64_t)(row * d + c), dp.thresh24);
          keepf[t] = keep ? dp.inv_keep : 0.f;
          rv *= keepf[t];
        }
        v[t] = x[row * d + c] + rv;
        sum += v[t];
      }
    }
    const float mean = warp_sum(sum) / (float)d;
    float var = 0.f;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int c = lane + 32 * t;
      if (c < d) {
        const float dv = v[t] - mean;
        var = fmaf(dv, dv, var);
      }
    }
    var = warp_sum(var) / (float)d;
    const float rstd = rsqrtf(var + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int c = lane + 32 * t;
      g[t] = 0.f;
      if (c < d) {
        const float dyv = dy[row * d + c];
        v[t] = (v[t] - mean) * rstd;  // xhat
        g[t] = dyv * gamma[c];
        m1 += g[t];
        m2 = fmaf(g[t], v[t], m2);
        pg[t] = fmaf(dyv, v[t], pg[t]);
        pb[t] += dyv;
      }
    }
    m1 = warp_sum(m1) / (float)d;
    m2 = warp_sum(m2) / (float)d;
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int c = lane + 32 * t;
      if (c < d) {
        const float dres = rstd * (g[t] - m1 - v[t] * m2);
        const float drv = dres * keepf[t];
        if (dx) dx[row * d + c] = dres;
        if (dr_bf16) dr_bf16[row * ld_bf16 + c] = __float2bfloat16_rn(drv);
        if (dr_f32) dr_f32[row * d + c] = drv;
        pr[t] += drv;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < NPL; ++t) {
    red[warp][0][lane + 32 * t] = pg[t];
    red[warp][1][lane + 32 * t] = pb[t];
    red[warp][2][lane + 32 * t] = pr[t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) {
    const int k = i / d, c = i - k * d;
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w][k][c];
    partial[((size_t)blockIdx.x * 3 + k) * d + c] = s;
  }
}

// ---- vectorised variants for d in {64, 128, 256}: 128-bit accesses, d/4 (<= 32) lanes per row so
// a warp covers 32 / LPR rows per pass and the statistics reduce over sub-warps.  Same math and the
// same dropout bits (element index row * d + column) as the generic kernels above.
template <int D>
struct LnVec {
  static constexpr int LPR = (D / 4) < 32 ? (D / 4) : 32;  // lanes per row
  static constexpr int VPL = D / (4 * LPR);                // float4 per lane
  static constexpr int RPW = 32 / LPR;                     // rows per warp
};

template <int LPR>
__device__ __forceinline__ float subwarp_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst) = o;
}

template <int D>
__global__ void __launch_bounds__(256)
residual_ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ r, long T,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           DropSpec dp, float eps, float* __restrict__ y_f32,
                           __nv_bfloat16* __restrict__ y_bf16, long ld_bf16) {
  using C = LnVec<D>;
  const int lane = threadIdx.x & 31;
  const int lr = lane % C::LPR;
  const long row = ((long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * C::RPW + lane / C::LPR;
  const bool live = row < T;  // dead sub-rows still join the shuffles
  if (dp.thresh24) dp.seed = resolve_seed(dp.seed);
  float v[C::VPL][4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < C::VPL; ++i) {
    const int c = (lr + i * C::LPR) * 4;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), rv = xv;
    if (live) {
      xv = *reinterpret_cast<const float4*>(x + row * D + c);
      rv = *reinterpret_cast<const float4*>(r + row * D + c);
    }
    float rr[4] = {rv.x, rv.y, rv.z, rv.w};
    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (dp.thresh24)
        rr[j] = dropout_keep(dp.seed, dp.site, (uint64_t)(row * D + c + j), dp.thresh24) ? rr[j] * dp.inv_keep : 0.f;
      v[i][j] = xx[j] + rr[j];
      sum += v[i][j];
    }
  }
  const float mean = subwarp_sum<C::LPR>(sum) / (float)D;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < C::VPL; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dv = v[i][j] - mean;
      var = fmaf(dv, dv, var);
    }
  var = subwarp_sum<C::LPR>(var) / (float)D;
  const float rstd = rsqrtf(var + eps);
  if (!live) return;
#pragma unroll
  for (int i = 0; i < C::VPL; ++i) {
    const int c = (lr + i * C::LPR) * 4;
    const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 bv = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
    if (y_f32) *reinterpret_cast<float4*>(y_f32 + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (y_bf16) store_bf16x4(y_bf16 + row * ld_bf16 + c, o);
  }
}

template <int D>
__global__ void __launch_bounds__(256)
residual_ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                           const float* __restrict__ r, long T, const float* __restrict__ gamma,
                           DropSpec dp, float eps, float* __restrict__ dx,
                           __nv_bfloat16* __restrict__ dr_bf16, long ld_bf16,
                           float* __restrict__ dr_f32, float* __restrict__ partial) {
  using C = LnVec<D>;
  __shared__ float red[8][3][D];
  if (dp.thresh24) dp.seed = resolve_seed(dp.seed);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int lr = lane % C::LPR, sub = lane / C::LPR;
  float pg[C::VPL][4], pb[C::VPL][4], pr[C::VPL][4], gam[C::VPL][4];
#pragma unroll
  for (int i = 0; i < C::VPL; ++i) {
    const float4 gv = __ldg(reinterpret_cast<const float4*>(gamma + (lr + i * C::LPR) * 4));
    gam[i][0] = gv.x; gam[i][1] = gv.y; gam[i][2] = gv.z; gam[i][3] = gv.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) pg[i][j] = pb[i][j] = pr[i][j] = 0.f;
  }
  const long groups = (T + C::RPW - 1) / C::RPW;
  // The loads of the NEXT row group are issued before this group's arithmetic.  With one group
  // per iteration a warp had 1.5 KB in flight and, at 76 registers (3 CTAs per SM), an SM 36 KB -
  // about half of what 1/148 of the HBM bandwidth needs at DRAM latency, which is where the
  // kernel sat (0.63 of the roofline; the 27-register forward kernel: 64 KB in flight, 0.98).
  const long gstride = (long)gridDim.x * nwarps;
  auto load_group = [&](long grp, float4 (&X)[C::VPL], float4 (&R)[C::VPL], float4 (&DY)[C::VPL]) {
    const long row = grp * C::RPW + sub;
    const bool live = grp < groups && row < T;
#pragma unroll
    for (int i = 0; i < C::VPL; ++i) {
      const int c = (lr + i * C::LPR) * 4;
      X[i] = R[i] = DY[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        X[i] = *reinterpret_cast<const float4*>(x + row * D + c);
        R[i] = *reinterpret_cast<const float4*>(r + row * D + c);
        DY[i] = *reinterpret_cast<const float4*>(dy + row * D + c);
      }
    }
  };
  float4 Xc[C::VPL], Rc[C::VPL], DYc[C::VPL], Xn[C::VPL], Rn[C::VPL], DYn[C::VPL];
  load_group((long)blockIdx.x * nwarps + warp, Xc, Rc, DYc);
  for (long grp = (long)blockIdx.x * nwarps + warp; grp < groups; grp += gstride) {
    load_group(grp + gstride, Xn, Rn, DYn);
    const long row = grp * C::RPW + sub;
    const bool live = row < T;
    float v[C::VPL][4], g[C::VPL][4], keepf[C::VPL][4], dyv[C::VPL][4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < C::VPL; ++i) {
      const int c = (lr + i * C::LPR) * 4;
      const float4 xv = Xc[i], rv = Rc[i], dv = DYc[i];
      const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, rr[4] = {rv.x, rv.y, rv.z, rv.w};
      dyv[i][0] = dv.x; dyv[i][1] = dv.y; dyv[i][2] = dv.z; dyv[i][3] = dv.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        keepf[i][j] = live ? 1.f : 0.f;
        if (dp.thresh24 && live)
          keepf[i][j] = dropout_keep(dp.seed, dp.site, (uint64_t)(row * D + c + j), dp.thresh24) ? dp.inv_keep : 0.f;
        v[i][j] = xx[j] + rr[j] * keepf[i][j];
        sum += v[i][j];
      }
    }
    const float mean = subwarp_sum<C::LPR>(sum) / (float)D;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < C::VPL; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dv = v[i][j] - mean;
        var = fmaf(dv, dv, var);
      }
    var = subwarp_sum<C::LPR>(var) / (float)D;
    const float rstd = rsqrtf(var + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < C::VPL; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i][j] = (v[i][j] - mean) * rstd;  // xhat
        g[i][j] = dyv[i][j] * gam[i][j];
        m1 += g[i][j];
        m2 = fmaf(g[i][j], v[i][j], m2);
        if (live) {
          pg[i][j] = fmaf(dyv[i][j], v[i][j], pg[i][j]);
          pb[i][j] += dyv[i][j];
        }
      }
    m1 = subwarp_sum<C::LPR>(m1) / (float)D;
    m2 = subwarp_sum<C::LPR>(m2) / (float)D;
    if (live) {
#pragma unroll
      for (int i = 0; i < C::VPL; ++i) {
        const int c = (lr + i * C::LPR) * 4;
        float dres[4], drv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dres[j] = rstd * (g[i][j] - m1 - v[i][j] * m2);
          drv[j] = dres[j] * keepf[i][j];
          pr[i][j] += drv[j];
        }
        if (dx) *reinterpret_cast<float4*>(dx + row * D + c) = make_float4(dres[0], dres[1], dres[2], dres[3]);
        if (dr_bf16) store_bf16x4(dr_bf16 + row * ld_bf16 + c, drv);
        if (dr_f32) *reinterpret_cast<float4*>(dr_f32 + row * D + c) = make_float4(drv[0], drv[1], drv[2], drv[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < C::VPL; ++i) {
      Xc[i] = Xn[i];
      Rc[i] = Rn[i];
      DYc[i] = DYn[i];
    }
  }
  // sub-rows of a warp hold the same columns: fold them (fixed order), then one writer per column
#pragma unroll
  for (int i = 0; i < C::VPL; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int o = C::LPR; o < 32; o <<= 1) {
        pg[i][j] += __shfl_xor_sync(0xffffffffu, pg[i][j], o);
        pb[i][j] += __shfl_xor_sync(0xffffffffu, pb[i][j], o);
        pr[i][j] += __shfl_xor_sync(0xffffffffu, pr[i][j], o);
      }
      if (sub == 0) {
        const int c = (lr + i * C::LPR) * 4 + j;
        red[warp][0][c] = pg[i][j];
        red[warp][1][c] = pb[i][j];
        red[warp][2][c] = pr[i][j];
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * D; i += blockDim.x) {
    const int k = i / D, c = i - k * D;
    float s = 0.f;
    for (int w = 0; w < nwarps; ++w) s += red[w][k][c];
    partial[((size_t)blockIdx.x * 3 + k) * D + c] = s;
  }
}

// (dgamma | dbeta | dbias)[c] = sum over blocks of partial[block][k][c]: one warp per column of
// the 3*d, lanes stride the blocks, fixed shuffle tree (deterministic)
__global__ void __launch_bounds__(256)
reduce_ln_partials_kernel(const float* __restrict__ partial, int P, int d,
                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                          float* __restrict__ dbias) {
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (col >= 3 * d) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int p = lane; p < P; p += 32) s += partial[(size_t)p * 3 * d + col];
  s = warp_sum(s);
  if (lane == 0) {
    const int k = col / d, c = col - k * d;
    float* out = k == 0 ? dgamma : (k == 1 ? dbeta : dbias);
    if (out) out[c] = s;
  }
}

// out[c] = sum_p partial[p][c] (fixed order -> deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partial, int P, long n, long stride,
                       float* __restrict__ out) {
  const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += partial[(size_t)p * stride + c];
  out[c] = s;
}

// same sum for MANY partials of FEW columns: one warp per column, lane l adds partials
// l, l+32, ... in order, then a fixed-shape shuffle tree (still deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_wide_kernel(const float* __restrict__ partial, int P, long n, long stride,
                            float* __restrict__ out) {
  const long c = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int p = lane; p < P; p += 32) s += partial[(size_t)p * stride + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
}

// The same sums for MANY partials (split-K over ~148 CTAs, one LayerNorm partial row per CTA of a
// 2,368-CTA grid): the one-thread-per-column loop above is a chain of P dependent-latency loads
// (13-14 us for a few KB of output).  Block = TX consecutive columns (coalesced) x TY partial
// lanes; lane ty adds partials ty, ty+TY, ... eight independent loads at a time, then the TY lane
// sums are added in lane order: a fixed association, deterministic.  Column c of segment
// c / seg goes to out0 / out1 / out2 (LayerNorm: dgamma | dbeta | dbias; plain sums: seg = n).
template <int TX, int TY>
__global__ void __launch_bounds__(TX * TY)
reduce_partials_2d_kernel(const float* __restrict__ partial, int P, long n, long stride, long seg,
                          float* __restrict__ out0, float* __restrict__ out1,
                          float* __restrict__ out2) {
  __shared__ float red[TY][TX + 1];
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const long c = (long)blockIdx.x * TX + tx;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < n) {
    const float* src = partial + c;
    int p = ty;
    for (; p + 7 * TY < P; p += 8 * TY) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (size_t)(p + u * TY) * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += v[u];
    }
    for (int u = 0; p < P; p += TY, ++u) a[u & 7] += __ldg(src + (size_t)p * stride);
  }
  red[ty][tx] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (ty == 0 && c < n) {
    float t = 0.f;
#pragma unroll 8
    for (int y = 0; y < TY; ++y) t += red[y][tx];
    const long k = c / seg;
    float* out = k == 0 ? out0 : (k == 1 ? out1 : out2);
    if (out) out[c - k * seg] = t;
  }
}

static void launch_reduce_partials_seg(const float* partial, int P, long n, long stride, long seg,
                                       float* out0, float* out1, float* out2, cudaStream_t st) {
  if (P >= 512)
    reduce_partials_2d_kernel<16, 64><<<ceil_div(n, 16), 1024, 0, st>>>(partial, P, n, stride, seg, out0, out1, out2);
  else if (n <= 8192 || P >= 64)
    reduce_partials_2d_kernel<32, 32><<<ceil_div(n, 32), 1024, 0, st>>>(partial, P, n, stride, seg, out0, out1, out2);
  else
    reduce_partials_2d_kernel<32, 8><<<ceil_div(n, 32), 256, 0, st>>>(partial, P, n, stride, seg, out0, out1, out2);
}

static void launch_reduce_partials(const float* partial, int P, long n, long stride, float* out,
                                   cudaStream_t st) {
  static const bool old = getenv("B4CP_REDUCE_1D") != nullptr;   // the round-1 kernels (comparison)
  if (P >= 16 && !old)
    launch_reduce_partials_seg(partial, P, n, stride, n, out, nullptr, nullptr, st);
  else if (P >= 64 && n <= 4096)
    reduce_partials_wide_kernel<<<ceil_div(n, 8), 256, 0, st>>>(partial, P, n, stride, out);
  else
    reduce_partials_kernel<<<ceil_div(n, 256), 256, 0, st>>>(partial, P, n, stride, out);
}

// Column sums with the thread shape fitted to the matrix: CGP column groups of 8 columns x
// 256 / CGP row lanes per CTA, `rows` rows per CTA chosen by the host so that the grid is ~8 CTAs
// per SM whatever T and n are (the fixed 32-group x 512-row shape left 19 of 32 column lanes idle
// at n = 104 and launched 56 CTAs for the head's 28,672 rows: ~20 us for 7-58 MB).
template <int CGP>
__global__ void __launch_bounds__(256)
colsum_partial_fit_kernel(const __nv_bfloat16* __restrict__ in, long T, int n, long ld, int rows,
                          float* __restrict__ partial) {
  constexpr int RL = 256 / CGP;
  __shared__ float red[RL][CGP][9];
  const int cx = threadIdx.x % CGP, ry = threadIdx.x / CGP;
  const int c0 = (blockIdx.y * CGP + cx) * 8;
  const long r0 = (long)blockIdx.x * rows;
  const long r1 = min(T, r0 + rows);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < n) {
#pragma unroll 4
    for (long r = r0 + ry; r < r1; r += RL) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + r * ld + c0));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[2 * k] += __uint_as_float(w[k] << 16);            // low half = even column
        s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[ry][cx][k] = s[k];
  __syncthreads();
  const int g = threadIdx.x >> 3, k = threadIdx.x & 7;
  const int c = (blockIdx.y * CGP + g) * 8 + k;
  if (g < CGP && c < n) {
    float t = 0.f;
#pragma unroll 8
    for (int y = 0; y < RL; ++y) t += red[y][g][k];   // fixed order: deterministic
    partial[(size_t)blockIdx.x * n + c] = t;
  }
}

struct ColsumPlan {
  int cgp, grid_y, rows, chunks;
};
static ColsumPlan colsum_plan(long T, int n) {
  ColsumPlan p;
  const int groups = ceil_div(n, 8);
  p.cgp = groups <= 8 ? 8 : (groups <= 16 ? 16 : 32);
  p.grid_y = ceil_div(groups, p.cgp);
  const int rl = 256 / p.cgp;
  long chunks = std::max(1L, (long)(148 * 8) / p.grid_y);
  long rows = std::max((long)std::max(64, 4 * rl), (T + chunks - 1) / chunks);
  rows = (rows + rl - 1) / rl * rl;
  p.rows = (int)rows;
  p.chunks = ceil_div(T, rows);
  return p;
}

// column sums of a bf16 matrix [T][ld] over rows -> partial[chunk][n]
static constexpr int CS_ROWS = 512;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ in, long T, int n, long ld,
                      float* __restrict__ partial) {
  __shared__ float red[4][64];
  const int cx = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cx;
  const long r0 = (long)blockIdx.x * CS_ROWS;
  const long r1 = min(T, r0 + CS_ROWS);
  float s = 0.f;
  if (c < n)
    for (long r = r0 + ry; r < r1; r += 4) s += bf2f(in[r * ld + c]);
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < n)
    partial[(size_t)blockIdx.x * n + c] = red[0][cx] + red[1][cx] + red[2][cx] + red[3][cx];
}

// same partials with 128-bit loads: thread = 8 consecutive columns x every 8th row of the chunk
__global__ void __launch_bounds__(256)
colsum_partial_vec_kernel(const __nv_bfloat16* __restrict__ in, long T, int n, long ld,
                          float* __restrict__ partial) {
  __shared__ float red[8][32][9];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c0 = (blockIdx.y * 32 + cx) * 8;
  const long r0 = (long)blockIdx.x * CS_ROWS;
  const long r1 = min(T, r0 + CS_ROWS);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < n) {
#pragma unroll 4
    for (long r = r0 + ry; r < r1; r += 8) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(in + r * ld + c0));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[2 * k] += __uint_as_float(w[k] << 16);            // low half = even column
        s[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[ry][cx][k] = s[k];
  __syncthreads();
  // 256 threads = 32 column groups x 8 columns: fixed-order sum over the 8 row lanes
  const int g = threadIdx.x >> 3, k = threadIdx.x & 7;
  const int c = (blockIdx.y * 32 + g) * 8 + k;
  if (c < n) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][g][k];
    partial[(size_t)blockIdx.x * n + c] = t;
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, long rows, int cols, long ld_in,
                     __nv_bfloat16* __restrict__ out, long ld_out) {
  const long total = rows * ld_out;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    out[i] = __float2bfloat16_rn(c < cols ? in[r * ld_in + c] : 0.f);
  }
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(float* __restrict__ out, long n, DropSpec dp) {
  if (dp.thresh24) dp.seed = resolve_seed(dp.seed);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x)
    out[i] = (!dp.thresh24 || dropout_keep(dp.seed, dp.site, (uint64_t)i, dp.thresh24)) ? dp.inv_keep : 0.f;
}

__global__ void __launch_bounds__(256)
dropout_apply_kernel(float* __restrict__ x, long n, DropSpec dp) {
  dp.seed = resolve_seed(dp.seed);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x)
    x[i] = dropout_keep(dp.seed, dp.site, (uint64_t)i, dp.thresh24) ? x[i] * dp.inv_keep : 0.f;
}

static DropSpec make_drop(float rate, uint64_t seed, uint32_t site) {
  DropSpec d;
  d.inv_keep = 1.0f / (1.0f - rate);
  d.thresh24 = (uint32_t)((double)rate * 16777216.0);
  d.seed = seed;
  d.site = site;
  return d;
}

static size_t attn_smem_fwd(int S, int dh, int threads) {
  return (size_t)3 * S * (dh + 2) * 2 + (size_t)S * 4 + (size_t)(threads / 32) * S * 4;
}
static size_t attn_smem_bwd(int S, int dh, int threads) {
  return (size_t)4 * S * (dh + 2) * 2 + (size_t)3 * S * 4 + (size_t)2 * (threads / 32) * S * 4;
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_attention_fwd(const void* qkv, const int32_t* ids_first, int B, int S, int H,
                                  int dh, void* out, float* lse, void* stream) {
  B4CP_CHECK_ARG(S >= 1 && S <= MAX_S, "attention: S=%d must be in [1,%d]", S, MAX_S);
  B4CP_CHECK_ARG(dh % 2 == 0 && dh >= 2 && dh <= 128, "attention: head depth %d unsupported", dh);
  if (B == 0) return 0;
  if (attention_mma_supported(S, dh)) {
    int rc = attention_mma_fwd(qkv, ids_first, B, S, H, dh, out, lse, (cudaStream_t)stream);
    if (rc) return rc;
    note_launches(1);
    B4CP_LAUNCH_CHECK();
    return 0;
  }
  const int threads = S <= 64 ? 128 : 256;
  const size_t smem = attn_smem_fwd(S, dh, threads);
  B4CP_CHECK_ARG(smem <= 227 * 1024, "attention: S=%d dh=%d needs %zu B smem", S, dh, smem);
  B4CP_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
  attention_fwd_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)qkv, ids_first, S, H, dh, (__nv_bfloat16*)out, lse);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_attention_bwd(const void* qkv, const void* out, const void* dout,
                                  const float* lse, const int32_t* ids_first, int B, int S, int H,
                                  int dh, void* dqkv, void* stream) {
  B4CP_CHECK_ARG(S >= 1 && S <= MAX_S, "attention: S=%d must be in [1,%d]", S, MAX_S);
  B4CP_CHECK_ARG(dh % 2 == 0 && dh >= 2 && dh <= 128, "attention: head depth %d unsupported", dh);
  if (B == 0) return 0;
  // 128 < S <= 256 on the tensor cores needs the saved forward output (delta = rowsum(dO o O));
  // without it the SIMT kernel below computes delta from a full sweep
  if (attention_mma_supported(S, dh) && (S <= 128 || out)) {
    int rc = attention_mma_bwd(qkv, out, dout, lse, ids_first, B, S, H, dh, dqkv, (cudaStream_t)stream);
    if (rc) return rc;
    note_launches(1);
    B4CP_LAUNCH_CHECK();
    return 0;
  }
  const int threads = S <= 64 ? 128 : 256;
  const size_t smem = attn_smem_bwd(S, dh, threads);
  B4CP_CHECK_ARG(smem <= 227 * 1024, "attention bwd: S=%d dh=%d needs %zu B smem", S, dh, smem);
  B4CP_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
  attention_bwd_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, lse, ids_first, S, H, dh,
      (__nv_bfloat16*)dqkv);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_residual_ln_fwd(const float* x, const float* r, long T, int d,
                                    const float* gamma, const float* beta, float dropout_rate,
                                    uint64_t seed, uint32_t site, float* y_f32, void* y_bf16,
                                    long ld_bf16, void* stream) {
  B4CP_CHECK_ARG(d >= 1 && d <= LN_MAX_PER_LANE * 32, "layernorm: d=%d must be <= 256", d);
  if (T == 0) return 0;
  const DropSpec dsp = make_drop(dropout_rate, seed, site);
  const int nb = ceil_div(T, 8);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* yb = (__nv_bfloat16*)y_bf16;
  const bool vec_ok = (((uintptr_t)x | (uintptr_t)r | (uintptr_t)y_f32 | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0 &&
                      ((uintptr_t)y_bf16 & 7) == 0 && ld_bf16 % 4 == 0;
  if (vec_ok && d == 64) residual_ln_fwd_vec_kernel<64><<<ceil_div(T, 8 * LnVec<64>::RPW), 256, 0, st>>>(x, r, T, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  else if (vec_ok && d == 128) residual_ln_fwd_vec_kernel<128><<<ceil_div(T, 8), 256, 0, st>>>(x, r, T, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  else if (vec_ok && d == 256) residual_ln_fwd_vec_kernel<256><<<ceil_div(T, 8), 256, 0, st>>>(x, r, T, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  else if (d <= 32) residual_ln_fwd_kernel<1><<<nb, 256, 0, st>>>(x, r, T, d, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  else if (d <= 64) residual_ln_fwd_kernel<2><<<nb, 256, 0, st>>>(x, r, T, d, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  else if (d <= 128) residual_ln_fwd_kernel<4><<<nb, 256, 0, st>>>(x, r, T, d, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  else residual_ln_fwd_kernel<8><<<nb, 256, 0, st>>>(x, r, T, d, gamma, beta, dsp, 1e-6f, y_f32, yb, ld_bf16);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

static constexpr int LN_BWD_BLOCKS = 148 * 16;

extern "C" long b4cp_residual_ln_bwd_workspace_bytes(int d) {
  return (long)LN_BWD_BLOCKS * 3 * d * sizeof(float);
}

extern "C" int b4cp_residual_ln_bwd(const float* dy, const float* x, const float* r, long T,
                                    int d, const float* gamma, float dropout_rate, uint64_t seed,
                                    uint32_t site, float* dx, void* dr_bf16, long ld_bf16,
                                    float* dr_f32, float* dgamma, float* dbeta, float* dbias,
                                    void* workspace, void* stream) {
  B4CP_CHECK_ARG(d >= 1 && d <= LN_MAX_PER_LANE * 32, "layernorm: d=%d must be <= 256", d);
  B4CP_CHECK_ARG(workspace, "layernorm bwd: workspace required");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  const int blocks = (int)std::min<long>(LN_BWD_BLOCKS, std::max<long>(1, ceil_div(T, 8)));
  const DropSpec dsp = make_drop(dropout_rate, seed, site);
  __nv_bfloat16* drb = (__nv_bfloat16*)dr_bf16;
  const bool vec_ok = (((uintptr_t)x | (uintptr_t)r | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma | (uintptr_t)dr_f32) & 15) == 0 &&
                      ((uintptr_t)dr_bf16 & 7) == 0 && ld_bf16 % 4 == 0;
  if (vec_ok && d == 64) residual_ln_bwd_vec_kernel<64><<<blocks, 256, 0, st>>>(dy, x, r, T, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  else if (vec_ok && d == 128) residual_ln_bwd_vec_kernel<128><<<blocks, 256, 0, st>>>(dy, x, r, T, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  else if (vec_ok && d == 256) residual_ln_bwd_vec_kernel<256><<<blocks, 256, 0, st>>>(dy, x, r, T, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  else if (d <= 32) residual_ln_bwd_kernel<1><<<blocks, 256, 0, st>>>(dy, x, r, T, d, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  else if (d <= 64) residual_ln_bwd_kernel<2><<<blocks, 256, 0, st>>>(dy, x, r, T, d, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  else if (d <= 128) residual_ln_bwd_kernel<4><<<blocks, 256, 0, st>>>(dy, x, r, T, d, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  else residual_ln_bwd_kernel<8><<<blocks, 256, 0, st>>>(dy, x, r, T, d, gamma, dsp, 1e-6f, dx, drb, ld_bf16, dr_f32, partial);
  if (dgamma || dbeta || dbias)
    if (blocks >= 16 && !getenv("B4CP_REDUCE_1D"))
      launch_reduce_partials_seg(partial, blocks, 3L * d, 3L * d, d, dgamma, dbeta, dbias, st);
    else
      reduce_ln_partials_kernel<<<ceil_div(3 * d, 8), 256, 0, st>>>(partial, blocks, d, dgamma, dbeta, dbias);
  note_launches(1 + ((dgamma || dbeta || dbias) ? 1 : 0));
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" long b4cp_colsum_workspace_bytes(long T, int n) {
  const long fit = T > 0 ? colsum_plan(T, n).chunks : 1;
  return std::max((long)ceil_div(T, CS_ROWS), fit) * n * sizeof(float);
}

extern "C" int b4cp_colsum_bf16(const void* in, long T, int n, long ld, float* out,
                                void* workspace, void* stream) {
  B4CP_CHECK_ARG(workspace, "colsum: workspace required");
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {
    B4CP_CUDA(cudaMemsetAsync(out, 0, (size_t)n * 4, st));
    return 0;
  }
  int chunks = ceil_div(T, CS_ROWS);
  static const bool old_shape = getenv("B4CP_COLSUM_FIXED") != nullptr;   // comparison
  if (ld % 8 == 0 && ((uintptr_t)in & 15) == 0 && ld >= (long)ceil_div(n, 8) * 8 && !old_shape) {
    const ColsumPlan cp = colsum_plan(T, n);
    chunks = cp.chunks;
    dim3 grid(cp.chunks, cp.grid_y);
    const __nv_bfloat16* src = (const __nv_bfloat16*)in;
    if (cp.cgp == 8) colsum_partial_fit_kernel<8><<<grid, 256, 0, st>>>(src, T, n, ld, cp.rows, (float*)workspace);
    else if (cp.cgp == 16) colsum_partial_fit_kernel<16><<<grid, 256, 0, st>>>(src, T, n, ld, cp.rows, (float*)workspace);
    else colsum_partial_fit_kernel<32><<<grid, 256, 0, st>>>(src, T, n, ld, cp.rows, (float*)workspace);
  } else if (ld % 8 == 0 && ((uintptr_t)in & 15) == 0 && ld >= (long)ceil_div(n, 8) * 8) {
    dim3 grid(chunks, ceil_div(n, 256));
    colsum_partial_vec_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, T, n, ld, (float*)workspace);
  } else {
    dim3 grid(chunks, ceil_div(n, 64));
    colsum_partial_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, T, n, ld, (float*)workspace);
  }
  launch_reduce_partials((const float*)workspace, chunks, n, n, out, st);
  note_launches(2);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_reduce_splits(const float* partials, int splits, long n, long split_stride,
                                  float* out, void* stream) {
  launch_reduce_partials(partials, splits, n, split_stride, out, (cudaStream_t)stream);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

namespace b4cp {
// sum of split-K partials + optional ReLU-backward gate, fp32 and/or bf16 outputs
__global__ void __launch_bounds__(256)
reduce_splits_ex_kernel(const float* __restrict__ partials, int splits, long M, int N,
                        long split_stride, const __nv_bfloat16* __restrict__ gate, long ld_gate,
                        float* __restrict__ out_f32, long ld_f32,
                        __nv_bfloat16* __restrict__ out_bf16, long ld_bf16) {
  const long total = M * N;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / N;
    const int c = (int)(i - r * N);
    float s = 0.f;
    for (int p = 0; p < splits; ++p) s += partials[(size_t)p * split_stride + i];
    if (gate && !(bf2f(gate[r * ld_gate + c]) > 0.f)) s = 0.f;
    if (out_f32) out_f32[r * ld_f32 + c] = s;
    if (out_bf16) out_bf16[r * ld_bf16 + c] = __float2bfloat16_rn(s);
  }
}
}  // namespace b4cp

extern "C" int b4cp_reduce_splits_ex(const float* partials, int splits, long M, int N,
                                     long split_stride, const void* gate, long ld_gate,
                                     float* out_f32, long ld_f32, void* out_bf16, long ld_bf16,
                                     void* stream) {
  if (M * N == 0) return 0;
  B4CP_CHECK_ARG(out_f32 == nullptr || ld_f32 >= N, "reduce_splits_ex: ld_f32=%ld < N=%d", ld_f32, N);
  const int blocks = (int)std::min<long>(ceil_div(M * N, 256), 148L * 16);
  reduce_splits_ex_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      partials, splits, M, N, split_stride, (const __nv_bfloat16*)gate, ld_gate, out_f32, ld_f32,
      (__nv_bfloat16*)out_bf16, ld_bf16);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_cast_f32_bf16(const float* in, long rows, int cols, long ld_in, void* out,
                                  long ld_out, void* stream) {
  if (rows * ld_out == 0) return 0;
  const int blocks = (int)std::min<long>(ceil_div(rows * ld_out, 256), 148L * 16);
  cast_f32_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, rows, cols, ld_in,
                                                                  (__nv_bfloat16*)out, ld_out);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_dropout_mask(float* out, long n, float dropout_rate, uint64_t seed,
                                 uint32_t site, void* stream) {
  if (n == 0) return 0;
  const int blocks = (int)std::min<long>(ceil_div(n, 256), 148L * 16);
  dropout_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, n,
                                                                 make_drop(dropout_rate, seed, site));
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_dropout_apply(float* x, long n, float dropout_rate, uint64_t seed,
                                  uint32_t site, void* stream) {
  if (n == 0 || !(dropout_rate > 0.f)) return 0;
  const DropSpec dsp = make_drop(dropout_rate, seed, site);
  if (!dsp.thresh24) return 0;
  const int blocks = (int)std::min<long>(ceil_div(n, 256), 148L * 16);
  dropout_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, dsp);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}
