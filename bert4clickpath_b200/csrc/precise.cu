// fp32-class ("parity") arithmetic for the paths that otherwise run on bf16 operands.
//
// The reference computes everything in fp32 (TensorFlow 2.3.1 Dense / softmax / attention ops:
// clickstream_transformer/transformer.py:64-97, :112-116, :139-167; head.py:35-45).  The fast
// path of this library feeds the tcgen05 tensor cores bf16 operands; this file supplies what the
// fp32-class mode needs to stay on the SAME tensor-core GEMM kernels and still meet the 1e-3 bar:
//
//   * b4cp_split_bf16x3: a fp32 matrix a is written as three bf16 copies along the contraction
//     axis, (hi | hi | lo) for the A operand and (hi | lo | hi) for the B operand, with
//     hi = bf16(a), lo = bf16(a - hi).  One ordinary bf16 GEMM over K' = 3K then accumulates
//     hi*hi + hi*lo + lo*hi in fp32: the dropped lo*lo term and the residual of the two-term
//     split are both <= 2^-16 relative, i.e. fp32-class products on the bf16 tensor pipe.
//   * fp32 masked self-attention (SIMT, one CTA per (sequence, head), exact division by sqrt(dh)
//     and expf as the reference's fp32 ops), forward and backward;
//   * fp32 column sums (bias gradients).
#include <algorithm>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

// ------------------------------------------------------------------------------ bf16 x 3 split
// K along the columns: in [rows][ld_in] (cols valid) -> out [rows][3*kp], kp = ld8(cols)
__global__ void __launch_bounds__(256)
split3_cols_kernel(const float* __restrict__ in, long rows, int cols, long ld_in,
                   __nv_bfloat16* __restrict__ out, int kp, int order) {
  const long total = rows * kp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / kp;
    const int c = (int)(i - r * kp);
    const float v = c < cols ? in[r * ld_in + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = out + r * 3L * kp + c;
    o[0] = hi;
    o[kp] = order ? lo : hi;
    o[2L * kp] = order ? hi : lo;
  }
}

// K along the rows: in [k][ld_in] (cols valid) -> out [3*kp][ld_out], kp = ld8(k); pad rows and
// pad columns zero (each K block is kp rows, as in the column layout, so both operands agree)
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* __restrict__ in, long k, long kp, int cols, long ld_in,
                   __nv_bfloat16* __restrict__ out, long ld_out, int order) {
  const long total = kp * ld_out;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long r = i / ld_out;
    const long c = i - r * ld_out;
    const float v = (c < cols && r < k) ? in[r * ld_in + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    out[i] = hi;
    out[i + total] = order ? lo : hi;
    out[i + 2 * total] = order ? hi : lo;
  }
}

// ------------------------------------------------------------------------- fp32 attention
static constexpr int PA_MAX_S = 256;

__device__ __forceinline__ void load_head_tile_f32(float* dst, const float* src, int S, int dh,
                                                   long ld) {
  for (int i = threadIdx.x; i < S * dh; i += blockDim.x) {
    const int r = i / dh, c = i - r * dh;
    dst[(size_t)r * (dh + 1) + c] = src[(size_t)r * ld + c];
  }
}

// grid = B*H.  qkv: fp32 [T][3d] (q | k | v), out: fp32 [T][d], lse: [B][H][S].
__global__ void __launch_bounds__(256)
attention_f32_fwd_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ ids, int S,
                         int H, int dh, float* __restrict__ out, float* __restrict__ lse_out) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int d = H * dh;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int st = dh + 1;
  float* sQ = reinterpret_cast<float*>(sm);
  float* sK = sQ + (size_t)S * st;
  float* sV = sK + (size_t)S * st;
  float* sPad = sV + (size_t)S * st;
  const int nwarps = blockDim.x >> 5;
  float* sP = sPad + S;  // [nwarps][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)b * S * 3 * d + h * dh;
  load_head_tile_f32(sQ, base, S, dh, 3L * d);
  load_head_tile_f32(sK, base + d, S, dh, 3L * d);
  load_head_tile_f32(sV, base + 2 * d, S, dh, 3L * d);
  for (int j = threadIdx.x; j < S; j += blockDim.x)
    sPad[j] = (ids[(size_t)b * S + j] == 0) ? -1e9f : 0.f;  // create_padding_mask * -1e9
  __syncthreads();
  const float sqrt_dh = sqrtf((float)dh);
  float* myP = sP + (size_t)warp * S;
  for (int i = warp; i < S; i += nwarps) {
    float z[PA_MAX_S / 32];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < PA_MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      z[t] = -INFINITY;
      if (j < S) {
        float acc = 0.f;
        const float* qr = sQ + (size_t)i * st;
        const float* kr = sK + (size_t)j * st;
        for (int c = 0; c < dh; ++c) acc = fmaf(qr[c], kr[c], acc);
        z[t] = __fdiv_rn(acc, sqrt_dh) + sPad[j];
        m = fmaxf(m, z[t]);
      }
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < PA_MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < S) {
        z[t] = expf(z[t] - m);
        sum += z[t];
      }
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int t = 0; t < PA_MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < S) myP[j] = __fdiv_rn(z[t], sum);
    }
    if (lane == 0 && lse_out) lse_out[((size_t)b * H + h) * S + i] = m + logf(sum);
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc = fmaf(myP[j], sV[(size_t)j * st + c], acc);
      out[((size_t)b * S + i) * d + h * dh + c] = acc;
    }
    __syncwarp();
  }
}

// dqkv: fp32 [T][3d].  Phase 1 (warp per query row): delta_i and dQ_i.  Phase 2 (warp per key
// row): dK_j, dV_j.  Probabilities are recomputed from the saved log-sum-exp.
__global__ void __launch_bounds__(256)
attention_f32_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                         const float* __restrict__ lse_in, const int32_t* __restrict__ ids, int S,
                         int H, int dh, float* __restrict__ dqkv) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int d = H * dh;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int st = dh + 1;
  float* sQ = reinterpret_cast<float*>(sm);
  float* sK = sQ + (size_t)S * st;
  float* sV = sK + (size_t)S * st;
  float* sDO = sV + (size_t)S * st;
  float* sPad = sDO + (size_t)S * st;
  float* sLse = sPad + S;
  float* sDelta = sLse + S;
  const int nwarps = blockDim.x >> 5;
  float* sP = sDelta + S;                 // [nwarps][S]
  float* sDZ = sP + (size_t)nwarps * S;   // [nwarps][S]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)b * S * 3 * d + h * dh;
  load_head_tile_f32(sQ, base, S, dh, 3L * d);
  load_head_tile_f32(sK, base + d, S, dh, 3L * d);
  load_head_tile_f32(sV, base + 2 * d, S, dh, 3L * d);
  load_head_tile_f32(sDO, dout + (size_t)b * S * d + h * dh, S, dh, (long)d);
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    sPad[j] = (ids[(size_t)b * S + j] == 0) ? -1e9f : 0.f;
    sLse[j] = lse_in[((size_t)b * H + h) * S + j];
  }
  __syncthreads();
  const float sqrt_dh = sqrtf((float)dh);
  float* myP = sP + (size_t)warp * S;
  float* myDZ = sDZ + (size_t)warp * S;
  float* dq_out = dqkv + (size_t)b * S * 3 * d + h * dh;

  for (int i = warp; i < S; i += nwarps) {
    float p[PA_MAX_S / 32], da[PA_MAX_S / 32];
    float delta = 0.f;
    const float lse_i = sLse[i];
#pragma unroll
    for (int t = 0; t < PA_MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      p[t] = 0.f;
      da[t] = 0.f;
      if (j < S) {
        float acc = 0.f, acc2 = 0.f;
        const float* qr = sQ + (size_t)i * st;
        const float* kr = sK + (size_t)j * st;
        const float* gr = sDO + (size_t)i * st;
        const float* vr = sV + (size_t)j * st;
        for (int c = 0; c < dh; ++c) {
          acc = fmaf(qr[c], kr[c], acc);
          acc2 = fmaf(gr[c], vr[c], acc2);
        }
        p[t] = expf(__fdiv_rn(acc, sqrt_dh) + sPad[j] - lse_i);
        da[t] = acc2;
        delta = fmaf(p[t], da[t], delta);
      }
    }
    delta = warp_sum(delta);
    if (lane == 0) sDelta[i] = delta;
#pragma unroll
    for (int t = 0; t < PA_MAX_S / 32; ++t) {
      const int j = lane + 32 * t;
      if (j < S) myDZ[j] = p[t] * (da[t] - delta);
    }
    __syncwarp();
    for (int c = lane; c < dh; c += 32) {
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc = fmaf(myDZ[j], sK[(size_t)j * st + c], acc);
      dq_out[(size_t)i * 3 * d + c] = __fdiv_rn(acc, sqrt_dh);
    }
    __syncwarp();
  }
  __syncthreads();

  for (int j = warp; j < S; j += nwarps) {
    const bool key_is_pad = sPad[j] != 0.f;  // uniform per warp
    if (!key_is_pad) {
#pragma unroll
      for (int t = 0; t < PA_MAX_S / 32; ++t) {
        const int i = lane + 32 * t;
        if (i < S) {
          float acc = 0.f, acc2 = 0.f;
          const float* qr = sQ + (size_t)i * st;
          const float* kr = sK + (size_t)j * st;
          const float* gr = sDO + (size_t)i * st;
          const float* vr = sV + (size_t)j * st;
          for (int c = 0; c < dh; ++c) {
            acc = fmaf(qr[c], kr[c], acc);
            acc2 = fmaf(gr[c], vr[c], acc2);
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
          dv = fmaf(myP[i], sDO[(size_t)i * st + c], dv);
          dk = fmaf(myDZ[i], sQ[(size_t)i * st + c], dk);
        }
      }
      dq_out[(size_t)j * 3 * d + d + c] = __fdiv_rn(dk, sqrt_dh);
      dq_out[(size_t)j * 3 * d + 2 * d + c] = dv;
    }
    __syncwarp();
  }
}

static size_t pa_smem_fwd(int S, int dh, int threads) {
  return (size_t)3 * S * (dh + 1) * 4 + (size_t)S * 4 + (size_t)(threads / 32) * S * 4;
}
static size_t pa_smem_bwd(int S, int dh, int threads) {
  return (size_t)4 * S * (dh + 1) * 4 + (size_t)3 * S * 4 + (size_t)2 * (threads / 32) * S * 4;
}

// ------------------------------------------------------------------------- fp32 column sums
static constexpr int PCS_ROWS = 512;
__global__ void __launch_bounds__(256)
colsum_f32_partial_kernel(const float* __restrict__ in, long T, int n, long ld,
                          float* __restrict__ partial) {
  __shared__ float red[4][64];
  const int cx = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cx;
  const long r0 = (long)blockIdx.x * PCS_ROWS;
  const long r1 = min(T, r0 + PCS_ROWS);
  float s = 0.f;
  if (c < n)
    for (long r = r0 + ry; r < r1; r += 4) s += in[r * ld + c];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < n)
    partial[(size_t)blockIdx.x * n + c] = red[0][cx] + red[1][cx] + red[2][cx] + red[3][cx];
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_split_bf16x3(const float* in, long rows, int cols, long ld_in, void* out,
                                 long ld_out, int k_along_rows, int order, void* stream) {
  B4CP_CHECK_ARG(in && out, "split_bf16x3: null operand");
  B4CP_CHECK_ARG(order == 0 || order == 1, "split_bf16x3: order must be 0 (hi|hi|lo) or 1 (hi|lo|hi)");
  if (rows == 0 || cols == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (k_along_rows) {
    B4CP_CHECK_ARG(ld_out >= cols && ld_out % 8 == 0, "split_bf16x3: ld_out=%ld must be a multiple of 8 >= cols", ld_out);
    const long kp = (rows + 7) / 8 * 8;
    const int blocks = (int)std::min<long>(ceil_div(kp * ld_out, 256), 148L * 16);
    split3_rows_kernel<<<blocks, 256, 0, st>>>(in, rows, kp, cols, ld_in, (__nv_bfloat16*)out, ld_out, order);
  } else {
    const int kp = (cols + 7) / 8 * 8;
    B4CP_CHECK_ARG(ld_out == 3L * kp, "split_bf16x3: ld_out=%ld must be 3 * ld8(cols) = %d", ld_out, 3 * kp);
    const int blocks = (int)std::min<long>(ceil_div(rows * kp, 256), 148L * 16);
    split3_cols_kernel<<<blocks, 256, 0, st>>>(in, rows, cols, ld_in, (__nv_bfloat16*)out, kp, order);
  }
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_attention_f32_fwd(const float* qkv, const int32_t* ids_first, int B, int S,
                                      int H, int dh, float* out, float* lse, void* stream) {
  B4CP_CHECK_ARG(S >= 1 && S <= PA_MAX_S, "attention_f32: S=%d must be in [1,%d]", S, PA_MAX_S);
  B4CP_CHECK_ARG(dh >= 1 && dh <= 128, "attention_f32: head depth %d unsupported", dh);
  if (B == 0) return 0;
  const int threads = S <= 64 ? 128 : 256;
  const size_t smem = pa_smem_fwd(S, dh, threads);
  B4CP_CHECK_ARG(smem <= 227 * 1024, "attention_f32: S=%d dh=%d needs %zu B smem", S, dh, smem);
  B4CP_CUDA(cudaFuncSetAttribute(attention_f32_fwd_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  attention_f32_fwd_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(qkv, ids_first, S, H, dh,
                                                                          out, lse);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_attention_f32_bwd(const float* qkv, const float* dout, const float* lse,
                                      const int32_t* ids_first, int B, int S, int H, int dh,
                                      float* dqkv, void* stream) {
  B4CP_CHECK_ARG(S >= 1 && S <= PA_MAX_S, "attention_f32: S=%d must be in [1,%d]", S, PA_MAX_S);
  B4CP_CHECK_ARG(dh >= 1 && dh <= 128, "attention_f32: head depth %d unsupported", dh);
  if (B == 0) return 0;
  int threads = S <= 64 ? 128 : 256;
  if (pa_smem_bwd(S, dh, threads) > 227 * 1024) threads = 128;  // fewer per-warp score rows
  const size_t smem = pa_smem_bwd(S, dh, threads);
  B4CP_CHECK_ARG(smem <= 227 * 1024, "attention_f32 bwd: S=%d dh=%d needs %zu B smem", S, dh, smem);
  B4CP_CUDA(cudaFuncSetAttribute(attention_f32_bwd_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  attention_f32_bwd_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(qkv, dout, lse, ids_first,
                                                                          S, H, dh, dqkv);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_colsum_f32(const float* in, long T, int n, long ld, float* out, void* workspace,
                               void* stream) {
  B4CP_CHECK_ARG(workspace, "colsum_f32: workspace required (b4cp_colsum_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  if (T == 0) {
    B4CP_CUDA(cudaMemsetAsync(out, 0, (size_t)n * 4, st));
    return 0;
  }
  const int chunks = ceil_div(T, PCS_ROWS);
  dim3 grid(chunks, ceil_div(n, 64));
  colsum_f32_partial_kernel<<<grid, 256, 0, st>>>(in, T, n, ld, (float*)workspace);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return b4cp_reduce_splits((const float*)workspace, chunks, n, n, out, stream);
}
