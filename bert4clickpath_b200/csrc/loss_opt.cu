// Row-wise softmax cross-entropy over MATERIALISED logits (small-vocabulary / validation path of
// the Cloze loss: examples/BERT4Rec/source/utils.py:56-134 + losses.py:31-98 in logits mode),
// the masked-mean reduction, recall@k / NDCG@k counters (utils.py:137-259) and Keras-semantics
// Adam (examples/BERT4Rec/source/main.py:87).  The large-vocabulary training path is the fused
// tcgen05 kernel in vocab_ce.cu; both share ce_loss_reduce / the metric counters.
#include <algorithm>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float nm = fmaxf(m, m2);
  if (nm == -INFINITY) {
    s = 0.f;
  } else {
    s = s * expf(m - nm) + s2 * expf(m2 - nm);
  }
  m = nm;
}

__global__ void __launch_bounds__(256)
ce_rows_stats_kernel(const float* __restrict__ logits, long ld, int V,
                     const int32_t* __restrict__ labels, float* __restrict__ lse,
                     float* __restrict__ tgt) {
  __shared__ float sm_m[8], sm_s[8];
  const long row = blockIdx.x;
  const float* z = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float x = z[v];
    if (x > m) {
      s = s * expf(m - x) + 1.f;
      m = x;
    } else {
      s += expf(x - m);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sm_m[warp] = m;
    sm_s[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = sm_m[0], S = sm_s[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) online_merge(M, S, sm_m[w], sm_s[w]);
    lse[row] = M + logf(S);
    const int t = labels[row];
    tgt[row] = (t >= 0 && t < V) ? z[t] : 0.f;
  }
}

// out[0] += 0 ; writes out[0] = sum over valid rows of (lse - tgt), out[1] = number of valid rows
__global__ void __launch_bounds__(1024)
ce_loss_reduce_kernel(const float* __restrict__ lse, const float* __restrict__ tgt,
                      const int32_t* __restrict__ labels, long M, const int32_t* __restrict__ n_global,
                      float* __restrict__ out) {
  // one CTA (the order of the sum is fixed); four rows' loads in flight per thread and shuffle
  // trees instead of ten block-wide barrier rounds: 16 -> ~4 us for 28,672 rows, on the critical
  // path between the vocabulary forward and its dX kernel
  __shared__ double s_loss[32];
  __shared__ int s_n[32];
  double a = 0.0;
  int n = 0;
  long i = threadIdx.x;
  for (; i + 3L * blockDim.x < M; i += 4L * blockDim.x) {
    int lb[4];
    float l[4], t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lb[u] = __ldg(labels + i + (long)u * blockDim.x);
      l[u] = __ldg(lse + i + (long)u * blockDim.x);
      t[u] = __ldg(tgt + i + (long)u * blockDim.x);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (lb[u] >= 0) {
        a += (double)(l[u] - t[u]);
        ++n;
      }
  }
  for (; i < M; i += blockDim.x)
    if (labels[i] >= 0) {
      a += (double)(lse[i] - tgt[i]);
      ++n;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    n += __shfl_xor_sync(0xffffffffu, n, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_loss[threadIdx.x >> 5] = a;
    s_n[threadIdx.x >> 5] = n;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? s_loss[threadIdx.x] : 0.0;
    n = threadIdx.x < (blockDim.x >> 5) ? s_n[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if (threadIdx.x == 0) {
      s_loss[0] = a;
      s_n[0] = n;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    out[0] = (float)s_loss[0];
    out[1] = n_global ? (float)*n_global : (float)s_n[0];
  }
}

// dz = (softmax(z) - onehot(label)) / n_valid for valid rows, 0 otherwise; bf16 [M][ld_dz]
__global__ void __launch_bounds__(256)
ce_rows_grad_kernel(const float* __restrict__ logits, long ld, int V,
                    const int32_t* __restrict__ labels, const float* __restrict__ lse,
                    const float* __restrict__ loss_stats, __nv_bfloat16* __restrict__ dz,
                    long ld_dz, float* __restrict__ probs, long ld_probs) {
  const long row = blockIdx.x;
  const float n = loss_stats ? loss_stats[1] : 1.f;
  const int t = labels ? labels[row] : 0;
  const float scale = (t >= 0 && n > 0.f) ? 1.f / n : 0.f;
  const float l = lse[row];
  const float* z = logits + row * ld;
  const long span = dz ? ld_dz : V;
  for (long v = threadIdx.x; v < span; v += blockDim.x) {
    float p = 0.f;
    if (v < V) p = expf(z[v] - l);
    if (probs && v < V) probs[row * ld_probs + v] = p;
    if (dz) {
      float g = 0.f;
      if (v < V) g = (p - (v == t ? 1.f : 0.f)) * scale;
      dz[row * ld_dz + v] = __float2bfloat16_rn(g);
    }
  }
}

// the same gradient in fp32, written IN PLACE over the logits (the fp32-class parity path keeps
// no bf16 copy of dZ): z[row][v] <- (softmax - onehot) / n_valid, pad columns [V, ld) <- 0
__global__ void __launch_bounds__(256)
ce_rows_grad_f32_kernel(float* __restrict__ logits, long ld, int V,
                        const int32_t* __restrict__ labels, const float* __restrict__ lse,
                        const float* __restrict__ loss_stats) {
  const long row = blockIdx.x;
  const float n = loss_stats[1];
  const int t = labels[row];
  const float scale = (t >= 0 && n > 0.f) ? 1.f / n : 0.f;
  const float l = lse[row];
  float* z = logits + row * ld;
  for (long v = threadIdx.x; v < ld; v += blockDim.x)
    z[v] = v < V ? (expf(z[v] - l) - (v == t ? 1.f : 0.f)) * scale : 0.f;
}

// counters[0] += hits, counters[1] += sum of 1/log2(rank+2) at the hit, counters[2] += n valid
__global__ void __launch_bounds__(256)
rank_metrics_kernel(const int32_t* __restrict__ topk_ids, long M, int k, long ld,
                    const int32_t* __restrict__ labels, float* __restrict__ counters) {
  __shared__ float s_h[256], s_g[256], s_n[256];
  float h = 0.f, g = 0.f, n = 0.f;
  for (long i = threadIdx.x; i < M; i += blockDim.x) {
    const int t = labels[i];
    if (t < 0) continue;
    n += 1.f;
    for (int r = 0; r < k; ++r) {
      if (topk_ids[i * ld + r] == t) {
        h += 1.f;
        g += 1.f / (logf((float)(r + 2)) / logf(2.0f));  // utils.py:211, :221-223
      }
    }
  }
  s_h[threadIdx.x] = h;
  s_g[threadIdx.x] = g;
  s_n[threadIdx.x] = n;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_h[threadIdx.x] += s_h[threadIdx.x + o];
      s_g[threadIdx.x] += s_g[threadIdx.x + o];
      s_n[threadIdx.x] += s_n[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counters[0] += s_h[0];
    counters[1] += s_g[0];
    counters[2] += s_n[0];
  }
}

// Keras Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
            float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
            const int* __restrict__ step_dev, int step_host, float grad_scale,
            __nv_bfloat16* __restrict__ shadow, int cols, long ld_shadow) {
  const int t = step_dev ? *step_dev : step_host;
  const float lr_t = lr * sqrtf(1.f - powf(b2, (float)t)) / (1.f - powf(b1, (float)t));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x) {
    const float g = grad[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    const float th = theta[i] - lr_t * mi / (sqrtf(vi) + eps);
    m[i] = mi;
    v[i] = vi;
    theta[i] = th;
    if (shadow) {
      const long r = i / cols;
      const int c = (int)(i - r * cols);
      shadow[r * ld_shadow + c] = __float2bfloat16_rn(th);
    }
  }
}

// 128-bit variant: n % 4 == 0, 16-byte aligned buffers, and (with a shadow) cols % 4 == 0 so that
// four consecutive parameters never straddle a shadow row.  Same arithmetic, element by element.
__global__ void __launch_bounds__(256)
adam_vec4_kernel(float4* __restrict__ theta, const float4* __restrict__ grad, float4* __restrict__ m,
                 float4* __restrict__ v, long n4, float lr, float b1, float b2, float eps,
                 const int* __restrict__ step_dev, int step_host, float grad_scale,
                 __nv_bfloat16* __restrict__ shadow, int cols, long ld_shadow) {
  const int t = step_dev ? *step_dev : step_host;
  const float lr_t = lr * sqrtf(1.f - powf(b2, (float)t)) / (1.f - powf(b1, (float)t));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4;
       i += (long)gridDim.x * blockDim.x) {
    const float4 g4 = grad[i];
    float4 m4 = m[i], v4 = v[i], th4 = theta[i];
    const float gs[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float th[4] = {th4.x, th4.y, th4.z, th4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = b1 * mm[j] + (1.f - b1) * gs[j];
      vv[j] = b2 * vv[j] + (1.f - b2) * gs[j] * gs[j];
      th[j] = th[j] - lr_t * mm[j] / (sqrtf(vv[j]) + eps);
    }
    m[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    v[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    theta[i] = make_float4(th[0], th[1], th[2], th[3]);
    if (shadow) {
      const long e = i * 4;
      const long r = e / cols;
      const int c = (int)(e - r * cols);
      __nv_bfloat162 a = __floats2bfloat162_rn(th[0], th[1]);
      __nv_bfloat162 b = __floats2bfloat162_rn(th[2], th[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(shadow + r * ld_shadow + c) = pk;
    }
  }
}

// ONE sweep over the whole flat parameter buffer (engine.ParamStore keeps every parameter, its
// gradient and both Adam moments in four parallel flat arrays): same arithmetic as adam_vec4_kernel,
// and for the float4s that fall inside a Dense kernel the bf16 shadow is refreshed as well.  The
// segment table (<= B4CP_ADAM_MAX_SEGS kernels) travels as a kernel parameter.
struct AdamSegs {
  int n;
  b4cp_adam_segment seg[B4CP_ADAM_MAX_SEGS];
};

__global__ void __launch_bounds__(256)
adam_flat_kernel(float4* __restrict__ theta, const float4* __restrict__ grad, float4* __restrict__ m,
                 float4* __restrict__ v, long n4, float lr, float b1, float b2, float eps,
                 const int* __restrict__ step_dev, int step_host, float grad_scale,
                 const __grid_constant__ AdamSegs segs) {
  const int t = step_dev ? *step_dev : step_host;
  const float lr_t = lr * sqrtf(1.f - powf(b2, (float)t)) / (1.f - powf(b1, (float)t));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4;
       i += (long)gridDim.x * blockDim.x) {
    const float4 g4 = grad[i];
    float4 m4 = m[i], v4 = v[i], th4 = theta[i];
    const float gs[4] = {g4.x * grad_scale, g4.y * grad_scale, g4.z * grad_scale, g4.w * grad_scale};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float th[4] = {th4.x, th4.y, th4.z, th4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = b1 * mm[j] + (1.f - b1) * gs[j];
      vv[j] = b2 * vv[j] + (1.f - b2) * gs[j] * gs[j];
      th[j] = th[j] - lr_t * mm[j] / (sqrtf(vv[j]) + eps);
    }
    m[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    v[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    theta[i] = make_float4(th[0], th[1], th[2], th[3]);
    // segments are sorted by offset and disjoint: find the one that holds element 4i, if any
    const long e = i * 4;
    int lo = 0, hi = segs.n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (segs.seg[mid].begin + segs.seg[mid].numel <= e) lo = mid + 1;
      else hi = mid;
    }
    if (lo < segs.n && segs.seg[lo].begin <= e) {
      const b4cp_adam_segment& sg = segs.seg[lo];
      __nv_bfloat16* shadow = (__nv_bfloat16*)sg.shadow_bf16;
      const long le = e - sg.begin;      // segment offsets are multiples of 4: all 4 lanes inside
      const long r = le / sg.cols;
      const int c = (int)(le - r * sg.cols);
      if ((sg.cols & 3) == 0 && (sg.ld_shadow & 3) == 0) {
        __nv_bfloat162 a = __floats2bfloat162_rn(th[0], th[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(th[2], th[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(shadow + r * sg.ld_shadow + c) = pk;
      } else {
        long rr = r;
        int cc = c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (le + j < sg.numel) shadow[rr * sg.ld_shadow + cc] = __float2bfloat16_rn(th[j]);
          if (++cc == sg.cols) {
            cc = 0;
            ++rr;
          }
        }
      }
    }
  }
}

__global__ void step_increment_kernel(int* step) { *step += 1; }

__global__ void __launch_bounds__(256)
sigmoid_kernel(const float* __restrict__ z, float* __restrict__ out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x)
    out[i] = 1.f / (1.f + expf(-z[i]));
}

// z = log(clip(p, lo, hi)) : the first half of K.sparse_categorical_crossentropy(from_logits=False)
__global__ void __launch_bounds__(256)
clip_log_kernel(const float* __restrict__ p, float* __restrict__ out, long n, float lo, float hi) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n;
       i += (long)gridDim.x * blockDim.x)
    out[i] = logf(fminf(fmaxf(p[i], lo), hi));
}

// MaskedLoss with K.binary_crossentropy (losses.py:31-98): stats = (sum of weighted masked
// item losses, number of unmasked items)
__global__ void __launch_bounds__(1024)
masked_bce_kernel(const float* __restrict__ y_true, const float* __restrict__ p, long n,
                  float label_pad, float pos_weight, int use_pos_weight, float* __restrict__ stats) {
  __shared__ double s_l[1024];
  __shared__ double s_n[1024];
  double a = 0.0, c = 0.0;
  const float eps = 1e-7f;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    const float yt = y_true[i];
    if (yt == label_pad) continue;
    const float pc = fminf(fmaxf(p[i], eps), 1.f - eps);
    float l = -(yt * logf(pc + eps) + (1.f - yt) * logf(1.f - pc + eps));
    if (use_pos_weight && yt == 1.f) l *= pos_weight;
    a += (double)l;
    c += 1.0;
  }
  s_l[threadIdx.x] = a;
  s_n[threadIdx.x] = c;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_l[threadIdx.x] += s_l[threadIdx.x + o];
      s_n[threadIdx.x] += s_n[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stats[0] = (float)s_l[0];
    stats[1] = (float)s_n[0];
  }
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_ce_rows_stats(const float* logits, long ld, long M, int V,
                                  const int32_t* labels, float* lse, float* tgt, void* stream) {
  if (M == 0) return 0;
  ce_rows_stats_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(logits, ld, V, labels, lse, tgt);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_ce_loss_reduce(const float* lse, const float* tgt, const int32_t* labels,
                                   long M, float* loss_stats, void* stream) {
  ce_loss_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lse, tgt, labels, M, nullptr, loss_stats);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_ce_loss_reduce_n(const float* lse, const float* tgt, const int32_t* labels,
                                     long M, const int32_t* n_global, float* loss_stats,
                                     void* stream) {
  B4CP_CHECK_ARG(n_global, "ce_loss_reduce_n: n_global required");
  ce_loss_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lse, tgt, labels, M, n_global, loss_stats);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_ce_rows_grad(const float* logits, long ld, long M, int V,
                                 const int32_t* labels, const float* lse, const float* loss_stats,
                                 void* dz_bf16, long ld_dz, float* probs, long ld_probs,
                                 void* stream) {
  if (M == 0) return 0;
  ce_rows_grad_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(
      logits, ld, V, labels, lse, loss_stats, (__nv_bfloat16*)dz_bf16, ld_dz, probs, ld_probs);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_ce_rows_grad_f32(float* logits, long ld, long M, int V, const int32_t* labels,
                                     const float* lse, const float* loss_stats, void* stream) {
  B4CP_CHECK_ARG(labels && lse && loss_stats, "ce_rows_grad_f32: labels, lse and loss_stats required");
  if (M == 0) return 0;
  ce_rows_grad_f32_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(logits, ld, V, labels, lse,
                                                                         loss_stats);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_rank_metrics(const int32_t* topk_ids, long M, int k, long ld,
                                 const int32_t* labels, float* counters, void* stream) {
  rank_metrics_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(topk_ids, M, k, ld, labels, counters);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_adam_step(float* theta, const float* grad, float* m, float* v, long n,
                              float lr, float beta1, float beta2, float eps, const int* step_dev,
                              int step_host, float grad_scale, void* shadow_bf16, int cols,
                              long ld_shadow, void* stream) {
  if (n == 0) return 0;
  B4CP_CHECK_ARG(step_dev || step_host >= 1, "adam: step must be >= 1");
  B4CP_CHECK_ARG(!shadow_bf16 || cols > 0, "adam: shadow needs cols");
  const bool vec = n % 4 == 0 && (((uintptr_t)theta | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0 &&
                   (!shadow_bf16 || (cols % 4 == 0 && ld_shadow % 4 == 0 && ((uintptr_t)shadow_bf16 & 7) == 0));
  if (vec) {
    const int vb = (int)std::min<long>(ceil_div(n / 4, 256), 148L * 8);
    adam_vec4_kernel<<<vb, 256, 0, (cudaStream_t)stream>>>(
        (float4*)theta, (const float4*)grad, (float4*)m, (float4*)v, n / 4, lr, beta1, beta2, eps,
        step_dev, step_host, grad_scale, (__nv_bfloat16*)shadow_bf16, cols, ld_shadow);
    note_launches(1);
    B4CP_LAUNCH_CHECK();
    return 0;
  }
  const int blocks = (int)std::min<long>(ceil_div(n, 256), 148L * 16);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(theta, grad, m, v, n, lr, beta1, beta2,
                                                        eps, step_dev, step_host, grad_scale,
                                                        (__nv_bfloat16*)shadow_bf16, cols, ld_shadow);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_adam_flat(float* theta, const float* grad, float* m, float* v, long n,
                              float lr, float beta1, float beta2, float eps, const int* step_dev,
                              int step_host, float grad_scale, const b4cp_adam_segment* h_segs,
                              int n_segs, void* stream) {
  if (n == 0) return 0;
  B4CP_CHECK_ARG(step_dev || step_host >= 1, "adam: step must be >= 1");
  B4CP_CHECK_ARG(n % 4 == 0 && (((uintptr_t)theta | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                 "adam_flat: buffers must be 16-byte aligned and n a multiple of 4");
  B4CP_CHECK_ARG(n_segs >= 0 && n_segs <= B4CP_ADAM_MAX_SEGS && (n_segs == 0 || h_segs),
                 "adam_flat: at most %d shadow segments", B4CP_ADAM_MAX_SEGS);
  AdamSegs segs;
  segs.n = n_segs;
  for (int i = 0; i < n_segs; ++i) {
    segs.seg[i] = h_segs[i];
    B4CP_CHECK_ARG(h_segs[i].begin % 4 == 0 && h_segs[i].cols > 0 && h_segs[i].shadow_bf16 &&
                       h_segs[i].begin + h_segs[i].numel <= n &&
                       (i == 0 || h_segs[i - 1].begin + h_segs[i - 1].numel <= h_segs[i].begin),
                   "adam_flat: segment %d is malformed (sorted, disjoint, 4-aligned offsets required)", i);
  }
  const int vb = (int)std::min<long>(ceil_div(n / 4, 256), 148L * 8);
  adam_flat_kernel<<<vb, 256, 0, (cudaStream_t)stream>>>((float4*)theta, (const float4*)grad, (float4*)m,
                                                         (float4*)v, n / 4, lr, beta1, beta2, eps, step_dev,
                                                         step_host, grad_scale, segs);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_zero(void* ptr, long bytes, void* stream) {
  if (bytes <= 0) return 0;
  B4CP_CHECK_ARG(ptr, "zero: null pointer");
  B4CP_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}

extern "C" int b4cp_step_increment(int* step_dev, void* stream) {
  step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_sigmoid(const float* z, float* out, long n, void* stream) {
  if (n == 0) return 0;
  const int blocks = (int)std::min<long>(ceil_div(n, 256), 148L * 16);
  sigmoid_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(z, out, n);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_clip_log(const float* p, float* out, long n, float lo, float hi, void* stream) {
  if (n == 0) return 0;
  const int blocks = (int)std::min<long>(ceil_div(n, 256), 148L * 16);
  clip_log_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, out, n, lo, hi);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

namespace b4cp {
// ---- backward of BinaryClassificationHead's Dense(1, sigmoid) + MaskedLoss(binary_crossentropy)
// (head.py:11,24-26; losses.py:31-98 with K.binary_crossentropy on probabilities):
//   l_i = -w_i (y log(pc + e) + (1 - y) log(1 - pc + e)),  pc = clip(p, e, 1 - e),  e = 1e-7,
//   w_i = pos_weight where y == 1 (if given), loss = sum l_i / n over labels != label_pad,
//   divided by (pos_weight + 1) / 2 when pos_weight is given (losses.py:94-96).
// dz_i = dl/dp * p (1 - p) / n (the clip passes a gradient only strictly inside (e, 1 - e)).
__device__ __forceinline__ float bce_dz_item(float yt, float pi, float nv, float label_pad,
                                             float pos_weight, int use_pos_weight) {
  const float eps = 1e-7f;
  float g = 0.f;
  if (yt != label_pad && nv > 0.f) {
    if (pi > eps && pi < 1.f - eps) {
      float dl = -(yt / (pi + eps) - (1.f - yt) / (1.f - pi + eps));
      if (use_pos_weight && yt == 1.f) dl *= pos_weight;
      g = dl * pi * (1.f - pi) / nv;
      // MaskedLoss divides the weighted mean by (pos_weight + negative_weight) / 2 (losses.py:94-96)
      if (use_pos_weight) g *= 2.f / (pos_weight + 1.f);
    }
  }
  return g;
}

__global__ void __launch_bounds__(256)
binary_head_dz_kernel(const float* __restrict__ y_true, const float* __restrict__ p, long n,
                      float label_pad, float pos_weight, int use_pos_weight,
                      const float* __restrict__ stats, float* __restrict__ dz) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  dz[i] = bce_dz_item(y_true[i], p[i], stats[1], label_pad, pos_weight, use_pos_weight);
}

// The same item-wise gradient over a (rows, cols) sigmoid output (MultiLabel_MultiClass_
// classification, head.py:50-69: every (row, class) cell is one item of the masked mean), written
// as the bf16 [rows][ld] operand of the dW / dx GEMMs (pad columns zeroed) and/or fp32 [rows][cols].
__global__ void __launch_bounds__(256)
sigmoid_bce_dz_kernel(const float* __restrict__ y_true, const float* __restrict__ p, long rows,
                      int cols, long ld, float label_pad, float pos_weight, int use_pos_weight,
                      const float* __restrict__ stats, float* __restrict__ dz_f32,
                      __nv_bfloat16* __restrict__ dz_bf16) {
  const float nv = stats[1];
  const long total = rows * ld;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
       idx += (long)gridDim.x * blockDim.x) {
    const long r = idx / ld;
    const int c = (int)(idx - r * ld);
    float g = 0.f;
    if (c < cols) {
      const long i = r * cols + c;
      g = bce_dz_item(y_true[i], p[i], nv, label_pad, pos_weight, use_pos_weight);
      if (dz_f32) dz_f32[i] = g;
    }
    if (dz_bf16) dz_bf16[idx] = __float2bfloat16_rn(g);
  }
}

// d_ab[i][c] = dz_i * w[c] (zeroed where the ReLU output ab[i][c] <= 0 when `gated`)
__global__ void __launch_bounds__(256)
binary_head_dx_kernel(const float* __restrict__ dz, const float* __restrict__ w, long M, int h,
                      const __nv_bfloat16* __restrict__ ab, long ld_ab, int gated,
                      float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long ld_bf16) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= M * h) return;
  const long i = idx / h;
  const int c = (int)(idx - i * h);
  float v = dz[i] * w[c];
  if (gated && !(__bfloat162float(ab[i * ld_ab + c]) > 0.f)) v = 0.f;
  if (out_f32) out_f32[i * h + c] = v;
  if (out_bf16) out_bf16[i * ld_bf16 + c] = __float2bfloat16_rn(v);
}

// dw[c] = sum_i ab[i][c] dz_i, db = sum_i dz_i: one block per 32 columns, warps stride the rows,
// the 8 warp partials are added in order (deterministic)
__global__ void __launch_bounds__(256)
binary_head_dw_kernel(const float* __restrict__ dz, const __nv_bfloat16* __restrict__ ab, long ld_ab,
                      long M, int h, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f, accb = 0.f;
  for (long i = warp; i < M; i += 8) {
    const float g = dz[i];
    if (c < h) acc += __bfloat162float(ab[i * ld_ab + c]) * g;
    if (blockIdx.x == 0 && lane == 0) accb += g;
  }
  part[warp][lane] = acc;
  if (lane == 0) part[warp][32] = accb;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    if (c < h) dw[c] = t;
    if (blockIdx.x == 0 && lane == 0) {
      float tb = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tb += part[k][32];
      *db = tb;
    }
  }
}
}  // namespace b4cp

extern "C" int b4cp_binary_head_bwd(const float* y_true, const float* probs, long M, float label_pad,
                                    float pos_weight, int use_pos_weight, const float* stats,
                                    const void* ab_bf16, long ld_ab, int h, const float* w_out,
                                    int gated, float* dz, float* dab_f32, void* dab_bf16,
                                    long ld_dab, float* dw, float* db, void* stream) {
  B4CP_CHECK_ARG(y_true && probs && stats && ab_bf16 && w_out && dz && dw && db && (dab_f32 || dab_bf16),
                 "binary_head_bwd: null argument");
  B4CP_CHECK_ARG(h >= 1, "binary_head_bwd: h=%d", h);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  binary_head_dz_kernel<<<ceil_div(M, 256), 256, 0, st>>>(y_true, probs, M, label_pad, pos_weight,
                                                          use_pos_weight, stats, dz);
  binary_head_dx_kernel<<<ceil_div(M * h, 256), 256, 0, st>>>(
      dz, w_out, M, h, (const __nv_bfloat16*)ab_bf16, ld_ab, gated, dab_f32,
      (__nv_bfloat16*)dab_bf16, ld_dab);
  binary_head_dw_kernel<<<ceil_div(h, 32), 256, 0, st>>>(dz, (const __nv_bfloat16*)ab_bf16, ld_ab, M,
                                                         h, dw, db);
  note_launches(3);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_sigmoid_bce_dz(const float* y_true, const float* probs, long rows, int cols,
                                   float label_pad, float pos_weight, int use_pos_weight,
                                   const float* stats, float* dz_f32, void* dz_bf16, long ld_bf16,
                                   void* stream) {
  B4CP_CHECK_ARG(y_true && probs && stats && (dz_f32 || dz_bf16), "sigmoid_bce_dz: null argument");
  B4CP_CHECK_ARG(cols >= 1 && (!dz_bf16 || ld_bf16 >= cols), "sigmoid_bce_dz: cols=%d ld=%ld", cols,
                 ld_bf16);
  if (rows == 0) return 0;
  const long ld = dz_bf16 ? ld_bf16 : cols;
  const int blocks = (int)std::min<long>(ceil_div(rows * ld, 256), 148L * 16);
  sigmoid_bce_dz_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      y_true, probs, rows, cols, ld, label_pad, pos_weight, use_pos_weight, stats, dz_f32,
      (__nv_bfloat16*)dz_bf16);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

namespace b4cp {
// counters += (sum mask, sum y*mask, sum round(p)*mask, tp, condition_true, predicted_true):
// the accumulators of clickstream_transformer/metrics.py (PositiveRate :12-20, PredictedPositives
// :36-45, F1Score :63-78).  mask = (y != label_pad); round = round-half-to-even (tf.round);
// tp / condition_true / predicted_true follow F1Score literally: int32 casts, NO mask (a padded
// label -1 is never "== 1", but a padded position still counts as predicted-true).
__global__ void __launch_bounds__(1024)
binary_metric_counts_kernel(const float* __restrict__ y_true, const float* __restrict__ p, long n,
                            float label_pad, float* __restrict__ counters) {
  __shared__ double sh[6][1024];
  double a[6] = {0, 0, 0, 0, 0, 0};
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    const float yt = y_true[i];
    const float m = yt != label_pad ? 1.f : 0.f;
    const float pr = rintf(p[i]);
    a[0] += m;
    a[1] += yt * m;
    a[2] += pr * m;
    const bool ct = (int)yt == 1, pt = (int)pr == 1;
    a[3] += (ct && pt) ? 1.0 : 0.0;
    a[4] += ct ? 1.0 : 0.0;
    a[5] += pt ? 1.0 : 0.0;
  }
  for (int j = 0; j < 6; ++j) sh[j][threadIdx.x] = a[j];
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int j = 0; j < 6; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 6) counters[threadIdx.x] += (float)sh[threadIdx.x][0];
}
}  // namespace b4cp

extern "C" int b4cp_binary_metric_counts(const float* y_true, const float* probs, long n,
                                         float label_pad, float* counters, void* stream) {
  B4CP_CHECK_ARG(y_true && probs && counters, "binary_metric_counts: null argument");
  if (n == 0) return 0;
  binary_metric_counts_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(y_true, probs, n, label_pad, counters);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_masked_bce(const float* y_true, const float* probs, long n, float label_pad,
                               float pos_weight, int use_pos_weight, float* stats, void* stream) {
  masked_bce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(y_true, probs, n, label_pad, pos_weight,
                                                          use_pos_weight, stats);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}
