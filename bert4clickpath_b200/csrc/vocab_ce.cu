// Fused output stage of the Cloze model: SoftMaxHead's Dense(V) (head.py:36,45) + softmax +
// sparse categorical cross-entropy (examples/BERT4Rec/source/utils.py:116-134, losses.py:31-98,
// logits mode) and its gradient, on tcgen05 tensor cores.  The (M x V) logits / probabilities /
// dlogits exist only as 128x128 tiles in TMEM and shared memory - never in HBM.
//
//   forward : per 128-row tile and vocabulary chunk, S = X W (+ b) in TMEM, epilogue warps keep a
//             running (max, sum exp) per row and pick the target logit; a merge kernel produces
//             lse[M] and tgt[M].
//   backward: one CTA owns a 128-wide vocabulary tile and sweeps every row tile: S is recomputed,
//             dZ = (exp(S - lse) - onehot) / n_valid is written to shared memory as bf16 in the
//             128B-swizzled UMMA layout and feeds two more MMAs: dW_tile += X^T dZ (accumulated in
//             TMEM across the whole sweep, written once) and dX_tile = dZ W^T (reduced into HBM
//             with red.global.add).  db is a warp-shuffle column reduction of dZ.
//
// X: bf16 [M][ldx] (K-major A operand).  W: bf16 [h][ldw], the Keras (in, out) kernel, used as the
// MN-major B operand of S and, through a second descriptor on the SAME shared-memory tile, as the
// K-major B operand of dX.  h in {64, 128} for the backward, {64,128,192,256} for the forward.
#include <algorithm>
#include <cstdlib>

#include "vocab_ce.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

// bias of this warp's 32 columns, pre-multiplied by log2(e); -inf beyond the vocabulary so that
// padded columns vanish from the softmax without per-element bound checks
__device__ __forceinline__ void stage_bias(float* sb_warp, const float* bias, int vbase, int V,
                                           int lane) {
  const int v = vbase + lane;
  sb_warp[lane] = v < V ? __ldg(bias + v) * LOG2E : -INFINITY;
  __syncwarp();
}

// =============================================================================== forward
// with_dx: besides the running (max, sum), the kernel accumulates, flash-attention style,
//   U[row][:] = sum_v exp2(z2[row][v] - m_row) * W[:, v]
// with a second MMA per tile (P' = exp2(z2 - m) as bf16 in shared memory times the SAME W tile read
// K-major), rescaled in registers whenever the running maximum moves.  dX of the cross-entropy is
// then (U / sum - W[:, label]) / n_valid - computed by vocab_ce_dx_kernel - with no atomics.
// TMEM columns: S0 [0,128) S1 [128,256) U0 [256,384) U1 [384,512).
__global__ void __launch_bounds__(NUM_THREADS, 1)
vocab_ce_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const VocabParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB;
  const int NST = p.fwd_stages;
  const bool with_dx = p.with_dx != 0;
  const int x_bytes = HB * VB_M * 128;
  const int w_bytes = 2 * HB * 64 * 128;
  const int p_bytes = 2 * VB_M * 128;
  uint8_t* sX = smem;
  uint8_t* sW = sX + x_bytes;
  uint8_t* sP = sW + (size_t)NST * w_bytes;                       // [2][vb 2][128 rows][128 B]
  float* sBias = reinterpret_cast<float*>(sP + (with_dx ? 2 * p_bytes : 0));  // [16 warps][32]
  float* sMax = sBias + 16 * 32;                                   // [2][4 cg][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMax + 2 * 4 * VB_M);
  uint64_t* x_full = bars;
  uint64_t* w_full = bars + 1;          // [4]
  uint64_t* w_empty = w_full + 4;       // [4]
  uint64_t* s_full = w_empty + 4;       // [2]
  uint64_t* s_empty = s_full + 2;       // [2]
  uint64_t* p_full = s_empty + 2;       // [2]
  uint64_t* p_empty = p_full + 2;       // [2]
  uint64_t* u_full = p_empty + 2;       // [2]
  uint64_t* u_empty = u_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * VB_M;
  const int chunk = blockIdx.y;
  const int t_begin = chunk * p.tiles_per_chunk;
  const int t_end = min(p.n_vtiles, t_begin + p.tiles_per_chunk);
  const int ntiles = t_end - t_begin;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmW);
    }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  } else if (warp == WARP_MMA && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], NUM_EPI_WARPS);
      mbar_init(&p_full[b], NUM_EPI_WARPS);
      mbar_init(&p_empty[b], 1);
      mbar_init(&u_full[b], 1);
      mbar_init(&u_empty[b], NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_U = tmem_base + 256;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      mbar_expect_tx(x_full, (uint32_t)x_bytes);
      for (int hb = 0; hb < HB; ++hb)
        tma_load_2d(sX + hb * (VB_M * 128), &tmX, x_full, hb * 64, m0);
      for (int t = 0; t < ntiles; ++t) {
        const int st = t % NST;
        mbar_wait(&w_empty[st], ((t / NST) & 1) ^ 1);
        mbar_expect_tx(&w_full[st], (uint32_t)w_bytes);
        const int v0 = (t_begin + t) * VB_N;
        uint8_t* dst = sW + (size_t)st * w_bytes;
        for (int vb = 0; vb < 2; ++vb)
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(dst + (vb * HB + hb) * 8192, &tmW, &w_full[st], v0 + vb * 64, hb * 64);
      }
    }
  } else if (warp == WARP_MMA) {
    if (lane == 0) {
      const uint32_t id_s = umma_idesc_bf16(VB_M, VB_N, 0, 1);
      const uint32_t id_u = umma_idesc_bf16(VB_M, p.h, 0, 0);
      const uint32_t aX = smem_u32(sX);
      auto issue_s = [&](int t) {
        const int st = t % NST, buf = t & 1;
        mbar_wait(&w_full[st], (t / NST) & 1);
        mbar_wait(&s_empty[buf], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t aW = smem_u32(sW + (size_t)st * w_bytes);
        for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_smem_desc(aX + hb * (VB_M * 128) + kk * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(aW + hb * 8192 + kk * 2048, HB * 8192, 1024);
            umma_bf16(tmem_base + buf * VB_N, da, db, id_s, (hb | kk) ? 1u : 0u);
          }
        }
        umma_commit(&s_full[buf]);
      };
      mbar_wait(x_full, 0);
      if (ntiles > 0) issue_s(0);
      for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) issue_s(t + 1);
        const int st = t % NST, buf = t & 1;
        if (with_dx) {
          mbar_wait(&p_full[buf], (t >> 1) & 1);
          mbar_wait(&u_empty[buf], ((t >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t aP = smem_u32(sP + (size_t)buf * p_bytes);
          const uint32_t aW = smem_u32(sW + (size_t)st * w_bytes);
          for (int vb = 0; vb < 2; ++vb) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = umma_smem_desc(aP + vb * (VB_M * 128) + kk * 32, 16, 1024);
              const uint64_t db = umma_smem_desc(aW + vb * HB * 8192 + kk * 32, 16, 1024);
              umma_bf16(T_U + buf * VB_N, da, db, id_u, (vb | kk) ? 1u : 0u);
            }
          }
          umma_commit(&u_full[buf]);
          umma_commit(&p_empty[buf]);
        }
        umma_commit(&w_empty[st]);
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    const int q = warp & 3;
    const int cg = warp >> 2;
    const int r_in_tile = q * 32 + lane;
    const int row = m0 + r_in_tile;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int label = row < p.M ? p.labels[row] : -1;
    const uint32_t sb = smem_u32(sBias + warp * 32);
    const uint32_t aMax = smem_u32(sMax), aP = smem_u32(sP);
    // running (max, sum) in the log2 domain: z2 = (x.w + b) * log2(e)
    float m_run = -INFINITY, s_run = 0.f, tgt2 = 0.f, alpha_prev = 1.f;
    bool have_tgt = false;
    float acc_u[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc_u[j] = 0.f;
    auto fold_u = [&](int t_done) {  // acc_u = acc_u * alpha + U(t_done)[:, cg*32 .. +32)
      const int ub = t_done & 1;
      mbar_wait(&u_full[ub], (t_done >> 1) & 1);
      tc_fence_after();
      uint32_t u[32];
      tmem_ld32(T_U + lane_base + (uint32_t)(ub * VB_N + cg * 32), u);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(&u_empty[ub]);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc_u[j] = fmaf(acc_u[j], alpha_prev, __uint_as_float(u[j]));
    };
    // raw bias of this lane's column (-inf past V); scaled by log2(e) only when it is consumed so
    // that the load stays in flight during the whole tile
    auto load_bias = [&](int t) -> float {
      const int v = (t_begin + t) * VB_N + cg * 32 + lane;
      return (t < ntiles && v < p.V) ? __ldg(p.bias + v) : -INFINITY;
    };
    float bias_next = load_bias(0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const int vbase = (t_begin + t) * VB_N + cg * 32;
      sts32f(sb + lane * 4, bias_next * LOG2E);
      __syncwarp();
      bias_next = load_bias(t + 1);  // in flight while this tile is processed
      mbar_wait(&s_full[buf], (t >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_base + (uint32_t)(buf * VB_N + cg * 32), r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(&s_empty[buf]);  // the accumulator is in registers: release the TMEM buffer
      float z[32];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = lds128f(sb + j * 4);
        z[j + 0] = fmaf(__uint_as_float(r[j + 0]), LOG2E, b4.x);
        z[j + 1] = fmaf(__uint_as_float(r[j + 1]), LOG2E, b4.y);
        z[j + 2] = fmaf(__uint_as_float(r[j + 2]), LOG2E, b4.z);
        z[j + 3] = fmaf(__uint_as_float(r[j + 3]), LOG2E, b4.w);
        cmax = fmaxf(cmax, fmaxf(fmaxf(z[j], z[j + 1]), fmaxf(z[j + 2], z[j + 3])));
      }
      __syncwarp();  // sb is rewritten by stage_bias of the next tile
      if (label >= vbase && label < vbase + 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (vbase + j == label) tgt2 = z[j];
        have_tgt = true;
      }
      if (with_dx) {
        // the 4 warps that share these rows must agree on one running maximum per row
        const uint32_t mx = aMax + (uint32_t)(buf * (4 * VB_M) + r_in_tile) * 4;
        sts32f(mx + cg * VB_M * 4, cmax);
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
        cmax = fmaxf(fmaxf(lds32f(mx), lds32f(mx + VB_M * 4)),
                     fmaxf(lds32f(mx + 2 * VB_M * 4), lds32f(mx + 3 * VB_M * 4)));
      }
      const float m_new = fmaxf(m_run, cmax);
      const float alpha = m_new > -INFINITY ? ex2(m_run - m_new) : 1.f;  // ex2(-inf) = 0 at start
      const float msub = m_new > -INFINITY ? m_new : 0.f;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        z[j + 0] = ex2(z[j + 0] - msub);
        z[j + 1] = ex2(z[j + 1] - msub);
        z[j + 2] = ex2(z[j + 2] - msub);
        z[j + 3] = ex2(z[j + 3] - msub);
        a0 += z[j + 0];
        a1 += z[j + 1];
        a2 += z[j + 2];
        a3 += z[j + 3];
      }
      s_run = fmaf(s_run, alpha, (a0 + a1) + (a2 + a3));
      m_run = m_new;
      if (with_dx) {
        mbar_wait(&p_empty[buf], ((t >> 1) & 1) ^ 1);
        const uint32_t blk = aP + (uint32_t)(buf * p_bytes + (cg >> 1) * (VB_M * 128) + r_in_tile * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(z[8 * k + 0], z[8 * k + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(z[8 * k + 2], z[8 * k + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(z[8 * k + 4], z[8 * k + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(z[8 * k + 6], z[8 * k + 7]);
          const int chunk16 = ((cg & 1) * 4 + k) ^ (r_in_tile & 7);
          sts128(blk + chunk16 * 16, *reinterpret_cast<uint32_t*>(&h0),
                 *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                 *reinterpret_cast<uint32_t*>(&h3));
        }
        fence_proxy_async_smem();
        mbar_arrive_warp(&p_full[buf]);
        if (t > 0) fold_u(t - 1);  // alpha_prev still holds the rescale of tile t-1
        alpha_prev = alpha;
      }
    }
    if (with_dx && ntiles > 0) fold_u(ntiles - 1);
    if (row < p.M) {
      const size_t slot = ((size_t)chunk * 4 + cg) * p.M + row;
      p.part_max[slot] = m_run;   // log2 domain
      p.part_sum[slot] = s_run;
      if (have_tgt) p.tgt[row] = tgt2 * LN2;
      if (with_dx) {
        float* dst = p.part_u + ((size_t)chunk * p.M + row) * p.h + cg * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) =
              make_float4(acc_u[j], acc_u[j + 1], acc_u[j + 2], acc_u[j + 3]);
      }
    }
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// partials are (max, sum) in the log2 domain; lse is returned in natural units
__global__ void __launch_bounds__(256)
vocab_ce_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                      int n_parts, int M, int V, const int32_t* __restrict__ labels,
                      float* __restrict__ lse, float* __restrict__ tgt) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float m = -INFINITY;
  for (int c = 0; c < n_parts; ++c) m = fmaxf(m, part_max[(size_t)c * M + row]);
  float s = 0.f;
  for (int c = 0; c < n_parts; ++c) {
    const float pm = part_max[(size_t)c * M + row];
    if (pm > -INFINITY) s += part_sum[(size_t)c * M + row] * exp2f(pm - m);
  }
  lse[row] = (m + log2f(s)) * LN2;
  // padded rows, and rows whose label another vocabulary shard owns, have no local target
  if (labels[row] < 0 || labels[row] >= V) tgt[row] = 0.f;
}

// dX[row][j] = gate * ( sum_c U_c[row][j] 2^(m_c - m) / s  -  W[j][label] ) / n_valid
// h = 128: one warp per row, lane l owns columns 4l..4l+3 (128-bit loads of the U partials, all
// chunk loads independent); the (max, sum) partials of the row (4 per chunk, <= 96) are spread
// over the lanes and combined with warp reductions.
static constexpr int DX_ROWS_PER_BLOCK = 8;
static constexpr int MAX_FWD_CHUNKS = 24;

__global__ void __launch_bounds__(32 * DX_ROWS_PER_BLOCK)
vocab_ce_dx_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                   const float* __restrict__ part_u, int n_chunks, int M, int h,
                   const int32_t* __restrict__ labels, const float* __restrict__ loss_stats,
                   const float* __restrict__ lse_global, int V,
                   const __nv_bfloat16* __restrict__ w, long ldw,
                   const __nv_bfloat16* __restrict__ gate, long ld_gate,
                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long ld_bf16) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * DX_ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= M) return;
  const int label = labels[row];
  const float n_valid = loss_stats[1];
  const bool live = label >= 0 && n_valid > 0.f;
  float pmx[3] = {-INFINITY, -INFINITY, -INFINITY};
  float m = -INFINITY, ssum = 1.f;
  if (live) {
    const int n_parts = 4 * n_chunks;
    float psm[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int c = lane + 32 * t;
      pmx[t] = c < n_parts ? __ldg(part_max + (size_t)c * M + row) : -INFINITY;
      psm[t] = c < n_parts ? __ldg(part_sum + (size_t)c * M + row) : 0.f;
      m = fmaxf(m, pmx[t]);
    }
    m = warp_max(m);
    ssum = 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t)
      if (pmx[t] > -INFINITY) ssum += psm[t] * exp2f(pmx[t] - m);
    ssum = warp_sum(ssum);
  }
  // vocabulary-parallel: normalise by the GLOBAL log-sum-exp; the one-hot term belongs to the
  // shard that owns the label (labels of other shards are remapped to >= V)
  const float norm = live ? (lse_global ? exp2f(m - __ldg(lse_global + row) * LOG2E) : 1.f / ssum) : 0.f;
  const float inv_n = live ? 1.f / n_valid : 0.f;
  // lane l owns columns 4l..4l+3 of every 128-column slab (h = 128 or 256)
  for (int j = lane * 4; j < h; j += 128) {
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int c = 0; c < n_chunks; ++c) {
        const int pc = c * 4;  // the 4 column groups of a chunk share one reference maximum
        const float mine = (pc >> 5) == 0 ? pmx[0] : ((pc >> 5) == 1 ? pmx[1] : pmx[2]);
        const float pm = __shfl_sync(0xffffffffu, mine, pc & 31);
        const float sc = pm > -INFINITY ? exp2f(pm - m) : 0.f;
        const float4 v = __ldg(reinterpret_cast<const float4*>(part_u + ((size_t)c * M + row) * h + j));
        u.x = fmaf(v.x, sc, u.x); u.y = fmaf(v.y, sc, u.y);
        u.z = fmaf(v.z, sc, u.z); u.w = fmaf(v.w, sc, u.w);
      }
      float wt[4] = {0.f, 0.f, 0.f, 0.f};
      if (label < V) {
#pragma unroll
        for (int i = 0; i < 4; ++i) wt[i] = __bfloat162float(w[(size_t)(j + i) * ldw + label]);
      }
      val.x = (u.x * norm - wt[0]) * inv_n;
      val.y = (u.y * norm - wt[1]) * inv_n;
      val.z = (u.z * norm - wt[2]) * inv_n;
      val.w = (u.w * norm - wt[3]) * inv_n;
    }
    if (gate) {
      const uint2 g = *reinterpret_cast<const uint2*>(gate + (size_t)row * ld_gate + j);
      const __nv_bfloat162 g0 = *reinterpret_cast<const __nv_bfloat162*>(&g.x);
      const __nv_bfloat162 g1 = *reinterpret_cast<const __nv_bfloat162*>(&g.y);
      if (!(__bfloat162float(g0.x) > 0.f)) val.x = 0.f;
      if (!(__bfloat162float(g0.y) > 0.f)) val.y = 0.f;
      if (!(__bfloat162float(g1.x) > 0.f)) val.z = 0.f;
      if (!(__bfloat162float(g1.y) > 0.f)) val.w = 0.f;
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + (size_t)row * h + j) = val;
    if (out_bf16) {
      __nv_bfloat162 o0 = __floats2bfloat162_rn(val.x, val.y);
      __nv_bfloat162 o1 = __floats2bfloat162_rn(val.z, val.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&o0);
      o.y = *reinterpret_cast<uint32_t*>(&o1);
      *reinterpret_cast<uint2*>(out_bf16 + (size_t)row * ld_bf16 + j) = o;
    }
  }
}

// =============================================================================== backward
// Shared-memory map (h = 128, HB = 2):
//   sW  : one vocabulary tile of W      [vb 2][hb HB][64 h-rows][128 B]      = 32 KB
//   sX  : 3 row tiles of X              [buf 3][hb HB][128 rows][128 B]      = 96 KB
//   sDZ : 2 dZ tiles (bf16)             [buf 2][vb 2][128 rows][128 B]       = 64 KB
// TMEM columns: S0 [0,128) S1 [128,256) dW [384,512).  (dX comes from the forward kernel.)
static constexpr int BWD_XBUF = 3;

__global__ void __launch_bounds__(NUM_THREADS, 1)
vocab_ce_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const VocabParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB, h = p.h;
  const int w_bytes = 2 * HB * 8192;
  const int x_bytes = HB * VB_M * 128;
  const int dz_bytes = 2 * VB_M * 128;
  uint8_t* sW = smem;
  uint8_t* sX = sW + w_bytes;
  uint8_t* sDZ = sX + (size_t)BWD_XBUF * x_bytes;
  float* sDB = reinterpret_cast<float*>(sDZ + 2 * dz_bytes);  // [2 warps][128]
  float* sBias = sDB + 2 * VB_N;                               // [16 warps][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 16 * 32);
  uint64_t* w_full = bars;                 // 1
  uint64_t* w_empty = bars + 1;            // 1
  uint64_t* x_full = bars + 2;             // 3
  uint64_t* x_empty = x_full + BWD_XBUF;   // 3
  uint64_t* s_full = x_empty + BWD_XBUF;   // 2
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* dz_full = s_empty + 2;         // 2
  uint64_t* dz_empty = dz_full + 2;        // 2
  uint64_t* dx_full = dz_empty + 2;        // 1
  uint64_t* dx_empty = dx_full + 1;        // 1
  uint64_t* dw_full = dx_empty + 1;        // 1
  uint64_t* dw_empty = dw_full + 1;        // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dw_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nm = p.n_mtiles;
  // vocabulary tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int n_my = (p.n_vtiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmW);
    }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  } else if (warp == WARP_MMA && lane == 0) {
    mbar_init(w_full, 1);
    mbar_init(w_empty, 1);
    for (int i = 0; i < BWD_XBUF; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], NUM_EPI_WARPS);
      mbar_init(&dz_full[i], NUM_EPI_WARPS);
      mbar_init(&dz_empty[i], 3);  // tcgen05.commit + the two column-sum warps
    }
    mbar_init(dx_full, 1);
    mbar_init(dx_empty, NUM_EPI_WARPS);
    mbar_init(dw_full, 1);
    mbar_init(dw_empty, NUM_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t T_S = tmem_base, T_DW = tmem_base + 384;

  if (warp == WARP_TMA) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      long it = 0;  // global row-tile iteration counter across vocabulary tiles
      for (int vt = 0; vt < n_my; ++vt) {
        const int v0 = ((int)blockIdx.x + vt * (int)gridDim.x) * VB_N;
        mbar_wait(w_empty, (vt & 1) ^ 1);
        mbar_expect_tx(w_full, (uint32_t)w_bytes);
        for (int vb = 0; vb < 2; ++vb)
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sW + (vb * HB + hb) * 8192, &tmW, w_full, v0 + vb * 64, hb * 64);
        for (int i = 0; i < nm; ++i, ++it) {
          const int xb = (int)(it % BWD_XBUF);
          mbar_wait(&x_empty[xb], (uint32_t)((it / BWD_XBUF) & 1) ^ 1);
          mbar_expect_tx(&x_full[xb], (uint32_t)x_bytes);
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d(sX + (size_t)xb * x_bytes + hb * (VB_M * 128), &tmX, &x_full[xb], hb * 64,
                        i * VB_M);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t id_s = umma_idesc_bf16(VB_M, VB_N, 0, 1);   // S  = X (K-major) * W (MN-major)
      const uint32_t id_dw = umma_idesc_bf16(h, VB_N, 1, 1);     // dW = X (MN-major) * dZ (MN-major)
      const uint32_t aW = smem_u32(sW);
      const uint32_t aDZ = smem_u32(sDZ);
      auto issue_s = [&](long it) {
        const int xb = (int)(it % BWD_XBUF), sb = (int)(it & 1);
        mbar_wait(&x_full[xb], (uint32_t)((it / BWD_XBUF) & 1));
        mbar_wait(&s_empty[sb], (uint32_t)((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t aX = smem_u32(sX + (size_t)xb * x_bytes);
        for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_smem_desc(aX + hb * (VB_M * 128) + kk * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(aW + hb * 8192 + kk * 2048, HB * 8192, 1024);
            umma_bf16(T_S + sb * VB_N, da, db, id_s, (hb | kk) ? 1u : 0u);
          }
        }
        umma_commit(&s_full[sb]);
      };
      long it = 0;
      for (int vt = 0; vt < n_my; ++vt) {
        mbar_wait(w_full, vt & 1);
        issue_s(it);
        for (int i = 0; i < nm; ++i, ++it) {
          if (i + 1 < nm) issue_s(it + 1);
          const int xb = (int)(it % BWD_XBUF), zb = (int)(it & 1);
          const uint32_t aX = smem_u32(sX + (size_t)xb * x_bytes);
          const uint32_t aZ = aDZ + zb * dz_bytes;
          mbar_wait(&dz_full[zb], (uint32_t)((it >> 1) & 1));
          if (i == 0) mbar_wait(dw_empty, (vt & 1) ^ 1);
          tc_fence_after();
          // dW[h x 128] (+)= X^T dZ : K = 128 rows, 8 steps of 16
#pragma unroll
          for (int kk = 0; kk < ((p.debug & 4) ? 0 : 8); ++kk) {
            const uint64_t da = umma_smem_desc(aX + kk * 2048, VB_M * 128, 1024);
            const uint64_t db = umma_smem_desc(aZ + kk * 2048, VB_M * 128, 1024);
            umma_bf16(T_DW, da, db, id_dw, (i | kk) ? 1u : 0u);
          }
          umma_commit(&dz_empty[zb]);
          umma_commit(&x_empty[xb]);
        }
        umma_commit(dw_full);
        umma_commit(w_empty);
      }
    }
  } else if (warp >= WARP_DB0) {
    // ------------------------------------------------------------------ bias gradient
    // db[v] = sum over rows of dZ[row][v], read back from the bf16 tiles the epilogue wrote.
    // lane <-> 4 columns (8 bytes); warp 18 sums rows 0..63, warp 19 rows 64..127.
    const int half = warp - WARP_DB0;
    const int vb = lane >> 4;
    const int chunk16 = (lane & 15) >> 1;
    const int sub = (lane & 1) * 8;
    long it = 0;
    for (int vt = 0; vt < n_my; ++vt) {
      const int v0 = ((int)blockIdx.x + vt * (int)gridDim.x) * VB_N;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int i = 0; i < nm; ++i, ++it) {
        const int zb = (int)(it & 1);
        mbar_wait(&dz_full[zb], (uint32_t)((it >> 1) & 1));
        const uint32_t blk = smem_u32(sDZ) + (uint32_t)(zb * dz_bytes + vb * (VB_M * 128));
#pragma unroll 8
        for (int rr = 0; rr < ((p.debug & 1) ? 0 : 64); ++rr) {
          const int r = half * 64 + rr;
          const uint2 u = lds64(blk + r * 128 + ((chunk16 ^ (r & 7)) << 4) + sub);
          a0 += __uint_as_float(u.x << 16);
          a1 += __uint_as_float(u.x & 0xFFFF0000u);
          a2 += __uint_as_float(u.y << 16);
          a3 += __uint_as_float(u.y & 0xFFFF0000u);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&dz_empty[zb]);
      }
      float* mine = sDB + half * VB_N + lane * 4;
      mine[0] = a0; mine[1] = a1; mine[2] = a2; mine[3] = a3;
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if (half == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = lane * 4 + k;
          if (v0 + col < p.V) p.db[v0 + col] = sDB[col] + sDB[VB_N + col];
        }
      }
      asm volatile("bar.sync 1, 64;" ::: "memory");
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    const int q = warp & 3;
    const int cg = warp >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t sb = smem_u32(sBias + warp * 32);
    const uint32_t aDZs = smem_u32(sDZ);
    const float n_valid = p.loss_stats[1];
    const float inv_n = n_valid > 0.f ? 1.f / n_valid : 0.f;
    const float log2_inv_n = n_valid > 0.f ? -log2f(n_valid) : 0.f;
    long it = 0;
    for (int vt = 0; vt < n_my; ++vt) {
      const int v0 = ((int)blockIdx.x + vt * (int)gridDim.x) * VB_N;
      const int vbase = v0 + cg * 32;
      {
        const int v = vbase + lane;
        sts32f(sb + lane * 4, v < p.V ? __ldg(p.bias + v) * LOG2E : -INFINITY);
        __syncwarp();
      }
      float b2[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = lds128f(sb + j * 4);
        b2[j] = b4.x; b2[j + 1] = b4.y; b2[j + 2] = b4.z; b2[j + 3] = b4.w;
      }
      __syncwarp();
      int label_next = r_in_tile < p.M ? __ldg(p.labels + r_in_tile) : -1;
      float lse_next = r_in_tile < p.M ? __ldg(p.lse + r_in_tile) : INFINITY;
      for (int i = 0; i < nm; ++i, ++it) {
        const int sbuf = (int)(it & 1), zb = sbuf;
        const int label = label_next;
        const float scale = label >= 0 ? inv_n : 0.f;
        // dZ = exp2(z2 - lse2 + log2(scale)): the 1/n_valid factor rides in the exponent.
        // -inf for padded rows and rows past M: they contribute exactly 0.
        const float lneg = scale > 0.f ? log2_inv_n - lse_next * LOG2E : -INFINITY;
        mbar_wait(&s_full[sbuf], (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(T_S + lane_base + (uint32_t)(sbuf * VB_N + cg * 32), r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(&s_empty[sbuf]);
        float g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          g[j] = (p.debug & 2) ? __uint_as_float(r[j]) : ex2(fmaf(__uint_as_float(r[j]), LOG2E, b2[j] + lneg));
        if (label >= vbase && label < vbase + 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (vbase + j == label) g[j] -= scale;
        }
        mbar_wait(&dz_empty[zb], (uint32_t)((it >> 1) & 1) ^ 1);
        // bf16 pack + swizzled store: 16-byte chunk index XOR (row & 7) inside the 128-B row
        const uint32_t blk = aDZs + (uint32_t)(zb * dz_bytes + (cg >> 1) * (VB_M * 128) + r_in_tile * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(g[8 * k + 0], g[8 * k + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(g[8 * k + 2], g[8 * k + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(g[8 * k + 4], g[8 * k + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(g[8 * k + 6], g[8 * k + 7]);
          const int chunk16 = ((cg & 1) * 4 + k) ^ (r_in_tile & 7);
          sts128(blk + chunk16 * 16, *reinterpret_cast<uint32_t*>(&h0),
                 *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                 *reinterpret_cast<uint32_t*>(&h3));
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the UMMA async proxy
        mbar_arrive_warp(&dz_full[zb]);
        {
          // prefetch the next row tile's statistics
          const int nrow = (i + 1) * VB_M + r_in_tile;
          const bool ok = (i + 1 < nm) && nrow < p.M;
          label_next = ok ? __ldg(p.labels + nrow) : -1;
          lse_next = ok ? __ldg(p.lse + nrow) : INFINITY;
        }
      }
      // dW tile: TMEM lane = input feature, columns = vocabulary entries of this tile
      mbar_wait(dw_full, vt & 1);
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(T_DW + lane_base + (uint32_t)(cg * 32), r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_warp(dw_empty);
        if (r_in_tile < h) {
          float* dst = p.dW + (size_t)r_in_tile * p.V + vbase;
          const int ncol = min(32, p.V - vbase);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) dst[j] = __uint_as_float(r[j]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Vocabulary chunks per row tile.  One CTA per SM is resident, so the launch runs in waves of 148
// CTAs and its critical path is waves x tiles-per-chunk tile iterations; every extra chunk adds
// one more (max, sum, U) partial per row for vocab_ce_dx_kernel to read (~2.5 tile iterations of
// HBM time).  Pick the count that minimises the sum instead of "about two waves" (448 CTAs = 3.03
// waves at the bench shape cost a whole extra wave).
static int fwd_chunks(int n_mtiles, int n_vtiles, int* tiles_per_chunk) {
  int best = 1;
  double best_cost = 1e30;
  for (int c = 1; c <= std::min(n_vtiles, MAX_FWD_CHUNKS); ++c) {
    const int tpc = (n_vtiles + c - 1) / c;
    const int real = (n_vtiles + tpc - 1) / tpc;
    if (real != c) continue;
    const long waves = ((long)n_mtiles * c + 147) / 148;
    const double cost = (double)waves * tpc + 2.5 * c;
    if (cost < best_cost) best_cost = cost, best = c;
  }
  *tiles_per_chunk = (n_vtiles + best - 1) / best;
  return best;
}

}  // namespace b4cp

using namespace b4cp;

extern "C" long b4cp_vocab_ce_workspace_bytes(long M, int V, int h) {
  int tpc;
  const int chunks = fwd_chunks(ceil_div(M, VB_M), ceil_div(V, VB_N), &tpc);
  return (long)(2 * 4 + h) * chunks * M * sizeof(float) + 256;
}

// B4CP_VOCAB_IMPL=ss selects the first-generation kernels of this file (h = 128 only for the
// gradient paths); the default is the TS-form generation of vocab_ce_ts.cu.
static bool use_ts_impl() {
  const char* e = getenv("B4CP_VOCAB_IMPL");
  return !(e && e[0] == 's');
}

static int check_vocab_args(const void* x, long ldx, long M, int h, const void* w, long ldw, int V,
                            int max_h) {
  B4CP_CHECK_ARG(x && w, "vocab_ce: null operand");
  B4CP_CHECK_ARG(M > 0 && V > 0, "vocab_ce: empty problem (M=%ld V=%d)", M, V);
  B4CP_CHECK_ARG(h % 64 == 0 && h >= 64 && h <= max_h,
                 "vocab_ce: head width h=%d must be a multiple of 64 in [64,%d]", h, max_h);
  B4CP_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "vocab_ce: leading dimensions must be multiples of 8");
  B4CP_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "vocab_ce: unaligned operand");
  return 0;
}

extern "C" int b4cp_vocab_ce_fwd(const void* x_bf16, long ldx, long M, int h, const void* w_bf16,
                                 long ldw, const float* bias, int V, const int32_t* labels,
                                 int want_dx, float* lse, float* tgt, void* workspace,
                                 void* stream) {
  int rc = check_vocab_args(x_bf16, ldx, M, h, w_bf16, ldw, V, 256);
  if (rc) return rc;
  const bool ts = use_ts_impl();
  B4CP_CHECK_ARG(!want_dx || h == 128 || (ts && h == 256),
                 "vocab_ce_fwd: the dX accumulation needs h in {128, 256} (got %d)", h);
  B4CP_CHECK_ARG(bias && labels && lse && tgt && workspace, "vocab_ce_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  VocabParams p = {};
  p.M = (int)M;
  p.V = V;
  p.h = h;
  p.HB = h / 64;
  p.n_mtiles = ceil_div(M, VB_M);
  p.n_vtiles = ceil_div(V, VB_N);
  p.n_chunks = fwd_chunks(p.n_mtiles, p.n_vtiles, &p.tiles_per_chunk);
  p.bias = bias;
  p.labels = labels;
  p.part_max = (float*)workspace;
  p.part_sum = p.part_max + (size_t)4 * p.n_chunks * M;
  p.part_u = p.part_sum + (size_t)4 * p.n_chunks * M;
  p.with_dx = want_dx ? 1 : 0;
  p.fwd_stages = (want_dx || p.HB <= 2) ? 3 : 2;
  p.tgt = tgt;
  CUtensorMap tmX, tmW;
  rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, VB_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  if (ts) {
    rc = launch_vocab_fwd_ts(tmX, tmW, p, st);
    if (rc) return rc;
  } else {
    const size_t smem = (size_t)p.HB * VB_M * 128 + (size_t)p.fwd_stages * 2 * p.HB * 8192 +
                        (want_dx ? 2 * (2 * VB_M * 128) : 0) + 16 * 32 * 4 + 2 * 4 * VB_M * 4 + 512 + 1024;
    B4CP_CUDA(cudaFuncSetAttribute(vocab_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    dim3 grid(p.n_mtiles, p.n_chunks);
    vocab_ce_fwd_kernel<<<grid, NUM_THREADS, smem, st>>>(tmX, tmW, p);
  }
  vocab_ce_merge_kernel<<<ceil_div(M, 256), 256, 0, st>>>(p.part_max, p.part_sum, 4 * p.n_chunks,
                                                          (int)M, V, labels, lse, tgt);
  note_launches(2);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_vocab_ce_dx(long M, int h, int V, const int32_t* labels,
                                const float* loss_stats, const float* lse_global,
                                const void* w_bf16, long ldw,
                                const void* gate_bf16, long ld_gate, float* out_f32, void* out_bf16,
                                long ld_bf16, const void* workspace, void* stream) {
  B4CP_CHECK_ARG(h == 128 || h == 256, "vocab_ce_dx: h=%d unsupported (128 or 256)", h);
  B4CP_CHECK_ARG(labels && loss_stats && w_bf16 && workspace && (out_f32 || out_bf16),
                 "vocab_ce_dx: null argument");
  if (M == 0) return 0;
  int tpc;
  const int chunks = fwd_chunks(ceil_div(M, VB_M), ceil_div(V, VB_N), &tpc);
  const float* part_max = (const float*)workspace;
  const float* part_sum = part_max + (size_t)4 * chunks * M;
  const float* part_u = part_sum + (size_t)4 * chunks * M;
  B4CP_CHECK_ARG(ld_gate % 4 == 0 && ld_bf16 % 4 == 0, "vocab_ce_dx: leading dimensions must be multiples of 4");
  vocab_ce_dx_kernel<<<ceil_div(M, DX_ROWS_PER_BLOCK), 32 * DX_ROWS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      part_max, part_sum, part_u, chunks, (int)M, h, labels, loss_stats, lse_global, V,
      (const __nv_bfloat16*)w_bf16, ldw, (const __nv_bfloat16*)gate_bf16, ld_gate, out_f32,
      (__nv_bfloat16*)out_bf16, ld_bf16);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_vocab_ce_bwd(const void* x_bf16, long ldx, long M, int h, const void* w_bf16,
                                 long ldw, const float* bias, int V, const int32_t* labels,
                                 const float* lse, const float* loss_stats, float* dW, float* db,
                                 void* stream) {
  int rc = check_vocab_args(x_bf16, ldx, M, h, w_bf16, ldw, V, 256);
  if (rc) return rc;
  const bool ts = use_ts_impl();
  B4CP_CHECK_ARG(h == 128 || (ts && h == 256),
                 "vocab_ce_bwd: head width h=%d unsupported (the fused backward needs h in {128, 256})", h);
  B4CP_CHECK_ARG(bias && labels && lse && loss_stats && dW && db, "vocab_ce_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  VocabParams p = {};
  p.M = (int)M;
  p.V = V;
  p.h = h;
  p.HB = h / 64;
  p.n_mtiles = ceil_div(M, VB_M);
  p.n_vtiles = ceil_div(V, VB_N);
  p.bias = bias;
  p.labels = labels;
  p.lse = lse;
  p.loss_stats = loss_stats;
  p.dW = dW;
  p.db = db;
  p.debug = getenv("B4CP_DEBUG_BWD") ? atoi(getenv("B4CP_DEBUG_BWD")) : 0;
  CUtensorMap tmX, tmW;
  rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, VB_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  if (ts) {
    rc = launch_vocab_bwd_ts(tmX, tmW, p, st);
    if (rc) return rc;
  } else {
    const size_t smem = (size_t)2 * p.HB * 8192 + (size_t)BWD_XBUF * p.HB * VB_M * 128 +
                        2 * (2 * VB_M * 128) + 2 * VB_N * 4 + 16 * 32 * 4 + 256 + 1024;
    B4CP_CUDA(cudaFuncSetAttribute(vocab_ce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    const int grid = std::min(148, p.n_vtiles);
    vocab_ce_bwd_kernel<<<grid, NUM_THREADS, smem, st>>>(tmX, tmW, p);
  }
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

namespace b4cp {
// labels of a vocabulary shard [v_begin, v_begin + v_count): padded rows stay -1, labels owned by
// the shard become local ids, labels owned elsewhere become a value >= v_count (row is valid, but
// no local column is its target)
__global__ void __launch_bounds__(256)
shard_labels_kernel(const int32_t* __restrict__ labels, long M, int v_begin, int v_count,
                    int32_t* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int t = labels[i];
  out[i] = t < 0 ? -1 : ((t >= v_begin && t < v_begin + v_count) ? t - v_begin : 0x3FFFFFFF);
}
// out[m] = log sum_r exp(parts[r][m])
__global__ void __launch_bounds__(256)
lse_merge_kernel(const float* __restrict__ parts, int R, long M, float* __restrict__ out) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= M) return;
  float m = -INFINITY;
  for (int r = 0; r < R; ++r) m = fmaxf(m, parts[(size_t)r * M + i]);
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += expf(parts[(size_t)r * M + i] - m);
  out[i] = m + logf(s);
}
}  // namespace b4cp

extern "C" int b4cp_shard_labels(const int32_t* labels, long M, int v_begin, int v_count,
                                 int32_t* out, void* stream) {
  if (M == 0) return 0;
  shard_labels_kernel<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(labels, M, v_begin, v_count, out);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_lse_merge(const float* parts, int n_parts, long M, float* out, void* stream) {
  if (M == 0) return 0;
  lse_merge_kernel<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(parts, n_parts, M, out);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}
