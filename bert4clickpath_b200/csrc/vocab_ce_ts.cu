// TS-form generation of the fused output stage (see vocab_ce.cu for the mathematics and the
// reference citations: head.py:36,45; examples/BERT4Rec/source/utils.py:116-134; losses.py:31-98).
//
// What changes against the SS-form kernels of vocab_ce.cu: the bf16 tile that feeds the SECOND
// tensor-core product of each kernel (P' = exp2(z2 - m) in the forward, dZ^T in the backward) is
// written by the epilogue warps straight into TMEM, over the fp32 accumulator columns they have
// just read, and is consumed as the TMEM A operand of `tcgen05.mma` (A from TMEM, B from shared
// memory).  It never touches shared memory, so the 128 B/clk shared-memory port only serves the
// B operands, and the 64 KB of staging buffers are gone - which is what lets a 256-wide head
// (V = 1M, h = 256) fit: X tile 64 KB + two 64 KB W stages.
//
//   forward : S = X W in TMEM (lanes = rows).  U = sum_v P'_v W_v^T accumulates IN TMEM over the
//             whole vocabulary chunk (no per-tile fold into registers): P' is taken relative to a
//             per-row reference maximum m_ref that only moves when the running maximum exceeds
//             it by more than 2^8 ("lazy rescale"; bf16 keeps fp32's exponent range, so P' <= 256
//             loses nothing); on that rare event the owning warps rescale their U columns in TMEM.
//   backward: the transposed problem.  S^T = W^T X^T (lanes = vocabulary entries, columns = rows),
//             dZ^T = exp2(z2 - lse2)/n - onehot goes back to TMEM, dW^T[v][:] += dZ^T X accumulates
//             in TMEM over the sweep of all row tiles (N = h up to 256 in one instruction).  The
//             bias gradient is a per-thread register sum (lane = vocabulary entry), and dW is
//             written with 128-byte coalesced rows.
#include <algorithm>
#include <cstdlib>

#include "vocab_ce.cuh"

namespace b4cp {

static constexpr int TS_THREADS = 576;  // 16 epilogue warps + TMA warp + MMA warp
static constexpr float RESCALE_TH = 8.f;
// Forward kernel template parameter PP: one pair in every PP groups of four exponentials is
// computed by ex2_poly2 instead of MUFU (2: a quarter of them - measured best; 1: half - the FMA
// pipe then becomes the limiter; 0: none).
static constexpr int FWD_POLY_DEFAULT = 2;
// (the backward kernel keeps MUFU for all exponentials: its -inf arguments - padded rows - must
// give exactly 0, and with two ping-pong epilogue groups the polynomial did not pay: 0.76 vs 0.73 ms)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// =============================================================================== forward
// TMEM columns: S0 [0,128) S1 [128,256) U [256, 256+h).  P'(t) overwrites S(t&1): epilogue warp
// (q, cg) owns lanes [32q,+32) x columns [32cg,+32) of S and writes its 32 probabilities as 16
// packed columns at [32cg, 32cg+16) - inside its own region, so no warp overwrites columns another
// warp still has to read.  K step k8 (vocabulary entries 16k8..16k8+15 of the tile) therefore
// reads A at column 32*(k8>>1) + 8*(k8&1).
//
// A CTA processes a list of SEGMENTS (vocab_ce.cuh: one under SCHED_GRID, the pieces of its tile
// range under SCHED_RANGES).  The W ring, the S / P' / U hand-shakes and their phases run on ONE
// tile counter across segments; per segment there is an X tile (ring of p.x_bufs buffers), a fresh
// set of row statistics, and at its end a drain of U from TMEM into the segment's partial slot
// (`u_drained` tells the MMA warp that U may be overwritten by the next segment's first product).
struct FwdSeg {
  int m, v0, len, slot;
  int rot;   // SCHED_GRID: tile i of the segment is vocabulary tile v0 + (i + rot) % len
  __device__ __forceinline__ int vtile(int i) const {
    int j = i + rot;
    if (j >= len) j -= len;
    return v0 + j;
  }
};

__device__ __forceinline__ void fwd_range(const VocabParams& p, long& L0, long& L1) {
  if (p.sched == SCHED_GRID) {
    const int t_begin = (int)blockIdx.y * p.tiles_per_chunk;
    const int t_end = min(p.n_vtiles, t_begin + p.tiles_per_chunk);
    L0 = 0;
    L1 = t_end - t_begin;
  } else {
    L0 = (long)blockIdx.x * p.range_q;
    L1 = min(p.total_tiles, L0 + p.range_q);
  }
}

__device__ __forceinline__ FwdSeg fwd_seg(const VocabParams& p, long L, long L1) {
  FwdSeg s;
  if (p.sched == SCHED_GRID) {
    s.m = (int)blockIdx.x;
    s.v0 = (int)blockIdx.y * p.tiles_per_chunk + (int)L;
    s.len = (int)(L1 - L);
    s.slot = (int)blockIdx.y;
    // The CTAs of a chunk sweep the same W tiles.  In exact lock step all of them ask the SAME
    // few L2 slices for the same 64 KB at the same moment; rotating each row tile's starting
    // point by a few tiles spreads the requests over the slices while the chunk's working set
    // (n_mtiles x stagger tiles) still sits in L2, so every W tile is still read from HBM once.
    s.rot = s.len > 0 ? (int)(((long)blockIdx.x * p.grid_stagger) % s.len) : 0;
  } else {
    s.rot = 0;
    s.m = (int)(L / p.n_vtiles);
    s.v0 = (int)(L - (long)s.m * p.n_vtiles);
    const long rem = L1 - L;
    const long left = p.n_vtiles - s.v0;
    s.len = (int)(rem < left ? rem : left);
    s.slot = (int)blockIdx.x - ranges_first_cta(p.range_q, p.n_vtiles, s.m);
  }
  return s;
}

// NSB: S accumulators in TMEM (3 at h <= 128, 2 at h = 256).  OPT: optimistic exponentials (below);
// they hold the S buffer until the row maxima are agreed, which only pays with three buffers.
template <int NSB, int PP, bool OPT, bool KSPLIT>
__global__ void __launch_bounds__(TS_THREADS, 1)
vocab_ce_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmX,
                       const __grid_constant__ CUtensorMap tmW, const VocabParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB;
  const int NST = p.fwd_stages;
  const int XB = p.x_bufs;
  const bool with_dx = p.with_dx != 0;
  const int x_bytes = HB * VB_M * 128;
  const int w_bytes = 2 * HB * 64 * 128;
  uint8_t* sX = smem;
  uint8_t* sW = sX + (size_t)XB * x_bytes;
  float* sBias = reinterpret_cast<float*>(sW + (size_t)NST * w_bytes);  // [16 warps][32]
  float* sMax = sBias + 16 * 32;                                         // [2][4 cg][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sMax + 2 * 4 * VB_M);
  uint64_t* x_full = bars;              // [2]
  uint64_t* x_empty = bars + 2;         // [2]
  uint64_t* w_full = x_empty + 2;       // [4]
  uint64_t* w_empty = w_full + 4;       // [4]
  uint64_t* s_full = w_empty + 4;       // [3]
  uint64_t* s_empty = s_full + 3;       // [3]
  uint64_t* p_full = s_empty + 3;       // [3]
  uint64_t* u_full = p_full + 3;        // [3]
  uint64_t* u_drained = u_full + 3;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_drained + 1);
  // S accumulators: three at h <= 128 (columns 0,128,256; U at 384), two at h = 256 (U at 256).
  // With three, S runs two tiles ahead of the epilogue: S(t+3) only has to follow U(t), so an
  // epilogue that takes about as long as the tile's two MMAs no longer stalls on every other tile.

  // warp index through shfl: provably warp-uniform, so the role branches are uniform control flow
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  long L_begin, L_end;
  fwd_range(p, L_begin, L_end);

  if (warp == WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmW);
    }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  } else if (warp == WARP_MMA && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], NUM_EPI_WARPS);
      mbar_init(&p_full[b], NUM_EPI_WARPS);
      mbar_init(&u_full[b], 1);
    }
    mbar_init(u_drained, NUM_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t T_U = tmem_base + (uint32_t)(NSB * VB_N);

  if (warp == WARP_TMA) {
    // whole warp, uniform control flow; the TMA instructions are predicated on an elected lane
    const uint32_t aX = smem_u32(sX), aW0 = smem_u32(sW);
    int t = 0;
    int sg = 0;
    for (long L = L_begin; L < L_end; ++sg) {
      const FwdSeg s = fwd_seg(p, L, L_end);
      const int xb = sg % XB;
      mbar_wait_all(&x_empty[xb], (uint32_t)((sg / XB) & 1) ^ 1);
      mbar_expect_tx_el(&x_full[xb], (uint32_t)x_bytes);
      for (int hb = 0; hb < HB; ++hb)
        tma_load_2d_el(aX + xb * x_bytes + hb * (VB_M * 128), &tmX, &x_full[xb], hb * 64, s.m * VB_M);
      for (int i = 0; i < s.len; ++i, ++t) {
        const int st = t % NST;
        if (p.l2_prefetch && (int)blockIdx.x % p.l2_prefetch_every == 0 &&
            i + p.l2_prefetch < s.len) {   // W tile i + PF of this segment -> L2
          const int vp = s.vtile(i + p.l2_prefetch) * VB_N;
          for (int vb = 0; vb < 2; ++vb)
            for (int hb = 0; hb < HB; ++hb) tma_prefetch_2d_el(&tmW, vp + vb * 64, hb * 64);
        }
        const int v0 = s.vtile(i) * VB_N;
        if (KSPLIT) {
          // h = 256: a W tile travels as two K halves (h rows [0,128) and [128,256)) through a
          // ring of FOUR 32 KB half stages - see the MMA warp
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int hi = 2 * t + hh, sl = hi & 3;
            mbar_wait_all(&w_empty[sl], (uint32_t)((hi >> 2) & 1) ^ 1);
            mbar_expect_tx_el(&w_full[sl], 32768u);
            const uint32_t dst = aW0 + (uint32_t)(sl * 32768);
            for (int vb = 0; vb < 2; ++vb)
              for (int hb2 = 0; hb2 < 2; ++hb2)
                tma_load_2d_el(dst + (vb * 2 + hb2) * 8192, &tmW, &w_full[sl], v0 + vb * 64,
                               (hh * 2 + hb2) * 64);
          }
        } else {
          mbar_wait_all(&w_empty[st], (uint32_t)((t / NST) & 1) ^ 1);
          mbar_expect_tx_el(&w_full[st], (uint32_t)w_bytes);
          const uint32_t dst = aW0 + (uint32_t)(st * w_bytes);
          for (int vb = 0; vb < 2; ++vb)
            for (int hb = 0; hb < HB; ++hb)
              tma_load_2d_el(dst + (vb * HB + hb) * 8192, &tmW, &w_full[st], v0 + vb * 64, hb * 64);
        }
      }
      L += s.len;
    }
  } else if (warp == WARP_MMA) {
    // The WHOLE warp runs this loop in uniform control flow; only the tcgen05 instructions are
    // predicated on an elected lane (see umma_bf16_el in common.cuh).
    const uint32_t id_s = umma_idesc_bf16(VB_M, VB_N, 0, 1);
    const uint32_t id_u = umma_idesc_bf16(VB_M, p.h, 0, 0);
    const uint32_t aX = smem_u32(sX);
    const uint32_t aW0 = smem_u32(sW);
    // descriptors differ only in their 14-bit start-address field: build once, add (bytes >> 4)
    const uint64_t dX = umma_smem_desc(aX, 16, 1024);                // X, K-major (A of S)
    const uint64_t dWs = umma_smem_desc(aW0, HB * 8192, 1024);       // W, MN-major (B of S)
    const uint64_t dWu = umma_smem_desc(aW0, 16, 1024);              // W, K-major (B of U)
    // two cursors over the segment list: the S queue and the U queue
    long Ls = L_begin, Lu = L_begin;
    int sgs = 0, sgu = 0, is = 0, iu = 0;         // segment index / tile index inside it
    int len_s = Ls < L_end ? fwd_seg(p, Ls, L_end).len : 0;
    int len_u = len_s;
    const int n_total = (int)(L_end - L_begin);
    // KSPLIT (h = 256): shared memory holds X (64 KB) and only 128 KB of W.  As two whole-tile
    // stages, the load of W(t+2) could not start before U(t) had finished with its stage, and
    // S(t+2) - which the epilogue needs next - waited a full 64 KB L2 -> SM transfer behind it.
    // The tile is therefore split along K = h into two 32 KB halves in a ring of four: U(t) is
    // issued as two N = 128 products (U columns [0,128) from half 0, [128,256) from half 1), the
    // first of which frees half 0 half a product earlier, and each half is a transfer half as long.
    const uint64_t dWsK = umma_smem_desc(aW0, 2 * 8192, 1024);      // KSPLIT: the vb blocks are 16 KB apart
    const uint32_t id_u128 = umma_idesc_bf16(VB_M, 128, 0, 0);
    auto issue_s = [&](int t) {
      const int st = t % NST, buf = t % NSB, xb = sgs % XB;
      tc_fence_after();
      const uint32_t wo = (uint32_t)(st * w_bytes) >> 4;
      const uint32_t xo = (uint32_t)(xb * x_bytes) >> 4;
      if (KSPLIT) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t ho = (uint32_t)(((2 * t + hh) & 3) * 32768) >> 4;
#pragma unroll
          for (int hb2 = 0; hb2 < 2; ++hb2)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_el(tmem_base + buf * VB_N,
                           dX + (xo + (uint32_t)(((hh * 2 + hb2) * (VB_M * 128) + kk * 32) >> 4)),
                           dWsK + (ho + (uint32_t)((hb2 * 8192 + kk * 2048) >> 4)), id_s,
                           (hh | hb2 | kk) ? 1u : 0u);
        }
      } else {
        for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_el(tmem_base + buf * VB_N, dX + (xo + (uint32_t)((hb * (VB_M * 128) + kk * 32) >> 4)),
                         dWs + (wo + (uint32_t)((hb * 8192 + kk * 2048) >> 4)), id_s, (hb | kk) ? 1u : 0u);
        }
      }
      umma_commit_el(&s_full[buf]);
      if (is == len_s - 1) umma_commit_el(&x_empty[xb]);   // last reader of this X tile
    };
    auto advance_s = [&]() {
      if (++is == len_s) {
        Ls += len_s;
        ++sgs;
        is = 0;
        len_s = Ls < L_end ? fwd_seg(p, Ls, L_end).len : 0;
      }
    };
    auto w_ready = [&](int t) -> bool {
      if (KSPLIT)
        return mbar_test_all(&w_full[(2 * t) & 3], (uint32_t)(((2 * t) >> 2) & 1)) &&
               mbar_test_all(&w_full[(2 * t + 1) & 3], (uint32_t)(((2 * t + 1) >> 2) & 1));
      return mbar_test_all(&w_full[t % NST], (uint32_t)((t / NST) & 1));
    };
    auto s_ready = [&](int t) -> bool {
      if (is == 0 && !mbar_test_all(&x_full[sgs % XB], (uint32_t)((sgs / XB) & 1))) return false;
      return w_ready(t) && mbar_test_all(&s_empty[t % NSB], (uint32_t)((t / NSB) & 1) ^ 1);
    };
    auto issue_u = [&](int t) {
      const int st = t % NST, buf = t % NSB;
      tc_fence_after();
      const uint32_t wo = (uint32_t)(st * w_bytes) >> 4;
      const uint32_t tP = tmem_base + buf * VB_N;
      if (KSPLIT) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int sl = (2 * t + hh) & 3;
          const uint32_t ho = (uint32_t)(sl * 32768) >> 4;
#pragma unroll
          for (int k8 = 0; k8 < 8; ++k8) {
            const int vb = k8 >> 2, kk = k8 & 3;
            umma_bf16_ts_el(T_U + hh * 128, tP + (k8 >> 1) * 32 + (k8 & 1) * 8,
                            dWu + (ho + (uint32_t)((vb * 2 * 8192 + kk * 32) >> 4)), id_u128,
                            (iu | k8) ? 1u : 0u);
          }
          umma_commit_el(&w_empty[sl]);   // this K half is free as soon as its product retires
        }
        umma_commit_el(&u_full[buf]);
        return;
      }
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        const int vb = k8 >> 2, kk = k8 & 3;
        umma_bf16_ts_el(T_U, tP + (k8 >> 1) * 32 + (k8 & 1) * 8,
                        dWu + (wo + (uint32_t)((vb * HB * 8192 + kk * 32) >> 4)), id_u, (iu | k8) ? 1u : 0u);
      }
      umma_commit_el(&u_full[buf]);
      umma_commit_el(&w_empty[st]);
    };
    if (!with_dx) {
      for (int t = 0; t < n_total; ++t) {
        if (is == 0) mbar_wait_all(&x_full[sgs % XB], (uint32_t)((sgs / XB) & 1));
        if (KSPLIT) {
          mbar_wait_all(&w_full[(2 * t) & 3], (uint32_t)(((2 * t) >> 2) & 1));
          mbar_wait_all(&w_full[(2 * t + 1) & 3], (uint32_t)(((2 * t + 1) >> 2) & 1));
        } else {
          mbar_wait_all(&w_full[t % NST], (uint32_t)((t / NST) & 1));
        }
        mbar_wait_all(&s_empty[t % NSB], (uint32_t)((t / NSB) & 1) ^ 1);
        issue_s(t);
        if (KSPLIT) {
          umma_commit_el(&w_empty[(2 * t) & 3]);
          umma_commit_el(&w_empty[(2 * t + 1) & 3]);
        } else {
          umma_commit_el(&w_empty[t % NST]);
        }
        advance_s();
      }
    } else {
      // Two queues, one issuing warp: S(ts) needs its W stage and a free accumulator, U(tu)
      // needs P'(tu).  Neither may hold the other up (a blocking wait for W(t+1) would delay
      // U(t), hence the release of W(t)'s stage, hence the load of W(t+2): loads and tensor
      // work would serialise).  Order constraint: S(t+2) after U(t) - the tensor pipe executes
      // in issue order, so S(t+NSB) then cannot overwrite P'(t) before U(t) has read it.
      int ts = 0, tu = 0;
      while (tu < n_total) {
        bool progressed = false;
        if (tu < ts && mbar_test_all(&p_full[tu % NSB], (uint32_t)((tu / NSB) & 1)) &&
            (iu != 0 || sgu == 0 || mbar_test_all(u_drained, (uint32_t)((sgu - 1) & 1)))) {
          issue_u(tu);
          ++tu;
          if (++iu == len_u) {
            Lu += len_u;
            ++sgu;
            iu = 0;
            len_u = Lu < L_end ? fwd_seg(p, Lu, L_end).len : 0;
          }
          progressed = true;
        }
        if (ts < n_total && ts <= tu + NSB - 1 && s_ready(ts)) {
          issue_s(ts);
          ++ts;
          advance_s();
          progressed = true;
        }
        if (!progressed) __nanosleep(20);
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    const int q = warp & 3;
    const int cg = warp >> 2;
    const int r_in_tile = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t sb = smem_u32(sBias + warp * 32);
    const uint32_t aMax = smem_u32(sMax);
    const int uw = p.h >> 2;  // U columns owned by this warp: [cg*uw, +uw), 32 or 64
    int t = 0;      // tiles processed by this CTA (< 2^31)
    int sg = 0;
    for (long L = L_begin; L < L_end; ++sg) {
      const FwdSeg sgm = fwd_seg(p, L, L_end);
      const int row = sgm.m * VB_M + r_in_tile;
      const int label = row < p.M ? p.labels[row] : -1;
      // statistics in the log2 domain: z2 = (x.w + b) * log2(e); s_run is relative to m_ref
      float m_run = -INFINITY, m_ref = -INFINITY, s_run = 0.f, tgt2 = 0.f;
      bool have_tgt = false;
      auto load_bias = [&](int i) -> float {
        if (i >= sgm.len) return -INFINITY;
        const int v = sgm.vtile(i) * VB_N + cg * 32 + lane;
        return v < p.V ? __ldg(p.bias + v) : -INFINITY;
      };
      float bias_next = load_bias(0);
      for (int i = 0; i < sgm.len; ++i, ++t) {
        const int buf = t % NSB;
        const uint32_t sph = (uint32_t)((t / NSB) & 1);   // accumulator buffer and its barrier phase
        const int vbase = sgm.vtile(i) * VB_N + cg * 32;
        sts32f(sb + lane * 4, bias_next * LOG2E);
        __syncwarp();
        bias_next = load_bias(i + 1);  // in flight while this tile is processed
        mbar_wait(&s_full[buf], sph);
        tc_fence_after();
        const uint32_t tS = tmem_base + lane_base + (uint32_t)(buf * VB_N + cg * 32);
        float z[32];
        // z2 = s * log2(e) + b2 for this thread's 32 scores (packed fp32x2 FMAs), optional maximum
        auto load_scores = [&](float& cmax) {
          uint32_t r[32];
          tmem_ld32(tS, r);
          tmem_ld_wait();
          const float2 l2 = make_float2(LOG2E, LOG2E);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = lds128f(sb + j * 4);
            const float2 za = __ffma2_rn(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])), l2,
                                         make_float2(b4.x, b4.y));
            const float2 zb = __ffma2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), l2,
                                         make_float2(b4.z, b4.w));
            z[j + 0] = za.x; z[j + 1] = za.y; z[j + 2] = zb.x; z[j + 3] = zb.y;
            cmax = fmaxf(cmax, fmaxf(fmaxf(za.x, za.y), fmaxf(zb.x, zb.y)));
          }
        };
        // z <- exp2(z - ref) in place; returns the sum of the 32 values
        auto exponentiate = [&](float ref) -> float {
          const float2 nr = make_float2(-ref, -ref);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float2 da = __fadd2_rn(make_float2(z[j + 0], z[j + 1]), nr);
            const float2 db = __fadd2_rn(make_float2(z[j + 2], z[j + 3]), nr);
            z[j + 0] = ex2(da.x); z[j + 1] = ex2(da.y);
            if (PP > 0 && (j / 4) % (PP > 0 ? PP : 1) == (PP > 0 ? PP : 1) - 1) {   // this pair on the FMA / ALU pipes
              const float2 e = ex2_poly2(db);
              z[j + 2] = e.x; z[j + 3] = e.y;
            } else {
              z[j + 2] = ex2(db.x); z[j + 3] = ex2(db.y);
            }
            acc0 = __fadd2_rn(acc0, make_float2(z[j + 0], z[j + 1]));
            acc1 = __fadd2_rn(acc1, make_float2(z[j + 2], z[j + 3]));
          }
          return (acc0.x + acc0.y) + (acc1.x + acc1.y);
        };
        float cmax = -INFINITY;
        load_scores(cmax);
        if (!OPT || !with_dx) {   // scores are in registers for good: release the accumulator now
          tc_fence_before();
          mbar_arrive_warp(&s_empty[buf]);
        }
        if (label >= vbase && label < vbase + 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (vbase + j == label) tgt2 = z[j];
          have_tgt = true;
        }
        float tile_sum;
        if (!with_dx) {
          // loss only: every thread keeps its own (max, sum) over its column slices
          const float m_new = fmaxf(m_run, cmax);
          const float alpha = m_ref > -INFINITY ? ex2(m_ref - m_new) : 0.f;
          s_run *= alpha;
          m_ref = m_run = m_new;
          tile_sum = exponentiate(m_ref);
        } else {
          // OPTIMISTIC exponentials: P' is taken against the reference maximum agreed on earlier
          // tiles, BEFORE the four warps that share these rows have exchanged this tile's
          // maxima.  The exchange (a 128-thread barrier) used to sit between the TMEM load and the
          // exponentials and re-aligned the four warps of a scheduler on every tile, so the TMEM
          // read phase (64 B/clk: 1,024 clk per tile) and the MUFU phase (1,024 clk) of a tile ran
          // back to back instead of overlapping across warps.  Only when some row's maximum has
          // moved past the reference by more than 2^RESCALE_TH (rare), or on the first tile of a
          // segment (no reference yet), are the scores read from TMEM again and redone.
          const bool first = (i == 0);
          const uint32_t mx = aMax + (uint32_t)((t & 1) * (4 * VB_M) + r_in_tile) * 4;
          sts32f(mx + cg * VB_M * 4, cmax);
          if (OPT && !first) tile_sum = exponentiate(m_ref);
          asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
          cmax = fmaxf(fmaxf(lds32f(mx), lds32f(mx + VB_M * 4)),
                       fmaxf(lds32f(mx + 2 * VB_M * 4), lds32f(mx + 3 * VB_M * 4)));
          const float m_new = fmaxf(m_run, cmax);
          m_run = m_new;
          bool redo = first;
          if (first) {
            m_ref = m_new;  // nothing accumulated yet in this segment
          } else {
            const bool jump = m_new > m_ref + RESCALE_TH;
            if (__any_sync(0xffffffffu, jump)) {
              // U(..t-1) must have retired before its columns are rescaled in TMEM; U(t) cannot
              // start before p_full(t), which this warp only signals after the rescale
              mbar_wait(&u_full[(t - 1) % NSB], (uint32_t)(((t - 1) / NSB) & 1));
              tc_fence_after();
              const float f = jump ? ex2(m_ref - m_new) : 1.f;
#pragma unroll 1
              for (int c = 0; c < uw; c += 8) {  // 8 columns at a time: keeps z[] in registers
                const uint32_t tU = T_U + lane_base + (uint32_t)(cg * uw + c);
                uint32_t u[8];
                tmem_ld8(tU, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = __float_as_uint(__uint_as_float(u[j]) * f);
                tmem_st8(tU, u);
              }
              tmem_st_wait();
              s_run *= f;
              if (jump) m_ref = m_new;
              redo = true;
            }
          }
          if (!OPT) {
            tile_sum = exponentiate(m_ref);   // statistics first, exponentials after the exchange
          } else if (redo) {   // S(t) is still in TMEM (released below): recompute against the new reference
            float unused = -INFINITY;
            load_scores(unused);
            tile_sum = exponentiate(m_ref);
          }
        }
        if (OPT && with_dx) {
          tc_fence_before();
          mbar_arrive_warp(&s_empty[buf]);
        }
        __syncwarp();   // sb is rewritten for the next tile
        s_run += tile_sum;
        if (with_dx) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(z[2 * j], z[2 * j + 1]);
          tmem_st16(tS, pk);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive_warp(&p_full[buf]);
        }
      }
      // ---- end of the segment: one (m_ref, sum, U) partial per row into slot sgm.slot
      if (with_dx && sgm.len > 0) {
        mbar_wait(&u_full[(t - 1) % NSB], (uint32_t)(((t - 1) / NSB) & 1));
        tc_fence_after();
      }
      if (row < p.M) {
        const size_t slot = ((size_t)sgm.slot * 4 + cg) * p.M + row;
        p.part_max[slot] = m_ref;   // log2 domain; the reference the sums are relative to
        p.part_sum[slot] = s_run;
        if (have_tgt) p.tgt[row] = tgt2 * LN2;
      }
      if (with_dx) {
        for (int c = 0; c < uw; c += 32) {
          uint32_t u[32];
          tmem_ld32(T_U + lane_base + (uint32_t)(cg * uw + c), u);
          tmem_ld_wait();
          if (row < p.M) {
            float* dst = p.part_u + ((size_t)sgm.slot * p.M + row) * p.h + cg * uw + c;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) =
                  make_float4(__uint_as_float(u[j]), __uint_as_float(u[j + 1]),
                              __uint_as_float(u[j + 2]), __uint_as_float(u[j + 3]));
          }
        }
        tc_fence_before();
        mbar_arrive_warp(u_drained);   // U may be overwritten by the next segment
      }
      L += sgm.len;
    }
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================== backward
// One CTA owns vocabulary tiles c, c+grid, ... and sweeps every row tile for each of them.
// Shared memory: sW one vocabulary tile [vb 2][hb HB][64 h-rows][128 B]; sX a ring of row tiles
// [hb HB][128 rows][128 B].  TMEM columns: S^T0 [0,128) S^T1 [128,256) dW^T [256, 256+h).
// Epilogue warp (q, cg): lanes = vocabulary entries [32q,+32) of the tile, columns = rows
// [32cg,+32) of the row tile; dZ^T goes back packed into columns [32cg, 32cg+16) of the S^T buffer.
// NG = 2: the 16 epilogue warps form two groups of 8 that take ALTERNATE row tiles (group g owns
// accumulator buffer g; a warp covers 64 columns instead of 32).  With one group all 16 warps wait
// for the same S^T tile, load it, compute and store in lockstep, so nobody issues during the TMEM
// load / store latencies; with two groups one computes while the other waits.
template <int NSB, int NG>
__global__ void __launch_bounds__(TS_THREADS, 1)
vocab_ce_bwd_ts_kernel(const __grid_constant__ CUtensorMap tmX,
                       const __grid_constant__ CUtensorMap tmW, const VocabParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB, h = p.h;
  const int XBUF = p.fwd_stages;
  const int w_bytes = 2 * HB * 8192;
  const int x_bytes = HB * VB_M * 128;
  uint8_t* sW = smem;
  uint8_t* sX = sW + w_bytes;
  float* sNeg = reinterpret_cast<float*>(sX + (size_t)XBUF * x_bytes);  // [16 warps][64]
  int32_t* sLab = reinterpret_cast<int32_t*>(sNeg + 16 * 64);           // [16 warps][64]
  float* sDB = reinterpret_cast<float*>(sLab + 16 * 64);                // [4 cg][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDB + 4 * VB_N);
  uint64_t* w_full = bars;            // 1
  uint64_t* w_empty = bars + 1;       // 1
  uint64_t* x_full = bars + 2;        // 4
  uint64_t* x_empty = x_full + 4;     // 4
  uint64_t* s_full = x_empty + 4;     // 3
  uint64_t* s_empty = s_full + 3;     // 3
  uint64_t* dz_full = s_empty + 3;    // 3
  uint64_t* dw_full = dz_full + 3;    // 1
  uint64_t* dw_empty = dw_full + 1;   // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dw_empty + 1);

  // warp index through shfl: provably warp-uniform, so the role branches are uniform control flow
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nm = p.n_mtiles;
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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], NUM_EPI_WARPS / NG);
      mbar_init(&dz_full[i], NUM_EPI_WARPS / NG);
    }
    mbar_init(dw_full, 1);
    mbar_init(dw_empty, NUM_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t T_S = tmem_base, T_DW = tmem_base + (uint32_t)(NSB * VB_M);

  if (warp == WARP_TMA) {
    // ------------------------------------------------------------------ TMA producer (whole
    // warp, uniform control flow; TMA instructions predicated on an elected lane)
    const uint32_t aW = smem_u32(sW), aX0 = smem_u32(sX);
    long it = 0;
    for (int vt = 0; vt < n_my; ++vt) {
      const int v0 = ((int)blockIdx.x + vt * (int)gridDim.x) * VB_N;
      mbar_wait_all(w_empty, (vt & 1) ^ 1);
      mbar_expect_tx_el(w_full, (uint32_t)w_bytes);
      for (int vb = 0; vb < 2; ++vb)
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d_el(aW + (vb * HB + hb) * 8192, &tmW, w_full, v0 + vb * 64, hb * 64);
      for (int i = 0; i < nm; ++i, ++it) {
        const int xb = (int)(it % XBUF);
        mbar_wait_all(&x_empty[xb], (uint32_t)((it / XBUF) & 1) ^ 1);
        mbar_expect_tx_el(&x_full[xb], (uint32_t)x_bytes);
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d_el(aX0 + (uint32_t)(xb * x_bytes) + hb * (VB_M * 128), &tmX, &x_full[xb],
                         hb * 64, i * VB_M);
      }
    }
  } else if (warp == WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer (whole warp,
    // uniform control flow; tcgen05 instructions predicated on an elected lane)
    const uint32_t id_s = umma_idesc_bf16(VB_N, VB_M, 1, 0);  // S^T = W^T (MN-major) X^T (K-major)
    const uint32_t id_dw = umma_idesc_bf16(VB_N, h, 0, 1);    // dW^T = dZ^T (TMEM) X (MN-major)
    const uint32_t aX0 = smem_u32(sX);
    const uint64_t dWa = umma_smem_desc(smem_u32(sW), HB * 8192, 1024);   // W^T, MN-major (A of S^T)
    const uint64_t dXk = umma_smem_desc(aX0, 16, 1024);                   // X, K-major (B of S^T)
    const uint64_t dXn = umma_smem_desc(aX0, VB_M * 128, 1024);           // X, MN-major (B of dW^T)
    auto s_ready = [&](long it) -> bool {
      return mbar_test_all(&x_full[it % XBUF], (uint32_t)((it / XBUF) & 1)) &&
             mbar_test_all(&s_empty[it % NSB], (uint32_t)((it / NSB) & 1) ^ 1);
    };
    auto issue_s = [&](long it) {
      const int xb = (int)(it % XBUF), sb = (int)(it % NSB);
      tc_fence_after();
      const uint32_t xo = (uint32_t)(xb * x_bytes) >> 4;
      for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_el(T_S + sb * VB_M, dWa + (uint32_t)((hb * 8192 + kk * 2048) >> 4),
                       dXk + (xo + (uint32_t)((hb * (VB_M * 128) + kk * 32) >> 4)), id_s,
                       (hb | kk) ? 1u : 0u);
      }
      umma_commit_el(&s_full[sb]);
    };
    // dW^T[128 v x h] (+)= dZ^T[128 v x 128 rows] X[128 rows x h]: 8 K steps of 16 rows
    auto issue_dw = [&](long it, int i) {
      const int xb = (int)(it % XBUF), zb = (int)(it % NSB);
      tc_fence_after();
      const uint32_t xo = (uint32_t)(xb * x_bytes) >> 4;
      const uint32_t tZ = T_S + zb * VB_M;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8)
        umma_bf16_ts_el(T_DW, tZ + (k8 >> 1) * 32 + (k8 & 1) * 8,
                        dXn + (xo + (uint32_t)((k8 * 2048) >> 4)), id_dw, (i | k8) ? 1u : 0u);
      umma_commit_el(&x_empty[xb]);
    };
    long it0 = 0;
    for (int vt = 0; vt < n_my; ++vt, it0 += nm) {
      mbar_wait_all(w_full, vt & 1);
      // two queues (see the forward kernel): S^T(is) needs its X tile and a free accumulator,
      // dW(iu) needs dZ^T(iu); S^T(i+2) is only issued after dW(i)
      int is = 0, iu = 0;
      while (iu < nm) {
        bool progressed = false;
        if (iu < is && mbar_test_all(&dz_full[(it0 + iu) % NSB], (uint32_t)(((it0 + iu) / NSB) & 1))) {
          if (iu == 0) mbar_wait_all(dw_empty, (vt & 1) ^ 1);
          issue_dw(it0 + iu, iu);
          ++iu;
          progressed = true;
        }
        if (is < nm && is <= iu + NSB - 1 && s_ready(it0 + is)) {
          issue_s(it0 + is);
          ++is;
          progressed = true;
        }
        if (!progressed) __nanosleep(20);
      }
      umma_commit_el(dw_full);
      umma_commit_el(w_empty);
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue (16 warps)
    const int q = warp & 3;
    const int cg = warp >> 2;
    static_assert(NG == 1 || NSB == 2, "two epilogue groups own one accumulator buffer each");
    const int v_local = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t aNeg = smem_u32(sNeg + warp * 64);
    const uint32_t aLab = smem_u32(sLab + warp * 64);
    constexpr int CW = 32 * NG;                           // columns (rows of the row tile) per warp
    const int grp = NG == 2 ? (cg & 1) : 0;               // epilogue group
    const int cb = NG == 2 ? (cg >> 1) * CW : cg * CW;    // first column of this warp
    const int uw = h >> 2;
    const float n_valid = p.loss_stats[1];
    const float inv_n = n_valid > 0.f ? 1.f / n_valid : 0.f;
    const float log2_inv_n = n_valid > 0.f ? -log2f(n_valid) : 0.f;
    long it0 = 0;
    for (int vt = 0; vt < n_my; ++vt, it0 += nm) {
      const int v0 = ((int)blockIdx.x + vt * (int)gridDim.x) * VB_N;
      const int v = v0 + v_local;
      const float b2 = v < p.V ? __ldg(p.bias + v) * LOG2E : -INFINITY;
      float db_acc = 0.f;
      // this group's first row tile of the vocabulary tile, then every NG-th
      const int i_first = NG == 2 ? (int)((it0 & 1) != grp) : 0;
      // statistics of rows (cb + lane [+ 32]) of the group's next row tile, prefetched a tile ahead
      int label_next[NG];
      float lse_next[NG];
#pragma unroll
      for (int u = 0; u < NG; ++u) {
        const int r0 = i_first * VB_M + cb + 32 * u + lane;
        const bool ok = i_first < nm && r0 < p.M;
        label_next[u] = ok ? __ldg(p.labels + r0) : -1;
        lse_next[u] = ok ? __ldg(p.lse + r0) : 0.f;
      }
      for (int i = i_first; i < nm; i += NG) {
        const long it = it0 + i;
        const int sbuf = (int)(it % NSB);
        // dZ = exp2(z2 - lse2 - log2 n): the 1/n_valid factor rides in the exponent; -inf for
        // padded rows and rows past M, which then contribute exactly 0
        bool hit_l = false;
#pragma unroll
        for (int u = 0; u < NG; ++u) {
          const float lneg = (label_next[u] >= 0 && n_valid > 0.f) ? log2_inv_n - lse_next[u] * LOG2E
                                                                    : -INFINITY;
          sts32f(aNeg + (32 * u + lane) * 4, lneg);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(aLab + (32 * u + lane) * 4), "r"(label_next[u]) : "memory");
          hit_l = hit_l || (unsigned)(label_next[u] - v0) < (unsigned)VB_N;
        }
        const bool hit = __any_sync(0xffffffffu, hit_l);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NG; ++u) {
          const int nrow = (i + NG) * VB_M + cb + 32 * u + lane;
          const bool ok = (i + NG < nm) && nrow < p.M;
          label_next[u] = ok ? __ldg(p.labels + nrow) : -1;
          lse_next[u] = ok ? __ldg(p.lse + nrow) : 0.f;
        }
        mbar_wait(&s_full[sbuf], (uint32_t)((it / NSB) & 1));
        tc_fence_after();
#pragma unroll
        for (int u = 0; u < NG; ++u) {
          const uint32_t tS = T_S + lane_base + (uint32_t)(sbuf * VB_M + cb + 32 * u);
          uint32_t r[32];
          tmem_ld32(tS, r);
          tmem_ld_wait();
          if (u == NG - 1) {   // the whole accumulator slice of this warp is in registers
            tc_fence_before();
            mbar_arrive_warp(&s_empty[sbuf]);
          }
          float g[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 n4 = lds128f(aNeg + (32 * u + j) * 4);
            g[j + 0] = ex2(fmaf(__uint_as_float(r[j + 0]), LOG2E, b2 + n4.x));
            g[j + 1] = ex2(fmaf(__uint_as_float(r[j + 1]), LOG2E, b2 + n4.y));
            g[j + 2] = ex2(fmaf(__uint_as_float(r[j + 2]), LOG2E, b2 + n4.z));
            g[j + 3] = ex2(fmaf(__uint_as_float(r[j + 3]), LOG2E, b2 + n4.w));
          }
          if (hit) {  // some row of this warp's group has its label inside this vocabulary tile
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              int lab;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(lab) : "r"(aLab + (32 * u + j) * 4) : "memory");
              if (lab == v) g[j] -= inv_n;
            }
          }
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 += g[j + 0];
            a1 += g[j + 1];
            a2 += g[j + 2];
            a3 += g[j + 3];
            pk[j / 2] = pack_bf16x2(g[j], g[j + 1]);
            pk[j / 2 + 1] = pack_bf16x2(g[j + 2], g[j + 3]);
          }
          db_acc += (a0 + a1) + (a2 + a3);
          tmem_st16(tS, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_warp(&dz_full[sbuf]);  // (its __syncwarp also orders the sNeg/sLab reuse)
      }
      // bias gradient: the 4 warps that share these vocabulary entries combine their row groups
      sDB[cg * VB_N + v_local] = db_acc;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
      if (cg == 0 && v < p.V)
        p.db[v] = (sDB[v_local] + sDB[VB_N + v_local]) + (sDB[2 * VB_N + v_local] + sDB[3 * VB_N + v_local]);
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
      // dW^T tile: lane = vocabulary entry, columns = input features -> each store instruction of
      // a warp writes 32 consecutive floats of one dW row
      mbar_wait(dw_full, vt & 1);
      tc_fence_after();
      for (int c = 0; c < uw; c += 32) {
        uint32_t r[32];
        tmem_ld32(T_DW + lane_base + (uint32_t)(cg * uw + c), r);
        tmem_ld_wait();
        if (v < p.V) {
          float* dst = p.dW + (size_t)(cg * uw + c) * p.V + v;
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)j * p.V] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      mbar_arrive_warp(dw_empty);
    }
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------- launchers
static size_t fwd_ts_smem(int HB, int stages, int xbufs) {
  return (size_t)xbufs * HB * VB_M * 128 + (size_t)stages * 2 * HB * 8192 + 16 * 32 * 4 +
         2 * 4 * VB_M * 4 + 256 + 1024;
}

int launch_vocab_fwd_ts(const CUtensorMap& tmX, const CUtensorMap& tmW, const VocabParams& p_in,
                        cudaStream_t st) {
  VocabParams p = p_in;
  // a CTA with several segments double-buffers the X tile when shared memory allows (h <= 128)
  p.x_bufs = (p.sched == SCHED_RANGES && p.h <= 128) ? 2 : 1;
  int stages = 4;
  while (stages > 2 && fwd_ts_smem(p.HB, stages, p.x_bufs) > 227 * 1024) --stages;
  B4CP_CHECK_ARG(fwd_ts_smem(p.HB, stages, p.x_bufs) <= 227 * 1024, "vocab_ce_fwd: h=%d does not fit", p.h);
  p.fwd_stages = stages;
  const size_t smem = fwd_ts_smem(p.HB, stages, p.x_bufs);
  dim3 grid(p.n_mtiles, p.n_chunks);
  if (p.sched == SCHED_RANGES) grid = dim3((unsigned)ceil_div(p.total_tiles, p.range_q), 1);
  // L2 prefetch of W tiles ahead of their TMA load: measured useless (one CTA per chunk prefetching:
  // 6.82 vs 6.88 ms at C4) to harmful (every CTA: 10.1 ms) - the exposed latency at h = 256 is the
  // L2 -> SM transfer of a stage, not an HBM miss.  Off; the switches remain for experiments.
  p.l2_prefetch = 0;
  p.l2_prefetch_every = 64;
  p.grid_stagger = 0;   // measured: 0 / 1 / 3 / 9 tiles all 7.1-7.2 ms at C4, 27 tiles 8.5 ms (HBM re-reads)
  if (const char* e = getenv("B4CP_FWD_STAGGER")) p.grid_stagger = std::max(0, atoi(e));   // developer switch
  if (const char* e = getenv("B4CP_FWD_PREFETCH")) p.l2_prefetch = atoi(e);   // developer switches
  if (const char* e = getenv("B4CP_FWD_PREFETCH_EVERY")) p.l2_prefetch_every = std::max(1, atoi(e));
  int pp = FWD_POLY_DEFAULT;
  if (const char* e = getenv("B4CP_FWD_POLY")) pp = atoi(e);   // developer switch: 0 / 2
  // optimistic exponentials: 0.88 -> 0.837 ms on their own at the C1 shape, but once a
  // quarter of the exponentials is off the MUFU unit the plain order is as fast (0.78 vs 0.80 ms
  // at C1, 6.81 vs 6.85 ms at C4) and releases the S accumulator earlier: off by default
  bool opt = false;
  if (const char* e = getenv("B4CP_FWD_OPT")) opt = atoi(e) != 0;   // developer switch
  // K-split half stages: 0.385 -> 0.339 ms at (M 7,424, V 54,293, h 256) where W sits in L2; no
  // gain under the lock-step grid at V = 1M (7.1 vs 6.9 ms), so only with the range schedule
  const bool ksplit = p.HB == 4 && stages == 2 && p.sched == SCHED_RANGES && !getenv("B4CP_FWD_NO_KSPLIT");
#define B4CP_LAUNCH_FWD_K(NSB_, PP_, OPT_, KS_)                                                     \
  do {                                                                                              \
    B4CP_CUDA(cudaFuncSetAttribute(vocab_ce_fwd_ts_kernel<NSB_, PP_, OPT_, KS_>,                    \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));       \
    vocab_ce_fwd_ts_kernel<NSB_, PP_, OPT_, KS_><<<grid, TS_THREADS, smem, st>>>(tmX, tmW, p);      \
  } while (0)
#define B4CP_LAUNCH_FWD(NSB_, PP_, OPT_)                                                            \
  do {                                                                                              \
    if (NSB_ == 2 && ksplit) B4CP_LAUNCH_FWD_K(NSB_, PP_, OPT_, (NSB_ == 2));                       \
    else B4CP_LAUNCH_FWD_K(NSB_, PP_, OPT_, false);                                                 \
  } while (0)
  if (p.h <= 128) {
    if (pp == 2 && opt) B4CP_LAUNCH_FWD(3, 2, true);
    else if (pp == 2) B4CP_LAUNCH_FWD(3, 2, false);
    else if (opt) B4CP_LAUNCH_FWD(3, 0, true);
    else B4CP_LAUNCH_FWD(3, 0, false);
  } else {
    if (pp == 2 && opt) B4CP_LAUNCH_FWD(2, 2, true);
    else if (pp == 2) B4CP_LAUNCH_FWD(2, 2, false);
    else if (opt) B4CP_LAUNCH_FWD(2, 0, true);
    else B4CP_LAUNCH_FWD(2, 0, false);
  }
#undef B4CP_LAUNCH_FWD
#undef B4CP_LAUNCH_FWD_K
  return 0;
}

static size_t bwd_ts_smem(int HB, int xbuf) {
  return (size_t)2 * HB * 8192 + (size_t)xbuf * HB * VB_M * 128 + 2 * 16 * 64 * 4 + 4 * VB_N * 4 +
         256 + 1024;
}

int launch_vocab_bwd_ts(const CUtensorMap& tmX, const CUtensorMap& tmW, const VocabParams& p_in,
                        cudaStream_t st) {
  VocabParams p = p_in;
  int xbuf = 4;
  while (xbuf > 2 && bwd_ts_smem(p.HB, xbuf) > 227 * 1024) --xbuf;
  B4CP_CHECK_ARG(bwd_ts_smem(p.HB, xbuf) <= 227 * 1024, "vocab_ce_bwd: h=%d does not fit", p.h);
  p.fwd_stages = xbuf;
  const int grid = std::min(148, p.n_vtiles);
  // two S^T accumulators: a third one (possible at h <= 128) measured slower here (0.84 vs 0.78 ms
  // at the bench shape) - the backward's epilogue is shorter than its two MMAs
  const char* pp = getenv("B4CP_BWD_GROUPS");
  if (pp && pp[0] == '1') {
    B4CP_CUDA(cudaFuncSetAttribute(vocab_ce_bwd_ts_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    vocab_ce_bwd_ts_kernel<2, 1><<<grid, TS_THREADS, bwd_ts_smem(p.HB, xbuf), st>>>(tmX, tmW, p);
  } else {
    B4CP_CUDA(cudaFuncSetAttribute(vocab_ce_bwd_ts_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    vocab_ce_bwd_ts_kernel<2, 2><<<grid, TS_THREADS, bwd_ts_smem(p.HB, xbuf), st>>>(tmX, tmW, p);
  }
  return 0;
}

}  // namespace b4cp
