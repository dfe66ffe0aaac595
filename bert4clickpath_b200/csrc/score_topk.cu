// Next-item scoring fused with an exact top-k (inference / evaluation path):
// SoftMaxHead's Dense(V) (head.py:36,45) followed by tf.math.top_k over the whole vocabulary
// (examples/BERT4Rec/source/utils.py:176, :245), without writing the (rows x V) score matrix.
//
// Ranking by logits x W + b is the ranking by softmax probabilities (monotone), with the
// tf.math.top_k tie rule: equal scores -> lower id first.
//
// grid = (row tiles of 128, vocabulary chunks).  Per CTA: X tile resident, W tiles streamed by TMA,
// S = X W on tcgen05 into a double-buffered TMEM accumulator; each of the 128 epilogue threads owns
// one row and keeps that row's k best (score, id) pairs of the chunk in a private binary min-heap
// laid out [slot][row] in shared memory (bank = row % 32: conflict-free for any access pattern).
// A candidate only touches the heap when it beats the current k-th best, which after the first few
// tiles happens ~k ln(V/k) times per row.  Per-chunk heaps go to a workspace; the exact merge
// (score desc, id asc) is done by topk_candidates_kernel (radix select on the (score, id) key).
//
// h = 256 (head [] -> V on d_model = 256) uses the WIDE instantiation: the X tile
// lives in TMEM (every epilogue thread loads its own row and writes it with tcgen05.st; the
// score MMA takes A from TMEM), shared memory holds two 64 KB W stages, and the heap ids are
// stored as 24 bits (u16 + u8 arrays) so that k = 100 still fits beside them.
#include <algorithm>
#include <climits>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

static constexpr int ST_M = 128, ST_N = 128;
static constexpr int ST_STAGES = 2;
static constexpr int ST_MAXK = 104;

struct ScoreParams {
  int M, V, h, HB, k;
  int n_vtiles, tiles_per_chunk, n_chunks;
  int id_base;  // added to every reported id (vocabulary shards)
  const float* bias;
  float* part_scores;  // [M][n_chunks][k]
  int32_t* part_ids;   // [M][n_chunks][k]
  int tile0;           // first vocabulary tile this launch covers (filter mode: after the seed)
  // filter mode: candidates (score > row threshold) are appended to per-row lists in HBM
  float* cand_scores;  // [M][cap]; slots [0,k) hold the seed's top-k, cand_scores[.][k-1] = threshold
  int32_t* cand_ids;   // [M][cap]
  int* cand_cnt;       // [M] appended so far
  int cap;
  const int* run_if;   // launch is a no-op unless *run_if != 0 (NULL: always run)
  int l2_prefetch;     // filter kernel: W tiles prefetched into L2 this many tiles ahead (0 = off)
};

static constexpr int ID24_EMPTY = 0xFFFFFF;

// ---- rare-path helpers, deliberately NOT inlined: the scan over a 128-wide score tile is fully
// unrolled (the accumulators are a register array), and an inlined heap update / append in each
// of its 32 copies blew the loop up to ~100 KB of code - every tile then streamed its
// instructions from L2 (ncu: stall_no_inst dominant, ~6 us per tile against 0.5 us of MMAs).
template <bool WIDE>
__device__ __forceinline__ int heap_ld_id(uint32_t aI, uint32_t aH, int slot) {
  int v;
  if (WIDE) {
    uint32_t lo, hi;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(aI + slot * (ST_M * 2)) : "memory");
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(hi) : "r"(aH + slot * ST_M) : "memory");
    v = (int)(lo | (hi << 16));
  } else {
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(aI + slot * (ST_M * 4)) : "memory");
  }
  return v;
}
template <bool WIDE>
__device__ __forceinline__ void heap_st_id(uint32_t aI, uint32_t aH, int slot, int v) {
  if (WIDE) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(aI + slot * (ST_M * 2)), "r"(v & 0xFFFF) : "memory");
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(aH + slot * ST_M), "r"((v >> 16) & 0xFF) : "memory");
  } else {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(aI + slot * (ST_M * 4)), "r"(v) : "memory");
  }
}
// replace the root (current k-th best) by (sc, id) and sift down; "worse" = lower score, or equal
// score and higher id.  Returns the new root score.
template <bool WIDE>
__device__ __noinline__ float heap_replace_root(uint32_t aS, uint32_t aI, uint32_t aH, int k, float sc,
                                                int id) {
  constexpr uint32_t SLOT = ST_M * 4;
  int pos = 0;
  while (true) {
    const int l = 2 * pos + 1, rr = l + 1;
    if (l >= k) break;
    float cs = lds32f(aS + l * SLOT);
    int ci = heap_ld_id<WIDE>(aI, aH, l), cpos = l;
    if (rr < k) {
      const float rs = lds32f(aS + rr * SLOT);
      const int ri = heap_ld_id<WIDE>(aI, aH, rr);
      if (rs < cs || (rs == cs && ri > ci)) {
        cs = rs;
        ci = ri;
        cpos = rr;
      }
    }
    if (!(cs < sc || (cs == sc && ci > id))) break;   // not better than the worse child: stop
    sts32f(aS + pos * SLOT, cs);
    heap_st_id<WIDE>(aI, aH, pos, ci);
    pos = cpos;
  }
  sts32f(aS + pos * SLOT, sc);
  heap_st_id<WIDE>(aI, aH, pos, id);
  return lds32f(aS);
}
// FILTER mode: stage (sc, id) in shared memory ([slot][row]); every FSTAGE candidates (or with
// force) one global atomic reserves their slots in the row's list.  Returns the new staged count.
static constexpr int FSTAGE = 16;
__device__ __noinline__ int filter_stage(uint32_t aS, uint32_t aI, int n_staged, float sc, int id,
                                         bool force, int* cnt, float* cand_scores, int32_t* cand_ids,
                                         int k, int cap) {
  if (!force) {
    sts32f(aS + (uint32_t)(n_staged * ST_M) * 4, sc);
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(aI + (uint32_t)(n_staged * ST_M) * 4), "r"(id) : "memory");
    if (++n_staged < FSTAGE) return n_staged;
  }
  const int base = k + atomicAdd(cnt, n_staged);
  for (int i = 0; i < n_staged; ++i) {
    int sid;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(sid) : "r"(aI + (uint32_t)(i * ST_M) * 4) : "memory");
    if (base + i < cap) {   // (an overflowing row is detected from its counter and redone)
      cand_scores[base + i] = lds32f(aS + (uint32_t)(i * ST_M) * 4);
      cand_ids[base + i] = sid;
    }
  }
  return 0;
}

// FILTER = true is the long-vocabulary variant: no heap; every score above the row's threshold (the
// k-th best of a seed range ranked beforehand) is appended to the row's candidate list in HBM.
template <bool WIDE, bool FILTER>
__global__ void __launch_bounds__(192, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const ScoreParams p, const __nv_bfloat16* __restrict__ x_rows, long ldx) {
  if (p.run_if && *p.run_if == 0) return;   // uniform: fallback launch that is not needed
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB, k = p.k;
  const int x_bytes = WIDE ? 0 : HB * ST_M * 128;   // WIDE: X lives in TMEM
  const int w_bytes = 2 * HB * 64 * 128;
  uint8_t* sX = smem;
  uint8_t* sW = sX + x_bytes;
  float* sHeapS = reinterpret_cast<float*>(sW + (size_t)ST_STAGES * w_bytes);  // [k][128]
  // ids: int32 [k][128], or (WIDE) u16 low halves [k][128] followed by u8 high bytes [k][128]
  // FILTER: no heap; the same two arrays are a [16][128] staging area for appended candidates
  int32_t* sHeapI = reinterpret_cast<int32_t*>(sHeapS + (FILTER ? (size_t)16 * ST_M : (size_t)k * ST_M));
  float* sBias = reinterpret_cast<float*>(
      reinterpret_cast<uint8_t*>(sHeapI) +
      (FILTER ? (size_t)16 * ST_M * 4 : (((size_t)k * ST_M * (WIDE ? 3 : 4) + 15) & ~(size_t)15)));  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 4 * 2 * ST_N);   // sBias: [4 warps][2][128]
  uint64_t* x_full = bars;
  uint64_t* w_full = bars + 1;
  uint64_t* w_empty = w_full + ST_STAGES;
  uint64_t* s_full = w_empty + ST_STAGES;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);
  constexpr int WARP_TMA = 4, WARP_MMA = 5;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * ST_M;
  const int chunk = blockIdx.y;
  const int t_begin = chunk * p.tiles_per_chunk;
  const int t_end = min(p.n_vtiles, t_begin + p.tiles_per_chunk);
  const int ntiles = t_end - t_begin;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmW);
    }
    tmem_alloc(tmem_slot, WIDE ? 512 : 256);
    tmem_relinquish();
  } else if (warp == WARP_MMA && lane == 0) {
    mbar_init(x_full, WIDE ? 4 : 1);   // WIDE: the 4 epilogue warps write X into TMEM
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 4);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == WARP_TMA) {
    {  // whole warp, uniform control flow; TMA instructions predicated on an elected lane
      if (!WIDE) {
        mbar_expect_tx_el(x_full, (uint32_t)x_bytes);
        for (int hb = 0; hb < HB; ++hb) tma_load_2d_el(smem_u32(sX + hb * (ST_M * 128)), &tmX, x_full, hb * 64, m0);
      }
      for (int t = 0; t < ntiles; ++t) {
        const int st = t % ST_STAGES;
        mbar_wait_all(&w_empty[st], ((t / ST_STAGES) & 1) ^ 1);
        mbar_expect_tx_el(&w_full[st], (uint32_t)w_bytes);
        const int v0 = (p.tile0 + t_begin + t) * ST_N;
        uint8_t* dst = sW + (size_t)st * w_bytes;
        for (int vb = 0; vb < 2; ++vb)
          for (int hb = 0; hb < HB; ++hb)
            tma_load_2d_el(smem_u32(dst + (vb * HB + hb) * 8192), &tmW, &w_full[st], v0 + vb * 64, hb * 64);
      }
    }
  } else if (warp == WARP_MMA) {
    {  // the WHOLE warp issues, in uniform control flow (see umma_bf16_el in common.cuh)
      const uint32_t idesc = umma_idesc_bf16(ST_M, ST_N, 0, 1);
      const uint32_t aX = smem_u32(sX);
      mbar_wait_all(x_full, 0);
      for (int t = 0; t < ntiles; ++t) {
        const int st = t % ST_STAGES, buf = t & 1;
        mbar_wait_all(&w_full[st], (t / ST_STAGES) & 1);
        mbar_wait_all(&s_empty[buf], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t aW = smem_u32(sW + (size_t)st * w_bytes);
        for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t db = umma_smem_desc(aW + hb * 8192 + kk * 2048, HB * 8192, 1024);
            if (WIDE) {  // A = X from TMEM: K step (hb, kk) = 16 bf16 = 8 packed columns
              umma_bf16_ts_el(tmem_base + buf * ST_N, tmem_base + 256 + (hb * 4 + kk) * 8, db, idesc,
                              (hb | kk) ? 1u : 0u);
            } else {
              const uint64_t da = umma_smem_desc(aX + hb * (ST_M * 128) + kk * 32, 16, 1024);
              umma_bf16_el(tmem_base + buf * ST_N, da, db, idesc, (hb | kk) ? 1u : 0u);
            }
          }
        }
        umma_commit_el(&w_empty[st]);
        umma_commit_el(&s_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = row
    const int r = warp * 32 + lane;
    if (WIDE) {
      // this thread's row of X (h = 256 bf16 = 128 packed columns) -> TMEM lanes of its warp
      const bool live = m0 + r < p.M;
      const uint4* src = reinterpret_cast<const uint4*>(x_rows + (size_t)(m0 + r) * ldx);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t w[32];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const uint4 v = live ? __ldg(src + c * 8 + q4) : make_uint4(0u, 0u, 0u, 0u);
          w[4 * q4] = v.x; w[4 * q4 + 1] = v.y; w[4 * q4 + 2] = v.z; w[4 * q4 + 3] = v.w;
        }
        tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(256 + c * 32), w);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_warp(x_full);
    }
    const uint32_t aS = smem_u32(sHeapS) + r * 4;
    const uint32_t aI = smem_u32(sHeapI) + r * (WIDE ? 2 : 4);                      // int32 or u16 low
    const uint32_t aH = smem_u32(sHeapI) + (uint32_t)k * ST_M * 2 + r;              // u8 high (WIDE)
    const uint32_t aB = smem_u32(sBias + warp * 2 * ST_N);   // this warp's private copy: no CTA-wide barrier per tile
    constexpr uint32_t SLOT = ST_M * 4;  // byte stride between heap score slots
    constexpr int EMPTY = WIDE ? ID24_EMPTY : INT_MAX;
    auto ld_id = [&](int slot) -> int {
      int v;
      if (WIDE) {
        uint32_t lo, hi;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(lo) : "r"(aI + slot * (ST_M * 2)) : "memory");
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(hi) : "r"(aH + slot * ST_M) : "memory");
        v = (int)(lo | (hi << 16));
      } else {
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(aI + slot * SLOT) : "memory");
      }
      return v;
    };
    auto st_id = [&](int slot, int v) {
      if (WIDE) {
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(aI + slot * (ST_M * 2)), "r"(v & 0xFFFF) : "memory");
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(aH + slot * ST_M), "r"((v >> 16) & 0xFF) : "memory");
      } else {
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(aI + slot * SLOT), "r"(v) : "memory");
      }
    };
    if (!FILTER) {
      for (int s = 0; s < k; ++s) {      // empty heap: k sentinels worse than any real score
        sts32f(aS + s * SLOT, -INFINITY);
        st_id(s, EMPTY);
      }
    }
    // FILTER: the threshold is the seed's k-th best score; a later id with an equal score loses the
    // tie (ids only grow), so "strictly greater" is exact.  Rows past M never pass.
    const long crow = (long)(m0 + r) * p.cap;
    float root_s = -INFINITY;
    if (FILTER) root_s = m0 + r < p.M ? p.cand_scores[crow + k - 1] : INFINITY;
    int n_staged = 0;   // FILTER: candidates staged in sHeapS / sHeapI ([slot][row])
    // bias of a tile (one value per thread, -inf past the vocabulary), fetched one tile ahead:
    // a global load per tile in front of the barrier below stalled every tile for a DRAM latency
    // lane l stages columns 4l..4l+3 of the tile's bias for ITS warp (every thread needs all 128)
    auto load_bias = [&](int t) -> float4 {
      const int v = (p.tile0 + t_begin + t) * ST_N + lane * 4;
      float4 b = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (t < ntiles) {
        if (v + 3 < p.V && (((uintptr_t)(p.bias + v)) & 15) == 0) {
          b = __ldg(reinterpret_cast<const float4*>(p.bias + v));
        } else {
          if (v < p.V) b.x = __ldg(p.bias + v);
          if (v + 1 < p.V) b.y = __ldg(p.bias + v + 1);
          if (v + 2 < p.V) b.z = __ldg(p.bias + v + 2);
          if (v + 3 < p.V) b.w = __ldg(p.bias + v + 3);
        }
      }
      return b;
    };
    float4 bias_next = load_bias(0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const int v0 = (p.tile0 + t_begin + t) * ST_N;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aB + (buf * ST_N + lane * 4) * 4),
                   "f"(bias_next.x), "f"(bias_next.y), "f"(bias_next.z), "f"(bias_next.w) : "memory");
      bias_next = load_bias(t + 1);
      __syncwarp();
      mbar_wait(&s_full[buf], (t >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < ST_N / 32; ++c) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * ST_N + c * 32), acc);
        tmem_ld_wait();
        if (c == ST_N / 32 - 1) {
          tc_fence_before();
          mbar_arrive_warp(&s_empty[buf]);
        }
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const float4 b4 = lds128f(aB + (buf * ST_N + c * 32 + j4) * 4);
          const float bj[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float sc = __uint_as_float(acc[j4 + u]) + bj[u];
            if (sc > root_s) {
              const int id = v0 + c * 32 + j4 + u;
              if (FILTER)
                n_staged = filter_stage(aS, aI, n_staged, sc, id, false, p.cand_cnt + m0 + r,
                                        p.cand_scores + crow, p.cand_ids + crow, k, p.cap);
              else
                root_s = heap_replace_root<WIDE>(aS, aI, aH, k, sc, id);
            }
          }
        }
      }
    }
    if (FILTER && n_staged > 0)
      filter_stage(aS, aI, n_staged, 0.f, 0, true, p.cand_cnt + m0 + r, p.cand_scores + crow,
                   p.cand_ids + crow, k, p.cap);
    const int row = m0 + r;
    if (!FILTER && row < p.M) {
      float* os = p.part_scores + ((size_t)row * p.n_chunks + chunk) * k;  // [row][chunk][k]
      int32_t* oi = p.part_ids + ((size_t)row * p.n_chunks + chunk) * k;
      for (int s = 0; s < k; ++s) {
        const int id = ld_id(s);
        os[s] = lds32f(aS + s * SLOT);
        oi[s] = id == EMPTY ? -1 : id + p.id_base;
      }
    }
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WIDE ? 512 : 256);
  }
}

// ---------------------------------------------------------------------------------------------
// FILTER kernel of the long-vocabulary path, second generation.  Same job as
// score_topk_kernel<., true> (append every score above the row's seed threshold to the row's
// candidate list), built like the fused vocabulary forward: 16 epilogue warps (warp (q, cg) owns
// TMEM lanes [32q, +32) x columns [32cg, +32) of a 128 x 128 score tile, 4 per scheduler), one TMA
// warp, one MMA warp, FOUR score accumulators in TMEM (no second product here, so all 512 columns
// hold scores and the tensor pipe runs three tiles ahead of the epilogue).  The common case of a
// thread - none of its 32 scores beats the threshold - is 16 packed adds, a max tree and one warp
// vote; survivors (~k V / seed per row over the whole sweep) go straight to the row's list with
// one global atomic each.  The first-generation kernel walked a row's 128 scores in ONE thread
// (4 epilogue warps) and measured 10 % tensor-pipe activity.
static constexpr int SF_THREADS = 576, SF_NSB = 4, SF_EPI_WARPS = 16;
static constexpr int SF_Q = 6;   // survivors a thread stages in shared memory before one global atomic

// Stage (sc, id) in this thread's shared-memory queue ([slot][thread]: conflict-free); when the
// queue is full (or with force) ONE global atomic reserves the slots in the row's list and the
// staged pairs are written out.  Not inlined: the caller tests 32 scores per tile in unrolled
// code, and survivors are rare (~k V / seed per row over the whole sweep).  Returns the new count.
__device__ __noinline__ int filter_stage16(uint32_t aQ, int n_staged, float sc, int id, bool force,
                                           int* cnt, float* cand_scores, int32_t* cand_ids, int k,
                                           int cap) {
  constexpr uint32_t SLOT = SF_EPI_WARPS * 32 * 8;   // bytes between queue slots
  if (!force) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(aQ + (uint32_t)n_staged * SLOT),
                 "r"(__float_as_uint(sc)), "r"(id) : "memory");
    if (++n_staged < SF_Q) return n_staged;
  }
  const int base = k + atomicAdd(cnt, n_staged);
  for (int i = 0; i < n_staged; ++i) {
    const uint2 e = lds64(aQ + (uint32_t)i * SLOT);
    if (base + i < cap) {   // (an overflowing row is detected from its counter and redone)
      cand_scores[base + i] = __uint_as_float(e.x);
      cand_ids[base + i] = (int32_t)e.y;
    }
  }
  return 0;
}

__global__ void __launch_bounds__(SF_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const ScoreParams p, int n_stages) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB, k = p.k, NST = n_stages;
  const int x_bytes = HB * ST_M * 128;
  const int w_bytes = 2 * HB * 64 * 128;
  uint8_t* sX = smem;
  uint8_t* sW = sX + x_bytes;
  float* sBias = reinterpret_cast<float*>(sW + (size_t)NST * w_bytes);   // [16 warps][32]
  uint2* sQueue = reinterpret_cast<uint2*>(sBias + SF_EPI_WARPS * 32);   // [SF_Q][512 threads]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQueue + SF_Q * SF_EPI_WARPS * 32);
  uint64_t* x_full = bars;
  uint64_t* w_full = bars + 1;            // [4]
  uint64_t* w_empty = w_full + 4;         // [4]
  uint64_t* s_full = w_empty + 4;         // [4]
  uint64_t* s_empty = s_full + SF_NSB;    // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + SF_NSB);
  constexpr int WARP_TMA = 16, WARP_MMA = 17;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * ST_M;
  const int t_begin = (int)blockIdx.y * p.tiles_per_chunk;
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
    for (int b = 0; b < SF_NSB; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], SF_EPI_WARPS);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == WARP_TMA) {
    // whole warp, uniform control flow; TMA instructions predicated on an elected lane
    const uint32_t aX = smem_u32(sX), aW0 = smem_u32(sW);
    mbar_expect_tx_el(x_full, (uint32_t)x_bytes);
    for (int hb = 0; hb < HB; ++hb) tma_load_2d_el(aX + hb * (ST_M * 128), &tmX, x_full, hb * 64, m0);
    for (int t = 0; t < ntiles; ++t) {
      const int st = t % NST;
      if (p.l2_prefetch && t + p.l2_prefetch < ntiles) {
        const int vp = (p.tile0 + t_begin + t + p.l2_prefetch) * ST_N;
        for (int vb = 0; vb < 2; ++vb)
          for (int hb = 0; hb < HB; ++hb) tma_prefetch_2d_el(&tmW, vp + vb * 64, hb * 64);
      }
      mbar_wait_all(&w_empty[st], (uint32_t)((t / NST) & 1) ^ 1);
      mbar_expect_tx_el(&w_full[st], (uint32_t)w_bytes);
      const int v0 = (p.tile0 + t_begin + t) * ST_N;
      const uint32_t dst = aW0 + (uint32_t)(st * w_bytes);
      for (int vb = 0; vb < 2; ++vb)
        for (int hb = 0; hb < HB; ++hb)
          tma_load_2d_el(dst + (vb * HB + hb) * 8192, &tmW, &w_full[st], v0 + vb * 64, hb * 64);
    }
  } else if (warp == WARP_MMA) {
    // the WHOLE warp issues, in uniform control flow (see umma_bf16_el in common.cuh)
    const uint32_t idesc = umma_idesc_bf16(ST_M, ST_N, 0, 1);
    const uint64_t dX = umma_smem_desc(smem_u32(sX), 16, 1024);               // X, K-major
    const uint64_t dW = umma_smem_desc(smem_u32(sW), HB * 8192, 1024);        // W, MN-major
    mbar_wait_all(x_full, 0);
    for (int t = 0; t < ntiles; ++t) {
      const int st = t % NST, buf = t % SF_NSB;
      mbar_wait_all(&w_full[st], (uint32_t)((t / NST) & 1));
      mbar_wait_all(&s_empty[buf], (uint32_t)((t / SF_NSB) & 1) ^ 1);
      tc_fence_after();
      const uint32_t wo = (uint32_t)(st * w_bytes) >> 4;
      for (int hb = 0; hb < HB; ++hb) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16_el(tmem_base + buf * ST_N, dX + (uint32_t)((hb * (ST_M * 128) + kk * 32) >> 4),
                       dW + (wo + (uint32_t)((hb * 8192 + kk * 2048) >> 4)), idesc, (hb | kk) ? 1u : 0u);
      }
      umma_commit_el(&w_empty[st]);
      umma_commit_el(&s_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    const int q = warp & 3, cg = warp >> 2;
    const int row = m0 + q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t sb = smem_u32(sBias + warp * 32);
    const long crow = (long)row * p.cap;
    // the threshold is the seed's k-th best score; a later id with an equal score loses the tie
    // (ids only grow), so "strictly greater" is exact.  Rows past M never pass.
    const float tau = row < p.M ? p.cand_scores[crow + k - 1] : INFINITY;
    const uint32_t aQ = smem_u32(sQueue) + (uint32_t)(warp * 32 + lane) * 8;
    int n_staged = 0;
    auto load_bias = [&](int t) -> float {
      const int v = (p.tile0 + t_begin + t) * ST_N + cg * 32 + lane;
      return (t < ntiles && v < p.V) ? __ldg(p.bias + v) : -INFINITY;
    };
    float bias_next = load_bias(0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t % SF_NSB;
      const int vbase = (p.tile0 + t_begin + t) * ST_N + cg * 32;
      sts32f(sb + lane * 4, bias_next);
      __syncwarp();
      bias_next = load_bias(t + 1);   // in flight while this tile is processed
      mbar_wait(&s_full[buf], (uint32_t)((t / SF_NSB) & 1));
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_base + (uint32_t)(buf * ST_N + cg * 32), r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(&s_empty[buf]);
      float sc[32];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = lds128f(sb + j * 4);
        const float2 a = __fadd2_rn(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])),
                                    make_float2(b4.x, b4.y));
        const float2 b = __fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])),
                                    make_float2(b4.z, b4.w));
        sc[j + 0] = a.x; sc[j + 1] = a.y; sc[j + 2] = b.x; sc[j + 3] = b.y;
        mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y)));
      }
      __syncwarp();   // sb is rewritten for the next tile
      if (__any_sync(0xffffffffu, mx > tau)) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (sc[j] > tau)
            n_staged = filter_stage16(aQ, n_staged, sc[j], vbase + j, false, p.cand_cnt + row,
                                      p.cand_scores + crow, p.cand_ids + crow, k, p.cap);
      }
    }
    if (n_staged > 0)
      filter_stage16(aQ, n_staged, 0.f, 0, true, p.cand_cnt + row, p.cand_scores + crow,
                     p.cand_ids + crow, k, p.cap);
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static size_t score_filter_smem(int HB, int stages) {
  return (size_t)HB * ST_M * 128 + (size_t)stages * 2 * HB * 8192 + SF_EPI_WARPS * 32 * 4 +
         (size_t)SF_Q * SF_EPI_WARPS * 32 * 8 + 256 + 1024;
}

// Vocabulary chunks per row tile: one CTA per SM is resident, so the critical path is
// waves-of-148 x tiles-per-chunk; each chunk adds k candidates per row to the merge.
static int score_chunks(int n_mtiles, int n_vtiles, int* tiles_per_chunk) {
  int best = 1;
  double best_cost = 1e30;
  for (int c = 1; c <= std::min(n_vtiles, 148); ++c) {
    const int tpc = (n_vtiles + c - 1) / c;
    const int real = (n_vtiles + tpc - 1) / tpc;
    if (real != c) continue;
    const long waves = ((long)n_mtiles * c + 147) / 148;
    const double cost = (double)waves * tpc + 1.0 * c;
    if (cost < best_cost) best_cost = cost, best = c;
  }
  *tiles_per_chunk = (n_vtiles + best - 1) / best;
  return best;
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_topk_candidates(const float* cand_scores, const int32_t* cand_ids, long ld,
                                    long rows, int n_cand, int V, int k, int32_t* out_ids,
                                    float* out_scores, long ld_out, void* stream);
extern "C" int b4cp_topk_candidates_redo(const float* cand_scores, const int32_t* cand_ids, long ld,
                                         long rows, int n_cand, int V, int k, int32_t* out_ids,
                                         float* out_scores, long ld_out, void* stream);
extern "C" int b4cp_topk_rows(const float* scores, long ld, long rows, int V, int k,
                              int32_t* out_ids, float* out_scores, long ld_out, void* stream);
extern "C" int b4cp_topk_candidates_counted(const float* cand_scores, const int32_t* cand_ids, long ld,
                                            long rows, int n_cand, const int* extra_count,
                                            int base_count, int V, int k, int32_t* out_ids,
                                            float* out_scores, long ld_out, void* stream);

namespace b4cp {

// ---- long vocabularies: seed + filter + merge -----------------------------------------------
// Above FILTER_MIN_V entries the per-row heaps of score_topk_kernel<., false> are the bottleneck
// (~k ln(V/k) insertions per row, each a chain of dependent shared-memory accesses).  Instead:
//   1. seed: the scores of the first FILTER_SEED entries are materialised (rows x 65536 fp32,
//      bounded row blocks) and ranked by the streaming top-k -> every row has k candidates and a
//      threshold tau = its k-th best seed score;
//   2. filter: score_topk_kernel<false, true> runs the tcgen05 product over the REST of the
//      vocabulary and appends every score > tau to the row's candidate list in HBM (one compare
//      per score, ~k (V/seed - 1) appends per row);
//   3. merge: topk_candidates ranks the seed winners + appended candidates (exact, ties -> lower id).
// Rows whose list would overflow (adversarially ordered vocabularies) are marked and redone by the
// heap kernel, which is launched behind a device-side flag.
static constexpr int FILTER_MIN_V = 262144;
static constexpr int FILTER_SEED = 65536;
static constexpr int FILTER_ROWS = 4096;   // rows of seed scores materialised at a time (1 GB)
static constexpr int TK_REDO_MARK = -2;    // = TK_REDO in topk.cu

static int filter_cap(int V, int k) {
  const long expect = (long)k * (V / FILTER_SEED) + k;   // appended + seed winners
  long cap = 1024;
  while (cap < 3 * expect && cap < 16384) cap <<= 1;
  return (int)cap;
}

struct FilterWs {
  float* seed;
  float* cand_scores;
  int32_t* cand_ids;
  int* cnt;
  int* flag;
  void* heap_ws;
  size_t bytes;
};

static size_t heap_ws_bytes(long M, int V, int k) {
  int tpc;
  const int chunks = score_chunks(ceil_div(M, ST_M), ceil_div(V, ST_N), &tpc);
  return (size_t)chunks * M * k * 8 + 256;
}

static FilterWs carve_filter_ws(void* base, long M, int V, int k) {
  FilterWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + o : nullptr;
    o += (bytes + 255) & ~(size_t)255;
    return r;
  };
  const long cap = filter_cap(V, k);
  w.seed = (float*)take((size_t)std::min<long>(M, FILTER_ROWS) * FILTER_SEED * 4);
  w.cand_scores = (float*)take((size_t)M * cap * 4);
  w.cand_ids = (int32_t*)take((size_t)M * cap * 4);
  w.cnt = (int*)take((size_t)M * 4);
  w.flag = (int*)take(256);
  w.heap_ws = take(heap_ws_bytes(M, V, k));
  w.bytes = o;
  return w;
}

__global__ void __launch_bounds__(256)
filter_mark_overflow_kernel(const int* __restrict__ cnt, long M, int room, int32_t* __restrict__ out_ids,
                            long ld_out, int* __restrict__ flag) {
  const long row = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (row < M && cnt[row] > room) {
    out_ids[row * ld_out] = TK_REDO_MARK;
    *flag = 1;
  }
}

static int launch_heap(const void* x_bf16, long ldx, long M, int h, const void* w_bf16, long ldw,
                       const float* bias, int V, int k, int id_base, void* workspace,
                       const int* run_if, ScoreParams* p_out, cudaStream_t st) {
  ScoreParams p = {};
  p.M = (int)M;
  p.V = V;
  p.h = h;
  p.HB = h / 64;
  p.k = k;
  p.n_vtiles = ceil_div(V, ST_N);
  p.n_chunks = score_chunks(ceil_div(M, ST_M), p.n_vtiles, &p.tiles_per_chunk);
  p.bias = bias;
  p.id_base = id_base;
  p.part_scores = (float*)workspace;
  p.part_ids = (int32_t*)(p.part_scores + (size_t)p.n_chunks * M * k);
  p.run_if = run_if;
  const bool wide = h == 256;
  CUtensorMap tmX, tmW;
  int rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, ST_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  const size_t smem = (wide ? 0 : (size_t)p.HB * ST_M * 128) + (size_t)ST_STAGES * 2 * p.HB * 8192 +
                      (size_t)k * ST_M * 4 + (((size_t)k * ST_M * (wide ? 3 : 4) + 15) & ~(size_t)15) +
                      4 * 2 * ST_N * 4 + 256 + 1024;
  B4CP_CHECK_ARG(smem <= 227 * 1024, "score_topk: k=%d h=%d needs %zu B of shared memory", k, h, smem);
  dim3 grid(ceil_div(M, ST_M), p.n_chunks);
  if (wide) {
    B4CP_CUDA(cudaFuncSetAttribute(score_topk_kernel<true, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    score_topk_kernel<true, false><<<grid, 192, smem, st>>>(tmX, tmW, p, (const __nv_bfloat16*)x_bf16, ldx);
  } else {
    B4CP_CUDA(cudaFuncSetAttribute(score_topk_kernel<false, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    score_topk_kernel<false, false><<<grid, 192, smem, st>>>(tmX, tmW, p, (const __nv_bfloat16*)x_bf16, ldx);
  }
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  *p_out = p;
  return 0;
}

static int score_topk_filter(const void* x_bf16, long ldx, long M, int h, const void* w_bf16,
                             long ldw, const float* bias, int V, int k, int32_t* out_ids,
                             float* out_scores, long ld_out, void* workspace, cudaStream_t st) {
  const FilterWs w = carve_filter_ws(workspace, M, V, k);
  const int cap = filter_cap(V, k);
  // (the merge reads only the k + cnt[row] filled slots of a list: no fill of the lists needed)
  B4CP_CUDA(cudaMemsetAsync(w.cnt, 0, (size_t)M * 4, st));
  B4CP_CUDA(cudaMemsetAsync(w.flag, 0, 4, st));
  // 1. seed: scores of vocabulary [0, FILTER_SEED) for a block of rows, ranked exactly
  const __nv_bfloat16* x = (const __nv_bfloat16*)x_bf16;
  for (long r0 = 0; r0 < M; r0 += FILTER_ROWS) {
    const int rows = (int)std::min<long>(FILTER_ROWS, M - r0);
    b4cp_gemm_epilogue ep = {};
    ep.alpha = 1.f;
    ep.bias = bias;
    ep.out_f32 = w.seed;
    ep.ld_f32 = FILTER_SEED;
    int rc = b4cp_gemm_bf16(x + r0 * ldx, 0, ldx, w_bf16, 1, ldw, rows, FILTER_SEED, h, 1, &ep, st);
    if (rc) return rc;
    rc = b4cp_topk_rows(w.seed, FILTER_SEED, rows, FILTER_SEED, k, w.cand_ids + r0 * cap,
                        w.cand_scores + r0 * cap, cap, st);
    if (rc) return rc;
  }
  // 2. filter the rest of the vocabulary against the per-row thresholds
  ScoreParams p = {};
  p.M = (int)M;
  p.V = V;
  p.h = h;
  p.HB = h / 64;
  p.k = k;
  p.tile0 = FILTER_SEED / ST_N;
  p.n_vtiles = ceil_div(V, ST_N) - p.tile0;
  p.n_chunks = score_chunks(ceil_div(M, ST_M), p.n_vtiles, &p.tiles_per_chunk);
  p.bias = bias;
  p.cand_scores = w.cand_scores;
  p.cand_ids = w.cand_ids;
  p.cand_cnt = w.cnt;
  p.cap = cap;
  CUtensorMap tmX, tmW;
  int rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, ST_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  dim3 grid(ceil_div(M, ST_M), p.n_chunks);
  if (getenv("B4CP_FILTER_GEN1")) {   // developer switch: the first-generation filter kernel
    const size_t smem = (size_t)p.HB * ST_M * 128 + (size_t)ST_STAGES * 2 * p.HB * 8192 +
                        (size_t)2 * 16 * ST_M * 4 + 4 * 2 * ST_N * 4 + 256 + 1024;
    B4CP_CUDA(cudaFuncSetAttribute(score_topk_kernel<false, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    score_topk_kernel<false, true><<<grid, 192, smem, st>>>(tmX, tmW, p, x, ldx);
  } else {
    // L2 prefetch of W tiles ahead of their TMA load: measured harmful when every CTA prefetches
    // (3.5 -> 5.2 ms at C5, batch 4,096) - off; the switch remains for experiments
    p.l2_prefetch = 0;
    if (const char* e = getenv("B4CP_FILTER_PREFETCH")) p.l2_prefetch = atoi(e);   // developer switch
    int stages = 4;
    while (stages > 2 && score_filter_smem(p.HB, stages) > 227 * 1024) --stages;
    B4CP_CHECK_ARG(score_filter_smem(p.HB, stages) <= 227 * 1024, "score_topk: h=%d does not fit", h);
    B4CP_CUDA(cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    score_filter_kernel<<<grid, SF_THREADS, score_filter_smem(p.HB, stages), st>>>(tmX, tmW, p, stages);
  }
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  // 3. exact merge of seed winners + appended candidates
  rc = b4cp_topk_candidates_counted(w.cand_scores, w.cand_ids, cap, M, cap, w.cnt, k, V, k, out_ids,
                                    out_scores, ld_out, st);
  if (rc) return rc;
  // 4. rows whose list overflowed are marked and redone by the heap kernel (device-side flag)
  filter_mark_overflow_kernel<<<ceil_div(M, 256), 256, 0, st>>>(w.cnt, M, cap - k, out_ids, ld_out, w.flag);
  note_launches(1);
  ScoreParams hp;
  rc = launch_heap(x_bf16, ldx, M, h, w_bf16, ldw, bias, V, k, 0, w.heap_ws, w.flag, &hp, st);
  if (rc) return rc;
  return b4cp_topk_candidates_redo(hp.part_scores, hp.part_ids, (long)hp.n_chunks * k, M,
                                   hp.n_chunks * k, V, k, out_ids, out_scores, ld_out, st);
}

}  // namespace b4cp

extern "C" long b4cp_score_topk_workspace_bytes(long M, int V, int k) {
  if (V >= FILTER_MIN_V) return (long)carve_filter_ws(nullptr, M, V, k).bytes;
  return (long)heap_ws_bytes(M, V, k);
}

extern "C" int b4cp_score_topk(const void* x_bf16, long ldx, long M, int h, const void* w_bf16,
                               long ldw, const float* bias, int V, int k, int id_base, int V_total,
                               int32_t* out_ids, float* out_scores, long ld_out, void* workspace,
                               void* stream) {
  B4CP_CHECK_ARG(x_bf16 && w_bf16 && bias && out_ids && workspace, "score_topk: null argument");
  B4CP_CHECK_ARG(M > 0 && V > 0, "score_topk: empty problem");
  B4CP_CHECK_ARG(h == 64 || h == 128 || h == 256, "score_topk: head width h=%d unsupported (64, 128, 256)", h);
  B4CP_CHECK_ARG(h != 256 || (V < ID24_EMPTY && ((uintptr_t)x_bf16 & 15) == 0),
                 "score_topk: h=256 needs V < 2^24 - 1 and 16-byte aligned rows");
  B4CP_CHECK_ARG(k >= 1 && k <= ST_MAXK, "score_topk: k=%d must be in [1,%d]", k, ST_MAXK);
  B4CP_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "score_topk: leading dimensions must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (V >= FILTER_MIN_V && id_base == 0 && (V_total == 0 || V_total == V))
    return score_topk_filter(x_bf16, ldx, M, h, w_bf16, ldw, bias, V, k, out_ids, out_scores, ld_out,
                             workspace, st);
  ScoreParams p;
  const int rc = launch_heap(x_bf16, ldx, M, h, w_bf16, ldw, bias, V, k, id_base, workspace, nullptr,
                             &p, st);
  if (rc) return rc;
  return b4cp_topk_candidates(p.part_scores, p.part_ids, (long)p.n_chunks * k, M, p.n_chunks * k,
                              V_total > 0 ? V_total : V, k, out_ids, out_scores, ld_out, stream);
}
