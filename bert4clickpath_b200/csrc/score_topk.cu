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
};

__global__ void __launch_bounds__(192, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const ScoreParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int HB = p.HB, k = p.k;
  const int x_bytes = HB * ST_M * 128;
  const int w_bytes = 2 * HB * 64 * 128;
  uint8_t* sX = smem;
  uint8_t* sW = sX + x_bytes;
  float* sHeapS = reinterpret_cast<float*>(sW + (size_t)ST_STAGES * w_bytes);  // [k][128]
  int32_t* sHeapI = reinterpret_cast<int32_t*>(sHeapS + (size_t)k * ST_M);       // [k][128]
  float* sBias = reinterpret_cast<float*>(sHeapI + (size_t)k * ST_M);            // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 2 * ST_N);
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
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  } else if (warp == WARP_MMA && lane == 0) {
    mbar_init(x_full, 1);
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
      mbar_expect_tx_el(x_full, (uint32_t)x_bytes);
      for (int hb = 0; hb < HB; ++hb) tma_load_2d_el(smem_u32(sX + hb * (ST_M * 128)), &tmX, x_full, hb * 64, m0);
      for (int t = 0; t < ntiles; ++t) {
        const int st = t % ST_STAGES;
        mbar_wait_all(&w_empty[st], ((t / ST_STAGES) & 1) ^ 1);
        mbar_expect_tx_el(&w_full[st], (uint32_t)w_bytes);
        const int v0 = (t_begin + t) * ST_N;
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
            const uint64_t da = umma_smem_desc(aX + hb * (ST_M * 128) + kk * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(aW + hb * 8192 + kk * 2048, HB * 8192, 1024);
            umma_bf16_el(tmem_base + buf * ST_N, da, db, idesc, (hb | kk) ? 1u : 0u);
          }
        }
        umma_commit_el(&w_empty[st]);
        umma_commit_el(&s_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = row
    const int r = warp * 32 + lane;
    const uint32_t aS = smem_u32(sHeapS) + r * 4, aI = smem_u32(sHeapI) + r * 4;
    const uint32_t aB = smem_u32(sBias);
    constexpr uint32_t SLOT = ST_M * 4;  // byte stride between heap slots
    for (int s = 0; s < k; ++s) {        // empty heap: k sentinels worse than any real score
      sts32f(aS + s * SLOT, -INFINITY);
      asm volatile("st.shared.s32 [%0], %1;" ::"r"(aI + s * SLOT), "r"(INT_MAX) : "memory");
    }
    float root_s = -INFINITY;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const int v0 = (t_begin + t) * ST_N;
      {  // bias of this tile (one value per thread), -inf past the vocabulary
        const int v = v0 + r;
        sts32f(aB + (buf * ST_N + r) * 4, v < p.V ? __ldg(p.bias + v) : -INFINITY);
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
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
              // replace the root (current k-th best) and sift down; "worse" = lower score, or
              // equal score and higher id
              const int id = v0 + c * 32 + j4 + u;
              int pos = 0;
              while (true) {
                const int l = 2 * pos + 1, rr = l + 1;
                if (l >= k) break;
                float cs = lds32f(aS + l * SLOT);
                int ci, cpos = l;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ci) : "r"(aI + l * SLOT) : "memory");
                if (rr < k) {
                  const float rs = lds32f(aS + rr * SLOT);
                  int ri;
                  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ri) : "r"(aI + rr * SLOT) : "memory");
                  if (rs < cs || (rs == cs && ri > ci)) {
                    cs = rs;
                    ci = ri;
                    cpos = rr;
                  }
                }
                // stop when the new item is not better than the worse child
                if (!(cs < sc || (cs == sc && ci > id))) break;
                sts32f(aS + pos * SLOT, cs);
                asm volatile("st.shared.s32 [%0], %1;" ::"r"(aI + pos * SLOT), "r"(ci) : "memory");
                pos = cpos;
              }
              sts32f(aS + pos * SLOT, sc);
              asm volatile("st.shared.s32 [%0], %1;" ::"r"(aI + pos * SLOT), "r"(id) : "memory");
              root_s = lds32f(aS);
            }
          }
        }
      }
    }
    const int row = m0 + r;
    if (row < p.M) {
      float* os = p.part_scores + ((size_t)row * p.n_chunks + chunk) * k;  // [row][chunk][k]
      int32_t* oi = p.part_ids + ((size_t)row * p.n_chunks + chunk) * k;
      for (int s = 0; s < k; ++s) {
        int id;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(id) : "r"(aI + s * SLOT) : "memory");
        os[s] = lds32f(aS + s * SLOT);
        oi[s] = id == INT_MAX ? -1 : id + p.id_base;
      }
    }
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
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

extern "C" long b4cp_score_topk_workspace_bytes(long M, int V, int k) {
  int tpc;
  const int chunks = score_chunks(ceil_div(M, ST_M), ceil_div(V, ST_N), &tpc);
  return (long)chunks * M * k * 8 + 256;
}

extern "C" int b4cp_score_topk(const void* x_bf16, long ldx, long M, int h, const void* w_bf16,
                               long ldw, const float* bias, int V, int k, int id_base, int V_total,
                               int32_t* out_ids, float* out_scores, long ld_out, void* workspace,
                               void* stream) {
  B4CP_CHECK_ARG(x_bf16 && w_bf16 && bias && out_ids && workspace, "score_topk: null argument");
  B4CP_CHECK_ARG(M > 0 && V > 0, "score_topk: empty problem");
  B4CP_CHECK_ARG(h == 64 || h == 128, "score_topk: head width h=%d unsupported (64 or 128)", h);
  B4CP_CHECK_ARG(k >= 1 && k <= ST_MAXK, "score_topk: k=%d must be in [1,%d]", k, ST_MAXK);
  B4CP_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "score_topk: leading dimensions must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
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
  CUtensorMap tmX, tmW;
  int rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, ST_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  const size_t smem = (size_t)p.HB * ST_M * 128 + (size_t)ST_STAGES * 2 * p.HB * 8192 +
                      (size_t)2 * k * ST_M * 4 + 2 * ST_N * 4 + 256 + 1024;
  B4CP_CHECK_ARG(smem <= 227 * 1024, "score_topk: k=%d h=%d needs %zu B of shared memory", k, h, smem);
  B4CP_CUDA(cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
  dim3 grid(ceil_div(M, ST_M), p.n_chunks);
  score_topk_kernel<<<grid, 192, smem, st>>>(tmX, tmW, p);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return b4cp_topk_candidates(p.part_scores, p.part_ids, (long)p.n_chunks * k, M, p.n_chunks * k,
                              V_total > 0 ? V_total : V, k, out_ids, out_scores, ld_out, stream);
}
