// Fused output stage of the Cloze model: SoftMaxHead's Dense(V) (head.py:36,45) + softmax +
// sparse categorical cross-entropy (examples/BERT4Rec/source/utils.py:116-134, losses.py:31-98,
// logits mode) and its gradient.  This file holds the C ABI, the chunk planner and the small
// kernels around the two tcgen05 kernels of vocab_ce_ts.cu (partial merge, dX finish, vocabulary
// shard helpers).  The (M x V) logits / probabilities / dlogits exist only as 128x128 tiles in
// TMEM - never in HBM.
//
// X: bf16 [M][ldx] (K-major A operand).  W: bf16 [h][ldw], the Keras (in, out) kernel, used as the
// MN-major B operand of S = X W and, through a second descriptor on the SAME shared-memory tile, as
// the K-major B operand of U = P' W^T.  h in {128, 256} for the gradient paths, {64,128,192,256}
// for the loss alone.
#include <algorithm>
#include <cstdlib>

#include "vocab_ce.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

// partials are (max, sum) in the log2 domain; lse is returned in natural units
__global__ void __launch_bounds__(256)
vocab_ce_merge_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                      const VocabParams p, int M, int V, const int32_t* __restrict__ labels,
                      float* __restrict__ lse, float* __restrict__ tgt) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  const int n_parts = 4 * vocab_row_slots(p, row);
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
                   const float* __restrict__ part_u, const VocabParams sp, int M, int h,
                   const int32_t* __restrict__ labels, const float* __restrict__ loss_stats,
                   const float* __restrict__ lse_global, int V,
                   const __nv_bfloat16* __restrict__ w, long ldw,
                   const __nv_bfloat16* __restrict__ gate, long ld_gate,
                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, long ld_bf16) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * DX_ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= M) return;
  const int n_chunks = vocab_row_slots(sp, row);   // partial slots this row's tile holds
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

// Vocabulary chunks per row tile (SCHED_GRID).  One CTA per SM is resident, so the launch runs in
// waves of 148 CTAs and its critical path is waves x tiles-per-chunk tile iterations; every extra
// chunk adds one more (max, sum, U) partial per row for vocab_ce_dx_kernel to read (~2.5 tile
// iterations of HBM time).  Pick the count that minimises the sum instead of "about two waves".
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

// The forward schedule for (M, V, h): fills sched / n_chunks (= partial slots per row, the
// workspace stride) / tiles_per_chunk / range_q / total_tiles.  SCHED_RANGES (vocab_ce.cuh) when
// the bf16 kernel fits L2 comfortably; the lock-step chunk grid otherwise.
static constexpr long RANGES_MAX_W_BYTES = 32L << 20;
static void plan_forward(long M, int V, int h, VocabParams* p) {
  p->n_mtiles = ceil_div(M, VB_M);
  p->n_vtiles = ceil_div(V, VB_N);
  p->total_tiles = (long)p->n_mtiles * p->n_vtiles;
  const char* force = getenv("B4CP_FWD_SCHED");   // developer switch: "grid" / "ranges"
  const bool fits = (long)p->n_vtiles * VB_N * h * 2 <= RANGES_MAX_W_BYTES;
  const bool ranges = force ? (force[0] == 'r') : fits;
  if (ranges) {
    // small M: fill the GPU, but keep a row tile within MAX_FWD_CHUNKS partial slots
    const long q_min = std::max<long>(1, (p->n_vtiles + MAX_FWD_CHUNKS - 3) / (MAX_FWD_CHUNKS - 2));
    long ctas = std::min<long>(148, std::max<long>(1, p->total_tiles / q_min));
    const long q = (p->total_tiles + ctas - 1) / ctas;
    int slots = 1;
    for (int m = 0; m < p->n_mtiles; ++m) slots = std::max(slots, ranges_slots(q, p->n_vtiles, m));
    if (slots <= MAX_FWD_CHUNKS) {
      p->sched = SCHED_RANGES;
      p->range_q = q;
      p->n_chunks = slots;
      p->tiles_per_chunk = 0;
      return;
    }
  }
  p->sched = SCHED_GRID;
  p->range_q = 0;
  p->n_chunks = fwd_chunks(p->n_mtiles, p->n_vtiles, &p->tiles_per_chunk);
}

}  // namespace b4cp

using namespace b4cp;

extern "C" long b4cp_vocab_ce_workspace_bytes(long M, int V, int h) {
  VocabParams p = {};
  plan_forward(M, V, h, &p);
  return (long)(2 * 4 + h) * p.n_chunks * M * sizeof(float) + 256;
}

// how the forward of (M, V, h) is scheduled: out[0] = schedule (0 grid / 1 ranges), out[1] = CTAs,
// out[2] = partial slots per row (workspace stride), out[3] = tiles per CTA range (ranges) or per
// chunk (grid).  Host-only; lets tests pin the planner's invariants without a device.
extern "C" int b4cp_vocab_ce_plan(long M, int V, int h, long* out) {
  B4CP_CHECK_ARG(out && M > 0 && V > 0 && h > 0, "vocab_ce_plan: bad argument");
  VocabParams p = {};
  plan_forward(M, V, h, &p);
  out[0] = p.sched;
  out[1] = p.sched == SCHED_RANGES ? ceil_div(p.total_tiles, p.range_q) : (long)p.n_mtiles * p.n_chunks;
  out[2] = p.n_chunks;
  out[3] = p.sched == SCHED_RANGES ? p.range_q : p.tiles_per_chunk;
  return 0;
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
  B4CP_CHECK_ARG(!want_dx || h == 128 || h == 256,
                 "vocab_ce_fwd: the dX accumulation needs h in {128, 256} (got %d)", h);
  B4CP_CHECK_ARG(bias && labels && lse && tgt && workspace, "vocab_ce_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  VocabParams p = {};
  p.M = (int)M;
  p.V = V;
  p.h = h;
  p.HB = h / 64;
  plan_forward(M, V, h, &p);
  p.bias = bias;
  p.labels = labels;
  p.part_max = (float*)workspace;
  p.part_sum = p.part_max + (size_t)4 * p.n_chunks * M;
  p.part_u = p.part_sum + (size_t)4 * p.n_chunks * M;
  p.with_dx = want_dx ? 1 : 0;
  p.tgt = tgt;
  CUtensorMap tmX, tmW;
  rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, VB_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  rc = launch_vocab_fwd_ts(tmX, tmW, p, st);
  if (rc) return rc;
  vocab_ce_merge_kernel<<<ceil_div(M, 256), 256, 0, st>>>(p.part_max, p.part_sum, p, (int)M, V,
                                                          labels, lse, tgt);
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
  VocabParams sp = {};
  sp.M = (int)M;
  sp.V = V;
  sp.h = h;
  plan_forward(M, V, h, &sp);
  const int chunks = sp.n_chunks;
  const float* part_max = (const float*)workspace;
  const float* part_sum = part_max + (size_t)4 * chunks * M;
  const float* part_u = part_sum + (size_t)4 * chunks * M;
  B4CP_CHECK_ARG(ld_gate % 4 == 0 && ld_bf16 % 4 == 0, "vocab_ce_dx: leading dimensions must be multiples of 4");
  vocab_ce_dx_kernel<<<ceil_div(M, DX_ROWS_PER_BLOCK), 32 * DX_ROWS_PER_BLOCK, 0, (cudaStream_t)stream>>>(
      part_max, part_sum, part_u, sp, (int)M, h, labels, loss_stats, lse_global, V,
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
  B4CP_CHECK_ARG(h == 128 || h == 256,
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
  CUtensorMap tmX, tmW;
  rc = make_tmap_bf16_2d(&tmX, x_bf16, (uint64_t)h, (uint64_t)M, (uint64_t)ldx * 2, 64, VB_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmW, w_bf16, (uint64_t)V, (uint64_t)h, (uint64_t)ldw * 2, 64, 64);
  if (rc) return rc;
  rc = launch_vocab_bwd_ts(tmX, tmW, p, st);
  if (rc) return rc;
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
