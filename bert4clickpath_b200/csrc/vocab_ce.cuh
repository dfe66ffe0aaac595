// Declarations shared by the host side of the fused vocabulary stage (vocab_ce.cu: C ABI, chunk
// planner, merge / dX kernels) and its two tcgen05 kernels (vocab_ce_ts.cu).
#pragma once
#include "common.cuh"

namespace b4cp {

static constexpr int VB_M = 128;   // rows per tile
static constexpr int VB_N = 128;   // vocabulary entries per tile
static constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of two values on the FMA / ALU pipes (no MUFU): Cody-Waite range reduction with the
// 1.5 * 2^23 magic constant (t = x + magic holds round(x) in its low mantissa bits, r = x - round(x)
// in [-0.5, 0.5]), a degree-4 polynomial for 2^r (max relative error 7.2e-6, far inside the bf16
// rounding of the probabilities it feeds) and the exponent spliced in with one integer
// multiply-add.  The MUFU unit does 4 ex2 per clock per scheduler; a tile of the fused vocabulary
// kernels needs 16,384 of them - as long as its two tensor-core products - so a share of the
// exponentials is computed here instead, in packed fp32x2 arithmetic.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float2 magic = make_float2(12582912.f, 12582912.f);
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 r = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(r, make_float2(0.009666373953223228f, 0.009666373953223228f),
                        make_float2(0.055838342756032944f, 0.055838342756032944f));
  p = __ffma2_rn(p, r, make_float2(0.2402234822511673f, 0.2402234822511673f));
  p = __ffma2_rn(p, r, make_float2(0.6931367516517639f, 0.6931367516517639f));
  p = __ffma2_rn(p, r, make_float2(1.f, 1.f));
  // 2^round(x): (bits(t) << 23) keeps exactly round(x) in the exponent field (the magic constant's
  // own bits shift out), added to the exponent of p in [0.70, 1.42]
  float2 y;
  y.x = __int_as_float(__float_as_int(t.x) * 0x800000 + __float_as_int(p.x));
  y.y = __int_as_float(__float_as_int(t.y) * 0x800000 + __float_as_int(p.y));
  return y;
}

// Forward schedules.  A SEGMENT is a run of consecutive vocabulary tiles of ONE row tile processed
// by one CTA; it ends with one (max, sum, U) partial per row in slot `slot` of that row tile.
//   SCHED_GRID   grid = row tiles x vocabulary chunks, one segment per CTA, slot = chunk.  CTAs of
//                a chunk sweep its W tiles in lock step, so every W tile is fetched from HBM once
//                however large the vocabulary is (C4: W = 512 MB >> L2).
//   SCHED_RANGES persistent: <= 148 CTAs, CTA k owns tiles [k Q, (k+1) Q) of the row-major
//                (row tile, vocabulary tile) space - no wave quantisation, one CTA set-up instead
//                of ~10, and partials only where a range cuts a row tile (<= 2 per row once
//                Q >= n_vtiles).  Used when W fits L2 (C1 / C2: 13.9 MB), where the CTAs need not
//                be aligned on the vocabulary.  slot = k - first_cta(row tile).
enum { SCHED_GRID = 0, SCHED_RANGES = 1 };

struct VocabParams {
  int M, V, h, HB;          // HB = h / 64
  int n_mtiles, n_vtiles;
  int tiles_per_chunk;      // SCHED_GRID: vocabulary tiles per CTA
  int n_chunks;             // partial slots per row (both schedules: the workspace stride)
  int sched;                // SCHED_GRID / SCHED_RANGES
  long range_q;             // SCHED_RANGES: tiles per CTA
  long total_tiles;         // n_mtiles * n_vtiles
  const float* bias;
  const int32_t* labels;
  // forward outputs
  float* part_max;          // [n_chunks][M]
  float* part_sum;          // [n_chunks][M]
  float* tgt;               // [M]
  float* part_u;            // [n_chunks][M][h] un-normalised sum_v exp2(z2 - m) W[:,v] (with_dx)
  int with_dx;              // forward also accumulates part_u (h = 128)
  int fwd_stages;
  int x_bufs;               // forward: X tile buffers in shared memory (1 or 2)
  int l2_prefetch;          // forward: W tiles prefetched into L2 this many tiles ahead (0 = off)
  int l2_prefetch_every;    // ... by the CTAs whose row tile index is a multiple of this
  int grid_stagger;         // SCHED_GRID: row tile m starts its sweep m * grid_stagger tiles in
  // backward inputs / outputs
  const float* lse;         // [M]
  const float* loss_stats;  // [2]: (sum, n_valid)
  float* dW;                // [h][V]
  float* db;                // [V]
};

// number of partial slots row tile m holds, and the CTA that writes its slot 0 (SCHED_RANGES)
__host__ __device__ inline int ranges_first_cta(long q, int n_vtiles, int m) {
  return (int)(((long)m * n_vtiles) / q);
}
__host__ __device__ inline int ranges_slots(long q, int n_vtiles, int m) {
  return (int)((((long)(m + 1) * n_vtiles - 1) / q) - ((long)m * n_vtiles) / q) + 1;
}
__host__ __device__ inline int vocab_row_slots(const VocabParams& p, int row) {
  return p.sched == SCHED_RANGES ? ranges_slots(p.range_q, p.n_vtiles, row / VB_M) : p.n_chunks;
}

// Warp roles (both kernels, 576 threads): warps 0-15 = epilogue, warp 16 = TMA producer + TMEM
// owner, warp 17 = MMA issuer.  Epilogue warp w reads TMEM lanes [32*(w%4), +32) (hardware
// restriction); each scheduler has 4 epilogue warps to hide latency.
static constexpr int NUM_EPI_WARPS = 16;
static constexpr int WARP_TMA = 16, WARP_MMA = 17;
static constexpr float LN2 = 0.6931471805599453f;

// TS-form launchers (vocab_ce_ts.cu)
int launch_vocab_fwd_ts(const CUtensorMap& tmX, const CUtensorMap& tmW, const VocabParams& p,
                        cudaStream_t st);
int launch_vocab_bwd_ts(const CUtensorMap& tmX, const CUtensorMap& tmW, const VocabParams& p,
                        cudaStream_t st);

}  // namespace b4cp
