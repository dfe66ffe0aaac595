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

struct VocabParams {
  int M, V, h, HB;          // HB = h / 64
  int n_mtiles, n_vtiles;
  int tiles_per_chunk;      // forward: vocabulary tiles per CTA
  int n_chunks;
  const float* bias;
  const int32_t* labels;
  // forward outputs
  float* part_max;          // [n_chunks][M]
  float* part_sum;          // [n_chunks][M]
  float* tgt;               // [M]
  float* part_u;            // [n_chunks][M][h] un-normalised sum_v exp2(z2 - m) W[:,v] (with_dx)
  int with_dx;              // forward also accumulates part_u (h = 128)
  int fwd_stages;
  // backward inputs / outputs
  const float* lse;         // [M]
  const float* loss_stats;  // [2]: (sum, n_valid)
  float* dW;                // [h][V]
  float* db;                // [V]
};

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
