// Generic bf16 x bf16 -> fp32 GEMM on tcgen05 tensor cores (sm_100a), operands fed by TMA.
//
//   C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// Replaces the tf.keras.layers.Dense MatMul+BiasAdd call sites of the reference
// (clickstream_transformer/transformer.py:112-116,139-141,158,163-167; head.py:35-45) and
// their autodiff transposes.  Operands may be K-major (k contiguous in memory) or MN-major
// (m / n contiguous), which covers X*W, dY*W^T and X^T*dY without materialised transposes.
//
// One CTA = one 128 x BN output tile (x one K split).  Warp roles: warp 0 = TMA producer and
// TMEM allocator, warp 1 = single-thread tcgen05.mma issuer, warps 2..5 = epilogue
// (TMEM -> registers -> global).  A `stages`-deep mbarrier ring connects producer and issuer.
#include <algorithm>
#include <cstring>

#include <cstdlib>
#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

static constexpr int BM = 128;
static constexpr int BK = 64;  // bf16 elements = one 128-byte swizzle row
static constexpr int A_STAGE_BYTES = BM * BK * 2;

struct GemmKernelParams {
  int M, N, K;
  int BN;
  int a_mn, b_mn;
  int stages;
  int k_tiles_total;
  int k_tiles_per_split;
  int tma_out_bf16, tma_out_f32;  // persistent kernel: outputs leave through TMA stores
  b4cp_gemm_epilogue ep;
};

// One 32-column chunk of one accumulator row: scale, bias, ReLU, ReLU-gate, residual add, stores.
// `sbias` points at this chunk's 32 bias values staged in shared memory (zero past N), or NULL.
__device__ __forceinline__ void epilogue_math(const b4cp_gemm_epilogue& ep, int row, int col0,
                                              int N, const uint32_t (&r)[32], const float* sbias,
                                              float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
  const int ncol = min(32, N - col0);
  if (sbias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = lds128f(smem_u32(sbias) + j * 4);
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (ep.gate) {  // dY * [gate > 0]  (ReLU backward)
    const __nv_bfloat16* g =
        reinterpret_cast<const __nv_bfloat16*>(ep.gate) + (size_t)row * ep.ld_gate + col0;
    if (ncol == 32 && ((reinterpret_cast<uintptr_t>(g) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(g + j));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // bf16 > 0  <=>  sign bit clear and magnitude bits non-zero
          if (!((w[k] & 0x8000u) == 0 && (w[k] & 0x7FFFu) != 0)) v[j + 2 * k] = 0.f;
          if (!((w[k] & 0x80000000u) == 0 && (w[k] & 0x7FFF0000u) != 0)) v[j + 2 * k + 1] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)  // predicated, fully unrolled: v[] stays in registers
        if (j < ncol && !(__bfloat162float(g[j]) > 0.f)) v[j] = 0.f;
    }
  }
  if (ep.addend) {
    const float* a = ep.addend + (size_t)row * ep.ld_addend + col0;
    if (ncol == 32 && ((reinterpret_cast<uintptr_t>(a) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(a + j));
        v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncol) v[j] += a[j];
    }
  }
}

// direct (per-thread row) stores of one finished chunk
__device__ __forceinline__ void epilogue_store_direct(const b4cp_gemm_epilogue& ep, float* out_f32,
                                                      int row, int col0, int N,
                                                      const float (&v)[32]) {
  const int ncol = min(32, N - col0);
  if (out_f32) {
    float* o = out_f32 + (size_t)row * ep.ld_f32 + col0;
    if (ncol == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncol) o[j] = v[j];
    }
  }
  if (ep.out_bf16) {
    __nv_bfloat16* o =
        reinterpret_cast<__nv_bfloat16*>(ep.out_bf16) + (size_t)row * ep.ld_bf16 + col0;
    if (ncol == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        uint4 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(o + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncol) o[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// Coalesced stores of one finished 32-row x 32-column chunk of a warp (lane = row): the chunk is
// transposed through a per-warp swizzled smem tile so that every store instruction writes whole
// rows segments (8 lanes x 16 B = one 128-byte fp32 row, 4 lanes x 16 B = one 64-byte bf16 row).
// Per-thread row stores write 16-byte halves of 32-byte sectors in separate instructions, which
// makes L2 fetch every output sector from DRAM before merging (measured: DRAM reads = output size).
// `stg` : 4096 B per warp, 128-byte aligned.  Rows >= M and chunks that are not 32 wide / 16-byte
// aligned take the direct path.
__device__ __forceinline__ void epilogue_store_coalesced(const b4cp_gemm_epilogue& ep,
                                                         float* out_f32, int row0, int M, int col0,
                                                         int N, const float (&v)[32],
                                                         uint32_t stg, int lane) {
  const int ncol = min(32, N - col0);
  const bool f_ok = out_f32 && ncol == 32 &&
                    ((reinterpret_cast<uintptr_t>(out_f32 + col0) | (uintptr_t)(ep.ld_f32 * 4)) & 15) == 0;
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(ep.out_bf16);
  const bool b_ok = ob && ncol == 32 &&
                    ((reinterpret_cast<uintptr_t>(ob + col0) | (uintptr_t)(ep.ld_bf16 * 2)) & 15) == 0;
  const int row = row0 + lane;
  if (out_f32) {
    if (f_ok) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        sts128(stg + lane * 128 + ((k ^ (lane & 7)) << 4), __float_as_uint(v[4 * k]),
               __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]),
               __float_as_uint(v[4 * k + 3]));
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + (lane >> 3), pc = lane & 7;
        const float4 q = lds128f(stg + r * 128 + ((pc ^ (r & 7)) << 4));
        if (row0 + r < M)
          *reinterpret_cast<float4*>(out_f32 + (size_t)(row0 + r) * ep.ld_f32 + col0 + pc * 4) = q;
      }
      __syncwarp();
    } else if (row < M) {
      b4cp_gemm_epilogue e2 = ep;
      e2.out_bf16 = nullptr;
      epilogue_store_direct(e2, out_f32, row, col0, N, v);
    }
  }
  if (ob) {
    if (b_ok) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * k + 0], v[8 * k + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * k + 2], v[8 * k + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * k + 4], v[8 * k + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * k + 6], v[8 * k + 7]);
        sts128(stg + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4), *reinterpret_cast<uint32_t*>(&h0),
               *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
               *reinterpret_cast<uint32_t*>(&h3));
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i * 8 + (lane >> 2), pc = lane & 3;
        const float4 q = lds128f(stg + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
        if (row0 + r < M)
          *reinterpret_cast<float4*>(ob + (size_t)(row0 + r) * ep.ld_bf16 + col0 + pc * 8) = q;
      }
      __syncwarp();
    } else if (row < M) {
      epilogue_store_direct(ep, nullptr, row, col0, N, v);
    }
  }
}

__device__ __forceinline__ void epilogue_chunk(const b4cp_gemm_epilogue& ep, float* out_f32,
                                               int row, int col0, int N, const uint32_t (&r)[32],
                                               const float* sbias) {
  float v[32];
  epilogue_math(ep, row, col0, N, r, sbias, v);
  epilogue_store_direct(ep, out_f32, row, col0, N, v);
}

// bias of the CTA's N tile staged in shared memory (zero beyond N) by `nthreads` threads
__device__ __forceinline__ void stage_bias_tile(float* sbias, const float* bias, int n0, int BN,
                                                int N, int tid, int nthreads) {
  for (int j = tid; j < BN; j += nthreads)
    sbias[j] = (bias && n0 + j < N) ? __ldg(bias + n0 + j) : 0.f;
}

__global__ void __launch_bounds__(192, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int BN = p.BN;
  const int b_stage_bytes = BN * BK * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* sbias = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));  // [BN]
  uint8_t* sStage = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(sbias + BN) + 127) & ~static_cast<uintptr_t>(127));  // [4][4096]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kt_begin = blockIdx.z * p.k_tiles_per_split;
  const int kt_end = min(p.k_tiles_total, kt_begin + p.k_tiles_per_split);
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < BN) tmem_cols <<= 1;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp >= 2) stage_bias_tile(sbias, p.ep.bias, n0, BN, p.N, threadIdx.x - 64, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    {  // whole warp, uniform control flow; TMA instructions predicated on an elected lane
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        mbar_wait_all(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = tiles + (size_t)stage * stage_bytes;
        uint8_t* sB = sA + A_STAGE_BYTES;
        mbar_expect_tx_el(&full_bar[stage], (uint32_t)stage_bytes);
        if (!p.a_mn) {
          tma_load_2d_el(smem_u32(sA), &tmA, &full_bar[stage], kt * BK, m0);
        } else {
          for (int h = 0; h < BM / 64; ++h)
            tma_load_2d_el(smem_u32(sA + h * (64 * BK * 2)), &tmA, &full_bar[stage], m0 + h * 64, kt * BK);
        }
        if (!p.b_mn) {
          tma_load_2d_el(smem_u32(sB), &tmB, &full_bar[stage], kt * BK, n0);
        } else {
          for (int h = 0; h < BN / 64; ++h)
            tma_load_2d_el(smem_u32(sB + h * (64 * BK * 2)), &tmB, &full_bar[stage], n0 + h * 64, kt * BK);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {  // the WHOLE warp issues, in uniform control flow (see umma_bf16_el in common.cuh)
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      // K-major: 8-row groups 1024 B apart; one UMMA_K (16 bf16) = 32 B along the row.
      // MN-major: 8 K-rows per 1024-B group, 64-wide M/N blocks one full box (64 x BK) apart;
      //           one UMMA_K = two groups = 2048 B.
      const uint32_t a_lbo = p.a_mn ? 64 * BK * 2 : 16, a_sbo = 1024;
      const uint32_t b_lbo = p.b_mn ? 64 * BK * 2 : 16, b_sbo = 1024;
      const uint32_t a_kstep = p.a_mn ? 2048 : 32;
      const uint32_t b_kstep = p.b_mn ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        mbar_wait_all(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(tiles + (size_t)stage * stage_bytes);
        const uint32_t sB = sA + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = umma_smem_desc(sA + k * a_kstep, a_lbo, a_sbo);
          const uint64_t db = umma_smem_desc(sB + k * b_kstep, b_lbo, b_sbo);
          umma_bf16_el(tmem_base, da, db, idesc, (kt > kt_begin || k > 0) ? 1u : 0u);
        }
        umma_commit_el(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_el(tmem_full_bar);
    }
  } else {
    // ---- epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int q = warp & 3;
    const int row0 = m0 + q * 32;
    const int row = row0 + lane;
    const uint32_t stg = smem_u32(sStage + (warp - 2) * 4096);
    const b4cp_gemm_epilogue& ep = p.ep;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    float* out_f32 = ep.out_f32 ? ep.out_f32 + (size_t)blockIdx.z * ep.split_stride : nullptr;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      const int col0 = n0 + c0;
      if (row0 >= p.M || col0 >= p.N) continue;  // warp-uniform
      float v[32];
      if (row < p.M) {
        epilogue_math(ep, row, col0, p.N, r, ep.bias ? sbias + c0 : nullptr, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      epilogue_store_coalesced(ep, out_f32, row0, p.M, col0, p.N, v, stg, lane);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------ tile-persistent variant
// More output tiles than SMs (the head MLP and its gradients): one CTA per SM walks the tiles
// t, t + grid, ... (row tile fastest, so concurrently running CTAs share the B column tile in
// L2).  The smem ring runs continuously across tiles and two TMEM accumulators ping-pong, so the
// epilogue of tile i overlaps the TMA loads and MMAs of tile i + 1 and the per-CTA set-up
// (TMEM allocation, barrier init, descriptor prefetch) is paid once per SM, not once per tile.
struct TileCoord {
  int m0, n0, z, kt_begin, kt_end;
};
__device__ __forceinline__ TileCoord tile_coord(const GemmKernelParams& p, int t, int n_m, int n_n) {
  TileCoord c;
  c.z = t / (n_m * n_n);
  const int rem = t - c.z * (n_m * n_n);
  const int nt = rem / n_m;
  c.m0 = (rem - nt * n_m) * BM;
  c.n0 = nt * p.BN;
  c.kt_begin = c.z * p.k_tiles_per_split;
  c.kt_end = min(p.k_tiles_total, c.kt_begin + p.k_tiles_per_split);
  return c;
}

static constexpr int TILES_EPI_WARPS = 16;  // four per TMEM lane quadrant, interleaved column chunks
static constexpr int TILES_THREADS = 64 + 32 * TILES_EPI_WARPS;

__global__ void __launch_bounds__(TILES_THREADS, 1)
gemm_umma_tiles_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const GemmKernelParams p, int n_m, int n_n, int total_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int BN = p.BN;
  const int b_stage_bytes = BN * BK * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* acc_full = empty_bar + p.stages;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sbias = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 1) + 15) & ~static_cast<uintptr_t>(15));  // [2][BN]
  uint8_t* sStage = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(sbias + 2 * BN) + 127) & ~static_cast<uintptr_t>(127));  // [8][4096]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], TILES_EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    {  // whole warp, uniform control flow; TMA instructions predicated on an elected lane
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord c = tile_coord(p, t, n_m, n_n);
        for (int kt = c.kt_begin; kt < c.kt_end; ++kt) {
          mbar_wait_all(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = tiles + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + A_STAGE_BYTES;
          mbar_expect_tx_el(&full_bar[stage], (uint32_t)stage_bytes);
          if (!p.a_mn) {
            tma_load_2d_el(smem_u32(sA), &tmA, &full_bar[stage], kt * BK, c.m0);
          } else {
            for (int h = 0; h < BM / 64; ++h)
              tma_load_2d_el(smem_u32(sA + h * (64 * BK * 2)), &tmA, &full_bar[stage], c.m0 + h * 64, kt * BK);
          }
          if (!p.b_mn) {
            tma_load_2d_el(smem_u32(sB), &tmB, &full_bar[stage], kt * BK, c.n0);
          } else {
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d_el(smem_u32(sB + h * (64 * BK * 2)), &tmB, &full_bar[stage], c.n0 + h * 64, kt * BK);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // the WHOLE warp issues, in uniform control flow (see umma_bf16_el in common.cuh)
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 64 * BK * 2 : 16, a_sbo = 1024;
      const uint32_t b_lbo = p.b_mn ? 64 * BK * 2 : 16, b_sbo = 1024;
      const uint32_t a_kstep = p.a_mn ? 2048 : 32;
      const uint32_t b_kstep = p.b_mn ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
        const TileCoord c = tile_coord(p, t, n_m, n_n);
        const int acc = iter & 1;
        mbar_wait_all(&acc_empty[acc], (uint32_t)((iter >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kt = c.kt_begin; kt < c.kt_end; ++kt) {
          mbar_wait_all(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(tiles + (size_t)stage * stage_bytes);
          const uint32_t sB = sA + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc(sA + k * a_kstep, a_lbo, a_sbo);
            const uint64_t db = umma_smem_desc(sB + k * b_kstep, b_lbo, b_sbo);
            umma_bf16_el(d_tmem, da, db, idesc, (kt > c.kt_begin || k > 0) ? 1u : 0u);
          }
          umma_commit_el(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_el(&acc_full[acc]);
      }
    }
  } else {
    // ---- epilogue warps 2..: TMEM lane quadrant warp % 4, column chunks half, half + NPQ, ...
    constexpr int NPQ = TILES_EPI_WARPS / 4;   // warps per quadrant
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int n_chunks = BN / 32;
    const int etid = threadIdx.x - 64;
    const b4cp_gemm_epilogue& ep = p.ep;
    int iter = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
      const TileCoord c = tile_coord(p, t, n_m, n_n);
      const int acc = iter & 1;
      float* sb = sbias + acc * BN;
      if (ep.bias) {
        stage_bias_tile(sb, ep.bias, c.n0, BN, p.N, etid, 32 * TILES_EPI_WARPS);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TILES_EPI_WARPS) : "memory");
      }
      mbar_wait(&acc_full[acc], (uint32_t)((iter >> 1) & 1));
      tc_fence_after();
      const int row0 = c.m0 + q * 32;
      const int row = row0 + lane;
      const uint32_t stg = smem_u32(sStage + (warp - 2) * 4096);
      float* out_f32 = ep.out_f32 ? ep.out_f32 + (size_t)c.z * ep.split_stride : nullptr;
      bool released = false;
      for (int ci = half; ci < n_chunks; ci += NPQ) {
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        tmem_ld_wait();
        if (ci + NPQ >= n_chunks) {  // this warp's last read: hand the accumulator back early
          tc_fence_before();
          mbar_arrive_warp(&acc_empty[acc]);
          released = true;
        }
        const int col0 = c.n0 + c0;
        if (row0 >= p.M || col0 >= p.N) continue;  // warp-uniform
        float v[32];
        if (row < p.M) {
          epilogue_math(ep, row, col0, p.N, r, ep.bias ? sb + c0 : nullptr, v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        epilogue_store_coalesced(ep, out_f32, row0, p.M, col0, p.N, v, stg, lane);
      }
      if (!released) {  // BN = 32: the second warp of the quadrant has no chunk
        tc_fence_before();
        mbar_arrive_warp(&acc_empty[acc]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------ persistent variant
// Skinny products (one N tile, K <= 192: every Dense of the encoder and most of the head MLP at
// large row counts) are HBM-bound and, one tile per CTA, dominated by per-CTA set-up.  Here a CTA
// keeps the whole B operand resident in shared memory, streams 128-row A tiles through a 2-stage
// ring and ping-pongs two TMEM accumulators so the epilogue of tile i overlaps the loads and MMAs
// of tile i+1.  8 epilogue warps: warp w drains TMEM lanes 32*(w%4).. and column half w/4.
static constexpr int PERSIST_MAX_KT = 3;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0,
                                             int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Epilogue warps of the persistent kernel.  The skinny products are bound by their epilogue (scale /
// bias / activation / pack / staging of 24 K outputs per 128 x 192 tile: ncu shows the 8-warp
// version at IPC 1.3 with its warps waiting on each other's latencies), so each TMEM lane quadrant
// gets PE_WARPS / 4 warps that split the tile's 32-column chunks.
static constexpr int PE_WARPS = 16;
static constexpr int PE_THREADS = 64 + 32 * PE_WARPS;

__global__ void __launch_bounds__(PE_THREADS, 1)
gemm_umma_persistent_kernel(const __grid_constant__ CUtensorMap tmA,
                            const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmOutB,
                            const __grid_constant__ CUtensorMap tmOutF, const GemmKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int BN = p.BN, KT = p.k_tiles_total;
  const int b_kt_bytes = BN * BK * 2;
  const int a_stage_bytes = KT * A_STAGE_BYTES;
  uint8_t* sB = smem;
  uint8_t* sA = sB + (size_t)KT * b_kt_bytes;
  // per-epilogue-warp staging tiles for the TMA stores: fp32 32x32 (4 KB, 128B swizzle) and bf16
  // 32x32 (2 KB, 64B swizzle)
  uint8_t* sStageF = sA + 2 * (size_t)a_stage_bytes;
  uint8_t* sStageB = sStageF + (p.tma_out_f32 ? PE_WARPS * 4096 : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStageB + (p.tma_out_bf16 ? PE_WARPS * 2048 : 0));
  uint64_t* b_full = bars;
  uint64_t* a_full = bars + 1;    // [2]
  uint64_t* a_empty = bars + 3;   // [2]
  uint64_t* t_full = bars + 5;    // [2]
  uint64_t* t_empty = bars + 7;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* sbias = reinterpret_cast<float*>(bars + 12);  // [BN]
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int n_mtiles = (p.M + BM - 1) / BM;
  const int n_my = (n_mtiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * BN) tmem_cols <<= 1;
  constexpr int WARP_TMA = PE_WARPS, WARP_MMA = PE_WARPS + 1;  // single-thread roles at the highest warp ids

  if (warp == WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  } else if (warp == WARP_MMA && lane == 0) {
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], PE_WARPS);
    }
    fence_barrier_init();
  }
  if (warp < 8) stage_bias_tile(sbias, p.ep.bias, 0, BN, p.N, threadIdx.x, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == WARP_TMA) {
    {  // whole warp, uniform control flow; TMA instructions predicated on an elected lane
      mbar_expect_tx_el(b_full, (uint32_t)(KT * b_kt_bytes));
      for (int kt = 0; kt < KT; ++kt) {
        uint8_t* dst = sB + (size_t)kt * b_kt_bytes;
        if (!p.b_mn) {
          tma_load_2d_el(smem_u32(dst), &tmB, b_full, kt * BK, 0);
        } else {
          for (int h = 0; h < BN / 64; ++h)
            tma_load_2d_el(smem_u32(dst + h * (64 * BK * 2)), &tmB, b_full, h * 64, kt * BK);
        }
      }
      for (int i = 0; i < n_my; ++i) {
        const int st = i & 1;
        const int m0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM;
        mbar_wait_all(&a_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx_el(&a_full[st], (uint32_t)a_stage_bytes);
        for (int kt = 0; kt < KT; ++kt) {
          uint8_t* dst = sA + (size_t)st * a_stage_bytes + (size_t)kt * A_STAGE_BYTES;
          if (!p.a_mn) {
            tma_load_2d_el(smem_u32(dst), &tmA, &a_full[st], kt * BK, m0);
          } else {
            for (int h = 0; h < BM / 64; ++h)
              tma_load_2d_el(smem_u32(dst + h * (64 * BK * 2)), &tmA, &a_full[st], m0 + h * 64, kt * BK);
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    {  // the WHOLE warp issues, in uniform control flow (see umma_bf16_el in common.cuh)
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 64 * BK * 2 : 16, b_lbo = p.b_mn ? 64 * BK * 2 : 16;
      const uint32_t a_kstep = p.a_mn ? 2048 : 32, b_kstep = p.b_mn ? 2048 : 32;
      mbar_wait_all(b_full, 0);
      for (int i = 0; i < n_my; ++i) {
        const int st = i & 1;
        mbar_wait_all(&a_full[st], (i >> 1) & 1);
        mbar_wait_all(&t_empty[st], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kt = 0; kt < KT; ++kt) {
          const uint32_t aA = smem_u32(sA + (size_t)st * a_stage_bytes + (size_t)kt * A_STAGE_BYTES);
          const uint32_t aB = smem_u32(sB + (size_t)kt * b_kt_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc(aA + k * a_kstep, a_lbo, 1024);
            const uint64_t db = umma_smem_desc(aB + k * b_kstep, b_lbo, 1024);
            umma_bf16_el(tmem_base + st * BN, da, db, idesc, (kt | k) ? 1u : 0u);
          }
        }
        umma_commit_el(&a_empty[st]);
        umma_commit_el(&t_full[st]);
      }
    }
  } else {
    const int q = warp & 3, part = warp >> 2;  // PE_WARPS / 4 warps per lane quadrant
    const b4cp_gemm_epilogue& ep = p.ep;
    const int chunks = BN / 32;
    constexpr int NP = PE_WARPS / 4;
    const int c_begin = (part * chunks) / NP, c_end = ((part + 1) * chunks) / NP;
    uint8_t* stgF = sStageF + warp * 4096;
    uint8_t* stgB = sStageB + warp * 2048;
    const bool staged = p.tma_out_bf16 || p.tma_out_f32;
    if (staged && lane == 0) {
      if (p.tma_out_bf16) tma_prefetch_desc(&tmOutB);
      if (p.tma_out_f32) tma_prefetch_desc(&tmOutF);
    }
    for (int i = 0; i < n_my; ++i) {
      const int st = i & 1;
      const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * BM + q * 32;
      const int row = row0 + lane;
      mbar_wait(&t_full[st], (i >> 1) & 1);
      tc_fence_after();
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * BN + c * 32), r);
        tmem_ld_wait();
        if (c == c_end - 1) {  // accumulator fully read: hand the TMEM buffer back early
          tc_fence_before();
          mbar_arrive_warp(&t_empty[st]);
        }
        if (c * 32 >= p.N || row0 >= p.M) continue;  // warp-uniform
        const float* sb = ep.bias ? sbias + c * 32 : nullptr;
        if (!staged) {
          if (row < p.M) epilogue_chunk(ep, ep.out_f32, row, c * 32, p.N, r, sb);
          continue;
        }
        float v[32];
        if (row < p.M) {
          epilogue_math(ep, row, c * 32, p.N, r, sb, v);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;  // clipped by the TMA store anyway
        }
        // the previous store of this warp must have finished READING the staging tiles
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        if (p.tma_out_f32) {
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16-byte chunk k of the 128-byte row, 128B swizzle
            sts128(smem_u32(stgF) + lane * 128 + ((k ^ (lane & 7)) << 4), __float_as_uint(v[4 * k]),
                   __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]),
                   __float_as_uint(v[4 * k + 3]));
        } else if (ep.out_f32 && row < p.M) {
          b4cp_gemm_epilogue e2 = ep;
          e2.out_bf16 = nullptr;
          epilogue_store_direct(e2, ep.out_f32, row, c * 32, p.N, v);
        }
        if (p.tma_out_bf16) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16-byte chunk k of the 64-byte row, 64B swizzle
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * k + 0], v[8 * k + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * k + 2], v[8 * k + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * k + 4], v[8 * k + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * k + 6], v[8 * k + 7]);
            sts128(smem_u32(stgB) + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4),
                   *reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                   *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
        } else if (ep.out_bf16 && row < p.M) {
          b4cp_gemm_epilogue e2 = ep;
          epilogue_store_direct(e2, nullptr, row, c * 32, p.N, v);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.tma_out_f32) tma_store_2d(&tmOutF, stgF, c * 32, row0);
          if (p.tma_out_bf16) tma_store_2d(&tmOutB, stgB, c * 32, row0);
          tma_store_commit();
        }
      }
      if (c_begin >= c_end) {
        tc_fence_before();
        mbar_arrive_warp(&t_empty[st]);
      }
    }
    if (staged && lane == 0) tma_store_wait_read();  // smem must outlive the last store's read
  }
  __syncthreads();
  if (warp == WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, CUtensorMapSwizzle swz,
                 const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer);

// 2-D bf16 tensor map, 128B swizzle, zero fill out of bounds.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, base, inner,
                      outer, row_stride_bytes, box_inner, box_outer);
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, CUtensorMapSwizzle swz,
                 const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point not available");
    return -2;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu stride=%llu",
                   (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
                   (unsigned long long)row_stride_bytes);
    return -3;
  }
  return 0;
}

static int pick_bn(int N, int b_mn) {
  int bn;
  if (N > 128) bn = 256;
  else if (N > 64) bn = 128;
  else if (N > 32) bn = 64;
  else if (N > 16) bn = 32;
  else bn = 16;
  if (b_mn && bn < 64) bn = 64;
  return bn;
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_gemm_bf16(const void* A, int a_mn, long lda, const void* B, int b_mn,
                              long ldb, int M, int N, int K, int splits,
                              const b4cp_gemm_epilogue* ep, void* stream) {
  B4CP_CHECK_ARG(A && B && ep, "gemm: null operand");
  B4CP_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  B4CP_CHECK_ARG((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0,
                 "gemm: leading dimensions must be multiples of 8 elements (lda=%ld ldb=%ld)", lda,
                 ldb);
  B4CP_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0,
                 "gemm: operands must be 16-byte aligned");
  B4CP_CHECK_ARG(ep->out_f32 || ep->out_bf16, "gemm: no output requested");
  GemmKernelParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.a_mn = a_mn ? 1 : 0;
  p.b_mn = b_mn ? 1 : 0;
  p.BN = pick_bn(N, p.b_mn);
  p.k_tiles_total = ceil_div(K, BK);
  if (splits < 1) splits = 1;
  const int splits_requested = splits;
  if (splits > p.k_tiles_total) splits = p.k_tiles_total;
  p.k_tiles_per_split = ceil_div(p.k_tiles_total, splits);
  splits = ceil_div(p.k_tiles_total, p.k_tiles_per_split);  // no empty K ranges
  if (splits < splits_requested && ep->out_f32) {
    // the caller will sum `splits_requested` partials: the ones no CTA writes must read as zero
    for (int z = splits; z < splits_requested; ++z)
      B4CP_CUDA(cudaMemsetAsync(ep->out_f32 + (size_t)z * ep->split_stride, 0,
                                ((size_t)(M - 1) * ep->ld_f32 + N) * sizeof(float),
                                (cudaStream_t)stream));
  }
  B4CP_CHECK_ARG(splits == 1 || (ep->out_f32 && !ep->out_bf16 && !ep->bias && !ep->relu &&
                                 !ep->gate && !ep->addend),
                 "gemm: split-K writes raw fp32 partials only");
  p.ep = *ep;
  const int stage_bytes = A_STAGE_BYTES + p.BN * BK * 2;
  const int stage_budget = 200 * 1024;
  int stages = stage_budget / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > 6) stages = 6;
  if (stages > p.k_tiles_per_split) stages = p.k_tiles_per_split < 2 ? 2 : p.k_tiles_per_split;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 64 + p.BN * 4 + 128 +
                      4 * 4096 + 1024;

  CUtensorMap tmA, tmB;
  int rc;
  if (!p.a_mn)
    rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM);
  else
    rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, BK);
  if (rc) return rc;
  if (!p.b_mn)
    rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, BK, p.BN);
  else
    rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, BK);
  if (rc) return rc;

  // persistent row-streaming variant for skinny products
  const int n_mtiles = ceil_div(M, BM);
  if (splits == 1 && N <= p.BN && p.k_tiles_total <= PERSIST_MAX_KT && n_mtiles >= 148 &&
      p.BN >= 32) {
    // outputs through TMA stores when their row strides / bases allow it (16-byte multiples)
    CUtensorMap tmOutB, tmOutF;
    memset(&tmOutB, 0, sizeof(tmOutB));
    memset(&tmOutF, 0, sizeof(tmOutF));
    p.tma_out_bf16 = p.tma_out_f32 = 0;
    if (ep->out_bf16 && (ep->ld_bf16 * 2) % 16 == 0 && ((uintptr_t)ep->out_bf16 & 15) == 0) {
      rc = make_tmap_2d(&tmOutB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B,
                        ep->out_bf16, (uint64_t)N, (uint64_t)M, (uint64_t)ep->ld_bf16 * 2, 32, 32);
      if (rc) return rc;
      p.tma_out_bf16 = 1;
    }
    if (ep->out_f32 && (ep->ld_f32 * 4) % 16 == 0 && ((uintptr_t)ep->out_f32 & 15) == 0) {
      rc = make_tmap_2d(&tmOutF, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B,
                        ep->out_f32, (uint64_t)N, (uint64_t)M, (uint64_t)ep->ld_f32 * 4, 32, 32);
      if (rc) return rc;
      p.tma_out_f32 = 1;
    }
    const size_t psmem = (size_t)p.k_tiles_total * (p.BN * BK * 2) +
                         2 * (size_t)p.k_tiles_total * A_STAGE_BYTES +
                         (p.tma_out_f32 ? PE_WARPS * 4096 : 0) + (p.tma_out_bf16 ? PE_WARPS * 2048 : 0) + 128 +
                         p.BN * 4 + 1024;
    if (psmem <= 227 * 1024) {
      static bool pattr = false;
      if (!pattr) {
        B4CP_CUDA(cudaFuncSetAttribute(gemm_umma_persistent_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        pattr = true;
      }
      const int grid = std::min(n_mtiles, 148);
      gemm_umma_persistent_kernel<<<grid, PE_THREADS, psmem, (cudaStream_t)stream>>>(tmA, tmB, tmOutB,
                                                                              tmOutF, p);
      note_launches(1);
      B4CP_LAUNCH_CHECK();
      return 0;
    }
  }
  // more tiles than SMs: tile-persistent kernel with accumulator ping-pong
  const int n_ntiles = ceil_div(N, p.BN);
  const long total_tiles = (long)n_mtiles * n_ntiles * splits;
  if (total_tiles > 148 && total_tiles < (1L << 30) && !getenv("B4CP_GEMM_NO_TILES")) {
    int tstages = (200 * 1024) / stage_bytes;
    if (tstages > 6) tstages = 6;
    if (tstages < 2) tstages = 2;
    p.stages = tstages;
    int tstages_fit = (int)((227 * 1024 - (TILES_EPI_WARPS * 4096 + 2 * p.BN * 4 + 1024 + 512)) / stage_bytes);
    if (tstages > tstages_fit) tstages = tstages_fit;
    p.stages = tstages;
    const size_t tsmem = (size_t)tstages * stage_bytes + (2 * tstages + 4) * 8 + 64 + 2 * p.BN * 4 +
                         128 + TILES_EPI_WARPS * 4096 + 1024;
    static bool tattr = false;
    if (!tattr) {
      B4CP_CUDA(cudaFuncSetAttribute(gemm_umma_tiles_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      tattr = true;
    }
    gemm_umma_tiles_kernel<<<148, TILES_THREADS, tsmem, (cudaStream_t)stream>>>(tmA, tmB, p, n_mtiles, n_ntiles,
                                                                     (int)total_tiles);
    note_launches(1);
    B4CP_LAUNCH_CHECK();
    return 0;
  }
  static bool attr_set = false;
  if (!attr_set) {
    B4CP_CUDA(cudaFuncSetAttribute(gemm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   227 * 1024));
    attr_set = true;
  }
  dim3 grid(ceil_div(M, BM), ceil_div(N, p.BN), splits);
  gemm_umma_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_gemm_splits_for(int M, int N, int K) {
  // enough K-splits to put roughly one wave of CTAs on 148 SMs
  const int bn = pick_bn(N, 1);
  const long tiles = (long)ceil_div(M, BM) * ceil_div(N, bn);
  const int kt = ceil_div(K, BK);
  long s = (148 + tiles - 1) / tiles;
  if (s > kt) s = kt;
  if (s < 1) s = 1;
  // the count b4cp_gemm_bf16 launches: equal K-tile ranges, none empty
  const int per = ceil_div(kt, (int)s);
  return ceil_div(kt, per);
}
