// Tensor-core path of the fused short-sequence masked self-attention (S <= 256, head depth 32
// or 64): transformer.py:64-97 (scaled_dot_product_attention) and :130-156 (split / merge heads).
// One CTA per (sequence, head); Q/K/V (and dO) live in shared memory as row-major bf16 tiles
// (staged with 128-bit loads; transposed operands come from ldmatrix.trans), the S x S score
// matrix only ever exists as mma.sync accumulator fragments in registers (these per-head products
// are 64x64x32: too small for a tcgen05 128-row tile, so the warp-level HMMA path is used here and
// tcgen05 is kept for the Dense layers and the vocabulary stage).
//
// Forward : S = Q K^T / sqrt(dh) + mask, P = softmax(S), O = P V, lse saved.
// Backward: pass A (query blocks): P, dP = dO V^T, delta = rowsum(P o dP), dZ = P o (dP - delta),
//           dQ = dZ K / sqrt(dh).  Pass B (key blocks): the transposed products give
//           dV = P^T dO and dK = dZ^T Q / sqrt(dh) with no atomics.
#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 "
      "{%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// A fragment (16 rows x 16 k) of a row-major bf16 tile with row stride `ld` elements
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __nv_bfloat16* tile, int ld, int r0,
                                       int k0, int g, int t) {
  const __nv_bfloat16* p0 = tile + (size_t)(r0 + g) * ld + k0 + t * 2;
  const __nv_bfloat16* p1 = p0 + 8 * ld;
  a[0] = *reinterpret_cast<const uint32_t*>(p0);
  a[1] = *reinterpret_cast<const uint32_t*>(p1);
  a[2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(p1 + 8);
}
// B fragment (16 k x 8 n) where B[k][n] = tile[n0 + n][k0 + k] (tile row-major, k contiguous)
__device__ __forceinline__ void load_b(uint32_t& b0, uint32_t& b1, const __nv_bfloat16* tile,
                                       int ld, int n0, int k0, int g, int t) {
  const __nv_bfloat16* p = tile + (size_t)(n0 + g) * ld + k0 + t * 2;
  b0 = *reinterpret_cast<const uint32_t*>(p);
  b1 = *reinterpret_cast<const uint32_t*>(p + 8);
}

// Two B fragments (16 k x 8 n each, n-tiles n0 and n0 + 8) where B[k][n] = tile[k0 + k][n0 + n]
// (tile row-major, n contiguous): ldmatrix.trans transposes the 8x8 blocks on the way in, so the
// P.V / dZ.K / P^T.dO / dZ^T.Q products read the row-major tiles directly (no transposed copies).
__device__ __forceinline__ void load_b_trans2(uint32_t (&b)[4], const __nv_bfloat16* tile, int ld,
                                              int k0, int n0, int lane) {
  const int row = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int col = n0 + (lane >> 4) * 8;
  const uint32_t addr = smem_u32(tile + (size_t)row * ld + col);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
               : "r"(addr));
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// cooperative 128-bit load of one head slice [S][DH] (global row stride ld elements, 16-byte
// aligned rows) into a row-major smem tile [SP][LDR]; rows >= S are zero
template <int DH, int SP>
__device__ __forceinline__ void stage_tile(const __nv_bfloat16* __restrict__ src, long ld, int S,
                                           __nv_bfloat16* rowm, int LDR) {
  constexpr int W = DH / 8;
#pragma unroll 2
  for (int i = threadIdx.x; i < SP * W; i += blockDim.x) {
    const int r = i / W, w = i - r * W;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < S) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * ld + 8 * w));
    *reinterpret_cast<uint4*>(rowm + (size_t)r * LDR + 8 * w) = v;
  }
}

// The Q / K / V (/ dO) slices of one (sequence, head), staged TOGETHER: every thread first issues
// all of its 128-bit global loads (6 or 8 independent requests at the C1 shape), then stores them.
// Tile by tile the loads of the next tile waited behind the shared-memory stores of the previous
// one, so a CTA paid three to four DRAM round trips before its first MMA; these kernels move
// 0.3 of the HBM roofline and are latency-, not flop-bound (ncu: profiles/r2_ncu_attn_*).
template <int DH, int SP, int NTILES, int THREADS>
__device__ __forceinline__ void stage_tiles(const __nv_bfloat16* const (&src)[NTILES],
                                            const long (&ld)[NTILES], int S,
                                            __nv_bfloat16* const (&dst)[NTILES], int LDR) {
  constexpr int W = DH / 8;
  constexpr int PER = (SP * W + THREADS - 1) / THREADS;   // uint4 per thread per tile
  if constexpr (PER * NTILES <= 16) {
    uint4 v[NTILES][PER];
#pragma unroll
    for (int k = 0; k < NTILES; ++k)
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int i = threadIdx.x + u * THREADS;
        const int r = i / W, w = i - r * W;
        v[k][u] = make_uint4(0u, 0u, 0u, 0u);
        if (i < SP * W && r < S) v[k][u] = __ldg(reinterpret_cast<const uint4*>(src[k] + (size_t)r * ld[k] + 8 * w));
      }
#pragma unroll
    for (int k = 0; k < NTILES; ++k)
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int i = threadIdx.x + u * THREADS;
        const int r = i / W, w = i - r * W;
        if (i < SP * W) *reinterpret_cast<uint4*>(dst[k] + (size_t)r * LDR + 8 * w) = v[k][u];
      }
  } else {
#pragma unroll
    for (int k = 0; k < NTILES; ++k) stage_tile<DH, SP>(src[k], ld[k], S, dst[k], LDR);
  }
}

template <int NKB, int DH>
struct AttnCfg {
  static constexpr int SP = 64 * NKB;       // padded sequence length
  static constexpr int LDR = DH + 8;        // row-major tiles: conflict-free fragment loads
  static constexpr int LDT = SP + 8;        // transposed tiles
  static constexpr int NT = SP / 8;         // 8-wide key tiles
  static constexpr int KS_D = DH / 16;      // k-steps over the head depth
  static constexpr int KS_S = SP / 16;      // k-steps over the sequence
  static constexpr int NT_D = DH / 8;
};

// ------------------------------------------------------------------------------- forward
template <int NKB, int DH>
__global__ void __launch_bounds__(128)
attention_mma_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ ids,
                         int S, int H, __nv_bfloat16* __restrict__ out,
                         float* __restrict__ lse_out) {
  using C = AttnCfg<NKB, DH>;
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sm);
  __nv_bfloat16* sK = sQ + C::SP * C::LDR;
  __nv_bfloat16* sV = sK + C::SP * C::LDR;
  float* sMask = reinterpret_cast<float*>(sV + C::SP * C::LDR);
  const int d = H * DH;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * d + h * DH;
  {
    const __nv_bfloat16* const src[3] = {base, base + d, base + 2 * d};
    const long lds[3] = {3L * d, 3L * d, 3L * d};
    __nv_bfloat16* const dst[3] = {sQ, sK, sV};
    stage_tiles<DH, C::SP, 3, 128>(src, lds, S, dst, C::LDR);
  }
  for (int j = threadIdx.x; j < C::SP; j += blockDim.x)
    sMask[j] = j >= S ? -INFINITY : (ids[(size_t)b * S + j] == 0 ? -1e9f : 0.f);
  __syncthreads();
  // scores are bf16 tensor-core products: multiplying by 1/sqrt(dh) instead of the reference's
  // exact division (transformer.py:86) moves them by at most 1 fp32 ulp and saves a ~12-instruction
  // division routine per score (it was ~25% of the backward's instructions)
  const float sqrt_dh = sqrtf((float)DH);
  const float rsqrt_dh = 1.f / sqrt_dh;
  for (int rb = warp; rb * 16 < S; rb += 4) {
    const int r0 = rb * 16;
    float sc[C::NT][4];
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KS_D; ++ks) {
      uint32_t a[4];
      load_a(a, sQ, C::LDR, r0, ks * 16, g, t);
#pragma unroll
      for (int nt = 0; nt < C::NT; ++nt) {
        uint32_t b0, b1;
        load_b(b0, b1, sK, C::LDR, nt * 8, ks * 16, g, t);
        mma_bf16(sc[nt], a, b0, b1);
      }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      const float k0 = sMask[nt * 8 + t * 2], k1 = sMask[nt * 8 + t * 2 + 1];
      sc[nt][0] = (sc[nt][0] * rsqrt_dh) + k0;
      sc[nt][1] = (sc[nt][1] * rsqrt_dh) + k1;
      sc[nt][2] = (sc[nt][2] * rsqrt_dh) + k0;
      sc[nt][3] = (sc[nt][3] * rsqrt_dh) + k1;
      m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
      m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      sc[nt][0] = __expf(sc[nt][0] - m0);
      sc[nt][1] = __expf(sc[nt][1] - m0);
      sc[nt][2] = __expf(sc[nt][2] - m1);
      sc[nt][3] = __expf(sc[nt][3] - m1);
      s0 += sc[nt][0] + sc[nt][1];
      s1 += sc[nt][2] + sc[nt][3];
    }
    s0 = quad_sum(s0);
    s1 = quad_sum(s1);
    const float i0 = 1.f / s0, i1 = 1.f / s1;
    float o[C::NT_D][4];
#pragma unroll
    for (int n2 = 0; n2 < C::NT_D; ++n2) o[n2][0] = o[n2][1] = o[n2][2] = o[n2][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KS_S; ++ks) {
      uint32_t a[4];
      a[0] = pack_bf16(sc[2 * ks][0] * i0, sc[2 * ks][1] * i0);
      a[1] = pack_bf16(sc[2 * ks][2] * i1, sc[2 * ks][3] * i1);
      a[2] = pack_bf16(sc[2 * ks + 1][0] * i0, sc[2 * ks + 1][1] * i0);
      a[3] = pack_bf16(sc[2 * ks + 1][2] * i1, sc[2 * ks + 1][3] * i1);
#pragma unroll
      for (int n2 = 0; n2 < C::NT_D; n2 += 2) {
        uint32_t bb[4];
        load_b_trans2(bb, sV, C::LDR, ks * 16, n2 * 8, lane);
        mma_bf16(o[n2], a, bb[0], bb[1]);
        mma_bf16(o[n2 + 1], a, bb[2], bb[3]);
      }
    }
    const int row0 = r0 + g, row1 = r0 + g + 8;
#pragma unroll
    for (int n2 = 0; n2 < C::NT_D; ++n2) {
      const int col = h * DH + n2 * 8 + t * 2;
      if (row0 < S)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * S + row0) * d + col) = pack_bf16(o[n2][0], o[n2][1]);
      if (row1 < S)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * S + row1) * d + col) = pack_bf16(o[n2][2], o[n2][3]);
    }
    if (lse_out && t == 0) {
      if (row0 < S) lse_out[((size_t)b * H + h) * S + row0] = m0 + logf(s0);
      if (row1 < S) lse_out[((size_t)b * H + h) * S + row1] = m1 + logf(s1);
    }
  }
}

// ------------------------------------------------------------------------------- backward
template <int NKB, int DH>
__global__ void __launch_bounds__(128)
attention_mma_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                         const float* __restrict__ lse_in, const int32_t* __restrict__ ids, int S,
                         int H, __nv_bfloat16* __restrict__ dqkv) {
  using C = AttnCfg<NKB, DH>;
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sm);
  __nv_bfloat16* sK = sQ + C::SP * C::LDR;
  __nv_bfloat16* sV = sK + C::SP * C::LDR;
  __nv_bfloat16* sDO = sV + C::SP * C::LDR;
  float* sMask = reinterpret_cast<float*>(sDO + C::SP * C::LDR);
  float* sLse = sMask + C::SP;
  float* sDelta = sLse + C::SP;
  const int d = H * DH;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * d + h * DH;
  {
    const __nv_bfloat16* const src[4] = {base, base + d, base + 2 * d, dout + (size_t)b * S * d + h * DH};
    const long lds[4] = {3L * d, 3L * d, 3L * d, (long)d};
    __nv_bfloat16* const dst[4] = {sQ, sK, sV, sDO};
    stage_tiles<DH, C::SP, 4, 128>(src, lds, S, dst, C::LDR);
  }
  for (int j = threadIdx.x; j < C::SP; j += blockDim.x) {
    sMask[j] = j >= S ? -INFINITY : (ids[(size_t)b * S + j] == 0 ? -1e9f : 0.f);
    sLse[j] = j < S ? lse_in[((size_t)b * H + h) * S + j] : INFINITY;  // rows past S: P = 0
    sDelta[j] = 0.f;
  }
  __syncthreads();
  // scores are bf16 tensor-core products: multiplying by 1/sqrt(dh) instead of the reference's
  // exact division (transformer.py:86) moves them by at most 1 fp32 ulp and saves a ~12-instruction
  // division routine per score (it was ~25% of the backward's instructions)
  const float sqrt_dh = sqrtf((float)DH);
  const float rsqrt_dh = 1.f / sqrt_dh;
  const float inv_sqrt = 1.f / sqrt_dh;
  __nv_bfloat16* dst = dqkv + (size_t)b * S * 3 * d + h * DH;

  // ---- pass A: query row blocks -> delta, dQ
  for (int rb = warp; rb * 16 < S; rb += 4) {
    const int r0 = rb * 16;
    float p[C::NT][4], dp[C::NT][4];
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < C::KS_D; ++ks) {
      uint32_t aq[4], ag[4];
      load_a(aq, sQ, C::LDR, r0, ks * 16, g, t);
      load_a(ag, sDO, C::LDR, r0, ks * 16, g, t);
#pragma unroll
      for (int nt = 0; nt < C::NT; ++nt) {
        uint32_t b0, b1;
        load_b(b0, b1, sK, C::LDR, nt * 8, ks * 16, g, t);
        mma_bf16(p[nt], aq, b0, b1);
        load_b(b0, b1, sV, C::LDR, nt * 8, ks * 16, g, t);
        mma_bf16(dp[nt], ag, b0, b1);
      }
    }
    const float l0 = sLse[r0 + g], l1 = sLse[r0 + g + 8];
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      const float k0 = sMask[nt * 8 + t * 2], k1 = sMask[nt * 8 + t * 2 + 1];
      p[nt][0] = __expf((p[nt][0] * rsqrt_dh) + k0 - l0);
      p[nt][1] = __expf((p[nt][1] * rsqrt_dh) + k1 - l0);
      p[nt][2] = __expf((p[nt][2] * rsqrt_dh) + k0 - l1);
      p[nt][3] = __expf((p[nt][3] * rsqrt_dh) + k1 - l1);
      d0 += p[nt][0] * dp[nt][0] + p[nt][1] * dp[nt][1];
      d1 += p[nt][2] * dp[nt][2] + p[nt][3] * dp[nt][3];
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
    if (t == 0) {
      sDelta[r0 + g] = d0;
      sDelta[r0 + g + 8] = d1;
    }
    float dq[C::NT_D][4];
#pragma unroll
    for (int n2 = 0; n2 < C::NT_D; ++n2) dq[n2][0] = dq[n2][1] = dq[n2][2] = dq[n2][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KS_S; ++ks) {
      uint32_t a[4];
      a[0] = pack_bf16(p[2 * ks][0] * (dp[2 * ks][0] - d0), p[2 * ks][1] * (dp[2 * ks][1] - d0));
      a[1] = pack_bf16(p[2 * ks][2] * (dp[2 * ks][2] - d1), p[2 * ks][3] * (dp[2 * ks][3] - d1));
      a[2] = pack_bf16(p[2 * ks + 1][0] * (dp[2 * ks + 1][0] - d0), p[2 * ks + 1][1] * (dp[2 * ks + 1][1] - d0));
      a[3] = pack_bf16(p[2 * ks + 1][2] * (dp[2 * ks + 1][2] - d1), p[2 * ks + 1][3] * (dp[2 * ks + 1][3] - d1));
#pragma unroll
      for (int n2 = 0; n2 < C::NT_D; n2 += 2) {
        uint32_t bb[4];
        load_b_trans2(bb, sK, C::LDR, ks * 16, n2 * 8, lane);
        mma_bf16(dq[n2], a, bb[0], bb[1]);
        mma_bf16(dq[n2 + 1], a, bb[2], bb[3]);
      }
    }
    const int row0 = r0 + g, row1 = r0 + g + 8;
#pragma unroll
    for (int n2 = 0; n2 < C::NT_D; ++n2) {
      const int col = n2 * 8 + t * 2;
      if (row0 < S)
        *reinterpret_cast<uint32_t*>(dst + (size_t)row0 * 3 * d + col) =
            pack_bf16(dq[n2][0] * inv_sqrt, dq[n2][1] * inv_sqrt);
      if (row1 < S)
        *reinterpret_cast<uint32_t*>(dst + (size_t)row1 * 3 * d + col) =
            pack_bf16(dq[n2][2] * inv_sqrt, dq[n2][3] * inv_sqrt);
    }
  }
  __syncthreads();

  // ---- pass B: key row blocks -> dK, dV (transposed score tiles: rows = keys, columns = queries)
  for (int kb = warp; kb * 16 < S; kb += 4) {
    const int j0 = kb * 16;
    float p[C::NT][4], dp[C::NT][4];
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < C::KS_D; ++ks) {
      uint32_t ak[4], av[4];
      load_a(ak, sK, C::LDR, j0, ks * 16, g, t);
      load_a(av, sV, C::LDR, j0, ks * 16, g, t);
#pragma unroll
      for (int nt = 0; nt < C::NT; ++nt) {
        uint32_t b0, b1;
        load_b(b0, b1, sQ, C::LDR, nt * 8, ks * 16, g, t);
        mma_bf16(p[nt], ak, b0, b1);
        load_b(b0, b1, sDO, C::LDR, nt * 8, ks * 16, g, t);
        mma_bf16(dp[nt], av, b0, b1);
      }
    }
    const float k0m = sMask[j0 + g], k1m = sMask[j0 + g + 8];
#pragma unroll
    for (int nt = 0; nt < C::NT; ++nt) {
      const int q0 = nt * 8 + t * 2;
      const float la = sLse[q0], lb = sLse[q0 + 1];
      const float da = sDelta[q0], db = sDelta[q0 + 1];
      p[nt][0] = __expf((p[nt][0] * rsqrt_dh) + k0m - la);
      p[nt][1] = __expf((p[nt][1] * rsqrt_dh) + k0m - lb);
      p[nt][2] = __expf((p[nt][2] * rsqrt_dh) + k1m - la);
      p[nt][3] = __expf((p[nt][3] * rsqrt_dh) + k1m - lb);
      dp[nt][0] = p[nt][0] * (dp[nt][0] - da);
      dp[nt][1] = p[nt][1] * (dp[nt][1] - db);
      dp[nt][2] = p[nt][2] * (dp[nt][2] - da);
      dp[nt][3] = p[nt][3] * (dp[nt][3] - db);
    }
    float dk[C::NT_D][4], dv[C::NT_D][4];
#pragma unroll
    for (int n2 = 0; n2 < C::NT_D; ++n2) {
      dk[n2][0] = dk[n2][1] = dk[n2][2] = dk[n2][3] = 0.f;
      dv[n2][0] = dv[n2][1] = dv[n2][2] = dv[n2][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < C::KS_S; ++ks) {
      uint32_t ap[4], az[4];
      ap[0] = pack_bf16(p[2 * ks][0], p[2 * ks][1]);
      ap[1] = pack_bf16(p[2 * ks][2], p[2 * ks][3]);
      ap[2] = pack_bf16(p[2 * ks + 1][0], p[2 * ks + 1][1]);
      ap[3] = pack_bf16(p[2 * ks + 1][2], p[2 * ks + 1][3]);
      az[0] = pack_bf16(dp[2 * ks][0], dp[2 * ks][1]);
      az[1] = pack_bf16(dp[2 * ks][2], dp[2 * ks][3]);
      az[2] = pack_bf16(dp[2 * ks + 1][0], dp[2 * ks + 1][1]);
      az[3] = pack_bf16(dp[2 * ks + 1][2], dp[2 * ks + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < C::NT_D; n2 += 2) {
        uint32_t bb[4];
        load_b_trans2(bb, sDO, C::LDR, ks * 16, n2 * 8, lane);
        mma_bf16(dv[n2], ap, bb[0], bb[1]);
        mma_bf16(dv[n2 + 1], ap, bb[2], bb[3]);
        load_b_trans2(bb, sQ, C::LDR, ks * 16, n2 * 8, lane);
        mma_bf16(dk[n2], az, bb[0], bb[1]);
        mma_bf16(dk[n2 + 1], az, bb[2], bb[3]);
      }
    }
    const int row0 = j0 + g, row1 = j0 + g + 8;
#pragma unroll
    for (int n2 = 0; n2 < C::NT_D; ++n2) {
      const int col = n2 * 8 + t * 2;
      if (row0 < S) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)row0 * 3 * d + d + col) =
            pack_bf16(dk[n2][0] * inv_sqrt, dk[n2][1] * inv_sqrt);
        *reinterpret_cast<uint32_t*>(dst + (size_t)row0 * 3 * d + 2 * d + col) = pack_bf16(dv[n2][0], dv[n2][1]);
      }
      if (row1 < S) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)row1 * 3 * d + d + col) =
            pack_bf16(dk[n2][2] * inv_sqrt, dk[n2][3] * inv_sqrt);
        *reinterpret_cast<uint32_t*>(dst + (size_t)row1 * 3 * d + 2 * d + col) = pack_bf16(dv[n2][2], dv[n2][3]);
      }
    }
  }
}

// =============================================================================== 128 < S <= 256
// The same products for sequences of up to 256 positions (S = 202 / 203).  The
// score tile of one 16-row block no longer fits in registers (32 key tiles x 4 accumulators, twice
// that in the backward), so keys (forward, dQ pass) / queries (dK, dV pass) are walked in two
// halves of 128: the forward keeps a running (max, sum) and rescales O, flash-attention style; the
// backward takes delta = rowsum(dO o O) from the saved forward output instead of a sweep over all
// keys.  256 threads per CTA (16 row blocks, two per warp).
static constexpr int LSP = 256;   // padded sequence length
static constexpr int LHALF = 128;

template <int DH>
__global__ void __launch_bounds__(256)
attention_mma_fwd_long_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ ids,
                              int S, int H, __nv_bfloat16* __restrict__ out,
                              float* __restrict__ lse_out) {
  constexpr int LDR = DH + 8, KS_D = DH / 16, NT_D = DH / 8, NT = LHALF / 8, KS_S = LHALF / 16;
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sm);
  __nv_bfloat16* sK = sQ + LSP * LDR;
  __nv_bfloat16* sV = sK + LSP * LDR;
  float* sMask = reinterpret_cast<float*>(sV + LSP * LDR);
  const int d = H * DH;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nwarps = blockDim.x >> 5;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * d + h * DH;
  stage_tile<DH, LSP>(base, 3L * d, S, sQ, LDR);
  stage_tile<DH, LSP>(base + d, 3L * d, S, sK, LDR);
  stage_tile<DH, LSP>(base + 2 * d, 3L * d, S, sV, LDR);
  for (int j = threadIdx.x; j < LSP; j += blockDim.x)
    sMask[j] = j >= S ? -INFINITY : (ids[(size_t)b * S + j] == 0 ? -1e9f : 0.f);
  __syncthreads();
  // scores are bf16 tensor-core products: multiplying by 1/sqrt(dh) instead of the reference's
  // exact division (transformer.py:86) moves them by at most 1 fp32 ulp and saves a ~12-instruction
  // division routine per score (it was ~25% of the backward's instructions)
  const float sqrt_dh = sqrtf((float)DH);
  const float rsqrt_dh = 1.f / sqrt_dh;
  for (int rb = warp; rb * 16 < S; rb += nwarps) {
    const int r0 = rb * 16;
    float o[NT_D][4];
#pragma unroll
    for (int n2 = 0; n2 < NT_D; ++n2) o[n2][0] = o[n2][1] = o[n2][2] = o[n2][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
#pragma unroll 1
    for (int kh = 0; kh < 2; ++kh) {
      const int kb0 = kh * LHALF;
      if (kb0 >= S) break;
      float sc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS_D; ++ks) {
        uint32_t a[4];
        load_a(a, sQ, LDR, r0, ks * 16, g, t);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint32_t b0, b1;
          load_b(b0, b1, sK, LDR, kb0 + nt * 8, ks * 16, g, t);
          mma_bf16(sc[nt], a, b0, b1);
        }
      }
      float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float k0 = sMask[kb0 + nt * 8 + t * 2], k1 = sMask[kb0 + nt * 8 + t * 2 + 1];
        sc[nt][0] = (sc[nt][0] * rsqrt_dh) + k0;
        sc[nt][1] = (sc[nt][1] * rsqrt_dh) + k1;
        sc[nt][2] = (sc[nt][2] * rsqrt_dh) + k0;
        sc[nt][3] = (sc[nt][3] * rsqrt_dh) + k1;
        c0 = fmaxf(c0, fmaxf(sc[nt][0], sc[nt][1]));
        c1 = fmaxf(c1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      const float n0 = fmaxf(m0, quad_max(c0)), n1 = fmaxf(m1, quad_max(c1));   // finite: key 0 exists
      const float a0 = __expf(m0 - n0), a1 = __expf(m1 - n1);                        // exp(-inf) = 0 at first
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sc[nt][0] = __expf(sc[nt][0] - n0);
        sc[nt][1] = __expf(sc[nt][1] - n0);
        sc[nt][2] = __expf(sc[nt][2] - n1);
        sc[nt][3] = __expf(sc[nt][3] - n1);
        s0 += sc[nt][0] + sc[nt][1];
        s1 += sc[nt][2] + sc[nt][3];
      }
      l0 = l0 * a0 + quad_sum(s0);
      l1 = l1 * a1 + quad_sum(s1);
      m0 = n0;
      m1 = n1;
#pragma unroll
      for (int n2 = 0; n2 < NT_D; ++n2) {
        o[n2][0] *= a0; o[n2][1] *= a0; o[n2][2] *= a1; o[n2][3] *= a1;
      }
#pragma unroll
      for (int ks = 0; ks < KS_S; ++ks) {
        uint32_t a[4];
        a[0] = pack_bf16(sc[2 * ks][0], sc[2 * ks][1]);
        a[1] = pack_bf16(sc[2 * ks][2], sc[2 * ks][3]);
        a[2] = pack_bf16(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
        a[3] = pack_bf16(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
#pragma unroll
        for (int n2 = 0; n2 < NT_D; n2 += 2) {
          uint32_t bb[4];
          load_b_trans2(bb, sV, LDR, kb0 + ks * 16, n2 * 8, lane);
          mma_bf16(o[n2], a, bb[0], bb[1]);
          mma_bf16(o[n2 + 1], a, bb[2], bb[3]);
        }
      }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int row0 = r0 + g, row1 = r0 + g + 8;
#pragma unroll
    for (int n2 = 0; n2 < NT_D; ++n2) {
      const int col = h * DH + n2 * 8 + t * 2;
      if (row0 < S)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * S + row0) * d + col) =
            pack_bf16(o[n2][0] * i0, o[n2][1] * i0);
      if (row1 < S)
        *reinterpret_cast<uint32_t*>(out + ((size_t)b * S + row1) * d + col) =
            pack_bf16(o[n2][2] * i1, o[n2][3] * i1);
    }
    if (lse_out && t == 0) {
      if (row0 < S) lse_out[((size_t)b * H + h) * S + row0] = m0 + logf(l0);
      if (row1 < S) lse_out[((size_t)b * H + h) * S + row1] = m1 + logf(l1);
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(256)
attention_mma_bwd_long_kernel(const __nv_bfloat16* __restrict__ qkv,
                              const __nv_bfloat16* __restrict__ fwd_out,
                              const __nv_bfloat16* __restrict__ dout,
                              const float* __restrict__ lse_in, const int32_t* __restrict__ ids,
                              int S, int H, __nv_bfloat16* __restrict__ dqkv) {
  constexpr int LDR = DH + 8, KS_D = DH / 16, NT_D = DH / 8, NT = LHALF / 8, KS_S = LHALF / 16;
  extern __shared__ __align__(16) uint8_t sm[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(sm);
  __nv_bfloat16* sK = sQ + LSP * LDR;
  __nv_bfloat16* sV = sK + LSP * LDR;
  __nv_bfloat16* sDO = sV + LSP * LDR;
  float* sMask = reinterpret_cast<float*>(sDO + LSP * LDR);
  float* sLse = sMask + LSP;
  float* sDelta = sLse + LSP;
  const int d = H * DH;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nwarps = blockDim.x >> 5;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * d + h * DH;
  stage_tile<DH, LSP>(base, 3L * d, S, sQ, LDR);
  stage_tile<DH, LSP>(base + d, 3L * d, S, sK, LDR);
  stage_tile<DH, LSP>(base + 2 * d, 3L * d, S, sV, LDR);
  stage_tile<DH, LSP>(dout + (size_t)b * S * d + h * DH, (long)d, S, sDO, LDR);
  for (int j = threadIdx.x; j < LSP; j += blockDim.x) {
    sMask[j] = j >= S ? -INFINITY : (ids[(size_t)b * S + j] == 0 ? -1e9f : 0.f);
    sLse[j] = j < S ? lse_in[((size_t)b * H + h) * S + j] : INFINITY;  // rows past S: P = 0
  }
  __syncthreads();
  // delta[r] = sum_c dO[r][c] * O[r][c]  (= rowsum(P o dP)); one warp per row
  for (int r = warp; r < LSP; r += nwarps) {
    float acc = 0.f;
    if (r < S) {
      const __nv_bfloat16* orow = fwd_out + ((size_t)b * S + r) * d + h * DH;
      for (int c = lane * 2; c < DH; c += 64) {
        const __nv_bfloat162 ov = *reinterpret_cast<const __nv_bfloat162*>(orow + c);
        const __nv_bfloat162 gv = *reinterpret_cast<const __nv_bfloat162*>(sDO + (size_t)r * LDR + c);
        acc += __bfloat162float(ov.x) * __bfloat162float(gv.x) + __bfloat162float(ov.y) * __bfloat162float(gv.y);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) sDelta[r] = acc;
  }
  __syncthreads();
  // scores are bf16 tensor-core products: multiplying by 1/sqrt(dh) instead of the reference's
  // exact division (transformer.py:86) moves them by at most 1 fp32 ulp and saves a ~12-instruction
  // division routine per score (it was ~25% of the backward's instructions)
  const float sqrt_dh = sqrtf((float)DH);
  const float rsqrt_dh = 1.f / sqrt_dh;
  const float inv_sqrt = 1.f / sqrt_dh;
  __nv_bfloat16* dst = dqkv + (size_t)b * S * 3 * d + h * DH;

  // ---- pass A: query row blocks -> dQ (keys in two halves)
  for (int rb = warp; rb * 16 < S; rb += nwarps) {
    const int r0 = rb * 16;
    float dq[NT_D][4];
#pragma unroll
    for (int n2 = 0; n2 < NT_D; ++n2) dq[n2][0] = dq[n2][1] = dq[n2][2] = dq[n2][3] = 0.f;
    const float l0 = sLse[r0 + g], l1 = sLse[r0 + g + 8];
    const float d0 = sDelta[r0 + g], d1 = sDelta[r0 + g + 8];
#pragma unroll 1
    for (int kh = 0; kh < 2; ++kh) {
      const int kb0 = kh * LHALF;
      if (kb0 >= S) break;
      float p[NT][4], dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < KS_D; ++ks) {
        uint32_t aq[4], ag[4];
        load_a(aq, sQ, LDR, r0, ks * 16, g, t);
        load_a(ag, sDO, LDR, r0, ks * 16, g, t);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint32_t b0, b1;
          load_b(b0, b1, sK, LDR, kb0 + nt * 8, ks * 16, g, t);
          mma_bf16(p[nt], aq, b0, b1);
          load_b(b0, b1, sV, LDR, kb0 + nt * 8, ks * 16, g, t);
          mma_bf16(dp[nt], ag, b0, b1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float k0 = sMask[kb0 + nt * 8 + t * 2], k1 = sMask[kb0 + nt * 8 + t * 2 + 1];
        p[nt][0] = __expf((p[nt][0] * rsqrt_dh) + k0 - l0) * (dp[nt][0] - d0);
        p[nt][1] = __expf((p[nt][1] * rsqrt_dh) + k1 - l0) * (dp[nt][1] - d0);
        p[nt][2] = __expf((p[nt][2] * rsqrt_dh) + k0 - l1) * (dp[nt][2] - d1);
        p[nt][3] = __expf((p[nt][3] * rsqrt_dh) + k1 - l1) * (dp[nt][3] - d1);
      }
#pragma unroll
      for (int ks = 0; ks < KS_S; ++ks) {
        uint32_t a[4];
        a[0] = pack_bf16(p[2 * ks][0], p[2 * ks][1]);
        a[1] = pack_bf16(p[2 * ks][2], p[2 * ks][3]);
        a[2] = pack_bf16(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        a[3] = pack_bf16(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
        for (int n2 = 0; n2 < NT_D; n2 += 2) {
          uint32_t bb[4];
          load_b_trans2(bb, sK, LDR, kb0 + ks * 16, n2 * 8, lane);
          mma_bf16(dq[n2], a, bb[0], bb[1]);
          mma_bf16(dq[n2 + 1], a, bb[2], bb[3]);
        }
      }
    }
    const int row0 = r0 + g, row1 = r0 + g + 8;
#pragma unroll
    for (int n2 = 0; n2 < NT_D; ++n2) {
      const int col = n2 * 8 + t * 2;
      if (row0 < S)
        *reinterpret_cast<uint32_t*>(dst + (size_t)row0 * 3 * d + col) =
            pack_bf16(dq[n2][0] * inv_sqrt, dq[n2][1] * inv_sqrt);
      if (row1 < S)
        *reinterpret_cast<uint32_t*>(dst + (size_t)row1 * 3 * d + col) =
            pack_bf16(dq[n2][2] * inv_sqrt, dq[n2][3] * inv_sqrt);
    }
  }

  // ---- pass B: key row blocks -> dK, dV (transposed tiles; queries in two halves)
  for (int kb = warp; kb * 16 < S; kb += nwarps) {
    const int j0 = kb * 16;
    float dk[NT_D][4], dv[NT_D][4];
#pragma unroll
    for (int n2 = 0; n2 < NT_D; ++n2) {
      dk[n2][0] = dk[n2][1] = dk[n2][2] = dk[n2][3] = 0.f;
      dv[n2][0] = dv[n2][1] = dv[n2][2] = dv[n2][3] = 0.f;
    }
    const float k0m = sMask[j0 + g], k1m = sMask[j0 + g + 8];
#pragma unroll 1
    for (int qh = 0; qh < 2; ++qh) {
      const int qb0 = qh * LHALF;
      if (qb0 >= S) break;
      float p[NT][4], dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        p[nt][0] = p[nt][1] = p[nt][2] = p[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < KS_D; ++ks) {
        uint32_t ak[4], av[4];
        load_a(ak, sK, LDR, j0, ks * 16, g, t);
        load_a(av, sV, LDR, j0, ks * 16, g, t);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint32_t b0, b1;
          load_b(b0, b1, sQ, LDR, qb0 + nt * 8, ks * 16, g, t);
          mma_bf16(p[nt], ak, b0, b1);
          load_b(b0, b1, sDO, LDR, qb0 + nt * 8, ks * 16, g, t);
          mma_bf16(dp[nt], av, b0, b1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int q0 = qb0 + nt * 8 + t * 2;
        const float la = sLse[q0], lb = sLse[q0 + 1];
        const float da = sDelta[q0], db = sDelta[q0 + 1];
        p[nt][0] = __expf((p[nt][0] * rsqrt_dh) + k0m - la);
        p[nt][1] = __expf((p[nt][1] * rsqrt_dh) + k0m - lb);
        p[nt][2] = __expf((p[nt][2] * rsqrt_dh) + k1m - la);
        p[nt][3] = __expf((p[nt][3] * rsqrt_dh) + k1m - lb);
        dp[nt][0] = p[nt][0] * (dp[nt][0] - da);
        dp[nt][1] = p[nt][1] * (dp[nt][1] - db);
        dp[nt][2] = p[nt][2] * (dp[nt][2] - da);
        dp[nt][3] = p[nt][3] * (dp[nt][3] - db);
      }
#pragma unroll
      for (int ks = 0; ks < KS_S; ++ks) {
        uint32_t ap[4], az[4];
        ap[0] = pack_bf16(p[2 * ks][0], p[2 * ks][1]);
        ap[1] = pack_bf16(p[2 * ks][2], p[2 * ks][3]);
        ap[2] = pack_bf16(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        ap[3] = pack_bf16(p[2 * ks + 1][2], p[2 * ks + 1][3]);
        az[0] = pack_bf16(dp[2 * ks][0], dp[2 * ks][1]);
        az[1] = pack_bf16(dp[2 * ks][2], dp[2 * ks][3]);
        az[2] = pack_bf16(dp[2 * ks + 1][0], dp[2 * ks + 1][1]);
        az[3] = pack_bf16(dp[2 * ks + 1][2], dp[2 * ks + 1][3]);
#pragma unroll
        for (int n2 = 0; n2 < NT_D; n2 += 2) {
          uint32_t bb[4];
          load_b_trans2(bb, sDO, LDR, qb0 + ks * 16, n2 * 8, lane);
          mma_bf16(dv[n2], ap, bb[0], bb[1]);
          mma_bf16(dv[n2 + 1], ap, bb[2], bb[3]);
          load_b_trans2(bb, sQ, LDR, qb0 + ks * 16, n2 * 8, lane);
          mma_bf16(dk[n2], az, bb[0], bb[1]);
          mma_bf16(dk[n2 + 1], az, bb[2], bb[3]);
        }
      }
    }
    const int row0 = j0 + g, row1 = j0 + g + 8;
#pragma unroll
    for (int n2 = 0; n2 < NT_D; ++n2) {
      const int col = n2 * 8 + t * 2;
      if (row0 < S) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)row0 * 3 * d + d + col) =
            pack_bf16(dk[n2][0] * inv_sqrt, dk[n2][1] * inv_sqrt);
        *reinterpret_cast<uint32_t*>(dst + (size_t)row0 * 3 * d + 2 * d + col) = pack_bf16(dv[n2][0], dv[n2][1]);
      }
      if (row1 < S) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)row1 * 3 * d + d + col) =
            pack_bf16(dk[n2][2] * inv_sqrt, dk[n2][3] * inv_sqrt);
        *reinterpret_cast<uint32_t*>(dst + (size_t)row1 * 3 * d + 2 * d + col) = pack_bf16(dv[n2][2], dv[n2][3]);
      }
    }
  }
}

template <int DH>
static int launch_fwd_long(const void* qkv, const int32_t* ids, int B, int S, int H, void* out,
                           float* lse, cudaStream_t st) {
  const size_t smem = (size_t)(3 * LSP * (DH + 8)) * 2 + LSP * 4;
  B4CP_CUDA(cudaFuncSetAttribute(attention_mma_fwd_long_kernel<DH>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_mma_fwd_long_kernel<DH><<<B * H, 256, smem, st>>>(
      (const __nv_bfloat16*)qkv, ids, S, H, (__nv_bfloat16*)out, lse);
  return 0;
}

template <int DH>
static int launch_bwd_long(const void* qkv, const void* fwd_out, const void* dout, const float* lse,
                           const int32_t* ids, int B, int S, int H, void* dqkv, cudaStream_t st) {
  const size_t smem = (size_t)(4 * LSP * (DH + 8)) * 2 + 3 * LSP * 4;
  B4CP_CUDA(cudaFuncSetAttribute(attention_mma_bwd_long_kernel<DH>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_mma_bwd_long_kernel<DH><<<B * H, 256, smem, st>>>(
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)fwd_out, (const __nv_bfloat16*)dout, lse, ids,
      S, H, (__nv_bfloat16*)dqkv);
  return 0;
}

template <int NKB, int DH>
static int launch_fwd(const void* qkv, const int32_t* ids, int B, int S, int H, void* out, float* lse,
                      cudaStream_t st) {
  using C = AttnCfg<NKB, DH>;
  const size_t smem = (size_t)(3 * C::SP * C::LDR) * 2 + C::SP * 4;
  B4CP_CUDA(cudaFuncSetAttribute(attention_mma_fwd_kernel<NKB, DH>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_mma_fwd_kernel<NKB, DH><<<B * H, 128, smem, st>>>(
      (const __nv_bfloat16*)qkv, ids, S, H, (__nv_bfloat16*)out, lse);
  return 0;
}

template <int NKB, int DH>
static int launch_bwd(const void* qkv, const void* dout, const float* lse, const int32_t* ids, int B,
                      int S, int H, void* dqkv, cudaStream_t st) {
  using C = AttnCfg<NKB, DH>;
  const size_t smem = (size_t)(4 * C::SP * C::LDR) * 2 + 3 * C::SP * 4;
  B4CP_CUDA(cudaFuncSetAttribute(attention_mma_bwd_kernel<NKB, DH>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_mma_bwd_kernel<NKB, DH><<<B * H, 128, smem, st>>>(
      (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, lse, ids, S, H, (__nv_bfloat16*)dqkv);
  return 0;
}

bool attention_mma_supported(int S, int dh) { return S <= LSP && (dh == 32 || dh == 64); }

// tcgen05 / TMEM / TMA kernels (attention_umma.cu): head depth 32, S <= 128
bool attention_umma_supported(int S, int H, int dh);
int attention_umma_fwd(const void* qkv, const int32_t* ids, int B, int S, int H, void* out,
                       float* lse, cudaStream_t st);

int attention_mma_fwd(const void* qkv, const int32_t* ids, int B, int S, int H, int dh, void* out,
                      float* lse, cudaStream_t st) {
  // (TMA needs 16-byte aligned base addresses; an offset view falls back to the kernels below)
  if (attention_umma_supported(S, H, dh) && (((uintptr_t)qkv | (uintptr_t)out) & 15) == 0)
    return attention_umma_fwd(qkv, ids, B, S, H, out, lse, st);
  if (S > 128)
    return dh == 32 ? launch_fwd_long<32>(qkv, ids, B, S, H, out, lse, st)
                    : launch_fwd_long<64>(qkv, ids, B, S, H, out, lse, st);
  const int nkb = S <= 64 ? 1 : 2;
  if (nkb == 1 && dh == 32) return launch_fwd<1, 32>(qkv, ids, B, S, H, out, lse, st);
  if (nkb == 1 && dh == 64) return launch_fwd<1, 64>(qkv, ids, B, S, H, out, lse, st);
  if (nkb == 2 && dh == 32) return launch_fwd<2, 32>(qkv, ids, B, S, H, out, lse, st);
  return launch_fwd<2, 64>(qkv, ids, B, S, H, out, lse, st);
}

bool attention_umma_bwd_supported(int S, int H, int dh);
int attention_umma_bwd(const void* qkv, const void* dout, const float* lse, const int32_t* ids,
                       int B, int S, int H, void* dqkv, cudaStream_t st);

int attention_mma_bwd(const void* qkv, const void* fwd_out, const void* dout, const float* lse,
                      const int32_t* ids, int B, int S, int H, int dh, void* dqkv, cudaStream_t st) {
  if (attention_umma_bwd_supported(S, H, dh) && (((uintptr_t)qkv | (uintptr_t)dout | (uintptr_t)dqkv) & 15) == 0)
    return attention_umma_bwd(qkv, dout, lse, ids, B, S, H, dqkv, st);
  if (S > 128) {
    if (!fwd_out) {
      set_last_error("attention_bwd: S=%d > 128 needs the forward output (delta = rowsum(dO o O))", S);
      return -1;
    }
    return dh == 32 ? launch_bwd_long<32>(qkv, fwd_out, dout, lse, ids, B, S, H, dqkv, st)
                    : launch_bwd_long<64>(qkv, fwd_out, dout, lse, ids, B, S, H, dqkv, st);
  }
  const int nkb = S <= 64 ? 1 : 2;
  if (nkb == 1 && dh == 32) return launch_bwd<1, 32>(qkv, dout, lse, ids, B, S, H, dqkv, st);
  if (nkb == 1 && dh == 64) return launch_bwd<1, 64>(qkv, dout, lse, ids, B, S, H, dqkv, st);
  if (nkb == 2 && dh == 32) return launch_bwd<2, 32>(qkv, dout, lse, ids, B, S, H, dqkv, st);
  return launch_bwd<2, 64>(qkv, dout, lse, ids, B, S, H, dqkv, st);
}

}  // namespace b4cp
