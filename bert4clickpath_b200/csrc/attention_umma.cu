// tcgen05 / TMEM / TMA path of the fused short-sequence masked self-attention, head depth 32:
// clickstream_transformer/transformer.py:64-97 (scaled_dot_product_attention with the additive
// -1e9 key padding mask of create_padding_mask :38-41) and :130-156 (split / merge heads).
//
// The per-(sequence, head) products of the reference's configurations are 52x52x32 - far below a
// 128-row UMMA tile - so a work ITEM packs them: for S <= 64, two sequences (64 TMEM lanes each)
// and the two heads that share a 64-column (128-byte) slice of the fused (q | k | v) rows; for
// 64 < S <= 128, one sequence and the same two heads.  The item's Q, K and V tiles (128 rows x 128
// bytes each, 128-byte swizzle) arrive with one 3-D TMA box per operand straight from the
// [B][S][3d] activation - rows past S and sequences past B are zero-filled by the TMA unit - into
// a 2-stage ring.
//
//   S_h  = Q_h K_h^T        tcgen05.mma, A and B K-major from shared memory (K = 32: two steps at
//                           byte offset 64 h inside the swizzled rows), 128 x 128 fp32 in TMEM.
//                           With two sequences per tile only the two diagonal 64 x 64 blocks are
//                           meaningful; the off-diagonal blocks are never read.
//   P_h  = exp2(S_h c + mask - max)   one thread per (query row, head) reads its row from TMEM
//                           (tcgen05.ld), and writes the un-normalised bf16 probabilities back to
//                           TMEM IN PLACE of the scores (tcgen05.st), zeros in the off-diagonal
//                           block: P never touches shared memory.
//   O_h  = P_h V            tcgen05.mma, A from TMEM (TS form), B = the V tile read MN-major.
//   out  = O_h / rowsum     read back with tcgen05.ld, scaled, packed to bf16, 64-byte row stores.
//
// Warp roles (320 threads, 2 CTAs per SM so that one CTA's softmax overlaps the other's tensor and
// TMA phases): warps 0-7 epilogue (warp w: TMEM lanes 32 (w % 4), head slot w / 4), warp 8 TMA
// producer + TMEM owner, warp 9 MMA issuer.  TMEM: 256 columns per CTA, per head [P (64 packed
// columns) | O (64 columns)] over the 128 columns S occupied.
//
// The backward kernel (same items, Q / K / V / dO tiles) computes both orientations on the tensor
// cores - S, dP = dO V^T with queries on lanes for dQ = dZ K, and S^T, dP^T = V dO^T with keys on
// lanes for dV = P^T dO and dK = dZ^T Q - so no transposition and no atomics are needed.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "../../include/b4cp.h"
#include "common.cuh"

namespace b4cp {

static constexpr int AU_THREADS = 320;
static constexpr int AU_WARP_TMA = 8, AU_WARP_MMA = 9;
static constexpr int AU_EPI_WARPS = 8;
static constexpr int AU_NST = 2;                      // stages of the operand ring
static constexpr int AU_TILE = 128 * 128;             // bytes: 128 rows x 64 bf16
static constexpr float AU_LOG2E = 1.4426950408889634f;
static constexpr float AU_LN2 = 0.6931471805599453f;
static constexpr float AU_PAD2 = -1.0e9f * 1.4426950408889634f;   // the -1e9 mask in log2 units

__device__ __forceinline__ float au_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t au_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void tma_load_3d_el(uint32_t smem_dst, const CUtensorMap* m,
                                               uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n\t}" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 32 lanes x 32 columns <- one register (zero fill without 32 live registers)
__device__ __forceinline__ void tmem_st32_same(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(v)
      : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// the 256 threads of the 8 epilogue warps
__device__ __forceinline__ void au_epi_sync() {
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

struct AttnUmmaParams {
  const int32_t* ids;     // [B][S] ids of the first feature (0 = pad)
  __nv_bfloat16* out;     // forward: [B*S][d]
  float* lse;             // [B][H][S]
  __nv_bfloat16* dqkv;    // backward: [B*S][3d]
  int B, S, H, d;
  int n_boxes;            // d / 64
  int n_items;            // ceil(B / SEQS) * n_boxes
  float scale2;           // log2(e) / sqrt(dh)
  float scale;            // 1 / sqrt(dh)
  int spin;               // experiments: epilogue warps poll their barriers without suspending
};

__device__ __forceinline__ void au_wait(uint64_t* bar, uint32_t parity, int spin) {
  if (spin) {
    while (!mbar_test(bar, parity)) {
    }
  } else {
    mbar_wait(bar, parity);
  }
}

// mask value (log2 units) of key j of a row whose 32-key group has validity bits `len` and pad
// bits `pad`
__device__ __forceinline__ float au_mask(uint32_t len, uint32_t pad, int j) {
  return ((len >> j) & 1u) ? (((pad >> j) & 1u) ? AU_PAD2 : 0.f) : -INFINITY;
}

// ================================================================================= forward
// SEQS = 2: S <= 64, lanes [0,64) hold sequence 2p, lanes [64,128) sequence 2p+1; a row's keys are
//           the 64 columns of its own diagonal block.
// SEQS = 1: 64 < S <= 128, all 128 columns are the row's keys.
template <int SEQS>
__global__ void __launch_bounds__(AU_THREADS, 2)
attention_umma_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AU_NST * 3 * AU_TILE);
  uint64_t* full = bars;             // [2] TMA -> MMA
  uint64_t* empty = bars + 2;        // [2] MMA (P V retired) -> TMA
  uint64_t* s_full = bars + 4;       // S of both heads in TMEM
  uint64_t* p_full = bars + 5;       // 8 warps: P written
  uint64_t* o_full = bars + 6;       // O of both heads in TMEM
  uint64_t* o_read = bars + 7;       // 8 warps: O read, TMEM free for the next item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  // per stage: keep bits of the item's keys [4 x 32] (SEQS = 2: sequence slot s, group g at 2s+g)
  // and, per sequence slot, whether every key of the sequence is a pad
  uint32_t* masks = reinterpret_cast<uint32_t*>(bars + 10);   // [AU_NST][8]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_mine = ((int)blockIdx.x < p.n_items)
                         ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x
                         : 0;

  if (warp == AU_WARP_TMA) {
    if (lane == 0) tma_prefetch_desc(&tmQKV);
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  } else if (warp == AU_WARP_MMA && lane == 0) {
    for (int s = 0; s < AU_NST; ++s) {
      mbar_init(&full[s], 2);   // the TMA bytes (expect_tx arrival) + the key masks
      mbar_init(&empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, AU_EPI_WARPS);
    mbar_init(o_full, 1);
    mbar_init(o_read, AU_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == AU_WARP_TMA) {
    const uint32_t a0 = smem_u32(smem);
    for (int i = 0; i < n_mine; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int grp = item / p.n_boxes, box = item - grp * p.n_boxes;
      const int st = i % AU_NST;
      mbar_wait_all(&empty[st], (uint32_t)((i / AU_NST) & 1) ^ 1);
      mbar_expect_tx_el(&full[st], 3u * AU_TILE);
      const uint32_t dst = a0 + (uint32_t)(st * 3 * AU_TILE);
#pragma unroll
      for (int o = 0; o < 3; ++o)
        tma_load_3d_el(dst + o * AU_TILE, &tmQKV, &full[st], o * p.d + box * 64, 0, grp * SEQS);
      // Key masks of the item, while its tiles are in flight.  keep: keys that take part in the
      // softmax.  A pad key's additive -1e9 (create_padding_mask) gives it exactly zero
      // probability next to any unpadded key, so it is simply left out; if EVERY key of the
      // sequence is a pad, fp32 absorbs the scores into the -1e9 and the reference's softmax is
      // uniform over the S keys: keep = all S keys, flagged so that the scores are ignored.
      uint32_t keep[4], len[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int b = SEQS == 2 ? grp * 2 + (g >> 1) : grp;
        const int j = (SEQS == 2 ? (g & 1) : g) * 32 + lane;
        const bool in = b < p.B && j < p.S;
        const int id = in ? __ldg(p.ids + (size_t)b * p.S + j) : 1;
        len[g] = __ballot_sync(0xffffffffu, in);
        keep[g] = __ballot_sync(0xffffffffu, in && id != 0);
      }
      uint32_t allpad[2];
      if (SEQS == 2) {
        allpad[0] = (keep[0] | keep[1]) == 0u;
        allpad[1] = (keep[2] | keep[3]) == 0u;
      } else {
        allpad[0] = allpad[1] = (keep[0] | keep[1] | keep[2] | keep[3]) == 0u;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (allpad[SEQS == 2 ? (g >> 1) : 0]) keep[g] = len[g];
      // a sequence past B: one kept key keeps the arithmetic of its (never stored) rows finite
      if (SEQS == 2) {
        if ((keep[0] | keep[1]) == 0u) keep[0] = 1u;
        if ((keep[2] | keep[3]) == 0u) keep[2] = 1u;
      } else if ((keep[0] | keep[1] | keep[2] | keep[3]) == 0u) {
        keep[0] = 1u;
      }
      if (lane < 4) masks[st * 8 + lane] = lane == 0 ? keep[0] : lane == 1 ? keep[1] : lane == 2 ? keep[2] : keep[3];
      if (lane >= 4 && lane < 6) masks[st * 8 + lane] = allpad[lane - 4];
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    }
  } else if (warp == AU_WARP_MMA) {
    const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t id_o = umma_idesc_bf16(128, 64, 0, 1);
    const uint32_t a0 = smem_u32(smem);
    const uint64_t dK0 = umma_smem_desc(a0, 16, 1024);       // K-major tiles (Q, K)
    const uint64_t dV0 = umma_smem_desc(a0, 8192, 1024);     // MN-major tile (V)
    for (int i = 0; i < n_mine; ++i) {
      const int st = i % AU_NST;
      const uint32_t so = (uint32_t)(st * 3 * AU_TILE);
      mbar_wait_all(&full[st], (uint32_t)((i / AU_NST) & 1));
      if (i > 0) mbar_wait_all(o_read, (uint32_t)((i - 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int hs = 0; hs < 2; ++hs)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16_el(tmem_base + hs * 128, dK0 + ((so + (uint32_t)(hs * 64 + kk * 32)) >> 4),
                       dK0 + ((so + (uint32_t)(AU_TILE + hs * 64 + kk * 32)) >> 4), id_s,
                       kk ? 1u : 0u);
      umma_commit_el(s_full);
      mbar_wait_all(p_full, (uint32_t)(i & 1));
      tc_fence_after();
#pragma unroll
      for (int hs = 0; hs < 2; ++hs)
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8)
          umma_bf16_ts_el(tmem_base + hs * 128 + 64, tmem_base + hs * 128 + k8 * 8,
                          dV0 + ((so + (uint32_t)(2 * AU_TILE + k8 * 2048)) >> 4), id_o,
                          k8 ? 1u : 0u);
      umma_commit_el(o_full);
      umma_commit_el(&empty[st]);
    }
  } else {
    // ------------------------------------------------------------ epilogue: thread = (row, head)
    const int q = warp & 3, hs = warp >> 2;
    const int row = 32 * q + lane;
    const int sq = SEQS == 2 ? (q >> 1) : 0;            // sequence slot of this row
    const int si = SEQS == 2 ? (row & 63) : row;        // query position
    const uint32_t t_head = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(hs * 128);
    constexpr int NG = SEQS == 2 ? 2 : 4;               // 32-key groups of a row
    const uint32_t t_keys = t_head + (uint32_t)(SEQS == 2 ? sq * 64 : 0);   // the row's scores
    const uint32_t t_p = t_head + (uint32_t)(SEQS == 2 ? sq * 32 : 0);      // its packed P
    for (int i = 0; i < n_mine; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int grp = item / p.n_boxes, box = item - grp * p.n_boxes;
      const int b = grp * SEQS + sq;
      const int head = box * 2 + hs;
      const bool seq_ok = b < p.B;
      // the producer's key masks of this stage (its arrival on full[st] released them)
      const int st = i % AU_NST;
      au_wait(&full[st], (uint32_t)((i / AU_NST) & 1), p.spin);
      uint32_t keep[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) keep[g] = masks[st * 8 + (SEQS == 2 ? sq * 2 : 0) + g];
      const bool allpad = masks[st * 8 + 4 + sq] != 0u;
      const float sc = allpad ? 0.f : p.scale2;
      const float bias = allpad ? AU_PAD2 : 0.f;
      au_wait(s_full, (uint32_t)(i & 1), p.spin);
      tc_fence_after();
      float sum;
      float m2;
      if constexpr (SEQS == 2) {
        // the row's 64 scores stay in registers: one TMEM read
        uint32_t ra[32], rb[32];
        tmem_ld32(t_keys, ra);
        tmem_ld32(t_keys + 32, rb);
        tmem_ld_wait();
        float m = -INFINITY;
        if (keep[0] == 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            m = fmaxf(m, ((keep[0] >> j) & 1u) ? __uint_as_float(ra[j]) : -INFINITY);
        }
        if (keep[1] == 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            m = fmaxf(m, ((keep[1] >> j) & 1u) ? __uint_as_float(rb[j]) : -INFINITY);
        }
        m2 = fmaf(m, sc, bias);
        const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(bias - m2, bias - m2);
        float2 acc = make_float2(0.f, 0.f);
        // 8 packed columns (16 keys) per store: the packed values never need more than 8
        // registers next to the 64 scores
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[8];
          const bool lo = c < 2;
          const uint32_t kp = lo ? keep[0] : keep[1];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = (c & 1) * 16 + 2 * u;
            const float2 z = __ffma2_rn(lo ? make_float2(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1]))
                                           : make_float2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1])),
                                        sc2, nb2);
            float e0 = au_ex2(z.x), e1 = au_ex2(z.y);
            if (kp != 0xffffffffu) {
              e0 = ((kp >> j) & 1u) ? e0 : 0.f;
              e1 = ((kp >> (j + 1)) & 1u) ? e1 : 0.f;
            }
            acc = __fadd2_rn(acc, make_float2(e0, e1));
            pk[u] = au_pack(e0, e1);
          }
          tmem_st8(t_p + c * 8, pk);
        }
        sum = acc.x + acc.y;
        tmem_st32_same(t_head + (uint32_t)((1 - sq) * 32), 0u);   // the other sequence's keys
      } else {
        // 128 scores per row: two sweeps over TMEM (maximum, then probabilities)
        float m = -INFINITY;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          uint32_t r[32];
          tmem_ld32(t_keys + g * 32, r);
          tmem_ld_wait();
          if (keep[g] == 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              m = fmaxf(m, ((keep[g] >> j) & 1u) ? __uint_as_float(r[j]) : -INFINITY);
          }
        }
        m2 = fmaf(m, sc, bias);
        const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(bias - m2, bias - m2);
        float2 acc = make_float2(0.f, 0.f);
        // In-place is safe inside a thread: group g is read from columns [32g, 32g+32) before its
        // 16 packed columns [16g, 16g+16) are written, and those lie below every group still to
        // be read.
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          uint32_t r[32], pk[16];
          tmem_ld32(t_keys + g * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 z = __ffma2_rn(make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sc2, nb2);
            float e0 = au_ex2(z.x), e1 = au_ex2(z.y);
            if (keep[g] != 0xffffffffu) {
              e0 = ((keep[g] >> j) & 1u) ? e0 : 0.f;
              e1 = ((keep[g] >> (j + 1)) & 1u) ? e1 : 0.f;
            }
            acc = __fadd2_rn(acc, make_float2(e0, e1));
            pk[j >> 1] = au_pack(e0, e1);
          }
          tmem_st16(t_p + g * 16, pk);
        }
        sum = acc.x + acc.y;
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_warp(p_full);
      const bool row_ok = seq_ok && si < p.S;
      if (row_ok && p.lse)
        p.lse[((size_t)b * p.H + head) * p.S + si] = (m2 + __log2f(sum)) * AU_LN2;
      const float inv = 1.f / sum;
      const float2 inv2 = make_float2(inv, inv);
      au_wait(o_full, (uint32_t)(i & 1), p.spin);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(t_head + 64 + hs * 32, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(o_read);
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.S + si) * p.d + box * 64 + hs * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[u] = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 2 * u]), __uint_as_float(o[8 * c + 2 * u + 1])), inv2);
          dst[c] = make_uint4(au_pack(v[0].x, v[0].y), au_pack(v[1].x, v[1].y),
                              au_pack(v[2].x, v[2].y), au_pack(v[3].x, v[3].y));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AU_WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ================================================================================= backward
// S <= 64: an item is the forward's - two sequences x the two heads of a 64-column box - with the
// Q, K, V and dO tiles in a 2-stage ring (64 KB per stage).  One CTA per SM, all 512 TMEM columns:
//
//   phase 1 (tensor):  S_h = Q_h K_h^T and dP_h = dO_h V_h^T for both heads, lanes = (sequence,
//                      query), four 128-column accumulators.
//   phase 2 (threads): thread (row, head) reads its 64 scores and 64 dP values, forms
//                      P = exp2(S c - lse), delta = sum_j P dP, dZ = P (dP - delta) / sqrt(dh) and
//                      writes the bf16 rows of P and dZ into shared memory, 128-byte swizzled, as
//                      per-sequence tiles [head][query][key] (16 KB each).
//   phase 3 (tensor):  the SAME bytes are read through two descriptors (as the vocabulary kernels
//                      read a W tile both ways): K-major, rows = (head, query), for
//                      dQ_s = dZ_s K_s; MN-major, rows = (head, key), for dK_s = dZ_s^T Q_s and
//                      dV_s = P_s^T dO_s - the transposed products need no transposition, no
//                      second orientation of the scores and no atomics.  B operands are the stage
//                      tiles of sequence s read MN-major over all 64 columns (a head's lanes use
//                      its own 32).  Six 64-column accumulators reuse the score columns.
//   phase 4 (threads): accumulators -> bf16 -> dqkv (64-byte row stores).
static constexpr int AB_STAGE = 4 * AU_TILE;          // Q, K, V, dO
static constexpr int AB_PZ = 4 * AU_TILE;             // P[2 sequences], dZ[2 sequences]

__global__ void __launch_bounds__(AU_THREADS, 1)
attention_umma_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV,
                          const __grid_constant__ CUtensorMap tmDO,
                          const __grid_constant__ CUtensorMap tmDQKV, const AttnUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sPZ = smem + AU_NST * AB_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPZ + AB_PZ);
  uint64_t* full = bars;             // [2] TMA bytes + key masks
  uint64_t* empty = bars + 2;        // [2] phase 3 retired
  uint64_t* s_full = bars + 4;       // phase 1 retired
  uint64_t* pz_full = bars + 5;      // 8 warps: P / dZ in shared memory, scores consumed
  uint64_t* acc_full = bars + 6;     // phase 3 retired
  uint64_t* acc_read = bars + 7;     // 8 warps: accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* masks = reinterpret_cast<uint32_t*>(bars + 10);   // [AU_NST][8]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_mine = ((int)blockIdx.x < p.n_items)
                         ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x
                         : 0;

  if (warp == AU_WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      tma_prefetch_desc(&tmDQKV);
    }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  } else if (warp == AU_WARP_MMA && lane == 0) {
    for (int s = 0; s < AU_NST; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pz_full, AU_EPI_WARPS);
    mbar_init(acc_full, 1);
    mbar_init(acc_read, AU_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == AU_WARP_TMA) {
    const uint32_t a0 = smem_u32(smem);
    for (int i = 0; i < n_mine; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int grp = item / p.n_boxes, box = item - grp * p.n_boxes;
      const int st = i % AU_NST;
      mbar_wait_all(&empty[st], (uint32_t)((i / AU_NST) & 1) ^ 1);
      mbar_expect_tx_el(&full[st], 4u * AU_TILE);
      const uint32_t dst = a0 + (uint32_t)(st * AB_STAGE);
#pragma unroll
      for (int o = 0; o < 3; ++o)
        tma_load_3d_el(dst + o * AU_TILE, &tmQKV, &full[st], o * p.d + box * 64, 0, grp * 2);
      tma_load_3d_el(dst + 3 * AU_TILE, &tmDO, &full[st], box * 64, 0, grp * 2);
      // key masks (see the forward kernel): [2 s + g] keep bits, [4 + s] all-pad flag
      uint32_t keep[4], len[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int b = grp * 2 + (g >> 1);
        const int j = (g & 1) * 32 + lane;
        const bool in = b < p.B && j < p.S;
        const int id = in ? __ldg(p.ids + (size_t)b * p.S + j) : 1;
        len[g] = __ballot_sync(0xffffffffu, in);
        keep[g] = __ballot_sync(0xffffffffu, in && id != 0);
      }
      uint32_t allpad[2];
      allpad[0] = (keep[0] | keep[1]) == 0u;
      allpad[1] = (keep[2] | keep[3]) == 0u;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (allpad[g >> 1]) keep[g] = len[g];
      if (lane < 4) masks[st * 8 + lane] = lane == 0 ? keep[0] : lane == 1 ? keep[1] : lane == 2 ? keep[2] : keep[3];
      if (lane >= 4 && lane < 6) masks[st * 8 + lane] = allpad[lane - 4];
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    }
  } else if (warp == AU_WARP_MMA) {
    const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t id_q = umma_idesc_bf16(128, 64, 0, 1);    // dQ: A K-major, B MN-major
    const uint32_t id_t = umma_idesc_bf16(128, 64, 1, 1);    // dK, dV: A MN-major, B MN-major
    const uint32_t a0 = smem_u32(smem);
    const uint32_t aPZ = smem_u32(sPZ);
    const uint64_t dKm = umma_smem_desc(a0, 16, 1024);        // K-major view of a stage tile
    const uint64_t dMn = umma_smem_desc(a0, 8192, 1024);      // MN-major view of a stage tile
    const uint64_t dPZk = umma_smem_desc(aPZ, 16, 1024);      // P / dZ tiles, rows = (head, query)
    const uint64_t dPZm = umma_smem_desc(aPZ, 8192, 1024);    // P / dZ tiles, rows = (head, key)
    for (int i = 0; i < n_mine; ++i) {
      const int st = i % AU_NST;
      const uint32_t so = (uint32_t)(st * AB_STAGE);
      mbar_wait_all(&full[st], (uint32_t)((i / AU_NST) & 1));
      if (i > 0) mbar_wait_all(acc_read, (uint32_t)((i - 1) & 1));
      tc_fence_after();
      // ---- phase 1
#pragma unroll
      for (int hs = 0; hs < 2; ++hs)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t ko = (uint32_t)(hs * 64 + kk * 32);
          umma_bf16_el(tmem_base + hs * 128, dKm + ((so + ko) >> 4),
                       dKm + ((so + AU_TILE + ko) >> 4), id_s, kk ? 1u : 0u);
          umma_bf16_el(tmem_base + 256 + hs * 128, dKm + ((so + 3 * AU_TILE + ko) >> 4),
                       dKm + ((so + 2 * AU_TILE + ko) >> 4), id_s, kk ? 1u : 0u);
        }
      umma_commit_el(s_full);
      mbar_wait_all(pz_full, (uint32_t)(i & 1));
      tc_fence_after();
      // ---- phase 3 (P[s] at aPZ + s * 16 KB, dZ[s] at aPZ + 32 KB + s * 16 KB)
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        const uint32_t rows = (uint32_t)(sq * 8192);              // sequence sq's 64 rows of a tile
        const uint32_t pP = (uint32_t)(sq * AU_TILE), pZ = (uint32_t)(2 * AU_TILE + sq * AU_TILE);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          // dQ_s = dZ_s K_s           (K = keys)
          umma_bf16_el(tmem_base + sq * 64, dPZk + ((pZ + (uint32_t)(kk * 32)) >> 4),
                       dMn + ((so + AU_TILE + rows + (uint32_t)(kk * 2048)) >> 4), id_q, kk ? 1u : 0u);
          // dK_s = dZ_s^T Q_s         (K = queries)
          umma_bf16_el(tmem_base + 128 + sq * 64, dPZm + ((pZ + (uint32_t)(kk * 2048)) >> 4),
                       dMn + ((so + rows + (uint32_t)(kk * 2048)) >> 4), id_t, kk ? 1u : 0u);
          // dV_s = P_s^T dO_s         (K = queries)
          umma_bf16_el(tmem_base + 256 + sq * 64, dPZm + ((pP + (uint32_t)(kk * 2048)) >> 4),
                       dMn + ((so + 3 * AU_TILE + rows + (uint32_t)(kk * 2048)) >> 4), id_t, kk ? 1u : 0u);
        }
      }
      umma_commit_el(acc_full);
      umma_commit_el(&empty[st]);
    }
  } else {
    // ------------------------------------------------------------ threads: (row, head slot)
    const int q = warp & 3, hs = warp >> 2;
    const int row = 32 * q + lane;
    const int sq = q >> 1;              // phase 2: sequence slot of this row
    const int qi = row & 63;            //          query position
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t aPZ = smem_u32(sPZ);
    // phase 4: TMEM lane = (head q >> 1, position (q & 1) * 32 + lane); warp group hs drains the
    // accumulators of sequence hs
    const int h4 = q >> 1, pos4 = (q & 1) * 32 + lane;
    // -lse (log2 units) of this thread's row, fetched one item ahead; -inf for rows that do not
    // exist -> P = 0
    auto fetch_nb = [&](int i) -> float {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int grp = item / p.n_boxes, box = item - grp * p.n_boxes;
      const int b = grp * 2 + sq;
      return (b < p.B && qi < p.S)
                 ? -__ldg(p.lse + ((size_t)b * p.H + box * 2 + hs) * p.S + qi) * AU_LOG2E
                 : -INFINITY;
    };
    float nb_next = n_mine > 0 ? fetch_nb(0) : 0.f;
    for (int i = 0; i < n_mine; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int grp = item / p.n_boxes, box = item - grp * p.n_boxes;
      const int st = i % AU_NST;
      float nb = nb_next;
      au_wait(&full[st], (uint32_t)((i / AU_NST) & 1), p.spin);
      const uint32_t keep0 = masks[st * 8 + sq * 2], keep1 = masks[st * 8 + sq * 2 + 1];
      float sc = p.scale2;
      if (masks[st * 8 + 4 + sq] != 0u) {   // every key a pad: uniform over the S keys
        sc = 0.f;
        if (nb != -INFINITY) nb = -__log2f((float)(__popc(keep0) + __popc(keep1)));
      }
      au_wait(s_full, (uint32_t)(i & 1), p.spin);
      tc_fence_after();
      uint32_t ra[32], rb[32], da[32], db[32];
      tmem_ld32(t_lane + hs * 128 + sq * 64, ra);
      tmem_ld32(t_lane + hs * 128 + sq * 64 + 32, rb);
      tmem_ld32(t_lane + 256 + hs * 128 + sq * 64, da);
      tmem_ld32(t_lane + 256 + hs * 128 + sq * 64 + 32, db);
      // the previous item's gradient tiles leave through TMA stores out of the P / dZ region:
      // their reads of shared memory must be over before it is written again
      if (i > 0) {
        if (threadIdx.x == 0) tma_store_wait_read();
        au_epi_sync();
      }
      tmem_ld_wait();
      // P (kept in place of the scores) and delta = sum_j P_j dP_j
      const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(nb, nb);
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        float2 z = __ffma2_rn(make_float2(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1])), sc2, nb2);
        float e0 = au_ex2(z.x), e1 = au_ex2(z.y);
        if (keep0 != 0xffffffffu) {
          e0 = ((keep0 >> j) & 1u) ? e0 : 0.f;
          e1 = ((keep0 >> (j + 1)) & 1u) ? e1 : 0.f;
        }
        ra[j] = __float_as_uint(e0);
        ra[j + 1] = __float_as_uint(e1);
        acc = __ffma2_rn(make_float2(e0, e1), make_float2(__uint_as_float(da[j]), __uint_as_float(da[j + 1])), acc);
        z = __ffma2_rn(make_float2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1])), sc2, nb2);
        e0 = au_ex2(z.x), e1 = au_ex2(z.y);
        if (keep1 != 0xffffffffu) {
          e0 = ((keep1 >> j) & 1u) ? e0 : 0.f;
          e1 = ((keep1 >> (j + 1)) & 1u) ? e1 : 0.f;
        }
        rb[j] = __float_as_uint(e0);
        rb[j + 1] = __float_as_uint(e1);
        acc = __ffma2_rn(make_float2(e0, e1), make_float2(__uint_as_float(db[j]), __uint_as_float(db[j + 1])), acc);
      }
      const float delta = acc.x + acc.y;
      // rows of P and dZ = P (dP - delta) / sqrt(dh) -> shared memory, 128-byte swizzle:
      // tile [head][query][key], 16-byte chunk c of row r at chunk position c ^ (r & 7)
      const uint32_t rowP = aPZ + (uint32_t)(sq * AU_TILE + hs * 8192 + (qi >> 3) * 1024 + (qi & 7) * 128);
      const uint32_t rowZ = rowP + 2 * AU_TILE;
      const float2 nd2 = make_float2(-delta, -delta), s2 = make_float2(p.scale, p.scale);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pw[4], zw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = (c & 3) * 8 + 2 * u;
          const float2 pv = c < 4 ? make_float2(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1]))
                                  : make_float2(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1]));
          const float2 dv = c < 4 ? make_float2(__uint_as_float(da[j]), __uint_as_float(da[j + 1]))
                                  : make_float2(__uint_as_float(db[j]), __uint_as_float(db[j + 1]));
          const float2 dz = __fmul2_rn(__fmul2_rn(pv, __fadd2_rn(dv, nd2)), s2);
          pw[u] = au_pack(pv.x, pv.y);
          zw[u] = au_pack(dz.x, dz.y);
        }
        const uint32_t off = (uint32_t)((c ^ (qi & 7)) << 4);
        sts128(rowP + off, pw[0], pw[1], pw[2], pw[3]);
        sts128(rowZ + off, zw[0], zw[1], zw[2], zw[3]);
      }
      fence_proxy_async_smem();     // generic-proxy stores -> visible to the tensor core's reads
      tc_fence_before();
      mbar_arrive_warp(pz_full);
      if (i + 1 < n_mine) nb_next = fetch_nb(i + 1);
      // ---- phase 4: accumulators -> bf16 tiles [sequence][position][64 columns] in the (now
      // free) P / dZ region -> three TMA stores (rows past S and sequences past B are clipped)
      au_wait(acc_full, (uint32_t)(i & 1), p.spin);
      tc_fence_after();
      uint32_t gq[32], gk[32], gv[32];
      tmem_ld32(t_lane + hs * 64 + h4 * 32, gq);
      tmem_ld32(t_lane + 128 + hs * 64 + h4 * 32, gk);
      tmem_ld32(t_lane + 256 + hs * 64 + h4 * 32, gv);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(acc_read);
      const uint32_t rowO = aPZ + (uint32_t)(hs * 8192 + (pos4 >> 3) * 1024 + (pos4 & 7) * 128);
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const uint32_t(&g)[32] = o == 0 ? gq : (o == 1 ? gk : gv);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(rowO + (uint32_t)(o * AU_TILE) + (uint32_t)(((h4 * 4 + c) ^ (pos4 & 7)) << 4),
                 au_pack(__uint_as_float(g[8 * c + 0]), __uint_as_float(g[8 * c + 1])),
                 au_pack(__uint_as_float(g[8 * c + 2]), __uint_as_float(g[8 * c + 3])),
                 au_pack(__uint_as_float(g[8 * c + 4]), __uint_as_float(g[8 * c + 5])),
                 au_pack(__uint_as_float(g[8 * c + 6]), __uint_as_float(g[8 * c + 7])));
      }
      fence_proxy_async_smem();
      au_epi_sync();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int o = 0; o < 3; ++o)
          tma_store_3d(&tmDQKV, aPZ + (uint32_t)(o * AU_TILE), o * p.d + box * 64, 0, grp * 2);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait_read();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AU_WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ backward, 64 < S <= 128
// One sequence per item; its two heads are processed one after the other ("sub-items" j = 2 i + h)
// because a 128-row sequence needs all 128 key columns of S_h and dP_h (256 TMEM columns per head)
// and 32 KB each for the P_h / dZ_h tiles.  TMEM: scores [0,256), accumulators dQ | dK | dV at
// [256,448) - separate, so phase 1 of sub-item j+1 runs under phase 4 of sub-item j.  The 8 thread
// warps all work on the same head: thread (query row, key half) owns 64 of the row's 128 keys;
// delta = sum_j P dP needs both halves, exchanged through shared memory.  P_h / dZ_h tiles are two
// 64-key blocks of [128 queries][64 keys] (16 KB each): K-major that is a 128 x 128 A operand for
// dQ_h = dZ_h K (the K steps 4..7 live in the second block), MN-major a 128-key x 128-query A
// operand (two 64-wide M blocks) for dK_h = dZ_h^T Q and dV_h = P_h^T dO.
__global__ void __launch_bounds__(AU_THREADS, 1)
attention_umma_bwd1_kernel(const __grid_constant__ CUtensorMap tmQKV,
                           const __grid_constant__ CUtensorMap tmDO, const AttnUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sPZ = smem + AU_NST * AB_STAGE;          // P_h: [0, 32 KB), dZ_h: [32 KB, 64 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPZ + AB_PZ);
  uint64_t* full = bars;             // [2] TMA bytes + key masks
  uint64_t* empty = bars + 2;        // [2] phase 3 of the item's second head retired
  uint64_t* s_full = bars + 4;       // phase 1 retired
  uint64_t* pz_full = bars + 5;      // 8 warps: P / dZ in shared memory, scores consumed
  uint64_t* acc_full = bars + 6;     // phase 3 retired
  uint64_t* acc_read = bars + 7;     // 8 warps: accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint32_t* masks = reinterpret_cast<uint32_t*>(bars + 10);   // [AU_NST][8]
  float* sDelta = reinterpret_cast<float*>(masks + AU_NST * 8);   // [2 halves][128 rows]

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_mine = ((int)blockIdx.x < p.n_items)
                         ? (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x
                         : 0;
  const int n_sub = 2 * n_mine;

  if (warp == AU_WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
    }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  } else if (warp == AU_WARP_MMA && lane == 0) {
    for (int s = 0; s < AU_NST; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(pz_full, AU_EPI_WARPS);
    mbar_init(acc_full, 1);
    mbar_init(acc_read, AU_EPI_WARPS);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == AU_WARP_TMA) {
    const uint32_t a0 = smem_u32(smem);
    for (int i = 0; i < n_mine; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = item / p.n_boxes, box = item - b * p.n_boxes;
      const int st = i % AU_NST;
      mbar_wait_all(&empty[st], (uint32_t)((i / AU_NST) & 1) ^ 1);
      mbar_expect_tx_el(&full[st], 4u * AU_TILE);
      const uint32_t dst = a0 + (uint32_t)(st * AB_STAGE);
#pragma unroll
      for (int o = 0; o < 3; ++o)
        tma_load_3d_el(dst + o * AU_TILE, &tmQKV, &full[st], o * p.d + box * 64, 0, b);
      tma_load_3d_el(dst + 3 * AU_TILE, &tmDO, &full[st], box * 64, 0, b);
      uint32_t keep[4], len[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int j = g * 32 + lane;
        const bool in = j < p.S;
        const int id = in ? __ldg(p.ids + (size_t)b * p.S + j) : 1;
        len[g] = __ballot_sync(0xffffffffu, in);
        keep[g] = __ballot_sync(0xffffffffu, in && id != 0);
      }
      const uint32_t allpad = (keep[0] | keep[1] | keep[2] | keep[3]) == 0u;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (allpad) keep[g] = len[g];
      if (lane < 4) masks[st * 8 + lane] = lane == 0 ? keep[0] : lane == 1 ? keep[1] : lane == 2 ? keep[2] : keep[3];
      if (lane == 4) masks[st * 8 + 4] = allpad;
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    }
  } else if (warp == AU_WARP_MMA) {
    const uint32_t id_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t id_q = umma_idesc_bf16(128, 64, 0, 1);
    const uint32_t id_t = umma_idesc_bf16(128, 64, 1, 1);
    const uint32_t a0 = smem_u32(smem);
    const uint32_t aPZ = smem_u32(sPZ);
    const uint64_t dKm = umma_smem_desc(a0, 16, 1024);
    const uint64_t dMn = umma_smem_desc(a0, 8192, 1024);
    const uint64_t dPZk = umma_smem_desc(aPZ, 16, 1024);        // rows = queries, K = keys
    const uint64_t dPZm = umma_smem_desc(aPZ, 16384, 1024);     // rows = keys (2 x 64), K = queries
    auto phase1 = [&](int j) {
      const int i = j >> 1, hs = j & 1, st = i % AU_NST;
      const uint32_t so = (uint32_t)(st * AB_STAGE);
      if (hs == 0) mbar_wait_all(&full[st], (uint32_t)((i / AU_NST) & 1));
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t ko = (uint32_t)(hs * 64 + kk * 32);
        umma_bf16_el(tmem_base, dKm + ((so + ko) >> 4), dKm + ((so + AU_TILE + ko) >> 4), id_s,
                     kk ? 1u : 0u);
        umma_bf16_el(tmem_base + 128, dKm + ((so + 3 * AU_TILE + ko) >> 4),
                     dKm + ((so + 2 * AU_TILE + ko) >> 4), id_s, kk ? 1u : 0u);
      }
      umma_commit_el(s_full);
    };
    if (n_sub > 0) phase1(0);
    for (int j = 0; j < n_sub; ++j) {
      const int i = j >> 1, hs = j & 1, st = i % AU_NST;
      const uint32_t so = (uint32_t)(st * AB_STAGE);
      mbar_wait_all(pz_full, (uint32_t)(j & 1));
      if (j > 0) mbar_wait_all(acc_read, (uint32_t)((j - 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t kq = (uint32_t)((kk >> 2) * 16384 + (kk & 3) * 32);   // K-major A: key step
        const uint32_t rows = (uint32_t)(kk * 2048);                          // 16 rows of a tile
        // dQ_h = dZ_h K
        umma_bf16_el(tmem_base + 256, dPZk + ((2u * AU_TILE + kq) >> 4),
                     dMn + ((so + AU_TILE + rows) >> 4), id_q, kk ? 1u : 0u);
        // dK_h = dZ_h^T Q
        umma_bf16_el(tmem_base + 320, dPZm + ((2u * AU_TILE + rows) >> 4),
                     dMn + ((so + rows) >> 4), id_t, kk ? 1u : 0u);
        // dV_h = P_h^T dO
        umma_bf16_el(tmem_base + 384, dPZm + (rows >> 4),
                     dMn + ((so + 3 * AU_TILE + rows) >> 4), id_t, kk ? 1u : 0u);
      }
      umma_commit_el(acc_full);
      if (hs == 1) umma_commit_el(&empty[st]);
      if (j + 1 < n_sub) phase1(j + 1);   // (its s_full arrives after the products above retire)
    }
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int row = 32 * q + lane;                     // query row (phase 2) / TMEM lane (phase 4)
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t aPZ = smem_u32(sPZ);
    auto fetch_nb = [&](int j) -> float {
      const int item = (int)blockIdx.x + (j >> 1) * (int)gridDim.x;
      const int b = item / p.n_boxes, box = item - b * p.n_boxes;
      return row < p.S ? -__ldg(p.lse + ((size_t)b * p.H + box * 2 + (j & 1)) * p.S + row) * AU_LOG2E
                       : -INFINITY;
    };
    float nb_next = n_sub > 0 ? fetch_nb(0) : 0.f;
    for (int j = 0; j < n_sub; ++j) {
      const int i = j >> 1, hs = j & 1, st = i % AU_NST;
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = item / p.n_boxes, box = item - b * p.n_boxes;
      float nb = nb_next;
      au_wait(&full[st], (uint32_t)((i / AU_NST) & 1), p.spin);
      const uint32_t keep0 = masks[st * 8 + half * 2], keep1 = masks[st * 8 + half * 2 + 1];
      float sc = p.scale2;
      if (masks[st * 8 + 4] != 0u) {   // every key a pad: uniform over the S keys
        sc = 0.f;
        if (nb != -INFINITY) nb = -__log2f((float)p.S);
      }
      au_wait(s_full, (uint32_t)(j & 1), p.spin);
      tc_fence_after();
      uint32_t ra[32], rb[32], da[32], db[32];
      tmem_ld32(t_lane + half * 64, ra);
      tmem_ld32(t_lane + half * 64 + 32, rb);
      tmem_ld32(t_lane + 128 + half * 64, da);
      tmem_ld32(t_lane + 128 + half * 64 + 32, db);
      tmem_ld_wait();
      const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(nb, nb);
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int jj = 0; jj < 32; jj += 2) {
        float2 z = __ffma2_rn(make_float2(__uint_as_float(ra[jj]), __uint_as_float(ra[jj + 1])), sc2, nb2);
        float e0 = au_ex2(z.x), e1 = au_ex2(z.y);
        if (keep0 != 0xffffffffu) {
          e0 = ((keep0 >> jj) & 1u) ? e0 : 0.f;
          e1 = ((keep0 >> (jj + 1)) & 1u) ? e1 : 0.f;
        }
        ra[jj] = __float_as_uint(e0);
        ra[jj + 1] = __float_as_uint(e1);
        acc = __ffma2_rn(make_float2(e0, e1), make_float2(__uint_as_float(da[jj]), __uint_as_float(da[jj + 1])), acc);
        z = __ffma2_rn(make_float2(__uint_as_float(rb[jj]), __uint_as_float(rb[jj + 1])), sc2, nb2);
        e0 = au_ex2(z.x), e1 = au_ex2(z.y);
        if (keep1 != 0xffffffffu) {
          e0 = ((keep1 >> jj) & 1u) ? e0 : 0.f;
          e1 = ((keep1 >> (jj + 1)) & 1u) ? e1 : 0.f;
        }
        rb[jj] = __float_as_uint(e0);
        rb[jj + 1] = __float_as_uint(e1);
        acc = __ffma2_rn(make_float2(e0, e1), make_float2(__uint_as_float(db[jj]), __uint_as_float(db[jj + 1])), acc);
      }
      // delta over both key halves (the other half's thread sits in warp +-4)
      sDelta[half * 128 + row] = acc.x + acc.y;
      au_epi_sync();
      const float delta = sDelta[row] + sDelta[128 + row];
      const uint32_t rowP = aPZ + (uint32_t)(half * 16384 + (row >> 3) * 1024 + (row & 7) * 128);
      const uint32_t rowZ = rowP + 2 * AU_TILE;
      const float2 nd2 = make_float2(-delta, -delta), s2 = make_float2(p.scale, p.scale);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t pw[4], zw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = (c & 3) * 8 + 2 * u;
          const float2 pv = c < 4 ? make_float2(__uint_as_float(ra[jj]), __uint_as_float(ra[jj + 1]))
                                  : make_float2(__uint_as_float(rb[jj]), __uint_as_float(rb[jj + 1]));
          const float2 dv = c < 4 ? make_float2(__uint_as_float(da[jj]), __uint_as_float(da[jj + 1]))
                                  : make_float2(__uint_as_float(db[jj]), __uint_as_float(db[jj + 1]));
          const float2 dz = __fmul2_rn(__fmul2_rn(pv, __fadd2_rn(dv, nd2)), s2);
          pw[u] = au_pack(pv.x, pv.y);
          zw[u] = au_pack(dz.x, dz.y);
        }
        const uint32_t off = (uint32_t)((c ^ (row & 7)) << 4);
        sts128(rowP + off, pw[0], pw[1], pw[2], pw[3]);
        sts128(rowZ + off, zw[0], zw[1], zw[2], zw[3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive_warp(pz_full);
      if (j + 1 < n_sub) nb_next = fetch_nb(j + 1);
      // ---- phase 4: this head's 32 columns of the accumulators (lane = query for dQ, key for
      // dK / dV); warp half 0 drains dQ and dK, half 1 dV
      au_wait(acc_full, (uint32_t)(j & 1), p.spin);
      tc_fence_after();
      uint32_t g0[32], g1[32];
      if (half == 0) {
        tmem_ld32(t_lane + 256 + hs * 32, g0);
        tmem_ld32(t_lane + 320 + hs * 32, g1);
      } else {
        tmem_ld32(t_lane + 384 + hs * 32, g0);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(acc_read);
      if (row < p.S) {
        __nv_bfloat16* dst = p.dqkv + ((size_t)b * p.S + row) * 3 * p.d + box * 64 + hs * 32;
        auto put = [&](const uint32_t(&g)[32], int o) {
          uint4* d4 = reinterpret_cast<uint4*>(dst + o * p.d);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            d4[c] = make_uint4(au_pack(__uint_as_float(g[8 * c + 0]), __uint_as_float(g[8 * c + 1])),
                               au_pack(__uint_as_float(g[8 * c + 2]), __uint_as_float(g[8 * c + 3])),
                               au_pack(__uint_as_float(g[8 * c + 4]), __uint_as_float(g[8 * c + 5])),
                               au_pack(__uint_as_float(g[8 * c + 6]), __uint_as_float(g[8 * c + 7])));
        };
        if (half == 0) {
          put(g0, 0);
          put(g1, 1);
        } else {
          put(g0, 2);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == AU_WARP_TMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================= host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn au_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 [B][S][width] activation (row stride ld elements) as a 3-D tensor: box = 64 columns x
// `box_rows` positions x `box_seqs` sequences, 128-byte swizzle, zero fill out of bounds
static int make_tmap_seq3d(CUtensorMap* map, const void* base, int width, int ld, int S, int B,
                           int box_rows, int box_seqs) {
  EncodeTiledFn fn = au_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point not available");
    return -2;
  }
  cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)S, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)S * ld * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, (cuuint32_t)box_seqs};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("attention: cuTensorMapEncodeTiled failed (%d): base=%p width=%d ld=%d S=%d B=%d",
                   (int)r, base, width, ld, S, B);
    return -3;
  }
  return 0;
}

static int au_spin() {
  static const int v = getenv("B4CP_ATTN_SPIN") ? atoi(getenv("B4CP_ATTN_SPIN")) : 0;
  return v;
}

static int au_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

bool attention_umma_supported(int S, int H, int dh) {
  static const bool off = getenv("B4CP_ATTN_MMA_SYNC") != nullptr;
  return !off && dh == 32 && (H % 2) == 0 && S >= 1 && S <= 128;
}

int attention_umma_fwd(const void* qkv, const int32_t* ids, int B, int S, int H, void* out,
                       float* lse, cudaStream_t st) {
  const int d = H * 32;
  const int seqs = S <= 64 ? 2 : 1;
  CUtensorMap tm;
  int rc = make_tmap_seq3d(&tm, qkv, 3 * d, 3 * d, S, B, seqs == 2 ? 64 : 128, seqs);
  if (rc) return rc;
  AttnUmmaParams p{};
  p.ids = ids;
  p.out = (__nv_bfloat16*)out;
  p.lse = lse;
  p.B = B;
  p.S = S;
  p.H = H;
  p.d = d;
  p.n_boxes = d / 64;
  p.n_items = ((B + seqs - 1) / seqs) * p.n_boxes;
  p.scale = 1.f / sqrtf(32.f);
  p.scale2 = AU_LOG2E / sqrtf(32.f);
  p.spin = au_spin();
  const int smem = AU_NST * 3 * AU_TILE + 1024 + 256;
  static const int grid_cap = getenv("B4CP_ATTN_GRID") ? atoi(getenv("B4CP_ATTN_GRID")) : 0;   // experiments
  const int grid = std::min(p.n_items, grid_cap > 0 ? grid_cap : 2 * au_num_sms());
  if (seqs == 2) {
    B4CP_CUDA(cudaFuncSetAttribute(attention_umma_fwd_kernel<2>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attention_umma_fwd_kernel<2><<<grid, AU_THREADS, smem, st>>>(tm, p);
  } else {
    B4CP_CUDA(cudaFuncSetAttribute(attention_umma_fwd_kernel<1>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attention_umma_fwd_kernel<1><<<grid, AU_THREADS, smem, st>>>(tm, p);
  }
  return 0;
}

bool attention_umma_bwd_supported(int S, int H, int dh) {
  return attention_umma_supported(S, H, dh);
}

int attention_umma_bwd(const void* qkv, const void* dout, const float* lse, const int32_t* ids,
                       int B, int S, int H, void* dqkv, cudaStream_t st) {
  const int d = H * 32;
  CUtensorMap tmQ, tmD;
  if (S > 64) {   // one sequence per item, heads in turn
    int rc1 = make_tmap_seq3d(&tmQ, qkv, 3 * d, 3 * d, S, B, 128, 1);
    if (rc1) return rc1;
    rc1 = make_tmap_seq3d(&tmD, dout, d, d, S, B, 128, 1);
    if (rc1) return rc1;
    AttnUmmaParams p1{};
    p1.ids = ids;
    p1.lse = const_cast<float*>(lse);
    p1.dqkv = (__nv_bfloat16*)dqkv;
    p1.B = B;
    p1.S = S;
    p1.H = H;
    p1.d = d;
    p1.n_boxes = d / 64;
    p1.n_items = B * p1.n_boxes;
    p1.scale = 1.f / sqrtf(32.f);
    p1.scale2 = AU_LOG2E / sqrtf(32.f);
    p1.spin = au_spin();
    const int smem1 = AU_NST * AB_STAGE + AB_PZ + 1024 + 256 + 1024;
    const int grid1 = std::min(p1.n_items, au_num_sms());
    B4CP_CUDA(cudaFuncSetAttribute(attention_umma_bwd1_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    attention_umma_bwd1_kernel<<<grid1, AU_THREADS, smem1, st>>>(tmQ, tmD, p1);
    return 0;
  }
  int rc = make_tmap_seq3d(&tmQ, qkv, 3 * d, 3 * d, S, B, 64, 2);
  if (rc) return rc;
  rc = make_tmap_seq3d(&tmD, dout, d, d, S, B, 64, 2);
  if (rc) return rc;
  CUtensorMap tmG;
  rc = make_tmap_seq3d(&tmG, dqkv, 3 * d, 3 * d, S, B, 64, 2);
  if (rc) return rc;
  AttnUmmaParams p{};
  p.ids = ids;
  p.lse = const_cast<float*>(lse);
  p.dqkv = (__nv_bfloat16*)dqkv;
  p.B = B;
  p.S = S;
  p.H = H;
  p.d = d;
  p.n_boxes = d / 64;
  p.n_items = ((B + 1) / 2) * p.n_boxes;
  p.scale = 1.f / sqrtf(32.f);
  p.scale2 = AU_LOG2E / sqrtf(32.f);
  p.spin = au_spin();
  const int smem = AU_NST * AB_STAGE + AB_PZ + 1024 + 256;
  const int grid = std::min(p.n_items, au_num_sms());
  B4CP_CUDA(cudaFuncSetAttribute(attention_umma_bwd_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attention_umma_bwd_kernel<<<grid, AU_THREADS, smem, st>>>(tmQ, tmD, tmG, p);
  return 0;
}

}  // namespace b4cp
