// Input embedding of the clickstream transformer and its backward.
//
// Forward (transformer.py:376-398): per-feature Embedding gather, concat on the last axis,
// * sqrt(d_model), + PE[:S], then the encoder's input dropout (transformer.py:263) - one kernel,
// 128-bit row loads, one 128-bit store per 4 outputs.  The multiply and the add are rounded
// separately (__fmul_rn / __fadd_rn, no FMA contraction) so the result is bit-identical to the
// two TensorFlow ops.
//
// Backward (TF autodiff -> IndexedSlices -> UnsortedSegmentSum): deterministic.  Token ids are
// stably radix-sorted (key = id, value = token index); the sorted list is cut into fixed 64-token
// tiles, one warp per tile sums every run of equal ids in token order, and runs that cross tile
// boundaries are finished from per-tile partial rows in tile order, so results are
// bit-reproducible run to run.
#include <algorithm>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

struct EmbedParams {
  const int32_t* ids[B4CP_MAX_FEATURES];
  const float* tables[B4CP_MAX_FEATURES];
  int dims[B4CP_MAX_FEATURES];
  int offs[B4CP_MAX_FEATURES + 1];
  int rows[B4CP_MAX_FEATURES];
  int F;
  int d_model;
  int S;
  long T;  // tokens = B*S
  const float* pe;
  float scale;
  float inv_keep;
  uint32_t thresh24;
  uint64_t seed;
  uint32_t site;
};

template <int VEC>
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const EmbedParams p, float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16) {
  const uint64_t seed_v = resolve_seed(p.seed);
  const long per_tok = p.d_model / VEC;
  const long total = p.T * per_tok;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long t = i / per_tok;
    const int c = (int)(i - t * per_tok) * VEC;
    int f = 0;
#pragma unroll
    for (int g = 1; g < B4CP_MAX_FEATURES; ++g)
      if (g < p.F && c >= p.offs[g]) f = g;
    const int id = __ldg(p.ids[f] + t);
    const int s = (int)(t % p.S);
    float e[VEC], pe[VEC], o[VEC];
    const bool id_ok = (unsigned)id < (unsigned)p.rows[f];
    const float* src = p.tables[f] + (size_t)(id_ok ? id : 0) * p.dims[f] + (c - p.offs[f]);
    const float* pes = p.pe + (size_t)s * p.d_model + c;
    if constexpr (VEC == 4) {
      const float4 ev = id_ok ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0, 0, 0, 0);
      const float4 pv = __ldg(reinterpret_cast<const float4*>(pes));
      e[0] = ev.x; e[1] = ev.y; e[2] = ev.z; e[3] = ev.w;
      pe[0] = pv.x; pe[1] = pv.y; pe[2] = pv.z; pe[3] = pv.w;
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        e[j] = id_ok ? __ldg(src + j) : 0.f;
        pe[j] = __ldg(pes + j);
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      o[j] = __fadd_rn(__fmul_rn(e[j], p.scale), pe[j]);
      if (p.thresh24) {
        const bool keep = dropout_keep(seed_v, p.site, (uint64_t)(t * p.d_model + c + j), p.thresh24);
        o[j] = keep ? __fmul_rn(o[j], p.inv_keep) : 0.f;
      }
    }
    const size_t off = (size_t)t * p.d_model + c;
    if constexpr (VEC == 4) {
      if (out) *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
      if (out_bf16) {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(out_bf16 + off) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (out) out[off + j] = o[j];
        if (out_bf16) out_bf16[off + j] = __float2bfloat16_rn(o[j]);
      }
    }
  }
}

// 128-bit path with U independent (id -> row) chains per thread: the id load and the row load it
// feeds are two dependent memory latencies, so one 16-byte load per thread in flight (the generic
// kernel above) keeps only ~32 KB per SM moving; U = 4 quadruples that.
template <int U>
__global__ void __launch_bounds__(256)
embed_fwd_vec_kernel(const EmbedParams p, int tok_shift, float* __restrict__ out,
                     __nv_bfloat16* __restrict__ out_bf16) {
  const uint64_t seed_v = resolve_seed(p.seed);
  const long per_tok = p.d_model / 4;
  const long total = p.T * per_tok;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
    long t[U];
    int c[U], f[U], id[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u * stride;
      ok[u] = i < total;
      t[u] = tok_shift >= 0 ? (i >> tok_shift) : i / per_tok;
      c[u] = (int)(i - t[u] * per_tok) * 4;
      f[u] = 0;
#pragma unroll
      for (int g = 1; g < B4CP_MAX_FEATURES; ++g)
        if (g < p.F && c[u] >= p.offs[g]) f[u] = g;
      id[u] = ok[u] ? __ldg(p.ids[f[u]] + t[u]) : -1;
    }
    float4 ev[U], pv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool id_ok = (unsigned)id[u] < (unsigned)p.rows[f[u]];
      ev[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      pv[u] = ev[u];
      if (ok[u]) {
        if (id_ok)
          ev[u] = __ldg(reinterpret_cast<const float4*>(
              p.tables[f[u]] + (size_t)id[u] * p.dims[f[u]] + (c[u] - p.offs[f[u]])));
        const int s = (int)(t[u] % p.S);
        pv[u] = __ldg(reinterpret_cast<const float4*>(p.pe + (size_t)s * p.d_model + c[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      float o[4] = {__fadd_rn(__fmul_rn(ev[u].x, p.scale), pv[u].x),
                    __fadd_rn(__fmul_rn(ev[u].y, p.scale), pv[u].y),
                    __fadd_rn(__fmul_rn(ev[u].z, p.scale), pv[u].z),
                    __fadd_rn(__fmul_rn(ev[u].w, p.scale), pv[u].w)};
      if (p.thresh24) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool keep = dropout_keep(seed_v, p.site, (uint64_t)(t[u] * p.d_model + c[u] + j), p.thresh24);
          o[j] = keep ? __fmul_rn(o[j], p.inv_keep) : 0.f;
        }
      }
      const size_t off = (size_t)t[u] * p.d_model + c[u];
      if (out) *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
      if (out_bf16) {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], o[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a);
        pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(out_bf16 + off) = pk;
      }
    }
  }
}

// =========================================================================== device scans
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*>= 32*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  const int res = smem[warp] + inc - v;
  if (total) *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const int* __restrict__ in, long n, int* __restrict__ tile_sums) {
  __shared__ int sm[40];
  const long base = (long)blockIdx.x * SCAN_TILE;
  int s = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long i = base + (long)threadIdx.x * SCAN_ITEMS + j;
    if (i < n) s += in[i];
  }
  int total;
  block_exclusive_scan(s, &total, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of up to SCAN_TILE*? values, looping
__global__ void __launch_bounds__(SCAN_THREADS)
scan_small_kernel(int* __restrict__ data, int n, int* __restrict__ total_out) {
  __shared__ int sm[40];
  int carry = 0;
  for (int base = 0; base < n; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int v = i < n ? data[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total, sm);
    if (i < n) data[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const int* __restrict__ in, int* __restrict__ out, long n,
                  const int* __restrict__ tile_offsets) {
  __shared__ int sm[40];
  const long base = (long)blockIdx.x * SCAN_TILE;
  int v[SCAN_ITEMS];
  int s = 0;
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long i = base + (long)threadIdx.x * SCAN_ITEMS + j;
    v[j] = i < n ? in[i] : 0;
    s += v[j];
  }
  int ex = block_exclusive_scan(s, nullptr, sm) + tile_offsets[blockIdx.x];
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    const long i = base + (long)threadIdx.x * SCAN_ITEMS + j;
    if (i < n) out[i] = ex;
    ex += v[j];
  }
}

// exclusive scan in -> out (may alias), optional device total; scratch >= ceil(n/SCAN_TILE) ints
static int exclusive_scan(const int* in, int* out, long n, int* scratch, int* total_out,
                          cudaStream_t st) {
  const int tiles = ceil_div(n, SCAN_TILE);
  scan_tile_sums_kernel<<<tiles, SCAN_THREADS, 0, st>>>(in, n, scratch);
  scan_small_kernel<<<1, SCAN_THREADS, 0, st>>>(scratch, tiles, total_out);
  scan_apply_kernel<<<tiles, SCAN_THREADS, 0, st>>>(in, out, n, scratch);
  note_launches(3);
  B4CP_LAUNCH_CHECK();
  return 0;
}

// =========================================================================== stable radix sort
// 8-bit LSD passes over (key = id, value = token index).  A block ranks a tile of
// SORT_ROUNDS*256 keys; within a round each warp holds 32 consecutive keys and ranks them with
// match.any; (round, warp) digit counts are prefix-summed in key order, which makes the pass
// stable and therefore the whole sort deterministic.
static constexpr int SORT_THREADS = 256;
static constexpr int SORT_ROUNDS = 8;
static constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;
static constexpr int SORT_SLOTS = SORT_ROUNDS * (SORT_THREADS / 32);

template <bool SCATTER>
__global__ void __launch_bounds__(SORT_THREADS)
radix_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, long n,
                  int shift, int* __restrict__ hist /*[256][nblocks]*/, int nblocks) {
  __shared__ uint16_t cnt[256][SORT_SLOTS + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 256 * (SORT_SLOTS + 2); i += SORT_THREADS)
    (&cnt[0][0])[i] = 0;
  __syncthreads();
  const long base = (long)blockIdx.x * SORT_TILE;
  uint32_t key[SORT_ROUNDS];
  uint32_t rank_in_group[SORT_ROUNDS];
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    const long i = base + r * SORT_THREADS + threadIdx.x;
    const bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0xFFFFFFFFu;
    const uint32_t dig = ok ? ((key[r] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, dig);
    rank_in_group[r] = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank_in_group[r] == 0) cnt[dig][r * (SORT_THREADS / 32) + warp] = (uint16_t)__popc(peers);
  }
  __syncthreads();
  {  // thread = digit: exclusive prefix over the (round, warp) slots
    const int dgt = threadIdx.x;
    uint32_t run = 0;
    for (int s = 0; s < SORT_SLOTS; ++s) {
      const uint32_t c = cnt[dgt][s];
      cnt[dgt][s] = (uint16_t)run;
      run += c;
    }
    if (!SCATTER) hist[(size_t)dgt * nblocks + blockIdx.x] = (int)run;
  }
  if (SCATTER) {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
      const long i = base + r * SORT_THREADS + threadIdx.x;
      if (i < n) {
        const uint32_t dig = (key[r] >> shift) & 255u;
        const long dst = (long)hist[(size_t)dig * nblocks + blockIdx.x] +
                         cnt[dig][r * (SORT_THREADS / 32) + warp] + rank_in_group[r];
        keys_out[dst] = key[r];
        vals_out[dst] = vals_in ? vals_in[i] : (uint32_t)i;
      }
    }
  }
}

// =========================================================================== segmented sum
__global__ void __launch_bounds__(256)
segment_flags_kernel(const uint32_t* __restrict__ sorted_keys, long n, int* __restrict__ flags) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1 : 0;
}

// seg_start[u] = sorted position where the u-th unique id begins (and seg_start[U] = n)
__global__ void __launch_bounds__(256)
segment_starts_kernel(const int* __restrict__ flags, const int* __restrict__ flag_scan, long n,
                      int* __restrict__ seg_start, const int* __restrict__ n_unique) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n && flags[i]) seg_start[flag_scan[i]] = (int)i;
  if (i == 0) seg_start[*n_unique] = (int)n;
}

__global__ void __launch_bounds__(256)
export_unique_kernel(const uint32_t* __restrict__ sorted_keys, const int* __restrict__ seg_start,
                     const int* __restrict__ n_unique, int* __restrict__ uniq_ids, long cap) {
  const long u = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (u < cap && u < *n_unique) uniq_ids[u] = (int)sorted_keys[seg_start[u]];
}

// ---------------------------------------------------------------------------------------------
// Tile-based segmented sum (default path).  The sorted token list is cut into fixed tiles of
// SEG_TILE positions, one warp per tile - no search, uniform work, coalesced key/token reads.  The
// warp walks its tile in order and keeps one running row (lane = columns lane, lane+32, ...):
//   * a run of equal ids that begins and ends inside the tile is scaled and written to the table;
//   * a run that continues from the previous tile (open left) or into the next (open right)
//     leaves an UNSCALED partial row: slot 0 = the tile's first run, slot 1 = its last run.
// segment_fixup_kernel then finishes every id whose tokens cross a tile boundary: the warp of the
// tile where the id BEGINS adds the partials of the following tiles in tile order (= token
// order), so the sum is bit-reproducible.  Ids spanning more than SEG_LONG tiles ([MASK], [CLS],
// [SEP], [PAD]: O(batch) duplicates) are queued for segment_long_kernel, where 32 warps sum fixed
// sub-ranges of the partial rows and the 32 results are added in order.
static constexpr int SEG_TILE = 64;
static constexpr int SEG_NV = 8;      // columns per lane: dim <= 256
static constexpr int SEG_LONG = 16;   // tiles

template <int NV>
__global__ void __launch_bounds__(256, NV <= 2 ? 3 : 1)   // d <= 64: <= 85 registers, 3 CTAs per SM
segment_tile_sum_kernel(const float* __restrict__ dout, int d_model, int off, int dim,
                        const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ stok,
                        long n, int rows, float scale, float inv_keep, uint32_t thresh24,
                        uint64_t seed, uint32_t site, float* __restrict__ table_grad,
                        float* __restrict__ partial) {
  seed = resolve_seed(seed);
  const int lane = threadIdx.x & 31;
  const long n_tiles = (n + SEG_TILE - 1) / SEG_TILE;
  for (long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_tiles;
       w += (long)gridDim.x * (blockDim.x >> 5)) {
    const long b = w * SEG_TILE;
    const int cnt = (int)min((long)SEG_TILE, n - b);
    uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu, t0 = 0, t1 = 0;
    if (lane < cnt) { k0 = skeys[b + lane]; t0 = stok[b + lane]; }
    if (lane + 32 < cnt) { k1 = skeys[b + 32 + lane]; t1 = stok[b + 32 + lane]; }
    const uint32_t k_first = __shfl_sync(0xffffffffu, k0, 0);
    const uint32_t k_last = cnt > 32 ? __shfl_sync(0xffffffffu, k1, cnt - 33)
                                     : __shfl_sync(0xffffffffu, k0, cnt - 1);
    const bool open_left = b > 0 && skeys[b - 1] == k_first;
    const bool open_right = b + cnt < n && skeys[b + cnt] == k_last;
    float acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.f;
    uint32_t run_key = k_first;
    bool run_is_first = true;
    auto flush = [&](bool is_last) {
      const bool part = (run_is_first && open_left) || (is_last && open_right);
      if (run_key < (uint32_t)rows) {
        float* dst = part ? partial + (size_t)(2 * w + (run_is_first ? 0 : 1)) * dim
                          : table_grad + (size_t)run_key * dim;
        const float sc = part ? 1.f : scale;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = lane + 32 * v;
          if (c < dim) dst[c] = part ? acc[v] : __fmul_rn(acc[v], sc);
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = 0.f;
    };
    // G tokens per step: all their row loads are issued before the first add (one warp keeps
    // G * NV 128-byte requests in flight instead of one row at a time).  The loads carry NO
    // branch: with `if (i < cnt && c < dim) x = load` every load sat in its own divergence region
    // and ncu showed 73 % of the stall samples on the first use of each loaded value, one DRAM
    // latency after another (82 us for the 213k tokens of a C1 batch, 8 % of the DRAM
    // throughput).  Lanes past the tile's end hold token 0 and columns past `dim` read column 0 -
    // valid addresses - and the values are zeroed afterwards.  (d <= 64 only: at 4 or 8 columns
    // per lane the same form needs 128 registers or spills, and measured 6 % slower at d = 256.)
    constexpr int G = NV <= 4 ? 8 : 4;
    for (int i0 = 0; i0 < cnt; i0 += G) {
      uint32_t kk[G];
      float g[G][NV];
      if constexpr (NV <= 2) {
        uint32_t tt[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int i = i0 + j;   // i0 is a multiple of G, G divides 32: one branch per step
          kk[j] = i0 < 32 ? __shfl_sync(0xffffffffu, k0, i & 31) : __shfl_sync(0xffffffffu, k1, i & 31);
          tt[j] = i0 < 32 ? __shfl_sync(0xffffffffu, t0, i & 31) : __shfl_sync(0xffffffffu, t1, i & 31);
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const float* src = dout + (size_t)tt[j] * d_model + off;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c = lane + 32 * v;
            g[j][v] = __ldg(src + (c < dim ? c : 0));
          }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const bool live = i0 + j < cnt;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c = lane + 32 * v;
            float x = g[j][v];
            if (thresh24) {
              const bool keep = dropout_keep(seed, site, (uint64_t)tt[j] * d_model + off + c, thresh24);
              x = keep ? __fmul_rn(x, inv_keep) : 0.f;
            }
            g[j][v] = (live && c < dim) ? x : 0.f;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const int i = i0 + j;   // i0 is a multiple of G, G divides 32: one branch per step
          kk[j] = i0 < 32 ? __shfl_sync(0xffffffffu, k0, i & 31) : __shfl_sync(0xffffffffu, k1, i & 31);
          const uint32_t ti = i0 < 32 ? __shfl_sync(0xffffffffu, t0, i & 31) : __shfl_sync(0xffffffffu, t1, i & 31);
          const float* src = dout + (size_t)ti * d_model + off;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c = lane + 32 * v;
            float x = 0.f;
            if (i < cnt && c < dim) {
              x = __ldg(src + c);
              if (thresh24) {
                const bool keep = dropout_keep(seed, site, (uint64_t)ti * d_model + off + c, thresh24);
                x = keep ? __fmul_rn(x, inv_keep) : 0.f;
              }
            }
            g[j][v] = x;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < G; ++j) {
        if (i0 + j < cnt) {
          if (kk[j] != run_key) {
            flush(false);
            run_key = kk[j];
            run_is_first = false;
          }
#pragma unroll
          for (int v = 0; v < NV; ++v) acc[v] = __fadd_rn(acc[v], g[j][v]);
        }
      }
    }
    flush(true);
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
segment_fixup_kernel(int dim, const uint32_t* __restrict__ skeys, long n, int rows, float scale,
                     const float* __restrict__ partial, float* __restrict__ table_grad,
                     int* __restrict__ long_count, int2* __restrict__ long_list, int long_cap) {
  const int lane = threadIdx.x & 31;
  const long n_tiles = (n + SEG_TILE - 1) / SEG_TILE;
  for (long a = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < n_tiles;
       a += (long)gridDim.x * (blockDim.x >> 5)) {
    const long b = a * SEG_TILE;
    const long e = min(n, b + SEG_TILE);
    if (e >= n) continue;
    const uint32_t K = skeys[e - 1];
    if (skeys[e] != K) continue;                                  // last run closed on the right
    const bool whole = skeys[b] == K;
    if (whole && b > 0 && skeys[b - 1] == K) continue;            // a continuation, not the start
    if (K >= (uint32_t)rows) continue;
    float acc[NV];
    const float* p0 = partial + (size_t)(2 * a + (whole ? 0 : 1)) * dim;
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = lane + 32 * v < dim ? p0[lane + 32 * v] : 0.f;
    long w = a + 1;
    bool is_long = false;
    while (true) {
      const float* pw = partial + (size_t)(2 * w) * dim;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (lane + 32 * v < dim) acc[v] = __fadd_rn(acc[v], pw[lane + 32 * v]);
      const long we = min(n, (w + 1) * SEG_TILE);
      const bool cont = we < n && skeys[we - 1] == K && skeys[we] == K;
      if (!cont) break;
      ++w;
      if (w - a > SEG_LONG) { is_long = true; break; }
    }
    if (is_long) {  // queue for the block-parallel kernel (queue order does not affect results)
      if (lane == 0) {
        const int slot = atomicAdd(long_count, 1);
        if (slot < long_cap) long_list[slot] = make_int2((int)a, (int)K);
      }
      continue;
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (lane + 32 * v < dim) table_grad[(size_t)K * dim + lane + 32 * v] = __fmul_rn(acc[v], scale);
  }
}

// one block per queued id: rows to add, in order: partial[a][slot], partial[a+1][0], ...,
// partial[b_end][0].  Warp q adds rows [q*per, (q+1)*per) in order; the 32 sums are added in order.
__global__ void __launch_bounds__(1024)
segment_long_kernel(int dim, const uint32_t* __restrict__ skeys, long n, float scale,
                    const float* __restrict__ partial, float* __restrict__ table_grad,
                    const int* __restrict__ long_count, const int2* __restrict__ long_list,
                    int long_cap) {
  __shared__ float part[32][256 + 1];
  __shared__ long s_end;
  const int n_long = min(*long_count, long_cap);
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  for (int li = blockIdx.x; li < n_long; li += gridDim.x) {
    const long a = long_list[li].x;
    const uint32_t K = (uint32_t)long_list[li].y;
    {  // upper bound of K in the sorted keys: 1024-ary search, two or three rounds
      long lo = (a + 1) * SEG_TILE, hi = n;   // skeys[lo] == K is known; answer in (lo, hi]
      while (hi - lo > 1) {
        const long step = (hi - lo + blockDim.x - 1) / blockDim.x;
        const long pos = lo + (long)threadIdx.x * step;
        const int below = __syncthreads_count(pos < hi && skeys[pos] <= K);   // monotone in pos
        const long nlo = lo + (long)(below - 1) * step;
        hi = min(hi, nlo + step);
        lo = nlo;
      }
      if (threadIdx.x == 0) s_end = hi;   // first position whose key is greater than K
    }
    __syncthreads();
    const long b_end = (s_end - 1) / SEG_TILE;
    const long R = b_end - a + 1;
    const long per = (R + 31) / 32;
    const bool whole = skeys[a * SEG_TILE] == K;
    const long r0 = q * per, r1 = min(R, r0 + per);
    for (int c = lane; c < dim; c += 32) {
      float acc = 0.f;
      for (long r = r0; r < r1; ++r) {
        const size_t prow = r == 0 ? (size_t)(2 * a + (whole ? 0 : 1)) : (size_t)(2 * (a + r));
        acc = __fadd_rn(acc, partial[prow * dim + c]);
      }
      part[q][c] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 32; ++w) acc = __fadd_rn(acc, part[w][c]);
      table_grad[(size_t)K * dim + c] = __fmul_rn(acc, scale);
    }
    __syncthreads();
  }
}

struct BwdWorkspace {
  uint32_t *keys0, *keys1, *vals0, *vals1;
  int *hist, *flags, *flag_scan, *seg_start, *scan_scratch, *counters;
  float* partial;
  int2* long_list;
  size_t bytes;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static BwdWorkspace carve_ws(void* base, long T, int max_dim) {
  BwdWorkspace w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    uint8_t* r = p ? p + o : nullptr;
    o += align256(bytes);
    return r;
  };
  const int nblocks = ceil_div(T, SORT_TILE);
  w.keys0 = (uint32_t*)take(T * 4);
  w.keys1 = (uint32_t*)take(T * 4);
  w.vals0 = (uint32_t*)take(T * 4);
  w.vals1 = (uint32_t*)take(T * 4);
  w.hist = (int*)take((size_t)256 * nblocks * 4);
  w.flags = (int*)take(T * 4);
  w.flag_scan = (int*)take(T * 4);
  w.seg_start = (int*)take((T + 1) * 4);
  w.scan_scratch = (int*)take((size_t)(ceil_div(std::max<long>(T, 256L * nblocks), SCAN_TILE) + 1) * 4);
  w.counters = (int*)take(64);
  // two partial rows (first / last run) per SEG_TILE tile
  w.partial = (float*)take((size_t)(2 * (T / SEG_TILE) + 4) * max_dim * 4);
  w.long_list = (int2*)take((size_t)(T / (SEG_TILE * SEG_LONG) + 2) * sizeof(int2));
  w.bytes = o;
  return w;
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_embed_fwd(const int32_t* const* h_ids, const float* const* h_tables,
                              const int* h_dims, const int* h_rows, int F, const float* pe,
                              int B, int S, float dropout_rate, uint64_t seed, uint32_t site,
                              float* out_f32, void* out_bf16, void* stream) {
  B4CP_CHECK_ARG(F >= 1 && F <= B4CP_MAX_FEATURES, "embed_fwd: F=%d out of range", F);
  B4CP_CHECK_ARG(out_f32 || out_bf16, "embed_fwd: no output");
  B4CP_CHECK_ARG(dropout_rate >= 0.f && dropout_rate < 1.f, "embed_fwd: bad dropout rate");
  if ((long)B * S == 0) return 0;
  EmbedParams p;
  p.F = F;
  int off = 0;
  bool vec4 = true;
  for (int f = 0; f < F; ++f) {
    p.ids[f] = h_ids[f];
    p.tables[f] = h_tables[f];
    p.dims[f] = h_dims[f];
    p.rows[f] = h_rows[f];
    p.offs[f] = off;
    off += h_dims[f];
    vec4 = vec4 && (h_dims[f] % 4 == 0) && (((uintptr_t)h_tables[f] & 15) == 0);
  }
  p.offs[F] = off;
  p.d_model = off;
  p.S = S;
  p.T = (long)B * S;
  p.pe = pe;
  p.scale = sqrtf((float)p.d_model);  // tf.math.sqrt(tf.cast(d_model, tf.float32))
  p.inv_keep = 1.0f / (1.0f - dropout_rate);
  p.thresh24 = (uint32_t)((double)dropout_rate * 16777216.0);
  p.seed = seed;
  p.site = site;
  vec4 = vec4 && (((uintptr_t)pe & 15) == 0) && (((uintptr_t)out_f32 & 15) == 0) &&
         (((uintptr_t)out_bf16 & 7) == 0);
  const long work = p.T * (p.d_model / (vec4 ? 4 : 1));
  const int blocks = (int)std::min<long>(ceil_div(work, 256), 148L * 16);
  if (vec4) {
    const int per_tok = p.d_model / 4;
    int tok_shift = -1;
    if ((per_tok & (per_tok - 1)) == 0) {
      tok_shift = 0;
      while ((1 << tok_shift) < per_tok) ++tok_shift;
    }
    const int vblocks = (int)std::min<long>(ceil_div(work, 256L * 4), 148L * 8);
    embed_fwd_vec_kernel<4><<<vblocks, 256, 0, (cudaStream_t)stream>>>(p, tok_shift, out_f32,
                                                                       (__nv_bfloat16*)out_bf16);
  } else
    embed_fwd_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, out_f32, (__nv_bfloat16*)out_bf16);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" long b4cp_embed_bwd_workspace_bytes(long tokens, int max_dim) {
  return (long)carve_ws(nullptr, tokens, max_dim).bytes;
}

// Stable LSD radix sort of (id, token index) for ONE feature into the workspace.  It depends on
// the ids alone - not on any gradient - so callers may run it early (b4cp_embed_sort on a side
// stream while the backward is still in the encoder layers) and finish with b4cp_embed_bwd_sorted.
static int embed_sort_impl(const int32_t* ids, long T, int rows, const BwdWorkspace& w, cudaStream_t st) {
  const int nblocks = ceil_div(T, SORT_TILE);
  int bits = 1;
  while ((1L << bits) < rows) ++bits;
  const int passes = (bits + 7) / 8;
  const uint32_t* kin = reinterpret_cast<const uint32_t*>(ids);
  const uint32_t* vin = nullptr;
  uint32_t* kout = w.keys0;
  uint32_t* vout = w.vals0;
  for (int ps = 0; ps < passes; ++ps) {
    radix_pass_kernel<false><<<nblocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, T, ps * 8,
                                                                w.hist, nblocks);
    if (exclusive_scan(w.hist, w.hist, 256L * nblocks, w.scan_scratch, nullptr, st)) return -1;
    radix_pass_kernel<true><<<nblocks, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, T, ps * 8,
                                                               w.hist, nblocks);
    note_launches(2);
    kin = kout;
    vin = vout;
    kout = (kout == w.keys0) ? w.keys1 : w.keys0;
    vout = (vout == w.vals0) ? w.vals1 : w.vals0;
  }
  B4CP_LAUNCH_CHECK();
  return 0;
}
// where embed_sort_impl leaves the sorted (keys, tokens): buffer 0 after an odd number of passes
static void sorted_buffers(int rows, const BwdWorkspace& w, const uint32_t** skeys, const uint32_t** stok) {
  int bits = 1;
  while ((1L << bits) < rows) ++bits;
  const int passes = (bits + 7) / 8;
  *skeys = (passes & 1) ? w.keys0 : w.keys1;
  *stok = (passes & 1) ? w.vals0 : w.vals1;
}

extern "C" int b4cp_embed_sort(const int32_t* ids, long tokens, int rows, int max_dim, void* workspace,
                               long workspace_bytes, void* stream) {
  B4CP_CHECK_ARG(tokens > 0 && tokens < (1L << 31), "embed_sort: bad token count %ld", tokens);
  BwdWorkspace w = carve_ws(workspace, tokens, max_dim);
  B4CP_CHECK_ARG(workspace && (long)w.bytes <= workspace_bytes,
                 "embed_sort: workspace too small (%ld < %zu)", workspace_bytes, w.bytes);
  return embed_sort_impl(ids, tokens, rows, w, (cudaStream_t)stream);
}

static int embed_bwd_impl(const float* dout, int d_model, int col_offset, int dim, const int32_t* ids,
                          long tokens, int rows, float dropout_rate, uint64_t seed, uint32_t site,
                          float* table_grad, int32_t* uniq_ids, int32_t* n_unique_out,
                          void* workspace, long workspace_bytes, bool presorted, void* stream);

extern "C" int b4cp_embed_bwd(const float* dout, int d_model, int col_offset, int dim,
                              const int32_t* ids, long tokens, int rows, float dropout_rate,
                              uint64_t seed, uint32_t site, float* table_grad, int32_t* uniq_ids,
                              int32_t* n_unique_out, void* workspace, long workspace_bytes,
                              void* stream) {
  return embed_bwd_impl(dout, d_model, col_offset, dim, ids, tokens, rows, dropout_rate, seed, site,
                        table_grad, uniq_ids, n_unique_out, workspace, workspace_bytes, false, stream);
}

extern "C" int b4cp_embed_bwd_sorted(const float* dout, int d_model, int col_offset, int dim, long tokens,
                                     int rows, float dropout_rate, uint64_t seed, uint32_t site,
                                     float* table_grad, void* workspace, long workspace_bytes,
                                     void* stream) {
  return embed_bwd_impl(dout, d_model, col_offset, dim, nullptr, tokens, rows, dropout_rate, seed, site,
                        table_grad, nullptr, nullptr, workspace, workspace_bytes, true, stream);
}

static int embed_bwd_impl(const float* dout, int d_model, int col_offset, int dim, const int32_t* ids,
                          long tokens, int rows, float dropout_rate, uint64_t seed, uint32_t site,
                          float* table_grad, int32_t* uniq_ids, int32_t* n_unique_out,
                          void* workspace, long workspace_bytes, bool presorted, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B4CP_CHECK_ARG(tokens > 0 && tokens < (1L << 31), "embed_bwd: bad token count %ld", tokens);
  BwdWorkspace w = carve_ws(workspace, tokens, dim);
  B4CP_CHECK_ARG(workspace && (long)w.bytes <= workspace_bytes,
                 "embed_bwd: workspace too small (%ld < %zu)", workspace_bytes, w.bytes);
  const long T = tokens;
  if (!presorted) {
    const int rc = embed_sort_impl(ids, T, rows, w, st);
    if (rc) return rc;
  }
  const uint32_t* kin;
  const uint32_t* vin;
  sorted_buffers(rows, w, &kin, &vin);
  B4CP_LAUNCH_CHECK();
  const uint32_t* skeys = kin;
  const uint32_t* stok = vin;
  int* n_unique = w.counters;
  int* long_count = w.counters + 2;
  const int tb = ceil_div(T, 256);
  B4CP_CHECK_ARG(dim <= 32 * SEG_NV, "embed_bwd: feature width %d > %d", dim, 32 * SEG_NV);
  B4CP_CUDA(cudaMemsetAsync(table_grad, 0, (size_t)rows * dim * sizeof(float), st));
  B4CP_CUDA(cudaMemsetAsync(long_count, 0, sizeof(int), st));
  const float scale = sqrtf((float)d_model);
  const float inv_keep = 1.0f / (1.0f - dropout_rate);
  const uint32_t thresh24 = (uint32_t)((double)dropout_rate * 16777216.0);
  const long n_tiles = (T + SEG_TILE - 1) / SEG_TILE;
  const int grid = (int)std::min<long>((n_tiles + 7) / 8, 148L * 16);
  const int long_cap = (int)(T / (SEG_TILE * SEG_LONG) + 2);
#define B4CP_SEG_LAUNCH(NV)                                                                        \
  do {                                                                                             \
    segment_tile_sum_kernel<NV><<<grid, 256, 0, st>>>(dout, d_model, col_offset, dim, skeys, stok, \
                                                      T, rows, scale, inv_keep, thresh24, seed,    \
                                                      site, table_grad, w.partial);                \
    segment_fixup_kernel<NV><<<grid, 256, 0, st>>>(dim, skeys, T, rows, scale, w.partial,          \
                                                   table_grad, long_count, w.long_list, long_cap); \
  } while (0)
  if (dim <= 64) B4CP_SEG_LAUNCH(2);
  else if (dim <= 128) B4CP_SEG_LAUNCH(4);
  else B4CP_SEG_LAUNCH(8);
#undef B4CP_SEG_LAUNCH
  segment_long_kernel<<<64, 1024, 0, st>>>(dim, skeys, T, scale, w.partial, table_grad, long_count,
                                           w.long_list, long_cap);
  note_launches(3);
  if (uniq_ids || n_unique_out) {  // optional export of the unique ids (not needed for the sums)
    segment_flags_kernel<<<tb, 256, 0, st>>>(skeys, T, w.flags);
    if (exclusive_scan(w.flags, w.flag_scan, T, w.scan_scratch, n_unique, st)) return -1;
    segment_starts_kernel<<<tb, 256, 0, st>>>(w.flags, w.flag_scan, T, w.seg_start, n_unique);
    note_launches(2);
  }
  if (uniq_ids) export_unique_kernel<<<tb, 256, 0, st>>>(skeys, w.seg_start, n_unique, uniq_ids, T);
  if (n_unique_out)
    B4CP_CUDA(cudaMemcpyAsync(n_unique_out, n_unique, sizeof(int), cudaMemcpyDeviceToDevice, st));
  B4CP_LAUNCH_CHECK();
  return 0;
}

// =========================================================================== output selection
// clickstream_transformer.py:260-297 (_gather_output_by_raw_value): positions whose first-feature
// token equals `value` ([MASK] = 1), in (b, s) order = flat token order.  row_index[i] is the
// token index of the i-th hit for i < count and -1 up to `capacity`.
namespace b4cp {
__global__ void __launch_bounds__(256)
select_flags_kernel(const int32_t* __restrict__ ids, long T, int value, int* __restrict__ flags) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < T) flags[i] = ids[i] == value ? 1 : 0;
}
__global__ void __launch_bounds__(256)
select_write_kernel(const int* __restrict__ flags, const int* __restrict__ pos, long T,
                    int32_t* __restrict__ row_index, long capacity, const int* __restrict__ count) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < T && flags[i] && pos[i] < capacity) row_index[pos[i]] = (int32_t)i;
  if (i < capacity && i >= *count) row_index[i] = -1;
}
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ x, int d, const int32_t* __restrict__ row_index,
                   long M, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                   long ld_bf16) {
  const long total = M * d;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long m = i / d;
    const int c = (int)(i - m * d);
    const int r = row_index[m];
    const float v = r >= 0 ? x[(size_t)r * d + c] : 0.f;  // zero vectors pad the ragged gather
    if (out_f32) out_f32[i] = v;
    if (out_bf16) out_bf16[m * ld_bf16 + c] = __float2bfloat16_rn(v);
  }
}
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ src, int d, const int32_t* __restrict__ row_index,
                    long M, float* __restrict__ dst) {
  const long total = M * d;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long m = i / d;
    const int c = (int)(i - m * d);
    const int r = row_index[m];
    if (r >= 0) dst[(size_t)r * d + c] = src[i];
  }
}
}  // namespace b4cp

extern "C" long b4cp_select_workspace_bytes(long tokens) {
  return (long)(2 * align256(tokens * 4) + align256((ceil_div(tokens, SCAN_TILE) + 1) * 4) + 256);
}

extern "C" int b4cp_select_masked(const int32_t* ids_first, long tokens, int value,
                                  int32_t* row_index, long capacity, int32_t* count_out,
                                  void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B4CP_CHECK_ARG(workspace && row_index && count_out, "select_masked: null argument");
  if (tokens == 0) {
    B4CP_CUDA(cudaMemsetAsync(count_out, 0, 4, st));
    if (capacity) B4CP_CUDA(cudaMemsetAsync(row_index, 0xFF, capacity * 4, st));
    return 0;
  }
  uint8_t* p = (uint8_t*)workspace;
  int* flags = (int*)p;
  int* pos = (int*)(p + align256(tokens * 4));
  int* scratch = (int*)(p + 2 * align256(tokens * 4));
  const long span = std::max(tokens, capacity);
  select_flags_kernel<<<ceil_div(tokens, 256), 256, 0, st>>>(ids_first, tokens, value, flags);
  if (exclusive_scan(flags, pos, tokens, scratch, count_out, st)) return -1;
  select_write_kernel<<<ceil_div(span, 256), 256, 0, st>>>(flags, pos, tokens, row_index, capacity,
                                                           count_out);
  note_launches(2);
  B4CP_LAUNCH_CHECK();
  return 0;
}

extern "C" int b4cp_gather_rows(const float* x, int d, const int32_t* row_index, long M,
                                float* out_f32, void* out_bf16, long ld_bf16, void* stream) {
  if (M == 0) return 0;
  const int blocks = (int)std::min<long>(ceil_div(M * d, 256), 148L * 16);
  gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, d, row_index, M, out_f32,
                                                                (__nv_bfloat16*)out_bf16, ld_bf16);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

/* dst must be zero-filled by the caller (rows that were not selected get no gradient) */
extern "C" int b4cp_scatter_rows(const float* src, int d, const int32_t* row_index, long M,
                                 float* dst, void* stream) {
  if (M == 0) return 0;
  const int blocks = (int)std::min<long>(ceil_div(M * d, 256), 148L * 16);
  scatter_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, d, row_index, M, dst);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

// Compacts the reference's padded label matrix (B, max_n_masked) float32 with LABEL_PAD = -1
// (input_pipeline.py:95-97, :198-214) into int32 labels in row-major order; out[i] = -1 beyond
// the number of valid labels.  The order equals the (b, s) order of the [MASK] positions.
namespace b4cp {
__global__ void __launch_bounds__(256)
label_flags_kernel(const float* __restrict__ labels, long n, float pad, int* __restrict__ flags) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) flags[i] = labels[i] != pad ? 1 : 0;
}
__global__ void __launch_bounds__(256)
label_write_kernel(const float* __restrict__ labels, const int* __restrict__ flags,
                   const int* __restrict__ pos, long n, int32_t* __restrict__ out, long capacity,
                   const int* __restrict__ count) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n && flags[i] && pos[i] < capacity) out[pos[i]] = (int32_t)labels[i];
  if (i < capacity && i >= *count) out[i] = -1;
}
}  // namespace b4cp

extern "C" int b4cp_compact_labels(const float* labels, long n, float label_pad, int32_t* out,
                                   long capacity, int32_t* count_out, void* workspace,
                                   void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  B4CP_CHECK_ARG(workspace && out && count_out, "compact_labels: null argument");
  if (n == 0) {
    B4CP_CUDA(cudaMemsetAsync(count_out, 0, 4, st));
    if (capacity) B4CP_CUDA(cudaMemsetAsync(out, 0xFF, capacity * 4, st));
    return 0;
  }
  uint8_t* p = (uint8_t*)workspace;
  int* flags = (int*)p;
  int* pos = (int*)(p + align256(n * 4));
  int* scratch = (int*)(p + 2 * align256(n * 4));
  label_flags_kernel<<<ceil_div(n, 256), 256, 0, st>>>(labels, n, label_pad, flags);
  if (exclusive_scan(flags, pos, n, scratch, count_out, st)) return -1;
  label_write_kernel<<<ceil_div(std::max(n, capacity), 256), 256, 0, st>>>(labels, flags, pos, n,
                                                                           out, capacity, count_out);
  note_launches(2);
  B4CP_LAUNCH_CHECK();
  return 0;
}
