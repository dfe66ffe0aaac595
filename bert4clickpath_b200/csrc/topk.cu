// Exact top-k over materialised score rows with the tf.math.top_k order: score descending,
// ties broken by the LOWER index (examples/BERT4Rec/source/utils.py:176, :245).
//
// One CTA per row.  Each (score, id) pair becomes one integer key
//     K = (~ordered(score) << idbits) | id          (smaller K = better rank, all K distinct)
// and the k smallest keys are found by an MSB-first radix select with 11-bit digits that stops
// as soon as the undecided bucket fits in shared memory; survivors are bitonic-sorted.  The id is
// part of the key, so any number of exact ties is resolved exactly.
#include <algorithm>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

static constexpr int TK_THREADS = 1024;
static constexpr int TK_BITS = 11;
static constexpr int TK_BINS = 1 << TK_BITS;
static constexpr int TK_MAXK = 256;
static constexpr int TK_CAP = 2048;               // undecided bucket must shrink below this
static constexpr int TK_SORT = 4096;              // >= TK_CAP + TK_MAXK, power of two

__device__ __forceinline__ uint32_t ordered_desc(float f) {
  if (f == 0.f) f = 0.f;  // -0 == +0 for the comparison TensorFlow does
  const uint32_t u = __float_as_uint(f);
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}
__device__ __forceinline__ float from_ordered_desc(uint32_t d) {
  const uint32_t asc = ~d;
  const uint32_t u = (asc & 0x80000000u) ? (asc & 0x7FFFFFFFu) : ~asc;
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const float* __restrict__ scores, const int32_t* __restrict__ cand_ids, long ld,
                 int V, int k, int idbits, int32_t* __restrict__ out_ids,
                 float* __restrict__ out_scores, long ld_out) {
  __shared__ int hist[TK_BINS];
  __shared__ unsigned long long buf[TK_SORT];
  __shared__ int scan_tmp[40];
  __shared__ int s_bin, s_below, s_nsel;
  const long row = blockIdx.x;
  const float* z = scores + row * ld;
  // candidate mode: element v carries the id cand[v] (negative = empty slot, skipped)
  const int32_t* cand = cand_ids ? cand_ids + row * ld : nullptr;
  const int total_bits = 32 + idbits;
  int n_valid = V;
  if (cand) {
    if (threadIdx.x == 0) s_nsel = 0;
    __syncthreads();
    int c = 0;
    for (int v = threadIdx.x; v < V; v += TK_THREADS) c += cand[v] >= 0 ? 1 : 0;
    if (c) atomicAdd(&s_nsel, c);
    __syncthreads();
    n_valid = s_nsel;
    __syncthreads();
  }
  int need = min(k, n_valid);
  if (need == 0) {  // nothing to rank
    for (int r = threadIdx.x; r < k; r += TK_THREADS) {
      out_ids[row * ld_out + r] = -1;
      if (out_scores) out_scores[row * ld_out + r] = -INFINITY;
    }
    return;
  }
  unsigned long long prefix = 0;
  int consumed = 0;
  int bin_count = V;
  while (true) {
    const int bits = min(TK_BITS, total_bits - consumed);
    const int shift = total_bits - consumed - bits;
    for (int i = threadIdx.x; i < TK_BINS; i += TK_THREADS) hist[i] = 0;
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += TK_THREADS) {
      const int id = cand ? cand[v] : v;
      if (id < 0) continue;
      const unsigned long long K = ((unsigned long long)ordered_desc(z[v]) << idbits) | (unsigned)id;
      if (consumed == 0 || (K >> (shift + bits)) == prefix)
        atomicAdd(&hist[(int)((K >> shift) & ((1u << bits) - 1u))], 1);
    }
    __syncthreads();
    // locate the bucket holding the need-th smallest key: thread t owns bins 2t, 2t+1
    {
      const int b0 = 2 * threadIdx.x;
      const int c0 = hist[b0], c1 = hist[b0 + 1];
      // block exclusive scan of (c0 + c1)
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      int inc = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (lane == 31) scan_tmp[warp] = inc;
      __syncthreads();
      if (warp == 0) {
        const int w = scan_tmp[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, winc, o);
          if (lane >= o) winc += n;
        }
        scan_tmp[lane] = winc - w;
      }
      __syncthreads();
      const int ex = scan_tmp[warp] + inc - (c0 + c1);
      if (ex < need && need <= ex + c0) {
        s_bin = b0;
        s_below = ex;
      } else if (ex + c0 < need && need <= ex + c0 + c1) {
        s_bin = b0 + 1;
        s_below = ex + c0;
      }
      __syncthreads();
    }
    const int b = s_bin;
    need -= s_below;
    bin_count = hist[b];
    prefix = (prefix << bits) | (unsigned)b;
    consumed += bits;
    __syncthreads();
    if (bin_count <= TK_CAP || consumed >= total_bits) break;
  }
  // collect winners (top bits < prefix) and the undecided bucket (== prefix)
  if (threadIdx.x == 0) s_nsel = 0;
  for (int i = threadIdx.x; i < TK_SORT; i += TK_THREADS) buf[i] = ~0ull;
  __syncthreads();
  const int rem = total_bits - consumed;
  for (int v = threadIdx.x; v < V; v += TK_THREADS) {
    const int id = cand ? cand[v] : v;
    if (id < 0) continue;
    const unsigned long long K = ((unsigned long long)ordered_desc(z[v]) << idbits) | (unsigned)id;
    if ((K >> rem) <= prefix) {
      const int slot = atomicAdd(&s_nsel, 1);
      if (slot < TK_SORT) buf[slot] = K;
    }
  }
  __syncthreads();
  // bitonic sort ascending
  for (int size = 2; size <= TK_SORT; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < TK_SORT / 2; i += TK_THREADS) {
        const int lo = (i / stride) * (stride << 1) + (i % stride);
        const int hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const unsigned long long a = buf[lo], c = buf[hi];
        if ((a > c) == asc) {
          buf[lo] = c;
          buf[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  const unsigned long long idmask = (1ull << idbits) - 1ull;
  for (int r = threadIdx.x; r < k; r += TK_THREADS) {
    const unsigned long long K = buf[r];
    const bool ok = r < n_valid && K != ~0ull;
    out_ids[row * ld_out + r] = ok ? (int32_t)(K & idmask) : -1;
    if (out_scores)
      out_scores[row * ld_out + r] = ok ? from_ordered_desc((uint32_t)(K >> idbits)) : -INFINITY;
  }
}

}  // namespace b4cp

using namespace b4cp;

extern "C" int b4cp_topk_rows(const float* scores, long ld, long rows, int V, int k,
                              int32_t* out_ids, float* out_scores, long ld_out, void* stream) {
  B4CP_CHECK_ARG(k >= 1 && k <= TK_MAXK, "topk: k=%d must be in [1,%d]", k, TK_MAXK);
  B4CP_CHECK_ARG(V >= 1, "topk: empty vocabulary");
  B4CP_CHECK_ARG(ld_out >= k, "topk: ld_out < k");
  if (rows == 0) return 0;
  int idbits = 1;
  while ((1L << idbits) < V) ++idbits;
  topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, (cudaStream_t)stream>>>(
      scores, nullptr, ld, V, k, idbits, out_ids, out_scores, ld_out);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

/* top-k over explicit (score, id) candidate lists: row r ranks cand_scores[r*ld .. +n_cand) whose
 * ids are cand_ids[...] (negative = empty).  Same order as b4cp_topk_rows; ids < V. */
extern "C" int b4cp_topk_candidates(const float* cand_scores, const int32_t* cand_ids, long ld,
                                    long rows, int n_cand, int V, int k, int32_t* out_ids,
                                    float* out_scores, long ld_out, void* stream) {
  B4CP_CHECK_ARG(k >= 1 && k <= TK_MAXK, "topk: k=%d must be in [1,%d]", k, TK_MAXK);
  B4CP_CHECK_ARG(cand_scores && cand_ids && out_ids, "topk_candidates: null argument");
  B4CP_CHECK_ARG(ld_out >= k, "topk: ld_out < k");
  if (rows == 0) return 0;
  int idbits = 1;
  while ((1L << idbits) < V) ++idbits;
  topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, (cudaStream_t)stream>>>(
      cand_scores, cand_ids, ld, n_cand, k, idbits, out_ids, out_scores, ld_out);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}
