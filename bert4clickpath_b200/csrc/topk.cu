// Exact top-k over materialised score rows with the tf.math.top_k order: score descending,
// ties broken by the LOWER index (examples/BERT4Rec/source/utils.py:176, :245).
//
// One CTA per row.  Each (score, id) pair becomes one integer key
//     K = (~ordered(score) << idbits) | id          (smaller K = better rank, all K distinct)
// and the k smallest keys are found by an MSB-first radix select with 11-bit digits that stops
// as soon as the undecided bucket fits in shared memory; survivors are bitonic-sorted.  The id is
// part of the key, so any number of exact ties is resolved exactly.
//
// Long rows (V >= 16384, the 1M-item catalogue of SURVEY.md C5) take a SINGLE streaming pass
// instead (topk_stream_kernel): 128-bit loads, a running threshold tau = the k-th best key seen so
// far, and a shared-memory candidate buffer that is compacted (rank by counting, keep k) whenever
// it holds more than k + 160 keys.  A random row admits ~k ln(V/256) candidates in total, so the
// row is read exactly once at HBM speed.  A row that would overflow the buffer
// (adversarially ordered scores) is flagged and redone by the radix-select kernel - the result is
// exact either way.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

static constexpr int TK_THREADS = 1024;
static constexpr int TK_BITS = 11;
static constexpr int TK_BINS = 1 << TK_BITS;
static constexpr int TK_MAXK = 256;
static constexpr int TK_CAP = 2048;               // undecided bucket must shrink below this
static constexpr int TK_SORT = 4096;              // >= TK_CAP + TK_MAXK, power of two
static constexpr int TK_REDO = -2;                // out_ids[row][0] marker: redo with radix select

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

// One row, whole CTA (every thread calls it with the same row).
__device__ __forceinline__ void
topk_row(const long row, const float* __restrict__ scores, const int32_t* __restrict__ cand_ids, long ld,
         int V, int k, int idbits, int32_t* __restrict__ out_ids, float* __restrict__ out_scores,
         long ld_out, const int* __restrict__ extra_count, int base_count) {
  // counted candidate lists: row r holds base_count + extra_count[r] entries (at most V)
  if (extra_count) V = min(V, base_count + max(extra_count[row], 0));
  __shared__ int hist[TK_BINS];
  __shared__ unsigned long long buf[TK_SORT];
  __shared__ int scan_tmp[40];
  __shared__ int s_bin, s_below, s_nsel;
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
    // stop once the undecided bucket is small; short (candidate-list) rows go on until it is
    // SMALL - the final sort is sized by what is collected, and its ~log^2 barrier-separated
    // stages dominated the merge of ~1.5K-entry lists when it always ran over TK_SORT keys
    if (bin_count <= (V <= 16384 ? TK_CAP / 8 : TK_CAP) || consumed >= total_bits) break;
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
  // bitonic sort ascending over the smallest power of two that holds what was collected (unused
  // slots carry ~0 and sort to the end)
  int sort_n = 2;
  while (sort_n < min(s_nsel, TK_SORT)) sort_n <<= 1;
  if (sort_n < k) {
    while (sort_n < k) sort_n <<= 1;
  }
  for (int size = 2; size <= sort_n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < sort_n / 2; i += TK_THREADS) {
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

__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const float* __restrict__ scores, const int32_t* __restrict__ cand_ids, long ld,
                 int V, int k, int idbits, int32_t* __restrict__ out_ids,
                 float* __restrict__ out_scores, long ld_out, int only_flagged,
                 const int* __restrict__ extra_count = nullptr, int base_count = 0) {
  // second launch after a single-pass kernel: only rows it flagged (out_ids[row][0] == -2)
  if (only_flagged && out_ids[(long)blockIdx.x * ld_out] != TK_REDO) return;
  topk_row(blockIdx.x, scores, cand_ids, ld, V, k, idbits, out_ids, out_scores, ld_out, extra_count,
           base_count);
}

// The redo pass after a single-pass kernel as a FIXED small grid: one CTA per row costs a CTA
// launch (1,024 threads, 50 KB of shared memory) per row just to find the flag clear - ~60 us for
// the 9,472 rows of a C1 ranking batch, a tenth of the whole top-k.  Here every CTA reads the
// flags of its rows with one parallel load, lists the flagged ones (almost always none) and
// redoes those.
static constexpr int TK_REDO_CHUNK = 64;   // rows whose flags one CTA inspects at a time

__global__ void __launch_bounds__(TK_THREADS)
topk_redo_kernel(const float* __restrict__ scores, long ld, long rows, int V, int k, int idbits,
                 int32_t* __restrict__ out_ids, float* __restrict__ out_scores, long ld_out) {
  __shared__ int s_n;
  __shared__ int s_rows[TK_REDO_CHUNK];
  for (long base = (long)blockIdx.x * TK_REDO_CHUNK; base < rows; base += (long)gridDim.x * TK_REDO_CHUNK) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const long mine = base + threadIdx.x;
    if (threadIdx.x < TK_REDO_CHUNK && mine < rows && out_ids[mine * ld_out] == TK_REDO)
      s_rows[atomicAdd(&s_n, 1)] = (int)(mine - base);
    __syncthreads();
    const int n = s_n;
    for (int i = 0; i < n; ++i) {
      const long row = base + s_rows[i];
      __syncthreads();
      topk_row(row, scores, nullptr, ld, V, k, idbits, out_ids, out_scores, ld_out, nullptr, 0);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ single-pass streaming top-k
static constexpr int ST_THREADS = 512;
static constexpr int ST_UNROLL = 4;                              // float4 loads in flight per thread
static constexpr int ST_BATCH = ST_THREADS * ST_UNROLL * 4;      // 8192 elements per full batch
static constexpr int ST_CAP = 4096;                              // candidate keys (+ ST_CAP scratch)
static constexpr int ST_SEED = 256;                              // elements that seed the threshold
static constexpr int ST_SLACK = 160;                             // compact above k + ST_SLACK keys
static constexpr int ST_RANK_MAX = 1024;                         // rank-by-counting up to this many
static constexpr int ST_MIN_V = 16384;

// ascending bitonic sort of n = 2^logn 64-bit keys in shared memory, whole block (shifts and
// masks only).  Only used when a batch admits more than ST_RANK_MAX candidates.
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* buf, int logn) {
  const int half = 1 << (logn - 1);
  for (int ls = 1; ls <= logn; ++ls) {          // size = 2^ls
    for (int lt = ls - 1; lt >= 0; --lt) {      // stride = 2^lt
      const int stride = 1 << lt;
      for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const int lo = ((i >> lt) << (lt + 1)) | (i & (stride - 1));
        const int hi = lo | stride;
        const bool asc = ((lo >> ls) & 1) == 0;
        const unsigned long long a = buf[lo], c = buf[hi];
        if ((a > c) == asc) {
          buf[lo] = c;
          buf[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Keep the min(c, k) smallest of the c keys in buf, sorted ascending in buf[0..); returns the new
// threshold (the k-th smallest, or "accept everything" while fewer than k keys exist).
// c <= ST_RANK_MAX: rank by counting - thread t owns keys t and t + 512 and counts the keys below
// each (all lanes read the same key: a shared-memory broadcast); ~6 c instructions per thread,
// against ~25 * 4 * log^2 instructions for a sort of the same buffer, which made the sorts
// 45% of all instructions of the first version of this kernel.
__device__ __forceinline__ unsigned long long st_compact(unsigned long long* buf, int c, int k,
                                                         int* s_count) {
  const int tid = threadIdx.x;
  const int n = min(c, k);
  if (c <= ST_RANK_MAX) {
    unsigned long long* dst = buf + ST_CAP;
    const unsigned long long my0 = tid < c ? buf[tid] : ~0ull;
    const unsigned long long my1 = tid + ST_THREADS < c ? buf[tid + ST_THREADS] : ~0ull;
    int r0 = 0, r1 = 0;
    if (tid < c) {
      if (c > ST_THREADS) {
#pragma unroll 4
        for (int i = 0; i < c; ++i) {
          const unsigned long long K = buf[i];
          r0 += K < my0 ? 1 : 0;
          r1 += K < my1 ? 1 : 0;
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < c; ++i) r0 += buf[i] < my0 ? 1 : 0;
      }
    }
    __syncthreads();
    if (tid < c && r0 < k) dst[r0] = my0;
    if (tid + ST_THREADS < c && r1 < k) dst[r1] = my1;
    __syncthreads();
    for (int i = tid; i < n; i += ST_THREADS) buf[i] = dst[i];
  } else {
    int logn = 1;
    while ((1 << logn) < c) ++logn;
    for (int i = c + tid; i < (1 << logn); i += ST_THREADS) buf[i] = ~0ull;
    __syncthreads();
    block_bitonic_sort(buf, logn);
  }
  __syncthreads();
  const unsigned long long tau = c >= k ? buf[k - 1] : ~0ull;
  if (tid == 0) *s_count = n;
  __syncthreads();
  return tau;
}

__global__ void __launch_bounds__(ST_THREADS, 2)
topk_stream_kernel(const float* __restrict__ scores, long ld, int V, int k, int idbits,
                   int32_t* __restrict__ out_ids, float* __restrict__ out_scores, long ld_out) {
  extern __shared__ __align__(16) unsigned long long st_buf[];  // [2 * ST_CAP]
  __shared__ int s_count;
  const long row = blockIdx.x;
  const float* z = scores + row * ld;
  const int tid = threadIdx.x;
  // the first ST_SEED elements seed the buffer and the threshold
  if (tid < ST_SEED)
    st_buf[tid] = ((unsigned long long)ordered_desc(__ldg(z + tid)) << idbits) | (unsigned)tid;
  __syncthreads();
  unsigned long long tau = st_compact(st_buf, ST_SEED, k, &s_count);
  // batches double from ST_SEED to ST_BATCH elements: while little has been seen the threshold is
  // loose and a batch as large as everything seen so far admits about k candidates
  auto load_batch = [&](int base, int lim, float (&x)[ST_UNROLL][4]) {
#pragma unroll
    for (int u = 0; u < ST_UNROLL; ++u) {
      const int idx = base + (u * ST_THREADS + tid) * 4;
      if (idx + 3 < lim) {
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(z + idx));
        x[u][0] = v4.x; x[u][1] = v4.y; x[u][2] = v4.z; x[u][3] = v4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[u][j] = idx + j < lim ? __ldg(z + idx + j) : -INFINITY;
      }
    }
  };
  float x[ST_UNROLL][4], xn[ST_UNROLL][4];
  int base = ST_SEED, bsz = ST_SEED;
  load_batch(base, min(V, base + bsz), x);
  while (base < V) {
    const int lim = min(V, base + bsz);
    const int nbsz = min(ST_BATCH, bsz * 2);
    if (lim < V) load_batch(lim, min(V, lim + nbsz), xn);   // in flight while x is filtered
    // fast filter on the float value (x >= threshold score; -0 == +0 as in the key), exact
    // 64-bit key comparison (ties -> lower id) only for the few that pass
    const float tau_f = from_ordered_desc((uint32_t)(tau >> idbits));
#pragma unroll
    for (int u = 0; u < ST_UNROLL; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (x[u][j] >= tau_f) {
          const int v = base + (u * ST_THREADS + tid) * 4 + j;
          const unsigned long long K = ((unsigned long long)ordered_desc(x[u][j]) << idbits) | (unsigned)v;
          if (K < tau && v < lim) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < ST_CAP) st_buf[slot] = K;
          }
        }
      }
    }
    __syncthreads();
    const int c = s_count;
    if (c > ST_CAP) {  // would have dropped candidates: hand the row to the radix-select kernel
      if (tid == 0) out_ids[row * ld_out] = TK_REDO;
      return;
    }
    if (c > k + ST_SLACK) tau = st_compact(st_buf, c, k, &s_count);
#pragma unroll
    for (int u = 0; u < ST_UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) x[u][j] = xn[u][j];
    base = lim;
    bsz = nbsz;
  }
  st_compact(st_buf, s_count, k, &s_count);   // sorted winners in st_buf[0..k)
  const unsigned long long idmask = (1ull << idbits) - 1ull;
  for (int r = tid; r < k; r += ST_THREADS) {
    const unsigned long long K = st_buf[r];
    out_ids[row * ld_out + r] = (int32_t)(K & idmask);
    if (out_scores) out_scores[row * ld_out + r] = from_ordered_desc((uint32_t)(K >> idbits));
  }
}

// ------------------------------------------------------------------ sampled-threshold streaming top-k
// One pass with a STATIC threshold taken from a strided sample of the row: the r-th best of n_s
// sampled scores (r ~ 4 k n_s / V) sits near population rank 4k, so admitting every score at
// least as good as it leaves ~4k candidates - no compaction and no block-wide barrier during the
// pass - and the exact top-k of those candidates finishes.  Correctness never depends on the
// sample: if fewer than k or more than SM_CAP candidates were admitted (astronomically unlikely for
// random order; certain for e.g. an all-ties row), the row is flagged and redone by the
// radix-select kernel.  This is the variant for rows of 16K..256K scores, where the running
// threshold of topk_stream_kernel spends more time compacting than reading.
static constexpr int SM_THREADS = 512;
static constexpr int SM_CAP = 4096;

template <int SPT>   // sampled scores per thread: n_s = 512 * SPT
__global__ void __launch_bounds__(SM_THREADS, 2)
topk_sample_kernel(const float* __restrict__ scores, long ld, int V, int k, int idbits, int rounds_target,
                   int32_t* __restrict__ out_ids, float* __restrict__ out_scores, long ld_out) {
  extern __shared__ __align__(16) unsigned long long sm_buf[];  // [2 * SM_CAP] (second half: scratch)
  __shared__ uint32_t s_wmin[SM_THREADS / 32];
  __shared__ int s_count;
  const long row = blockIdx.x;
  const float* z = scores + row * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NS = SM_THREADS * SPT;
  // ---- 1. threshold = the sample's r-th best distinct key (ordered_desc: smaller = better)
  uint32_t key[SPT];
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    const long j = (long)(tid + SM_THREADS * q) * V / NS;
    key[q] = ordered_desc(__ldg(z + j));
  }
  if (tid == 0) s_count = 0;
  uint32_t tau_od = 0;
  int removed = 0;
  while (removed < rounds_target) {
    uint32_t m = 0xFFFFFFFFu;
#pragma unroll
    for (int q = 0; q < SPT; ++q) m = min(m, key[q]);
    m = __reduce_min_sync(0xffffffffu, m);
    if (lane == 0) s_wmin[warp] = m;
    __syncthreads();
    m = s_wmin[lane & (SM_THREADS / 32 - 1)];
    m = __reduce_min_sync(0xffffffffu, m);
    if (m == 0xFFFFFFFFu) break;   // sample exhausted (cannot happen for V >= 16384 real scores)
    int mine = 0;
#pragma unroll
    for (int q = 0; q < SPT; ++q)
      if (key[q] == m) {
        key[q] = 0xFFFFFFFFu;
        ++mine;
      }
    removed += __syncthreads_count(mine > 0);   // >= 1 per round; ties at m are removed together
    tau_od = m;
  }
  const float tau_f = from_ordered_desc(tau_od);
  // ---- 2. one pass: admit every score >= tau
  auto load_batch = [&](int base, float (&x)[4][4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + (u * SM_THREADS + tid) * 4;
      if (idx + 3 < V) {
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(z + idx));
        x[u][0] = v4.x; x[u][1] = v4.y; x[u][2] = v4.z; x[u][3] = v4.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[u][j] = idx + j < V ? __ldg(z + idx + j) : -INFINITY;
      }
    }
  };
  constexpr int BATCH = SM_THREADS * 16;
  float x[4][4], xn[4][4];
  load_batch(0, x);
  for (int base = 0; base < V; base += BATCH) {
    if (base + BATCH < V) load_batch(base + BATCH, xn);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (x[u][j] >= tau_f) {
          const int v = base + (u * SM_THREADS + tid) * 4 + j;
          if (v < V) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < SM_CAP)
              sm_buf[slot] = ((unsigned long long)ordered_desc(x[u][j]) << idbits) | (unsigned)v;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) x[u][j] = xn[u][j];
  }
  __syncthreads();
  const int c = s_count;
  if (c > SM_CAP || c < k) {   // threshold too loose / too tight: exact redo by the radix kernel
    if (tid == 0) out_ids[row * ld_out] = TK_REDO;
    return;
  }
  // ---- 3. exact top-k of the admitted candidates (sorted into sm_buf[0..k))
  __syncthreads();
  {
    unsigned long long* buf = sm_buf;
    if (c <= ST_RANK_MAX) {
      unsigned long long* dst = buf + SM_CAP;
      const unsigned long long my0 = tid < c ? buf[tid] : ~0ull;
      const unsigned long long my1 = tid + SM_THREADS < c ? buf[tid + SM_THREADS] : ~0ull;
      int r0 = 0, r1 = 0;
      if (tid < c) {
        if (c > SM_THREADS) {
#pragma unroll 4
          for (int i = 0; i < c; ++i) {
            const unsigned long long K = buf[i];
            r0 += K < my0 ? 1 : 0;
            r1 += K < my1 ? 1 : 0;
          }
        } else {
#pragma unroll 4
          for (int i = 0; i < c; ++i) r0 += buf[i] < my0 ? 1 : 0;
        }
      }
      __syncthreads();
      if (tid < c && r0 < k) dst[r0] = my0;
      if (tid + SM_THREADS < c && r1 < k) dst[r1] = my1;
      __syncthreads();
      for (int i = tid; i < k; i += SM_THREADS) buf[i] = dst[i];
    } else {
      int logn = 1;
      while ((1 << logn) < c) ++logn;
      for (int i = c + tid; i < (1 << logn); i += SM_THREADS) buf[i] = ~0ull;
      __syncthreads();
      block_bitonic_sort(buf, logn);
    }
    __syncthreads();
  }
  const unsigned long long idmask = (1ull << idbits) - 1ull;
  for (int r = tid; r < k; r += SM_THREADS) {
    const unsigned long long K = sm_buf[r];
    out_ids[row * ld_out + r] = (int32_t)(K & idmask);
    if (out_scores) out_scores[row * ld_out + r] = from_ordered_desc((uint32_t)(K >> idbits));
  }
}


// ------------------------------------------------------ sampled threshold, rows of 16K..128K scores
// Same algorithm as topk_sample_kernel, re-shaped for rows that take only ~10 us to stream (the
// C1 vocabulary: 54,293 scores = 217 KB), where everything that is not streaming shows:
//   * 256 threads and 34 KB of shared memory per CTA -> FOUR rows per SM, so the threshold and
//     finish phases of one row hide under the streaming of three others (the 512-thread kernel
//     keeps two, and sat at 0.34 of the HBM roofline at V = 54K against 0.77 at V = 1M);
//   * the sample is 512 float4 (2,048 scores in 512 sectors instead of 2,048);
//   * the r-th best sample key is found WITHOUT block-wide rounds: every warp peels the r best
//     distinct keys of its own 256 samples with warp reductions only, one barrier, then warp 0
//     peels the r best distinct keys of those 8 r candidates (the r-th best of the whole sample is
//     among them).  The 512-thread kernel pays two block barriers per round, ~3 us per row.
// The threshold only has to be roughly right: too few (< k) or too many (> SS_CAP) admitted scores
// flag the row for the exact radix-select redo, as before.
static constexpr int SS_THREADS = 256;
static constexpr int SS_CAP = 2048;
static constexpr int SS_V4 = 2;                       // float4 samples per thread
static constexpr int SS_RMAX = 64;                    // rounds_target is clamped to [8, 64]
static constexpr int SS_NS = SS_THREADS * SS_V4 * 4;  // 2,048 sampled scores

__global__ void __launch_bounds__(SS_THREADS, 4)
topk_sample_small_kernel(const float* __restrict__ scores, long ld, int V, int k, int idbits,
                         int rounds_tight, int rounds_loose, int32_t* __restrict__ out_ids,
                         float* __restrict__ out_scores, long ld_out) {
  extern __shared__ __align__(16) unsigned long long sm_buf[];  // [2 * SS_CAP] (second half: scratch)
  constexpr int NW = SS_THREADS / 32;
  __shared__ uint32_t s_cand[NW * SS_RMAX];
  __shared__ uint32_t s_tau;
  __shared__ int s_count;
  const long row = blockIdx.x;
  const float* z = scores + row * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Streaming pass helpers.  ncu (profiles/r2c_ncu_topk_small_summary.txt) showed the first version
  // of this kernel ISSUE-bound (73 % issue-active, 55 % DRAM): 218 instructions per warp and batch
  // of 16 scores per thread - a convergence barrier + compare + branch per score, bounds checks
  // around every load, 32 register moves for the double buffer.  Now: full batches carry no
  // bounds checks (the < 4,096-score tail is a separate scalar loop), the two register sets
  // alternate (no moves), and a thread tests the MAXIMUM of its 16 scores once per batch - 0.4 %
  // of the scores pass, so 94 % of the tests skip the per-score code.
  constexpr int BATCH = SS_THREADS * 16;
  const int n_full = V / BATCH;
  auto load_full = [&](int bi, float4 (&q)[4]) {
    const float4* p = reinterpret_cast<const float4*>(z) + (size_t)bi * (BATCH / 4) + tid;
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(p + u * SS_THREADS);
  };
  float tau_f = 0.f;
  auto admit = [&](float xv, int v) {
    if (xv >= tau_f) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < SS_CAP) sm_buf[slot] = ((unsigned long long)ordered_desc(xv) << idbits) | (unsigned)v;
    }
  };
  auto scan16 = [&](const float4 (&q)[4], int base) {
    float m = fmaxf(fmaxf(q[0].x, q[0].y), fmaxf(q[0].z, q[0].w));
#pragma unroll
    for (int u = 1; u < 4; ++u) m = fmaxf(m, fmaxf(fmaxf(q[u].x, q[u].y), fmaxf(q[u].z, q[u].w)));
    if (m >= tau_f) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v0 = base + (u * SS_THREADS + tid) * 4;
        admit(q[u].x, v0);
        admit(q[u].y, v0 + 1);
        admit(q[u].z, v0 + 2);
        admit(q[u].w, v0 + 3);
      }
    }
  };
  // Two attempts.  The first aims the threshold at population rank ~2k: ~2k admitted scores fit
  // the one-key-per-thread rank count (two barriers) instead of a 512-key bitonic sort (45), and
  // the search peels half as many keys.  It admits fewer than k scores on a few percent of the
  // rows; those run once more at rank ~6k, reading the row from L2.  Only a row that fails both
  // (or overflows the buffer) is left to the radix-select redo.
  int c = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int R = attempt == 0 ? rounds_tight : rounds_loose;
    // ---- 1. threshold = (about) the sample's R-th best distinct key (ordered_desc: smaller = better)
    uint32_t key[SS_V4 * 4];
    {
      const long V4 = V >> 2;
      constexpr int NV = SS_THREADS * SS_V4;
#pragma unroll
      for (int q = 0; q < SS_V4; ++q) {
        const long j4 = (long)(tid + SS_THREADS * q) * V4 / NV;
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(z) + j4);
        key[4 * q + 0] = ordered_desc(v4.x);
        key[4 * q + 1] = ordered_desc(v4.y);
        key[4 * q + 2] = ordered_desc(v4.z);
        key[4 * q + 3] = ordered_desc(v4.w);
      }
    }
    if (tid == 0) s_count = 0;
    // the first batch of the streaming pass is requested now: its DRAM latency runs under the
    // threshold search instead of after it
    float4 qa[4], qb[4];
    if (n_full > 0) load_full(0, qa);
    for (int r = 0; r < R; ++r) {
      uint32_t m = 0xFFFFFFFFu;
#pragma unroll
      for (int q = 0; q < SS_V4 * 4; ++q) m = min(m, key[q]);
      m = __reduce_min_sync(0xffffffffu, m);
      if (lane == 0) s_cand[warp * R + r] = m;
#pragma unroll
      for (int q = 0; q < SS_V4 * 4; ++q)
        if (key[q] == m) key[q] = 0xFFFFFFFFu;
    }
    __syncthreads();
    if (warp == 0) {
      constexpr int PER_LANE = NW * SS_RMAX / 32;
      uint32_t cd[PER_LANE];
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) {
        const int idx = lane + 32 * i;
        cd[i] = idx < NW * R ? s_cand[idx] : 0xFFFFFFFFu;
      }
      uint32_t tau = 0xFFFFFFFFu;
      for (int r = 0; r < R; ++r) {
        uint32_t m = 0xFFFFFFFFu;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) m = min(m, cd[i]);
        m = __reduce_min_sync(0xffffffffu, m);
        if (m == 0xFFFFFFFFu) break;   // fewer than R distinct sample keys (e.g. an all-ties row)
        tau = m;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i)
          if (cd[i] == m) cd[i] = 0xFFFFFFFFu;
      }
      if (lane == 0) s_tau = tau;
    }
    __syncthreads();
    tau_f = from_ordered_desc(s_tau);
    // ---- 2. one pass: admit every score >= tau
    int bi = 0;
    for (; bi + 1 < n_full; bi += 2) {
      load_full(bi + 1, qb);
      scan16(qa, bi * BATCH);
      if (bi + 2 < n_full) load_full(bi + 2, qa);
      scan16(qb, (bi + 1) * BATCH);
    }
    if (bi < n_full) scan16(qa, bi * BATCH);   // odd number of full batches: the last sits in qa
    for (int v = n_full * BATCH + tid; v < V; v += SS_THREADS) admit(__ldg(z + v), v);
    __syncthreads();
    c = s_count;
    if (c >= k || rounds_loose <= rounds_tight) break;   // enough (or too many: decided below)
    __syncthreads();                                      // s_count is reset by the next attempt
  }
  if (c > SS_CAP || c < k) {   // threshold too loose / too tight: exact redo by the radix kernel
    if (tid == 0) out_ids[row * ld_out] = TK_REDO;
    return;
  }
  // ---- 3. exact top-k of the admitted candidates (sorted into sm_buf[0..k)).  Up to 256: rank by
  // counting, one key per thread.  More (second attempts, k > 128): counting ranks of 400 keys is
  // 160 K 64-bit comparisons - more issue slots than the whole streaming pass (0.40 of the HBM
  // roofline at k = 100 when every row did that) - so those take a bitonic sort, 45 barrier steps
  // but a fifth of the instructions (0.53).
  {
    unsigned long long* buf = sm_buf;
    if (c <= SS_THREADS) {   // rank by counting, one key per thread: c^2 comparisons
      unsigned long long* dst = buf + SS_CAP;
      const unsigned long long my = tid < c ? buf[tid] : ~0ull;
      int rk = 0;
#pragma unroll 4
      for (int i = 0; i < c; ++i) rk += buf[i] < my ? 1 : 0;
      __syncthreads();
      if (tid < c && rk < k) dst[rk] = my;
      __syncthreads();
      for (int i = tid; i < k; i += SS_THREADS) buf[i] = dst[i];
    } else {
      int logn = 1;
      while ((1 << logn) < c) ++logn;
      for (int i = c + tid; i < (1 << logn); i += SS_THREADS) buf[i] = ~0ull;
      __syncthreads();
      block_bitonic_sort(buf, logn);
    }
    __syncthreads();
  }
  const unsigned long long idmask = (1ull << idbits) - 1ull;
  for (int r = tid; r < k; r += SS_THREADS) {
    const unsigned long long K = sm_buf[r];
    out_ids[row * ld_out + r] = (int32_t)(K & idmask);
    if (out_scores) out_scores[row * ld_out + r] = from_ordered_desc((uint32_t)(K >> idbits));
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
  if (V >= ST_MIN_V && ld % 4 == 0 && ((uintptr_t)scores & 15) == 0) {
    // sampled static threshold (measured faster: 0.34 vs 0.16 of the HBM roofline at V = 54K,
    // 0.77 vs 0.66 at V = 1M, k = 100) except for very small k on very long rows, where the
    // running threshold of topk_stream_kernel admits almost nothing (0.79 vs 0.77 at k = 10)
    const char* mode = getenv("B4CP_TOPK_STREAM");
    bool sampled = !(V >= 262144 && k <= 16);
    if (mode && mode[0] == 'a') sampled = false;
    if (mode && mode[0] == 's') sampled = true;
    if (sampled) {
      const bool big = V >= 131072;
      const int ns = SM_THREADS * (big ? 16 : 4);
      int r = (int)((4L * k * ns + V - 1) / V);
      r = std::max(8, std::min(r, 64));
      const size_t smem = (size_t)2 * SM_CAP * sizeof(unsigned long long);
      static const bool small_ok = !(getenv("B4CP_TOPK_SMALL") && getenv("B4CP_TOPK_SMALL")[0] == '0');
      if (!big && small_ok) {
        // four rows per SM, warp-local threshold search, two attempts (ranks ~2k, then ~6k)
        static_assert(SS_NS == SM_THREADS * 4, "the small kernel samples as many scores");
        const int r_tight = std::max(8, std::min((int)((2L * k * ns + V - 1) / V), SS_RMAX));
        const int r_loose = std::max(r_tight, std::min((int)((6L * k * ns + V - 1) / V), SS_RMAX));
        const size_t smem_s = (size_t)2 * SS_CAP * sizeof(unsigned long long);
        topk_sample_small_kernel<<<(unsigned)rows, SS_THREADS, smem_s, (cudaStream_t)stream>>>(
            scores, ld, V, k, idbits, r_tight, r_loose, out_ids, out_scores, ld_out);
      } else if (big) {
        B4CP_CUDA(cudaFuncSetAttribute(topk_sample_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        topk_sample_kernel<16><<<(unsigned)rows, SM_THREADS, smem, (cudaStream_t)stream>>>(
            scores, ld, V, k, idbits, r, out_ids, out_scores, ld_out);
      } else {
        B4CP_CUDA(cudaFuncSetAttribute(topk_sample_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        topk_sample_kernel<4><<<(unsigned)rows, SM_THREADS, smem, (cudaStream_t)stream>>>(
            scores, ld, V, k, idbits, r, out_ids, out_scores, ld_out);
      }
    } else {
      const size_t smem = (size_t)2 * ST_CAP * sizeof(unsigned long long);
      B4CP_CUDA(cudaFuncSetAttribute(topk_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
      topk_stream_kernel<<<(unsigned)rows, ST_THREADS, smem, (cudaStream_t)stream>>>(
          scores, ld, V, k, idbits, out_ids, out_scores, ld_out);
    }
    const unsigned redo_grid = (unsigned)std::min<long>(ceil_div(rows, (long)TK_REDO_CHUNK), 2 * 148L);
    topk_redo_kernel<<<redo_grid, TK_THREADS, 0, (cudaStream_t)stream>>>(
        scores, ld, rows, V, k, idbits, out_ids, out_scores, ld_out);
    note_launches(2);
    B4CP_LAUNCH_CHECK();
    return 0;
  }
  topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, (cudaStream_t)stream>>>(
      scores, nullptr, ld, V, k, idbits, out_ids, out_scores, ld_out, 0);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

/* b4cp_topk_candidates for the rows marked for a redo only (out_ids[row][0] == -2): the fallback
 * leg of the long-vocabulary path of b4cp_score_topk; other rows keep their results. */
extern "C" int b4cp_topk_candidates_redo(const float* cand_scores, const int32_t* cand_ids, long ld,
                                         long rows, int n_cand, int V, int k, int32_t* out_ids,
                                         float* out_scores, long ld_out, void* stream) {
  B4CP_CHECK_ARG(k >= 1 && k <= TK_MAXK, "topk: k=%d must be in [1,%d]", k, TK_MAXK);
  B4CP_CHECK_ARG(cand_scores && cand_ids && out_ids, "topk_candidates: null argument");
  if (rows == 0) return 0;
  int idbits = 1;
  while ((1L << idbits) < V) ++idbits;
  topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, (cudaStream_t)stream>>>(
      cand_scores, cand_ids, ld, n_cand, k, idbits, out_ids, out_scores, ld_out, 1);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}

/* b4cp_topk_candidates over COUNTED lists: row r holds base_count + extra_count[r] entries (capped
 * at n_cand); slots past that are never read, so they need not be initialised. */
extern "C" int b4cp_topk_candidates_counted(const float* cand_scores, const int32_t* cand_ids, long ld,
                                            long rows, int n_cand, const int* extra_count,
                                            int base_count, int V, int k, int32_t* out_ids,
                                            float* out_scores, long ld_out, void* stream) {
  B4CP_CHECK_ARG(k >= 1 && k <= TK_MAXK, "topk: k=%d must be in [1,%d]", k, TK_MAXK);
  B4CP_CHECK_ARG(cand_scores && cand_ids && out_ids && extra_count, "topk_candidates: null argument");
  B4CP_CHECK_ARG(ld_out >= k, "topk: ld_out < k");
  if (rows == 0) return 0;
  int idbits = 1;
  while ((1L << idbits) < V) ++idbits;
  topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, (cudaStream_t)stream>>>(
      cand_scores, cand_ids, ld, n_cand, k, idbits, out_ids, out_scores, ld_out, 0, extra_count, base_count);
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
      cand_scores, cand_ids, ld, n_cand, k, idbits, out_ids, out_scores, ld_out, 0);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}
