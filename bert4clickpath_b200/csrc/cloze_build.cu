// On-device Cloze batch builder: the step immediately before the hot path.
// Restates examples/BERT4Rec/source/input_pipeline.py:21-32, :59-133 (mask selection and label
// extraction), :198-214 (padding) and clickstream_transformer.py:38-63 (chaining) on int32 ids
// held in HBM as one CSR of sessions, so a training step needs no host work beyond choosing
// which sessions form the batch:
//   TRAIN: drop the last item; n = clip(int(float32(len) * float32(p)), 0, max_masked) distinct positions become
//          [MASK]; labels (label-vocabulary id = input id - label_offset, as float32) follow in
//          ascending position; EVAL: only the last position is masked.
//   ids row = [CLS] [SEP] items... [PAD]... [SEP]  (the sequence is padded BEFORE chaining),
//   labels row padded with label_pad.
// The reference draws the positions with tf.random.shuffle(range)[:n]; its RNG stream cannot be
// matched, so the n positions with the smallest 64-bit keys key(seed, session, pos) are taken
// (ties by position) - a uniformly random n-subset, and a pure function of its arguments that
// oracle/clickpath_oracle.py restates for bit-exact parity.
// HBM-bound integer work: one CTA per session, keys ranked by counting in shared memory
// (L <= CB_MAX_LEN positions, L^2 / 128 compares per thread); traffic = the session's items in,
// one ids row and one labels row out.
#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

constexpr int CB_THREADS = 128;
constexpr int CB_MAX_LEN = 2048;

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct ClozeBuildParams {
  const int32_t* items;
  const long long* offsets;
  const int32_t* session_idx;
  int L, Mmax, train, max_masked;
  double masked_percentage;
  unsigned long long seed;
  int cls_id, sep_id, mask_id, pad_id, label_offset;
  float label_pad;
  int32_t* ids;
  float* labels;
  int32_t* n_masked;
  int32_t* status;
};

__global__ void __launch_bounds__(CB_THREADS) cloze_build_kernel(const ClozeBuildParams p) {
  extern __shared__ unsigned long long s_key[];            // [L]
  unsigned char* s_flag = (unsigned char*)(s_key + p.L);   // [L]
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const long long sess = p.session_idx[b];
  const long long beg = p.offsets[sess];
  int len = (int)(p.offsets[sess + 1] - beg);
  if (p.train) len -= 1;                                   // input_pipeline.py:101-104
  // :68-70 / :118-121 - the reference multiplies in float32 (tf.cast(size, tf.float32) * p) and
  // truncates; float32 and float64 disagree for some (len, p), e.g. 90 * 0.7 -> 63 vs 62
  int n = p.train ? (int)__fmul_rn((float)len, (float)p.masked_percentage) : 1;
  if (p.train) n = max(0, min(n, p.max_masked));
  const int S = p.L + 3;
  int32_t* row = p.ids + (long)b * S;
  float* lab = p.labels + (long)b * p.Mmax;
  if (len < (p.train ? 0 : 1) || len > p.L || n > p.Mmax) {   // caller sized the batch wrongly
    if (tid == 0) atomicExch(p.status, b + 1);
    for (int i = tid; i < S; i += CB_THREADS) row[i] = p.pad_id;
    for (int i = tid; i < p.Mmax; i += CB_THREADS) lab[i] = p.label_pad;
    return;
  }
  const unsigned long long base = splitmix64(p.seed + (unsigned long long)sess);
  for (int i = tid; i < len; i += CB_THREADS) s_key[i] = splitmix64(base + (unsigned long long)i);
  __syncthreads();
  for (int i = tid; i < len; i += CB_THREADS) {
    bool masked;
    if (p.train) {
      const unsigned long long ki = s_key[i];
      int rank = 0;
      for (int j = 0; j < len; ++j) {
        const unsigned long long kj = s_key[j];
        rank += (kj < ki) || (kj == ki && j < i);
      }
      masked = rank < n;
    } else {
      masked = (i == len - 1);
    }
    s_flag[i] = masked ? 1 : 0;
  }
  __syncthreads();
  if (tid == 0) {
    row[0] = p.cls_id;
    row[1] = p.sep_id;
    row[S - 1] = p.sep_id;
    atomicAdd(p.n_masked, n);
  }
  for (int i = tid; i < p.L; i += CB_THREADS) {
    int v = p.pad_id;
    if (i < len) {
      v = p.items[beg + i];
      if (s_flag[i]) {
        int slot = 0;
        for (int j = 0; j < i; ++j) slot += s_flag[j];
        lab[slot] = (float)(v - p.label_offset);
        v = p.mask_id;
      }
    }
    row[2 + i] = v;
  }
  for (int i = n + tid; i < p.Mmax; i += CB_THREADS) lab[i] = p.label_pad;
}

}  // namespace b4cp

using namespace b4cp;

extern "C" unsigned long long b4cp_cloze_position_key(unsigned long long seed,
                                                      unsigned long long session,
                                                      unsigned long long pos) {
  return splitmix64(splitmix64(seed + session) + pos);
}

extern "C" int b4cp_cloze_build(const int32_t* items, const long long* offsets,
                                const int32_t* session_idx, int B, int L, int Mmax, int train,
                                double masked_percentage, int max_masked, unsigned long long seed,
                                int cls_id, int sep_id, int mask_id, int pad_id, int label_offset,
                                float label_pad, int32_t* ids, float* labels, int32_t* n_masked,
                                int32_t* status, void* stream) {
  B4CP_CHECK_ARG(items && offsets && session_idx && ids && labels && n_masked && status,
                 "cloze_build: null argument");
  B4CP_CHECK_ARG(B >= 0 && L >= 1 && L <= CB_MAX_LEN && Mmax >= 1,
                 "cloze_build: B=%d L=%d (max %d) Mmax=%d", B, L, CB_MAX_LEN, Mmax);
  if (B == 0) return 0;
  ClozeBuildParams p;
  p.items = items; p.offsets = offsets; p.session_idx = session_idx;
  p.L = L; p.Mmax = Mmax; p.train = train ? 1 : 0; p.max_masked = max_masked;
  p.masked_percentage = masked_percentage; p.seed = seed;
  p.cls_id = cls_id; p.sep_id = sep_id; p.mask_id = mask_id;
  p.pad_id = pad_id; p.label_offset = label_offset;
  p.label_pad = label_pad; p.ids = ids; p.labels = labels; p.n_masked = n_masked; p.status = status;
  const size_t smem = (size_t)L * 9;
  cloze_build_kernel<<<B, CB_THREADS, smem, (cudaStream_t)stream>>>(p);
  note_launches(1);
  B4CP_LAUNCH_CHECK();
  return 0;
}
