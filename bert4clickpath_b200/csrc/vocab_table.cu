// Host-side vocabulary lookup: tf.lookup.StaticVocabularyTable(KeyValueTensorInitializer(keys,
// range(len(keys))), num_oov_buckets=1) as the reference builds it
// (clickstream_transformer/clickstream_transformer.py:247-258): key j -> j (the first occurrence
// of a duplicated key wins, as with the Python dict it replaces), any other string -> len(keys).
//
// The reference feeds STRING tensors to its model and the lookup runs inside the TensorFlow
// graph; here it is the one per-token host step between a caller's string batch and the device.
// A Python dict costs ~150 ns per token - 30 ms for one C1 batch of 4,096 x 52 tokens, ten times
// the training step it feeds.  This is an open-addressing table over UCS4 code points (NumPy's
// '<U' layout: fixed width, NUL padded) probed by a few host threads: ~1 ms for the same batch.
// No device code in this file; it needs no GPU.
#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"
#include "../../include/b4cp.h"

namespace b4cp {

struct Slot {
  int32_t key;     // key index, -1 = empty
  uint32_t hash;   // full hash of that key: a probe touches the key row only on a hash match
};

struct VocabTable {
  std::vector<uint32_t> rows;   // keys as fixed-width rows of `width` code points, NUL padded
  std::vector<Slot> slots;      // open addressing, linear probing, load <= 1/2
  uint32_t mask = 0;
  int32_t n_keys = 0;
  int width = 1;
};

static inline int token_len(const uint32_t* s, int width) {
  int n = width;
  while (n > 0 && s[n - 1] == 0) --n;   // NumPy pads with NULs (and cannot hold trailing NULs)
  return n;
}

static inline uint32_t hash_cp(const uint32_t* s, int n) {
  uint32_t h = 2166136261u;              // FNV-1a over the code points, then a finaliser
  for (int i = 0; i < n; ++i) {
    h ^= s[i];
    h *= 16777619u;
  }
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

static inline bool same_key(const VocabTable& t, int32_t j, const uint32_t* s, int n) {
  if (n > t.width) return false;
  const uint32_t* k = t.rows.data() + (size_t)j * t.width;
  for (int i = 0; i < n; ++i)
    if (k[i] != s[i]) return false;
  return n == t.width || k[n] == 0;
}

static inline int32_t find(const VocabTable& t, const uint32_t* s, int n) {
  const uint32_t h = hash_cp(s, n);
  uint32_t p = h & t.mask;
  while (true) {
    const Slot sl = t.slots[p];
    if (sl.key < 0) return t.n_keys;     // the single OOV bucket
    if (sl.hash == h && same_key(t, sl.key, s, n)) return sl.key;
    p = (p + 1) & t.mask;
  }
}

}  // namespace b4cp

using namespace b4cp;

extern "C" void* b4cp_vocab_table_create(const uint32_t* keys_ucs4, long n_keys, int width) {
  if (n_keys < 0 || width < 1 || (n_keys > 0 && !keys_ucs4) || n_keys >= (1L << 30)) {
    set_last_error("vocab_table_create: bad arguments (n_keys=%ld width=%d)", n_keys, width);
    return nullptr;
  }
  VocabTable* t = new VocabTable();
  t->n_keys = (int32_t)n_keys;
  t->width = width;
  t->rows.assign(keys_ucs4, keys_ucs4 + (size_t)n_keys * width);
  uint32_t cap = 16;
  while (cap < 2 * (uint64_t)n_keys + 2) cap <<= 1;
  t->mask = cap - 1;
  t->slots.assign(cap, Slot{-1, 0});
  for (long j = 0; j < n_keys; ++j) {
    const uint32_t* s = t->rows.data() + (size_t)j * width;
    const int n = token_len(s, width);
    const uint32_t h = hash_cp(s, n);
    uint32_t p = h & t->mask;
    bool dup = false;
    while (t->slots[p].key >= 0) {
      if (t->slots[p].hash == h && same_key(*t, t->slots[p].key, s, n)) {
        dup = true;                           // duplicate key: the first occurrence keeps it
        break;
      }
      p = (p + 1) & t->mask;
    }
    if (!dup) t->slots[p] = Slot{(int32_t)j, h};
  }
  return t;
}

extern "C" void b4cp_vocab_table_destroy(void* table) { delete static_cast<VocabTable*>(table); }

extern "C" long b4cp_vocab_table_size(const void* table) {
  return table ? (long)static_cast<const VocabTable*>(table)->n_keys + 1 : -1;
}

extern "C" int b4cp_vocab_table_lookup(const void* table, const uint32_t* tokens_ucs4, long n_tokens,
                                       int width, int32_t* out_ids, int n_threads) {
  B4CP_CHECK_ARG(table != nullptr, "vocab_table_lookup: null table");
  B4CP_CHECK_ARG(n_tokens >= 0 && width >= 1, "vocab_table_lookup: bad shape (%ld x %d)", n_tokens, width);
  if (n_tokens == 0) return 0;
  B4CP_CHECK_ARG(tokens_ucs4 && out_ids, "vocab_table_lookup: null buffer");
  const VocabTable& t = *static_cast<const VocabTable*>(table);
  auto work = [&](long a, long b) {
    for (long i = a; i < b; ++i) {
      const uint32_t* s = tokens_ucs4 + i * width;
      out_ids[i] = find(t, s, token_len(s, width));
    }
  };
  const long min_per_thread = 8192;
  int nt = (int)std::max(1L, std::min<long>(std::max(1, n_threads), n_tokens / min_per_thread));
  if (nt <= 1) {
    work(0, n_tokens);
    return 0;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  const long per = (n_tokens + nt - 1) / nt;
  for (int k = 1; k < nt; ++k) pool.emplace_back(work, std::min(n_tokens, k * per), std::min(n_tokens, (k + 1) * per));
  work(0, std::min(n_tokens, per));
  for (auto& th : pool) th.join();
  return 0;
}
