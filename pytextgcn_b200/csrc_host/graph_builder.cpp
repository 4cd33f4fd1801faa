// Host-side word-word PMI edge builder (C ABI, no CUDA): the native replacement for the reference's
// Cython `compute_word_word_edges` (textgcn/lib/clib/graphbuilder.pyx:23-66) behind
// Text2GraphTransformer.fit_transform (textgcn/lib/text2graph.py:156-160).
//
// Same results, different algorithm:
//  * the reference walks every window and every position pair inside it (O(L * w^2) increments per
//    document) into a dense packed V(V+1)/2 uint32 array (4*V^2 bytes with the PMI field; 32-bit index
//    overflow for V >= 65,536 -- graphbuilder.pyx:44,134,224,250);
//  * here every position pair (a <= b, b - a < w) of a document is visited ONCE and weighted by the
//    number of windows that contain both positions (closed form below), O(L * w) per document, and
//    counts live in open-addressing hash tables keyed by the 64-bit pair id, one table per thread
//    (huge-page backed, key and counter in one slot, the probes of a position issued as a prefetched batch).
//    The tables are merged in parallel over KEY RANGES (splitters from a sample of the keys, so hub words do not
//    unbalance them): every thread deals its table into the ranges, each range is sorted and reduced on its own, the
//    PMI pass runs per range, and the ranges -- ordered by key -- are the edge list in the reference's order.
//    Memory is O(#distinct co-occurring pairs); indices are 64-bit.
// Window rule restated from graphbuilder.pyx:94-113: window starts j = 0 .. seq_len - w; a window is
// skipped (and all later ones) as soon as its LAST slot is padding, except j = 0 which is always
// taken; inside a window pairs (k <= l) are counted while both tokens are not padding.  With `len`
// real tokens (padding only at the tail) the taken windows are j = 0 .. jmax, jmax = max(0, len - w),
// and the pair of positions (a <= b) lies in the windows max(0, b-w+1) <= j <= min(a, jmax).
// PMI arithmetic is done in the reference's types: float divisions, double libc log, float result
// (graphbuilder.pyx:146-147,156-164), threshold 1e-10f (graphbuilder.pyx:20).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <new>
#include <thread>
#include <vector>

#include <sys/mman.h>

#include "../../include/textgcn_host.h"

namespace {

// The pair table is far larger than the caches and probed at random: with 4 KB pages every probe is also a TLB miss.
// Anonymous mapping + MADV_HUGEPAGE (transparent huge pages in "madvise" mode) removes most of them; where huge pages are
// unavailable the advice is ignored and the mapping behaves like malloc'ed memory.
template <typename T>
struct HugeArray {
  T* p = nullptr; size_t n = 0;
  HugeArray() = default;
  HugeArray(const HugeArray&) = delete;
  HugeArray& operator=(const HugeArray&) = delete;
  ~HugeArray() { release(); }
  void allocate(size_t count) {
    release();
    const size_t bytes = ((count * sizeof(T) + (2u << 20) - 1) / (2u << 20)) * (2u << 20);
    void* q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (q == MAP_FAILED) throw std::bad_alloc();
#ifdef MADV_HUGEPAGE
    madvise(q, bytes, MADV_HUGEPAGE);
#endif
    p = static_cast<T*>(q); n = count;
  }
  void release() {
    if (p) munmap(p, ((n * sizeof(T) + (2u << 20) - 1) / (2u << 20)) * (2u << 20));
    p = nullptr; n = 0;
  }
  void swap(HugeArray& o) { std::swap(p, o.p); std::swap(n, o.n); }
  T* begin() const { return p; }
  T* end() const { return p + n; }
  size_t size() const { return n; }
  T& operator[](size_t i) const { return p[i]; }
};

struct PairTable {              // open addressing, linear probing, key = i * V + j (i <= j), 0xFFFF.. = empty
  struct Slot { uint64_t key; uint32_t val; uint32_t pad; };       // key and counter share a cache line: one miss per probe
  HugeArray<Slot> slots;
  uint64_t mask = 0, used = 0;
  static constexpr uint64_t EMPTY = ~0ull;
  explicit PairTable(uint64_t cap_pow2 = 1u << 16) { reset(cap_pow2); }
  void reset(uint64_t cap) {
    slots.allocate(cap);
    for (Slot& sl : slots) sl = Slot{EMPTY, 0, 0};
    mask = cap - 1; used = 0;
  }
  static inline uint64_t hash(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
  void grow() {
    HugeArray<Slot> old;
    old.swap(slots);
    reset((mask + 1) * 2);
    for (const Slot& sl : old) if (sl.key != EMPTY) add(sl.key, sl.val);
  }
  inline void prefetch(uint64_t k) const { __builtin_prefetch(&slots[hash(k) & mask], 1, 1); }
  inline void add(uint64_t k, uint32_t c) {
    if ((used + 1) * 10 > (mask + 1) * 7) grow();
    uint64_t h = hash(k) & mask;
    while (true) {
      Slot& sl = slots[h];
      if (sl.key == k) { sl.val += c; return; }             // uint32 wrap-around like the reference's counters
      if (sl.key == EMPTY) { sl.key = k; sl.val = c; ++used; return; }
      h = (h + 1) & mask;
    }
  }
  void release() { slots.release(); }
};

struct Result {                 // the edge list as the key ranges produced it (concatenated by tgcn_ww_fetch)
  std::vector<std::vector<int32_t>> coo;     // per range [n_edges_r][2]
  std::vector<std::vector<float>> w;         // per range [n_edges_r]
  uint64_t n_windows = 0;
};

constexpr int PF_MAX = 64;

void count_docs(const int32_t* X, int64_t d0, int64_t d1, int64_t seq_len, int64_t w, uint64_t V, PairTable& tab,
                uint64_t& n_windows, int& bad) {
  for (int64_t d = d0; d < d1; ++d) {
    const int32_t* x = X + d * seq_len;
    int64_t len = 0;
    while (len < seq_len && x[len] != -1) ++len;            // padding only at the tail (text2graph.py:40-44)
    const int64_t wcap = std::min(w, seq_len);
    const int64_t jmax = std::max<int64_t>(0, len - wcap);
    n_windows += (uint64_t)(jmax + 1);                       // window 0 always counts, even for an empty document
    uint64_t keys[PF_MAX]; uint32_t cnts[PF_MAX];
    for (int64_t a = 0; a < len; ++a) {
      const int64_t xa = x[a];
      if (xa < 0 || (uint64_t)xa >= V) { bad = 1; continue; }
      const int64_t bend = std::min(len, a + wcap);
      // the pairs of position a, in two passes: keys + prefetch of their slots, then the increments (the table is far
      // larger than the caches, so the probes are cache misses; issued together they overlap)
      for (int64_t b0 = a; b0 < bend; b0 += PF_MAX) {
        const int64_t b1 = std::min(bend, b0 + PF_MAX);
        int m = 0;
        for (int64_t b = b0; b < b1; ++b) {
          const int64_t xb = x[b];
          if (xb < 0 || (uint64_t)xb >= V) { bad = 1; continue; }
          const int64_t jlo = std::max<int64_t>(0, b - wcap + 1), jhi = std::min(a, jmax);
          if (jhi < jlo) continue;
          const uint64_t lo = (uint64_t)std::min(xa, xb), hi = (uint64_t)std::max(xa, xb);
          keys[m] = lo * V + hi; cnts[m] = (uint32_t)(jhi - jlo + 1);
          tab.prefetch(keys[m]);
          ++m;
        }
        for (int q = 0; q < m; ++q) tab.add(keys[q], cnts[q]);
      }
    }
  }
}

}  // namespace

extern "C" {

// Builds the edge list; returns an opaque handle (NULL on bad input), *n_edges_out = number of DIRECTED edges.
void* tgcn_ww_build(const int32_t* X, int64_t n_docs, int64_t seq_len, int64_t n_vocab, int64_t window_size,
                    int32_t n_threads, int64_t* n_edges_out, uint64_t* n_windows_out) try {
  if (!X || n_docs < 0 || seq_len <= 0 || n_vocab <= 0 || window_size <= 0 || !n_edges_out) return nullptr;
  const uint64_t V = (uint64_t)n_vocab;
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  T = (int)std::max<int64_t>(1, std::min<int64_t>(T, std::max<int64_t>(1, n_docs / 64)));
  const bool timing = std::getenv("TGCN_HOST_TIMING") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  std::vector<PairTable> tabs((size_t)T);     // (PairTable is not copyable: sized once, never resized)
  std::vector<uint64_t> nwin((size_t)T, 0);
  std::vector<int> bad((size_t)T, 0);
  std::vector<std::vector<uint64_t>> samples((size_t)T);
  std::atomic<bool> failed(false);
  std::vector<std::thread> th;
  const int64_t per = (n_docs + T - 1) / T;
  for (int t = 0; t < T; ++t) {
    const int64_t d0 = std::min<int64_t>(n_docs, t * per), d1 = std::min<int64_t>(n_docs, d0 + per);
    th.emplace_back([&, t, d0, d1]() {
      try {
        count_docs(X, d0, d1, seq_len, window_size, V, tabs[t], nwin[t], bad[t]);
        size_t seen = 0;                                      // sample of this table's keys for the range splitters below
        for (const PairTable::Slot& sl : tabs[t].slots)
          if (sl.key != PairTable::EMPTY && (seen++ % 61) == 0) samples[t].push_back(sl.key);
      } catch (...) { failed = true; }                        // out of memory: reported as NULL, never thrown across the C ABI
    });
  }
  for (auto& t : th) t.join();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[tgcn_ww_build] %s: %.3f s\n", what, std::chrono::duration<double>(now - t_start).count());
    t_start = now;
  };
  lap("count pairs");
  if (failed) return nullptr;
  for (int t = 0; t < T; ++t) if (bad[t]) return nullptr;    // token id outside [0, n_vocab)
  uint64_t n_windows = 0;
  for (uint64_t v : nwin) n_windows += v;
  // ---- merge, in parallel over KEY RANGES ----
  // Splitters from a sample of the keys (balanced even when a few hub words own most pairs); every thread deals
  // the entries of ITS table into the ranges, then range r gathers its pieces from all tables, sorts them and adds
  // equal keys.  Ranges are ordered by key, so the concatenation is the reference's upper-triangle row-major order.
  typedef std::pair<uint64_t, uint32_t> KV;
  const int R = T == 1 ? 1 : T * 4;
  std::vector<uint64_t> split;                                // R - 1 ascending splitters; range r = [split[r-1], split[r])
  if (R > 1) {
    std::vector<uint64_t> sample;
    for (auto& sm : samples) sample.insert(sample.end(), sm.begin(), sm.end());
    std::sort(sample.begin(), sample.end());
    for (int r = 1; r < R && !sample.empty(); ++r) split.push_back(sample[sample.size() * (size_t)r / (size_t)R]);
    split.erase(std::unique(split.begin(), split.end()), split.end());
  }
  lap("splitters");
  const int NR = (int)split.size() + 1;
  std::vector<std::vector<KV>> piece((size_t)T * NR);          // piece[t * NR + r]
  auto run_parallel = [&](int n_tasks, const std::function<void(int)>& fn) {
    std::atomic<int> next(0);
    std::vector<std::thread> ws;
    const int W = std::min(T, n_tasks);
    for (int k = 0; k < W; ++k) ws.emplace_back([&]() {
      try { for (int q; (q = next.fetch_add(1)) < n_tasks;) fn(q); } catch (...) { failed = true; }
    });
    for (auto& x : ws) x.join();
  };
  run_parallel(T, [&](int t) {
    PairTable& tb = tabs[t];
    for (const PairTable::Slot& sl : tb.slots) {
      if (sl.key == PairTable::EMPTY) continue;
      const int r = (int)(std::upper_bound(split.begin(), split.end(), sl.key) - split.begin());
      piece[(size_t)t * NR + r].emplace_back(sl.key, sl.val);
    }
    tb.release();
  });
  lap("deal into ranges");
  // p_i = c_ii / n_windows   (graphbuilder.pyx:146-147); a diagonal key lives in exactly one range: no write conflicts
  const uint32_t nw32 = (uint32_t)n_windows;                  // the reference keeps n_windows in an unsigned int
  std::vector<float> p((size_t)V, 0.0f);
  std::vector<std::vector<KV>> merged((size_t)NR);
  run_parallel(NR, [&](int r) {
    std::vector<KV>& all = merged[r];
    size_t total = 0;
    for (int t = 0; t < T; ++t) total += piece[(size_t)t * NR + r].size();
    all.reserve(total);
    for (int t = 0; t < T; ++t) {
      std::vector<KV>& pc = piece[(size_t)t * NR + r];
      all.insert(all.end(), pc.begin(), pc.end());
      std::vector<KV>().swap(pc);
    }
    std::sort(all.begin(), all.end(), [](const KV& a, const KV& b) { return a.first < b.first; });
    size_t m = 0;
    for (size_t q = 0; q < all.size(); ++q) {
      if (m > 0 && all[m - 1].first == all[q].first) all[m - 1].second += all[q].second;   // uint32 wrap like the reference
      else all[m++] = all[q];
    }
    all.resize(m);
    for (const KV& kv : all) {
      const uint64_t i = kv.first / V, j = kv.first % V;
      if (i == j) p[i] = (float)kv.second / (float)nw32;
    }
  });
  lap("sort + merge ranges");
  // PMI per range (needs every p_i: second pass), then the ranges are concatenated in key order
  const float EPSILON = 1e-10f;
  auto* res = new Result();
  res->n_windows = n_windows;
  res->coo.resize((size_t)NR);
  res->w.resize((size_t)NR);
  run_parallel(NR, [&](int r) {
    std::vector<int32_t> co;                                  // locals, moved into the result at the end: the vector
    std::vector<float> wv;                                    // headers of neighbouring ranges would share cache lines
    co.reserve(4 * merged[r].size()); wv.reserve(2 * merged[r].size());      // upper bound: every pair becomes two edges
    for (const KV& kv : merged[r]) {                          // sorted by i, then j: upper-triangle row-major order
      const uint64_t i = kv.first / V, j = kv.first % V;
      if (i == j) continue;
      const float p_ij = (float)kv.second / (float)nw32;
      if (p_ij == 0 || p[i] == 0 || p[j] == 0) continue;      // graphbuilder.pyx:157-160
      const float pmi = (float)std::log((double)(p_ij / (p[i] * p[j])));
      if (pmi > EPSILON) {
        co.push_back((int32_t)i); co.push_back((int32_t)j); wv.push_back(pmi);
        co.push_back((int32_t)j); co.push_back((int32_t)i); wv.push_back(pmi);
      }
    }
    res->coo[r] = std::move(co);
    res->w[r] = std::move(wv);
  });
  lap("pmi");
  if (failed) { delete res; return nullptr; }
  size_t n_e = 0;
  for (int r = 0; r < NR; ++r) n_e += res->w[r].size();
  *n_edges_out = (int64_t)n_e;
  if (n_windows_out) *n_windows_out = n_windows;
  return res;
} catch (...) {                                               // allocation failure on the calling thread
  return nullptr;
}

int tgcn_ww_fetch(void* handle, int32_t* coo_out, float* w_out) {
  if (!handle || !coo_out || !w_out) return 1;
  auto* r = static_cast<Result*>(handle);
  const int NR = (int)r->w.size();
  std::vector<size_t> off((size_t)NR + 1, 0);
  for (int q = 0; q < NR; ++q) off[q + 1] = off[q] + r->w[q].size();
  std::atomic<int> next(0);                                  // the ranges are copied by a few threads (first touch of the
  std::vector<std::thread> ws;                               // caller's arrays is the cost, not the copy)
  const int W = std::max(1, std::min(NR, (int)std::thread::hardware_concurrency()));
  for (int k = 0; k < W; ++k) ws.emplace_back([&]() {
    for (int q; (q = next.fetch_add(1)) < NR;) {
      if (r->w[q].empty()) continue;
      std::memcpy(coo_out + 2 * off[q], r->coo[q].data(), r->coo[q].size() * sizeof(int32_t));
      std::memcpy(w_out + off[q], r->w[q].data(), r->w[q].size() * sizeof(float));
    }
  });
  for (auto& x : ws) x.join();
  return 0;
}

void tgcn_ww_free(void* handle) { delete static_cast<Result*>(handle); }

// c_ij in the reference's packed upper-triangular layout (graphbuilder.pyx:214-226) for the KAT of
// textgcn/test/test_cfunc.py:81-99 (small vocabularies only: writes V(V+1)/2 counters).
int tgcn_ww_counts_packed(const int32_t* X, int64_t n_docs, int64_t seq_len, int64_t n_vocab, int64_t window_size,
                          uint32_t* c_ij_out, uint64_t* n_windows_out) {
  if (!X || !c_ij_out || n_vocab <= 0 || seq_len <= 0 || window_size <= 0) return 1;
  const uint64_t V = (uint64_t)n_vocab;
  PairTable tab;
  uint64_t nw = 0; int bad = 0;
  count_docs(X, 0, n_docs, seq_len, window_size, V, tab, nw, bad);
  if (bad) return 2;
  std::memset(c_ij_out, 0, sizeof(uint32_t) * (V * (V + 1) / 2));
  for (const PairTable::Slot& sl : tab.slots) {
    if (sl.key == PairTable::EMPTY) continue;
    const uint64_t i = sl.key / V, j = sl.key % V;             // i <= j
    c_ij_out[i * V + j - (i + 1) * i / 2] = sl.val;
  }
  if (n_windows_out) *n_windows_out = nw;
  return 0;
}

}  // extern "C"
