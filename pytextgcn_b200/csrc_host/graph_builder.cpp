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
//    counts live in open-addressing hash tables keyed by the 64-bit pair id, one table per thread,
//    merged at the end.  Memory is O(#distinct co-occurring pairs); indices are 64-bit.
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
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/textgcn_host.h"

namespace {

struct PairTable {              // open addressing, linear probing, key = i * V + j (i <= j), 0xFFFF.. = empty
  std::vector<uint64_t> keys;
  std::vector<uint32_t> vals;
  uint64_t mask = 0, used = 0;
  static constexpr uint64_t EMPTY = ~0ull;
  explicit PairTable(uint64_t cap_pow2 = 1u << 16) { reset(cap_pow2); }
  void reset(uint64_t cap) { keys.assign(cap, EMPTY); vals.assign(cap, 0); mask = cap - 1; used = 0; }
  static inline uint64_t hash(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
  void grow() {
    std::vector<uint64_t> ok; std::vector<uint32_t> ov;
    ok.swap(keys); ov.swap(vals);
    reset((mask + 1) * 2);
    for (size_t i = 0; i < ok.size(); ++i) if (ok[i] != EMPTY) add(ok[i], ov[i]);
  }
  inline void add(uint64_t k, uint32_t c) {
    if ((used + 1) * 10 > (mask + 1) * 7) grow();
    uint64_t h = hash(k) & mask;
    while (true) {
      if (keys[h] == k) { vals[h] += c; return; }          // uint32 wrap-around like the reference's counters
      if (keys[h] == EMPTY) { keys[h] = k; vals[h] = c; ++used; return; }
      h = (h + 1) & mask;
    }
  }
};

struct Result {
  std::vector<int32_t> coo;     // [n_edges][2]
  std::vector<float> w;         // [n_edges]
  uint64_t n_windows = 0;
};

void count_docs(const int32_t* X, int64_t d0, int64_t d1, int64_t seq_len, int64_t w, uint64_t V, PairTable& tab,
                uint64_t& n_windows, int& bad) {
  for (int64_t d = d0; d < d1; ++d) {
    const int32_t* x = X + d * seq_len;
    int64_t len = 0;
    while (len < seq_len && x[len] != -1) ++len;            // padding only at the tail (text2graph.py:40-44)
    const int64_t wcap = std::min(w, seq_len);
    const int64_t jmax = std::max<int64_t>(0, len - wcap);
    n_windows += (uint64_t)(jmax + 1);                       // window 0 always counts, even for an empty document
    for (int64_t a = 0; a < len; ++a) {
      const int64_t xa = x[a];
      if (xa < 0 || (uint64_t)xa >= V) { bad = 1; continue; }
      const int64_t bend = std::min(len, a + wcap);
      for (int64_t b = a; b < bend; ++b) {
        const int64_t xb = x[b];
        if (xb < 0 || (uint64_t)xb >= V) { bad = 1; continue; }
        const int64_t jlo = std::max<int64_t>(0, b - wcap + 1), jhi = std::min(a, jmax);
        if (jhi < jlo) continue;
        const uint64_t lo = (uint64_t)std::min(xa, xb), hi = (uint64_t)std::max(xa, xb);
        tab.add(lo * V + hi, (uint32_t)(jhi - jlo + 1));
      }
    }
  }
}

}  // namespace

extern "C" {

// Builds the edge list; returns an opaque handle (NULL on bad input), *n_edges_out = number of DIRECTED edges.
void* tgcn_ww_build(const int32_t* X, int64_t n_docs, int64_t seq_len, int64_t n_vocab, int64_t window_size,
                    int32_t n_threads, int64_t* n_edges_out, uint64_t* n_windows_out) {
  if (!X || n_docs < 0 || seq_len <= 0 || n_vocab <= 0 || window_size <= 0 || !n_edges_out) return nullptr;
  const uint64_t V = (uint64_t)n_vocab;
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  T = (int)std::max<int64_t>(1, std::min<int64_t>(T, std::max<int64_t>(1, n_docs / 64)));
  std::vector<PairTable> tabs((size_t)T);
  std::vector<uint64_t> nwin((size_t)T, 0);
  std::vector<int> bad((size_t)T, 0);
  std::vector<std::thread> th;
  const int64_t per = (n_docs + T - 1) / T;
  for (int t = 0; t < T; ++t) {
    const int64_t d0 = std::min<int64_t>(n_docs, t * per), d1 = std::min<int64_t>(n_docs, d0 + per);
    th.emplace_back([&, t, d0, d1]() { count_docs(X, d0, d1, seq_len, window_size, V, tabs[t], nwin[t], bad[t]); });
  }
  for (auto& t : th) t.join();
  for (int t = 0; t < T; ++t) if (bad[t]) return nullptr;    // token id outside [0, n_vocab)
  uint64_t n_windows = 0;
  for (uint64_t v : nwin) n_windows += v;
  // merge: gather (key, count) of every table, sort by key, add equal keys
  std::vector<std::pair<uint64_t, uint32_t>> all;
  size_t total = 0;
  for (auto& tb : tabs) total += tb.used;
  all.reserve(total);
  for (auto& tb : tabs) {
    for (size_t i = 0; i < tb.keys.size(); ++i) if (tb.keys[i] != PairTable::EMPTY) all.emplace_back(tb.keys[i], tb.vals[i]);
    std::vector<uint64_t>().swap(tb.keys); std::vector<uint32_t>().swap(tb.vals);
  }
  std::sort(all.begin(), all.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  size_t m = 0;
  for (size_t i = 0; i < all.size(); ++i) {
    if (m > 0 && all[m - 1].first == all[i].first) all[m - 1].second += all[i].second;
    else all[m++] = all[i];
  }
  all.resize(m);
  // p_i = c_ii / n_windows   (graphbuilder.pyx:146-147)
  const uint32_t nw32 = (uint32_t)n_windows;                  // the reference keeps n_windows in an unsigned int
  std::vector<float> p((size_t)V, 0.0f);
  for (const auto& kv : all) {
    const uint64_t i = kv.first / V, j = kv.first % V;
    if (i == j) p[i] = (float)kv.second / (float)nw32;
  }
  auto* res = new Result();
  res->n_windows = n_windows;
  const float EPSILON = 1e-10f;
  for (const auto& kv : all) {                                // sorted by i, then j: upper-triangle row-major order
    const uint64_t i = kv.first / V, j = kv.first % V;
    if (i == j) continue;
    const float p_ij = (float)kv.second / (float)nw32;
    if (p_ij == 0 || p[i] == 0 || p[j] == 0) continue;        // graphbuilder.pyx:157-160
    const float pmi = (float)std::log((double)(p_ij / (p[i] * p[j])));
    if (pmi > EPSILON) {
      res->coo.push_back((int32_t)i); res->coo.push_back((int32_t)j); res->w.push_back(pmi);
      res->coo.push_back((int32_t)j); res->coo.push_back((int32_t)i); res->w.push_back(pmi);
    }
  }
  *n_edges_out = (int64_t)res->w.size();
  if (n_windows_out) *n_windows_out = n_windows;
  return res;
}

int tgcn_ww_fetch(void* handle, int32_t* coo_out, float* w_out) {
  if (!handle || !coo_out || !w_out) return 1;
  auto* r = static_cast<Result*>(handle);
  std::memcpy(coo_out, r->coo.data(), r->coo.size() * sizeof(int32_t));
  std::memcpy(w_out, r->w.data(), r->w.size() * sizeof(float));
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
  for (size_t s = 0; s < tab.keys.size(); ++s) {
    if (tab.keys[s] == PairTable::EMPTY) continue;
    const uint64_t i = tab.keys[s] / V, j = tab.keys[s] % V;   // i <= j
    c_ij_out[i * V + j - (i + 1) * i / 2] = tab.vals[s];
  }
  if (n_windows_out) *n_windows_out = nw;
  return 0;
}

}  // extern "C"
