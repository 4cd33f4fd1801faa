/*
 * CPU oracle for the word-word PMI edge builder -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of textgcn/lib/clib/graphbuilder.pyx (the reference's only native
 * component), kept deliberately in the reference's own dense O(V^2) form so it can be read
 * against it line by line:
 *   sliding_window      graphbuilder.pyx:71-115   (position-pair counts per window, diagonal
 *                                                  included, windows stop at the first one whose
 *                                                  last slot is padding except window 0)
 *   edges_from_counts   graphbuilder.pyx:118-211  (p_i = c_ii / n_windows, pmi = log(p_ij/(p_i p_j))
 *                                                  in float, keep pmi > 1e-10, symmetric COO emitted
 *                                                  as (i,j),(j,i) in upper-triangle row-major order)
 *   packed index helpers graphbuilder.pyx:214-259
 * Pinned by the reference's own known-answer test (textgcn/test/test_cfunc.py:81-99, exact c_ij)
 * and by golden vectors generated with the reference builder itself (tests/golden/, made by
 * oracle/make_golden.py from oracle/_ref).  Only tests/ may load this library; the product
 * builder is pytextgcn_b200/csrc_host/graph_builder.cpp (sparse, threaded, no V < 65536 limit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* graphbuilder.pyx:214-226 */
static uint64_t sym_diag_idx(uint64_t row, uint64_t col, uint64_t n) {
  if (row >= col) return col * n + row - ((col + 1) * col / 2);
  return row * n + col - ((row + 1) * row / 2);
}

static uint64_t sym_diag_size(uint64_t n) { return n * (n + 1) / 2; }

/* graphbuilder.pyx:71-115.  c_ij must hold sym_diag_size(n_vocab) zeroed uint32.  Returns n_windows. */
uint32_t oracle_sliding_window(const int32_t* X, uint32_t* c_ij, uint32_t window_size, uint32_t n_vocab,
                               uint32_t n_documents, uint32_t seq_len) {
  uint32_t n_windows = 0;
  /* window_size > seq_len: the reference's unsigned `seq_len - window_size + 1` wraps and the loops read
   * past the row (undefined behaviour, graphbuilder.pyx:96-108).  Both this oracle and the product builder
   * define that case as window_size = seq_len (one window per document over all its tokens). */
  if (window_size > seq_len) window_size = seq_len;
  for (uint32_t i = 0; i < n_documents; ++i) {
    for (uint32_t j = 0; j + window_size < seq_len + 1; ++j) {              /* range(seq_len - window_size + 1) */
      if (X[(uint64_t)i * seq_len + j + window_size - 1] == -1 && j != 0) break;  /* :98-100 */
      n_windows += 1;
      for (uint32_t k = j; k < j + window_size; ++k) {
        for (uint32_t l = k; l < j + window_size; ++l) {
          if (X[(uint64_t)i * seq_len + k] != -1 && X[(uint64_t)i * seq_len + l] != -1) {
            uint32_t a = (uint32_t)X[(uint64_t)i * seq_len + k], b = (uint32_t)X[(uint64_t)i * seq_len + l];
            c_ij[sym_diag_idx(a, b, n_vocab)] += 1;                          /* :103-113 */
          } else {
            break;
          }
        }
      }
    }
  }
  return n_windows;
}

uint64_t oracle_sym_diag_size(uint64_t n) { return sym_diag_size(n); }

/* graphbuilder.pyx:118-211.  Two passes like the reference: count, then emit.  The caller passes
 * coo/weights == NULL to get the edge count, then buffers of 2*n_edges int32 / n_edges float. */
uint64_t oracle_edges_from_counts(const uint32_t* c_ij, uint32_t n_vocab, uint32_t n_windows, int32_t* coo,
                                  float* weights) {
  const float EPSILON = 1e-10f;                                              /* :20 */
  float* p = (float*)malloc(sizeof(float) * (n_vocab ? n_vocab : 1));
  for (uint32_t i = 0; i < n_vocab; ++i) p[i] = (float)c_ij[sym_diag_idx(i, i, n_vocab)] / (float)n_windows;  /* :146-147 */
  uint64_t k = 0;
  for (uint32_t i = 0; i + 1 < n_vocab; ++i) {
    for (uint32_t j = i + 1; j < n_vocab; ++j) {
      float p_ij = (float)c_ij[sym_diag_idx(i, j, n_vocab)] / (float)n_windows;    /* :156 */
      if (p_ij == 0 || p[i] == 0 || p[j] == 0) continue;                     /* :157-160 */
      float pmi = (float)log((double)(p_ij / (p[i] * p[j])));                /* :161: float ratio, libc double log, float result */
      if (pmi > EPSILON) {
        if (coo) {
          coo[2 * k] = (int32_t)i; coo[2 * k + 1] = (int32_t)j; weights[k] = pmi; ++k;   /* :183-186 */
          coo[2 * k] = (int32_t)j; coo[2 * k + 1] = (int32_t)i; weights[k] = pmi; ++k;   /* :188-191 */
        } else {
          k += 2;
        }
      }
    }
  }
  free(p);
  return k;
}
