/*
 * textgcn_host.h -- C ABI of the host-side (CPU, no CUDA) native graph builder, libtextgcn_host.so.
 *
 * Replaces the reference's only native entry point, the Cython function
 *     compute_word_word_edges(int[:, ::1] X, n_vocab, n_documents, seq_len, window_size=20, n_jobs=1, verbose=0)
 *         -> (int32[E, 2] COO, float32[E] PMI weights)                    textgcn/lib/clib/graphbuilder.pyx:23-66
 * called by Text2GraphTransformer.fit_transform (textgcn/lib/text2graph.py:156-160).  Same result bit for bit
 * (edges as (i,j),(j,i) pairs in upper-triangle row-major order, PMI = log(p_ij / (p_i p_j)) > 0 only), different
 * algorithm: closed-form sliding-window counts into per-thread hash tables instead of a packed V x V array, so memory
 * is O(#co-occurring pairs), vocabularies >= 65,536 work, and `n_threads` is honoured.
 *
 * Conventions: every pointer is a HOST pointer; X is row-major int32 [n_docs][seq_len] with token ids in [0, n_vocab),
 * -1 = padding at the tail of short documents (skipped, as the reference does: text2graph.py:40-44,
 * test_cfunc.py:83-86); the caller owns all output buffers;
 * a handle returned by tgcn_ww_build must be released with tgcn_ww_free.  Python binding: pytextgcn_b200/graphbuilder.py.
 */
#ifndef TEXTGCN_HOST_H
#define TEXTGCN_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Counts the co-occurrences of all windows, computes the PMI edge list and keeps it behind an opaque handle.
 * Returns NULL on bad input (null X, non-positive sizes, a token id >= n_vocab or < -1) or when memory runs out.
 * n_threads <= 0: all hardware threads.  *n_edges_out = number of DIRECTED edges (2 per unordered pair);
 * *n_windows_out (optional) = number of sliding windows (the denominator of the probabilities, graphbuilder.pyx:150). */
void* tgcn_ww_build(const int32_t* X, int64_t n_docs, int64_t seq_len, int64_t n_vocab, int64_t window_size,
                    int32_t n_threads, int64_t* n_edges_out, uint64_t* n_windows_out);

/* Copies the result out: coo_out int32 [n_edges][2], w_out float [n_edges].  0 = ok, 1 = null argument. */
int tgcn_ww_fetch(void* handle, int32_t* coo_out, float* w_out);

void tgcn_ww_free(void* handle);

/* Raw pair counts c_ij in the reference's packed upper-triangular layout (graphbuilder.pyx:214-226), V (V + 1) / 2
 * uint32 counters, diagonal = occurrences of word i: the quantity the reference's known-answer test pins
 * (textgcn/test/test_cfunc.py:81-99).  Small vocabularies only.  0 = ok, 1 = bad argument, 2 = token id out of range. */
int tgcn_ww_counts_packed(const int32_t* X, int64_t n_docs, int64_t seq_len, int64_t n_vocab, int64_t window_size,
                          uint32_t* c_ij_out, uint64_t* n_windows_out);

#ifdef __cplusplus
}
#endif
#endif /* TEXTGCN_HOST_H */
