/*
 * textgcn_b200.h -- C ABI of the B200-native TextGCN training hot path.
 *
 * The reference (BeFranke/PyTextGCN) has no native boundary for this path: it is plain
 * Python, `GCNConv(in, out, add_self_loops=True)` constructed at textgcn/lib/models.py:11-15
 * and called as `layer(x, g.edge_index, g.edge_attr)` at textgcn/lib/models.py:20, looped by
 * flat_amazon.py:99-117.  The entry points below are what a binding for that path would
 * bind (ctypes stub: INTEGRATION.md); each comment names the reference step it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - nothing here allocates device memory: the caller owns all buffers (torch does, in the
 *     Python host) and passes workspaces whose size comes from the `*_workspace_bytes` calls;
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued, never synchronised,
 *     unless the comment says "host sync";
 *   - return value: 0 = ok, otherwise a TGCN_E* code; tgcn_last_error() gives the text
 *     (thread-local).  The Python host turns non-zero into RuntimeError, which is what the
 *     reference's sweeps catch (old/h_o_train.py:129-131);
 *   - dense matrices are row-major with an explicit leading dimension (in elements);
 *   - built only for sm_100a.  There is no CPU path.
 */
#ifndef TEXTGCN_B200_H
#define TEXTGCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGCN_OK        0
#define TGCN_EINVAL    1   /* bad argument (shape, alignment, null pointer) */
#define TGCN_ECUDA     2   /* a CUDA runtime call or launch failed */
#define TGCN_EWORKSPACE 3  /* workspace too small */
#define TGCN_EINDEX    4   /* edge_index holds a node id outside [0, n_nodes) */

/* dense element types for SpMM operands */
#define TGCN_F32  0
#define TGCN_BF16 1

/* epilogue activation (reference: identity -- models.py:22 is commented out) */
#define TGCN_ACT_NONE 0
#define TGCN_ACT_RELU 1

/* dropout modes of the SpMM epilogue (reference: F.dropout, models.py:23) */
#define TGCN_DROP_NONE   0   /* eval / p == 0 */
#define TGCN_DROP_MASK   1   /* keep-mask supplied by the caller (uint8 per element) */
#define TGCN_DROP_PHILOX 2   /* keep = philox4x32-10(seed, element index) >= p, regenerated in backward */

const char* tgcn_last_error(void);
int tgcn_version(void);
/* number of kernels this library has launched (or captured into a CUDA graph) so far in this process */
uint64_t tgcn_launch_count(void);
/* number of SMs / compute capability of the current device (host sync; used by the host to size grids) */
int tgcn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * (1) Graph upload: COO -> CSR of A_hat = D^-1/2 (A + I) D^-1/2, bit-exact against gcn_norm.
 * Replaces: gcn_norm / add_remaining_self_loops, executed by GCNConv on EVERY call because
 * cached=False (models.py:11-15,20; [PyG-1.6.3] gcn_conv.py).  Here: once per graph.
 *
 * edge_src[e*idx_stride] = edge_index[0][e] (source j), edge_dst[e*idx_stride] =
 * edge_index[1][e] (target i); idx_stride lets the non-contiguous `coo.T` view the reference
 * emits (text2graph.py:171,192) be consumed without a copy (src = base, dst = base+1,
 * stride 2).  edge_w may be NULL (all ones).
 * CSR rows are TARGET nodes; row i lists its in-edges in original edge order, then its self
 * loop; colidx = source node.  Degree = sequential fp32 sum in that order, dis =
 * 1/sqrt(deg) (IEEE, inf -> 0), val = (dis[src]*w)*dis[dst] -- all bit-identical to torch CPU.
 * Pre-existing self loops are dropped and their weight becomes the loop weight (as PyG).
 * Outputs: rowptr[n_nodes+1], colidx/val capacity n_edges+n_nodes (nnz = rowptr[n_nodes]),
 * dis[n_nodes], edge_slot (optional, n_edges+n_nodes): CSR slot of gcn_norm's k-th edge
 * (original edges 0..E-1, then loops), -1 for dropped loops.
 * status_out (device int32[2]): [0] = 0 ok / TGCN_EINDEX, [1] = nnz.  The caller reads it
 * after synchronising.
 */
int tgcn_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, size_t* bytes_out);
int tgcn_csr_from_coo_gcn_norm(const int64_t* edge_src, const int64_t* edge_dst, int64_t idx_stride,
                               const float* edge_w, int64_t n_edges, int64_t n_nodes,
                               int32_t* rowptr, int32_t* colidx, float* val, float* dis,
                               int32_t* edge_slot, int32_t* status_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Row-chunk work list for the SpMM kernels: rows longer than `chunk_nnz` are split so hub
 * word rows do not serialise one warp.  chunks[k] = {row, begin, end, slot}; slot == -1 means
 * "row owned entirely by this chunk, write C directly", otherwise the index of the partial
 * row in the split-row scratch.  n_chunks/n_partials are returned through counts_out (device
 * int32[4]: n_chunks, n_partial_slots, n_split_rows, max_row_nnz).
 * Capacity needed: n_rows + nnz/chunk_nnz + 1 chunks.
 * Works on any row range [row_begin, row_end) of the CSR (1D row partition). */
int tgcn_spmm_plan(const int32_t* rowptr, int64_t row_begin, int64_t row_end, int32_t chunk_nnz,
                   int32_t* chunks /* [cap][4] */, int64_t chunk_capacity,
                   int32_t* split_rows /* [cap][3] = {row, first_slot, n_slots} */,
                   int32_t* slot_owner /* [n_slots]: split-row index owning each scratch slot */,
                   int32_t* counts_out, void* workspace, size_t workspace_bytes, void* stream);
int tgcn_spmm_plan_workspace_bytes(int64_t n_rows, size_t* bytes_out);

/* ------------------------------------------------------------------------------------------
 * (2) SpMM with fused epilogue:  C[r,:] = epi( sum_k val[k] * B[colidx[k],:] )  for the rows
 * named by the plan.  Replaces GCNConv.propagate (index_select + mul + scatter_add, atomics)
 * + bias add + F.dropout  (models.py:20,23).  Deterministic: no atomics.
 *   epi(z) = dropout(act(z + bias)),  dropout per TGCN_DROP_*; element index for Philox is
 *   r*F + c with r the GLOBAL row id, so forward and backward agree and so do all ranks.
 * Optional fused projection (layer 2's thin X*W done on the row while it is in registers):
 *   P[r,:] = epi(...)[r,:] @ W_proj[F, n_proj]   (replaces torch.matmul(x, weight) of the
 *   NEXT layer, models.py:20).
 * B / C dtype: TGCN_F32 or TGCN_BF16 (accumulation always fp32).
 * scratch: (n_partial_slots x F) fp32 partial rows for split rows.
 */
typedef struct {
  const int32_t* rowptr; const int32_t* colidx; const float* val;   /* CSR of A_hat */
  const int32_t* chunks; int32_t n_chunks;                           /* plan */
  const int32_t* split_rows; int32_t n_split_rows;
  const int32_t* slot_owner;    /* from tgcn_spmm_plan: split-row index owning each scratch slot */
  int32_t* split_counters;      /* int32[n_split_rows], zero once, self-resetting: the last-arriving chunk of a split
                                   row adds the partial rows in slot order inside the kernel (required when
                                   n_split_rows > 0) */
  float* scratch;                                                    /* partial rows */
  const void* B; int64_t ldb; int32_t b_dtype;
  void* C; int64_t ldc; int32_t c_dtype;                             /* may be NULL if only P wanted */
  int32_t F;                                                         /* feature width */
  int64_t c_row_offset;   /* C/P/keep-mask row r is stored at r - c_row_offset (row-partitioned output) */
  const float* bias; int32_t bias_len;   /* [bias_len] or NULL; bias_len <= F (0 means F): columns past it are padding */
  int32_t act;
  int32_t drop_mode; float drop_p; const uint8_t* keep_mask; int64_t ldmask;
  uint64_t philox_seed; uint64_t philox_offset;
  const int64_t* philox_offset_dev;   /* optional device counter added to philox_offset (CUDA-graph replays) */
  int64_t philox_row_offset;          /* added to the CSR row id in the Philox element index (row shards: global id of local row 0) */
  const float* W_proj; int32_t n_proj; float* P; int64_t ldp;       /* optional projection */
  /* optional fused Adam/AMSGrad (backward of layer 1 with X = I: output row r IS dW1[r], so the update of
   * W1[r] runs on the row while it is in registers; replaces optimizer.step() for W1, flat_amazon.py:106).
   * adam_hyper_dev = float[2] written by tgcn_adam_prepare; C may be NULL (gradient not materialised);
   * adam_param_mirror_mc: multicast mapping of the parameter (row-partitioned mode) or NULL. */
  float* adam_param; float* adam_exp_avg; float* adam_exp_avg_sq; float* adam_max_exp_avg_sq; int64_t adam_ld;
  const float* adam_hyper_dev; float adam_beta1; float adam_beta2; float adam_eps; void* adam_param_mirror_mc;
  /* optional dense-tile contributions computed by tgcn_spmm_tc (2b): when tc_part != NULL the CSR holds only the
   * entries OUTSIDE the dense tiles and, before the epilogue, row r adds part[s][tc_rank[r] % 128][:] for the slots
   * s in [tc_slot_ptr[tc_rank[r] / 128], tc_slot_ptr[tc_rank[r] / 128 + 1]), in slot order. */
  const float* tc_part; int64_t tc_ld; const int32_t* tc_rank; const int32_t* tc_slot_ptr;
  /* optional partial rows computed elsewhere (bipartite exchange of the multi-GPU path: the contributions of the
   * other ranks' documents to this rank's word rows): local row r < raw_rows adds raw_in[k * raw_stride + r * raw_ld + :]
   * for k = 0 .. n_raw - 1, in that order, before the dense-tile partials and the epilogue. */
  const float* raw_in; int64_t raw_ld; int64_t raw_stride; int32_t n_raw; int64_t raw_rows;
  /* word-block exchange fused into the kernels (multi-GPU, see (6)):
   * adam_mirror_rows > 0: only the first adam_mirror_rows local rows (the rank's words) are repeated to the multicast
   *   mapping (0 = all rows);
   * c_scatter_bases != NULL (device array of peer-mapped base pointers, fp32 C only): output row r is stored to
   *   bases[r / c_scatter_rows] + (c_scatter_row0 + r % c_scatter_rows) * ldc instead of C + r * ldc -- the
   *   all-to-all of the partial word rows happens in the epilogue's stores (plain peer stores over NVLink). */
  int64_t adam_mirror_rows;
  const uint64_t* c_scatter_bases; int64_t c_scatter_rows; int64_t c_scatter_row0;
} tgcn_spmm_args;
int tgcn_spmm(const tgcn_spmm_args* args, void* stream);

/* (2b) Dense-tile part of a hybrid propagation on the tensor cores (csrc/spmm_tc.cu).  The nodes are ranked by
 * degree; the 128 x 16 blocks of A_hat (in rank space) that are dense enough are stored as dense TF32 hi/lo tiles and
 * multiplied with tcgen05.mma (3xTF32: fp32 accuracy), everything else stays in the CSR that tgcn_spmm gathers.
 * tgcn_spmm_tc packs the operand (transpose to K-major + hi/lo split, `Bt` workspace of
 * tgcn_spmm_tc_workspace_elems floats) and writes one 128 x F partial result per unit into part[unit slot]; the
 * following tgcn_spmm call (tc_part/tc_rank/tc_slot_ptr) adds them to the gathered remainder and runs the epilogue.
 * Replaces GCNConv.propagate for the dense blocks (models.py:20).  Plan: pytextgcn_b200/tc_plan.py. */
typedef struct {
  const float* A_tiles;     /* [n_tiles][128][16] fp32 values of A_hat, 64-byte swizzle pre-applied (see spmm_tc.cu) */
  const int32_t* tile_kb;   /* [n_tiles] column block (16 ranks) of each tile */
  int32_t n_tiles;
  const int32_t* units;     /* [n_units][4] = {tile_begin, tile_end, slot, row_block}; empty units (begin == end) allowed */
  int32_t n_units;
  const int32_t* perm;      /* [n_col_blocks * 16] node id of each rank, -1 past the last node */
  int32_t n_col_blocks;
} tgcn_tc_plan;
int tgcn_spmm_tc(const tgcn_tc_plan* plan, const float* B, int64_t ldb, int32_t F, float* Bt, float* part, int64_t ldp,
                 void* stream);
int tgcn_spmm_tc_workspace_elems(int32_t F, int64_t n_col_blocks, int64_t* bt_elems_out);

/* ------------------------------------------------------------------------------------------
 * (3) Masked log-softmax / NLL and its gradient, one pass over the logits.
 * Replaces: gcn(g)[mask] boolean gather + CrossEntropyLoss(mean) + its backward
 * (flat_amazon.py:82,101-102,105).  y is never read where mask == 0 (labels may be -1 there,
 * perlabel_amazon.py:108-109).  loss_out[0] = mean NLL, loss_out[1] = #masked rows (as float).
 * dZ (optional) = (softmax - onehot)/n_mask on masked rows, 0 elsewhere.
 * pred_out (optional, int32[n_rows]) = argmax per row; correct_out (optional int32[1]) +=
 * #masked rows with argmax == y.  n_mask_total > 0 overrides the divisor (row-partitioned
 * runs pass the global count); partial sums land in partial_out (fp64[2]: sum nll, count).
 */
int tgcn_masked_nll(const float* Z, int64_t ldz, int64_t n_rows, int32_t n_classes,
                    const int64_t* y, const uint8_t* mask, int64_t n_mask_total,
                    float* loss_out, double* partial_out,
                    float* dZ, int64_t lddz, int32_t* pred_out, int32_t* correct_out,
                    void* dZ_mirror_mc /* multicast mapping of dZ or NULL: see (6) */,
                    const uint8_t* mask2, int32_t* correct2_out /* optional second accuracy count from the same pass:
                       #rows of mask2 with argmax == y (the eval pass scores validation AND training rows, flat_amazon.py:113-114) */,
                    void* workspace, size_t workspace_bytes, void* stream);
int tgcn_masked_nll_workspace_bytes(int64_t n_rows, size_t* bytes_out);

/* ------------------------------------------------------------------------------------------
 * (4) Dense backward of layer 2's projection + dropout, one pass over the rows:
 *   dW2 = H1d^T G2  (H x C),  db_hidden = colsum(dZ1),  db_out = colsum(dZ2)
 *   dZ1 = (G2 W2^T) .* keep/(1-p) .* act'(H1d)
 * Replaces the autograd of torch.matmul / F.dropout / bias (flat_amazon.py:105).
 * Deterministic two-stage reduction (per-CTA partials in workspace, fixed-order final sum).
 */
typedef struct {
  const float* G2; int64_t ldg2;        /* [n_rows, C]  = A_hat dZ2 */
  const void* H1d; int64_t ldh; int32_t h_dtype;   /* [n_rows, H] dropped hidden activations (forward output) */
  const float* W2;                      /* [H, C] */
  const float* dZ2; int64_t lddz2;      /* [n_rows, C] for db_out (may be NULL) */
  int64_t n_rows; int64_t row_offset;   /* global row id of local row 0 (Philox index) */
  int32_t H; int32_t C;
  int32_t act; int32_t drop_mode; float drop_p; const uint8_t* keep_mask; int64_t ldmask;
  uint64_t philox_seed; uint64_t philox_offset; const int64_t* philox_offset_dev;
  void* dZ1; int64_t lddz1; int32_t dz1_dtype;      /* out [n_rows, H] */
  void* dZ1_mirror_mc;                              /* multicast mapping of dZ1 (fp32) or NULL: see (6) */
  float* dW2; float* db_hidden; float* db_out;      /* out [H*C], [H], [C] */
  int64_t dZ1_mirror_rows;                          /* > 0: only the first rows are repeated to the multicast mapping (0 = all) */
} tgcn_dense_bwd_args;
int tgcn_dense_bwd(const tgcn_dense_bwd_args* args, void* workspace, size_t workspace_bytes, void* stream);
int tgcn_dense_bwd_workspace_bytes(int32_t H, int32_t C, size_t* bytes_out);

/* Thin projection P = X W (+ bias) (X [n,K] fp32/bf16, W [K,M], bias [M] or NULL); replaces torch.matmul(x, weight) of a
 * hidden->classes layer when it is not fused into the producing SpMM.  With bias it is layer 2 in the
 * propagate-first order (A_hat H) W2 + b2, used when there are more classes than hidden units (perlevel_dbpedia.py:
 * 219 classes, hidden 32): the propagation then moves hidden-wide instead of classes-wide rows. */
int tgcn_project(const void* X, int64_t ldx, int32_t x_dtype, int64_t n_rows, int32_t K,
                 const float* W, int32_t M, const float* bias, float* P, int64_t ldp, void* P_mirror_mc, void* stream);
/* The same projection with the training forward's dropout applied to X while it is loaded (the keep decision of the
 * tgcn_spmm epilogue: keep-mask bytes, or Philox keyed by (seed, offset, (row + philox_row_offset) * K + col)) and the
 * dropped block optionally written to Xd for the backward pass: P = dropout(X) W (+ bias), Xd = dropout(X).  One pass
 * instead of tgcn_dropout_apply + tgcn_project when the pre-dropout activation is shared with an eval forward.
 * fp32 X takes the row-per-lane kernel (csrc/dense_bwd.cu k_project_rows); bf16 X the 4-row kernel (no dropout). */
typedef struct {
  const void* X; int64_t ldx; int32_t x_dtype; int64_t n_rows; int32_t K;
  const float* W; int32_t M; const float* bias;
  float* P; int64_t ldp; void* P_mirror_mc;
  int32_t drop_mode; float drop_p; const uint8_t* keep_mask; int64_t ldmask;
  uint64_t philox_seed; uint64_t philox_offset; const int64_t* philox_offset_dev; int64_t philox_row_offset;
  float* Xd; int64_t ldxd;
  int64_t mirror_rows;                              /* > 0: only the first rows are repeated to P_mirror_mc (0 = all) */
} tgcn_project_args;
int tgcn_project_ex(const tgcn_project_args* args, void* stream);
/* out[c] = sum_r X[r, c] (deterministic two-stage sum): the bias gradient db1 = colsum(dZ1) in the propagate-first order */
int tgcn_colsum(const float* X, int64_t ldx, int64_t n_rows, int32_t F, float* out, void* workspace, size_t workspace_bytes,
                void* stream);
int tgcn_colsum_workspace_bytes(int32_t F, size_t* bytes_out);

/* Y = dropout(X) with exactly the keep decision of the tgcn_spmm epilogue (same mode / seed / offset / element index
 * (row + philox_row_offset) * F + col).  Used to reuse the pre-dropout hidden activation A_hat (X W1) + b1 of an eval
 * forward as the training forward of the next epoch: W1/b1 do not change in between (flat_amazon.py:100-110) and
 * F.dropout follows the product (models.py:20-23), so the result is bit-identical to recomputing the propagation. */
int tgcn_dropout_apply(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int32_t F,
                       int32_t drop_mode, float drop_p, const uint8_t* keep_mask, int64_t ldmask,
                       uint64_t philox_seed, uint64_t philox_offset, const int64_t* philox_offset_dev,
                       int64_t philox_row_offset, void* stream);

/* Hierarchy-feature prologue/epilogue for X = [I | F] (text2graph.py:237-241;
 * perlevel_dbpedia.py:140-141):  XW[r,:] = W1[r,:] (+ F[r-n_vocab,:] @ W1[N:, :] on doc rows);
 * and its transpose for dW1[N:, :] = F^T G1[docs].  Fdoc is dense [n_docs, Cprev]. */
int tgcn_hier_forward(const float* W1, int64_t ldw, int64_t n_nodes, int64_t n_vocab,
                      const float* Fdoc, int64_t ldf, int32_t c_prev, int32_t H,
                      float* XW, int64_t ldxw, void* stream);
int tgcn_hier_backward(const float* G1, int64_t ldg, int64_t n_nodes, int64_t n_vocab,
                       const float* Fdoc, int64_t ldf, int32_t c_prev, int32_t H,
                       float* dW_tail /* [c_prev, H] */, void* workspace, size_t workspace_bytes, void* stream);
int tgcn_hier_backward_workspace_bytes(int32_t c_prev, int32_t H, size_t* bytes_out);

/* ------------------------------------------------------------------------------------------
 * (5) Fused Adam / AMSGrad step (flat_amazon.py:89,106: Adam(lr, amsgrad=True), betas
 * (0.9, 0.999), eps 1e-8, no weight decay; torch semantics:
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  vhat = amsgrad ? max(vmax, v) : v;
 *   p -= lr/(1-b1^t) * m / (sqrt(vhat)/sqrt(1-b2^t) + eps) ).
 * `step` is the 1-based step count AFTER increment; when step_dev (device int64) is not NULL
 * it is read instead, so a captured CUDA graph can be replayed with a moving step count
 * (tgcn_increment_step bumps it inside the graph).  vmax may be NULL when amsgrad == 0.
 */
int tgcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                   int64_t n, float lr, float beta1, float beta2, float eps, int32_t amsgrad,
                   int64_t step, const int64_t* step_dev, void* param_mirror_mc, void* stream);
int tgcn_increment_step(int64_t* step_dev, void* stream);
/* step_dev += 1 and hyper_dev[0..1] = {lr/(1-b1^t), sqrt(1-b2^t)} for kernels that fuse the update */
int tgcn_adam_prepare(int64_t* step_dev, float* hyper_dev, float lr, float beta1, float beta2, void* stream);
/* the same update for up to 4 SMALL tensors in one launch (b1, W2, b2); the arrays are HOST arrays of device pointers */
int tgcn_adam_step_small(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, float* const* max_exp_avg_sq, const int64_t* sizes, float lr,
                         float beta1, float beta2, float eps, int32_t amsgrad, int64_t step, const int64_t* step_dev,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * (6) Exchange step of the 1D row partition over NVLink peer memory.  Replaces the
 * ncclAllGather between layers: the local slice [src, src+bytes) is stored into every peer's
 * symmetric buffer at dst_offset_bytes (peer_bases_host: HOST array of `world` device pointers, the
 * peer mappings of the same symmetric allocation), or once through the NVSwitch multicast mapping
 * of that allocation when multicast_base != NULL (multimem.st: the switch replicates).  The caller
 * follows it with a cross-rank barrier before any rank reads.  The producer kernels of the exchanged
 * buffers (tgcn_adam_step -> W1, tgcn_project -> P, tgcn_masked_nll -> dZ2, tgcn_dense_bwd -> dZ1)
 * take an optional `*_mirror_mc` pointer: the multicast mapping of their output, to which they repeat
 * every store with multimem.st -- compute and exchange in ONE kernel, only the barrier remains.  tgcn_sum_slots adds `n_slots`
 * vectors in slot order (the all-reduce of the small gradients: every rank pushes its vector into
 * slot `rank` of every peer, then sums the slots locally -- bit-identical on all ranks).
 * Word-block partition (documents >> words; pytextgcn_b200/dist_bipartite.py): the same mirrors restricted to the
 * producer's first `*_mirror_rows` rows (its words) are the all-gather of the word block, and tgcn_spmm with
 * c_scatter_bases stores every output row into the slot buffer of the rank that owns it (the all-to-all of the partial
 * word rows, plain peer stores); the consuming tgcn_spmm adds the slots through raw_in.
 */
int tgcn_peer_push(const void* src, void* const* peer_bases_host, int32_t world, int32_t rank, int64_t bytes,
                   int64_t dst_offset_bytes, void* multicast_base, void* stream);
int tgcn_sum_slots(const float* slots, int32_t n_slots, int64_t slot_stride, int64_t n, float* out, void* stream);

/* Utility kernels the host uses on the path */
int tgcn_count_mask(const uint8_t* mask, int64_t n, int32_t* count_out, void* stream);
int tgcn_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TEXTGCN_B200_H */
