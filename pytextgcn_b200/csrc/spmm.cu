// CSR SpMM with fused epilogue for the symmetric-normalised adjacency A_hat.
//   C[r,:] = dropout(act( sum_k val[k] * B[colidx[k],:] + bias ))      (+ optional P = C @ W_proj)
// Replaces GCNConv.propagate (index_select + mul + scatter_add with atomics) + bias add
// (textgcn/lib/models.py:20, [PyG-1.6.3] message_passing.py) and F.dropout (models.py:23).
// Also used for the backward pass (A_hat is symmetric for TextGCN graphs; in general the
// caller passes the CSR of the transpose).
//
// Work decomposition: one warp per row CHUNK (tgcn_spmm_plan splits rows longer than
// chunk_nnz, so hub word rows with 10^4+ neighbours do not serialise a warp).  Inside a warp
// LPR lanes cover one gathered B row with 16-byte vector loads (VPL vectors per lane) and
// 32/LPR non-zeros are processed side by side; the (col,val) stream is loaded coalesced, 32
// entries per warp load, and broadcast by shuffle.  Split rows write fp32 partial rows to a
// scratch buffer and the chunk that arrives last adds them in slot order: deterministic, no
// floating-point atomics.
#include "spmm_common.cuh"

// Measured alternatives that were removed (profiles/r02_ab_variants.json, 20NG-shape, B200): (col,val) as one 8-byte
// broadcast load instead of two shuffles (+6 %), L1::no_allocate gathers (+36 %), exact lanes-per-row for class-wide
// rows (+1 %), 128-byte row pitch (-2.6 %), shared-memory staged panels fed by per-row cp.async.bulk (2.5-4x slower:
// ~46 clk per 800-byte bulk copy per SM).  Kept: 32-bit row-pitch arithmetic in the gather loop (-4.4 %).

namespace tgcn {

template <typename TB, int LPR, int VPL, int EPI>
__global__ void __launch_bounds__(256, (VPL * Vec<TB>::E <= 8) ? 4 : 2) k_spmm(const SpmmParams p) {
  constexpr int E = Vec<TB>::E;
  constexpr int NZP = 32 / LPR;          // non-zeros processed side by side in a warp
  // gathers kept in flight per lane group: 8 for the wide rows (halves the critical path of a
  // 2048-entry chunk and lifts the L2 gather rate ~7%), 4 for the narrow ones (8 costs occupancy there)
  constexpr int U = (LPR == 32) ? ((64 / (VPL * E)) > 8 ? 8 : ((64 / (VPL * E)) < 2 ? 2 : (64 / (VPL * E)))) : 4;
  extern __shared__ __align__(16) float smem[];
  const int warps_per_block = blockDim.x >> 5;
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float* smem_w = smem;                                   // [F][pad4(n_proj)] (+32 floats slack) when staged
  if (EPI == EPI_PROJ && p.wproj_in_smem) {
    const int Ms = (p.n_proj + 3) & ~3;
    for (int i = threadIdx.x; i < p.F * Ms + 32; i += blockDim.x) {
      const int c = i / Ms, m = i - c * Ms;
      smem_w[i] = (c < p.F && m < p.n_proj) ? p.W_proj[c * p.n_proj + m] : 0.0f;
    }
    __syncthreads();
  }
  const int sub = lane / LPR;                   // which of the NZP side-by-side non-zeros
  const int l = lane % LPR;
  const TB* __restrict__ B = reinterpret_cast<const TB*>(p.B);
  const int F = p.F;
  const uint32_t ldb_bytes = (uint32_t)(p.ldb * (int64_t)sizeof(TB));   // host checks the pitch fits 32 bits
  bool active[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) active[v] = ((l + v * LPR) * E) < F;

  // One chunk per warp, CTAs handed out by the hardware scheduler in list order (longest chunks
  // first).  A persistent grid-stride variant measured 4% slower on B200 (static round-robin
  // cannot absorb the run-time variance of L2-bound chunks), so the simple form stays.
  const int chunk_id = blockIdx.x * warps_per_block + wib;
  if (chunk_id >= p.n_chunks) return;
  {
  const int4 ch = __ldg(p.chunks + chunk_id);   // {row, begin, end, slot}
  float acc[VPL][E];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < E; ++i) acc[v][i] = 0.0f;

  for (int base = ch.y; base < ch.z; base += 32) {
    const int k = base + lane;
    int mc = 0; float mv = 0.0f;
    if (k < ch.z) { mc = __ldg(p.colidx + k); mv = __ldg(p.val + k); }
    const int cnt = min(32, ch.z - base);
    // NZP non-zeros per step, U steps in flight
    for (int j = 0; j < cnt; j += NZP * U) {
      int cc[U]; float vv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int srcl = j + u * NZP + sub;         // may run past cnt: then mv == 0 and mc == 0 (row 0 is valid memory)
        cc[u] = __shfl_sync(0xffffffffu, mc, srcl & 31);
        vv[u] = __shfl_sync(0xffffffffu, mv, srcl & 31);
        if (srcl >= cnt) vv[u] = 0.0f;
      }
      float x[U][VPL][E];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // one IMAD.WIDE.U32 per gather: lane base + column * row pitch in bytes
        const TB* brow = reinterpret_cast<const TB*>(reinterpret_cast<const char*>(B) +
                                                     (uint64_t)(uint32_t)cc[u] * (uint64_t)ldb_bytes);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          if (active[v] && vv[u] != 0.0f) Vec<TB>::load(brow + (l + v * LPR) * E, x[u][v]);
          else {
#pragma unroll
            for (int i = 0; i < E; ++i) x[u][v][i] = 0.0f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int i = 0; i < E; ++i) acc[v][i] = fmaf(vv[u], x[u][v][i], acc[v][i]);
    }
  }
  // fold the NZP side-by-side partial rows into lanes [0, LPR)
#pragma unroll
  for (int o = 16; o >= LPR; o >>= 1)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < E; ++i) acc[v][i] += __shfl_down_sync(0xffffffffu, acc[v][i], o);

  if (ch.w >= 0) {
    // split row: raw fp32 partial into the scratch slot of this chunk
    if (lane < LPR) {
      float* s = p.scratch + (int64_t)ch.w * F;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c0 = (l + v * LPR) * E;
        if (c0 < F) {
#pragma unroll
          for (int q = 0; q < E / 4; ++q)
            *reinterpret_cast<float4*>(s + c0 + 4 * q) = make_float4(acc[v][4 * q], acc[v][4 * q + 1], acc[v][4 * q + 2], acc[v][4 * q + 3]);
        }
      }
    }
    // The chunk that arrives LAST adds the partial rows in slot order (the order is fixed, so the
    // result does not depend on which chunk that is) and runs the epilogue: no second kernel.
    const int sr = __ldg(p.slot_owner + ch.w);
    const int first = __ldg(p.split_rows + 3 * sr + 1), n = __ldg(p.split_rows + 3 * sr + 2);
    __threadfence();
    __syncwarp();                                     // every lane's partial is fenced before lane 0 publishes the arrival
    int arrived = 0;
    if (lane == 0) arrived = atomicAdd(p.split_counters + sr, 1);
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived != n - 1) return;
    __threadfence();
    if (lane == 0) p.split_counters[sr] = 0;          // self-resetting: ready for the next launch
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < E; ++i) acc[v][i] = 0.0f;
    if (lane < LPR) {
      for (int sl = 0; sl < n; ++sl) {
        const float* src = p.scratch + (int64_t)(first + sl) * F;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c0 = (l + v * LPR) * E;
          if (c0 < F) {
#pragma unroll
            for (int q = 0; q < E / 4; ++q) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(src + c0 + 4 * q));   // L2: written by other SMs
              acc[v][4 * q] += t.x; acc[v][4 * q + 1] += t.y; acc[v][4 * q + 2] += t.z; acc[v][4 * q + 3] += t.w;
            }
          }
        }
      }
    }
  }
  if constexpr (E == 4) {
    if (p.raw_in != nullptr && (ch.x - p.c_row_offset) < p.raw_rows && lane < LPR) {
      // contributions computed by the other ranks (bipartite exchange), fixed slot order
      const float* src = p.raw_in + (ch.x - p.c_row_offset) * p.raw_ld;
      for (int k = 0; k < p.n_raw; ++k, src += p.raw_stride) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c0 = (l + v * LPR) * E;
          if (c0 < F) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(src + c0));
            acc[v][0] += t.x; acc[v][1] += t.y; acc[v][2] += t.z; acc[v][3] += t.w;
          }
        }
      }
    }
    if (p.tc_part != nullptr) {
      // hybrid propagation: the dense blocks of this row were multiplied on the tensor cores (spmm_tc.cu); their
      // partial rows are added here in slot order (fixed order: deterministic), once per row (last chunk of a split row)
      const int rk = __ldg(p.tc_rank + ch.x);
      const int s0 = __ldg(p.tc_slot_ptr + (rk >> 7)), s1 = __ldg(p.tc_slot_ptr + (rk >> 7) + 1);
      if (lane < LPR) {
        for (int sl = s0; sl < s1; ++sl) {
          const float* src = p.tc_part + ((int64_t)sl * 128 + (rk & 127)) * p.tc_ld;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int c0 = (l + v * LPR) * E;
            if (c0 < F) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(src + c0));
              acc[v][0] += t.x; acc[v][1] += t.y; acc[v][2] += t.z; acc[v][3] += t.w;
            }
          }
        }
      }
    }
  }
  row_epilogue<LPR, VPL, E, EPI>(p, ch.x, lane, acc, smem_w);
  }
}

// ---- spmm plan: chunk list ------------------------------------------------------------
// Single CTA, three running prefix sums (chunks, partial slots, split rows) over the rows.
__global__ void __launch_bounds__(1024) k_plan(const int32_t* __restrict__ rowptr, int64_t row_begin, int64_t row_end,
                                               int32_t chunk_nnz, int4* __restrict__ chunks, int64_t cap,
                                               int32_t* __restrict__ split_rows, int32_t* __restrict__ slot_owner,
                                               int32_t* __restrict__ counts) {
  __shared__ int s_warp[3][32];
  __shared__ int s_base[3];
  __shared__ int s_maxlen;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { s_base[0] = s_base[1] = s_base[2] = 0; s_maxlen = 0; }
  __syncthreads();
  int my_max = 0;
  for (int64_t r0 = row_begin; r0 < row_end; r0 += blockDim.x) {
    const int64_t r = r0 + tid;
    int len = 0, nch = 0, b = 0;
    if (r < row_end) { b = rowptr[r]; len = rowptr[r + 1] - b; nch = max(1, (len + chunk_nnz - 1) / chunk_nnz); }
    my_max = max(my_max, len);
    int v[3] = {nch, nch > 1 ? nch : 0, nch > 1 ? 1 : 0};
    int inc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int x = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      inc[q] = x;
      if (lane == 31) s_warp[q][wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        int x = (lane < (blockDim.x >> 5)) ? s_warp[q][lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        s_warp[q][lane] = x;   // inclusive over warps
      }
    }
    __syncthreads();
    int excl[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) excl[q] = s_base[q] + (wid ? s_warp[q][wid - 1] : 0) + inc[q] - v[q];
    if (r < row_end) {
      if (nch == 1) {
        if (excl[0] < cap) chunks[excl[0]] = make_int4((int)r, b, b + len, -1);
      } else {
        // even split keeps the chunks of a hub row the same size
        const int per = (len + nch - 1) / nch;
        for (int c = 0; c < nch; ++c) {
          const int cb = b + c * per, ce = min(b + len, cb + per);
          if (excl[0] + c < cap) chunks[excl[0] + c] = make_int4((int)r, cb, ce, excl[1] + c);
          if (slot_owner) slot_owner[excl[1] + c] = excl[2];
        }
        split_rows[3 * excl[2]] = (int)r; split_rows[3 * excl[2] + 1] = excl[1]; split_rows[3 * excl[2] + 2] = nch;
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int nw = blockDim.x >> 5;
#pragma unroll
      for (int q = 0; q < 3; ++q) s_base[q] += s_warp[q][nw - 1];
    }
    __syncthreads();
  }
  atomicMax(&s_maxlen, my_max);
  __syncthreads();
  if (tid == 0) { counts[0] = s_base[0]; counts[1] = s_base[1]; counts[2] = s_base[2]; counts[3] = s_maxlen; }
}

template <typename TB, int LPR, int VPL, int EPI>
static int launch_spmm_t(const SpmmParams& p, cudaStream_t stream) {
  const int threads = 256, wpb = threads / 32;
  size_t smem = 0;
  if (EPI == EPI_PROJ && p.wproj_in_smem) smem = ((size_t)p.F * ((p.n_proj + 3) & ~3) + 32) * sizeof(float);
  if (smem > 48 * 1024) {
    TGCN_CUDA(cudaFuncSetAttribute(k_spmm<TB, LPR, VPL, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (p.n_chunks > 0) {
    const int64_t grid = cdiv(p.n_chunks, wpb);
    k_spmm<TB, LPR, VPL, EPI><<<(unsigned)grid, threads, smem, stream>>>(p);
    TGCN_LAUNCH_CHECK();
  }
  return TGCN_OK;
}

template <typename TB, int LPR, int VPL>
static int launch_spmm(const SpmmParams& p, cudaStream_t stream) {
  if (p.ad_p) {
    if constexpr (std::is_same<TB, float>::value) return launch_spmm_t<TB, LPR, VPL, EPI_ADAM>(p, stream);
    set_error("spmm: the fused Adam epilogue needs an fp32 operand");
    return TGCN_EINVAL;
  }
  if (p.P) return launch_spmm_t<TB, LPR, VPL, EPI_PROJ>(p, stream);
  return launch_spmm_t<TB, LPR, VPL, EPI_PLAIN>(p, stream);
}

template <typename TB>
static int dispatch_spmm(const SpmmParams& p, cudaStream_t stream) {
  constexpr int E = Vec<TB>::E;
  const int nvec = (p.F + E - 1) / E;   // 16-byte vectors per dense row
  if (nvec <= 4) return launch_spmm<TB, 4, 1>(p, stream);
  if (nvec <= 8) return launch_spmm<TB, 8, 1>(p, stream);
  if (nvec <= 16) return launch_spmm<TB, 16, 1>(p, stream);
  if (nvec <= 32) return launch_spmm<TB, 32, 1>(p, stream);
  if (nvec <= 64) return launch_spmm<TB, 32, 2>(p, stream);
  if (nvec <= 128) return launch_spmm<TB, 32, 4>(p, stream);
  set_error("spmm: feature width %d too large (max %d)", p.F, 128 * E);
  return TGCN_EINVAL;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_spmm_plan_workspace_bytes(int64_t n_rows, size_t* bytes_out) {
  TGCN_CHECK_ARG(bytes_out != nullptr, "bytes_out is null");
  (void)n_rows;
  *bytes_out = 0;   // the plan kernel keeps its running sums in shared memory
  return TGCN_OK;
}

extern "C" int tgcn_spmm_plan(const int32_t* rowptr, int64_t row_begin, int64_t row_end, int32_t chunk_nnz,
                              int32_t* chunks, int64_t chunk_capacity, int32_t* split_rows, int32_t* slot_owner,
                              int32_t* counts_out, void* workspace, size_t workspace_bytes, void* stream_) {
  (void)workspace; (void)workspace_bytes;
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(rowptr && chunks && split_rows && counts_out, "spmm_plan: null pointer");
  TGCN_CHECK_ARG(row_begin >= 0 && row_end >= row_begin, "spmm_plan: bad row range");
  TGCN_CHECK_ARG(chunk_nnz >= 32, "spmm_plan: chunk_nnz must be >= 32");
  k_plan<<<1, 1024, 0, stream>>>(rowptr, row_begin, row_end, chunk_nnz, reinterpret_cast<int4*>(chunks), chunk_capacity,
                                 split_rows, slot_owner, counts_out);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_spmm(const tgcn_spmm_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(a != nullptr, "spmm: args null");
  TGCN_CHECK_ARG(a->rowptr && a->colidx && a->val && a->chunks, "spmm: CSR/plan pointer null");
  TGCN_CHECK_ARG(a->B != nullptr, "spmm: B null");
  TGCN_CHECK_ARG(a->C != nullptr || a->P != nullptr || a->adam_param != nullptr, "spmm: no output requested");
  TGCN_CHECK_ARG(a->F > 0, "spmm: F must be > 0");
  TGCN_CHECK_ARG(a->b_dtype == TGCN_F32 || a->b_dtype == TGCN_BF16, "spmm: bad b_dtype");
  TGCN_CHECK_ARG(a->c_dtype == TGCN_F32 || a->c_dtype == TGCN_BF16, "spmm: bad c_dtype");
  const int eb = a->b_dtype == TGCN_F32 ? 4 : 8;
  TGCN_CHECK_ARG(a->F % eb == 0, "spmm: F (%d) must be a multiple of %d for this dtype (pad the operand)", a->F, eb);
  TGCN_CHECK_ARG(a->ldb % eb == 0 && ((uintptr_t)a->B % 16) == 0, "spmm: B must be 16-byte aligned with ldb %% %d == 0", eb);
  if (a->C) {
    const int ec = a->c_dtype == TGCN_F32 ? 4 : 2;
    TGCN_CHECK_ARG(a->ldc % ec == 0 && ((uintptr_t)a->C % 16) == 0, "spmm: C must be 16-byte aligned with ldc %% %d == 0", ec);
    TGCN_CHECK_ARG(a->ldc >= a->F, "spmm: ldc < F");
  }
  TGCN_CHECK_ARG(a->n_split_rows == 0 || (a->split_rows && a->scratch && a->slot_owner && a->split_counters),
                 "spmm: split rows need split_rows, scratch, slot_owner and split_counters");
  TGCN_CHECK_ARG(a->drop_mode >= TGCN_DROP_NONE && a->drop_mode <= TGCN_DROP_PHILOX, "spmm: bad drop_mode");
  TGCN_CHECK_ARG(a->drop_mode == TGCN_DROP_NONE || (a->drop_p >= 0.0f && a->drop_p < 1.0f), "spmm: dropout p must be in [0,1)");
  TGCN_CHECK_ARG(a->drop_mode != TGCN_DROP_MASK || a->keep_mask, "spmm: TGCN_DROP_MASK needs keep_mask");
  TGCN_CHECK_ARG(a->act == TGCN_ACT_NONE || a->act == TGCN_ACT_RELU, "spmm: bad act");
  TGCN_CHECK_ARG(a->P == nullptr || (a->W_proj && a->n_proj > 0 && a->ldp >= a->n_proj), "spmm: bad projection arguments");

  SpmmParams p;
  if (int rc = fill_spmm_params(a, &p)) return rc;
  TGCN_CHECK_ARG(a->ldb * (a->b_dtype == TGCN_F32 ? 4 : 2) < (int64_t)1 << 31, "spmm: row pitch must fit 32 bits");
  if (a->b_dtype == TGCN_F32) return dispatch_spmm<float>(p, stream);
  return dispatch_spmm<__nv_bfloat16>(p, stream);
}
