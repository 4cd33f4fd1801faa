// Panel SpMM that stages the gathered rows of the dense operand in shared memory (sm_100a).
//
// Same operation and epilogues as tgcn_spmm (GCNConv.propagate + bias + F.dropout,
// textgcn/lib/models.py:20,23; fused Adam for dW1, flat_amazon.py:106), different data movement.
// tgcn_spmm moves one 4*F-byte operand row from L2 into an SM for EVERY non-zero (17.2 GB per
// launch at 20NG-shape) and is bound by that L2->SM fill path.  Here one CTA owns a
// PANEL of consecutive chunks of the length-sorted chunk list (tgcn_spmm_plan: one chunk = one row
// or one piece of a hub row) and walks the sorted UNION of the columns its chunks touch, tile by
// tile:
//   * producer warps copy the operand rows of a tile (tile_cols rows of 4*F bytes) from L2 into a
//     ring of shared-memory stages -- one cp.async.bulk per row (TMA unit, complete_tx byte counting)
//     or 16-byte cp.async by all lanes (LDGSTS, L1 bypassed) -- signalled by one mbarrier per stage;
//   * every consumer warp owns rows_per_warp chunks; for each tile it reads the (slot, value) pairs
//     of its chunks that fall into the tile from its own pre-sorted entry stream (one broadcast
//     8-byte load per entry), reads the operand row from shared memory (LDS.128) and accumulates in
//     registers; when done with a tile it arrives on the stage's "empty" mbarrier.
// An operand row needed by k chunks of the panel crosses L2->SM once instead of k times (popular
// word columns are shared by many rows: tools/staged_reuse.py reports the ratio for a graph), and
// the per-non-zero traffic moves to the shared-memory read port (128 B/clk/SM vs ~64 B/clk/SM fills).
// Split rows and all epilogues go through finish_row() exactly as in tgcn_spmm: deterministic, no
// floating-point atomics.
//
// Plan layout (built once per graph by pytextgcn_b200/staged_plan.py; emulated on the CPU in
// tests/test_staged_plan.py):
//   panel p      = chunks [p*R, (p+1)*R) of the chunk list, R = warps_per_panel * rows_per_warp;
//                  consumer warp w owns chunks (p*W + w)*rows_per_warp + {0 .. rows_per_warp-1}
//   ucols        = for every panel the ascending distinct column ids of its chunks;
//                  panel_ucol_ptr[p] .. panel_ucol_ptr[p+1]; tile t = entries [t*tile_cols, ...)
//   stream       = int2 entries; warp_stream_ptr[p*W + w] = start of that warp's stream.  Per tile:
//                  a header {n0, n1} (number of entries of the warp's first / second chunk in this
//                  tile) followed by n0 + n1 entries {slot within the tile, value bits}.
#include "spmm_common.cuh"

namespace tgcn {

struct StagedParams {
  const int32_t* __restrict__ panel_ucol_ptr; const int32_t* __restrict__ ucols;
  const int64_t* __restrict__ warp_stream_ptr; const int2* __restrict__ stream;
  int32_t n_panels, warps_per_panel, n_producers, tile_cols, n_stages;
  uint32_t row_bytes, stage_bytes;
};

constexpr int STAGED_MAX_TILE_COLS = 128;   // producer lanes prefetch up to 4 column ids each

// ---- mbarrier / bulk-copy primitives (PTX; shared::cta addresses as 32-bit) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// L2 -> shared memory copy of `bytes` (multiple of 16, both addresses 16-byte aligned) by the TMA unit;
// completion is reported to `bar` as `bytes` of its transaction count.
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Asynchronous 16-byte copy L2 -> shared memory (LDGSTS, L1 bypassed) and its completion hook: the
// mbarrier receives one arrival (already counted in its expected total) once all earlier cp.async
// operations of the executing thread have landed.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// (c0, c1) += a * (x0, x1) as ONE packed instruction (FFMA2, sm_100): same IEEE fma per element, half the
// issue slots of two scalar FFMAs -- the consumer loop is issue-sensitive (about 20 instructions per non-zero).
__device__ __forceinline__ void fma2(float& c0, float& c1, float a, float x0, float x1) {
  unsigned long long c, x;
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("{\n .reg .b64 av;\n mov.b64 av, {%2, %2};\n fma.rn.f32x2 %0, av, %1, %0;\n}" : "+l"(c) : "l"(x), "f"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(c));
}

constexpr int PROD_BULK = 0;     // one cp.async.bulk (TMA unit) per operand row, issued lane by lane
constexpr int PROD_LDGSTS = 1;   // the 32 lanes of a producer warp copy a row with 16-byte cp.async

// One CTA per panel: warps [0, W) consume, warps [W, W+NP) produce.  VPL = 16-byte vectors per lane
// of one operand row (F <= 128: 1, F <= 256: 2); RPW = chunks per consumer warp (1 or 2).
template <int VPL, int RPW, int EPI, int PROD>
__global__ void __launch_bounds__(1024, 1) k_spmm_staged(const SpmmParams p, const StagedParams sp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int W = sp.warps_per_panel, NP = sp.n_producers;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = sp.n_stages;
  unsigned char* stages = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * sp.stage_bytes);   // full[S], empty[S]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S);
  if (threadIdx.x == 0) {
    const uint32_t full_count = (PROD == PROD_BULK) ? (uint32_t)NP : (uint32_t)NP * 32u;
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, full_count); mbar_init(empty0 + 8 * s, (uint32_t)W); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int panel = blockIdx.x;
  const int u0 = __ldg(sp.panel_ucol_ptr + panel), u1 = __ldg(sp.panel_ucol_ptr + panel + 1);
  const int KC = sp.tile_cols;
  const int n_tiles = (u1 - u0 + KC - 1) / KC;

  if (wib >= W) {
    // ---------------- producers ----------------
    // Producer pj owns the 32-slot groups pj, pj + NP, ... of every tile (slot = position of the column
    // in the tile); the column ids of the NEXT tile are loaded before waiting for the current stage.
    constexpr int MAXG = STAGED_MAX_TILE_COLS / 32;
    const int pj = wib - W;
    const float* __restrict__ B = reinterpret_cast<const float*>(p.B);
    int cnext[MAXG];
#pragma unroll
    for (int i = 0; i < MAXG; ++i) {
      const int s = (i * NP + pj) * 32 + lane;
      cnext[i] = (s < KC && u0 + s < u1) ? __ldg(sp.ucols + u0 + s) : 0;
    }
    int stage = 0; uint32_t parity = 0;
    for (int t = 0; t < n_tiles; ++t) {
      int ccur[MAXG];
      const int base_next = u0 + (t + 1) * KC;
#pragma unroll
      for (int i = 0; i < MAXG; ++i) {
        ccur[i] = cnext[i];
        const int s = (i * NP + pj) * 32 + lane;
        cnext[i] = (s < KC && base_next + s < u1) ? __ldg(sp.ucols + base_next + s) : 0;
      }
      const int nc = min(KC, u1 - u0 - t * KC);
      const uint32_t full = full0 + 8 * stage;
      const uint32_t dst0 = smem_u32(stages + (size_t)stage * sp.stage_bytes);
      mbar_wait(empty0 + 8 * stage, parity ^ 1u);          // all consumers left the tile that used this stage
      if (PROD == PROD_BULK) {
        if (lane == 0) {
          int mine = 0;                                     // operand rows this producer copies into the tile
#pragma unroll
          for (int i = 0; i < MAXG; ++i) mine += max(0, min(32, nc - (i * NP + pj) * 32));
          mbar_arrive_expect_tx(full, (uint32_t)mine * sp.row_bytes);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {
          const int s = (i * NP + pj) * 32 + lane;
          if (s < nc) bulk_copy_g2s(dst0 + (uint32_t)s * sp.row_bytes, B + (int64_t)ccur[i] * p.ldb, sp.row_bytes, full);
        }
      } else {
        const int nvec = (int)(sp.row_bytes >> 4);          // 16-byte pieces per operand row
#pragma unroll
        for (int i = 0; i < MAXG; ++i) {
          const int g0 = (i * NP + pj) * 32;
          const int ng = min(32, nc - g0);
          for (int k = 0; k < ng; ++k) {
            const int col = __shfl_sync(0xffffffffu, ccur[i], k);
            const float* src = B + (int64_t)col * p.ldb;
            const uint32_t dst = dst0 + (uint32_t)(g0 + k) * sp.row_bytes;
            for (int c = lane; c < nvec; c += 32) cp_async16(dst + 16u * c, src + 4 * c);
          }
        }
        cp_async_arrive_noinc(full);                        // every lane: fires when its own copies have landed
      }
      if (++stage == S) { stage = 0; parity ^= 1u; }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int F = p.F;
  int4 ch[RPW];
  bool have[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t vid = ((int64_t)panel * W + wib) * RPW + r;
    have[r] = vid < p.n_chunks;
    ch[r] = have[r] ? __ldg(p.chunks + vid) : make_int4(0, 0, 0, -1);
  }
  float acc[RPW][VPL][4];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][v][i] = 0.0f;
  // Byte offset of this lane's v-th 16-byte piece inside an operand row.  Lanes past the end of the
  // row re-read a piece an active lane of the same instruction reads (a broadcast, no extra
  // shared-memory wavefront); what they accumulate is never stored (finish_row checks c0 < F).
  uint32_t loff[VPL];
  {
    const int nvec = F >> 2;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int nact = max(1, min(32, nvec - 32 * v));
      loff[v] = (uint32_t)(min(32 * v, nvec - 1) + (lane % nact)) * 16u;
    }
  }

  // Entry stream of this warp: every lane reads the same 8-byte entry (one broadcast request, served
  // from L1 after the first touch of a 128-byte line); the lines of the next tiles are prefetched.
  const int2* __restrict__ st = sp.stream + __ldg(sp.warp_stream_ptr + (int64_t)panel * W + wib);
  int q = 0;
  int stage = 0; uint32_t parity = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const int2 hdr = __ldg(st + q); ++q;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(st + q + 32));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(st + q + 48));
    mbar_wait(full0 + 8 * stage, parity);                 // the tile's operand rows have landed
    const unsigned char* tile[VPL];                       // this lane's pieces of slot 0 of the stage
#pragma unroll
    for (int v = 0; v < VPL; ++v) tile[v] = stages + (size_t)stage * sp.stage_bytes + loff[v];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int cnt = (r == 0) ? hdr.x : hdr.y;
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const int2 e = __ldg(st + q + j);                 // {slot within the tile, value bits}
        const float a = __int_as_float(e.y);
        const uint32_t row = (uint32_t)e.x * sp.row_bytes;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const float4 x = *reinterpret_cast<const float4*>(tile[v] + row);
          fma2(acc[r][v][0], acc[r][v][1], a, x.x, x.y);
          fma2(acc[r][v][2], acc[r][v][3], a, x.z, x.w);
        }
      }
      q += cnt;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * stage);       // this warp is done reading the stage
    if (++stage == S) { stage = 0; parity ^= 1u; }
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r)
    if (have[r]) finish_row<32, VPL, 4, EPI>(p, ch[r], lane, acc[r], nullptr);
}

template <int VPL, int RPW, int EPI, int PROD>
static int launch_staged_t(const SpmmParams& p, StagedParams sp, cudaStream_t stream) {
  static int max_optin = 0;                               // same for every device of a B200 box
  if (max_optin == 0) {
    int dev = 0;
    TGCN_CUDA(cudaGetDevice(&dev));
    TGCN_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  sp.row_bytes = (uint32_t)p.F * 4u;
  sp.stage_bytes = (uint32_t)sp.tile_cols * sp.row_bytes;
  const size_t bar_bytes = 2 * 8 * 8;                      // full[] + empty[], up to 8 stages
  int S = (int)(((size_t)max_optin - bar_bytes - 128) / sp.stage_bytes);
  if (S > 8) S = 8;
  TGCN_CHECK_ARG(S >= 2, "spmm_staged: a tile of %d columns x %u bytes does not fit twice in %d bytes of shared memory",
                 sp.tile_cols, sp.row_bytes, max_optin);
  sp.n_stages = S;
  const size_t smem = (size_t)S * sp.stage_bytes + bar_bytes;
  TGCN_CUDA(cudaFuncSetAttribute(k_spmm_staged<VPL, RPW, EPI, PROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = (sp.warps_per_panel + sp.n_producers) * 32;
  k_spmm_staged<VPL, RPW, EPI, PROD><<<(unsigned)sp.n_panels, threads, smem, stream>>>(p, sp);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

template <int VPL, int RPW>
static int launch_staged(const SpmmParams& p, const StagedParams& sp, int prod, cudaStream_t stream) {
  if (p.ad_p) {
    return prod == PROD_BULK ? launch_staged_t<VPL, RPW, EPI_ADAM, PROD_BULK>(p, sp, stream)
                             : launch_staged_t<VPL, RPW, EPI_ADAM, PROD_LDGSTS>(p, sp, stream);
  }
  return prod == PROD_BULK ? launch_staged_t<VPL, RPW, EPI_PLAIN, PROD_BULK>(p, sp, stream)
                           : launch_staged_t<VPL, RPW, EPI_PLAIN, PROD_LDGSTS>(p, sp, stream);
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_spmm_staged(const tgcn_spmm_args* a, const tgcn_staged_plan* pl, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(a != nullptr && pl != nullptr, "spmm_staged: args or plan null");
  TGCN_CHECK_ARG(a->chunks && a->n_chunks >= 0, "spmm_staged: chunk list null");
  TGCN_CHECK_ARG(pl->panel_ucol_ptr && pl->ucols && pl->warp_stream_ptr && pl->stream, "spmm_staged: plan pointer null");
  TGCN_CHECK_ARG(pl->rows_per_warp == 1 || pl->rows_per_warp == 2, "spmm_staged: rows_per_warp must be 1 or 2");
  TGCN_CHECK_ARG(pl->n_producers >= 1 && pl->n_producers <= 4, "spmm_staged: n_producers must be in [1, 4]");
  TGCN_CHECK_ARG(pl->warps_per_panel >= 1 && pl->warps_per_panel + pl->n_producers <= 32,
                 "spmm_staged: warps_per_panel + n_producers must be at most 32");
  TGCN_CHECK_ARG(pl->producer_mode == PROD_BULK || pl->producer_mode == PROD_LDGSTS, "spmm_staged: bad producer_mode");
  TGCN_CHECK_ARG(pl->tile_cols >= 1 && pl->tile_cols <= STAGED_MAX_TILE_COLS, "spmm_staged: tile_cols must be in [1, %d]", STAGED_MAX_TILE_COLS);
  TGCN_CHECK_ARG((int64_t)pl->n_panels * pl->warps_per_panel * pl->rows_per_warp >= a->n_chunks &&
                 (int64_t)(pl->n_panels - 1) * pl->warps_per_panel * pl->rows_per_warp < (a->n_chunks > 0 ? a->n_chunks : 1),
                 "spmm_staged: the plan (%d panels x %d warps x %d rows) does not match the chunk list (%d chunks)",
                 pl->n_panels, pl->warps_per_panel, pl->rows_per_warp, a->n_chunks);
  TGCN_CHECK_ARG(a->B != nullptr && a->b_dtype == TGCN_F32, "spmm_staged: B must be an fp32 operand");
  TGCN_CHECK_ARG(a->C != nullptr || a->adam_param != nullptr, "spmm_staged: no output requested");
  TGCN_CHECK_ARG(a->P == nullptr, "spmm_staged: the fused projection is not available in this kernel");
  TGCN_CHECK_ARG(a->c_dtype == TGCN_F32 || a->c_dtype == TGCN_BF16, "spmm_staged: bad c_dtype");
  TGCN_CHECK_ARG(a->F > 0 && a->F % 4 == 0 && a->F <= 256, "spmm_staged: F (%d) must be a multiple of 4, at most 256", a->F);
  TGCN_CHECK_ARG(a->ldb % 4 == 0 && ((uintptr_t)a->B % 16) == 0, "spmm_staged: B must be 16-byte aligned with ldb %% 4 == 0");
  if (a->C) {
    const int ec = a->c_dtype == TGCN_F32 ? 4 : 2;
    TGCN_CHECK_ARG(a->ldc % ec == 0 && ((uintptr_t)a->C % 16) == 0 && a->ldc >= a->F, "spmm_staged: bad C alignment or ldc");
  }
  TGCN_CHECK_ARG(a->n_split_rows == 0 || (a->split_rows && a->scratch && a->slot_owner && a->split_counters),
                 "spmm_staged: split rows need split_rows, scratch, slot_owner and split_counters");
  TGCN_CHECK_ARG(a->drop_mode >= TGCN_DROP_NONE && a->drop_mode <= TGCN_DROP_PHILOX, "spmm_staged: bad drop_mode");
  TGCN_CHECK_ARG(a->drop_mode == TGCN_DROP_NONE || (a->drop_p >= 0.0f && a->drop_p < 1.0f), "spmm_staged: dropout p must be in [0,1)");
  TGCN_CHECK_ARG(a->drop_mode != TGCN_DROP_MASK || a->keep_mask, "spmm_staged: TGCN_DROP_MASK needs keep_mask");
  TGCN_CHECK_ARG(a->act == TGCN_ACT_NONE || a->act == TGCN_ACT_RELU, "spmm_staged: bad act");
  if (a->n_chunks == 0 || pl->n_panels == 0) return TGCN_OK;

  SpmmParams p;
  if (int rc = fill_spmm_params(a, &p)) return rc;
  StagedParams sp;
  sp.panel_ucol_ptr = pl->panel_ucol_ptr; sp.ucols = pl->ucols;
  sp.warp_stream_ptr = pl->warp_stream_ptr; sp.stream = reinterpret_cast<const int2*>(pl->stream);
  sp.n_panels = pl->n_panels; sp.warps_per_panel = pl->warps_per_panel; sp.n_producers = pl->n_producers; sp.tile_cols = pl->tile_cols;
  sp.n_stages = 0; sp.row_bytes = 0; sp.stage_bytes = 0;
  const bool wide = a->F > 128;
  const int prod = pl->producer_mode;
  if (pl->rows_per_warp == 1) return wide ? launch_staged<2, 1>(p, sp, prod, stream) : launch_staged<1, 1>(p, sp, prod, stream);
  return wide ? launch_staged<2, 2>(p, sp, prod, stream) : launch_staged<1, 2>(p, sp, prod, stream);
}
