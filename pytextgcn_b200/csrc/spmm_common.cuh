// Pieces of the SpMM kernel (spmm.cu): launch parameters, 16-byte operand loads and the fused row epilogue.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace tgcn {

struct SpmmParams {
  const int32_t* __restrict__ rowptr; const int32_t* __restrict__ colidx; const float* __restrict__ val;
  const int4* __restrict__ chunks; int32_t n_chunks;
  const int32_t* __restrict__ split_rows; int32_t n_split_rows;
  const int32_t* __restrict__ slot_owner; int32_t* split_counters;   // the last chunk of a split row reduces it in-kernel
  float* scratch;
  const void* __restrict__ B; int64_t ldb;
  void* C; int64_t ldc; int32_t c_dtype;
  int32_t F;
  int64_t c_row_offset;
  const float* __restrict__ bias; int32_t bias_len;
  int32_t act;
  int32_t drop_mode; float drop_p; float drop_scale; const uint8_t* __restrict__ keep_mask; int64_t ldmask;
  uint64_t philox_seed; uint64_t philox_offset; const int64_t* __restrict__ philox_offset_dev; int64_t philox_row_offset;
  const float* __restrict__ W_proj; int32_t n_proj; float* P; int64_t ldp;
  int32_t wproj_in_smem;
  // fused Adam/AMSGrad on the finished row (backward of layer 1 with X = I: the row IS dW1[row])
  float* ad_p; float* ad_m; float* ad_v; float* ad_x; int64_t ad_ld; const float* __restrict__ ad_hyp;
  float ad_b1, ad_b2, ad_eps; float* ad_mirror;
  // dense-tile partial results of tgcn_spmm_tc, added to the row before the epilogue (hybrid propagation)
  const float* __restrict__ tc_part; int64_t tc_ld; const int32_t* __restrict__ tc_rank; const int32_t* __restrict__ tc_slot_ptr;
  // partial rows of other ranks (bipartite exchange), added in slot order to local rows < raw_rows before the epilogue
  const float* __restrict__ raw_in; int64_t raw_ld, raw_stride, raw_rows; int32_t n_raw;
  // word-block exchange: rows the Adam mirror covers; all-to-all of the output rows through peer-mapped bases
  int64_t ad_mirror_rows;
  const uint64_t* __restrict__ c_scatter; int32_t c_scatter_rows; int64_t c_scatter_row0;
};

// ---- loads of 16 bytes of the dense operand -> 4 (fp32) or 8 (bf16) floats ----
template <typename TB> struct Vec;
template <> struct Vec<float> {
  static constexpr int E = 4;
  __device__ __forceinline__ static void load(const float* p, float (&x)[4]) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int E = 8;
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&x)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = __uint_as_float(w[i] << 16);
      x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// Epilogue on one finished row held as: lane l (< LPR) owns elements
// [ (l + v*LPR)*E , +E ) for v < VPL.  All 32 lanes call this (lanes >= LPR idle in the
// element part but take part in the projection).
constexpr int EPI_PLAIN = 0, EPI_PROJ = 1, EPI_ADAM = 2;

template <int LPR, int VPL, int E, int EPI>
__device__ __forceinline__ void row_epilogue(const SpmmParams& p, int64_t row, int lane, float (&acc)[VPL][E],
                                             const float* smem_w) {
  const int F = p.F;
  const int64_t lrow = row - p.c_row_offset;
  const uint64_t ph_off = p.philox_offset + (p.philox_offset_dev ? (uint64_t)__ldg(p.philox_offset_dev) : 0ull);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c0 = (lane + v * LPR) * E;
    float (&z)[E] = acc[v];
    if (lane < LPR && c0 < F) {
      if (p.bias) {
#pragma unroll
        for (int i = 0; i < E; ++i) if (c0 + i < p.bias_len) z[i] += __ldg(p.bias + c0 + i);
      }
      if (p.act == TGCN_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < E; ++i) z[i] = fmaxf(z[i], 0.0f);
      }
      if (p.drop_mode == TGCN_DROP_MASK) {
        const uint8_t* m = p.keep_mask + lrow * p.ldmask + c0;
#pragma unroll
        for (int i = 0; i < E; ++i) z[i] = m[i] ? z[i] * p.drop_scale : 0.0f;
      } else if (p.drop_mode == TGCN_DROP_PHILOX) {
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
          uint64_t e4 = ((uint64_t)(row + p.philox_row_offset) * (uint64_t)F + (uint64_t)(c0 + 4 * q)) >> 2;
          uint4 r = philox_quad(e4, p.philox_seed, ph_off);
          z[4 * q + 0] = (u01(r.x) >= p.drop_p) ? z[4 * q + 0] * p.drop_scale : 0.0f;
          z[4 * q + 1] = (u01(r.y) >= p.drop_p) ? z[4 * q + 1] * p.drop_scale : 0.0f;
          z[4 * q + 2] = (u01(r.z) >= p.drop_p) ? z[4 * q + 2] * p.drop_scale : 0.0f;
          z[4 * q + 3] = (u01(r.w) >= p.drop_p) ? z[4 * q + 3] * p.drop_scale : 0.0f;
        }
      }
      if (p.C) {
        if (p.c_dtype == TGCN_F32) {
          float* c = reinterpret_cast<float*>(p.C) + lrow * p.ldc + c0;
          if (p.c_scatter) {     // this row belongs to another rank's slot buffer (peer store over NVLink)
            const int d = (int)lrow / p.c_scatter_rows;
            c = reinterpret_cast<float*>(__ldg(p.c_scatter + d)) + (p.c_scatter_row0 + ((int)lrow - d * p.c_scatter_rows)) * p.ldc + c0;
          }
#pragma unroll
          for (int q = 0; q < E / 4; ++q)
            *reinterpret_cast<float4*>(c + 4 * q) = make_float4(z[4 * q], z[4 * q + 1], z[4 * q + 2], z[4 * q + 3]);
        } else {
          __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(p.C) + lrow * p.ldc + c0;
#pragma unroll
          for (int q = 0; q < E / 2; ++q)
            *reinterpret_cast<__nv_bfloat162*>(c + 2 * q) = __floats2bfloat162_rn(z[2 * q], z[2 * q + 1]);
          // the projection consumes exactly what the next layer would read back
#pragma unroll
          for (int i = 0; i < E; ++i) z[i] = __bfloat162float(__float2bfloat16_rn(z[i]));
        }
      }
      if constexpr (EPI == EPI_ADAM) {
        // z = dW1[row, c0 .. c0+E): Adam/AMSGrad right here, while the gradient row is in registers
        // (torch.optim.Adam arithmetic, same as k_adam); the gradient itself is stored only if C != NULL.
        const float step_size = __ldg(p.ad_hyp), bc2s = __ldg(p.ad_hyp + 1);
        const int64_t off = lrow * p.ad_ld + c0;
#pragma unroll
        for (int q = 0; q < E / 4; ++q) {
          float4 P4 = *reinterpret_cast<const float4*>(p.ad_p + off + 4 * q);
          float4 M4 = *reinterpret_cast<const float4*>(p.ad_m + off + 4 * q);
          float4 V4 = *reinterpret_cast<const float4*>(p.ad_v + off + 4 * q);
          float4 X4 = p.ad_x ? *reinterpret_cast<const float4*>(p.ad_x + off + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
          float* pp = &P4.x; float* mm = &M4.x; float* vv = &V4.x; float* xx = &X4.x;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float g = z[4 * q + k];
            mm[k] = mm[k] * p.ad_b1 + (1.0f - p.ad_b1) * g;
            vv[k] = vv[k] * p.ad_b2 + (1.0f - p.ad_b2) * (g * g);
            float vh = vv[k];
            if (p.ad_x) { xx[k] = fmaxf(xx[k], vv[k]); vh = xx[k]; }
            pp[k] = pp[k] - step_size * (mm[k] / (sqrtf(vh) / bc2s + p.ad_eps));
          }
          *reinterpret_cast<float4*>(p.ad_p + off + 4 * q) = P4;
          *reinterpret_cast<float4*>(p.ad_m + off + 4 * q) = M4;
          *reinterpret_cast<float4*>(p.ad_v + off + 4 * q) = V4;
          if (p.ad_x) *reinterpret_cast<float4*>(p.ad_x + off + 4 * q) = X4;
          if (p.ad_mirror && lrow < p.ad_mirror_rows) multimem_st_v4(p.ad_mirror + off + 4 * q, P4.x, P4.y, P4.z, P4.w);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) z[i] = 0.0f;     // lanes outside the row contribute nothing to the projection
    }
  }
  if constexpr (EPI == EPI_PROJ) {
    // P[row, m] = sum_c z[c] * W[c, m].  Every lane multiplies ITS columns (still in registers) into
    // 16 outputs at a time; a transposing butterfly (16 shuffles) then leaves output m0 + lane/2 in
    // every lane pair.
    const int M = p.n_proj;
    const int Ms = p.wproj_in_smem ? ((M + 3) & ~3) : M;      // row stride of W (padded in shared memory)
    const float* W = p.wproj_in_smem ? smem_w : p.W_proj;
    for (int m0 = 0; m0 < M; m0 += 16) {
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = 0.0f;
      const int mc = min(16, M - m0);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c0 = (lane + v * LPR) * E;
        if (lane < LPR && c0 < F) {
#pragma unroll
          for (int i = 0; i < E; ++i) {
            const float zi = acc[v][i];
            const float* wr = W + (int64_t)(c0 + i) * Ms + m0;
            if (p.wproj_in_smem) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                if (j < mc) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wr + j);
                  o[j] = fmaf(zi, w4.x, o[j]); o[j + 1] = fmaf(zi, w4.y, o[j + 1]);
                  o[j + 2] = fmaf(zi, w4.z, o[j + 2]); o[j + 3] = fmaf(zi, w4.w, o[j + 3]);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (j < mc) o[j] = fmaf(zi, __ldg(wr + j), o[j]);
            }
          }
        }
      }
#pragma unroll
      for (int off = 16; off >= 2; off >>= 1) {      // 16 values -> 1 value per lane (8+4+2+1 shuffles)
        const bool upper = (lane & off) != 0;
        const int half = off >> 1;
#pragma unroll
        for (int j = 0; j < half; ++j) {
          const float send = upper ? o[j] : o[j + half];
          const float keep = upper ? o[j + half] : o[j];
          o[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      o[0] += __shfl_xor_sync(0xffffffffu, o[0], 1);
      const int m = lane >> 1;
      if ((lane & 1) == 0 && m < mc) p.P[lrow * p.ldp + m0 + m] = o[0];
    }
  }
}

// tgcn_spmm_args -> kernel parameters; checks the fused-Adam arguments
static inline int fill_spmm_params(const tgcn_spmm_args* a, SpmmParams* p) {
  p->rowptr = a->rowptr; p->colidx = a->colidx; p->val = a->val;
  p->chunks = reinterpret_cast<const int4*>(a->chunks); p->n_chunks = a->n_chunks;
  p->split_rows = a->split_rows; p->n_split_rows = a->n_split_rows;
  p->slot_owner = a->slot_owner; p->split_counters = a->split_counters;
  p->scratch = a->scratch;
  p->B = a->B; p->ldb = a->ldb;
  p->C = a->C; p->ldc = a->ldc; p->c_dtype = a->c_dtype;
  p->F = a->F; p->c_row_offset = a->c_row_offset;
  p->bias = a->bias; p->bias_len = (a->bias_len > 0 && a->bias_len <= a->F) ? a->bias_len : a->F; p->act = a->act;
  p->drop_mode = (a->drop_mode != TGCN_DROP_NONE && a->drop_p > 0.0f) ? a->drop_mode : TGCN_DROP_NONE;
  p->drop_p = a->drop_p; p->drop_scale = 1.0f / (1.0f - a->drop_p);
  p->keep_mask = a->keep_mask; p->ldmask = a->ldmask;
  p->philox_seed = a->philox_seed; p->philox_offset = a->philox_offset; p->philox_offset_dev = a->philox_offset_dev; p->philox_row_offset = a->philox_row_offset;
  p->W_proj = a->W_proj; p->n_proj = a->n_proj; p->P = a->P; p->ldp = a->ldp;
  p->ad_p = a->adam_param; p->ad_m = a->adam_exp_avg; p->ad_v = a->adam_exp_avg_sq; p->ad_x = a->adam_max_exp_avg_sq;
  p->ad_ld = a->adam_ld; p->ad_hyp = a->adam_hyper_dev; p->ad_b1 = a->adam_beta1; p->ad_b2 = a->adam_beta2; p->ad_eps = a->adam_eps;
  p->ad_mirror = (float*)a->adam_param_mirror_mc;
  p->tc_part = a->tc_part; p->tc_ld = a->tc_ld; p->tc_rank = a->tc_rank; p->tc_slot_ptr = a->tc_slot_ptr;
  p->ad_mirror_rows = a->adam_mirror_rows > 0 ? a->adam_mirror_rows : INT64_MAX;
  p->c_scatter = a->c_scatter_bases; p->c_scatter_rows = (int32_t)a->c_scatter_rows; p->c_scatter_row0 = a->c_scatter_row0;
  if (p->c_scatter) {
    TGCN_CHECK_ARG(a->C && a->c_dtype == TGCN_F32 && a->c_scatter_rows > 0 && a->c_scatter_rows < INT32_MAX && a->c_scatter_row0 >= 0,
                   "spmm: c_scatter_bases needs an fp32 C (for ldc), c_scatter_rows > 0 and c_scatter_row0 >= 0");
  }
  p->raw_in = a->n_raw > 0 ? a->raw_in : nullptr; p->raw_ld = a->raw_ld; p->raw_stride = a->raw_stride; p->raw_rows = a->raw_rows; p->n_raw = a->n_raw;
  if (p->raw_in) {
    TGCN_CHECK_ARG(p->raw_ld % 4 == 0 && p->raw_ld >= a->F && p->raw_stride % 4 == 0 && ((uintptr_t)p->raw_in & 15) == 0 && a->b_dtype == TGCN_F32,
                   "spmm: raw_in needs fp32 operands and 16-byte aligned rows (raw_ld, raw_stride multiples of 4)");
  }
  if (p->tc_part) {
    TGCN_CHECK_ARG(p->tc_rank && p->tc_slot_ptr && p->tc_ld % 4 == 0 && p->tc_ld >= a->F && ((uintptr_t)p->tc_part & 15) == 0,
                   "spmm: dense-tile partials need tc_rank, tc_slot_ptr and a 16-byte aligned buffer with tc_ld %% 4 == 0");
    TGCN_CHECK_ARG(a->b_dtype == TGCN_F32, "spmm: dense-tile partials need an fp32 operand");
  }
  if (p->ad_p) {
    TGCN_CHECK_ARG(p->ad_m && p->ad_v && p->ad_hyp, "spmm: fused Adam needs exp_avg, exp_avg_sq and the hyper buffer");
    TGCN_CHECK_ARG(a->P == nullptr && a->b_dtype == TGCN_F32 && p->ad_ld % 4 == 0 && p->ad_ld >= a->F,
                   "spmm: fused Adam needs fp32 operands, no projection and adam_ld %% 4 == 0");
    TGCN_CHECK_ARG((((uintptr_t)p->ad_p | (uintptr_t)p->ad_m | (uintptr_t)p->ad_v | (uintptr_t)(p->ad_x ? p->ad_x : p->ad_p)) & 15) == 0,
                   "spmm: fused Adam buffers must be 16-byte aligned");
  }
  p->wproj_in_smem = (p->P && (size_t)p->F * ((p->n_proj + 3) & ~3) * sizeof(float) <= 64 * 1024) ? 1 : 0;
  return TGCN_OK;
}

}  // namespace tgcn
