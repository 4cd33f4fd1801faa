// Dense part of the backward pass between the two SpMMs, one pass over the node rows:
//   dW2       = H1d^T G2                 (H x C)   -- weight gradient of layer 2
//   db_out    = colsum(dZ2)              (C)       -- bias gradient of layer 2
//   dZ1       = (G2 W2^T) .* keep/(1-p) .* act'    -- gradient entering layer 1's propagation
//   db_hidden = colsum(dZ1)              (H)       -- bias gradient of layer 1
// Replaces the autograd of torch.matmul(x, weight), `out += bias` and F.dropout
// (flat_amazon.py:105 through [PyG-1.6.3] GCNConv.forward and models.py:23).
// Deterministic: per-CTA partial sums in a workspace, reduced in CTA order by a second kernel.
//
// The contraction is H x C <= 256 x 219 per row: far too thin for tensor cores, so it is
// register-tiled FMA over shared-memory tiles of 32 rows.
#include "common.cuh"

namespace tgcn {

constexpr int DB_ROWS = 32;       // rows per tile
constexpr int DB_THREADS = 256;
constexpr int DB_MAXNB = 4;       // 4x4 dW2 blocks per thread

struct DenseBwdParams {
  const float* __restrict__ G2; int64_t ldg2;
  const void* __restrict__ H1d; int64_t ldh; int32_t h_dtype;
  const float* __restrict__ W2;
  const float* __restrict__ dZ2; int64_t lddz2;
  int64_t n_rows; int64_t row_offset;
  int32_t H, C;
  int32_t act, drop_mode; float drop_p, drop_scale; const uint8_t* __restrict__ keep_mask; int64_t ldmask;
  uint64_t philox_seed, philox_offset; const int64_t* __restrict__ philox_offset_dev;
  void* dZ1; int64_t lddz1; int32_t dz1_dtype;
  float* dZ1_mirror;   // multicast mapping of dZ1 (fp32 only) or NULL
  int64_t dZ1_mirror_rows;   // rows the mirror covers
  // workspace partials
  float* part_dW2;   // [n_cta_x][H*C]
  float* part_dbh;   // [n_cta_x][H]
  float* part_dbo;   // [n_cta_x][C]
  int32_t cb_per_y;  // c-blocks (of 4 classes) handled per blockIdx.y
  int32_t n_tiles;
  int32_t w2_in_smem;  // 0: W2 too large for shared memory, read it through L1/L2 instead
};

__device__ __forceinline__ float load_h(const void* H1d, int h_dtype, int64_t idx) {
  return h_dtype == TGCN_F32 ? reinterpret_cast<const float*>(H1d)[idx]
                             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(H1d)[idx]);
}

// Lane l of a warp owns the hidden quads q = l + 32*k (h = 4q .. 4q+3), k < KQ.
template <int KQ>   // KQ = ceil(H/128) upper bound
__global__ void __launch_bounds__(DB_THREADS, 2) k_dense_bwd(const DenseBwdParams p) {
  extern __shared__ __align__(16) float smem[];
  bool mirrored = false;                          // this thread stored to the multicast mapping (fence at the end)
  const int H = p.H, C = p.C;                    // H % 4 == 0 (host pads)
  const int Cp = (C + 3) & ~3;
  float* W2t = smem;                             // [C][H]   transposed weights (h contiguous) when staged
  float* Hs = W2t + (p.w2_in_smem ? C * H : 0);  // [DB_ROWS][H]
  float* G2s = Hs + DB_ROWS * H;                 // [DB_ROWS][Cp]   row-major, for the dW2 outer products
  float* G2t = G2s + DB_ROWS * Cp;               // [Cp][DB_ROWS]   transposed, for the dZ1 contraction
  float* dZs = G2t + Cp * DB_ROWS;               // [DB_ROWS][Cp]   dZ2 tile for db_out
  float* red = dZs + DB_ROWS * Cp;               // [8][H] cross-warp reduction scratch
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool do_rows = (blockIdx.y == 0);        // dZ1 / db are produced once, by the y == 0 slice
  const int HQ = H >> 2;

  if (p.w2_in_smem)
    for (int i = tid; i < H * C; i += DB_THREADS) { const int h = i / C, c = i - h * C; W2t[c * H + h] = p.W2[i]; }

  // dW2 register tile: block id -> (hb, cb) with cb restricted to this y-slice
  const int cb0 = blockIdx.y * p.cb_per_y;
  const int CBl = min(p.cb_per_y, (Cp >> 2) - cb0);
  const int n_blocks = HQ * CBl;
  float dw[DB_MAXNB][16];
#pragma unroll
  for (int b = 0; b < DB_MAXNB; ++b)
#pragma unroll
    for (int i = 0; i < 16; ++i) dw[b][i] = 0.0f;
  float dbh[KQ][4];
#pragma unroll
  for (int k = 0; k < KQ; ++k) dbh[k][0] = dbh[k][1] = dbh[k][2] = dbh[k][3] = 0.0f;
  constexpr int DBO_MAX = 2;   // thread t accumulates db_out[t + k*DB_THREADS], k < DBO_MAX (host checks C <= 512)
  float dbo[DBO_MAX] = {0.0f, 0.0f};
  const uint64_t ph_off = p.philox_offset + (p.philox_offset_dev ? (uint64_t)__ldg(p.philox_offset_dev) : 0ull);
  __syncthreads();

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int64_t r0 = (int64_t)tile * DB_ROWS;
    // ---- stage the tile: warp w loads rows 4w .. 4w+3 with 16-byte loads ----
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = wid * 4 + q;
      const int64_t row = r0 + r;
      for (int hq = lane; hq < HQ; hq += 32) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.n_rows) {
          if (p.h_dtype == TGCN_F32) {
            v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.H1d) + row * p.ldh) + hq);
          } else {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.H1d) + row * p.ldh) + hq);
            v = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                            __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
          }
        }
        *reinterpret_cast<float4*>(Hs + r * H + 4 * hq) = v;
      }
    }
    for (int i = tid; i < DB_ROWS * Cp; i += DB_THREADS) {
      const int r = i / Cp, c = i - r * Cp;
      const int64_t row = r0 + r;
      const bool ok = (row < p.n_rows && c < C);
      const float g = ok ? p.G2[row * p.ldg2 + c] : 0.0f;
      G2s[i] = g;
      G2t[c * DB_ROWS + r] = g;
      if (p.dZ2) dZs[i] = ok ? p.dZ2[row * p.lddz2 + c] : 0.0f;
    }
    __syncthreads();

    if (do_rows && p.dZ2) {
#pragma unroll
      for (int k = 0; k < DBO_MAX; ++k) {
        const int c = tid + k * DB_THREADS;
        if (c < C) {
#pragma unroll 8
          for (int r = 0; r < DB_ROWS; ++r) dbo[k] += dZs[r * Cp + c];
        }
      }
    }

    // ---- dW2 += Hs^T G2s  (4x4 register blocks) ----
#pragma unroll
    for (int b = 0; b < DB_MAXNB; ++b) {
      const int blk = tid + b * DB_THREADS;
      if (blk < n_blocks) {
        const int hb = blk / CBl, cb = blk - hb * CBl + cb0;
#pragma unroll 8
        for (int r = 0; r < DB_ROWS; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(Hs + r * H + hb * 4);
          const float4 g = *reinterpret_cast<const float4*>(G2s + r * Cp + cb * 4);
          dw[b][0] = fmaf(a.x, g.x, dw[b][0]);  dw[b][1] = fmaf(a.x, g.y, dw[b][1]);
          dw[b][2] = fmaf(a.x, g.z, dw[b][2]);  dw[b][3] = fmaf(a.x, g.w, dw[b][3]);
          dw[b][4] = fmaf(a.y, g.x, dw[b][4]);  dw[b][5] = fmaf(a.y, g.y, dw[b][5]);
          dw[b][6] = fmaf(a.y, g.z, dw[b][6]);  dw[b][7] = fmaf(a.y, g.w, dw[b][7]);
          dw[b][8] = fmaf(a.z, g.x, dw[b][8]);  dw[b][9] = fmaf(a.z, g.y, dw[b][9]);
          dw[b][10] = fmaf(a.z, g.z, dw[b][10]); dw[b][11] = fmaf(a.z, g.w, dw[b][11]);
          dw[b][12] = fmaf(a.w, g.x, dw[b][12]); dw[b][13] = fmaf(a.w, g.y, dw[b][13]);
          dw[b][14] = fmaf(a.w, g.z, dw[b][14]); dw[b][15] = fmaf(a.w, g.w, dw[b][15]);
        }
      }
    }

    // ---- dZ1 rows, narrow hidden layer (H <= 32): lane = (row q = lane/8, hidden quad hq = lane%8), so
    //      all 32 lanes work on the warp's 4 rows (the general mapping below would leave 24 idle) ----
    if (do_rows && p.dZ1 && HQ <= 8) {
      const int rb = wid * 4, q = lane >> 3, hq = lane & 7;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (hq < HQ) {
        for (int c = 0; c < C; ++c) {
          const float g = G2t[c * DB_ROWS + rb + q];
          float4 w;
          if (p.w2_in_smem) {
            w = *reinterpret_cast<const float4*>(W2t + c * H + 4 * hq);
          } else {
            const float* wp = p.W2 + (int64_t)(4 * hq) * C + c;
            w = make_float4(__ldg(wp), __ldg(wp + C), __ldg(wp + 2 * C), __ldg(wp + 3 * C));
          }
          v[0] = fmaf(g, w.x, v[0]); v[1] = fmaf(g, w.y, v[1]); v[2] = fmaf(g, w.z, v[2]); v[3] = fmaf(g, w.w, v[3]);
        }
      }
      const int r = rb + q;
      const int64_t row = r0 + r;
      if (row < p.n_rows && hq < HQ) {
        if (p.act == TGCN_ACT_RELU) {
          const float4 hv = *reinterpret_cast<const float4*>(Hs + r * H + 4 * hq);
          v[0] = hv.x > 0.0f ? v[0] * p.drop_scale : 0.0f; v[1] = hv.y > 0.0f ? v[1] * p.drop_scale : 0.0f;
          v[2] = hv.z > 0.0f ? v[2] * p.drop_scale : 0.0f; v[3] = hv.w > 0.0f ? v[3] * p.drop_scale : 0.0f;
        } else if (p.drop_mode == TGCN_DROP_MASK) {
          const uint8_t* m = p.keep_mask + row * p.ldmask + 4 * hq;
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = m[i] ? v[i] * p.drop_scale : 0.0f;
        } else if (p.drop_mode == TGCN_DROP_PHILOX) {
          const uint64_t e4 = ((uint64_t)(row + p.row_offset) * (uint64_t)H + (uint64_t)(4 * hq)) >> 2;
          const uint4 rr = philox_quad(e4, p.philox_seed, ph_off);
          v[0] = (u01(rr.x) >= p.drop_p) ? v[0] * p.drop_scale : 0.0f;
          v[1] = (u01(rr.y) >= p.drop_p) ? v[1] * p.drop_scale : 0.0f;
          v[2] = (u01(rr.z) >= p.drop_p) ? v[2] * p.drop_scale : 0.0f;
          v[3] = (u01(rr.w) >= p.drop_p) ? v[3] * p.drop_scale : 0.0f;
        }
        if (p.dz1_dtype == TGCN_F32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dZ1) + row * p.lddz1 + 4 * hq) = make_float4(v[0], v[1], v[2], v[3]);
          if (p.dZ1_mirror && row < p.dZ1_mirror_rows) { multimem_st_v4(p.dZ1_mirror + row * p.lddz1 + 4 * hq, v[0], v[1], v[2], v[3]); mirrored = true; }
        } else {
          __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.dZ1) + row * p.lddz1 + 4 * hq);
          o[0] = __floats2bfloat162_rn(v[0], v[1]);
          o[1] = __floats2bfloat162_rn(v[2], v[3]);
        }
      } else {
        v[0] = v[1] = v[2] = v[3] = 0.0f;
      }
      // db_hidden: add the 4 rows of this warp (lanes hq, hq+8, hq+16, hq+24), fixed order
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float t = v[i];
        t += __shfl_down_sync(0xffffffffu, t, 16);
        t += __shfl_down_sync(0xffffffffu, t, 8);
        if (lane < 8) dbh[0][i] += t;
      }
    } else
    // ---- dZ1 rows: warp `wid` owns rows 4*wid .. 4*wid+3; lane owns hidden quads lane + 32k ----
    if (do_rows && p.dZ1) {
      float dz[4][KQ][4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < KQ; ++k) dz[q][k][0] = dz[q][k][1] = dz[q][k][2] = dz[q][k][3] = 0.0f;
      const int rb = wid * 4;
      for (int c = 0; c < C; ++c) {
        const float4 g = *reinterpret_cast<const float4*>(G2t + c * DB_ROWS + rb);   // 4 rows, broadcast
        const float gq[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int hq = lane + 32 * k;
          if (hq < HQ) {
            float4 w;
            if (p.w2_in_smem) {
              w = *reinterpret_cast<const float4*>(W2t + c * H + 4 * hq);
            } else {
              const float* wp = p.W2 + (int64_t)(4 * hq) * C + c;
              w = make_float4(__ldg(wp), __ldg(wp + C), __ldg(wp + 2 * C), __ldg(wp + 3 * C));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dz[q][k][0] = fmaf(gq[q], w.x, dz[q][k][0]); dz[q][k][1] = fmaf(gq[q], w.y, dz[q][k][1]);
              dz[q][k][2] = fmaf(gq[q], w.z, dz[q][k][2]); dz[q][k][3] = fmaf(gq[q], w.w, dz[q][k][3]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = rb + q;
        const int64_t row = r0 + r;
        if (row >= p.n_rows) continue;
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          const int hq = lane + 32 * k;
          if (hq >= HQ) continue;
          float v[4] = {dz[q][k][0], dz[q][k][1], dz[q][k][2], dz[q][k][3]};
          if (p.act == TGCN_ACT_RELU) {
            // forward was dropout(relu(z)): the output is > 0 iff kept and z > 0
            const float4 hv = *reinterpret_cast<const float4*>(Hs + r * H + 4 * hq);
            v[0] = hv.x > 0.0f ? v[0] * p.drop_scale : 0.0f; v[1] = hv.y > 0.0f ? v[1] * p.drop_scale : 0.0f;
            v[2] = hv.z > 0.0f ? v[2] * p.drop_scale : 0.0f; v[3] = hv.w > 0.0f ? v[3] * p.drop_scale : 0.0f;
          } else if (p.drop_mode == TGCN_DROP_MASK) {
            const uint8_t* m = p.keep_mask + row * p.ldmask + 4 * hq;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = m[i] ? v[i] * p.drop_scale : 0.0f;
          } else if (p.drop_mode == TGCN_DROP_PHILOX) {
            const uint64_t e4 = ((uint64_t)(row + p.row_offset) * (uint64_t)H + (uint64_t)(4 * hq)) >> 2;
            const uint4 rr = philox_quad(e4, p.philox_seed, ph_off);
            v[0] = (u01(rr.x) >= p.drop_p) ? v[0] * p.drop_scale : 0.0f;
            v[1] = (u01(rr.y) >= p.drop_p) ? v[1] * p.drop_scale : 0.0f;
            v[2] = (u01(rr.z) >= p.drop_p) ? v[2] * p.drop_scale : 0.0f;
            v[3] = (u01(rr.w) >= p.drop_p) ? v[3] * p.drop_scale : 0.0f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) dbh[k][i] += v[i];
          if (p.dz1_dtype == TGCN_F32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dZ1) + row * p.lddz1 + 4 * hq) = make_float4(v[0], v[1], v[2], v[3]);
            if (p.dZ1_mirror && row < p.dZ1_mirror_rows) { multimem_st_v4(p.dZ1_mirror + row * p.lddz1 + 4 * hq, v[0], v[1], v[2], v[3]); mirrored = true; }
          } else {
            __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(p.dZ1) + row * p.lddz1 + 4 * hq);
            o[0] = __floats2bfloat162_rn(v[0], v[1]);
            o[1] = __floats2bfloat162_rn(v[2], v[3]);
          }
        }
      }
    }
    __syncthreads();
  }

  if (mirrored) __threadfence_system();      // only the threads that stored to the multicast mapping
  // ---- per-CTA partials ----
  float* my_dw = p.part_dW2 + (int64_t)blockIdx.x * H * C;
#pragma unroll
  for (int b = 0; b < DB_MAXNB; ++b) {
    const int blk = tid + b * DB_THREADS;
    if (blk < n_blocks) {
      const int hb = blk / CBl, cb = blk - hb * CBl + cb0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int h = hb * 4 + i, c = cb * 4 + j;
          if (h < H && c < C) my_dw[h * C + c] = dw[b][i * 4 + j];
        }
    }
  }
  if (do_rows) {
    // db_hidden: fixed-order sum over the 8 warps
#pragma unroll
    for (int k = 0; k < KQ; ++k) {
      const int hq = lane + 32 * k;
      if (hq < HQ) *reinterpret_cast<float4*>(red + wid * H + 4 * hq) = make_float4(dbh[k][0], dbh[k][1], dbh[k][2], dbh[k][3]);
    }
    __syncthreads();
    for (int h = tid; h < H; h += DB_THREADS) {
      float s = 0.0f;
      for (int w = 0; w < DB_THREADS / 32; ++w) s += red[w * H + h];
      p.part_dbh[(int64_t)blockIdx.x * H + h] = s;
    }
#pragma unroll
    for (int k = 0; k < DBO_MAX; ++k)
      if (tid + k * DB_THREADS < C) p.part_dbo[(int64_t)blockIdx.x * C + tid + k * DB_THREADS] = dbo[k];
  }
}

// Reduces the three per-CTA partial arrays in ONE launch: a warp per output element, lanes stride over the
// CTA partials and a fixed shuffle tree adds the 32 lane sums -- same order every run (deterministic).
__global__ void __launch_bounds__(256) k_reduce_partials3(const float* __restrict__ p0, int64_t n0, float* __restrict__ o0,
                                                          const float* __restrict__ p1, int64_t n1, float* __restrict__ o1,
                                                          const float* __restrict__ p2, int64_t n2, float* __restrict__ o2,
                                                          int n_parts) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n0 + n1 + n2) return;
  const float* part; float* out; int64_t n, i;
  if (w < n0) { part = p0; out = o0; n = n0; i = w; }
  else if (w < n0 + n1) { part = p1; out = o1; n = n1; i = w - n0; }
  else { part = p2; out = o2; n = n2; i = w - n0 - n1; }
  float s = 0.0f;
  for (int c = lane; c < n_parts; c += 32) s += part[(int64_t)c * n + i];
  s = warp_sum(s);
  if (lane == 0) out[i] = s;
}

// thin projection P = X W: a warp takes 4 rows per step (row values staged transposed [k][4] in shared
// memory so one 16-byte broadcast read feeds 4 FMAs per weight read); W staged in shared memory when it fits
constexpr int PJ_ROWS = 4;
__global__ void __launch_bounds__(256) k_project(const void* __restrict__ X, int64_t ldx, int x_dtype, int64_t n_rows, int K,
                                                 const float* __restrict__ W, int M, const float* __restrict__ bias,
                                                 float* __restrict__ P, int64_t ldp, int w_in_smem, float* __restrict__ mirror) {
  extern __shared__ __align__(16) float smem[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ws = smem;
  float* rowbuf = smem + (w_in_smem ? ((K * M + 3) & ~3) : 0) + wid * (K * PJ_ROWS);   // [K][4]
  if (w_in_smem) {
    for (int i = threadIdx.x; i < K * M; i += blockDim.x) Ws[i] = W[i];
  }
  __syncthreads();
  const float* Wp = w_in_smem ? Ws : W;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t n_groups = (n_rows + PJ_ROWS - 1) / PJ_ROWS;
  for (int64_t grp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid; grp < n_groups; grp += warps) {
    const int64_t row0 = grp * PJ_ROWS;
    for (int i = lane; i < K * PJ_ROWS; i += 32) {
      const int r = i / K, k = i - r * K;                       // coalesced along k within a row
      const int64_t row = row0 + r;
      rowbuf[k * PJ_ROWS + r] = (row < n_rows) ? load_h(X, x_dtype, row * ldx + k) : 0.0f;
    }
    __syncwarp();
    for (int m = lane; m < M; m += 32) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (int k = 0; k < K; ++k) {
        const float w = Wp[(int64_t)k * M + m];
        const float4 x = *reinterpret_cast<const float4*>(rowbuf + k * PJ_ROWS);
        s0 = fmaf(x.x, w, s0); s1 = fmaf(x.y, w, s1); s2 = fmaf(x.z, w, s2); s3 = fmaf(x.w, w, s3);
      }
      if (bias) { const float b = __ldg(bias + m); s0 += b; s1 += b; s2 += b; s3 += b; }
      if (row0 + 0 < n_rows) { P[(row0 + 0) * ldp + m] = s0; if (mirror) multimem_st_f32(mirror + (row0 + 0) * ldp + m, s0); }
      if (row0 + 1 < n_rows) { P[(row0 + 1) * ldp + m] = s1; if (mirror) multimem_st_f32(mirror + (row0 + 1) * ldp + m, s1); }
      if (row0 + 2 < n_rows) { P[(row0 + 2) * ldp + m] = s2; if (mirror) multimem_st_f32(mirror + (row0 + 2) * ldp + m, s2); }
      if (row0 + 3 < n_rows) { P[(row0 + 3) * ldp + m] = s3; if (mirror) multimem_st_f32(mirror + (row0 + 3) * ldp + m, s3); }
    }
    __syncwarp();
  }
  if (mirror) __threadfence_system();
}

// column sums of X[n_rows, F]: per-CTA partial rows (a CTA walks a contiguous row range, thread = column), reduced in
// CTA order by k_reduce_partials3 -- same order every run
__global__ void __launch_bounds__(256) k_colsum_partial(const float* __restrict__ X, int64_t ldx, int64_t n_rows, int F,
                                                        float* __restrict__ part) {
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(n_rows, r0 + per);
  for (int c = threadIdx.x; c < F; c += blockDim.x) {
    float s = 0.0f;
    for (int64_t r = r0; r < r1; ++r) s += X[r * ldx + c];
    part[(int64_t)blockIdx.x * F + c] = s;
  }
}

// Thin projection, lane = row:  P[r, c0 .. c0 + 4*N4) = dropout(X[r, :K]) W[:, c0 ..] (+ bias).
// A warp owns 32 consecutive rows.  The X block of the warp goes through shared memory TRANSPOSED in chunks of 32
// columns ([k][33]: every lane then reads its own row's value without bank conflicts), W sits in shared memory and is
// read as 16-byte BROADCASTS (all lanes the same address): per k one 4-byte load + N4 broadcasts feed 4*N4 FMAs per lane
// -- ~0.8 instructions per row and k at 20 classes, against 2 shared-memory loads per 4 FMAs in the row-blocked kernel
// above (which stays for bf16 operands).  The dropout of the training forward (keep-mask bytes or Philox, the decision of
// the SpMM epilogue) is applied while the block is loaded and the dropped block is written out (Xd) for the backward pass,
// so a pre-dropout activation shared with the preceding eval forward needs no separate dropout pass.
struct ProjParams {
  const float* __restrict__ X; int64_t ldx; int64_t n_rows; int32_t K;
  const float* __restrict__ W; int32_t M; const float* __restrict__ bias;
  float* P; int64_t ldp; float* mirror; int64_t mirror_rows;
  int32_t drop_mode; float drop_p, drop_scale; const uint8_t* __restrict__ keep_mask; int64_t ldmask;
  uint64_t philox_seed, philox_offset; const int64_t* __restrict__ philox_offset_dev; int64_t philox_row_offset; int32_t philox_F;
  float* Xd; int64_t ldxd;
};
constexpr int PR_KC = 32;     // columns of X per shared-memory chunk
template <int N4>
__global__ void __launch_bounds__(256) k_project_rows(const ProjParams p) {
  extern __shared__ __align__(16) float smem[];
  bool mirrored = false;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * (4 * N4);                 // first output column of this CTA's column tile
  const int Kp = (p.K + PR_KC - 1) / PR_KC * PR_KC;
  float* Ws = smem;                                      // [Kp][4*N4], zero past K / past M
  float* Xs = smem + Kp * (4 * N4) + wid * (PR_KC * 33); // [32 k][33]
  for (int i = threadIdx.x; i < Kp * (4 * N4); i += blockDim.x) {
    const int k = i / (4 * N4), c = i - k * (4 * N4);
    Ws[i] = (k < p.K && c0 + c < p.M) ? p.W[(int64_t)k * p.M + c0 + c] : 0.0f;
  }
  __syncthreads();
  const uint64_t ph_off = p.philox_offset + (p.philox_offset_dev ? (uint64_t)__ldg(p.philox_offset_dev) : 0ull);
  const int64_t n_blocks = (p.n_rows + 31) / 32;
  for (int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + wid; blk < n_blocks; blk += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t row0 = blk * 32;
    float acc[4 * N4];
#pragma unroll
    for (int i = 0; i < 4 * N4; ++i) acc[i] = 0.0f;
    // software pipeline: the 8 float4 of the NEXT chunk are in flight while the current chunk is multiplied
    float4 nxt[8];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = i * 32 + lane, r = t >> 3, kq = t & 7;
        const int64_t row = row0 + r;
        const int k = k0 + 4 * kq;
        nxt[i] = (row < p.n_rows && k < p.K) ? __ldg(reinterpret_cast<const float4*>(p.X + row * p.ldx + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    fetch(0);
    for (int k0 = 0; k0 < p.K; k0 += PR_KC) {
      // ---- 32 rows x 32 columns (8 float4 per lane): dropout, write-out, transpose into shared memory ----
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = i * 32 + lane, r = t >> 3, kq = t & 7;
        const int64_t row = row0 + r;
        const int k = k0 + 4 * kq;
        float4 v = nxt[i];
        if (row < p.n_rows && k < p.K) {
          if (p.drop_mode == TGCN_DROP_MASK) {
            const uint8_t* m = p.keep_mask + row * p.ldmask + k;
            v.x = m[0] ? v.x * p.drop_scale : 0.0f; v.y = m[1] ? v.y * p.drop_scale : 0.0f;
            v.z = m[2] ? v.z * p.drop_scale : 0.0f; v.w = m[3] ? v.w * p.drop_scale : 0.0f;
          } else if (p.drop_mode == TGCN_DROP_PHILOX) {
            const uint64_t e4 = ((uint64_t)(row + p.philox_row_offset) * (uint64_t)p.philox_F + (uint64_t)k) >> 2;
            const uint4 rr = philox_quad(e4, p.philox_seed, ph_off);
            v.x = (u01(rr.x) >= p.drop_p) ? v.x * p.drop_scale : 0.0f; v.y = (u01(rr.y) >= p.drop_p) ? v.y * p.drop_scale : 0.0f;
            v.z = (u01(rr.z) >= p.drop_p) ? v.z * p.drop_scale : 0.0f; v.w = (u01(rr.w) >= p.drop_p) ? v.w * p.drop_scale : 0.0f;
          }
          if (p.Xd && blockIdx.y == 0) *reinterpret_cast<float4*>(p.Xd + row * p.ldxd + k) = v;
        }
        float* d = Xs + (4 * kq) * 33 + r;
        d[0] = v.x; d[33] = v.y; d[66] = v.z; d[99] = v.w;
      }
      __syncwarp();
      if (k0 + PR_KC < p.K) fetch(k0 + PR_KC);
      const float* wk = Ws + k0 * (4 * N4);
#pragma unroll 4
      for (int k = 0; k < PR_KC; ++k) {
        const float x = Xs[k * 33 + lane];
#pragma unroll
        for (int c4 = 0; c4 < N4; ++c4) {
          const float4 w = *reinterpret_cast<const float4*>(wk + k * (4 * N4) + 4 * c4);
          acc[4 * c4] = fmaf(x, w.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(x, w.y, acc[4 * c4 + 1]);
          acc[4 * c4 + 2] = fmaf(x, w.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(x, w.w, acc[4 * c4 + 3]);
        }
      }
      __syncwarp();
    }
    const int64_t row = row0 + lane;
    if (row < p.n_rows) {
#pragma unroll
      for (int c4 = 0; c4 < N4; ++c4) {
        const int c = c0 + 4 * c4;
        if (c < p.ldp) {      // ldp is a multiple of 4 >= M: padding columns receive the zeros of the padded W
          float4 o = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
          if (p.bias) {
            if (c < p.M) o.x += __ldg(p.bias + c);
            if (c + 1 < p.M) o.y += __ldg(p.bias + c + 1);
            if (c + 2 < p.M) o.z += __ldg(p.bias + c + 2);
            if (c + 3 < p.M) o.w += __ldg(p.bias + c + 3);
          }
          *reinterpret_cast<float4*>(p.P + row * p.ldp + c) = o;
          if (p.mirror && row < p.mirror_rows) { multimem_st_v4(p.mirror + row * p.ldp + c, o.x, o.y, o.z, o.w); mirrored = true; }
        }
      }
    }
  }
  if (mirrored) __threadfence_system();
}

template <int N4>
static int launch_project_rows(const ProjParams& p, int col_tiles, cudaStream_t stream) {
  const int Kp = (p.K + PR_KC - 1) / PR_KC * PR_KC;
  const size_t smem = ((size_t)Kp * 4 * N4 + 8 * (size_t)PR_KC * 33) * sizeof(float);
  if (smem > 48 * 1024) TGCN_CUDA(cudaFuncSetAttribute(k_project_rows<N4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_blocks = cdiv(p.n_rows, 32);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n_blocks, 8), (int64_t)sm_count() * 4));
  k_project_rows<N4><<<dim3(gx, col_tiles), 256, smem, stream>>>(p);
  return TGCN_OK;
}

struct DbLayout { size_t off_dw, off_dbh, off_dbo, total; int n_cta; };
static DbLayout db_layout(int H, int C, int n_cta) {
  DbLayout L; size_t off = 0;
  L.n_cta = n_cta;
  L.off_dw = off; off += align_up((size_t)n_cta * H * C * 4, 256);
  L.off_dbh = off; off += align_up((size_t)n_cta * H * 4, 256);
  L.off_dbo = off; off += align_up((size_t)n_cta * C * 4, 256);
  L.total = off;
  return L;
}
static int db_grid_x() { return 2 * sm_count(); }

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_dense_bwd_workspace_bytes(int32_t H, int32_t C, size_t* bytes_out) {
  TGCN_CHECK_ARG(bytes_out && H > 0 && C > 0, "dense_bwd_workspace_bytes: bad arguments");
  *bytes_out = db_layout(H, C, db_grid_x()).total;
  return TGCN_OK;
}

extern "C" int tgcn_dense_bwd(const tgcn_dense_bwd_args* a, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(a != nullptr, "dense_bwd: args null");
  TGCN_CHECK_ARG(a->G2 && a->H1d && a->W2, "dense_bwd: input pointer null");
  TGCN_CHECK_ARG(a->dW2 != nullptr, "dense_bwd: dW2 null");
  TGCN_CHECK_ARG(a->n_rows > 0 && a->H > 0 && a->C > 0, "dense_bwd: bad shape");
  TGCN_CHECK_ARG(a->H <= 512, "dense_bwd: hidden width %d > 512 not supported", a->H);
  TGCN_CHECK_ARG(a->C <= 512, "dense_bwd: %d classes > 512 not supported", a->C);
  TGCN_CHECK_ARG(a->ldg2 >= a->C && a->ldh >= a->H, "dense_bwd: leading dimension too small");
  TGCN_CHECK_ARG(a->dZ1 == nullptr || a->lddz1 >= a->H, "dense_bwd: lddz1 < H");
  TGCN_CHECK_ARG(a->drop_mode != TGCN_DROP_MASK || a->act == TGCN_ACT_RELU || a->keep_mask || a->drop_p == 0.0f,
                 "dense_bwd: TGCN_DROP_MASK needs keep_mask");
  const int H = a->H, C = a->C;
  const int n_tiles = (int)cdiv(a->n_rows, DB_ROWS);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(db_grid_x(), cdiv(n_tiles, 3)));   // >= 3 tiles per CTA
  DbLayout L = db_layout(H, C, db_grid_x());
  if (!workspace || workspace_bytes < L.total) {
    set_error("dense_bwd workspace too small: need %zu bytes, got %zu", L.total, workspace_bytes);
    return TGCN_EWORKSPACE;
  }
  DenseBwdParams p;
  p.G2 = a->G2; p.ldg2 = a->ldg2; p.H1d = a->H1d; p.ldh = a->ldh; p.h_dtype = a->h_dtype;
  p.W2 = a->W2; p.dZ2 = a->dZ2; p.lddz2 = a->lddz2;
  p.n_rows = a->n_rows; p.row_offset = a->row_offset; p.H = H; p.C = C;
  p.act = a->act;
  p.drop_mode = (a->drop_mode != TGCN_DROP_NONE && a->drop_p > 0.0f) ? a->drop_mode : TGCN_DROP_NONE;
  p.drop_p = a->drop_p; p.drop_scale = p.drop_mode == TGCN_DROP_NONE ? 1.0f : 1.0f / (1.0f - a->drop_p);
  p.keep_mask = a->keep_mask; p.ldmask = a->ldmask;
  p.philox_seed = a->philox_seed; p.philox_offset = a->philox_offset; p.philox_offset_dev = a->philox_offset_dev;
  p.dZ1 = a->dZ1; p.lddz1 = a->lddz1; p.dz1_dtype = a->dz1_dtype;
  p.dZ1_mirror = (a->dz1_dtype == TGCN_F32) ? (float*)a->dZ1_mirror_mc : nullptr;
  p.dZ1_mirror_rows = a->dZ1_mirror_rows > 0 ? a->dZ1_mirror_rows : INT64_MAX;
  p.part_dW2 = (float*)((char*)workspace + L.off_dw);
  p.part_dbh = (float*)((char*)workspace + L.off_dbh);
  p.part_dbo = (float*)((char*)workspace + L.off_dbo);
  p.n_tiles = n_tiles;
  TGCN_CHECK_ARG(H % 4 == 0, "dense_bwd: H (%d) must be a multiple of 4 (pad the hidden width)", H);
  TGCN_CHECK_ARG(a->ldh % 4 == 0 && ((uintptr_t)a->H1d % 16) == 0, "dense_bwd: H1d must be 16-byte aligned with ldh %% 4 == 0");
  TGCN_CHECK_ARG(a->dZ1 == nullptr || (a->lddz1 % 4 == 0 && ((uintptr_t)a->dZ1 % 16) == 0),
                 "dense_bwd: dZ1 must be 16-byte aligned with lddz1 %% 4 == 0");
  const int Cp = (C + 3) & ~3;
  const int HB = H / 4, CB = Cp / 4;
  int cb_per_y = std::max(1, (DB_THREADS * DB_MAXNB) / HB);
  cb_per_y = std::min(cb_per_y, CB);
  p.cb_per_y = cb_per_y;
  const int gy = (CB + cb_per_y - 1) / cb_per_y;
  const size_t smem_rest = ((size_t)DB_ROWS * H + 3 * (size_t)DB_ROWS * Cp + (size_t)(DB_THREADS / 32) * H) * sizeof(float);
  const size_t smem_w2 = (size_t)H * C * sizeof(float);
  p.w2_in_smem = (smem_rest + smem_w2 <= 220 * 1024) ? 1 : 0;     // strided global reads of W2 are far slower than losing the 2nd CTA/SM
  size_t smem = smem_rest + (p.w2_in_smem ? smem_w2 : 0);
  TGCN_CHECK_ARG(smem <= 227 * 1024, "dense_bwd: H=%d C=%d needs %zu bytes of shared memory (> 227 KB)", H, C, smem);
  const int kq = (H / 4 + 31) / 32;
  dim3 grid(gx, gy);
#define TGCN_DB_LAUNCH(K)                                                                                       \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      TGCN_CUDA(cudaFuncSetAttribute(k_dense_bwd<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    k_dense_bwd<K><<<grid, DB_THREADS, smem, stream>>>(p);                                                      \
  } while (0)
  if (kq <= 1) TGCN_DB_LAUNCH(1);
  else if (kq <= 2) TGCN_DB_LAUNCH(2);
  else TGCN_DB_LAUNCH(4);
#undef TGCN_DB_LAUNCH
  TGCN_LAUNCH_CHECK();
  TGCN_CHECK_ARG(a->db_out == nullptr || a->dZ2 != nullptr, "dense_bwd: db_out needs dZ2");
  const int64_t n0 = (int64_t)H * C, n1 = a->db_hidden ? H : 0, n2 = a->db_out ? C : 0;
  const int T = 256;
  k_reduce_partials3<<<(unsigned)cdiv((n0 + n1 + n2) * 32, T), T, 0, stream>>>(p.part_dW2, n0, a->dW2, p.part_dbh, n1, a->db_hidden,
                                                                              p.part_dbo, n2, a->db_out, gx);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_colsum_workspace_bytes(int32_t F, size_t* bytes_out) {
  TGCN_CHECK_ARG(bytes_out && F > 0, "colsum_workspace_bytes: bad arguments");
  *bytes_out = (size_t)sm_count() * 4 * F * sizeof(float);
  return TGCN_OK;
}

extern "C" int tgcn_colsum(const float* X, int64_t ldx, int64_t n_rows, int32_t F, float* out, void* workspace,
                           size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(X && out && n_rows > 0 && F > 0 && ldx >= F, "colsum: bad arguments");
  const int grid = (int)std::min<int64_t>((int64_t)sm_count() * 4, cdiv(n_rows, 64));
  if (!workspace || workspace_bytes < (size_t)grid * F * sizeof(float)) {
    set_error("colsum workspace too small");
    return TGCN_EWORKSPACE;
  }
  k_colsum_partial<<<grid, 256, 0, stream>>>(X, ldx, n_rows, F, (float*)workspace);
  TGCN_LAUNCH_CHECK();
  k_reduce_partials3<<<(unsigned)cdiv((int64_t)F * 32, 256), 256, 0, stream>>>((const float*)workspace, F, out, nullptr, 0, nullptr,
                                                                              nullptr, 0, nullptr, grid);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_project_ex(const tgcn_project_args* a, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(a && a->X && a->W && a->P, "project: null pointer");
  TGCN_CHECK_ARG(a->n_rows > 0 && a->K > 0 && a->M > 0 && a->ldx >= a->K && a->ldp >= a->M, "project: bad shape");
  TGCN_CHECK_ARG(a->drop_mode >= TGCN_DROP_NONE && a->drop_mode <= TGCN_DROP_PHILOX, "project: bad drop_mode");
  const bool drop = a->drop_mode != TGCN_DROP_NONE && a->drop_p > 0.0f;
  const int Kp4 = (a->K + 3) & ~3;
  const bool rows_ok = a->x_dtype == TGCN_F32 && a->ldx % 4 == 0 && a->ldx >= Kp4 && a->ldp % 4 == 0 && (((uintptr_t)a->X | (uintptr_t)a->P) & 15) == 0 &&
                       (a->P_mirror_mc == nullptr || ((uintptr_t)a->P_mirror_mc & 15) == 0);
  if (!rows_ok || (size_t)((a->K + 31) / 32 * 32) * 32 * 4 > 150 * 1024) {
    TGCN_CHECK_ARG(a->mirror_rows <= 0 || a->mirror_rows >= a->n_rows, "project: mirror_rows needs fp32 X with 16-byte aligned rows");
    TGCN_CHECK_ARG(!drop && a->Xd == nullptr, "project: the fused dropout needs fp32 X with 16-byte aligned rows (ldx %% 4 == 0, ldx >= pad4(K))");
    return tgcn_project(a->X, a->ldx, a->x_dtype, a->n_rows, a->K, a->W, a->M, a->bias, a->P, a->ldp, a->P_mirror_mc, stream_);
  }
  TGCN_CHECK_ARG(!drop || a->drop_p < 1.0f, "project: dropout p must be in [0,1)");
  TGCN_CHECK_ARG(a->drop_mode != TGCN_DROP_MASK || !drop || a->keep_mask, "project: TGCN_DROP_MASK needs keep_mask");
  TGCN_CHECK_ARG(a->Xd == nullptr || (a->ldxd % 4 == 0 && a->ldxd >= Kp4 && ((uintptr_t)a->Xd & 15) == 0), "project: bad Xd");
  ProjParams p;
  p.X = (const float*)a->X; p.ldx = a->ldx; p.n_rows = a->n_rows; p.K = a->K; p.W = a->W; p.M = a->M; p.bias = a->bias;
  p.P = a->P; p.ldp = a->ldp; p.mirror = (float*)a->P_mirror_mc; p.mirror_rows = a->mirror_rows > 0 ? a->mirror_rows : INT64_MAX;
  p.drop_mode = drop ? a->drop_mode : TGCN_DROP_NONE; p.drop_p = a->drop_p; p.drop_scale = drop ? 1.0f / (1.0f - a->drop_p) : 1.0f;
  p.keep_mask = a->keep_mask; p.ldmask = a->ldmask; p.philox_seed = a->philox_seed; p.philox_offset = a->philox_offset;
  p.philox_offset_dev = a->philox_offset_dev; p.philox_row_offset = a->philox_row_offset; p.philox_F = a->K;
  p.Xd = drop ? a->Xd : nullptr; p.ldxd = a->ldxd;
  const int Mp = (a->M + 3) & ~3;
  const int n4_all = Mp / 4;
  const int n4 = std::min(n4_all, 8);                    // <= 32 output columns per thread; more: column tiles (grid.y)
  const int col_tiles = (n4_all + n4 - 1) / n4;
  int rc;
  switch (n4) {
    case 1: rc = launch_project_rows<1>(p, col_tiles, stream); break;
    case 2: rc = launch_project_rows<2>(p, col_tiles, stream); break;
    case 3: rc = launch_project_rows<3>(p, col_tiles, stream); break;
    case 4: rc = launch_project_rows<4>(p, col_tiles, stream); break;
    case 5: rc = launch_project_rows<5>(p, col_tiles, stream); break;
    case 6: rc = launch_project_rows<6>(p, col_tiles, stream); break;
    case 7: rc = launch_project_rows<7>(p, col_tiles, stream); break;
    default: rc = launch_project_rows<8>(p, col_tiles, stream); break;
  }
  if (rc) return rc;
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_project(const void* X, int64_t ldx, int32_t x_dtype, int64_t n_rows, int32_t K,
                            const float* W, int32_t M, const float* bias, float* P, int64_t ldp, void* P_mirror_mc, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(X && W && P, "project: null pointer");
  TGCN_CHECK_ARG(n_rows > 0 && K > 0 && M > 0 && ldx >= K && ldp >= M, "project: bad shape");
  const int threads = 256, wpb = threads / 32;
  const int w_in_smem = ((size_t)K * M * 4 <= 96 * 1024) ? 1 : 0;
  size_t smem = ((w_in_smem ? ((K * M + 3) & ~3) : 0) + (size_t)wpb * K * PJ_ROWS) * sizeof(float);
  if (smem > 48 * 1024) TGCN_CUDA(cudaFuncSetAttribute(k_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(cdiv(cdiv(n_rows, PJ_ROWS), wpb), (int64_t)sm_count() * 8);
  k_project<<<grid, threads, smem, stream>>>(X, ldx, x_dtype, n_rows, K, W, M, bias, P, ldp, w_in_smem, (float*)P_mirror_mc);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}
