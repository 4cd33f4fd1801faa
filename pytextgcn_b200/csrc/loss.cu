// Masked log-softmax / NLL (+ gradient, argmax, #correct) in one pass over the logits.
// Replaces `criterion(gcn(g)[g.train_mask], g.y[g.train_mask])` with
// CrossEntropyLoss(reduction='mean') and its backward (flat_amazon.py:82,101-102,105), and the
// host-side argmax/accuracy of the eval half of the epoch (flat_amazon.py:111-114).
// y is never read where mask == 0 (labels may be -1 there, perlabel_amazon.py:108-109).
#include "common.cuh"

namespace tgcn {

// one warp per row
// row_hit[row]: bit 0 = row in `mask` and argmax == y, bit 1 = row in `mask2` and argmax == y (second accuracy count of
// the eval pass: flat_amazon.py:113-114 scores the validation AND the training rows from the same logits)
__global__ void __launch_bounds__(256) k_masked_nll(const float* __restrict__ Z, int64_t ldz, int64_t n_rows, int C,
                                                    const int64_t* __restrict__ y, const uint8_t* __restrict__ mask,
                                                    const uint8_t* __restrict__ mask2,
                                                    float inv_n, float* __restrict__ dZ, int64_t lddz,
                                                    int32_t* __restrict__ pred, float* __restrict__ row_nll,
                                                    int32_t* __restrict__ row_hit, float* __restrict__ dZ_mirror,
                                                    uint32_t* __restrict__ done_counter) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *done_counter = 0;   // arrival counter of the reduction kernel that follows
  if (row >= n_rows) return;
  const bool m = mask ? (mask[row] != 0) : true;
  const bool m2 = (mask2 != nullptr && row_hit != nullptr) ? (mask2[row] != 0) : false;
  const bool need_fwd = m || m2 || pred;
  const float* z = Z + row * ldz;
  if (!need_fwd) {
    if (dZ) for (int c = lane; c < C; c += 32) { dZ[row * lddz + c] = 0.0f; if (dZ_mirror) multimem_st_f32(dZ_mirror + row * lddz + c, 0.0f); }
    if (lane == 0) { row_nll[row] = 0.0f; if (row_hit) row_hit[row] = 0; }
    return;
  }
  float mx = -INFINITY; int arg = 0;
  for (int c = lane; c < C; c += 32) {
    float v = z[c];
    if (v > mx) { mx = v; arg = c; }
  }
  // warp argmax, lowest index wins ties (numpy argmax, flat_amazon.py:111)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
    if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
  }
  if (pred && lane == 0) pred[row] = arg;
  if (!m) {
    if (dZ) for (int c = lane; c < C; c += 32) { dZ[row * lddz + c] = 0.0f; if (dZ_mirror) multimem_st_f32(dZ_mirror + row * lddz + c, 0.0f); }
    if (lane == 0) { row_nll[row] = 0.0f; if (row_hit) row_hit[row] = (m2 && arg == (int)y[row]) ? 2 : 0; }
    return;
  }
  float se = 0.0f;
  for (int c = lane; c < C; c += 32) se += expf(z[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const int64_t yi = y[row];
  if (lane == 0) {
    row_nll[row] = (yi >= 0 && yi < C) ? (lse - z[yi]) : 0.0f;
    if (row_hit) row_hit[row] = (arg == (int)yi) ? (m2 ? 3 : 1) : 0;
  }
  if (dZ) {
    for (int c = lane; c < C; c += 32) {
      float pr = expf(z[c] - lse);
      const float gz = (pr - (c == (int)yi ? 1.0f : 0.0f)) * inv_n;
      dZ[row * lddz + c] = gz;
      if (dZ_mirror) multimem_st_f32(dZ_mirror + row * lddz + c, gz);
    }
    if (dZ_mirror) __threadfence_system();
  }
}

// fixed-order reduction of the per-row terms (deterministic, fp64): stage 1 = NLL_PARTS CTAs over contiguous
// slices, stage 2 = one CTA adding the slice partials in slice order.
constexpr int NLL_PARTS = 64;
__device__ __forceinline__ void block_reduce3(double& s, int& cnt, int& hit, double* s_sum, int* s_cnt, int* s_hit) {
  s_sum[threadIdx.x] = s; s_cnt[threadIdx.x] = cnt; s_hit[threadIdx.x] = hit;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
      s_hit[threadIdx.x] += s_hit[threadIdx.x + o];
    }
    __syncthreads();
  }
  s = s_sum[0]; cnt = s_cnt[0]; hit = s_hit[0];
}

// stage 1: every CTA reduces a contiguous slice; stage 2: the CTA that arrives LAST (device counter) adds the slice
// partials in slice order -- same sum whichever CTA that is -- and writes the results.  One launch.
__global__ void __launch_bounds__(256) k_nll_reduce(const float* __restrict__ row_nll, const int32_t* __restrict__ row_hit,
                                                    const uint8_t* __restrict__ mask, int64_t n_rows,
                                                    double* __restrict__ part_sum, int32_t* __restrict__ part_cnt,
                                                    uint32_t* __restrict__ done_counter, int64_t n_mask_total,
                                                    float* __restrict__ loss_out, double* __restrict__ partial_out,
                                                    int32_t* __restrict__ correct_out, int32_t* __restrict__ correct2_out) {
  __shared__ double s_sum[256];
  __shared__ int s_cnt[256];
  __shared__ int s_hit[256];
  __shared__ int s_hit2[256];
  __shared__ bool s_last;
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(n_rows, r0 + per);
  double s = 0.0; int cnt = 0, hit = 0, hit2 = 0;
  for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    s += (double)row_nll[r];
    cnt += mask ? (mask[r] != 0) : 1;
    if (row_hit) { const int h = row_hit[r]; hit += h & 1; hit2 += (h >> 1) & 1; }
  }
  s_hit2[threadIdx.x] = hit2;
  block_reduce3(s, cnt, hit, s_sum, s_cnt, s_hit);
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_hit2[threadIdx.x] += s_hit2[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part_sum[blockIdx.x] = s; part_cnt[3 * blockIdx.x] = cnt; part_cnt[3 * blockIdx.x + 1] = hit; part_cnt[3 * blockIdx.x + 2] = s_hit2[0];
    __threadfence();
    s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  double ts = 0.0; int tc = 0, th = 0, th2 = 0;
  for (unsigned i = 0; i < gridDim.x; ++i) {
    ts += __ldcg(part_sum + i); tc += __ldcg(part_cnt + 3 * i); th += __ldcg(part_cnt + 3 * i + 1); th2 += __ldcg(part_cnt + 3 * i + 2);
  }
  const double n = n_mask_total > 0 ? (double)n_mask_total : (double)tc;
  if (loss_out) { loss_out[0] = (float)(ts / n); loss_out[1] = (float)tc; }
  if (partial_out) { partial_out[0] = ts; partial_out[1] = (double)tc; }
  if (correct_out) correct_out[0] = th;
  if (correct2_out) correct2_out[0] = th2;
}

__global__ void k_count_mask(const uint8_t* __restrict__ mask, int64_t n, int32_t* __restrict__ out) {
  __shared__ int s[1024];
  int c = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) c += mask[i] != 0;
  s[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_masked_nll_workspace_bytes(int64_t n_rows, size_t* bytes_out) {
  TGCN_CHECK_ARG(bytes_out != nullptr && n_rows >= 0, "masked_nll_workspace_bytes: bad arguments");
  *bytes_out = align_up((size_t)n_rows * 4, 256) * 2 + 4096;   // row terms + row hits + slice partials
  return TGCN_OK;
}

extern "C" int tgcn_masked_nll(const float* Z, int64_t ldz, int64_t n_rows, int32_t C,
                               const int64_t* y, const uint8_t* mask, int64_t n_mask_total,
                               float* loss_out, double* partial_out, float* dZ, int64_t lddz,
                               int32_t* pred_out, int32_t* correct_out, void* dZ_mirror_mc,
                               const uint8_t* mask2, int32_t* correct2_out,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(Z && y, "masked_nll: Z / y null");
  TGCN_CHECK_ARG(n_rows > 0 && C > 0 && ldz >= C, "masked_nll: bad shape");
  TGCN_CHECK_ARG(dZ == nullptr || lddz >= C, "masked_nll: lddz < C");
  TGCN_CHECK_ARG(dZ == nullptr || n_mask_total > 0, "masked_nll: the gradient needs the global mask count (n_mask_total > 0)");
  TGCN_CHECK_ARG(correct2_out == nullptr || (mask2 != nullptr && correct_out != nullptr), "masked_nll: correct2_out needs mask2 and correct_out");
  size_t need = align_up((size_t)n_rows * 4, 256) * 2 + 4096;
  if (!workspace || workspace_bytes < need) {
    set_error("masked_nll workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return TGCN_EWORKSPACE;
  }
  float* row_nll = (float*)workspace;
  int32_t* row_hit = (int32_t*)((char*)workspace + align_up((size_t)n_rows * 4, 256));
  const float inv_n = n_mask_total > 0 ? 1.0f / (float)n_mask_total : 0.0f;
  const int T = 256;
  // slice partials + the arrival counter live in the last 4 KB of the caller's workspace (nothing is allocated here)
  static_assert(NLL_PARTS * (sizeof(double) + 3 * sizeof(int32_t)) + 16 <= 4096, "partials must fit the tail pad");
  double* part_sum = (double*)((char*)workspace + 2 * align_up((size_t)n_rows * 4, 256));
  int32_t* part_cnt = (int32_t*)(part_sum + NLL_PARTS);
  uint32_t* done = (uint32_t*)(part_cnt + 3 * NLL_PARTS);
  k_masked_nll<<<(unsigned)cdiv(n_rows * 32, T), T, 0, stream>>>(Z, ldz, n_rows, C, y, mask, correct2_out ? mask2 : nullptr, inv_n, dZ, lddz,
                                                                  pred_out, row_nll, correct_out ? row_hit : nullptr,
                                                                  dZ ? (float*)dZ_mirror_mc : nullptr, done);
  TGCN_LAUNCH_CHECK();
  const int parts = (int)std::min<int64_t>(NLL_PARTS, std::max<int64_t>(1, n_rows / 1024));
  k_nll_reduce<<<parts, 256, 0, stream>>>(row_nll, correct_out ? row_hit : nullptr, mask, n_rows, part_sum, part_cnt, done, n_mask_total,
                                          loss_out, partial_out, correct_out, correct2_out);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_count_mask(const uint8_t* mask, int64_t n, int32_t* count_out, void* stream_) {
  TGCN_CHECK_ARG(mask && count_out && n >= 0, "count_mask: bad arguments");
  k_count_mask<<<1, 1024, 0, (cudaStream_t)stream_>>>(mask, n, count_out);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}
