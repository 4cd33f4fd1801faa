// Error plumbing, device info, fused Adam/AMSGrad, the [I | F] hierarchy-feature kernels and
// small utility kernels of the textgcn_b200 C ABI.
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace tgcn {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- Adam / AMSGrad (torch.optim.Adam semantics; flat_amazon.py:89,106) ----
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, float* __restrict__ vmax, int64_t n, float lr,
                                              float b1, float b2, float eps, int amsgrad, int64_t step_host,
                                              const int64_t* __restrict__ step_dev, float* __restrict__ mirror) {
  __shared__ float s_hyp[2];
  if (threadIdx.x == 0) {
    const double t = (double)(step_dev ? *step_dev : step_host);
    const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
    s_hyp[0] = (float)((double)lr / bc1);
    s_hyp[1] = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_hyp[0], bc2s = s_hyp[1];
  const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)(vmax ? vmax : p) | (uintptr_t)(mirror ? mirror : p)) & 15) == 0);
  int64_t start_tail = 0;
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 P = reinterpret_cast<float4*>(p)[i];
      const float4 G = reinterpret_cast<const float4*>(g)[i];
      float4 M = reinterpret_cast<float4*>(m)[i];
      float4 V = reinterpret_cast<float4*>(v)[i];
      float4 X = amsgrad ? reinterpret_cast<float4*>(vmax)[i] : make_float4(0, 0, 0, 0);
      float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &V.x; float* xx = &X.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mm[k] = mm[k] * b1 + omb1 * gg[k];
        vv[k] = vv[k] * b2 + omb2 * (gg[k] * gg[k]);
        float vh = vv[k];
        if (amsgrad) { xx[k] = fmaxf(xx[k], vv[k]); vh = xx[k]; }
        const float denom = sqrtf(vh) / bc2s + eps;
        pp[k] = pp[k] - step_size * (mm[k] / denom);
      }
      reinterpret_cast<float4*>(p)[i] = P;
      if (mirror) multimem_st_v4(mirror + 4 * i, P.x, P.y, P.z, P.w);   // updated rows land in every rank's copy
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
      if (amsgrad) reinterpret_cast<float4*>(vmax)[i] = X;
    }
    start_tail = n4 << 2;
  }
  for (int64_t i = start_tail + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = m[i] * b1 + omb1 * gi;
    const float vi = v[i] * b2 + omb2 * (gi * gi);
    float vh = vi;
    if (amsgrad) { vh = fmaxf(vmax[i], vi); vmax[i] = vh; }
    m[i] = mi; v[i] = vi;
    const float pn = p[i] - step_size * (mi / (sqrtf(vh) / bc2s + eps));
    p[i] = pn;
    if (mirror) multimem_st_f32(mirror + i, pn);
  }
  if (mirror) __threadfence_system();
}

// up to 4 small tensors in one launch (b1, W2, b2): blockIdx.y selects the tensor
struct AdamSmall { float* p[4]; const float* g[4]; float* m[4]; float* v[4]; float* x[4]; int64_t n[4]; };
__global__ void __launch_bounds__(256) k_adam_small(AdamSmall t, float lr, float b1, float b2, float eps, int amsgrad,
                                                    int64_t step_host, const int64_t* __restrict__ step_dev) {
  __shared__ float s_hyp[2];
  if (threadIdx.x == 0) {
    const double st = (double)(step_dev ? *step_dev : step_host);
    s_hyp[0] = (float)((double)lr / (1.0 - pow((double)b1, st)));
    s_hyp[1] = (float)sqrt(1.0 - pow((double)b2, st));
  }
  __syncthreads();
  const int k = blockIdx.y;
  float* p = t.p[k]; const float* g = t.g[k]; float* m = t.m[k]; float* v = t.v[k]; float* x = t.x[k];
  const float step_size = s_hyp[0], bc2s = s_hyp[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n[k]; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] * b1 + (1.0f - b1) * gi;
    const float vi = v[i] * b2 + (1.0f - b2) * (gi * gi);
    float vh = vi;
    if (amsgrad) { vh = fmaxf(x[i], vi); x[i] = vh; }
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vh) / bc2s + eps));
  }
}

__global__ void k_increment(int64_t* c) { *c += 1; }

// step += 1 and the two step-dependent Adam scalars for kernels that fuse the update (SpMM epilogue)
__global__ void k_adam_prepare(int64_t* step, float* hyp, float lr, float b1, float b2) {
  const int64_t t = *step + 1;
  *step = t;
  hyp[0] = (float)((double)lr / (1.0 - pow((double)b1, (double)t)));
  hyp[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}

__global__ void k_cast_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i]);
}

// ---- X = [I | F]: XW = W1[:N] (+ F @ W1[N:] on document rows) ----  text2graph.py:237-241
__global__ void __launch_bounds__(256) k_hier_forward(const float* __restrict__ W1, int64_t ldw, int64_t N, int64_t n_vocab,
                                                      const float* __restrict__ Fd, int64_t ldf, int c_prev, int H,
                                                      float* __restrict__ XW, int64_t ldxw) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* tail = W1 + N * ldw;
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int h = h0 + lane;
    float s = (h < H) ? W1[row * ldw + h] : 0.0f;
    if (row >= n_vocab) {
      const float* f = Fd + (row - n_vocab) * ldf;
      for (int c0 = 0; c0 < c_prev; c0 += 32) {      // lanes read 32 features at once; non-zeros found by ballot
        const int c = c0 + lane;
        const float fv = (c < c_prev) ? f[c] : 0.0f;
        unsigned msk = __ballot_sync(0xffffffffu, fv != 0.0f);
        while (msk) {
          const int src = __ffs(msk) - 1;
          msk &= msk - 1;
          const float fc = __shfl_sync(0xffffffffu, fv, src);
          if (h < H) s = fmaf(fc, tail[(int64_t)(c0 + src) * ldw + h], s);
        }
      }
    }
    if (h < H) XW[row * ldxw + h] = s;
  }
}

// dW1[N:, :] = F^T G1[docs].  Every WARP owns a strided subset of the documents and a private
// (c_prev x H) accumulator in shared memory (no block barrier per document); lanes cover H, the
// non-zeros of the F row are found by ballot (one-hot rows: a single update).  Per-warp partials are
// added in warp order at the end, per-CTA partials in CTA order by k_reduce_parts: deterministic.
__global__ void __launch_bounds__(256) k_hier_backward(const float* __restrict__ G1, int64_t ldg, int64_t n_vocab, int64_t n_docs,
                                                       const float* __restrict__ Fd, int64_t ldf, int c_prev, int H,
                                                       float* __restrict__ part, int warps_with_acc) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nw = warps_with_acc;                 // warps that own an accumulator (smem permitting)
  const int sz = c_prev * H;
  for (int i = tid; i < nw * sz; i += blockDim.x) smem[i] = 0.0f;
  __syncthreads();
  if (wid < nw) {
    float* acc = smem + wid * sz;
    const int64_t stride = (int64_t)gridDim.x * nw;
    for (int64_t d = (int64_t)blockIdx.x * nw + wid; d < n_docs; d += stride) {
      const float* g = G1 + (n_vocab + d) * ldg;
      for (int c0 = 0; c0 < c_prev; c0 += 32) {
        const int c = c0 + lane;
        const float f = (c < c_prev) ? Fd[d * ldf + c] : 0.0f;
        unsigned m = __ballot_sync(0xffffffffu, f != 0.0f);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float fv = __shfl_sync(0xffffffffu, f, src);
          float* a = acc + (c0 + src) * H;
          for (int h = lane; h < H; h += 32) a[h] = fmaf(fv, g[h], a[h]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < sz; i += blockDim.x) {
    float s = 0.0f;
    for (int w = 0; w < nw; ++w) s += smem[w * sz + i];
    part[(int64_t)blockIdx.x * sz + i] = s;
  }
}

// Y = dropout(X): the SpMM epilogue's dropout step as a stand-alone pass (same keep decision per element: keep-mask
// byte, or Philox keyed by (seed, offset, global row * F + col)).  Lets a pre-dropout activation computed once
// (eval forward of epoch k) serve the training forward of epoch k+1, whose A_hat (X W1) + b1 is identical because
// W1/b1 do not change in between (flat_amazon.py:100-110; dropout comes after the product, models.py:20-23).
__global__ void __launch_bounds__(256) k_dropout_apply(const float* __restrict__ X, int64_t ldx, float* __restrict__ Y, int64_t ldy,
                                                       int64_t n_rows, int F, int drop_mode, float drop_p, float scale,
                                                       const uint8_t* __restrict__ keep, int64_t ldmask, uint64_t seed,
                                                       uint64_t offset, const int64_t* __restrict__ offset_dev, int64_t row_offset) {
  const int FQ = F >> 2;
  const int64_t total = n_rows * FQ;
  const uint64_t ph_off = offset + (offset_dev ? (uint64_t)__ldg(offset_dev) : 0ull);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / FQ;
    const int c0 = (int)(i - row * FQ) << 2;
    float4 v = __ldg(reinterpret_cast<const float4*>(X + row * ldx + c0));
    if (drop_mode == TGCN_DROP_MASK) {
      const uint8_t* m = keep + row * ldmask + c0;
      v.x = m[0] ? v.x * scale : 0.0f; v.y = m[1] ? v.y * scale : 0.0f;
      v.z = m[2] ? v.z * scale : 0.0f; v.w = m[3] ? v.w * scale : 0.0f;
    } else if (drop_mode == TGCN_DROP_PHILOX) {
      const uint64_t e4 = ((uint64_t)(row + row_offset) * (uint64_t)F + (uint64_t)c0) >> 2;
      const uint4 r = philox_quad(e4, seed, ph_off);
      v.x = (u01(r.x) >= drop_p) ? v.x * scale : 0.0f; v.y = (u01(r.y) >= drop_p) ? v.y * scale : 0.0f;
      v.z = (u01(r.z) >= drop_p) ? v.z * scale : 0.0f; v.w = (u01(r.w) >= drop_p) ? v.w * scale : 0.0f;
    }
    *reinterpret_cast<float4*>(Y + row * ldy + c0) = v;
  }
}

__global__ void k_reduce_parts(const float* __restrict__ part, int n_parts, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int c = 0; c < n_parts; ++c) s += part[(int64_t)c * n + i];
  out[i] = s;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" const char* tgcn_last_error(void) { return g_err; }
extern "C" int tgcn_version(void) { return 100; }
extern "C" uint64_t tgcn_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int tgcn_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  TGCN_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  TGCN_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  TGCN_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  TGCN_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm) *sm = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return TGCN_OK;
}

extern "C" int tgcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                              int64_t n, float lr, float beta1, float beta2, float eps, int32_t amsgrad,
                              int64_t step, const int64_t* step_dev, void* param_mirror_mc, void* stream_) {
  TGCN_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
  TGCN_CHECK_ARG(!amsgrad || max_exp_avg_sq, "adam_step: amsgrad needs max_exp_avg_sq");
  TGCN_CHECK_ARG(n >= 0, "adam_step: n < 0");
  TGCN_CHECK_ARG(step_dev != nullptr || step >= 1, "adam_step: step must be >= 1");
  if (n == 0) return TGCN_OK;
  const int T = 256;
  const int64_t blocks = std::min<int64_t>(cdiv(cdiv(n, 4), T), (int64_t)sm_count() * 8);
  k_adam<<<(unsigned)std::max<int64_t>(blocks, 1), T, 0, (cudaStream_t)stream_>>>(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n, lr,
                                                                                 beta1, beta2, eps, amsgrad, step, step_dev, (float*)param_mirror_mc);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_adam_step_small(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                                    float* const* exp_avg_sq, float* const* max_exp_avg_sq, const int64_t* sizes, float lr,
                                    float beta1, float beta2, float eps, int32_t amsgrad, int64_t step, const int64_t* step_dev,
                                    void* stream_) {
  TGCN_CHECK_ARG(n_tensors >= 1 && n_tensors <= 4, "adam_step_small: 1..4 tensors");
  TGCN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && sizes, "adam_step_small: null pointer");
  TGCN_CHECK_ARG(!amsgrad || max_exp_avg_sq, "adam_step_small: amsgrad needs max_exp_avg_sq");
  TGCN_CHECK_ARG(step_dev != nullptr || step >= 1, "adam_step_small: step must be >= 1");
  AdamSmall t;
  int64_t nmax = 0;
  for (int i = 0; i < 4; ++i) {
    const bool on = i < n_tensors;
    t.p[i] = on ? params[i] : nullptr; t.g[i] = on ? grads[i] : nullptr; t.m[i] = on ? exp_avg[i] : nullptr;
    t.v[i] = on ? exp_avg_sq[i] : nullptr; t.x[i] = (on && amsgrad) ? max_exp_avg_sq[i] : nullptr; t.n[i] = on ? sizes[i] : 0;
    if (on) { TGCN_CHECK_ARG(t.p[i] && t.g[i] && t.m[i] && t.v[i] && (!amsgrad || t.x[i]), "adam_step_small: null tensor"); }
    nmax = std::max(nmax, t.n[i]);
  }
  if (nmax == 0) return TGCN_OK;
  dim3 grid((unsigned)std::min<int64_t>(cdiv(nmax, 256), 64), n_tensors);
  k_adam_small<<<grid, 256, 0, (cudaStream_t)stream_>>>(t, lr, beta1, beta2, eps, amsgrad, step, step_dev);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_adam_prepare(int64_t* step_dev, float* hyper_dev, float lr, float beta1, float beta2, void* stream_) {
  TGCN_CHECK_ARG(step_dev && hyper_dev, "adam_prepare: null pointer");
  k_adam_prepare<<<1, 1, 0, (cudaStream_t)stream_>>>(step_dev, hyper_dev, lr, beta1, beta2);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_increment_step(int64_t* step_dev, void* stream_) {
  TGCN_CHECK_ARG(step_dev != nullptr, "increment_step: null pointer");
  k_increment<<<1, 1, 0, (cudaStream_t)stream_>>>(step_dev);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream_) {
  TGCN_CHECK_ARG(src && dst && n >= 0, "cast: bad arguments");
  if (n == 0) return TGCN_OK;
  const int T = 256;
  const int64_t blocks = std::min<int64_t>(cdiv(n, T), (int64_t)sm_count() * 16);
  k_cast_bf16<<<(unsigned)blocks, T, 0, (cudaStream_t)stream_>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_dropout_apply(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n_rows, int32_t F,
                                  int32_t drop_mode, float drop_p, const uint8_t* keep_mask, int64_t ldmask,
                                  uint64_t philox_seed, uint64_t philox_offset, const int64_t* philox_offset_dev,
                                  int64_t philox_row_offset, void* stream_) {
  TGCN_CHECK_ARG(X && Y && n_rows >= 0 && F > 0, "dropout_apply: bad arguments");
  TGCN_CHECK_ARG(F % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldx >= F && ldy >= F && (((uintptr_t)X | (uintptr_t)Y) & 15) == 0,
                 "dropout_apply: F and the row pitches must be multiples of 4 floats, the buffers 16-byte aligned");
  TGCN_CHECK_ARG(drop_mode >= TGCN_DROP_NONE && drop_mode <= TGCN_DROP_PHILOX, "dropout_apply: bad drop_mode");
  TGCN_CHECK_ARG(drop_mode == TGCN_DROP_NONE || (drop_p >= 0.0f && drop_p < 1.0f), "dropout_apply: p must be in [0,1)");
  TGCN_CHECK_ARG(drop_mode != TGCN_DROP_MASK || keep_mask, "dropout_apply: TGCN_DROP_MASK needs keep_mask");
  if (n_rows == 0) return TGCN_OK;
  const int mode = (drop_mode != TGCN_DROP_NONE && drop_p > 0.0f) ? drop_mode : TGCN_DROP_NONE;
  const int T = 256;
  const int64_t blocks = std::min<int64_t>(cdiv(n_rows * (F / 4), T), (int64_t)sm_count() * 16);
  k_dropout_apply<<<(unsigned)blocks, T, 0, (cudaStream_t)stream_>>>(X, ldx, Y, ldy, n_rows, F, mode, drop_p, 1.0f / (1.0f - drop_p),
                                                                     keep_mask, ldmask, philox_seed, philox_offset, philox_offset_dev,
                                                                     philox_row_offset);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_hier_forward(const float* W1, int64_t ldw, int64_t N, int64_t n_vocab, const float* Fdoc, int64_t ldf,
                                 int32_t c_prev, int32_t H, float* XW, int64_t ldxw, void* stream_) {
  TGCN_CHECK_ARG(W1 && XW && (c_prev == 0 || Fdoc), "hier_forward: null pointer");
  TGCN_CHECK_ARG(N > 0 && n_vocab >= 0 && n_vocab <= N && H > 0 && c_prev >= 0, "hier_forward: bad shape");
  TGCN_CHECK_ARG(ldw >= H && ldxw >= H && (c_prev == 0 || ldf >= c_prev), "hier_forward: leading dimension too small");
  const int T = 256;
  k_hier_forward<<<(unsigned)cdiv(N * 32, T), T, 0, (cudaStream_t)stream_>>>(W1, ldw, N, n_vocab, Fdoc, ldf, c_prev, H, XW, ldxw);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

static int hier_grid(int64_t n_docs) { return (int)std::min<int64_t>(std::max<int64_t>(n_docs / 64, 1), (int64_t)sm_count() * 2); }

extern "C" int tgcn_hier_backward_workspace_bytes(int32_t c_prev, int32_t H, size_t* bytes_out) {
  TGCN_CHECK_ARG(bytes_out && c_prev > 0 && H > 0, "hier_backward_workspace_bytes: bad arguments");
  *bytes_out = (size_t)sm_count() * 2 * c_prev * H * sizeof(float);
  return TGCN_OK;
}

extern "C" int tgcn_hier_backward(const float* G1, int64_t ldg, int64_t N, int64_t n_vocab, const float* Fdoc, int64_t ldf,
                                  int32_t c_prev, int32_t H, float* dW_tail, void* workspace, size_t workspace_bytes,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(G1 && Fdoc && dW_tail, "hier_backward: null pointer");
  TGCN_CHECK_ARG(N > 0 && n_vocab >= 0 && n_vocab < N && H > 0 && c_prev > 0, "hier_backward: bad shape");
  const int64_t n_docs = N - n_vocab;
  const int grid = hier_grid(n_docs);
  const size_t need = (size_t)grid * c_prev * H * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_error("hier_backward workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return TGCN_EWORKSPACE;
  }
  const size_t sz_bytes = (size_t)c_prev * H * sizeof(float);
  TGCN_CHECK_ARG(sz_bytes <= 200 * 1024, "hier_backward: c_prev*H too large for shared memory");
  int nw = (int)std::min<size_t>(8, (200 * 1024) / sz_bytes);
  if (nw < 1) nw = 1;
  const size_t smem = (size_t)nw * sz_bytes;
  if (smem > 48 * 1024) TGCN_CUDA(cudaFuncSetAttribute(k_hier_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_hier_backward<<<grid, 256, smem, stream>>>(G1, ldg, n_vocab, n_docs, Fdoc, ldf, c_prev, H, (float*)workspace, nw);
  TGCN_LAUNCH_CHECK();
  const int T = 256;
  k_reduce_parts<<<(unsigned)cdiv((int64_t)c_prev * H, T), T, 0, stream>>>((const float*)workspace, grid, (int64_t)c_prev * H, dW_tail);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}
