// Graph upload: COO (edge_index, edge_attr) -> CSR of A_hat = D^-1/2 (A+I) D^-1/2.
// Replaces gcn_norm / add_remaining_self_loops, which the reference re-runs inside every
// GCNConv call (textgcn/lib/models.py:11-15,20; [PyG-1.6.3] gcn_conv.py::gcn_norm).
// Bit-exact against torch CPU: the degree is a sequential fp32 sum in original edge order
// (self loop last), dis = 1/sqrt(deg) with IEEE div/sqrt, val = (dis[src]*w)*dis[dst].
//
// One-off per graph (not in the per-epoch path).  The stable sort by target row is CUB's
// DeviceRadixSort (a CUDA-toolkit library primitive, like cuBLAS for a plain GEMM); every
// other step is a kernel below.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace tgcn {

// keys: target row for every kept edge, n_nodes (sentinel, sorts last) for dropped self
// loops; ids: position in gcn_norm's edge list (0..E-1 original, E..E+N-1 loops).
__global__ void k_make_keys(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t stride,
                            int64_t E, int64_t N, int32_t* __restrict__ keys, int32_t* __restrict__ ids,
                            int32_t* __restrict__ loop_edge, int32_t* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = E + N;
  if (i >= total) return;
  ids[i] = (int32_t)i;
  if (i < E) {
    int64_t s = src[i * stride], d = dst[i * stride];
    if (s < 0 || s >= N || d < 0 || d >= N) {
      status[0] = TGCN_EINDEX;
      keys[i] = (int32_t)N;
      return;
    }
    if (s == d) {
      keys[i] = (int32_t)N;                 // dropped; its weight becomes the loop weight
      atomicMax(&loop_edge[s], (int32_t)i); // last occurrence wins (torch CPU index_put order)
    } else {
      keys[i] = (int32_t)d;
    }
  } else {
    keys[i] = (int32_t)(i - E);
  }
}

__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// rowptr from sorted keys: every row has >= 1 entry (its self loop), so row r starts at the
// first position whose key is r.
__global__ void k_rowptr(const int32_t* __restrict__ keys, int64_t total, int64_t N,
                         int32_t* __restrict__ rowptr, int32_t* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int32_t k = keys[i];
  int32_t kprev = (i == 0) ? -1 : keys[i - 1];
  if (k != kprev) {
    if (k <= N) rowptr[k] = (int32_t)i;     // k == N: start of the dropped tail == nnz
    if (k == N) status[1] = (int32_t)i;
  }
  if (i == total - 1 && k < N) { rowptr[N] = (int32_t)total; status[1] = (int32_t)total; }
}

__device__ __forceinline__ float entry_weight(int32_t id, int64_t E, const float* __restrict__ w,
                                              const int32_t* __restrict__ loop_edge) {
  if (id < E) return w ? w[id] : 1.0f;
  int32_t le = loop_edge[id - E];
  return (le >= 0 && w) ? w[le] : 1.0f;
}

// One warp per row: loads 32 weights at a time (coalesced over the sorted id list, gathered
// over w), then adds them one by one in lane order so the fp32 rounding sequence is exactly
// the sequential scatter_add_ of torch CPU.
__global__ void k_degree(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ ids, int64_t E, int64_t N,
                         const float* __restrict__ w, const int32_t* __restrict__ loop_edge,
                         float* __restrict__ dis) {
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= N) return;
  int32_t b = rowptr[row], e = rowptr[row + 1];
  float deg = 0.0f;
  for (int32_t base = b; base < e; base += 32) {
    int32_t k = base + lane;
    float wv = 0.0f;
    if (k < e) wv = entry_weight(ids[k], E, w, loop_edge);
    int cnt = min(32, e - base);
    for (int j = 0; j < cnt; ++j) deg = __fadd_rn(deg, __shfl_sync(0xffffffffu, wv, j));
  }
  if (lane == 0) {
    float d = __fdiv_rn(1.0f, __fsqrt_rn(deg));   // == torch CPU deg.pow_(-0.5), bit for bit
    if (isinf(d)) d = 0.0f;                        // masked_fill_(== inf, 0); NaN stays NaN
    dis[row] = d;
  }
}

__global__ void k_values(const int32_t* __restrict__ keys, const int32_t* __restrict__ ids, int64_t total,
                         int64_t E, int64_t N, const int64_t* __restrict__ src, int64_t stride,
                         const float* __restrict__ w, const int32_t* __restrict__ loop_edge,
                         const float* __restrict__ dis, int32_t* __restrict__ colidx, float* __restrict__ val,
                         int32_t* __restrict__ edge_slot) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int32_t id = ids[i];
  int32_t row = keys[i];
  if (row >= N) {                 // dropped self loop (sentinel key sorts last)
    if (edge_slot) edge_slot[id] = -1;
    return;
  }
  int32_t s = (id < E) ? (int32_t)src[(int64_t)id * stride] : (int32_t)(id - E);
  float wv = entry_weight(id, E, w, loop_edge);
  colidx[i] = s;
  val[i] = __fmul_rn(__fmul_rn(dis[s], wv), dis[row]);   // (dis[row']*w')*dis[col'], left to right
  if (edge_slot) edge_slot[id] = (int32_t)i;
}

struct CsrWs {
  int32_t *keys_in, *keys_out, *ids_in, *ids_out, *loop_edge;
  void* cub_tmp; size_t cub_bytes; size_t total_bytes;
};

static int csr_ws_layout(int64_t N, int64_t E, void* base, CsrWs* ws) {
  int64_t total = E + N;
  size_t cub_bytes = 0;
  int end_bit = 1;
  while ((1ll << end_bit) <= N) ++end_bit;   // keys in [0, N]
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                  (const int32_t*)nullptr, (int32_t*)nullptr, total,
                                                  0, end_bit, (cudaStream_t)0);
  if (e != cudaSuccess) { set_error("cub workspace query failed: %s", cudaGetErrorString(e)); return TGCN_ECUDA; }
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { void* r = p ? p + off : nullptr; off += align_up(bytes, 256); return r; };
  ws->keys_in = (int32_t*)take(total * 4);
  ws->keys_out = (int32_t*)take(total * 4);
  ws->ids_in = (int32_t*)take(total * 4);
  ws->ids_out = (int32_t*)take(total * 4);
  ws->loop_edge = (int32_t*)take((size_t)N * 4);
  ws->cub_tmp = take(cub_bytes);
  ws->cub_bytes = cub_bytes;
  ws->total_bytes = off;
  return TGCN_OK;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_csr_workspace_bytes(int64_t n_nodes, int64_t n_edges, size_t* bytes_out) {
  TGCN_CHECK_ARG(bytes_out != nullptr, "bytes_out is null");
  TGCN_CHECK_ARG(n_nodes > 0 && n_edges >= 0, "n_nodes must be > 0 and n_edges >= 0");
  TGCN_CHECK_ARG(n_nodes + n_edges < (int64_t)0x7fffffff, "n_edges + n_nodes must fit int32");
  CsrWs ws;
  int rc = csr_ws_layout(n_nodes, n_edges, nullptr, &ws);
  if (rc) return rc;
  *bytes_out = ws.total_bytes;
  return TGCN_OK;
}

extern "C" int tgcn_csr_from_coo_gcn_norm(const int64_t* edge_src, const int64_t* edge_dst, int64_t idx_stride,
                                          const float* edge_w, int64_t E, int64_t N,
                                          int32_t* rowptr, int32_t* colidx, float* val, float* dis,
                                          int32_t* edge_slot, int32_t* status_out,
                                          void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(N > 0 && E >= 0, "n_nodes must be > 0 and n_edges >= 0");
  TGCN_CHECK_ARG(N + E < (int64_t)0x7fffffff, "n_edges + n_nodes must fit int32");
  TGCN_CHECK_ARG(E == 0 || (edge_src && edge_dst), "edge_src/edge_dst null");
  TGCN_CHECK_ARG(idx_stride >= 1, "idx_stride must be >= 1");
  TGCN_CHECK_ARG(rowptr && colidx && val && dis && status_out, "output pointer null");
  CsrWs ws;
  int rc = csr_ws_layout(N, E, workspace, &ws);
  if (rc) return rc;
  if (workspace == nullptr || workspace_bytes < ws.total_bytes) {
    set_error("csr workspace too small: need %zu bytes, got %zu", ws.total_bytes, workspace_bytes);
    return TGCN_EWORKSPACE;
  }
  const int64_t total = E + N;
  const int T = 256;
  TGCN_CUDA(cudaMemsetAsync(status_out, 0, 2 * sizeof(int32_t), stream));
  k_fill_i32<<<(unsigned)cdiv(N, T), T, 0, stream>>>(ws.loop_edge, N, -1);
  TGCN_LAUNCH_CHECK();
  k_make_keys<<<(unsigned)cdiv(total, T), T, 0, stream>>>(edge_src, edge_dst, idx_stride, E, N, ws.keys_in, ws.ids_in,
                                                          ws.loop_edge, status_out);
  TGCN_LAUNCH_CHECK();
  int end_bit = 1;
  while ((1ll << end_bit) <= N) ++end_bit;
  size_t cub_bytes = ws.cub_bytes;
  TGCN_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, cub_bytes, (const int32_t*)ws.keys_in, ws.keys_out,
                                            (const int32_t*)ws.ids_in, ws.ids_out, total, 0, end_bit, stream));
  k_rowptr<<<(unsigned)cdiv(total, T), T, 0, stream>>>(ws.keys_out, total, N, rowptr, status_out);
  TGCN_LAUNCH_CHECK();
  k_degree<<<(unsigned)cdiv(N * 32, T), T, 0, stream>>>(rowptr, ws.ids_out, E, N, edge_w, ws.loop_edge, dis);
  TGCN_LAUNCH_CHECK();
  k_values<<<(unsigned)cdiv(total, T), T, 0, stream>>>(ws.keys_out, ws.ids_out, total, E, N, edge_src, idx_stride, edge_w,
                                                       ws.loop_edge, dis, colidx, val, edge_slot);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}
