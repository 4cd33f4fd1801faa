// Exchange step of the row-partitioned path over NVLink peer memory: instead of an
// ncclAllGather, every rank STORES its freshly produced row slice straight into the operand buffer
// of every peer -- one multimem.st per 16 bytes through the NVSwitch multicast address when the
// fabric offers it (the switch replicates: each rank sends its slice once), plain peer stores
// otherwise -- followed by a device-side barrier (torch symmetric-memory signal pads).  The buffers
// are symmetric allocations whose peer / multicast addresses come from the host
// (torch.distributed._symmetric_memory rendezvous); nothing is allocated here.
#include "common.cuh"

namespace tgcn {

struct PeerPtrs { void* p[16]; };

__device__ __forceinline__ void multimem_st_v4(float4* addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256) k_peer_push(const float4* __restrict__ src, PeerPtrs peers, int world, int rank,
                                                   int64_t n_vec, int64_t dst_off_vec, float4* mc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const float4 v = src[i];
    if (mc) {
      multimem_st_v4(mc + dst_off_vec + i, v);
    } else {
      for (int p = 0; p < world; ++p)
        if (p != rank) reinterpret_cast<float4*>(peers.p[p])[dst_off_vec + i] = v;
    }
  }
  __threadfence_system();
}

__global__ void k_sum_slots(const float* __restrict__ slots, int n_slots, int64_t slot_stride, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int k = 0; k < n_slots; ++k) s += slots[(int64_t)k * slot_stride + i];     // rank order: identical on every rank
  out[i] = s;
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_peer_push(const void* src, void* const* peer_bases_host, int32_t world, int32_t rank, int64_t bytes,
                              int64_t dst_offset_bytes, void* multicast_base, void* stream_) {
  TGCN_CHECK_ARG(src != nullptr && (peer_bases_host != nullptr || multicast_base != nullptr), "peer_push: null pointer");
  TGCN_CHECK_ARG(world >= 1 && world <= 16 && rank >= 0 && rank < world, "peer_push: bad world/rank");
  TGCN_CHECK_ARG(bytes >= 0 && bytes % 16 == 0 && dst_offset_bytes % 16 == 0 && ((uintptr_t)src % 16) == 0,
                 "peer_push: sizes and offsets must be multiples of 16 bytes");
  if (bytes == 0 || world == 1) return TGCN_OK;
  PeerPtrs pp;
  for (int i = 0; i < 16; ++i) pp.p[i] = (peer_bases_host && i < world) ? peer_bases_host[i] : nullptr;
  const int64_t n_vec = bytes / 16;
  const int T = 256;
  const int64_t blocks = std::min<int64_t>(cdiv(n_vec, T), (int64_t)sm_count() * 4);
  k_peer_push<<<(unsigned)blocks, T, 0, (cudaStream_t)stream_>>>(reinterpret_cast<const float4*>(src), pp, world, rank, n_vec,
                                                                 dst_offset_bytes / 16, reinterpret_cast<float4*>(multicast_base));
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}

extern "C" int tgcn_sum_slots(const float* slots, int32_t n_slots, int64_t slot_stride, int64_t n, float* out, void* stream_) {
  TGCN_CHECK_ARG(slots && out && n_slots >= 1 && n >= 0 && slot_stride >= n, "sum_slots: bad arguments");
  if (n == 0) return TGCN_OK;
  const int T = 256;
  k_sum_slots<<<(unsigned)cdiv(n, T), T, 0, (cudaStream_t)stream_>>>(slots, n_slots, slot_stride, n, out);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}
