// Dense-tile part of the hybrid propagation C = A_hat B on the 5th-generation tensor cores
// (tcgen05.mma with the accumulator in tensor memory, operands staged by bulk asynchronous copies).
//
// Replaces, for the DENSE blocks of A_hat, the per-non-zero gathers of tgcn_spmm (GCNConv.propagate:
// index_select + mul + scatter_add, textgcn/lib/models.py:20).  Why: the gather kernel moves one 4*F-byte
// operand row from L2 into an SM for every non-zero and is bound by that L2->SM fill path (12.2 clk per
// non-zero per SM at F = 200).  Word-word PMI graphs are far from uniformly sparse: once the
// nodes are ranked by degree, the blocks that pair hub words with anything hold half of the non-zeros at
// 5-40 % density.  A 128 x 32 block of that kind is cheaper as a dense matrix product: its operand tile
// (32 rows of B) crosses L2->SM once for 128 output rows, and the multiply-adds run on the tensor pipe,
// which the gather kernel leaves idle.
//
// Arithmetic: fp32 accuracy from TF32 tensor-core passes (3xTF32).  a = a_hi + a_lo, b = b_hi + b_lo with
// *_hi = cvt.rna.tf32(*) and *_lo = * - *_hi (exact); D += a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32
// accumulation in TMEM.  Dropped: a_lo b_lo and the truncation of the *_lo parts, each <= 2^-21 relative
// per product -- below the fp32 rounding of the 349-term sums themselves (tests/test_gpu_spmm_tc.py).
//
// Data (built once per graph by pytextgcn_b200/tc_plan.py, rank = position in the degree order):
//   A_tiles  [n_tiles][128 rows][16 cols] fp32 values of A_hat, each 64-byte row stored with the 64-byte
//            shared-memory swizzle already applied (16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)), so a
//            tile is ONE 8 KB bulk copy; four "splitter" warps turn it into the hi / lo operand tiles in
//            tensor memory -- half the HBM / L2->SM bytes of shipping hi and lo separately;
//   tile_kb  [n_tiles] column block (16 ranks) of each tile; tiles of one row block are consecutive;
//   units    {tile_begin, tile_end, slot, row_block}: <= 96 tiles of one row block; a unit's 128 x F partial
//            result goes to part[slot]; the slots of a row block are consecutive and tgcn_spmm's epilogue
//            adds them, in slot order, to the gathered remainder of the row (deterministic);
//   Bt       [n_col_blocks][2 (hi, lo)][Fp features][16 ranks] fp32, same swizzle: the operand transposed
//            to K-major and split into hi/lo by k_tc_pack at every launch (B changes every step).
//
// Kernel k_tc_mma: persistent, one CTA per SM, 320 threads.  warp 0 = producer (cp.async.bulk into a 5-6 stage
// ring {A values, Bt tile}, mbarrier complete_tx; A tiles prefetched into L2 ahead of time), warp 1 = MMA issuer
// (one thread; A operand from TMEM, B operand from shared memory; tcgen05.commit releases the stage and, after
// the last tile of a unit, publishes the accumulator), warps 2-5 = epilogue (tcgen05.ld 32 lanes x 32 columns
// per warp -> fp32 partial rows in global memory), warps 6-9 = splitters (A values from shared memory -> TF32
// hi + residual lo, tcgen05.st into TMEM).  Shared-memory bandwidth bounds 3xTF32 (every pass re-reads its
// operands): with A in TMEM only the Bt tiles are read by the tensor core.  Two accumulators in TMEM: the
// epilogue of unit i overlaps the MMAs of unit i+1.
#include "common.cuh"

namespace tgcn {

constexpr int TC_M = 128;
constexpr int TC_K = 16;             // 64-byte operand rows (SWIZZLE_64B): small stages, so that 4-5 of them are in flight
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_PREFETCH = 8;       // A tiles requested into L2 this many tiles ahead of their bulk copy
constexpr int TC_ABUF = 3;           // {A hi, A lo} operand tiles in TMEM: the splitters run up to 2 tiles ahead of the tensor pipe
constexpr int TC_THREADS = 320;
constexpr uint32_t TC_A_PART = TC_M * TC_K * 4;      // 8 KB: the values of a tile, or one of its (hi, lo) operand tiles
constexpr uint32_t TC_TMEM_COLS = 512;               // all of the SM's tensor memory: accumulator(s) + two {A hi, A lo} tiles

struct TcParams {
  const float* __restrict__ A_tiles;
  const int32_t* __restrict__ tile_kb;
  const int4* __restrict__ units;
  int32_t n_units;
  const float* __restrict__ Bt;
  float* part; int64_t ldp;
  int32_t Fp;      // MMA N: F rounded up to a multiple of 16 (also the TMEM column stride of the accumulators)
  int32_t n_acc;   // accumulators in TMEM: 2 (epilogue of unit i overlaps the MMAs of unit i+1) when 2 Fp + 64 <= 512, else 1
  int32_t F;       // columns stored
  int32_t n_stages;
  int32_t n_tiles;
};

// ---- PTX helpers: mbarrier, bulk copy, tcgen05 ----
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem], TF32 inputs, fp32 accumulate; issued by ONE thread for the CTA.  The A operand is
// read from tensor memory (lane = row, one 32-bit column per K element): the A tile is then not re-read from shared
// memory by each of the three passes -- shared-memory bandwidth is what bounds 3xTF32
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
               "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor: K-major operand, 64-byte rows, SWIZZLE_64B, 8-row groups 512 bytes apart
// (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor; version 1 = Blackwell)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major layouts)
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version
  d |= (uint64_t)4 << 61;                            // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void tc_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
#define TC_LD_REGS8(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : TC_LD_REGS8(v, 0), TC_LD_REGS8(v, 8), TC_LD_REGS8(v, 16), TC_LD_REGS8(v, 24) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : TC_LD_REGS8(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_mma(const TcParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // SWIZZLE_128B operands need 1024-byte aligned tiles: align the dynamic shared memory by hand
  const uint32_t raw = tc_smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t b_part = (uint32_t)p.Fp * (TC_K * 4u);        // bytes of one of (hi, lo) of an operand tile
  const int S = p.n_stages;
  const uint32_t stage_bytes = TC_A_PART + 2u * b_part;        // ring stage: A values, Bt hi, Bt lo
  const uint32_t bars = base + (uint32_t)S * stage_bytes;      // 8-byte mbarriers
  const uint32_t full0 = bars, empty0 = bars + 8 * TC_MAX_STAGES, afull0 = empty0 + 8 * TC_MAX_STAGES, aempty0 = afull0 + 8 * TC_ABUF,
                 tfull0 = aempty0 + 8 * TC_ABUF, tempty0 = tfull0 + 16;
  const uint32_t holder = tempty0 + 16;                        // TMEM base address written by tcgen05.alloc
  volatile uint32_t* holder_ptr = reinterpret_cast<volatile uint32_t*>(tc_smem_raw + (holder - raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      tc_mbar_init(full0 + 8 * s, 1);          // producer's arrive.expect_tx (+ the bytes of both bulk copies)
      tc_mbar_init(empty0 + 8 * s, 1 + 128);   // tcgen05.commit (Bt read) + the 128 splitter threads (A values read)
    }
    for (int a = 0; a < TC_ABUF; ++a) {
      tc_mbar_init(afull0 + 8 * a, 128);       // splitter threads: hi / lo tiles stored to TMEM
      tc_mbar_init(aempty0 + 8 * a, 1);        // tcgen05.commit: hi / lo tiles read
    }
    for (int a = 0; a < 2; ++a) { tc_mbar_init(tfull0 + 8 * a, 1); tc_mbar_init(tempty0 + 8 * a, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // one warp allocates all 512 columns of this SM's tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(holder), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder_ptr;
  const uint32_t tmem_a0 = tmem_base + (uint32_t)(p.n_acc * p.Fp);   // TC_ABUF {hi[16 cols], lo[16 cols]} A tiles behind the accumulators

  if (warp == 0) {
    // ---------------- producer: one thread issues two bulk copies per tile ----------------
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int4 un = __ldg(p.units + u);
        for (int t = un.x; t < min(un.y, un.x + TC_PREFETCH); ++t) tc_prefetch_l2(p.A_tiles + (int64_t)t * (TC_M * TC_K), TC_A_PART);
        for (int t = un.x; t < un.y; ++t, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          if (t + TC_PREFETCH < p.n_tiles) tc_prefetch_l2(p.A_tiles + (int64_t)(t + TC_PREFETCH) * (TC_M * TC_K), TC_A_PART);
          tc_mbar_wait(empty0 + 8 * s, ph ^ 1u);               // MMAs and splitters are done with this stage
          const uint32_t st = base + s * stage_bytes;
          tc_mbar_arrive_expect_tx(full0 + 8 * s, stage_bytes);
          tc_bulk_g2s(st, p.A_tiles + (int64_t)t * (TC_M * TC_K), TC_A_PART, full0 + 8 * s);
          const int kb = __ldg(p.tile_kb + t);
          tc_bulk_g2s(st + TC_A_PART, reinterpret_cast<const char*>(p.Bt) + (int64_t)kb * (2 * b_part), 2u * b_part, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: a single thread drives the tensor core ----------------
    if (lane == 0) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D fp32, A/B TF32, both K-major, N = Fp, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Fp >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      int it = 0, ui = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const int4 un = __ldg(p.units + u);
        if (un.x >= un.y) continue;
        const int as = (p.n_acc == 2) ? (ui & 1) : 0;
        const uint32_t aph = (uint32_t)((p.n_acc == 2) ? (ui >> 1) : ui) & 1u;
        tc_mbar_wait(tempty0 + 8 * as, aph ^ 1u);              // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.Fp);
        for (int t = un.x; t < un.y; ++t, ++it) {
          const int s = it % S, a = it % TC_ABUF;
          const uint32_t ph = (uint32_t)(it / S) & 1u, sph = (uint32_t)(it / TC_ABUF) & 1u;
          tc_mbar_wait(full0 + 8 * s, ph);                     // the Bt tile has landed
          tc_mbar_wait(afull0 + 8 * a, sph);                   // the splitters have stored A hi / lo to TMEM
          tc_fence_after();
          const uint32_t st = base + s * stage_bytes;
          const uint32_t a_hi = tmem_a0 + (uint32_t)a * (2 * TC_K), a_lo = a_hi + TC_K;
          const uint64_t b_hi = tc_smem_desc(st + TC_A_PART), b_lo = tc_smem_desc(st + TC_A_PART + b_part);
#pragma unroll
          for (int ks = 0; ks < TC_K / 8; ++ks) {              // 8 TF32 columns per instruction
            const uint64_t adv = (uint64_t)(ks * 2);           // B: +32 bytes on the 16-byte start-address field
            const uint32_t ac = (uint32_t)(ks * 8);            // A: +8 TMEM columns
            tc_mma_tf32_ts(d_tmem, a_hi + ac, b_hi + adv, idesc, (t > un.x || ks > 0) ? 1u : 0u);
            tc_mma_tf32_ts(d_tmem, a_hi + ac, b_lo + adv, idesc, 1u);
            tc_mma_tf32_ts(d_tmem, a_lo + ac, b_hi + adv, idesc, 1u);
          }
          tc_commit(empty0 + 8 * s);                           // ring stage reusable once these MMAs are done
          tc_commit(aempty0 + 8 * a);                          // so are the hi / lo tiles
        }
        tc_commit(tfull0 + 8 * as);                            // accumulator complete
        ++ui;
      }
    }
  } else if (warp >= 6) {
    // ---------------- splitters: A values (shared memory) -> TF32 hi + residual lo (tensor memory) ----------------
    // thread = row of the tile = TMEM lane (warps 6..9 own the lane quarters 2, 3, 0, 1); a 64-byte row is 4 chunks of
    // 16 bytes, logical chunk c stored at chunk c ^ ((row >> 1) & 3) (the swizzle keeps these row-wise reads conflict-free)
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t xr = (uint32_t)((row >> 1) & 3);
    int it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int4 un = __ldg(p.units + u);
      for (int t = un.x; t < un.y; ++t, ++it) {
        const int s = it % S, a = it % TC_ABUF;
        const uint32_t ph = (uint32_t)(it / S) & 1u, sph = (uint32_t)(it / TC_ABUF) & 1u;
        tc_mbar_wait(full0 + 8 * s, ph);                       // the tile's values have landed
        const uint32_t src = base + s * stage_bytes + (uint32_t)row * (TC_K * 4);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + (((uint32_t)c ^ xr) << 4)));
          const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi[4 * c + j]) : "f"(f[j]));
            lo[4 * c + j] = __float_as_uint(f[j] - __uint_as_float(hi[4 * c + j]));
          }
        }
        tc_mbar_arrive(empty0 + 8 * s);                        // this thread is done with the stage's A values
        tc_mbar_wait(aempty0 + 8 * a, sph ^ 1u);               // the MMAs that read the previous tiles in this TMEM buffer are done
        tc_fence_after();
        const uint32_t ta = tmem_a0 + (uint32_t)a * (2 * TC_K) + ((uint32_t)(q * 32) << 16);
        tc_st16(ta, hi);
        tc_st16(ta + TC_K, lo);
        tc_wait_st();
        tc_fence_before();
        tc_mbar_arrive(afull0 + 8 * a);
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> partial rows in global memory ----------------
    const int q = warp & 3;                                    // TMEM lanes [32q, 32q + 32) belong to warps with id % 4 == q
    int ui = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const int4 un = __ldg(p.units + u);
      if (un.x >= un.y) continue;
      const int as = (p.n_acc == 2) ? (ui & 1) : 0;
      const uint32_t aph = (uint32_t)((p.n_acc == 2) ? (ui >> 1) : ui) & 1u;
      tc_mbar_wait(tfull0 + 8 * as, aph);
      tc_fence_after();
      const int row = q * 32 + lane;
      float* dst = p.part + ((int64_t)un.z * TC_M + row) * p.ldp;
      const uint32_t taddr = tmem_base + (uint32_t)(as * p.Fp) + ((uint32_t)(q * 32) << 16);
      int c = 0;
      for (; c + 32 <= p.Fp; c += 32) {
        uint32_t v[32];
        tc_ld32(taddr + (uint32_t)c, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c + 4 * j < p.F)
            *reinterpret_cast<float4*>(dst + c + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
      for (; c < p.Fp; c += 8) {
        uint32_t v[8];
        tc_ld8(taddr + (uint32_t)c, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (c + 4 * j < p.F)
            *reinterpret_cast<float4*>(dst + c + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                      __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
      tc_fence_before();
      tc_mbar_arrive(tempty0 + 8 * as);                        // 128 arrivals hand the accumulator back
      ++ui;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// Bt[kb][hi|lo][f][k] = split(B[perm[16 kb + k]][f]), swizzled: transposes the operand to K-major (the feature index
// becomes the tile row, the 16 ranks of a column block its 64-byte row) and splits it into TF32 hi / residual lo.
__global__ void __launch_bounds__(256) k_tc_pack(const float* __restrict__ B, int64_t ldb, const int32_t* __restrict__ perm,
                                                 int F, int Fp, float* __restrict__ Bt) {
  extern __shared__ float tc_pack_smem[];          // [TC_K][F + 1]
  const int kb = blockIdx.x;
  const int FS = F + 1;
  const int FQ = F >> 2;
  for (int i = threadIdx.x; i < TC_K * FQ; i += blockDim.x) {
    const int k = i / FQ, fq = i - k * FQ;
    const int node = __ldg(perm + kb * TC_K + k);                   // -1: rank past the last node (zero column)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (node >= 0) v = __ldg(reinterpret_cast<const float4*>(B + (int64_t)node * ldb) + fq);
    float* d = tc_pack_smem + k * FS + 4 * fq;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  char* out_hi = reinterpret_cast<char*>(Bt) + (int64_t)kb * (2 * Fp * TC_K * 4);
  char* out_lo = out_hi + Fp * TC_K * 4;
  for (int i = threadIdx.x; i < Fp * (TC_K / 4); i += blockDim.x) {
    const int f = i / (TC_K / 4), ch = i % (TC_K / 4);
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = (f < F) ? tc_pack_smem[(4 * ch + j) * FS + f] : 0.0f;
      uint32_t h;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
      hi[j] = __uint_as_float(h);
      lo[j] = v - hi[j];
    }
    const int off = f * (TC_K * 4) + ((ch ^ ((f >> 1) & 3)) << 4);      // SWIZZLE_64B: 16-byte chunk ^= bits [1,3) of the row
    *reinterpret_cast<float4*>(out_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(out_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

}  // namespace tgcn

using namespace tgcn;

extern "C" int tgcn_spmm_tc_workspace_elems(int32_t F, int64_t n_col_blocks, int64_t* bt_elems_out) {
  TGCN_CHECK_ARG(bt_elems_out && F > 0 && n_col_blocks >= 0, "spmm_tc_workspace_elems: bad arguments");
  const int64_t Fp = (F + 15) / 16 * 16;
  *bt_elems_out = n_col_blocks * 2 * Fp * TC_K;
  return TGCN_OK;
}

extern "C" int tgcn_spmm_tc(const tgcn_tc_plan* plan, const float* B, int64_t ldb, int32_t F, float* Bt, float* part,
                            int64_t ldp, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TGCN_CHECK_ARG(plan && B && Bt && part, "spmm_tc: null pointer");
  TGCN_CHECK_ARG(plan->A_tiles && plan->tile_kb && plan->units && plan->perm, "spmm_tc: plan pointer null");
  TGCN_CHECK_ARG(F >= 8 && F <= 256 && F % 4 == 0, "spmm_tc: F (%d) must be a multiple of 4 in [8, 256]", F);
  TGCN_CHECK_ARG(ldb % 4 == 0 && ldb >= F && ((uintptr_t)B & 15) == 0, "spmm_tc: B must be 16-byte aligned with ldb %% 4 == 0");
  TGCN_CHECK_ARG(ldp % 4 == 0 && ldp >= F && ((uintptr_t)part & 15) == 0, "spmm_tc: part must be 16-byte aligned with ldp %% 4 == 0");
  TGCN_CHECK_ARG((((uintptr_t)plan->A_tiles | (uintptr_t)Bt) & 127) == 0, "spmm_tc: A_tiles / Bt must be 128-byte aligned");
  TGCN_CHECK_ARG(plan->n_units >= 0 && plan->n_col_blocks > 0, "spmm_tc: bad plan sizes");
  if (plan->n_units == 0) return TGCN_OK;
  const int Fp = (F + 15) / 16 * 16;      // tcgen05.mma with A in TMEM and M = 128: N must be a multiple of 16
  {
    const size_t smem = (size_t)TC_K * (F + 1) * sizeof(float);
    if (smem > 48 * 1024) TGCN_CUDA(cudaFuncSetAttribute(k_tc_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_pack<<<(unsigned)plan->n_col_blocks, 256, smem, stream>>>(B, ldb, plan->perm, F, Fp, Bt);
    TGCN_LAUNCH_CHECK();
  }
  TcParams p;
  p.A_tiles = plan->A_tiles; p.tile_kb = plan->tile_kb; p.units = reinterpret_cast<const int4*>(plan->units);
  p.n_units = plan->n_units; p.Bt = Bt; p.part = part; p.ldp = ldp; p.Fp = Fp; p.F = F;
  const size_t stage = (size_t)TC_A_PART + 2 * (size_t)Fp * TC_K * 4;
  const size_t fixed = 1024 /* alignment slack */ + 512 /* barriers */;
  const int n_stages = (int)std::min<size_t>(TC_MAX_STAGES, (227 * 1024 - fixed) / stage);
  p.n_acc = (2 * Fp + TC_ABUF * 2 * TC_K <= (int)TC_TMEM_COLS) ? 2 : 1;
  TGCN_CHECK_ARG(n_stages >= 2, "spmm_tc: F = %d does not leave room for two pipeline stages", F);
  const size_t smem = n_stages * stage + fixed;
  p.n_stages = n_stages; p.n_tiles = plan->n_tiles;
  TGCN_CUDA(cudaFuncSetAttribute(k_tc_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = std::min<int>(sm_count(), plan->n_units);
  k_tc_mma<<<grid, TC_THREADS, smem, stream>>>(p);
  TGCN_LAUNCH_CHECK();
  return TGCN_OK;
}
