// k_l2_candidates.cu -- all-pairs squared-L2 candidate generation on the 5th-gen tensor cores.
//
// Replaces the inner loop of ArrayMatcherBruteForce<float,L2<float>>::SearchNeighbours (the a6 row
// of SURVEY.md 8a; contract visible at src/utils/matcher_hnsw.h:133-191): for every query
// descriptor of image J, the distances to all descriptors of image I.
//
// Formulation.  With the operands built by k_view_prepare,
//     D[q, i] = opQ_J[q, :] . opD_I[i, :] = ||a_i||^2 + ||b_q||^2 - 2 a_i.b_q
// is a plain K-major x K-major GEMM with K = pad16(dim) + 16 (the norm terms ride in one extra
// UMMA K-step).  fp16 operands, fp32 accumulation in TMEM.
//
// The GEMM is only the CANDIDATE generator: the epilogue reduces every 16-column chunk of a
// query row to its minimum (FMNMX3 tree), packs the chunk id into the 12 low mantissa bits and
// keeps the 4 smallest packed keys per query in registers.  k_rerank then recomputes the
// candidate chunks exactly (upstream float order) and certifies the result against the 4th key
// (DESIGN.md "certification"); what cannot be certified goes to k_exact_scan.  Bit-exact output
// therefore never depends on tensor-core rounding.
//
// CTA organisation (one persistent CTA per SM, 384 threads):
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : tcgen05.mma issuer (one elected lane)
//   warp 2   : TMEM allocator
//   warp 3   : idle
//   warps 4-11: epilogue; warp w owns TMEM lanes 32*(w%4).., query block (w-4)/4
// A work item is (pair, 256-query super-block).  The 2x128 query rows stay resident in shared
// memory (A operand, M = 128 each); the database image streams through a ring of 128-row x 64-col
// TMA boxes (B operand, N = 128).  Two accumulator stages of 2 x (128 lanes x 128 fp32 columns)
// fill the 512 TMEM columns, so the MMAs of tile t+1 overlap the epilogue of tile t.
#include "r3d_internal.cuh"

#include <cstdlib>

namespace r3d {

namespace {

constexpr int kMaxStages = 12;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (4 + kEpiWarps);
constexpr uint32_t kBoxBytes = kTileRows * kKBlock * 2;  // 16384
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCols = 128;                        // fp32 columns per (stage, query block)
constexpr uint32_t kKeySentinel = 0x7f7fffffu;            // FLT_MAX

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled
// above the wait by the compiler.
__device__ __forceinline__ void tc_wait_ld(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :: "memory");
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// one lane of a converged warp (all 32 lanes must execute this)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (SM100 "version 1"):
//   start address >> 4 | LBO (ignored for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) |
//   version = 1 | layout type = SWIZZLE_128B (2)
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO | version | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ uint64_t make_desc(uint32_t lo) { return ((uint64_t)kDescHi << 32) | (uint64_t)lo; }
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, N = 128, M = 128.
constexpr uint32_t kInstrDesc = (1u << 4) | ((uint32_t)(kTileRows >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);

// minimum of kChunk accumulator columns, packed with the chunk id, inserted into the sorted key set
__device__ __forceinline__ void chunk_update(const uint32_t* v, uint32_t chunk_id, uint32_t keep_mask,
                                             float (&key)[kNumKeys]) {
  static_assert(kChunk == 8 || kChunk == 16, "chunk width");
  float m = fmin3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
  m = fmin3(m, __uint_as_float(v[3]), __uint_as_float(v[4]));
  m = fmin3(m, __uint_as_float(v[5]), __uint_as_float(v[6]));
  if (kChunk == 16) {
    m = fmin3(m, __uint_as_float(v[7]), __uint_as_float(v[8]));
    m = fmin3(m, __uint_as_float(v[9]), __uint_as_float(v[10]));
    m = fmin3(m, __uint_as_float(v[11]), __uint_as_float(v[12]));
    m = fmin3(m, __uint_as_float(v[13]), __uint_as_float(v[14]));
  }
  m = fminf(m, __uint_as_float(v[kChunk - 1]));
  float x = __uint_as_float((__float_as_uint(m) & keep_mask) | chunk_id);
#pragma unroll
  for (int i = 0; i < kNumKeys - 1; ++i) {  // sorted insertion network: 2 FMNMX per level
    const float hi = fmaxf(key[i], x);
    key[i] = fminf(key[i], x);
    x = hi;
  }
  key[kNumKeys - 1] = fminf(key[kNumKeys - 1], x);
}

struct SmemLayout {
  uint32_t q_base, d_base, bars;  // byte offsets from the 1024-aligned base
};

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
k_l2_candidates(const CUtensorMap* __restrict__ tmapQ, const CUtensorMap* __restrict__ tmapD,
                const CUtensorMap* __restrict__ tmapDh, const PairDesc* __restrict__ pairs, const WorkItem* __restrict__ items, uint32_t n_items,
                uint32_t* __restrict__ keys_out, uint32_t nkb, uint32_t ksteps, uint32_t n_stages) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = base;                                   // kQB * nkb boxes
  const uint32_t d_base = q_base + kQB * nkb * kBoxBytes;          // n_stages boxes
  const uint32_t bar_base = d_base + n_stages * kBoxBytes;         // 8-byte barriers
  const uint32_t bar_full = bar_base;                              // [kMaxStages]
  const uint32_t bar_empty = bar_base + 8 * kMaxStages;            // [kMaxStages]
  const uint32_t bar_qfull = bar_base + 16 * kMaxStages;
  const uint32_t bar_qempty = bar_qfull + 8;
  const uint32_t bar_tfull = bar_qempty + 8;                       // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                      // [2]
  const uint32_t tmem_slot = bar_tempty + 16;                      // uint32
  unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(gen_base + (tmem_slot - base));

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  // Cluster of 2 (optional): both CTAs stream the SAME database image (their work items are two
  // super-blocks of one pair); each CTA loads half of every database box and multicasts it to
  // both, which halves the L2 -> shared-memory traffic, the measured limiter of this kernel.
  const uint32_t ncta = cluster_nctarank();
  const uint32_t crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << ncta) - 1u);
  const uint32_t cluster_id = blockIdx.x / ncta;
  const uint32_t n_clusters = gridDim.x / ncta;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, ncta);  // released by the MMA warps of every CTA in the cluster
    }
    mbar_init(bar_qfull, 1);
    mbar_init(bar_qempty, 1);
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (ncta > 1) cluster_sync_all();  // peer barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // The whole warp runs the control flow (warp-uniform); one elected lane issues the copies.
    uint32_t stage = 0, phase = 0, qphase = 0;
    for (uint32_t it = cluster_id * ncta + crank; it < n_items; it += n_clusters * ncta) {
      const WorkItem wi = items[it];
      const PairDesc pd = pairs[wi.pair];
      const uint32_t sb = wi.sb & 0x7fffffffu;
      const CUtensorMap* mq = tmapQ + pd.slotJ;
      const CUtensorMap* md = (ncta > 1 ? tmapDh : tmapD) + pd.slotI;
      const uint32_t nboxes = (pd.nI_pad / kTileRows) * nkb;
      // Database boxes do not depend on the query tiles: run the ring ahead (it fills as the
      // previous item's MMAs retire) before blocking on the query buffer.
      const uint32_t ahead = nboxes < n_stages ? nboxes : n_stages;
      uint32_t b = 0, t = 0, kb = 0;
      for (;;) {
        if (b == ahead) {
          mbar_wait(bar_qempty, qphase ^ 1u);  // previous item's MMAs no longer read the query tiles
          qphase ^= 1u;
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_qfull, kQB * nkb * kBoxBytes);
            for (uint32_t qb = 0; qb < (uint32_t)kQB; ++qb)
              for (uint32_t k2 = 0; k2 < nkb; ++k2)
                tma_load_2d(q_base + (qb * nkb + k2) * kBoxBytes, mq, (int)(k2 * kKBlock),
                            (int)(sb * kSuperRows + qb * kTileRows), bar_qfull);
          }
          __syncwarp();
        }
        if (b == nboxes) break;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, kBoxBytes);  // own half + the peer's half
          if (ncta > 1)
            tma_load_2d_mc(d_base + stage * kBoxBytes + crank * (kBoxBytes / 2), md, (int)(kb * kKBlock),
                           (int)(t * kTileRows + crank * (kTileRows / 2)), bar_full + 8 * stage, cmask);
          else
            tma_load_2d(d_base + stage * kBoxBytes, md, (int)(kb * kKBlock), (int)(t * kTileRows),
                        bar_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        ++b;
        if (++kb == nkb) { kb = 0; ++t; }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    // Warp-uniform control flow; the tcgen05 instructions are issued by one elected lane.  All
    // descriptor words are uniform values (shared-memory offsets + loop counters).
    uint32_t stage = 0, phase = 0, acc = 0, accphase = 0, qf = 0;
    for (uint32_t it = cluster_id * ncta + crank; it < n_items; it += n_clusters * ncta) {
      const WorkItem wi = items[it];
      const PairDesc pd = pairs[wi.pair];
      const uint32_t ntiles = pd.nI_pad / kTileRows;
      mbar_wait(bar_qfull, qf);
      qf ^= 1u;
      tc_fence_after();
      for (uint32_t t = 0; t < ntiles; ++t) {
        mbar_wait(bar_tempty + 8 * acc, accphase ^ 1u);  // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t d0 = tmem_base + (acc * kQB + 0) * kAccCols;
        const uint32_t d1 = tmem_base + (acc * kQB + 1) * kAccCols;
        uint32_t ks_left = ksteps;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = desc_lo(d_base + stage * kBoxBytes);
            const uint32_t a0_lo = desc_lo(q_base + (0 * nkb + kb) * kBoxBytes);
            const uint32_t a1_lo = desc_lo(q_base + (1 * nkb + kb) * kBoxBytes);
            const uint32_t ks_here = ks_left < 4u ? ks_left : 4u;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
              if (k < ks_here) {  // +2 per K-step: 32 bytes inside the 128-byte swizzle row
                const uint32_t accum = (kb | k) != 0u ? 1u : 0u;
                tc_mma_f16(d0, make_desc(a0_lo + 2 * k), make_desc(b_lo + 2 * k), kInstrDesc, accum);
                tc_mma_f16(d1, make_desc(a1_lo + 2 * k), make_desc(b_lo + 2 * k), kInstrDesc, accum);
              }
            }
            if (ncta > 1) tc_commit_mc(bar_empty + 8 * stage, cmask);  // ring slot free in BOTH CTAs' view
            else tc_commit(bar_empty + 8 * stage);                     // frees the ring slot when these MMAs retire
            if (kb + 1 == nkb) tc_commit(bar_tfull + 8 * acc);  // accumulator stage complete -> epilogue
          }
          __syncwarp();
          ks_left -= 4u;
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        acc ^= 1u;
        if (acc == 0) accphase ^= 1u;
      }
      if (elect_one()) tc_commit(bar_qempty);  // query tiles may be overwritten
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const uint32_t ew = warp - 4u;
    const uint32_t qb = ew >> 2;
    const uint32_t lane_quarter = warp & 3u;  // TMEM lanes this warp may touch
    uint32_t acc = 0, accphase = 0;
    for (uint32_t it = cluster_id * ncta + crank; it < n_items; it += n_clusters * ncta) {
      const WorkItem wi = items[it];
      const PairDesc pd = pairs[wi.pair];
      const uint32_t ntiles = pd.nI_pad / kTileRows;
      float key[kNumKeys];
#pragma unroll
      for (int i = 0; i < kNumKeys; ++i) key[i] = __uint_as_float(kKeySentinel);
      const uint32_t keep_mask = ~((1u << pd.chunk_bits) - 1u);
      for (uint32_t t = 0; t < ntiles; ++t) {
        mbar_wait(bar_tfull + 8 * acc, accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((lane_quarter * 32u) << 16) + (acc * kQB + qb) * kAccCols;
        const uint32_t chunk0 = t * (kTileRows / kChunk);
        uint32_t va[32], vb[32];
        constexpr uint32_t kCpl = 32 / kChunk;  // chunks per 32-column TMEM load
        tc_ld32(taddr, va);
        tc_wait_ld(va);
        tc_ld32(taddr + 32, vb);
#pragma unroll
        for (uint32_t c = 0; c < kCpl; ++c) chunk_update(va + c * kChunk, chunk0 + c, keep_mask, key);
        tc_wait_ld(vb);
        tc_ld32(taddr + 64, va);
#pragma unroll
        for (uint32_t c = 0; c < kCpl; ++c) chunk_update(vb + c * kChunk, chunk0 + kCpl + c, keep_mask, key);
        tc_wait_ld(va);
        tc_ld32(taddr + 96, vb);
#pragma unroll
        for (uint32_t c = 0; c < kCpl; ++c) chunk_update(va + c * kChunk, chunk0 + 2 * kCpl + c, keep_mask, key);
        tc_wait_ld(vb);
        // all TMEM reads of this stage are done: hand it back before the last chunks
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
#pragma unroll
        for (uint32_t c = 0; c < kCpl; ++c) chunk_update(vb + c * kChunk, chunk0 + 3 * kCpl + c, keep_mask, key);
        acc ^= 1u;
        if (acc == 0) accphase ^= 1u;
      }
      if (wi.sb & 0x80000000u) continue;  // padding item (odd super-block count): nothing to store
      const uint32_t row = wi.sb * kSuperRows + qb * kTileRows + lane_quarter * 32u + lane;
      uint4 o0, o1;
      o0.x = __float_as_uint(key[0]); o0.y = __float_as_uint(key[1]);
      o0.z = __float_as_uint(key[2]); o0.w = __float_as_uint(key[3]);
      o1.x = __float_as_uint(key[4]); o1.y = __float_as_uint(key[5]);
      o1.z = kKeySentinel; o1.w = kKeySentinel;
      uint4* dst = (uint4*)keys_out + (size_t)(pd.q_ofs + row) * (kKeyStride / 4);
      dst[0] = o0;
      dst[1] = o1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (ncta > 1) cluster_sync_all();  // no CTA may exit while its peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

static int ring_stages(int nkb) {
  // as deep as the 227 KB of shared memory allow next to the resident query tiles
  int stages = (int)((232448 - 2048 - (size_t)kQB * nkb * kBoxBytes) / kBoxBytes);
  if (stages > kMaxStages) stages = kMaxStages;
  const char* e = getenv("R3D_K1_STAGES");
  if (e && atoi(e) >= 2 && atoi(e) < stages) stages = atoi(e);
  return stages;
}

size_t l2_candidates_smem_bytes(int kp_cols) {
  const int nkb = (kp_cols + kKBlock - 1) / kKBlock;
  return 1024 + (size_t)(kQB * nkb + ring_stages(nkb)) * kBoxBytes + 8 * (2 * kMaxStages + 2 + 4) + 16;
}

int launch_l2_candidates(r3d_ctx* ctx, DeviceWorker& w, const PairDesc* d_pairs, const WorkItem* d_items,
                         uint32_t n_items, uint32_t* d_keys, int kp_cols, int ksteps, int cluster) {
  if (n_items == 0) return R3D_OK;
  const int nkb = (kp_cols + kKBlock - 1) / kKBlock;
  if (nkb > kMaxKBlocks) return fail(ctx, R3D_ERR_UNSUPPORTED, "descriptor dimension too large for the tensor-core path");
  const int stages = ring_stages(nkb);
  const size_t smem = l2_candidates_smem_bytes(kp_cols);
  R3D_CUDA_TRY(ctx, cudaFuncSetAttribute(k_l2_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  if (cluster != 2) cluster = 1;
  uint32_t grid = (uint32_t)w.sm_count / cluster * cluster;
  const uint32_t need = (n_items + cluster - 1) / cluster * cluster;
  if (need < grid) grid = need;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = w.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R3D_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, k_l2_candidates, (const CUtensorMap*)w.d_tmapQ, (const CUtensorMap*)w.d_tmapD,
                                       (const CUtensorMap*)w.d_tmapDh, d_pairs, d_items, n_items, d_keys, (uint32_t)nkb,
                                       (uint32_t)ksteps, (uint32_t)stages));
  return R3D_OK;
}

}  // namespace r3d
