// tc_ptx.cuh -- inline-PTX wrappers for the sm_100a tensor-core path (mbarrier, TMA, tcgen05, clusters)
// and the candidate-key epilogue shared by the k_l2_candidates kernels.
#pragma once
#include "r3d_internal.cuh"

namespace r3d {
namespace tcx {

constexpr uint32_t kBoxBytes = kTileRows * kKBlock * 2;  // one 128-row x 64-column fp16 TMA box: 16384
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
template <bool kVote = false>
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
  // A key that is not below the current largest kept key leaves the set unchanged (the network would carry it
  // through every level).  After t chunks a lane inserts with probability ~ kNumKeys / t, so most chunks need no
  // insertion in ANY lane of the warp: one vote skips the 11-instruction network (warp-uniform branch).
  if (kVote && !__any_sync(0xffffffffu, x < key[kNumKeys - 1])) return;
#pragma unroll
  for (int i = 0; i < kNumKeys - 1; ++i) {  // sorted insertion network: 2 FMNMX per level
    const float hi = fmaxf(key[i], x);
    key[i] = fminf(key[i], x);
    x = hi;
  }
  key[kNumKeys - 1] = fminf(key[kNumKeys - 1], x);
}


// ---- cluster-remote mbarrier arrive (CTA pair protocols) ----------------------------------------
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
// RELAXED on purpose: a release at cluster scope compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR per
// arrive (measured: +65 % kernel time).  The hand-shakes that use it publish no generic-proxy writes:
// "TMEM stage drained" follows tcgen05.wait::ld + tcgen05.fence::before_thread_sync, and "operands
// landed" forwards an mbarrier completion of TMA (async-proxy) writes that the tensor core reads.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// ---- cta_group::2 variants -----------------------------------------------------------------------
__device__ __forceinline__ void tc2_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

}  // namespace tcx
}  // namespace r3d
