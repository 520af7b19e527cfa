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
#include "tc_ptx.cuh"

#include <cstdlib>

namespace r3d {

using namespace tcx;

namespace {

constexpr int kMaxStages = 12;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (4 + kEpiWarps);
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCols = 128;                        // fp32 columns per (stage, query block)

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
