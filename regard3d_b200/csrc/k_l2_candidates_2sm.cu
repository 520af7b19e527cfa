// k_l2_candidates_2sm.cu -- the candidate kernel on CTA PAIRS (tcgen05 cta_group::2).
//
// Why: a single-CTA kernel (round 1, removed; profiles/r01_k_l2_candidates_1sm_hunt.md) is bound by shared-memory operand reads: an
// M128 x N128 x K16 SS-MMA fetches 8 KB per 64 tensor cycles = the 128 B/clk port
// (profiles/r01_k_l2_candidates.md).  With cta_group::2 one instruction computes M = 256: each SM of
// the pair contributes its own 128 query rows (A) and HALF of the database tile (B, 128 of 256
// rows), and receives a 128 x 256 accumulator.  Per SM that is 8 KB of operands per 128 tensor
// cycles -- half the shared-memory traffic, half the TMA/L2 traffic per flop.
//
// Work item = (pair, 128-query block); a cluster owns two consecutive items of ONE pair.
// Per CTA: A = 1 query block (nkb boxes, resident), B ring = 128-row boxes of its half of every
// 256-row database tile, TMEM = 2 accumulator stages x 256 columns.
// Protocol (leader = CTA rank 0 issues every MMA):
//   full[s]    local TMA complete                      (each CTA, count 1 + tx)
//   pfull[s]   peer's full[s] forwarded to the leader  (leader, count 1; remote arrive by the peer's warp 1)
//   empty[s]   ring slot free in both CTAs             (multicast tcgen05.commit from the leader)
//   qfull / pqfull / qempty  same scheme for the resident query tiles
//   tfull[a]   accumulator stage ready in both CTAs    (multicast commit)
//   tempty[a]  both epilogues drained the stage        (leader, count 2 x 8 warps; the peer arrives remotely)
// Epilogue: warps 4-7 take accumulator columns 0-127, warps 8-11 columns 128-255 of the same query
// rows; the two partial key sets of a row are merged through shared memory at the end of an item.
#include "r3d_internal.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>

namespace r3d {

using namespace tcx;

namespace {

constexpr int kMaxStages2 = 12;
// Epilogue warps per CTA: 8 (two per TMEM lane quarter, 128 accumulator columns each) or 16 (four per quarter, 64
// columns each).  The epilogue's cost per tile does not shrink with the descriptor dimension while the MMA time does
// (5 K-steps at D = 64 against 10 at D = 144), and it is bound by the latency of its TMEM-load / min-tree chains, not
// by ALU throughput: short descriptors get twice the warps to hide it.
constexpr uint32_t kTmemCols2 = 512;
constexpr uint32_t kTileN = 256;  // database rows per tile (both halves)
// kind::f16, D = f32, A = B = f16 K-major, N = 256, M = 256 (cta_group::2)
constexpr uint32_t kInstrDesc2 = (1u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

}  // namespace

template <bool kVote, int kEpiWarps2, bool kDrain>
__global__ void __launch_bounds__(32 * (4 + kEpiWarps2), 1)
k_l2_candidates_2sm(const CUtensorMap* __restrict__ tmapQ, const CUtensorMap* __restrict__ tmapD,
                    const PairDesc* __restrict__ pairs, const WorkItem* __restrict__ items, uint32_t n_items,
                    uint32_t* __restrict__ keys_out, uint32_t nkb, uint32_t ksteps, uint32_t n_stages,
                    uint32_t n_qbuf) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // n_qbuf (1 or 2) x nkb boxes: this CTA's 128 query rows.  With two buffers the next work item's query block is
  // loaded while the current item's last tiles are still in the tensor pipe (no bubble at item boundaries).
  const uint32_t q_base = base;
  const uint32_t q_bytes = nkb * kBoxBytes;
  const uint32_t d_base = q_base + n_qbuf * q_bytes;              // n_stages boxes
  const uint32_t bar_base = d_base + n_stages * kBoxBytes;
  const uint32_t bar_full = bar_base;                             // [kMaxStages2]
  const uint32_t bar_pfull = bar_full + 8 * kMaxStages2;          // [kMaxStages2] (leader)
  const uint32_t bar_empty = bar_pfull + 8 * kMaxStages2;         // [kMaxStages2]
  const uint32_t bar_qfull = bar_empty + 8 * kMaxStages2;         // [2]
  const uint32_t bar_pqfull = bar_qfull + 16;                     // [2] (leader)
  const uint32_t bar_qempty = bar_pqfull + 16;                    // [2]
  const uint32_t bar_tfull = bar_qempty + 16;                     // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                     // [2] (leader)
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t key_xchg = (tmem_slot + 16 + 15u) & ~15u;        // (column groups - 1) x 128 rows x 8 u32: partial key sets
  unsigned char* gen_base = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(gen_base + (tmem_slot - base));
  uint32_t* xchg = (uint32_t*)(gen_base + (key_xchg - base));

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const uint32_t cluster_id = blockIdx.x >> 1;
  const uint32_t n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < (uint32_t)kMaxStages2; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (uint32_t qb = 0; qb < 2; ++qb) {
      mbar_init(bar_qfull + 8 * qb, 1);
      mbar_init(bar_pqfull + 8 * qb, 1);
      mbar_init(bar_qempty + 8 * qb, 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2 * kEpiWarps2);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) =====================================
    uint32_t stage = 0, phase = 0, qi = 0;
    for (uint32_t it = cluster_id * 2 + crank; it < n_items; it += n_clusters * 2, ++qi) {
      const WorkItem wi = items[it];
      const PairDesc pd = pairs[wi.pair];
      const CUtensorMap* mq = tmapQ + pd.slotJ;
      const CUtensorMap* md = tmapD + pd.slotI;
      const uint32_t nboxes = (pd.nI_pad / kTileN) * nkb;
      const uint32_t ahead = nboxes < n_stages ? nboxes : n_stages;
      // query buffer of this item and how often it has been used before (barrier parity)
      const uint32_t qb = n_qbuf == 2 ? (qi & 1u) : 0u;
      const uint32_t quse = n_qbuf == 2 ? (qi >> 1) : qi;
      // one buffer: run the database ring ahead first, the buffer is released by the previous item's last MMA;
      // two buffers: the buffer was released an item ago, load it before anything else
      const uint32_t q_at = n_qbuf == 2 ? 0u : ahead;
      uint32_t b = 0, t = 0, kb = 0;
      for (;;) {
        if (b == q_at) {
          mbar_wait(bar_qempty + 8 * qb, (quse & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_qfull + 8 * qb, nkb * kBoxBytes);
            for (uint32_t k2 = 0; k2 < nkb; ++k2)
              tma_load_2d(q_base + qb * q_bytes + k2 * kBoxBytes, mq, (int)(k2 * kKBlock), (int)(wi.sb * kTileRows),
                          bar_qfull + 8 * qb);
          }
          __syncwarp();
        }
        if (b == nboxes) break;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, kBoxBytes);
          // this CTA's half of the 256-row database tile
          tma_load_2d(d_base + stage * kBoxBytes, md, (int)(kb * kKBlock), (int)(t * kTileN + crank * kTileRows),
                      bar_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        ++b;
        if (++kb == nkb) { kb = 0; ++t; }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, acc = 0, accphase = 0, qi = 0;
    if (leader) {
      // ============================ MMA issuer (leader CTA, cta_group::2) ============================
      for (uint32_t it = cluster_id * 2; it < n_items; it += n_clusters * 2, ++qi) {
        const WorkItem wi = items[it];
        const PairDesc pd = pairs[wi.pair];
        const uint32_t ntiles = pd.nI_pad / kTileN;
        const uint32_t qb = n_qbuf == 2 ? (qi & 1u) : 0u;
        const uint32_t qf = (n_qbuf == 2 ? (qi >> 1) : qi) & 1u;
        mbar_wait(bar_qfull + 8 * qb, qf);
        mbar_wait(bar_pqfull + 8 * qb, qf);
        tc_fence_after();
        for (uint32_t t = 0; t < ntiles; ++t) {
          mbar_wait(bar_tempty + 8 * acc, accphase ^ 1u);  // both epilogues drained this accumulator stage
          tc_fence_after();
          const uint32_t d0 = tmem_base + acc * kTileN;
          uint32_t ks_left = ksteps;
          for (uint32_t kb = 0; kb < nkb; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            mbar_wait(bar_pfull + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b_lo = desc_lo(d_base + stage * kBoxBytes);
              const uint32_t a_lo = desc_lo(q_base + qb * q_bytes + kb * kBoxBytes);
              const uint32_t ks_here = ks_left < 4u ? ks_left : 4u;
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) {
                if (k < ks_here)
                  tc2_mma_f16(d0, make_desc(a_lo + 2 * k), make_desc(b_lo + 2 * k), kInstrDesc2, (kb | k) != 0u ? 1u : 0u);
              }
              tc2_commit_mc(bar_empty + 8 * stage, 3);
              if (kb + 1 == nkb) tc2_commit_mc(bar_tfull + 8 * acc, 3);
            }
            __syncwarp();
            ks_left -= 4u;
            if (++stage == n_stages) { stage = 0; phase ^= 1u; }
          }
          acc ^= 1u;
          if (acc == 0) accphase ^= 1u;
        }
        if (elect_one()) tc2_commit_mc(bar_qempty + 8 * qb, 3);
        __syncwarp();
      }
    } else {
      // ============ peer CTA: forward "my operands have landed" to the leader's barriers ============
      const uint32_t r_pqfull = mapa_shared(bar_pqfull, 0);
      for (uint32_t it = cluster_id * 2 + 1; it < n_items; it += n_clusters * 2, ++qi) {
        const WorkItem wi = items[it];
        const PairDesc pd = pairs[wi.pair];
        const uint32_t nboxes = (pd.nI_pad / kTileN) * nkb;
        const uint32_t qb = n_qbuf == 2 ? (qi & 1u) : 0u;
        const uint32_t qf = (n_qbuf == 2 ? (qi >> 1) : qi) & 1u;
        mbar_wait(bar_qfull + 8 * qb, qf);
        if (elect_one()) mbar_arrive_remote(r_pqfull + 8 * qb);
        __syncwarp();
        for (uint32_t b = 0; b < nboxes; ++b) {
          mbar_wait(bar_full + 8 * stage, phase);
          if (elect_one()) mbar_arrive_remote(mapa_shared(bar_pfull + 8 * stage, 0));
          __syncwarp();
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ======================================= epilogue (both CTAs) =======================================
    constexpr uint32_t kColGroups = kEpiWarps2 / 4;          // warps that share a TMEM lane quarter
    constexpr uint32_t kColsPerWarp = kTileN / kColGroups;   // accumulator columns of one warp: 128 or 64
    constexpr uint32_t kLoads = kColsPerWarp / 32;
    const uint32_t ew = warp - 4u;
    const uint32_t half = ew >> 2;            // column group: accumulator columns [half * kColsPerWarp, +kColsPerWarp)
    const uint32_t lane_quarter = warp & 3u;
    const uint32_t r_tempty0 = mapa_shared(bar_tempty, 0);
    uint32_t acc = 0, accphase = 0;
    for (uint32_t it = cluster_id * 2 + crank; it < n_items; it += n_clusters * 2) {
      const WorkItem wi = items[it];
      const PairDesc pd = pairs[wi.pair];
      const uint32_t ntiles = pd.nI_pad / kTileN;
      float key[kNumKeys];
#pragma unroll
      for (int i = 0; i < kNumKeys; ++i) key[i] = __uint_as_float(kKeySentinel);
      const uint32_t keep_mask = ~((1u << pd.chunk_bits) - 1u);
      for (uint32_t t = 0; t < ntiles; ++t) {
        mbar_wait(bar_tfull + 8 * acc, accphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((lane_quarter * 32u) << 16) + acc * kTileN + half * kColsPerWarp;
        // accumulator column j of the 256-wide tile is database row t*256 + j (rows 0-127 from the leader's
        // half, 128-255 from the peer's)
        const uint32_t chunk0 = (t * kTileN + half * kColsPerWarp) / kChunk;
        constexpr uint32_t kCpl = 32 / kChunk;
        if constexpr (kDrain) {
          // drain first: all of this warp's columns go to registers, the accumulator stage is handed back to the MMA
          // issuer after ONE TMEM round trip, and the min / insertion arithmetic overlaps the next tiles' MMAs
          uint32_t v[kLoads][32];
#pragma unroll
          for (uint32_t sblk = 0; sblk < kLoads; ++sblk) tc_ld32(taddr + 32 * sblk, v[sblk]);
#pragma unroll
          for (uint32_t sblk = 0; sblk < kLoads; ++sblk) tc_wait_ld(v[sblk]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(r_tempty0 + 8 * acc);  // the LEADER's tempty collects both CTAs
#pragma unroll
          for (uint32_t sblk = 0; sblk < kLoads; ++sblk)
#pragma unroll
            for (uint32_t c = 0; c < kCpl; ++c) chunk_update<kVote>(v[sblk] + c * kChunk, chunk0 + sblk * kCpl + c, keep_mask, key);
        } else {
          uint32_t v[2][32];
          tc_ld32(taddr, v[0]);
#pragma unroll
          for (uint32_t sblk = 0; sblk < kLoads; ++sblk) {
            tc_wait_ld(v[sblk & 1u]);
            if (sblk + 1 < kLoads) tc_ld32(taddr + 32 * (sblk + 1), v[(sblk + 1) & 1u]);
            if (sblk + 1 == kLoads) {  // the stage is drained: release it before the last block's arithmetic
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_remote(r_tempty0 + 8 * acc);  // the LEADER's tempty collects both CTAs
            }
#pragma unroll
            for (uint32_t c = 0; c < kCpl; ++c) chunk_update<kVote>(v[sblk & 1u] + c * kChunk, chunk0 + sblk * kCpl + c, keep_mask, key);
          }
        }
        acc ^= 1u;
        if (acc == 0) accphase ^= 1u;
      }
      // merge the column groups of every query row: groups 1.. hand their keys over in shared memory
      const uint32_t r = lane_quarter * 32u + lane;  // row inside the 128-query block
      if (half != 0) {
#pragma unroll
        for (int i = 0; i < kNumKeys; ++i) xchg[((half - 1u) * 128u + r) * 8 + i] = __float_as_uint(key[i]);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps2 * 32) : "memory");
      if (half == 0) {
#pragma unroll
        for (uint32_t g = 0; g + 1 < kColGroups; ++g) {
#pragma unroll
          for (int i = 0; i < kNumKeys; ++i) {
            float x = __uint_as_float(xchg[(g * 128u + r) * 8 + i]);
#pragma unroll
            for (int j = 0; j < kNumKeys - 1; ++j) {
              const float hi = fmaxf(key[j], x);
              key[j] = fminf(key[j], x);
              x = hi;
            }
            key[kNumKeys - 1] = fminf(key[kNumKeys - 1], x);
          }
        }
        const uint32_t row = wi.sb * kTileRows + r;
        uint4 o0, o1;
        o0.x = __float_as_uint(key[0]); o0.y = __float_as_uint(key[1]);
        o0.z = __float_as_uint(key[2]); o0.w = __float_as_uint(key[3]);
        o1.x = __float_as_uint(key[4]); o1.y = __float_as_uint(key[5]);
        o1.z = kKeySentinel; o1.w = kKeySentinel;
        uint4* dst = (uint4*)keys_out + (size_t)(pd.q_ofs + row) * (kKeyStride / 4);
        dst[0] = o0;
        dst[1] = o1;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps2 * 32) : "memory");  // xchg may be reused
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
  }
}

static int ring_stages2(int nkb, int n_qbuf) {
  int stages = (int)((232448 - 8192 - (size_t)n_qbuf * nkb * kBoxBytes) / kBoxBytes);
  if (stages > kMaxStages2) stages = kMaxStages2;
  const char* e = getenv("R3D_K1_STAGES");
  if (e && atoi(e) >= 2 && atoi(e) < stages) stages = atoi(e);
  return stages;
}

int launch_l2_candidates_2sm(r3d_ctx* ctx, DeviceWorker& w, const PairDesc* d_pairs, const WorkItem* d_items,
                             uint32_t n_items, uint32_t* d_keys, int kp_cols, int ksteps) {
  if (n_items == 0) return R3D_OK;
  if (n_items & 1u) return fail(ctx, R3D_ERR_INVALID, "2-SM candidate kernel needs an even number of work items");
  const int nkb = (kp_cols + kKBlock - 1) / kKBlock;
  if (nkb > kMaxKBlocks) return fail(ctx, R3D_ERR_UNSUPPORTED, "descriptor dimension too large for the tensor-core path");
  static const int want_qbuf = []() {  // R3D_K1_QBUF=1: single query buffer (A/B switch)
    const char* e = getenv("R3D_K1_QBUF");
    return (e && atoi(e) == 1) ? 1 : 2;
  }();
  int n_qbuf = want_qbuf;
  if (ring_stages2(nkb, n_qbuf) < 4) n_qbuf = 1;  // very wide descriptors: keep the ring deep enough instead
  // Epilogue shape (A/B in profiles/r02_chunk_ab.md): 8 warps everywhere -- 16 change nothing at D = 64 and cost 5 % at
  // D >= 128; R3D_K1_EPI = 8 | 16 forces one
  static const int force_epi = []() {
    const char* e = getenv("R3D_K1_EPI");
    return e ? atoi(e) : 0;
  }();
  const int epi = force_epi == 16 ? 16 : 8;
  const int stages = ring_stages2(nkb, n_qbuf);
  const size_t smem = 1024 + (size_t)(n_qbuf * nkb + stages) * kBoxBytes + 8 * (3 * kMaxStages2 + 6 + 4) + 32 + 16 +
                      (size_t)(epi / 4 - 1) * 128 * 8 * 4;
  static const bool vote = []() {  // R3D_K1_VOTE=0: always run the insertion network (A/B switch)
    const char* e = getenv("R3D_K1_VOTE");
    return !(e && atoi(e) == 0);
  }();
  // R3D_K1_DRAIN = 0 | 1: pipelined TMEM loads (two 32-column blocks in flight) or drain-first.  Drain-first gains
  // 1-2 % where the tile is MMA-bound (>= 7 K-steps) and loses 3 % at D = 64
  static const int force_drain = []() {
    const char* e = getenv("R3D_K1_DRAIN");
    return e ? atoi(e) : -1;
  }();
  const int drain = force_drain >= 0 ? force_drain : (ksteps > 6 ? 1 : 0);
  auto kernel = drain ? (epi == 16 ? (vote ? k_l2_candidates_2sm<true, 16, true> : k_l2_candidates_2sm<false, 16, true>)
                                   : (vote ? k_l2_candidates_2sm<true, 8, true> : k_l2_candidates_2sm<false, 8, true>))
                      : (epi == 16 ? (vote ? k_l2_candidates_2sm<true, 16, false> : k_l2_candidates_2sm<false, 16, false>)
                                   : (vote ? k_l2_candidates_2sm<true, 8, false> : k_l2_candidates_2sm<false, 8, false>));
  R3D_CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  uint32_t grid = (uint32_t)w.sm_count / 2 * 2;
  if (n_items < grid) grid = n_items;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(32 * (4 + epi));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = w.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R3D_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kernel, (const CUtensorMap*)w.d_tmapQ, (const CUtensorMap*)w.d_tmapD,
                                       d_pairs, d_items, n_items, d_keys, (uint32_t)nkb, (uint32_t)ksteps, (uint32_t)stages,
                                       (uint32_t)n_qbuf));
  return R3D_OK;
}

}  // namespace r3d
