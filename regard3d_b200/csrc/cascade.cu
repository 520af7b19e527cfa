// cascade.cu -- OpenMVG's CASCADE_HASHING_L2 matcher (BASELINE config 4 names it) on the GPU.
//
// Replaces: openMVG::matching_image_collection::Cascade_Hashing_Matcher_Regions::Match and matching::CascadeHasher
// (OpenMVG 1.4, un-vendored; the pair loop has the shape of /root/reference/src/R3DComputeMatches.cpp:437-488 and the
// same tail -- ratio^2 test, IndMatch(i in I, j in J), getDeduplicated, coordinate de-duplication -- as the
// brute-force path).  Algorithm restated in SURVEY.md A.8 / oracle/oracle_cascade.cpp:
//   Init(dim)     hash length = descriptor dimension: dim primary + 6 x 10 secondary N(0,1) projections drawn on the
//                 HOST with <random> (std::mt19937(default_seed), one std::normal_distribution<>) -- the library the
//                 reference links, so the table is the reference's table
//   zero mean     mean over the used views (ascending id) of each view's mean descriptor: float running sums in row
//                 order, one thread per dimension (k_cascade_view_mean / k_cascade_zero_mean) -- order preserved
//   hash          bit j = (P (d - mean))_j > 0, one float accumulator per projection over k = 0..dim-1, product and sum
//                 rounded separately (k_cascade_hash); bucket id = the group's 10 sign bits, first = MSB
//   buckets       per (view, group): ids in ascending order per bucket = the reference's push_back order
//                 (k_cascade_buckets: counting sort + per-bucket ordering)
//   match         one warp per query of J against the buckets of I (k_cascade_match): first occurrences through a
//                 per-warp bitmap in shared memory, Hamming histogram = the reference's counting sort, cut at the 10th
//                 candidate in (Hamming, arrival) order, exact L2 in the upstream accumulation order, top-2 by
//                 (distance, id) = std::partial_sort of pairs, then the common emit_result()
// HBM/L2-bound integer work: per query ~60-120 candidate ids (4 B) + hash codes (dim/8 B) and 10 descriptors.
#include "match_device.cuh"
#include "r3d_cascade.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <string>
#include <vector>

namespace r3d {

namespace {

constexpr int kGroups = 6, kBitsPerBucket = 10, kBuckets = 1 << kBitsPerBucket, kTop = 10;
constexpr int kMatchWarps = 8;
constexpr uint32_t kMaxCascadeRows = 65536;  // per-warp "seen" bitmap: 8 KB of shared memory

struct MeanJob { const void* desc; uint32_t n; uint32_t pad_; };

template <int DTYPE>
__device__ __forceinline__ float desc_elem(const void* desc, size_t idx) {
  return DTYPE == 0 ? ((const float*)desc)[idx] : (float)((const uint8_t*)desc)[idx];
}

// CascadeHasher::GetZeroMeanDescriptor of one view: thread j = dimension j, rows in order
template <int DTYPE>
__global__ void k_cascade_view_mean(const MeanJob* __restrict__ jobs, uint32_t dim, float* __restrict__ means) {
  const MeanJob jb = jobs[blockIdx.x];
  const uint32_t j = threadIdx.x;
  if (j >= dim) return;
  float acc = 0.f;
  for (uint32_t i = 0; i < jb.n; ++i) acc = __fadd_rn(acc, desc_elem<DTYPE>(jb.desc, (size_t)i * dim + j));
  means[(size_t)blockIdx.x * dim + j] = jb.n ? __fdiv_rn(acc, (float)jb.n) : 0.f;
}

__global__ void k_cascade_zero_mean(const float* __restrict__ means, uint32_t n_views, uint32_t dim, float* __restrict__ out) {
  const uint32_t j = threadIdx.x;
  if (j >= dim) return;
  float acc = 0.f;
  for (uint32_t v = 0; v < n_views; ++v) acc = __fadd_rn(acc, means[(size_t)v * dim + j]);
  out[j] = n_views ? __fdiv_rn(acc, (float)n_views) : 0.f;
}

// Hash codes + bucket ids of kHashRows descriptors per CTA, all views of the job in ONE launch: CTA b belongs to the view
// v with first_block[v] <= b < first_block[v + 1].  projT: [dim][np] (np = dim + 60), thread p owns projection p (and
// p + blockDim.x).
constexpr int kHashRows = 8;
struct HashJob { const void* desc; uint32_t* code; uint16_t* bucket; uint32_t n; uint32_t first_block; };
template <int DTYPE>
__global__ void __launch_bounds__(256) k_cascade_hash(const HashJob* __restrict__ jobs, uint32_t n_jobs, uint32_t dim,
                                                      const float* __restrict__ projT, const float* __restrict__ zero_mean,
                                                      uint32_t words) {
  extern __shared__ float s_d[];                 // kHashRows x dim centred descriptors
  __shared__ uint32_t s_code[kHashRows][8];
  __shared__ uint32_t s_bucket[kHashRows][kGroups];
  uint32_t lo = 0, hi = n_jobs;                  // last job whose first_block <= blockIdx.x
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (jobs[mid].first_block <= blockIdx.x) lo = mid; else hi = mid;
  }
  const HashJob jb = jobs[lo];
  const void* desc = jb.desc;
  const uint32_t n = jb.n;
  uint32_t* code = jb.code;
  uint16_t* bucket = jb.bucket;
  const uint32_t row0 = (blockIdx.x - jb.first_block) * kHashRows;
  const uint32_t np = dim + kGroups * kBitsPerBucket;
  for (uint32_t t = threadIdx.x; t < kHashRows * dim; t += blockDim.x) {
    const uint32_t r = t / dim, k = t % dim;
    s_d[t] = row0 + r < n ? __fsub_rn(desc_elem<DTYPE>(desc, (size_t)(row0 + r) * dim + k), zero_mean[k]) : 0.f;
  }
  for (uint32_t t = threadIdx.x; t < kHashRows * 8; t += blockDim.x) s_code[t / 8][t % 8] = 0u;
  for (uint32_t t = threadIdx.x; t < kHashRows * kGroups; t += blockDim.x) s_bucket[t / kGroups][t % kGroups] = 0u;
  __syncthreads();
  for (uint32_t p = threadIdx.x; p < np; p += blockDim.x) {
    float acc[kHashRows];
#pragma unroll
    for (int r = 0; r < kHashRows; ++r) acc[r] = 0.f;
    for (uint32_t k = 0; k < dim; ++k) {
      const float pv = __ldg(projT + (size_t)k * np + p);
#pragma unroll
      for (int r = 0; r < kHashRows; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(pv, s_d[r * dim + k]));
    }
#pragma unroll
    for (int r = 0; r < kHashRows; ++r) {
      if (!(acc[r] > 0.f)) continue;
      if (p < dim) {
        atomicOr(&s_code[r][p >> 5], 1u << (p & 31u));
      } else {
        const uint32_t g = (p - dim) / kBitsPerBucket, k = (p - dim) % kBitsPerBucket;
        atomicOr(&s_bucket[r][g], 1u << (kBitsPerBucket - 1 - k));
      }
    }
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < kHashRows * words; t += blockDim.x) {
    const uint32_t r = t / words, wd = t % words;
    if (row0 + r < n) code[(size_t)(row0 + r) * words + wd] = s_code[r][wd];
  }
  for (uint32_t t = threadIdx.x; t < kHashRows * kGroups; t += blockDim.x) {
    const uint32_t r = t / kGroups, g = t % kGroups;
    if (row0 + r < n) bucket[(size_t)(row0 + r) * kGroups + g] = (uint16_t)s_bucket[r][g];
  }
}

// Buckets of one (view, group): offsets[1025] + ids in ascending order inside every bucket.
__global__ void __launch_bounds__(256) k_cascade_buckets(const CascadeView* __restrict__ views) {
  __shared__ uint32_t s_cnt[kBuckets];
  __shared__ uint32_t s_ofs[kBuckets + 1];
  __shared__ uint32_t s_warp[8];
  const CascadeView cv = views[blockIdx.y];
  const uint32_t g = blockIdx.x, n = cv.n;
  uint32_t* ofs = cv.bk_ofs + (size_t)g * (kBuckets + 1);
  uint32_t* ids = cv.bk_ids + (size_t)g * n;
  for (uint32_t b = threadIdx.x; b < kBuckets; b += blockDim.x) s_cnt[b] = 0u;
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) atomicAdd(&s_cnt[cv.bucket[(size_t)j * kGroups + g]], 1u);
  __syncthreads();
  {  // exclusive scan of 1024 counters: 4 per thread
    const uint32_t b0 = threadIdx.x * 4;
    const uint32_t c0 = s_cnt[b0], c1 = s_cnt[b0 + 1], c2 = s_cnt[b0 + 2], c3 = s_cnt[b0 + 3];
    uint32_t tot = c0 + c1 + c2 + c3, incl = tot;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)(threadIdx.x & 31u) >= o) incl += u;
    }
    if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (uint32_t wv = 0; wv < (threadIdx.x >> 5); ++wv) base += s_warp[wv];
    const uint32_t e = base + incl - tot;
    s_ofs[b0] = e; s_ofs[b0 + 1] = e + c0; s_ofs[b0 + 2] = e + c0 + c1; s_ofs[b0 + 3] = e + c0 + c1 + c2;
    if (threadIdx.x == blockDim.x - 1) s_ofs[kBuckets] = e + tot;
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b <= kBuckets; b += blockDim.x) ofs[b] = s_ofs[b];
  for (uint32_t b = threadIdx.x; b < kBuckets; b += blockDim.x) s_cnt[b] = 0u;  // now: fill cursors
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
    const uint32_t b = cv.bucket[(size_t)j * kGroups + g];
    ids[s_ofs[b] + atomicAdd(&s_cnt[b], 1u)] = j;
  }
  __syncthreads();
  // order inside the buckets: small ones by insertion (one thread), large ones by an ordered re-scan (one warp)
  for (uint32_t b = threadIdx.x; b < kBuckets; b += blockDim.x) {
    const uint32_t lo = s_ofs[b], hi = s_ofs[b + 1];
    if (hi - lo > 32u) continue;
    for (uint32_t a = lo + 1; a < hi; ++a) {
      const uint32_t v = ids[a];
      uint32_t c = a;
      while (c > lo && ids[c - 1] > v) { ids[c] = ids[c - 1]; --c; }
      ids[c] = v;
    }
  }
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t b = warp; b < kBuckets; b += blockDim.x / 32) {
    const uint32_t lo = s_ofs[b], hi = s_ofs[b + 1];
    if (hi - lo <= 32u) continue;
    uint32_t at = lo;
    for (uint32_t base = 0; base < n; base += 32) {
      const uint32_t j = base + lane;
      const bool hit = j < n && cv.bucket[(size_t)j * kGroups + g] == b;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) ids[at + __popc(m & ((1u << lane) - 1u))] = j;
      at += __popc(m);
    }
  }
}

// One warp per query (a descriptor of view J) against the hashed view I.  Dynamic shared memory per warp: the "seen"
// bitmap over I's descriptors (kept all-zero between queries), the Hamming histogram, the list of first occurrences
// (Hamming << 16 | id, arrival order) that spares pass 2 its gathers, the selected ids and their distances.
constexpr uint32_t kListCap = 256;
template <int DTYPE>
__global__ void __launch_bounds__(kMatchWarps * 32, 4) k_cascade_match(const PairDesc* __restrict__ pairs, const uint2* __restrict__ cidx,
                                                                       const CascadeView* __restrict__ views, uint32_t dim, float ratio2,
                                                                       uint32_t bitmap_words, uint32_t* __restrict__ counters,
                                                                       uint2* __restrict__ matches) {
  extern __shared__ uint32_t s_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t hist_words = (dim + 1 + 31u) & ~31u;
  uint32_t* bitmap = s_raw + (size_t)warp * (bitmap_words + hist_words + kListCap + 64);
  uint32_t* hist = bitmap + bitmap_words;
  uint32_t* list = hist + hist_words;
  uint32_t* sel = list + kListCap;       // ids of the (<= 10) selected candidates
  float* seld = reinterpret_cast<float*>(sel + 32);
  for (uint32_t t = lane; t < bitmap_words; t += 32) bitmap[t] = 0u;
  __syncwarp();
  const uint32_t pair = blockIdx.y;
  const PairDesc pd = pairs[pair];
  const uint2 ci = cidx[pair];
  const CascadeView vI = views[ci.x], vJ = views[ci.y];
  const uint32_t words = vI.words;
  const size_t rb = row_bytes(DTYPE, dim);
  // uint8 rows of 16 * 2^k bytes: `parts` lanes share one exact distance (integer sum, order-free: exact_l2's own rule)
  const uint32_t parts = (DTYPE == 1 && (dim == 16 || dim == 32 || dim == 64 || dim == 128 || dim == 256)) ? dim / 16 : 0;
  uint32_t stat_raw = 0, stat_distinct = 0;  // candidates walked / distinct candidates, over this warp's queries
  for (uint32_t q = blockIdx.x * kMatchWarps + warp; q < pd.nJ; q += gridDim.x * kMatchWarps) {
    // the query's six buckets in I
    uint32_t beg[kGroups], len[kGroups], total = 0;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      const uint32_t bid = vJ.bucket[(size_t)q * kGroups + g];
      const uint32_t* o = vI.bk_ofs + (size_t)g * (kBuckets + 1) + bid;
      beg[g] = o[0];
      len[g] = o[1] - o[0];
      total += len[g];
    }
    if (total <= 2u) continue;  // "not at least NN candidates" (raw count)
    uint32_t qc[8];
    if (words == 4) {
      const uint4 c = *reinterpret_cast<const uint4*>(vJ.code + (size_t)q * 4);
      qc[0] = c.x; qc[1] = c.y; qc[2] = c.z; qc[3] = c.w;
      qc[4] = qc[5] = qc[6] = qc[7] = 0u;
    } else {
#pragma unroll
      for (int wd = 0; wd < 8; ++wd) qc[wd] = (uint32_t)wd < words ? vJ.code[(size_t)q * words + wd] : 0u;
    }
    for (uint32_t t = lane; t < hist_words; t += 32) hist[t] = 0u;
    __syncwarp();
    // candidate s of the concatenated bucket lists -> id
    auto candidate = [&](uint32_t s) -> uint32_t {
      uint32_t r = s;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        if (r < len[g]) return vI.bk_ids[(size_t)g * vI.n + beg[g] + r];
        r -= len[g];
      }
      return 0u;
    };
    auto hamming = [&](uint32_t id) -> uint32_t {
      if (words == 4) {
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(vI.code + (size_t)id * 4));
        return __popc(qc[0] ^ c.x) + __popc(qc[1] ^ c.y) + __popc(qc[2] ^ c.z) + __popc(qc[3] ^ c.w);
      }
      uint32_t h = 0;
#pragma unroll
      for (int wd = 0; wd < 8; ++wd)
        if ((uint32_t)wd < words) h += __popc(qc[wd] ^ vI.code[(size_t)id * words + wd]);
      return h;
    };
    // ---- pass 1: first occurrences -> Hamming histogram (+ the list, when it fits) ----
    const bool use_list = total <= kListCap;
    uint32_t n_list = 0;
    for (uint32_t base = 0; base < total; base += 32) {
      const uint32_t s = base + lane;
      const bool act = s < total;
      const uint32_t id = act ? candidate(s) : 0xffffffffu;
      const unsigned same = __match_any_sync(0xffffffffu, id);
      bool fresh = false;
      if (act && (uint32_t)(__ffs(same) - 1) == lane) {
        const uint32_t bit = 1u << (id & 31u);
        fresh = (atomicOr(&bitmap[id >> 5], bit) & bit) == 0u;
      }
      const uint32_t h = fresh ? hamming(id) : 0u;
      if (fresh) atomicAdd(&hist[h], 1u);
      if (use_list) {
        const unsigned fm = __ballot_sync(0xffffffffu, fresh);
        if (fresh) list[n_list + __popc(fm & lt)] = (h << 16) | id;
        n_list += __popc(fm);
      }
    }
    __syncwarp();
    // ---- cut: the Hamming distance h* of the 10th distinct candidate ----
    uint32_t distinct = 0, hstar = 0, below = 0;
    {
      const uint32_t per = hist_words / 32;  // bins per lane
      uint32_t mine = 0;
      for (uint32_t t = 0; t < per; ++t) mine += hist[lane * per + t];
      uint32_t incl = mine;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += u;
      }
      distinct = __shfl_sync(0xffffffffu, incl, 31);
      const uint32_t target = distinct < (uint32_t)kTop ? distinct : (uint32_t)kTop;
      const unsigned reach = __ballot_sync(0xffffffffu, incl >= target && target > 0u);
      const int owner = reach ? __ffs(reach) - 1 : 0;
      uint32_t h_l = 0, b_l = 0;
      if ((int)lane == owner) {
        uint32_t run = incl - mine;
        for (uint32_t t = 0; t < per; ++t) {
          const uint32_t c = hist[lane * per + t];
          if (run + c >= target) { h_l = lane * per + t; b_l = run; break; }
          run += c;
        }
      }
      hstar = __shfl_sync(0xffffffffu, h_l, owner);
      below = __shfl_sync(0xffffffffu, b_l, owner);
    }
    const uint32_t target = distinct < (uint32_t)kTop ? distinct : (uint32_t)kTop;
    const uint32_t need_eq = target - below;
    stat_raw += total;
    stat_distinct += distinct;
    // ---- pass 2: every first occurrence again, in arrival order: clear its bit (the bitmap ends all-zero), take it
    //      when it lies below the cut or among the first arrivals on it ----
    uint32_t n_sel = 0, eq_seen = 0;
    const uint32_t walk = use_list ? n_list : total;
    for (uint32_t base = 0; base < walk; base += 32) {
      const uint32_t s = base + lane;
      const bool act = s < walk;
      uint32_t id = 0xffffffffu, h = 0xffffffffu;
      bool fresh = false;
      if (use_list) {
        if (act) {
          const uint32_t e = list[s];
          id = e & 0xffffu;
          h = e >> 16;
          fresh = true;
          atomicAnd(&bitmap[id >> 5], ~(1u << (id & 31u)));
        }
      } else {
        id = act ? candidate(s) : 0xffffffffu;
        const unsigned same = __match_any_sync(0xffffffffu, id);
        if (act && (uint32_t)(__ffs(same) - 1) == lane) {
          const uint32_t bit = 1u << (id & 31u);
          fresh = (atomicAnd(&bitmap[id >> 5], ~bit) & bit) != 0u;
        }
        h = fresh ? hamming(id) : 0xffffffffu;
      }
      const bool is_eq = fresh && h == hstar;
      const unsigned eqm = __ballot_sync(0xffffffffu, is_eq);
      const bool take = fresh && distinct >= 2u && (h < hstar || (is_eq && eq_seen + __popc(eqm & lt) < need_eq));
      eq_seen += __popc(eqm);
      const unsigned tm = __ballot_sync(0xffffffffu, take);
      if (take) sel[n_sel + __popc(tm & lt)] = id;
      n_sel += __popc(tm);
    }
    __syncwarp();
    if (n_sel < 2u) continue;
    // ---- exact distances of the selected candidates (upstream accumulation order), top-2 by (d, id) ----
    float d = FLT_MAX;
    uint32_t id = 0xffffffffu;
    const unsigned char* qrow = (const unsigned char*)pd.descJ + (size_t)q * rb;
    if (parts) {
      const uint32_t piece = lane % parts, per_pass = 32 / parts;
      const uint4 qv = *reinterpret_cast<const uint4*>(qrow + piece * 16);
      for (uint32_t c0 = 0; c0 < n_sel; c0 += per_pass) {
        const uint32_t c = c0 + lane / parts;
        uint32_t isum = 0u;
        if (c < n_sel) {
          const uint4 dv = __ldg(reinterpret_cast<const uint4*>((const unsigned char*)pd.descI + (size_t)sel[c] * rb + piece * 16));
          uint32_t ad = __vabsdiffu4(qv.x, dv.x); isum = __dp4a(ad, ad, isum);
          ad = __vabsdiffu4(qv.y, dv.y); isum = __dp4a(ad, ad, isum);
          ad = __vabsdiffu4(qv.z, dv.z); isum = __dp4a(ad, ad, isum);
          ad = __vabsdiffu4(qv.w, dv.w); isum = __dp4a(ad, ad, isum);
        }
        for (uint32_t o = parts >> 1; o >= 1; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
        if (c < n_sel && piece == 0) seld[c] = (float)isum;
      }
      __syncwarp();
      if (lane < n_sel) { id = sel[lane]; d = seld[lane]; }
    } else if (lane < n_sel) {
      id = sel[lane];
      d = exact_l2<DTYPE>(qrow, (const unsigned char*)pd.descI + (size_t)id * rb, dim);
    }
    const bool have = lane < n_sel;
    // lexicographic minimum over the lanes that hold a candidate, twice
    auto warp_min = [&](bool mine_ok, float dv, uint32_t iv, float* od, uint32_t* oi) -> int {
      float bd = dv;
      uint32_t bi = iv;
      int bl = mine_ok ? (int)lane : 64;
      for (int o = 16; o >= 1; o >>= 1) {
        const float xd = __shfl_xor_sync(0xffffffffu, bd, o);
        const uint32_t xi = __shfl_xor_sync(0xffffffffu, bi, o);
        const int xl = __shfl_xor_sync(0xffffffffu, bl, o);
        const bool other_better = xl < 64 && (bl >= 64 || vi_less(xd, xi, bd, bi));
        if (other_better) { bd = xd; bi = xi; bl = xl; }
      }
      *od = bd; *oi = bi;
      return bl;
    };
    Top2 t;
    const int l1 = warp_min(have, d, id, &t.d1, &t.i1);
    warp_min(have && (int)lane != l1, d, id, &t.d2, &t.i2);
    if (lane == 0) emit_result(pd, pair, q, t, ratio2, counters, matches, nullptr);
    __syncwarp();
  }
  if (lane == 0 && stat_raw) {  // reported through r3d_match_timing (third_chunk_queries / fifth_chunk_queries slots)
    atomicAdd(&counters[4], stat_raw);
    atomicAdd(&counters[3], stat_distinct);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

void cascade_release_view(DeviceWorker& w, ViewDev& v) {
  if (v.d_cascade) pool_release(w, v.d_cascade);
  v.d_cascade = nullptr;
  v.cascade_epoch = 0;
}

int cascade_prepare(r3d_ctx* ctx, DeviceWorker& w, const std::vector<uint32_t>& used) {
  const auto t_begin = std::chrono::steady_clock::now();
  R3D_CUDA_TRY(ctx, cudaSetDevice(w.device));
  if (used.empty()) return R3D_OK;
  uint32_t dim = 0;
  int dtype = -1;
  for (uint32_t id : used) {
    auto it = w.views.find(id);
    if (it == w.views.end()) return fail(ctx, R3D_ERR_INVALID, "cascade hashing: view " + std::to_string(id) + " was not uploaded");
    const ViewDev& v = it->second;
    if (v.n == 0) continue;
    if (dtype < 0) { dim = v.dim; dtype = (int)v.dtype; }
    if (v.dim != dim || (int)v.dtype != dtype) return fail(ctx, R3D_ERR_UNSUPPORTED, "cascade hashing: mixed descriptor types");
    if (v.n > kMaxCascadeRows) return fail(ctx, R3D_ERR_UNSUPPORTED, "cascade hashing: more than 65536 features in a view");
  }
  if (dtype < 0) return R3D_OK;  // nothing but empty views
  if (dim > 256 || dim == 0) return fail(ctx, R3D_ERR_UNSUPPORTED, "cascade hashing: descriptor dimension above 256");
  const uint32_t np = dim + kGroups * kBitsPerBucket, words = (dim + 31) / 32;
  // (1) projection table, [k][p]
  if (w.cascade_dim != dim) {
    std::vector<float> P((size_t)np * dim), PT((size_t)np * dim);
    {
      std::mt19937 gen(std::mt19937::default_seed);
      std::normal_distribution<> nd(0, 1);
      for (uint32_t i = 0; i < dim; ++i)
        for (uint32_t j = 0; j < dim; ++j) P[(size_t)i * dim + j] = (float)nd(gen);
      for (int g = 0; g < kGroups; ++g)
        for (int j = 0; j < kBitsPerBucket; ++j)
          for (uint32_t k = 0; k < dim; ++k) P[(size_t)(dim + g * kBitsPerBucket + j) * dim + k] = (float)nd(gen);
    }
    for (uint32_t p = 0; p < np; ++p)
      for (uint32_t k = 0; k < dim; ++k) PT[(size_t)k * np + p] = P[(size_t)p * dim + k];
    if (w.d_cascade_proj) pool_release(w, w.d_cascade_proj);
    w.d_cascade_proj = (float*)pool_alloc(w, PT.size() * sizeof(float));
    if (!w.d_cascade_proj) return fail(ctx, R3D_ERR_NOMEM, "cascade hashing: projection table");
    R3D_CUDA_TRY(ctx, cudaMemcpyAsync(w.d_cascade_proj, PT.data(), PT.size() * sizeof(float), cudaMemcpyHostToDevice, w.stream));
    R3D_CUDA_TRY(ctx, cudaStreamSynchronize(w.stream));
    w.cascade_dim = dim;
  }
  // (2) zero-mean descriptor over the used views, ascending id
  const uint32_t nu = (uint32_t)used.size();
  float* d_means = (float*)pool_alloc(w, ((size_t)nu + 1) * dim * sizeof(float));
  MeanJob* d_jobs = (MeanJob*)pool_alloc(w, (size_t)nu * sizeof(MeanJob));
  CascadeView* d_views = (CascadeView*)pool_alloc(w, (size_t)nu * sizeof(CascadeView));
  if (!d_means || !d_jobs || !d_views) return fail(ctx, R3D_ERR_NOMEM, "cascade hashing: scratch");
  std::vector<MeanJob> jobs(nu);
  for (uint32_t k = 0; k < nu; ++k) {
    const ViewDev& v = w.views.find(used[k])->second;
    jobs[k] = MeanJob{v.d_desc, v.n, 0u};
  }
  R3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(MeanJob), cudaMemcpyHostToDevice, w.stream));
  const uint32_t mean_threads = (dim + 31) & ~31u;
  if (dtype == 0) k_cascade_view_mean<0><<<nu, mean_threads, 0, w.stream>>>(d_jobs, dim, d_means);
  else k_cascade_view_mean<1><<<nu, mean_threads, 0, w.stream>>>(d_jobs, dim, d_means);
  float* d_zero = d_means + (size_t)nu * dim;
  k_cascade_zero_mean<<<1, mean_threads, 0, w.stream>>>(d_means, nu, dim, d_zero);
  R3D_CUDA_TRY(ctx, cudaGetLastError());
  // (3) per view: tables, hash codes; (4) buckets of all views in one launch
  const uint64_t epoch = ++ctx->cascade_epoch_counter;
  std::vector<CascadeView> hv(nu);
  std::vector<HashJob> hjobs;
  uint32_t total_blocks = 0;
  void* to_release = nullptr;
  for (uint32_t k = 0; k < nu; ++k) {
    ViewDev& v = w.views.find(used[k])->second;
    const size_t b_code = align16((size_t)v.n * words * 4), b_ofs = align16((size_t)kGroups * (kBuckets + 1) * 4),
                 b_ids = align16((size_t)kGroups * v.n * 4), b_bucket = align16((size_t)v.n * kGroups * 2);
    const size_t bytes = b_code + b_ofs + b_ids + b_bucket + 16;
    if (!v.d_cascade || v.cascade_bytes < bytes) {
      if (v.d_cascade) pool_release(w, v.d_cascade);
      v.d_cascade = pool_alloc(w, bytes);
      if (!v.d_cascade) return fail(ctx, R3D_ERR_NOMEM, "cascade hashing: view tables");
      v.cascade_bytes = bytes;
    }
    unsigned char* base = (unsigned char*)v.d_cascade;
    CascadeView& cv = hv[k];
    cv.code = (uint32_t*)base;
    cv.bk_ofs = (uint32_t*)(base + b_code);
    cv.bk_ids = (uint32_t*)(base + b_code + b_ofs);
    cv.bucket = (uint16_t*)(base + b_code + b_ofs + b_ids);
    cv.n = v.n;
    cv.words = words;
    v.cascade_epoch = epoch;
    v.cascade_index = k;
    if (v.n) {
      hjobs.push_back(HashJob{v.d_desc, cv.code, cv.bucket, v.n, total_blocks});
      total_blocks += (v.n + kHashRows - 1) / kHashRows;
    }
  }
  if (!hjobs.empty()) {
    HashJob* d_hjobs = (HashJob*)pool_alloc(w, hjobs.size() * sizeof(HashJob));
    if (!d_hjobs) return fail(ctx, R3D_ERR_NOMEM, "cascade hashing: job table");
    R3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_hjobs, hjobs.data(), hjobs.size() * sizeof(HashJob), cudaMemcpyHostToDevice, w.stream));
    const size_t smem = (size_t)kHashRows * dim * sizeof(float);
    if (dtype == 0) k_cascade_hash<0><<<total_blocks, 256, smem, w.stream>>>(d_hjobs, (uint32_t)hjobs.size(), dim, w.d_cascade_proj, d_zero, words);
    else k_cascade_hash<1><<<total_blocks, 256, smem, w.stream>>>(d_hjobs, (uint32_t)hjobs.size(), dim, w.d_cascade_proj, d_zero, words);
    to_release = d_hjobs;
  }
  R3D_CUDA_TRY(ctx, cudaGetLastError());
  R3D_CUDA_TRY(ctx, cudaMemcpyAsync(d_views, hv.data(), hv.size() * sizeof(CascadeView), cudaMemcpyHostToDevice, w.stream));
  k_cascade_buckets<<<dim3(kGroups, nu), 256, 0, w.stream>>>(d_views);
  R3D_CUDA_TRY(ctx, cudaGetLastError());
  R3D_CUDA_TRY(ctx, cudaStreamSynchronize(w.stream));  // hv / jobs are locals
  pool_release(w, d_means);
  pool_release(w, d_jobs);
  if (to_release) pool_release(w, to_release);
  if (w.d_cascade_views) pool_release(w, w.d_cascade_views);
  w.d_cascade_views = d_views;
  w.cascade_epoch = epoch;
  static const bool dbg = getenv("R3D_DEBUG_TIMING") != nullptr;
  if (dbg)
    fprintf(stderr, "[r3d] cascade hashing of %u views: %.2f ms\n", nu,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  return R3D_OK;
}

bool cascade_ready(const DeviceWorker& w, const uint32_t* pairs, uint64_t n_pairs) {
  if (!w.d_cascade_views || !w.cascade_epoch) return false;
  for (uint64_t p = 0; p < 2 * n_pairs; ++p) {
    auto it = w.views.find(pairs[p]);
    if (it == w.views.end()) return false;
    if (it->second.n && it->second.cascade_epoch != w.cascade_epoch) return false;
  }
  return true;
}

int launch_cascade_match(r3d_ctx* ctx, DeviceWorker& w, const PairDesc* d_pairs, const uint2* d_cidx, uint32_t n_pairs, uint32_t max_nJ,
                         uint32_t max_nI, uint32_t dim, int dtype, float ratio2, uint32_t* d_counters, uint2* d_matches) {
  if (!n_pairs || !max_nJ) return R3D_OK;
  const uint32_t bitmap_words = ((max_nI + 31) / 32 + 31) & ~31u;
  const uint32_t hist_words = (dim + 1 + 31u) & ~31u;
  const size_t smem = (size_t)kMatchWarps * (bitmap_words + hist_words + kListCap + 64) * sizeof(uint32_t);
  const uint32_t gx = std::min<uint32_t>((max_nJ + kMatchWarps - 1) / kMatchWarps, 64u);
  for (uint32_t p0 = 0; p0 < n_pairs; p0 += 65535u) {
    const uint32_t np = std::min(65535u, n_pairs - p0);
    if (dtype == 0) {
      R3D_CUDA_TRY(ctx, cudaFuncSetAttribute(k_cascade_match<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_cascade_match<0><<<dim3(gx, np), kMatchWarps * 32, smem, w.stream>>>(d_pairs + p0, d_cidx + p0, w.d_cascade_views, dim, ratio2,
                                                                            bitmap_words, d_counters, d_matches);
    } else {
      R3D_CUDA_TRY(ctx, cudaFuncSetAttribute(k_cascade_match<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_cascade_match<1><<<dim3(gx, np), kMatchWarps * 32, smem, w.stream>>>(d_pairs + p0, d_cidx + p0, w.d_cascade_views, dim, ratio2,
                                                                            bitmap_words, d_counters, d_matches);
    }
  }
  R3D_CUDA_TRY(ctx, cudaGetLastError());
  return R3D_OK;
}

}  // namespace r3d

// ---- C ABI -------------------------------------------------------------------------------------------
extern "C" int r3d_cascade_prepare(r3d_ctx* ctx, const uint32_t* view_ids, uint32_t n_views) try {
  if (!ctx || (!view_ids && n_views)) return r3d::fail(ctx, R3D_ERR_INVALID, "r3d_cascade_prepare: bad arguments");
  std::set<uint32_t> s(view_ids, view_ids + n_views);
  std::vector<uint32_t> used(s.begin(), s.end());
  for (auto& w : ctx->workers) {
    const int rc = r3d::cascade_prepare(ctx, w, used);
    if (rc) return rc;
  }
  return R3D_OK;
} catch (const std::exception& e) {
  return r3d::fail(ctx, R3D_ERR_NOMEM, std::string("r3d_cascade_prepare: ") + e.what());
}

extern "C" int r3d_debug_cascade_view(r3d_ctx* ctx, uint32_t view_id, uint32_t* code, uint16_t* bucket, uint32_t* bk_ofs, uint32_t* bk_ids) {
  if (!ctx) return R3D_ERR_INVALID;
  r3d::DeviceWorker& w = ctx->workers[0];
  auto it = w.views.find(view_id);
  if (it == w.views.end() || !it->second.d_cascade) return r3d::fail(ctx, R3D_ERR_INVALID, "r3d_debug_cascade_view: view has no hash tables");
  const r3d::ViewDev& v = it->second;
  R3D_CUDA_TRY(ctx, cudaSetDevice(w.device));
  r3d::CascadeView cv;
  R3D_CUDA_TRY(ctx, cudaMemcpy(&cv, w.d_cascade_views + v.cascade_index, sizeof(cv), cudaMemcpyDeviceToHost));
  if (code) R3D_CUDA_TRY(ctx, cudaMemcpy(code, cv.code, (size_t)v.n * cv.words * 4, cudaMemcpyDeviceToHost));
  if (bucket) R3D_CUDA_TRY(ctx, cudaMemcpy(bucket, cv.bucket, (size_t)v.n * 6 * 2, cudaMemcpyDeviceToHost));
  if (bk_ofs) R3D_CUDA_TRY(ctx, cudaMemcpy(bk_ofs, cv.bk_ofs, (size_t)6 * 1025 * 4, cudaMemcpyDeviceToHost));
  if (bk_ids) R3D_CUDA_TRY(ctx, cudaMemcpy(bk_ids, cv.bk_ids, (size_t)6 * v.n * 4, cudaMemcpyDeviceToHost));
  return R3D_OK;
}
