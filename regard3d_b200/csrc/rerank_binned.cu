// rerank_binned.cu -- exact re-rank of the candidate chunks, organised for data reuse.
//
// Stage A of the certification needs, for every query, the exact distances to the 2x16
// database rows of its two best candidate chunks.  Done query-by-query that is 32 scattered row
// reads per query (L2-bound: 18 KB per query at D=144).  Here the (query, chunk) incidences are
// binned by chunk first, so one warp stages the 16 rows of a chunk in shared memory ONCE and
// serves every query that selected it (~2*N_J/n_chunks of them):
//   k_bin_count  : per (pair, chunk) histogram, remembers each incidence's slot
//   k_bin_scan   : exclusive scan per pair
//   k_bin_fill   : per-chunk query lists
//   k_bin_rerank : warp per (pair, chunk): exact top-2 of the chunk for each listed query
//   k_bin_merge  : thread per query: merge the two partial top-2s, certify against key[2],
//                  ratio test + emit, or defer to k_rerank_list (stages B/C, then exact scan)
// All distances are "exact" in the sense of match_kernels.cu (upstream float accumulation order).
#include "r3d_internal.cuh"
#include "match_device.cuh"

namespace r3d {

struct Part {  // partial top-2 of one (query, chunk)
  float d1, d2;
  uint32_t i1, i2;
};

__global__ void __launch_bounds__(256) k_bin_count(const PairDesc* __restrict__ pairs,
                                                   const uint32_t* __restrict__ keys, uint32_t cstride,
                                                   uint32_t* __restrict__ cnt, uint32_t* __restrict__ slot,
                                                   uint32_t dim, float ratio2, int allow_reject) {
  const uint32_t pair = blockIdx.y;
  const PairDesc pd = pairs[pair];
  if (!pd.use_tc) return;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= pd.nJ) return;
  const uint32_t cmask = (1u << pd.chunk_bits) - 1u;
  const uint2 k = __ldg((const uint2*)(keys + (size_t)(pd.q_ofs + q) * kKeyStride));
  const uint32_t nchunks = pd.nI_pad / kChunk;
  if (stage_a_skips_query(pd, k.x, k.y, dim, ratio2, allow_reject)) return;  // cannot pass the ratio test
  uint32_t c0 = k.x & cmask, c1 = k.y & cmask;
  if (c0 >= nchunks) c0 = 0;  // sentinel keys (fewer than 2 real chunks): any valid chunk keeps the
  if (c1 >= nchunks) c1 = 0;  // bookkeeping regular; certification handles the rest
  slot[(size_t)(pd.q_ofs + q) * 2 + 0] = atomicAdd(&cnt[(size_t)pair * cstride + c0], 1u);
  slot[(size_t)(pd.q_ofs + q) * 2 + 1] = atomicAdd(&cnt[(size_t)pair * cstride + c1], 1u);
}

// in-place exclusive scan of cnt[pair][0..nchunks); one block per pair
__global__ void __launch_bounds__(256) k_bin_scan(const PairDesc* __restrict__ pairs, uint32_t cstride,
                                                  uint32_t* __restrict__ cnt) {
  __shared__ uint32_t part[256];
  const uint32_t pair = blockIdx.x;
  const PairDesc pd = pairs[pair];
  if (!pd.use_tc) return;
  const uint32_t nchunks = pd.nI_pad / kChunk;
  uint32_t* c = cnt + (size_t)pair * cstride;
  const uint32_t per = (nchunks + 255u) / 256u;
  const uint32_t b = threadIdx.x * per, e = min(b + per, nchunks);
  uint32_t s = 0;
  for (uint32_t i = b; i < e; ++i) s += c[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (uint32_t o = 1; o < 256; o <<= 1) {
    uint32_t v = 0;
    if (threadIdx.x >= o) v = part[threadIdx.x - o];
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t v = c[i];
    c[i] = run;
    run += v;
  }
  if (threadIdx.x == 255) c[nchunks] = part[255];  // total = end of the last chunk's list (cstride > nchunks)
}

__global__ void __launch_bounds__(256) k_bin_fill(const PairDesc* __restrict__ pairs,
                                                  const uint32_t* __restrict__ keys, uint32_t cstride,
                                                  const uint32_t* __restrict__ ofs, const uint32_t* __restrict__ slot,
                                                  uint32_t* __restrict__ list, uint32_t dim, float ratio2,
                                                  int allow_reject) {
  const uint32_t pair = blockIdx.y;
  const PairDesc pd = pairs[pair];
  if (!pd.use_tc) return;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= pd.nJ) return;
  const uint32_t cmask = (1u << pd.chunk_bits) - 1u;
  const uint2 k = __ldg((const uint2*)(keys + (size_t)(pd.q_ofs + q) * kKeyStride));
  const uint32_t nchunks = pd.nI_pad / kChunk;
  if (stage_a_skips_query(pd, k.x, k.y, dim, ratio2, allow_reject)) return;
  uint32_t c0 = k.x & cmask, c1 = k.y & cmask;
  if (c0 >= nchunks) c0 = 0;
  if (c1 >= nchunks) c1 = 0;
  uint32_t* l = list + (size_t)pd.q_ofs * 2;
  l[ofs[(size_t)pair * cstride + c0] + slot[(size_t)(pd.q_ofs + q) * 2 + 0]] = q * 2u + 0u;
  l[ofs[(size_t)pair * cstride + c1] + slot[(size_t)(pd.q_ofs + q) * 2 + 1]] = q * 2u + 1u;
}

// warp per (pair, chunk).  Shared memory: kBinWarps x kChunk rows x row_stride bytes, staged once per chunk and
// shared by every query that selected it; lane mapping: see the comment inside the kernel.
constexpr int kBinWarps = 8;

template <int DTYPE>
__global__ void __launch_bounds__(kBinWarps * 32) k_bin_rerank(const PairDesc* __restrict__ pairs, uint32_t cstride,
                                                               const uint32_t* __restrict__ ofs,
                                                               const uint32_t* __restrict__ list, uint32_t dim,
                                                               uint32_t row_stride, Part* __restrict__ parts) {
  extern __shared__ __align__(16) unsigned char smem_rows[];
  const uint32_t pair = blockIdx.y;
  const PairDesc pd = pairs[pair];
  if (!pd.use_tc) return;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t chunk = blockIdx.x * kBinWarps + warp;
  const uint32_t nchunks = pd.nI_pad / kChunk;
  if (chunk >= nchunks) return;
  const uint32_t* o = ofs + (size_t)pair * cstride;
  const uint32_t beg = o[chunk];
  const uint32_t end = o[chunk + 1];
  if (beg == end) return;
  const size_t rb = row_bytes(DTYPE, dim);
  unsigned char* rows = smem_rows + (size_t)warp * kChunk * row_stride;
  if ((rb & 15u) == 0) {  // stage the 16 rows with 128-bit loads (rows beyond nI are zero; masked below)
    const uint32_t vec_per_row = (uint32_t)(rb >> 4);
    const uint32_t total = kChunk * vec_per_row;
    for (uint32_t v = lane; v < total; v += 32) {
      const uint32_t r = v / vec_per_row, c = v - r * vec_per_row;
      const uint32_t col = chunk * kChunk + r;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (col < pd.nI) val = __ldg((const uint4*)((const char*)pd.descI + (size_t)col * rb) + c);
      *(uint4*)(rows + (size_t)r * row_stride + (size_t)c * 16) = val;
    }
  } else {  // rows are only 4-byte aligned (rb % 4 == 0 is guaranteed by the launcher)
    const uint32_t w_per_row = (uint32_t)(rb >> 2);
    const uint32_t total = kChunk * w_per_row;
    for (uint32_t v = lane; v < total; v += 32) {
      const uint32_t r = v / w_per_row, c = v - r * w_per_row;
      const uint32_t col = chunk * kChunk + r;
      uint32_t val = 0u;
      if (col < pd.nI) val = __ldg((const uint32_t*)((const char*)pd.descI + (size_t)col * rb) + c);
      *(uint32_t*)(rows + (size_t)r * row_stride + (size_t)c * 4) = val;
    }
  }
  __syncwarp();
  // Mapping: lane = (query slot, row): 32 / kChunk queries per pass, every lane accumulates ONE exact distance
  // (its query against row `r` of the chunk) in the strict upstream order; the chunk's top-2 of a query is a
  // butterfly merge over its kChunk lanes ((value, index)-lexicographic, hence order independent).  Cost per pass is
  // 1 / kChunk of a "lane = query, kChunk rows per lane" pass, so sparsely selected chunks -- the normal case once
  // the early rejection has removed the queries that cannot match -- no longer pay for 32 query slots.
  static_assert(kChunk == 8 || kChunk == 16, "lane mapping");
  constexpr uint32_t kQPerPass = 32 / kChunk;
  const uint32_t r = lane % kChunk, qslot = lane / kChunk;
  const unsigned char* arow = rows + (size_t)r * row_stride;
  uint32_t row_norm = 0u;  // uint8 only: |a_r|^2
  if (DTYPE != 0) {
    for (uint32_t g = 0; g < (dim >> 2); ++g) {
      const uint32_t a = *(const uint32_t*)(arow + (size_t)g * 4);
      row_norm = __dp4a(a, a, row_norm);
    }
  }
  const uint32_t* l = list + (size_t)pd.q_ofs * 2;
  const uint32_t col0 = chunk * kChunk;
  const uint32_t nvalid = (pd.nI > col0) ? min((uint32_t)kChunk, pd.nI - col0) : 0u;
  for (uint32_t e0 = beg; e0 < end; e0 += kQPerPass) {
    const uint32_t e = e0 + qslot;
    const bool live = e < end;
    const uint32_t qs = live ? __ldg(l + e) : __ldg(l + beg);  // dead slots shadow a valid entry
    const char* qrow = (const char*)pd.descJ + (size_t)(qs >> 1) * rb;
    float acc = 0.f;
    if (DTYPE == 0) {
      const uint32_t groups = dim >> 2;
      if ((rb & 15u) == 0) {
#pragma unroll 4
        for (uint32_t g = 0; g < groups; ++g) {
          const float4 qv = __ldg((const float4*)qrow + g);
          const float4 a = *(const float4*)(arow + (size_t)g * 16);
          acc = acc4(acc, __fsub_rn(qv.x, a.x), __fsub_rn(qv.y, a.y), __fsub_rn(qv.z, a.z), __fsub_rn(qv.w, a.w));
        }
      } else {
        for (uint32_t g = 0; g < groups; ++g) {
          const float* qf = (const float*)qrow + 4 * g;
          const float* af = (const float*)(arow + (size_t)g * 16);
          acc = acc4(acc, __fsub_rn(__ldg(qf), af[0]), __fsub_rn(__ldg(qf + 1), af[1]), __fsub_rn(__ldg(qf + 2), af[2]),
                     __fsub_rn(__ldg(qf + 3), af[3]));
        }
      }
      for (uint32_t k = groups * 4; k < dim; ++k) {
        const float df = __fsub_rn(__ldg((const float*)qrow + k), *(const float*)(arow + (size_t)k * 4));
        acc = __fadd_rn(acc, __fmul_rn(df, df));
      }
    } else {
      // uint8 descriptors: every partial sum of the upstream float accumulation is an integer below 2^24
      // (dim <= 240: 240 * 255^2 < 2^24), so the float result IS the integer sum -- computed here as
      // |q|^2 + |a|^2 - 2 q.a with IDP4A instead of I2F / FSUB / FMUL / FADD per element
      uint32_t dot = 0u, qn = 0u;
      if ((rb & 15u) == 0) {
        const uint32_t g16 = (uint32_t)(rb >> 4);
        for (uint32_t g = 0; g < g16; ++g) {
          const uint4 qv = __ldg((const uint4*)qrow + g);
          const uint4 a = *(const uint4*)(arow + (size_t)g * 16);
          qn = __dp4a(qv.x, qv.x, qn); qn = __dp4a(qv.y, qv.y, qn); qn = __dp4a(qv.z, qv.z, qn); qn = __dp4a(qv.w, qv.w, qn);
          dot = __dp4a(qv.x, a.x, dot); dot = __dp4a(qv.y, a.y, dot); dot = __dp4a(qv.z, a.z, dot); dot = __dp4a(qv.w, a.w, dot);
        }
      } else {
        const uint32_t groups = dim >> 2;  // rb % 4 == 0 is guaranteed by the launcher
        for (uint32_t g = 0; g < groups; ++g) {
          const uint32_t qv = __ldg((const uint32_t*)qrow + g);
          qn = __dp4a(qv, qv, qn);
          dot = __dp4a(qv, *(const uint32_t*)(arow + (size_t)g * 4), dot);
        }
      }
      acc = (float)(qn + row_norm - 2u * dot);
    }
    Top2 t;
    t.d1 = t.d2 = FLT_MAX; t.i1 = t.i2 = 0xffffffffu;
    if (r < nvalid) { t.d1 = acc; t.i1 = col0 + r; }
#pragma unroll
    for (uint32_t o = kChunk / 2; o >= 1; o >>= 1) {
      Top2 u;
      u.d1 = __shfl_xor_sync(0xffffffffu, t.d1, o);
      u.d2 = __shfl_xor_sync(0xffffffffu, t.d2, o);
      u.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
      u.i2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
      t = top2_merge(t, u);
    }
    if (live && r == 0) {
      Part p;
      p.d1 = t.d1; p.d2 = t.d2; p.i1 = t.i1; p.i2 = t.i2;
      parts[(size_t)(pd.q_ofs + (qs >> 1)) * 2 + (qs & 1u)] = p;
    }
  }
}

__global__ void __launch_bounds__(256) k_bin_merge(const PairDesc* __restrict__ pairs,
                                                   const uint32_t* __restrict__ keys, const Part* __restrict__ parts,
                                                   uint32_t dim, float ratio2, uint32_t* counters, uint2* matches,
                                                   uint2* list2, uint2* fallback, float4* nn) {
  const uint32_t pair = blockIdx.y;
  const PairDesc pd = pairs[pair];
  if (!pd.use_tc) return;
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= pd.nJ) return;
  const uint32_t cmask = (1u << pd.chunk_bits) - 1u;
  const uint4 k = __ldg((const uint4*)(keys + (size_t)(pd.q_ofs + q) * kKeyStride));
  if (stage_a_skips_query(pd, k.x, k.y, dim, ratio2, nn == nullptr)) {  // no match can come out of this query
    atomicAdd(&counters[5], 1u);
    return;
  }
  const Part a = parts[(size_t)(pd.q_ofs + q) * 2 + 0];
  const Part b = parts[(size_t)(pd.q_ofs + q) * 2 + 1];
  Top2 ta, tb;
  ta.d1 = a.d1; ta.d2 = a.d2; ta.i1 = a.i1; ta.i2 = a.i2;
  tb.d1 = b.d1; tb.d2 = b.d2; tb.i1 = b.i1; tb.i2 = b.i2;
  const uint32_t nchunks = pd.nI_pad / kChunk;
  // identical chunks (only possible through the sentinel remap) must not be merged twice
  const bool dup = ((k.x & cmask) >= nchunks) || ((k.y & cmask) >= nchunks) || ((k.x & cmask) == (k.y & cmask));
  bool ok = false;
  Top2 t = ta;
  if (!dup) {
    t = top2_merge(ta, tb);
    const double pack_rel = ldexp(1.0, (int)pd.chunk_bits - 23);
    const double gamma = (double)(dim + 16) * (1.0 / 16777216.0);
    ok = key_lower_bound(k.z, pd.eps_abs, gamma, pack_rel) > (double)t.d2;
  }
  if (ok) {
    emit_result(pd, pair, q, t, ratio2, counters, matches, nn);
  } else if (dup) {  // fewer than two real candidate chunks (tiny database): exact scan
    const uint32_t s = atomicAdd(&counters[1], 1u);
    fallback[s] = make_uint2(pair, q);
  } else {
    const uint32_t s = atomicAdd(&counters[4], 1u);
    list2[s] = make_uint2(pair, q);
  }
}

// ------------------------------------------------------------------------------------------------
int launch_rerank_binned(r3d_ctx* ctx, DeviceWorker& w, const PairDesc* d_pairs, uint32_t n_pairs, uint32_t max_nJ,
                         uint32_t cstride, const uint32_t* d_keys, uint32_t dim, int dtype, float ratio2,
                         uint32_t* d_cnt, uint32_t* d_slot, uint32_t* d_list, void* d_parts, uint32_t* d_counters,
                         uint2* d_matches, uint2* d_list2, uint2* d_fallback, float4* d_nn) {
  if (n_pairs == 0 || max_nJ == 0) return R3D_OK;
  const size_t rb = dtype == 0 ? (size_t)dim * 4 : (size_t)dim;
  if (rb & 3) return fail(ctx, R3D_ERR_UNSUPPORTED, "binned re-rank needs row bytes % 4 == 0");
  // Row pitch in shared memory: a quarter-warp (8 lanes) reads 16 bytes of 8 DIFFERENT rows per LDS.128, so the pitch
  // in 4-byte words must be 4 * odd (mod 32) for the eight 4-bank groups to be distinct (a pitch of 576 B = 144
  // words = 16 mod 32 was a 4-way bank conflict and made the kernel L1/shared-pipe bound at 98 %).
  uint32_t row_stride = (uint32_t)((rb + 15) / 16 * 16);
  while (((row_stride / 4) % 32) % 8 != 4) row_stride += 16;
  const size_t smem = (size_t)kBinWarps * kChunk * row_stride;
  R3D_CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, (size_t)n_pairs * cstride * sizeof(uint32_t), w.stream));
  dim3 gq((max_nJ + 255) / 256, n_pairs);
  const int allow_reject = d_nn == nullptr ? 1 : 0;
  k_bin_count<<<gq, 256, 0, w.stream>>>(d_pairs, d_keys, cstride, d_cnt, d_slot, dim, ratio2, allow_reject);
  k_bin_scan<<<n_pairs, 256, 0, w.stream>>>(d_pairs, cstride, d_cnt);
  k_bin_fill<<<gq, 256, 0, w.stream>>>(d_pairs, d_keys, cstride, d_cnt, d_slot, d_list, dim, ratio2, allow_reject);
  dim3 gc((cstride + kBinWarps - 1) / kBinWarps, n_pairs);
  if (dtype == 0) {
    R3D_CUDA_TRY(ctx, cudaFuncSetAttribute(k_bin_rerank<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bin_rerank<0><<<gc, kBinWarps * 32, smem, w.stream>>>(d_pairs, cstride, d_cnt, d_list, dim, row_stride, (Part*)d_parts);
  } else {
    R3D_CUDA_TRY(ctx, cudaFuncSetAttribute(k_bin_rerank<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bin_rerank<1><<<gc, kBinWarps * 32, smem, w.stream>>>(d_pairs, cstride, d_cnt, d_list, dim, row_stride, (Part*)d_parts);
  }
  k_bin_merge<<<gq, 256, 0, w.stream>>>(d_pairs, d_keys, (const Part*)d_parts, dim, ratio2, d_counters, d_matches,
                                        d_list2, d_fallback, d_nn);
  R3D_CUDA_TRY(ctx, cudaGetLastError());
  return R3D_OK;
}

}  // namespace r3d
