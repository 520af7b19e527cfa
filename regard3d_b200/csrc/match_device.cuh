// match_device.cuh -- device helpers shared by the matching kernels (exact distance in the upstream
// accumulation order, top-2 bookkeeping, certification bound, result emission).
#pragma once
#include "r3d_internal.cuh"

#include <cfloat>

namespace r3d {

// ------------------------------------------------------------------------------------------------
// exact distance, upstream accumulation order
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float acc4(float acc, float d0, float d1, float d2, float d3) {
  const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)),
                            __fmul_rn(d3, d3));
  return __fadd_rn(acc, s);
}

template <int DTYPE>
__device__ __forceinline__ float exact_l2(const void* __restrict__ qrow, const void* __restrict__ drow,
                                           uint32_t dim) {
  float acc = 0.f;
  if (DTYPE == 0) {
    const float* q = (const float*)qrow;
    const float* d = (const float*)drow;
    uint32_t k = 0;
    if ((dim & 3u) == 0) {
      const float4* q4 = (const float4*)q;
      const float4* d4 = (const float4*)d;
      const uint32_t g = dim >> 2;
#pragma unroll 4
      for (uint32_t t = 0; t < g; ++t) {
        const float4 a = q4[t];  // may live in shared memory
        const float4 b = __ldg(d4 + t);
        acc = acc4(acc, __fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z), __fsub_rn(a.w, b.w));
      }
      k = dim;
    } else {
      for (; k + 3 < dim; k += 4)
        acc = acc4(acc, __fsub_rn(q[k], d[k]), __fsub_rn(q[k + 1], d[k + 1]), __fsub_rn(q[k + 2], d[k + 2]),
                   __fsub_rn(q[k + 3], d[k + 3]));
    }
    for (; k < dim; ++k) {
      const float df = __fsub_rn(q[k], d[k]);
      acc = __fadd_rn(acc, __fmul_rn(df, df));
    }
  } else {
    const uint8_t* q = (const uint8_t*)qrow;
    const uint8_t* d = (const uint8_t*)drow;
    uint32_t k = 0;
    if ((dim & 3u) == 0 && dim <= 256u) {  // 255^2 * 256 < 2^24: beyond that the float accumulation may round
      const uint32_t* q4 = (const uint32_t*)q;
      const uint32_t* d4 = (const uint32_t*)d;
      const uint32_t g = dim >> 2;
#pragma unroll 4
      // integer form: all partial sums of the upstream float accumulation are integers < 2^24 (dim <= 256),
      // so the float result equals the exact integer sum of squared byte differences
      uint32_t isum = 0u;
      for (uint32_t t = 0; t < g; ++t) {
        const uint32_t ad = __vabsdiffu4(q4[t], __ldg(d4 + t));
        isum = __dp4a(ad, ad, isum);
      }
      acc = (float)isum;
      k = dim;
    } else {
      for (; k + 3 < dim; k += 4)
        acc = acc4(acc, (float)((int)q[k] - (int)d[k]), (float)((int)q[k + 1] - (int)d[k + 1]),
                   (float)((int)q[k + 2] - (int)d[k + 2]), (float)((int)q[k + 3] - (int)d[k + 3]));
    }
    for (; k < dim; ++k) {
      const float df = (float)((int)q[k] - (int)d[k]);
      acc = __fadd_rn(acc, __fmul_rn(df, df));
    }
  }
  return acc;
}

__device__ __forceinline__ size_t row_bytes(int dtype, uint32_t dim) { return dtype == 0 ? (size_t)dim * 4 : (size_t)dim; }

// lexicographic (value, index) "less" -- symmetric tie-break so butterfly merges agree on all lanes
__device__ __forceinline__ bool vi_less(float a, uint32_t ia, float b, uint32_t ib) {
  return (a < b) || (a == b && ia < ib);
}

struct Top2 {
  float d1, d2;
  uint32_t i1, i2;
};

__device__ __forceinline__ void top2_insert(Top2& t, float d, uint32_t i) {
  if (vi_less(d, i, t.d1, t.i1)) {
    t.d2 = t.d1; t.i2 = t.i1; t.d1 = d; t.i1 = i;
  } else if (vi_less(d, i, t.d2, t.i2)) {
    t.d2 = d; t.i2 = i;
  }
}

__device__ __forceinline__ Top2 top2_merge(const Top2& a, const Top2& b) {
  Top2 r;
  if (vi_less(a.d1, a.i1, b.d1, b.i1)) {
    r.d1 = a.d1; r.i1 = a.i1;
    if (vi_less(a.d2, a.i2, b.d1, b.i1)) { r.d2 = a.d2; r.i2 = a.i2; } else { r.d2 = b.d1; r.i2 = b.i1; }
  } else {
    r.d1 = b.d1; r.i1 = b.i1;
    if (vi_less(b.d2, b.i2, a.d1, a.i1)) { r.d2 = b.d2; r.i2 = b.i2; } else { r.d2 = a.d1; r.i2 = a.i1; }
  }
  return r;
}

__device__ __forceinline__ Top2 top2_warp_reduce(Top2 t) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    Top2 b;
    b.d1 = __shfl_xor_sync(0xffffffffu, t.d1, o);
    b.d2 = __shfl_xor_sync(0xffffffffu, t.d2, o);
    b.i1 = __shfl_xor_sync(0xffffffffu, t.i1, o);
    b.i2 = __shfl_xor_sync(0xffffffffu, t.i2, o);
    t = top2_merge(t, b);
  }
  return t;
}

// Lower bound on the ORACLE-ORDER float distance of any database column whose chunk has packed
// key `key` (see DESIGN.md "certification").
__device__ __forceinline__ double key_lower_bound(uint32_t key, float eps_abs, double gamma, double pack_rel) {
  const double kv = (double)__uint_as_float(key);
  double lb = kv - fabs(kv) * pack_rel - (double)eps_abs;
  if (lb > 0.0) lb = lb * (1.0 - gamma);
  return lb;
}

// Early rejection (matching mode only; r3d_search_neighbours needs every neighbour): the ratio test
// d1 < fl(ratio2 * d2) cannot pass when a LOWER bound of the nearest distance already reaches ratio2 times an
// UPPER bound of the second-nearest.  E1 >= LB(key0) because key0 is the smallest chunk key; E2 <= UB(key1)
// because the minima of the two best chunks are two distinct database rows whose exact distances are at most
// UB(key0) <= UB(key1).  Same two-sided error model as the certification: |key value - real distance| <=
// eps_abs + |key| * pack_rel (tests/test_gpu_match.py::test_candidate_error_bound_holds), float recipe within
// gamma of the real number, one more rounding for the product.  Such a query needs no exact distance at all.
__device__ __forceinline__ bool ratio_test_cannot_pass(uint32_t key0, uint32_t key1, float eps_abs, double gamma,
                                                       double pack_rel, float ratio2) {
  const double lb = key_lower_bound(key0, eps_abs, gamma, pack_rel);
  const double kv1 = (double)__uint_as_float(key1);
  double ub = (kv1 + fabs(kv1) * pack_rel + (double)eps_abs) * (1.0 + gamma);
  if (!(ub > 0.0)) ub = 0.0;
  return lb >= (double)ratio2 * ub * (1.0 + 1.2e-7);
}

// the (query, chunk) bookkeeping of stage A and this test, shared by k_bin_count / k_bin_fill / k_bin_merge so that
// the three kernels take the same decision
__device__ __forceinline__ bool stage_a_skips_query(const PairDesc& pd, uint32_t key0, uint32_t key1, uint32_t dim,
                                                    float ratio2, int allow_reject) {
  if (!allow_reject) return false;
  const uint32_t cmask = (1u << pd.chunk_bits) - 1u;
  const uint32_t nchunks = pd.nI_pad / kChunk;
  const uint32_t c0 = key0 & cmask, c1 = key1 & cmask;
  if (c0 >= nchunks || c1 >= nchunks || c0 == c1) return false;  // sentinel keys: the regular path sorts it out
  const double pack_rel = ldexp(1.0, (int)pd.chunk_bits - 23);
  const double gamma = (double)(dim + 16) * (1.0 / 16777216.0);
  return ratio_test_cannot_pass(key0, key1, pd.eps_abs, gamma, pack_rel, ratio2);
}

// A query that passes the ratio test appends IndMatch(i in I, j = q in J) to ITS PAIR's segment of the
// dense match array (segment = the pair's query rows, so it can never overflow); the per-pair counters
// live behind the 16 scalar counters.  k_pack_matches then packs the segments for the host copy: the
// matches arrive on the host already bucketed by pair.
constexpr uint32_t kPairCounterBase = 16;
__device__ __forceinline__ void emit_result(const PairDesc& pd, uint32_t pair, uint32_t q, const Top2& t,
                                            float ratio2, uint32_t* counters, uint2* matches, float4* nn) {
  if (nn) {
    nn[pd.q_ofs + q] = make_float4(__uint_as_float(t.i1), __uint_as_float(t.i2), t.d1, t.d2);
  }
  if (matches && t.d1 < __fmul_rn(ratio2, t.d2)) {  // NNdistanceRatio: strict, float
    const uint32_t slot = atomicAdd(&counters[kPairCounterBase + pair], 1u);
    matches[pd.q_ofs + slot] = make_uint2(t.i1, q);
  }
}


}  // namespace r3d
