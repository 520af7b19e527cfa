// oracle_cascade.cpp -- CPU ORACLE (test infrastructure; see oracle.h): OpenMVG's CASCADE_HASHING_L2 matcher.
//
// BASELINE config 4 names "cascade-hashing ANN"; Regard3D itself never selects it (no CASCADE anywhere under
// /root/reference/src; the GUI offers FLANN / KGraph / MRPT / HNSW, src/R3DComputeMatches.cpp:2035-2062), it is
// OpenMVG 1.4's matching_image_collection/Cascade_Hashing_Matcher_Regions + matching/cascade_hasher.hpp, un-vendored.
// PARITY UNPINNED: restated from the published algorithm (Cheng et al., "Fast and accurate image matching with
// cascade hashing for 3D reconstruction", CVPR 2014) and the upstream structure as summarised in SURVEY.md A.8:
//   CascadeHasher::Init(dim)        hash length = descriptor dimension; primary projection dim x dim, 6 groups of
//                                   10 secondary projections (10 x dim), all N(0,1) drawn row by row from
//                                   std::mt19937(default_seed) through ONE std::normal_distribution<>
//   zero-mean descriptor            mean over the used views (ascending view id) of each view's mean descriptor, float
//   CreateHashedDescriptions        bit j = (P (d - mean))_j > 0; bucket id of group g = its 10 sign bits, first = MSB;
//                                   buckets hold descriptor ids in ascending order
//   Match_HashedDescriptions        query = view J, database = view I: candidates = the 6 buckets of the query's ids
//                                   (skip the query when the raw count is <= NN = 2), first occurrences only,
//                                   counting sort by Hamming distance (stable), exact L2 of the first 10,
//                                   std::partial_sort of (distance, id) pairs, keep 2 when at least 2
//   then NNdistanceRatio(Square(ratio)), IndMatch(i in I, j in J), getDeduplicated, coordinate de-duplication --
//   the same tail as the brute-force matcher (oracle_match.cpp).
// Deviation (fidelity, documented in DESIGN.md): Eigen evaluates the projections as a vectorised float gemv whose
// summation order is build-dependent; here every projection is ONE float accumulator over k = 0 .. dim-1, product and
// sum rounded separately -- the order the GPU reproduces.
#include "oracle_internal.hpp"

#include <algorithm>
#include <cstring>
#include <random>
#include <set>
#include <omp.h>

namespace orc {
namespace {

constexpr int kGroups = 6, kBitsPerBucket = 10, kBuckets = 1 << kBitsPerBucket, kTopCandidates = 10, kNN = 2;

struct Hasher {
  uint32_t dim = 0;
  std::vector<float> primary;              // dim x dim, row-major
  std::vector<float> secondary[kGroups];   // 10 x dim each
  void init(uint32_t d) {
    dim = d;
    std::mt19937 gen(std::mt19937::default_seed);
    std::normal_distribution<> nd(0, 1);
    primary.resize((size_t)d * d);
    for (uint32_t i = 0; i < d; ++i)
      for (uint32_t j = 0; j < d; ++j) primary[(size_t)i * d + j] = (float)nd(gen);
    for (int g = 0; g < kGroups; ++g) {
      secondary[g].resize((size_t)kBitsPerBucket * d);
      for (int j = 0; j < kBitsPerBucket; ++j)
        for (uint32_t k = 0; k < d; ++k) secondary[g][(size_t)j * d + k] = (float)nd(gen);
    }
  }
};

struct Hashed {
  uint32_t n = 0, words = 0;
  std::vector<uint32_t> code;                 // n x words
  std::vector<uint16_t> bucket_id;            // n x 6
  std::vector<std::vector<uint32_t>> bucket;  // 6 * 1024 lists of descriptor ids
};

inline float elem(const void* desc, int dtype, size_t idx) {
  return dtype == 0 ? ((const float*)desc)[idx] : (float)((const uint8_t*)desc)[idx];
}

// CascadeHasher::GetZeroMeanDescriptor: running float sums over the rows, then one division
void mean_rows(const void* desc, int dtype, uint32_t n, uint32_t dim, float* out) {
  for (uint32_t j = 0; j < dim; ++j) out[j] = 0.f;
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < dim; ++j) out[j] += elem(desc, dtype, (size_t)i * dim + j);
  for (uint32_t j = 0; j < dim; ++j) out[j] = out[j] / (float)n;
}

float project(const float* row, const float* d, uint32_t dim) {
  float acc = 0.f;
  for (uint32_t k = 0; k < dim; ++k) acc += row[k] * d[k];
  return acc;
}

void hash_view(const Hasher& H, const void* desc, int dtype, uint32_t n, const float* zero_mean, Hashed& out) {
  const uint32_t dim = H.dim;
  out.n = n;
  out.words = (dim + 31) / 32;
  out.code.assign((size_t)n * out.words, 0u);
  out.bucket_id.assign((size_t)n * kGroups, 0);
  out.bucket.assign((size_t)kGroups * kBuckets, {});
  std::vector<float> d(dim);
  for (uint32_t i = 0; i < n; ++i) {
    for (uint32_t k = 0; k < dim; ++k) d[k] = elem(desc, dtype, (size_t)i * dim + k) - zero_mean[k];
    for (uint32_t j = 0; j < dim; ++j)
      if (project(&H.primary[(size_t)j * dim], d.data(), dim) > 0) out.code[(size_t)i * out.words + j / 32] |= 1u << (j & 31);
    for (int g = 0; g < kGroups; ++g) {
      uint16_t id = 0;
      for (int k = 0; k < kBitsPerBucket; ++k)
        id = (uint16_t)((id << 1) + (project(&H.secondary[g][(size_t)k * dim], d.data(), dim) > 0 ? 1 : 0));
      out.bucket_id[(size_t)i * kGroups + g] = id;
    }
  }
  for (int g = 0; g < kGroups; ++g)
    for (uint32_t j = 0; j < n; ++j) out.bucket[(size_t)g * kBuckets + out.bucket_id[(size_t)j * kGroups + g]].push_back(j);
}

// Match_HashedDescriptions(query = J, database = I) + the ratio test; emits IndMatch(i in I, j in J)
void match_hashed(const Hashed& hq, const void* descQ, const Hashed& hd, const void* descD, uint32_t dim, int dtype,
                  float f_dist_ratio, std::vector<orc_indmatch>& out) {
  const float fratio = f_dist_ratio * f_dist_ratio;
  std::vector<uint32_t> candidates;
  std::vector<uint8_t> used(hd.n, 0);
  std::vector<std::vector<uint32_t>> by_hamming(dim + 1);
  std::vector<std::pair<float, int>> eucl;
  for (uint32_t q = 0; q < hq.n; ++q) {
    candidates.clear();
    for (int g = 0; g < kGroups; ++g) {
      const auto& b = hd.bucket[(size_t)g * kBuckets + hq.bucket_id[(size_t)q * kGroups + g]];
      for (uint32_t id : b) {
        candidates.push_back(id);
        used[id] = 0;
      }
    }
    if (candidates.size() <= (size_t)kNN) continue;
    for (auto& v : by_hamming) v.clear();
    for (uint32_t id : candidates) {
      if (used[id]) continue;
      used[id] = 1;
      uint32_t h = 0;
      for (uint32_t w = 0; w < hq.words; ++w) h += (uint32_t)__builtin_popcount(hq.code[(size_t)q * hq.words + w] ^ hd.code[(size_t)id * hd.words + w]);
      by_hamming[h].push_back(id);
    }
    eucl.clear();
    for (uint32_t h = 0; h <= dim && eucl.size() < (size_t)kTopCandidates; ++h)
      for (size_t k = 0; k < by_hamming[h].size() && eucl.size() < (size_t)kTopCandidates; ++k) {
        const uint32_t id = by_hamming[h][k];
        const float dist = dtype == 0 ? l2_f32((const float*)descD + (size_t)id * dim, (const float*)descQ + (size_t)q * dim, dim)
                                      : l2_u8((const uint8_t*)descD + (size_t)id * dim, (const uint8_t*)descQ + (size_t)q * dim, dim);
        eucl.emplace_back(dist, (int)id);
      }
    if (eucl.size() < (size_t)kNN) continue;
    std::partial_sort(eucl.begin(), eucl.begin() + kNN, eucl.end());
    if (eucl[0].first < fratio * eucl[1].first) out.push_back(orc_indmatch{(uint32_t)eucl[0].second, q});
  }
}

struct IndMatchLess2 {
  bool operator()(const orc_indmatch& a, const orc_indmatch& b) const { return (a.i < b.i) || (a.i == b.i && a.j < b.j); }
};

}  // namespace

void cascade_match_pairs(const void* const* descs, const float* const* xys, const uint32_t* ns, uint32_t n_views,
                         uint32_t dim, int dtype, const uint32_t* pairs, uint64_t P, float ratio,
                         std::map<std::pair<uint32_t, uint32_t>, std::vector<orc_indmatch>>& out, int n_threads) {
  (void)n_views;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  Hasher H;
  H.init(dim);
  std::set<uint32_t> used_index;
  std::map<uint32_t, std::vector<uint32_t>> map_pairs;
  for (uint64_t p = 0; p < P; ++p) {
    used_index.insert(pairs[2 * p]);
    used_index.insert(pairs[2 * p + 1]);
    map_pairs[pairs[2 * p]].push_back(pairs[2 * p + 1]);
  }
  // zero-mean descriptor: mean of the per-view means (a view without features contributes a zero row)
  std::vector<float> zero_mean(dim, 0.f);
  {
    std::vector<float> rows((size_t)used_index.size() * dim, 0.f);
    size_t r = 0;
    for (uint32_t v : used_index) {
      if (ns[v] > 0) mean_rows(descs[v], dtype, ns[v], dim, &rows[r * dim]);
      ++r;
    }
    if (!used_index.empty()) mean_rows(rows.data(), 0, (uint32_t)used_index.size(), dim, zero_mean.data());
  }
  std::map<uint32_t, Hashed> hashed;
  {
    std::vector<uint32_t> ids(used_index.begin(), used_index.end());
    for (uint32_t v : ids) hashed[v];
#pragma omp parallel for schedule(dynamic) num_threads(n_threads)
    for (int k = 0; k < (int)ids.size(); ++k) hash_view(H, descs[ids[k]], dtype, ns[ids[k]], zero_mean.data(), hashed[ids[k]]);
  }
  for (const auto& kv : map_pairs) {
    const uint32_t I = kv.first;
    const auto& index_to_compare = kv.second;
    if (ns[I] == 0) continue;
#pragma omp parallel for schedule(dynamic) num_threads(n_threads)
    for (int j = 0; j < (int)index_to_compare.size(); ++j) {
      const uint32_t J = index_to_compare[j];
      std::vector<orc_indmatch> v;
      match_hashed(hashed[J], descs[J], hashed[I], descs[I], dim, dtype, ratio, v);
      std::set<orc_indmatch, IndMatchLess2> s(v.begin(), v.end());  // IndMatch::getDeduplicated
      v.assign(s.begin(), s.end());
      coord_dedup(v, xys[I], xys[J]);
#pragma omp critical
      {
        if (!v.empty()) out.insert({{I, J}, std::move(v)});
      }
    }
  }
}

}  // namespace orc

extern "C" {

int64_t orc_cascade_match_pairs(const void* const* descs, const float* const* xys, const uint32_t* ns, uint32_t n_views,
                                uint32_t dim, int dtype, const uint32_t* pairs, uint64_t P, float ratio,
                                uint64_t* pair_ofs, orc_indmatch* out, uint64_t cap, int n_threads) {
  std::map<std::pair<uint32_t, uint32_t>, std::vector<orc_indmatch>> res;
  orc::cascade_match_pairs(descs, xys, ns, n_views, dim, dtype, pairs, P, ratio, res, n_threads);
  uint64_t ofs = 0;
  for (uint64_t p = 0; p < P; ++p) {
    pair_ofs[p] = ofs;
    auto it = res.find({pairs[2 * p], pairs[2 * p + 1]});
    if (it == res.end()) continue;
    if (ofs + it->second.size() > cap) return -1;
    std::memcpy(out + ofs, it->second.data(), it->second.size() * sizeof(orc_indmatch));
    ofs += it->second.size();
  }
  pair_ofs[P] = ofs;
  return (int64_t)ofs;
}

/* the projections (tests compare the product's host-side table with this one): primary dim x dim, then 6 x 10 x dim */
void orc_cascade_projections(uint32_t dim, float* out) {
  orc::Hasher H;
  H.init(dim);
  std::memcpy(out, H.primary.data(), H.primary.size() * sizeof(float));
  for (int g = 0; g < orc::kGroups; ++g)
    std::memcpy(out + (size_t)dim * dim + (size_t)g * orc::kBitsPerBucket * dim, H.secondary[g].data(),
                H.secondary[g].size() * sizeof(float));
}
}
