// oracle_match.cpp -- CPU ORACLE (test infrastructure; see oracle.h header comment).
//
// Restates the putative-matching stage that R3DComputeMatches::computeMatches() delegates to
// OpenMVG (reference call sites: src/R3DComputeMatches.cpp:2035-2048; the reference's own verbatim
// copy of the upstream pair loop is src/R3DComputeMatches.cpp:423-491).  Upstream semantics restated
// from OpenMVG 1.4 (un-vendored):
//   matching/metric.hpp              L2<T>::operator()      -> l2_f32 / l2_u8
//   matching/matcher_brute_force.hpp ArrayMatcherBruteForce -> search_neighbours
//   matching/matching_filters.hpp    NNdistanceRatio        -> inside match_distance_ratio
//   matching/regions_matcher.hpp     MatchDistanceRatio     -> match_distance_ratio
//   matching/indMatch.hpp            getDeduplicated        -> std::set<IndMatch>
//   matching/indMatchDecoratorXY.hpp IndMatchDecorator      -> coord_dedup
// PARITY UNPINNED (no reference tests / golden vectors exist).
#include "oracle.h"
#include "oracle_internal.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <vector>
#include <omp.h>

namespace orc {

// ---- A.1  openMVG::matching::L2<float> (referenced at src/R3DComputeMatches.cpp:290-291) ----
float l2_f32(const float* a, const float* b, size_t size) {
  float result = 0.0f;
  float diff0, diff1, diff2, diff3;
  const float* last = a + size;
  const float* lastgroup = last - 3;
  while (a < lastgroup) {
    diff0 = a[0] - b[0];
    diff1 = a[1] - b[1];
    diff2 = a[2] - b[2];
    diff3 = a[3] - b[3];
    result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
    a += 4;
    b += 4;
  }
  while (a < last) {
    diff0 = *a++ - *b++;
    result += diff0 * diff0;
  }
  return result;
}

// L2<unsigned char>: Accumulator<unsigned char>::Type = float; the difference is formed in int
// (integral promotion) and converted to float.
float l2_u8(const uint8_t* a, const uint8_t* b, size_t size) {
  float result = 0.0f;
  float diff0, diff1, diff2, diff3;
  const uint8_t* last = a + size;
  const uint8_t* lastgroup = last - 3;
  while (a < lastgroup) {
    diff0 = (float)(a[0] - b[0]);
    diff1 = (float)(a[1] - b[1]);
    diff2 = (float)(a[2] - b[2]);
    diff3 = (float)(a[3] - b[3]);
    result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
    a += 4;
    b += 4;
  }
  while (a < last) {
    diff0 = (float)(*a++ - *b++);
    result += diff0 * diff0;
  }
  return result;
}

namespace {
struct Packet {  // upstream sort_index_helper "packet": ordered by value only
  float val;
  int index;
  bool operator<(const Packet& r) const { return val < r.val; }
};
}  // namespace

// ---- A.2  ArrayMatcherBruteForce<Scalar, L2>::SearchNeighbours(query, nbQuery, idx, dist, NN) ----
// Result layout (contract visible in the reference's plug-ins, src/utils/matcher_hnsw.h:133-191):
// entry q*NN+k = IndMatch(i_=q, j_=dbIndex_k), ascending squared distance.
bool search_neighbours(const void* db, uint32_t n_db, const void* q, uint32_t nq, uint32_t dim,
                       int dtype, int NN, int32_t* idx, float* dist, int n_threads) {
  if (db == nullptr || (uint32_t)NN > n_db || nq < 1) return false;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel num_threads(n_threads)
  {
    std::vector<float> d(n_db);
    std::vector<Packet> packets(n_db);
#pragma omp for schedule(static)
    for (int64_t qi = 0; qi < (int64_t)nq; ++qi) {
      if (dtype == 0) {
        const float* qr = (const float*)q + (size_t)qi * dim;
        const float* dbf = (const float*)db;
        for (uint32_t i = 0; i < n_db; ++i) d[i] = l2_f32(qr, dbf + (size_t)i * dim, dim);
      } else {
        const uint8_t* qr = (const uint8_t*)q + (size_t)qi * dim;
        const uint8_t* dbu = (const uint8_t*)db;
        for (uint32_t i = 0; i < n_db; ++i) d[i] = l2_u8(qr, dbu + (size_t)i * dim, dim);
      }
      const int maxMinFound = std::min<int>(NN, (int)n_db);
      for (uint32_t i = 0; i < n_db; ++i) { packets[i].val = d[i]; packets[i].index = (int)i; }
      std::partial_sort(packets.begin(), packets.begin() + maxMinFound, packets.end());
      for (int k = 0; k < maxMinFound; ++k) {
        idx[qi * NN + k] = packets[k].index;
        dist[qi * NN + k] = packets[k].val;
      }
    }
  }
  return true;
}

// ---- A.3  IndMatchDecorator<float> (indMatchDecoratorXY.hpp) ----
// The upstream comparator is NOT a strict weak ordering; the outcome therefore depends on the
// std::set range-constructor's probing order.  Kept literally (including that quirk).
namespace {
struct DecoratedMatch {
  float x1, y1, x2, y2;
  orc_indmatch index;
  friend bool operator==(const DecoratedMatch& m1, const DecoratedMatch& m2) {
    return (m1.x1 == m2.x1 && m1.y1 == m2.y1 && m1.x2 == m2.x2 && m1.y2 == m2.y2);
  }
  friend bool operator<(const DecoratedMatch& m1, const DecoratedMatch& m2) {
    if (m1 == m2) return false;
    if (m1.x1 < m2.x1)
      return m1.y1 < m2.y1;
    else if (m1.x1 > m2.x1)
      return m1.y1 < m2.y1;
    return m1.x1 < m2.x1;
  }
};
struct IndMatchLess {
  bool operator()(const orc_indmatch& a, const orc_indmatch& b) const {
    return (a.i < b.i) || (a.i == b.i && a.j < b.j);
  }
};
}  // namespace

void coord_dedup(std::vector<orc_indmatch>& m, const float* xyI, const float* xyJ) {
  std::vector<DecoratedMatch> dec;
  dec.reserve(m.size());
  for (const auto& im : m) {
    DecoratedMatch d;
    d.x1 = xyI[2 * (size_t)im.i];
    d.y1 = xyI[2 * (size_t)im.i + 1];
    d.x2 = xyJ[2 * (size_t)im.j];
    d.y2 = xyJ[2 * (size_t)im.j + 1];
    d.index = im;
    dec.push_back(d);
  }
  std::set<DecoratedMatch> dedup(dec.begin(), dec.end());
  dec.assign(dedup.begin(), dedup.end());
  m.resize(dec.size());
  for (size_t i = 0; i < dec.size(); ++i) m[i] = dec[i].index;
}

// ---- A.2  RegionsMatcherT<MatcherT>::MatchDistanceRatio (called at R3DComputeMatches.cpp:479) ----
void match_distance_ratio(const void* descI, const float* xyI, uint32_t nI, const void* descJ,
                          const float* xyJ, uint32_t nJ, uint32_t dim, int dtype, float f_dist_ratio,
                          std::vector<orc_indmatch>& out, int n_threads) {
  out.clear();
  if (nI == 0 || nJ == 0) return;  // regions_.RegionCount()==0 / query empty
  const int NN = 2;
  std::vector<int32_t> idx((size_t)nJ * NN);
  std::vector<float> dist((size_t)nJ * NN);
  if (!search_neighbours(descI, nI, descJ, nJ, dim, dtype, NN, idx.data(), dist.data(), n_threads))
    return;
  // NNdistanceRatio(first, last, NN, ok, b_squared_metric ? Square(ratio) : ratio); BRUTE_FORCE_L2
  // passes b_squared_metric = true.
  const float fratio = f_dist_ratio * f_dist_ratio;
  std::vector<orc_indmatch> v;
  for (uint32_t q = 0; q < nJ; ++q) {
    if (dist[(size_t)q * NN] < fratio * dist[(size_t)q * NN + 1])
      v.push_back(orc_indmatch{(uint32_t)idx[(size_t)q * NN], q});  // (i_ = index in I, j_ = index in J)
  }
  // IndMatch::getDeduplicated
  std::set<orc_indmatch, IndMatchLess> s(v.begin(), v.end());
  v.assign(s.begin(), s.end());
  // IndMatchDecorator<float>(matches, posI, posJ).getDeduplicated
  coord_dedup(v, xyI, xyJ);
  out.swap(v);
}

// ---- a4  Matcher_Regions::Match (pair loop shape: src/R3DComputeMatches.cpp:437-488) ----
void match_pairs(const void* const* descs, const float* const* xys, const uint32_t* ns,
                 uint32_t n_views, uint32_t dim, int dtype, const uint32_t* pairs, uint64_t P,
                 float ratio, std::map<std::pair<uint32_t, uint32_t>, std::vector<orc_indmatch>>& out,
                 int n_threads) {
  (void)n_views;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  std::map<uint32_t, std::vector<uint32_t>> map_pairs;
  for (uint64_t p = 0; p < P; ++p) map_pairs[pairs[2 * p]].push_back(pairs[2 * p + 1]);
  for (const auto& kv : map_pairs) {
    const uint32_t I = kv.first;
    const auto& index_to_compare = kv.second;
    if (ns[I] == 0) continue;
#pragma omp parallel for schedule(dynamic) num_threads(n_threads)
    for (int j = 0; j < (int)index_to_compare.size(); ++j) {
      const uint32_t J = index_to_compare[j];
      if (ns[J] == 0) continue;
      std::vector<orc_indmatch> v;
      // the per-query loop inside SearchNeighbours is itself an omp-for upstream; nested
      // parallelism is off by default, so it runs single-threaded inside this team.
      match_distance_ratio(descs[I], xys[I], ns[I], descs[J], xys[J], ns[J], dim, dtype, ratio, v, 1);
#pragma omp critical
      {
        if (!v.empty()) out.insert({{I, J}, std::move(v)});
      }
    }
  }
}

}  // namespace orc

// ------------------------------------------------------------------------------------------------
extern "C" {

float orc_l2_f32(const float* a, const float* b, uint64_t n) { return orc::l2_f32(a, b, n); }
float orc_l2_u8(const uint8_t* a, const uint8_t* b, uint64_t n) { return orc::l2_u8(a, b, n); }

int orc_search_neighbours(const void* db, uint32_t n_db, const void* q, uint32_t nq, uint32_t dim,
                          int dtype, int32_t* idx, float* dist, int n_threads) {
  return orc::search_neighbours(db, n_db, q, nq, dim, dtype, 2, idx, dist, n_threads) ? 0 : 1;
}

int64_t orc_match_distance_ratio(const void* descI, const float* xyI, uint32_t nI, const void* descJ,
                                 const float* xyJ, uint32_t nJ, uint32_t dim, int dtype, float ratio,
                                 orc_indmatch* out, int n_threads) {
  std::vector<orc_indmatch> v;
  orc::match_distance_ratio(descI, xyI, nI, descJ, xyJ, nJ, dim, dtype, ratio, v, n_threads);
  std::memcpy(out, v.data(), v.size() * sizeof(orc_indmatch));
  return (int64_t)v.size();
}

int64_t orc_match_pairs(const void* const* descs, const float* const* xys, const uint32_t* ns,
                        uint32_t n_views, uint32_t dim, int dtype, const uint32_t* pairs, uint64_t P,
                        float ratio, uint64_t* pair_ofs, orc_indmatch* out, uint64_t cap,
                        int n_threads) {
  std::map<std::pair<uint32_t, uint32_t>, std::vector<orc_indmatch>> res;
  orc::match_pairs(descs, xys, ns, n_views, dim, dtype, pairs, P, ratio, res, n_threads);
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

int64_t orc_coord_dedup(orc_indmatch* m, int64_t n, const float* xyI, const float* xyJ) {
  std::vector<orc_indmatch> v(m, m + n);
  orc::coord_dedup(v, xyI, xyJ);
  std::memcpy(m, v.data(), v.size() * sizeof(orc_indmatch));
  return (int64_t)v.size();
}

int orc_num_threads(void) { return omp_get_max_threads(); }
}
