// ArrayMatcher_b200.h -- the finest-grained drop-in: openMVG::matching::ArrayMatcher<Scalar, Metric> on libr3dgpu.
//
// Regard3D's own NN plug-ins implement this interface (src/utils/matcher_hnsw.h:53-191, matcher_kgraph.h:116-251,
// matcher_mrpt.h:77-251): Build(dataset, nbRows, dimension) borrows a row-major array, SearchNeighbours(query,
// nbQuery, &indices, &distances, NN) fills nbQuery * NN entries, entry q * NN + k = IndMatch(q, dbIndex_k), ascending
// SQUARED L2 distance (the metric OpenMVG's L2 functor returns).  This adaptor forwards to r3d_upload_regions /
// r3d_search_neighbours (exact brute force on the B200, bit-identical to ArrayMatcherBruteForce); it is per pair and
// therefore far below the batched r3d_match_pairs path in throughput -- a compatibility shim, not the fast path.
//
// Header only and free of OpenMVG includes: instantiate it with the OpenMVG base and IndMatch types
//     using GpuMatcher = r3d_shim::ArrayMatcher_b200<float, openMVG::matching::L2<float>,
//                                                    openMVG::matching::ArrayMatcher<float, openMVG::matching::L2<float>>,
//                                                    openMVG::matching::IndMatch>;
// (tests/test_cpp_adaptors.py compiles it against a stand-in of those two types.)
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/r3dgpu.h"

namespace r3d_shim {

template <typename Scalar, typename Metric, typename Base, typename IndMatchT>
class ArrayMatcher_b200 : public Base {
 public:
  using DistanceType = typename Metric::ResultType;
  ArrayMatcher_b200() {
    if (r3d_create(nullptr, 0, &ctx_) != R3D_OK) {  // no B200: every call below returns false (no CPU fallback)
      ctx_ = nullptr;
      const char* e = r3d_last_error(nullptr);
      error_ = e ? e : "r3d_create failed";
    }
  }
  ~ArrayMatcher_b200() override { r3d_destroy(ctx_); }
  ArrayMatcher_b200(const ArrayMatcher_b200&) = delete;
  ArrayMatcher_b200& operator=(const ArrayMatcher_b200&) = delete;

  // matcher_hnsw.h:53-68 -- the data is copied to the device here (the reference's plug-ins borrow the pointer)
  bool Build(const Scalar* dataset, int nbRows, int dimension) override {
    if (!ctx_ || nbRows < 1 || dimension < 1 || !dataset) return false;
    dimension_ = dimension;
    nbRows_ = nbRows;
    return r3d_upload_regions(ctx_, kDbView, dataset, (uint32_t)nbRows, (uint32_t)dimension, dtype(), nullptr) == R3D_OK;
  }

  // single query, nearest neighbour (matcher_hnsw.h:81-120)
  bool SearchNeighbour(const Scalar* query, int* indice, DistanceType* distance) override {
    if (!indice || !distance) return false;
    std::vector<IndMatchT> idx;
    std::vector<DistanceType> dist;
    if (!SearchNeighbours(query, 1, &idx, &dist, 1)) return false;
    *indice = (int)idx[0].j_;
    *distance = dist[0];
    return true;
  }

  // matcher_hnsw.h:133-191.  NN <= 2: the library returns the exact two nearest rows.
  bool SearchNeighbours(const Scalar* query, int nbQuery, std::vector<IndMatchT>* pvec_indices,
                        std::vector<DistanceType>* pvec_distances, size_t NN) override {
    if (!ctx_ || !query || nbQuery < 1 || !pvec_indices || !pvec_distances) return false;
    if (NN < 1 || NN > 2 || (size_t)nbRows_ < NN) return false;
    if (nbRows_ < 2) return false;  // the library's 2-NN contract needs two database rows
    if (r3d_upload_regions(ctx_, kQueryView, query, (uint32_t)nbQuery, (uint32_t)dimension_, dtype(), nullptr) != R3D_OK)
      return false;
    std::vector<int32_t> idx(2 * (size_t)nbQuery);
    std::vector<float> dist(2 * (size_t)nbQuery);
    if (r3d_search_neighbours(ctx_, kDbView, kQueryView, idx.data(), dist.data()) != R3D_OK) return false;
    pvec_indices->reserve(pvec_indices->size() + (size_t)nbQuery * NN);
    pvec_distances->reserve(pvec_distances->size() + (size_t)nbQuery * NN);
    for (int q = 0; q < nbQuery; ++q)
      for (size_t k = 0; k < NN; ++k) {
        pvec_indices->emplace_back((uint32_t)q, (uint32_t)idx[2 * (size_t)q + k]);
        pvec_distances->push_back((DistanceType)dist[2 * (size_t)q + k]);
      }
    return true;
  }

  const std::string& lastError() const { return error_; }

 private:
  static int dtype() { return sizeof(Scalar) == 1 ? R3D_U8 : R3D_F32; }
  static constexpr uint32_t kDbView = 0, kQueryView = 1;
  r3d_ctx* ctx_ = nullptr;
  int dimension_ = 0, nbRows_ = 0;
  std::string error_;
};

}  // namespace r3d_shim
