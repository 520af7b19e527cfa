// oracle_internal.hpp -- CPU ORACLE internals (test infrastructure; see oracle.h).
#pragma once
#include "oracle.h"
#include <cstddef>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

namespace orc {
float l2_f32(const float* a, const float* b, size_t size);
float l2_u8(const uint8_t* a, const uint8_t* b, size_t size);
bool search_neighbours(const void* db, uint32_t n_db, const void* q, uint32_t nq, uint32_t dim,
                       int dtype, int NN, int32_t* idx, float* dist, int n_threads);
void coord_dedup(std::vector<orc_indmatch>& m, const float* xyI, const float* xyJ);
void match_distance_ratio(const void* descI, const float* xyI, uint32_t nI, const void* descJ,
                          const float* xyJ, uint32_t nJ, uint32_t dim, int dtype, float ratio,
                          std::vector<orc_indmatch>& out, int n_threads);
void match_pairs(const void* const* descs, const float* const* xys, const uint32_t* ns,
                 uint32_t n_views, uint32_t dim, int dtype, const uint32_t* pairs, uint64_t P,
                 float ratio, std::map<std::pair<uint32_t, uint32_t>, std::vector<orc_indmatch>>& out,
                 int n_threads);
}  // namespace orc
