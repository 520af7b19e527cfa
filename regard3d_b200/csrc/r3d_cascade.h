// r3d_cascade.h -- cascade-hashing matcher (cascade.cu), internal interface used by match_host.cu / context.cu.
#pragma once
#include "r3d_internal.cuh"

#include <vector>

namespace r3d {
// hash every view of `used` (ascending ids) under the zero-mean descriptor of exactly that set
int cascade_prepare(r3d_ctx* ctx, DeviceWorker& w, const std::vector<uint32_t>& used);
// true when every non-empty view of the pair list carries tables of the worker's current epoch
bool cascade_ready(const DeviceWorker& w, const uint32_t* pairs, uint64_t n_pairs);
void cascade_release_view(DeviceWorker& w, ViewDev& v);
// d_cidx[k] = (row of view I, row of view J) in the worker's CascadeView table
int launch_cascade_match(r3d_ctx* ctx, DeviceWorker& w, const PairDesc* d_pairs, const uint2* d_cidx, uint32_t n_pairs, uint32_t max_nJ,
                         uint32_t max_nI, uint32_t dim, int dtype, float ratio2, uint32_t* d_counters, uint2* d_matches);
}  // namespace r3d
