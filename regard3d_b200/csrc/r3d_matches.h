// r3d_matches.h -- PairWiseMatches container shared by the host translation units.
#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/r3dgpu.h"

struct r3d_matches {
  std::vector<uint32_t> pairs;   // 2 per pair, sorted by (I,J) -- std::map<Pair, IndMatches> order
  std::vector<uint64_t> ofs;     // n_pairs + 1
  std::vector<r3d_indmatch> m;
};
