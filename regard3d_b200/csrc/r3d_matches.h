// r3d_matches.h -- PairWiseMatches container shared by the host translation units.
#pragma once
#include <stdint.h>
#include <vector>
#include "../../include/r3dgpu.h"

// std::map<Pair, IndMatches> flattened: pairs sorted by (I,J), one IndMatches vector per pair.
// The vectors are the ones the batch tails produced: assembling a result moves them, it never copies matches.
struct r3d_matches {
  std::vector<uint32_t> pairs;                       // 2 per pair
  std::vector<std::vector<r3d_indmatch>> per;        // per[k] = matches of pair k
  uint64_t total = 0;
  void push(uint32_t I, uint32_t J, std::vector<r3d_indmatch>&& v) {
    pairs.push_back(I);
    pairs.push_back(J);
    total += v.size();
    per.push_back(std::move(v));
  }
};
