// r3d_matches.h -- PairWiseMatches container shared by the host translation units.
#pragma once
#include <stdint.h>
#include <memory>
#include <vector>
#include "../../include/r3dgpu.h"

// std::map<Pair, IndMatches> flattened: pairs sorted by (I,J); the IndMatches of pair k are the span per[k] inside
// one of the slabs.  A matching batch hands over ONE slab (the buffer its matches were copied into from the device,
// de-duplicated in place), so assembling a result neither copies matches nor allocates per pair.
struct r3d_span {
  const r3d_indmatch* p = nullptr;
  size_t n = 0;
  const r3d_indmatch* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  const r3d_indmatch& operator[](size_t k) const { return p[k]; }
  const r3d_indmatch* begin() const { return p; }
  const r3d_indmatch* end() const { return p + n; }
};
typedef std::shared_ptr<void> r3d_slab;  // type-erased owner of the storage some spans point into

struct r3d_matches {
  std::vector<uint32_t> pairs;  // 2 per pair
  std::vector<r3d_span> per;    // per[k] = matches of pair k
  std::vector<r3d_slab> slabs;  // storage the spans point into
  uint64_t total = 0;
  void push(uint32_t I, uint32_t J, std::vector<r3d_indmatch>&& v) {  // own slab for this pair
    auto s = std::make_shared<std::vector<r3d_indmatch>>(std::move(v));
    slabs.push_back(s);
    push_span(I, J, r3d_span{s->data(), s->size()});
  }
  void push_span(uint32_t I, uint32_t J, r3d_span sp) {  // the caller has added the slab
    pairs.push_back(I);
    pairs.push_back(J);
    total += sp.n;
    per.push_back(sp);
  }
};
