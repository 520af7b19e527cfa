// match_post.cpp -- host-side, order-dependent tail of RegionsMatcherT::MatchDistanceRatio
// (upstream matching/regions_matcher.hpp, invoked by the reference at src/R3DComputeMatches.cpp:479):
//   IndMatch::getDeduplicated            -> sort by (i_,j_) + unique
//   IndMatchDecorator<float>::getDeduplicated -> std::set with the upstream comparator
// The second step's comparator is not a strict weak ordering, so its result is defined only by
// the std::set range-insertion algorithm; it therefore runs on the host with the very container
// the reference uses.  The set's nodes come from a per-thread monotonic arena
// (std::pmr): the allocator does not take part in the tree algorithm, so the result is the one of
// std::set<.., std::allocator>, without a malloc/free per match.  O(#matches log #matches) per pair.
#include <algorithm>
#include <memory_resource>
#include <set>
#include <vector>

#include "../../include/r3dgpu.h"

namespace r3d {

namespace {
struct XYMatch {
  float x1, y1, x2, y2;
  r3d_indmatch im;
};
inline bool same_xy(const XYMatch& a, const XYMatch& b) {
  return a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2;
}
// upstream IndMatchDecoratorStruct::operator< ("lexicographical ordering", verbatim semantics)
struct XYLess {
  bool operator()(const XYMatch& m1, const XYMatch& m2) const {
    if (same_xy(m1, m2)) return false;
    if (m1.x1 < m2.x1) return m1.y1 < m2.y1;
    if (m1.x1 > m2.x1) return m1.y1 < m2.y1;
    return m1.x1 < m2.x1;
  }
};
}  // namespace

// in place on m[0..n); returns the new count
size_t post_process_pair(r3d_indmatch* m, size_t n, const float* xyI, const float* xyJ, bool coord_dedup) {
  std::sort(m, m + n, [](const r3d_indmatch& a, const r3d_indmatch& b) {
    return a.i < b.i || (a.i == b.i && a.j < b.j);
  });
  n = (size_t)(std::unique(m, m + n, [](const r3d_indmatch& a, const r3d_indmatch& b) { return a.i == b.i && a.j == b.j; }) - m);
  if (!coord_dedup || !xyI || !xyJ) return n;
  thread_local std::vector<XYMatch> dec;  // scratch reused across pairs: no allocation per pair
  dec.resize(n);
  for (size_t k = 0; k < n; ++k) {
    dec[k].x1 = xyI[2 * (size_t)m[k].i];
    dec[k].y1 = xyI[2 * (size_t)m[k].i + 1];
    dec[k].x2 = xyJ[2 * (size_t)m[k].j];
    dec[k].y2 = xyJ[2 * (size_t)m[k].j + 1];
    dec[k].im = m[k];
  }
  thread_local std::vector<unsigned char> arena_buf;
  const size_t need = dec.size() * (sizeof(XYMatch) + 48) + 1024;  // red-black node = 32-byte header + payload
  if (arena_buf.size() < need) arena_buf.resize(need + need / 2);
  std::pmr::monotonic_buffer_resource arena(arena_buf.data(), arena_buf.size());
  {
    std::pmr::set<XYMatch, XYLess> uniq(dec.begin(), dec.end(), XYLess(), &arena);
    n = 0;
    for (const auto& d : uniq) m[n++] = d.im;
  }
  return n;
}

}  // namespace r3d
