// match_post.cpp -- host-side, order-dependent tail of RegionsMatcherT::MatchDistanceRatio
// (upstream matching/regions_matcher.hpp, invoked by the reference at src/R3DComputeMatches.cpp:479):
//   IndMatch::getDeduplicated            -> sort by (i_,j_) + unique
//   IndMatchDecorator<float>::getDeduplicated -> std::set with the upstream comparator
// The second step's comparator is not a strict weak ordering, so its result is defined only by
// the std::set range-insertion algorithm; it therefore runs on the host, as an exact replay of that algorithm
// (libstdc++'s, the reference's toolchain).  O(#matches log #matches) per pair.
#include <algorithm>
#include <cstdint>
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

// ---- the std::set range insertion, replayed on a compact node array ------------------------------------
// std::set<XYMatch, XYLess>(first, last) in libstdc++ is, element by element,
//     _M_insert_unique_(end(), v):  hint == end():  size() > 0 && less(key(rightmost), v)  -> append right of rightmost
//                                   else _M_get_insert_unique_pos(v): descend (left iff less(v, node)), then with
//                                   j = the in-order predecessor of the landing slot (none if it is the leftmost
//                                   slot): insert iff no predecessor or less(key(j), v)
//     _M_insert_(x, p, v):          insert_left = (p == header || less(v, key(p)))   [x is always null here]
//     _Rb_tree_insert_and_rebalance (tree.cc): the textbook red-black fix-up with libstdc++'s rotation cases.
// Because XYLess is not a strict weak ordering the RESULT depends on exactly this procedure (tree shape included),
// so it is replayed step by step -- only the container changes: 32-bit links into one contiguous node array
// (40-byte nodes, no allocator, no virtual calls) instead of 56-byte heap nodes.  tests/test_io_and_abi.py pins the
// replay against std::set (through the oracle) on adversarial inputs (many equal x1 / y1 / full duplicates).
namespace {
// XYLess(a, b) reduces to (a.x1 != b.x1) && (a.y1 < b.y1): with equal x1 it returns a.x1 < b.x1 = false whether or
// not the other coordinates agree, with different x1 the "same coordinates" test is false and both branches return
// a.y1 < b.y1.  So only (x1, y1) of image I take part, and the test is branch-free.
// Nodes are 16 bytes with 16-bit links whenever a pair has fewer than 65 535 matches (always, in practice): the
// whole tree of a 3 000-match pair then sits in the L1 data cache; the IndMatch payload lives in a side array.
template <typename Idx>
struct RbNodeT {
  float x1, y1;
  Idx parent, child[2];  // child[0] = left, child[1] = right
  Idx red;
};

template <typename Idx>
struct RbTreeT {
  typedef RbNodeT<Idx> RbNode;
  static constexpr Idx kNil = (Idx)~(Idx)0;
  std::vector<RbNode> n;  // n[0] is the header: parent = root, left = leftmost, right = rightmost
  std::vector<r3d_indmatch> payload;  // payload[k] belongs to node k
  static bool less(const RbNode& a, const RbNode& b) { return (a.x1 != b.x1) & (a.y1 < b.y1); }
  void reset(size_t cap) {
    n.clear();
    n.reserve(cap + 1);
    payload.clear();
    payload.reserve(cap + 1);
    payload.push_back(r3d_indmatch{0, 0});
    RbNode h{};
    h.parent = kNil; h.child[0] = 0; h.child[1] = 0; h.red = 1;
    n.push_back(h);
  }
  Idx& root() { return n[0].parent; }
  void rotate_left(Idx x) {
    const Idx y = n[x].child[1];
    n[x].child[1] = n[y].child[0];
    if (n[y].child[0] != kNil) n[n[y].child[0]].parent = x;
    n[y].parent = n[x].parent;
    if (x == root()) root() = y;
    else if (x == n[n[x].parent].child[0]) n[n[x].parent].child[0] = y;
    else n[n[x].parent].child[1] = y;
    n[y].child[0] = x;
    n[x].parent = y;
  }
  void rotate_right(Idx x) {
    const Idx y = n[x].child[0];
    n[x].child[0] = n[y].child[1];
    if (n[y].child[1] != kNil) n[n[y].child[1]].parent = x;
    n[y].parent = n[x].parent;
    if (x == root()) root() = y;
    else if (x == n[n[x].parent].child[1]) n[n[x].parent].child[1] = y;
    else n[n[x].parent].child[0] = y;
    n[y].child[1] = x;
    n[x].parent = y;
  }
  Idx decrement(Idx x) const {  // _Rb_tree_decrement for a non-header node that is not the leftmost
    if (n[x].child[0] != kNil) {
      Idx y = n[x].child[0];
      while (n[y].child[1] != kNil) y = n[y].child[1];
      return y;
    }
    Idx y = n[x].parent;
    while (x == n[y].child[0]) { x = y; y = n[y].parent; }
    return y;
  }
  void insert_and_rebalance(bool insert_left, Idx x, Idx p) {
    n[x].parent = p; n[x].child[0] = kNil; n[x].child[1] = kNil; n[x].red = 1;
    if (insert_left) {
      n[p].child[0] = x;  // also makes leftmost = x when p is the header
      if (p == 0) { n[0].parent = x; n[0].child[1] = x; }
      else if (p == n[0].child[0]) n[0].child[0] = x;
    } else {
      n[p].child[1] = x;
      if (p == n[0].child[1]) n[0].child[1] = x;
    }
    while (x != root() && n[n[x].parent].red) {
      const Idx xp = n[x].parent, xpp = n[xp].parent;
      if (xp == n[xpp].child[0]) {
        const Idx y = n[xpp].child[1];
        if (y != kNil && n[y].red) {
          n[xp].red = 0; n[y].red = 0; n[xpp].red = 1;
          x = xpp;
        } else {
          if (x == n[xp].child[1]) { x = xp; rotate_left(x); }
          n[n[x].parent].red = 0;
          n[xpp].red = 1;
          rotate_right(xpp);
        }
      } else {
        const Idx y = n[xpp].child[0];
        if (y != kNil && n[y].red) {
          n[xp].red = 0; n[y].red = 0; n[xpp].red = 1;
          x = xpp;
        } else {
          if (x == n[xp].child[0]) { x = xp; rotate_right(x); }
          n[n[x].parent].red = 0;
          n[xpp].red = 1;
          rotate_left(xpp);
        }
      }
    }
    n[root()].red = 0;
  }
  // _M_insert_unique_(end(), v)
  void insert_unique_hint_end(const RbNode& v, const r3d_indmatch& im) {
    const Idx count = (Idx)(n.size() - 1);
    Idx p;
    if (count > 0 && less(n[n[0].child[1]], v)) {
      p = n[0].child[1];  // {0, rightmost}
    } else {           // _M_get_insert_unique_pos
      Idx x = count ? root() : kNil, y = 0;
      bool comp = true;
      while (x != kNil) {
        y = x;
        comp = less(v, n[x]);
        x = n[x].child[comp ? 0 : 1];  // branch-free descent
      }
      Idx j = y;
      bool check = true;
      if (comp) {
        if (y == n[0].child[0] || count == 0) check = false;  // j == begin(): insert
        else j = decrement(y);
      }
      if (check && !less(n[j], v)) return;  // an "equivalent" key is already there
      p = y;
    }
    const bool insert_left = (p == 0) || less(v, n[p]);
    n.push_back(v);
    payload.push_back(im);
    insert_and_rebalance(insert_left, (Idx)(n.size() - 1), p);
  }
  template <typename F>
  void in_order(F&& f) const {
    if (n.size() <= 1) return;
    Idx x = n[0].child[0];  // leftmost
    for (;;) {
      f(payload[x]);
      if (n[x].child[1] != kNil) {  // _Rb_tree_increment
        x = n[x].child[1];
        while (n[x].child[0] != kNil) x = n[x].child[0];
      } else {
        Idx y = n[x].parent;
        while (y != 0 && x == n[y].child[1]) { x = y; y = n[y].parent; }
        if (y == 0) {
          // came up from the root's right spine (or x is the root without a right child): done
          // (libstdc++ uses the header trick "if (x->right != y) x = y"; an explicit end test is equivalent here)
          return;
        }
        x = y;
      }
    }
  }
};
}  // namespace

template <typename Idx>
static size_t coord_dedup_replay(r3d_indmatch* m, size_t n, const float* xyI) {
  thread_local RbTreeT<Idx> tree;  // arrays reused across pairs: no allocation per pair
  tree.reset(n);
  for (size_t k = 0; k < n; ++k) {
    typename RbTreeT<Idx>::RbNode v{};
    v.x1 = xyI[2 * (size_t)m[k].i];
    v.y1 = xyI[2 * (size_t)m[k].i + 1];
    tree.insert_unique_hint_end(v, m[k]);
  }
  size_t out = 0;
  tree.in_order([&](const r3d_indmatch& im) { m[out++] = im; });
  return out;
}

// ascending (i, j): LSD radix sort on the 64-bit key i << 32 | j, 8 bits per pass, passes whose digit is the same in
// every key are skipped (indices rarely need more than 2 bytes each)
static void sort_ij(r3d_indmatch* m, size_t n) {
  if (n < 64) {
    std::sort(m, m + n, [](const r3d_indmatch& a, const r3d_indmatch& b) { return a.i < b.i || (a.i == b.i && a.j < b.j); });
    return;
  }
  thread_local std::vector<uint64_t> ka, kb;
  ka.resize(n);
  kb.resize(n);
  uint64_t all_or = 0, all_and = ~0ull;
  for (size_t k = 0; k < n; ++k) {
    const uint64_t key = ((uint64_t)m[k].i << 32) | m[k].j;
    ka[k] = key;
    all_or |= key;
    all_and &= key;
  }
  uint64_t* src = ka.data();
  uint64_t* dst = kb.data();
  for (int pass = 0; pass < 8; ++pass) {
    const int sh = 8 * pass;
    if ((((all_or ^ all_and) >> sh) & 0xffu) == 0) continue;  // this byte is identical in every key
    size_t hist[257] = {0};
    for (size_t k = 0; k < n; ++k) hist[((src[k] >> sh) & 0xffu) + 1]++;
    for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
    for (size_t k = 0; k < n; ++k) dst[hist[(src[k] >> sh) & 0xffu]++] = src[k];
    std::swap(src, dst);
  }
  for (size_t k = 0; k < n; ++k) { m[k].i = (uint32_t)(src[k] >> 32); m[k].j = (uint32_t)src[k]; }
}

// in place on m[0..n); returns the new count
size_t post_process_pair(r3d_indmatch* m, size_t n, const float* xyI, const float* xyJ, bool coord_dedup) {
  sort_ij(m, n);
  n = (size_t)(std::unique(m, m + n, [](const r3d_indmatch& a, const r3d_indmatch& b) { return a.i == b.i && a.j == b.j; }) - m);
  if (!coord_dedup || !xyI || !xyJ) return n;
  if (n < 65000) return coord_dedup_replay<uint16_t>(m, n, xyI);
  return coord_dedup_replay<uint32_t>(m, n, xyI);
}

}  // namespace r3d

// host-only diagnostic (no GPU needed): the post-processing of one pair, for the CPU test that pins the tree replay
extern "C" int64_t r3d_debug_post_process(r3d_indmatch* m, int64_t n, const float* xyI, const float* xyJ, int coord_dedup) {
  if (!m || n < 0) return -1;
  return (int64_t)r3d::post_process_pair(m, (size_t)n, xyI, xyJ, coord_dedup != 0);
}
