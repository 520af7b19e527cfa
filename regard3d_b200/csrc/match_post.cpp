// match_post.cpp -- host-side, order-dependent tail of RegionsMatcherT::MatchDistanceRatio
// (upstream matching/regions_matcher.hpp, invoked by the reference at src/R3DComputeMatches.cpp:479):
//   IndMatch::getDeduplicated            -> sort by (i_,j_) + unique
//   IndMatchDecorator<float>::getDeduplicated -> std::set with the upstream comparator
// The second step's comparator is not a strict weak ordering, so its result is defined only by
// the std::set range-insertion algorithm; it therefore runs on the host, as an exact replay of that algorithm
// (libstdc++'s, the reference's toolchain).  O(#matches log #matches) per pair.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/r3dgpu.h"

namespace r3d {

constexpr int kPostLanes = 4;  // keep in sync with r3d_internal.cuh
struct ViewRankRef { const uint32_t* yrank; const uint8_t* xshared; uint32_t n_slots; };  // idem

// upstream IndMatchDecoratorStruct::operator< ("lexicographical ordering"), verbatim semantics, on records
// (x1, y1, x2, y2, IndMatch):
//     if (m1 == m2) return false;                       // all four coordinates equal
//     if (m1.x1 < m2.x1) return m1.y1 < m2.y1;
//     if (m1.x1 > m2.x1) return m1.y1 < m2.y1;
//     return m1.x1 < m2.x1;                             // equal x1: false
// i.e. less(a, b) == (a.x1 < b.x1 || a.x1 > b.x1) && (a.y1 < b.y1): only the position in image I takes part
// (NOT `x1 != x1'`: an unordered (NaN) x1 fails both branches and lands on `return m1.x1 < m2.x1` = false).

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
// XYLess(a, b) reduces to (a.x1 <> b.x1) && (a.y1 < b.y1): with equal x1 it returns a.x1 < b.x1 = false whether or
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
  // (x1 < | x1 >) rather than x1 != : with a NaN x1 the upstream comparator falls through to `m1.x1 < m2.x1` = false
  static bool less(const RbNode& a, const RbNode& b) { return ((a.x1 < b.x1) | (a.x1 > b.x1)) & (a.y1 < b.y1); }
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
  // _M_insert_unique_(end(), v), cut into begin / one descent level / finish so that several independent trees can
  // be advanced in lockstep by one thread (the descent is a chain of dependent loads; interleaving hides its latency)
  struct Cursor { Idx x, y; bool comp; };
  // returns true when a descent is needed; false: the hint path applied and the element has been appended
  bool begin_insert(const RbNode& v, const r3d_indmatch& im, Cursor& c) {
    const Idx count = (Idx)(n.size() - 1);
    if (count > 0 && less(n[n[0].child[1]], v)) {  // {0, rightmost}
      const Idx p = n[0].child[1];
      link(v, im, p);
      return false;
    }
    c.x = count ? root() : kNil;
    c.y = 0;
    c.comp = true;
    return true;
  }
  bool descending(const Cursor& c) const { return c.x != kNil; }
  void step(const RbNode& v, Cursor& c) const {  // one level of _M_get_insert_unique_pos
    c.y = c.x;
    c.comp = less(v, n[c.x]);
    c.x = n[c.x].child[c.comp ? 0 : 1];
  }
  void finish_insert(const RbNode& v, const r3d_indmatch& im, const Cursor& c) {
    const Idx count = (Idx)(n.size() - 1);
    Idx j = c.y;
    bool check = true;
    if (c.comp) {
      if (c.y == n[0].child[0] || count == 0) check = false;  // j == begin(): insert
      else j = decrement(c.y);
    }
    if (check && !less(n[j], v)) return;  // an "equivalent" key is already there
    link(v, im, c.y);
  }
  void link(const RbNode& v, const r3d_indmatch& im, Idx p) {  // _M_insert_(0, p, v)
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

// up to kLanes pairs advanced in lockstep by the calling thread; counts[t] is updated in place
constexpr int kLanes = kPostLanes;
template <typename Idx>
static void coord_dedup_replay(int lanes, r3d_indmatch* const* ms, size_t* counts, const float* const* xyIs) {
  typedef RbTreeT<Idx> Tree;
  thread_local Tree trees[kLanes];  // arrays reused across pairs: no allocation per pair
  typename Tree::Cursor cur[kLanes];
  typename Tree::RbNode val[kLanes];
  size_t pos[kLanes];
  bool busy[kLanes], done[kLanes];
  int active = 0;
  for (int t = 0; t < lanes; ++t) {
    trees[t].reset(counts[t]);
    pos[t] = 0;
    busy[t] = false;
    done[t] = counts[t] == 0;
    if (!done[t]) ++active;
  }
  while (active) {
    for (int t = 0; t < lanes; ++t) {
      if (done[t]) continue;
      Tree& tr = trees[t];
      const r3d_indmatch* m = ms[t];
      if (busy[t]) {
        if (tr.descending(cur[t])) { tr.step(val[t], cur[t]); continue; }
        tr.finish_insert(val[t], m[pos[t] - 1], cur[t]);
        busy[t] = false;
      }
      // next element of this pair; an element with the same i as its predecessor shares its coordinates and its fate is
      // "rejected" either way (predecessor accepted: an equal key is present; rejected: same descent on the same tree)
      size_t k = pos[t];
      const size_t n = counts[t];
      while (k < n && k > 0 && m[k].i == m[k - 1].i) ++k;
      if (k >= n) { done[t] = true; --active; continue; }
      val[t].x1 = xyIs[t][2 * (size_t)m[k].i];
      val[t].y1 = xyIs[t][2 * (size_t)m[k].i + 1];
      pos[t] = k + 1;
      busy[t] = tr.begin_insert(val[t], m[k], cur[t]);
    }
  }
  for (int t = 0; t < lanes; ++t) {
    size_t out = 0;
    r3d_indmatch* m = ms[t];
    trees[t].in_order([&](const r3d_indmatch& im) { m[out++] = im; });
    counts[t] = out;
  }
}

// ---- per-view tables for the O(1) landing-slot replay ---------------------------------------------------
// yrank[i]  : dense rank of keypoint i's y coordinate among the view's keypoints (equal y share a rank)
// xshared[i]: some OTHER keypoint of the view has the same x coordinate
static inline uint32_t float_key(float f) {  // monotonic float -> uint32 (f finite; -0 is folded into +0)
  f += 0.0f;
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static void sort_by_key(std::vector<uint64_t>& a, std::vector<uint64_t>& tmp) {  // LSD radix on the high 32 bits
  const size_t n = a.size();
  tmp.resize(n);
  uint64_t* src = a.data();
  uint64_t* dst = tmp.data();
  for (int pass = 0; pass < 4; ++pass) {
    const int sh = 32 + 8 * pass;
    size_t hist[257] = {0};
    for (size_t k = 0; k < n; ++k) hist[((src[k] >> sh) & 0xffu) + 1]++;
    for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
    for (size_t k = 0; k < n; ++k) dst[hist[(src[k] >> sh) & 0xffu]++] = src[k];
    std::swap(src, dst);
  }  // 4 passes: the result is back in a
}
void build_view_ranks(const float* xy, uint32_t n, std::vector<uint32_t>& yrank, std::vector<uint8_t>& xshared,
                      uint32_t* n_slots) {
  yrank.clear();
  xshared.clear();
  *n_slots = 0;
  // non-finite positions: the rank tables assume totally ordered coordinates; leave them empty and the tails fall back
  // to the classic replay, which evaluates the comparator exactly like std::set does (NaN compares false everywhere)
  for (size_t k = 0; k < 2 * (size_t)n; ++k)
    if (!(xy[k] - xy[k] == 0.0f)) return;
  yrank.assign(n, 0);
  xshared.assign(n, 0);
  std::vector<uint64_t> a(n), tmp;
  for (uint32_t i = 0; i < n; ++i) a[i] = ((uint64_t)float_key(xy[2 * (size_t)i + 1]) << 32) | i;
  sort_by_key(a, tmp);
  uint32_t rank = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t i = (uint32_t)a[k];
    if (k > 0 && xy[2 * (size_t)i + 1] != xy[2 * (size_t)(uint32_t)a[k - 1] + 1]) ++rank;
    yrank[i] = rank;
  }
  *n_slots = n ? rank + 1 : 0;
  for (uint32_t i = 0; i < n; ++i) a[i] = ((uint64_t)float_key(xy[2 * (size_t)i]) << 32) | i;
  sort_by_key(a, tmp);
  for (uint32_t k = 0; k + 1 < n; ++k) {
    const uint32_t i = (uint32_t)a[k], j = (uint32_t)a[k + 1];
    if (xy[2 * (size_t)i] == xy[2 * (size_t)j]) { xshared[i] = 1; xshared[j] = 1; }
  }
}

// The replay without the descents.  The set's in-order sequence is always strictly increasing in y (an element is
// only linked where its in-order predecessor has a smaller y -- see the insertion rule), so for an element v:
//   * a present element with the SAME y: the search path reaches it, turns right, and every later predecessor has
//     y >= v.y                                                                                   -> rejected;
//   * otherwise the descent is the plain BST descent by y EXCEPT at a node with the same x and a larger y, where
//     less(v, node) is false and the descent turns right (and v is then rejected).  If no other keypoint of the view
//     shares v's x (xshared == 0) that cannot happen: v lands in THE slot between its y-predecessor and y-successor
//     among the present elements -- the right child of the predecessor when that is free, else the left child of
//     the successor -- and is accepted (its predecessor has a different x).
// Predecessor / successor come from a bitset over the view's y ranks; elements with xshared (a few percent) take the
// classic descent on the very same tree.  Same tree, same rotations, same result -- minus ~12 dependent loads per match.
template <typename Idx>
static size_t ranked_replay(r3d_indmatch* m, size_t n, const float* xyI, const uint32_t* yrank, const uint8_t* xshared,
                            uint32_t n_slots) {
  typedef RbTreeT<Idx> Tree;
  thread_local Tree tree;
  thread_local std::vector<uint64_t> occ;
  thread_local std::vector<Idx> slot_node;
  tree.reset(n);
  const size_t words = ((size_t)n_slots + 63) / 64 + 1;
  occ.assign(words, 0ull);
  if (slot_node.size() < n_slots) slot_node.resize(n_slots);
  for (size_t k = 0; k < n; ++k) {
    if (k > 0 && m[k].i == m[k - 1].i) continue;  // same keypoint as its predecessor: rejected either way (see above)
    const uint32_t i = m[k].i, s = yrank[i];
    if ((occ[s >> 6] >> (s & 63)) & 1ull) continue;
    typename Tree::RbNode v{};
    v.x1 = xyI[2 * (size_t)i];
    v.y1 = xyI[2 * (size_t)i + 1];
    const size_t before = tree.n.size();
    if (xshared[i]) {
      typename Tree::Cursor c;
      if (tree.begin_insert(v, m[k], c)) {
        while (tree.descending(c)) tree.step(v, c);
        tree.finish_insert(v, m[k], c);
      }
    } else {
      // predecessor: highest occupied rank below s ; successor: lowest occupied rank above s
      long pred = -1, succ = -1;
      {
        long wi = (long)(s >> 6);
        uint64_t w = occ[wi] & ((1ull << (s & 63)) - 1ull);
        for (;;) {
          if (w) { pred = wi * 64 + 63 - __builtin_clzll(w); break; }
          if (--wi < 0) break;
          w = occ[wi];
        }
        wi = (long)(s >> 6);
        w = (s & 63) == 63 ? 0ull : (occ[wi] & ~((2ull << (s & 63)) - 1ull));
        for (;;) {
          if (w) { succ = wi * 64 + __builtin_ctzll(w); break; }
          if (++wi >= (long)words) break;
          w = occ[wi];
        }
      }
      Idx parent = 0;  // header: empty tree
      if (pred >= 0 && tree.n[slot_node[pred]].child[1] == Tree::kNil) parent = slot_node[pred];
      else if (succ >= 0) parent = slot_node[succ];
      else if (pred >= 0) parent = slot_node[pred];  // unreachable in a consistent tree; keeps the code total
      tree.link(v, m[k], parent);
    }
    if (tree.n.size() != before) {
      occ[s >> 6] |= 1ull << (s & 63);
      slot_node[s] = (Idx)(tree.n.size() - 1);
    }
  }
  size_t out = 0;
  for (size_t wi = 0; wi < words; ++wi) {  // in-order == increasing y rank
    uint64_t w = occ[wi];
    while (w) {
      const int b = __builtin_ctzll(w);
      w &= w - 1;
      m[out++] = tree.payload[slot_node[wi * 64 + b]];
    }
  }
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

// `lanes` (<= kPostLanes) pairs, each in place on ms[t][0..counts[t]); counts[] receives the new sizes.
// xyIs[t] == nullptr or coord_dedup == false: only the (i, j) de-duplication.  ranks (optional): per lane the
// build_view_ranks() tables of view I -- with them the coordinate step runs without tree descents.
void post_process_pairs(int lanes, r3d_indmatch* const* ms, size_t* counts, const float* const* xyIs, const float* const* xyJs,
                        bool coord_dedup, const ViewRankRef* ranks) {
  bool all_xy = coord_dedup;
  size_t nmax = 0;
  for (int t = 0; t < lanes; ++t) {
    // the device hands most pairs over already sorted (k_pack_matches); the check makes the order a performance
    // matter only, never a correctness one
    if (!std::is_sorted(ms[t], ms[t] + counts[t],
                        [](const r3d_indmatch& a, const r3d_indmatch& b) { return a.i < b.i || (a.i == b.i && a.j < b.j); }))
      sort_ij(ms[t], counts[t]);
    counts[t] = (size_t)(std::unique(ms[t], ms[t] + counts[t],
                                     [](const r3d_indmatch& a, const r3d_indmatch& b) { return a.i == b.i && a.j == b.j; }) - ms[t]);
    all_xy = all_xy && xyIs[t] && xyJs[t];
    nmax = std::max(nmax, counts[t]);
  }
  if (!coord_dedup) return;
  if (ranks) {
    for (int t = 0; t < lanes; ++t) {
      if (!xyIs[t] || !xyJs[t]) continue;
      if (!ranks[t].yrank) { post_process_pairs(1, ms + t, counts + t, xyIs + t, xyJs + t, true, nullptr); continue; }
      counts[t] = counts[t] < 65000
                      ? ranked_replay<uint16_t>(ms[t], counts[t], xyIs[t], ranks[t].yrank, ranks[t].xshared, ranks[t].n_slots)
                      : ranked_replay<uint32_t>(ms[t], counts[t], xyIs[t], ranks[t].yrank, ranks[t].xshared, ranks[t].n_slots);
    }
    return;
  }
  if (!all_xy) {  // mixed: one pair at a time
    for (int t = 0; t < lanes; ++t)
      if (xyIs[t] && xyJs[t]) post_process_pairs(1, ms + t, counts + t, xyIs + t, xyJs + t, true, nullptr);
    return;
  }
  if (nmax < 65000) coord_dedup_replay<uint16_t>(lanes, ms, counts, xyIs);
  else coord_dedup_replay<uint32_t>(lanes, ms, counts, xyIs);
}

size_t post_process_pair(r3d_indmatch* m, size_t n, const float* xyI, const float* xyJ, bool coord_dedup) {
  post_process_pairs(1, &m, &n, &xyI, &xyJ, coord_dedup, nullptr);
  return n;
}

}  // namespace r3d

// host-only diagnostic (no GPU needed): the post-processing of one pair, for the CPU test that pins the tree replay
extern "C" int64_t r3d_debug_post_process(r3d_indmatch* m, int64_t n, const float* xyI, const float* xyJ, int coord_dedup) {
  if (!m || n < 0) return -1;
  return (int64_t)r3d::post_process_pair(m, (size_t)n, xyI, xyJ, coord_dedup != 0);
}

extern "C" int r3d_debug_post_process_many(int lanes, r3d_indmatch* const* ms, uint64_t* counts, const float* const* xyIs,
                                           const float* const* xyJs, int coord_dedup) {
  if (lanes < 1 || lanes > r3d::kPostLanes || !ms || !counts) return -1;
  size_t c[r3d::kPostLanes];
  for (int t = 0; t < lanes; ++t) c[t] = (size_t)counts[t];
  // mode 1: lockstep classic replay ; mode 2: ranked replay with tables built here from xyI (n_keypoints needed)
  r3d::post_process_pairs(lanes, ms, c, xyIs, xyJs, coord_dedup != 0, nullptr);
  for (int t = 0; t < lanes; ++t) counts[t] = c[t];
  return 0;
}

// host-only diagnostic: the ranked (descent-free) replay of one pair, tables built from xyI (n_keypoints rows)
extern "C" int64_t r3d_debug_post_process_ranked(r3d_indmatch* m, int64_t n, const float* xyI, uint32_t n_keypoints,
                                                 const float* xyJ) {
  if (!m || n < 0 || !xyI || !xyJ) return -1;
  std::vector<uint32_t> yrank;
  std::vector<uint8_t> xshared;
  uint32_t n_slots = 0;
  r3d::build_view_ranks(xyI, n_keypoints, yrank, xshared, &n_slots);
  r3d::ViewRankRef ref{yrank.empty() ? nullptr : yrank.data(), xshared.empty() ? nullptr : xshared.data(), n_slots};
  size_t c = (size_t)n;
  r3d::post_process_pairs(1, &m, &c, &xyI, &xyJ, true, &ref);
  return (int64_t)c;
}
