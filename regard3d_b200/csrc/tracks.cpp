// tracks.cpp -- openMVG::tracks::TracksBuilder (Build / Filter / ExportToSTL) and TracksUtilsMap::GetTracksInImages,
// the two track utilities Regard3D calls directly (src/threads/PreviewGeneratorThread.cpp:345-358) and the first step of
// every OpenMVG SfM engine it drives.  Host code, like the reference's: one union-find pass over the
// matches.  Upstream (tracks/tracks.hpp, tracks/union_find.hpp) keys a track by the root its union-by-rank / path-
// compression forest ends with; the same union sequence and rules are followed here, on flat arrays:
//   nodes        (view, feature) pairs that occur in a match, numbered in sorted order.  Upstream builds a std::set and a
//                sorted flat map; here a bitmap per view + prefix counts give the same numbering without sorting.
//   unions       pairs in map order, matches in list order; rank[i] < rank[j] ? i under j : j under i (+ rank bump)
//   filter       a track that holds two features of one image, or fewer than min_length images, disappears
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "../../include/r3dgpu.h"
#include "r3d_matches.h"
#include "r3d_sfm.h"

struct r3d_tracks {
  std::vector<uint32_t> ids;     // track ids (upstream: the root's node index), ascending = std::map order
  std::vector<uint64_t> ofs;     // ids.size() + 1
  std::vector<uint32_t> views, feats;  // per track: its (view, feature) pairs, ascending view
};

namespace {

struct UnionFind {
  std::vector<uint32_t> parent, rank, size;
  void init(uint32_t n) {
    parent.resize(n); rank.assign(n, 0); size.assign(n, 1);
    for (uint32_t i = 0; i < n; ++i) parent[i] = i;
  }
  uint32_t find(uint32_t i) {  // full path compression, iteratively
    uint32_t r = i;
    while (parent[r] != r) r = parent[r];
    while (parent[i] != r) { const uint32_t nx = parent[i]; parent[i] = r; i = nx; }
    return r;
  }
  void unite(uint32_t i, uint32_t j) {
    i = find(i); j = find(j);
    if (i == j) return;
    if (rank[i] < rank[j]) { parent[i] = j; size[j] += size[i]; }
    else { parent[j] = i; size[i] += size[j]; if (rank[i] == rank[j]) ++rank[i]; }
  }
};

}  // namespace

extern "C" {

int r3d_tracks_build(const r3d_matches* m, uint32_t min_length, r3d_tracks** out) try {
  if (!m || !out) return R3D_ERR_INVALID;
  *out = nullptr;
  const uint64_t P = m->pairs.size() / 2;
  // ---- node numbering: bitmap of used features per view ----
  uint32_t n_views = 0;
  for (uint64_t p = 0; p < P; ++p) n_views = std::max(n_views, std::max(m->pairs[2 * p], m->pairs[2 * p + 1]) + 1);
  std::vector<uint32_t> max_feat(n_views, 0);
  std::vector<uint8_t> seen(n_views, 0);
  for (uint64_t p = 0; p < P; ++p) {
    const uint32_t I = m->pairs[2 * p], J = m->pairs[2 * p + 1];
    for (const r3d_indmatch& e : m->per[p]) {
      max_feat[I] = std::max(max_feat[I], e.i); seen[I] = 1;
      max_feat[J] = std::max(max_feat[J], e.j); seen[J] = 1;
    }
  }
  std::vector<uint64_t> word_ofs((size_t)n_views + 1, 0);
  for (uint32_t v = 0; v < n_views; ++v) word_ofs[v + 1] = word_ofs[v] + (seen[v] ? (max_feat[v] / 64 + 1) : 0);
  std::vector<uint64_t> bits(word_ofs[n_views], 0);
  for (uint64_t p = 0; p < P; ++p) {
    const uint32_t I = m->pairs[2 * p], J = m->pairs[2 * p + 1];
    uint64_t* bi = bits.data() + word_ofs[I];
    uint64_t* bj = bits.data() + word_ofs[J];
    for (const r3d_indmatch& e : m->per[p]) {
      bi[e.i >> 6] |= 1ull << (e.i & 63);
      bj[e.j >> 6] |= 1ull << (e.j & 63);
    }
  }
  std::vector<uint32_t> word_rank(bits.size() + 1, 0);  // nodes before each word, in (view, feature) order
  for (size_t k = 0; k < bits.size(); ++k) word_rank[k + 1] = word_rank[k] + (uint32_t)__builtin_popcountll(bits[k]);
  const uint32_t n_nodes = word_rank[bits.size()];
  auto node_of = [&](uint32_t v, uint32_t f) {
    const uint64_t wd = word_ofs[v] + (f >> 6);
    return word_rank[wd] + (uint32_t)__builtin_popcountll(bits[wd] & ((1ull << (f & 63)) - 1));
  };
  // ---- unions ----
  UnionFind uf;
  uf.init(n_nodes);
  for (uint64_t p = 0; p < P; ++p) {
    const uint32_t I = m->pairs[2 * p], J = m->pairs[2 * p + 1];
    for (const r3d_indmatch& e : m->per[p]) uf.unite(node_of(I, e.i), node_of(J, e.j));
  }
  for (uint32_t k = 0; k < n_nodes; ++k) uf.find(k);
  // node -> view (walk the bitmaps once)
  std::vector<uint32_t> node_view(n_nodes), node_feat(n_nodes);
  {
    uint32_t k = 0;
    for (uint32_t v = 0; v < n_views; ++v)
      for (uint64_t wd = word_ofs[v]; wd < word_ofs[v + 1]; ++wd) {
        uint64_t b = bits[wd];
        while (b) {
          const int t = __builtin_ctzll(b);
          node_view[k] = v;
          node_feat[k] = (uint32_t)((wd - word_ofs[v]) * 64 + t);
          ++k;
          b &= b - 1;
        }
      }
  }
  // ---- Filter: image-id collisions (nodes of a track arrive in ascending view order: a repeat is adjacent in the
  //      per-root "last view seen"), too short tracks ----
  const uint32_t kNone = 0xffffffffu;
  std::vector<uint32_t> last_view(n_nodes, kNone), n_imgs(n_nodes, 0);
  std::vector<uint8_t> bad(n_nodes, 0);
  for (uint32_t k = 0; k < n_nodes; ++k) {
    const uint32_t r = uf.parent[k];
    if (bad[r]) continue;
    if (last_view[r] == node_view[k]) bad[r] = 1;
    else { last_view[r] = node_view[k]; ++n_imgs[r]; }
  }
  for (uint32_t r = 0; r < n_nodes; ++r)
    if (n_imgs[r] && n_imgs[r] < min_length) bad[r] = 1;
  // ---- export: tracks in root order, nodes in (view, feature) order; 1-node sets are not tracks ----
  std::vector<uint32_t> cnt(n_nodes, 0);
  for (uint32_t k = 0; k < n_nodes; ++k) {
    const uint32_t r = uf.parent[k];
    if (!bad[r] && uf.size[r] > 1) ++cnt[r];
  }
  r3d_tracks* t = new r3d_tracks();
  std::vector<uint64_t> start(n_nodes, 0);
  uint64_t total = 0;
  for (uint32_t r = 0; r < n_nodes; ++r)
    if (cnt[r]) { t->ids.push_back(r); t->ofs.push_back(total); start[r] = total; total += cnt[r]; }
  t->ofs.push_back(total);
  t->views.resize(total);
  t->feats.resize(total);
  for (uint32_t k = 0; k < n_nodes; ++k) {
    const uint32_t r = uf.parent[k];
    if (!bad[r] && uf.size[r] > 1) { t->views[start[r]] = node_view[k]; t->feats[start[r]] = node_feat[k]; ++start[r]; }
  }
  *out = t;
  return R3D_OK;
} catch (const std::bad_alloc&) { return R3D_ERR_NOMEM; } catch (...) { return R3D_ERR_INVALID; }

uint64_t r3d_tracks_count(const r3d_tracks* t) { return t ? t->ids.size() : 0; }

int r3d_tracks_get(const r3d_tracks* t, uint64_t k, uint32_t* track_id, const uint32_t** views, const uint32_t** feats, uint32_t* n) {
  if (!t || k >= t->ids.size()) return R3D_ERR_INVALID;
  if (track_id) *track_id = t->ids[k];
  if (views) *views = t->views.data() + t->ofs[k];
  if (feats) *feats = t->feats.data() + t->ofs[k];
  if (n) *n = (uint32_t)(t->ofs[k + 1] - t->ofs[k]);
  return R3D_OK;
}

// TracksUtilsMap::GetTracksInImages: the tracks that contain EVERY listed view, restricted to those views
int r3d_tracks_in_images(const r3d_tracks* t, const uint32_t* view_ids, uint32_t n, r3d_tracks** out) try {
  if (!t || !out || (n && !view_ids)) return R3D_ERR_INVALID;
  std::vector<uint32_t> want(view_ids, view_ids + n);
  std::sort(want.begin(), want.end());
  want.erase(std::unique(want.begin(), want.end()), want.end());
  r3d_tracks* r = new r3d_tracks();
  uint64_t total = 0;
  for (size_t k = 0; k < t->ids.size(); ++k) {
    const uint32_t* v = t->views.data() + t->ofs[k];
    const uint32_t* f = t->feats.data() + t->ofs[k];
    const size_t len = (size_t)(t->ofs[k + 1] - t->ofs[k]);
    size_t found = 0;
    uint32_t tmp_v[64], tmp_f[64];
    std::vector<uint32_t> bv, bf;
    for (uint32_t wv : want) {
      const uint32_t* it = std::lower_bound(v, v + len, wv);
      if (it == v + len || *it != wv) break;
      if (found < 64) { tmp_v[found] = wv; tmp_f[found] = f[it - v]; } else { bv.push_back(wv); bf.push_back(f[it - v]); }
      ++found;
    }
    if (found == 0 || found != want.size()) continue;
    r->ids.push_back(t->ids[k]);
    r->ofs.push_back(total);
    for (size_t q = 0; q < std::min<size_t>(found, 64); ++q) { r->views.push_back(tmp_v[q]); r->feats.push_back(tmp_f[q]); }
    r->views.insert(r->views.end(), bv.begin(), bv.end());
    r->feats.insert(r->feats.end(), bf.begin(), bf.end());
    total += found;
  }
  r->ofs.push_back(total);
  *out = r;
  return R3D_OK;
} catch (const std::bad_alloc&) { return R3D_ERR_NOMEM; } catch (...) { return R3D_ERR_INVALID; }

void r3d_tracks_free(r3d_tracks* t) { delete t; }

}  // extern "C"
