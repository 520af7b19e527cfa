// matches.cpp -- PairWiseMatches container + matching::Save / matching::Load (text format).
// Reference call sites: Save(map, "matches.putative.txt") src/R3DComputeMatches.cpp:2064,
// Save(map, matchesFFilename_) :2120; consumers that Load it: src/threads/R3DTriangulationThread.cpp:222,
// :411 and src/threads/PreviewGeneratorThread.cpp:337-338.  Format: SURVEY.md Appendix B.3.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "r3d_matches.h"

namespace r3d { void pool_parallel_for(int n_threads, size_t n, const std::function<void(size_t)>& f); }

extern "C" {

uint64_t r3d_matches_num_pairs(const r3d_matches* m) { return m ? m->pairs.size() / 2 : 0; }
uint64_t r3d_matches_total(const r3d_matches* m) { return m ? m->total : 0; }

int r3d_matches_get_pair(const r3d_matches* m, uint64_t k, uint32_t* I, uint32_t* J, const r3d_indmatch** matches,
                         uint64_t* count) {
  if (!m || k >= m->pairs.size() / 2) return R3D_ERR_INVALID;
  if (I) *I = m->pairs[2 * k];
  if (J) *J = m->pairs[2 * k + 1];
  if (matches) *matches = m->per[k].data();
  if (count) *count = m->per[k].size();
  return R3D_OK;
}

int r3d_matches_from_csr(const uint32_t* pairs, uint64_t n_pairs, const uint64_t* pair_ofs, const r3d_indmatch* matches,
                         r3d_matches** out) try {
  if (!out || (n_pairs && (!pairs || !pair_ofs))) return R3D_ERR_INVALID;
  std::vector<uint64_t> order(n_pairs);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
    return pairs[2 * a] < pairs[2 * b] || (pairs[2 * a] == pairs[2 * b] && pairs[2 * a + 1] < pairs[2 * b + 1]);
  });
  r3d_matches* m = new r3d_matches();
  for (uint64_t t = 0; t < n_pairs; ++t) {
    const uint64_t p = order[t];
    if (pair_ofs[p + 1] == pair_ofs[p]) continue;  // empty pairs are never in the map
    if (!m->pairs.empty() && m->pairs[m->pairs.size() - 2] == pairs[2 * p] && m->pairs.back() == pairs[2 * p + 1]) continue;
    m->push(pairs[2 * p], pairs[2 * p + 1], std::vector<r3d_indmatch>(matches + pair_ofs[p], matches + pair_ofs[p + 1]));
  }
  *out = m;
  return R3D_OK;
} catch (const std::bad_alloc&) { return R3D_ERR_NOMEM; } catch (...) { return R3D_ERR_INVALID; }

/* Flat copy of the map (what a host gather across ranks ships): pairs_out 2 x num_pairs view ids in map order,
 * ofs_out num_pairs + 1 prefix offsets, matches_out total entries.  Any output may be NULL. */
int r3d_matches_export_csr(const r3d_matches* m, uint32_t* pairs_out, uint64_t* ofs_out, r3d_indmatch* matches_out) try {
  if (!m) return R3D_ERR_INVALID;
  const uint64_t P = m->pairs.size() / 2;
  if (pairs_out && P) std::memcpy(pairs_out, m->pairs.data(), sizeof(uint32_t) * 2 * P);
  std::vector<uint64_t> ofs(P + 1, 0);
  for (uint64_t k = 0; k < P; ++k) ofs[k + 1] = ofs[k] + m->per[k].n;
  if (ofs_out) std::memcpy(ofs_out, ofs.data(), sizeof(uint64_t) * (P + 1));
  if (matches_out && P) {
    // slabs of pairs of ~4 MB each, copied by the persistent host pool (a multi-GPU gather exports hundreds of MB)
    std::vector<uint64_t> cut{0};
    for (uint64_t k = 1; k <= P; ++k)
      if (k == P || (ofs[k] - ofs[cut.back()]) * sizeof(r3d_indmatch) >= ((size_t)4 << 20)) cut.push_back(k);
    const std::function<void(size_t)> body = [&](size_t c) {
      for (uint64_t k = cut[c]; k < cut[c + 1]; ++k)
        if (m->per[k].n) std::memcpy(matches_out + ofs[k], m->per[k].p, sizeof(r3d_indmatch) * m->per[k].n);
    };
    r3d::pool_parallel_for(16, cut.size() - 1, body);
  }
  return R3D_OK;
} catch (...) { return R3D_ERR_NOMEM; }

void r3d_free_matches(r3d_matches* m) { delete m; }

int r3d_save_matches_txt(const r3d_matches* m, const char* path) try {
  if (!m || !path) return R3D_ERR_INVALID;
  std::ofstream stream(path);
  if (!stream.is_open()) return R3D_ERR_IO;
  const uint64_t P = m->pairs.size() / 2;
  for (uint64_t k = 0; k < P; ++k) {
    stream << m->pairs[2 * k] << " " << m->pairs[2 * k + 1] << '\n' << m->per[k].size() << '\n';
    for (const r3d_indmatch& im : m->per[k]) stream << im.i << " " << im.j << "\n";
  }
  return stream.good() ? R3D_OK : R3D_ERR_IO;
} catch (...) { return R3D_ERR_IO; }

int r3d_load_matches_txt(const char* path, r3d_matches** out) try {
  if (!path || !out) return R3D_ERR_INVALID;
  std::ifstream stream(path);
  if (!stream.is_open()) return R3D_ERR_IO;
  stream.seekg(0, std::ios::end);
  const uint64_t file_bytes = (uint64_t)std::max<std::streamoff>(0, stream.tellg());
  stream.seekg(0, std::ios::beg);
  std::map<std::pair<uint32_t, uint32_t>, std::vector<r3d_indmatch>> mp;
  uint32_t I, J;
  uint64_t number;
  while (stream >> I >> J >> number) {
    // a match line is at least "i j\n" = 4 bytes: a count the rest of the file cannot hold is a corrupt file, not
    // an allocation request
    if (number > file_bytes / 4) return R3D_ERR_IO;
    std::vector<r3d_indmatch> v(number);
    for (uint64_t k = 0; k < number; ++k)
      if (!(stream >> v[k].i >> v[k].j)) return R3D_ERR_IO;  // truncated inside pair (I, J)
    mp[{I, J}] = std::move(v);
  }
  r3d_matches* m = new r3d_matches();
  for (auto& kv : mp)
    m->push(kv.first.first, kv.first.second, std::move(kv.second));  // matching::Load keeps what the file lists
  *out = m;
  return R3D_OK;
} catch (const std::bad_alloc&) { return R3D_ERR_NOMEM; } catch (...) { return R3D_ERR_IO; }

}  // extern "C"
