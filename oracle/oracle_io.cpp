// oracle_io.cpp -- CPU ORACLE (test infrastructure; see oracle.h).
// On-disk interfaces of the path (names fixed by src/R3DProject.cpp:848-871):
//   <img>.feat  text, "x y scale orientation" per line, default ostream precision
//               (writer: src/keypointSet.hpp:61-67 -> upstream saveFeatsToFile; reader used by the
//               GUI: src/threads/PreviewGeneratorThread.cpp:313-321)
//   <img>.desc  binary, size_t count + count*dim float32 (upstream saveDescsToBinFile; loaded at
//               src/R3DComputeMatches.cpp:2040)
//   matches.*.txt  "I J\ncount\n i j\n..." in std::map order (upstream matching::Save, called at
//               src/R3DComputeMatches.cpp:2064, :2120, :2196, :2224)
#include "oracle.h"
#include <fstream>
#include <map>
#include <vector>

extern "C" {

int orc_save_feat(const char* path, const float* f, uint32_t n) {
  std::ofstream file(path);
  if (!file.is_open()) return 1;
  for (uint32_t i = 0; i < n; ++i)
    file << f[4 * i] << " " << f[4 * i + 1] << " " << f[4 * i + 2] << " " << f[4 * i + 3] << "\n";
  return file.good() ? 0 : 1;
}

int orc_load_feat(const char* path, float* f, uint32_t cap, uint32_t* n) {
  std::ifstream file(path);
  if (!file.is_open()) return 1;
  uint32_t k = 0;
  float x, y, s, o;
  while (file >> x >> y >> s >> o) {
    if (k >= cap) return 2;
    f[4 * k] = x; f[4 * k + 1] = y; f[4 * k + 2] = s; f[4 * k + 3] = o;
    ++k;
  }
  *n = k;
  return 0;
}

int orc_save_desc_f32(const char* path, const float* d, uint64_t n, uint32_t dim) {
  std::ofstream file(path, std::ios::out | std::ios::binary);
  if (!file.is_open()) return 1;
  const std::size_t card = (std::size_t)n;
  file.write((const char*)&card, sizeof(std::size_t));
  file.write((const char*)d, (std::streamsize)(n * dim * sizeof(float)));
  return file.good() ? 0 : 1;
}

int orc_load_desc_f32(const char* path, float* d, uint64_t cap_rows, uint32_t dim, uint64_t* n) {
  std::ifstream file(path, std::ios::in | std::ios::binary);
  if (!file.is_open()) return 1;
  std::size_t card = 0;
  file.read((char*)&card, sizeof(std::size_t));
  if (!file.good()) return 1;
  *n = card;
  if (card > cap_rows) return 2;
  file.read((char*)d, (std::streamsize)(card * dim * sizeof(float)));
  return file.good() || file.eof() ? 0 : 1;
}

int orc_save_matches_txt(const char* path, const uint32_t* pairs, uint64_t P,
                         const uint64_t* pair_ofs, const orc_indmatch* m) {
  // PairWiseMatches is a std::map keyed by (I,J): iterate in key order, skip empty pairs (they
  // are never inserted: src/R3DComputeMatches.cpp:483-486).
  std::map<std::pair<uint32_t, uint32_t>, uint64_t> order;
  for (uint64_t p = 0; p < P; ++p)
    if (pair_ofs[p + 1] > pair_ofs[p]) order[{pairs[2 * p], pairs[2 * p + 1]}] = p;
  std::ofstream stream(path);
  if (!stream.is_open()) return 1;
  for (const auto& kv : order) {
    const uint64_t p = kv.second;
    stream << kv.first.first << " " << kv.first.second << '\n'
           << (pair_ofs[p + 1] - pair_ofs[p]) << '\n';
    for (uint64_t k = pair_ofs[p]; k < pair_ofs[p + 1]; ++k) stream << m[k].i << " " << m[k].j << "\n";
  }
  return stream.good() ? 0 : 1;
}
}
