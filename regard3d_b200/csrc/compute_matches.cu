// compute_matches.cu -- file-level twin of R3DComputeMatches::computeMatches()
// (src/R3DComputeMatches.cpp:1667-2256), for the steps after feature extraction:
//   Regions_Provider::load            (:2040)  <img>.feat / <img>.desc  -> r3d_upload_regions
//   exhaustivePairs(N)                (:2042)
//   Matcher_Regions::Match            (:2048)  -> r3d_match_pairs
//   Save(matches.putative.txt)        (:2064)
//   Robust_model_estimation(F, AC)    (:2113)  -> r3d_filter_pairs
//   Save(matchesFFilename_)           (:2120)
//   Robust_model_estimation(E, AC)    (:2169)  -> r3d_filter_pairs + poor-overlap removal (:2173-2191), Save(:2196)
//   Robust_model_estimation(H, AC)    (:2216)  -> r3d_filter_pairs, Save(:2224)
// Progress fractions as the reference emits them (0.7 before matching :2000, 0.8 before the F filter
// :2107).  File formats: SURVEY.md Appendix B.  Feature extraction (AKAZE + LIOP on the CPU thread
// pool, src/threads/R3DFeaturesThread.cpp) stays with the caller: it is a "next" row.
#include "r3d_internal.cuh"

#include <chrono>
#include <fstream>
#include <string>
#include <vector>

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// <basename>.feat : "x y scale orientation" per line (src/keypointSet.hpp:61-67 -> saveFeatsToFile)
bool load_feat(const std::string& path, std::vector<float>& xy) {
  std::ifstream f(path);
  if (!f.is_open()) return false;
  float x, y, s, o;
  xy.clear();
  while (f >> x >> y >> s >> o) {
    xy.push_back(x);
    xy.push_back(y);
  }
  return true;
}

// <basename>.desc : size_t count + count * dim float32 (saveDescsToBinFile)
bool load_desc(const std::string& path, uint32_t dim, std::vector<float>& d, uint64_t* n) {
  std::ifstream f(path, std::ios::in | std::ios::binary);
  if (!f.is_open()) return false;
  std::size_t card = 0;
  f.read((char*)&card, sizeof(std::size_t));
  if (!f.good()) return false;
  d.resize((size_t)card * dim);
  if (card) f.read((char*)d.data(), (std::streamsize)(d.size() * sizeof(float)));
  if (card && f.gcount() != (std::streamsize)(d.size() * sizeof(float))) return false;
  *n = card;
  return true;
}

}  // namespace

using namespace r3d;

// PairWiseMatchingToAdjacencyMatrixSVG (upstream graph_stats / svgDrawer; called at src/R3DComputeMatches.cpp:2074,
// :2238): an N x N grid, one blue square per pair that has matches (column J, row I), and the two 0..N axes.
static bool write_adjacency_svg(uint32_t n_images, const r3d_matches* m, const std::string& path) {
  if (!m || r3d_matches_num_pairs(m) == 0) return true;  // upstream writes nothing for an empty map
  const float scale = 5.0f;
  std::ofstream f(path.c_str());
  if (!f.is_open()) return false;
  const float W = (float)(n_images + 3) * 5.0f;
  f << "<?xml version=\"1.0\" standalone=\"yes\"?>\n<!-- SVG graphic -->\n<svg xmlns='http://www.w3.org/2000/svg'"
    << " width=\"" << W << "px\" height=\"" << W << "px\" preserveAspectRatio=\"xMinYMin meet\" viewBox=\"0 0 " << W << ' ' << W
    << "\" version=\"1.1\">\n";
  const uint64_t P = r3d_matches_num_pairs(m);
  for (uint64_t k = 0; k < P; ++k) {
    uint32_t I, J;
    uint64_t cnt;
    r3d_matches_get_pair(m, k, &I, &J, nullptr, &cnt);
    if (!cnt || I >= n_images || J >= n_images) continue;
    f << "<rect x=\"" << J * scale << "\" y=\"" << I * scale << "\" width=\"" << scale / 2.0f << "\" height=\"" << scale / 2.0f
      << "\" fill=\"blue\"><title>(" << J << ',' << I << ' ' << cnt << ")</title></rect>\n";
  }
  const float a = (float)(n_images + 1) * scale, b = (float)n_images * scale;
  auto text = [&](float x, float y, const std::string& t) {
    f << "<text x=\"" << x << "\" y=\"" << y << "\" font-size=\"" << scale << "\" fill=\"black\">" << t << "</text>\n";
  };
  auto line = [&](float x0, float y0, float x1, float y1) {
    f << "<polyline points=\"" << x0 << ',' << y0 << ' ' << x1 << ',' << y1 << "\" stroke=\"black\" stroke-width=\"1\"/>\n";
  };
  text(a, scale, "0");
  text(a, b - scale, std::to_string(n_images));
  line(a, 2 * scale, a, b - 2 * scale);
  text(scale, a, "0");
  text(b - scale, a, std::to_string(n_images));
  line(2 * scale, a, b - 2 * scale, a);
  f << "</svg>\n";
  return f.good();
}

extern "C" int r3d_compute_matches(r3d_ctx* ctx, const r3d_cm_params* params, const r3d_cm_paths* paths,
                                   r3d_progress_cb cb, void* user, r3d_cm_stats* stats) {
  if (!ctx || !params || !paths || !paths->matches_dir || !paths->image_basenames || !paths->views)
    return fail(ctx, R3D_ERR_INVALID, "r3d_compute_matches: bad arguments");
  const uint32_t N = paths->n_views;
  const uint32_t dim = params->descriptor_dim ? params->descriptor_dim : 144;  // R3D_AKAZE_LIOP_Regions
  const std::string dir(paths->matches_dir);
  if (stats) {
    stats->putative_pairs = stats->putative_matches = stats->f_pairs = stats->f_matches = 0;
    stats->h_pairs = stats->h_matches = 0;
    stats->e_pairs = stats->e_matches = 0;
    stats->seconds_load = stats->seconds_match = stats->seconds_filter = 0;
  }
  // ---- regions ------------------------------------------------------------------------------------
  double t0 = now_s();
  int rc = r3d_clear_regions(ctx);
  if (rc) return rc;
  std::vector<float> xy, desc;
  for (uint32_t v = 0; v < N; ++v) {
    const std::string base = dir + "/" + paths->image_basenames[v];
    uint64_t n = 0;
    if (!load_feat(base + ".feat", xy) || !load_desc(base + ".desc", dim, desc, &n))
      return fail(ctx, R3D_ERR_IO, "r3d_compute_matches: cannot read regions of " + base);  // "Invalid regions"
    if (xy.size() / 2 != n) return fail(ctx, R3D_ERR_IO, "r3d_compute_matches: .feat/.desc count mismatch for " + base);
    rc = r3d_upload_regions(ctx, v, desc.data(), (uint32_t)n, dim, R3D_F32, xy.data());
    if (rc) return rc;
    if (stats && stats->number_of_keypoints && v < stats->n_views) stats->number_of_keypoints[v] = (uint32_t)n;
  }
  if (stats) stats->seconds_load = now_s() - t0;
  // ---- putative matches ---------------------------------------------------------------------------
  if (cb) cb(0.7f, "Computing matches", user);
  t0 = now_s();
  std::vector<uint32_t> pairs;
  for (uint32_t i = 0; i < N; ++i)
    for (uint32_t j = i + 1; j < N; ++j) { pairs.push_back(i); pairs.push_back(j); }
  r3d_matches* put = nullptr;
  // every matchingAlgorithm value of the reference (0 FLANN, 1-3 KGraph, 4 brute force, 5 MRPT, 6-8 HNSW:
  // src/R3DComputeMatches.cpp:2036-2062) maps to the exact brute-force matcher: the ANN variants are
  // approximations of it
  // matchingAlgorithm 0..8 (FLANN / KGraph / MRPT / HNSW / brute force) all map to the exact matcher; the one extension is
  // R3D_MATCHING_CASCADE_HASHING = OpenMVG's CASCADE_HASHING_L2, which the reference's switch does not offer
  const uint32_t mflags = params->matching_algorithm == R3D_MATCHING_CASCADE_HASHING ? R3D_MATCH_CASCADE_HASHING : R3D_MATCH_DEFAULT;
  rc = r3d_match_pairs(ctx, pairs.data(), pairs.size() / 2, params->dist_ratio, mflags, &put);
  if (rc) return rc;
  if (stats) {
    stats->seconds_match = now_s() - t0;
    stats->putative_pairs = r3d_matches_num_pairs(put);
    stats->putative_matches = r3d_matches_total(put);
  }
  rc = r3d_save_matches_txt(put, (dir + "/matches.putative.txt").c_str());
  if (rc) { r3d_free_matches(put); return fail(ctx, R3D_ERR_IO, "r3d_compute_matches: cannot save matches.putative.txt"); }
  if (params->svg_output) write_adjacency_svg(N, put, dir + "/PutativeAdjacencyMatrix.svg");  // :2074-2076
  // ---- geometric filtering -------------------------------------------------------------------------
  if (params->compute_fundamental) {
    if (cb) cb(0.8f, "Calculate fundamental matrix", user);
    t0 = now_s();
    r3d_matches* fm = nullptr;
    rc = r3d_filter_pairs(ctx, R3D_MODEL_F, 4.0, 2048, put, paths->views, N, &fm);  // maxResidualError 4.0, 2048 iterations
    if (rc) { r3d_free_matches(put); return rc; }
    if (stats) {
      stats->seconds_filter = now_s() - t0;
      stats->f_pairs = r3d_matches_num_pairs(fm);
      stats->f_matches = r3d_matches_total(fm);
    }
    const std::string fpath = paths->matches_f_filename ? std::string(paths->matches_f_filename) : dir + "/matches.f.txt";
    rc = r3d_save_matches_txt(fm, fpath.c_str());
    if (params->svg_output) write_adjacency_svg(N, fm, dir + "/GeometricAdjacencyMatrix.svg");  // :2238-2240
    r3d_free_matches(fm);
    if (rc) { r3d_free_matches(put); return fail(ctx, R3D_ERR_IO, "r3d_compute_matches: cannot save " + fpath); }
  }
  if (params->compute_essential) {  // src/R3DComputeMatches.cpp:2130-2204
    if (cb) cb(0.9f, "Calculate essential matrix", user);
    r3d_matches* em = nullptr;
    rc = r3d_filter_pairs(ctx, R3D_MODEL_E, 4.0, 2048, put, paths->views, N, &em);
    if (rc) { r3d_free_matches(put); return rc; }
    // "Perform an additional check to remove pairs with poor overlap" (:2173-2191): drop a pair when it keeps
    // fewer than 50 matches or less than 30 % (float ratio) of its putative matches
    {
      std::vector<uint32_t> kp;
      std::vector<uint64_t> kofs(1, 0);
      std::vector<r3d_indmatch> km;
      const uint64_t ne = r3d_matches_num_pairs(em), np = r3d_matches_num_pairs(put);
      uint64_t q = 0;  // both maps are sorted by (I,J): merge walk
      for (uint64_t k = 0; k < ne; ++k) {
        uint32_t I, J, PI = 0, PJ = 0;
        const r3d_indmatch* mm;
        const r3d_indmatch* pm;
        uint64_t cnt, pcnt = 0;
        r3d_matches_get_pair(em, k, &I, &J, &mm, &cnt);
        for (; q < np; ++q) {
          r3d_matches_get_pair(put, q, &PI, &PJ, &pm, &pcnt);
          if (PI == I && PJ == J) break;
        }
        const float ratio = cnt / (float)pcnt;
        if (cnt < 50 || ratio < .3f) continue;
        kp.push_back(I);
        kp.push_back(J);
        km.insert(km.end(), mm, mm + cnt);
        kofs.push_back(km.size());
      }
      r3d_free_matches(em);
      em = nullptr;
      rc = r3d_matches_from_csr(kp.data(), kp.size() / 2, kofs.data(), km.data(), &em);
      if (rc) { r3d_free_matches(put); return fail(ctx, rc, "r3d_compute_matches: essential overlap filter"); }
    }
    if (stats) {
      stats->e_pairs = r3d_matches_num_pairs(em);
      stats->e_matches = r3d_matches_total(em);
    }
    const std::string epath = paths->matches_e_filename ? std::string(paths->matches_e_filename) : dir + "/matches.e.txt";
    rc = r3d_save_matches_txt(em, epath.c_str());
    r3d_free_matches(em);
    if (rc) { r3d_free_matches(put); return fail(ctx, R3D_ERR_IO, "r3d_compute_matches: cannot save " + epath); }
  }
  if (params->compute_homography) {  // src/R3DComputeMatches.cpp:2206-2233
    if (cb) cb(0.95f, "Calculate homography matrix", user);
    r3d_matches* hm = nullptr;
    rc = r3d_filter_pairs(ctx, R3D_MODEL_H, 4.0, 2048, put, paths->views, N, &hm);
    if (rc) { r3d_free_matches(put); return rc; }
    if (stats) {
      stats->h_pairs = r3d_matches_num_pairs(hm);
      stats->h_matches = r3d_matches_total(hm);
    }
    const std::string hpath = paths->matches_h_filename ? std::string(paths->matches_h_filename) : dir + "/matches.h.txt";
    rc = r3d_save_matches_txt(hm, hpath.c_str());
    r3d_free_matches(hm);
    if (rc) { r3d_free_matches(put); return fail(ctx, R3D_ERR_IO, "r3d_compute_matches: cannot save " + hpath); }
  }
  r3d_free_matches(put);
  if (cb) cb(1.0f, "Done", user);
  return R3D_OK;
}
