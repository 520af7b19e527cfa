// oracle_liop.cpp -- CPU ORACLE (test infrastructure; see oracle.h) of the LIOP-144 descriptor stage,
// Regard3DFeatures::extractLIOPFeatures (src/Regard3DFeatures.cpp:719-861).
//
// Two parts with different parity status:
//  (1) liop_process(): restatement of r3d_vl_liopdesc_new / r3d_vl_liopdesc_process
//      (src/thirdparty/liop/vl_liop.c:318-394 tables, :434-575 descriptor, quick sort of
//      src/thirdparty/liop/vl_qsort-def.h:123-162).  The reference file compiles standalone here, so
//      this part is PINNED: oracle/_ref/libvlliop_ref.so (built from the reference source where it lies,
//      oracle/Makefile) must agree bit for bit (tests/test_oracle_liop.py), and its outputs on seeded
//      patches are committed as golden fixtures (tests/golden/liop_ref_v1.npz).
//  (2) liop_patch(): the 41x41 patch the reference obtains with OpenCV -- cv::warpAffine(INTER_LINEAR |
//      WARP_INVERSE_MAP, constant border 0) + cv::GaussianBlur(sigma 1.2) (src/Regard3DFeatures.cpp:766-806).
//      OpenCV is an un-vendored dependency (src/CMakeLists.txt:198, README: 4.0); its published algorithm is
//      restated (imgwarp.cpp WarpAffineInvoker + remapBilinear: 1/32-pixel fixed-point coordinates, float weight
//      table; smooth.cpp/filter.cpp: 11-tap separable kernel, BORDER_REFLECT_101) and pinned against
//      cv2 4.13 run in the build container (tests/golden/liop_patch_cv2_v1.npz).  OpenCV's own result depends
//      on the CPU it dispatches to (FMA in the AVX2 row filter), so that comparison carries a tolerance of a
//      few float ulp; the warp alone is bit-exact.
#include "oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

namespace {

constexpr int kSide = 41;             // patchSize = 2 * patchResolution + 1 (Regard3DFeatures.cpp:726-730)
constexpr int kNeigh = 4;             // DEFAULT_NUM_NEIGHBOURS
constexpr int kBinsSpatial = 6;       // DEFAULT_NUM_SPATIAL_BINS
constexpr double kRadius = 6.0;       // DEFAULT_RADIUS
constexpr float kIntensityThreshold = -(5.0 / 255);

struct LiopTables {
  std::vector<uint32_t> pixels;       // patchPixels
  std::vector<double> sx, sy;         // neighSamplesX / neighSamplesY, kNeigh per patch pixel
};

// r3d_vl_liopdesc_new (vl_liop.c:318-394), sideLength = 41
const LiopTables& tables() {
  static LiopTables T = [] {
    LiopTables t;
    const long center = (kSide - 1) / 2;
    const float radius = (float)kRadius;  // the parameter is a float
    const double tt = center - radius + 0.6;
    const long t2 = (long)(tt * tt);
    for (long y = 0; y < kSide; ++y)
      for (long x = 0; x < kSide; ++x) {
        const long dx = x - center, dy = y - center;
        if (x == 0 && y == 0) continue;
        if (dx * dx + dy * dy <= t2) t.pixels.push_back((uint32_t)(x + y * kSide));
      }
    const size_t n = t.pixels.size();
    t.sx.assign(n * kNeigh, 0.0);
    t.sy.assign(n * kNeigh, 0.0);
    for (size_t i = 0; i < n; ++i) {
      const double dangle = 2 * M_PI / (double)kNeigh;
      const long pixel = (long)t.pixels[i];
      const double x = (double)((pixel % kSide) - center);
      const double y = (double)((pixel / kSide) - center);
      const double angle0 = atan2(y, x);
      for (int k = 0; k < kNeigh; ++k) {
        t.sx[k + kNeigh * i] = x + radius * cos(angle0 + dangle * k) + center;
        t.sy[k + kNeigh * i] = y + radius * sin(angle0 + dangle * k) + center;
      }
    }
    return t;
  }();
  return T;
}

// vl_qsort-def.h:123-162 (Lomuto partition around the middle element, "<= 0" goes low), on a permutation compared
// through the float difference of the keyed intensities (patch_cmp / neigh_cmp, vl_liop.c:239-273).  The order of
// equal keys is a property of exactly this procedure; the sub-ranges are independent, so an explicit stack replaces
// the recursion.
template <typename Idx>
void vl_qsort_perm(Idx* perm, const float* key, long size) {
  if (size < 1) return;
  std::vector<std::pair<long, long>> stack;
  stack.push_back({0, size - 1});
  while (!stack.empty()) {
    const long begin = stack.back().first, end = stack.back().second;
    stack.pop_back();
    long pivot = (end + begin) / 2;
    std::swap(perm[pivot], perm[end]);
    pivot = end;
    long lowPart = begin;
    for (long i = begin; i < end; ++i) {
      if (key[perm[i]] - key[perm[pivot]] <= 0) {
        std::swap(perm[lowPart], perm[i]);
        lowPart++;
      }
    }
    std::swap(perm[lowPart], perm[pivot]);
    pivot = lowPart;
    if (pivot > begin) stack.push_back({begin, pivot - 1});
    if (pivot < end) stack.push_back({pivot + 1, end});
  }
}

long vl_floor_d(double x) {
  const long xi = (long)x;
  if (x >= 0 || (double)xi == x) return xi;
  return xi - 1;
}

}  // namespace

uint32_t liop_patch_size() { return (uint32_t)tables().pixels.size(); }

// r3d_vl_liopdesc_process (vl_liop.c:434-575): desc[144] from a 41x41 float patch
void liop_process(const float* patch, float* desc) {
  const LiopTables& T = tables();
  const long n = (long)T.pixels.size();
  const int dimension = 24 * kBinsSpatial;
  std::memset(desc, 0, sizeof(float) * dimension);
  std::vector<float> inten(n);
  std::vector<uint32_t> perm(n);
  for (long i = 0; i < n; ++i) {
    inten[i] = patch[T.pixels[i]];
    perm[i] = (uint32_t)i;
  }
  vl_qsort_perm(perm.data(), inten.data(), n);
  float threshold;
  if (kIntensityThreshold < 0) {
    const long i = perm[0], t = perm[n - 1];
    threshold = -kIntensityThreshold * (inten[t] - inten[i]);
  } else {
    threshold = kIntensityThreshold;
  }
  const long numPermutations = 24;
  const long spatialBinArea = n / kBinsSpatial;
  long spatialBinEnd = spatialBinArea, spatialBinIndex = 0, offset = 0;
  for (long i = 0; i < n; ++i) {
    if (i >= spatialBinEnd && spatialBinIndex < kBinsSpatial - 1) {
      spatialBinEnd += spatialBinArea;
      spatialBinIndex++;
      offset += numPermutations;
    }
    const double* sx = T.sx.data() + kNeigh * perm[i];
    const double* sy = T.sy.data() + kNeigh * perm[i];
    float nI[kNeigh];
    unsigned long nP[kNeigh];
    for (int t = 0; t < kNeigh; ++t) {
      const double x = sx[t], y = sy[t];
      const long ix = vl_floor_d(x), iy = vl_floor_d(y);
      const double wx = x - ix, wy = y - iy;
      double a = 0, b = 0, c = 0, d = 0;
      const int L = kSide;
      if (ix >= 0 && iy >= 0) a = patch[ix + iy * L];
      if (ix < L - 1 && iy >= 0) b = patch[ix + 1 + iy * L];
      if (ix >= 0 && iy < L - 1) c = patch[ix + (iy + 1) * L];
      if (ix < L - 1 && iy < L - 1) d = patch[ix + 1 + (iy + 1) * L];
      nP[t] = (unsigned long)t;
      nI[t] = (float)((1.0 - wy) * (a + (b - a) * wx) + wy * (c + (d - c) * wx));
    }
    vl_qsort_perm(nP, nI, kNeigh);
    // get_permutation_index (vl_liop.c:222-235)
    long permIndex = 0;
    for (long a = 0; a < kNeigh; ++a) {
      permIndex = permIndex * (kNeigh - a) + (long)nP[a];
      for (long b = a + 1; b < kNeigh; ++b)
        if (nP[b] > nP[a]) nP[b]--;
    }
    float weight = 0;
    for (int k = 0; k < kNeigh; ++k)
      for (int t = k + 1; t < kNeigh; ++t) {
        const float a = nI[k], b = nI[t];
        weight += (a > b + threshold || b > a + threshold);
      }
    desc[permIndex + offset] += weight;
  }
  float norm = 0;
  for (int i = 0; i < dimension; ++i) norm += desc[i] * desc[i];
  // `float norm ... norm = VL_MAX(sqrt(norm), 1e-12)`: the double square root is stored back into the FLOAT
  norm = (float)(std::sqrt((double)norm) > 1e-12 ? std::sqrt((double)norm) : 1e-12);
  for (int i = 0; i < dimension; ++i) desc[i] /= norm;
}

// cv::getGaussianKernel(11, 1.2, CV_32F) of OpenCV 4.x (bit-exact softdouble kernel, rounded to float);
// ksize = cvRound(1.2 * 4 * 2 + 1) | 1 = 11 for a CV_32F image (smooth.cpp createGaussianKernels)
static const float kGauss11[11] = {0x1.d9b2eep-15f, 0x1.50eab6p-10f, 0x1.dea402p-7f, 0x1.538cacp-4f, 0x1.e1217cp-3f, 0x1.546e7ep-2f,
                                   0x1.e1217cp-3f,  0x1.538cacp-4f,  0x1.dea402p-7f, 0x1.50eab6p-10f, 0x1.d9b2eep-15f};

static inline int cv_round(double v) { return (int)std::lrint(v); }  // saturate_cast<int>(double): round half to even
static inline int reflect101(int p, int len) {                        // BORDER_REFLECT_101 (borderInterpolate)
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
  return p;
}

// The 2x3 inverse map of Regard3DFeatures.cpp:766-800 (float arithmetic, cos/sin in double)
void liop_affine(float x, float y, float kp_size, float kp_angle, float kpSizeFactor, float* M) {
  const int patchResolution = 20;
  const float angle = -90.0f - kp_angle;
  const float scale = kp_size / (float)kSide * kpSizeFactor;
  const float alpha = (float)(scale * std::cos(angle * M_PI / 180.0f));
  const float beta = (float)(scale * std::sin(angle * M_PI / 180.0f));
  const float trans_x = x - (float)patchResolution;
  const float trans_y = y - (float)patchResolution;
  M[0] = alpha;
  M[1] = beta;
  M[2] = beta * trans_y + alpha * trans_x - beta * y + (1.0f - alpha) * x;
  M[3] = -beta;
  M[4] = alpha;
  M[5] = alpha * trans_y - beta * trans_x + beta * x + (1.0f - alpha) * y;
}

// cv::warpAffine(img, patch, M, Size(41,41), INTER_LINEAR | WARP_INVERSE_MAP) [constant border 0], CV_32F
void liop_warp(const float* img, int w, int h, const float* Mf, float* patch) {
  double M[6];
  for (int i = 0; i < 6; ++i) M[i] = (double)Mf[i];
  const int AB_BITS = 10, AB_SCALE = 1 << AB_BITS, INTER_BITS = 5, INTER_TAB_SIZE = 1 << INTER_BITS;
  const int round_delta = AB_SCALE / INTER_TAB_SIZE / 2;
  for (int dy = 0; dy < kSide; ++dy) {
    const int X0 = cv_round((M[1] * dy + M[2]) * AB_SCALE) + round_delta;
    const int Y0 = cv_round((M[4] * dy + M[5]) * AB_SCALE) + round_delta;
    for (int dx = 0; dx < kSide; ++dx) {
      const int adelta = cv_round(M[0] * dx * AB_SCALE), bdelta = cv_round(M[3] * dx * AB_SCALE);
      const int X = (X0 + adelta) >> (AB_BITS - INTER_BITS), Y = (Y0 + bdelta) >> (AB_BITS - INTER_BITS);
      int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
      sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);  // saturate_cast<short>
      sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
      const int ax = X & (INTER_TAB_SIZE - 1), ay = Y & (INTER_TAB_SIZE - 1);
      // initInterTab2D(INTER_LINEAR): tab[ay][ax][k1][k2] = vy[k1] * vx[k2], v[0] = 1 - t/32, v[1] = t/32 (float)
      const float fx = (float)ax * (1.f / INTER_TAB_SIZE), fy = (float)ay * (1.f / INTER_TAB_SIZE);
      const float vx0 = 1.f - fx, vx1 = fx, vy0 = 1.f - fy, vy1 = fy;
      const float w0 = vy0 * vx0, w1 = vy0 * vx1, w2 = vy1 * vx0, w3 = vy1 * vx1;
      float v;
      if (sx >= w || sx + 1 < 0 || sy >= h || sy + 1 < 0) {
        v = 0.f;
      } else {
        auto at = [&](int xx, int yy) -> float { return (xx >= 0 && xx < w && yy >= 0 && yy < h) ? img[(size_t)yy * w + xx] : 0.f; };
        const float v0 = at(sx, sy), v1 = at(sx + 1, sy), v2 = at(sx, sy + 1), v3 = at(sx + 1, sy + 1);
        v = v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
      }
      patch[dy * kSide + dx] = v;
    }
  }
}

// cv::GaussianBlur(patch, patch, Size(0,0), 1.2): separable 11-tap filter, BORDER_REFLECT_101; canonical (non-FMA)
// evaluation order of filter.cpp: rows tap by tap, columns in the symmetric form k0*c + sum k_j*(up_j + down_j)
void liop_blur(const float* in, float* out) {
  float tmp[kSide * kSide];
  for (int y = 0; y < kSide; ++y)
    for (int x = 0; x < kSide; ++x) {
      float s = in[y * kSide + reflect101(x - 5, kSide)] * kGauss11[0];
      for (int k = 1; k < 11; ++k) s = s + in[y * kSide + reflect101(x - 5 + k, kSide)] * kGauss11[k];
      tmp[y * kSide + x] = s;
    }
  for (int y = 0; y < kSide; ++y)
    for (int x = 0; x < kSide; ++x) {
      float s = kGauss11[5] * tmp[y * kSide + x];
      for (int k = 1; k <= 5; ++k)
        s = s + kGauss11[5 + k] * (tmp[reflect101(y + k, kSide) * kSide + x] + tmp[reflect101(y - k, kSide) * kSide + x]);
      out[y * kSide + x] = s;
    }
}

// Regard3DFeatures::extractLIOPFeatures, per keypoint (x, y, size, angle), in keypoint order (the reference pushes
// results in thread-completion order under `omp critical`, :838-851 -- a nondeterminism not restated)
void liop_describe(const float* img, int w, int h, const float* kps, uint64_t n, float kpSizeFactor, float* desc,
                   float* patches_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; ++i) {
    float M[6], warped[kSide * kSide], patch[kSide * kSide];
    liop_affine(kps[4 * i], kps[4 * i + 1], kps[4 * i + 2], kps[4 * i + 3], kpSizeFactor, M);
    liop_warp(img, w, h, M, warped);
    liop_blur(warped, patch);
    if (patches_out) std::memcpy(patches_out + (size_t)i * kSide * kSide, patch, sizeof(patch));
    liop_process(patch, desc + (size_t)i * 144);
  }
}

}  // namespace orc

extern "C" {
uint32_t orc_liop_patch_size(void) { return orc::liop_patch_size(); }
void orc_liop_process(const float* patch41, float* desc144) { orc::liop_process(patch41, desc144); }
void orc_liop_affine(float x, float y, float size, float angle, float factor, float* M) { orc::liop_affine(x, y, size, angle, factor, M); }
void orc_liop_warp(const float* img, int w, int h, const float* M, float* patch41) { orc::liop_warp(img, w, h, M, patch41); }
void orc_liop_blur(const float* in41, float* out41) { orc::liop_blur(in41, out41); }
void orc_liop_describe(const float* img, int w, int h, const float* kps, uint64_t n, float factor, float* desc, float* patches) {
  orc::liop_describe(img, w, h, kps, n, factor, desc, patches);
}
}
