// =====================================================================================
// linemod_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A scalar C++ restatement of the LINEMOD matching path that the reference ROS package
// birlrobotics/linemod_pose_estimation enters through
//     rgbdDetector::linemod_detection -> cv::linemod::Detector::match   (src/rgbdDetector.cpp:31-34)
// and of the template-extraction / persistence calls it makes
//     Detector::addTemplate  (src/renderer.cpp:308, src/renderer_only_image.cpp:266)
//     Detector::read/readClass (src/rgbdDetector.cpp:1668-1680), write/writeClass (src/renderer.cpp:56-70).
//
// The arithmetic lives in a third-party dependency that is NOT vendored in /root/reference:
// OpenCV 2.4.x  modules/objdetect/src/linemod.cpp (+ normal_lut.i), de-facto pinned to 2.4.8 by the
// Ubuntu-14.04 / ROS-Indigo toolchain the reference builds on (CMakeLists.txt:22,188).  It cannot be
// compiled here (no OpenCV C++ in the image), and the reference ships no tests or golden vectors.
//
//                      *** PARITY UNPINNED ***
// This file follows the published algorithm as written down in SURVEY.md Appendix A ("[OCV]" = the
// upstream function each routine restates).  The imgproc primitives it relies on (GaussianBlur 7x7,
// Sobel 3x3, phase/fastAtan2, pyrDown, medianBlur 5, NN resize, erode, distanceTransform) are pinned
// against the in-container cv2 4.13 build by tests/test_oracle_primitives.py and tests/golden/.
// SIMILARITY_LUT and NORMAL_LUT are recalled/generated data and therefore injectable.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  The product (linemod_pose_estimation_b200/) never links or calls it.
//
// Build: see oracle/Makefile  (g++ -O3 -msse4.1 -ffp-contract=off -pthread -shared -fPIC)
// =====================================================================================
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include <emmintrin.h>
#include <smmintrin.h>
#include <tmmintrin.h>

namespace {

// ----------------------------------------------------------------------------- data model (A.1)
struct Feature { int x, y, label; };
struct Template {
  int width = 0, height = 0, pyramid_level = 0;
  std::vector<Feature> features;
};
typedef std::vector<Template> TemplatePyramid;  // index l*M + m

enum ModalityType { MOD_COLOR_GRADIENT = 0, MOD_DEPTH_NORMAL = 1 };

struct ModalityDesc {
  int type;
  // ColorGradient: weak_threshold, num_features, strong_threshold      ([OCV] ColorGradient::ColorGradient)
  float weak_threshold = 10.0f;
  float strong_threshold = 55.0f;
  // DepthNormal: distance_threshold, difference_threshold, num_features, extract_threshold
  int distance_threshold = 2000;
  int difference_threshold = 50;
  int extract_threshold = 2;
  int num_features = 63;
};

struct MatchRec {  // layout shared with the python binding
  int32_t x, y, template_id, class_index;
  float similarity;
};

struct CandRec {  // coarse candidates (pre-refinement), for the parity taps
  int32_t class_index, template_id, pos, raw;
};

struct RawRec {  // survivor in the product's exchange format (lm_raw_match): integer score + feature count
  uint32_t order_key, coarse_pos;
  int32_t x, y;
  uint32_t score, nf;
  int32_t template_id, class_index;
};

// [OCV] Match::operator< / operator==  (A.1)
static inline bool match_less(const MatchRec& a, const MatchRec& b) {
  if (a.similarity != b.similarity) return a.similarity > b.similarity;
  return a.template_id < b.template_id;
}
static inline bool match_equal(const MatchRec& a, const MatchRec& b) {
  return a.x == b.x && a.y == b.y && a.similarity == b.similarity && a.class_index == b.class_index;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
  return p;
}

// ----------------------------------------------------------------------------- imgproc primitives
// cv::GaussianBlur(src, 7x7, sigma 0, BORDER_REPLICATE) on 8UC3 (A.2-1): fixed-point separable filter,
// taps {8,28,56,72,56,28,8}/256 per pass, one rounding at the end: (sum + 2^15) >> 16.
static void gaussian7_u8(const uint8_t* src, int rows, int cols, int ch, uint8_t* dst) {
  static const int k[7] = {8, 28, 56, 72, 56, 28, 8};
  std::vector<int> tmp((size_t)rows * cols * ch);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x)
      for (int c = 0; c < ch; ++c) {
        int s = 0;
        for (int i = 0; i < 7; ++i) s += k[i] * src[((size_t)y * cols + clampi(x + i - 3, 0, cols - 1)) * ch + c];
        tmp[((size_t)y * cols + x) * ch + c] = s;
      }
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x)
      for (int c = 0; c < ch; ++c) {
        int s = 0;
        for (int j = 0; j < 7; ++j) s += k[j] * tmp[((size_t)clampi(y + j - 3, 0, rows - 1) * cols + x) * ch + c];
        dst[((size_t)y * cols + x) * ch + c] = (uint8_t)((s + 32768) >> 16);
      }
}

// cv::Sobel(src, CV_16S, dx,dy, ksize 3, BORDER_REPLICATE) (A.2-2)
static void sobel3_u8(const uint8_t* src, int rows, int cols, int ch, int16_t* dx, int16_t* dy) {
  for (int y = 0; y < rows; ++y) {
    int ym = clampi(y - 1, 0, rows - 1), yp = clampi(y + 1, 0, rows - 1);
    for (int x = 0; x < cols; ++x) {
      int xm = clampi(x - 1, 0, cols - 1), xp = clampi(x + 1, 0, cols - 1);
      for (int c = 0; c < ch; ++c) {
#define P(yy, xx) (int)src[((size_t)(yy)*cols + (xx)) * ch + c]
        int gx = (P(ym, xp) + 2 * P(y, xp) + P(yp, xp)) - (P(ym, xm) + 2 * P(y, xm) + P(yp, xm));
        int gy = (P(yp, xm) + 2 * P(yp, x) + P(yp, xp)) - (P(ym, xm) + 2 * P(ym, x) + P(ym, xp));
#undef P
        dx[((size_t)y * cols + x) * ch + c] = (int16_t)gx;
        dy[((size_t)y * cols + x) * ch + c] = (int16_t)gy;
      }
    }
  }
}

// cv::fastAtan2 as used by cv::phase(..., angleInDegrees=true), non-FMA evaluation (A.2-4).
static inline float fast_atan2_deg(float y, float x) {
  static const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
  static const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
  static const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
  static const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
  float ax = std::fabs(x), ay = std::fabs(y);
  float mn = ax < ay ? ax : ay, mx = ax < ay ? ay : ax;
  volatile float den = mx + (float)DBL_EPSILON;  // volatile: keep every intermediate an f32 rounding
  float c = mn / den;
  float c2 = c * c;
  float a = p7 * c2;
  a = (a + p5) * c2;
  a = (a + p3) * c2;
  a = (a + p1) * c;
  if (ax < ay) a = 90.f - a;
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// cv::pyrDown on 8UC<ch>: 5x5 [1 4 6 4 1]^2, (sum + 128) >> 8, BORDER_REFLECT_101, even samples (A.2-7)
static void pyrdown_u8(const uint8_t* src, int rows, int cols, int ch, uint8_t* dst) {
  static const int k[5] = {1, 4, 6, 4, 1};
  int orows = rows / 2, ocols = cols / 2;
  for (int y = 0; y < orows; ++y)
    for (int x = 0; x < ocols; ++x)
      for (int c = 0; c < ch; ++c) {
        int s = 0;
        for (int j = 0; j < 5; ++j) {
          int sy = reflect101(2 * y + j - 2, rows);
          int rs = 0;
          for (int i = 0; i < 5; ++i) rs += k[i] * src[((size_t)sy * cols + reflect101(2 * x + i - 2, cols)) * ch + c];
          s += k[j] * rs;
        }
        dst[((size_t)y * ocols + x) * ch + c] = (uint8_t)((s + 128) >> 8);
      }
}

// cv::resize(..., INTER_NN) to exactly half size: dst(y,x) = src(2y,2x)  (A.3 pyramid)
static void nn_half_u8(const uint8_t* src, int rows, int cols, uint8_t* dst) {
  int orows = rows / 2, ocols = cols / 2;
  for (int y = 0; y < orows; ++y)
    for (int x = 0; x < ocols; ++x) dst[(size_t)y * ocols + x] = src[(size_t)(2 * y) * cols + 2 * x];
}

// cv::medianBlur(src, dst, 5) on 8UC1, replicate border (A.3)
static void median5_u8(const uint8_t* src, int rows, int cols, uint8_t* dst) {
  uint8_t w[25];
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      int n = 0;
      for (int j = -2; j <= 2; ++j) {
        const uint8_t* row = src + (size_t)clampi(y + j, 0, rows - 1) * cols;
        for (int i = -2; i <= 2; ++i) w[n++] = row[clampi(x + i, 0, cols - 1)];
      }
      std::nth_element(w, w + 12, w + 25);
      dst[(size_t)y * cols + x] = w[12];
    }
}

// cv::erode(src, dst, Mat(), Point(-1,-1), iterations, BORDER_REPLICATE): 3x3 rectangular minimum (A.11)
static void erode3_u8(const uint8_t* src, int rows, int cols, int iterations, uint8_t* dst) {
  std::vector<uint8_t> a(src, src + (size_t)rows * cols), b((size_t)rows * cols);
  for (int it = 0; it < iterations; ++it) {
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) {
        uint8_t m = 255;
        for (int j = -1; j <= 1; ++j)
          for (int i = -1; i <= 1; ++i)
            m = std::min(m, a[(size_t)clampi(y + j, 0, rows - 1) * cols + clampi(x + i, 0, cols - 1)]);
        b[(size_t)y * cols + x] = m;
      }
    a.swap(b);
  }
  std::memcpy(dst, a.data(), (size_t)rows * cols);
}

// cv::distanceTransform(src, dst, CV_DIST_C, 3): two-pass 3x3 chamfer, HV = DIAG = 1 (fixed point 2^16),
// image border initialised to INT_MAX>>2 (i.e. "no zero pixel outside the image")  (A.11)
static void distance_transform_c3(const uint8_t* src, int rows, int cols, float* dst) {
  const int SHIFT = 16, ONE = 1 << SHIFT, INIT = INT_MAX >> 2;
  int tc = cols + 2;
  std::vector<int> tmp((size_t)(rows + 2) * tc, INIT);
  for (int y = 0; y < rows; ++y) {
    int* t = &tmp[(size_t)(y + 1) * tc + 1];
    for (int x = 0; x < cols; ++x) {
      if (!src[(size_t)y * cols + x]) {
        t[x] = 0;
      } else {
        int t0 = t[x - tc - 1] + ONE;
        int v = t[x - tc] + ONE;
        if (t0 > v) t0 = v;
        v = t[x - tc + 1] + ONE;
        if (t0 > v) t0 = v;
        v = t[x - 1] + ONE;
        if (t0 > v) t0 = v;
        t[x] = t0;
      }
    }
  }
  const float scale = 1.f / ONE;
  for (int y = rows - 1; y >= 0; --y) {
    int* t = &tmp[(size_t)(y + 1) * tc + 1];
    for (int x = cols - 1; x >= 0; --x) {
      int t0 = t[x];
      if (t0 > ONE) {
        int v = t[x + tc + 1] + ONE;
        if (t0 > v) t0 = v;
        v = t[x + tc] + ONE;
        if (t0 > v) t0 = v;
        v = t[x + tc - 1] + ONE;
        if (t0 > v) t0 = v;
        v = t[x + 1] + ONE;
        if (t0 > v) t0 = v;
        t[x] = t0;
      }
      dst[(size_t)y * cols + x] = (float)t0 * scale;
    }
  }
}

// ----------------------------------------------------------------------------- ColorGradient (A.2)
// [OCV] hysteresisGradient
static void hysteresis_gradient(const float* magnitude, const float* angle, int rows, int cols, float threshold,
                                uint8_t* quantized) {
  std::vector<uint8_t> q((size_t)rows * cols);
  const float scale = (float)(16.0 / 360.0);
  for (size_t i = 0; i < q.size(); ++i) {
    long r = lrintf(angle[i] * scale);  // cvRound: round-half-even, then saturate_cast<uchar>
    q[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
  }
  std::memset(&q[0], 0, cols);
  std::memset(&q[(size_t)(rows - 1) * cols], 0, cols);
  for (int r = 0; r < rows; ++r) q[(size_t)r * cols] = q[(size_t)r * cols + cols - 1] = 0;
  for (int r = 1; r < rows - 1; ++r)
    for (int c = 1; c < cols - 1; ++c) q[(size_t)r * cols + c] &= 7;

  std::memset(quantized, 0, (size_t)rows * cols);
  for (int r = 1; r < rows - 1; ++r)
    for (int c = 1; c < cols - 1; ++c) {
      if (magnitude[(size_t)r * cols + c] > threshold) {
        int hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = -1; j <= 1; ++j)
          for (int i = -1; i <= 1; ++i) hist[q[(size_t)(r + j) * cols + c + i]]++;
        int max_votes = 0, index = -1;
        for (int i = 0; i < 8; ++i)
          if (max_votes < hist[i]) { index = i; max_votes = hist[i]; }
        if (max_votes >= 5) quantized[(size_t)r * cols + c] = (uint8_t)(1 << index);
      }
    }
}

// [OCV] quantizedOrientations
static void quantized_orientations(const uint8_t* bgr, int rows, int cols, float weak_threshold, float* magnitude,
                                   uint8_t* quantized, float* angle_out /*nullable*/) {
  size_t n = (size_t)rows * cols;
  std::vector<uint8_t> smoothed(n * 3);
  gaussian7_u8(bgr, rows, cols, 3, smoothed.data());
  std::vector<int16_t> dx(n * 3), dy(n * 3);
  sobel3_u8(smoothed.data(), rows, cols, 3, dx.data(), dy.data());
  std::vector<float> angle(n);
  for (size_t i = 0; i < n; ++i) {
    int m0 = dx[3 * i] * dx[3 * i] + dy[3 * i] * dy[3 * i];
    int m1 = dx[3 * i + 1] * dx[3 * i + 1] + dy[3 * i + 1] * dy[3 * i + 1];
    int m2 = dx[3 * i + 2] * dx[3 * i + 2] + dy[3 * i + 2] * dy[3 * i + 2];
    int sel, mag;
    if (m0 >= m1 && m0 >= m2) { sel = 0; mag = m0; }
    else if (m1 >= m0 && m1 >= m2) { sel = 1; mag = m1; }
    else { sel = 2; mag = m2; }
    magnitude[i] = (float)mag;
    angle[i] = fast_atan2_deg((float)dy[3 * i + sel], (float)dx[3 * i + sel]);
  }
  if (angle_out) std::memcpy(angle_out, angle.data(), n * sizeof(float));
  hysteresis_gradient(magnitude, angle.data(), rows, cols, weak_threshold * weak_threshold, quantized);
}

// ----------------------------------------------------------------------------- DepthNormal (A.3)
// [OCV] quantizedNormals (before the median filter)
static void quantized_normals_raw(const uint16_t* depth, int rows, int cols, int distance_threshold,
                                  int difference_threshold, const uint8_t* normal_lut, uint8_t* dst) {
  std::memset(dst, 0, (size_t)rows * cols);
  const int r = 5;
  static const int OI[8] = {-r, 0, +r, -r, +r, -r, 0, +r};
  static const int OJ[8] = {-r, -r, -r, 0, 0, +r, +r, +r};
  for (int y = r; y < rows - r - 1; ++y)
    for (int x = r; x < cols - r - 1; ++x) {
      long d = depth[(size_t)y * cols + x];
      uint8_t out = 0;
      if (d < distance_threshold) {
        long A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
        for (int k = 0; k < 8; ++k) {
          long i = OI[k], j = OJ[k];
          long delta = (long)depth[(size_t)(y + j) * cols + (x + i)] - d;
          long f = std::labs(delta) < difference_threshold ? 1 : 0;
          long fi = f * i, fj = f * j;
          A0 += fi * i; A1 += fi * j; A3 += fj * j;
          b0 += fi * delta; b1 += fj * delta;
        }
        long det = A0 * A3 - A1 * A1;
        long ddx = A3 * b0 - A1 * b1;
        long ddy = -A1 * b0 + A0 * b1;
        float nx = (float)(1150 * ddx);
        float ny = (float)(1150 * ddy);
        float nz = (float)(-det * d);
        volatile float xx = nx * nx, yy = ny * ny, zz = nz * nz;  // no contraction, left-to-right
        volatile float s2 = xx + yy;
        s2 = s2 + zz;
        float s = sqrtf(s2);
        if (s > 0) {
          float inv = 1.0f / s;
          nx *= inv; ny *= inv; nz *= inv;
          volatile float t1 = nx * 10.0f, t2 = ny * 10.0f, t3 = nz * 20.0f;
          int v1 = (int)(t1 + 10.0f);
          int v2 = (int)(t2 + 10.0f);
          int v3 = (int)(t3 + 20.0f);
          int flat = (v3 * 20 + v2) * 20 + v1;  // NORMAL_LUT[v3][v2][v1]; out-of-table (upstream UB) -> 0
          out = (flat >= 0 && flat < 8000) ? normal_lut[flat] : 0;
        }
      }
      dst[(size_t)y * cols + x] = out;
    }
}

// ----------------------------------------------------------------------------- spread / response / linearize
// [OCV] spread (A.4)
static void spread_T(const uint8_t* src, int rows, int cols, int T, uint8_t* dst) {
  std::memset(dst, 0, (size_t)rows * cols);
  for (int r = 0; r < T; ++r)
    for (int c = 0; c < T; ++c)
      for (int y = 0; y < rows - r; ++y) {
        const uint8_t* s = src + (size_t)(y + r) * cols + c;
        uint8_t* d = dst + (size_t)y * cols;
        for (int x = 0; x < cols - c; ++x) d[x] |= s[x];
      }
}

// [OCV] computeResponseMaps (A.5)
static void response_maps(const uint8_t* spread, size_t n, const uint8_t* lut, uint8_t* resp /*8*n*/) {
  for (int ori = 0; ori < 8; ++ori) {
    const uint8_t* lo = lut + 32 * ori;
    const uint8_t* hi = lo + 16;
    uint8_t* out = resp + ori * n;
    for (size_t i = 0; i < n; ++i) out[i] = std::max(lo[spread[i] & 15], hi[spread[i] >> 4]);
  }
}

// Flat layout of one orientation's linear memories: T*T rows of W*H bytes, followed by a zero tail so that every
// read the reference can perform from an in-bounds feature stays inside defined memory.
static inline size_t lm_plane_stride(int T, int W, int H) {
  size_t wh = (size_t)W * H;
  return ((size_t)T * T * wh + wh + 16 * (size_t)W + 16 + 15) & ~(size_t)15;  // planes start 16-byte aligned
}

// [OCV] linearize (A.6)
static void linearize_T(const uint8_t* resp, int rows, int cols, int T, uint8_t* plane) {
  int W = cols / T, H = rows / T;
  size_t idx = 0;
  for (int rs = 0; rs < T; ++rs)
    for (int cs = 0; cs < T; ++cs)
      for (int r = rs; r < rows; r += T)
        for (int c = cs; c < cols; c += T) plane[idx++] = resp[(size_t)r * cols + c];
  (void)W; (void)H;
}

// ----------------------------------------------------------------------------- detector state
class BandPool;
struct FastScratch {  // working buffers of the baseline front end (linemod_fast.inc), kept between frames
  std::vector<uint8_t> img, next, sm, qbin, raw, mask;
  std::vector<uint16_t> depth;
};

struct LevelMod {  // per (level, modality) products of the front end, kept for the parity taps
  int rows = 0, cols = 0, T = 0, W = 0, H = 0;
  std::vector<uint8_t> quant_raw;  // unmasked quantisation (CG "angle" / DN "normal")
  std::vector<uint8_t> quantized;  // after mask
  std::vector<float> magnitude;    // CG only
  std::vector<uint8_t> spread;
  std::vector<uint8_t> response;   // 8 * rows*cols
  std::vector<uint8_t> lm;         // 8 * plane_stride
  size_t plane_stride = 0;
};

struct Detector {
  std::vector<int> T;
  std::vector<ModalityDesc> mods;
  std::map<std::string, std::vector<TemplatePyramid> > classes;
  uint8_t sim_lut[256];
  uint8_t normal_lut[8000];
  int threads = 1;
  bool fast = false;   // front end by linemod_fast.inc (the CPU baseline's SSE / threaded routines) instead of the plain code
  std::shared_ptr<BandPool> pool;
  std::vector<FastScratch> scratch;
  // last frame
  std::vector<LevelMod> front;  // index l*M+m
  std::vector<CandRec> last_cands;
  std::vector<MatchRec> last_presort;
  std::vector<RawRec> last_raw;
  std::string err;
  int levels() const { return (int)T.size(); }
  int M() const { return (int)mods.size(); }
};

// Default SIMILARITY_LUT ([OCV] linemod.cpp, literal table; see DESIGN.md "LUTs" for the recall caveat):
// LUT[32*i + 16*h + v] = max over set bits b of v (j = 4h+b) of max(0, 4 - |i - j|).
static void default_similarity_lut(uint8_t* lut) {
  for (int i = 0; i < 8; ++i)
    for (int h = 0; h < 2; ++h)
      for (int v = 0; v < 16; ++v) {
        int best = 0;
        for (int b = 0; b < 4; ++b)
          if (v & (1 << b)) best = std::max(best, std::max(0, 4 - std::abs(i - (4 * h + b))));
        lut[32 * i + 16 * h + v] = (uint8_t)best;
      }
}

// Default NORMAL_LUT stand-in (upstream normal_lut.i is not recoverable): azimuthal 8-sector code.
static void default_normal_lut(uint8_t* lut) {
  for (int v3 = 0; v3 < 20; ++v3)
    for (int v2 = 0; v2 < 20; ++v2)
      for (int v1 = 0; v1 < 20; ++v1) {
        double ang = std::atan2((double)(v2 - 10), (double)(v1 - 10)) * 180.0 / 3.14159265358979323846;
        int s = (int)std::lround(ang / 45.0);
        s = ((s % 8) + 8) % 8;
        lut[(v3 * 20 + v2) * 20 + v1] = (uint8_t)(1 << s);
      }
}

// ----------------------------------------------------------------------------- quantised pyramids
struct Source {
  const uint8_t* data; int rows, cols, type; size_t step;  // type: 0 = 8UC3, 1 = 16UC1, 2 = 8UC1
};

struct QuantPyr {  // [OCV] ColorGradientPyramid / DepthNormalPyramid
  ModalityDesc desc;
  int rows = 0, cols = 0, level = 0;
  int num_features = 0, extract_threshold = 0;
  std::vector<uint8_t> src;        // CG: BGR image at this level
  std::vector<uint8_t> mask;       // empty = no mask
  std::vector<float> magnitude;    // CG
  std::vector<uint8_t> quant;      // CG angle (quantised) / DN normal
};

static void cg_update(QuantPyr& q) {
  size_t n = (size_t)q.rows * q.cols;
  q.magnitude.resize(n);
  q.quant.resize(n);
  quantized_orientations(q.src.data(), q.rows, q.cols, q.desc.weak_threshold, q.magnitude.data(), q.quant.data(),
                         nullptr);
}

// [OCV] Modality::process -> ColorGradientPyramid / DepthNormalPyramid constructors
static bool pyr_process(const Detector& det, const ModalityDesc& desc, const Source& s, const Source* mask,
                        QuantPyr& q, std::string& err) {
  q.desc = desc;
  q.rows = s.rows; q.cols = s.cols; q.level = 0;
  q.num_features = desc.num_features;
  q.extract_threshold = desc.extract_threshold;
  size_t n = (size_t)s.rows * s.cols;
  if (mask && mask->data) {
    if (mask->rows != s.rows || mask->cols != s.cols || mask->type != 2) { err = "mask size/type mismatch"; return false; }
    q.mask.resize(n);
    for (int y = 0; y < s.rows; ++y) std::memcpy(&q.mask[(size_t)y * s.cols], mask->data + y * mask->step, s.cols);
  }
  if (desc.type == MOD_COLOR_GRADIENT) {
    if (s.type != 0) { err = "ColorGradient needs an 8UC3 source"; return false; }
    q.src.resize(n * 3);
    for (int y = 0; y < s.rows; ++y) std::memcpy(&q.src[(size_t)y * s.cols * 3], s.data + y * s.step, (size_t)s.cols * 3);
    cg_update(q);
  } else {
    if (s.type != 1) { err = "DepthNormal needs a 16UC1 source"; return false; }
    std::vector<uint16_t> depth(n);
    for (int y = 0; y < s.rows; ++y) std::memcpy(&depth[(size_t)y * s.cols], s.data + y * s.step, (size_t)s.cols * 2);
    std::vector<uint8_t> raw(n);
    quantized_normals_raw(depth.data(), s.rows, s.cols, desc.distance_threshold, desc.difference_threshold,
                          det.normal_lut, raw.data());
    q.quant.resize(n);
    median5_u8(raw.data(), s.rows, s.cols, q.quant.data());
  }
  return true;
}

// [OCV] ColorGradientPyramid::pyrDown / DepthNormalPyramid::pyrDown
static void pyr_down(QuantPyr& q) {
  int orows = q.rows / 2, ocols = q.cols / 2;
  q.num_features /= 2;
  ++q.level;
  if (q.desc.type == MOD_COLOR_GRADIENT) {
    std::vector<uint8_t> next((size_t)orows * ocols * 3);
    pyrdown_u8(q.src.data(), q.rows, q.cols, 3, next.data());
    q.src.swap(next);
  } else {
    q.extract_threshold /= 2;
    std::vector<uint8_t> next((size_t)orows * ocols);
    nn_half_u8(q.quant.data(), q.rows, q.cols, next.data());
    q.quant.swap(next);
  }
  if (!q.mask.empty()) {
    std::vector<uint8_t> nm((size_t)orows * ocols);
    nn_half_u8(q.mask.data(), q.rows, q.cols, nm.data());
    q.mask.swap(nm);
  }
  q.rows = orows; q.cols = ocols;
  if (q.desc.type == MOD_COLOR_GRADIENT) cg_update(q);
}

// [OCV] QuantizedPyramid::quantize: dst = zeros; quant.copyTo(dst, mask)
static void pyr_quantize(const QuantPyr& q, std::vector<uint8_t>& dst) {
  dst = q.quant;
  if (!q.mask.empty())
    for (size_t i = 0; i < dst.size(); ++i)
      if (!q.mask[i]) dst[i] = 0;
}

// ----------------------------------------------------------------------------- template extraction (A.11)
struct Candidate { Feature f; float score; };
static inline bool cand_less(const Candidate& a, const Candidate& b) { return a.score > b.score; }

static inline int get_label(int quantized) {
  switch (quantized) {
    case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3;
    case 16: return 4; case 32: return 5; case 64: return 6; case 128: return 7;
    default: return -1;
  }
}

// [OCV] QuantizedPyramid::selectScatteredFeatures
static void select_scattered(const std::vector<Candidate>& cands, std::vector<Feature>& feats, size_t num_features,
                             float distance) {
  feats.clear();
  float distance_sq = distance * distance;
  int i = 0;
  while (feats.size() < num_features) {
    const Candidate& c = cands[i];
    bool keep = true;
    for (int j = 0; j < (int)feats.size() && keep; ++j) {
      const Feature& f = feats[j];
      keep = (float)((c.f.x - f.x) * (c.f.x - f.x) + (c.f.y - f.y) * (c.f.y - f.y)) >= distance_sq;
    }
    if (keep) feats.push_back(c.f);
    if (++i == (int)cands.size()) {
      i = 0;
      distance -= 1.0f;
      distance_sq = distance * distance;
    }
  }
}

// [OCV] ColorGradientPyramid::extractTemplate (2.4.x: features restricted to the 1-px silhouette ring)
static bool cg_extract(const QuantPyr& q, Template& t) {
  std::vector<uint8_t> local;
  bool no_mask = q.mask.empty();
  size_t n = (size_t)q.rows * q.cols;
  if (!no_mask) {
    local.resize(n);
    erode3_u8(q.mask.data(), q.rows, q.cols, 1, local.data());
    for (size_t i = 0; i < n; ++i) {
      int d = (int)q.mask[i] - (int)local[i];
      local[i] = (uint8_t)(d < 0 ? 0 : d);
    }
  }
  std::vector<Candidate> cands;
  float thr_sq = q.desc.strong_threshold * q.desc.strong_threshold;
  for (int r = 0; r < q.rows; ++r)
    for (int c = 0; c < q.cols; ++c) {
      size_t i = (size_t)r * q.cols + c;
      if (no_mask || local[i]) {
        uint8_t qv = q.quant[i];
        if (qv > 0) {
          float score = q.magnitude[i];
          if (score > thr_sq) {
            Candidate cd; cd.f.x = c; cd.f.y = r; cd.f.label = get_label(qv); cd.score = score;
            cands.push_back(cd);
          }
        }
      }
    }
  if (cands.size() < (size_t)q.num_features) return false;
  std::stable_sort(cands.begin(), cands.end(), cand_less);
  float distance = (float)(cands.size() / (size_t)q.num_features + 1);
  select_scattered(cands, t.features, q.num_features, distance);
  t.width = -1; t.height = -1; t.pyramid_level = q.level;
  return true;
}

// [OCV] DepthNormalPyramid::extractTemplate
static bool dn_extract(const QuantPyr& q, Template& t) {
  std::vector<uint8_t> local;
  bool no_mask = q.mask.empty();
  size_t n = (size_t)q.rows * q.cols;
  if (!no_mask) {
    local.resize(n);
    erode3_u8(q.mask.data(), q.rows, q.cols, 2, local.data());
  }
  std::vector<uint8_t> temp(n, 0);
  std::vector<float> dist[8];
  for (int i = 0; i < 8; ++i) {
    for (size_t k = 0; k < n; ++k) {
      if (no_mask || local[k]) temp[k] = (uint8_t)(1 << i);
      temp[k] &= q.quant[k];
    }
    dist[i].resize(n);
    distance_transform_c3(temp.data(), q.rows, q.cols, dist[i].data());
  }
  int label_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<Candidate> cands;
  for (int r = 0; r < q.rows; ++r)
    for (int c = 0; c < q.cols; ++c) {
      size_t i = (size_t)r * q.cols + c;
      if (no_mask || local[i]) {
        uint8_t qv = q.quant[i];
        if (qv != 0 && qv != 255) {
          int label = get_label(qv);
          if (label < 0) continue;  // cannot happen with one-hot LUT data
          float score = dist[label][i];
          if (score >= (float)q.extract_threshold) {
            Candidate cd; cd.f.x = c; cd.f.y = r; cd.f.label = label; cd.score = score;
            cands.push_back(cd);
            ++label_counts[label];
          }
        }
      }
    }
  if (cands.size() < (size_t)q.num_features) return false;
  for (size_t i = 0; i < cands.size(); ++i) cands[i].score /= (float)label_counts[cands[i].f.label];
  std::stable_sort(cands.begin(), cands.end(), cand_less);
  float area;
  if (no_mask) area = (float)n;
  else { size_t nz = 0; for (size_t k = 0; k < n; ++k) nz += local[k] != 0; area = (float)nz; }
  float distance = sqrtf(area) / sqrtf((float)q.num_features) + 1.5f;
  select_scattered(cands, t.features, q.num_features, distance);
  t.width = -1; t.height = -1; t.pyramid_level = q.level;
  return true;
}

// [OCV] cropTemplates
static void crop_templates(TemplatePyramid& tp, int bb[4]) {
  int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
  for (size_t i = 0; i < tp.size(); ++i)
    for (size_t j = 0; j < tp[i].features.size(); ++j) {
      int x = tp[i].features[j].x << tp[i].pyramid_level;
      int y = tp[i].features[j].y << tp[i].pyramid_level;
      min_x = std::min(min_x, x); min_y = std::min(min_y, y);
      max_x = std::max(max_x, x); max_y = std::max(max_y, y);
    }
  if (min_x % 2 == 1) --min_x;
  if (min_y % 2 == 1) --min_y;
  for (size_t i = 0; i < tp.size(); ++i) {
    Template& t = tp[i];
    t.width = (max_x - min_x) >> t.pyramid_level;
    t.height = (max_y - min_y) >> t.pyramid_level;
    int ox = min_x >> t.pyramid_level, oy = min_y >> t.pyramid_level;
    for (size_t j = 0; j < t.features.size(); ++j) { t.features[j].x -= ox; t.features[j].y -= oy; }
  }
  bb[0] = min_x; bb[1] = min_y; bb[2] = max_x - min_x; bb[3] = max_y - min_y;
}

// ----------------------------------------------------------------------------- matching (A.7 - A.10)
// [OCV] accessLinearMemory, against the flat plane layout
static inline const uint8_t* access_lm(const LevelMod& lm, const Feature& f) {
  int T = lm.T;
  size_t grid = (size_t)(f.y % T) * T + (f.x % T);
  return lm.lm.data() + (size_t)f.label * lm.plane_stride + grid * ((size_t)lm.W * lm.H) + (size_t)(f.y / T) * lm.W + f.x / T;
}

// [OCV] similarity: dst is H*W u8, zero-initialised by the caller
static void similarity(const LevelMod& lm, const Template& t, uint8_t* dst) {
  int T = lm.T, W = lm.W, H = lm.H;
  int wf = (t.width - 1) / T + 1, hf = (t.height - 1) / T + 1;
  int span_x = W - wf, span_y = H - hf;
  int P = span_y * W + span_x + 1;
  for (size_t i = 0; i < t.features.size(); ++i) {
    const Feature& f = t.features[i];
    if (f.x < 0 || f.x >= lm.cols || f.y < 0 || f.y >= lm.rows) continue;
    const uint8_t* p = access_lm(lm, f);
    // [OCV] "dst[j] += lm_ptr[j]" with _mm_add_epi8 over unaligned 16-byte blocks (upstream's HAVE_SSE2 branch), scalar tail
    int j = 0;
    for (; j + 16 <= P; j += 16)
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + j),
                       _mm_add_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(dst + j)),
                                    _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + j))));
    for (; j < P; ++j) dst[j] = (uint8_t)(dst[j] + p[j]);
  }
}

// [OCV] similarityLocal: dst is 16x16 u8, zero-initialised by the caller
static void similarity_local(const LevelMod& lm, const Template& t, uint8_t* dst, int cx, int cy) {
  int T = lm.T, W = lm.W;
  int offset_x = (cx / T - 8) * T, offset_y = (cy / T - 8) * T;
  for (size_t i = 0; i < t.features.size(); ++i) {
    Feature f = t.features[i];
    f.x += offset_x; f.y += offset_y;
    if (f.x < 0 || f.y < 0 || f.x >= lm.cols || f.y >= lm.rows) continue;
    const uint8_t* p = access_lm(lm, f);
    for (int row = 0; row < 16; ++row)
      for (int col = 0; col < 16; ++col) dst[row * 16 + col] = (uint8_t)(dst[row * 16 + col] + p[(size_t)row * W + col]);
  }
}

// [OCV] Detector::matchClass body for one template.  Appends surviving matches to `out`
// and (optionally) the coarse candidates to `cands`.
static void match_template(const Detector& det, int class_index, int template_id, const TemplatePyramid& tp,
                           float threshold, std::vector<MatchRec>& out, std::vector<CandRec>* cands,
                           uint16_t* coarse_map /*nullable, H*W*/, std::vector<RawRec>* raws = nullptr,
                           uint32_t order_key = 0) {
  const int L = det.levels(), M = det.M();
  const int lowest_start = (int)tp.size() - M;
  const int lowest_T = det.T.back();
  const LevelMod& l0 = det.front[(size_t)(L - 1) * M];
  const int W = l0.W, H = l0.H;
  std::vector<uint16_t> total((size_t)W * H, 0);
  std::vector<uint8_t> sim((size_t)W * H);
  int num_features = 0;
  for (int m = 0; m < M; ++m) {
    const Template& t = tp[lowest_start + m];
    if (t.features.size() > 63) continue;  // asserted at insertion time
    num_features += (int)t.features.size();
    std::fill(sim.begin(), sim.end(), 0);
    similarity(det.front[(size_t)(L - 1) * M + m], t, sim.data());
    for (size_t j = 0; j < total.size(); ++j) total[j] = (uint16_t)(total[j] + sim[j]);
  }
  if (coarse_map) std::memcpy(coarse_map, total.data(), total.size() * 2);

  int raw_threshold = (int)(2 * num_features + (threshold / 100.f) * (2 * num_features) + 0.5f);
  std::vector<MatchRec> candidates;
  std::vector<RawRec> rr;  // integer view of `candidates`, kept in lock-step
  for (int r = 0; r < H; ++r)
    for (int c = 0; c < W; ++c) {
      int raw = total[(size_t)r * W + c];
      if (raw > raw_threshold) {
        int offset = lowest_T / 2 + (lowest_T % 2 - 1);
        MatchRec mr;
        mr.x = c * lowest_T + offset;
        mr.y = r * lowest_T + offset;
        mr.similarity = (raw * 100.f) / (4 * num_features) + 0.5f;
        mr.class_index = class_index;
        mr.template_id = template_id;
        candidates.push_back(mr);
        RawRec q = {order_key, (uint32_t)(r * W + c), mr.x, mr.y, (uint32_t)raw, (uint32_t)num_features, template_id, class_index};
        rr.push_back(q);
        if (cands) { CandRec cr = {class_index, template_id, r * W + c, raw}; cands->push_back(cr); }
      }
    }

  for (int l = L - 2; l >= 0; --l) {
    int T = det.T[l];
    int start = l * M;
    const LevelMod& lv = det.front[(size_t)l * M];
    int border = 8 * T;
    int offset = T / 2 + (T % 2 - 1);
    int max_x = lv.cols - tp[start].width - border;
    int max_y = lv.rows - tp[start].height - border;
    uint8_t s8[256];
    uint16_t tot[256];
    for (size_t k = 0; k < candidates.size(); ++k) {
      MatchRec& m2 = candidates[k];
      int x = m2.x * 2 + 1, y = m2.y * 2 + 1;
      x = std::max(x, border); y = std::max(y, border);
      x = std::min(x, max_x); y = std::min(y, max_y);
      int nf = 0;
      std::memset(tot, 0, sizeof(tot));
      for (int m = 0; m < M; ++m) {
        const Template& t = tp[start + m];
        nf += (int)t.features.size();
        std::memset(s8, 0, sizeof(s8));
        similarity_local(det.front[(size_t)l * M + m], t, s8, x, y);
        for (int j = 0; j < 256; ++j) tot[j] = (uint16_t)(tot[j] + s8[j]);
      }
      int best_score = 0, best_r = -1, best_c = -1;
      for (int r = 0; r < 16; ++r)
        for (int c = 0; c < 16; ++c) {
          int sc = tot[r * 16 + c];
          if (sc > best_score) { best_score = sc; best_r = r; best_c = c; }
        }
      m2.x = (x / T - 8 + best_c) * T + offset;
      m2.y = (y / T - 8 + best_r) * T + offset;
      m2.similarity = (best_score * 100.f) / (4 * nf);
      rr[k].x = m2.x; rr[k].y = m2.y; rr[k].score = (uint32_t)best_score; rr[k].nf = (uint32_t)nf;
    }
    size_t w = 0;
    for (size_t k = 0; k < candidates.size(); ++k)
      if (!(candidates[k].similarity < threshold)) { rr[w] = rr[k]; candidates[w++] = candidates[k]; }
    candidates.resize(w);
    rr.resize(w);
  }
  out.insert(out.end(), candidates.begin(), candidates.end());
  if (raws) raws->insert(raws->end(), rr.begin(), rr.end());
}

// Front end of [OCV] Detector::match: quantise every level / modality, spread, response maps, linearize.
static bool build_front(Detector& det, const Source* srcs, int nsrc, const Source* masks, int nmasks) {
  const int L = det.levels(), M = det.M();
  if (nsrc != M) { det.err = "sources.size() != modalities.size()"; return false; }
  if (nmasks != 0 && nmasks != M) { det.err = "masks.size() != modalities.size()"; return false; }
  std::vector<QuantPyr> q(M);
  for (int m = 0; m < M; ++m)
    if (!pyr_process(det, det.mods[m], srcs[m], nmasks ? &masks[m] : nullptr, q[m], det.err)) return false;
  det.front.assign((size_t)L * M, LevelMod());
  for (int l = 0; l < L; ++l) {
    int T = det.T[l];
    for (int m = 0; m < M; ++m) {
      if (l > 0) pyr_down(q[m]);
      LevelMod& lm = det.front[(size_t)l * M + m];
      lm.rows = q[m].rows; lm.cols = q[m].cols; lm.T = T;
      size_t n = (size_t)lm.rows * lm.cols;
      if (n % 16 != 0) { det.err = "(rows*cols) % 16 != 0"; return false; }
      if (lm.rows % T != 0 || lm.cols % T != 0) { det.err = "rows % T != 0 or cols % T != 0"; return false; }
      lm.W = lm.cols / T; lm.H = lm.rows / T;
      lm.quant_raw = q[m].quant;
      lm.magnitude = q[m].magnitude;
      pyr_quantize(q[m], lm.quantized);
      lm.spread.resize(n);
      spread_T(lm.quantized.data(), lm.rows, lm.cols, T, lm.spread.data());
      lm.response.resize(8 * n);
      response_maps(lm.spread.data(), n, det.sim_lut, lm.response.data());
      lm.plane_stride = lm_plane_stride(T, lm.W, lm.H);
      lm.lm.assign(8 * lm.plane_stride, 0);
      for (int o = 0; o < 8; ++o)
        linearize_T(lm.response.data() + o * n, lm.rows, lm.cols, T, lm.lm.data() + o * lm.plane_stride);
    }
  }
  return true;
}

#include "linemod_fast.inc"

static bool build_front_any(Detector& det, const Source* srcs, int nsrc, const Source* masks, int nmasks) {
  if (!det.fast) return build_front(det, srcs, nsrc, masks, nmasks);
  if (!det.pool || det.pool->size() != det.threads) det.pool.reset(new BandPool(det.threads));
  return build_front_fast(det, *det.pool, srcs, nsrc, masks, nmasks);
}

static int class_index_of(const Detector& det, const std::string& id) {
  int i = 0;
  for (auto it = det.classes.begin(); it != det.classes.end(); ++it, ++i)
    if (it->first == id) return i;
  return -1;
}

// Matching half of [OCV] Detector::match, on the front end built by build_front().
static void run_match(Detector& det, float threshold, const char* const* class_ids, int n_ids, bool keep_cands,
                      std::vector<MatchRec>& matches) {
  matches.clear();
  det.last_cands.clear();
  std::vector<std::pair<int, const std::vector<TemplatePyramid>*> > todo;
  if (n_ids == 0) {
    int ci = 0;
    for (auto it = det.classes.begin(); it != det.classes.end(); ++it, ++ci) todo.push_back(std::make_pair(ci, &it->second));
  } else {
    for (int i = 0; i < n_ids; ++i) {
      auto it = det.classes.find(class_ids[i]);
      if (it != det.classes.end()) todo.push_back(std::make_pair(class_index_of(det, class_ids[i]), &it->second));
    }
  }
  det.last_raw.clear();
  uint32_t order_base = 0;  // position in the iteration order (== canonical index when all classes are matched)
  for (size_t k = 0; k < todo.size(); ++k) {
    int ci = todo[k].first;
    const std::vector<TemplatePyramid>& tps = *todo[k].second;
    int n = (int)tps.size();
    std::vector<std::vector<MatchRec> > per((size_t)n);
    std::vector<std::vector<CandRec> > perc((size_t)n);
    std::vector<std::vector<RawRec> > perr((size_t)n);
    // Templates are independent given the frame's linear memories; the reference loop is sequential, the
    // "all host cores" baseline hands out chunks of 8 templates to det.threads workers (order restored below).
    std::atomic<int> next(0);
    auto worker = [&]() {
      for (;;) {
        int t0 = next.fetch_add(8);
        if (t0 >= n) break;
        for (int t = t0; t < std::min(n, t0 + 8); ++t)
          match_template(det, ci, t, tps[t], threshold, per[t], keep_cands ? &perc[t] : nullptr, nullptr,
                         keep_cands ? &perr[t] : nullptr, order_base + (uint32_t)t);
      }
    };
    if (det.threads <= 1) worker();
    else {
      std::vector<std::thread> pool;
      for (int w = 0; w < det.threads; ++w) pool.emplace_back(worker);
      for (auto& th : pool) th.join();
    }
    for (int t = 0; t < n; ++t) {
      matches.insert(matches.end(), per[t].begin(), per[t].end());
      if (keep_cands) det.last_cands.insert(det.last_cands.end(), perc[t].begin(), perc[t].end());
      if (keep_cands) det.last_raw.insert(det.last_raw.end(), perr[t].begin(), perr[t].end());
    }
    order_base += (uint32_t)n;
  }
  det.last_presort = matches;
  std::sort(matches.begin(), matches.end(), match_less);
  matches.erase(std::unique(matches.begin(), matches.end(), match_equal), matches.end());
}

}  // namespace

// ===================================================================================== C interface
extern "C" {

typedef struct {
  const void* data; int32_t rows, cols, type; size_t step;
} orc_image;  // type: 0 = 8UC3 (BGR), 1 = 16UC1 (depth, mm), 2 = 8UC1 (mask)

typedef struct {
  int32_t type;            // 0 ColorGradient, 1 DepthNormal
  float weak_threshold, strong_threshold;
  int32_t distance_threshold, difference_threshold, extract_threshold;
  int32_t num_features;
} orc_modality;

static Source to_source(const orc_image& im) {
  Source s; s.data = (const uint8_t*)im.data; s.rows = im.rows; s.cols = im.cols; s.type = im.type; s.step = im.step;
  return s;
}

void* orc_create(const int32_t* T, int levels, const orc_modality* mods, int M) {
  Detector* d = new Detector();
  d->T.assign(T, T + levels);
  for (int m = 0; m < M; ++m) {
    ModalityDesc md;
    md.type = mods[m].type;
    md.weak_threshold = mods[m].weak_threshold; md.strong_threshold = mods[m].strong_threshold;
    md.distance_threshold = mods[m].distance_threshold; md.difference_threshold = mods[m].difference_threshold;
    md.extract_threshold = mods[m].extract_threshold; md.num_features = mods[m].num_features;
    d->mods.push_back(md);
  }
  default_similarity_lut(d->sim_lut);
  default_normal_lut(d->normal_lut);
  return d;
}
void orc_destroy(void* h) { delete (Detector*)h; }
const char* orc_last_error(void* h) { return ((Detector*)h)->err.c_str(); }
void orc_set_threads(void* h, int n) { ((Detector*)h)->threads = n < 1 ? 1 : n; }
void orc_set_fast(void* h, int on) { ((Detector*)h)->fast = on != 0; }
int orc_max_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}
void orc_set_similarity_lut(void* h, const uint8_t* lut) { std::memcpy(((Detector*)h)->sim_lut, lut, 256); }
void orc_get_similarity_lut(void* h, uint8_t* lut) { std::memcpy(lut, ((Detector*)h)->sim_lut, 256); }
void orc_set_normal_lut(void* h, const uint8_t* lut) { std::memcpy(((Detector*)h)->normal_lut, lut, 8000); }
void orc_get_normal_lut(void* h, uint8_t* lut) { std::memcpy(lut, ((Detector*)h)->normal_lut, 8000); }

// [OCV] Detector::addTemplate up to the point where the pyramid joins the class: quantise, extract per level and
// modality, cropTemplates.  Reads the detector only (thread-safe).  0 = ok, -1 = some level lacks candidates, -2 = error.
static int extract_pyramid(const Detector& det, const orc_image* srcs, const orc_image* mask, TemplatePyramid& tp, int box[4],
                           std::string& err) {
  const int L = det.levels(), M = det.M();
  tp.assign((size_t)L * M, Template());
  Source msk; if (mask && mask->data) msk = to_source(*mask);
  for (int m = 0; m < M; ++m) {
    QuantPyr q;
    if (!pyr_process(det, det.mods[m], to_source(srcs[m]), (mask && mask->data) ? &msk : nullptr, q, err)) return -2;
    for (int l = 0; l < L; ++l) {
      if (l > 0) pyr_down(q);
      bool ok = det.mods[m].type == MOD_COLOR_GRADIENT ? cg_extract(q, tp[(size_t)l * M + m]) : dn_extract(q, tp[(size_t)l * M + m]);
      if (!ok) return -1;
    }
  }
  crop_templates(tp, box);
  return 0;
}

// [OCV] Modality::process(src, mask) of modality m, `level` pyrDown() calls, then QuantizedPyramid::quantize(dst) and
// extractTemplate(templ): quant_out (nullable) rows>>level x cols>>level; hdr = {width, height, level, n}; feats (nullable)
// n (x, y, label) triples.  Returns 1 / 0 = extractTemplate's bool, -2 on error.
int orc_modality_process(void* h, int m, const orc_image* src, const orc_image* mask, int level, uint8_t* quant_out,
                         int32_t* hdr, int32_t* feats) {
  Detector& det = *(Detector*)h;
  if (m < 0 || m >= det.M() || level < 0) { det.err = "modality / level out of range"; return -2; }
  Source msk; if (mask && mask->data) msk = to_source(*mask);
  QuantPyr q;
  if (!pyr_process(det, det.mods[m], to_source(*src), (mask && mask->data) ? &msk : nullptr, q, det.err)) return -2;
  for (int l = 0; l < level; ++l) pyr_down(q);
  if (quant_out) {
    std::vector<uint8_t> dst;
    pyr_quantize(q, dst);
    std::memcpy(quant_out, dst.data(), dst.size());
  }
  Template t;
  const bool ok = det.mods[m].type == MOD_COLOR_GRADIENT ? cg_extract(q, t) : dn_extract(q, t);
  hdr[0] = t.width; hdr[1] = t.height; hdr[2] = t.pyramid_level; hdr[3] = ok ? (int)t.features.size() : 0;
  if (ok && feats)
    for (size_t j = 0; j < t.features.size(); ++j) { feats[3 * j] = t.features[j].x; feats[3 * j + 1] = t.features[j].y; feats[3 * j + 2] = t.features[j].label; }
  return ok ? 1 : 0;
}

// [OCV] Detector::addTemplate.  Returns template_id, -1 if any level lacks candidates, -2 on error.
int orc_add_template(void* h, const orc_image* srcs, int nsrc, const char* class_id, const orc_image* mask,
                     int32_t* bb /*nullable [x,y,w,h]*/) {
  Detector& det = *(Detector*)h;
  if (nsrc != det.M()) { det.err = "sources.size() != modalities.size()"; return -2; }
  std::vector<TemplatePyramid>& tps = det.classes[class_id];
  int template_id = (int)tps.size();
  TemplatePyramid tp;
  int box[4];
  const int rc = extract_pyramid(det, srcs, mask, tp, box, det.err);
  if (rc != 0) return rc;
  if (bb) { bb[0] = box[0]; bb[1] = box[1]; bb[2] = box[2]; bb[3] = box[3]; }
  tps.push_back(tp);
  return template_id;
}

// The trainer's loop (/root/reference/src/renderer.cpp:239-329: render a view, addTemplate) for n views with the scalar
// rasteriser of render_oracle.cpp, det.threads views at a time; templates join the class in view order, so the result
// equals n sequential render + orc_add_template calls.  Modalities must be (ColorGradient, DepthNormal) in that order or
// a prefix / single one of them.  tids[v] = template_id or -1; returns the number of templates added, -2 on error.
struct OrcCameraDesc { int32_t width, height; double fx, fy, near_, far_; };
int orc_render(const float* tris, int n_tri, const OrcCameraDesc* cam, const double T[3], const double up[3], uint8_t* bgr,
               uint16_t* depth, uint8_t* mask, int32_t rect[4]);
int orc_train_views(void* h, const float* tris, int n_tri, const OrcCameraDesc* cam, const double* T, const double* up, int n_views,
                    const char* class_id, int32_t* tids) {
  Detector& det = *(Detector*)h;
  const int M = det.M();
  std::vector<TemplatePyramid> pyr((size_t)n_views);
  std::vector<int> status((size_t)n_views, -2);
  std::atomic<int> next(0);
  const size_t px = (size_t)cam->width * cam->height;
  auto worker = [&]() {
    std::vector<uint8_t> bgr(px * 3), mask(px);
    std::vector<uint16_t> depth(px);
    std::string err;
    for (;;) {
      const int v = next.fetch_add(1);
      if (v >= n_views) break;
      int32_t rect[4];
      if (orc_render(tris, n_tri, cam, T + 3 * v, up + 3 * v, bgr.data(), depth.data(), mask.data(), rect) != 0) continue;
      orc_image srcs[4], mk;
      mk.data = mask.data(); mk.rows = cam->height; mk.cols = cam->width; mk.type = 2; mk.step = (size_t)cam->width;
      for (int m = 0; m < M; ++m) {
        const bool cg = det.mods[m].type == MOD_COLOR_GRADIENT;
        srcs[m].data = cg ? (const void*)bgr.data() : (const void*)depth.data();
        srcs[m].rows = cam->height; srcs[m].cols = cam->width; srcs[m].type = cg ? 0 : 1;
        srcs[m].step = (size_t)cam->width * (cg ? 3 : 2);
      }
      int box[4];
      status[v] = extract_pyramid(det, srcs, &mk, pyr[v], box, err);
    }
  };
  if (det.threads <= 1) worker();
  else {
    std::vector<std::thread> pool;
    for (int w = 0; w < det.threads; ++w) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  std::vector<TemplatePyramid>& tps = det.classes[class_id];
  int added = 0;
  for (int v = 0; v < n_views; ++v) {
    if (tids) tids[v] = -1;
    if (status[v] != 0) continue;
    if (tids) tids[v] = (int)tps.size();
    tps.push_back(pyr[v]);
    ++added;
  }
  return added;
}

// [OCV] Detector::addSyntheticTemplate, flat encoding: per template (L*M of them) {width,height,level,nfeat},
// then all features as {x,y,label} triples in the same order.
int orc_add_synthetic_template(void* h, const char* class_id, int n_templ, const int32_t* hdr /*4*n_templ*/,
                               const int32_t* feats) {
  Detector& det = *(Detector*)h;
  if (n_templ != det.levels() * det.M()) { det.err = "template pyramid size mismatch"; return -2; }
  TemplatePyramid tp((size_t)n_templ);
  size_t k = 0;
  for (int i = 0; i < n_templ; ++i) {
    tp[i].width = hdr[4 * i]; tp[i].height = hdr[4 * i + 1]; tp[i].pyramid_level = hdr[4 * i + 2];
    int nf = hdr[4 * i + 3];
    if (nf > 63) { det.err = "features.size() > 63"; return -2; }
    tp[i].features.resize(nf);
    for (int j = 0; j < nf; ++j, ++k) {
      tp[i].features[j].x = feats[3 * k]; tp[i].features[j].y = feats[3 * k + 1]; tp[i].features[j].label = feats[3 * k + 2];
    }
  }
  std::vector<TemplatePyramid>& tps = det.classes[class_id];
  tps.push_back(tp);
  return (int)tps.size() - 1;
}

int orc_num_classes(void* h) { return (int)((Detector*)h)->classes.size(); }
int orc_num_templates(void* h, const char* class_id /*nullable = all*/) {
  Detector& det = *(Detector*)h;
  int n = 0;
  for (auto it = det.classes.begin(); it != det.classes.end(); ++it)
    if (!class_id || it->first == class_id) n += (int)it->second.size();
  return n;
}
const char* orc_class_id(void* h, int index) {
  Detector& det = *(Detector*)h;
  int i = 0;
  for (auto it = det.classes.begin(); it != det.classes.end(); ++it, ++i)
    if (i == index) return it->first.c_str();
  return nullptr;
}
// Template export: hdr gets 4 ints per template of the pyramid; returns total feature count. feats may be null.
int orc_get_template(void* h, const char* class_id, int template_id, int32_t* hdr, int32_t* feats) {
  Detector& det = *(Detector*)h;
  auto it = det.classes.find(class_id);
  if (it == det.classes.end() || template_id < 0 || template_id >= (int)it->second.size()) return -1;
  const TemplatePyramid& tp = it->second[template_id];
  int total = 0;
  for (size_t i = 0; i < tp.size(); ++i) {
    if (hdr) { hdr[4 * i] = tp[i].width; hdr[4 * i + 1] = tp[i].height; hdr[4 * i + 2] = tp[i].pyramid_level; hdr[4 * i + 3] = (int)tp[i].features.size(); }
    for (size_t j = 0; j < tp[i].features.size(); ++j, ++total)
      if (feats) { feats[3 * total] = tp[i].features[j].x; feats[3 * total + 1] = tp[i].features[j].y; feats[3 * total + 2] = tp[i].features[j].label; }
  }
  return total;
}

// Front end only (quantise -> spread -> response -> linearize); results via orc_debug_fetch.
int orc_build_front(void* h, const orc_image* srcs, int nsrc, const orc_image* masks, int nmasks) {
  Detector& det = *(Detector*)h;
  std::vector<Source> s, mk;
  for (int i = 0; i < nsrc; ++i) s.push_back(to_source(srcs[i]));
  for (int i = 0; i < nmasks; ++i) mk.push_back(to_source(masks[i]));
  return build_front_any(det, s.data(), nsrc, mk.data(), nmasks) ? 0 : -2;
}

// Matching only, on the front end of the last orc_build_front / orc_match call.  Returns #matches (or -2);
// *out is malloc'ed, free with orc_free.
long orc_match_only(void* h, float threshold, const char* const* class_ids, int n_ids, int keep_cands, MatchRec** out) {
  Detector& det = *(Detector*)h;
  if (det.front.empty()) { det.err = "no front end built"; return -2; }
  std::vector<MatchRec> matches;
  run_match(det, threshold, class_ids, n_ids, keep_cands != 0, matches);
  *out = (MatchRec*)std::malloc(std::max<size_t>(1, matches.size()) * sizeof(MatchRec));
  std::memcpy(*out, matches.data(), matches.size() * sizeof(MatchRec));
  return (long)matches.size();
}

// [OCV] Detector::match
long orc_match(void* h, const orc_image* srcs, int nsrc, float threshold, const char* const* class_ids, int n_ids,
               const orc_image* masks, int nmasks, int keep_cands, MatchRec** out) {
  if (orc_build_front(h, srcs, nsrc, masks, nmasks) != 0) return -2;
  return orc_match_only(h, threshold, class_ids, n_ids, keep_cands, out);
}
void orc_free(void* p) { std::free(p); }

long orc_last_presort(void* h, MatchRec* dst /*nullable*/) {
  Detector& det = *(Detector*)h;
  if (dst) std::memcpy(dst, det.last_presort.data(), det.last_presort.size() * sizeof(MatchRec));
  return (long)det.last_presort.size();
}
// Survivors of the last match (keep_cands != 0) in the product's exchange format, emission order.
long orc_last_raw(void* h, RawRec* dst /*nullable*/) {
  Detector& det = *(Detector*)h;
  if (dst) std::memcpy(dst, det.last_raw.data(), det.last_raw.size() * sizeof(RawRec));
  return (long)det.last_raw.size();
}
long orc_last_candidates(void* h, CandRec* dst /*nullable*/) {
  Detector& det = *(Detector*)h;
  if (dst) std::memcpy(dst, det.last_cands.data(), det.last_cands.size() * sizeof(CandRec));
  return (long)det.last_cands.size();
}

// Coarse u16 similarity map of one template on the last front end (H*W of the lowest level).
int orc_coarse_map(void* h, const char* class_id, int template_id, uint16_t* dst) {
  Detector& det = *(Detector*)h;
  auto it = det.classes.find(class_id);
  if (it == det.classes.end() || template_id < 0 || template_id >= (int)it->second.size() || det.front.empty()) return -1;
  std::vector<MatchRec> tmp;
  match_template(det, 0, template_id, it->second[template_id], 1e9f, tmp, nullptr, dst);
  return 0;
}

// Parity taps.  stage: 0 quantized(u8) 1 spread(u8) 2 response(8 x u8) 3 linear memories(8 x plane_stride u8)
//               4 CG magnitude(f32) 5 unmasked quantisation(u8).  Returns byte count (dst may be null).
long orc_debug_fetch(void* h, int stage, int level, int modality, void* dst) {
  Detector& det = *(Detector*)h;
  if (det.front.empty() || level < 0 || level >= det.levels() || modality < 0 || modality >= det.M()) return -1;
  const LevelMod& lm = det.front[(size_t)level * det.M() + modality];
  const void* p = nullptr; size_t n = 0;
  switch (stage) {
    case 0: p = lm.quantized.data(); n = lm.quantized.size(); break;
    case 1: p = lm.spread.data(); n = lm.spread.size(); break;
    case 2: p = lm.response.data(); n = lm.response.size(); break;
    case 3: p = lm.lm.data(); n = lm.lm.size(); break;
    case 4: p = lm.magnitude.data(); n = lm.magnitude.size() * sizeof(float); break;
    case 5: p = lm.quant_raw.data(); n = lm.quant_raw.size(); break;
    default: return -1;
  }
  if (dst && n) std::memcpy(dst, p, n);
  return (long)n;
}
int orc_level_geometry(void* h, int level, int modality, int32_t* out /*rows,cols,T,W,H*/, size_t* plane_stride) {
  Detector& det = *(Detector*)h;
  if (det.front.empty()) return -1;
  const LevelMod& lm = det.front[(size_t)level * det.M() + modality];
  out[0] = lm.rows; out[1] = lm.cols; out[2] = lm.T; out[3] = lm.W; out[4] = lm.H;
  *plane_stride = lm.plane_stride;
  return 0;
}

// --------------------------------------------------------------------- primitive entry points (pinned vs cv2)
void orc_prim_gaussian7(const uint8_t* src, int rows, int cols, int ch, uint8_t* dst) { gaussian7_u8(src, rows, cols, ch, dst); }
void orc_prim_sobel3(const uint8_t* src, int rows, int cols, int ch, int16_t* dx, int16_t* dy) { sobel3_u8(src, rows, cols, ch, dx, dy); }
void orc_prim_phase_deg(const float* x, const float* y, size_t n, float* out) { for (size_t i = 0; i < n; ++i) out[i] = fast_atan2_deg(y[i], x[i]); }
void orc_prim_pyrdown(const uint8_t* src, int rows, int cols, int ch, uint8_t* dst) { pyrdown_u8(src, rows, cols, ch, dst); }
void orc_prim_nn_half(const uint8_t* src, int rows, int cols, uint8_t* dst) { nn_half_u8(src, rows, cols, dst); }
void orc_prim_median5(const uint8_t* src, int rows, int cols, uint8_t* dst) { median5_u8(src, rows, cols, dst); }
void orc_prim_median5_fast(const uint8_t* src, int rows, int cols, uint8_t* dst) { fast_median5_rows(src, rows, cols, 0, rows, dst); }
void orc_prim_erode3(const uint8_t* src, int rows, int cols, int iterations, uint8_t* dst) { erode3_u8(src, rows, cols, iterations, dst); }
void orc_prim_distance_c3(const uint8_t* src, int rows, int cols, float* dst) { distance_transform_c3(src, rows, cols, dst); }
void orc_prim_cg_quantize(const uint8_t* bgr, int rows, int cols, float weak, float* magnitude, uint8_t* quantized, float* angle) {
  quantized_orientations(bgr, rows, cols, weak, magnitude, quantized, angle);
}
void orc_prim_dn_quantize(void* h, const uint16_t* depth, int rows, int cols, int distance_threshold, int difference_threshold,
                          uint8_t* raw, uint8_t* out) {
  quantized_normals_raw(depth, rows, cols, distance_threshold, difference_threshold, ((Detector*)h)->normal_lut, raw);
  median5_u8(raw, rows, cols, out);
}
void orc_prim_spread(const uint8_t* src, int rows, int cols, int T, uint8_t* dst) { spread_T(src, rows, cols, T, dst); }

// Final ordering stage alone ([OCV] Detector::match tail: std::sort + std::unique), in place; returns new length.
long orc_sort_unique(MatchRec* recs, long n) {
  std::sort(recs, recs + n, match_less);
  return (long)(std::unique(recs, recs + n, match_equal) - recs);
}

}  // extern "C"
