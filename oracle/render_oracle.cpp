// render_oracle.cpp -- CPU restatement of the template-generation front half: the view-sphere
// iterator and the mesh renderer that feed Detector::addTemplate in /root/reference/src/renderer.cpp:239-329.
//
// TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
//
// PINNING: the reference renders with `object_recognition_renderer` (ORK Renderer3d / RendererIterator over OpenGL +
// assimp), an external dependency that is not vendored under /root/reference (CMakeLists.txt
// find_package(object_recognition_renderer)), so it cannot be run here.  What the reference does ship is the output of
// one training run: config/data/boxNew_longDistance_linemod_xtion_renderer_params.yml -- per template the pose (R, T),
// the centre-depth offset D and the render rectangle -- for config/stl/boxNew.stl.  tests/golden/make_renderer_golden.py
// turns it into tests/golden/renderer_params_boxnew.npz and tests/test_oracle_render.py checks against it:
//   * view_params / the iteration order (ORK RendererIterator, renderer/src/utils.cpp, restated from memory): all 2 652
//     recorded poses are reproduced to 1e-15 (T = -camera position; R rows = left, -up, -view direction) in iteration order;
//   * the rasteriser (this project's own specification of "pinhole z-buffer of a triangle mesh", below; OpenGL's
//     fixed-function output is not reproducible bit for bit): silhouette bounding boxes equal the 2 652 recorded
//     rectangles (ORK grows its box by one pixel per side and keeps GL's bottom-left origin) and the centre depth
//     agrees with D within 1 mm.
//   Shading (BGR) is NOT pinned: STL carries no material and the GL lighting set-up is unknown; ColorGradient templates
//   only use the silhouette ring, where the contrast to the black background dominates.
// The CUDA rasteriser must agree with this scalar one bit-exactly on every pixel (tests/test_gpu_train.py).
//
// Rasteriser specification (all arithmetic IEEE f32, no contraction, evaluation order as written):
//   camera      eye = T, looks at the object origin, `up` as given (gluLookAt): f = normalize(-T), s = normalize(f x up),
//               u = s x f.  Camera frame (OpenCV convention): Xc = s.(P-eye), Yc = -u.(P-eye), Zc = f.(P-eye); the
//               3x4 matrix [R|t] is built in double and rounded to f32 once.
//   vertex      Pc_i = ((R_i0*x + R_i1*y) + R_i2*z) + t_i;  px = (fx*Xc)/Zc + W/2;  py = (fy*Yc)/Zc + H/2
//               (principal point at the image centre, as K at renderer.cpp:273).
//   triangle    skipped when any Zc <= near or its screen area is 0.  Pixel (ix, iy) is sampled at its centre
//               (ix+0.5, iy+0.5); covered when the three edge functions have the sign of the area (zero counts as inside).
//   depth       perspective-correct: b_i = w_i/area, iz = (b0*(1/z0) + b1*(1/z1)) + b2*(1/z2), z = 1/iz; fragments with
//               z > 0.99*far are dropped (ORK's max_allowed_z).  The nearest fragment wins; ties go to the lower
//               triangle index.
//   outputs     depth u16 = round-half-even(z * 1000) millimetres (cv::Mat::convertTo(CV_16UC1, 1e3)), 0 = background;
//               mask u8 = 255 where covered; BGR = (v, v, v) with v = (int)(40 + 200*|nz|), nz the z component of the
//               winning triangle's unit normal in the camera frame (a head-light Lambert term; STL carries no material);
//               rect = tight bounding box of the mask (x, y, w, h), all zero when nothing was drawn.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct ViewSphere {
  int32_t n_points, angle_min, angle_max, angle_step;
  float radius_min, radius_max, radius_step;
};
struct Camera {
  int32_t width, height;
  double fx, fy, near_, far_;
};

void normalize3f(float& x, float& y, float& z) {
  float n = std::sqrt(x * x + y * y + z * z);
  x /= n; y /= n; z /= n;
}
void normalize3d(double* v) {
  double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] /= n; v[1] /= n; v[2] /= n;
}
void cross3d(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// ORK RendererIterator::operator++ : angle innermost, then radius, then the sphere point.
struct ViewState { int index; float radius; int angle; };
std::vector<ViewState> enumerate_views(const ViewSphere& vs) {
  std::vector<ViewState> out;
  if (vs.n_points <= 0 || vs.angle_step <= 0 || !(vs.radius_step > 0.f)) return out;
  ViewState s = {0, vs.radius_min, vs.angle_min};
  while (s.index < vs.n_points) {
    out.push_back(s);
    s.angle += vs.angle_step;
    if (s.angle > vs.angle_max) {
      s.angle = vs.angle_min;
      s.radius += vs.radius_step;
      // The sweep tolerates the f32 accumulation error: the reference's shipped renderer_params.yml (radius 0.5 .. 1.0
      // step 0.1) holds a sixth radius 1.0000001192092896 (tests/golden/renderer_params_boxnew.npz).
      if (s.radius > vs.radius_max + 1e-6f) {
        s.radius = vs.radius_min;
        ++s.index;
      }
    }
  }
  return out;
}

// ORK RendererIterator::view_params: camera position T on the sphere and the up vector rotated in plane.
void view_params(const ViewSphere& vs, const ViewState& st, double T[3], double up[3]) {
  float angle_rad = (float)(st.angle * 3.14159265358979323846 / 180.);
  float inc = (float)(3.14159265358979323846 * (3 - std::sqrt(5.0)));
  float off = 2.0f / (float)vs.n_points;
  float y = st.index * off - 1.0f + (off / 2.0f);
  float r = std::sqrt(1.0f - y * y);
  float phi = st.index * inc;
  float x = std::cos(phi) * r;
  float z = std::sin(phi) * r;
  float lat = std::acos(z), lon;
  if ((std::fabs(std::sin(lat)) < 1e-5) || (std::fabs(y / std::sin(lat)) > 1))
    lon = 0;
  else
    lon = std::asin(y / std::sin(lat));
  x *= st.radius; y *= st.radius; z *= st.radius;
  float x_up = st.radius * std::cos(lon) * std::sin(lat - 1e-5) - x;
  float y_up = st.radius * std::sin(lon) * std::sin(lat - 1e-5) - y;
  float z_up = st.radius * std::cos(lat - 1e-5) - z;
  normalize3f(x_up, y_up, z_up);
  float x_right = -y_up * z + z_up * y;
  float y_right = x_up * z - z_up * x;
  float z_right = -x_up * y + y_up * x;
  normalize3f(x_right, y_right, z_right);
  float x_new_up = x_up * std::cos(angle_rad) + x_right * std::sin(angle_rad);
  float y_new_up = y_up * std::cos(angle_rad) + y_right * std::sin(angle_rad);
  float z_new_up = z_up * std::cos(angle_rad) + z_right * std::sin(angle_rad);
  T[0] = x; T[1] = y; T[2] = z;
  double u0[3] = {x_new_up, y_new_up, z_new_up};
  double l[3];
  cross3d(u0, T, l);
  normalize3d(l);
  cross3d(T, l, up);
  normalize3d(up);
}

// gluLookAt(eye = T, centre = 0, up) in the OpenCV camera convention, rounded to f32: Pc = R * Po + t.
bool look_at(const double T[3], const double up[3], float R[9], float t[3]) {
  double f[3] = {-T[0], -T[1], -T[2]};
  double nf = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  if (!(nf > 0)) return false;
  f[0] /= nf; f[1] /= nf; f[2] /= nf;
  double s[3], u[3];
  cross3d(f, up, s);
  double ns = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  if (!(ns > 0)) return false;
  s[0] /= ns; s[1] /= ns; s[2] /= ns;
  cross3d(s, f, u);
  double Rd[9] = {s[0], s[1], s[2], -u[0], -u[1], -u[2], f[0], f[1], f[2]};
  for (int i = 0; i < 3; ++i) {
    double ti = -(Rd[3 * i] * T[0] + Rd[3 * i + 1] * T[1] + Rd[3 * i + 2] * T[2]);
    t[i] = (float)ti;
    for (int j = 0; j < 3; ++j) R[3 * i + j] = (float)Rd[3 * i + j];
  }
  return true;
}

inline float edge(float ax, float ay, float bx, float by, float px, float py) {
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}

void render(const float* tris, int n_tri, const Camera& cam, const float R[9], const float t[3], uint8_t* bgr,
            uint16_t* depth, uint8_t* mask, int32_t rect[4]) {
  const int W = cam.width, H = cam.height;
  const float fx = (float)cam.fx, fy = (float)cam.fy, cx = (float)W / 2.0f, cy = (float)H / 2.0f;
  const float z_near = (float)cam.near_, z_max = (float)cam.far_ * 0.99f;
  std::vector<uint64_t> zb((size_t)W * H, ~0ull);
  std::vector<float> nz_abs((size_t)n_tri, 0.f);
  for (int k = 0; k < n_tri; ++k) {
    float X[3], Y[3], Z[3], px[3], py[3];
    bool ok = true;
    for (int v = 0; v < 3; ++v) {
      const float* p = tris + 9 * k + 3 * v;
      X[v] = ((R[0] * p[0] + R[1] * p[1]) + R[2] * p[2]) + t[0];
      Y[v] = ((R[3] * p[0] + R[4] * p[1]) + R[5] * p[2]) + t[1];
      Z[v] = ((R[6] * p[0] + R[7] * p[1]) + R[8] * p[2]) + t[2];
      if (!(Z[v] > z_near)) ok = false;
    }
    {  // unit normal in the camera frame (shading)
      float ax = X[1] - X[0], ay = Y[1] - Y[0], az = Z[1] - Z[0];
      float bx = X[2] - X[0], by = Y[2] - Y[0], bz = Z[2] - Z[0];
      float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
      float nn = std::sqrt((nx * nx + ny * ny) + nz * nz);
      nz_abs[k] = nn > 0.f ? std::fabs(nz / nn) : 0.f;
    }
    if (!ok) continue;
    for (int v = 0; v < 3; ++v) {
      px[v] = (fx * X[v]) / Z[v] + cx;
      py[v] = (fy * Y[v]) / Z[v] + cy;
    }
    const float area = edge(px[0], py[0], px[1], py[1], px[2], py[2]);
    if (area == 0.f || area != area) continue;
    const float minx = std::fmin(px[0], std::fmin(px[1], px[2])), maxx = std::fmax(px[0], std::fmax(px[1], px[2]));
    const float miny = std::fmin(py[0], std::fmin(py[1], py[2])), maxy = std::fmax(py[0], std::fmax(py[1], py[2]));
    if (!(maxx >= 0.f) || !(maxy >= 0.f) || !(minx <= (float)W) || !(miny <= (float)H)) continue;
    const int x_lo = (int)std::fmax(0.f, std::floor(minx)), x_hi = (int)std::fmin((float)(W - 1), std::floor(maxx));
    const int y_lo = (int)std::fmax(0.f, std::floor(miny)), y_hi = (int)std::fmin((float)(H - 1), std::floor(maxy));
    const float iz0 = 1.0f / Z[0], iz1 = 1.0f / Z[1], iz2 = 1.0f / Z[2];
    for (int iy = y_lo; iy <= y_hi; ++iy)
      for (int ix = x_lo; ix <= x_hi; ++ix) {
        const float sx = (float)ix + 0.5f, sy = (float)iy + 0.5f;
        const float w0 = edge(px[1], py[1], px[2], py[2], sx, sy);
        const float w1 = edge(px[2], py[2], px[0], py[0], sx, sy);
        const float w2 = edge(px[0], py[0], px[1], py[1], sx, sy);
        const bool inside = area > 0.f ? (w0 >= 0.f && w1 >= 0.f && w2 >= 0.f) : (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
        if (!inside) continue;
        const float b0 = w0 / area, b1 = w1 / area, b2 = w2 / area;
        const float iz = (b0 * iz0 + b1 * iz1) + b2 * iz2;
        const float z = 1.0f / iz;
        if (!(z > 0.f) || z > z_max) continue;
        uint32_t zbits;
        std::memcpy(&zbits, &z, 4);
        const uint64_t key = ((uint64_t)zbits << 32) | (uint32_t)k;
        uint64_t& cell = zb[(size_t)iy * W + ix];
        if (key < cell) cell = key;
      }
  }
  int x0 = W, y0 = H, x1 = -1, y1 = -1;
  for (int iy = 0; iy < H; ++iy)
    for (int ix = 0; ix < W; ++ix) {
      const size_t i = (size_t)iy * W + ix;
      const uint64_t key = zb[i];
      if (key == ~0ull) {
        depth[i] = 0; mask[i] = 0; bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = 0;
        continue;
      }
      const uint32_t zbits = (uint32_t)(key >> 32);
      float z;
      std::memcpy(&z, &zbits, 4);
      float mm = std::nearbyint(z * 1000.0f);  // round-half-even, like cvRound
      depth[i] = mm > 65535.f ? 65535 : (uint16_t)mm;
      mask[i] = 255;
      const uint8_t v = (uint8_t)(int)(40.0f + 200.0f * nz_abs[(uint32_t)key]);
      bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = v;
      if (ix < x0) x0 = ix;
      if (ix > x1) x1 = ix;
      if (iy < y0) y0 = iy;
      if (iy > y1) y1 = iy;
    }
  if (x1 < 0) { rect[0] = rect[1] = rect[2] = rect[3] = 0; }
  else { rect[0] = x0; rect[1] = y0; rect[2] = x1 - x0 + 1; rect[3] = y1 - y0 + 1; }
}

// depth_diff of /root/reference/src/rgbdDetector.cpp:236-283 (method A): mean |template - scene| over pixels where the
// template mask and the low byte-saturated scene depth are both non-zero, in metres.
double depth_diff(const uint16_t* scene, int scene_cols, const uint16_t* templ, const uint8_t* templ_mask, int templ_cols,
                  int x, int y, int tx, int ty, int w, int h) {
  double sum = 0.0;
  int num = 0;
  for (int r = 0; r < h; ++r)
    for (int c = 0; c < w; ++c) {
      const uint16_t s = scene[(size_t)(y + r) * scene_cols + x + c];
      const uint16_t tv = templ[(size_t)(ty + r) * templ_cols + tx + c];
      const uint8_t smask = s > 255 ? 255 : (uint8_t)s;  // convertTo(CV_8UC1) saturates
      if ((templ_mask[(size_t)(ty + r) * templ_cols + tx + c] & smask) == 0) continue;
      // depth_template - depth_roi on CV_16U saturates at 0, stored into the CV_16S destination header that the
      // assignment replaces: the values read back through short* are the u16 bit patterns reinterpreted.
      const uint16_t d = tv > s ? (uint16_t)(tv - s) : 0;
      const int16_t sd = (int16_t)d;
      sum += (double)std::abs((int)sd);
      ++num;
    }
  return sum / (num * 1000.0);
}

}  // namespace

extern "C" {

int orc_view_count(const ViewSphere* vs) { return (int)enumerate_views(*vs).size(); }

// state out: [point index, angle]; radius out.
int orc_view_params(const ViewSphere* vs, int index, double T[3], double up[3], int32_t state[2], float* radius) {
  std::vector<ViewState> v = enumerate_views(*vs);
  if (index < 0 || index >= (int)v.size()) return -1;
  view_params(*vs, v[index], T, up);
  if (state) { state[0] = v[index].index; state[1] = v[index].angle; }
  if (radius) *radius = v[index].radius;
  return 0;
}

// All views at once: T / up [n][3], state [n][2] = (point index, angle), radius [n]; returns n.
int orc_view_list(const ViewSphere* vs, double* T, double* up, int32_t* state, float* radius) {
  std::vector<ViewState> v = enumerate_views(*vs);
  for (size_t i = 0; i < v.size(); ++i) {
    view_params(*vs, v[i], T + 3 * i, up + 3 * i);
    state[2 * i] = v[i].index; state[2 * i + 1] = v[i].angle;
    radius[i] = v[i].radius;
  }
  return (int)v.size();
}

int orc_look_at(const double T[3], const double up[3], float R[9], float t[3]) { return look_at(T, up, R, t) ? 0 : -1; }

int orc_render(const float* tris, int n_tri, const Camera* cam, const double T[3], const double up[3], uint8_t* bgr,
               uint16_t* depth, uint8_t* mask, int32_t rect[4]) {
  float R[9], t[3];
  if (!look_at(T, up, R, t)) return -1;
  render(tris, n_tri, *cam, R, t, bgr, depth, mask, rect);
  return 0;
}

double orc_depth_diff(const uint16_t* scene, int scene_cols, const uint16_t* templ, const uint8_t* templ_mask, int templ_cols,
                      int x, int y, int tx, int ty, int w, int h) {
  return depth_diff(scene, scene_cols, templ, templ_mask, templ_cols, x, y, tx, ty, w, h);
}

}  // extern "C"
