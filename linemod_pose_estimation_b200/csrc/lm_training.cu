// lm_training.cu -- host side of template generation behind the C ABI: meshes and the view sphere, the
// rasteriser driver, batched addTemplate (lm_train_views / lm_add_templates_batch), the depth hypothesis check and the
// pose table (renderer_params.yml).  Kernels: lm_render.cu, lm_train.cu; shared workspace types: lm_detector_internal.hpp.
//
// Reference loop: /root/reference/src/renderer.cpp:239-329 (render + addTemplate per view), :72-123 (pose table writer);
// src/rgbdDetector.cpp:1681-1749 (pose table reader), :147-283 (depth_diff over rendered hypotheses).
#include <fstream>
#include <sstream>

#include "lm_detector_internal.hpp"
#include "lm_yaml.hpp"

// ================================================================================================ template generation
// meshes, the view sphere, rendering, batched addTemplate and the depth hypothesis check.
// Reference loop: /root/reference/src/renderer.cpp:239-329; depth check: src/rgbdDetector.cpp:147-283.
struct lm_mesh {
  std::vector<float> tris;  // n x 3 vertices x (x, y, z)
  int n = 0;
  int device = -1;          // where d_tris lives (uploaded on first use)
  void* d_tris = nullptr;
};

namespace {

const int kTrainBatch = 32;  // views per batch: bounds the image pools (1.8 MB per 640x480 view) and the key pool

bool parse_stl(const std::string& buf, std::vector<float>& tris, std::string& err) {
  tris.clear();
  if (buf.size() >= 84) {  // binary: 80-byte header, u32 count, 50 bytes per facet (normal, 3 vertices, attribute)
    uint32_t n;
    std::memcpy(&n, buf.data() + 80, 4);
    if ((uint64_t)84 + (uint64_t)50 * n == buf.size()) {
      tris.resize((size_t)n * 9);
      for (uint32_t i = 0; i < n; ++i) std::memcpy(&tris[(size_t)i * 9], buf.data() + 84 + (size_t)50 * i + 12, 36);
      return true;
    }
  }
  size_t pos = 0;  // ASCII: every "vertex x y z"
  while ((pos = buf.find("vertex", pos)) != std::string::npos) {
    pos += 6;
    const char* p = buf.c_str() + pos;
    for (int k = 0; k < 3; ++k) {
      char* end = nullptr;
      double v = std::strtod(p, &end);
      if (end == p) { err = "malformed vertex in ASCII STL"; return false; }
      tris.push_back((float)v);
      p = end;
    }
    pos = (size_t)(p - buf.c_str());
  }
  if (tris.empty() || tris.size() % 9 != 0) { err = "not an STL file (no complete facets found)"; return false; }
  return true;
}

struct SphereShape { int n_angles, n_radii; std::vector<float> radii; };
SphereShape sphere_shape(const lm_view_sphere& vs) {
  SphereShape sh;
  sh.n_angles = vs.angle_max >= vs.angle_min ? (vs.angle_max - vs.angle_min) / vs.angle_step + 1 : 1;
  float r = vs.radius_min;  // the iterator accumulates the radius in f32
  // ... and tolerates the accumulation error: the reference's shipped renderer_params.yml (0.5 .. 1.0 step 0.1) holds a
  // sixth radius 1.0000001192092896
  do { sh.radii.push_back(r); r += vs.radius_step; } while (!(r > vs.radius_max + 1e-6f) && sh.radii.size() < (1u << 20));
  sh.n_radii = (int)sh.radii.size();
  return sh;
}
bool sphere_valid(const lm_view_sphere* vs) {
  return vs && vs->n_points > 0 && vs->angle_step > 0 && vs->radius_step > 0.f;
}

void unit3f(float& x, float& y, float& z) {
  const float n = std::sqrt(x * x + y * y + z * z);
  x /= n; y /= n; z /= n;
}
void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
bool unit3(double v[3]) {
  const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (!(n > 0)) return false;
  v[0] /= n; v[1] /= n; v[2] /= n;
  return true;
}

// ORK RendererIterator::view_params (golden-spiral point `point` of n_points, in-plane rotation `angle_deg`, radius).
void sphere_view(int n_points, int point, int angle_deg, float radius, double T[3], double up[3]) {
  const double kPi = 3.14159265358979323846;
  const float angle_rad = (float)(angle_deg * kPi / 180.);
  const float inc = (float)(kPi * (3 - std::sqrt(5.0)));
  const float off = 2.0f / (float)n_points;
  float y = point * off - 1.0f + (off / 2.0f);
  const float r = std::sqrt(1.0f - y * y);
  const float phi = point * inc;
  float x = std::cos(phi) * r, z = std::sin(phi) * r;
  const float lat = std::acos(z);
  float lon = 0;
  if (!((std::fabs(std::sin(lat)) < 1e-5) || (std::fabs(y / std::sin(lat)) > 1))) lon = std::asin(y / std::sin(lat));
  x *= radius; y *= radius; z *= radius;
  float ux = radius * std::cos(lon) * std::sin(lat - 1e-5) - x;
  float uy = radius * std::sin(lon) * std::sin(lat - 1e-5) - y;
  float uz = radius * std::cos(lat - 1e-5) - z;
  unit3f(ux, uy, uz);
  float rx = -uy * z + uz * y, ry = ux * z - uz * x, rz = -ux * y + uy * x;
  unit3f(rx, ry, rz);
  const float ca = std::cos(angle_rad), sa = std::sin(angle_rad);
  const double u0[3] = {ux * ca + rx * sa, uy * ca + ry * sa, uz * ca + rz * sa};
  T[0] = x; T[1] = y; T[2] = z;
  double left[3];
  cross3(u0, T, left);
  unit3(left);
  cross3(T, left, up);
  unit3(up);
}

// gluLookAt(eye = T, centre = origin, up) expressed in the OpenCV camera convention: Pc = R * Po + t.
bool look_at(const double T[3], const double up[3], double R[9], double t[3]) {
  double f[3] = {-T[0], -T[1], -T[2]};
  if (!unit3(f)) return false;
  double s[3], u[3];
  cross3(f, up, s);
  if (!unit3(s)) return false;
  cross3(s, f, u);
  const double Rd[9] = {s[0], s[1], s[2], -u[0], -u[1], -u[2], f[0], f[1], f[2]};
  for (int i = 0; i < 3; ++i) {
    t[i] = -(Rd[3 * i] * T[0] + Rd[3 * i + 1] * T[1] + Rd[3 * i + 2] * T[2]);
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Rd[3 * i + j];
  }
  return true;
}

int mesh_on_device(lm_detector* d, const lm_mesh* mesh_c, const float** out) {
  lm_mesh* mesh = const_cast<lm_mesh*>(mesh_c);
  if (mesh->d_tris && mesh->device != d->device) {
    cudaSetDevice(mesh->device); cudaFree(mesh->d_tris); mesh->d_tris = nullptr; cudaSetDevice(d->device);
  }
  if (!mesh->d_tris) {
    CU(cudaMalloc(&mesh->d_tris, std::max<size_t>(36, mesh->tris.size() * sizeof(float))));
    CU(cudaMemcpy(mesh->d_tris, mesh->tris.data(), mesh->tris.size() * sizeof(float), cudaMemcpyHostToDevice));
    mesh->device = d->device;
  }
  *out = (const float*)mesh->d_tris;
  return LM_OK;
}

int check_camera(const lm_camera* cam) {
  if (!cam || cam->width <= 0 || cam->height <= 0 || cam->width > 8191 || cam->height > 8191 || !(cam->fx > 0) || !(cam->fy > 0) ||
      !(cam->near_ > 0) || !(cam->far_ > cam->near_))
    return lm_fail(LM_E_INVALID, "bad camera (width/height 1..8191, fx, fy > 0, 0 < near < far)");
  return LM_OK;
}

// Enqueues the rasteriser for n views (T, up: 3 doubles each) on stream s.  The images land in the device buffers given
// (one tightly packed image per view; a null buffer skips that output), the mask rectangles in d->train.rects as
// (x_min, y_min, x_max, y_max) per view.
int render_batch(lm_detector* d, const float* d_tris, int n_tri, const lm_camera& cam, const double* T, const double* up, int n,
                 uint8_t* d_bgr, uint16_t* d_depth, uint8_t* d_mask, cudaStream_t s) {
  TrainWs& ws = d->train;
  const size_t px = (size_t)cam.width * cam.height;
  if (ws.zbuf.ensure(px * 8 * n) != LM_OK || ws.nz_abs.ensure(std::max<size_t>(4, (size_t)n * n_tri * 4)) != LM_OK ||
      ws.views.ensure(sizeof(RenderView) * n) != LM_OK || ws.rects.ensure(16 * (size_t)n) != LM_OK ||
      ws.h_stage.ensure(sizeof(RenderView) * n) != LM_OK)
    return LM_E_CUDA;
  RenderView* hv = ws.h_stage.as<RenderView>();
  for (int v = 0; v < n; ++v) {
    double R[9], t[3];
    if (!look_at(T + 3 * v, up + 3 * v, R, t)) return lm_fail(LM_E_INVALID, "view %d: degenerate camera position / up vector", v);
    for (int i = 0; i < 9; ++i) hv[v].R[i] = (float)R[i];
    for (int i = 0; i < 3; ++i) hv[v].t[i] = (float)t[i];
  }
  CU(cudaMemcpyAsync(ws.views.p, hv, sizeof(RenderView) * n, cudaMemcpyHostToDevice, s));
  RenderCamera rc;
  rc.width = cam.width; rc.height = cam.height;
  rc.fx = (float)cam.fx; rc.fy = (float)cam.fy;
  rc.cx = (float)cam.width / 2.0f; rc.cy = (float)cam.height / 2.0f;
  rc.z_near = (float)cam.near_; rc.z_max = (float)cam.far_ * 0.99f;
  RenderTargets rt;
  rt.bgr = d_bgr; rt.depth = d_depth; rt.mask = d_mask;
  rt.bgr_stride = px * 3; rt.depth_stride = px; rt.mask_stride = px;
  rt.rect = ws.rects.as<int>();
  launch_raster(d_tris, n_tri, ws.views.as<RenderView>(), n, rc, ws.zbuf.as<unsigned long long>(), ws.nz_abs.as<float>(), rt, s);
  CU(cudaGetLastError());
  return LM_OK;
}

lm_rect rect_of(const int r[4]) {  // (x_min, y_min, x_max, y_max) -> cv::Rect, empty -> zeros
  lm_rect o = {0, 0, 0, 0};
  if (r[2] >= 0) { o.x = r[0]; o.y = r[1]; o.width = r[2] - r[0] + 1; o.height = r[3] - r[1] + 1; }
  return o;
}

uint32_t pow2_at_least(uint32_t v) {
  uint32_t p = 64;
  while (p < v) p <<= 1;
  return p;
}

// addTemplate for n views whose sources / masks are device resident (d_src[v * M + m], d_mask[v]); rects = bounding boxes
// of the masks (x_min, y_min, x_max, y_max), host.  Appends the successful templates in view order.
int train_device_batch(lm_detector* d, int rows, int cols, int n, const void* const* d_src, const uint8_t* const* d_mask,
                       const int* rects, const char* class_id, int32_t* tids, lm_rect* bbs) {
  const int L = d->model.levels(), M = d->model.M();
  TrainWs& ws = d->train;
  if (upload_luts(d) != LM_OK) return LM_E_CUDA;
  const int lanes = std::min(n, LM_LANES);
  for (int i = 0; i < lanes; ++i) {
    Lane& ln = d->lane[i];
    if (ensure_quant_ws(d, ln, rows, cols) != LM_OK) return LM_E_CUDA;
    ln.lm_ready = false; ln.front_valid = false;
    for (int l = 0; l < L; ++l)
      if (ws.pb[i][l].ensure((size_t)(rows >> l) * (cols >> l)) != LM_OK ||
          ws.runs[i][l].ensure((size_t)(rows >> l) * (cols >> l) * 2) != LM_OK)
        return LM_E_CUDA;
    if (!ws.ev[i]) CU(cudaEventCreateWithFlags(&ws.ev[i], cudaEventDisableTiming));
  }
  // segment table: (view, level, modality); the candidates of a level lie inside the decimated bounding box of the mask
  const int S = n * L * M;
  if (ws.h_segs.ensure(sizeof(TrainSeg) * S) != LM_OK || ws.segs.ensure(sizeof(TrainSeg) * S) != LM_OK ||
      ws.feats.ensure((size_t)S * 64 * 4) != LM_OK || ws.h_feats.ensure((size_t)S * 64 * 4) != LM_OK)
    return LM_E_CUDA;
  TrainSeg* hs = ws.h_segs.as<TrainSeg>();
  size_t total = 0;
  for (int v = 0; v < n; ++v) {
    const int* r = rects + 4 * v;
    for (int l = 0; l < L; ++l) {
      uint32_t bound = 0;
      if (r[2] >= 0) {
        const int add = (1 << l) - 1;
        const int w = (r[2] >> l) - ((r[0] + add) >> l) + 1, h = (r[3] >> l) - ((r[1] + add) >> l) + 1;
        if (w > 0 && h > 0) bound = (uint32_t)w * (uint32_t)h;
      }
      for (int m = 0; m < M; ++m) {
        TrainSeg& sg = hs[(v * L + l) * M + m];
        std::memset(&sg, 0, sizeof(sg));
        sg.cap = pow2_at_least(bound);
        sg.off = (uint32_t)total;
        total += sg.cap;
        sg.cols = cols >> l;
        sg.type = d->model.mods[m].type;
        sg.nf = d->model.mods[m].num_features >> l;  // num_features /= 2 per level
      }
    }
  }
  if (total >= (1ull << 32)) return lm_fail(LM_E_INVALID, "training batch too large");
  if (ws.pool.ensure(total * 8) != LM_OK) return LM_E_CUDA;
  cudaStream_t s0 = d->lane[0].stream;
  CU(cudaMemcpyAsync(ws.segs.p, hs, sizeof(TrainSeg) * S, cudaMemcpyHostToDevice, s0));
  CU(cudaEventRecord(ws.ev[0], s0));
  for (int i = 1; i < lanes; ++i) CU(cudaStreamWaitEvent(d->lane[i].stream, ws.ev[0], 0));
  for (int v = 0; v < n; ++v) {
    const int li = v % lanes;
    Lane& ln = d->lane[li];
    for (int m = 0; m < M; ++m) { ln.src_ptr[0][m] = d_src[v * M + m]; ln.has_mask[m] = false; }
    ln.launches = 0;
    if (begin_chunk(d, ln, 1, 0, ln.stream) != LM_OK) return LM_E_CUDA;
    if (run_quantize(d, ln, 1, ln.stream) != LM_OK) return LM_E_CUDA;
    for (int m = 0; m < M; ++m) {
      const lm_modality_desc& md = d->model.mods[m];
      TrainViewParams tp;
      std::memset(&tp, 0, sizeof(tp));
      tp.n_levels = L; tp.cols0 = cols; tp.mask0 = d_mask[v];
      tp.thr_sq = md.strong_threshold * md.strong_threshold;
      int blocks = 0, ext = md.extract_threshold;
      for (int l = 0; l < L; ++l) {
        if (l > 0) ext /= 2;
        tp.extract_threshold[l] = ext;
        TrainLevel& lv = tp.lv[l];
        lv.quant = ln.quant_raw[l][m].as<uint8_t>();
        lv.mag = md.type == LM_COLOR_GRADIENT ? ln.mag[l][m].as<float>() : nullptr;
        lv.rows = rows >> l; lv.cols = cols >> l;
        lv.seg = (v * L + l) * M + m;
        lv.block_begin = blocks;
        blocks += train_blocks(lv.rows, lv.cols);
        tp.pb[l] = ws.pb[li][l].as<uint8_t>();
        tp.runs[l] = ws.runs[li][l].as<uint16_t>();
      }
      if (md.type == LM_COLOR_GRADIENT) launch_train_cg(tp, blocks, ws.segs.as<TrainSeg>(), ws.pool.as<unsigned long long>(), ln.stream);
      else launch_train_dn(tp, blocks, ws.segs.as<TrainSeg>(), ws.pool.as<unsigned long long>(), ln.stream);
    }
  }
  for (int i = 1; i < lanes; ++i) {
    CU(cudaEventRecord(ws.ev[i], d->lane[i].stream));
    CU(cudaStreamWaitEvent(s0, ws.ev[i], 0));
  }
  launch_train_finish(ws.segs.as<TrainSeg>(), S, ws.pool.as<unsigned long long>(), ws.feats.as<uint32_t>(), s0);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(hs, ws.segs.p, sizeof(TrainSeg) * S, cudaMemcpyDeviceToHost, s0));
  CU(cudaMemcpyAsync(ws.h_feats.p, ws.feats.p, (size_t)S * 64 * 4, cudaMemcpyDeviceToHost, s0));
  if (cudaStreamSynchronize(s0) != cudaSuccess) return lm_fail(LM_E_CUDA, "training kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
  // host tail: [OCV] cropTemplates + bookkeeping, in view order
  std::vector<TemplatePyramid>& tps = d->model.classes[class_id];  // the reference creates the class entry up front
  refresh_class_cache(d);
  ++d->model.version;
  const uint32_t* hf = ws.h_feats.as<uint32_t>();
  for (int v = 0; v < n; ++v) {
    bool ok = true;
    for (int i = 0; i < L * M; ++i) {
      const int ns = hs[v * L * M + i].n_sel;
      if (ns == -2) return lm_fail(LM_E_STATE, "training candidate pool overflow (view %d)", v);
      if (ns < 0) ok = false;
    }
    tids[v] = -1;
    if (bbs) { lm_rect z = {0, 0, 0, 0}; bbs[v] = z; }
    if (!ok) continue;
    TemplatePyramid tp((size_t)L * M);
    for (int l = 0; l < L; ++l)
      for (int m = 0; m < M; ++m) {
        const int sidx = (v * L + l) * M + m;
        Template& t = tp[(size_t)l * M + m];
        t.pyramid_level = l; t.width = -1; t.height = -1;
        t.features.resize((size_t)hs[sidx].n_sel);
        for (int k = 0; k < hs[sidx].n_sel; ++k) {
          const uint32_t w = hf[(size_t)sidx * 64 + k];
          t.features[k].x = (int)(w & 8191u); t.features[k].y = (int)((w >> 13) & 8191u); t.features[k].label = (int)(w >> 26);
        }
      }
    const lm_rect bb = crop_templates(tp);
    if (bbs) bbs[v] = bb;
    tps.push_back(tp);
    tids[v] = (int)tps.size() - 1;
  }
  return LM_OK;
}

int check_train_model(lm_detector* d, int rows, int cols) {
  const int L = d->model.levels();
  if ((rows >> (L - 1)) <= 0 || (cols >> (L - 1)) <= 0) return lm_fail(LM_E_INVALID, "image too small for %d pyramid levels", L);
  if (rows > 8191 || cols > 8191) return lm_fail(LM_E_INVALID, "training images are limited to 8191 x 8191");
  for (int m = 0; m < d->model.M(); ++m)
    if (d->model.mods[m].num_features > LM_MAX_FEATURES || d->model.mods[m].num_features < 1)
      return lm_fail(LM_E_INVALID, "num_features must be 1..63");
  return LM_OK;
}

}  // namespace

extern "C" {

int lm_mesh_create(const float* triangles, int n_triangles, lm_mesh** out) {
  if (!out || n_triangles < 0 || (n_triangles > 0 && !triangles)) return lm_fail(LM_E_INVALID, "NULL argument");
  lm_mesh* m = new lm_mesh();
  m->n = n_triangles;
  m->tris.assign(triangles, triangles + (size_t)n_triangles * 9);
  *out = m;
  return LM_OK;
}

int lm_mesh_load_stl(const char* path, lm_mesh** out) {
  if (!path || !out) return lm_fail(LM_E_INVALID, "NULL argument");
  std::ifstream f(path, std::ios::binary);
  if (!f) return lm_fail(LM_E_IO, "cannot open %s", path);
  std::stringstream ss;
  ss << f.rdbuf();
  std::vector<float> tris;
  std::string err;
  if (!parse_stl(ss.str(), tris, err)) return lm_fail(LM_E_IO, "%s: %s", path, err.c_str());
  lm_mesh* m = new lm_mesh();
  m->n = (int)(tris.size() / 9);
  m->tris.swap(tris);
  *out = m;
  return LM_OK;
}

int lm_mesh_num_triangles(const lm_mesh* mesh) { return mesh ? mesh->n : 0; }
int lm_mesh_get_triangles(const lm_mesh* mesh, float* dst) {
  if (!mesh || !dst) return lm_fail(LM_E_INVALID, "NULL argument");
  std::memcpy(dst, mesh->tris.data(), mesh->tris.size() * sizeof(float));
  return LM_OK;
}
void lm_mesh_destroy(lm_mesh* mesh) {
  if (!mesh) return;
  if (mesh->d_tris) { cudaSetDevice(mesh->device); cudaFree(mesh->d_tris); }
  delete mesh;
}

int lm_view_count(const lm_view_sphere* vs) {
  if (!sphere_valid(vs)) return lm_fail(LM_E_INVALID, "bad view sphere (n_points, angle_step, radius_step must be positive)");
  const SphereShape sh = sphere_shape(*vs);
  const long long n = (long long)vs->n_points * sh.n_angles * sh.n_radii;
  if (n > 0x7fffffffLL) return lm_fail(LM_E_INVALID, "view sphere too large");
  return (int)n;
}

int lm_view_params(const lm_view_sphere* vs, int index, double T[3], double up[3], float* radius, int32_t* point_index,
                   int32_t* angle_deg) {
  if (!sphere_valid(vs) || !T || !up) return lm_fail(LM_E_INVALID, "bad view sphere / NULL argument");
  const SphereShape sh = sphere_shape(*vs);
  const int per_point = sh.n_angles * sh.n_radii;
  if (index < 0 || index / per_point >= vs->n_points) return lm_fail(LM_E_NOTFOUND, "view index %d out of range", index);
  const int point = index / per_point, rem = index % per_point;
  const float r = sh.radii[rem / sh.n_angles];
  const int angle = vs->angle_min + (rem % sh.n_angles) * vs->angle_step;
  sphere_view(vs->n_points, point, angle, r, T, up);
  if (radius) *radius = r;
  if (point_index) *point_index = point;
  if (angle_deg) *angle_deg = angle;
  return LM_OK;
}

int lm_view_pose(const double T[3], const double up[3], double R[9], double t[3]) {
  if (!T || !up || !R || !t) return lm_fail(LM_E_INVALID, "NULL argument");
  if (!look_at(T, up, R, t)) return lm_fail(LM_E_INVALID, "degenerate camera position / up vector");
  return LM_OK;
}

int lm_render_views(lm_detector* d, const lm_mesh* mesh, const lm_camera* cam, const double* T, const double* up,
                    int n_views, uint8_t* bgr, uint16_t* depth, uint8_t* mask, lm_rect* rects) {
  if (!d || !mesh || !T || !up || n_views < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  if (check_camera(cam) != LM_OK) return LM_E_INVALID;
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  const float* d_tris = nullptr;
  if (mesh_on_device(d, mesh, &d_tris) != LM_OK) return LM_E_CUDA;
  TrainWs& ws = d->train;
  const size_t px = (size_t)cam->width * cam->height;
  cudaStream_t s = d->lane[0].stream;
  for (int v0 = 0; v0 < n_views; v0 += kTrainBatch) {
    const int n = std::min(kTrainBatch, n_views - v0);
    if (ws.src[0].ensure(px * 3 * n) != LM_OK || ws.src[1].ensure(px * 2 * n) != LM_OK || ws.mask.ensure(px * n) != LM_OK ||
        ws.h_rects.ensure(16 * (size_t)n) != LM_OK)
      return LM_E_CUDA;
    int rc = render_batch(d, d_tris, mesh->n, *cam, T + 3 * (size_t)v0, up + 3 * (size_t)v0, n, bgr ? ws.src[0].as<uint8_t>() : nullptr,
                          depth ? ws.src[1].as<uint16_t>() : nullptr, mask ? ws.mask.as<uint8_t>() : nullptr, s);
    if (rc != LM_OK) return rc;
    if (bgr) CU(cudaMemcpyAsync(bgr + px * 3 * v0, ws.src[0].p, px * 3 * n, cudaMemcpyDeviceToHost, s));
    if (depth) CU(cudaMemcpyAsync(depth + px * v0, ws.src[1].p, px * 2 * n, cudaMemcpyDeviceToHost, s));
    if (mask) CU(cudaMemcpyAsync(mask + px * v0, ws.mask.p, px * n, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ws.h_rects.p, ws.rects.p, 16 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return lm_fail(LM_E_CUDA, "rasteriser failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rects)
      for (int v = 0; v < n; ++v) rects[v0 + v] = rect_of(ws.h_rects.as<int>() + 4 * v);
  }
  return LM_OK;
}

int lm_train_views(lm_detector* d, const lm_mesh* mesh, const lm_camera* cam, const double* T, const double* up,
                   int n_views, const char* class_id, int32_t* template_ids, lm_rect* bounding_boxes, lm_rect* mask_rects,
                   uint16_t* centre_depth_mm) {
  if (!d || !mesh || !T || !up || !class_id || !template_ids || n_views < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  if (check_camera(cam) != LM_OK) return LM_E_INVALID;
  const int M = d->model.M();
  const int rows = cam->height, cols = cam->width;
  if (check_train_model(d, rows, cols) != LM_OK) return LM_E_INVALID;
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  const float* d_tris = nullptr;
  if (mesh_on_device(d, mesh, &d_tris) != LM_OK) return LM_E_CUDA;
  TrainWs& ws = d->train;
  const size_t px = (size_t)rows * cols;
  cudaStream_t s = d->lane[0].stream;
  d->model.classes[class_id];  // Detector::addTemplate creates the class entry even when every view fails
  refresh_class_cache(d);
  ++d->model.version;
  for (int v0 = 0; v0 < n_views; v0 += kTrainBatch) {
    const int n = std::min(kTrainBatch, n_views - v0);
    // one rendered image pool per source type; modalities of the same type share it
    if (ws.src[0].ensure(px * 3 * n) != LM_OK || ws.src[1].ensure(px * 2 * n) != LM_OK || ws.mask.ensure(px * n) != LM_OK ||
        ws.h_rects.ensure(16 * (size_t)n + 2 * (size_t)n) != LM_OK)
      return LM_E_CUDA;
    int rc = render_batch(d, d_tris, mesh->n, *cam, T + 3 * (size_t)v0, up + 3 * (size_t)v0, n, ws.src[0].as<uint8_t>(),
                          ws.src[1].as<uint16_t>(), ws.mask.as<uint8_t>(), s);
    if (rc != LM_OK) return rc;
    CU(cudaMemcpyAsync(ws.h_rects.p, ws.rects.p, 16 * (size_t)n, cudaMemcpyDeviceToHost, s));
    uint16_t* h_centre = reinterpret_cast<uint16_t*>(ws.h_rects.as<uint8_t>() + 16 * (size_t)n);
    if (centre_depth_mm)  // one strided copy: the centre pixel of every view's depth image
      CU(cudaMemcpy2DAsync(h_centre, 2, ws.src[1].as<uint16_t>() + (size_t)(rows / 2) * cols + cols / 2, px * 2, 2, n,
                           cudaMemcpyDeviceToHost, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return lm_fail(LM_E_CUDA, "rasteriser failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (centre_depth_mm) std::memcpy(centre_depth_mm + v0, h_centre, 2 * (size_t)n);
    std::vector<int> rects(ws.h_rects.as<int>(), ws.h_rects.as<int>() + 4 * n);
    std::vector<const void*> srcs((size_t)n * M);
    std::vector<const uint8_t*> masks((size_t)n);
    for (int v = 0; v < n; ++v) {
      for (int m = 0; m < M; ++m)
        srcs[(size_t)v * M + m] = d->model.mods[m].type == LM_COLOR_GRADIENT ? (const void*)(ws.src[0].as<uint8_t>() + px * 3 * v)
                                                                             : (const void*)(ws.src[1].as<uint16_t>() + px * v);
      masks[v] = ws.mask.as<uint8_t>() + px * v;
      if (mask_rects) mask_rects[v0 + v] = rect_of(&rects[4 * v]);
    }
    rc = train_device_batch(d, rows, cols, n, srcs.data(), masks.data(), rects.data(), class_id, template_ids + v0,
                            bounding_boxes ? bounding_boxes + v0 : nullptr);
    if (rc != LM_OK) return rc;
  }
  return LM_OK;
}

int lm_add_templates_batch(lm_detector* d, const lm_image* sources, const lm_image* masks, int n_views, int n_sources,
                           const char* class_id, int32_t* template_ids, lm_rect* bounding_boxes) {
  if (!d || !class_id || !template_ids || n_views < 0 || (n_views > 0 && (!sources || !masks))) return lm_fail(LM_E_INVALID, "NULL argument");
  const int M = d->model.M();
  if (n_sources != M) return lm_fail(LM_E_INVALID, "sources.size() == modalities.size() violated (%d vs %d)", n_sources, M);
  if (n_views == 0) return LM_OK;
  const int rows = sources[0].rows, cols = sources[0].cols;
  for (int v = 0; v < n_views; ++v) {
    for (int m = 0; m < M; ++m) {
      const lm_image& im = sources[(size_t)v * M + m];
      if (!im.data || im.rows != rows || im.cols != cols || im.type != expected_src_type(d->model.mods[m]))
        return lm_fail(LM_E_INVALID, "view %d source %d: size / type mismatch", v, m);
    }
    const lm_image& mk = masks[v];
    if (!mk.data || mk.type != LM_8UC1 || mk.rows != rows || mk.cols != cols)
      return lm_fail(LM_E_INVALID, "view %d: an object mask of the sources' size is required", v);
  }
  if (check_train_model(d, rows, cols) != LM_OK) return LM_E_INVALID;
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  TrainWs& ws = d->train;
  const size_t px = (size_t)rows * cols;
  Lane& l0 = d->lane[0];
  cudaStream_t s = l0.stream;
  d->model.classes[class_id];
  refresh_class_cache(d);
  ++d->model.version;
  for (int v0 = 0; v0 < n_views; v0 += kTrainBatch) {
    const int n = std::min(kTrainBatch, n_views - v0);
    size_t stage = 0;
    for (int m = 0; m < M; ++m) {
      const size_t bytes = src_row_bytes(expected_src_type(d->model.mods[m]), cols) * rows;
      if (ws.src[m].ensure(bytes * n) != LM_OK) return LM_E_CUDA;
      stage += ((bytes + 255) & ~(size_t)255) * n;
    }
    stage += ((px + 255) & ~(size_t)255) * n;
    if (ws.mask.ensure(px * n) != LM_OK || ws.rects.ensure(16 * (size_t)n) != LM_OK || ws.h_rects.ensure(16 * (size_t)n) != LM_OK ||
        l0.stage_in.ensure(stage) != LM_OK)
      return LM_E_CUDA;
    std::vector<const void*> srcs((size_t)n * M);
    std::vector<const uint8_t*> dmasks((size_t)n);
    size_t off = 0;
    for (int v = 0; v < n; ++v) {
      for (int m = 0; m < M; ++m) {
        const size_t bytes = src_row_bytes(expected_src_type(d->model.mods[m]), cols) * rows;
        uint8_t* dst = ws.src[m].as<uint8_t>() + bytes * v;
        if (upload_image(l0, sources[(size_t)(v0 + v) * M + m], dst, &off) != LM_OK) return LM_E_CUDA;
        srcs[(size_t)v * M + m] = dst;
      }
      uint8_t* dm = ws.mask.as<uint8_t>() + px * v;
      if (upload_image(l0, masks[v0 + v], dm, &off) != LM_OK) return LM_E_CUDA;
      dmasks[v] = dm;
      launch_mask_rect(dm, cols, rows, ws.rects.as<int>() + 4 * v, s);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ws.h_rects.p, ws.rects.p, 16 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return lm_fail(LM_E_CUDA, "mask upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    std::vector<int> rects(ws.h_rects.as<int>(), ws.h_rects.as<int>() + 4 * n);
    int rc = train_device_batch(d, rows, cols, n, srcs.data(), dmasks.data(), rects.data(), class_id, template_ids + v0,
                                bounding_boxes ? bounding_boxes + v0 : nullptr);
    if (rc != LM_OK) return rc;
  }
  return LM_OK;
}

int lm_depth_diff_batch(lm_detector* d, const lm_image* scene, const lm_mesh* mesh, const lm_camera* cam, const double* T,
                        const double* up, const int32_t* x, const int32_t* y, int n, double* out) {
  if (!d || !scene || !scene->data || !mesh || !T || !up || !x || !y || !out || n < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  if (scene->type != LM_16UC1) return lm_fail(LM_E_INVALID, "scene depth must be LM_16UC1");
  if (check_camera(cam) != LM_OK) return LM_E_INVALID;
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  const float* d_tris = nullptr;
  if (mesh_on_device(d, mesh, &d_tris) != LM_OK) return LM_E_CUDA;
  TrainWs& ws = d->train;
  Lane& l0 = d->lane[0];
  cudaStream_t s = l0.stream;
  const size_t px = (size_t)cam->width * cam->height, spx = (size_t)scene->rows * scene->cols;
  if (ws.scene.ensure(spx * 2) != LM_OK || l0.stage_in.ensure(spx * 2 + 256) != LM_OK) return LM_E_CUDA;
  size_t off = 0;
  if (upload_image(l0, *scene, ws.scene.p, &off) != LM_OK) return LM_E_CUDA;
  for (int v0 = 0; v0 < n; v0 += kTrainBatch) {
    const int nb = std::min(kTrainBatch, n - v0);
    if (ws.src[1].ensure(px * 2 * nb) != LM_OK || ws.mask.ensure(px * nb) != LM_OK || ws.h_rects.ensure(16 * (size_t)nb + 16 * (size_t)nb) != LM_OK ||
        ws.diff.ensure(16 * (size_t)nb) != LM_OK)
      return LM_E_CUDA;
    int rc = render_batch(d, d_tris, mesh->n, *cam, T + 3 * (size_t)v0, up + 3 * (size_t)v0, nb, nullptr, ws.src[1].as<uint16_t>(),
                          ws.mask.as<uint8_t>(), s);
    if (rc != LM_OK) return rc;
    CU(cudaMemcpyAsync(ws.h_rects.p, ws.rects.p, 16 * (size_t)nb, cudaMemcpyDeviceToHost, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return lm_fail(LM_E_CUDA, "rasteriser failed: %s", cudaGetErrorString(cudaGetLastError()));
    const int* hr = ws.h_rects.as<int>();
    for (int v = 0; v < nb; ++v) {
      const lm_rect r = rect_of(hr + 4 * v);
      const int xs = x[v0 + v], ys = y[v0 + v];
      if (r.width > 0 && (xs < 0 || ys < 0 || xs + r.width > scene->cols || ys + r.height > scene->rows))
        return lm_fail(LM_E_INVALID, "hypothesis %d: the %dx%d template crop at (%d, %d) leaves the %dx%d scene image", v0 + v, r.width,
                    r.height, xs, ys, scene->cols, scene->rows);
      launch_depth_diff(ws.scene.as<uint16_t>(), scene->cols, ws.src[1].as<uint16_t>() + px * v, ws.mask.as<uint8_t>() + px * v,
                        cam->width, xs, ys, r.x, r.y, r.width, r.height, ws.diff.as<unsigned long long>() + 2 * v, s);
    }
    CU(cudaGetLastError());
    unsigned long long* hd = reinterpret_cast<unsigned long long*>(ws.h_rects.as<uint8_t>() + 16 * (size_t)nb);
    CU(cudaMemcpyAsync(hd, ws.diff.p, 16 * (size_t)nb, cudaMemcpyDeviceToHost, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return lm_fail(LM_E_CUDA, "depth_diff failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (int v = 0; v < nb; ++v) out[v0 + v] = (double)hd[2 * v] / ((double)hd[2 * v + 1] * 1000.0);
  }
  return LM_OK;
}

}  // extern "C"

// ================================================================================================ pose table
// writeLinemodTemplateParams (/root/reference/src/renderer.cpp:72-123) / readLinemodTemplateParams
// (src/rgbdDetector.cpp:1681-1749): cv::FileStorage YAML, one "Template i" map per template + the renderer_* scalars.
namespace {

void write_matrix(lmyaml::Writer& w, const char* key, int rows, int cols, const double* d, const float* f) {
  w.key(key);
  w.begin_map_tagged("!!opencv-matrix");
  w.key("rows"); w.write_int(rows);
  w.key("cols"); w.write_int(cols);
  w.key("dt"); w.write_string(d ? "d" : "f");
  w.key("data");
  w.begin_seq(true);
  for (int i = 0; i < rows * cols; ++i) {
    if (d) w.write_double(d[i]);
    else w.write_float(f[i]);
  }
  w.end_seq();
  w.end_map();
}

bool read_matrix(const lmyaml::Node& n, int count, double* d, float* f) {
  const lmyaml::Node& data = n["data"];
  if (data.kind != lmyaml::Node::SEQ || (int)data.size() != count) return false;
  for (int i = 0; i < count; ++i) {
    if (d) d[i] = data.num(i);
    else f[i] = (float)data.num(i);
  }
  return true;
}

}  // namespace

extern "C" {

int lm_write_renderer_params(const char* path, const lm_template_pose* poses, size_t n, const lm_renderer_params* p) {
  if (!path || (n && !poses) || !p) return lm_fail(LM_E_INVALID, "NULL argument");
  lmyaml::Writer w;
  for (size_t i = 0; i < n; ++i) {
    const lm_template_pose& t = poses[i];
    w.key("Template " + std::to_string(i));
    w.begin_map();
    w.key("ID"); w.write_int((int)i);
    write_matrix(w, "R", 3, 3, t.R, nullptr);
    write_matrix(w, "T", 3, 1, t.T, nullptr);
    write_matrix(w, "K", 3, 3, nullptr, t.K);
    w.key("D"); w.write_double(t.D);
    w.key("Ori_dist"); w.write_double(t.ori_dist);
    w.key("Rect");
    w.begin_seq(true);
    w.write_int(t.rect.x); w.write_int(t.rect.y); w.write_int(t.rect.width); w.write_int(t.rect.height);
    w.end_seq();
    w.end_map();
  }
  w.key("renderer_n_points"); w.write_int(p->n_points);
  w.key("renderer_angle_step"); w.write_int(p->angle_step);
  w.key("renderer_radius_min"); w.write_double(p->radius_min);
  w.key("renderer_radius_max"); w.write_double(p->radius_max);
  w.key("renderer_radius_step"); w.write_double(p->radius_step);
  w.key("renderer_width"); w.write_int(p->width);
  w.key("renderer_height"); w.write_int(p->height);
  w.key("renderer_focal_length_x"); w.write_double(p->fx);
  w.key("renderer_focal_length_y"); w.write_double(p->fy);
  w.key("renderer_near"); w.write_double(p->near_);
  w.key("renderer_far"); w.write_double(p->far_);
  std::string err;
  if (!w.save(path, err)) return lm_fail(LM_E_IO, "%s", err.c_str());
  return LM_OK;
}

int lm_read_renderer_params(const char* path, lm_template_pose** out_poses, size_t* out_n, lm_renderer_params* p) {
  if (!path || !out_poses || !out_n || !p) return lm_fail(LM_E_INVALID, "NULL argument");
  lmyaml::Node root;
  std::string err;
  if (!lmyaml::parse_file(path, root, err)) return lm_fail(LM_E_IO, "%s: %s", path, err.c_str());
  std::vector<lm_template_pose> poses;
  for (size_t i = 0;; ++i) {  // the reference reads "Template 0", "Template 1", ... until the first missing key
    const lmyaml::Node& t = root["Template " + std::to_string(i)];
    if (t.empty()) break;
    lm_template_pose ps;
    std::memset(&ps, 0, sizeof(ps));
    const lmyaml::Node& rc = t["Rect"];
    if (!read_matrix(t["R"], 9, ps.R, nullptr) || !read_matrix(t["T"], 3, ps.T, nullptr) || !read_matrix(t["K"], 9, nullptr, ps.K) ||
        !t["D"].as_double(ps.D) || !t["Ori_dist"].as_double(ps.ori_dist) || rc.kind != lmyaml::Node::SEQ || rc.size() != 4)
      return lm_fail(LM_E_IO, "%s: malformed entry \"Template %zu\"", path, i);
    ps.rect.x = (int)rc.num(0); ps.rect.y = (int)rc.num(1); ps.rect.width = (int)rc.num(2); ps.rect.height = (int)rc.num(3);
    poses.push_back(ps);
  }
  std::memset(p, 0, sizeof(*p));
  bool ok = root["renderer_n_points"].as_int(p->n_points) && root["renderer_angle_step"].as_int(p->angle_step) &&
            root["renderer_radius_min"].as_double(p->radius_min) && root["renderer_radius_max"].as_double(p->radius_max) &&
            root["renderer_radius_step"].as_double(p->radius_step) && root["renderer_width"].as_int(p->width) &&
            root["renderer_height"].as_int(p->height) && root["renderer_focal_length_x"].as_double(p->fx) &&
            root["renderer_focal_length_y"].as_double(p->fy) && root["renderer_near"].as_double(p->near_) &&
            root["renderer_far"].as_double(p->far_);
  if (!ok) return lm_fail(LM_E_IO, "%s: renderer_* parameters missing", path);
  *out_n = poses.size();
  *out_poses = nullptr;
  if (!poses.empty()) {
    *out_poses = (lm_template_pose*)std::malloc(poses.size() * sizeof(lm_template_pose));
    if (!*out_poses) return lm_fail(LM_E_INVALID, "out of memory");
    std::memcpy(*out_poses, poses.data(), poses.size() * sizeof(lm_template_pose));
  }
  return LM_OK;
}

void lm_free_poses(lm_template_pose* poses) { std::free(poses); }

}  // extern "C"
