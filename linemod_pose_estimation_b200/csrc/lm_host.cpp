// lm_host.cpp -- template extraction (host half of addTemplate) and templates.yml persistence.  See lm_host.hpp.
#include "lm_host.hpp"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "lm_yaml.hpp"

namespace lm {

const char* modality_name(int type) { return type == LM_COLOR_GRADIENT ? "ColorGradient" : "DepthNormal"; }
bool modality_from_name(const std::string& name, int& type) {
  if (name == "ColorGradient") { type = LM_COLOR_GRADIENT; return true; }
  if (name == "DepthNormal") { type = LM_DEPTH_NORMAL; return true; }
  return false;
}
lm_modality_desc default_modality(int type) {
  lm_modality_desc d;
  d.type = type;
  d.weak_threshold = 10.0f; d.strong_threshold = 55.0f;
  d.distance_threshold = 2000; d.difference_threshold = 50; d.extract_threshold = 2;
  d.num_features = 63;
  return d;
}

// ------------------------------------------------------------------------------------------------ extraction
namespace {

struct Scored {
  Feature f;
  float score;
};
struct ByScoreDesc {
  bool operator()(const Scored& a, const Scored& b) const { return a.score > b.score; }
};

inline int label_of(uint8_t q) {  // [OCV] getLabel: one-hot byte -> bit index
  return (q && !(q & (q - 1))) ? __builtin_ctz(q) : -1;
}

// 3x3 minimum with replicated border, `iterations` passes ([OCV] cv::erode default kernel, BORDER_REPLICATE).
void erode_3x3(std::vector<uint8_t>& img, int rows, int cols, int iterations) {
  std::vector<uint8_t> rowmin((size_t)rows * cols);
  for (int it = 0; it < iterations; ++it) {
    for (int y = 0; y < rows; ++y) {
      const uint8_t* s = &img[(size_t)y * cols];
      uint8_t* d = &rowmin[(size_t)y * cols];
      for (int x = 0; x < cols; ++x) {
        uint8_t l = s[x > 0 ? x - 1 : 0], r = s[x + 1 < cols ? x + 1 : cols - 1];
        d[x] = std::min(s[x], std::min(l, r));
      }
    }
    for (int y = 0; y < rows; ++y) {
      const uint8_t* u = &rowmin[(size_t)(y > 0 ? y - 1 : 0) * cols];
      const uint8_t* c = &rowmin[(size_t)y * cols];
      const uint8_t* b = &rowmin[(size_t)(y + 1 < rows ? y + 1 : rows - 1) * cols];
      uint8_t* d = &img[(size_t)y * cols];
      for (int x = 0; x < cols; ++x) d[x] = std::min(c[x], std::min(u[x], b[x]));
    }
  }
}

// Chessboard distance to the nearest zero pixel ([OCV] cv::distanceTransform(CV_DIST_C, 3)): forward / backward 3x3
// chamfer in 16.16 fixed point over a frame whose outside is "infinitely far" (INT_MAX >> 2), result scaled back.
void chessboard_distance(const uint8_t* src, int rows, int cols, std::vector<float>& dst) {
  const int kOne = 1 << 16, kFar = INT_MAX >> 2;
  const int stride = cols + 2;
  std::vector<int> d((size_t)(rows + 2) * stride, kFar);
  for (int y = 0; y < rows; ++y) {
    int* row = &d[(size_t)(y + 1) * stride + 1];
    const int* up = row - stride;
    for (int x = 0; x < cols; ++x) {
      if (!src[(size_t)y * cols + x]) { row[x] = 0; continue; }
      int best = std::min(std::min(up[x - 1], up[x]), std::min(up[x + 1], row[x - 1]));
      row[x] = best + kOne;
    }
  }
  dst.resize((size_t)rows * cols);
  const float scale = 1.0f / kOne;
  for (int y = rows - 1; y >= 0; --y) {
    int* row = &d[(size_t)(y + 1) * stride + 1];
    const int* dn = row + stride;
    for (int x = cols - 1; x >= 0; --x) {
      int v = row[x];
      if (v > kOne) {
        int best = std::min(std::min(dn[x + 1], dn[x]), std::min(dn[x - 1], row[x + 1])) + kOne;
        if (best < v) { v = best; row[x] = v; }
      }
      dst[(size_t)y * cols + x] = (float)v * scale;
    }
  }
}

// [OCV] QuantizedPyramid::selectScatteredFeatures
void select_scattered(const std::vector<Scored>& cands, size_t want, float distance, std::vector<Feature>& out) {
  out.clear();
  float dist_sq = distance * distance;
  size_t i = 0;
  while (out.size() < want) {
    const Feature& c = cands[i].f;
    bool far_enough = true;
    for (size_t j = 0; j < out.size() && far_enough; ++j) {
      int dx = c.x - out[j].x, dy = c.y - out[j].y;
      far_enough = (float)(dx * dx + dy * dy) >= dist_sq;
    }
    if (far_enough) out.push_back(c);
    if (++i == cands.size()) {  // wrap: relax the spacing and sweep again
      i = 0;
      distance -= 1.0f;
      dist_sq = distance * distance;
    }
  }
}

}  // namespace

bool extract_color_gradient(const uint8_t* quantized, const float* magnitude, const uint8_t* mask, int rows, int cols,
                            float strong_threshold, int num_features, int level, Template& out) {
  const size_t n = (size_t)rows * cols;
  std::vector<uint8_t> ring;  // mask minus its erosion: the 1-px silhouette ring (OpenCV 2.4 behaviour)
  if (mask) {
    ring.assign(mask, mask + n);
    erode_3x3(ring, rows, cols, 1);
    for (size_t i = 0; i < n; ++i) ring[i] = mask[i] > ring[i] ? (uint8_t)(mask[i] - ring[i]) : 0;
  }
  const float thr_sq = strong_threshold * strong_threshold;
  std::vector<Scored> cands;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      size_t i = (size_t)r * cols + c;
      if (mask && !ring[i]) continue;
      uint8_t q = quantized[i];
      if (!q || !(magnitude[i] > thr_sq)) continue;
      Scored s;
      s.f.x = c; s.f.y = r; s.f.label = label_of(q); s.score = magnitude[i];
      cands.push_back(s);
    }
  if (cands.size() < (size_t)num_features) return false;
  std::stable_sort(cands.begin(), cands.end(), ByScoreDesc());
  float distance = (float)(cands.size() / (size_t)num_features + 1);
  select_scattered(cands, (size_t)num_features, distance, out.features);
  out.width = -1; out.height = -1; out.pyramid_level = level;
  return true;
}

bool extract_depth_normal(const uint8_t* normal, const uint8_t* mask, int rows, int cols, int num_features,
                          int extract_threshold, int level, Template& out) {
  const size_t n = (size_t)rows * cols;
  std::vector<uint8_t> inner;  // mask eroded twice: features right on the border are unreliable
  if (mask) {
    inner.assign(mask, mask + n);
    erode_3x3(inner, rows, cols, 2);
  }
  std::vector<float> dist[8];
  std::vector<uint8_t> plane(n, 0);
  for (int b = 0; b < 8; ++b) {
    // the reference reuses one temp image across labels: temp.setTo(1<<b, mask); temp &= normal
    for (size_t i = 0; i < n; ++i) {
      if (!mask || inner[i]) plane[i] = (uint8_t)(1 << b);
      plane[i] &= normal[i];
    }
    chessboard_distance(plane.data(), rows, cols, dist[b]);
  }
  int per_label[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<Scored> cands;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      size_t i = (size_t)r * cols + c;
      if (mask && !inner[i]) continue;
      uint8_t q = normal[i];
      if (q == 0 || q == 255) continue;
      int label = label_of(q);
      if (label < 0) continue;
      float score = dist[label][i];
      if (score >= (float)extract_threshold) {
        Scored s;
        s.f.x = c; s.f.y = r; s.f.label = label; s.score = score;
        cands.push_back(s);
        ++per_label[label];
      }
    }
  if (cands.size() < (size_t)num_features) return false;
  for (size_t i = 0; i < cands.size(); ++i) cands[i].score /= (float)per_label[cands[i].f.label];
  std::stable_sort(cands.begin(), cands.end(), ByScoreDesc());
  size_t area_px = n;
  if (mask) { area_px = 0; for (size_t i = 0; i < n; ++i) area_px += inner[i] != 0; }
  float distance = sqrtf((float)area_px) / sqrtf((float)num_features) + 1.5f;
  select_scattered(cands, (size_t)num_features, distance, out.features);
  out.width = -1; out.height = -1; out.pyramid_level = level;
  return true;
}

lm_rect crop_templates(TemplatePyramid& tp) {
  int min_x = INT_MAX, min_y = INT_MAX, max_x = INT_MIN, max_y = INT_MIN;
  for (const Template& t : tp)
    for (const Feature& f : t.features) {
      int x = f.x << t.pyramid_level, y = f.y << t.pyramid_level;
      min_x = std::min(min_x, x); max_x = std::max(max_x, x);
      min_y = std::min(min_y, y); max_y = std::max(max_y, y);
    }
  if (min_x % 2 == 1) --min_x;  // the reference keeps the origin even so that every level shifts by whole pixels
  if (min_y % 2 == 1) --min_y;
  for (Template& t : tp) {
    t.width = (max_x - min_x) >> t.pyramid_level;
    t.height = (max_y - min_y) >> t.pyramid_level;
    int sx = min_x >> t.pyramid_level, sy = min_y >> t.pyramid_level;
    for (Feature& f : t.features) { f.x -= sx; f.y -= sy; }
  }
  lm_rect r = {min_x, min_y, max_x - min_x, max_y - min_y};
  return r;
}

// ------------------------------------------------------------------------------------------------ persistence
namespace {

using lmyaml::Node;

bool node_int(const Node& n, const char* what, int& v, std::string& err) {
  if (!n.as_int(v)) { err = std::string("missing or non-numeric '") + what + "'"; return false; }
  return true;
}

// [OCV] Detector::read
bool read_header(const Node& root, HostModel& model, std::string& err) {
  int levels = 0;
  if (!node_int(root["pyramid_levels"], "pyramid_levels", levels, err)) return false;
  const Node& tn = root["T"];
  if (tn.kind != Node::SEQ) { err = "missing 'T'"; return false; }
  model.T.clear();
  for (size_t i = 0; i < tn.size(); ++i) model.T.push_back((int)std::lrint(tn.num(i)));
  if ((int)model.T.size() != levels) { err = "pyramid_levels does not match T"; return false; }
  if (levels < 1 || levels > LM_MAX_LEVELS) { err = "unsupported pyramid_levels"; return false; }
  model.mods.clear();
  const Node& mn = root["modalities"];
  if (mn.kind != Node::SEQ) { err = "missing 'modalities'"; return false; }
  for (size_t i = 0; i < mn.size(); ++i) {
    const Node& m = mn.at(i);
    int type;
    if (!modality_from_name(m["type"].sval, type)) { err = "unknown modality type '" + m["type"].sval + "'"; return false; }
    lm_modality_desc d = default_modality(type);
    double v;
    if (type == LM_COLOR_GRADIENT) {
      if (m["weak_threshold"].as_double(v)) d.weak_threshold = (float)v;
      if (m["strong_threshold"].as_double(v)) d.strong_threshold = (float)v;
    } else {
      m["distance_threshold"].as_int(d.distance_threshold);
      m["difference_threshold"].as_int(d.difference_threshold);
      m["extract_threshold"].as_int(d.extract_threshold);
    }
    m["num_features"].as_int(d.num_features);
    model.mods.push_back(d);
  }
  if (model.mods.empty() || model.mods.size() > LM_MAX_MODALITIES) { err = "unsupported number of modalities"; return false; }
  model.classes.clear();
  ++model.version;
  return true;
}

// Every template pyramid that enters a HostModel from outside (templates.yml, class files, the binary cache) passes this:
// the packer (ensure_pack) and the kernels rely on pyramid size == levels * M, labels 0..7, |x|, |y| <= 4095 and at most
// LM_MAX_FEATURES features -- a malformed or foreign file must end in LM_E_IO, never in a device fault or silent garbage.
bool validate_pyramid(const HostModel& model, const TemplatePyramid& tp, std::string& err) {
  if (tp.size() != (size_t)model.levels() * (size_t)model.M()) {
    err = "template pyramid has " + std::to_string(tp.size()) + " templates, expected levels * modalities = " +
          std::to_string(model.levels() * model.M());
    return false;
  }
  for (const Template& t : tp) {
    if (t.features.size() > LM_MAX_FEATURES) { err = "features.size() <= 63 violated (" + std::to_string(t.features.size()) + ")"; return false; }
    if (t.width < 0 || t.height < 0 || t.width > 8191 || t.height > 8191) { err = "template width / height out of range"; return false; }
    for (const Feature& f : t.features) {
      if (f.label < 0 || f.label > 7) { err = "feature label " + std::to_string(f.label) + " outside 0..7"; return false; }
      if (f.x < -4095 || f.x > 4095 || f.y < -4095 || f.y > 4095) { err = "feature coordinate outside +-4095"; return false; }
    }
  }
  return true;
}

// [OCV] Detector::readClass (class_id_override empty)
bool read_class(const Node& cn, HostModel& model, std::string& err) {
  const Node& mods = cn["modalities"];
  if (mods.kind != Node::SEQ || mods.size() != model.mods.size()) { err = "class modalities do not match the detector"; return false; }
  for (size_t i = 0; i < mods.size(); ++i)
    if (mods.at(i).sval != modality_name(model.mods[i].type)) { err = "class modality '" + mods.at(i).sval + "' does not match the detector"; return false; }
  int levels = 0;
  if (!node_int(cn["pyramid_levels"], "pyramid_levels", levels, err)) return false;
  if (levels != model.levels()) { err = "class pyramid_levels does not match the detector"; return false; }
  std::string class_id = cn["class_id"].sval;
  if (model.classes.count(class_id)) { err = "detector already has class '" + class_id + "'"; return false; }
  const Node& tps = cn["template_pyramids"];
  std::vector<TemplatePyramid> out(tps.kind == Node::SEQ ? tps.size() : 0);
  for (size_t i = 0; i < out.size(); ++i) {
    const Node& tpn = tps.at(i);
    int tid = -1;
    if (!node_int(tpn["template_id"], "template_id", tid, err)) return false;
    if (tid != (int)i) { err = "template_id out of sequence"; return false; }
    const Node& tn = tpn["templates"];
    size_t nt = tn.kind == Node::SEQ ? tn.size() : 0;
    out[i].resize(nt);
    for (size_t j = 0; j < nt; ++j) {
      const Node& t = tn.at(j);
      Template& dst = out[i][j];
      if (!node_int(t["width"], "width", dst.width, err) || !node_int(t["height"], "height", dst.height, err) ||
          !node_int(t["pyramid_level"], "pyramid_level", dst.pyramid_level, err))
        return false;
      const Node& fn = t["features"];
      size_t nf = fn.kind == Node::SEQ ? fn.size() : 0;
      dst.features.resize(nf);
      for (size_t k = 0; k < nf; ++k) {
        const Node& f = fn.at(k);
        if (f.kind != Node::SEQ || f.size() != 3) { err = "feature is not an [x, y, label] triple"; return false; }
        dst.features[k].x = (int)std::lrint(f.num(0));
        dst.features[k].y = (int)std::lrint(f.num(1));
        dst.features[k].label = (int)std::lrint(f.num(2));
      }
    }
    if (!validate_pyramid(model, out[i], err)) { err = "class '" + class_id + "' template " + std::to_string(i) + ": " + err; return false; }
  }
  model.classes[class_id].swap(out);
  ++model.version;
  return true;
}

// [OCV] Detector::write
void write_header(const HostModel& model, lmyaml::Writer& w) {
  w.key("pyramid_levels"); w.write_int(model.levels());
  w.key("T"); w.begin_seq(true);
  for (int t : model.T) w.write_int(t);
  w.end_seq();
  w.key("modalities"); w.begin_seq(false);
  for (const lm_modality_desc& d : model.mods) {
    w.begin_map();
    w.key("type"); w.write_string(modality_name(d.type));
    if (d.type == LM_COLOR_GRADIENT) {
      w.key("weak_threshold"); w.write_float(d.weak_threshold);
      w.key("num_features"); w.write_int(d.num_features);
      w.key("strong_threshold"); w.write_float(d.strong_threshold);
    } else {
      w.key("distance_threshold"); w.write_int(d.distance_threshold);
      w.key("difference_threshold"); w.write_int(d.difference_threshold);
      w.key("num_features"); w.write_int(d.num_features);
      w.key("extract_threshold"); w.write_int(d.extract_threshold);
    }
    w.end_map();
  }
  w.end_seq();
}

// [OCV] Detector::writeClass
void write_class(const HostModel& model, const std::string& class_id, const std::vector<TemplatePyramid>& tps,
                 lmyaml::Writer& w) {
  w.key("class_id"); w.write_string(class_id);
  w.key("modalities"); w.begin_seq(true);
  for (const lm_modality_desc& d : model.mods) w.write_string(modality_name(d.type));
  w.end_seq();
  w.key("pyramid_levels"); w.write_int(model.levels());
  w.key("template_pyramids"); w.begin_seq(false);
  for (size_t i = 0; i < tps.size(); ++i) {
    w.begin_map();
    w.key("template_id"); w.write_int((int)i);
    w.key("templates"); w.begin_seq(false);
    for (const Template& t : tps[i]) {
      w.begin_map();
      w.key("width"); w.write_int(t.width);
      w.key("height"); w.write_int(t.height);
      w.key("pyramid_level"); w.write_int(t.pyramid_level);
      w.key("features"); w.begin_seq(false);
      for (const Feature& f : t.features) {
        w.begin_seq(true);
        w.write_int(f.x); w.write_int(f.y); w.write_int(f.label);
        w.end_seq();
      }
      w.end_seq();
      w.end_map();
    }
    w.end_seq();
    w.end_map();
  }
  w.end_seq();
}

}  // namespace

bool load_detector_yaml(const std::string& path, HostModel& model, std::string& err) {
  Node root;
  if (!lmyaml::parse_file(path, root, err)) return false;
  if (!read_header(root, model, err)) { err = path + ": " + err; return false; }
  const Node& classes = root["classes"];
  if (classes.kind == Node::SEQ)
    for (size_t i = 0; i < classes.size(); ++i)
      if (!read_class(classes.at(i), model, err)) { err = path + ": " + err; return false; }
  return true;
}

bool save_detector_yaml(const HostModel& model, const std::string& path, std::string& err) {
  lmyaml::Writer w;
  write_header(model, w);
  w.key("classes"); w.begin_seq(false);
  for (const auto& kv : model.classes) {
    w.begin_map();
    write_class(model, kv.first, kv.second, w);
    w.end_map();
  }
  w.end_seq();
  return w.save(path, err);
}

bool load_class_file(const std::string& path, HostModel& model, std::string& err) {
  Node root;
  if (!lmyaml::parse_file(path, root, err)) return false;
  if (!read_class(root, model, err)) { err = path + ": " + err; return false; }
  return true;
}

bool save_class_file(const HostModel& model, const std::string& class_id, const std::string& path, std::string& err) {
  auto it = model.classes.find(class_id);
  if (it == model.classes.end()) { err = "unknown class '" + class_id + "'"; return false; }
  lmyaml::Writer w;
  write_class(model, it->first, it->second, w);
  return w.save(path, err);
}

// ------------------------------------------------------------------------------------------------ binary cache
// the reference's service re-parses templates.yml on every request
// (/root/reference/src/linemod_ensenso_detect_3_mult_detect_service.cpp:1784,1851 -> readLinemod).  The cache holds the
// same model as flat little-endian arrays: a detector loads from it with a few large reads instead of a YAML parse.
//   header  : "LMB2CACH", u32 version (1), u32 levels, u32 modalities, u32 classes, u64 payload bytes, u64 FNV-1a of payload
//   payload : i32 T[levels]; lm_modality_desc mods[modalities];
//             per class: u32 name length, name bytes, u32 templates; per template pyramid, per (level, modality):
//             i32 width, height, pyramid_level, n_features, then n_features x (i16 x, i16 y, u8 label) packed in 5 bytes
namespace {
const char kCacheMagic[8] = {'L', 'M', 'B', '2', 'C', 'A', 'C', 'H'};
uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}
template <class T> void put(std::vector<uint8_t>& b, const T& v) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
  b.insert(b.end(), p, p + sizeof(T));
}
struct Cursor {
  const uint8_t* p; const uint8_t* end; bool ok = true;
  template <class T> T get() {
    T v = T();
    if ((size_t)(end - p) < sizeof(T)) { ok = false; return v; }
    std::memcpy(&v, p, sizeof(T)); p += sizeof(T);
    return v;
  }
};
}  // namespace

bool save_model_cache(const HostModel& model, const std::string& path, std::string& err) {
  std::vector<uint8_t> pay;
  for (int t : model.T) put<int32_t>(pay, t);
  for (const lm_modality_desc& m : model.mods) put(pay, m);
  for (const auto& kv : model.classes) {
    put<uint32_t>(pay, (uint32_t)kv.first.size());
    pay.insert(pay.end(), kv.first.begin(), kv.first.end());
    put<uint32_t>(pay, (uint32_t)kv.second.size());
    for (const TemplatePyramid& tp : kv.second) {
      if ((int)tp.size() != model.levels() * model.M()) { err = "class '" + kv.first + "' holds a template pyramid of the wrong size"; return false; }
      for (const Template& t : tp) {
        put<int32_t>(pay, t.width); put<int32_t>(pay, t.height); put<int32_t>(pay, t.pyramid_level);
        put<int32_t>(pay, (int32_t)t.features.size());
        for (const Feature& f : t.features) {
          if (f.x < -32768 || f.x > 32767 || f.y < -32768 || f.y > 32767 || f.label < 0 || f.label > 255) { err = "feature outside the cache's value range"; return false; }
          put<int16_t>(pay, (int16_t)f.x); put<int16_t>(pay, (int16_t)f.y); put<uint8_t>(pay, (uint8_t)f.label);
        }
      }
    }
  }
  std::vector<uint8_t> hdr;
  hdr.insert(hdr.end(), kCacheMagic, kCacheMagic + 8);
  put<uint32_t>(hdr, 1u); put<uint32_t>(hdr, (uint32_t)model.levels()); put<uint32_t>(hdr, (uint32_t)model.M());
  put<uint32_t>(hdr, (uint32_t)model.classes.size());
  put<uint64_t>(hdr, (uint64_t)pay.size()); put<uint64_t>(hdr, fnv1a(pay.data(), pay.size()));
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { err = "cannot open '" + path + "' for writing"; return false; }
  bool ok = std::fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size() && std::fwrite(pay.data(), 1, pay.size(), f) == pay.size();
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) err = "short write to '" + path + "'";
  return ok;
}

bool load_model_cache(const std::string& path, HostModel& model, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open '" + path + "'"; return false; }
  uint8_t hdr[40];
  if (std::fread(hdr, 1, sizeof(hdr), f) != sizeof(hdr) || std::memcmp(hdr, kCacheMagic, 8) != 0) {
    std::fclose(f); err = path + ": not a linemod_b200 template cache"; return false;
  }
  uint32_t version, levels, M, n_classes; uint64_t bytes, sum;
  std::memcpy(&version, hdr + 8, 4); std::memcpy(&levels, hdr + 12, 4); std::memcpy(&M, hdr + 16, 4);
  std::memcpy(&n_classes, hdr + 20, 4); std::memcpy(&bytes, hdr + 24, 8); std::memcpy(&sum, hdr + 32, 8);
  if (version != 1 || levels < 1 || levels > LM_MAX_LEVELS || M < 1 || M > LM_MAX_MODALITIES || bytes > (1ull << 34)) {
    std::fclose(f); err = path + ": unsupported cache header"; return false;
  }
  std::vector<uint8_t> pay((size_t)bytes);
  const bool read_ok = std::fread(pay.data(), 1, pay.size(), f) == pay.size();
  std::fclose(f);
  if (!read_ok || fnv1a(pay.data(), pay.size()) != sum) { err = path + ": truncated or corrupted cache (checksum mismatch)"; return false; }
  Cursor c{pay.data(), pay.data() + pay.size()};
  HostModel fresh;
  for (uint32_t l = 0; l < levels; ++l) fresh.T.push_back(c.get<int32_t>());
  for (uint32_t m = 0; m < M; ++m) fresh.mods.push_back(c.get<lm_modality_desc>());
  for (uint32_t k = 0; k < n_classes && c.ok; ++k) {
    const uint32_t len = c.get<uint32_t>();
    if (!c.ok || (size_t)(c.end - c.p) < len) { c.ok = false; break; }
    std::string id(reinterpret_cast<const char*>(c.p), len);
    c.p += len;
    const uint32_t n_templates = c.get<uint32_t>();
    // a pyramid takes at least 16 bytes per template header: bound the count by what the payload can still hold
    if (!c.ok || (uint64_t)n_templates * 16ull * levels * M > (uint64_t)(c.end - c.p)) { c.ok = false; break; }
    std::vector<TemplatePyramid>& tps = fresh.classes[id];
    tps.resize(n_templates);
    for (uint32_t t = 0; t < n_templates && c.ok; ++t) {
      TemplatePyramid& tp = tps[t];
      tp.resize((size_t)levels * M);
      for (Template& tm : tp) {
        tm.width = c.get<int32_t>(); tm.height = c.get<int32_t>(); tm.pyramid_level = c.get<int32_t>();
        const int32_t nf = c.get<int32_t>();
        if (!c.ok || nf < 0 || nf > LM_MAX_FEATURES || (size_t)(c.end - c.p) < (size_t)nf * 5) { c.ok = false; break; }
        tm.features.resize((size_t)nf);
        for (Feature& ft : tm.features) { ft.x = c.get<int16_t>(); ft.y = c.get<int16_t>(); ft.label = c.get<uint8_t>(); }
      }
    }
  }
  if (!c.ok || c.p != c.end) { err = path + ": malformed cache payload"; return false; }
  for (const auto& kv : fresh.classes)
    for (size_t t = 0; t < kv.second.size(); ++t)
      if (!validate_pyramid(fresh, kv.second[t], err)) { err = path + ": class '" + kv.first + "' template " + std::to_string(t) + ": " + err; return false; }
  for (int T : fresh.T) if (T < 1 || T > 16) { err = path + ": unsupported T"; return false; }
  for (const lm_modality_desc& md : fresh.mods)
    if (md.type != LM_COLOR_GRADIENT && md.type != LM_DEPTH_NORMAL) { err = path + ": unknown modality type"; return false; }
  fresh.version = model.version + 1;
  model = fresh;
  return true;
}

}  // namespace lm
