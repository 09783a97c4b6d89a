// linemod_b200.hpp -- header-only C++ facade over the C ABI (linemod_b200.h) with the method set of
// cv::linemod::Detector as the reference ROS package drives it:
//
//   cv::linemod::Detector(modalities, T_pyramid), ColorGradient(), DepthNormal()   /root/reference/src/renderer.cpp:179-185
//   detector->match(sources, threshold, matches, class_ids, noArray())             src/rgbdDetector.cpp:31-34
//   detector->addTemplate(sources, "obj", mask)                                    src/renderer.cpp:308
//   readLinemod(filename): read(fs.root()) + readClass per "classes" entry         src/rgbdDetector.cpp:1668-1680
//   writeLinemod(detector, filename): write(fs) + writeClass per class             src/renderer.cpp:56-70
//   detector->getTemplates(class_id, template_id)                                  src/linemod_ensenso_detect_3_mult_detect_service.cpp:351,741-744
//   detector->classIds(), numTemplates()                                           src/linemod_carmine_detect.cpp:319, src/renderer.cpp:61
//
// Types keep OpenCV's names and field order (Feature, Template, Match with the same operator< / operator==), images are
// borrowed views (linemod_b200::Image) so no OpenCV headers are needed; define LINEMOD_B200_WITH_OPENCV before
// including this file to get cv::Mat / cv::Rect overloads (see INTEGRATION.md for the three-line change in the
// reference).  Failed preconditions that OpenCV reports with CV_Assert -> cv::Exception surface here as
// linemod_b200::Exception carrying the LM_E_* code and the library's message.
#ifndef LINEMOD_B200_HPP_
#define LINEMOD_B200_HPP_

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "linemod_b200.h"

#ifdef LINEMOD_B200_WITH_OPENCV
#include <opencv2/core/core.hpp>
#endif

namespace linemod_b200 {

class Exception : public std::runtime_error {
 public:
  Exception(int code_, const std::string& what) : std::runtime_error(what), code(code_) {}
  int code;  // LM_E_*
};

namespace detail {
inline int check(int rc) {
  if (rc < 0) throw Exception(rc, lm_last_error());
  return rc;
}
}  // namespace detail

// cv::linemod::Feature
struct Feature {
  int x, y, label;
  Feature() : x(0), y(0), label(0) {}
  Feature(int x_, int y_, int label_) : x(x_), y(y_), label(label_) {}
};

// cv::linemod::Template
struct Template {
  int width, height, pyramid_level;
  std::vector<Feature> features;
  Template() : width(0), height(0), pyramid_level(0) {}
};

// cv::linemod::Match, including its ordering (similarity descending, then template_id ascending) and its equality
// (x, y, similarity, class_id -- template_id is NOT compared).
struct Match {
  int x, y;
  float similarity;
  std::string class_id;
  int template_id;
  Match() : x(0), y(0), similarity(0), template_id(0) {}
  Match(int x_, int y_, float s, const std::string& c, int t) : x(x_), y(y_), similarity(s), class_id(c), template_id(t) {}
  bool operator<(const Match& rhs) const {
    if (similarity != rhs.similarity) return similarity > rhs.similarity;
    return template_id < rhs.template_id;
  }
  bool operator==(const Match& rhs) const {
    return x == rhs.x && y == rhs.y && similarity == rhs.similarity && class_id == rhs.class_id;
  }
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};

// Borrowed view of host pixels (the role cv::Mat plays in the reference's calls).  step = bytes between rows, so a
// cropped ROI (mat_rgb(crop), ..._service.cpp:324-326) is passed without copying.
struct Image {
  const void* data;
  int rows, cols, type;  // LM_8UC3 (BGR) / LM_16UC1 (depth, mm) / LM_8UC1 (mask)
  size_t step;
  Image() : data(nullptr), rows(0), cols(0), type(LM_8UC1), step(0) {}
  Image(const void* d, int r, int c, int t, size_t s = 0) : data(d), rows(r), cols(c), type(t), step(s) {
    if (step == 0) step = (size_t)cols * (type == LM_8UC3 ? 3 : type == LM_16UC1 ? 2 : 1);
  }
  bool empty() const { return data == nullptr; }
#ifdef LINEMOD_B200_WITH_OPENCV
  Image(const cv::Mat& m) : data(m.data), rows(m.rows), cols(m.cols), step(m.step[0]) {  // NOLINT: implicit on purpose
    if (m.empty()) { data = nullptr; type = LM_8UC1; return; }
    if (m.type() == CV_8UC3) type = LM_8UC3;
    else if (m.type() == CV_16UC1) type = LM_16UC1;
    else if (m.type() == CV_8UC1) type = LM_8UC1;
    else throw Exception(LM_E_INVALID, "unsupported cv::Mat type (CV_8UC3, CV_16UC1 or CV_8UC1 expected)");
  }
#endif
  lm_image c() const {
    lm_image im;
    im.data = data; im.rows = rows; im.cols = cols; im.type = type; im.step = step;
    return im;
  }
};

// cv::linemod::QuantizedPyramid, the object Modality::process returns ([OCV] linemod.cpp ColorGradientPyramid /
// DepthNormalPyramid): quantize(dst), extractTemplate(templ), pyrDown().  All pyramid levels the image allows (up to
// LM_MAX_LEVELS) were quantised by the CUDA front end when process() ran; pyrDown() steps to the next one.
class QuantizedPyramid {
 public:
  explicit QuantizedPyramid(lm_qpyramid* q) : q_(q), level_(0) {}
  ~QuantizedPyramid() { lm_qpyramid_destroy(q_); }
  QuantizedPyramid(const QuantizedPyramid&) = delete;
  QuantizedPyramid& operator=(const QuantizedPyramid&) = delete;

  // size of the current level (what quantize() writes)
  void size(int& rows, int& cols) const { detail::check(lm_qpyramid_size(q_, level_, &rows, &cols)); }
  // quantize into caller-owned CV_8UC1 memory of size()
  void quantize(uint8_t* dst, size_t step = 0) const {
    int r = 0, c = 0;
    size(r, c);
    lm_image im;
    im.data = dst; im.rows = r; im.cols = c; im.type = LM_8UC1; im.step = step ? step : (size_t)c;
    detail::check(lm_qpyramid_quantize(q_, level_, &im));
  }
  void quantize(std::vector<uint8_t>& dst) const {
    int r = 0, c = 0;
    size(r, c);
    dst.assign((size_t)r * c, 0);
    quantize(dst.data());
  }
#ifdef LINEMOD_B200_WITH_OPENCV
  void quantize(cv::Mat& dst) const {
    int r = 0, c = 0;
    size(r, c);
    dst.create(r, c, CV_8UC1);
    quantize(dst.data, dst.step[0]);
  }
#endif
  bool extractTemplate(Template& templ) const {
    lm_template_hdr hdr;
    int32_t f[3 * LM_MAX_FEATURES];
    const int ok = detail::check(lm_qpyramid_extract(q_, level_, &hdr, f));
    if (!ok) return false;
    templ.width = hdr.width; templ.height = hdr.height; templ.pyramid_level = hdr.pyramid_level;
    templ.features.resize((size_t)hdr.num_features);
    for (int j = 0; j < hdr.num_features; ++j) templ.features[(size_t)j] = Feature(f[3 * j], f[3 * j + 1], f[3 * j + 2]);
    return true;
  }
  void pyrDown() {
    if (level_ + 1 >= lm_qpyramid_levels(q_)) throw Exception(LM_E_INVALID, "QuantizedPyramid: no further level (image too small or LM_MAX_LEVELS reached)");
    ++level_;
  }

 private:
  lm_qpyramid* q_;
  int level_;
};

// cv::linemod::Modality and its two implementations: parameter carriers; process() runs the CUDA front end.
class Modality {
 public:
  virtual ~Modality() {}
  virtual std::string name() const = 0;
  virtual lm_modality_desc desc() const = 0;
  // Modality::process(src, mask): src CV_8UC3 (ColorGradient) / CV_16UC1 (DepthNormal), mask CV_8UC1 or empty
#ifdef LINEMOD_B200_WITH_OPENCV
  typedef cv::Ptr<QuantizedPyramid> PyramidPtr;
#else
  typedef std::shared_ptr<QuantizedPyramid> PyramidPtr;
#endif
  PyramidPtr process(const Image& src, const Image& mask = Image()) const {
    const lm_modality_desc d = desc();
    const lm_image s = src.c(), m = mask.c();
    int levels = LM_MAX_LEVELS;
    while (levels > 1 && ((src.rows >> (levels - 1)) < 16 || (src.cols >> (levels - 1)) < 16)) --levels;
    lm_qpyramid* q = nullptr;
    detail::check(lm_modality_process(&d, &s, mask.empty() ? nullptr : &m, levels, nullptr, &q));
    return PyramidPtr(new QuantizedPyramid(q));
  }
  static std::shared_ptr<Modality> create(const std::string& modality_type);
#ifdef LINEMOD_B200_WITH_OPENCV
  // Modality::read / write / create(FileNode): { type: "ColorGradient", weak_threshold, num_features, strong_threshold } or
  // { type: "DepthNormal", distance_threshold, difference_threshold, num_features, extract_threshold }
  virtual void read(const cv::FileNode& fn) = 0;
  virtual void write(cv::FileStorage& fs) const = 0;
  static cv::Ptr<Modality> create(const cv::FileNode& fn);
#endif
};

class ColorGradient : public Modality {
 public:
  ColorGradient() : weak_threshold(10.0f), num_features(63), strong_threshold(55.0f) {}
  ColorGradient(float weak, size_t nf, float strong) : weak_threshold(weak), num_features(nf), strong_threshold(strong) {}
  std::string name() const override { return "ColorGradient"; }
  lm_modality_desc desc() const override {
    lm_modality_desc d = {LM_COLOR_GRADIENT, weak_threshold, strong_threshold, 2000, 50, 2, (int32_t)num_features};
    return d;
  }
#ifdef LINEMOD_B200_WITH_OPENCV
  void read(const cv::FileNode& fn) override {
    if ((std::string)fn["type"] != name()) throw Exception(LM_E_IO, "modality node is not a ColorGradient");
    weak_threshold = (float)fn["weak_threshold"];
    num_features = (size_t)(int)fn["num_features"];
    strong_threshold = (float)fn["strong_threshold"];
  }
  void write(cv::FileStorage& fs) const override {
    fs << "type" << name() << "weak_threshold" << weak_threshold << "num_features" << (int)num_features << "strong_threshold" << strong_threshold;
  }
#endif
  float weak_threshold;
  size_t num_features;
  float strong_threshold;
};

class DepthNormal : public Modality {
 public:
  DepthNormal() : distance_threshold(2000), difference_threshold(50), num_features(63), extract_threshold(2) {}
  DepthNormal(int distance, int difference, size_t nf, int extract)
      : distance_threshold(distance), difference_threshold(difference), num_features(nf), extract_threshold(extract) {}
  std::string name() const override { return "DepthNormal"; }
  lm_modality_desc desc() const override {
    lm_modality_desc d = {LM_DEPTH_NORMAL, 10.0f, 55.0f, distance_threshold, difference_threshold, extract_threshold,
                          (int32_t)num_features};
    return d;
  }
#ifdef LINEMOD_B200_WITH_OPENCV
  void read(const cv::FileNode& fn) override {
    if ((std::string)fn["type"] != name()) throw Exception(LM_E_IO, "modality node is not a DepthNormal");
    distance_threshold = (int)fn["distance_threshold"];
    difference_threshold = (int)fn["difference_threshold"];
    num_features = (size_t)(int)fn["num_features"];
    extract_threshold = (int)fn["extract_threshold"];
  }
  void write(cv::FileStorage& fs) const override {
    fs << "type" << name() << "distance_threshold" << distance_threshold << "difference_threshold" << difference_threshold
       << "num_features" << (int)num_features << "extract_threshold" << extract_threshold;
  }
#endif
  int distance_threshold, difference_threshold;
  size_t num_features;
  int extract_threshold;
};

inline std::shared_ptr<Modality> Detector_from_desc_impl(const lm_modality_desc& d) {
  if (d.type == LM_COLOR_GRADIENT)
    return std::make_shared<ColorGradient>(d.weak_threshold, (size_t)d.num_features, d.strong_threshold);
  return std::make_shared<DepthNormal>(d.distance_threshold, d.difference_threshold, (size_t)d.num_features, d.extract_threshold);
}

inline std::shared_ptr<Modality> Modality::create(const std::string& modality_type) {
  if (modality_type == "ColorGradient") return std::make_shared<ColorGradient>();
  if (modality_type == "DepthNormal") return std::make_shared<DepthNormal>();
  throw Exception(LM_E_INVALID, "unknown modality '" + modality_type + "'");
}

#ifdef LINEMOD_B200_WITH_OPENCV
inline cv::Ptr<Modality> Modality::create(const cv::FileNode& fn) {
  const std::string type = (std::string)fn["type"];
  cv::Ptr<Modality> m;
  if (type == "ColorGradient") m = cv::Ptr<Modality>(new ColorGradient());
  else if (type == "DepthNormal") m = cv::Ptr<Modality>(new DepthNormal());
  else throw Exception(LM_E_INVALID, "unknown modality '" + type + "'");
  m->read(fn);
  return m;
}
#endif

// cv::linemod::Detector
class Detector {
 public:
  typedef std::vector<Template> TemplatePyramid;

  // Empty detector, to be filled by read() exactly like `new cv::linemod::Detector` + read(fs.root()).
  Detector() : h_(nullptr) {}
  Detector(const std::vector<std::shared_ptr<Modality> >& modalities, const std::vector<int>& T_pyramid)
      : h_(nullptr), modalities_(modalities) {
    std::vector<lm_modality_desc> mods;
    for (size_t i = 0; i < modalities.size(); ++i) mods.push_back(modalities[i]->desc());
    std::vector<int32_t> T(T_pyramid.begin(), T_pyramid.end());
    detail::check(lm_create(T.data(), (int)T.size(), mods.data(), (int)mods.size(), &h_));
  }
  ~Detector() { lm_destroy(h_); }
  Detector(const Detector&) = delete;
  Detector& operator=(const Detector&) = delete;

  // Detector::match.  quantized_images (nullable) receives pyramidLevels()*modalities images, index l*M+m, each
  // (rows>>l) x (cols>>l) bytes.  masks: empty or one LM_8UC1 image per modality.
  void match(const std::vector<Image>& sources, float threshold, std::vector<Match>& matches,
             const std::vector<std::string>& class_ids = std::vector<std::string>(),
             std::vector<std::vector<uint8_t> >* quantized_images = nullptr,
             const std::vector<Image>& masks = std::vector<Image>()) const {
    need_handle();
    std::vector<lm_image> src, msk;
    for (size_t i = 0; i < sources.size(); ++i) src.push_back(sources[i].c());
    for (size_t i = 0; i < masks.size(); ++i) msk.push_back(masks[i].c());
    std::vector<const char*> ids;
    for (size_t i = 0; i < class_ids.size(); ++i) ids.push_back(class_ids[i].c_str());
    std::vector<lm_image_out> qout;
    if (quantized_images && !sources.empty()) {
      const int L = pyramidLevels(), M = lm_num_modalities(h_);
      quantized_images->assign((size_t)L * M, std::vector<uint8_t>());
      for (int l = 0; l < L; ++l)
        for (int m = 0; m < M; ++m) {
          std::vector<uint8_t>& q = (*quantized_images)[(size_t)l * M + m];
          const int r = sources[0].rows >> l, c = sources[0].cols >> l;
          q.assign((size_t)r * c, 0);
          lm_image_out o;
          o.data = q.data(); o.rows = r; o.cols = c; o.type = LM_8UC1; o.step = (size_t)c;
          qout.push_back(o);
        }
    }
    lm_match_rec* recs = nullptr;
    size_t n = 0;
    detail::check(lm_match(h_, src.data(), (int)src.size(), threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size(),
                           msk.empty() ? nullptr : msk.data(), (int)msk.size(), qout.empty() ? nullptr : qout.data(),
                           &recs, &n));
    matches.clear();
    matches.reserve(n);
    for (size_t i = 0; i < n; ++i)
      matches.push_back(Match(recs[i].x, recs[i].y, recs[i].similarity, lm_class_id(h_, recs[i].class_index),
                              recs[i].template_id));
    lm_free_matches(recs);
  }

  // Detector::addTemplate: template_id, or -1 when some pyramid level lacks candidate features.
  int addTemplate(const std::vector<Image>& sources, const std::string& class_id, const Image& object_mask,
                  Rect* bounding_box = nullptr) {
    need_handle();
    std::vector<lm_image> src;
    for (size_t i = 0; i < sources.size(); ++i) src.push_back(sources[i].c());
    lm_image mask = object_mask.c();
    lm_rect bb = {0, 0, 0, 0};
    int id = lm_add_template(h_, src.data(), (int)src.size(), class_id.c_str(), object_mask.empty() ? nullptr : &mask, &bb);
    if (id < -1) throw Exception(id + 100, lm_last_error());
    if (bounding_box) { bounding_box->x = bb.x; bounding_box->y = bb.y; bounding_box->width = bb.width; bounding_box->height = bb.height; }
    cache_.clear();
    return id;
  }

  // Detector::addSyntheticTemplate
  int addSyntheticTemplate(const std::vector<Template>& templates, const std::string& class_id) {
    need_handle();
    std::vector<lm_template_hdr> hdr(templates.size());
    std::vector<int32_t> feats;
    for (size_t i = 0; i < templates.size(); ++i) {
      hdr[i].width = templates[i].width; hdr[i].height = templates[i].height;
      hdr[i].pyramid_level = templates[i].pyramid_level; hdr[i].num_features = (int32_t)templates[i].features.size();
      for (size_t j = 0; j < templates[i].features.size(); ++j) {
        feats.push_back(templates[i].features[j].x); feats.push_back(templates[i].features[j].y);
        feats.push_back(templates[i].features[j].label);
      }
    }
    if (feats.empty()) feats.push_back(0);
    cache_.clear();
    return detail::check(lm_add_synthetic_template(h_, class_id.c_str(), (int)templates.size(), hdr.data(), feats.data()));
  }

  // Detector::getTemplates: the pyramid (index l*M+m) of one template; the reference stays valid until the next
  // addTemplate / read on this detector.
  const std::vector<Template>& getTemplates(const std::string& class_id, int template_id) const {
    need_handle();
    std::pair<std::string, int> key(class_id, template_id);
    std::map<std::pair<std::string, int>, TemplatePyramid>::const_iterator it = cache_.find(key);
    if (it != cache_.end()) return it->second;
    const int n = pyramidLevels() * lm_num_modalities(h_);
    std::vector<lm_template_hdr> hdr((size_t)n);
    int total = detail::check(lm_get_templates(h_, class_id.c_str(), template_id, hdr.data(), nullptr));
    std::vector<int32_t> feats((size_t)total * 3 + 3);
    detail::check(lm_get_templates(h_, class_id.c_str(), template_id, hdr.data(), feats.data()));
    TemplatePyramid tp((size_t)n);
    size_t k = 0;
    for (int i = 0; i < n; ++i) {
      tp[i].width = hdr[i].width; tp[i].height = hdr[i].height; tp[i].pyramid_level = hdr[i].pyramid_level;
      for (int j = 0; j < hdr[i].num_features; ++j, ++k)
        tp[i].features.push_back(Feature(feats[3 * k], feats[3 * k + 1], feats[3 * k + 2]));
    }
    return cache_[key] = tp;
  }

  int numTemplates() const { return h_ ? lm_num_templates(h_, nullptr) : 0; }
  int numTemplates(const std::string& class_id) const { return h_ ? lm_num_templates(h_, class_id.c_str()) : 0; }
  int numClasses() const { return h_ ? lm_num_classes(h_) : 0; }
  std::vector<std::string> classIds() const {
    std::vector<std::string> ids;
    for (int i = 0; i < numClasses(); ++i) ids.push_back(lm_class_id(h_, i));
    return ids;
  }
  int pyramidLevels() const { return h_ ? lm_pyramid_levels(h_) : 0; }
  int getT(int pyramid_level) const { need_handle(); return detail::check(lm_get_T(h_, pyramid_level)); }
  const std::vector<std::shared_ptr<Modality> >& getModalities() const { return modalities_; }

  // readLinemod(filename): Detector::read(fs.root()) followed by readClass for every entry of "classes".
  void read(const std::string& filename) {
    lm_detector* fresh = nullptr;
    detail::check(lm_create_from_yaml(filename.c_str(), &fresh));
    lm_destroy(h_);
    h_ = fresh;
    cache_.clear();
    modalities_.clear();
    for (int m = 0; m < lm_num_modalities(h_); ++m) {
      lm_modality_desc d;
      detail::check(lm_get_modality(h_, m, &d));
      modalities_.push_back(from_desc(d));
    }
  }
  // writeLinemod(detector, filename): Detector::write(fs) + "classes" [ { writeClass } ... ].
  void write(const std::string& filename) const { need_handle(); detail::check(lm_write_yaml(h_, filename.c_str())); }
  // Detector::readClasses / writeClasses (one FileStorage file per class, gzip when the name ends in .gz).
  void readClasses(const std::vector<std::string>& class_ids, const std::string& format = "templates_%s.yml.gz") {
    need_handle();
    std::vector<const char*> ids;
    for (size_t i = 0; i < class_ids.size(); ++i) ids.push_back(class_ids[i].c_str());
    detail::check(lm_read_classes(h_, ids.data(), (int)ids.size(), format.c_str()));
    cache_.clear();
  }
  void writeClasses(const std::string& format = "templates_%s.yml.gz") const {
    need_handle();
    detail::check(lm_write_classes(h_, format.c_str()));
  }

#ifdef LINEMOD_B200_WITH_OPENCV
  // ---- the OpenCV spellings the reference uses, so that its call sites compile unchanged against this class:
  //   cv::Ptr<cv::linemod::Detector> detector_(new cv::linemod::Detector(modalities, T))      src/renderer.cpp:179-185
  //   detector->read(fs.root()); detector->readClass(*i)  (readLinemod)                         src/rgbdDetector.cpp:1668-1680
  //   detector->write(fs); detector->writeClass(ids[i], fs)  (writeLinemod)                     src/renderer.cpp:56-70
  //   linemod_detector->match(sources, threshold, matches, std::vector<String>(), noArray())   src/rgbdDetector.cpp:33
  Detector(const std::vector<cv::Ptr<Modality> >& modalities, const std::vector<int>& T_pyramid) : h_(nullptr) {
    std::vector<lm_modality_desc> mods;
    for (size_t i = 0; i < modalities.size(); ++i) {
      mods.push_back(modalities[i]->desc());
      modalities_.push_back(from_desc(mods.back()));
    }
    std::vector<int32_t> T(T_pyramid.begin(), T_pyramid.end());
    detail::check(lm_create(T.data(), (int)T.size(), mods.data(), (int)mods.size(), &h_));
  }

  // [OCV] Detector::read: "pyramid_levels", "T", "modalities" [ { type, parameters } ... ] of a templates.yml root node.
  // Replaces the detector's configuration and drops its classes, like upstream.
  void read(const cv::FileNode& fn) {
    const int levels = (int)fn["pyramid_levels"];
    std::vector<int> T;
    fn["T"] >> T;
    if (levels < 1 || (int)T.size() != levels) throw Exception(LM_E_IO, "templates file: pyramid_levels / T mismatch");
    std::vector<lm_modality_desc> mods;
    const cv::FileNode mn = fn["modalities"];
    for (cv::FileNodeIterator it = mn.begin(), e = mn.end(); it != e; ++it) {
      const cv::FileNode m = *it;
      const std::string type = (std::string)m["type"];
      lm_modality_desc d = Modality::create(type)->desc();
      if (d.type == LM_COLOR_GRADIENT) {
        if (!m["weak_threshold"].empty()) d.weak_threshold = (float)m["weak_threshold"];
        if (!m["strong_threshold"].empty()) d.strong_threshold = (float)m["strong_threshold"];
      } else {
        if (!m["distance_threshold"].empty()) d.distance_threshold = (int)m["distance_threshold"];
        if (!m["difference_threshold"].empty()) d.difference_threshold = (int)m["difference_threshold"];
        if (!m["extract_threshold"].empty()) d.extract_threshold = (int)m["extract_threshold"];
      }
      if (!m["num_features"].empty()) d.num_features = (int)m["num_features"];
      mods.push_back(d);
    }
    std::vector<int32_t> T32(T.begin(), T.end());
    lm_detector* fresh = nullptr;
    detail::check(lm_create(T32.data(), levels, mods.data(), (int)mods.size(), &fresh));
    lm_destroy(h_);
    h_ = fresh;
    cache_.clear();
    modalities_.clear();
    for (size_t i = 0; i < mods.size(); ++i) modalities_.push_back(from_desc(mods[i]));
  }

  // [OCV] Detector::readClass: one entry of "classes".  The CV_Asserts of upstream (modalities and pyramid_levels match the
  // detector, the class is new, template ids are consecutive) become Exceptions.  Returns the class id.
  std::string readClass(const cv::FileNode& fn, const std::string& class_id_override = "") {
    need_handle();
    const cv::FileNode mn = fn["modalities"];
    if (mn.size() != modalities_.size()) throw Exception(LM_E_IO, "class modalities do not match the detector");
    size_t k = 0;
    for (cv::FileNodeIterator it = mn.begin(), e = mn.end(); it != e; ++it, ++k)
      if ((std::string)(*it) != modalities_[k]->name()) throw Exception(LM_E_IO, "class modality '" + (std::string)(*it) + "' does not match the detector");
    if ((int)fn["pyramid_levels"] != pyramidLevels()) throw Exception(LM_E_IO, "class pyramid_levels does not match the detector");
    const std::string class_id = class_id_override.empty() ? (std::string)fn["class_id"] : class_id_override;
    if (numTemplates(class_id) != 0) throw Exception(LM_E_IO, "detector already has class '" + class_id + "'");
    const cv::FileNode tps = fn["template_pyramids"];
    int expected_id = 0;
    for (cv::FileNodeIterator it = tps.begin(), e = tps.end(); it != e; ++it, ++expected_id) {
      const cv::FileNode tpn = *it;
      if ((int)tpn["template_id"] != expected_id) throw Exception(LM_E_IO, "template_id out of sequence");
      std::vector<Template> tp;
      const cv::FileNode tn = tpn["templates"];
      for (cv::FileNodeIterator jt = tn.begin(), je = tn.end(); jt != je; ++jt) {
        const cv::FileNode t = *jt;
        Template tm;
        tm.width = (int)t["width"]; tm.height = (int)t["height"]; tm.pyramid_level = (int)t["pyramid_level"];
        const cv::FileNode feats = t["features"];
        for (cv::FileNodeIterator ft = feats.begin(), fe = feats.end(); ft != fe; ++ft) {
          const cv::FileNode f = *ft;
          tm.features.push_back(Feature((int)f[0], (int)f[1], (int)f[2]));
        }
        tp.push_back(tm);
      }
      if (addSyntheticTemplate(tp, class_id) != expected_id) throw Exception(LM_E_IO, "template_id out of sequence");
    }
    return class_id;
  }

  // [OCV] Detector::write: the detector's configuration into an open FileStorage (same keys as lm_write_yaml emits).
  void write(cv::FileStorage& fs) const {
    need_handle();
    fs << "pyramid_levels" << pyramidLevels();
    std::vector<int> T;
    for (int l = 0; l < pyramidLevels(); ++l) T.push_back(getT(l));
    fs << "T" << T;
    fs << "modalities" << "[";
    for (int m = 0; m < lm_num_modalities(h_); ++m) {
      lm_modality_desc d;
      detail::check(lm_get_modality(h_, m, &d));
      fs << "{";
      if (d.type == LM_COLOR_GRADIENT) {
        fs << "type" << "ColorGradient" << "weak_threshold" << d.weak_threshold << "num_features" << (int)d.num_features
           << "strong_threshold" << d.strong_threshold;
      } else {
        fs << "type" << "DepthNormal" << "distance_threshold" << (int)d.distance_threshold << "difference_threshold"
           << (int)d.difference_threshold << "num_features" << (int)d.num_features << "extract_threshold" << (int)d.extract_threshold;
      }
      fs << "}";
    }
    fs << "]";
  }

  // [OCV] Detector::writeClass: class_id, modalities, pyramid_levels, template_pyramids [ { template_id, templates [ ... ] } ]
  void writeClass(const std::string& class_id, cv::FileStorage& fs) const {
    need_handle();
    if (numTemplates(class_id) == 0 && numClasses() == 0) throw Exception(LM_E_NOTFOUND, "unknown class '" + class_id + "'");
    fs << "class_id" << class_id;
    fs << "modalities" << "[:";
    for (size_t i = 0; i < modalities_.size(); ++i) fs << modalities_[i]->name();
    fs << "]";
    fs << "pyramid_levels" << pyramidLevels();
    fs << "template_pyramids" << "[";
    const int n = numTemplates(class_id);
    for (int i = 0; i < n; ++i) {
      const std::vector<Template>& tp = getTemplates(class_id, i);
      fs << "{";
      fs << "template_id" << i;
      fs << "templates" << "[";
      for (size_t j = 0; j < tp.size(); ++j) {
        fs << "{";
        fs << "width" << tp[j].width << "height" << tp[j].height << "pyramid_level" << tp[j].pyramid_level;
        fs << "features" << "[";
        for (size_t k = 0; k < tp[j].features.size(); ++k)
          fs << "[:" << tp[j].features[k].x << tp[j].features[k].y << tp[j].features[k].label << "]";
        fs << "]";
        fs << "}";
      }
      fs << "]";
      fs << "}";
    }
    fs << "]";
  }

  // Detector::match with OpenCV's own parameter list (class_ids given, quantized_images as an OutputArrayOfArrays)
  void match(const std::vector<cv::Mat>& sources, float threshold, std::vector<Match>& matches,
             const std::vector<std::string>& class_ids, cv::OutputArrayOfArrays quantized_images,
             const std::vector<cv::Mat>& masks = std::vector<cv::Mat>()) const {
    std::vector<Image> src(sources.begin(), sources.end()), msk(masks.begin(), masks.end());
    std::vector<std::vector<uint8_t> > q;
    match(src, threshold, matches, class_ids, quantized_images.needed() ? &q : nullptr, msk);
    if (quantized_images.needed()) {
      const int M = lm_num_modalities(h_);
      quantized_images.create(1, (int)q.size(), CV_8U);
      for (size_t i = 0; i < q.size(); ++i) {
        const int l = (int)i / M, r = sources[0].rows >> l, c = sources[0].cols >> l;
        quantized_images.create(r, c, CV_8UC1, (int)i);
        cv::Mat dst = quantized_images.getMat((int)i);
        for (int y = 0; y < r; ++y) std::memcpy(dst.data + (size_t)y * dst.step[0], q[i].data() + (size_t)y * c, (size_t)c);
      }
    }
  }

  // cv::Mat spellings of the two calls the reference makes with images.
  void match(const std::vector<cv::Mat>& sources, float threshold, std::vector<Match>& matches,
             const std::vector<std::string>& class_ids = std::vector<std::string>(),
             std::vector<cv::Mat>* quantized_images = nullptr,
             const std::vector<cv::Mat>& masks = std::vector<cv::Mat>()) const {
    std::vector<Image> src(sources.begin(), sources.end()), msk(masks.begin(), masks.end());
    std::vector<std::vector<uint8_t> > q;
    match(src, threshold, matches, class_ids, quantized_images ? &q : nullptr, msk);
    if (quantized_images) {
      quantized_images->clear();
      const int M = lm_num_modalities(h_);
      for (size_t i = 0; i < q.size(); ++i) {
        const int l = (int)i / M;
        quantized_images->push_back(cv::Mat(sources[0].rows >> l, sources[0].cols >> l, CV_8UC1, q[i].data()).clone());
      }
    }
  }
  int addTemplate(const std::vector<cv::Mat>& sources, const std::string& class_id, const cv::Mat& object_mask,
                  cv::Rect* bounding_box = nullptr) {
    std::vector<Image> src(sources.begin(), sources.end());
    Rect bb;
    int id = addTemplate(src, class_id, Image(object_mask), &bb);
    if (bounding_box) *bounding_box = cv::Rect(bb.x, bb.y, bb.width, bb.height);
    return id;
  }
#endif

  // The trainer's loop (src/renderer.cpp:239-329): render every view (T, up: camera position and up vector per view,
  // see ViewSphere) of `mesh` and addTemplate it, all on the GPU.  Returns the template ids (-1 where a view failed).
  std::vector<int> trainViews(const class Mesh& mesh, const lm_camera& camera, const std::vector<double>& T,
                              const std::vector<double>& up, const std::string& class_id,
                              std::vector<Rect>* mask_rects = nullptr, std::vector<uint16_t>* centre_depth_mm = nullptr);

  lm_detector* handle() const { return h_; }  // for the batch / multi-query / multi-GPU entry points of the C ABI

 private:
  static std::shared_ptr<Modality> from_desc(const lm_modality_desc& d);
  void need_handle() const {
    if (!h_) throw Exception(LM_E_STATE, "empty Detector: construct it with modalities or read() a templates.yml first");
  }
  lm_detector* h_;
  std::vector<std::shared_ptr<Modality> > modalities_;
  mutable std::map<std::pair<std::string, int>, TemplatePyramid> cache_;
};

inline std::shared_ptr<Modality> Detector::from_desc(const lm_modality_desc& d) { return Detector_from_desc_impl(d); }

// ------------------------------------------------------------------------------------------------ several GPUs, one caller
// lm_group: `prototype`'s model cloned onto `devices`, one worker thread per device.  Frames: every device holds all
// templates and takes its share of a batch's frames over its own PCIe link; Templates: the template set is sharded and every
// device sees every frame (the north-star layout).  Both return exactly what the prototype would.
class DetectorGroup {
 public:
  enum Mode { Frames = LM_GROUP_FRAMES, Templates = LM_GROUP_TEMPLATES };
  DetectorGroup(const Detector& prototype, const std::vector<int>& devices, Mode mode = Frames) : g_(nullptr), proto_(prototype.handle()) {
    detail::check(lm_group_create(prototype.handle(), devices.data(), (int)devices.size(), (int)mode, &g_));
  }
  // the 2-D grid: devices.size() / template_shards sets of `template_shards` template shards, frames dealt out to the sets
  DetectorGroup(const Detector& prototype, const std::vector<int>& devices, int template_shards) : g_(nullptr), proto_(prototype.handle()) {
    detail::check(lm_group_create_grid(prototype.handle(), devices.data(), (int)devices.size(), template_shards, &g_));
  }
  ~DetectorGroup() { lm_group_destroy(g_); }
  DetectorGroup(const DetectorGroup&) = delete;
  DetectorGroup& operator=(const DetectorGroup&) = delete;
  int size() const { return lm_group_size(g_); }
  void setOption(const std::string& key, int value) { detail::check(lm_group_set_option(g_, key.c_str(), value)); }
  // Detector::match on a batch of frames (frames[f] = the sources of frame f): matches[f] as Detector::match returns them.
  void matchBatch(const std::vector<std::vector<Image> >& frames, float threshold, std::vector<std::vector<Match> >& matches,
                  const std::vector<std::string>& class_ids = std::vector<std::string>()) const {
    std::vector<lm_image> src;
    for (size_t f = 0; f < frames.size(); ++f)
      for (size_t m = 0; m < frames[f].size(); ++m) src.push_back(frames[f][m].c());
    std::vector<const char*> ids;
    for (size_t i = 0; i < class_ids.size(); ++i) ids.push_back(class_ids[i].c_str());
    lm_query q = {threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size()};
    std::vector<size_t> offs(frames.size() + 1, 0);
    lm_match_rec* recs = nullptr;
    lm_image none = {nullptr, 0, 0, 0, 0};
    detail::check(lm_group_match_batch_multi(g_, src.empty() ? &none : src.data(), (int)frames.size(),
                                             frames.empty() ? 0 : (int)frames[0].size(), &q, 1, &recs, offs.data()));
    matches.assign(frames.size(), std::vector<Match>());
    for (size_t f = 0; f < frames.size(); ++f)
      for (size_t i = offs[f]; i < offs[f + 1]; ++i)
        matches[f].push_back(Match(recs[i].x, recs[i].y, recs[i].similarity, lm_class_id(proto_, recs[i].class_index), recs[i].template_id));
    lm_free_matches(recs);
  }
  void match(const std::vector<Image>& sources, float threshold, std::vector<Match>& matches,
             const std::vector<std::string>& class_ids = std::vector<std::string>()) const {
    std::vector<std::vector<Match> > out;
    matchBatch(std::vector<std::vector<Image> >(1, sources), threshold, out, class_ids);
    matches.swap(out[0]);
  }
  lm_group* handle() const { return g_; }

 private:
  lm_group* g_;
  const lm_detector* proto_;
};

// lm_stream: Detector::match over a continuous stream of host frames.  push() enqueues frames (their buffers must stay
// unchanged until pop() has returned them), pop() hands out finished frames in push order, each as Detector::match would
// return it.  While a FrameStream lives, its detector refuses other matching calls.
class FrameStream {
 public:
  FrameStream(const Detector& det, float threshold, const std::vector<std::string>& class_ids = std::vector<std::string>())
      : s_(nullptr), det_(det.handle()) {
    std::vector<const char*> ids;
    for (size_t i = 0; i < class_ids.size(); ++i) ids.push_back(class_ids[i].c_str());
    lm_query q = {threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size()};
    detail::check(lm_stream_open(det.handle(), &q, 1, &s_));
  }
  ~FrameStream() { lm_stream_close(s_); }
  FrameStream(const FrameStream&) = delete;
  FrameStream& operator=(const FrameStream&) = delete;
  void push(const std::vector<std::vector<Image> >& frames) {
    std::vector<lm_image> src;
    for (size_t f = 0; f < frames.size(); ++f)
      for (size_t m = 0; m < frames[f].size(); ++m) src.push_back(frames[f][m].c());
    if (!frames.empty()) detail::check(lm_stream_push(s_, src.data(), (int)frames.size(), (int)frames[0].size()));
  }
  int inFlight() const { return lm_stream_in_flight(s_); }
  // finished frames, oldest first; wait_all: everything pushed so far
  void pop(std::vector<std::vector<Match> >& matches, bool wait_all = false) {
    const int cap = std::max(1, inFlight());
    std::vector<size_t> offs((size_t)cap + 1, 0);
    lm_match_rec* recs = nullptr;
    int n = 0;
    detail::check(lm_stream_pop(s_, wait_all ? 1 : 0, cap, &recs, offs.data(), &n));
    matches.assign((size_t)n, std::vector<Match>());
    for (int f = 0; f < n; ++f)
      for (size_t i = offs[(size_t)f]; i < offs[(size_t)f + 1]; ++i)
        matches[(size_t)f].push_back(Match(recs[i].x, recs[i].y, recs[i].similarity, lm_class_id(det_, recs[i].class_index), recs[i].template_id));
    lm_free_matches(recs);
  }

 private:
  lm_stream* s_;
  const lm_detector* det_;
};

// ------------------------------------------------------------------------------------------------ training helpers
// Renderer3d(stl_file) of the reference's trainer (src/renderer.cpp:239): a triangle mesh in the object frame, metres.
class Mesh {
 public:
  explicit Mesh(const std::string& stl_file) : m_(nullptr) { detail::check(lm_mesh_load_stl(stl_file.c_str(), &m_)); }
  Mesh(const float* triangles, int n_triangles) : m_(nullptr) { detail::check(lm_mesh_create(triangles, n_triangles, &m_)); }
  ~Mesh() { lm_mesh_destroy(m_); }
  int numTriangles() const { return lm_mesh_num_triangles(m_); }
  lm_mesh* handle() const { return m_; }

 private:
  Mesh(const Mesh&);
  Mesh& operator=(const Mesh&);
  lm_mesh* m_;
};

// RendererIterator (src/renderer.cpp:242-246): n_points on a sphere x in-plane angles x radii, in the reference's order.
class ViewSphere {
 public:
  ViewSphere(int n_points, int angle_step, float radius_min, float radius_max, float radius_step, int angle_min = -80,
             int angle_max = 80) {
    vs_.n_points = n_points; vs_.angle_min = angle_min; vs_.angle_max = angle_max; vs_.angle_step = angle_step;
    vs_.radius_min = radius_min; vs_.radius_max = radius_max; vs_.radius_step = radius_step;
  }
  int size() const { return detail::check(lm_view_count(&vs_)); }  // n_templates()
  // camera position T and up vector of view `index`; D_obj = radius
  void view(int index, double T[3], double up[3], float* radius = nullptr) const {
    detail::check(lm_view_params(&vs_, index, T, up, radius, nullptr, nullptr));
  }
  // all views, flattened (3 doubles per view)
  void views(std::vector<double>& T, std::vector<double>& up, std::vector<float>* radii = nullptr) const {
    const int n = size();
    T.assign((size_t)n * 3, 0.0); up.assign((size_t)n * 3, 0.0);
    if (radii) radii->assign((size_t)n, 0.f);
    for (int i = 0; i < n; ++i) view(i, &T[3 * (size_t)i], &up[3 * (size_t)i], radii ? &(*radii)[(size_t)i] : nullptr);
  }
  const lm_view_sphere& c() const { return vs_; }

 private:
  lm_view_sphere vs_;
};

inline std::vector<int> Detector::trainViews(const Mesh& mesh, const lm_camera& camera, const std::vector<double>& T,
                                             const std::vector<double>& up, const std::string& class_id,
                                             std::vector<Rect>* mask_rects, std::vector<uint16_t>* centre_depth_mm) {
  need_handle();
  if (T.size() != up.size() || T.size() % 3 != 0) throw Exception(LM_E_INVALID, "T and up must hold 3 doubles per view");
  const int n = (int)(T.size() / 3);
  std::vector<int32_t> ids((size_t)n, -1);
  std::vector<lm_rect> rects((size_t)n);
  std::vector<uint16_t> centre((size_t)n);
  const double zero3[3] = {0, 0, 0};
  detail::check(lm_train_views(h_, mesh.handle(), &camera, n ? T.data() : zero3, n ? up.data() : zero3, n, class_id.c_str(),
                               ids.data(), nullptr, rects.data(), centre.data()));
  cache_.clear();
  if (mask_rects) {
    mask_rects->clear();
    for (int i = 0; i < n; ++i) mask_rects->push_back(Rect(rects[i].x, rects[i].y, rects[i].width, rects[i].height));
  }
  if (centre_depth_mm) *centre_depth_mm = centre;
  return std::vector<int>(ids.begin(), ids.end());
}

// writeLinemodTemplateParams (src/renderer.cpp:72-123) / readLinemodTemplateParams (src/rgbdDetector.cpp:1681-1749)
inline void writeRendererParams(const std::string& filename, const std::vector<lm_template_pose>& poses,
                                const lm_renderer_params& params) {
  detail::check(lm_write_renderer_params(filename.c_str(), poses.empty() ? nullptr : poses.data(), poses.size(), &params));
}
inline void readRendererParams(const std::string& filename, std::vector<lm_template_pose>& poses, lm_renderer_params& params) {
  lm_template_pose* p = nullptr;
  size_t n = 0;
  detail::check(lm_read_renderer_params(filename.c_str(), &p, &n, &params));
  poses.assign(p, p + n);
  lm_free_poses(p);
}

}  // namespace linemod_b200

#endif  // LINEMOD_B200_HPP_
