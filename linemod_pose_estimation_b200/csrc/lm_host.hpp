// lm_host.hpp -- host-side data model of the detector: templates, class map, template extraction, persistence.
//
// Mirrors the types the reference's callers touch ([OCV] cv::linemod::{Feature, Template, Detector};
// used at /root/reference/src/linemod_ensenso_detect_3_mult_detect_service.cpp:351,741-744 and
// src/rgbdDetector.cpp:46-48).  Nothing here computes on images with the CPU except the inherently sequential
// feature selection of addTemplate (stable sort + greedy scattered selection), which consumes quantised maps
// produced by the CUDA front end.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/linemod_b200.h"

namespace lm {

struct Feature {
  int x, y, label;
};
struct Template {
  int width = 0, height = 0, pyramid_level = 0;
  std::vector<Feature> features;
};
typedef std::vector<Template> TemplatePyramid;                             // index l*M + m
typedef std::map<std::string, std::vector<TemplatePyramid> > TemplatesMap;  // std::map order == reference order

struct HostModel {
  std::vector<int> T;                  // T_at_level
  std::vector<lm_modality_desc> mods;  // modalities
  TemplatesMap classes;                // class_templates
  uint64_t version = 1;                // bumped on every template change (device pack invalidation)
  int levels() const { return (int)T.size(); }
  int M() const { return (int)mods.size(); }
};

const char* modality_name(int type);  // "ColorGradient" / "DepthNormal"
bool modality_from_name(const std::string& name, int& type);
lm_modality_desc default_modality(int type);

// [OCV] ColorGradientPyramid::extractTemplate on GPU-quantised inputs. mask may be null (no mask).
bool extract_color_gradient(const uint8_t* quantized, const float* magnitude, const uint8_t* mask, int rows, int cols,
                            float strong_threshold, int num_features, int level, Template& out);
// [OCV] DepthNormalPyramid::extractTemplate
bool extract_depth_normal(const uint8_t* normal, const uint8_t* mask, int rows, int cols, int num_features,
                          int extract_threshold, int level, Template& out);
// [OCV] cropTemplates
lm_rect crop_templates(TemplatePyramid& tp);

// Persistence (formats: SURVEY.md App. B).  All return false and set err on failure.
bool load_detector_yaml(const std::string& path, HostModel& model, std::string& err);       // readLinemod()
bool save_detector_yaml(const HostModel& model, const std::string& path, std::string& err);  // writeLinemod()
bool load_class_file(const std::string& path, HostModel& model, std::string& err);           // readClasses(): one file
bool save_class_file(const HostModel& model, const std::string& class_id, const std::string& path, std::string& err);

// Binary template cache: the same model as flat arrays, checksummed; loads without a YAML parse.
bool save_model_cache(const HostModel& model, const std::string& path, std::string& err);
bool load_model_cache(const std::string& path, HostModel& model, std::string& err);

}  // namespace lm
