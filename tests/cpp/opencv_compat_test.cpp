// opencv_compat_test.cpp -- the reference's own call sites, verbatim, compiled against include/linemod_b200.hpp with
// LINEMOD_B200_WITH_OPENCV and the OpenCV 2.4 API stub under tests/cpp/opencv_stub (the image has no OpenCV C++):
//   readLinemod          /root/reference/src/rgbdDetector.cpp:1668-1680   Detector::read(FileNode) + readClass(FileNode)
//   writeLinemod         /root/reference/src/renderer.cpp:56-70           Detector::write(FileStorage) + writeClass
//   detector construction  /root/reference/src/renderer.cpp:179-185       cv::Ptr modalities, cv::Ptr<Detector>
//   linemod_detection    /root/reference/src/rgbdDetector.cpp:31-34       match(sources, threshold, matches, vector<String>(), noArray())
// The only line a maintainer adds is the namespace alias below.
//   opencv_compat_test host   persistence round trip through the FileStorage API, no CUDA device needed
//   opencv_compat_test gpu    linemod_detection on cv::Mat frames
#define LINEMOD_B200_WITH_OPENCV
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <opencv2/core/core.hpp>

#include "../../include/linemod_b200.hpp"

namespace cv { namespace linemod = ::linemod_b200; }  // was: OpenCV's own cv::linemod (opencv2/objdetect/objdetect.hpp)
using namespace cv;
using namespace std;

#define REQUIRE(cond)                                                             \
  do {                                                                            \
    if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
  } while (0)

// ---------------------------------------------------------------- verbatim: src/rgbdDetector.cpp:1668-1680 (rgbdDetector:: dropped)
cv::Ptr<cv::linemod::Detector> readLinemod(const std::string& filename)
{
  //cv::Ptr<cv::linemod::Detector> detector = cv::makePtr<cv::linemod::Detector>();
  cv::Ptr<cv::linemod::Detector> detector(new cv::linemod::Detector);
  cv::FileStorage fs(filename, cv::FileStorage::READ);
  detector->read(fs.root());

  cv::FileNode fn = fs["classes"];
  for (cv::FileNodeIterator i = fn.begin(), iend = fn.end(); i != iend; ++i)
    detector->readClass(*i);

  return detector;
}

// ---------------------------------------------------------------- verbatim: src/renderer.cpp:56-70
static void writeLinemod(const cv::Ptr<cv::linemod::Detector>& detector, const std::string& filename)
{
  cv::FileStorage fs(filename, cv::FileStorage::WRITE);
  detector->write(fs);

  std::vector<cv::String> ids = detector->classIds();
  fs << "classes" << "[";
  for (int i = 0; i < (int)ids.size(); ++i)
  {
    fs << "{";
    detector->writeClass(ids[i], fs);
    fs << "}"; // current class
  }
  fs << "]"; // classes
}

// ---------------------------------------------------------------- verbatim: src/rgbdDetector.cpp:31-34 (rgbdDetector:: dropped)
void linemod_detection(Ptr<linemod::Detector> linemod_detector,const vector<Mat>& sources,const float& threshold,std::vector<linemod::Match>& matches)
{
    linemod_detector->match (sources,threshold,matches,std::vector<String>(),noArray());
}

static cv::Ptr<cv::linemod::Detector> make_detector() {
  // ---------------------------------------------------------------- verbatim: src/renderer.cpp:179-185
    std::vector< cv::Ptr<cv::linemod::Modality> > modalities;
    modalities.push_back(cv::Ptr<cv::linemod::ColorGradient>(new cv::linemod::ColorGradient));
    modalities.push_back(cv::Ptr<cv::linemod::DepthNormal>(new cv::linemod::DepthNormal));
    std::vector<int> ensenso_T;
    ensenso_T.push_back(5);
    ensenso_T.push_back(8);
    cv::Ptr<cv::linemod::Detector> detector_(new cv::linemod::Detector(modalities,ensenso_T));
  return detector_;
}

static uint32_t rng_state = 99u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static std::vector<linemod::Template> synthetic_pyramid(int w, int h) {
  std::vector<linemod::Template> tp(4);
  for (int l = 0; l < 2; ++l)
    for (int m = 0; m < 2; ++m) {
      linemod::Template& t = tp[l * 2 + m];
      t.width = w >> l; t.height = h >> l; t.pyramid_level = l;
      const int nf = l == 0 ? 63 : 31;
      for (int i = 0; i < nf; ++i) t.features.push_back(linemod::Feature((int)(rnd() % (w >> l)), (int)(rnd() % (h >> l)), (int)(rnd() % 8)));
    }
  return tp;
}

static void same_templates(const linemod::Detector& a, const linemod::Detector& b) {
  REQUIRE(a.classIds() == b.classIds() && a.numTemplates() == b.numTemplates());
  REQUIRE(a.pyramidLevels() == b.pyramidLevels() && a.getT(0) == b.getT(0) && a.getT(1) == b.getT(1));
  std::vector<String> ids = a.classIds();
  for (size_t c = 0; c < ids.size(); ++c)
    for (int t = 0; t < a.numTemplates(ids[c]); ++t) {
      const std::vector<linemod::Template>&x = a.getTemplates(ids[c], t), &y = b.getTemplates(ids[c], t);
      REQUIRE(x.size() == y.size());
      for (size_t i = 0; i < x.size(); ++i) {
        REQUIRE(x[i].width == y[i].width && x[i].height == y[i].height && x[i].pyramid_level == y[i].pyramid_level);
        REQUIRE(x[i].features.size() == y[i].features.size());
        for (size_t k = 0; k < x[i].features.size(); ++k)
          REQUIRE(x[i].features[k].x == y[i].features[k].x && x[i].features[k].y == y[i].features[k].y && x[i].features[k].label == y[i].features[k].label);
      }
    }
}

static int run_host(const std::string& dir) {
  cv::Ptr<cv::linemod::Detector> det = make_detector();
  REQUIRE(det->getModalities().size() == 2 && det->getModalities()[1]->name() == "DepthNormal");
  for (int i = 0; i < 5; ++i) REQUIRE(det->addSyntheticTemplate(synthetic_pyramid(100 + 8 * i, 90), "obj") == i);
  REQUIRE(det->addSyntheticTemplate(synthetic_pyramid(64, 72), "another") == 0);
  writeLinemod(det, dir + "/templates.yml");
  cv::Ptr<cv::linemod::Detector> back = readLinemod(dir + "/templates.yml");
  same_templates(*det, *back);
  // the same model through the library's own YAML writer / reader (the on-disk format) agrees with the FileStorage walk
  det->write(dir + "/templates_disk.yml");
  linemod::Detector disk;
  disk.read(dir + "/templates_disk.yml");
  same_templates(*back, disk);
  // upstream's CV_Asserts: a class may be read once; modalities and pyramid_levels of a class must match the detector
  cv::FileStorage fs(dir + "/templates.yml", cv::FileStorage::READ);
  bool threw = false;
  try { back->readClass(*fs["classes"].begin()); } catch (const linemod::Exception&) { threw = true; }
  REQUIRE(threw);
  REQUIRE(back->readClass(*fs["classes"].begin(), "renamed") == "renamed" && back->numTemplates("renamed") == back->numTemplates("another"));
  std::vector< cv::Ptr<cv::linemod::Modality> > one;
  one.push_back(cv::Ptr<cv::linemod::ColorGradient>(new cv::linemod::ColorGradient));
  cv::linemod::Detector rgb_only(one, std::vector<int>(2, 4));
  threw = false;
  try { rgb_only.readClass(*fs["classes"].begin()); } catch (const linemod::Exception&) { threw = true; }
  REQUIRE(threw);
  std::printf("ok host\n");
  return 0;
}

static int run_gpu(const std::string& dir) {
  cv::Ptr<cv::linemod::Detector> det = make_detector();
  // a textured box on a plane, as cv::Mat
  const int rows = 480, cols = 640;
  cv::Mat bgr(rows, cols, CV_8UC3), depth(rows, cols, CV_16UC1), mask(rows, cols, CV_8UC1);
  const int ox = 200, oy = 150, bw = 130, bh = 110;
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      uchar* p = bgr.data + (size_t)y * bgr.step[0] + 3 * x;
      unsigned short* d = reinterpret_cast<unsigned short*>(depth.data + (size_t)y * depth.step[0]) + x;
      p[0] = (uchar)(90 + (x * 40) / cols); p[1] = (uchar)(100 + (y * 30) / rows); p[2] = 110;
      *d = (unsigned short)(900 + y / 4);
      mask.data[(size_t)y * mask.step[0] + x] = 0;
      const int u = x - ox, v = y - oy;
      if (u >= 0 && u < bw && v >= 0 && v < bh) {   // a textured box with two depth facets on a tilted ground plane
        mask.data[(size_t)y * mask.step[0] + x] = 255;
        const int cell = ((u / 12) + (v / 12)) & 1, stripe = (u / 7) % 3;
        p[0] = (uchar)(cell ? 30 : 10); p[1] = (uchar)(stripe == 0 ? 250 : 225); p[2] = (uchar)(cell ? 20 : 40);
        *d = (unsigned short)(600 + (u * 3) / 2 + ((v / 20) % 2 ? v : -v) / 2);
      }
    }
  std::vector<cv::Mat> sources;
  sources.push_back(bgr);
  sources.push_back(depth);
  cv::Rect bb;
  REQUIRE(det->addTemplate(sources, "obj", mask, &bb) == 0 && bb.width > 60);
  std::vector<linemod::Match> matches;
  linemod_detection(det, sources, 90.f, matches);
  REQUIRE(!matches.empty() && matches[0].class_id == "obj" && matches[0].similarity >= 99.f);
  // quantized_images as an OutputArrayOfArrays, like OpenCV's signature
  std::vector<cv::Mat> quantized;
  std::vector<linemod::Match> again;
  det->match(sources, 90.f, again, std::vector<String>(), quantized);
  REQUIRE(again.size() == matches.size() && quantized.size() == 4 && quantized[0].rows == 480 && quantized[2].cols == 320);
  // [OCV] Detector::addTemplate spelled out through Modality::process / QuantizedPyramid, the way upstream implements it:
  // the extracted templates are the stored ones before cropTemplates shifted them by the bounding-box origin, and
  // quantize() is the (masked) quantised image match() reports for the same mask.
  std::vector<cv::Mat> masks(2, mask), masked_q;
  det->match(sources, 90.f, again, std::vector<String>(), masked_q, masks);
  const size_t M = det->getModalities().size();
  for (size_t m = 0; m < M; ++m) {
    cv::Ptr<cv::linemod::QuantizedPyramid> qp = det->getModalities()[m]->process(sources[m], mask);
    for (int l = 0; l < det->pyramidLevels(); ++l) {
      if (l > 0) qp->pyrDown();
      cv::Mat q;
      qp->quantize(q);
      const cv::Mat& want = masked_q[(size_t)l * M + m];
      REQUIRE(q.rows == (rows >> l) && q.cols == (cols >> l) && want.rows == q.rows && want.cols == q.cols);
      for (int y = 0; y < q.rows; ++y) REQUIRE(std::memcmp(q.data + (size_t)y * q.step[0], want.data + (size_t)y * want.step[0], (size_t)q.cols) == 0);
      cv::linemod::Template t;
      REQUIRE(qp->extractTemplate(t) && t.width == -1 && t.height == -1 && t.pyramid_level == l);
      const cv::linemod::Template& stored = det->getTemplates("obj", 0)[(size_t)l * M + m];
      REQUIRE(t.features.size() == stored.features.size() && !t.features.empty());
      for (size_t k = 0; k < t.features.size(); ++k)
        REQUIRE(t.features[k].x == stored.features[k].x + (bb.x >> l) && t.features[k].y == stored.features[k].y + (bb.y >> l) &&
                t.features[k].label == stored.features[k].label);
    }
  }
  // Modality::write / create(FileNode) round trip
  {
    cv::FileStorage out(dir + "/modality.yml", cv::FileStorage::WRITE);
    out << "m" << "{";
    det->getModalities()[1]->write(out);
    out << "}";
    out.release();
    cv::FileStorage in(dir + "/modality.yml", cv::FileStorage::READ);
    cv::Ptr<cv::linemod::Modality> back = cv::linemod::Modality::create(in["m"]);
    REQUIRE(back->name() == "DepthNormal" && back->desc().extract_threshold == det->getModalities()[1]->desc().extract_threshold);
  }
  std::printf("ok gpu (%zu matches)\n", matches.size());
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: opencv_compat_test host|gpu <tmpdir>\n"); return 2; }
  try {
    return std::strcmp(argv[1], "gpu") == 0 ? run_gpu(argv[2]) : run_host(argv[2]);
  } catch (const linemod::Exception& e) {
    std::fprintf(stderr, "linemod_b200::Exception %d: %s\n", e.code, e.what());
    return 1;
  }
}
