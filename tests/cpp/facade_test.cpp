// facade_test.cpp -- exercises include/linemod_b200.hpp the way the reference drives cv::linemod::Detector
// (/root/reference/src/renderer.cpp:179-185,308; src/rgbdDetector.cpp:31-34,1668-1680).
//   facade_test host <tmpdir>   template bookkeeping + persistence, no CUDA device needed
//   facade_test gpu  <tmpdir>   addTemplate + match on the GPU
// Prints "ok <mode>" and exits 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/linemod_b200.hpp"

namespace lm = linemod_b200;

#define REQUIRE(cond)                                                             \
  do {                                                                            \
    if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
  } while (0)

static std::shared_ptr<lm::Detector> make_detector() {
  std::vector<std::shared_ptr<lm::Modality> > modalities;
  modalities.push_back(std::make_shared<lm::ColorGradient>());
  modalities.push_back(std::make_shared<lm::DepthNormal>());
  std::vector<int> T;
  T.push_back(5);
  T.push_back(8);
  return std::make_shared<lm::Detector>(modalities, T);
}

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static std::vector<lm::Template> synthetic_pyramid(int w, int h) {
  std::vector<lm::Template> tp(4);  // index l*M+m, L = 2, M = 2
  for (int l = 0; l < 2; ++l)
    for (int m = 0; m < 2; ++m) {
      lm::Template& t = tp[l * 2 + m];
      t.width = w >> l; t.height = h >> l; t.pyramid_level = l;
      const int nf = l == 0 ? 63 : 31;
      for (int i = 0; i < nf; ++i) t.features.push_back(lm::Feature((int)(rnd() % (w >> l)), (int)(rnd() % (h >> l)), (int)(rnd() % 8)));
    }
  return tp;
}

static int run_host(const std::string& dir) {
  std::shared_ptr<lm::Detector> det = make_detector();
  REQUIRE(det->pyramidLevels() == 2 && det->getT(0) == 5 && det->getT(1) == 8);
  REQUIRE(det->getModalities().size() == 2 && det->getModalities()[0]->name() == "ColorGradient");
  std::vector<lm::Template> a = synthetic_pyramid(120, 100), b = synthetic_pyramid(80, 90);
  REQUIRE(det->addSyntheticTemplate(a, "obj") == 0);
  REQUIRE(det->addSyntheticTemplate(b, "obj") == 1);
  REQUIRE(det->addSyntheticTemplate(b, "another") == 0);
  REQUIRE(det->numTemplates() == 3 && det->numTemplates("obj") == 2 && det->numClasses() == 2);
  std::vector<std::string> ids = det->classIds();
  REQUIRE(ids.size() == 2 && ids[0] == "another" && ids[1] == "obj");  // std::map order
  const std::vector<lm::Template>& got = det->getTemplates("obj", 1);
  REQUIRE(got.size() == 4 && got[0].width == 80 && got[3].features.size() == 31);
  for (size_t i = 0; i < got.size(); ++i)
    for (size_t j = 0; j < got[i].features.size(); ++j)
      REQUIRE(got[i].features[j].x == b[i].features[j].x && got[i].features[j].y == b[i].features[j].y &&
              got[i].features[j].label == b[i].features[j].label);
  // writeLinemod / readLinemod round trip
  const std::string path = dir + "/templates.yml";
  det->write(path);
  lm::Detector loaded;
  REQUIRE(loaded.numTemplates() == 0);
  loaded.read(path);
  REQUIRE(loaded.numTemplates() == 3 && loaded.pyramidLevels() == 2 && loaded.getT(1) == 8);
  REQUIRE(loaded.getModalities().size() == 2 && loaded.getModalities()[1]->name() == "DepthNormal");
  const std::vector<lm::Template>& again = loaded.getTemplates("obj", 0);
  REQUIRE(again.size() == 4 && again[0].features.size() == 63 && again[0].features[5].x == a[0].features[5].x);
  // per-class files
  det->writeClasses(dir + "/cls_%s.yml.gz");
  std::shared_ptr<lm::Detector> fresh = make_detector();
  fresh->readClasses(ids, dir + "/cls_%s.yml.gz");
  REQUIRE(fresh->numTemplates("obj") == 2 && fresh->numTemplates("another") == 1);
  // CV_Assert-style failures become exceptions
  bool threw = false;
  try { std::vector<lm::Template> bad(3); det->addSyntheticTemplate(bad, "obj"); } catch (const lm::Exception& e) { threw = e.code == LM_E_INVALID; }
  REQUIRE(threw);
  threw = false;
  try { lm::Detector empty; std::vector<lm::Match> m; empty.match(std::vector<lm::Image>(), 90.f, m); } catch (const lm::Exception& e) { threw = e.code == LM_E_STATE; }
  REQUIRE(threw);
  threw = false;
  try { lm::Detector missing; missing.read(dir + "/does_not_exist.yml"); } catch (const lm::Exception& e) { threw = e.code == LM_E_IO; }
  REQUIRE(threw);
  // Match ordering / equality as std::sort + std::unique see them
  lm::Match m1(1, 2, 95.f, "obj", 7), m2(1, 2, 95.f, "obj", 3), m3(4, 2, 96.f, "obj", 9);
  REQUIRE(m3 < m1 && m2 < m1 && m1 == m2 && !(m1 == m3));
  // the trainer's view sphere and pose table (src/renderer.cpp:242-246, 72-123): host-only
  lm::ViewSphere sphere(150, 10, 0.5f, 1.0f, 0.1f);
  REQUIRE(sphere.size() == 150 * 17 * 6);
  double T[3], up[3];
  float radius = 0;
  sphere.view(3672, T, up, &radius);   // first template of the reference's shipped renderer_params.yml (T = -camera position)
  REQUIRE(std::fabs(T[0] - 2.0922734402120113e-03) < 1e-12 && std::fabs(T[1] + 2.5666666030883789e-01) < 1e-12 &&
          std::fabs(T[2] + 4.2908957600593567e-01) < 1e-12 && radius == 0.5f);
  std::vector<lm_template_pose> poses(2);
  std::memset(poses.data(), 0, sizeof(lm_template_pose) * 2);
  for (int k = 0; k < 2; ++k) {
    REQUIRE(lm_view_pose(T, up, poses[k].R, poses[k].T) == LM_OK);
    for (int j = 0; j < 3; ++j) poses[k].T[j] = -T[j];
    poses[k].K[0] = 535.566011f; poses[k].K[2] = 320.f; poses[k].K[4] = 537.168115f; poses[k].K[5] = 240.f; poses[k].K[8] = 1.f;
    poses[k].D = 0.047 + k; poses[k].ori_dist = radius; poses[k].rect.x = 253; poses[k].rect.width = 134 + k;
  }
  REQUIRE(std::fabs(poses[0].R[6] + 4.1845467435124026e-03) < 1e-12 && std::fabs(poses[0].R[0] - 9.7591209808210677e-01) < 1e-12);
  lm_renderer_params rp = {150, 10, 0.5, 1.0, 0.1, 640, 480, 535.566011, 537.168115, 0.1, 1000.0};
  lm::writeRendererParams(dir + "/renderer_params.yml", poses, rp);
  std::vector<lm_template_pose> back;
  lm_renderer_params rp2;
  lm::readRendererParams(dir + "/renderer_params.yml", back, rp2);
  REQUIRE(back.size() == 2 && std::memcmp(back.data(), poses.data(), sizeof(lm_template_pose) * 2) == 0);
  REQUIRE(rp2.n_points == 150 && rp2.width == 640 && rp2.far_ == 1000.0);
  std::printf("ok host\n");
  return 0;
}

// A textured box on a tilted ground plane: enough gradient and normal structure for both modalities.
static void scene(int rows, int cols, int ox, int oy, std::vector<uint8_t>& bgr, std::vector<uint16_t>& depth,
                  std::vector<uint8_t>& mask, int bw, int bh) {
  bgr.assign((size_t)rows * cols * 3, 0); depth.assign((size_t)rows * cols, 0); mask.assign((size_t)rows * cols, 0);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      uint8_t* p = &bgr[((size_t)y * cols + x) * 3];
      p[0] = (uint8_t)(90 + (x * 40) / cols); p[1] = (uint8_t)(100 + (y * 30) / rows); p[2] = 110;
      depth[(size_t)y * cols + x] = (uint16_t)(900 + y / 4);
      const int u = x - ox, v = y - oy;
      if (u >= 0 && u < bw && v >= 0 && v < bh) {
        mask[(size_t)y * cols + x] = 255;
        const int cell = ((u / 12) + (v / 12)) & 1, stripe = (u / 7) % 3;
        p[0] = (uint8_t)(cell ? 30 : 10); p[1] = (uint8_t)(stripe == 0 ? 250 : 225); p[2] = (uint8_t)(cell ? 20 : 40);
        depth[(size_t)y * cols + x] = (uint16_t)(600 + (u * 3) / 2 + ((v / 20) % 2 ? v : -v) / 2);
      }
    }
}

static int run_gpu(const std::string& dir) {
  std::shared_ptr<lm::Detector> det = make_detector();
  const int bw = 96, bh = 88;
  std::vector<uint8_t> bgr, mask;
  std::vector<uint16_t> depth;
  scene(240, 240, 70, 60, bgr, depth, mask, bw, bh);
  std::vector<lm::Image> sources;
  sources.push_back(lm::Image(bgr.data(), 240, 240, LM_8UC3));
  sources.push_back(lm::Image(depth.data(), 240, 240, LM_16UC1));
  lm::Rect bb;
  int id = det->addTemplate(sources, "obj", lm::Image(mask.data(), 240, 240, LM_8UC1), &bb);
  REQUIRE(id == 0);
  REQUIRE(bb.width > 40 && bb.height > 40 && bb.x >= 60 && bb.y >= 50);
  // the same object elsewhere in a 640x480 frame
  const int ox = 301, oy = 177;
  scene(480, 640, ox, oy, bgr, depth, mask, bw, bh);
  std::vector<lm::Image> frame;
  frame.push_back(lm::Image(bgr.data(), 480, 640, LM_8UC3));
  frame.push_back(lm::Image(depth.data(), 480, 640, LM_16UC1));
  std::vector<lm::Match> matches;
  std::vector<std::vector<uint8_t> > quantized;
  det->match(frame, 80.f, matches, std::vector<std::string>(), &quantized);
  REQUIRE(!matches.empty());
  REQUIRE(quantized.size() == 4 && quantized[0].size() == 640u * 480u && quantized[2].size() == 320u * 240u);
  for (size_t i = 1; i < matches.size(); ++i) REQUIRE(!(matches[i] < matches[i - 1]));  // sorted
  const lm::Match& best = matches[0];
  REQUIRE(best.class_id == "obj" && best.template_id == 0 && best.similarity >= 80.f);
  // Match.x/y is where the template's (0,0) lands: the bounding-box corner shifted by the planted offset
  REQUIRE(std::abs(best.x - (bb.x - 70 + ox)) <= 8 && std::abs(best.y - (bb.y - 60 + oy)) <= 8);
  // class filter with an unknown id yields nothing; persistence keeps matching identical
  std::vector<lm::Match> none;
  det->match(frame, 80.f, none, std::vector<std::string>(1, "nope"));
  REQUIRE(none.empty());
  det->write(dir + "/gpu_templates.yml");
  lm::Detector loaded;
  loaded.read(dir + "/gpu_templates.yml");
  std::vector<lm::Match> again;
  loaded.match(frame, 80.f, again);
  REQUIRE(again.size() == matches.size());
  for (size_t i = 0; i < again.size(); ++i)
    REQUIRE(again[i] == matches[i] && again[i].template_id == matches[i].template_id);
  // the trainer's loop on a box mesh: every view rendered and added on the GPU, then one rendered view is found again
  const float hx = 0.06f, hy = 0.04f, hz = 0.03f;
  const float v[8][3] = {{-hx, -hy, -hz}, {hx, -hy, -hz}, {-hx, hy, -hz}, {hx, hy, -hz}, {-hx, -hy, hz}, {hx, -hy, hz}, {-hx, hy, hz}, {hx, hy, hz}};
  const int faces[6][4] = {{0, 2, 3, 1}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 6, 7, 3}, {0, 4, 6, 2}, {1, 3, 7, 5}};
  std::vector<float> tris;
  for (int f = 0; f < 6; ++f) {
    const int order[6] = {0, 1, 2, 0, 2, 3};
    for (int k = 0; k < 6; ++k)
      for (int c = 0; c < 3; ++c) tris.push_back(v[faces[f][order[k]]][c]);
  }
  lm::Mesh mesh(tris.data(), 12);
  REQUIRE(mesh.numTriangles() == 12);
  lm_camera cam = {640, 480, 535.566011, 537.168115, 0.1, 1000.0};
  lm::ViewSphere sphere(20, 40, 0.4f, 0.5f, 0.1f);
  std::vector<double> T, up;
  sphere.views(T, up);
  std::shared_ptr<lm::Detector> trained = make_detector();
  std::vector<lm::Rect> rects;
  std::vector<uint16_t> centre;
  std::vector<int> ids = trained->trainViews(mesh, cam, T, up, "box", &rects, &centre);
  REQUIRE((int)ids.size() == sphere.size() && trained->numTemplates("box") > 100);
  int pick = -1;
  for (size_t i = 37; i < ids.size() && pick < 0; ++i)
    if (ids[i] >= 0) pick = (int)i;
  REQUIRE(pick >= 0 && rects[pick].width > 40 && centre[pick] > 300 && centre[pick] < 500);
  std::vector<uint8_t> rb(640 * 480 * 3), rm(640 * 480);
  std::vector<uint16_t> rd(640 * 480);
  lm_rect rr;
  REQUIRE(lm_render_views(trained->handle(), mesh.handle(), &cam, &T[3 * pick], &up[3 * pick], 1, rb.data(), rd.data(), rm.data(), &rr) == LM_OK);
  REQUIRE(rr.x == rects[pick].x && rr.width == rects[pick].width);
  std::vector<lm::Image> rframe;
  rframe.push_back(lm::Image(rb.data(), 480, 640, LM_8UC3));
  rframe.push_back(lm::Image(rd.data(), 480, 640, LM_16UC1));
  std::vector<lm::Match> found;
  trained->match(rframe, 90.f, found);
  REQUIRE(!found.empty() && found[0].similarity >= 99.f);
  bool self = false;
  for (size_t i = 0; i < found.size() && found[i].similarity == found[0].similarity; ++i) self = self || found[i].template_id == ids[pick];
  REQUIRE(self);
  // several GPUs behind one caller (lm_group): both modes return the single-handle lists (members share device 0 here)
  {
    std::vector<int> devs(3, 0);
    std::vector<std::vector<lm::Image> > batch(5, rframe);
    batch[1] = frame; batch[3] = frame;
    std::vector<lm::Match> want_r, want_f;
    trained->match(rframe, 90.f, want_r);
    trained->match(frame, 90.f, want_f);
    for (int mode = 0; mode < 2; ++mode) {
      lm::DetectorGroup group(*trained, devs, mode == 0 ? lm::DetectorGroup::Frames : lm::DetectorGroup::Templates);
      REQUIRE(group.size() == 3);
      group.setOption("batch_frames", 2);
      std::vector<std::vector<lm::Match> > got;
      group.matchBatch(batch, 90.f, got);
      REQUIRE(got.size() == 5);
      for (size_t f = 0; f < got.size(); ++f) {
        const std::vector<lm::Match>& want = (f == 1 || f == 3) ? want_f : want_r;
        REQUIRE(got[f].size() == want.size());
        for (size_t i = 0; i < want.size(); ++i) REQUIRE(got[f][i] == want[i] && got[f][i].template_id == want[i].template_id);
      }
      std::vector<lm::Match> one;
      group.match(rframe, 90.f, one);
      REQUIRE(one.size() == want_r.size());
    }
  }
  // a stream of host frames (lm_stream): frames pushed in pieces, finished frames popped in order, equal to match() per frame;
  // while the stream is open the detector refuses other matching calls
  {
    std::vector<std::vector<lm::Image> > batch(11, rframe);
    batch[1] = frame; batch[3] = frame; batch[10] = frame;
    std::vector<lm::Match> want_r, want_f;
    trained->match(rframe, 90.f, want_r);
    trained->match(frame, 90.f, want_f);
    REQUIRE(lm_set_option(trained->handle(), "stream_frames", 4) == LM_OK);
    {
      lm::FrameStream stream(*trained, 90.f);
      bool refused = false;
      try { std::vector<lm::Match> tmp; trained->match(rframe, 90.f, tmp); } catch (const lm::Exception& e) { refused = e.code == LM_E_STATE; }
      REQUIRE(refused);
      std::vector<std::vector<lm::Match> > got, part;
      stream.push(std::vector<std::vector<lm::Image> >(batch.begin(), batch.begin() + 3));
      stream.push(std::vector<std::vector<lm::Image> >(batch.begin() + 3, batch.begin() + 9));
      stream.pop(part);                       // whatever is ready
      got.insert(got.end(), part.begin(), part.end());
      stream.push(std::vector<std::vector<lm::Image> >(batch.begin() + 9, batch.end()));
      REQUIRE(stream.inFlight() == (int)(batch.size() - got.size()));
      stream.pop(part, true);                 // the rest
      got.insert(got.end(), part.begin(), part.end());
      REQUIRE(got.size() == batch.size() && stream.inFlight() == 0);
      for (size_t f = 0; f < got.size(); ++f) {
        const std::vector<lm::Match>& want = (f == 1 || f == 3 || f == 10) ? want_f : want_r;
        REQUIRE(got[f].size() == want.size());
        for (size_t i = 0; i < want.size(); ++i) REQUIRE(got[f][i] == want[i] && got[f][i].template_id == want[i].template_id);
      }
    }
    std::vector<lm::Match> again;
    trained->match(rframe, 90.f, again);      // the stream is closed: the handle answers again
    REQUIRE(again.size() == want_r.size());
    REQUIRE(lm_set_option(trained->handle(), "stream_frames", 16) == LM_OK);
  }
  std::printf("ok gpu (%zu matches, best %.2f at %d,%d)\n", matches.size(), best.similarity, best.x, best.y);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: facade_test host|gpu <tmpdir>\n"); return 2; }
  try {
    return std::strcmp(argv[1], "gpu") == 0 ? run_gpu(argv[2]) : run_host(argv[2]);
  } catch (const lm::Exception& e) {
    std::fprintf(stderr, "linemod_b200::Exception %d: %s\n", e.code, e.what());
    return 1;
  }
}
