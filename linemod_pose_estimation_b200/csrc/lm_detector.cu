// lm_detector.cu -- host orchestration behind the C ABI (include/linemod_b200.h): workspaces, template packing,
// the per-frame kernel sequence of Detector::match and the match-list finalisation.
//
// Reference surface mirrored: cv::linemod::Detector as driven by /root/reference/src/rgbdDetector.cpp:31-34 (match),
// src/renderer.cpp:179-185,308 (construction, addTemplate), src/rgbdDetector.cpp:1668-1680 / src/renderer.cpp:56-70
// (persistence).  Everything that touches pixels runs in the CUDA kernels of lm_frontend.cu / lm_match.cu; the host
// only stages buffers, packs template records and orders the (few) surviving matches.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "lm_host.hpp"
#include "lm_kernels.cuh"

using namespace lm;
using namespace lmk;

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return fail(LM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                               \
  } while (0)

// ------------------------------------------------------------------------------------------------ buffers
struct DevBuf {  // grow-only device allocation
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, bool* grew = nullptr) {
    if (grew) *grew = false;
    if (bytes <= cap) return LM_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    CU(cudaMalloc(&p, want));
    cap = want;
    if (grew) *grew = true;
    return LM_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {  // grow-only page-locked host allocation
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return LM_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    CU(cudaMallocHost(&p, bytes + 256));
    cap = bytes + 256;
    return LM_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct LevelGeom {
  int rows = 0, cols = 0, T = 0, W = 0, H = 0;
  size_t plane_stride = 0;
};
// Nibble-packed rows of a level start on 32-bit words when W and W*H are multiples of 8 (one word = 8 positions).
static bool level_nibble_aligned(const LevelGeom& g) { return (g.W % 8) == 0 && ((size_t)g.W * g.H) % 8 == 0; }

static size_t plane_stride_of(int T, int W, int H) {
  size_t wh = (size_t)W * H;
  return ((size_t)T * T * wh + wh + 16 * (size_t)W + 16 + 15) & ~(size_t)15;  // same rule as the oracle (App. D-2)
}
static const size_t kLmSlack = 8192;  // tail slack: vector loads of partially filled passes may over-read

// Frames in flight per handle: lm_match_batch* pipelines this many frames (H2D copy, kernels, D2H copy of different
// frames overlap), lm_match_device_multi_lane exposes them to callers that manage their own streams.
static const int LM_LANES = 8;

// One in-flight frame: stream, events, device workspace, pinned staging.
struct Lane {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // modalities quantise concurrently: modality m > 0 runs on side[m-1], forked from / joined into the frame's stream
  cudaStream_t side[LM_MAX_MODALITIES - 1] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[LM_MAX_MODALITIES - 1] = {nullptr, nullptr, nullptr};
  int rows = 0, cols = 0;     // geometry of the quantisation workspace
  bool lm_ready = false;      // LM buffers sized + zero-tailed for (rows, cols)
  bool front_valid = false;
  bool debug_taps_written = false;
  bool bytes_valid[LM_MAX_LEVELS] = {false, false, false, false};    // byte planes written by the last front end
  bool nibbles_valid[LM_MAX_LEVELS] = {false, false, false, false};  // nibble planes written by the last front end
  std::vector<LevelGeom> geom;
  // per modality
  DevBuf src[LM_MAX_MODALITIES];       // level-0 source (BGR / depth)
  const void* src_ptr[LM_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};  // own buffer or caller's device ptr
  DevBuf mask0[LM_MAX_MODALITIES];
  bool has_mask[LM_MAX_MODALITIES] = {false, false, false, false};
  // per (level, modality)
  DevBuf bgr[LM_MAX_LEVELS][LM_MAX_MODALITIES];       // CG pyramid sources for level >= 1
  DevBuf smoothed[LM_MAX_MODALITIES], qunf[LM_MAX_MODALITIES], dn_raw[LM_MAX_MODALITIES];  // scratch, reused per level
  DevBuf mag[LM_MAX_LEVELS][LM_MAX_MODALITIES];
  DevBuf quant_raw[LM_MAX_LEVELS][LM_MAX_MODALITIES];
  DevBuf quantized[LM_MAX_LEVELS][LM_MAX_MODALITIES];
  DevBuf spread[LM_MAX_LEVELS][LM_MAX_MODALITIES], response[LM_MAX_LEVELS][LM_MAX_MODALITIES];  // parity taps only
  DevBuf lmem[LM_MAX_LEVELS];                          // [M][8][plane_stride] + slack
  DevBuf lmn[LM_MAX_LEVELS];                           // the same planes nibble-packed (two positions per byte): what the
                                                       // matching kernels read when the level's rows are word-aligned
  // The GPU work of one frame (front end, header memset, coarse, refine) as an instantiated CUDA graph: the batch and
  // device-resident paths replay it instead of ~20 runtime calls per frame.  Valid while `gkey` matches.
  struct GraphKey {
    const void* plan; const void* plan_recs; const void* cand; const void* result; const void* src[LM_MAX_MODALITIES];
    uint64_t model_version; int rows, cols, n_q, n_tiles, variant, prune, frontend, shard_rank, shard_world; uint32_t cand_cap, out_cap;
    float thr[LM_MAX_QUERIES];
  };
  cudaGraphExec_t gexec = nullptr;
  GraphKey gkey;
  int graph_launches = 0;
  bool graph_broken = false;  // capture failed once on this lane: stay on the eager path
  // matching
  DevBuf cand, work, work_order, dump, dbg_recs;
  DevBuf mod_bits;  // per modality: orientation bits set in the coarsest level's spread image (front end -> coarse kernel hint)
  struct Ref {  // this lane's result block inside the detector-wide allocation (lm_detector::results_all)
    void* p = nullptr;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  } result;
  uint32_t cand_cap = 0, out_cap = 0;
  PinBuf stage_in, stage_out;
  // last-call bookkeeping
  float ms[5] = {0, 0, 0, 0, 0};
  int launches = 0;
  uint64_t work_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<lm_match_rec> presort;

  int init() {
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (int i = 0; i < 6; ++i) CU(cudaEventCreate(&ev[i]));
    CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    if (mod_bits.ensure(sizeof(unsigned int) * LM_MAX_MODALITIES) != LM_OK) return LM_E_CUDA;
    for (int i = 0; i < LM_MAX_MODALITIES - 1; ++i) {
      CU(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
    }
    return LM_OK;
  }
  void destroy() {
    for (int m = 0; m < LM_MAX_MODALITIES; ++m) {
      src[m].release(); mask0[m].release(); smoothed[m].release(); qunf[m].release(); dn_raw[m].release();
      for (int l = 0; l < LM_MAX_LEVELS; ++l) {
        bgr[l][m].release(); mag[l][m].release(); quant_raw[l][m].release(); quantized[l][m].release();
        spread[l][m].release(); response[l][m].release();
      }
    }
    for (int l = 0; l < LM_MAX_LEVELS; ++l) lmem[l].release();
    for (int l = 0; l < LM_MAX_LEVELS; ++l) lmn[l].release();
    cand.release(); work.release(); work_order.release(); dump.release(); dbg_recs.release(); mod_bits.release();
    stage_in.release(); stage_out.release();
    if (gexec) cudaGraphExecDestroy(gexec);
    for (int i = 0; i < 6; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    if (ev_fork) cudaEventDestroy(ev_fork);
    for (int i = 0; i < LM_MAX_MODALITIES - 1; ++i) {
      if (ev_join[i]) cudaEventDestroy(ev_join[i]);
      if (side[i]) cudaStreamDestroy(side[i]);
    }
    if (stream) cudaStreamDestroy(stream);
  }
};

// Device-resident template records for one frame geometry.
struct Pack {
  uint64_t version = 0;
  int rows = 0, cols = 0, shard_rank = 0, shard_world = 1, variant = -1;
  int n = 0;        // templates on this shard
  int max_P = 0;
  DevBuf ctpl, foff;
  DevBuf rtpl[LM_MAX_LEVELS], rfeats[LM_MAX_LEVELS];
  std::vector<CoarseTpl> h_ctpl;
  std::vector<uint32_t> h_foff;         // host copy of the coarse feature offsets (tile records are built from it)
  std::vector<uint64_t> coarse_bytes;   // per template: in-bounds features x positions (B_coarse)
  uint64_t coarse_bytes_all = 0;
  uint64_t refine_bytes_per_cand = 0;   // approximate (first template); exact per candidate is computed at finalise
  std::vector<uint32_t> refine_nf;      // per template: sum over refine levels of features (x256 = bytes / candidate)
  struct ClassRange { std::string id; int class_index; std::vector<uint32_t> local; std::vector<uint32_t> global_pos; };
  std::vector<ClassRange> classes;      // canonical order
  // Device-side description of one (multi-query) request: work items and coarse tiles.  Cached by the class lists.
  struct Plan { DevBuf items, tiles, recs; int n_items = 0, n_tiles = 0, rec_words = 0; uint64_t coarse_bytes = 0, evals = 0; double refine_nf_sum = 0; };
  std::map<std::string, Plan> plans;
  void clear_filtered() {
    for (auto& kv : plans) { kv.second.items.release(); kv.second.tiles.release(); kv.second.recs.release(); }
    plans.clear();
  }
  void release() {
    clear_filtered();
    ctpl.release(); foff.release();
    for (int l = 0; l < LM_MAX_LEVELS; ++l) { rtpl[l].release(); rfeats[l].release(); }
  }
};

// Workspace of the batched trainer / renderer (lm_train_views, lm_add_templates_batch, lm_render_views, lm_depth_diff_batch).
struct TrainWs {
  DevBuf zbuf, nz_abs, views, rects;             // rasteriser: u64 z-buffer per view, per (view, triangle) shading, poses
  DevBuf src[LM_MAX_MODALITIES], mask;           // per-view source images and masks of a batch, tightly packed
  DevBuf segs, pool, feats;                      // TrainSeg table, candidate key pool, selected features [seg][64]
  DevBuf pb[LM_LANES][LM_MAX_LEVELS];            // DepthNormal scratch per lane and level
  DevBuf runs[LM_LANES][LM_MAX_LEVELS];          // DepthNormal run tables per lane and level (u16)
  DevBuf scene, diff;                            // lm_depth_diff_batch: scene depth, [n][2] sums / counts
  PinBuf h_rects, h_segs, h_feats, h_stage;
  cudaEvent_t ev[LM_LANES] = {};
  void release() {
    zbuf.release(); nz_abs.release(); views.release(); rects.release(); mask.release();
    for (int m = 0; m < LM_MAX_MODALITIES; ++m) src[m].release();
    segs.release(); pool.release(); feats.release(); scene.release(); diff.release();
    for (int i = 0; i < LM_LANES; ++i)
      for (int l = 0; l < LM_MAX_LEVELS; ++l) { pb[i][l].release(); runs[i][l].release(); }
    h_rects.release(); h_segs.release(); h_feats.release(); h_stage.release();
    for (int i = 0; i < LM_LANES; ++i) if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
  }
};

struct lm_detector {
  HostModel model;
  TrainWs train;
  int device = -1;
  bool cuda_ready = false;
  uint8_t sim_lut[256];
  uint8_t normal_lut[8000];
  DevBuf d_resp_all, d_normal_lut;
  // result blocks of all lanes, contiguous (lane stride result_stride): a sharded caller exchanges the survivors of
  // LM_LANES frames in flight with ONE collective over this region and no staging copies
  DevBuf results_all;
  size_t result_stride = 0;
  uint32_t out_cap = 0, device_out_cap = 2048;
  bool luts_dirty = true;
  Lane lane[LM_LANES];
  Pack pack;
  int shard_rank = 0, shard_world = 1;
  int debug_taps = 0, coarse_variant = 0, refine_variant = 0, timing = 1, frontend_variant = 0, prune = 1, graphs = 1;
  int mod_order = 2;  // coarse kernel: 0 = modalities in template order, 1 = reversed, 2 = chosen per frame (default)
  std::vector<std::string> class_id_cache;
};

// ------------------------------------------------------------------------------------------------ LUT defaults
// SIMILARITY_LUT ([OCV] linemod.cpp): LUT[32*i + 16*h + v] = max over set bits b of v of max(0, 4 - |i - (4h+b)|).
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
// NORMAL_LUT stand-in ([OCV] normal_lut.i is not recoverable): 8 azimuthal sectors of (v1-10, v2-10).
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

static int upload_luts(lm_detector* d) {
  if (!d->luts_dirty) return LM_OK;
  uint32_t resp_all[256];
  for (int v = 0; v < 256; ++v) {
    uint32_t packed = 0;
    for (int ori = 0; ori < 8; ++ori) {
      uint8_t r = std::max(d->sim_lut[32 * ori + (v & 15)], d->sim_lut[32 * ori + 16 + (v >> 4)]);
      packed |= (uint32_t)r << (4 * ori);
    }
    resp_all[v] = packed;
  }
  if (d->d_resp_all.ensure(sizeof(resp_all)) != LM_OK) return LM_E_CUDA;
  if (d->d_normal_lut.ensure(8000) != LM_OK) return LM_E_CUDA;
  CU(cudaMemcpy(d->d_resp_all.p, resp_all, sizeof(resp_all), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d->d_normal_lut.p, d->normal_lut, 8000, cudaMemcpyHostToDevice));
  d->luts_dirty = false;
  for (int i = 0; i < LM_LANES; ++i) d->lane[i].front_valid = false;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ workspace
static size_t src_row_bytes(int type, int cols) { return type == LM_8UC3 ? (size_t)cols * 3 : (type == LM_16UC1 ? (size_t)cols * 2 : (size_t)cols); }
static int expected_src_type(const lm_modality_desc& m) { return m.type == LM_COLOR_GRADIENT ? LM_8UC3 : LM_16UC1; }

// Buffers needed by quantisation at (rows, cols); no divisibility requirements (addTemplate uses this alone).
static int ensure_quant_ws(lm_detector* d, Lane& ln, int rows, int cols) {
  const int L = d->model.levels(), M = d->model.M();
  if (ln.rows != rows || ln.cols != cols) { ln.lm_ready = false; ln.front_valid = false; }
  ln.rows = rows; ln.cols = cols;
  for (int m = 0; m < M; ++m) {
    const bool cg = d->model.mods[m].type == LM_COLOR_GRADIENT;
    size_t n0 = (size_t)rows * cols;
    if (ln.src[m].ensure(src_row_bytes(expected_src_type(d->model.mods[m]), cols) * rows) != LM_OK) return LM_E_CUDA;
    if (cg) {
      if (ln.smoothed[m].ensure(n0 * 3) != LM_OK || ln.qunf[m].ensure(n0) != LM_OK) return LM_E_CUDA;
    } else if (ln.dn_raw[m].ensure(n0) != LM_OK) return LM_E_CUDA;
    for (int l = 0; l < L; ++l) {
      size_t n = (size_t)(rows >> l) * (cols >> l);
      if (n == 0) return fail(LM_E_INVALID, "image too small for %d pyramid levels", L);
      if (cg && l > 0 && ln.bgr[l][m].ensure(n * 3) != LM_OK) return LM_E_CUDA;
      if (cg && ln.mag[l][m].ensure(n * sizeof(float)) != LM_OK) return LM_E_CUDA;
      if (ln.quant_raw[l][m].ensure(n) != LM_OK || ln.quantized[l][m].ensure(n) != LM_OK) return LM_E_CUDA;
    }
  }
  return LM_OK;
}

// Linear-memory buffers for matching at (rows, cols): enforces the reference's CV_Asserts on the geometry.
static int ensure_lm_ws(lm_detector* d, Lane& ln, int rows, int cols) {
  const int L = d->model.levels(), M = d->model.M();
  if (ln.lm_ready && ln.rows == rows && ln.cols == cols && (int)ln.geom.size() == L) return LM_OK;
  std::vector<LevelGeom> geom(L);
  for (int l = 0; l < L; ++l) {
    LevelGeom& g = geom[l];
    g.rows = rows >> l; g.cols = cols >> l; g.T = d->model.T[l];
    if (g.T < 1 || g.T > 16) return fail(LM_E_INVALID, "unsupported T=%d at level %d (1..16)", g.T, l);
    if (g.rows <= 0 || g.cols <= 0) return fail(LM_E_INVALID, "image too small for %d pyramid levels", L);
    if (((size_t)g.rows * g.cols) % 16 != 0)
      return fail(LM_E_INVALID, "(rows * cols) %% 16 != 0 at level %d (%dx%d)", l, g.cols, g.rows);  // computeResponseMaps
    if (g.rows % g.T != 0 || g.cols % g.T != 0)
      return fail(LM_E_INVALID, "rows %% T != 0 or cols %% T != 0 at level %d (%dx%d, T=%d)", l, g.cols, g.rows, g.T);  // linearize
    if (g.cols > 4095 || g.rows > 4095) return fail(LM_E_INVALID, "images larger than 4095 px are not supported");
    g.W = g.cols / g.T; g.H = g.rows / g.T;
    g.plane_stride = plane_stride_of(g.T, g.W, g.H);
  }
  if (ensure_quant_ws(d, ln, rows, cols) != LM_OK) return LM_E_CUDA;
  for (int l = 0; l < L; ++l) {
    size_t bytes = (size_t)M * 8 * geom[l].plane_stride + kLmSlack;
    if (ln.lmem[l].ensure(bytes) != LM_OK) return LM_E_CUDA;
    CU(cudaMemsetAsync(ln.lmem[l].p, 0, ln.lmem[l].cap, ln.stream));  // zero tails (and slack) once per geometry
  }
  for (int l = 0; l < L; ++l) {
    size_t bytes = ((size_t)M * 8 * geom[l].plane_stride + kLmSlack) / 2;
    if (ln.lmn[l].ensure(bytes) != LM_OK) return LM_E_CUDA;
    CU(cudaMemsetAsync(ln.lmn[l].p, 0, ln.lmn[l].cap, ln.stream));
  }
  ln.geom.swap(geom);
  ln.lm_ready = true;
  ln.front_valid = false;
  return LM_OK;
}

static int ensure_tap_ws(lm_detector* d, Lane& ln) {
  const int L = d->model.levels(), M = d->model.M();
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      size_t n = (size_t)ln.geom[l].rows * ln.geom[l].cols;
      if (ln.spread[l][m].ensure(n) != LM_OK || ln.response[l][m].ensure(8 * n) != LM_OK) return LM_E_CUDA;
    }
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ uploads
static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Host image -> tightly packed device buffer.  Pinned sources go straight to the copy engine; pageable ones are
// packed into the lane's pinned staging area first (offset *stage_off, advanced).
static int upload_image(Lane& ln, const lm_image& im, void* dst, size_t* stage_off) {
  const size_t rb = src_row_bytes(im.type, im.cols);
  const size_t total = rb * im.rows;
  if (is_pinned(im.data)) {
    if (im.step == rb) CU(cudaMemcpyAsync(dst, im.data, total, cudaMemcpyHostToDevice, ln.stream));  // one linear DMA
    else CU(cudaMemcpy2DAsync(dst, rb, im.data, im.step, rb, im.rows, cudaMemcpyHostToDevice, ln.stream));
    return LM_OK;
  }
  uint8_t* st = ln.stage_in.as<uint8_t>() + *stage_off;
  if (im.step == rb) std::memcpy(st, im.data, total);
  else
    for (int y = 0; y < im.rows; ++y) std::memcpy(st + (size_t)y * rb, (const uint8_t*)im.data + (size_t)y * im.step, rb);
  CU(cudaMemcpyAsync(dst, st, total, cudaMemcpyHostToDevice, ln.stream));
  *stage_off += (total + 255) & ~(size_t)255;
  return LM_OK;
}

static int check_sources(lm_detector* d, const lm_image* sources, int n_sources, const lm_image* masks, int n_masks) {
  const int M = d->model.M();
  if (n_sources != M) return fail(LM_E_INVALID, "sources.size() (%d) != modalities.size() (%d)", n_sources, M);
  if (n_masks != 0 && n_masks != M) return fail(LM_E_INVALID, "masks.size() (%d) != modalities.size() (%d)", n_masks, M);
  for (int m = 0; m < M; ++m) {
    if (!sources[m].data) return fail(LM_E_INVALID, "source %d is empty", m);
    if (sources[m].type != expected_src_type(d->model.mods[m]))
      return fail(LM_E_INVALID, "source %d: %s needs a %s image", m, modality_name(d->model.mods[m].type),
                  d->model.mods[m].type == LM_COLOR_GRADIENT ? "CV_8UC3" : "CV_16UC1");
    if (sources[m].rows != sources[0].rows || sources[m].cols != sources[0].cols)
      return fail(LM_E_INVALID, "sources differ in size");
    if (sources[m].step < src_row_bytes(sources[m].type, sources[m].cols)) return fail(LM_E_INVALID, "source %d: step too small", m);
    if (n_masks && masks[m].data) {
      if (masks[m].type != LM_8UC1 || masks[m].rows != sources[m].rows || masks[m].cols != sources[m].cols)
        return fail(LM_E_INVALID, "mask %d: size/type mismatch (mask.size() == source.size())", m);
    }
  }
  return LM_OK;
}

static int upload_frame(lm_detector* d, Lane& ln, const lm_image* sources, const lm_image* masks, int n_masks) {
  const int M = d->model.M();
  size_t need = 0;
  for (int m = 0; m < M; ++m) {
    need += ((src_row_bytes(sources[m].type, sources[m].cols) * sources[m].rows) + 255) & ~(size_t)255;
    if (n_masks && masks[m].data) need += (((size_t)masks[m].rows * masks[m].cols) + 255) & ~(size_t)255;
  }
  if (ln.stage_in.ensure(need) != LM_OK) return LM_E_CUDA;
  size_t off = 0;
  for (int m = 0; m < M; ++m) {
    if (upload_image(ln, sources[m], ln.src[m].p, &off) != LM_OK) return LM_E_CUDA;
    ln.src_ptr[m] = ln.src[m].p;
    ln.has_mask[m] = n_masks && masks[m].data;
    if (ln.has_mask[m]) {
      if (ln.mask0[m].ensure((size_t)masks[m].rows * masks[m].cols) != LM_OK) return LM_E_CUDA;
      if (upload_image(ln, masks[m], ln.mask0[m].p, &off) != LM_OK) return LM_E_CUDA;
    }
  }
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ front end
// [OCV] Modality::process + QuantizedPyramid::pyrDown for every level: fills quant_raw[l][m] (and mag[l][m]).
// Stage-by-stage kernels (lm_frontend.cu): the A/B reference of the fused path, selected by option frontend_variant=1.
static int run_quantize_staged(lm_detector* d, Lane& ln, cudaStream_t s) {
  const int L = d->model.levels(), M = d->model.M();
  for (int m = 0; m < M; ++m) {
    const lm_modality_desc& md = d->model.mods[m];
    for (int l = 0; l < L; ++l) {
      const int rows = ln.rows >> l, cols = ln.cols >> l;
      if (md.type == LM_COLOR_GRADIENT) {
        const uint8_t* src = l == 0 ? (const uint8_t*)ln.src_ptr[m] : ln.bgr[l][m].as<uint8_t>();
        if (l > 0) {
          const uint8_t* prev = l == 1 ? (const uint8_t*)ln.src_ptr[m] : ln.bgr[l - 1][m].as<uint8_t>();
          launch_pyrdown_u8c3(prev, ln.rows >> (l - 1), ln.cols >> (l - 1), ln.bgr[l][m].as<uint8_t>(), s);
          ++ln.launches;
        }
        launch_gauss7_u8c3(src, rows, cols, ln.smoothed[m].as<uint8_t>(), s);
        launch_cg_grad(ln.smoothed[m].as<uint8_t>(), rows, cols, ln.mag[l][m].as<float>(), ln.qunf[m].as<uint8_t>(), s);
        launch_cg_hysteresis(ln.qunf[m].as<uint8_t>(), ln.mag[l][m].as<float>(), rows, cols,
                             md.weak_threshold * md.weak_threshold, ln.quant_raw[l][m].as<uint8_t>(), s);
        ln.launches += 3;
      } else {
        if (l == 0) {
          launch_dn_normals((const uint16_t*)ln.src_ptr[m], rows, cols, md.distance_threshold, md.difference_threshold,
                            d->d_normal_lut.as<uint8_t>(), ln.dn_raw[m].as<uint8_t>(), s);
          launch_median5_u8(ln.dn_raw[m].as<uint8_t>(), rows, cols, ln.quant_raw[0][m].as<uint8_t>(), s);
          ln.launches += 2;
        } else {
          launch_nn_half_u8(ln.quant_raw[l - 1][m].as<uint8_t>(), ln.rows >> (l - 1), ln.cols >> (l - 1),
                            ln.quant_raw[l][m].as<uint8_t>(), s);
          ++ln.launches;
        }
      }
    }
  }
  CU(cudaGetLastError());
  return LM_OK;
}

// Production path (lm_frontend_fused.cu): per ColorGradient modality the pyrDown chain plus ONE launch covering every
// level, per DepthNormal modality ONE launch.
static int run_quantize(lm_detector* d, Lane& ln, cudaStream_t main_stream) {
  if (d->frontend_variant == 1) return run_quantize_staged(d, ln, main_stream);
  const int L = d->model.levels(), M = d->model.M();
  if (M > 1) CU(cudaEventRecord(ln.ev_fork, main_stream));
  for (int m = 0; m < M; ++m) {
    const lm_modality_desc& md = d->model.mods[m];
    cudaStream_t s = m == 0 ? main_stream : ln.side[m - 1];
    if (m > 0) CU(cudaStreamWaitEvent(s, ln.ev_fork, 0));
    if (md.type == LM_COLOR_GRADIENT) {
      CgParams cp;
      std::memset(&cp, 0, sizeof(cp));
      cp.n_levels = L;
      cp.thr_sq = md.weak_threshold * md.weak_threshold;
      int total = 0;
      for (int l = 0; l < L; ++l) {
        const int rows = ln.rows >> l, cols = ln.cols >> l;
        if (l > 0) {
          const uint8_t* prev = l == 1 ? (const uint8_t*)ln.src_ptr[m] : ln.bgr[l - 1][m].as<uint8_t>();
          launch_pyrdown_fast(prev, ln.rows >> (l - 1), ln.cols >> (l - 1), ln.bgr[l][m].as<uint8_t>(), s);
          ++ln.launches;
        }
        CgLevel& lv = cp.lv[l];
        lv.src = l == 0 ? (const uint8_t*)ln.src_ptr[m] : ln.bgr[l][m].as<uint8_t>();
        lv.mag = ln.mag[l][m].as<float>();
        lv.quant = ln.quant_raw[l][m].as<uint8_t>();
        lv.rows = rows; lv.cols = cols; lv.block_begin = total;
        total += cg_fused_blocks(rows, cols, &lv.blocks_x);
      }
      launch_cg_fused(cp, total, s);
      ++ln.launches;
    } else {
      DnParams dp;
      std::memset(&dp, 0, sizeof(dp));
      dp.depth = (const uint16_t*)ln.src_ptr[m];
      dp.lut = d->d_normal_lut.as<uint8_t>();
      dp.rows = ln.rows; dp.cols = ln.cols; dp.n_levels = L;
      dp.distance_threshold = md.distance_threshold; dp.difference_threshold = md.difference_threshold;
      for (int l = 0; l < L; ++l) dp.quant[l] = ln.quant_raw[l][m].as<uint8_t>();
      launch_dn_fused(dp, s);
      ++ln.launches;
    }
    if (m > 0) {
      CU(cudaEventRecord(ln.ev_join[m - 1], s));
      CU(cudaStreamWaitEvent(main_stream, ln.ev_join[m - 1], 0));
    }
  }
  CU(cudaGetLastError());
  return LM_OK;
}

// The refinement kernel reads nibble planes when every refinement level has word-aligned rows (refine_variant 0), byte
// planes otherwise.
static bool refine_nibbles(const lm_detector* d, const Lane& ln) {
  if (d->refine_variant != 0) return false;
  for (size_t l = 0; l + 1 < ln.geom.size(); ++l)
    if (!level_nibble_aligned(ln.geom[l])) return false;
  return true;
}

// [OCV] Detector::match front half: quantize (mask) -> spread -> computeResponseMaps -> linearize per level/modality.
static int run_front(lm_detector* d, Lane& ln, cudaStream_t s) {
  const int L = d->model.levels(), M = d->model.M();
  CU(cudaMemsetAsync(ln.mod_bits.p, 0, sizeof(unsigned int) * LM_MAX_MODALITIES, s));
  if (run_quantize(d, ln, s) != LM_OK) return LM_E_CUDA;
  const bool taps = d->debug_taps != 0;
  if (taps && ensure_tap_ws(d, ln) != LM_OK) return LM_E_CUDA;
  // which planes this front end produces per level (staged A/B path: byte planes everywhere, nibbles packed from them)
  bool nib[LM_MAX_LEVELS] = {false, false, false, false}, byt[LM_MAX_LEVELS] = {true, true, true, true};
  if (d->frontend_variant == 1) {
    for (int l = 0; l < L; ++l) {
      const LevelGeom& g = ln.geom[l];
      for (int m = 0; m < M; ++m) {
        launch_spread_lm(ln.quant_raw[l][m].as<uint8_t>(), ln.has_mask[m] ? ln.mask0[m].as<uint8_t>() : nullptr, ln.cols, l,
                         g.rows, g.cols, g.T, d->d_resp_all.as<uint32_t>(), ln.quantized[l][m].as<uint8_t>(),
                         taps ? ln.spread[l][m].as<uint8_t>() : nullptr, taps ? ln.response[l][m].as<uint8_t>() : nullptr,
                         ln.lmem[l].as<uint8_t>() + (size_t)m * 8 * g.plane_stride, g.plane_stride, s);
        ++ln.launches;
      }
    }
  } else {
    SpreadParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.resp_all = d->d_resp_all.as<uint32_t>();
    int total = 0, max_T = 1;
    for (int l = 0; l < L; ++l) {
      const LevelGeom& g = ln.geom[l];
      max_T = std::max(max_T, g.T);
      // nibble planes are written directly when every (orientation, phase) row starts on a word; the byte planes only
      // when something reads them: parity taps, the byte A/B kernels, or a level whose rows are not word-aligned
      nib[l] = l == L - 1 ? level_nibble_aligned(g) : refine_nibbles(d, ln);
      byt[l] = taps || !nib[l] || (l == L - 1 && d->coarse_variant == 1);
      for (int m = 0; m < M; ++m) {
        SpreadEntry& e = sp.e[sp.n++];
        e.qraw = ln.quant_raw[l][m].as<uint8_t>();
        e.mask0 = ln.has_mask[m] ? ln.mask0[m].as<uint8_t>() : nullptr;
        e.quantized = ln.quantized[l][m].as<uint8_t>();
        e.spread = taps ? ln.spread[l][m].as<uint8_t>() : nullptr;
        e.response = taps ? ln.response[l][m].as<uint8_t>() : nullptr;
        e.lm = byt[l] ? ln.lmem[l].as<uint8_t>() + (size_t)m * 8 * g.plane_stride : nullptr;
        e.lm_nib = nib[l] ? ln.lmn[l].as<uint8_t>() + (size_t)m * 4 * g.plane_stride : nullptr;
        e.bits = l == L - 1 ? ln.mod_bits.as<unsigned int>() + m : nullptr;
        e.plane_stride = g.plane_stride;
        e.rows = g.rows; e.cols = g.cols; e.T = g.T; e.W = g.W; e.H = g.H; e.level = l; e.mask_cols0 = ln.cols;
        e.block_begin = total;
        total += spread_all_blocks(g.W, g.H, &e.blocks_x);
      }
    }
    if (!launch_spread_all(sp, total, max_T, s)) return fail(LM_E_INVALID, "T=%d needs too much shared memory", max_T);
    ++ln.launches;
  }
  for (int l = 0; l < L; ++l) {
    // nibble planes the matching kernels will read but the spread kernel could not write directly: pack the byte planes
    const bool wanted = l == L - 1 ? d->coarse_variant != 1 : refine_nibbles(d, ln);
    if (wanted && !nib[l]) {
      launch_pack_nibbles(ln.lmem[l].as<uint8_t>(), ln.lmn[l].as<uint8_t>(), (size_t)M * 8 * ln.geom[l].plane_stride, s);
      ++ln.launches;
      nib[l] = true;
    }
    ln.bytes_valid[l] = byt[l];
    ln.nibbles_valid[l] = nib[l];
  }
  CU(cudaGetLastError());
  ln.front_valid = true;
  ln.debug_taps_written = taps;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ template pack
static int ensure_pack(lm_detector* d, const Lane& ln) {
  Pack& pk = d->pack;
  const HostModel& md = d->model;
  if (pk.version == md.version && pk.rows == ln.rows && pk.cols == ln.cols && pk.shard_rank == d->shard_rank &&
      pk.shard_world == d->shard_world && pk.variant == d->coarse_variant)
    return LM_OK;
  // all lanes must be idle before the shared records are replaced
  for (int i = 0; i < LM_LANES; ++i)
    if (d->lane[i].stream) CU(cudaStreamSynchronize(d->lane[i].stream));
  const int L = md.levels(), M = md.M();
  const LevelGeom& gc = ln.geom[L - 1];
  std::vector<CoarseTpl> ctpl;
  std::vector<uint32_t> foff;
  std::vector<RefineTpl> rtpl[LM_MAX_LEVELS];
  std::vector<uint32_t> rfeats[LM_MAX_LEVELS];
  pk.clear_filtered();
  pk.classes.clear(); pk.coarse_bytes.clear(); pk.refine_nf.clear();
  pk.coarse_bytes_all = 0; pk.max_P = 0;
  uint32_t canonical = 0;
  int class_index = 0;
  for (auto it = md.classes.begin(); it != md.classes.end(); ++it, ++class_index) {
    Pack::ClassRange cr;
    cr.id = it->first; cr.class_index = class_index;
    const std::vector<TemplatePyramid>& tps = it->second;
    for (size_t tid = 0; tid < tps.size(); ++tid, ++canonical) {
      if ((int)(canonical % (uint32_t)d->shard_world) != d->shard_rank) continue;
      const TemplatePyramid& tp = tps[tid];
      if ((int)tp.size() != L * M) return fail(LM_E_INVALID, "template pyramid of class '%s' has %zu templates, expected %d", it->first.c_str(), tp.size(), L * M);
      CoarseTpl ct;
      std::memset(&ct, 0, sizeof(ct));
      ct.feat_begin = (uint32_t)foff.size();
      ct.order_key = canonical; ct.template_id = (int)tid; ct.class_index = class_index;
      const int lowest = (L - 1) * M;
      // [OCV] similarity(): geometry of the sliding window, from each modality's own template
      uint64_t bytes = 0;
      int P_all = -1;
      for (int m = 0; m < M; ++m) {
        const Template& t = tp[lowest + m];
        ct.nf += (uint32_t)t.features.size();
        int wf = (t.width - 1) / gc.T + 1, hf = (t.height - 1) / gc.T + 1;
        int span_x = gc.W - wf, span_y = gc.H - hf;
        int P = span_y * gc.W + span_x + 1;
        if (P > gc.W * gc.H) P = gc.W * gc.H;
        // cropTemplates gives every modality the same width/height, so P is shared; a hand-made pyramid that
        // violates this is rejected rather than silently mis-scored.
        if (m == 0) P_all = P;
        else if (P != P_all) return fail(LM_E_INVALID, "class '%s' template %zu: modalities disagree on width/height", it->first.c_str(), tid);
        std::vector<uint32_t> grp[4];
        for (const Feature& f : t.features) {
          if (f.x < 0 || f.x >= gc.cols || f.y < 0 || f.y >= gc.rows) continue;  // "Discard feature if out of bounds"
          size_t a = (size_t)m * 8 * gc.plane_stride + (size_t)f.label * gc.plane_stride +
                     (size_t)((f.y % gc.T) * gc.T + (f.x % gc.T)) * ((size_t)gc.W * gc.H) + (size_t)(f.y / gc.T) * gc.W + f.x / gc.T;
          // window-alignment class of the feature (a compile-time constant in the kernels): byte planes -> word
          // offset in the 16-byte chunk; nibble planes -> the same with a = nibble index
          const int q = d->coarse_variant == 1 ? (int)((a & 15) >> 2) : (int)((a >> 3) & 3);
          grp[q].push_back((uint32_t)a);
        }
        for (int q = 0; q < 4; ++q) {
          ct.cnt[m][q] = (uint8_t)grp[q].size();
          foff.insert(foff.end(), grp[q].begin(), grp[q].end());
          if (P > 0) bytes += (uint64_t)grp[q].size() * (uint64_t)P;
        }
      }
      ct.P = P_all;
      pk.max_P = std::max(pk.max_P, ct.P);
      uint32_t rnf = 0;
      for (int l = 0; l < L - 1; ++l) {
        RefineTpl rt;
        std::memset(&rt, 0, sizeof(rt));
        rt.feat_begin = (uint32_t)rfeats[l].size();
        rt.width = tp[l * M].width; rt.height = tp[l * M].height;
        for (int m = 0; m < M; ++m) {
          const Template& t = tp[l * M + m];
          rt.cnt[m] = (uint16_t)t.features.size();
          rt.nf += (uint32_t)t.features.size();
          for (const Feature& f : t.features)
            rfeats[l].push_back((uint32_t)(f.x + 4096) | ((uint32_t)(f.y + 4096) << 13) | ((uint32_t)f.label << 26));
        }
        rnf += rt.nf;
        rtpl[l].push_back(rt);
      }
      cr.local.push_back((uint32_t)ctpl.size());
      cr.global_pos.push_back(canonical);
      ctpl.push_back(ct);
      pk.coarse_bytes.push_back(bytes);
      pk.coarse_bytes_all += bytes;
      pk.refine_nf.push_back(rnf);
    }
    pk.classes.push_back(cr);
  }
  pk.n = (int)ctpl.size();
  auto up = [&](DevBuf& b, const void* src, size_t bytes) -> int {
    if (b.ensure(bytes + 64) != LM_OK) return LM_E_CUDA;
    if (bytes) CU(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return LM_OK;
  };
  foff.resize(foff.size() + 8, 0);  // the 4-way unrolled loop never reads past the template, padding is for safety
  if (up(pk.ctpl, ctpl.data(), ctpl.size() * sizeof(CoarseTpl)) != LM_OK) return LM_E_CUDA;
  if (up(pk.foff, foff.data(), foff.size() * 4) != LM_OK) return LM_E_CUDA;
  for (int l = 0; l < L - 1; ++l) {
    if (up(pk.rtpl[l], rtpl[l].data(), rtpl[l].size() * sizeof(RefineTpl)) != LM_OK) return LM_E_CUDA;
    if (up(pk.rfeats[l], rfeats[l].data(), rfeats[l].size() * 4) != LM_OK) return LM_E_CUDA;
  }
  pk.h_ctpl.swap(ctpl);
  pk.h_foff.swap(foff);
  pk.version = md.version; pk.rows = ln.rows; pk.cols = ln.cols;
  pk.shard_rank = d->shard_rank; pk.shard_world = d->shard_world; pk.variant = d->coarse_variant;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ matching
struct Query {
  float threshold;
  const char* const* class_ids;
  int n_ids;
};

static const size_t kFirstChunkRecords = 1024;  // records fetched together with the header in one D2H copy
static const int kMaxQueries = LM_MAX_QUERIES;  // (class list, threshold) queries answered from one front end

// Self-contained tile records of the production coarse kernel (layout: lm_kernels.cuh).
static int build_tile_records(const Pack& pk, const std::vector<WorkItem>& items, const std::vector<uint2>& tiles,
                              int pass_pos, int M, std::vector<uint32_t>& recs, int* rec_words) {
  const int hdr = coarse_record_header_words();
  int max_feat = 0;
  for (const WorkItem& it : items) {
    const CoarseTpl& ct = pk.h_ctpl[it.tglob];
    int n = 0;
    for (int m = 0; m < M; ++m) for (int g = 0; g < 4; ++g) n += ct.cnt[m][g];
    max_feat = std::max(max_feat, n);
  }
  const int words = (hdr + max_feat + 3) & ~3;
  if (words > coarse_record_max_words()) return fail(LM_E_INVALID, "template with too many features for a tile record");
  *rec_words = words;
  recs.assign((size_t)words * tiles.size(), 0u);
  for (size_t t = 0; t < tiles.size(); ++t) {
    uint32_t* r = &recs[t * (size_t)words];
    const WorkItem& it = items[tiles[t].x];
    const CoarseTpl& ct = pk.h_ctpl[it.tglob];
    const int j0 = (int)tiles[t].y * pass_pos;
    int n = 0;
    for (int m = 0; m < M; ++m) {
      uint32_t c4 = 0;
      for (int g = 0; g < 4; ++g) { c4 |= (uint32_t)ct.cnt[m][g] << (8 * g); n += ct.cnt[m][g]; }
      r[8 + m] = c4;
    }
    r[0] = tiles[t].x; r[1] = it.tglob; r[2] = (ct.nf & 0x0fffffffu) | (it.order & 0xf0000000u); r[3] = (uint32_t)n;
    r[4] = (uint32_t)j0; r[5] = (uint32_t)std::min(pass_pos, ct.P - j0); r[6] = it.order;
    for (int f = 0; f < n; ++f) {
      const uint32_t a = pk.h_foff[ct.feat_begin + f] + (uint32_t)j0;  // nibble index of lane 0's window
      r[hdr + f] = ((a >> 1) & ~15u) | (a & 7u);
    }
  }
  return LM_OK;
}

// [OCV] Detector::match: "if (class_ids.empty()) match all templates else only the requested class IDs" (in the
// order requested, unknown ids skipped) -- for every query of the request.  The plan lists the work items (template
// + emission order key tagged with the query index) and the non-empty 512-position tiles of the coarse kernel,
// heaviest first; it depends on the class lists only, so it is built once per distinct request and cached.
static int get_plan(lm_detector* d, const Query* qs, int n_q, Pack::Plan** out) {
  Pack& pk = d->pack;
  std::string key;
  for (int q = 0; q < n_q; ++q) {
    for (int i = 0; i < qs[q].n_ids; ++i) {
      if (!qs[q].class_ids[i]) return fail(LM_E_INVALID, "class_ids[%d] is NULL", i);
      key += qs[q].class_ids[i];
      key += '\n';
    }
    key += '\x01';
  }
  auto hit = pk.plans.find(key);
  if (hit != pk.plans.end()) { *out = &hit->second; return LM_OK; }
  std::vector<WorkItem> items;
  struct Tile { uint2 t; uint64_t cost; };
  std::vector<Tile> tiles;
  Pack::Plan plan;
  const int pass_pos = coarse_positions_per_pass(d->coarse_variant);
  auto add_class = [&](const Pack::ClassRange& cr, uint32_t base, uint32_t first_global, int q) {
    for (size_t k = 0; k < cr.local.size(); ++k) {
      const uint32_t local = cr.local[k];
      WorkItem it;
      it.tglob = local;
      it.order = (base + (cr.global_pos[k] - first_global)) | ((uint32_t)q << 28);
      const CoarseTpl& ct = pk.h_ctpl[local];
      uint64_t nfeat = 0;
      for (int m = 0; m < LM_MAX_MODALITIES; ++m) for (int g = 0; g < 4; ++g) nfeat += ct.cnt[m][g];
      for (int pass = 0; pass * pass_pos < ct.P; ++pass) {
        Tile t;
        t.t.x = (uint32_t)items.size(); t.t.y = (uint32_t)pass;
        t.cost = nfeat * (uint64_t)std::min(pass_pos, ct.P - pass * pass_pos);
        tiles.push_back(t);
      }
      items.push_back(it);
      plan.coarse_bytes += pk.coarse_bytes[local];
      plan.refine_nf_sum += pk.refine_nf[local];
    }
  };
  for (int q = 0; q < n_q; ++q) {
    if (qs[q].n_ids == 0) {
      for (const Pack::ClassRange& cr : pk.classes) add_class(cr, 0, 0, q);  // order key = canonical index
    } else {
      uint32_t base = 0;  // position in the filtered iteration (a class may be listed more than once)
      for (int i = 0; i < qs[q].n_ids; ++i) {
        auto it = d->model.classes.find(qs[q].class_ids[i]);
        if (it == d->model.classes.end()) continue;
        uint32_t first_global = 0;
        for (auto jt = d->model.classes.begin(); jt != it; ++jt) first_global += (uint32_t)jt->second.size();
        for (const Pack::ClassRange& cr : pk.classes)
          if (cr.id == qs[q].class_ids[i]) add_class(cr, base, first_global, q);
        base += (uint32_t)it->second.size();
      }
    }
  }
  if (items.size() >= (1u << 28)) return fail(LM_E_INVALID, "too many templates in one request");
  std::stable_sort(tiles.begin(), tiles.end(), [](const Tile& a, const Tile& b) { return a.cost > b.cost; });
  std::vector<uint2> tl(tiles.size());
  for (size_t i = 0; i < tiles.size(); ++i) tl[i] = tiles[i].t;
  plan.n_items = (int)items.size();
  plan.n_tiles = (int)tl.size();
  plan.evals = (uint64_t)items.size();
  std::vector<uint32_t> recs;
  if (d->coarse_variant == 0) {
    if (build_tile_records(pk, items, tl, pass_pos, d->model.M(), recs, &plan.rec_words) != LM_OK) return LM_E_INVALID;
  }
  Pack::Plan& dst = pk.plans[key];
  dst = plan;
  if (!recs.empty()) {
    if (dst.recs.ensure(recs.size() * 4 + 64) != LM_OK) return LM_E_CUDA;
    CU(cudaMemcpy(dst.recs.p, recs.data(), recs.size() * 4, cudaMemcpyHostToDevice));
  }
  if (dst.items.ensure(items.size() * sizeof(WorkItem) + 64) != LM_OK || dst.tiles.ensure(tl.size() * sizeof(uint2) + 64) != LM_OK) return LM_E_CUDA;
  if (!items.empty()) CU(cudaMemcpy(dst.items.p, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice));
  if (!tl.empty()) CU(cudaMemcpy(dst.tiles.p, tl.data(), tl.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  *out = &dst;
  return LM_OK;
}

// Device pointer of the plan's tile records (variant 0) or null.
static const uint32_t* plan_recs(const Pack::Plan& plan) { return plan.recs.as<uint32_t>(); }

// Result block in device memory: [16 B kernel statistics][ResultHeader][out_cap x lm_raw_match].  The statistics
// (u64 (feature, position) pairs the coarse kernel gathered) sit in front so that one memset clears both and one D2H
// copy brings both; the public block (lm_match_device) starts at the header.
static const size_t kStatsBytes = 16;
static size_t result_bytes(const Lane& ln) { return sizeof(ResultHeader) + (size_t)ln.out_cap * sizeof(lm_raw_match); }
static uint8_t* block_ptr(const Lane& ln) { return ln.result.as<uint8_t>() + kStatsBytes; }

static int ensure_match_buffers(lm_detector* d, Lane& ln, uint32_t cand_cap, uint32_t out_cap) {
  if (cand_cap > ln.cand_cap) {
    if (ln.cand.ensure((size_t)cand_cap * sizeof(Cand)) != LM_OK) return LM_E_CUDA;
    ln.cand_cap = cand_cap;
  }
  if (out_cap > d->out_cap || d->results_all.p == nullptr) {
    // every lane's block moves: nothing may be in flight (callers' streams included); blocks whose results have not
    // been downloaded yet (other lanes of a batch) keep their contents
    CU(cudaDeviceSynchronize());
    const size_t old_stride = d->result_stride;
    DevBuf old = d->results_all;
    d->results_all = DevBuf();
    d->out_cap = std::max(out_cap, d->out_cap);
    d->result_stride = (kStatsBytes + sizeof(ResultHeader) + (size_t)d->out_cap * sizeof(lm_raw_match) + 255) & ~(size_t)255;
    if (d->results_all.ensure(d->result_stride * LM_LANES) != LM_OK) { d->results_all = old; return LM_E_CUDA; }
    CU(cudaMemset(d->results_all.p, 0, d->results_all.cap));
    for (int i = 0; i < LM_LANES; ++i) {
      uint8_t* fresh = d->results_all.as<uint8_t>() + (size_t)i * d->result_stride;
      if (old.p) CU(cudaMemcpy(fresh, old.as<uint8_t>() + (size_t)i * old_stride, std::min(old_stride, d->result_stride), cudaMemcpyDeviceToDevice));
      d->lane[i].result.p = fresh;
      d->lane[i].out_cap = d->out_cap;
    }
    old.release();
  }
  if (ln.stage_out.ensure(kStatsBytes + result_bytes(ln)) != LM_OK) return LM_E_CUDA;
  return LM_OK;
}

// Enqueues the request's coarse similarity + refinement on stream s: one launch each, whatever the number of queries.
static int enqueue_match(lm_detector* d, Lane& ln, const Pack::Plan& plan, const Query* qs, int n_q, cudaStream_t s,
                         cudaEvent_t ev_mid) {
  const HostModel& md = d->model;
  const int L = md.levels(), M = md.M();
  Pack& pk = d->pack;
  const LevelGeom& gc = ln.geom[L - 1];
  ResultHeader* d_hdr = reinterpret_cast<ResultHeader*>(block_ptr(ln));
  lm_raw_match* d_out = reinterpret_cast<lm_raw_match*>(block_ptr(ln) + sizeof(ResultHeader));
  CU(cudaMemsetAsync(ln.result.p, 0, kStatsBytes + sizeof(ResultHeader), s));
  QueryThresholds qt;
  RefineParams rp;
  std::memset(&rp, 0, sizeof(rp));
  std::memset(&qt, 0, sizeof(qt));
  for (int q = 0; q < n_q; ++q) { qt.v[q] = qs[q].threshold; rp.threshold[q] = qs[q].threshold; }
  launch_similarity_coarse(d->coarse_variant, ln.lmem[L - 1].as<uint8_t>(), ln.lmn[d->model.levels() - 1].as<uint8_t>(), pk.foff.as<uint32_t>(),
                           pk.ctpl.as<CoarseTpl>(), plan.items.as<WorkItem>(), plan.tiles.as<uint2>(), plan_recs(plan), plan.rec_words,
                           plan.n_tiles, qt, M, (d->prune & 1) | (d->mod_order << 8), ln.cand.as<Cand>(), d_hdr,
                           ln.result.as<unsigned long long>(), ln.cand_cap, nullptr, 0, s,
                           d->frontend_variant == 0 ? ln.mod_bits.as<unsigned int>() : nullptr);
  if (plan.n_tiles > 0) ++ln.launches;
  if (ev_mid) CU(cudaEventRecord(ev_mid, s));
  rp.levels = L; rp.M = M; rp.coarse_T = gc.T; rp.coarse_W = gc.W;
  for (int l = 0; l < L - 1; ++l) {
    const LevelGeom& g = ln.geom[l];
    rp.level[l].lm = ln.lmem[l].as<uint8_t>();
    rp.level[l].lmn = ln.lmn[l].as<uint8_t>();
    rp.level[l].tpl = pk.rtpl[l].as<RefineTpl>();
    rp.level[l].feats = pk.rfeats[l].as<uint32_t>();
    rp.level[l].plane_stride = g.plane_stride;
    rp.level[l].rows = g.rows; rp.level[l].cols = g.cols; rp.level[l].T = g.T; rp.level[l].W = g.W;
  }
  launch_refine(refine_nibbles(d, ln), rp, pk.ctpl.as<CoarseTpl>(), plan.items.as<WorkItem>(), ln.cand.as<Cand>(), ln.cand_cap,
                d_hdr, d_out, ln.out_cap, s);
  ++ln.launches;
  CU(cudaGetLastError());
  return LM_OK;
}

// Front end + matching of one frame whose sources are already in device memory (ln.src_ptr), enqueued on s.  Replays
// the lane's CUDA graph when one matching this request exists, records a new one otherwise; falls back to plain
// launches when graphs are switched off, the parity taps are on, or a capture ever failed on this lane.
static int enqueue_frame(lm_detector* d, Lane& ln, const Pack::Plan& plan, const Query* qs, int n_q, cudaStream_t s) {
  bool masks = false;
  for (int m = 0; m < d->model.M(); ++m) masks = masks || ln.has_mask[m];
  if (!d->graphs || ln.graph_broken || d->debug_taps || masks) {
    if (run_front(d, ln, s) != LM_OK) return LM_E_CUDA;
    return enqueue_match(d, ln, plan, qs, n_q, s, nullptr);
  }
  Lane::GraphKey key;
  std::memset(&key, 0, sizeof(key));
  key.plan = &plan; key.plan_recs = plan.recs.p; key.n_tiles = plan.n_tiles; key.cand = ln.cand.p; key.result = ln.result.p;
  key.shard_rank = d->shard_rank; key.shard_world = d->shard_world;
  for (int m = 0; m < d->model.M(); ++m) key.src[m] = ln.src_ptr[m];
  key.model_version = d->model.version; key.rows = ln.rows; key.cols = ln.cols; key.n_q = n_q;
  key.variant = d->coarse_variant + 16 * d->refine_variant; key.prune = d->prune | (d->mod_order << 8); key.frontend = d->frontend_variant;
  key.cand_cap = ln.cand_cap; key.out_cap = ln.out_cap;
  for (int q = 0; q < n_q; ++q) key.thr[q] = qs[q].threshold;
  if (ln.gexec == nullptr || std::memcmp(&key, &ln.gkey, sizeof(key)) != 0) {
    if (ln.gexec) { cudaGraphExecDestroy(ln.gexec); ln.gexec = nullptr; }
    cudaGraph_t graph = nullptr;
    set_programmatic_launch(false);
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    int rc = LM_OK;
    if (e == cudaSuccess) {
      ln.launches = 0;
      rc = run_front(d, ln, s);
      if (rc == LM_OK) rc = enqueue_match(d, ln, plan, qs, n_q, s, nullptr);
      e = cudaStreamEndCapture(s, &graph);
    }
    set_programmatic_launch(true);
    if (e == cudaSuccess && rc == LM_OK && graph != nullptr) e = cudaGraphInstantiate(&ln.gexec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc != LM_OK || ln.gexec == nullptr) {  // not fatal: this lane keeps to plain launches
      cudaGetLastError();
      ln.gexec = nullptr;
      ln.graph_broken = true;
      if (run_front(d, ln, s) != LM_OK) return LM_E_CUDA;
      return enqueue_match(d, ln, plan, qs, n_q, s, nullptr);
    }
    ln.gkey = key;
    ln.graph_launches = ln.launches;
  }
  CU(cudaGraphLaunch(ln.gexec, s));
  ln.launches = ln.graph_launches;
  ln.front_valid = true;
  ln.debug_taps_written = false;
  return LM_OK;
}

// [OCV] Match::operator< and operator==
static inline bool match_less(const lm_match_rec& a, const lm_match_rec& b) {
  if (a.similarity != b.similarity) return a.similarity > b.similarity;
  return a.template_id < b.template_id;
}
static inline bool match_equal(const lm_match_rec& a, const lm_match_rec& b) {
  return a.x == b.x && a.y == b.y && a.similarity == b.similarity && a.class_index == b.class_index;
}

// Raw survivor records -> the reference's match list: restore matchClass's emission order (class iteration order,
// template_id, coarse raster position), convert scores to percentages in f32 exactly as the reference does, then the
// same libstdc++ std::sort + std::unique ([OCV] Detector::match tail).
static void finalize_records(int levels, std::vector<lm_raw_match>& raw, std::vector<lm_match_rec>& presort,
                             std::vector<lm_match_rec>& out) {
  std::sort(raw.begin(), raw.end(), [](const lm_raw_match& a, const lm_raw_match& b) {
    return a.order_key != b.order_key ? a.order_key < b.order_key : a.coarse_pos < b.coarse_pos;
  });
  presort.resize(raw.size());
  for (size_t i = 0; i < raw.size(); ++i) {
    const lm_raw_match& r = raw[i];
    lm_match_rec m;
    m.x = r.x; m.y = r.y; m.template_id = r.template_id; m.class_index = r.class_index;
    float sim = ((int)r.score * 100.f) / (4 * (int)r.nf);
    if (levels == 1) sim = sim + 0.5f;  // the coarse score carries +0.5f, refined scores do not (App. D-3)
    m.similarity = sim;
    presort[i] = m;
  }
  out = presort;
  std::sort(out.begin(), out.end(), match_less);
  out.erase(std::unique(out.begin(), out.end(), match_equal), out.end());
}

// One D2H copy brings the header + the first records; long lists need a second copy.
// Result download, part 1: header + leading records into the lane's pinned staging block.  The pipelined paths enqueue it
// right behind the frame's kernels, so that by the time the host comes back to this lane the records are already there.
static int enqueue_download(Lane& ln, cudaStream_t s) {
  const size_t first = std::min<size_t>(kFirstChunkRecords, ln.out_cap);
  CU(cudaMemcpyAsync(ln.stage_out.p, ln.result.p, kStatsBytes + sizeof(ResultHeader) + first * sizeof(lm_raw_match), cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(ln.ev[5], s));
  return LM_OK;
}

static int download_records(Lane& ln, cudaStream_t s, std::vector<lm_raw_match>& raw, bool* overflow, uint32_t* n_cands,
                            bool pre_enqueued = false) {
  const size_t first = std::min<size_t>(kFirstChunkRecords, ln.out_cap);
  uint8_t* host = ln.stage_out.as<uint8_t>();
  if (!pre_enqueued && enqueue_download(ln, s) != LM_OK) return LM_E_CUDA;
  CU(cudaEventSynchronize(ln.ev[5]));
  ln.work_stats[6] = *reinterpret_cast<const unsigned long long*>(host);
  host += kStatsBytes;
  ResultHeader h = *reinterpret_cast<ResultHeader*>(host);
  *overflow = h.overflow != 0 || h.count > ln.out_cap;
  *n_cands = h.n_cands;
  if (*overflow) return LM_OK;
  if (h.count > first) {
    CU(cudaMemcpyAsync(host + sizeof(ResultHeader) + first * sizeof(lm_raw_match),
                       block_ptr(ln) + sizeof(ResultHeader) + first * sizeof(lm_raw_match),
                       (h.count - first) * sizeof(lm_raw_match), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  const lm_raw_match* recs = reinterpret_cast<const lm_raw_match*>(host + sizeof(ResultHeader));
  raw.assign(recs, recs + h.count);
  return LM_OK;
}

static void collect_timings(Lane& ln) {
  for (int i = 0; i < 5; ++i) {
    float t = 0;
    if (cudaEventElapsedTime(&t, ln.ev[i], ln.ev[i + 1]) != cudaSuccess) { cudaGetLastError(); t = 0; }
    ln.ms[i] = t;
  }
}

// Splits the request's survivors by query tag (order_key >> 28) and finalises each query's list.
static void finalize_queries(lm_detector* d, Lane& ln, std::vector<lm_raw_match>& raw, int n_q, std::vector<lm_match_rec>* out) {
  std::vector<lm_raw_match> part;
  for (int q = 0; q < n_q; ++q) {
    part.clear();
    for (const lm_raw_match& r : raw)
      if ((int)(r.order_key >> 28) == q) part.push_back(r);
    finalize_records(d->model.levels(), part, ln.presort, out[q]);
  }
}

// Matching on an already-built front end; buffers grow and the request is re-run on overflow (exactness over speed).
static int match_front(lm_detector* d, Lane& ln, const Query* queries, int n_q, std::vector<lm_match_rec>* out) {
  if (n_q < 1 || n_q > kMaxQueries) return fail(LM_E_INVALID, "number of queries must be 1..%d", kMaxQueries);
  int rc = ensure_pack(d, ln);
  if (rc != LM_OK) return rc;
  Pack::Plan* plan = nullptr;
  rc = get_plan(d, queries, n_q, &plan);
  if (rc != LM_OK) return rc;
  uint32_t cand_cap = std::max<uint32_t>(ln.cand_cap, 1u << 16), out_cap = std::max<uint32_t>(ln.out_cap, 1u << 14);
  std::vector<lm_raw_match> raw;
  uint32_t n_cands = 0;
  for (int attempt = 0;; ++attempt) {
    if (ensure_match_buffers(d, ln, cand_cap, out_cap) != LM_OK) return LM_E_CUDA;
    if (attempt > 0) CU(cudaEventRecord(ln.ev[2], ln.stream));
    if (enqueue_match(d, ln, *plan, queries, n_q, ln.stream, ln.ev[3]) != LM_OK) return LM_E_CUDA;
    CU(cudaEventRecord(ln.ev[4], ln.stream));
    bool overflow = false;
    if (download_records(ln, ln.stream, raw, &overflow, &n_cands) != LM_OK) return LM_E_CUDA;
    if (!overflow) break;
    if (attempt >= 8) return fail(LM_E_CUDA, "match buffers overflowed repeatedly");
    if (n_cands > ln.cand_cap) cand_cap = std::max<uint32_t>(n_cands + n_cands / 4, cand_cap * 2);
    out_cap = std::max<uint32_t>(out_cap * 4, std::min<uint32_t>(n_cands + 1024, 1u << 26));
  }
  // work accounting: B_coarse from the plan; B_refine = candidates x mean refine features x 256 bytes
  ln.work_stats[1] = plan->coarse_bytes;
  if (d->coarse_variant != 0) ln.work_stats[6] = plan->coarse_bytes;  // the A/B kernels always gather everything
  ln.work_stats[4] = n_cands;
  ln.work_stats[5] = (uint64_t)plan->n_items * (uint64_t)(ln.geom.back().W * ln.geom.back().H);
  ln.work_stats[2] = plan->n_items ? (uint64_t)((double)n_cands * (plan->refine_nf_sum / plan->n_items) * 256.0) : 0;
  ln.work_stats[3] = 20ull * raw.size();
  finalize_queries(d, ln, raw, n_q, out);
  return LM_OK;
}

static int copy_out(const std::vector<lm_match_rec>& v, lm_match_rec** out_matches, size_t* out_n) {
  lm_match_rec* p = (lm_match_rec*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(lm_match_rec));
  if (!p) return fail(LM_E_INVALID, "out of host memory");
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(lm_match_rec));
  *out_matches = p; *out_n = v.size();
  return LM_OK;
}

static int validate_template(const HostModel& md, int n_templates, const lm_template_hdr* hdr, const int32_t* feats) {
  if (n_templates != md.levels() * md.M()) return fail(LM_E_INVALID, "template pyramid has %d templates, expected levels*modalities = %d", n_templates, md.levels() * md.M());
  size_t k = 0;
  for (int i = 0; i < n_templates; ++i) {
    if (hdr[i].num_features < 0 || hdr[i].num_features > LM_MAX_FEATURES) return fail(LM_E_INVALID, "features.size() <= 63 violated (%d)", hdr[i].num_features);
    for (int j = 0; j < hdr[i].num_features; ++j, ++k) {
      int x = feats[3 * k], y = feats[3 * k + 1], label = feats[3 * k + 2];
      if (label < 0 || label > 7) return fail(LM_E_INVALID, "feature label %d outside 0..7", label);
      if (x < -4096 || x > 4095 || y < -4096 || y > 4095) return fail(LM_E_INVALID, "feature coordinate outside +-4095");
    }
  }
  return LM_OK;
}

// Binds the handle to the current CUDA device on first use by a compute entry point.  Host-only calls (persistence,
// template bookkeeping, lm_finalize_raw) never get here; everything that touches pixels does, and fails loudly when
// no device is usable -- there is no CPU path.
static int set_device(lm_detector* d) {
  if (!d->cuda_ready) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      return fail(LM_E_CUDA, "no CUDA device available (%s); this library has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    CU(cudaGetDevice(&d->device));
    for (int i = 0; i < LM_LANES; ++i)
      if (d->lane[i].init() != LM_OK) return LM_E_CUDA;
    d->cuda_ready = true;
  }
  CU(cudaSetDevice(d->device));
  return LM_OK;
}

static int create_common(lm_detector* d) {
  default_similarity_lut(d->sim_lut);
  default_normal_lut(d->normal_lut);
  d->luts_dirty = true;
  return LM_OK;
}

static void refresh_class_cache(lm_detector* d) {
  d->class_id_cache.clear();
  for (auto& kv : d->model.classes) d->class_id_cache.push_back(kv.first);
}

// ================================================================================================ C ABI
extern "C" {

const char* lm_last_error(void) { return g_err.c_str(); }

void* lm_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void lm_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int lm_create(const int32_t* T, int levels, const lm_modality_desc* mods, int M, lm_detector** out) {
  if (!out) return fail(LM_E_INVALID, "out is NULL");
  *out = nullptr;
  if (levels < 1 || levels > LM_MAX_LEVELS || !T) return fail(LM_E_INVALID, "pyramid levels must be 1..%d", LM_MAX_LEVELS);
  if (M < 1 || M > LM_MAX_MODALITIES || !mods) return fail(LM_E_INVALID, "modalities must be 1..%d", LM_MAX_MODALITIES);
  for (int m = 0; m < M; ++m) {
    if (mods[m].type != LM_COLOR_GRADIENT && mods[m].type != LM_DEPTH_NORMAL) return fail(LM_E_INVALID, "unknown modality type %d", mods[m].type);
    if (mods[m].num_features < 1 || mods[m].num_features > LM_MAX_FEATURES) return fail(LM_E_INVALID, "num_features must be 1..63");
  }
  lm_detector* d = new lm_detector();
  d->model.T.assign(T, T + levels);
  d->model.mods.assign(mods, mods + M);
  int rc = create_common(d);
  if (rc != LM_OK) { delete d; return rc; }
  *out = d;
  return LM_OK;
}

int lm_create_from_yaml(const char* path, lm_detector** out) {
  if (!out || !path) return fail(LM_E_INVALID, "NULL argument");
  *out = nullptr;
  lm_detector* d = new lm_detector();
  std::string err;
  if (!load_detector_yaml(path, d->model, err)) { delete d; return fail(LM_E_IO, "%s", err.c_str()); }
  for (auto& kv : d->model.classes)
    for (auto& tp : kv.second) {
      if ((int)tp.size() != d->model.levels() * d->model.M()) { delete d; return fail(LM_E_IO, "%s: class '%s' has a template pyramid of the wrong size", path, kv.first.c_str()); }
      for (auto& t : tp) if (t.features.size() > LM_MAX_FEATURES) { delete d; return fail(LM_E_IO, "%s: features.size() <= 63 violated", path); }
    }
  int rc = create_common(d);
  if (rc != LM_OK) { delete d; return rc; }
  refresh_class_cache(d);
  *out = d;
  return LM_OK;
}

int lm_create_from_cache(const char* path, lm_detector** out) {
  if (!out || !path) return fail(LM_E_INVALID, "NULL argument");
  *out = nullptr;
  lm_detector* d = new lm_detector();
  std::string err;
  if (!load_model_cache(path, d->model, err)) { delete d; return fail(LM_E_IO, "%s", err.c_str()); }
  int rc = create_common(d);
  if (rc != LM_OK) { delete d; return rc; }
  refresh_class_cache(d);
  *out = d;
  return LM_OK;
}

int lm_write_cache(const lm_detector* d, const char* path) {
  if (!d || !path) return fail(LM_E_INVALID, "NULL argument");
  std::string err;
  if (!save_model_cache(d->model, path, err)) return fail(LM_E_IO, "%s", err.c_str());
  return LM_OK;
}

int lm_write_yaml(const lm_detector* d, const char* path) {
  if (!d || !path) return fail(LM_E_INVALID, "NULL argument");
  std::string err;
  if (!save_detector_yaml(d->model, path, err)) return fail(LM_E_IO, "%s", err.c_str());
  return LM_OK;
}

static std::string format_name(const char* format, const std::string& id) {
  char buf[4096];
  snprintf(buf, sizeof(buf), format, id.c_str());
  return buf;
}

int lm_read_classes(lm_detector* d, const char* const* class_ids, int n_ids, const char* format) {
  if (!d || (n_ids && !class_ids)) return fail(LM_E_INVALID, "NULL argument");
  const char* fmt = format ? format : "templates_%s.yml.gz";
  for (int i = 0; i < n_ids; ++i) {
    std::string err;
    if (!load_class_file(format_name(fmt, class_ids[i]), d->model, err)) return fail(LM_E_IO, "%s", err.c_str());
  }
  refresh_class_cache(d);
  return LM_OK;
}

int lm_write_classes(const lm_detector* d, const char* format) {
  if (!d) return fail(LM_E_INVALID, "NULL argument");
  const char* fmt = format ? format : "templates_%s.yml.gz";
  for (auto& kv : d->model.classes) {
    std::string err;
    if (!save_class_file(d->model, kv.first, format_name(fmt, kv.first), err)) return fail(LM_E_IO, "%s", err.c_str());
  }
  return LM_OK;
}

void lm_destroy(lm_detector* d) {
  if (!d) return;
  if (d->cuda_ready) {
    cudaSetDevice(d->device);
    for (int i = 0; i < LM_LANES; ++i) {
      if (d->lane[i].stream) cudaStreamSynchronize(d->lane[i].stream);
      d->lane[i].destroy();
    }
    d->pack.release();
    d->train.release();
    d->results_all.release();
    d->d_resp_all.release(); d->d_normal_lut.release();
  }
  delete d;
}

int lm_device(const lm_detector* d) { return d ? d->device : -1; }
int lm_pyramid_levels(const lm_detector* d) { return d->model.levels(); }
int lm_get_T(const lm_detector* d, int level) {
  if (level < 0 || level >= d->model.levels()) return fail(LM_E_INVALID, "level out of range");
  return d->model.T[level];
}
int lm_num_modalities(const lm_detector* d) { return d->model.M(); }
int lm_get_modality(const lm_detector* d, int m, lm_modality_desc* out) {
  if (m < 0 || m >= d->model.M() || !out) return fail(LM_E_INVALID, "modality out of range");
  *out = d->model.mods[m];
  return LM_OK;
}
int lm_num_classes(const lm_detector* d) { return (int)d->model.classes.size(); }
int lm_num_templates(const lm_detector* d, const char* class_id) {
  int n = 0;
  if (class_id) {
    auto it = d->model.classes.find(class_id);
    return it == d->model.classes.end() ? 0 : (int)it->second.size();
  }
  for (auto& kv : d->model.classes) n += (int)kv.second.size();
  return n;
}
const char* lm_class_id(const lm_detector* d, int class_index) {
  if (d->class_id_cache.size() != d->model.classes.size()) refresh_class_cache(const_cast<lm_detector*>(d));
  if (class_index < 0 || class_index >= (int)d->class_id_cache.size()) return nullptr;
  return d->class_id_cache[class_index].c_str();
}

int lm_get_templates(const lm_detector* d, const char* class_id, int template_id, lm_template_hdr* hdr, int32_t* feats) {
  if (!class_id) return fail(LM_E_INVALID, "class_id is NULL");
  auto it = d->model.classes.find(class_id);
  if (it == d->model.classes.end()) return fail(LM_E_NOTFOUND, "unknown class '%s'", class_id);
  if (template_id < 0 || template_id >= (int)it->second.size()) return fail(LM_E_NOTFOUND, "class '%s' has no template %d", class_id, template_id);
  const TemplatePyramid& tp = it->second[template_id];
  int total = 0;
  for (size_t i = 0; i < tp.size(); ++i) {
    if (hdr) { hdr[i].width = tp[i].width; hdr[i].height = tp[i].height; hdr[i].pyramid_level = tp[i].pyramid_level; hdr[i].num_features = (int)tp[i].features.size(); }
    for (const Feature& f : tp[i].features) {
      if (feats) { feats[3 * total] = f.x; feats[3 * total + 1] = f.y; feats[3 * total + 2] = f.label; }
      ++total;
    }
  }
  return total;
}

int lm_add_synthetic_template(lm_detector* d, const char* class_id, int n_templates, const lm_template_hdr* hdr,
                              const int32_t* feats) {
  if (!d || !class_id || !hdr || !feats) return fail(LM_E_INVALID, "NULL argument");
  int rc = validate_template(d->model, n_templates, hdr, feats);
  if (rc != LM_OK) return rc;
  TemplatePyramid tp((size_t)n_templates);
  size_t k = 0;
  for (int i = 0; i < n_templates; ++i) {
    tp[i].width = hdr[i].width; tp[i].height = hdr[i].height; tp[i].pyramid_level = hdr[i].pyramid_level;
    tp[i].features.resize(hdr[i].num_features);
    for (int j = 0; j < hdr[i].num_features; ++j, ++k) {
      tp[i].features[j].x = feats[3 * k]; tp[i].features[j].y = feats[3 * k + 1]; tp[i].features[j].label = feats[3 * k + 2];
    }
  }
  std::vector<TemplatePyramid>& tps = d->model.classes[class_id];
  tps.push_back(tp);
  ++d->model.version;
  refresh_class_cache(d);
  return (int)tps.size() - 1;
}

int lm_add_template(lm_detector* d, const lm_image* sources, int n_sources, const char* class_id,
                    const lm_image* object_mask, lm_rect* bounding_box) {
  if (!d || !sources || !class_id) return fail(LM_E_INVALID, "NULL argument") - 100;
  if (set_device(d) != LM_OK) return LM_E_CUDA - 100;
  const int L = d->model.levels(), M = d->model.M();
  int rc = check_sources(d, sources, n_sources, nullptr, 0);
  if (rc != LM_OK) return rc - 100;
  const int rows = sources[0].rows, cols = sources[0].cols;
  const bool has_mask = object_mask && object_mask->data;
  if (has_mask && (object_mask->type != LM_8UC1 || object_mask->rows != rows || object_mask->cols != cols))
    return fail(LM_E_INVALID, "object_mask size/type mismatch") - 100;
  Lane& ln = d->lane[0];
  if (upload_luts(d) != LM_OK) return LM_E_CUDA - 100;
  if (ensure_quant_ws(d, ln, rows, cols) != LM_OK) return LM_E_CUDA - 100;
  ln.lm_ready = false; ln.front_valid = false;
  if (upload_frame(d, ln, sources, nullptr, 0) != LM_OK) return LM_E_CUDA - 100;
  ln.launches = 0;
  if (run_quantize(d, ln, ln.stream) != LM_OK) return LM_E_CUDA - 100;
  // download quantised maps (+ CG magnitudes) of every level
  size_t total = 0;
  std::vector<size_t> qoff((size_t)L * M), moff((size_t)L * M, 0);
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      size_t n = (size_t)(rows >> l) * (cols >> l);
      qoff[l * M + m] = total; total += (n + 255) & ~(size_t)255;
      if (d->model.mods[m].type == LM_COLOR_GRADIENT) { moff[l * M + m] = total; total += (n * 4 + 255) & ~(size_t)255; }
    }
  if (ln.stage_out.ensure(total) != LM_OK) return LM_E_CUDA - 100;
  uint8_t* host = ln.stage_out.as<uint8_t>();
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      size_t n = (size_t)(rows >> l) * (cols >> l);
      if (cudaMemcpyAsync(host + qoff[l * M + m], ln.quant_raw[l][m].p, n, cudaMemcpyDeviceToHost, ln.stream) != cudaSuccess) return fail(LM_E_CUDA, "D2H failed") - 100;
      if (d->model.mods[m].type == LM_COLOR_GRADIENT &&
          cudaMemcpyAsync(host + moff[l * M + m], ln.mag[l][m].p, n * 4, cudaMemcpyDeviceToHost, ln.stream) != cudaSuccess)
        return fail(LM_E_CUDA, "D2H failed") - 100;
    }
  if (cudaStreamSynchronize(ln.stream) != cudaSuccess) return fail(LM_E_CUDA, "quantisation kernels failed: %s", cudaGetErrorString(cudaGetLastError())) - 100;

  std::vector<lm_image> qimgs((size_t)L * M);
  std::vector<const float*> mags((size_t)L * M, nullptr);
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      lm_image& q = qimgs[l * M + m];
      q.data = host + qoff[l * M + m]; q.rows = rows >> l; q.cols = cols >> l; q.type = LM_8UC1; q.step = (size_t)(cols >> l);
      if (d->model.mods[m].type == LM_COLOR_GRADIENT) mags[l * M + m] = reinterpret_cast<const float*>(host + moff[l * M + m]);
    }
  int tid = lm_add_template_from_quantized(d, qimgs.data(), mags.data(), class_id, object_mask, bounding_box);
  return tid < -1 ? tid - 100 : tid;
}

int lm_add_template_from_quantized(lm_detector* d, const lm_image* quantized, const float* const* magnitudes,
                                   const char* class_id, const lm_image* object_mask, lm_rect* bounding_box) {
  if (!d || !quantized || !magnitudes || !class_id) return fail(LM_E_INVALID, "NULL argument");
  const int L = d->model.levels(), M = d->model.M();
  const int rows = quantized[0].rows, cols = quantized[0].cols;
  const bool has_mask = object_mask && object_mask->data;
  if (has_mask && (object_mask->type != LM_8UC1 || object_mask->rows != rows || object_mask->cols != cols))
    return fail(LM_E_INVALID, "object_mask size/type mismatch");
  std::vector<std::vector<uint8_t> > qbuf((size_t)L * M);
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      const lm_image& q = quantized[l * M + m];
      if (!q.data || q.type != LM_8UC1 || q.rows != (rows >> l) || q.cols != (cols >> l)) return fail(LM_E_INVALID, "quantized[%d] must be a %dx%d CV_8UC1 image", l * M + m, cols >> l, rows >> l);
      if (d->model.mods[m].type == LM_COLOR_GRADIENT && !magnitudes[l * M + m]) return fail(LM_E_INVALID, "magnitudes[%d] missing", l * M + m);
      qbuf[l * M + m].resize((size_t)q.rows * q.cols);
      for (int y = 0; y < q.rows; ++y) std::memcpy(&qbuf[l * M + m][(size_t)y * q.cols], (const uint8_t*)q.data + (size_t)y * q.step, q.cols);
    }
  // mask pyramid: [OCV] pyrDown() NN-resizes the mask, i.e. plain index decimation
  std::vector<std::vector<uint8_t> > masks((size_t)L);
  if (has_mask) {
    masks[0].resize((size_t)rows * cols);
    for (int y = 0; y < rows; ++y) std::memcpy(&masks[0][(size_t)y * cols], (const uint8_t*)object_mask->data + (size_t)y * object_mask->step, cols);
    for (int l = 1; l < L; ++l) {
      int pc = cols >> (l - 1), r = rows >> l, c = cols >> l;
      masks[l].resize((size_t)r * c);
      for (int y = 0; y < r; ++y)
        for (int x = 0; x < c; ++x) masks[l][(size_t)y * c + x] = masks[l - 1][(size_t)(2 * y) * pc + 2 * x];
    }
  }
  std::vector<TemplatePyramid>& tps = d->model.classes[class_id];  // the reference creates the class entry up front
  refresh_class_cache(d);
  ++d->model.version;
  TemplatePyramid tp((size_t)L * M);
  for (int m = 0; m < M; ++m) {
    const lm_modality_desc& md = d->model.mods[m];
    int nf = md.num_features, ext = md.extract_threshold;
    for (int l = 0; l < L; ++l) {
      if (l > 0) { nf /= 2; ext /= 2; }
      const int r = rows >> l, c = cols >> l;
      const uint8_t* mk = has_mask ? masks[l].data() : nullptr;
      bool ok = md.type == LM_COLOR_GRADIENT
                    ? extract_color_gradient(qbuf[l * M + m].data(), magnitudes[l * M + m], mk, r, c, md.strong_threshold, nf, l, tp[l * M + m])
                    : extract_depth_normal(qbuf[l * M + m].data(), mk, r, c, nf, ext, l, tp[l * M + m]);
      if (!ok) return -1;
    }
  }
  lm_rect bb = crop_templates(tp);
  if (bounding_box) *bounding_box = bb;
  tps.push_back(tp);
  return (int)tps.size() - 1;
}

int lm_set_shard(lm_detector* d, int rank, int world) {
  if (!d || world < 1 || rank < 0 || rank >= world) return fail(LM_E_INVALID, "bad shard %d/%d", rank, world);
  d->shard_rank = rank; d->shard_world = world;
  return LM_OK;
}

int lm_set_similarity_lut(lm_detector* d, const uint8_t lut[256]) {
  for (int i = 0; i < 256; ++i)
    if (lut[i] > 4) return fail(LM_E_INVALID, "similarity LUT entries must be <= 4 (u8 accumulation of 63 features)");
  std::memcpy(d->sim_lut, lut, 256);
  d->luts_dirty = true;
  return LM_OK;
}
int lm_get_similarity_lut(const lm_detector* d, uint8_t lut[256]) { std::memcpy(lut, d->sim_lut, 256); return LM_OK; }
int lm_set_normal_lut(lm_detector* d, const uint8_t lut[8000]) { std::memcpy(d->normal_lut, lut, 8000); d->luts_dirty = true; return LM_OK; }
int lm_get_normal_lut(const lm_detector* d, uint8_t lut[8000]) { std::memcpy(lut, d->normal_lut, 8000); return LM_OK; }

int lm_set_option(lm_detector* d, const char* key, int value) {
  if (!d || !key) return fail(LM_E_INVALID, "NULL argument");
  std::string k(key);
  if (k == "debug_taps") d->debug_taps = value;
  else if (k == "coarse_variant") { d->coarse_variant = value; for (int i = 0; i < LM_LANES; ++i) d->lane[i].front_valid = false; }
  else if (k == "timing") d->timing = value;
  else if (k == "prune") d->prune = value;
  else if (k == "mod_order") d->mod_order = value & 3;
  else if (k == "refine_variant") { d->refine_variant = value; for (int i = 0; i < LM_LANES; ++i) d->lane[i].front_valid = false; }
  else if (k == "graphs") d->graphs = value;
  else if (k == "coarse_grid_limit") {  // process-wide; recorded graphs hold the old grid
    set_coarse_grid_limit(value);
    for (int i = 0; i < LM_LANES; ++i)
      if (d->lane[i].gexec) { cudaGraphExecDestroy(d->lane[i].gexec); d->lane[i].gexec = nullptr; }
  }
  else if (k == "device_out_cap") d->device_out_cap = (uint32_t)std::max(16, value);
  else if (k == "frontend_variant") { d->frontend_variant = value; for (int i = 0; i < LM_LANES; ++i) d->lane[i].front_valid = false; }
  else return fail(LM_E_INVALID, "unknown option '%s'", key);
  return LM_OK;
}

// ---------------------------------------------------------------------------------------------- match
static int front_from_host(lm_detector* d, Lane& ln, const lm_image* sources, int n_sources, const lm_image* masks, int n_masks,
                           bool run = true) {
  int rc = check_sources(d, sources, n_sources, masks, n_masks);
  if (rc != LM_OK) return rc;
  if (set_device(d) != LM_OK || upload_luts(d) != LM_OK) return LM_E_CUDA;
  const int rows = sources[0].rows, cols = sources[0].cols;
  rc = ensure_lm_ws(d, ln, rows, cols);
  if (rc != LM_OK) return rc;
  ln.launches = 0;
  std::memset(ln.work_stats, 0, sizeof(ln.work_stats));
  CU(cudaEventRecord(ln.ev[0], ln.stream));
  if (upload_frame(d, ln, sources, masks, n_masks) != LM_OK) return LM_E_CUDA;
  CU(cudaEventRecord(ln.ev[1], ln.stream));
  if (run) {
    if (run_front(d, ln, ln.stream) != LM_OK) return LM_E_CUDA;
    CU(cudaEventRecord(ln.ev[2], ln.stream));
  }
  // B_front: sources read once + linear memories written once
  uint64_t bf = 0;
  for (int m = 0; m < n_sources; ++m) bf += src_row_bytes(sources[m].type, cols) * rows;
  for (size_t l = 0; l < ln.geom.size(); ++l) bf += (uint64_t)n_sources * 8 * ln.geom[l].rows * ln.geom[l].cols;
  ln.work_stats[0] = bf;
  return LM_OK;
}

int lm_build_front(lm_detector* d, const lm_image* sources, int n_sources, const lm_image* masks, int n_masks) {
  if (!d || !sources) return fail(LM_E_INVALID, "NULL argument");
  Lane& ln = d->lane[0];
  int rc = front_from_host(d, ln, sources, n_sources, masks, n_masks);
  if (rc != LM_OK) return rc;
  CU(cudaStreamSynchronize(ln.stream));
  return LM_OK;
}

static int to_queries(const lm_query* in, int n, Query* out) {
  if (!in || n < 1 || n > kMaxQueries) return fail(LM_E_INVALID, "number of queries must be 1..%d", kMaxQueries);
  for (int q = 0; q < n; ++q) {
    if (in[q].n_ids < 0 || (in[q].n_ids > 0 && !in[q].class_ids)) return fail(LM_E_INVALID, "query %d: bad class id list", q);
    out[q].threshold = in[q].threshold; out[q].class_ids = in[q].class_ids; out[q].n_ids = in[q].n_ids;
  }
  return LM_OK;
}

int lm_match_multi(lm_detector* d, const lm_image* sources, int n_sources, const lm_query* queries, int n_queries,
                   const lm_image* masks, int n_masks, lm_image_out* quantized_out, lm_match_rec** out_matches,
                   size_t* out_offsets) {
  if (!d || !sources || !out_matches || !out_offsets) return fail(LM_E_INVALID, "NULL argument");
  *out_matches = nullptr;
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  Lane& ln = d->lane[0];
  rc = front_from_host(d, ln, sources, n_sources, masks, n_masks);
  if (rc != LM_OK) return rc;
  std::vector<lm_match_rec> out[kMaxQueries];
  rc = match_front(d, ln, qs, n_queries, out);
  if (rc != LM_OK) return rc;
  collect_timings(ln);
  if (quantized_out) {
    const int L = d->model.levels(), M = d->model.M();
    for (int l = 0; l < L; ++l)
      for (int m = 0; m < M; ++m) {
        lm_image_out& q = quantized_out[l * M + m];
        const LevelGeom& g = ln.geom[l];
        if (!q.data || q.rows != g.rows || q.cols != g.cols || q.type != LM_8UC1 || q.step < (size_t)g.cols)
          return fail(LM_E_INVALID, "quantized_out[%d] must be a %dx%d CV_8UC1 image", l * M + m, g.cols, g.rows);
        CU(cudaMemcpy2D(q.data, q.step, ln.quantized[l][m].p, g.cols, g.cols, g.rows, cudaMemcpyDeviceToHost));
      }
  }
  std::vector<lm_match_rec> all;
  out_offsets[0] = 0;
  for (int q = 0; q < n_queries; ++q) {
    all.insert(all.end(), out[q].begin(), out[q].end());
    out_offsets[q + 1] = all.size();
  }
  size_t n = 0;
  return copy_out(all, out_matches, &n);
}

int lm_match(lm_detector* d, const lm_image* sources, int n_sources, float threshold, const char* const* class_ids,
             int n_ids, const lm_image* masks, int n_masks, lm_image_out* quantized_out, lm_match_rec** out_matches,
             size_t* out_n) {
  if (!out_n) return fail(LM_E_INVALID, "NULL argument");
  *out_n = 0;
  lm_query q = {threshold, class_ids, n_ids};
  size_t offs[2] = {0, 0};
  int rc = lm_match_multi(d, sources, n_sources, &q, 1, masks, n_masks, quantized_out, out_matches, offs);
  if (rc == LM_OK) *out_n = offs[1];
  return rc;
}

// Frames pipelined over the two lanes: while lane A's kernels run, lane B's frame is packed / copied to the device.
// out_offsets: n_frames * n_q + 1 prefix offsets, frame-major.
static int match_batch_impl(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, const Query* qs, int n_q,
                            lm_match_rec** out_matches, size_t* out_offsets) {
  *out_matches = nullptr;
  if (n_q < 1 || n_q > kMaxQueries) return fail(LM_E_INVALID, "number of queries must be 1..%d", kMaxQueries);
  std::vector<lm_match_rec> all;
  out_offsets[0] = 0;
  bool busy[LM_LANES] = {};
  auto finish = [&](int li, int frame) -> int {
    Lane& ln = d->lane[li];
    std::vector<lm_raw_match> raw;
    bool overflow = false;
    uint32_t n_cands = 0;
    if (download_records(ln, ln.stream, raw, &overflow, &n_cands, true) != LM_OK) return LM_E_CUDA;
    std::vector<lm_match_rec> out[kMaxQueries];
    if (overflow) {  // rare: redo this frame alone with growing buffers
      int rc = match_front(d, ln, qs, n_q, out);
      if (rc != LM_OK) return rc;
    } else finalize_queries(d, ln, raw, n_q, out);
    for (int q = 0; q < n_q; ++q) {
      all.insert(all.end(), out[q].begin(), out[q].end());
      out_offsets[(size_t)frame * n_q + q + 1] = all.size();
    }
    return LM_OK;
  };
  const bool prof = getenv("LM_HOST_PROFILE") != nullptr;
  double t_fin = 0, t_up = 0, t_plan = 0, t_enq = 0;
  auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int f = 0; f < n_frames; ++f) {
    const int li = f % LM_LANES;
    Lane& ln = d->lane[li];
    double t0 = prof ? now() : 0;
    // The upload of frame f is queued on the lane's stream BEFORE the host waits for the lane's previous frame: stream order
    // keeps it behind that frame's kernels and download, and the copy engine always has the next frame waiting instead of
    // idling until the host comes back (the frame's H2D copy is what bounds the end-to-end rate).  The previous frame's
    // linear memories and result block are untouched until the kernels of frame f are enqueued below.
    int rc = front_from_host(d, ln, sources + (size_t)f * n_sources, n_sources, nullptr, 0, false);  // upload only
    if (rc != LM_OK) return rc;
    double t1 = prof ? now() : 0;
    if (busy[li]) { rc = finish(li, f - LM_LANES); if (rc != LM_OK) return rc; busy[li] = false; }
    double t2 = prof ? now() : 0;
    rc = ensure_pack(d, ln);
    if (rc != LM_OK) return rc;
    Pack::Plan* plan = nullptr;
    rc = get_plan(d, qs, n_q, &plan);
    if (rc != LM_OK) return rc;
    if (ensure_match_buffers(d, ln, std::max<uint32_t>(ln.cand_cap, 1u << 16), std::max<uint32_t>(ln.out_cap, 1u << 14)) != LM_OK) return LM_E_CUDA;
    double t3 = prof ? now() : 0;
    if (enqueue_frame(d, ln, *plan, qs, n_q, ln.stream) != LM_OK) return LM_E_CUDA;
    CU(cudaEventRecord(ln.ev[4], ln.stream));
    if (enqueue_download(ln, ln.stream) != LM_OK) return LM_E_CUDA;
    busy[li] = true;
    if (prof) { double t4 = now(); t_up += t1 - t0; t_fin += t2 - t1; t_plan += t3 - t2; t_enq += t4 - t3; }
  }
  if (prof && n_frames)
    fprintf(stderr, "[lm host profile] per frame us: finish %.1f upload %.1f pack/plan %.1f enqueue %.1f\n", t_fin / n_frames,
            t_up / n_frames, t_plan / n_frames, t_enq / n_frames);
  for (int f = std::max(0, n_frames - LM_LANES); f < n_frames; ++f)
    if (busy[f % LM_LANES]) { int rc = finish(f % LM_LANES, f); if (rc != LM_OK) return rc; busy[f % LM_LANES] = false; }
  size_t n = 0;
  return copy_out(all, out_matches, &n);
}

int lm_match_batch(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, float threshold,
                   const char* const* class_ids, int n_ids, lm_match_rec** out_matches, size_t* out_offsets) {
  if (!d || !sources || !out_matches || !out_offsets || n_frames < 0) return fail(LM_E_INVALID, "NULL argument");
  const Query query = {threshold, class_ids, n_ids};
  return match_batch_impl(d, sources, n_frames, n_sources, &query, 1, out_matches, out_offsets);
}

int lm_match_batch_multi(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                         int n_queries, lm_match_rec** out_matches, size_t* out_offsets) {
  if (!d || !sources || !out_matches || !out_offsets || n_frames < 0) return fail(LM_E_INVALID, "NULL argument");
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  return match_batch_impl(d, sources, n_frames, n_sources, qs, n_queries, out_matches, out_offsets);
}

void lm_free_matches(lm_match_rec* m) { std::free(m); }

int lm_match_device_multi_lane(lm_detector* d, int lane_index, const void* const* d_sources, int n_sources, int rows,
                               int cols, const lm_query* queries, int n_queries, void* stream, const void** d_records,
                               size_t* record_bytes_capacity) {
  if (!d || !d_sources || !d_records) return fail(LM_E_INVALID, "NULL argument");
  if (lane_index < 0 || lane_index >= LM_LANES) return fail(LM_E_INVALID, "lane must be 0..%d", LM_LANES - 1);
  if (n_sources != d->model.M()) return fail(LM_E_INVALID, "sources.size() (%d) != modalities.size() (%d)", n_sources, d->model.M());
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  if (set_device(d) != LM_OK || upload_luts(d) != LM_OK) return LM_E_CUDA;
  Lane& ln = d->lane[lane_index];
  cudaStream_t s = (cudaStream_t)stream;
  const bool fresh_ws = !(ln.lm_ready && ln.rows == rows && ln.cols == cols);
  rc = ensure_lm_ws(d, ln, rows, cols);
  if (rc != LM_OK) return rc;
  if (fresh_ws) CU(cudaStreamSynchronize(ln.stream));  // workspace memsets were enqueued on the lane's own stream
  const bool replay = d->graphs && !ln.graph_broken && !d->debug_taps;
  for (int m = 0; m < n_sources; ++m) {
    ln.has_mask[m] = false;
    if (replay) {  // the recorded graph reads the lane's own source buffers: one device-to-device copy per source
      const size_t bytes = src_row_bytes(expected_src_type(d->model.mods[m]), cols) * rows;
      CU(cudaMemcpyAsync(ln.src[m].p, d_sources[m], bytes, cudaMemcpyDeviceToDevice, s));
      ln.src_ptr[m] = ln.src[m].p;
    } else ln.src_ptr[m] = d_sources[m];
  }
  ln.launches = 0;
  rc = ensure_pack(d, ln);
  if (rc != LM_OK) return rc;
  Pack::Plan* plan = nullptr;
  rc = get_plan(d, qs, n_queries, &plan);
  if (rc != LM_OK) return rc;
  if (ensure_match_buffers(d, ln, std::max<uint32_t>(ln.cand_cap, 1u << 18), std::max<uint32_t>(ln.out_cap, d->device_out_cap)) != LM_OK) return LM_E_CUDA;
  if (enqueue_frame(d, ln, *plan, qs, n_queries, s) != LM_OK) return LM_E_CUDA;
  *d_records = block_ptr(ln);
  if (record_bytes_capacity) *record_bytes_capacity = result_bytes(ln);
  return LM_OK;
}

int lm_match_device_multi(lm_detector* d, const void* const* d_sources, int n_sources, int rows, int cols,
                          const lm_query* queries, int n_queries, void* stream, const void** d_records,
                          size_t* record_bytes_capacity) {
  return lm_match_device_multi_lane(d, 0, d_sources, n_sources, rows, cols, queries, n_queries, stream, d_records,
                                    record_bytes_capacity);
}

int lm_device_result_region(lm_detector* d, const void** base, size_t* lane_stride, int* n_lanes) {
  if (!d || !base || !lane_stride || !n_lanes) return fail(LM_E_INVALID, "NULL argument");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  if (d->results_all.p == nullptr &&
      ensure_match_buffers(d, d->lane[0], std::max<uint32_t>(d->lane[0].cand_cap, 1u << 18), d->device_out_cap) != LM_OK)
    return LM_E_CUDA;
  *base = d->results_all.as<uint8_t>() + kStatsBytes;
  *lane_stride = d->result_stride;
  *n_lanes = LM_LANES;
  return LM_OK;
}

int lm_copy_result_block(lm_detector* d, int lane_index, void* d_dst, size_t bytes, void* stream) {
  if (!d || !d_dst) return fail(LM_E_INVALID, "NULL argument");
  if (lane_index < 0 || lane_index >= LM_LANES || d->lane[lane_index].result.p == nullptr) return fail(LM_E_STATE, "lane %d has no result block yet", lane_index);
  const Lane& ln = d->lane[lane_index];
  if (bytes > result_bytes(ln)) return fail(LM_E_INVALID, "block is only %zu bytes", result_bytes(ln));
  CU(cudaMemcpyAsync(d_dst, block_ptr(ln), bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return LM_OK;
}

int lm_match_device_stream(lm_detector* d, const void* const* d_sources, int n_frames, int n_sources, int rows, int cols,
                           const lm_query* queries, int n_queries, void* const* streams, int n_streams, void* d_stage,
                           size_t stage_slot_bytes) {
  if (!d || !d_sources || !streams || n_frames < 0) return fail(LM_E_INVALID, "NULL argument");
  if (n_streams < 1 || n_streams > LM_LANES) return fail(LM_E_INVALID, "number of streams must be 1..%d", LM_LANES);
  for (int f = 0; f < n_frames; ++f) {
    const int lane = f % n_streams;
    const void* rec = nullptr;
    int rc = lm_match_device_multi_lane(d, lane, d_sources + (size_t)f * n_sources, n_sources, rows, cols, queries, n_queries,
                                        streams[lane], &rec, nullptr);
    if (rc != LM_OK) return rc;
    if (d_stage) {
      rc = lm_copy_result_block(d, lane, static_cast<uint8_t*>(d_stage) + (size_t)f * stage_slot_bytes, stage_slot_bytes, streams[lane]);
      if (rc != LM_OK) return rc;
    }
  }
  return LM_OK;
}

int lm_match_device(lm_detector* d, const void* const* d_sources, int n_sources, int rows, int cols, float threshold,
                    const char* const* class_ids, int n_ids, void* stream, const void** d_records,
                    size_t* record_bytes_capacity) {
  lm_query q = {threshold, class_ids, n_ids};
  return lm_match_device_multi(d, d_sources, n_sources, rows, cols, &q, 1, stream, d_records, record_bytes_capacity);
}

int lm_upload_images(lm_detector* d, const lm_image* images, int n, void* const* d_dst, void* stream) {
  if (!d || n < 0 || (n > 0 && (!images || !d_dst))) return fail(LM_E_INVALID, "NULL argument");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    const lm_image& im = images[i];
    if (!im.data || !d_dst[i]) return fail(LM_E_INVALID, "image %d: NULL data / destination", i);
    const size_t rb = src_row_bytes(im.type, im.cols);
    if (is_pinned(im.data)) {
      if (im.step == rb) CU(cudaMemcpyAsync(d_dst[i], im.data, rb * im.rows, cudaMemcpyHostToDevice, s));
      else CU(cudaMemcpy2DAsync(d_dst[i], rb, im.data, im.step, rb, im.rows, cudaMemcpyHostToDevice, s));
    } else {  // pageable: the copy returns once the source has been read
      CU(cudaStreamSynchronize(s));
      CU(cudaMemcpy2D(d_dst[i], rb, im.data, im.step, rb, im.rows, cudaMemcpyHostToDevice));
    }
  }
  return LM_OK;
}

int lm_finalize_raw(const lm_detector* d, const lm_raw_match* raw, size_t n_raw, lm_match_rec** out_matches, size_t* out_n) {
  if (!d || (!raw && n_raw) || !out_matches || !out_n) return fail(LM_E_INVALID, "NULL argument");
  std::vector<lm_raw_match> r(raw, raw + n_raw);
  std::vector<lm_match_rec> presort, out;
  finalize_records(d->model.levels(), r, presort, out);
  return copy_out(out, out_matches, out_n);
}

int lm_finalize_gathered(const lm_detector* d, const void* blocks, int world, int n_frames, size_t block_bytes,
                         size_t rank_stride, uint32_t capacity_records, int n_queries, lm_match_rec** out_matches,
                         size_t* out_offsets, uint8_t* frame_status) {
  if (!d || !blocks || !out_matches || !out_offsets || !frame_status || world < 1 || n_frames < 0 || n_queries < 1 ||
      n_queries > kMaxQueries || block_bytes < sizeof(ResultHeader))
    return fail(LM_E_INVALID, "bad argument");
  const uint8_t* base = static_cast<const uint8_t*>(blocks);
  const int levels = d->model.levels();
  std::vector<lm_match_rec> all, presort, out;
  std::vector<lm_raw_match> raw, part;
  out_offsets[0] = 0;
  for (int f = 0; f < n_frames; ++f) {
    raw.clear();
    bool over = false, dev_overflow = false;
    for (int r = 0; r < world; ++r) {  // every rank's header first: a frame is finalised only when all blocks are whole
      ResultHeader h;
      std::memcpy(&h, base + (size_t)r * rank_stride + (size_t)f * block_bytes, sizeof(h));
      over = over || h.count > capacity_records;
      dev_overflow = dev_overflow || h.overflow != 0;
    }
    const uint8_t status = over ? 1 : (dev_overflow ? 2 : 0);
    for (int r = 0; r < world && status == 0; ++r) {
      const uint8_t* blk = base + (size_t)r * rank_stride + (size_t)f * block_bytes;
      ResultHeader h;
      std::memcpy(&h, blk, sizeof(h));
      const size_t at = raw.size();
      raw.resize(at + h.count);
      if (h.count) std::memcpy(&raw[at], blk + sizeof(ResultHeader), (size_t)h.count * sizeof(lm_raw_match));
    }
    frame_status[f] = status;
    for (int q = 0; q < n_queries; ++q) {
      if (status == 0) {
        part.clear();
        for (const lm_raw_match& m : raw)
          if ((int)(m.order_key >> 28) == q) part.push_back(m);
        finalize_records(levels, part, presort, out);
        all.insert(all.end(), out.begin(), out.end());
      }
      out_offsets[(size_t)f * n_queries + q + 1] = all.size();
    }
  }
  size_t n = 0;
  return copy_out(all, out_matches, &n);
}

// ---------------------------------------------------------------------------------------------- match clustering
// Restates rgbdDetector::{rcd_voting, cluster_filter, similarity_score_calc, nonMaximaSuppressionUsingIOU, computeIoU}
// (/root/reference/src/rgbdDetector.cpp:36-84, 133-145, 462-574) on the match records this library returns.
namespace {
struct ClusterTmp {
  std::vector<int> index;
  std::vector<uint32_t> members;  // indices into the match list, in arrival order
  double score = 0;
  lm_rect rect = {0, 0, 0, 0};
  bool checked = false;
};
float compute_iou(const lm_rect& r1, const lm_rect& r2) {
  const int r1_minX = r1.x, r1_maxX = r1.x + r1.width - 1, r1_minY = r1.y, r1_maxY = r1.y + r1.height - 1;
  const int r2_minX = r2.x, r2_maxX = r2.x + r2.width - 1, r2_minY = r2.y, r2_maxY = r2.y + r2.height - 1;
  const int minX = std::max(r1_minX, r2_minX), maxX = std::min(r1_maxX, r2_maxX);
  const int minY = std::max(r1_minY, r2_minY), maxY = std::min(r1_maxY, r2_maxY);
  const bool x_inter = (minX >= r1_minX && minX <= r1_maxX) || (minX >= r2_minX && minX <= r2_maxX);
  const bool y_inter = (minY >= r1_minY && minY <= r1_maxY) || (minY >= r2_minY && minY <= r2_maxY);
  float inter_area = 0.0f;
  if (x_inter && y_inter) inter_area = (float)((maxX - minX + 1) * (maxY - minY + 1));
  const float union_area = (float)(r1.width * r1.height + r2.width * r2.height) - inter_area;
  return inter_area / union_area;
}
}  // namespace

int lm_cluster_matches(const lm_match_rec* matches, size_t n_matches, const double* obj_origin_dists, const lm_rect* rects,
                       size_t n_templates, const lm_cluster_params* p, lm_cluster** out_clusters, size_t* out_n,
                       uint32_t** out_match_index) {
  if ((!matches && n_matches) || !obj_origin_dists || !rects || !p || !out_clusters || !out_n || !out_match_index)
    return fail(LM_E_INVALID, "NULL argument");
  if (p->vote_row_col_step <= 0 || !(p->renderer_radius_step > 0)) return fail(LM_E_INVALID, "voting steps must be positive");
  *out_clusters = nullptr; *out_match_index = nullptr; *out_n = 0;
  // rcd_voting: std::map<std::vector<int>, std::vector<Match>> keyed by (row bin, column bin, depth bin)
  std::map<std::vector<int>, ClusterTmp> bins;
  const float depth_step = (float)p->renderer_radius_step;
  for (size_t i = 0; i < n_matches; ++i) {
    const lm_match_rec& m = matches[i];
    if (m.template_id < 0 || (size_t)m.template_id >= n_templates) return fail(LM_E_INVALID, "match %zu: template_id %d outside the pose tables", i, m.template_id);
    const float depth = (float)obj_origin_dists[m.template_id];
    std::vector<int> index(3);
    index[0] = m.y / p->vote_row_col_step;
    index[1] = m.x / p->vote_row_col_step;
    index[2] = (int)((depth - p->renderer_radius_min) / depth_step);
    ClusterTmp& c = bins[index];
    c.index = index;
    c.members.push_back((uint32_t)i);
  }
  // cluster_filter (intent: erase bins with size <= thresh) + cluster_scoring (mean similarity) in map order
  std::vector<ClusterTmp> clusters;
  for (auto& kv : bins) {
    ClusterTmp& c = kv.second;
    if ((int)c.members.size() <= p->cluster_threshold) continue;
    double sum = 0.0;
    int num = 0;
    for (uint32_t i : c.members) { sum += matches[i].similarity; ++num; }
    c.score = sum / num;
    // nonMaximaSuppressionUsingIOU: mean position of the matches, mean size of their templates (integer division)
    int X = 0, Y = 0, Wd = 0, Ht = 0;
    for (uint32_t i : c.members) {
      X += matches[i].x; Y += matches[i].y;
      Wd += rects[matches[i].template_id].width; Ht += rects[matches[i].template_id].height;
    }
    const int n = (int)c.members.size();
    c.rect.x = X / n; c.rect.y = Y / n; c.rect.width = Wd / n; c.rect.height = Ht / n;
    clusters.push_back(c);
  }
  std::sort(clusters.begin(), clusters.end(), [](const ClusterTmp& a, const ClusterTmp& b) { return a.score > b.score; });
  for (size_t i = 0; i < clusters.size(); ++i) {
    if (clusters[i].checked) continue;
    for (size_t j = i + 1; j < clusters.size(); ++j)
      if (!clusters[j].checked && (double)compute_iou(clusters[i].rect, clusters[j].rect) > p->iou_threshold) clusters[j].checked = true;
  }
  size_t n_out = 0, n_idx = 0;
  for (const ClusterTmp& c : clusters) if (!c.checked) { ++n_out; n_idx += c.members.size(); }
  lm_cluster* oc = (lm_cluster*)std::malloc(std::max<size_t>(1, n_out) * sizeof(lm_cluster));
  uint32_t* oi = (uint32_t*)std::malloc(std::max<size_t>(1, n_idx) * sizeof(uint32_t));
  if (!oc || !oi) { std::free(oc); std::free(oi); return fail(LM_E_INVALID, "out of host memory"); }
  size_t k = 0, pos = 0;
  for (const ClusterTmp& c : clusters) {
    if (c.checked) continue;
    lm_cluster& o = oc[k++];
    o.index[0] = c.index[0]; o.index[1] = c.index[1]; o.index[2] = c.index[2];
    o.score = c.score; o.rect = c.rect; o.first = (uint32_t)pos; o.count = (uint32_t)c.members.size();
    for (uint32_t i : c.members) oi[pos++] = i;
  }
  *out_clusters = oc; *out_match_index = oi; *out_n = n_out;
  return LM_OK;
}

void lm_free_clusters(lm_cluster* clusters, uint32_t* match_index) { std::free(clusters); std::free(match_index); }

// ---------------------------------------------------------------------------------------------- parity taps
int lm_level_geometry(lm_detector* d, int level, int32_t out[5], size_t* plane_stride) {
  Lane& ln = d->lane[0];
  if (!ln.lm_ready || level < 0 || level >= (int)ln.geom.size()) return fail(LM_E_STATE, "no front end built");
  const LevelGeom& g = ln.geom[level];
  out[0] = g.rows; out[1] = g.cols; out[2] = g.T; out[3] = g.W; out[4] = g.H;
  if (plane_stride) *plane_stride = g.plane_stride;
  return LM_OK;
}

long lm_debug_fetch(lm_detector* d, int stage, int level, int modality, void* dst) {
  Lane& ln = d->lane[0];
  if (!ln.front_valid) return fail(LM_E_STATE, "no front end built");
  if (level < 0 || level >= d->model.levels() || modality < 0 || modality >= d->model.M()) return fail(LM_E_INVALID, "level/modality out of range");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  const LevelGeom& g = ln.geom[level];
  const size_t n = (size_t)g.rows * g.cols;
  const void* src = nullptr;
  size_t bytes = 0;
  switch (stage) {
    case LM_STAGE_QUANTIZED: src = ln.quantized[level][modality].p; bytes = n; break;
    case LM_STAGE_QUANT_RAW: src = ln.quant_raw[level][modality].p; bytes = n; break;
    case LM_STAGE_MAGNITUDE:
      if (d->model.mods[modality].type != LM_COLOR_GRADIENT) return fail(LM_E_INVALID, "magnitude exists for ColorGradient only");
      src = ln.mag[level][modality].p; bytes = n * 4; break;
    case LM_STAGE_SPREAD:
    case LM_STAGE_RESPONSE:
      if (!ln.debug_taps_written) return fail(LM_E_STATE, "enable lm_set_option(det, \"debug_taps\", 1) before matching");
      src = stage == LM_STAGE_SPREAD ? ln.spread[level][modality].p : ln.response[level][modality].p;
      bytes = stage == LM_STAGE_SPREAD ? n : 8 * n; break;
    case LM_STAGE_LINEAR:
      bytes = 8 * g.plane_stride;
      if (!ln.bytes_valid[level]) {  // only the packed planes exist: unpack them
        if (dst) {
          std::vector<uint8_t> packed(bytes / 2);
          if (cudaStreamSynchronize(ln.stream) != cudaSuccess ||
              cudaMemcpy(packed.data(), ln.lmn[level].as<uint8_t>() + (size_t)modality * 4 * g.plane_stride, bytes / 2, cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(LM_E_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
          uint8_t* o = static_cast<uint8_t*>(dst);
          for (size_t i = 0; i < bytes / 2; ++i) { o[2 * i] = packed[i] & 15; o[2 * i + 1] = packed[i] >> 4; }
        }
        return (long)bytes;
      }
      src = ln.lmem[level].as<uint8_t>() + (size_t)modality * 8 * g.plane_stride; break;
    case LM_STAGE_LINEAR_PACKED:
      if (!ln.nibbles_valid[level]) return fail(LM_E_STATE, "level %d has no packed planes (rows not word-aligned, or a byte kernel variant is selected)", level);
      src = ln.lmn[level].as<uint8_t>() + (size_t)modality * 4 * g.plane_stride; bytes = 4 * g.plane_stride; break;
    default: return fail(LM_E_INVALID, "unknown stage %d", stage);
  }
  if (dst) {
    if (cudaStreamSynchronize(ln.stream) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
      return fail(LM_E_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return (long)bytes;
}

int lm_debug_coarse_map(lm_detector* d, const char* class_id, int template_id, uint16_t* dst) {
  Lane& ln = d->lane[0];
  if (!ln.front_valid) return fail(LM_E_STATE, "no front end built");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  int prc = ensure_pack(d, ln);
  if (prc != LM_OK) return prc;
  Pack& pk = d->pack;
  int local = -1;
  for (const Pack::ClassRange& cr : pk.classes)
    if (cr.id == class_id)
      for (size_t k = 0; k < cr.local.size(); ++k)
        if (pk.h_ctpl[cr.local[k]].template_id == template_id) local = (int)cr.local[k];
  if (local < 0) return fail(LM_E_NOTFOUND, "class '%s' template %d is not on this shard", class_id, template_id);
  const LevelGeom& gc = ln.geom.back();
  const int WH = gc.W * gc.H;
  if (ensure_match_buffers(d, ln, std::max<uint32_t>(ln.cand_cap, 1u << 16), std::max<uint32_t>(ln.out_cap, 1u << 14)) != LM_OK) return LM_E_CUDA;
  const int pass_pos = coarse_positions_per_pass(d->coarse_variant);
  const int P = pk.h_ctpl[local].P;
  std::vector<uint2> tl;
  for (int pass = 0; pass * pass_pos < P; ++pass) tl.push_back(make_uint2(0u, (uint32_t)pass));
  if (ln.dump.ensure((size_t)WH * 2) != LM_OK || ln.work.ensure(64) != LM_OK || ln.work_order.ensure(tl.size() * sizeof(uint2) + 64) != LM_OK) return LM_E_CUDA;
  WorkItem it = {(uint32_t)local, 0u};
  CU(cudaMemsetAsync(ln.dump.p, 0, (size_t)WH * 2, ln.stream));
  CU(cudaMemcpyAsync(ln.work.p, &it, sizeof(it), cudaMemcpyHostToDevice, ln.stream));
  if (!tl.empty()) CU(cudaMemcpyAsync(ln.work_order.p, tl.data(), tl.size() * sizeof(uint2), cudaMemcpyHostToDevice, ln.stream));
  CU(cudaMemsetAsync(ln.result.p, 0, kStatsBytes + sizeof(ResultHeader), ln.stream));
  std::vector<uint32_t> recs;
  int rec_words = 0;
  if (d->coarse_variant == 0) {
    std::vector<WorkItem> one(1, it);
    if (build_tile_records(pk, one, tl, pass_pos, d->model.M(), recs, &rec_words) != LM_OK) return LM_E_INVALID;
    if (ln.dbg_recs.ensure(recs.size() * 4 + 64) != LM_OK) return LM_E_CUDA;
    if (!recs.empty()) CU(cudaMemcpyAsync(ln.dbg_recs.p, recs.data(), recs.size() * 4, cudaMemcpyHostToDevice, ln.stream));
  }
  // threshold 1e30 -> raw threshold saturates: nothing becomes a candidate, the kernel only dumps its accumulators
  QueryThresholds qt;
  for (int q = 0; q < LM_MAX_QUERIES; ++q) qt.v[q] = 1e30f;
  launch_similarity_coarse(d->coarse_variant, ln.lmem[d->model.levels() - 1].as<uint8_t>(), ln.lmn[d->model.levels() - 1].as<uint8_t>(),
                           pk.foff.as<uint32_t>(), pk.ctpl.as<CoarseTpl>(),
                           ln.work.as<WorkItem>(), ln.work_order.as<uint2>(), ln.dbg_recs.as<uint32_t>(), rec_words,
                           (int)tl.size(), qt, d->model.M(),
                           0, ln.cand.as<Cand>(), reinterpret_cast<ResultHeader*>(block_ptr(ln)), nullptr, 0, ln.dump.as<uint16_t>(), WH, ln.stream);
  CU(cudaMemcpyAsync(dst, ln.dump.p, (size_t)WH * 2, cudaMemcpyDeviceToHost, ln.stream));
  CU(cudaStreamSynchronize(ln.stream));
  return LM_OK;
}

long lm_debug_presort(lm_detector* d, lm_match_rec* dst) {
  Lane& ln = d->lane[0];
  if (dst && !ln.presort.empty()) std::memcpy(dst, ln.presort.data(), ln.presort.size() * sizeof(lm_match_rec));
  return (long)ln.presort.size();
}

int lm_last_timings(const lm_detector* d, float ms[5], int* kernel_launches) {
  const Lane& ln = d->lane[0];
  for (int i = 0; i < 5; ++i) ms[i] = ln.ms[i];
  if (kernel_launches) *kernel_launches = ln.launches;
  return LM_OK;
}
int lm_last_work(const lm_detector* d, uint64_t out[8]) {
  for (int i = 0; i < 8; ++i) out[i] = d->lane[0].work_stats[i];
  return LM_OK;
}

}  // extern "C"

// ================================================================================================ template generation
// meshes, the view sphere, rendering, batched addTemplate and the depth hypothesis check.
// Reference loop: /root/reference/src/renderer.cpp:239-329; depth check: src/rgbdDetector.cpp:147-283.
#include <fstream>
#include <sstream>

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
    return fail(LM_E_INVALID, "bad camera (width/height 1..8191, fx, fy > 0, 0 < near < far)");
  return LM_OK;
}

// Enqueues the rasteriser for views [v0, v0 + n) on stream s: images land in d->train.{src of the modality types, mask},
// rectangles in d->train.rects (x_min, y_min, x_max, y_max per view).  want_* select the targets.
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
    if (!look_at(T + 3 * v, up + 3 * v, R, t)) return fail(LM_E_INVALID, "view %d: degenerate camera position / up vector", v);
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
  if (total >= (1ull << 32)) return fail(LM_E_INVALID, "training batch too large");
  if (ws.pool.ensure(total * 8) != LM_OK) return LM_E_CUDA;
  cudaStream_t s0 = d->lane[0].stream;
  CU(cudaMemcpyAsync(ws.segs.p, hs, sizeof(TrainSeg) * S, cudaMemcpyHostToDevice, s0));
  CU(cudaEventRecord(ws.ev[0], s0));
  for (int i = 1; i < lanes; ++i) CU(cudaStreamWaitEvent(d->lane[i].stream, ws.ev[0], 0));
  for (int v = 0; v < n; ++v) {
    const int li = v % lanes;
    Lane& ln = d->lane[li];
    for (int m = 0; m < M; ++m) { ln.src_ptr[m] = d_src[v * M + m]; ln.has_mask[m] = false; }
    ln.launches = 0;
    if (run_quantize(d, ln, ln.stream) != LM_OK) return LM_E_CUDA;
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
  if (cudaStreamSynchronize(s0) != cudaSuccess) return fail(LM_E_CUDA, "training kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
  // host tail: [OCV] cropTemplates + bookkeeping, in view order
  std::vector<TemplatePyramid>& tps = d->model.classes[class_id];  // the reference creates the class entry up front
  refresh_class_cache(d);
  ++d->model.version;
  const uint32_t* hf = ws.h_feats.as<uint32_t>();
  for (int v = 0; v < n; ++v) {
    bool ok = true;
    for (int i = 0; i < L * M; ++i) {
      const int ns = hs[v * L * M + i].n_sel;
      if (ns == -2) return fail(LM_E_STATE, "training candidate pool overflow (view %d)", v);
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
  if ((rows >> (L - 1)) <= 0 || (cols >> (L - 1)) <= 0) return fail(LM_E_INVALID, "image too small for %d pyramid levels", L);
  if (rows > 8191 || cols > 8191) return fail(LM_E_INVALID, "training images are limited to 8191 x 8191");
  for (int m = 0; m < d->model.M(); ++m)
    if (d->model.mods[m].num_features > LM_MAX_FEATURES || d->model.mods[m].num_features < 1)
      return fail(LM_E_INVALID, "num_features must be 1..63");
  return LM_OK;
}

}  // namespace

extern "C" {

int lm_mesh_create(const float* triangles, int n_triangles, lm_mesh** out) {
  if (!out || n_triangles < 0 || (n_triangles > 0 && !triangles)) return fail(LM_E_INVALID, "NULL argument");
  lm_mesh* m = new lm_mesh();
  m->n = n_triangles;
  m->tris.assign(triangles, triangles + (size_t)n_triangles * 9);
  *out = m;
  return LM_OK;
}

int lm_mesh_load_stl(const char* path, lm_mesh** out) {
  if (!path || !out) return fail(LM_E_INVALID, "NULL argument");
  std::ifstream f(path, std::ios::binary);
  if (!f) return fail(LM_E_IO, "cannot open %s", path);
  std::stringstream ss;
  ss << f.rdbuf();
  std::vector<float> tris;
  std::string err;
  if (!parse_stl(ss.str(), tris, err)) return fail(LM_E_IO, "%s: %s", path, err.c_str());
  lm_mesh* m = new lm_mesh();
  m->n = (int)(tris.size() / 9);
  m->tris.swap(tris);
  *out = m;
  return LM_OK;
}

int lm_mesh_num_triangles(const lm_mesh* mesh) { return mesh ? mesh->n : 0; }
int lm_mesh_get_triangles(const lm_mesh* mesh, float* dst) {
  if (!mesh || !dst) return fail(LM_E_INVALID, "NULL argument");
  std::memcpy(dst, mesh->tris.data(), mesh->tris.size() * sizeof(float));
  return LM_OK;
}
void lm_mesh_destroy(lm_mesh* mesh) {
  if (!mesh) return;
  if (mesh->d_tris) { cudaSetDevice(mesh->device); cudaFree(mesh->d_tris); }
  delete mesh;
}

int lm_view_count(const lm_view_sphere* vs) {
  if (!sphere_valid(vs)) return fail(LM_E_INVALID, "bad view sphere (n_points, angle_step, radius_step must be positive)");
  const SphereShape sh = sphere_shape(*vs);
  const long long n = (long long)vs->n_points * sh.n_angles * sh.n_radii;
  if (n > 0x7fffffffLL) return fail(LM_E_INVALID, "view sphere too large");
  return (int)n;
}

int lm_view_params(const lm_view_sphere* vs, int index, double T[3], double up[3], float* radius, int32_t* point_index,
                   int32_t* angle_deg) {
  if (!sphere_valid(vs) || !T || !up) return fail(LM_E_INVALID, "bad view sphere / NULL argument");
  const SphereShape sh = sphere_shape(*vs);
  const int per_point = sh.n_angles * sh.n_radii;
  if (index < 0 || index / per_point >= vs->n_points) return fail(LM_E_NOTFOUND, "view index %d out of range", index);
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
  if (!T || !up || !R || !t) return fail(LM_E_INVALID, "NULL argument");
  if (!look_at(T, up, R, t)) return fail(LM_E_INVALID, "degenerate camera position / up vector");
  return LM_OK;
}

int lm_render_views(lm_detector* d, const lm_mesh* mesh, const lm_camera* cam, const double* T, const double* up,
                    int n_views, uint8_t* bgr, uint16_t* depth, uint8_t* mask, lm_rect* rects) {
  if (!d || !mesh || !T || !up || n_views < 0) return fail(LM_E_INVALID, "NULL argument");
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
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LM_E_CUDA, "rasteriser failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rects)
      for (int v = 0; v < n; ++v) rects[v0 + v] = rect_of(ws.h_rects.as<int>() + 4 * v);
  }
  return LM_OK;
}

int lm_train_views(lm_detector* d, const lm_mesh* mesh, const lm_camera* cam, const double* T, const double* up,
                   int n_views, const char* class_id, int32_t* template_ids, lm_rect* bounding_boxes, lm_rect* mask_rects,
                   uint16_t* centre_depth_mm) {
  if (!d || !mesh || !T || !up || !class_id || !template_ids || n_views < 0) return fail(LM_E_INVALID, "NULL argument");
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
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LM_E_CUDA, "rasteriser failed: %s", cudaGetErrorString(cudaGetLastError()));
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
  if (!d || !class_id || !template_ids || n_views < 0 || (n_views > 0 && (!sources || !masks))) return fail(LM_E_INVALID, "NULL argument");
  const int M = d->model.M();
  if (n_sources != M) return fail(LM_E_INVALID, "sources.size() == modalities.size() violated (%d vs %d)", n_sources, M);
  if (n_views == 0) return LM_OK;
  const int rows = sources[0].rows, cols = sources[0].cols;
  for (int v = 0; v < n_views; ++v) {
    for (int m = 0; m < M; ++m) {
      const lm_image& im = sources[(size_t)v * M + m];
      if (!im.data || im.rows != rows || im.cols != cols || im.type != expected_src_type(d->model.mods[m]))
        return fail(LM_E_INVALID, "view %d source %d: size / type mismatch", v, m);
    }
    const lm_image& mk = masks[v];
    if (!mk.data || mk.type != LM_8UC1 || mk.rows != rows || mk.cols != cols)
      return fail(LM_E_INVALID, "view %d: an object mask of the sources' size is required", v);
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
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LM_E_CUDA, "mask upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    std::vector<int> rects(ws.h_rects.as<int>(), ws.h_rects.as<int>() + 4 * n);
    int rc = train_device_batch(d, rows, cols, n, srcs.data(), dmasks.data(), rects.data(), class_id, template_ids + v0,
                                bounding_boxes ? bounding_boxes + v0 : nullptr);
    if (rc != LM_OK) return rc;
  }
  return LM_OK;
}

int lm_depth_diff_batch(lm_detector* d, const lm_image* scene, const lm_mesh* mesh, const lm_camera* cam, const double* T,
                        const double* up, const int32_t* x, const int32_t* y, int n, double* out) {
  if (!d || !scene || !scene->data || !mesh || !T || !up || !x || !y || !out || n < 0) return fail(LM_E_INVALID, "NULL argument");
  if (scene->type != LM_16UC1) return fail(LM_E_INVALID, "scene depth must be LM_16UC1");
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
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LM_E_CUDA, "rasteriser failed: %s", cudaGetErrorString(cudaGetLastError()));
    const int* hr = ws.h_rects.as<int>();
    for (int v = 0; v < nb; ++v) {
      const lm_rect r = rect_of(hr + 4 * v);
      const int xs = x[v0 + v], ys = y[v0 + v];
      if (r.width > 0 && (xs < 0 || ys < 0 || xs + r.width > scene->cols || ys + r.height > scene->rows))
        return fail(LM_E_INVALID, "hypothesis %d: the %dx%d template crop at (%d, %d) leaves the %dx%d scene image", v0 + v, r.width,
                    r.height, xs, ys, scene->cols, scene->rows);
      launch_depth_diff(ws.scene.as<uint16_t>(), scene->cols, ws.src[1].as<uint16_t>() + px * v, ws.mask.as<uint8_t>() + px * v,
                        cam->width, xs, ys, r.x, r.y, r.width, r.height, ws.diff.as<unsigned long long>() + 2 * v, s);
    }
    CU(cudaGetLastError());
    unsigned long long* hd = reinterpret_cast<unsigned long long*>(ws.h_rects.as<uint8_t>() + 16 * (size_t)nb);
    CU(cudaMemcpyAsync(hd, ws.diff.p, 16 * (size_t)nb, cudaMemcpyDeviceToHost, s));
    if (cudaStreamSynchronize(s) != cudaSuccess) return fail(LM_E_CUDA, "depth_diff failed: %s", cudaGetErrorString(cudaGetLastError()));
    for (int v = 0; v < nb; ++v) out[v0 + v] = (double)hd[2 * v] / ((double)hd[2 * v + 1] * 1000.0);
  }
  return LM_OK;
}

}  // extern "C"

// ================================================================================================ pose table
// writeLinemodTemplateParams (/root/reference/src/renderer.cpp:72-123) / readLinemodTemplateParams
// (src/rgbdDetector.cpp:1681-1749): cv::FileStorage YAML, one "Template i" map per template + the renderer_* scalars.
#include "lm_yaml.hpp"

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
  if (!path || (n && !poses) || !p) return fail(LM_E_INVALID, "NULL argument");
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
  if (!w.save(path, err)) return fail(LM_E_IO, "%s", err.c_str());
  return LM_OK;
}

int lm_read_renderer_params(const char* path, lm_template_pose** out_poses, size_t* out_n, lm_renderer_params* p) {
  if (!path || !out_poses || !out_n || !p) return fail(LM_E_INVALID, "NULL argument");
  lmyaml::Node root;
  std::string err;
  if (!lmyaml::parse_file(path, root, err)) return fail(LM_E_IO, "%s: %s", path, err.c_str());
  std::vector<lm_template_pose> poses;
  for (size_t i = 0;; ++i) {  // the reference reads "Template 0", "Template 1", ... until the first missing key
    const lmyaml::Node& t = root["Template " + std::to_string(i)];
    if (t.empty()) break;
    lm_template_pose ps;
    std::memset(&ps, 0, sizeof(ps));
    const lmyaml::Node& rc = t["Rect"];
    if (!read_matrix(t["R"], 9, ps.R, nullptr) || !read_matrix(t["T"], 3, ps.T, nullptr) || !read_matrix(t["K"], 9, nullptr, ps.K) ||
        !t["D"].as_double(ps.D) || !t["Ori_dist"].as_double(ps.ori_dist) || rc.kind != lmyaml::Node::SEQ || rc.size() != 4)
      return fail(LM_E_IO, "%s: malformed entry \"Template %zu\"", path, i);
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
  if (!ok) return fail(LM_E_IO, "%s: renderer_* parameters missing", path);
  *out_n = poses.size();
  *out_poses = nullptr;
  if (!poses.empty()) {
    *out_poses = (lm_template_pose*)std::malloc(poses.size() * sizeof(lm_template_pose));
    if (!*out_poses) return fail(LM_E_INVALID, "out of memory");
    std::memcpy(*out_poses, poses.data(), poses.size() * sizeof(lm_template_pose));
  }
  return LM_OK;
}

void lm_free_poses(lm_template_pose* poses) { std::free(poses); }

}  // extern "C"
