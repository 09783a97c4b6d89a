// lm_detector_internal.hpp -- types shared by the translation units behind the C ABI (lm_detector.cu: lifecycle, matching,
// parity taps; lm_training.cu: meshes, rendering, batched addTemplate, pose table): buffers, workspace lanes, the template
// pack and the lm_detector handle itself.  Not installed; include/linemod_b200.h is the public surface.
#pragma once
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
// Records the message lm_last_error() returns on this thread and hands the code back (defined in lm_detector.cu).
int lm_fail(int code, const char* fmt, ...);
#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) return lm_fail(LM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
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
static inline bool level_nibble_aligned(const LevelGeom& g) { return (g.W % 8) == 0 && ((size_t)g.W * g.H) % 8 == 0; }

static inline size_t plane_stride_of(int T, int W, int H) {
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

// ------------------------------------------------------------------------------------------------ shared helpers
// (defined in lm_detector.cu)
int set_device(lm_detector* d);                                         // binds the handle to its CUDA device; no CPU path
int upload_luts(lm_detector* d);
int ensure_quant_ws(lm_detector* d, Lane& ln, int rows, int cols);
int run_quantize(lm_detector* d, Lane& ln, cudaStream_t main_stream);  // [OCV] Modality::process + pyrDown, every level
int upload_image(Lane& ln, const lm_image& im, void* dst, size_t* stage_off);
bool is_pinned(const void* p);
size_t src_row_bytes(int type, int cols);
int expected_src_type(const lm_modality_desc& m);
void refresh_class_cache(lm_detector* d);
