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
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
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

// Grow-only device allocation holding `frames` equally sized per-frame regions, `stride` bytes apart (multiple of 256):
// the kernels of the matching path address frame f of a chunk as base + f * stride.
struct FrameBuf {
  DevBuf buf;
  size_t stride = 0;
  int frames = 0;
  int ensure(size_t bytes_per_frame, int n_frames, bool* grew = nullptr) {
    if (grew) *grew = false;
    size_t st = (bytes_per_frame + 255) & ~(size_t)255;
    if (buf.p && st <= stride && n_frames <= frames) return LM_OK;
    st = std::max(st, stride);
    n_frames = std::max(n_frames, frames);
    buf.release();
    stride = 0; frames = 0;
    if (buf.ensure(st * (size_t)n_frames) != LM_OK) return LM_E_CUDA;
    stride = st; frames = n_frames;
    if (grew) *grew = true;
    return LM_OK;
  }
  void release() { buf.release(); stride = 0; frames = 0; }
  template <class T> T* as(int f = 0) const { return reinterpret_cast<T*>(static_cast<uint8_t*>(buf.p) + (size_t)f * stride); }
  template <class T> size_t stride_in() const { return stride / sizeof(T); }
  size_t bytes() const { return buf.cap; }
};

struct LevelGeom {
  int rows = 0, cols = 0, T = 0, W = 0, H = 0;
  size_t plane_stride = 0;   // positions per orientation plane in the reference's flat order (byte planes, parity taps)
  int Hh = 0;                // != 0: the level's nibble planes are column-blocked (lm_kernels.cuh tiled_nibble_index), H + 16
  size_t nib_plane = 0;      // positions per orientation plane of the nibble planes (plane_stride, or tiled_plane_stride)
};
// Nibble-packed rows of a level start on 32-bit words when W and W*H are multiples of 8 (one word = 8 positions).
static inline bool level_nibble_aligned(const LevelGeom& g) { return (g.W % 8) == 0 && ((size_t)g.W * g.H) % 8 == 0; }

static inline size_t plane_stride_of(int T, int W, int H) {
  size_t wh = (size_t)W * H;
  return ((size_t)T * T * wh + wh + 16 * (size_t)W + 16 + 15) & ~(size_t)15;  // same rule as the oracle (App. D-2)
}
static const size_t kLmSlack = 8192;  // tail slack: vector loads of partially filled passes may over-read

// Workspace lanes per handle.  A lane processes one CHUNK of frames at a time (one launch set: front end, coarse
// similarity and refinement kernels each cover every frame of the chunk); lm_match_batch* pipelines chunks over
// `batch_lanes` of them so that the host->device copies, the kernels and the result download of consecutive chunks
// overlap, and lm_match_device_multi_lane / lm_match_device_stream expose them to callers that manage their own streams.
static const int LM_LANES = 8;
static const int LM_GRAPH_SLOTS = 6;  // recorded launch geometries per lane: 1, 2, 4, 8, 16, 32 frames

// One in-flight chunk: stream, events, device workspace, pinned staging.
struct Lane {
  cudaStream_t stream = nullptr;
  cudaStream_t user_stream = nullptr;   // the caller's stream of the lane's last device-resident chunk
  bool user_stream_valid = false;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // modalities quantise concurrently: modality m > 0 runs on side[m-1], forked from / joined into the chunk's stream
  cudaStream_t side[LM_MAX_MODALITIES - 1] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[LM_MAX_MODALITIES - 1] = {nullptr, nullptr, nullptr};
  int rows = 0, cols = 0;     // geometry of the quantisation workspace
  int frames = 0;             // frame slots the workspace holds
  bool lm_ready = false;      // LM buffers sized + zero-tailed for (rows, cols, frames)
  bool front_valid = false;
  bool debug_taps_written = false;
  bool bytes_valid[LM_MAX_LEVELS] = {false, false, false, false};    // byte planes written by the last front end
  bool nibbles_valid[LM_MAX_LEVELS] = {false, false, false, false};  // nibble planes written by the last front end
  std::vector<LevelGeom> geom;
  // per modality
  FrameBuf src[LM_MAX_MODALITIES];     // level-0 sources uploaded from the host (BGR / depth)
  const void* src_ptr[LM_MAX_BATCH][LM_MAX_MODALITIES] = {};  // sources of the current chunk: own slots or the caller's device pointers
  int n_frames = 0;                    // frames of the current / last chunk
  DevBuf mask0[LM_MAX_MODALITIES];     // single-frame requests only
  bool has_mask[LM_MAX_MODALITIES] = {false, false, false, false};
  // per (level, modality)
  FrameBuf bgr[LM_MAX_LEVELS][LM_MAX_MODALITIES];       // CG pyramid sources for level >= 1
  FrameBuf mag[LM_MAX_LEVELS][LM_MAX_MODALITIES];
  FrameBuf quant_raw[LM_MAX_LEVELS][LM_MAX_MODALITIES];
  FrameBuf quantized[LM_MAX_LEVELS][LM_MAX_MODALITIES];
  DevBuf spread[LM_MAX_LEVELS][LM_MAX_MODALITIES], response[LM_MAX_LEVELS][LM_MAX_MODALITIES];  // parity taps only (frame 0)
  FrameBuf lmem[LM_MAX_LEVELS];                        // [M][8][plane_stride] byte planes + slack: only for the parity taps and
                                                       // for levels whose nibble rows are not word-aligned (packed from here)
  FrameBuf lmn[LM_MAX_LEVELS];                         // [M][8][plane_stride / 2] nibble-packed planes + slack: what the
                                                       // matching kernels read
  DevBuf ctl;                                          // BatchCtl: frame table, dispenser / counters (k_begin_chunk)
  // The GPU work of one chunk (front end, coarse, refine) as an instantiated CUDA graph per launch geometry (frames in
  // the grid, rounded up to a power of two): replayed instead of ~20 runtime calls per chunk.  Valid while `key` matches.
  struct GraphKey {
    const void* plan; const void* plan_recs; const void* cand; const void* result; const void* lmn0; const void* ctl;
    uint64_t model_version; int rows, cols, n_q, n_tiles, prune, shard_rank, shard_world, ws_frames; uint32_t cand_cap, out_cap;
    float thr[LM_MAX_QUERIES];
  };
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    GraphKey key;
    int launches = 0;
  } graph[LM_GRAPH_SLOTS];
  bool graph_broken = false;  // capture failed once on this lane: stay on the eager path
  // matching
  DevBuf cand, work, work_order, dump, dbg_recs;
  FrameBuf result;            // per frame: [16 B statistics][ResultHeader][out_cap x lm_raw_match]
  uint32_t cand_cap = 0, out_cap = 0;
  uint32_t head_records = 256;  // records downloaded together with each frame's header (follows the survivor counts)
  PinBuf stage_in, stage_out;
  // last-call bookkeeping (frame 0 of the last chunk)
  float ms[5] = {0, 0, 0, 0, 0};
  int launches = 0;
  uint64_t work_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<lm_match_rec> presort;

  void drop_graphs() {
    for (int i = 0; i < LM_GRAPH_SLOTS; ++i)
      if (graph[i].exec) { cudaGraphExecDestroy(graph[i].exec); graph[i].exec = nullptr; }
  }
  int init() {
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (int i = 0; i < 6; ++i) CU(cudaEventCreate(&ev[i]));
    CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    if (ctl.ensure(sizeof(BatchCtl)) != LM_OK) return LM_E_CUDA;
    CU(cudaMemset(ctl.p, 0, sizeof(BatchCtl)));
    for (int i = 0; i < LM_MAX_MODALITIES - 1; ++i) {
      CU(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
      CU(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
    }
    return LM_OK;
  }
  void destroy() {
    for (int m = 0; m < LM_MAX_MODALITIES; ++m) {
      src[m].release(); mask0[m].release();
      for (int l = 0; l < LM_MAX_LEVELS; ++l) {
        bgr[l][m].release(); mag[l][m].release(); quant_raw[l][m].release(); quantized[l][m].release();
        spread[l][m].release(); response[l][m].release();
      }
    }
    for (int l = 0; l < LM_MAX_LEVELS; ++l) lmem[l].release();
    for (int l = 0; l < LM_MAX_LEVELS; ++l) lmn[l].release();
    cand.release(); work.release(); work_order.release(); dump.release(); dbg_recs.release(); ctl.release();
    result.release();
    stage_in.release(); stage_out.release();
    drop_graphs();
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
  int rows = 0, cols = 0, shard_rank = 0, shard_world = 1;
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
  struct Plan { DevBuf items, recs; int n_items = 0, n_tiles = 0, n_full = 0, rec_words = 0, max_feat = 0; uint64_t coarse_bytes = 0, evals = 0; double refine_nf_sum = 0; };
  std::map<std::string, Plan> plans;
  void clear_filtered() {
    for (auto& kv : plans) { kv.second.items.release(); kv.second.recs.release(); }
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

// Host threads that turn downloaded survivor records into the reference's ordered match lists (two std::sort passes per
// frame and query) while the calling thread keeps the copy engine and the launch queue fed.  Started on first use.
struct FinalizePool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable wake, idle;
  std::deque<std::function<void()> > jobs;
  int busy = 0;
  bool stop = false;
  void start(int n) {
    while ((int)threads.size() < n)
      threads.emplace_back([this]() {
        for (;;) {
          std::function<void()> job;
          {
            std::unique_lock<std::mutex> lk(mu);
            wake.wait(lk, [this]() { return stop || !jobs.empty(); });
            if (stop && jobs.empty()) return;
            job.swap(jobs.front());
            jobs.pop_front();
            ++busy;
          }
          job();
          {
            std::lock_guard<std::mutex> lk(mu);
            if (--busy == 0 && jobs.empty()) idle.notify_all();
          }
        }
      });
  }
  void submit(std::function<void()> job) {
    { std::lock_guard<std::mutex> lk(mu); jobs.push_back(std::move(job)); }
    wake.notify_one();
  }
  void wait_all() {
    std::unique_lock<std::mutex> lk(mu);
    idle.wait(lk, [this]() { return busy == 0 && jobs.empty(); });
  }
  ~FinalizePool() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    wake.notify_all();
    for (auto& t : threads) t.join();
  }
};

struct lm_detector {
  HostModel model;
  FinalizePool finalizers;
  int finalize_threads = 4;  // host threads ordering the match lists of batched calls (0: on the calling thread)
  TrainWs train;
  int device = -1;
  bool cuda_ready = false;
  uint8_t sim_lut[256];
  uint8_t normal_lut[8000];
  DevBuf d_resp_all, d_normal_lut;
  uint32_t device_out_cap = 2048;  // records per frame block on the device-resident paths
  uint32_t cand_per_frame = 1u << 16;  // coarse candidates a chunk may produce, per frame of the chunk
  bool luts_dirty = true;
  Lane lane[LM_LANES];
  Pack pack;
  int shard_rank = 0, shard_world = 1;
  int debug_taps = 0, timing = 0, graphs = 1;
  int prune = 3;         // exact early termination: bit 0 in the coarse kernel, bit 1 in the refinement kernel
  int batch_frames = 8;  // frames per chunk on the batched paths (lm_match_batch*, lm_match_device_stream)
  int batch_lanes = 4;   // chunks in flight on the batched host path
  int stream_frames = 16;  // frames per chunk of an lm_stream: no fill and drain per call to pay, so larger launch sets win
  int mod_order = 2;  // coarse kernel: 0 = modalities in template order, 1 = reversed, 2 = chosen per frame (default)
  int refine_tiled = 1;  // refinement levels with W % 16 == 0 and H >= 16 keep their nibble planes column-blocked
  int coarse_share = 1;  // coarse tail passes of <= 128 positions are scored for eight frames per warp
  int dn_count = 1;      // DepthNormal's medianBlur(5) by counting when the NORMAL_LUT is one-hot (A/B switch)
  bool stream_open = false;  // an lm_stream owns the lanes: other matching calls are refused until it is closed
  std::vector<std::string> class_id_cache;
};

// ------------------------------------------------------------------------------------------------ shared helpers
// (defined in lm_detector.cu)
int set_device(lm_detector* d);                                         // binds the handle to its CUDA device; no CPU path
int upload_luts(lm_detector* d);
int ensure_quant_ws(lm_detector* d, Lane& ln, int rows, int cols, int frames = 1);
// Installs the frame table of the lane's current chunk (ln.src_ptr[0 .. n_frames)) and zeroes the chunk's counters and the
// first `result_blocks` result headers: must precede the chunk's kernels on `s`.
int begin_chunk(lm_detector* d, Lane& ln, int n_frames, int result_blocks, cudaStream_t s);
// [OCV] Modality::process + pyrDown, every level, for grid_frames frame slots (frames beyond the table's count idle)
int run_quantize(lm_detector* d, Lane& ln, int grid_frames, cudaStream_t main_stream);
int upload_image(Lane& ln, const lm_image& im, void* dst, size_t* stage_off);
bool is_pinned(const void* p);
size_t src_row_bytes(int type, int cols);
int expected_src_type(const lm_modality_desc& m);
void refresh_class_cache(lm_detector* d);

// ------------------------------------------------------------------------------------------------ lm_group.cu <-> lm_detector.cu
extern "C" {
// lm_match_batch_multi that hands back every frame's un-ordered survivor records instead of finalised lists
int lm_internal_match_batch_raw(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                                int n_queries, std::vector<std::vector<lm_raw_match> >* raw_frames);
// raw records of one frame (all shards) -> per query the reference's sorted / de-duplicated list
void lm_internal_finalize(int levels, std::vector<lm_raw_match>& raw, int n_queries, std::vector<lm_match_rec>* out);
// a new handle with the same model (templates, modalities, T), tables and options, not yet bound to a device
lm_detector* lm_internal_clone(const lm_detector* src);
}
