// lm_kernels.cuh -- device-side records and kernel launchers of the LINEMOD hot path (sm_100a).
//
// Kernel <-> reference function map ([OCV] = OpenCV 2.4.x modules/objdetect/src/linemod.cpp, the code behind
// cv::linemod::Detector::match called at /root/reference/src/rgbdDetector.cpp:33):
//   production path
//   k_pyrdown_fast                              [OCV] ColorGradientPyramid::pyrDown -> cv::pyrDown            
//   k_cg_fused                                  [OCV] quantizedOrientations + hysteresisGradient, all levels (a1, a2)
//   k_dn_fused                                  [OCV] quantizedNormals + medianBlur(5) + DepthNormalPyramid::pyrDown (a4, a5)
//   k_spread_all                                [OCV] quantize(mask) + spread + computeResponseMaps + linearize (a6-a8)
//   k_similarity_coarse_rec63 / _rec            [OCV] similarity + addSimilarities + matchClass coarse scan  (a9-a12); _rec63: requests
//                                               whose tiles have <= 63 features (u8 sums only), _rec: the general body (u16 totals)
//   k_refine_nib                                [OCV] similarityLocal + matchClass refinement loop          (a13)
//   template generation (/root/reference/src/renderer.cpp:239-329, src/rgbdDetector.cpp:147-283)
//   k_raster_tris, k_raster_resolve             RendererIterator::render / renderDepthOnly (ORK) per oracle/render_oracle.cpp
//   k_train_cg, k_train_dn_pb/dist/keys         [OCV] ColorGradientPyramid / DepthNormalPyramid::extractTemplate candidates
//   k_train_sort, k_train_select                [OCV] std::stable_sort + QuantizedPyramid::selectScatteredFeatures
//   k_depth_diff, k_mask_rect                   rgbdDetector::depth_diff; mask bounding boxes
//   k_begin_chunk                               per launch set: frame table (source pointers of the chunk's frames), zeroed
//                                               result headers / counters -- everything that changes between replays of a
//                                               lane's CUDA graph
//   k_pack_nibbles                              byte planes -> nibble planes for levels whose rows are not word-aligned
//
// Every kernel of the matching path takes a CHUNK of frames: blockIdx.y / .z (front end), the virtual tile index (coarse)
// or the candidate's frame tag (refinement) select the frame, whose buffers lie at base + frame * stride; the level-0
// sources come from the device-resident FrameTable.  A single-frame call is a chunk of one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/linemod_b200.h"

#define LM_MAX_QUERIES 8

namespace lmk {

struct CoarseTpl {                        // one per template (canonical order), coarse (lowest) pyramid level
  uint32_t feat_begin;                    // first entry in the coarse feature-offset array
  uint8_t cnt[LM_MAX_MODALITIES][4];      // in-bounds features per modality, grouped by (offset & 15) >> 2
  int32_t P;                              // template_positions = span_y*W + span_x + 1 (<= 0: nothing to score)
  uint32_t nf;                            // sum over modalities of features.size() (threshold denominator)
  uint32_t order_key;                     // canonical order index (class map order, then template_id)
  int32_t template_id, class_index;
};

struct RefineTpl {                        // one per (level < L-1, template)
  uint32_t feat_begin;                    // first entry in that level's packed feature array
  uint16_t cnt[LM_MAX_MODALITIES];        // all features per modality
  int32_t width, height;                  // of the level's first-modality template (clamp window)
  uint32_t nf;                            // sum of features.size() over modalities
};

struct RefineLevel {
  const uint8_t* lmn;                     // [M][8][plane_stride / 2] nibble-packed planes of frame 0
  unsigned long long frame_stride;        // bytes between the planes of consecutive frames of the chunk
  const RefineTpl* tpl;
  const uint32_t* feats;                  // (x + 4096) | (y + 4096) << 13 | label << 26
  unsigned long long plane_stride;        // positions (nibbles) per orientation plane, in the layout the level is stored in
  int rows, cols, T, W;
  unsigned int inv_T;                     // ceil(2^32 / T) for T >= 2: x / T = umulhi(x, inv_T) for the coordinates that occur
  int Hh;                              // 0: flat planes (the reference's linearize order).  Otherwise the COLUMN-BLOCKED
                                          // layout of refinement levels (see tiled_nibble_index): rows per column block
};

// Column-blocked layout of a refinement level's nibble planes (levels with W % 16 == 0 and H >= 16).  similarityLocal
// reads 16 x 16 windows: in the reference's flat order the 16 rows of a window lie W/2 bytes apart (16 cache lines per
// feature); here a (T^2 phase) matrix is cut into blocks of 16 columns and the rows of a block are stored back to back,
// 8 bytes each, so the 16 rows of a window's 16-column chunk are 128 contiguous bytes.  A phase holds W/16 column blocks of
// Hh = H + 16 rows: rows H .. H+15 repeat rows 0 .. 15 of the NEXT phase (zero after the last one), which is what the
// reference's flat over-read past the bottom of a phase matrix sees; a window's second chunk past the
// last column block is column block 0 one row down, again the flat order's successor.
__host__ __device__ inline size_t tiled_nibble_index(int W, int Hh, int phase, int row, int col) {
  return (size_t)phase * ((size_t)W * Hh) + (size_t)(col >> 4) * ((size_t)Hh * 16) + (size_t)row * 16 + (size_t)(col & 15);
}
static inline size_t tiled_plane_stride(int T, int W, int H) {   // nibbles per orientation plane incl. a 256-nibble zero run
  return ((size_t)T * T * W * (size_t)(H + 16) + 256 + 31) & ~(size_t)31;
}

struct RefineParams {
  RefineLevel level[LM_MAX_LEVELS];       // index = pyramid level (only 0 .. L-2 used)
  int levels, M, coarse_T, coarse_W;
  int prune;                              // exact early termination of hopeless candidates (warp-per-candidate path)
  int mod_order;                          // order of the modalities in the sum: 0 template order, 1 reversed, 2 per frame (mod_bits)
  float threshold[LM_MAX_QUERIES];        // per query of the request
};

struct QueryThresholds {
  float v[LM_MAX_QUERIES];
};

struct WorkItem {                         // one template of one query of the request
  uint32_t tglob;                         // index into the packed template records of this shard
  uint32_t order;                         // emission order key: position in the query's iteration | query << 28
};

struct Cand {                             // coarse candidate: raw score above the template's raw threshold
  uint32_t tglob, pos;                    // template index in the pack; raster position at the coarsest level (bits 0-23,
                                          // <= 4095^2) | frame of the chunk << 24
  uint32_t raw_nf;                        // raw score (u16) | number of features behind it << 16
  uint32_t order;                         // the work item's emission order key (query << 28 | position in the iteration)
};

struct ResultHeader {                     // one per frame, zeroed before every request (k_begin_chunk)
  uint32_t count;                         // surviving matches written (may exceed the capacity: overflow)
  uint32_t reserved;
  uint32_t overflow, n_cands;             // n_cands: coarse candidates of this frame (counted by the refinement kernel)
};

#define LM_MAX_BATCH 32                   // frames per launch set (chunk)

struct FrameTable {                       // the frames of one launch set
  int32_t n_frames, pad[3];
  const void* src[LM_MAX_BATCH][LM_MAX_MODALITIES];  // level-0 sources, tightly packed rows, device memory
};

struct BatchCtl {                         // device-resident, one per workspace lane; rewritten by k_begin_chunk
  FrameTable ft;
  uint32_t next_tile, n_cands, overflow, pad;        // coarse kernel: tile dispenser, chunk-wide candidate count
  unsigned int mod_bits[LM_MAX_BATCH][LM_MAX_MODALITIES];  // per frame and modality: orientation bits set in the coarsest
                                                           // level's spread image (front end -> coarse kernel hint)
};

// ------------------------------------------------------------------------------------------------ front end
// Buffers of frame f of a chunk lie at base + f * stride (strides in ELEMENTS of the pointer's type); the level-0 source
// of modality `modality` is ctl->ft.src[f][modality].  Blocks of frames >= ctl->ft.n_frames exit at once, so one recorded
// launch geometry (a lane's CUDA graph) serves every chunk size up to the grid's frame count.
struct CgLevel {
  const uint8_t* src;  // BGR source of this level for levels >= 1 (output of k_pyrdown_fast); level 0 comes from the table
  float* mag;          // [rows][cols] squared gradient magnitude (exact integer in f32)
  uint8_t* quant;      // [rows][cols] one-hot quantised orientation (unmasked)
  unsigned long long src_stride, mag_stride, quant_stride;
  int rows, cols, block_begin, blocks_x;
};
struct CgParams {
  CgLevel lv[LM_MAX_LEVELS];
  const BatchCtl* ctl;
  int n_levels, modality;
  float thr_sq;  // weak_threshold^2
};
struct DnParams {
  const BatchCtl* ctl;
  const uint8_t* lut;  // NORMAL_LUT, 8000 bytes
  uint8_t* quant[LM_MAX_LEVELS];
  unsigned long long quant_stride[LM_MAX_LEVELS];
  int rows, cols, n_levels, modality, distance_threshold, difference_threshold;
};
struct SpreadEntry {
  const uint8_t* qraw;   // unmasked quantisation of this (level, modality)
  const uint8_t* mask0;  // level-0 mask or null (single-frame requests only)
  uint8_t* quantized;    // masked quantisation (Detector::match's quantized_images)
  uint8_t* spread;       // parity tap or null (frame 0 only)
  uint8_t* response;     // parity tap or null (frame 0 only)
  uint8_t* lm;           // this modality's 8 orientation byte planes, or null when only lm_nib is needed
  uint8_t* lm_nib;       // this modality's 8 nibble-packed planes (nib_plane / 2 bytes each), or null
  unsigned long long nib_plane;  // positions per orientation plane of lm_nib (plane_stride, or tiled_plane_stride)
  int tiled_Hh;          // lm_nib is column-blocked with this many rows per block (0: flat), see tiled_nibble_index
  int count_bits;        // coarsest level: ctl->mod_bits[frame][modality] += orientation bits set in the spread image;
                         // the coarse kernel starts with the modality that has fewer (lower responses, earlier pruning)
  unsigned long long plane_stride;
  unsigned long long qraw_stride, quantized_stride, lm_stride, lm_nib_stride;  // per frame, bytes
  int rows, cols, T, W, H, level, modality, mask_cols0, block_begin, blocks_x;
};
struct SpreadParams {
  SpreadEntry e[LM_MAX_LEVELS * LM_MAX_MODALITIES];
  const uint32_t* resp_all;
  BatchCtl* ctl;
  int n;
};
// Per launch set, outside the lane's CUDA graph: installs the frame table, zeroes the coarse kernel's dispenser and
// counters and the (statistics + header) prefix of the first n_blocks result blocks (results + i * result_stride).
void launch_begin_chunk(const FrameTable& ft, BatchCtl* ctl, uint8_t* results, size_t result_stride, int n_blocks,
                        cudaStream_t s);
// src: level-0 source of modality `modality` when null (frame table), else the previous pyramid level at src + f * src_stride
void launch_pyrdown_fast(const BatchCtl* ctl, int modality, const uint8_t* src, size_t src_stride, int rows, int cols,
                         uint8_t* dst, size_t dst_stride, int n_frames, cudaStream_t s);
int cg_fused_blocks(int rows, int cols, int* blocks_x);
void launch_cg_fused(const CgParams& p, int total_blocks, int n_frames, cudaStream_t s);
void launch_dn_fused(const DnParams& p, int n_frames, cudaStream_t s);
int spread_all_blocks(int T, int W, int H, int* blocks_x);
bool launch_spread_all(const SpreadParams& p, int total_blocks, int n_frames, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ matching
// Coarse similarity of a chunk of frames against the request's tile records, one launch.  Virtual tile v = frame *
// n_tiles + tile (frame-major: at any time nearly every SM reads the same frame's linear memories, and the tail of one
// frame -- the dependent-load chain of the tiles that survive every pruning test -- overlaps the head of the next).
//
// Tile record (rec_words 32-bit words each, 16-byte aligned), one per (work item, 1 024-position pass), heaviest first:
//   [0] work item  [1] template index in the pack  [2] nf | query << 28  [3] number of feature words
//   [4] j0 = first position of the pass  [5] positions in the pass  [6] the item's order key  [7] 0  [8..11] per modality: 4 class sizes, u8 each
//   [12..] feature words, modality-major, grouped by class Q = (a >> 3) & 3 where a = nibble index of the window of
//   lane 0: ((a >> 1) & ~15) | (a & 7)  -- aligned chunk byte offset | nibble shift
//
// `prune` bits: bit 0 = exact early termination; bit 8 = sum the modalities in reverse order; bit 9 = pick the order per
// frame: reversed when mod_bits[f][M-1] < mod_bits[f][0] (fewer orientation bits = lower responses = earlier termination).
// dump (nullable, single frame): u16 totals [work item][dump_stride] of every scored position (parity tap; disables pruning).
struct CoarseParams {
  const uint8_t* lmn;                     // coarsest level's nibble planes of frame 0
  unsigned long long lmn_stride;          // bytes between frames
  const uint32_t* recs;
  BatchCtl* ctl;
  Cand* cand;                             // chunk-wide candidate list
  unsigned long long* touched;            // (feature, position) pairs gathered by the launch (statistics), nullable
  uint16_t* dump;
  QueryThresholds thr;
  uint32_t cand_cap;
  int rec_words, n_tiles, M, prune, dump_stride;
  int max_feat;                           // largest feature count of a tile: <= 63 selects the u8-only kernel
  int n_full;                             // records [0, n_full) are full tiles (a frame per warp); [n_full, n_tiles) are
                                          // shared tiles of at most 128 positions (eight frames per warp, four lanes each)
};
void set_programmatic_launch(bool enabled);  // per thread; disabled while launches are recorded into a CUDA graph
void set_coarse_grid_limit(int blocks);     // process-wide; 0 = no limit
void set_coarse_narrow(int mode);           // process-wide A/B switch: 0 general kernel only, 1 u8-only kernel (default), 2 u8-only at 2 CTAs/SM
int coarse_positions_per_pass();
int coarse_record_header_words();
int coarse_record_max_words();
void launch_similarity_coarse(const CoarseParams& p, int max_frames, cudaStream_t s);
// n_bytes (multiple of 16) of byte planes -> n_bytes / 2 of nibble-packed planes, for n_frames frames
void launch_pack_nibbles(const uint8_t* lm_bytes, size_t bytes_stride, uint8_t* lm_nibbles, size_t nib_stride, size_t n_bytes,
                         const BatchCtl* ctl, int n_frames, cudaStream_t s);
// Refinement of every candidate of the chunk; survivors of frame f go to the result block results + f * result_stride
// ([16 B statistics][ResultHeader][out_cap x lm_raw_match]).
void launch_refine(const RefineParams& p, const CoarseTpl* ctpl, const Cand* cand, uint32_t cand_cap, BatchCtl* ctl,
                   uint8_t* results, size_t result_stride, uint32_t out_cap, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ rendering (lm_render.cu)
struct RenderView {                       // Pc = R * Po + t, OpenCV camera convention (x right, y down, z forward)
  float R[9], t[3];
};
struct RenderCamera {
  int width, height;
  float fx, fy, cx, cy, z_near, z_max;    // z_max = 0.99 * far: fragments beyond it are dropped
};
struct RenderTargets {                    // per view v: base + v * stride (elements); null pointers are skipped
  uint8_t* bgr; uint16_t* depth; uint8_t* mask;
  size_t bgr_stride, depth_stride, mask_stride;
  int* rect;                              // [n_views][4] = x_min, y_min, x_max, y_max of the mask (x_max < 0: empty)
};
void launch_raster(const float* tris, int n_tri, const RenderView* views, int n_views, const RenderCamera& cam,
                   unsigned long long* zbuf, float* nz_abs, const RenderTargets& out, cudaStream_t s);
void launch_mask_rect(const uint8_t* mask, int W, int H, int* rect, cudaStream_t s);
void launch_depth_diff(const uint16_t* scene, int scene_cols, const uint16_t* templ, const uint8_t* tmask, int templ_cols,
                       int x, int y, int tx, int ty, int w, int h, unsigned long long* out, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ batched extraction (lm_train.cu)
struct TrainSeg {                         // one (view, level, modality) of a training batch
  uint32_t off, cap;                      // key pool region (cap: power of two >= any possible candidate count)
  uint32_t count, area;                   // device: candidates appended, pixels of the twice-eroded mask (DepthNormal)
  uint32_t per_label[8];                  // device: DepthNormal candidates per bin
  int32_t cols, nf, type, n_sel;          // level width, features wanted, LM_COLOR_GRADIENT / LM_DEPTH_NORMAL; device:
                                          // features selected, -1 = too few candidates, -2 = pool overflow
  uint32_t flags;                         // device: bit 0 = the eroded normal map holds a value that is not one-hot
                                          // (injected LUTs only): distances by ring search instead of run tables
};
struct TrainLevel {
  const uint8_t* quant;                   // unmasked quantisation of this level
  const float* mag;                       // ColorGradient only
  int rows, cols, seg, block_begin;
};
struct TrainViewParams {
  TrainLevel lv[LM_MAX_LEVELS];
  uint8_t* pb[LM_MAX_LEVELS];             // DepthNormal scratch: normal bin where the eroded mask is set, else 0
  uint16_t* runs[LM_MAX_LEVELS];          // DepthNormal scratch: horizontal distance to the nearest pixel with another value
  const uint8_t* mask0;                   // level-0 object mask
  int extract_threshold[LM_MAX_LEVELS];
  int n_levels, cols0;
  float thr_sq;                           // strong_threshold^2
};
int train_blocks(int rows, int cols);
void launch_train_cg(const TrainViewParams& p, int total_blocks, TrainSeg* segs, unsigned long long* pool, cudaStream_t s);
void launch_train_dn(const TrainViewParams& p, int total_blocks, TrainSeg* segs, unsigned long long* pool, cudaStream_t s);
void launch_train_finish(TrainSeg* segs, int n_segs, unsigned long long* pool, uint32_t* out_feats, cudaStream_t s);

}  // namespace lmk
