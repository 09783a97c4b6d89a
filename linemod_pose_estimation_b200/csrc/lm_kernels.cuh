// lm_kernels.cuh -- device-side records and kernel launchers of the LINEMOD hot path (sm_100a).
//
// Kernel <-> reference function map ([OCV] = OpenCV 2.4.x modules/objdetect/src/linemod.cpp, the code behind
// cv::linemod::Detector::match called at /root/reference/src/rgbdDetector.cpp:33):
//   production path
//   k_pyrdown_fast                              [OCV] ColorGradientPyramid::pyrDown -> cv::pyrDown            
//   k_cg_fused                                  [OCV] quantizedOrientations + hysteresisGradient, all levels (a1, a2)
//   k_dn_fused                                  [OCV] quantizedNormals + medianBlur(5) + DepthNormalPyramid::pyrDown (a4, a5)
//   k_spread_all                                [OCV] quantize(mask) + spread + computeResponseMaps + linearize (a6-a8)
//   k_similarity_coarse_rec                     [OCV] similarity + addSimilarities + matchClass coarse scan  (a9-a12)
//   k_refine_nib                                [OCV] similarityLocal + matchClass refinement loop          (a13)
//   template generation (/root/reference/src/renderer.cpp:239-329, src/rgbdDetector.cpp:147-283)
//   k_raster_tris, k_raster_resolve             RendererIterator::render / renderDepthOnly (ORK) per oracle/render_oracle.cpp
//   k_train_cg, k_train_dn_pb/dist/keys         [OCV] ColorGradientPyramid / DepthNormalPyramid::extractTemplate candidates
//   k_train_sort, k_train_select                [OCV] std::stable_sort + QuantizedPyramid::selectScatteredFeatures
//   k_depth_diff, k_mask_rect                   rgbdDetector::depth_diff; mask bounding boxes
//   A/B references and fallbacks (same results, selected with lm_set_option)
//   k_gauss7_u8c3, k_cg_grad, k_cg_hysteresis, k_pyrdown_u8c3, k_dn_normals, k_median5_u8, k_nn_half_u8, k_spread_lm
//                                               the front end stage by stage            (frontend_variant = 1)
//   k_similarity_coarse, k_similarity_coarse_nib<4>   coarse scan on byte / nibble planes without tile records (coarse_variant = 1 / 2)
//   k_refine                                    refinement on byte planes (refine_variant = 1, or rows not word-aligned)
//   k_pack_nibbles                              byte planes -> nibble planes when the spread kernel could not write them
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
  const uint8_t* lm;                      // [M][8][plane_stride] byte planes (valid when the byte kernel is used)
  const uint8_t* lmn;                     // [M][8][plane_stride / 2] nibble-packed planes (valid when the nibble kernel is used)
  const RefineTpl* tpl;
  const uint32_t* feats;                  // (x + 4096) | (y + 4096) << 13 | label << 26
  unsigned long long plane_stride;
  int rows, cols, T, W;
};

struct RefineParams {
  RefineLevel level[LM_MAX_LEVELS];       // index = pyramid level (only 0 .. L-2 used)
  int levels, M, coarse_T, coarse_W;
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
  uint32_t tglob, pos;                    // template index in the pack, raster position at the coarsest level
  uint32_t raw_nf;                        // raw score (u16) | number of features behind it << 16
  uint32_t order;                         // the work item's emission order key (query << 28 | position in the iteration)
};

struct ResultHeader {                     // zeroed before every request
  uint32_t count;                         // surviving matches written (may exceed the capacity: overflow)
  uint32_t next_tile;                     // coarse kernel's tile dispenser
  uint32_t overflow, n_cands;
};

// ------------------------------------------------------------------------------------------------ front end
void launch_gauss7_u8c3(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s);
void launch_cg_grad(const uint8_t* smoothed, int rows, int cols, float* mag, uint8_t* qunf, cudaStream_t s);
void launch_cg_hysteresis(const uint8_t* qunf, const float* mag, int rows, int cols, float threshold_sq, uint8_t* quant,
                          cudaStream_t s);
void launch_pyrdown_u8c3(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s);
void launch_dn_normals(const uint16_t* depth, int rows, int cols, int distance_threshold, int difference_threshold,
                       const uint8_t* normal_lut, uint8_t* out, cudaStream_t s);
void launch_median5_u8(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s);
void launch_nn_half_u8(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s);
// mask0: level-0 mask (nullable) sampled at (y << level, x << level); resp_all: 256 x u32, 8 response nibbles per
// spread value.  quantized_out always written; spread_out / response_out only when non-null (parity taps).
void launch_spread_lm(const uint8_t* quant_raw, const uint8_t* mask0, int mask_cols0, int level, int rows, int cols,
                      int T, const uint32_t* resp_all, uint8_t* quantized_out, uint8_t* spread_out,
                      uint8_t* response_out, uint8_t* lm, size_t plane_stride, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ fused front end
struct CgLevel {
  const uint8_t* src;  // BGR source of this level (level >= 1: output of k_pyrdown_u8c3)
  float* mag;          // [rows][cols] squared gradient magnitude (exact integer in f32)
  uint8_t* quant;      // [rows][cols] one-hot quantised orientation (unmasked)
  int rows, cols, block_begin, blocks_x;
};
struct CgParams {
  CgLevel lv[LM_MAX_LEVELS];
  int n_levels;
  float thr_sq;  // weak_threshold^2
};
struct DnParams {
  const uint16_t* depth;
  const uint8_t* lut;  // NORMAL_LUT, 8000 bytes
  uint8_t* quant[LM_MAX_LEVELS];
  int rows, cols, n_levels, distance_threshold, difference_threshold;
};
struct SpreadEntry {
  const uint8_t* qraw;   // unmasked quantisation of this (level, modality)
  const uint8_t* mask0;  // level-0 mask or null
  uint8_t* quantized;    // masked quantisation (Detector::match's quantized_images)
  uint8_t* spread;       // parity tap or null
  uint8_t* response;     // parity tap or null
  uint8_t* lm;           // this modality's 8 orientation byte planes, or null (coarsest level when only lm_nib is needed)
  uint8_t* lm_nib;       // coarsest level: this modality's 8 nibble-packed planes (plane_stride / 2 bytes each), or null
  unsigned int* bits;    // coarsest level: += orientation bits set in this modality's spread image (null: not counted);
                         // the coarse kernel starts with the modality that has fewer (lower responses, earlier pruning)
  unsigned long long plane_stride;
  int rows, cols, T, W, H, level, mask_cols0, block_begin, blocks_x;
};
struct SpreadParams {
  SpreadEntry e[LM_MAX_LEVELS * LM_MAX_MODALITIES];
  const uint32_t* resp_all;
  int n;
};
void launch_pyrdown_fast(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s);
int cg_fused_blocks(int rows, int cols, int* blocks_x);
void launch_cg_fused(const CgParams& p, int total_blocks, cudaStream_t s);
void launch_dn_fused(const DnParams& p, cudaStream_t s);
int spread_all_blocks(int W, int H, int* blocks_x);
bool launch_spread_all(const SpreadParams& p, int total_blocks, int max_T, cudaStream_t s);

// ------------------------------------------------------------------------------------------------ matching
// tiles: (work item, pass) pairs with at least one position, heaviest first.  dump (nullable): u16 totals,
// [work item][dump_stride], written for every scored position (parity tap).
// variant: 0 = nibble-packed linear memories, self-contained tile records, exact early termination (production;
// prune = 0 switches the early termination off, `touched` (nullable) accumulates the (feature, position) pairs gathered); 2 = nibble-packed, two overlapping vector loads per feature; 1 = byte linear memories.  2 and 1 are the
// A/B references and read the (items, tiles, tpl, foff) arrays; 0 reads `recs`.
// lmc: byte planes, lmn: nibble-packed planes of the coarsest level.
//
// Tile record of variant 0 (rec_words 32-bit words each, 16-byte aligned):
//   [0] work item  [1] template index in the pack  [2] nf | query << 28  [3] number of feature words
//   [4] j0 = first position of the pass  [5] positions in the pass  [6] the item's order key  [7] 0  [8..11] per modality: 4 class sizes, u8 each
//   [12..] feature words, modality-major, grouped by class Q = (a >> 3) & 3 where a = nibble index of the window of
//   lane 0: ((a >> 1) & ~15) | (a & 7)  -- aligned chunk byte offset | nibble shift
void set_programmatic_launch(bool enabled);  // per thread; disabled while launches are recorded into a CUDA graph
void set_coarse_grid_limit(int blocks);     // process-wide; 0 = no limit
int coarse_positions_per_pass(int variant);
int coarse_record_header_words();
int coarse_record_max_words();
void launch_similarity_coarse(int variant, const uint8_t* lmc, const uint8_t* lmn, const uint32_t* foff,
                              const CoarseTpl* tpl, const WorkItem* items, const uint2* tiles, const uint32_t* recs,
                              int rec_words, int n_tiles, const QueryThresholds& thr, int M, int prune, Cand* cand,
                              ResultHeader* hdr, unsigned long long* touched, uint32_t cand_cap, uint16_t* dump,
                              int dump_stride, cudaStream_t s, const unsigned int* mod_bits = nullptr);
// `prune` bits for variant 0: bit 0 = exact early termination; bit 8 = sum the modalities in reverse order; bit 9 = pick
// the order per frame: reversed when mod_bits[M-1] < mod_bits[0] (mod_bits[m] = orientation bits set in modality m's
// spread image at the coarsest level, counted by the front end: fewer bits = lower responses = earlier termination).
// n_bytes (multiple of 16) of byte planes -> n_bytes / 2 of nibble-packed planes
void launch_pack_nibbles(const uint8_t* lm_bytes, uint8_t* lm_nibbles, size_t n_bytes, cudaStream_t s);
void launch_refine(bool nibble_planes, const RefineParams& p, const CoarseTpl* ctpl, const WorkItem* items, const Cand* cand,
                   uint32_t cand_cap, ResultHeader* hdr, lm_raw_match* out, uint32_t out_cap, cudaStream_t s);

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
