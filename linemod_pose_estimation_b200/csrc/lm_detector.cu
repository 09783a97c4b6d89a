// lm_detector.cu -- host orchestration behind the C ABI (include/linemod_b200.h): workspaces, template packing,
// the per-frame kernel sequence of Detector::match and the match-list finalisation.
//
// Reference surface mirrored: cv::linemod::Detector as driven by /root/reference/src/rgbdDetector.cpp:31-34 (match),
// src/renderer.cpp:179-185,308 (construction, addTemplate), src/rgbdDetector.cpp:1668-1680 / src/renderer.cpp:56-70
// (persistence).  Everything that touches pixels runs in the CUDA kernels of lm_frontend_fused.cu / lm_match.cu; the host
// only stages buffers, packs template records and orders the (few) surviving matches.
#include "lm_detector_internal.hpp"

#include <cctype>
#include <deque>
#include <memory>

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
int lm_fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}


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

int upload_luts(lm_detector* d) {
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
  // byte 8000 behind the table: every entry is 0 or a single bit, so k_dn_fused may take medianBlur(5) by counting (it reads
  // the verdict on the device: recorded CUDA graphs stay valid across lm_set_normal_lut)
  uint8_t lut_and_flag[8004] = {0};
  std::memcpy(lut_and_flag, d->normal_lut, 8000);
  bool one_hot = d->dn_count != 0;
  for (int i = 0; i < 8000 && one_hot; ++i) one_hot = (d->normal_lut[i] & (d->normal_lut[i] - 1)) == 0;
  lut_and_flag[8000] = one_hot ? 1 : 0;
  if (d->d_normal_lut.ensure(sizeof(lut_and_flag)) != LM_OK) return LM_E_CUDA;
  CU(cudaMemcpy(d->d_resp_all.p, resp_all, sizeof(resp_all), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d->d_normal_lut.p, lut_and_flag, sizeof(lut_and_flag), cudaMemcpyHostToDevice));
  d->luts_dirty = false;
  for (int i = 0; i < LM_LANES; ++i) d->lane[i].front_valid = false;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ workspace
size_t src_row_bytes(int type, int cols) { return type == LM_8UC3 ? (size_t)cols * 3 : (type == LM_16UC1 ? (size_t)cols * 2 : (size_t)cols); }
int expected_src_type(const lm_modality_desc& m) { return m.type == LM_COLOR_GRADIENT ? LM_8UC3 : LM_16UC1; }

// Buffers needed by quantisation at (rows, cols) for `frames` frame slots; no divisibility requirements (addTemplate
// uses this alone).
int ensure_quant_ws(lm_detector* d, Lane& ln, int rows, int cols, int frames) {
  const int L = d->model.levels(), M = d->model.M();
  if (frames < 1 || frames > LM_MAX_BATCH) return lm_fail(LM_E_INVALID, "a chunk holds 1..%d frames", LM_MAX_BATCH);
  if (ln.rows != rows || ln.cols != cols || ln.frames < frames) { ln.lm_ready = false; ln.front_valid = false; }
  frames = std::max(frames, ln.rows == rows && ln.cols == cols ? ln.frames : 0);
  bool grew = false, g = false;
  for (int m = 0; m < M; ++m) {
    const bool cg = d->model.mods[m].type == LM_COLOR_GRADIENT;
    if (ln.src[m].ensure(src_row_bytes(expected_src_type(d->model.mods[m]), cols) * rows, frames, &g) != LM_OK) return LM_E_CUDA;
    grew = grew || g;
    for (int l = 0; l < L; ++l) {
      size_t n = (size_t)(rows >> l) * (cols >> l);
      if (n == 0) return lm_fail(LM_E_INVALID, "image too small for %d pyramid levels", L);
      if (cg && l > 0) { if (ln.bgr[l][m].ensure(n * 3, frames, &g) != LM_OK) return LM_E_CUDA; grew = grew || g; }
      if (cg) { if (ln.mag[l][m].ensure(n * sizeof(float), frames, &g) != LM_OK) return LM_E_CUDA; grew = grew || g; }
      if (ln.quant_raw[l][m].ensure(n, frames, &g) != LM_OK) return LM_E_CUDA;
      grew = grew || g;
      if (ln.quantized[l][m].ensure(n, frames, &g) != LM_OK) return LM_E_CUDA;
      grew = grew || g;
    }
  }
  if (grew) ln.drop_graphs();  // recorded launches hold the old addresses
  ln.rows = rows; ln.cols = cols; ln.frames = frames;
  return LM_OK;
}

// Everything enqueued with this lane's buffers is done: its own streams and the caller's stream of the last device-resident
// chunk.  (Not cudaDeviceSynchronize: another handle on the same device may be recording a CUDA graph in another thread,
// and a device-wide synchronisation is illegal while any stream captures.)
static int lane_quiesce(Lane& ln) {
  if (ln.stream) CU(cudaStreamSynchronize(ln.stream));
  for (int i = 0; i < LM_MAX_MODALITIES - 1; ++i)
    if (ln.side[i]) CU(cudaStreamSynchronize(ln.side[i]));
  if (ln.user_stream_valid) CU(cudaStreamSynchronize(ln.user_stream));
  return LM_OK;
}

// Does this request need the reference's byte planes of a level (besides the nibble-packed ones the kernels read)?
static bool level_needs_bytes(const lm_detector* d, const LevelGeom& g) { return d->debug_taps != 0 || !level_nibble_aligned(g); }

// Linear-memory buffers for matching at (rows, cols): enforces the reference's CV_Asserts on the geometry.
static int ensure_lm_ws(lm_detector* d, Lane& ln, int rows, int cols, int frames) {
  const int L = d->model.levels(), M = d->model.M();
  bool bytes_ok = true;
  // refinement levels (every level above the coarsest) are column-blocked when their geometry allows it
  auto tiled_level = [&](int l, int W, int H) { return d->refine_tiled != 0 && l < L - 1 && W % 16 == 0 && H >= 16; };
  if (ln.lm_ready && ln.rows == rows && ln.cols == cols && (int)ln.geom.size() == L)
    for (int l = 0; l < L; ++l) {
      if (level_needs_bytes(d, ln.geom[l]) && ln.lmem[l].frames < (level_nibble_aligned(ln.geom[l]) ? 1 : ln.frames)) bytes_ok = false;
      if ((ln.geom[l].Hh != 0) != tiled_level(l, ln.geom[l].W, ln.geom[l].H)) bytes_ok = false;   // option changed: rebuild
    }
  if (ln.lm_ready && ln.rows == rows && ln.cols == cols && (int)ln.geom.size() == L && ln.frames >= frames && bytes_ok) return LM_OK;
  std::vector<LevelGeom> geom(L);
  for (int l = 0; l < L; ++l) {
    LevelGeom& g = geom[l];
    g.rows = rows >> l; g.cols = cols >> l; g.T = d->model.T[l];
    if (g.T < 1 || g.T > 16) return lm_fail(LM_E_INVALID, "unsupported T=%d at level %d (1..16)", g.T, l);
    if (g.rows <= 0 || g.cols <= 0) return lm_fail(LM_E_INVALID, "image too small for %d pyramid levels", L);
    if (((size_t)g.rows * g.cols) % 16 != 0)
      return lm_fail(LM_E_INVALID, "(rows * cols) %% 16 != 0 at level %d (%dx%d)", l, g.cols, g.rows);  // computeResponseMaps
    if (g.rows % g.T != 0 || g.cols % g.T != 0)
      return lm_fail(LM_E_INVALID, "rows %% T != 0 or cols %% T != 0 at level %d (%dx%d, T=%d)", l, g.cols, g.rows, g.T);  // linearize
    if (g.cols > 4095 || g.rows > 4095) return lm_fail(LM_E_INVALID, "images larger than 4095 px are not supported");
    g.W = g.cols / g.T; g.H = g.rows / g.T;
    g.plane_stride = plane_stride_of(g.T, g.W, g.H);
    g.Hh = tiled_level(l, g.W, g.H) ? g.H + 16 : 0;
    g.nib_plane = g.Hh ? lmk::tiled_plane_stride(g.T, g.W, g.H) : g.plane_stride;
  }
  // everything that reads or wrote the old buffers must be done (callers' streams included) before they move
  if (lane_quiesce(ln) != LM_OK) return LM_E_CUDA;
  if (ensure_quant_ws(d, ln, rows, cols, frames) != LM_OK) return LM_E_CUDA;
  frames = ln.frames;
  ln.drop_graphs();
  for (int l = 0; l < L; ++l) {
    const size_t bytes = (size_t)M * 8 * geom[l].plane_stride + kLmSlack;
    if (level_needs_bytes(d, geom[l])) {
      // parity taps look at frame 0 only; unaligned levels are packed from byte planes for every frame
      if (ln.lmem[l].ensure(bytes, level_nibble_aligned(geom[l]) ? 1 : frames) != LM_OK) return LM_E_CUDA;
      CU(cudaMemsetAsync(ln.lmem[l].buf.p, 0, ln.lmem[l].bytes(), ln.stream));  // zero tails (and slack) once per geometry
    }
    if (ln.lmn[l].ensure(((size_t)M * 8 * geom[l].nib_plane + kLmSlack) / 2, frames) != LM_OK) return LM_E_CUDA;
    CU(cudaMemsetAsync(ln.lmn[l].buf.p, 0, ln.lmn[l].bytes(), ln.stream));   // zero tails, halo rows of the last phase, slack
  }
  CU(cudaStreamSynchronize(ln.stream));
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
// cudaPointerGetAttributes costs a microsecond or two and the batched paths ask four times per frame, nearly always about the
// same few buffers: a small per-thread cache in front of it.  A stale entry (the address freed and handed out again with the
// other kind of memory) only picks the other copy path -- cudaMemcpyAsync is correct from pageable memory too, and staging
// pinned memory is merely slower.
bool is_pinned(const void* p) {
  struct Entry { const void* p; bool pinned; };
  static thread_local Entry cache[64] = {};
  Entry& e = cache[(reinterpret_cast<uintptr_t>(p) >> 12) & 63u];
  if (p != nullptr && e.p == p) return e.pinned;
  cudaPointerAttributes a;
  bool pinned = false;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) cudaGetLastError();
  else pinned = a.type == cudaMemoryTypeHost;
  e.p = p; e.pinned = pinned;
  return pinned;
}

// Host image -> tightly packed device buffer.  Pinned sources go straight to the copy engine; pageable ones are
// packed into the lane's pinned staging area first (offset *stage_off, advanced).  The staging area belongs to the lane's
// current chunk: callers wait for the lane's previous chunk before they upload the next one.
int upload_image(Lane& ln, const lm_image& im, void* dst, size_t* stage_off) {
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
  if (n_sources != M) return lm_fail(LM_E_INVALID, "sources.size() (%d) != modalities.size() (%d)", n_sources, M);
  if (n_masks != 0 && n_masks != M) return lm_fail(LM_E_INVALID, "masks.size() (%d) != modalities.size() (%d)", n_masks, M);
  for (int m = 0; m < M; ++m) {
    if (!sources[m].data) return lm_fail(LM_E_INVALID, "source %d is empty", m);
    if (sources[m].type != expected_src_type(d->model.mods[m]))
      return lm_fail(LM_E_INVALID, "source %d: %s needs a %s image", m, modality_name(d->model.mods[m].type),
                  d->model.mods[m].type == LM_COLOR_GRADIENT ? "CV_8UC3" : "CV_16UC1");
    if (sources[m].rows != sources[0].rows || sources[m].cols != sources[0].cols)
      return lm_fail(LM_E_INVALID, "sources differ in size");
    if (sources[m].step < src_row_bytes(sources[m].type, sources[m].cols)) return lm_fail(LM_E_INVALID, "source %d: step too small", m);
    if (n_masks && masks[m].data) {
      if (masks[m].type != LM_8UC1 || masks[m].rows != sources[m].rows || masks[m].cols != sources[m].cols)
        return lm_fail(LM_E_INVALID, "mask %d: size/type mismatch (mask.size() == source.size())", m);
    }
  }
  return LM_OK;
}

// Host frames (sources[f * M + m]) -> the lane's source slots 0 .. n_frames-1; masks for single-frame requests only.
static int upload_frames(lm_detector* d, Lane& ln, const lm_image* sources, int n_frames, const lm_image* masks, int n_masks) {
  const int M = d->model.M();
  size_t need = 0;
  for (int f = 0; f < n_frames; ++f)
    for (int m = 0; m < M; ++m) {
      const lm_image& im = sources[(size_t)f * M + m];
      if (!is_pinned(im.data)) need += ((src_row_bytes(im.type, im.cols) * im.rows) + 255) & ~(size_t)255;
    }
  for (int m = 0; m < M && n_masks; ++m)
    if (masks[m].data) need += (((size_t)masks[m].rows * masks[m].cols) + 255) & ~(size_t)255;
  if (need && ln.stage_in.ensure(need) != LM_OK) return LM_E_CUDA;
  size_t off = 0;
  for (int f = 0; f < n_frames; ++f)
    for (int m = 0; m < M; ++m) {
      if (upload_image(ln, sources[(size_t)f * M + m], ln.src[m].as<uint8_t>(f), &off) != LM_OK) return LM_E_CUDA;
      ln.src_ptr[f][m] = ln.src[m].as<uint8_t>(f);
    }
  for (int m = 0; m < M; ++m) {
    ln.has_mask[m] = n_masks && masks[m].data;
    if (ln.has_mask[m]) {
      if (ln.mask0[m].ensure((size_t)masks[m].rows * masks[m].cols) != LM_OK) return LM_E_CUDA;
      if (upload_image(ln, masks[m], ln.mask0[m].p, &off) != LM_OK) return LM_E_CUDA;
    }
  }
  ln.n_frames = n_frames;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ front end
int begin_chunk(lm_detector* d, Lane& ln, int n_frames, int result_blocks, cudaStream_t s) {
  FrameTable ft;
  std::memset(&ft, 0, sizeof(ft));
  ft.n_frames = n_frames;
  for (int f = 0; f < n_frames; ++f)
    for (int m = 0; m < d->model.M(); ++m) ft.src[f][m] = ln.src_ptr[f][m];
  launch_begin_chunk(ft, ln.ctl.as<BatchCtl>(), ln.result.as<uint8_t>(), ln.result.stride, ln.result.buf.p ? result_blocks : 0, s);
  ln.n_frames = n_frames;
  ++ln.launches;
  CU(cudaGetLastError());
  return LM_OK;
}

// [OCV] Modality::process + QuantizedPyramid::pyrDown for every level: fills quant_raw[l][m] (and mag[l][m]) of every
// frame of the chunk.  Per ColorGradient modality the pyrDown chain plus ONE launch covering every level, per DepthNormal
// modality ONE launch; the modalities run concurrently on forked streams.
int run_quantize(lm_detector* d, Lane& ln, int grid_frames, cudaStream_t main_stream) {
  const int L = d->model.levels(), M = d->model.M();
  const BatchCtl* ctl = ln.ctl.as<BatchCtl>();
  if (M > 1) CU(cudaEventRecord(ln.ev_fork, main_stream));
  for (int m = 0; m < M; ++m) {
    const lm_modality_desc& md = d->model.mods[m];
    cudaStream_t s = m == 0 ? main_stream : ln.side[m - 1];
    if (m > 0) CU(cudaStreamWaitEvent(s, ln.ev_fork, 0));
    if (md.type == LM_COLOR_GRADIENT) {
      CgParams cp;
      std::memset(&cp, 0, sizeof(cp));
      cp.ctl = ctl; cp.n_levels = L; cp.modality = m;
      cp.thr_sq = md.weak_threshold * md.weak_threshold;
      int total = 0;
      for (int l = 0; l < L; ++l) {
        const int rows = ln.rows >> l, cols = ln.cols >> l;
        if (l > 0) {
          launch_pyrdown_fast(ctl, m, l == 1 ? nullptr : ln.bgr[l - 1][m].as<uint8_t>(), l == 1 ? 0 : ln.bgr[l - 1][m].stride,
                              ln.rows >> (l - 1), ln.cols >> (l - 1), ln.bgr[l][m].as<uint8_t>(), ln.bgr[l][m].stride, grid_frames, s);
          ++ln.launches;
        }
        CgLevel& lv = cp.lv[l];
        lv.src = l == 0 ? nullptr : ln.bgr[l][m].as<uint8_t>();
        lv.src_stride = l == 0 ? 0 : ln.bgr[l][m].stride;
        lv.mag = ln.mag[l][m].as<float>(); lv.mag_stride = ln.mag[l][m].stride_in<float>();
        lv.quant = ln.quant_raw[l][m].as<uint8_t>(); lv.quant_stride = ln.quant_raw[l][m].stride;
        lv.rows = rows; lv.cols = cols; lv.block_begin = total;
        total += cg_fused_blocks(rows, cols, &lv.blocks_x);
      }
      launch_cg_fused(cp, total, grid_frames, s);
      ++ln.launches;
    } else {
      DnParams dp;
      std::memset(&dp, 0, sizeof(dp));
      dp.ctl = ctl; dp.modality = m;
      dp.lut = d->d_normal_lut.as<uint8_t>();
      dp.rows = ln.rows; dp.cols = ln.cols; dp.n_levels = L;
      dp.distance_threshold = md.distance_threshold; dp.difference_threshold = md.difference_threshold;
      for (int l = 0; l < L; ++l) { dp.quant[l] = ln.quant_raw[l][m].as<uint8_t>(); dp.quant_stride[l] = ln.quant_raw[l][m].stride; }
      launch_dn_fused(dp, grid_frames, s);
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

// [OCV] Detector::match front half: quantize (mask) -> spread -> computeResponseMaps -> linearize per level/modality,
// for grid_frames frame slots of the lane (the frame table's count decides which of them do anything).
static int run_front(lm_detector* d, Lane& ln, int grid_frames, cudaStream_t s) {
  const int L = d->model.levels(), M = d->model.M();
  if (run_quantize(d, ln, grid_frames, s) != LM_OK) return LM_E_CUDA;
  const bool taps = d->debug_taps != 0;
  if (taps && ensure_tap_ws(d, ln) != LM_OK) return LM_E_CUDA;
  SpreadParams sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.resp_all = d->d_resp_all.as<uint32_t>();
  sp.ctl = ln.ctl.as<BatchCtl>();
  int total = 0, max_T = 1;
  bool direct[LM_MAX_LEVELS] = {false, false, false, false}, byt[LM_MAX_LEVELS] = {false, false, false, false};
  for (int l = 0; l < L; ++l) {
    const LevelGeom& g = ln.geom[l];
    max_T = std::max(max_T, g.T);
    // nibble planes are written directly when every (orientation, phase) row starts on a word; otherwise the byte planes
    // are written and packed.  The byte planes are also written for the parity taps (frame 0 of the chunk only then).
    direct[l] = level_nibble_aligned(g);
    byt[l] = taps || !direct[l];
    for (int m = 0; m < M; ++m) {
      SpreadEntry& e = sp.e[sp.n++];
      e.qraw = ln.quant_raw[l][m].as<uint8_t>(); e.qraw_stride = ln.quant_raw[l][m].stride;
      e.mask0 = ln.has_mask[m] ? ln.mask0[m].as<uint8_t>() : nullptr;
      e.quantized = ln.quantized[l][m].as<uint8_t>(); e.quantized_stride = ln.quantized[l][m].stride;
      e.spread = taps ? ln.spread[l][m].as<uint8_t>() : nullptr;
      e.response = taps ? ln.response[l][m].as<uint8_t>() : nullptr;
      e.lm = byt[l] ? ln.lmem[l].as<uint8_t>() + (size_t)m * 8 * g.plane_stride : nullptr;
      e.lm_stride = ln.lmem[l].frames > 1 ? ln.lmem[l].stride : 0;
      e.lm_nib = direct[l] ? ln.lmn[l].as<uint8_t>() + (size_t)m * 4 * g.nib_plane : nullptr;
      e.lm_nib_stride = ln.lmn[l].stride;
      e.nib_plane = g.nib_plane; e.tiled_Hh = g.Hh;
      e.count_bits = l == L - 1;
      e.plane_stride = g.plane_stride;
      e.rows = g.rows; e.cols = g.cols; e.T = g.T; e.W = g.W; e.H = g.H; e.level = l; e.modality = m; e.mask_cols0 = ln.cols;
      e.block_begin = total;
      total += spread_all_blocks(g.T, g.W, g.H, &e.blocks_x);
    }
  }
  // byte planes that hold one frame only (taps on an aligned level) can be written by a single-frame launch only
  for (int l = 0; l < L; ++l)
    if (byt[l] && ln.lmem[l].frames < grid_frames && grid_frames > 1)
      return lm_fail(LM_E_STATE, "parity taps are available for single-frame requests only");
  if (!launch_spread_all(sp, total, grid_frames, s)) return lm_fail(LM_E_INVALID, "T=%d needs too much shared memory", max_T);
  ++ln.launches;
  for (int l = 0; l < L; ++l) {
    if (!direct[l]) {
      launch_pack_nibbles(ln.lmem[l].as<uint8_t>(), ln.lmem[l].stride, ln.lmn[l].as<uint8_t>(), ln.lmn[l].stride,
                          (size_t)M * 8 * ln.geom[l].plane_stride, ln.ctl.as<BatchCtl>(), grid_frames, s);
      ++ln.launches;
    }
    ln.bytes_valid[l] = byt[l];
    ln.nibbles_valid[l] = true;
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
      pk.shard_world == d->shard_world)
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
      if ((int)tp.size() != L * M) return lm_fail(LM_E_INVALID, "template pyramid of class '%s' has %zu templates, expected %d", it->first.c_str(), tp.size(), L * M);
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
        else if (P != P_all) return lm_fail(LM_E_INVALID, "class '%s' template %zu: modalities disagree on width/height", it->first.c_str(), tid);
        std::vector<uint32_t> grp[4];
        for (const Feature& f : t.features) {
          if (f.x < 0 || f.x >= gc.cols || f.y < 0 || f.y >= gc.rows) continue;  // "Discard feature if out of bounds"
          size_t a = (size_t)m * 8 * gc.plane_stride + (size_t)f.label * gc.plane_stride +
                     (size_t)((f.y % gc.T) * gc.T + (f.x % gc.T)) * ((size_t)gc.W * gc.H) + (size_t)(f.y / gc.T) * gc.W + f.x / gc.T;
          // window-alignment class of the feature (a compile-time constant in the kernel): word offset of the window in
          // its 16-byte chunk of the nibble-packed plane, a = nibble index
          const int q = (int)((a >> 3) & 3);
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
  if (up(pk.ctpl, ctpl.data(), ctpl.size() * sizeof(CoarseTpl)) != LM_OK) return LM_E_CUDA;
  for (int l = 0; l < L - 1; ++l) {
    if (up(pk.rtpl[l], rtpl[l].data(), rtpl[l].size() * sizeof(RefineTpl)) != LM_OK) return LM_E_CUDA;
    if (up(pk.rfeats[l], rfeats[l].data(), rfeats[l].size() * 4) != LM_OK) return LM_E_CUDA;
  }
  pk.h_ctpl.swap(ctpl);
  pk.h_foff.swap(foff);
  pk.version = md.version; pk.rows = ln.rows; pk.cols = ln.cols;
  pk.shard_rank = d->shard_rank; pk.shard_world = d->shard_world;
  return LM_OK;
}

// ------------------------------------------------------------------------------------------------ matching
struct Query {
  float threshold;
  const char* const* class_ids;
  int n_ids;
};

static const int kMaxQueries = LM_MAX_QUERIES;  // (class list, threshold) queries answered from one front end

// Self-contained tile records of the production coarse kernel (layout: lm_kernels.cuh).
static int build_tile_records(const Pack& pk, const std::vector<WorkItem>& items, const std::vector<uint2>& tiles,
                              int pass_pos, int M, std::vector<uint32_t>& recs, int* rec_words, int* max_feat_out) {
  const int hdr = coarse_record_header_words();
  int max_feat = 0;
  for (const WorkItem& it : items) {
    const CoarseTpl& ct = pk.h_ctpl[it.tglob];
    int n = 0;
    for (int m = 0; m < M; ++m) for (int g = 0; g < 4; ++g) n += ct.cnt[m][g];
    max_feat = std::max(max_feat, n);
  }
  const int words = (hdr + max_feat + 3) & ~3;
  if (words > coarse_record_max_words()) return lm_fail(LM_E_INVALID, "template with too many features for a tile record");
  *rec_words = words;
  *max_feat_out = max_feat;
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
      if (!qs[q].class_ids[i]) return lm_fail(LM_E_INVALID, "class_ids[%d] is NULL", i);
      key += qs[q].class_ids[i];
      key += '\n';
    }
    key += '\x01';
  }
  auto hit = pk.plans.find(key);
  if (hit != pk.plans.end()) { *out = &hit->second; return LM_OK; }
  std::vector<WorkItem> items;
  struct Tile { uint2 t; uint64_t cost; int positions; };
  std::vector<Tile> tiles;
  Pack::Plan plan;
  const int pass_pos = coarse_positions_per_pass();
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
        t.positions = std::min(pass_pos, ct.P - pass * pass_pos);
        t.cost = nfeat * (uint64_t)t.positions;
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
  if (items.size() >= (1u << 28)) return lm_fail(LM_E_INVALID, "too many templates in one request");
  // full tiles first, heaviest first; then the tiles of at most 128 positions (the tail pass of templates whose span exceeds
  // one pass), which the coarse kernel scores for eight frames per warp
  const bool share = d->coarse_share != 0;
  std::stable_sort(tiles.begin(), tiles.end(), [share](const Tile& a, const Tile& b) {
    const bool sa = share && a.positions <= 128, sb = share && b.positions <= 128;
    if (sa != sb) return sb;
    return a.cost > b.cost;
  });
  std::vector<uint2> tl(tiles.size());
  for (size_t i = 0; i < tiles.size(); ++i) {
    tl[i] = tiles[i].t;
    if (!(share && tiles[i].positions <= 128)) plan.n_full = (int)i + 1;
  }
  plan.n_items = (int)items.size();
  plan.n_tiles = (int)tl.size();
  plan.evals = (uint64_t)items.size();
  std::vector<uint32_t> recs;
  if (build_tile_records(pk, items, tl, pass_pos, d->model.M(), recs, &plan.rec_words, &plan.max_feat) != LM_OK) return LM_E_INVALID;
  Pack::Plan& dst = pk.plans[key];
  dst = plan;
  if (dst.recs.ensure(recs.size() * 4 + 64) != LM_OK) return LM_E_CUDA;
  if (!recs.empty()) CU(cudaMemcpy(dst.recs.p, recs.data(), recs.size() * 4, cudaMemcpyHostToDevice));
  *out = &dst;
  return LM_OK;
}

// Result block of a frame in device memory: [16 B kernel statistics][ResultHeader][out_cap x lm_raw_match].  The
// statistics (u64 (feature, position) pairs the coarse kernel gathered for the whole chunk, in frame 0's block) sit in
// front so that k_begin_chunk clears both and one D2H copy brings both; the public block (lm_match_device) starts at the
// header.  The blocks of a lane's frames are ln.result.stride apart.
static const size_t kStatsBytes = 16;
static size_t result_bytes(const Lane& ln) { return sizeof(ResultHeader) + (size_t)ln.out_cap * sizeof(lm_raw_match); }
static uint8_t* block_ptr(const Lane& ln, int frame = 0) { return ln.result.as<uint8_t>(frame) + kStatsBytes; }
// Records fetched together with the header in one D2H copy: 256 to start with, then what the lane's recent frames needed
// (ln.head_records follows the largest survivor count seen, so that a workload with long lists settles on one copy per chunk).
static size_t head_records(const Lane& ln) { return std::min<size_t>(std::max<uint32_t>(ln.head_records, 256u), ln.out_cap); }
static size_t head_bytes(const Lane& ln) {
  return (kStatsBytes + sizeof(ResultHeader) + head_records(ln) * sizeof(lm_raw_match) + 255) & ~(size_t)255;
}

static int ensure_match_buffers(lm_detector* d, Lane& ln, uint32_t cand_cap, uint32_t out_cap, int frames) {
  if (cand_cap > ln.cand_cap) {
    if (lane_quiesce(ln) != LM_OK) return LM_E_CUDA;  // a previous chunk (possibly on a caller's stream) may still read the list
    if (ln.cand.ensure((size_t)cand_cap * sizeof(Cand)) != LM_OK) return LM_E_CUDA;
    ln.cand_cap = cand_cap;
    ln.drop_graphs();
  }
  out_cap = std::max(out_cap, ln.out_cap);
  if (out_cap > ln.out_cap || ln.result.frames < frames) {
    if (lane_quiesce(ln) != LM_OK) return LM_E_CUDA;
    bool grew = false;
    if (ln.result.ensure(kStatsBytes + sizeof(ResultHeader) + (size_t)out_cap * sizeof(lm_raw_match), frames, &grew) != LM_OK) return LM_E_CUDA;
    ln.out_cap = out_cap;
    CU(cudaMemset(ln.result.buf.p, 0, ln.result.bytes()));
    ln.drop_graphs();
  }
  // pinned staging for the heads of a whole chunk at the largest head size (the head follows the survivor counts)
  if (ln.stage_out.ensure((kStatsBytes + result_bytes(ln) + 256) * (size_t)ln.result.frames) != LM_OK) return LM_E_CUDA;
  return LM_OK;
}

// Enqueues the chunk's coarse similarity + refinement on stream s: one launch each, whatever the number of frames and
// queries.
static int enqueue_match(lm_detector* d, Lane& ln, const Pack::Plan& plan, const Query* qs, int n_q, int grid_frames,
                         cudaStream_t s, cudaEvent_t ev_mid) {
  const HostModel& md = d->model;
  const int L = md.levels(), M = md.M();
  Pack& pk = d->pack;
  const LevelGeom& gc = ln.geom[L - 1];
  CoarseParams cp;
  RefineParams rp;
  std::memset(&rp, 0, sizeof(rp));
  std::memset(&cp, 0, sizeof(cp));
  for (int q = 0; q < n_q; ++q) { cp.thr.v[q] = qs[q].threshold; rp.threshold[q] = qs[q].threshold; }
  cp.lmn = ln.lmn[L - 1].as<uint8_t>(); cp.lmn_stride = ln.lmn[L - 1].stride;
  cp.recs = plan.recs.as<uint32_t>(); cp.rec_words = plan.rec_words; cp.n_tiles = plan.n_tiles; cp.max_feat = plan.max_feat; cp.n_full = plan.n_full;
  cp.ctl = ln.ctl.as<BatchCtl>();
  cp.cand = ln.cand.as<Cand>(); cp.cand_cap = ln.cand_cap;
  cp.touched = ln.result.as<unsigned long long>();
  cp.M = M; cp.prune = (d->prune & 1) | (d->mod_order << 8);
  launch_similarity_coarse(cp, grid_frames, s);
  if (plan.n_tiles > 0) ++ln.launches;
  if (ev_mid) CU(cudaEventRecord(ev_mid, s));
  rp.levels = L; rp.M = M; rp.coarse_T = gc.T; rp.coarse_W = gc.W;
  rp.prune = (d->prune & 2) != 0;
  rp.mod_order = d->mod_order == 1 ? 1 : 0;  // measured: the coarse kernel's per-frame choice (2) costs the refinement 40 %
  for (int l = 0; l < L - 1; ++l) {
    const LevelGeom& g = ln.geom[l];
    rp.level[l].lmn = ln.lmn[l].as<uint8_t>();
    rp.level[l].frame_stride = ln.lmn[l].stride;
    rp.level[l].tpl = pk.rtpl[l].as<RefineTpl>();
    rp.level[l].feats = pk.rfeats[l].as<uint32_t>();
    rp.level[l].plane_stride = g.nib_plane;
    rp.level[l].rows = g.rows; rp.level[l].cols = g.cols; rp.level[l].T = g.T; rp.level[l].W = g.W; rp.level[l].Hh = g.Hh;
    rp.level[l].inv_T = g.T >= 2 ? 0xffffffffu / (unsigned)g.T + 1u : 0u;
  }
  launch_refine(rp, pk.ctpl.as<CoarseTpl>(), ln.cand.as<Cand>(), ln.cand_cap, ln.ctl.as<BatchCtl>(), ln.result.as<uint8_t>(),
                ln.result.stride, ln.out_cap, s);
  ++ln.launches;
  CU(cudaGetLastError());
  return LM_OK;
}

static int graph_slot_of(int n_frames, int* grid_frames) {
  int slot = 0, g = 1;
  while (g < n_frames) { g <<= 1; ++slot; }
  *grid_frames = g;
  return slot;
}

// Front end + matching of the lane's current chunk (ln.src_ptr[0 .. n_frames), already in device memory), enqueued on s:
// k_begin_chunk, then the lane's CUDA graph for this launch geometry (recorded on first use), or plain launches when
// graphs are switched off, the parity taps or per-stage timing are on, masks are used, or a capture ever failed here.
static int enqueue_chunk(lm_detector* d, Lane& ln, const Pack::Plan& plan, const Query* qs, int n_q, int n_frames, cudaStream_t s,
                         bool stage_events = false) {
  bool masks = false;
  for (int m = 0; m < d->model.M(); ++m) masks = masks || ln.has_mask[m];
  int grid_frames = 1;
  const int slot = graph_slot_of(n_frames, &grid_frames);
  grid_frames = std::min(grid_frames, ln.frames);
  ln.launches = 0;
  if (begin_chunk(d, ln, n_frames, n_frames, s) != LM_OK) return LM_E_CUDA;
  if (!d->graphs || ln.graph_broken || d->debug_taps || masks || stage_events || slot >= LM_GRAPH_SLOTS) {
    if (stage_events) CU(cudaEventRecord(ln.ev[1], s));
    if (run_front(d, ln, n_frames, s) != LM_OK) return LM_E_CUDA;
    if (stage_events) CU(cudaEventRecord(ln.ev[2], s));
    if (enqueue_match(d, ln, plan, qs, n_q, n_frames, s, stage_events ? ln.ev[3] : nullptr) != LM_OK) return LM_E_CUDA;
    if (stage_events) CU(cudaEventRecord(ln.ev[4], s));
    return LM_OK;
  }
  Lane::GraphSlot& gs = ln.graph[slot];
  Lane::GraphKey key;
  std::memset(&key, 0, sizeof(key));
  key.plan = &plan; key.plan_recs = plan.recs.p; key.n_tiles = plan.n_tiles; key.cand = ln.cand.p; key.result = ln.result.buf.p;
  key.lmn0 = ln.lmn[0].buf.p; key.ctl = ln.ctl.p; key.ws_frames = ln.frames;
  key.shard_rank = d->shard_rank; key.shard_world = d->shard_world;
  key.model_version = d->model.version; key.rows = ln.rows; key.cols = ln.cols; key.n_q = n_q;
  key.prune = d->prune | (d->mod_order << 8);
  key.cand_cap = ln.cand_cap; key.out_cap = ln.out_cap;
  for (int q = 0; q < n_q; ++q) key.thr[q] = qs[q].threshold;
  if (gs.exec == nullptr || std::memcmp(&key, &gs.key, sizeof(key)) != 0) {
    if (gs.exec) { cudaGraphExecDestroy(gs.exec); gs.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    set_programmatic_launch(false);
    const int launches_before = ln.launches;
    cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
    int rc = LM_OK;
    if (e == cudaSuccess) {
      rc = run_front(d, ln, grid_frames, s);
      if (rc == LM_OK) rc = enqueue_match(d, ln, plan, qs, n_q, grid_frames, s, nullptr);
      e = cudaStreamEndCapture(s, &graph);
    }
    set_programmatic_launch(true);
    if (e == cudaSuccess && rc == LM_OK && graph != nullptr) e = cudaGraphInstantiate(&gs.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc != LM_OK || gs.exec == nullptr) {  // not fatal: this lane keeps to plain launches
      cudaGetLastError();
      gs.exec = nullptr;
      ln.graph_broken = true;
      ln.launches = launches_before;
      if (run_front(d, ln, n_frames, s) != LM_OK) return LM_E_CUDA;
      return enqueue_match(d, ln, plan, qs, n_q, n_frames, s, nullptr);
    }
    gs.key = key;
    gs.launches = ln.launches - launches_before;
    ln.launches = launches_before;
  }
  CU(cudaGraphLaunch(gs.exec, s));
  ln.launches += gs.launches;
  ln.front_valid = true;
  ln.debug_taps_written = false;
  for (size_t l = 0; l < ln.geom.size(); ++l) { ln.nibbles_valid[l] = true; ln.bytes_valid[l] = !level_nibble_aligned(ln.geom[l]); }
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
  // Emission order: (order key, coarse raster position) is unique per record of a query, so any sort gives the same order;
  // 16-byte (key, index) pairs sort faster than the 32-byte records themselves (long lists: BASELINE config 5).
  if (raw.size() > 64) {
    std::vector<std::pair<uint64_t, uint32_t> > keys(raw.size());
    for (size_t i = 0; i < raw.size(); ++i) keys[i] = std::make_pair(((uint64_t)raw[i].order_key << 32) | raw[i].coarse_pos, (uint32_t)i);
    std::sort(keys.begin(), keys.end());
    std::vector<lm_raw_match> ordered(raw.size());
    for (size_t i = 0; i < raw.size(); ++i) ordered[i] = raw[keys[i].second];
    raw.swap(ordered);
  } else {
    std::sort(raw.begin(), raw.end(), [](const lm_raw_match& a, const lm_raw_match& b) {
      return a.order_key != b.order_key ? a.order_key < b.order_key : a.coarse_pos < b.coarse_pos;
    });
  }
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

// Result download, part 1: (statistics + header + leading records) of every frame of the chunk into the lane's pinned
// staging block, one strided copy.  The pipelined paths enqueue it right behind the chunk's kernels, so that by the time
// the host comes back to this lane the records are already there.
static int enqueue_download(Lane& ln, int n_frames, cudaStream_t s) {
  const size_t hb = head_bytes(ln);
  if (n_frames == 1) CU(cudaMemcpyAsync(ln.stage_out.p, ln.result.buf.p, hb, cudaMemcpyDeviceToHost, s));
  else CU(cudaMemcpy2DAsync(ln.stage_out.p, hb, ln.result.buf.p, ln.result.stride, hb, (size_t)n_frames, cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(ln.ev[5], s));
  return LM_OK;
}

// Result download, part 2 (after ln.ev[5]): the records of the chunk's n frames out of the staging block.  Frames with
// more survivors than the head holds get the rest with one extra copy each, all enqueued before ONE synchronisation.
struct FrameRecords {
  std::vector<lm_raw_match> raw;
  bool overflow = false;
  uint32_t n_cands = 0;
};
static int collect_chunk(Lane& ln, int n, cudaStream_t s, std::vector<FrameRecords>& out) {
  const size_t first = head_records(ln);
  const size_t hb = head_bytes(ln);
  uint32_t longest = 0;
  out.resize((size_t)n);
  ln.work_stats[6] = *reinterpret_cast<const unsigned long long*>(ln.stage_out.as<uint8_t>());
  bool extra = false;
  for (int f = 0; f < n; ++f) {
    const uint8_t* host = ln.stage_out.as<uint8_t>() + (size_t)f * hb + kStatsBytes;
    ResultHeader h;
    std::memcpy(&h, host, sizeof(h));
    FrameRecords& fr = out[(size_t)f];
    fr.overflow = h.overflow != 0 || h.count > ln.out_cap;
    fr.n_cands = h.n_cands;
    fr.raw.clear();
    longest = std::max(longest, std::min(h.count, ln.out_cap));
    if (fr.overflow) continue;
    const lm_raw_match* recs = reinterpret_cast<const lm_raw_match*>(host + sizeof(ResultHeader));
    fr.raw.assign(recs, recs + std::min<size_t>(h.count, first));
    if (h.count > first) {
      fr.raw.resize(h.count);
      CU(cudaMemcpyAsync(fr.raw.data() + first, block_ptr(ln, f) + sizeof(ResultHeader) + first * sizeof(lm_raw_match),
                         (h.count - first) * sizeof(lm_raw_match), cudaMemcpyDeviceToHost, s));
      extra = true;
    }
  }
  if (extra) CU(cudaStreamSynchronize(s));
  // the next chunk's head: room for the longest list just seen plus a quarter, in steps of 256 records; shrinks slowly
  const uint32_t want = ((longest + longest / 4 + 255u) / 256u) * 256u;
  ln.head_records = want > ln.head_records ? want : std::max(want, ln.head_records - 256u * (ln.head_records > 256u));
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

static const uint32_t kCandPerFrame = 1u << 16, kOutPerFrame = 1u << 12;

// One frame (slot `frame` of the lane's uploaded / referenced sources), blocking; buffers grow and the request is re-run
// on overflow (exactness over speed).  The single-frame calls land here, and the batched paths for the rare frame
// whose survivors did not fit.
static int match_one(lm_detector* d, Lane& ln, int frame, const Query* queries, int n_q, std::vector<lm_match_rec>* out,
                     bool stage_events, std::vector<lm_raw_match>* raw_out = nullptr) {
  if (n_q < 1 || n_q > kMaxQueries) return lm_fail(LM_E_INVALID, "number of queries must be 1..%d", kMaxQueries);
  int rc = ensure_pack(d, ln);
  if (rc != LM_OK) return rc;
  Pack::Plan* plan = nullptr;
  rc = get_plan(d, queries, n_q, &plan);
  if (rc != LM_OK) return rc;
  if (frame != 0)
    for (int m = 0; m < d->model.M(); ++m) ln.src_ptr[0][m] = ln.src_ptr[frame][m];
  uint32_t cand_cap = std::max<uint32_t>(ln.cand_cap, kCandPerFrame), out_cap = std::max<uint32_t>(ln.out_cap, kOutPerFrame);
  std::vector<lm_raw_match> raw;
  uint32_t n_cands = 0;
  for (int attempt = 0;; ++attempt) {
    if (ensure_match_buffers(d, ln, cand_cap, out_cap, 1) != LM_OK) return LM_E_CUDA;
    if (enqueue_chunk(d, ln, *plan, queries, n_q, 1, ln.stream, stage_events) != LM_OK) return LM_E_CUDA;
    if (enqueue_download(ln, 1, ln.stream) != LM_OK) return LM_E_CUDA;
    CU(cudaEventSynchronize(ln.ev[5]));
    std::vector<FrameRecords> got;
    if (collect_chunk(ln, 1, ln.stream, got) != LM_OK) return LM_E_CUDA;
    raw.swap(got[0].raw);
    n_cands = got[0].n_cands;
    if (!got[0].overflow) break;
    if (attempt >= 8) return lm_fail(LM_E_CUDA, "match buffers overflowed repeatedly");
    // the candidate count of a truncated list is the chunk-wide counter: read it back
    BatchCtl ctl_head;
    CU(cudaMemcpy(&ctl_head.next_tile, &ln.ctl.as<BatchCtl>()->next_tile, 16, cudaMemcpyDeviceToHost));
    n_cands = std::max(n_cands, ctl_head.n_cands);
    if (n_cands > ln.cand_cap) cand_cap = std::max<uint32_t>(n_cands + n_cands / 4, cand_cap * 2);
    out_cap = std::max<uint32_t>(out_cap * 4, std::min<uint32_t>(n_cands + 1024, 1u << 26));
  }
  // work accounting: B_coarse from the plan; B_refine = candidates x mean refine features x 256 bytes
  ln.work_stats[1] = plan->coarse_bytes;
  ln.work_stats[4] = n_cands;
  ln.work_stats[5] = (uint64_t)plan->n_items * (uint64_t)(ln.geom.back().W * ln.geom.back().H);
  ln.work_stats[2] = plan->n_items ? (uint64_t)((double)n_cands * (plan->refine_nf_sum / plan->n_items) * 256.0) : 0;
  ln.work_stats[3] = 20ull * raw.size();
  ln.work_stats[7] = 1;
  if (raw_out) raw_out->swap(raw);
  else finalize_queries(d, ln, raw, n_q, out);
  return LM_OK;
}

static int copy_out(const std::vector<lm_match_rec>& v, lm_match_rec** out_matches, size_t* out_n) {
  lm_match_rec* p = (lm_match_rec*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(lm_match_rec));
  if (!p) return lm_fail(LM_E_INVALID, "out of host memory");
  if (!v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(lm_match_rec));
  *out_matches = p; *out_n = v.size();
  return LM_OK;
}

static int validate_template(const HostModel& md, int n_templates, const lm_template_hdr* hdr, const int32_t* feats) {
  if (n_templates != md.levels() * md.M()) return lm_fail(LM_E_INVALID, "template pyramid has %d templates, expected levels*modalities = %d", n_templates, md.levels() * md.M());
  size_t k = 0;
  for (int i = 0; i < n_templates; ++i) {
    if (hdr[i].num_features < 0 || hdr[i].num_features > LM_MAX_FEATURES) return lm_fail(LM_E_INVALID, "features.size() <= 63 violated (%d)", hdr[i].num_features);
    for (int j = 0; j < hdr[i].num_features; ++j, ++k) {
      int x = feats[3 * k], y = feats[3 * k + 1], label = feats[3 * k + 2];
      if (label < 0 || label > 7) return lm_fail(LM_E_INVALID, "feature label %d outside 0..7", label);
      if (x < -4096 || x > 4095 || y < -4096 || y > 4095) return lm_fail(LM_E_INVALID, "feature coordinate outside +-4095");
    }
  }
  return LM_OK;
}

// Binds the handle to the current CUDA device on first use by a compute entry point.  Host-only calls (persistence,
// template bookkeeping, lm_finalize_raw) never get here; everything that touches pixels does, and fails loudly when
// no device is usable -- there is no CPU path.
int set_device(lm_detector* d) {
  if (!d->cuda_ready) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      return lm_fail(LM_E_CUDA, "no CUDA device available (%s); this library has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
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

void refresh_class_cache(lm_detector* d) {
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
  if (!out) return lm_fail(LM_E_INVALID, "out is NULL");
  *out = nullptr;
  if (levels < 1 || levels > LM_MAX_LEVELS || !T) return lm_fail(LM_E_INVALID, "pyramid levels must be 1..%d", LM_MAX_LEVELS);
  if (M < 1 || M > LM_MAX_MODALITIES || !mods) return lm_fail(LM_E_INVALID, "modalities must be 1..%d", LM_MAX_MODALITIES);
  for (int m = 0; m < M; ++m) {
    if (mods[m].type != LM_COLOR_GRADIENT && mods[m].type != LM_DEPTH_NORMAL) return lm_fail(LM_E_INVALID, "unknown modality type %d", mods[m].type);
    if (mods[m].num_features < 1 || mods[m].num_features > LM_MAX_FEATURES) return lm_fail(LM_E_INVALID, "num_features must be 1..63");
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
  if (!out || !path) return lm_fail(LM_E_INVALID, "NULL argument");
  *out = nullptr;
  lm_detector* d = new lm_detector();
  std::string err;
  bool ok = false;
  try { ok = load_detector_yaml(path, d->model, err); } catch (const std::exception& e) { err = std::string(path) + ": " + e.what(); }
  if (!ok) { delete d; return lm_fail(LM_E_IO, "%s", err.c_str()); }  // every pyramid was validated by the loader
  int rc = create_common(d);
  if (rc != LM_OK) { delete d; return rc; }
  refresh_class_cache(d);
  *out = d;
  return LM_OK;
}

int lm_create_from_cache(const char* path, lm_detector** out) {
  if (!out || !path) return lm_fail(LM_E_INVALID, "NULL argument");
  *out = nullptr;
  lm_detector* d = new lm_detector();
  std::string err;
  bool ok = false;
  try { ok = load_model_cache(path, d->model, err); } catch (const std::exception& e) { err = std::string(path) + ": " + e.what(); }
  if (!ok) { delete d; return lm_fail(LM_E_IO, "%s", err.c_str()); }
  int rc = create_common(d);
  if (rc != LM_OK) { delete d; return rc; }
  refresh_class_cache(d);
  *out = d;
  return LM_OK;
}

int lm_write_cache(const lm_detector* d, const char* path) {
  if (!d || !path) return lm_fail(LM_E_INVALID, "NULL argument");
  std::string err;
  if (!save_model_cache(d->model, path, err)) return lm_fail(LM_E_IO, "%s", err.c_str());
  return LM_OK;
}

int lm_write_yaml(const lm_detector* d, const char* path) {
  if (!d || !path) return lm_fail(LM_E_INVALID, "NULL argument");
  std::string err;
  if (!save_detector_yaml(d->model, path, err)) return lm_fail(LM_E_IO, "%s", err.c_str());
  return LM_OK;
}

static std::string format_name(const char* format, const std::string& id) {
  char buf[4096];
  snprintf(buf, sizeof(buf), format, id.c_str());
  return buf;
}

int lm_read_classes(lm_detector* d, const char* const* class_ids, int n_ids, const char* format) {
  if (!d || (n_ids && !class_ids)) return lm_fail(LM_E_INVALID, "NULL argument");
  const char* fmt = format ? format : "templates_%s.yml.gz";
  for (int i = 0; i < n_ids; ++i) {
    std::string err;
    bool ok = false;
    try { ok = load_class_file(format_name(fmt, class_ids[i]), d->model, err); } catch (const std::exception& e) { err = e.what(); }
    if (!ok) return lm_fail(LM_E_IO, "%s", err.c_str());
  }
  refresh_class_cache(d);
  return LM_OK;
}

int lm_write_classes(const lm_detector* d, const char* format) {
  if (!d) return lm_fail(LM_E_INVALID, "NULL argument");
  const char* fmt = format ? format : "templates_%s.yml.gz";
  for (auto& kv : d->model.classes) {
    std::string err;
    if (!save_class_file(d->model, kv.first, format_name(fmt, kv.first), err)) return lm_fail(LM_E_IO, "%s", err.c_str());
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
    d->d_resp_all.release(); d->d_normal_lut.release();
  }
  delete d;
}

int lm_device(const lm_detector* d) { return d ? d->device : -1; }
int lm_pyramid_levels(const lm_detector* d) { return d->model.levels(); }
int lm_get_T(const lm_detector* d, int level) {
  if (level < 0 || level >= d->model.levels()) return lm_fail(LM_E_INVALID, "level out of range");
  return d->model.T[level];
}
int lm_num_modalities(const lm_detector* d) { return d->model.M(); }
int lm_get_modality(const lm_detector* d, int m, lm_modality_desc* out) {
  if (m < 0 || m >= d->model.M() || !out) return lm_fail(LM_E_INVALID, "modality out of range");
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
  if (!class_id) return lm_fail(LM_E_INVALID, "class_id is NULL");
  auto it = d->model.classes.find(class_id);
  if (it == d->model.classes.end()) return lm_fail(LM_E_NOTFOUND, "unknown class '%s'", class_id);
  if (template_id < 0 || template_id >= (int)it->second.size()) return lm_fail(LM_E_NOTFOUND, "class '%s' has no template %d", class_id, template_id);
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
  if (!d || !class_id || !hdr || !feats) return lm_fail(LM_E_INVALID, "NULL argument");
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

// Quantised maps (+ ColorGradient magnitudes) of every pyramid level of ONE frame: upload, quantisation kernels, download
// into the lane's pinned staging buffer.  Image l*M+m at host + qoff / moff.
struct QuantHost {
  uint8_t* host = nullptr;
  std::vector<size_t> qoff, moff;
};
static int quantize_to_host(lm_detector* d, const lm_image* sources, int n_sources, QuantHost& qh) {
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  const int L = d->model.levels(), M = d->model.M();
  int rc = check_sources(d, sources, n_sources, nullptr, 0);
  if (rc != LM_OK) return rc;
  const int rows = sources[0].rows, cols = sources[0].cols;
  Lane& ln = d->lane[0];
  if (upload_luts(d) != LM_OK) return LM_E_CUDA;
  if (ensure_quant_ws(d, ln, rows, cols) != LM_OK) return LM_E_CUDA;
  ln.lm_ready = false; ln.front_valid = false;
  if (upload_frames(d, ln, sources, 1, nullptr, 0) != LM_OK) return LM_E_CUDA;
  ln.launches = 0;
  if (begin_chunk(d, ln, 1, 0, ln.stream) != LM_OK) return LM_E_CUDA;
  if (run_quantize(d, ln, 1, ln.stream) != LM_OK) return LM_E_CUDA;
  size_t total = 0;
  qh.qoff.assign((size_t)L * M, 0); qh.moff.assign((size_t)L * M, 0);
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      size_t n = (size_t)(rows >> l) * (cols >> l);
      qh.qoff[l * M + m] = total; total += (n + 255) & ~(size_t)255;
      if (d->model.mods[m].type == LM_COLOR_GRADIENT) { qh.moff[l * M + m] = total; total += (n * 4 + 255) & ~(size_t)255; }
    }
  if (ln.stage_out.ensure(total) != LM_OK) return LM_E_CUDA;
  qh.host = ln.stage_out.as<uint8_t>();
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      size_t n = (size_t)(rows >> l) * (cols >> l);
      if (cudaMemcpyAsync(qh.host + qh.qoff[l * M + m], ln.quant_raw[l][m].buf.p, n, cudaMemcpyDeviceToHost, ln.stream) != cudaSuccess) return lm_fail(LM_E_CUDA, "D2H failed");
      if (d->model.mods[m].type == LM_COLOR_GRADIENT &&
          cudaMemcpyAsync(qh.host + qh.moff[l * M + m], ln.mag[l][m].buf.p, n * 4, cudaMemcpyDeviceToHost, ln.stream) != cudaSuccess)
        return lm_fail(LM_E_CUDA, "D2H failed");
    }
  if (cudaStreamSynchronize(ln.stream) != cudaSuccess) return lm_fail(LM_E_CUDA, "quantisation kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
  return LM_OK;
}

int lm_add_template(lm_detector* d, const lm_image* sources, int n_sources, const char* class_id,
                    const lm_image* object_mask, lm_rect* bounding_box) {
  if (!d || !sources || !class_id) return lm_fail(LM_E_INVALID, "NULL argument") - 100;
  const int L = d->model.levels(), M = d->model.M();
  const bool has_mask = object_mask && object_mask->data;
  if (has_mask && n_sources > 0 && (object_mask->type != LM_8UC1 || object_mask->rows != sources[0].rows || object_mask->cols != sources[0].cols))
    return lm_fail(LM_E_INVALID, "object_mask size/type mismatch") - 100;
  QuantHost qh;
  int rc = quantize_to_host(d, sources, n_sources, qh);
  if (rc != LM_OK) return rc - 100;
  const int rows = sources[0].rows, cols = sources[0].cols;
  std::vector<lm_image> qimgs((size_t)L * M);
  std::vector<const float*> mags((size_t)L * M, nullptr);
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      lm_image& q = qimgs[l * M + m];
      q.data = qh.host + qh.qoff[l * M + m]; q.rows = rows >> l; q.cols = cols >> l; q.type = LM_8UC1; q.step = (size_t)(cols >> l);
      if (d->model.mods[m].type == LM_COLOR_GRADIENT) mags[l * M + m] = reinterpret_cast<const float*>(qh.host + qh.moff[l * M + m]);
    }
  int tid = lm_add_template_from_quantized(d, qimgs.data(), mags.data(), class_id, object_mask, bounding_box);
  return tid < -1 ? tid - 100 : tid;
}

// ---- cv::linemod::Modality::process / QuantizedPyramid ([OCV] linemod.cpp: ColorGradientPyramid, DepthNormalPyramid).
// Host-side state of one processed image: per level the unmasked quantisation (+ magnitude) the CUDA front end produced
// and the decimated mask; quantize() and extractTemplate() of the reference read exactly these.
struct lm_qpyramid {
  lm_modality_desc mod;
  int rows = 0, cols = 0, levels = 0;
  std::vector<std::vector<uint8_t> > quant, mask;  // [level]; mask[level] empty when process() got no mask
  std::vector<std::vector<float> > mag;            // [level], ColorGradient only
};

int lm_modality_process(const lm_modality_desc* mod, const lm_image* src, const lm_image* mask, int levels,
                        const uint8_t* normal_lut, lm_qpyramid** out) {
  if (!mod || !src || !out) return lm_fail(LM_E_INVALID, "NULL argument");
  *out = nullptr;
  if (levels < 1 || levels > LM_MAX_LEVELS) return lm_fail(LM_E_INVALID, "levels must be 1..%d", LM_MAX_LEVELS);
  const bool has_mask = mask && mask->data;
  if (has_mask && (mask->type != LM_8UC1 || mask->rows != src->rows || mask->cols != src->cols))
    return lm_fail(LM_E_INVALID, "mask size/type mismatch (mask.size() == src.size())");
  if ((src->rows >> (levels - 1)) < 1 || (src->cols >> (levels - 1)) < 1) return lm_fail(LM_E_INVALID, "image too small for %d levels", levels);
  // a private single-modality detector carries the parameters, the LUT and the quantisation workspace of this call
  int32_t T[LM_MAX_LEVELS] = {1, 1, 1, 1};
  lm_detector* d = nullptr;
  int rc = lm_create(T, levels, mod, 1, &d);
  if (rc != LM_OK) return rc;
  if (normal_lut) lm_set_normal_lut(d, normal_lut);
  QuantHost qh;
  rc = quantize_to_host(d, src, 1, qh);
  if (rc != LM_OK) { lm_destroy(d); return rc; }
  lm_qpyramid* q = new lm_qpyramid();
  q->mod = *mod; q->rows = src->rows; q->cols = src->cols; q->levels = levels;
  q->quant.resize((size_t)levels); q->mask.resize((size_t)levels); q->mag.resize((size_t)levels);
  for (int l = 0; l < levels; ++l) {
    const size_t n = (size_t)(q->rows >> l) * (q->cols >> l);
    q->quant[(size_t)l].assign(qh.host + qh.qoff[(size_t)l], qh.host + qh.qoff[(size_t)l] + n);
    if (mod->type == LM_COLOR_GRADIENT) {
      const float* mg = reinterpret_cast<const float*>(qh.host + qh.moff[(size_t)l]);
      q->mag[(size_t)l].assign(mg, mg + n);
    }
  }
  lm_destroy(d);
  if (has_mask) {  // [OCV] pyrDown(): the mask is NN-resized, i.e. plain index decimation
    q->mask[0].resize((size_t)q->rows * q->cols);
    for (int y = 0; y < q->rows; ++y) std::memcpy(&q->mask[0][(size_t)y * q->cols], (const uint8_t*)mask->data + (size_t)y * mask->step, (size_t)q->cols);
    for (int l = 1; l < levels; ++l) {
      const int pc = q->cols >> (l - 1), r = q->rows >> l, c = q->cols >> l;
      q->mask[(size_t)l].resize((size_t)r * c);
      for (int y = 0; y < r; ++y)
        for (int x = 0; x < c; ++x) q->mask[(size_t)l][(size_t)y * c + x] = q->mask[(size_t)l - 1][(size_t)(2 * y) * pc + 2 * x];
    }
  }
  *out = q;
  return LM_OK;
}

void lm_qpyramid_destroy(lm_qpyramid* q) { delete q; }
int lm_qpyramid_levels(const lm_qpyramid* q) { return q ? q->levels : 0; }

int lm_qpyramid_size(const lm_qpyramid* q, int level, int* rows, int* cols) {
  if (!q || level < 0 || level >= q->levels) return lm_fail(LM_E_INVALID, "level out of range");
  if (rows) *rows = q->rows >> level;
  if (cols) *cols = q->cols >> level;
  return LM_OK;
}

// [OCV] {ColorGradient,DepthNormal}Pyramid::quantize: dst = zeros; quantised.copyTo(dst, mask)
int lm_qpyramid_quantize(const lm_qpyramid* q, int level, lm_image* dst) {
  if (!q || !dst || !dst->data) return lm_fail(LM_E_INVALID, "NULL argument");
  if (level < 0 || level >= q->levels) return lm_fail(LM_E_INVALID, "level out of range");
  const int r = q->rows >> level, c = q->cols >> level;
  if (dst->type != LM_8UC1 || dst->rows != r || dst->cols != c || dst->step < (size_t)c) return lm_fail(LM_E_INVALID, "dst must be a %dx%d CV_8UC1 image", c, r);
  const uint8_t* src = q->quant[(size_t)level].data();
  const uint8_t* mk = q->mask[(size_t)level].empty() ? nullptr : q->mask[(size_t)level].data();
  for (int y = 0; y < r; ++y) {
    uint8_t* o = (uint8_t*)dst->data + (size_t)y * dst->step;
    for (int x = 0; x < c; ++x) o[x] = (!mk || mk[(size_t)y * c + x]) ? src[(size_t)y * c + x] : 0;
  }
  return LM_OK;
}

// [OCV] QuantizedPyramid::extractTemplate after `level` pyrDown() calls (num_features and extract_threshold halved per
// level).  Returns 1 (hdr = {-1, -1, level, n}, features = n (x, y, label) triples in level coordinates), 0 when the
// level lacks candidates (the reference returns false), < 0 = LM_E_*.
int lm_qpyramid_extract(const lm_qpyramid* q, int level, lm_template_hdr* hdr, int32_t* features) {
  if (!q || !hdr) return lm_fail(LM_E_INVALID, "NULL argument");
  if (level < 0 || level >= q->levels) return lm_fail(LM_E_INVALID, "level out of range");
  int nf = q->mod.num_features, ext = q->mod.extract_threshold;
  for (int l = 0; l < level; ++l) { nf /= 2; ext /= 2; }
  const int r = q->rows >> level, c = q->cols >> level;
  const uint8_t* mk = q->mask[(size_t)level].empty() ? nullptr : q->mask[(size_t)level].data();
  Template t;
  const bool ok = q->mod.type == LM_COLOR_GRADIENT
                      ? extract_color_gradient(q->quant[(size_t)level].data(), q->mag[(size_t)level].data(), mk, r, c, q->mod.strong_threshold, nf, level, t)
                      : extract_depth_normal(q->quant[(size_t)level].data(), mk, r, c, nf, ext, level, t);
  hdr->width = -1; hdr->height = -1; hdr->pyramid_level = level; hdr->num_features = ok ? (int)t.features.size() : 0;
  if (!ok) return 0;
  if (features)
    for (size_t j = 0; j < t.features.size(); ++j) {
      features[3 * j] = t.features[j].x; features[3 * j + 1] = t.features[j].y; features[3 * j + 2] = t.features[j].label;
    }
  return 1;
}

int lm_add_template_from_quantized(lm_detector* d, const lm_image* quantized, const float* const* magnitudes,
                                   const char* class_id, const lm_image* object_mask, lm_rect* bounding_box) {
  if (!d || !quantized || !magnitudes || !class_id) return lm_fail(LM_E_INVALID, "NULL argument");
  const int L = d->model.levels(), M = d->model.M();
  const int rows = quantized[0].rows, cols = quantized[0].cols;
  const bool has_mask = object_mask && object_mask->data;
  if (has_mask && (object_mask->type != LM_8UC1 || object_mask->rows != rows || object_mask->cols != cols))
    return lm_fail(LM_E_INVALID, "object_mask size/type mismatch");
  std::vector<std::vector<uint8_t> > qbuf((size_t)L * M);
  for (int l = 0; l < L; ++l)
    for (int m = 0; m < M; ++m) {
      const lm_image& q = quantized[l * M + m];
      if (!q.data || q.type != LM_8UC1 || q.rows != (rows >> l) || q.cols != (cols >> l)) return lm_fail(LM_E_INVALID, "quantized[%d] must be a %dx%d CV_8UC1 image", l * M + m, cols >> l, rows >> l);
      if (d->model.mods[m].type == LM_COLOR_GRADIENT && !magnitudes[l * M + m]) return lm_fail(LM_E_INVALID, "magnitudes[%d] missing", l * M + m);
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
  if (!d || world < 1 || rank < 0 || rank >= world) return lm_fail(LM_E_INVALID, "bad shard %d/%d", rank, world);
  d->shard_rank = rank; d->shard_world = world;
  return LM_OK;
}

int lm_set_similarity_lut(lm_detector* d, const uint8_t lut[256]) {
  for (int i = 0; i < 256; ++i)
    if (lut[i] > 4) return lm_fail(LM_E_INVALID, "similarity LUT entries must be <= 4 (u8 accumulation of 63 features)");
  std::memcpy(d->sim_lut, lut, 256);
  d->luts_dirty = true;
  return LM_OK;
}
int lm_get_similarity_lut(const lm_detector* d, uint8_t lut[256]) { std::memcpy(lut, d->sim_lut, 256); return LM_OK; }
int lm_set_normal_lut(lm_detector* d, const uint8_t lut[8000]) { std::memcpy(d->normal_lut, lut, 8000); d->luts_dirty = true; return LM_OK; }
int lm_get_normal_lut(const lm_detector* d, uint8_t lut[8000]) { std::memcpy(lut, d->normal_lut, 8000); return LM_OK; }

// OpenCV ships NORMAL_LUT as text (modules/objdetect/src/normal_lut.i: a brace-initialised unsigned char [20][20][20]); a
// user who holds OpenCV injects the real table with one call.  Every integer after the first '{' is an entry.
int lm_load_normal_lut_file(lm_detector* d, const char* path) {
  if (!d || !path) return lm_fail(LM_E_INVALID, "NULL argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return lm_fail(LM_E_IO, "cannot open '%s'", path);
  std::string text;
  char buf[65536];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
  std::fclose(f);
  size_t i = text.find('{');
  if (i == std::string::npos) return lm_fail(LM_E_IO, "%s: no '{' -- not a brace-initialised table", path);
  std::vector<uint8_t> v;
  while (i < text.size()) {
    const char c = text[i];
    if (c == '/' && i + 1 < text.size() && text[i + 1] == '/') { while (i < text.size() && text[i] != '\n') ++i; continue; }
    if (c == '/' && i + 1 < text.size() && text[i + 1] == '*') { const size_t e = text.find("*/", i + 2); i = e == std::string::npos ? text.size() : e + 2; continue; }
    if (c >= '0' && c <= '9') {
      unsigned long val = 0;
      if (c == '0' && i + 1 < text.size() && (text[i + 1] == 'x' || text[i + 1] == 'X')) {
        i += 2;
        while (i < text.size() && std::isxdigit((unsigned char)text[i])) { val = val * 16 + (unsigned long)(std::isdigit((unsigned char)text[i]) ? text[i] - '0' : (std::tolower(text[i]) - 'a' + 10)); ++i; }
      } else {
        while (i < text.size() && text[i] >= '0' && text[i] <= '9') { val = val * 10 + (unsigned long)(text[i] - '0'); ++i; }
      }
      if (val > 255) return lm_fail(LM_E_IO, "%s: entry %zu is %lu (> 255)", path, v.size(), val);
      v.push_back((uint8_t)val);
      continue;
    }
    ++i;
  }
  if (v.size() != 8000) return lm_fail(LM_E_IO, "%s: %zu entries, NORMAL_LUT has 20 * 20 * 20 = 8000", path, v.size());
  return lm_set_normal_lut(d, v.data());
}

int lm_set_option(lm_detector* d, const char* key, int value) {
  if (!d || !key) return lm_fail(LM_E_INVALID, "NULL argument");
  std::string k(key);
  if (d->stream_open && k != "timing" && k != "finalize_threads")   // an open lm_stream holds plans, graphs and workspace geometry
    return lm_fail(LM_E_STATE, "option '%s' cannot change while the handle has an open lm_stream", key);
  if (k == "debug_taps") d->debug_taps = value;
  else if (k == "timing") d->timing = value;
  else if (k == "prune") d->prune = value;
  else if (k == "mod_order") d->mod_order = value & 3;
  else if (k == "refine_tiled") d->refine_tiled = value != 0;   // takes effect with the next request (workspace rebuilt)
  else if (k == "dn_count") {           // A/B switch: 0 keeps the 99-exchange median network for one-hot tables too
    d->dn_count = value != 0;
    d->luts_dirty = true;
  }
  else if (k == "coarse_share") {       // A/B switch: tail passes of <= 128 positions scored for eight frames per warp
    d->coarse_share = value != 0;
    d->pack.plans.clear();              // plans order their tiles by it (the lanes' graphs are keyed by the plan)
    for (int i = 0; i < LM_LANES; ++i) d->lane[i].drop_graphs();
  }
  else if (k == "graphs") d->graphs = value;
  else if (k == "batch_frames") {
    if (value < 1 || value > LM_MAX_BATCH) return lm_fail(LM_E_INVALID, "batch_frames must be 1..%d", LM_MAX_BATCH);
    d->batch_frames = value;
  }
  else if (k == "stream_frames") {      // frames per chunk of an lm_stream opened afterwards
    if (value < 1 || value > LM_MAX_BATCH) return lm_fail(LM_E_INVALID, "stream_frames must be 1..%d", LM_MAX_BATCH);
    d->stream_frames = value;
  }
  else if (k == "batch_lanes") {
    if (value < 1 || value > LM_LANES) return lm_fail(LM_E_INVALID, "batch_lanes must be 1..%d", LM_LANES);
    d->batch_lanes = value;
  }
  else if (k == "coarse_grid_limit") {  // process-wide; recorded graphs hold the old grid
    set_coarse_grid_limit(value);
    for (int i = 0; i < LM_LANES; ++i) d->lane[i].drop_graphs();
  }
  else if (k == "coarse_narrow") {      // process-wide A/B switch; recorded graphs hold the old kernel
    set_coarse_narrow(value < 0 || value > 2 ? 1 : value);
    for (int i = 0; i < LM_LANES; ++i) d->lane[i].drop_graphs();
  }
  else if (k == "finalize_threads") d->finalize_threads = std::max(0, std::min(value, 16));
  else if (k == "device_out_cap") d->device_out_cap = (uint32_t)std::max(16, value);
  else if (k == "cand_per_frame") d->cand_per_frame = (uint32_t)std::max(1024, value);
  else return lm_fail(LM_E_INVALID, "unknown option '%s'", key);
  return LM_OK;
}

// ---------------------------------------------------------------------------------------------- match
// Validates and uploads n_frames host frames (sources[f * M + m]) into the lane's source slots; nothing else is enqueued.
static int frames_from_host(lm_detector* d, Lane& ln, const lm_image* sources, int n_frames, int n_sources, const lm_image* masks,
                            int n_masks) {
  for (int f = 0; f < n_frames; ++f) {
    int rc = check_sources(d, sources + (size_t)f * n_sources, n_sources, f == 0 ? masks : nullptr, f == 0 ? n_masks : 0);
    if (rc != LM_OK) return rc;
    if (sources[(size_t)f * n_sources].rows != sources[0].rows || sources[(size_t)f * n_sources].cols != sources[0].cols)
      return lm_fail(LM_E_INVALID, "frames of a batch differ in size");
  }
  if (set_device(d) != LM_OK || upload_luts(d) != LM_OK) return LM_E_CUDA;
  const int rows = sources[0].rows, cols = sources[0].cols;
  int rc = ensure_lm_ws(d, ln, rows, cols, n_frames);
  if (rc != LM_OK) return rc;
  std::memset(ln.work_stats, 0, sizeof(ln.work_stats));
  if (d->timing) CU(cudaEventRecord(ln.ev[0], ln.stream));
  if (upload_frames(d, ln, sources, n_frames, masks, n_masks) != LM_OK) return LM_E_CUDA;
  // B_front: sources read once + linear memories written once
  uint64_t bf = 0;
  for (int m = 0; m < n_sources; ++m) bf += src_row_bytes(sources[m].type, cols) * rows;
  for (size_t l = 0; l < ln.geom.size(); ++l) bf += (uint64_t)n_sources * 8 * ln.geom[l].rows * ln.geom[l].cols;
  ln.work_stats[0] = bf;
  return LM_OK;
}

int lm_build_front(lm_detector* d, const lm_image* sources, int n_sources, const lm_image* masks, int n_masks) {
  if (!d || !sources) return lm_fail(LM_E_INVALID, "NULL argument");
  if (d->stream_open) return lm_fail(LM_E_STATE, "the handle has an open lm_stream: close it before other matching calls");
  Lane& ln = d->lane[0];
  int rc = frames_from_host(d, ln, sources, 1, n_sources, masks, n_masks);
  if (rc != LM_OK) return rc;
  ln.launches = 0;
  if (begin_chunk(d, ln, 1, 0, ln.stream) != LM_OK) return LM_E_CUDA;
  if (run_front(d, ln, 1, ln.stream) != LM_OK) return LM_E_CUDA;
  CU(cudaStreamSynchronize(ln.stream));
  return LM_OK;
}

static int to_queries(const lm_query* in, int n, Query* out) {
  if (!in || n < 1 || n > kMaxQueries) return lm_fail(LM_E_INVALID, "number of queries must be 1..%d", kMaxQueries);
  for (int q = 0; q < n; ++q) {
    if (in[q].n_ids < 0 || (in[q].n_ids > 0 && !in[q].class_ids)) return lm_fail(LM_E_INVALID, "query %d: bad class id list", q);
    out[q].threshold = in[q].threshold; out[q].class_ids = in[q].class_ids; out[q].n_ids = in[q].n_ids;
  }
  return LM_OK;
}

int lm_match_multi(lm_detector* d, const lm_image* sources, int n_sources, const lm_query* queries, int n_queries,
                   const lm_image* masks, int n_masks, lm_image_out* quantized_out, lm_match_rec** out_matches,
                   size_t* out_offsets) {
  if (!d || !sources || !out_matches || !out_offsets) return lm_fail(LM_E_INVALID, "NULL argument");
  if (d->stream_open) return lm_fail(LM_E_STATE, "the handle has an open lm_stream: close it before other matching calls");
  *out_matches = nullptr;
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  Lane& ln = d->lane[0];
  rc = frames_from_host(d, ln, sources, 1, n_sources, masks, n_masks);
  if (rc != LM_OK) return rc;
  std::vector<lm_match_rec> out[kMaxQueries];
  rc = match_one(d, ln, 0, qs, n_queries, out, d->timing != 0);
  if (rc != LM_OK) return rc;
  if (d->timing) collect_timings(ln);
  if (quantized_out) {
    const int L = d->model.levels(), M = d->model.M();
    for (int l = 0; l < L; ++l)
      for (int m = 0; m < M; ++m) {
        lm_image_out& q = quantized_out[l * M + m];
        const LevelGeom& g = ln.geom[l];
        if (!q.data || q.rows != g.rows || q.cols != g.cols || q.type != LM_8UC1 || q.step < (size_t)g.cols)
          return lm_fail(LM_E_INVALID, "quantized_out[%d] must be a %dx%d CV_8UC1 image", l * M + m, g.cols, g.rows);
        CU(cudaMemcpy2D(q.data, q.step, ln.quantized[l][m].as<uint8_t>(), g.cols, g.cols, g.rows, cudaMemcpyDeviceToHost));
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
  if (!out_n) return lm_fail(LM_E_INVALID, "NULL argument");
  *out_n = 0;
  lm_query q = {threshold, class_ids, n_ids};
  size_t offs[2] = {0, 0};
  int rc = lm_match_multi(d, sources, n_sources, &q, 1, masks, n_masks, quantized_out, out_matches, offs);
  if (rc == LM_OK) *out_n = offs[1];
  return rc;
}

// Pre-sizes a lane for chunks of `frames` frames of this geometry: workspace, template pack, plan, result blocks.
static int prepare_lane(lm_detector* d, Lane& ln, int rows, int cols, int frames, const Query* qs, int n_q, uint32_t out_cap,
                        Pack::Plan** plan) {
  int rc = ensure_lm_ws(d, ln, rows, cols, frames);
  if (rc != LM_OK) return rc;
  rc = ensure_pack(d, ln);
  if (rc != LM_OK) return rc;
  rc = get_plan(d, qs, n_q, plan);
  if (rc != LM_OK) return rc;
  const uint32_t cand_cap = std::max<uint32_t>(ln.cand_cap, d->cand_per_frame * (uint32_t)std::max(frames, ln.frames));
  return ensure_match_buffers(d, ln, cand_cap, std::max<uint32_t>(ln.out_cap, out_cap), std::max(frames, ln.frames));
}

// Frames in chunks of `batch_frames`, chunks pipelined over `batch_lanes` workspace lanes: while one chunk's kernels run,
// the next chunk's frames are copied to the device and the previous chunk's survivors are ordered on the host.  Every
// kernel launch covers a whole chunk.  BatchPipe is that pipeline: lm_match_batch* runs one to completion per call,
// lm_stream keeps one alive between calls so that the device never drains at a call boundary.
struct BatchPipe {
  lm_detector* d = nullptr;
  Query qs[kMaxQueries];
  int n_q = 0, F = 8, NL = 4, ws_frames = 8;
  bool use_pool = false;
  // raw mode (nullable): instead of finalised lists, frame f's un-ordered survivor records go to (*raw_frames)[f] -- what a
  // template-sharded caller merges across shards before the reference's sort + unique
  std::vector<std::vector<lm_raw_match> >* raw_frames = nullptr;
  // per frame the n_q result lists (contiguous: the finalisers write lists[q]) of the frames [base, base + lists.size()).
  // A deque: finalizer jobs hold pointers into its elements, which stay put while the container grows at the back and
  // shrinks at the front.
  std::deque<std::vector<std::vector<lm_match_rec> > > lists;
  long long base = 0, submitted = 0, finished = 0, chunks = 0;   // frame / chunk counters since the pipe was set up
  struct Pending { long long first = -1; int n = 0; const Pack::Plan* plan = nullptr; } pending[LM_LANES];
  std::vector<FrameRecords> got;
  bool prof = false;
  double t_fin = 0, t_up = 0, t_enq = 0;
  static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

  int setup(lm_detector* det, const Query* queries, int n_queries, int total_frames, bool pool, int chunk_frames) {
    d = det; n_q = n_queries;
    for (int q = 0; q < n_q; ++q) qs[q] = queries[q];
    F = std::max(1, std::min(chunk_frames, LM_MAX_BATCH));
    NL = std::max(1, std::min(d->batch_lanes, LM_LANES));
    ws_frames = total_frames > 0 ? std::min(F, total_frames) : F;
    use_pool = pool && d->finalize_threads > 0;
    if (use_pool) d->finalizers.start(std::min(d->finalize_threads, 16));
    prof = getenv("LM_HOST_PROFILE") != nullptr;
    return LM_OK;
  }
  std::vector<lm_match_rec>* list_of(long long frame) { return lists[(size_t)(frame - base)].data(); }

  // The lane's chunk is complete on the device: its survivors become lists (ordered here or on a finalizer thread).
  int finish(int li) {
    Lane& ln = d->lane[li];
    const Pending pd = pending[li];
    pending[li].first = -1;
    CU(cudaEventSynchronize(ln.ev[5]));
    if (d->timing) collect_timings(ln);  // per-stage events of this chunk (plain launches): lm_last_timings of lane 0
    std::vector<int> redo;
    uint64_t cands = 0, survivors = 0;
    if (collect_chunk(ln, pd.n, ln.stream, got) != LM_OK) return LM_E_CUDA;
    for (int f = 0; f < pd.n; ++f) {
      cands += got[(size_t)f].n_cands; survivors += got[(size_t)f].raw.size();
      if (got[(size_t)f].overflow) { redo.push_back(f); continue; }
      if (raw_frames) (*raw_frames)[(size_t)(pd.first + f)].swap(got[(size_t)f].raw);
      else if (use_pool) {  // ordered on a finalizer thread while this thread goes on feeding the device
        std::shared_ptr<std::vector<lm_raw_match> > raw = std::make_shared<std::vector<lm_raw_match> >();
        raw->swap(got[(size_t)f].raw);
        std::vector<lm_match_rec>* dst = list_of(pd.first + f);
        const int levels = d->model.levels(), nq = n_q;
        d->finalizers.submit([raw, dst, levels, nq]() { lm_internal_finalize(levels, *raw, nq, dst); });
      } else finalize_queries(d, ln, got[(size_t)f].raw, n_q, list_of(pd.first + f));
    }
    // work accounting of the chunk (lm_last_work reads lane 0): B_coarse of all its frames, candidates, evals, frames
    ln.work_stats[1] = pd.plan->coarse_bytes * (uint64_t)pd.n;
    ln.work_stats[4] = cands;
    ln.work_stats[5] = (uint64_t)pd.plan->n_items * (uint64_t)(ln.geom.back().W * ln.geom.back().H) * (uint64_t)pd.n;
    ln.work_stats[2] = pd.plan->n_items ? (uint64_t)((double)cands * (pd.plan->refine_nf_sum / pd.plan->n_items) * 256.0) : 0;
    ln.work_stats[3] = 20ull * survivors;
    ln.work_stats[7] = (uint64_t)pd.n;
    for (int f : redo) {  // rare: this frame alone with growing buffers (its sources are still in the lane's slot f)
      int rc = match_one(d, ln, f, qs, n_q, raw_frames ? nullptr : list_of(pd.first + f), false,
                         raw_frames ? &(*raw_frames)[(size_t)(pd.first + f)] : nullptr);
      if (rc != LM_OK) return rc;
    }
    finished = pd.first + pd.n;   // chunks finish in submission order
    return LM_OK;
  }

  // One chunk of n <= F host frames (fs[f * n_sources + m]): waits for the lane's previous chunk, uploads, enqueues.
  int submit(const lm_image* fs, int n, int n_sources) {
    const int li = (int)(chunks % NL);
    Lane& ln = d->lane[li];
    double t0 = prof ? now() : 0;
    // the lane's previous chunk must be done before its source slots and pinned staging are overwritten
    if (pending[li].first >= 0) { int rc = finish(li); if (rc != LM_OK) return rc; }
    double t1 = prof ? now() : 0;
    int rc = check_sources(d, fs, n_sources, nullptr, 0);
    if (rc != LM_OK) return rc;
    Pack::Plan* plan = nullptr;
    rc = prepare_lane(d, ln, fs[0].rows, fs[0].cols, ws_frames, qs, n_q, kOutPerFrame, &plan);
    if (rc != LM_OK) return rc;
    rc = frames_from_host(d, ln, fs, n, n_sources, nullptr, 0);
    if (rc != LM_OK) return rc;
    double t2 = prof ? now() : 0;
    if (!raw_frames)
      for (int f = 0; f < n; ++f) lists.emplace_back((size_t)n_q);
    if (enqueue_chunk(d, ln, *plan, qs, n_q, n, ln.stream, d->timing != 0) != LM_OK) return LM_E_CUDA;
    if (enqueue_download(ln, n, ln.stream) != LM_OK) return LM_E_CUDA;
    pending[li].first = submitted; pending[li].n = n; pending[li].plan = plan;
    submitted += n; ++chunks;
    if (prof) { double t3 = now(); t_fin += t1 - t0; t_up += t2 - t1; t_enq += t3 - t2; }
    return LM_OK;
  }
  // the oldest chunk in flight, or -1
  int oldest() const {
    int best = -1;
    for (int li = 0; li < LM_LANES; ++li)
      if (pending[li].first >= 0 && (best < 0 || pending[li].first < pending[best].first)) best = li;
    return best;
  }
  int drain() {   // in submission order
    for (int li = oldest(); li >= 0; li = oldest()) { int rc = finish(li); if (rc != LM_OK) return rc; }
    return LM_OK;
  }
  int finish_ready() {   // without blocking: the oldest chunks whose survivors have already reached the host
    for (int li = oldest(); li >= 0; li = oldest()) {
      const cudaError_t e = cudaEventQuery(d->lane[li].ev[5]);
      if (e == cudaErrorNotReady) break;
      if (e != cudaSuccess) return lm_fail(LM_E_CUDA, "CUDA error: %s", cudaGetErrorString(e));
      int rc = finish(li);
      if (rc != LM_OK) return rc;
    }
    return LM_OK;
  }
  void abandon() {   // error paths: nothing of this pipe may stay in flight or keep pointers into `lists`
    for (int li = 0; li < LM_LANES; ++li)
      if (pending[li].first >= 0) { lane_quiesce(d->lane[li]); pending[li].first = -1; }
    if (use_pool) d->finalizers.wait_all();
  }
};

// out_offsets: n_frames * n_q + 1 prefix offsets, frame-major.  raw_frames: see BatchPipe.
static int match_batch_impl(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, const Query* qs, int n_q,
                            lm_match_rec** out_matches, size_t* out_offsets,
                            std::vector<std::vector<lm_raw_match> >* raw_frames = nullptr) {
  if (out_matches) *out_matches = nullptr;
  if (n_q < 1 || n_q > kMaxQueries) return lm_fail(LM_E_INVALID, "number of queries must be 1..%d", kMaxQueries);
  if (out_offsets) out_offsets[0] = 0;
  if (raw_frames) raw_frames->assign((size_t)n_frames, std::vector<lm_raw_match>());
  if (n_frames == 0) { size_t n = 0; return raw_frames ? LM_OK : copy_out(std::vector<lm_match_rec>(), out_matches, &n); }
  if (d->stream_open) return lm_fail(LM_E_STATE, "the handle has an open lm_stream: close it before other matching calls");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  BatchPipe pipe;
  pipe.raw_frames = raw_frames;
  pipe.setup(d, qs, n_q, n_frames, !raw_frames && n_frames > 1, d->batch_frames);
  struct Guard {  // nothing may outlive this call, whichever way it returns
    BatchPipe* p;
    ~Guard() { p->abandon(); }
  } guard = {&pipe};
  // Chunk boundaries.  The first chunks of a call ramp up (2, 2, 4, ... frames): the GPU starts on the call's first frames
  // after two frames' worth of host->device copy instead of a whole chunk's, which is the bubble between two calls.
  const int F = pipe.F;
  for (int at = 0, step = std::min(F, 2), k = 0; at < n_frames; ++k) {
    const int n = std::min(step, n_frames - at);
    int rc = pipe.submit(sources + (size_t)at * n_sources, n, n_sources);
    if (rc != LM_OK) return rc;
    at += n;
    if (k >= 1 && step < F) step = std::min(F, step * 2);
  }
  int rc = pipe.drain();
  if (rc != LM_OK) return rc;
  if (pipe.prof)
    fprintf(stderr, "[lm host profile] per frame us: finish %.1f upload %.1f enqueue %.1f\n", pipe.t_fin / n_frames, pipe.t_up / n_frames,
            pipe.t_enq / n_frames);
  if (raw_frames) return LM_OK;
  if (pipe.use_pool) d->finalizers.wait_all();
  std::vector<lm_match_rec> all;
  for (size_t f = 0; f < pipe.lists.size(); ++f)
    for (int q = 0; q < n_q; ++q) {
      all.insert(all.end(), pipe.lists[f][(size_t)q].begin(), pipe.lists[f][(size_t)q].end());
      out_offsets[f * (size_t)n_q + (size_t)q + 1] = all.size();
    }
  size_t n = 0;
  return copy_out(all, out_matches, &n);
}

// ---------------------------------------------------------------------------------------------- streams of host frames
struct lm_stream {
  lm_detector* d = nullptr;
  BatchPipe pipe;
  std::vector<std::vector<std::string> > ids;       // the queries' class ids, owned
  std::vector<std::vector<const char*> > id_ptrs;
  int n_sources = 0;
  int rows = 0, cols = 0;   // geometry of the frames in flight
  bool failed = false;
};

int lm_stream_open(lm_detector* d, const lm_query* queries, int n_queries, lm_stream** out) {
  if (!d || !out) return lm_fail(LM_E_INVALID, "NULL argument");
  *out = nullptr;
  if (d->stream_open) return lm_fail(LM_E_STATE, "the handle already has an open lm_stream");
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  std::unique_ptr<lm_stream> s(new lm_stream());
  s->d = d;
  s->ids.resize((size_t)n_queries); s->id_ptrs.resize((size_t)n_queries);
  for (int q = 0; q < n_queries; ++q) {
    for (int i = 0; i < qs[q].n_ids; ++i) {
      if (!qs[q].class_ids[i]) return lm_fail(LM_E_INVALID, "class_ids[%d] is NULL", i);
      s->ids[(size_t)q].push_back(qs[q].class_ids[i]);
    }
    for (const std::string& id : s->ids[(size_t)q]) s->id_ptrs[(size_t)q].push_back(id.c_str());
    qs[q].class_ids = s->id_ptrs[(size_t)q].data();
  }
  s->pipe.setup(d, qs, n_queries, 0, true, d->stream_frames);
  d->stream_open = true;
  *out = s.release();
  return LM_OK;
}

int lm_stream_push(lm_stream* s, const lm_image* sources, int n_frames, int n_sources) {
  if (!s || (!sources && n_frames > 0) || n_frames < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  if (s->failed) return lm_fail(LM_E_STATE, "the stream failed earlier: close it");
  if (n_sources != s->d->model.M())
    return lm_fail(LM_E_INVALID, "sources.size() (%d) != modalities.size() (%d)", n_sources, s->d->model.M());
  if (set_device(s->d) != LM_OK) return LM_E_CUDA;
  if (n_frames > 0 && (sources[0].rows != s->rows || sources[0].cols != s->cols)) {
    // another frame size re-packs the templates and rebuilds the lanes' workspaces: nothing of the old size may be in flight
    int rc = s->pipe.drain();
    if (rc != LM_OK) { s->failed = true; s->pipe.abandon(); return rc; }
    s->rows = sources[0].rows; s->cols = sources[0].cols;
  }
  for (int f = 1; f < n_frames; ++f)
    if (sources[(size_t)f * n_sources].rows != sources[0].rows || sources[(size_t)f * n_sources].cols != sources[0].cols)
      return lm_fail(LM_E_INVALID, "frames of a push differ in size");
  for (int at = 0; at < n_frames;) {
    // a stream's very first chunks ramp up like a batch call's; after that every chunk is full
    int n = std::min(s->pipe.F, n_frames - at);
    if (s->pipe.chunks < 2) n = std::min(n, 2);
    else if (s->pipe.chunks == 2) n = std::min(n, 4);
    int rc = s->pipe.submit(sources + (size_t)at * n_sources, n, n_sources);
    if (rc != LM_OK) { s->failed = true; s->pipe.abandon(); return rc; }
    at += n;
  }
  return LM_OK;
}

int lm_stream_pop(lm_stream* s, int wait_all, int max_frames, lm_match_rec** out_matches, size_t* out_offsets, int* n_frames_out) {
  if (!s || !out_matches || !out_offsets || !n_frames_out) return lm_fail(LM_E_INVALID, "NULL argument");
  *out_matches = nullptr; *n_frames_out = 0; out_offsets[0] = 0;
  if (s->failed) return lm_fail(LM_E_STATE, "the stream failed earlier: close it");
  if (set_device(s->d) != LM_OK) return LM_E_CUDA;
  BatchPipe& p = s->pipe;
  int rc = wait_all ? p.drain() : p.finish_ready();
  if (rc != LM_OK) { s->failed = true; p.abandon(); return rc; }
  if (p.use_pool) s->d->finalizers.wait_all();
  const long long ready = std::min<long long>(p.finished - p.base, std::max(0, max_frames));
  std::vector<lm_match_rec> all;
  for (long long f = 0; f < ready; ++f)
    for (int q = 0; q < p.n_q; ++q) {
      const std::vector<lm_match_rec>& l = p.lists[(size_t)f][(size_t)q];
      all.insert(all.end(), l.begin(), l.end());
      out_offsets[f * p.n_q + q + 1] = all.size();
    }
  p.lists.erase(p.lists.begin(), p.lists.begin() + (size_t)ready);
  p.base += ready;
  *n_frames_out = (int)ready;
  size_t n = 0;
  return copy_out(all, out_matches, &n);
}

int lm_stream_in_flight(const lm_stream* s) { return s ? (int)(s->pipe.submitted - s->pipe.base) : 0; }

void lm_stream_close(lm_stream* s) {
  if (!s) return;
  if (set_device(s->d) == LM_OK) s->pipe.abandon();
  s->d->stream_open = false;
  delete s;
}

// Internal entry points of lm_group.cu (a handle per device, driven from the group's worker threads).
int lm_internal_match_batch_raw(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                                int n_queries, std::vector<std::vector<lm_raw_match> >* raw_frames) {
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  return match_batch_impl(d, sources, n_frames, n_sources, qs, n_queries, nullptr, nullptr, raw_frames);
}
void lm_internal_finalize(int levels, std::vector<lm_raw_match>& raw, int n_queries, std::vector<lm_match_rec>* out) {
  std::vector<lm_raw_match> part;
  std::vector<lm_match_rec> presort;
  for (int q = 0; q < n_queries; ++q) {
    part.clear();
    for (const lm_raw_match& r : raw)
      if ((int)(r.order_key >> 28) == q) part.push_back(r);
    finalize_records(levels, part, presort, out[q]);
  }
}
lm_detector* lm_internal_clone(const lm_detector* src) {
  lm_detector* d = new lm_detector();
  d->model = src->model;
  std::memcpy(d->sim_lut, src->sim_lut, sizeof(d->sim_lut));
  std::memcpy(d->normal_lut, src->normal_lut, sizeof(d->normal_lut));
  d->luts_dirty = true;
  d->device_out_cap = src->device_out_cap; d->cand_per_frame = src->cand_per_frame;
  d->prune = src->prune; d->graphs = src->graphs; d->mod_order = src->mod_order;
  d->batch_frames = src->batch_frames; d->batch_lanes = src->batch_lanes; d->finalize_threads = src->finalize_threads;
  d->refine_tiled = src->refine_tiled; d->coarse_share = src->coarse_share; d->stream_frames = src->stream_frames; d->dn_count = src->dn_count;
  refresh_class_cache(d);
  return d;
}

int lm_match_batch(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, float threshold,
                   const char* const* class_ids, int n_ids, lm_match_rec** out_matches, size_t* out_offsets) {
  if (!d || !sources || !out_matches || !out_offsets || n_frames < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  const Query query = {threshold, class_ids, n_ids};
  return match_batch_impl(d, sources, n_frames, n_sources, &query, 1, out_matches, out_offsets);
}

int lm_match_batch_multi(lm_detector* d, const lm_image* sources, int n_frames, int n_sources, const lm_query* queries,
                         int n_queries, lm_match_rec** out_matches, size_t* out_offsets) {
  if (!d || !sources || !out_matches || !out_offsets || n_frames < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  return match_batch_impl(d, sources, n_frames, n_sources, qs, n_queries, out_matches, out_offsets);
}

void lm_free_matches(lm_match_rec* m) { std::free(m); }

// A chunk of device-resident frames on one lane and the caller's stream: nothing is uploaded, copied or synchronised.
static int device_chunk(lm_detector* d, int lane_index, const void* const* d_sources, int n_frames, int n_sources, int rows,
                        int cols, const Query* qs, int n_q, int ws_frames, cudaStream_t s) {
  if (d->stream_open) return lm_fail(LM_E_STATE, "the handle has an open lm_stream: close it before other matching calls");
  if (lane_index < 0 || lane_index >= LM_LANES) return lm_fail(LM_E_INVALID, "lane must be 0..%d", LM_LANES - 1);
  if (n_sources != d->model.M()) return lm_fail(LM_E_INVALID, "sources.size() (%d) != modalities.size() (%d)", n_sources, d->model.M());
  if (n_frames < 1 || n_frames > LM_MAX_BATCH) return lm_fail(LM_E_INVALID, "a chunk holds 1..%d frames", LM_MAX_BATCH);
  if (set_device(d) != LM_OK || upload_luts(d) != LM_OK) return LM_E_CUDA;
  Lane& ln = d->lane[lane_index];
  Pack::Plan* plan = nullptr;
  int rc = prepare_lane(d, ln, rows, cols, std::max(ws_frames, n_frames), qs, n_q, d->device_out_cap, &plan);
  if (rc != LM_OK) return rc;
  for (int f = 0; f < n_frames; ++f)
    for (int m = 0; m < n_sources; ++m) {
      if (!d_sources[(size_t)f * n_sources + m]) return lm_fail(LM_E_INVALID, "frame %d: source %d is NULL", f, m);
      ln.src_ptr[f][m] = d_sources[(size_t)f * n_sources + m];
    }
  for (int m = 0; m < n_sources; ++m) ln.has_mask[m] = false;
  ln.user_stream = s; ln.user_stream_valid = true;
  return enqueue_chunk(d, ln, *plan, qs, n_q, n_frames, s);
}

int lm_match_device_multi_lane(lm_detector* d, int lane_index, const void* const* d_sources, int n_sources, int rows,
                               int cols, const lm_query* queries, int n_queries, void* stream, const void** d_records,
                               size_t* record_bytes_capacity) {
  if (!d || !d_sources || !d_records) return lm_fail(LM_E_INVALID, "NULL argument");
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  rc = device_chunk(d, lane_index, d_sources, 1, n_sources, rows, cols, qs, n_queries, 1, (cudaStream_t)stream);
  if (rc != LM_OK) return rc;
  Lane& ln = d->lane[lane_index];
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

int lm_device_result_region(lm_detector* d, int lane_index, const void** base, size_t* frame_stride, int* n_frames) {
  if (!d || !base || !frame_stride || !n_frames) return lm_fail(LM_E_INVALID, "NULL argument");
  if (lane_index < 0 || lane_index >= LM_LANES) return lm_fail(LM_E_INVALID, "lane must be 0..%d", LM_LANES - 1);
  const Lane& ln = d->lane[lane_index];
  if (ln.result.buf.p == nullptr) return lm_fail(LM_E_STATE, "lane %d has no result blocks yet", lane_index);
  *base = block_ptr(ln);
  *frame_stride = ln.result.stride;
  *n_frames = ln.result.frames;
  return LM_OK;
}

int lm_copy_result_block(lm_detector* d, int lane_index, void* d_dst, size_t bytes, void* stream) {
  if (!d || !d_dst) return lm_fail(LM_E_INVALID, "NULL argument");
  if (lane_index < 0 || lane_index >= LM_LANES || d->lane[lane_index].result.buf.p == nullptr) return lm_fail(LM_E_STATE, "lane %d has no result block yet", lane_index);
  const Lane& ln = d->lane[lane_index];
  if (bytes > result_bytes(ln)) return lm_fail(LM_E_INVALID, "block is only %zu bytes", result_bytes(ln));
  CU(cudaMemcpyAsync(d_dst, block_ptr(ln), bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return LM_OK;
}

int lm_match_device_stream(lm_detector* d, const void* const* d_sources, int n_frames, int n_sources, int rows, int cols,
                           const lm_query* queries, int n_queries, void* const* streams, int n_streams, void* d_stage,
                           size_t stage_slot_bytes) {
  if (!d || !d_sources || !streams || n_frames < 0) return lm_fail(LM_E_INVALID, "NULL argument");
  if (n_streams < 1 || n_streams > LM_LANES) return lm_fail(LM_E_INVALID, "number of streams must be 1..%d", LM_LANES);
  Query qs[kMaxQueries];
  int rc = to_queries(queries, n_queries, qs);
  if (rc != LM_OK) return rc;
  const int F = std::max(1, std::min(d->batch_frames, LM_MAX_BATCH));
  for (int first = 0, c = 0; first < n_frames; first += F, ++c) {
    const int lane = c % n_streams, n = std::min(F, n_frames - first);
    cudaStream_t s = (cudaStream_t)streams[lane];
    rc = device_chunk(d, lane, d_sources + (size_t)first * n_sources, n, n_sources, rows, cols, qs, n_queries, std::min(F, n_frames), s);
    if (rc != LM_OK) return rc;
    if (d_stage) {  // heads of the chunk's record blocks -> consecutive slots of the caller's exchange buffer, one strided copy
      const Lane& ln = d->lane[lane];
      if (stage_slot_bytes > result_bytes(ln)) return lm_fail(LM_E_INVALID, "block is only %zu bytes", result_bytes(ln));
      CU(cudaMemcpy2DAsync(static_cast<uint8_t*>(d_stage) + (size_t)first * stage_slot_bytes, stage_slot_bytes, block_ptr(ln),
                           ln.result.stride, stage_slot_bytes, (size_t)n, cudaMemcpyDeviceToDevice, s));
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
  if (!d || n < 0 || (n > 0 && (!images || !d_dst))) return lm_fail(LM_E_INVALID, "NULL argument");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n; ++i) {
    const lm_image& im = images[i];
    if (!im.data || !d_dst[i]) return lm_fail(LM_E_INVALID, "image %d: NULL data / destination", i);
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
  if (!d || (!raw && n_raw) || !out_matches || !out_n) return lm_fail(LM_E_INVALID, "NULL argument");
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
    return lm_fail(LM_E_INVALID, "bad argument");
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
    return lm_fail(LM_E_INVALID, "NULL argument");
  if (p->vote_row_col_step <= 0 || !(p->renderer_radius_step > 0)) return lm_fail(LM_E_INVALID, "voting steps must be positive");
  *out_clusters = nullptr; *out_match_index = nullptr; *out_n = 0;
  // rcd_voting: std::map<std::vector<int>, std::vector<Match>> keyed by (row bin, column bin, depth bin)
  std::map<std::vector<int>, ClusterTmp> bins;
  const float depth_step = (float)p->renderer_radius_step;
  for (size_t i = 0; i < n_matches; ++i) {
    const lm_match_rec& m = matches[i];
    if (m.template_id < 0 || (size_t)m.template_id >= n_templates) return lm_fail(LM_E_INVALID, "match %zu: template_id %d outside the pose tables", i, m.template_id);
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
  if (!oc || !oi) { std::free(oc); std::free(oi); return lm_fail(LM_E_INVALID, "out of host memory"); }
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
  if (!ln.lm_ready || level < 0 || level >= (int)ln.geom.size()) return lm_fail(LM_E_STATE, "no front end built");
  const LevelGeom& g = ln.geom[level];
  out[0] = g.rows; out[1] = g.cols; out[2] = g.T; out[3] = g.W; out[4] = g.H;
  if (plane_stride) *plane_stride = g.plane_stride;
  return LM_OK;
}

// The nibble planes of (level, modality) of frame 0 in the reference's flat order, two positions per byte: a plain download,
// or -- for a column-blocked refinement level -- a download put back into linearize's order on the host (parity taps only).
static int fetch_flat_nibbles(Lane& ln, int level, int modality, std::vector<uint8_t>& flat) {
  const LevelGeom& g = ln.geom[level];
  flat.assign(4 * g.plane_stride, 0);
  if (cudaStreamSynchronize(ln.stream) != cudaSuccess) return lm_fail(LM_E_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
  const uint8_t* src = ln.lmn[level].as<uint8_t>() + (size_t)modality * 4 * g.nib_plane;
  if (!g.Hh) {
    if (cudaMemcpy(flat.data(), src, flat.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
      return lm_fail(LM_E_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return LM_OK;
  }
  std::vector<uint8_t> blocked(4 * g.nib_plane);
  if (cudaMemcpy(blocked.data(), src, blocked.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
    return lm_fail(LM_E_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
  const size_t WH = (size_t)g.W * g.H;
  for (int o = 0; o < 8; ++o) {
    const uint8_t* b = blocked.data() + (size_t)o * (g.nib_plane / 2);
    uint8_t* f = flat.data() + (size_t)o * (g.plane_stride / 2);
    for (int ph = 0; ph < g.T * g.T; ++ph)
      for (int r = 0; r < g.H; ++r)
        for (int c = 0; c < g.W; ++c) {
          const size_t ti = tiled_nibble_index(g.W, g.Hh, ph, r, c), fi = (size_t)ph * WH + (size_t)r * g.W + c;
          const uint8_t v = (uint8_t)((b[ti >> 1] >> (4 * (ti & 1))) & 15);
          f[fi >> 1] |= (uint8_t)(v << (4 * (fi & 1)));
        }
    // the halo rows must repeat the next phase's first rows (zero below the last phase) and the plane must end in zeros:
    // a violation is reported as a value no response can have, so that the parity tests see it
    bool ok = true;
    for (int ph = 0; ph < g.T * g.T && ok; ++ph)
      for (int r = 0; r < 16 && ok; ++r)
        for (int c = 0; c < g.W && ok; ++c) {
          const size_t hi = tiled_nibble_index(g.W, g.Hh, ph, g.H + r, c);
          const uint8_t hv = (uint8_t)((b[hi >> 1] >> (4 * (hi & 1))) & 15);
          uint8_t want = 0;
          if (ph + 1 < g.T * g.T) {
            const size_t ni = tiled_nibble_index(g.W, g.Hh, ph + 1, r, c);
            want = (uint8_t)((b[ni >> 1] >> (4 * (ni & 1))) & 15);
          }
          ok = hv == want;
        }
    for (size_t i = (size_t)g.T * g.T * g.W * g.Hh; i < g.nib_plane && ok; ++i) ok = ((b[i >> 1] >> (4 * (i & 1))) & 15) == 0;
    if (!ok) f[0] |= 0x0f;
  }
  return LM_OK;
}

long lm_debug_fetch(lm_detector* d, int stage, int level, int modality, void* dst) {
  Lane& ln = d->lane[0];
  if (!ln.front_valid) return lm_fail(LM_E_STATE, "no front end built");
  if (level < 0 || level >= d->model.levels() || modality < 0 || modality >= d->model.M()) return lm_fail(LM_E_INVALID, "level/modality out of range");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  const LevelGeom& g = ln.geom[level];
  const size_t n = (size_t)g.rows * g.cols;
  const void* src = nullptr;
  size_t bytes = 0;
  switch (stage) {
    case LM_STAGE_QUANTIZED: src = ln.quantized[level][modality].buf.p; bytes = n; break;
    case LM_STAGE_QUANT_RAW: src = ln.quant_raw[level][modality].buf.p; bytes = n; break;
    case LM_STAGE_MAGNITUDE:
      if (d->model.mods[modality].type != LM_COLOR_GRADIENT) return lm_fail(LM_E_INVALID, "magnitude exists for ColorGradient only");
      src = ln.mag[level][modality].buf.p; bytes = n * 4; break;
    case LM_STAGE_SPREAD:
    case LM_STAGE_RESPONSE:
      if (!ln.debug_taps_written) return lm_fail(LM_E_STATE, "enable lm_set_option(det, \"debug_taps\", 1) before matching");
      src = stage == LM_STAGE_SPREAD ? ln.spread[level][modality].p : ln.response[level][modality].p;
      bytes = stage == LM_STAGE_SPREAD ? n : 8 * n; break;
    case LM_STAGE_LINEAR:
      bytes = 8 * g.plane_stride;
      if (!ln.bytes_valid[level]) {  // only the packed planes exist: unpack them
        if (dst) {
          std::vector<uint8_t> packed;
          if (fetch_flat_nibbles(ln, level, modality, packed) != LM_OK) return LM_E_CUDA;
          uint8_t* o = static_cast<uint8_t*>(dst);
          for (size_t i = 0; i < bytes / 2; ++i) { o[2 * i] = packed[i] & 15; o[2 * i + 1] = packed[i] >> 4; }
        }
        return (long)bytes;
      }
      src = ln.lmem[level].as<uint8_t>() + (size_t)modality * 8 * g.plane_stride; break;
    case LM_STAGE_LINEAR_PACKED:
      if (!ln.nibbles_valid[level]) return lm_fail(LM_E_STATE, "level %d has no packed planes", level);
      bytes = 4 * g.plane_stride;
      if (g.Hh) {   // column-blocked on the device: handed out in the reference's flat order
        if (dst) {
          std::vector<uint8_t> packed;
          if (fetch_flat_nibbles(ln, level, modality, packed) != LM_OK) return LM_E_CUDA;
          std::memcpy(dst, packed.data(), bytes);
        }
        return (long)bytes;
      }
      src = ln.lmn[level].as<uint8_t>() + (size_t)modality * 4 * g.plane_stride; break;
    default: return lm_fail(LM_E_INVALID, "unknown stage %d", stage);
  }
  if (dst) {
    if (cudaStreamSynchronize(ln.stream) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
      return lm_fail(LM_E_CUDA, "debug fetch failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return (long)bytes;
}

int lm_debug_coarse_map(lm_detector* d, const char* class_id, int template_id, uint16_t* dst) {
  Lane& ln = d->lane[0];
  if (!ln.front_valid) return lm_fail(LM_E_STATE, "no front end built");
  if (set_device(d) != LM_OK) return LM_E_CUDA;
  int prc = ensure_pack(d, ln);
  if (prc != LM_OK) return prc;
  Pack& pk = d->pack;
  int local = -1;
  for (const Pack::ClassRange& cr : pk.classes)
    if (cr.id == class_id)
      for (size_t k = 0; k < cr.local.size(); ++k)
        if (pk.h_ctpl[cr.local[k]].template_id == template_id) local = (int)cr.local[k];
  if (local < 0) return lm_fail(LM_E_NOTFOUND, "class '%s' template %d is not on this shard", class_id, template_id);
  const LevelGeom& gc = ln.geom.back();
  const int WH = gc.W * gc.H;
  if (ensure_match_buffers(d, ln, std::max<uint32_t>(ln.cand_cap, kCandPerFrame), std::max<uint32_t>(ln.out_cap, kOutPerFrame), 1) != LM_OK) return LM_E_CUDA;
  const int pass_pos = coarse_positions_per_pass();
  const int P = pk.h_ctpl[local].P;
  std::vector<uint2> tl;
  for (int pass = 0; pass * pass_pos < P; ++pass) tl.push_back(make_uint2(0u, (uint32_t)pass));
  if (ln.dump.ensure((size_t)WH * 2) != LM_OK) return LM_E_CUDA;
  WorkItem it = {(uint32_t)local, 0u};
  CU(cudaMemsetAsync(ln.dump.p, 0, (size_t)WH * 2, ln.stream));
  if (begin_chunk(d, ln, 1, 1, ln.stream) != LM_OK) return LM_E_CUDA;  // frame 0 = the front end built last
  std::vector<uint32_t> recs;
  int rec_words = 0, max_feat = 0;
  std::vector<WorkItem> one(1, it);
  if (build_tile_records(pk, one, tl, pass_pos, d->model.M(), recs, &rec_words, &max_feat) != LM_OK) return LM_E_INVALID;
  if (ln.dbg_recs.ensure(recs.size() * 4 + 64) != LM_OK) return LM_E_CUDA;
  if (!recs.empty()) CU(cudaMemcpyAsync(ln.dbg_recs.p, recs.data(), recs.size() * 4, cudaMemcpyHostToDevice, ln.stream));
  // threshold 1e30 -> raw threshold saturates: nothing becomes a candidate, the kernel only dumps its accumulators
  CoarseParams cp;
  std::memset(&cp, 0, sizeof(cp));
  for (int q = 0; q < LM_MAX_QUERIES; ++q) cp.thr.v[q] = 1e30f;
  cp.lmn = ln.lmn[d->model.levels() - 1].as<uint8_t>(); cp.lmn_stride = ln.lmn[d->model.levels() - 1].stride;
  cp.recs = ln.dbg_recs.as<uint32_t>(); cp.rec_words = rec_words; cp.n_tiles = (int)tl.size(); cp.max_feat = max_feat; cp.n_full = cp.n_tiles;
  cp.ctl = ln.ctl.as<BatchCtl>(); cp.cand = ln.cand.as<Cand>(); cp.cand_cap = 0; cp.M = d->model.M();
  cp.dump = ln.dump.as<uint16_t>(); cp.dump_stride = WH;
  launch_similarity_coarse(cp, 1, ln.stream);
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
