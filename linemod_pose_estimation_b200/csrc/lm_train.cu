// lm_train.cu -- batched template extraction on the GPU.
//
// Restates, for many training views at a time, the per-view tail of [OCV] Detector::addTemplate that the reference
// runs once per rendered view at /root/reference/src/renderer.cpp:308 (and src/renderer_only_image.cpp:266):
//   ColorGradientPyramid::extractTemplate   candidates = silhouette ring (mask - erode(mask)) with a quantised orientation and
//                                           magnitude > strong_threshold^2, score = magnitude
//   DepthNormalPyramid::extractTemplate     candidates = mask eroded twice, one-hot normal bin, chessboard distance to the
//                                           bin's border >= extract_threshold, score = distance / candidates of that bin
//   QuantizedPyramid::selectScatteredFeatures  stable sort by score (descending), greedy selection with a minimum spacing
//                                           that is relaxed by one pixel per sweep
// The quantised maps come from the production front end (k_cg_fused / k_dn_fused) run on the view; one segment = one
// (view, level, modality).  Results are bit-identical to the sequential host path (lm_host.cpp) and to the oracle.
//
//   k_train_cg       thread per pixel, every level in one grid: ring test (3x3 minimum of the decimated mask, replicated
//                    border), candidate key = ~f32 bits(magnitude) << 32 | raster << 3 | label, appended with one atomic
//   k_train_dn_pb    pb = (5x5 minimum of the decimated mask != 0) ? normal : 0, and the eroded area
//   k_train_dn_runs  per pixel the horizontal distance to the nearest pixel of its row with another pb value
//   k_train_dn_dist  chessboard distance of a candidate pixel to the nearest pixel outside its bin's plane (exact
//                    L-infinity distance = what cv::distanceTransform(DIST_C, 3) computes; frames without any such pixel
//                    take the reference's "far border" value): min over rows of max(row offset, run-table distance) when
//                    the map is one-hot (always, with the stock NORMAL_LUT), an expanding ring search otherwise;
//                    per-bin candidate counts
//   k_train_dn_keys  score = distance / bin count -> sort key                                  (once per batch)
//   k_train_sort     block per segment: bitonic sort of the 64-bit keys (shared memory up to 4096 keys, global above).
//                    The raster index in the low word makes every key unique, so "stable sort by score" is a plain sort
//   k_train_select   block per segment: the greedy scattered selection, 256 candidates tested per step against the
//                    features chosen so far; accepted candidates are committed in candidate order and the later ones of
//                    the step re-tested against each newly accepted feature, which reproduces the sequential loop exactly
#include "lm_kernels.cuh"

namespace lmk {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int level_of_block(const TrainViewParams& P, int b) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < LM_MAX_LEVELS; ++i)
    if (i < P.n_levels && b >= P.lv[i].block_begin) l = i;
  return l;
}

// decimated mask of level l ([OCV] pyrDown NN-resizes the mask: plain index decimation), clamped = replicated border
__device__ __forceinline__ uint8_t mask_at(const TrainViewParams& P, int l, int rows, int cols, int y, int x) {
  y = min(max(y, 0), rows - 1); x = min(max(x, 0), cols - 1);
  return __ldg(P.mask0 + (size_t)(y << l) * P.cols0 + (x << l));
}

__global__ void __launch_bounds__(256) k_train_cg(const TrainViewParams P, TrainSeg* __restrict__ segs,
                                                  unsigned long long* __restrict__ pool) {
  const int l = level_of_block(P, blockIdx.x);
  const TrainLevel lv = P.lv[l];
  const int i = (blockIdx.x - lv.block_begin) * 256 + threadIdx.x;
  if (i >= lv.rows * lv.cols) return;
  const int y = i / lv.cols, x = i - y * lv.cols;
  const uint8_t m = mask_at(P, l, lv.rows, lv.cols, y, x);
  if (!m) return;
  uint8_t er = 255;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) er = min(er, mask_at(P, l, lv.rows, lv.cols, y + dy, x + dx));
  if (!(m > er)) return;  // not on the silhouette ring
  const uint8_t q = lv.quant[i];
  const float mag = lv.mag[i];
  if (!q || !(mag > P.thr_sq)) return;
  TrainSeg& sg = segs[lv.seg];
  const uint32_t idx = atomicAdd(&sg.count, 1u);
  if (idx < sg.cap)
    pool[sg.off + idx] = ((unsigned long long)(~__float_as_uint(mag)) << 32) | ((uint32_t)i << 3) | (uint32_t)(__ffs(q) - 1);
}

__global__ void __launch_bounds__(256) k_train_dn_pb(const TrainViewParams P, TrainSeg* __restrict__ segs) {
  const int l = level_of_block(P, blockIdx.x);
  const TrainLevel lv = P.lv[l];
  const int i = (blockIdx.x - lv.block_begin) * 256 + threadIdx.x;
  bool inner = false;
  if (i < lv.rows * lv.cols) {
    const int y = i / lv.cols, x = i - y * lv.cols;
    uint8_t er = 255;  // two 3x3 erosions with a replicated border == one clamped 5x5 minimum
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) er = min(er, mask_at(P, l, lv.rows, lv.cols, y + dy, x + dx));
    inner = er != 0;
    const uint8_t q = inner ? lv.quant[i] : (uint8_t)0;
    P.pb[l][i] = q;
    if (q & (q - 1)) atomicOr(&segs[lv.seg].flags, 1u);  // more than one bit: only with an injected NORMAL_LUT
  }
  const unsigned n = __popc(__ballot_sync(kFull, inner));
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(&segs[lv.seg].area, n);
}

constexpr int kRunInf = 0xffff;

// Horizontal distance (in pixels, >= 1) from every non-zero pb pixel to the nearest pixel of its row holding another
// value; a side on which the run of equal values reaches the frame edge has no such pixel (the frame border is not a
// zero of the plane), kRunInf when that holds on both sides.
__global__ void __launch_bounds__(256) k_train_dn_runs(const TrainViewParams P) {
  const int l = level_of_block(P, blockIdx.x);
  const TrainLevel lv = P.lv[l];
  const int i = (blockIdx.x - lv.block_begin) * 256 + threadIdx.x;
  if (i >= lv.rows * lv.cols) return;
  const uint8_t* __restrict__ pb = P.pb[l];
  const uint8_t q = pb[i];
  int h = 0;
  if (q) {
    const int cols = lv.cols;
    const int y = i / cols, x = i - y * cols;
    const uint8_t* row = pb + (size_t)y * cols;
    int left = kRunInf, right = kRunInf;
    for (int c = 1; c <= x; ++c)
      if (row[x - c] != q) { left = c; break; }
    for (int c = 1; x + c < cols && c < left; ++c)
      if (row[x + c] != q) { right = c; break; }
    h = min(left, right);
  }
  P.runs[l][i] = (uint16_t)h;
}

__global__ void __launch_bounds__(256) k_train_dn_dist(const TrainViewParams P, TrainSeg* __restrict__ segs,
                                                       unsigned long long* __restrict__ pool) {
  const int l = level_of_block(P, blockIdx.x);
  const TrainLevel lv = P.lv[l];
  const int i = (blockIdx.x - lv.block_begin) * 256 + threadIdx.x;
  if (i >= lv.rows * lv.cols) return;
  const uint8_t* __restrict__ pb = P.pb[l];
  const uint8_t q = pb[i];
  if (q == 0 || q == 255 || (q & (q - 1))) return;  // outside the eroded mask, or not a one-hot bin
  const int rows = lv.rows, cols = lv.cols;
  const int y = i / cols, x = i - y * cols;
  // exact chessboard distance to the nearest in-frame pixel whose plane value is zero
  const int k_out = max(max(x, y), max(cols - 1 - x, rows - 1 - y));  // beyond this the ring is outside the frame
  int d = 0;
  const bool one_hot_map = (segs[lv.seg].flags & 1u) == 0;
  if (one_hot_map) {
    // Every value of the map is zero or one-hot, so "outside bin q's plane" == "another value" and the distance separates:
    // d = min over rows y' of max(|y' - y|, horizontal distance in row y' from column x to another value), the latter
    // read from the run table (0 when (y', x) itself holds another value).  O(d) reads instead of the ring search's O(d^2).
    const uint16_t* __restrict__ runs = P.runs[l];
    int best = runs[i];
    for (int k = 1; k < best && (y - k >= 0 || y + k < rows); ++k) {
      if (y - k >= 0) {
        const size_t j = (size_t)(y - k) * cols + x;
        best = min(best, max(k, pb[j] == q ? (int)runs[j] : 0));
      }
      if (y + k < rows) {
        const size_t j = (size_t)(y + k) * cols + x;
        best = min(best, max(k, pb[j] == q ? (int)runs[j] : 0));
      }
    }
    d = best >= kRunInf ? 0 : best;  // 0: no pixel of another value anywhere -> the "far border" value below
  }
  for (int k = 1; k <= k_out && !d && !one_hot_map; ++k) {
    bool zero = false;
    const int y0 = y - k, y1 = y + k, x0 = x - k, x1 = x + k;
    const int xa = max(x0, 0), xb = min(x1, cols - 1);
    if (y0 >= 0) {
      const uint8_t* r = pb + (size_t)y0 * cols;
      for (int xx = xa; xx <= xb; ++xx) zero |= !(r[xx] & q);
    }
    if (y1 < rows && !zero) {
      const uint8_t* r = pb + (size_t)y1 * cols;
      for (int xx = xa; xx <= xb; ++xx) zero |= !(r[xx] & q);
    }
    if (!zero) {
      const int ya = max(y0 + 1, 0), yb = min(y1 - 1, rows - 1);
      if (x0 >= 0)
        for (int yy = ya; yy <= yb; ++yy) zero |= !(pb[(size_t)yy * cols + x0] & q);
      if (x1 < cols && !zero)
        for (int yy = ya; yy <= yb; ++yy) zero |= !(pb[(size_t)yy * cols + x1] & q);
    }
    if (zero) d = k;
  }
  // 16.16 fixed point like the reference's chamfer; a plane without any zero sees only the "infinitely far" frame
  const int kOne = 1 << 16, kFar = 0x7fffffff >> 2;
  const int fixed = d ? d * kOne : kFar + kOne * min(min(x + 1, y + 1), min(cols - x, rows - y));
  const float score = __fmul_rn((float)fixed, 1.0f / 65536.0f);
  if (!(score >= (float)P.extract_threshold[l])) return;
  TrainSeg& sg = segs[lv.seg];
  const int label = __ffs(q) - 1;
  atomicAdd(&sg.per_label[label], 1u);
  const uint32_t idx = atomicAdd(&sg.count, 1u);
  if (idx < sg.cap) pool[sg.off + idx] = ((unsigned long long)__float_as_uint(score) << 32) | ((uint32_t)i << 3) | (uint32_t)label;
}

// DepthNormal segments: (distance bits | lo) -> (~bits(distance / candidates of the label) | lo)
__global__ void __launch_bounds__(256) k_train_dn_keys(const TrainSeg* __restrict__ segs, int n_segs,
                                                       unsigned long long* __restrict__ pool) {
  const int s = blockIdx.y;
  if (s >= n_segs) return;
  const TrainSeg sg = segs[s];
  if (sg.type != LM_DEPTH_NORMAL) return;
  const uint32_t n = min(sg.count, sg.cap);
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const unsigned long long k = pool[sg.off + i];
    const uint32_t lo = (uint32_t)k;
    const float score = __fdiv_rn(__uint_as_float((uint32_t)(k >> 32)), (float)sg.per_label[lo & 7]);
    pool[sg.off + i] = ((unsigned long long)(~__float_as_uint(score)) << 32) | lo;
  }
}

constexpr int kSortThreads = 512;
constexpr int kSortSmem = 4096;

__device__ __forceinline__ void bitonic_pass(unsigned long long* a, uint32_t n_pad, uint32_t k, uint32_t j) {
  for (uint32_t t = threadIdx.x; t < n_pad / 2; t += kSortThreads) {
    const uint32_t lo = 2 * t - (t & (j - 1));  // t with a zero bit inserted at log2(j)
    const uint32_t hi = lo + j;
    const unsigned long long u = a[lo], v = a[hi];
    const bool up = (lo & k) == 0;
    if ((u > v) == up) { a[lo] = v; a[hi] = u; }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSortThreads) k_train_sort(const TrainSeg* __restrict__ segs,
                                                             unsigned long long* __restrict__ pool) {
  __shared__ unsigned long long sm[kSortSmem];
  const TrainSeg sg = segs[blockIdx.x];
  const uint32_t n = min(sg.count, sg.cap);
  if (n < 2) return;
  uint32_t n_pad = 2;
  while (n_pad < n) n_pad <<= 1;  // <= cap (a power of two)
  unsigned long long* g = pool + sg.off;
  const bool in_smem = n_pad <= (uint32_t)kSortSmem;
  unsigned long long* a = in_smem ? sm : g;
  if (in_smem) {
    for (uint32_t t = threadIdx.x; t < n_pad; t += kSortThreads) sm[t] = t < n ? g[t] : ~0ull;
  } else {
    for (uint32_t t = n + threadIdx.x; t < n_pad; t += kSortThreads) g[t] = ~0ull;
  }
  __syncthreads();
  for (uint32_t k = 2; k <= n_pad; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) bitonic_pass(a, n_pad, k, j);
  if (in_smem)
    for (uint32_t t = threadIdx.x; t < n; t += kSortThreads) g[t] = sm[t];
}

constexpr int kSelectThreads = 256;

// Block per segment.  The reference's loop looks at one candidate at a time; here 256 consecutive candidates are tested
// against the features chosen so far in one step, and the (rare) acceptances are committed in candidate order: the first
// passing candidate is taken, the later ones of the step are re-tested against it, and so on -- the same decisions as the
// sequential loop.  A sweep over 14 000 DepthNormal candidates is 55 steps instead of 14 000.
__global__ void __launch_bounds__(kSelectThreads) k_train_select(TrainSeg* __restrict__ segs, int n_segs,
                                                                const unsigned long long* __restrict__ pool,
                                                                uint32_t* __restrict__ out_feats) {
  __shared__ int s_x[64], s_y[64];
  __shared__ unsigned s_ballot[kSelectThreads / 32];
  __shared__ int s_ax, s_ay;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = blockIdx.x;
  if (s >= n_segs) return;
  const TrainSeg sg = segs[s];
  const int n = (int)min(sg.count, sg.cap), want = sg.nf;
  if (sg.count > sg.cap || n < want || want <= 0 || want > 63) {  // [OCV] extractTemplate returns false
    if (tid == 0) segs[s].n_sel = sg.count > sg.cap ? -2 : -1;
    return;
  }
  float distance;
  if (sg.type == LM_COLOR_GRADIENT) distance = (float)(n / want + 1);
  else distance = __fadd_rn(__fdiv_rn(__fsqrt_rn((float)sg.area), __fsqrt_rn((float)want)), 1.5f);
  float dist_sq = __fmul_rn(distance, distance);
  const unsigned long long* keys = pool + sg.off;
  int n_sel = 0, i = 0;  // block-uniform
  while (n_sel < want) {
    const int cnt = min(kSelectThreads, n - i);
    const bool have = tid < cnt;
    const uint32_t lo = have ? (uint32_t)keys[i + tid] : 0u;
    const int raster = (int)(lo >> 3);
    const int y = raster / sg.cols, x = raster - y * sg.cols;
    bool ok = have;
    for (int j = 0; j < n_sel && ok; ++j) {
      const int dx = x - s_x[j], dy = y - s_y[j];
      ok = (float)(dx * dx + dy * dy) >= dist_sq;
    }
    for (;;) {  // commit the passing candidates of this step in order
      const unsigned b = __ballot_sync(kFull, ok);
      if (lane == 0) s_ballot[warp] = b;
      __syncthreads();
      int first = -1;
#pragma unroll
      for (int w = kSelectThreads / 32 - 1; w >= 0; --w)
        if (s_ballot[w]) first = w * 32 + __ffs(s_ballot[w]) - 1;
      if (first < 0) break;  // block-uniform
      if (tid == first) {
        s_ax = x; s_ay = y;
        s_x[n_sel] = x; s_y[n_sel] = y;
        out_feats[(size_t)s * 64 + n_sel] = (uint32_t)x | ((uint32_t)y << 13) | ((lo & 7u) << 26);
      }
      __syncthreads();
      ++n_sel;
      if (n_sel == want) break;
      const int dx = x - s_ax, dy = y - s_ay;
      ok = ok && tid > first && (float)(dx * dx + dy * dy) >= dist_sq;
    }
    __syncthreads();  // s_ballot / s_x are rewritten by the next step
    i += cnt;
    if (i == n) {  // wrap: relax the spacing by one pixel and sweep again
      i = 0;
      distance = __fsub_rn(distance, 1.0f);
      dist_sq = __fmul_rn(distance, distance);
    }
  }
  if (tid == 0) segs[s].n_sel = n_sel;
}

}  // namespace

int train_blocks(int rows, int cols) { return (rows * cols + 255) / 256; }

void launch_train_cg(const TrainViewParams& p, int total_blocks, TrainSeg* segs, unsigned long long* pool, cudaStream_t s) {
  k_train_cg<<<total_blocks, 256, 0, s>>>(p, segs, pool);
}
void launch_train_dn(const TrainViewParams& p, int total_blocks, TrainSeg* segs, unsigned long long* pool, cudaStream_t s) {
  k_train_dn_pb<<<total_blocks, 256, 0, s>>>(p, segs);
  k_train_dn_runs<<<total_blocks, 256, 0, s>>>(p);
  k_train_dn_dist<<<total_blocks, 256, 0, s>>>(p, segs, pool);
}
void launch_train_finish(TrainSeg* segs, int n_segs, unsigned long long* pool, uint32_t* out_feats, cudaStream_t s) {
  if (n_segs <= 0) return;
  k_train_dn_keys<<<dim3(8, (unsigned)n_segs), 256, 0, s>>>(segs, n_segs, pool);
  k_train_sort<<<n_segs, kSortThreads, 0, s>>>(segs, pool);
  k_train_select<<<n_segs, kSelectThreads, 0, s>>>(segs, n_segs, pool, out_feats);
}

}  // namespace lmk
