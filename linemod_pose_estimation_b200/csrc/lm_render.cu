// lm_render.cu -- z-buffer rasteriser for template generation and hypothesis checks.
//
// Replaces, for this library, the renders that /root/reference/src/renderer.cpp:239-252 (RendererIterator::render) and
// src/rgbdDetector.cpp:165 (renderDepthOnly) obtain from `object_recognition_renderer` (OpenGL + assimp; an external
// dependency that is not part of the reference tree).  The specification -- pinhole camera looking at the object
// origin, pixel-centre sampling, perspective-correct depth, nearest fragment wins with ties to the lower triangle
// index, depth in u16 millimetres, mask 255, head-light grey shading -- is written out in oracle/render_oracle.cpp,
// whose scalar implementation the CUDA one below matches bit for bit (tests/test_gpu_train.py).
//
//   k_raster_tris     one warp per (triangle, view): every lane repeats the (cheap) vertex transform and set-up, then the
//                     lanes stride the columns of the triangle's bounding box row by row; a covered sample issues one
//                     64-bit atomicMin of (f32 depth bits << 32 | triangle index) -- positive floats order like their
//                     bit patterns, so the z-buffer result does not depend on the order the triangles arrive in.
//   k_raster_resolve  one thread per pixel of every view: depth / mask / BGR of the winning fragment and the mask's
//                     bounding box (warp-reduced, four atomics per warp that saw a covered pixel).
// Arithmetic is IEEE f32 without contraction (-fmad=false), in the order the specification states.
#include "lm_kernels.cuh"

namespace lmk {
namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float px, float py) {
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}

__global__ void __launch_bounds__(256) k_raster_tris(const float* __restrict__ tris, int n_tri,
                                                     const RenderView* __restrict__ views, int n_views, RenderCamera cam,
                                                     unsigned long long* __restrict__ zbuf, float* __restrict__ nz_abs) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)n_tri * n_views) return;
  const int v = (int)(wid / n_tri), k = (int)(wid % n_tri);
  const RenderView vw = views[v];
  const int W = cam.width, H = cam.height;
  float X[3], Y[3], Z[3], px[3], py[3];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float* p = tris + 9 * (size_t)k + 3 * i;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    X[i] = ((vw.R[0] * x + vw.R[1] * y) + vw.R[2] * z) + vw.t[0];
    Y[i] = ((vw.R[3] * x + vw.R[4] * y) + vw.R[5] * z) + vw.t[1];
    Z[i] = ((vw.R[6] * x + vw.R[7] * y) + vw.R[8] * z) + vw.t[2];
    if (!(Z[i] > cam.z_near)) ok = false;
  }
  if (lane == 0) {  // unit normal in the camera frame -> shading term of this (view, triangle)
    const float ax = X[1] - X[0], ay = Y[1] - Y[0], az = Z[1] - Z[0];
    const float bx = X[2] - X[0], by = Y[2] - Y[0], bz = Z[2] - Z[0];
    const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const float nn = sqrtf((nx * nx + ny * ny) + nz * nz);
    nz_abs[(size_t)v * n_tri + k] = nn > 0.f ? fabsf(nz / nn) : 0.f;
  }
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    px[i] = (cam.fx * X[i]) / Z[i] + cam.cx;
    py[i] = (cam.fy * Y[i]) / Z[i] + cam.cy;
  }
  const float area = edge_fn(px[0], py[0], px[1], py[1], px[2], py[2]);
  if (area == 0.f || area != area) return;
  const float minx = fminf(px[0], fminf(px[1], px[2])), maxx = fmaxf(px[0], fmaxf(px[1], px[2]));
  const float miny = fminf(py[0], fminf(py[1], py[2])), maxy = fmaxf(py[0], fmaxf(py[1], py[2]));
  if (!(maxx >= 0.f) || !(maxy >= 0.f) || !(minx <= (float)W) || !(miny <= (float)H)) return;
  const int x_lo = (int)fmaxf(0.f, floorf(minx)), x_hi = (int)fminf((float)(W - 1), floorf(maxx));
  const int y_lo = (int)fmaxf(0.f, floorf(miny)), y_hi = (int)fminf((float)(H - 1), floorf(maxy));
  const float iz0 = 1.0f / Z[0], iz1 = 1.0f / Z[1], iz2 = 1.0f / Z[2];
  unsigned long long* zb = zbuf + (size_t)v * W * H;
  for (int iy = y_lo; iy <= y_hi; ++iy) {
    const float sy = (float)iy + 0.5f;
    for (int ix = x_lo + lane; ix <= x_hi; ix += 32) {
      const float sx = (float)ix + 0.5f;
      const float w0 = edge_fn(px[1], py[1], px[2], py[2], sx, sy);
      const float w1 = edge_fn(px[2], py[2], px[0], py[0], sx, sy);
      const float w2 = edge_fn(px[0], py[0], px[1], py[1], sx, sy);
      const bool inside = area > 0.f ? (w0 >= 0.f && w1 >= 0.f && w2 >= 0.f) : (w0 <= 0.f && w1 <= 0.f && w2 <= 0.f);
      if (!inside) continue;
      const float b0 = w0 / area, b1 = w1 / area, b2 = w2 / area;
      const float iz = (b0 * iz0 + b1 * iz1) + b2 * iz2;
      const float z = 1.0f / iz;
      if (!(z > 0.f) || z > cam.z_max) continue;
      const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned)k;
      atomicMin(zb + (size_t)iy * W + ix, key);
    }
  }
}

__global__ void __launch_bounds__(256) k_raster_resolve(const unsigned long long* __restrict__ zbuf,
                                                        const float* __restrict__ nz_abs, int n_tri, int W, int H,
                                                        int n_views, RenderTargets out) {
  const int v = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = W * H;
  const bool in_img = i < n;
  bool covered = false;
  int ix = 0, iy = 0;
  if (in_img) {
    const unsigned long long key = zbuf[(size_t)v * n + i];
    uint16_t d = 0;
    uint8_t m = 0, g = 0;
    if (key != ~0ull) {
      covered = true;
      const float z = __uint_as_float((unsigned)(key >> 32));
      const float mm = rintf(z * 1000.0f);
      d = mm > 65535.f ? (uint16_t)65535 : (uint16_t)mm;
      m = 255;
      g = (uint8_t)(int)(40.0f + 200.0f * nz_abs[(size_t)v * n_tri + (unsigned)key]);
      iy = i / W; ix = i - iy * W;
    }
    if (out.depth) out.depth[(size_t)v * out.depth_stride + i] = d;
    if (out.mask) out.mask[(size_t)v * out.mask_stride + i] = m;
    if (out.bgr) {
      uint8_t* p = out.bgr + (size_t)v * out.bgr_stride + 3 * (size_t)i;
      p[0] = g; p[1] = g; p[2] = g;
    }
  }
  const int big = 1 << 30;
  const int x0 = __reduce_min_sync(kFull, covered ? ix : big), y0 = __reduce_min_sync(kFull, covered ? iy : big);
  const int x1 = __reduce_max_sync(kFull, covered ? ix : -1), y1 = __reduce_max_sync(kFull, covered ? iy : -1);
  if ((threadIdx.x & 31) == 0 && x1 >= 0) {
    int* r = out.rect + 4 * v;  // x_min, y_min, x_max, y_max (initialised to W, H, -1, -1)
    atomicMin(r + 0, x0); atomicMin(r + 1, y0); atomicMax(r + 2, x1); atomicMax(r + 3, y1);
  }
}

__global__ void k_rect_init(int* rect, int n_views, int W, int H) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n_views) { rect[4 * v] = W; rect[4 * v + 1] = H; rect[4 * v + 2] = -1; rect[4 * v + 3] = -1; }
}

// Bounding box of a caller-provided mask (non-zero pixels), same [x_min, y_min, x_max, y_max] encoding.
__global__ void __launch_bounds__(256) k_mask_rect(const uint8_t* __restrict__ mask, int W, int H, int* rect) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool covered = i < W * H && mask[i] != 0;
  const int iy = covered ? i / W : 0, ix = covered ? i - iy * W : 0;
  const int big = 1 << 30;
  const int x0 = __reduce_min_sync(kFull, covered ? ix : big), y0 = __reduce_min_sync(kFull, covered ? iy : big);
  const int x1 = __reduce_max_sync(kFull, covered ? ix : -1), y1 = __reduce_max_sync(kFull, covered ? iy : -1);
  if ((threadIdx.x & 31) == 0 && x1 >= 0) {
    atomicMin(rect + 0, x0); atomicMin(rect + 1, y0); atomicMax(rect + 2, x1); atomicMax(rect + 3, y1);
  }
}

// depth_diff of /root/reference/src/rgbdDetector.cpp:236-283: sum of |template - scene| (u16 saturating subtract read
// back as s16) and the count of pixels where the template mask and the byte-saturated scene depth are both non-zero.
__global__ void __launch_bounds__(256) k_depth_diff(const uint16_t* __restrict__ scene, int scene_cols,
                                                    const uint16_t* __restrict__ templ, const uint8_t* __restrict__ tmask,
                                                    int templ_cols, int x, int y, int tx, int ty, int w, int h,
                                                    unsigned long long* out /* [sum, count] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned sum = 0, cnt = 0;
  if (i < w * h) {
    const int r = i / w, c = i - r * w;
    const unsigned s = scene[(size_t)(y + r) * scene_cols + x + c];
    const unsigned tv = templ[(size_t)(ty + r) * templ_cols + tx + c];
    const unsigned smask = s > 255u ? 255u : s;
    if (tmask[(size_t)(ty + r) * templ_cols + tx + c] & smask) {
      const unsigned d = tv > s ? tv - s : 0u;
      const int sd = (int)(short)(unsigned short)d;
      sum = (unsigned)(sd < 0 ? -sd : sd);
      cnt = 1;
    }
  }
  sum = __reduce_add_sync(kFull, sum);
  cnt = __reduce_add_sync(kFull, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(out, (unsigned long long)sum); atomicAdd(out + 1, (unsigned long long)cnt); }
}

}  // namespace

void launch_raster(const float* tris, int n_tri, const RenderView* views, int n_views, const RenderCamera& cam,
                   unsigned long long* zbuf, float* nz_abs, const RenderTargets& out, cudaStream_t s) {
  const size_t n = (size_t)cam.width * cam.height;
  cudaMemsetAsync(zbuf, 0xff, n * sizeof(unsigned long long) * n_views, s);
  k_rect_init<<<(n_views + 127) / 128, 128, 0, s>>>(out.rect, n_views, cam.width, cam.height);
  const long long warps = (long long)n_tri * n_views;
  if (warps > 0) k_raster_tris<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(tris, n_tri, views, n_views, cam, zbuf, nz_abs);
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_views);
  k_raster_resolve<<<grid, 256, 0, s>>>(zbuf, nz_abs, n_tri, cam.width, cam.height, n_views, out);
}

void launch_mask_rect(const uint8_t* mask, int W, int H, int* rect, cudaStream_t s) {
  k_rect_init<<<1, 32, 0, s>>>(rect, 1, W, H);
  k_mask_rect<<<(W * H + 255) / 256, 256, 0, s>>>(mask, W, H, rect);
}

void launch_depth_diff(const uint16_t* scene, int scene_cols, const uint16_t* templ, const uint8_t* tmask, int templ_cols,
                       int x, int y, int tx, int ty, int w, int h, unsigned long long* out, cudaStream_t s) {
  cudaMemsetAsync(out, 0, 2 * sizeof(unsigned long long), s);
  if (w > 0 && h > 0)
    k_depth_diff<<<(w * h + 255) / 256, 256, 0, s>>>(scene, scene_cols, templ, tmask, templ_cols, x, y, tx, ty, w, h, out);
}

}  // namespace lmk
