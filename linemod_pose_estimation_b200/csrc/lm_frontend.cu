// lm_frontend.cu -- per-frame front end of Detector::match on sm_100a: quantisation of both modalities, pyramid,
// OR-spreading, response maps and linearisation.  Integer / byte work, HBM- and L2-bound; no tensor cores.
//
// Every kernel restates one [OCV] routine (OpenCV 2.4.x objdetect/linemod.cpp + the imgproc primitives it calls);
// the exact arithmetic each must reproduce is SURVEY.md Appendix A.  Floating point is evaluated with explicit
// round-to-nearest intrinsics (no FMA contraction) so results are bit-identical to the SSE2 reference path.
#include <float.h>

#include "lm_kernels.cuh"

namespace lmk {

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
  return p;
}

// ---------------------------------------------------------------------------------------------- GaussianBlur 7x7
// [OCV] quantizedOrientations: GaussianBlur(src, smoothed, Size(7,7), 0, 0, BORDER_REPLICATE)  (A.2-1)
// 8-bit fixed point, taps {8,28,56,72,56,28,8}, single rounding (sum + 2^15) >> 16.
constexpr int G_TW = 64, G_TH = 16;

__global__ void __launch_bounds__(256) k_gauss7_u8c3(const uint8_t* __restrict__ src, int rows, int cols,
                                                     uint8_t* __restrict__ dst) {
  __shared__ uint8_t s_in[G_TH + 6][(G_TW + 6) * 3];
  __shared__ uint16_t s_h[G_TH + 6][G_TW * 3];
  const int x0 = blockIdx.x * G_TW, y0 = blockIdx.y * G_TH;
  const int tid = threadIdx.x;
  constexpr int IN_W = (G_TW + 6) * 3;
  for (int i = tid; i < (G_TH + 6) * IN_W; i += 256) {
    int r = i / IN_W, rem = i - r * IN_W;
    int cx = rem / 3, c = rem - cx * 3;
    int gy = clampi(y0 + r - 3, 0, rows - 1), gx = clampi(x0 + cx - 3, 0, cols - 1);
    s_in[r][rem] = src[((size_t)gy * cols + gx) * 3 + c];
  }
  __syncthreads();
  for (int i = tid; i < (G_TH + 6) * G_TW * 3; i += 256) {
    int r = i / (G_TW * 3), rem = i - r * (G_TW * 3);
    const uint8_t* p = &s_in[r][rem];
    int s = 8 * (p[0] + p[18]) + 28 * (p[3] + p[15]) + 56 * (p[6] + p[12]) + 72 * p[9];
    s_h[r][rem] = (uint16_t)s;
  }
  __syncthreads();
  for (int i = tid; i < G_TH * G_TW * 3; i += 256) {
    int r = i / (G_TW * 3), rem = i - r * (G_TW * 3);
    int gy = y0 + r, gx = x0 + rem / 3;
    if (gy < rows && gx < cols) {
      int s = 8 * ((int)s_h[r][rem] + s_h[r + 6][rem]) + 28 * ((int)s_h[r + 1][rem] + s_h[r + 5][rem]) +
              56 * ((int)s_h[r + 2][rem] + s_h[r + 4][rem]) + 72 * (int)s_h[r + 3][rem];
      dst[((size_t)gy * cols + x0) * 3 + rem] = (uint8_t)((s + 32768) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------- Sobel + phase
// cv::fastAtan2 (degrees) exactly as the SSE2 path evaluates it: every operation a separate f32 rounding (A.2-4).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
  const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
  const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
  const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
  float ax = fabsf(x), ay = fabsf(y);
  float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
  float c = __fdiv_rn(mn, __fadd_rn(mx, (float)DBL_EPSILON));
  float c2 = __fmul_rn(c, c);
  float a = __fmul_rn(p7, c2);
  a = __fmul_rn(__fadd_rn(a, p5), c2);
  a = __fmul_rn(__fadd_rn(a, p3), c2);
  a = __fmul_rn(__fadd_rn(a, p1), c);
  if (ax < ay) a = __fsub_rn(90.f, a);
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// [OCV] quantizedOrientations: Sobel dx/dy (16S, replicate) on the smoothed image, channel of maximum magnitude,
// phase in degrees; plus the first half of hysteresisGradient: 16-bin rounding (round-half-even), zeroed border ring,
// "& 7" in the interior.  Outputs: magnitude (f32, exact integer) and the unfiltered quantisation.
constexpr int S_TW = 64, S_TH = 16;

__global__ void __launch_bounds__(256) k_cg_grad(const uint8_t* __restrict__ sm, int rows, int cols,
                                                 float* __restrict__ mag, uint8_t* __restrict__ qunf) {
  __shared__ uint8_t s[S_TH + 2][(S_TW + 2) * 3];
  const int x0 = blockIdx.x * S_TW, y0 = blockIdx.y * S_TH;
  const int tid = threadIdx.x;
  constexpr int IN_W = (S_TW + 2) * 3;
  for (int i = tid; i < (S_TH + 2) * IN_W; i += 256) {
    int r = i / IN_W, rem = i - r * IN_W;
    int cx = rem / 3, c = rem - cx * 3;
    int gy = clampi(y0 + r - 1, 0, rows - 1), gx = clampi(x0 + cx - 1, 0, cols - 1);
    s[r][rem] = sm[((size_t)gy * cols + gx) * 3 + c];
  }
  __syncthreads();
  for (int i = tid; i < S_TH * S_TW; i += 256) {
    int r = i / S_TW, x = i - r * S_TW;
    int gy = y0 + r, gx = x0 + x;
    if (gy >= rows || gx >= cols) continue;
    int best_m = 0, best_dx = 0, best_dy = 0;
    int m[3], dxs[3], dys[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint8_t* a = &s[r][x * 3 + c];      // row above, column x-1
      const uint8_t* b = &s[r + 1][x * 3 + c];  // same row
      const uint8_t* d = &s[r + 2][x * 3 + c];  // row below
      int dx = (a[6] + 2 * b[6] + d[6]) - (a[0] + 2 * b[0] + d[0]);
      int dy = (d[0] + 2 * d[3] + d[6]) - (a[0] + 2 * a[3] + a[6]);
      dxs[c] = dx; dys[c] = dy; m[c] = dx * dx + dy * dy;
    }
    if (m[0] >= m[1] && m[0] >= m[2]) { best_m = m[0]; best_dx = dxs[0]; best_dy = dys[0]; }
    else if (m[1] >= m[0] && m[1] >= m[2]) { best_m = m[1]; best_dx = dxs[1]; best_dy = dys[1]; }
    else { best_m = m[2]; best_dx = dxs[2]; best_dy = dys[2]; }
    float angle = fast_atan2_deg((float)best_dy, (float)best_dx);
    int q = __float2int_rn(__fmul_rn(angle, (float)(16.0 / 360.0)));
    q = clampi(q, 0, 255);
    bool border = gy == 0 || gy == rows - 1 || gx == 0 || gx == cols - 1;
    q = border ? 0 : (q & 7);
    size_t o = (size_t)gy * cols + gx;
    mag[o] = (float)best_m;
    qunf[o] = (uint8_t)q;
  }
}

// [OCV] hysteresisGradient, second half: 3x3 vote over the unfiltered bins where magnitude > weak^2; a bin needs
// >= 5 of 9 votes; ties resolved towards the lowest bin (strict '<' scan).
__global__ void __launch_bounds__(256) k_cg_hysteresis(const uint8_t* __restrict__ qunf, const float* __restrict__ mag,
                                                       int rows, int cols, float thr, uint8_t* __restrict__ quant) {
  int gx = blockIdx.x * 64 + (threadIdx.x & 63), gy = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (gx >= cols || gy >= rows) return;
  size_t o = (size_t)gy * cols + gx;
  uint8_t out = 0;
  if (gy >= 1 && gy < rows - 1 && gx >= 1 && gx < cols - 1 && mag[o] > thr) {
    unsigned hist = 0;  // 8 x 4-bit counters
#pragma unroll
    for (int j = -1; j <= 1; ++j)
#pragma unroll
      for (int i = -1; i <= 1; ++i) hist += 1u << (4 * qunf[o + j * cols + i]);
    int max_votes = 0, index = -1;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      int v = (hist >> (4 * b)) & 15;
      if (max_votes < v) { index = b; max_votes = v; }
    }
    if (max_votes >= 5) out = (uint8_t)(1 << index);
  }
  quant[o] = out;
}

// ---------------------------------------------------------------------------------------------- pyrDown
// [OCV] ColorGradientPyramid::pyrDown -> cv::pyrDown: 5x5 binomial, (sum + 128) >> 8, BORDER_REFLECT_101 (A.2-7)
__global__ void __launch_bounds__(256) k_pyrdown_u8c3(const uint8_t* __restrict__ src, int rows, int cols,
                                                      uint8_t* __restrict__ dst) {
  const int orows = rows / 2, ocols = cols / 2;
  int e = blockIdx.x * 256 + threadIdx.x;  // element within an output row (x*3 + c)
  int y = blockIdx.y;
  if (e >= ocols * 3 || y >= orows) return;
  int x = e / 3, c = e - x * 3;
  const int k[5] = {1, 4, 6, 4, 1};
  int s = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint8_t* row = src + (size_t)reflect101(2 * y + j - 2, rows) * cols * 3 + c;
    int rs = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) rs += k[i] * row[reflect101(2 * x + i - 2, cols) * 3];
    s += k[j] * rs;
  }
  dst[((size_t)y * ocols) * 3 + e] = (uint8_t)((s + 128) >> 8);
}

// ---------------------------------------------------------------------------------------------- DepthNormal
// [OCV] quantizedNormals (before medianBlur): 8 taps at +-5 px, bilateral plane fit in int64, normalisation in f32
// (mul/add/sqrt/div each rounded separately), NORMAL_LUT[v3][v2][v1] with the flat-index / out-of-table -> 0 rule.
__global__ void __launch_bounds__(256) k_dn_normals(const uint16_t* __restrict__ depth, int rows, int cols,
                                                    int distance_threshold, int difference_threshold,
                                                    const uint8_t* __restrict__ lut, uint8_t* __restrict__ out) {
  int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= cols || y >= rows) return;
  uint8_t res = 0;
  const int r = 5;
  if (y >= r && y < rows - r - 1 && x >= r && x < cols - r - 1) {
    long long d = depth[(size_t)y * cols + x];
    if (d < distance_threshold) {
      long long A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = k < 4 ? k : k + 1;  // skip the centre of the 3x3 offset grid
        const int i = (kk % 3 - 1) * r, j = (kk / 3 - 1) * r;
        long long delta = (long long)depth[(size_t)(y + j) * cols + (x + i)] - d;
        long long f = (delta < 0 ? -delta : delta) < difference_threshold ? 1 : 0;
        long long fi = f * i, fj = f * j;
        A0 += fi * i; A1 += fi * j; A3 += fj * j;
        b0 += fi * delta; b1 += fj * delta;
      }
      long long det = A0 * A3 - A1 * A1;
      long long ddx = A3 * b0 - A1 * b1;
      long long ddy = -A1 * b0 + A0 * b1;
      float nx = __ll2float_rn(1150 * ddx);
      float ny = __ll2float_rn(1150 * ddy);
      float nz = __ll2float_rn(-det * d);
      float s2 = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
      float s = __fsqrt_rn(s2);
      if (s > 0) {
        float inv = __fdiv_rn(1.0f, s);
        nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
        int v1 = __float2int_rz(__fadd_rn(__fmul_rn(nx, 10.0f), 10.0f));
        int v2 = __float2int_rz(__fadd_rn(__fmul_rn(ny, 10.0f), 10.0f));
        int v3 = __float2int_rz(__fadd_rn(__fmul_rn(nz, 20.0f), 20.0f));
        int flat = (v3 * 20 + v2) * 20 + v1;
        res = (flat >= 0 && flat < 8000) ? lut[flat] : 0;
      }
    }
  }
  out[(size_t)y * cols + x] = res;
}

// [OCV] quantizedNormals tail: medianBlur(dst, dst, 5), replicate border.  Rank selection: the median of 25 is the
// element with exactly 12 elements ordered before it (ties broken by window index).
__global__ void __launch_bounds__(256) k_median5_u8(const uint8_t* __restrict__ src, int rows, int cols,
                                                    uint8_t* __restrict__ dst) {
  __shared__ uint8_t s[8 + 4][64 + 4];
  const int x0 = blockIdx.x * 64, y0 = blockIdx.y * 8;
  for (int i = threadIdx.x; i < 12 * 68; i += 256) {
    int r = i / 68, c = i - r * 68;
    s[r][c] = src[(size_t)clampi(y0 + r - 2, 0, rows - 1) * cols + clampi(x0 + c - 2, 0, cols - 1)];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 64; i += 256) {
    int r = i >> 6, c = i & 63;
    int gy = y0 + r, gx = x0 + c;
    if (gy >= rows || gx >= cols) continue;
    uint8_t w[25];
#pragma unroll
    for (int j = 0; j < 5; ++j)
#pragma unroll
      for (int k = 0; k < 5; ++k) w[j * 5 + k] = s[r + j][c + k];
    // fast path: the windows of a quantised-normal image are mostly constant
    uint8_t med = w[12];
    bool uniform = true;
#pragma unroll
    for (int k = 0; k < 25; ++k) uniform &= (w[k] == med);
    if (!uniform) {
#pragma unroll
      for (int a = 0; a < 25; ++a) {
        int before = 0;
#pragma unroll
        for (int b = 0; b < 25; ++b) before += (w[b] < w[a]) || (w[b] == w[a] && b < a);
        if (before == 12) med = w[a];
      }
    }
    dst[(size_t)gy * cols + gx] = med;
  }
}

// [OCV] DepthNormalPyramid::pyrDown: resize(normal, next, size/2, INTER_NN) == src(2y, 2x)
__global__ void __launch_bounds__(256) k_nn_half_u8(const uint8_t* __restrict__ src, int rows, int cols,
                                                    uint8_t* __restrict__ dst) {
  const int orows = rows / 2, ocols = cols / 2;
  int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= ocols || y >= orows) return;
  dst[(size_t)y * ocols + x] = src[(size_t)(2 * y) * cols + 2 * x];
}

// ---------------------------------------------------------------------------------------------- spread -> LM
// [OCV] quantize (mask) + spread + computeResponseMaps + linearize fused.  One block owns grid row `a` (image rows
// a*T .. a*T+T-1) and a run of CW grid cells; it ORs the T x T neighbourhood in shared memory (separable), maps
// every spread byte to its 8 responses with one table lookup and stores them straight into the T^2-strided linear
// memories, 4 consecutive cells per 32-bit store.
template <int CW>
__global__ void __launch_bounds__(256) k_spread_lm(const uint8_t* __restrict__ qraw, const uint8_t* __restrict__ mask0,
                                                   int mask_cols0, int level, int rows, int cols, int T, int W, int H,
                                                   const uint32_t* __restrict__ resp_all,
                                                   uint8_t* __restrict__ quantized_out, uint8_t* __restrict__ spread_out,
                                                   uint8_t* __restrict__ response_out, uint8_t* __restrict__ lm,
                                                   size_t plane_stride) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int TWp = CW * T;         // tile width in pixels
  const int IW = TWp + T - 1;     // with right halo
  const int IH = 2 * T - 1;       // with bottom halo
  uint32_t* s_resp = reinterpret_cast<uint32_t*>(smem);  // 256 entries
  uint8_t* sq = smem + 1024;                             // [IH][IW]   masked quantisation
  uint8_t* sh = sq + ((IH * IW + 15) & ~15);             // [IH][TWp]  horizontal OR
  uint8_t* sp = sh + ((IH * TWp + 15) & ~15);            // [T][TWp]   spread
  const int c0 = blockIdx.x * CW;  // first grid cell
  const int a = blockIdx.y;
  const int px0 = c0 * T, py0 = a * T;
  const int tid = threadIdx.x;
  s_resp[tid] = resp_all[tid];
  for (int i = tid; i < IH * IW; i += 256) {
    int r = i / IW, x = i - r * IW;
    int gy = py0 + r, gx = px0 + x;
    uint8_t v = 0;
    if (gy < rows && gx < cols) {
      v = qraw[(size_t)gy * cols + gx];
      if (mask0 && !mask0[(size_t)(gy << level) * mask_cols0 + (gx << level)]) v = 0;
      if (r < T && x < TWp) quantized_out[(size_t)gy * cols + gx] = v;
    }
    sq[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < IH * TWp; i += 256) {
    int r = i / TWp, x = i - r * TWp;
    const uint8_t* p = sq + r * IW + x;
    uint8_t v = 0;
    for (int c = 0; c < T; ++c) v |= p[c];
    sh[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < T * TWp; i += 256) {
    int r = i / TWp, x = i - r * TWp;
    uint8_t v = 0;
    for (int k = 0; k < T; ++k) v |= sh[(r + k) * TWp + x];
    sp[i] = v;
    int gy = py0 + r, gx = px0 + x;
    if (spread_out && gy < rows && gx < cols) spread_out[(size_t)gy * cols + gx] = v;
  }
  __syncthreads();
  const size_t WH = (size_t)W * H;
  const int ncell = min(CW, W - c0);
  if ((W & 3) == 0) {
    const int items = T * T * (CW / 4);
    for (int it = tid; it < items; it += 256) {
      int b4 = it % (CW / 4), g = it / (CW / 4);  // g = rs*T + cs
      if (b4 * 4 >= ncell) continue;
      int rs = g / T, cs = g - rs * T;
      const uint8_t* p = sp + rs * TWp + cs + T * (b4 * 4);
      uint32_t r0 = s_resp[p[0]], r1 = s_resp[p[T]], r2 = s_resp[p[2 * T]], r3 = s_resp[p[3 * T]];
      size_t o = (size_t)g * WH + (size_t)a * W + c0 + b4 * 4;
#pragma unroll
      for (int ori = 0; ori < 8; ++ori) {
        uint32_t v = ((r0 >> (4 * ori)) & 15) | (((r1 >> (4 * ori)) & 15) << 8) | (((r2 >> (4 * ori)) & 15) << 16) |
                     (((r3 >> (4 * ori)) & 15) << 24);
        *reinterpret_cast<uint32_t*>(lm + ori * plane_stride + o) = v;
      }
    }
  } else {
    const int items = T * T * CW;
    for (int it = tid; it < items; it += 256) {
      int b = it % CW, g = it / CW;
      if (b >= ncell) continue;
      int rs = g / T, cs = g - rs * T;
      uint32_t r0 = s_resp[sp[rs * TWp + cs + T * b]];
      size_t o = (size_t)g * WH + (size_t)a * W + c0 + b;
#pragma unroll
      for (int ori = 0; ori < 8; ++ori) lm[ori * plane_stride + o] = (uint8_t)((r0 >> (4 * ori)) & 15);
    }
  }
  if (response_out) {
    for (int i = tid; i < T * TWp; i += 256) {
      int r = i / TWp, x = i - r * TWp;
      int gy = py0 + r, gx = px0 + x;
      if (gy < rows && gx < cols) {
        uint32_t r0 = s_resp[sp[i]];
#pragma unroll
        for (int ori = 0; ori < 8; ++ori)
          response_out[(size_t)ori * rows * cols + (size_t)gy * cols + gx] = (uint8_t)((r0 >> (4 * ori)) & 15);
      }
    }
  }
}

template <int CW>
size_t spread_smem_bytes(int T) {
  int TWp = CW * T, IW = TWp + T - 1, IH = 2 * T - 1;
  return 1024 + ((IH * IW + 15) & ~15) + ((IH * TWp + 15) & ~15) + (size_t)T * TWp;
}

}  // namespace

// ================================================================================================ launchers
void launch_gauss7_u8c3(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s) {
  dim3 grid((cols + G_TW - 1) / G_TW, (rows + G_TH - 1) / G_TH);
  k_gauss7_u8c3<<<grid, 256, 0, s>>>(src, rows, cols, dst);
}
void launch_cg_grad(const uint8_t* smoothed, int rows, int cols, float* mag, uint8_t* qunf, cudaStream_t s) {
  dim3 grid((cols + S_TW - 1) / S_TW, (rows + S_TH - 1) / S_TH);
  k_cg_grad<<<grid, 256, 0, s>>>(smoothed, rows, cols, mag, qunf);
}
void launch_cg_hysteresis(const uint8_t* qunf, const float* mag, int rows, int cols, float threshold_sq, uint8_t* quant,
                          cudaStream_t s) {
  dim3 grid((cols + 63) / 64, (rows + 3) / 4);
  k_cg_hysteresis<<<grid, 256, 0, s>>>(qunf, mag, rows, cols, threshold_sq, quant);
}
void launch_pyrdown_u8c3(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s) {
  dim3 grid(((cols / 2) * 3 + 255) / 256, rows / 2);
  k_pyrdown_u8c3<<<grid, 256, 0, s>>>(src, rows, cols, dst);
}
void launch_dn_normals(const uint16_t* depth, int rows, int cols, int distance_threshold, int difference_threshold,
                       const uint8_t* normal_lut, uint8_t* out, cudaStream_t s) {
  dim3 grid((cols + 63) / 64, (rows + 3) / 4);
  k_dn_normals<<<grid, 256, 0, s>>>(depth, rows, cols, distance_threshold, difference_threshold, normal_lut, out);
}
void launch_median5_u8(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s) {
  dim3 grid((cols + 63) / 64, (rows + 7) / 8);
  k_median5_u8<<<grid, 256, 0, s>>>(src, rows, cols, dst);
}
void launch_nn_half_u8(const uint8_t* src, int rows, int cols, uint8_t* dst, cudaStream_t s) {
  dim3 grid((cols / 2 + 255) / 256, rows / 2);
  k_nn_half_u8<<<grid, 256, 0, s>>>(src, rows, cols, dst);
}

void launch_spread_lm(const uint8_t* quant_raw, const uint8_t* mask0, int mask_cols0, int level, int rows, int cols,
                      int T, const uint32_t* resp_all, uint8_t* quantized_out, uint8_t* spread_out,
                      uint8_t* response_out, uint8_t* lm, size_t plane_stride, cudaStream_t s) {
  const int W = cols / T, H = rows / T;
  const size_t limit = 48 * 1024;
#define LM_SPREAD_CASE(CWV)                                                                                       \
  if (spread_smem_bytes<CWV>(T) <= limit) {                                                                       \
    dim3 grid((W + CWV - 1) / CWV, H);                                                                            \
    k_spread_lm<CWV><<<grid, 256, spread_smem_bytes<CWV>(T), s>>>(quant_raw, mask0, mask_cols0, level, rows, cols, \
                                                                  T, W, H, resp_all, quantized_out, spread_out,   \
                                                                  response_out, lm, plane_stride);                \
    return;                                                                                                       \
  }
  LM_SPREAD_CASE(64)
  LM_SPREAD_CASE(32)
  LM_SPREAD_CASE(16)
  LM_SPREAD_CASE(8)
  LM_SPREAD_CASE(4)
#undef LM_SPREAD_CASE
}

}  // namespace lmk
