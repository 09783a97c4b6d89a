// lm_frontend_fused.cu -- the front end: four launches per CHUNK of frames (pyrDown, ColorGradient, DepthNormal, spread).
//
//   k_cg_fused    [OCV] quantizedOrientations + hysteresisGradient for every pyramid level in one grid:
//                 GaussianBlur 7x7 -> Sobel 3x3 -> max-magnitude channel -> fastAtan2 -> 16-bin rounding -> 3x3 vote,
//                 all staged through shared memory (the per-level source of levels >= 1 comes from k_pyrdown_u8c3).
//   k_dn_fused    [OCV] quantizedNormals incl. medianBlur(5) and DepthNormalPyramid::pyrDown: plane fit + LUT, a
//                 99-exchange median-of-25 network on packed u16x2 (VIMNMX.U16x2), NN decimation to all levels.
//   k_spread_all  [OCV] quantize(mask) + spread + computeResponseMaps + linearize for every (level, modality).
//
// The front end is O(pixels) and tiny next to a B200 (7.7 MB of algorithmic traffic per 640x480 frame): one frame's
// grids (150 - 900 CTAs) are less than one wave, so a launch costs the latency of its dependent shared-memory phases
// whatever its size.  Hence (i) the stages are fused (no global round trips between them) and (ii) every kernel takes a
// chunk of frames -- blockIdx.y / .z is the frame, buffers at base + frame * stride, level-0 sources from the device
// frame table -- so that a launch fills the machine for several waves.
#include <float.h>

#include <type_traits>

#include "lm_kernels.cuh"
#include "lm_median_net.h"

namespace lmk {

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// cv::fastAtan2 in degrees, every operation a separate f32 rounding.
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

// ---------------------------------------------------------------------------------------------- ColorGradient
// One CTA (320 threads) = a 64 x 16 pixel tile of one pyramid level, halo 5 (blur 3 + Sobel 1 + vote 1).  All stages
// work on packed data:
//   A  source tile -> smem as 32-bit words (replicate-extended at the image border)
//   B  vertical 7-tap blur on byte pairs (u16x2 lanes: 255 * 256 fits), sliding window down a word column
//   C  horizontal 7-tap blur on the u16 sums, (sum + 2^15) >> 16, written planar (one byte plane per channel)
//   C' BORDER_REPLICATE of the *smoothed* image for Sobel: out-of-image entries take the value at the clamped coordinate
//   D  Sobel 3x3 for four pixels at a time on u16x2 lanes, max-magnitude channel, fastAtan2, 16-bin rounding;
//      every pixel leaves a one-hot vote nibble 1 << 4q
//   E  3x3 vote = sum of nine nibble words; at most one bin can reach 5 of 9 votes, found with one add + mask + ffs
constexpr int C_TW = 64, C_TH = 16, C_THREADS = 320;
constexpr int C_SRC_ROWS = C_TH + 10, C_SRC_WORDS = 57;   // 57 words = 228 bytes >= 1 + 74 * 3
constexpr int C_V_ROWS = C_TH + 4, C_V_COLS = C_SRC_WORDS * 4;
constexpr int C_SM_W = 72;                                // smoothed plane row: 68 pixels + pad, 18 words
constexpr int C_OH_ROWS = C_TH + 2, C_OH_W = 68;

__global__ void __launch_bounds__(C_THREADS) k_cg_fused(const CgParams P) {
  __shared__ __align__(16) uint32_t s_src[C_SRC_ROWS][C_SRC_WORDS];   // source bytes; pixel x0-5+i, channel c at byte 1+3i+c
  __shared__ __align__(16) uint16_t s_v[C_V_ROWS][C_V_COLS];         // vertical blur sums, same byte indexing
  __shared__ __align__(16) uint8_t s_sm[3][C_V_ROWS][C_SM_W];         // smoothed, planar; entry ix <-> x0-2+ix
  // s_oh / s_mag live in the space of s_src / s_v, which are dead once stage C is done
  uint32_t (*s_oh)[C_OH_W] = reinterpret_cast<uint32_t (*)[C_OH_W]>(&s_v[0][0]);       // [18][68] vote nibbles, jx <-> x0-1+jx
  float (*s_mag)[C_TW] = reinterpret_cast<float (*)[C_TW]>(&s_src[0][0]);              // [16][64]
  static_assert(sizeof(uint32_t) * C_OH_ROWS * C_OH_W <= sizeof(s_v), "s_oh must fit in s_v");
  static_assert(sizeof(float) * C_TH * C_TW <= sizeof(s_src), "s_mag must fit in s_src");

  const int frame = blockIdx.y;
  if (frame >= P.ctl->ft.n_frames) return;
  int lvl = 0;
  while (lvl + 1 < P.n_levels && (int)blockIdx.x >= P.lv[lvl + 1].block_begin) ++lvl;
  const uint8_t* __restrict__ src = lvl == 0 ? static_cast<const uint8_t*>(P.ctl->ft.src[frame][P.modality])
                                             : P.lv[lvl].src + (size_t)frame * P.lv[lvl].src_stride;
  float* __restrict__ mag_out = P.lv[lvl].mag + (size_t)frame * P.lv[lvl].mag_stride;
  uint8_t* __restrict__ quant_out = P.lv[lvl].quant + (size_t)frame * P.lv[lvl].quant_stride;
  const int rows = P.lv[lvl].rows, cols = P.lv[lvl].cols;
  const int b = blockIdx.x - P.lv[lvl].block_begin;
  const int x0 = (b % P.lv[lvl].blocks_x) * C_TW, y0 = (b / P.lv[lvl].blocks_x) * C_TH;
  const int tid = threadIdx.x;

  // ---- A: source tile
  if ((cols & 3) == 0 && x0 >= C_TW && x0 + 71 <= cols && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    const uint8_t* base = src + (size_t)3 * x0 - 16;  // 4-byte aligned: cols % 4 == 0 and x0 % 64 == 0
    // item i = r * C_SRC_WORDS + w, i += C_THREADS: (r, w) advanced without a division per item
    constexpr int kDr = C_THREADS / C_SRC_WORDS, kDw = C_THREADS % C_SRC_WORDS;
    for (int r = tid / C_SRC_WORDS, w = tid % C_SRC_WORDS; r < C_SRC_ROWS;) {
      const int gy = clampi(y0 - 5 + r, 0, rows - 1);
      s_src[r][w] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)gy * cols * 3) + w);
      r += kDr; w += kDw;
      if (w >= C_SRC_WORDS) { w -= C_SRC_WORDS; ++r; }
    }
  } else {
    uint8_t* sb = reinterpret_cast<uint8_t*>(&s_src[0][0]);
    for (int i = tid; i < C_SRC_ROWS * (C_TW + 10); i += C_THREADS) {
      const int r = i / (C_TW + 10), px = i - r * (C_TW + 10);
      const int gy = clampi(y0 - 5 + r, 0, rows - 1), gx = clampi(x0 - 5 + px, 0, cols - 1);
      const uint8_t* g = src + ((size_t)gy * cols + gx) * 3;
      uint8_t* d = sb + r * (C_SRC_WORDS * 4) + 1 + 3 * px;
      d[0] = g[0]; d[1] = g[1]; d[2] = g[2];
    }
  }
  __syncthreads();

  // ---- B: vertical blur, thread = (word column, group of 5 output rows); output row o uses source rows o .. o+6
  if (tid < C_SRC_WORDS * 4) {
    const int w = tid % C_SRC_WORDS, g = tid / C_SRC_WORDS;
    uint32_t e[11], o[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const uint32_t v = s_src[5 * g + k][w];
      e[k] = v & 0x00ff00ffu;
      o[k] = (v >> 8) & 0x00ff00ffu;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const uint32_t E = 8u * (e[k] + e[k + 6]) + 28u * (e[k + 1] + e[k + 5]) + 56u * (e[k + 2] + e[k + 4]) + 72u * e[k + 3];
      const uint32_t O = 8u * (o[k] + o[k + 6]) + 28u * (o[k + 1] + o[k + 5]) + 56u * (o[k + 2] + o[k + 4]) + 72u * o[k + 3];
      uint2 st;
      st.x = __byte_perm(E, O, 0x5410);  // bytes 4w, 4w+1
      st.y = __byte_perm(E, O, 0x7632);  // bytes 4w+2, 4w+3
      *reinterpret_cast<uint2*>(&s_v[5 * g + k][4 * w]) = st;
    }
  }
  __syncthreads();

  // ---- C: horizontal blur, item = (row, channel, run of 4 pixels); smoothed ix uses source pixels ix .. ix+6
  // item it = (r * 3 + c) * 17 + run, it += C_THREADS = 18 * 17 + 14: (r, c, run) advanced without divisions per item
  for (int run = tid % 17, c = (tid / 17) % 3, r = tid / 51; r < C_V_ROWS;) {
    const uint16_t* v = &s_v[r][1 + 12 * run + c];
    uint32_t t[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) t[k] = v[3 * k];
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t sum = 8u * (t[k] + t[k + 6]) + 28u * (t[k + 1] + t[k + 5]) + 56u * (t[k + 2] + t[k + 4]) + 72u * t[k + 3];
      out |= ((sum + 32768u) >> 16) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(&s_sm[c][r][4 * run]) = out;
    static_assert(C_THREADS == 18 * 17 + 14, "stage C's index update assumes 320 threads");
    run += 14; r += 6;                       // + 18 (row, channel) pairs = + 6 rows ...
    if (run >= 17) { run -= 17; ++c; }       // ... + 1 pair when the run index wraps
    if (c >= 3) { c -= 3; ++r; }
  }
  __syncthreads();

  // ---- C': Sobel's BORDER_REPLICATE acts on the smoothed image (only tiles that touch the image border)
  if (y0 < 2 || y0 + C_TH + 2 > rows || x0 < 2 || x0 + C_TW + 2 > cols) {
    // only the entries outside the image are touched: columns left of x = 0 / right of x = cols - 1 (all 20 rows, corners
    // included), then rows above y = 0 / below y = rows - 1 (in-image columns).  Every source entry is in-image: never rewritten.
    const int n_left = max(0, 2 - x0), n_right = max(0, min(68, x0 - 2 + 68 - cols));
    const int n_top = max(0, 2 - y0), n_bot = max(0, min(C_V_ROWS, y0 - 2 + C_V_ROWS - rows));
    const int nc = n_left + n_right, nr = n_top + n_bot;
    for (int i = tid; i < 3 * C_V_ROWS * nc; i += C_THREADS) {
      const int j = i % nc, rc = i / nc;
      const int r = rc % C_V_ROWS, c = rc / C_V_ROWS;
      const int ix = j < n_left ? j : 68 - n_right + (j - n_left);
      const int cy = clampi(y0 - 2 + r, 0, rows - 1), cx = clampi(x0 - 2 + ix, 0, cols - 1);
      s_sm[c][r][ix] = s_sm[c][cy - (y0 - 2)][cx - (x0 - 2)];
    }
    const int w_in = 68 - nc;   // in-image columns: ix in [n_left, 68 - n_right)
    for (int i = tid; i < 3 * nr * w_in; i += C_THREADS) {
      const int k = i % w_in, rc = i / w_in;
      const int j = rc % nr, c = rc / nr;
      const int r = j < n_top ? j : C_V_ROWS - n_bot + (j - n_top);
      const int ix = n_left + k;
      const int cy = clampi(y0 - 2 + r, 0, rows - 1);
      s_sm[c][r][ix] = s_sm[c][cy - (y0 - 2)][ix];
    }
    __syncthreads();
  }

  // ---- D: Sobel + orientation, item = (row jr <-> y0-1+jr, quad q <-> pixels jx = 4q .. 4q+3 <-> x0-1+jx)
  float mag_keep[4];
  int keep_r = -1, keep_q = 0;
  if (tid < C_OH_ROWS * 17) {
    const int q = tid % 17, jr = tid / 17;
    int dxs[3][4], dys[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t* ra = reinterpret_cast<const uint32_t*>(&s_sm[c][jr][0]) + q;
      const uint32_t* rb = reinterpret_cast<const uint32_t*>(&s_sm[c][jr + 1][0]) + q;
      const uint32_t* rc = reinterpret_cast<const uint32_t*>(&s_sm[c][jr + 2][0]) + q;
      const uint32_t a0 = ra[0], a1 = ra[1], b0 = rb[0], b1 = rb[1], c0 = rc[0], c1 = rc[1];
      const uint32_t M = 0x00ff00ffu;
      const uint32_t ae0 = a0 & M, ao0 = (a0 >> 8) & M, ae1 = a1 & M, ao1 = (a1 >> 8) & M;
      const uint32_t be0 = b0 & M, bo0 = (b0 >> 8) & M, be1 = b1 & M, bo1 = (b1 >> 8) & M;
      const uint32_t ce0 = c0 & M, co0 = (c0 >> 8) & M, ce1 = c1 & M, co1 = (c1 >> 8) & M;
      // vertical (1,2,1): s[j] for byte columns 0..5 as u16x2: (s0,s2) (s1,s3) (s4,-) (s5,-)
      const uint32_t se0 = ae0 + 2u * be0 + ce0, so0 = ao0 + 2u * bo0 + co0;
      const uint32_t se1 = ae1 + 2u * be1 + ce1, so1 = ao1 + 2u * bo1 + co1;
      dxs[c][0] = (int)(se0 >> 16) - (int)(se0 & 0xffffu);
      dxs[c][1] = (int)(so0 >> 16) - (int)(so0 & 0xffffu);
      dxs[c][2] = (int)(se1 & 0xffffu) - (int)(se0 >> 16);
      dxs[c][3] = (int)(so1 & 0xffffu) - (int)(so0 >> 16);
      // horizontal (1,2,1) of the rows above and below: h[j] = r[j] + 2 r[j+1] + r[j+2]
      const uint32_t ha_e = ae0 + 2u * ao0 + __byte_perm(ae0, ae1, 0x5432);  // (h0, h2)
      const uint32_t ha_o = ao0 + 2u * __byte_perm(ae0, ae1, 0x5432) + __byte_perm(ao0, ao1, 0x5432);  // (h1, h3)
      const uint32_t hc_e = ce0 + 2u * co0 + __byte_perm(ce0, ce1, 0x5432);
      const uint32_t hc_o = co0 + 2u * __byte_perm(ce0, ce1, 0x5432) + __byte_perm(co0, co1, 0x5432);
      dys[c][0] = (int)(hc_e & 0xffffu) - (int)(ha_e & 0xffffu);
      dys[c][1] = (int)(hc_o & 0xffffu) - (int)(ha_o & 0xffffu);
      dys[c][2] = (int)(hc_e >> 16) - (int)(ha_e >> 16);
      dys[c][3] = (int)(hc_o >> 16) - (int)(ha_o >> 16);
    }
    const int gy = y0 - 1 + jr;
    uint32_t oh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gx = x0 - 1 + 4 * q + k;
      const int m0 = dxs[0][k] * dxs[0][k] + dys[0][k] * dys[0][k];
      const int m1 = dxs[1][k] * dxs[1][k] + dys[1][k] * dys[1][k];
      const int m2 = dxs[2][k] * dxs[2][k] + dys[2][k] * dys[2][k];
      int bm, bdx, bdy;
      if (m0 >= m1 && m0 >= m2) { bm = m0; bdx = dxs[0][k]; bdy = dys[0][k]; }
      else if (m1 >= m0 && m1 >= m2) { bm = m1; bdx = dxs[1][k]; bdy = dys[1][k]; }
      else { bm = m2; bdx = dxs[2][k]; bdy = dys[2][k]; }
      const float angle = fast_atan2_deg((float)bdy, (float)bdx);
      const int qq = clampi(__float2int_rn(__fmul_rn(angle, (float)(16.0 / 360.0))), 0, 255) & 7;
      const bool border = gy <= 0 || gy >= rows - 1 || gx <= 0 || gx >= cols - 1;  // [OCV] first / last row and column are zeroed
      oh[k] = border ? 1u : (1u << (4 * qq));
      mag_keep[k] = (float)bm;
    }
    *reinterpret_cast<uint4*>(&s_oh[jr][4 * q]) = make_uint4(oh[0], oh[1], oh[2], oh[3]);  // s_v is dead: every thread passed stage C
    keep_r = jr; keep_q = q;
  }
  // magnitudes of the tile's own pixels: to global memory (addTemplate reads them) and to s_mag (aliases s_src, dead)
  if (keep_r >= 1 && keep_r <= C_TH) {
    const int gy = y0 + keep_r - 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int jx = 4 * keep_q + k;
      const int gx = x0 - 1 + jx;
      if (jx >= 1 && jx <= C_TW) {
        s_mag[keep_r - 1][jx - 1] = mag_keep[k];
        if (gy < rows && gx < cols) mag_out[(size_t)gy * cols + gx] = mag_keep[k];
      }
    }
  }
  __syncthreads();

  // ---- E: hysteresis vote, thread = four consecutive pixels of a row
  if (tid < C_TH * (C_TW / 4)) {
    const int eq = tid % (C_TW / 4), er = tid / (C_TW / 4);
    const int gy = y0 + er, gx0 = x0 + 4 * eq;
    uint32_t cs[6];
    {
      const uint4 r0 = *reinterpret_cast<const uint4*>(&s_oh[er][4 * eq]);
      const uint4 r1 = *reinterpret_cast<const uint4*>(&s_oh[er + 1][4 * eq]);
      const uint4 r2 = *reinterpret_cast<const uint4*>(&s_oh[er + 2][4 * eq]);
      const uint2 t0 = *reinterpret_cast<const uint2*>(&s_oh[er][4 * eq + 4]);
      const uint2 t1 = *reinterpret_cast<const uint2*>(&s_oh[er + 1][4 * eq + 4]);
      const uint2 t2 = *reinterpret_cast<const uint2*>(&s_oh[er + 2][4 * eq + 4]);
      cs[0] = r0.x + r1.x + r2.x; cs[1] = r0.y + r1.y + r2.y; cs[2] = r0.z + r1.z + r2.z; cs[3] = r0.w + r1.w + r2.w;
      cs[4] = t0.x + t1.x + t2.x; cs[5] = t0.y + t1.y + t2.y;
    }
    const float4 mg = *reinterpret_cast<const float4*>(&s_mag[er][4 * eq]);
    const float mags[4] = {mg.x, mg.y, mg.z, mg.w};
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gx = gx0 + k;
      const uint32_t hist = cs[k] + cs[k + 1] + cs[k + 2];             // nine 4-bit vote counters
      const uint32_t five = (hist + 0x33333333u) & 0x88888888u;        // counters >= 5: at most one (9 votes in all)
      const bool inner = gy >= 1 && gy < rows - 1 && gx >= 1 && gx < cols - 1;
      if (inner && mags[k] > P.thr_sq && five != 0) packed |= (1u << ((__ffs((int)five) - 1) >> 2)) << (8 * k);
    }
    uint8_t* qrow = quant_out + (size_t)gy * cols;
    if (gy < rows) {
      if ((cols & 3) == 0 && gx0 + 3 < cols) *reinterpret_cast<uint32_t*>(qrow + gx0) = packed;  // quant strides are multiples of 4
      else
        for (int k = 0; k < 4; ++k)
          if (gx0 + k < cols) qrow[gx0 + k] = (uint8_t)(packed >> (8 * k));
    }
  }
}

// ---------------------------------------------------------------------------------------------- pyrDown (BGR)
// [OCV] ColorGradientPyramid::pyrDown -> cv::pyrDown: 5x5 (1 4 6 4 1)^2, reflect-101, even samples, (sum + 128) >> 8.
// 64 x 8 output pixels per CTA: source region to smem as words (reflected at the image border), vertical taps on
// byte pairs (u16x2 lanes, 16 * 255 fits), horizontal taps on the u16 sums.
constexpr int Y_TW = 64, Y_TH = 8;
constexpr int Y_SRC_ROWS = 2 * Y_TH + 3, Y_SRC_WORDS = 99;  // 99 words = 396 bytes >= 2 + 131 * 3

__device__ __forceinline__ int reflect101_f(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * len - 2 - p;
  return p;
}

__global__ void __launch_bounds__(256) k_pyrdown_fast(const BatchCtl* __restrict__ ctl, int modality,
                                                      const uint8_t* __restrict__ src0, size_t src_stride, int rows, int cols,
                                                      uint8_t* __restrict__ dst0, size_t dst_stride) {
  __shared__ __align__(16) uint32_t s_src[Y_SRC_ROWS][Y_SRC_WORDS];   // region pixel i <-> source x 2*x0-2+i, channel c at byte 2+3i+c
  __shared__ __align__(16) uint16_t s_v[Y_TH][Y_SRC_WORDS * 4];
  const int frame = blockIdx.z;
  if (frame >= ctl->ft.n_frames) return;
  const uint8_t* __restrict__ src = src0 ? src0 + (size_t)frame * src_stride : static_cast<const uint8_t*>(ctl->ft.src[frame][modality]);
  uint8_t* __restrict__ dst = dst0 + (size_t)frame * dst_stride;
  const int orows = rows / 2, ocols = cols / 2;
  const int x0 = blockIdx.x * Y_TW, y0 = blockIdx.y * Y_TH;
  const int tid = threadIdx.x;
  if ((cols & 3) == 0 && x0 >= Y_TW && 2 * x0 + 130 <= cols && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
    const uint8_t* base = src + (size_t)6 * x0 - 8;  // 4-byte aligned
    for (int i = tid; i < Y_SRC_ROWS * Y_SRC_WORDS; i += 256) {
      const int r = i / Y_SRC_WORDS, w = i - r * Y_SRC_WORDS;
      const int gy = reflect101_f(2 * y0 - 2 + r, rows);
      s_src[r][w] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)gy * cols * 3) + w);
    }
  } else {
    uint8_t* sb = reinterpret_cast<uint8_t*>(&s_src[0][0]);
    for (int i = tid; i < Y_SRC_ROWS * 131; i += 256) {
      const int r = i / 131, px = i - r * 131;
      const int gy = reflect101_f(2 * y0 - 2 + r, rows), gx = reflect101_f(2 * x0 - 2 + px, cols);
      const uint8_t* g = src + ((size_t)gy * cols + gx) * 3;
      uint8_t* d = sb + r * (Y_SRC_WORDS * 4) + 2 + 3 * px;
      d[0] = g[0]; d[1] = g[1]; d[2] = g[2];
    }
  }
  __syncthreads();
  // vertical taps: output row oy uses region rows 2*oy .. 2*oy+4; thread = (word column, group of 4 output rows)
  if (tid < Y_SRC_WORDS * 2) {
    const int w = tid % Y_SRC_WORDS, g = tid / Y_SRC_WORDS;
    uint32_t e[11], o[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const uint32_t v = s_src[8 * g + k][w];
      e[k] = v & 0x00ff00ffu;
      o[k] = (v >> 8) & 0x00ff00ffu;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t E = (e[2 * k] + e[2 * k + 4]) + 4u * (e[2 * k + 1] + e[2 * k + 3]) + 6u * e[2 * k + 2];
      const uint32_t O = (o[2 * k] + o[2 * k + 4]) + 4u * (o[2 * k + 1] + o[2 * k + 3]) + 6u * o[2 * k + 2];
      uint2 st;
      st.x = __byte_perm(E, O, 0x5410);
      st.y = __byte_perm(E, O, 0x7632);
      *reinterpret_cast<uint2*>(&s_v[4 * g + k][4 * w]) = st;
    }
  }
  __syncthreads();
  // horizontal taps: output ox uses region pixels 2*ox .. 2*ox+4; item = (row, channel, run of 4 output pixels)
  for (int it = tid; it < Y_TH * 3 * 16; it += 256) {
    const int run = it & 15, rc = it >> 4;
    const int c = rc % 3, oy = rc / 3;
    const int gy = y0 + oy;
    if (gy >= orows) continue;
    const uint16_t* v = &s_v[oy][2 + 24 * run + c];
    uint32_t t[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) t[k] = v[3 * k];
    uint8_t* out = dst + ((size_t)gy * ocols + x0 + 4 * run) * 3 + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t sum = (t[2 * k] + t[2 * k + 4]) + 4u * (t[2 * k + 1] + t[2 * k + 3]) + 6u * t[2 * k + 2];
      if (x0 + 4 * run + k < ocols) out[3 * k] = (uint8_t)((sum + 128u) >> 8);
    }
  }
}

// ---------------------------------------------------------------------------------------------- DepthNormal
// [OCV] quantizedNormals accumulates the 2x2 normal equations in `long`.  With difference_threshold <= 200 (default
// 50) every intermediate fits 32 bits -- |b| <= 30 * 199, A <= 150, |1150 * ddx| <= 1150 * 250 * 5970 < 2^31,
// det * d <= 22500 * 65535 < 2^31 -- so FAST evaluates the same integers in int32 (identical values, identical
// int -> float conversions); larger thresholds take the 64-bit path.
template <bool FAST>
__device__ __forceinline__ uint8_t dn_normal_at(const uint16_t* __restrict__ depth, int rows, int cols, int y, int x,
                                                int distance_threshold, int difference_threshold,
                                                const uint8_t* __restrict__ lut) {
  typedef typename std::conditional<FAST, int, long long>::type acc_t;
  const int r = 5;
  if (!(y >= r && y < rows - r - 1 && x >= r && x < cols - r - 1)) return 0;
  const uint16_t* pc = depth + (size_t)y * cols + x;
  const int d = pc[0];
  if (!(d < distance_threshold)) return 0;
  acc_t A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int kk = k < 4 ? k : k + 1;
    const int i = (kk % 3 - 1) * r, j = (kk / 3 - 1) * r;
    const int delta = (int)pc[j * cols + i] - d;
    if (abs(delta) < difference_threshold) {
      A0 += i * i; A1 += i * j; A3 += j * j;
      b0 += (acc_t)(i * delta); b1 += (acc_t)(j * delta);
    }
  }
  const acc_t det = A0 * A3 - A1 * A1;
  const acc_t ddx = A3 * b0 - A1 * b1;
  const acc_t ddy = -A1 * b0 + A0 * b1;
  float nx, ny, nz;
  if (FAST) {
    nx = __int2float_rn((int)(1150 * ddx));
    ny = __int2float_rn((int)(1150 * ddy));
    nz = __int2float_rn((int)(-det * d));
  } else {
    nx = __ll2float_rn((long long)(1150 * ddx));
    ny = __ll2float_rn((long long)(1150 * ddy));
    nz = __ll2float_rn((long long)(-det * d));
  }
  const float s = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  if (!(s > 0)) return 0;
  const float inv = __fdiv_rn(1.0f, s);
  nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
  const int v1 = __float2int_rz(__fadd_rn(__fmul_rn(nx, 10.0f), 10.0f));
  const int v2 = __float2int_rz(__fadd_rn(__fmul_rn(ny, 10.0f), 10.0f));
  const int v3 = __float2int_rz(__fadd_rn(__fmul_rn(nz, 20.0f), 20.0f));
  const int flat = (v3 * 20 + v2) * 20 + v1;
  return (flat >= 0 && flat < 8000) ? lut[flat] : (uint8_t)0;
}

constexpr int D_TW = 64, D_TH = 16;

__device__ __forceinline__ void cswap_u16x2(uint32_t& a, uint32_t& b) {
  uint32_t lo, hi;
  asm("min.u16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
  asm("max.u16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
  a = lo; b = hi;
}

constexpr int D_RW = D_TW + 4;                                           // tile columns incl. halo 2
constexpr int D_COUNT_WORDS = 2 * (D_TH + 4) * D_RW + 2 * D_TH * (D_RW + 4);   // staging of the counting median (20 KB)

template <bool FAST>
__device__ __forceinline__ void dn_tile(const DnParams& P, int bx, int by, int frame, uint32_t* smem) {
  uint8_t (*s_raw)[D_TW + 8] = reinterpret_cast<uint8_t (*)[D_TW + 8]>(smem);   // [D_TH + 4][D_TW + 8]
  const int rows = P.rows, cols = P.cols;
  const uint16_t* __restrict__ depth = static_cast<const uint16_t*>(P.ctl->ft.src[frame][P.modality]);
  const int x0 = bx * D_TW, y0 = by * D_TH;
  const int tid = threadIdx.x;
  // raw quantised normals for the tile + halo 2; medianBlur's BORDER_REPLICATE = value at the clamped coordinate
  for (int i = tid; i < (D_TH + 4) * (D_TW + 4); i += 256) {
    const int r = i / (D_TW + 4), c = i - r * (D_TW + 4);
    const int gy = clampi(y0 - 2 + r, 0, rows - 1), gx = clampi(x0 - 2 + c, 0, cols - 1);
    s_raw[r][c] = dn_normal_at<FAST>(depth, rows, cols, gy, gx, P.distance_threshold, P.difference_threshold, P.lut);
  }
  __syncthreads();
  // median of 25 for two horizontally adjacent pixels at once (one per 16-bit lane)
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int r = (tid >> 5) + 8 * half, c = (tid & 31) * 2;
    uint32_t p[25];
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      // six consecutive bytes starting at an even column: three aligned 16-bit loads
      const uint16_t* row = reinterpret_cast<const uint16_t*>(&s_raw[r + dy][c]);
      const uint32_t h0 = row[0], h1 = row[1], h2 = row[2];
      const uint32_t b0 = h0 & 255u, b1 = h0 >> 8, b2 = h1 & 255u, b3 = h1 >> 8, b4 = h2 & 255u, b5 = h2 >> 8;
      p[dy * 5 + 0] = b0 | (b1 << 16);
      p[dy * 5 + 1] = b1 | (b2 << 16);
      p[dy * 5 + 2] = b2 | (b3 << 16);
      p[dy * 5 + 3] = b3 | (b4 << 16);
      p[dy * 5 + 4] = b4 | (b5 << 16);
    }
#define LM_CSWAP(a, b) cswap_u16x2(p[a], p[b]);
    LM_MEDIAN25_NET(LM_CSWAP)
#undef LM_CSWAP
    const uint32_t med = p[12];
    const int gy = y0 + r;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int gx = x0 + c + k;
      if (gy >= rows || gx >= cols) continue;
      const uint8_t v = (uint8_t)(k ? (med >> 16) : (med & 0xffffu));
      P.quant[0][(size_t)frame * P.quant_stride[0] + (size_t)gy * cols + gx] = v;
      // [OCV] DepthNormalPyramid::pyrDown: level l is the NN decimation src(2^l y, 2^l x) of the level-0 map
#pragma unroll
      for (int l = 1; l < LM_MAX_LEVELS; ++l) {
        const int mask = (1 << l) - 1;
        if (l >= P.n_levels || (gy & mask) || (gx & mask)) break;
        const int lr = rows >> l, lc = cols >> l;
        if ((gy >> l) < lr && (gx >> l) < lc) P.quant[l][(size_t)frame * P.quant_stride[l] + (size_t)(gy >> l) * lc + (gx >> l)] = v;
      }
    }
  }
}

// medianBlur(5) by COUNTING, for one-hot normal codes (every NORMAL_LUT entry is 0 or a single bit: true of upstream's table
// and of the stand-in; the host checks the table and leaves the verdict behind it, byte 8000).  A code is one of nine values
// 0 < 1 < 2 < 4 < ... < 128 = rank k 0..8; the median of a 5 x 5 window is the smallest rank whose cumulative count reaches
// 13.  Per pixel one 5-bit counter per rank (25 fits), ranks 0..5 in one word and 6..8 in a second; the 25-fold sum is
// separable (five rows, then five columns, sliding); prefix sums of the counters are ONE multiplication by 0x02108421 (no
// field overflows: every prefix is <= 25), "prefix >= 13" is bit 4 of (prefix + 3), the first such field is the median's
// rank.  About a third of the instructions of the 99-exchange network (which stays for injected tables that are not one-hot).
template <bool FAST>
__device__ __forceinline__ void dn_tile_count(const DnParams& P, int bx, int by, int frame, uint32_t* smem) {
  constexpr int RW = D_RW;
  uint32_t (*s_vs)[D_TH][RW + 4] = reinterpret_cast<uint32_t (*)[D_TH][RW + 4]>(smem);                        // [2]: sums over five rows
  uint32_t (*s_inc)[D_TH + 4][RW] = reinterpret_cast<uint32_t (*)[D_TH + 4][RW]>(smem + 2 * D_TH * (RW + 4));  // [2]: 1 << 5k in word k / 6
  const int rows = P.rows, cols = P.cols;
  const uint16_t* __restrict__ depth = static_cast<const uint16_t*>(P.ctl->ft.src[frame][P.modality]);
  const int x0 = bx * D_TW, y0 = by * D_TH;
  const int tid = threadIdx.x;
  // raw quantised normals for the tile + halo 2; medianBlur's BORDER_REPLICATE = value at the clamped coordinate
  for (int i = tid; i < (D_TH + 4) * RW; i += 256) {
    const int r = i / RW, c = i - r * RW;
    const int gy = clampi(y0 - 2 + r, 0, rows - 1), gx = clampi(x0 - 2 + c, 0, cols - 1);
    const uint32_t code = dn_normal_at<FAST>(depth, rows, cols, gy, gx, P.distance_threshold, P.difference_threshold, P.lut);
    const int k = __ffs((int)code);                          // 0 for code 0, j + 1 for 1 << j
    s_inc[0][r][c] = k < 6 ? 1u << (5 * k) : 0u;
    s_inc[1][r][c] = k < 6 ? 0u : 1u << (5 * (k - 6));
  }
  __syncthreads();
  // five-row sums, sliding down a column: task = (word, column, half of the 16 output rows)
  for (int t = tid; t < 2 * 2 * RW; t += 256) {
    const int half = t / (2 * RW), rem = t - half * (2 * RW);
    const int w = rem / RW, c = rem - w * RW;
    const int r0 = 8 * half;
    uint32_t a[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) a[j] = s_inc[w][r0 + j][c];
    uint32_t sum = a[0] + a[1] + a[2] + a[3] + a[4];
    s_vs[w][r0][c] = sum;
#pragma unroll
    for (int o = 1; o < 8; ++o) {
      sum += a[o + 4] - a[o - 1];
      s_vs[w][r0 + o][c] = sum;
    }
  }
  __syncthreads();
  // four consecutive pixels of a row per thread: five-column sums (sliding), then the median's rank
  const int r = tid >> 4, c = (tid & 15) * 4;
  uint32_t h[2][4];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const uint4 v0 = *reinterpret_cast<const uint4*>(&s_vs[w][r][c]);
    const uint4 v1 = *reinterpret_cast<const uint4*>(&s_vs[w][r][c + 4]);
    h[w][0] = v0.x + v0.y + v0.z + v0.w + v1.x;
    h[w][1] = h[w][0] - v0.x + v1.y;
    h[w][2] = h[w][1] - v0.y + v1.z;
    h[w][3] = h[w][2] - v0.z + v1.w;
  }
  const int gy = y0 + r;
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t p0 = h[0][k] * 0x02108421u;               // field j: count of ranks <= j (j = 0..5)
    const uint32_t t0 = (p0 + 0x06318C63u) & 0x21084210u;    // bit 5j+4: that count >= 13
    uint32_t rank;
    if (t0) rank = ((uint32_t)(__ffs((int)t0) - 1) * 13u) >> 6;             // 4, 9, .. 29 -> 0 .. 5
    else {
      const uint32_t p1 = (h[1][k] + ((p0 >> 25) & 31u)) * 0x421u;         // ranks 6..8 on top of the count of ranks <= 5
      const uint32_t t1 = (p1 + 0xC63u) & 0x4210u;
      rank = 6u + (((uint32_t)(__ffs((int)t1) - 1) * 13u) >> 6);
    }
    packed |= (rank ? 1u << (rank - 1) : 0u) << (8 * k);
  }
  if (gy < rows) {
    uint8_t* q0 = P.quant[0] + (size_t)frame * P.quant_stride[0] + (size_t)gy * cols;
    const int gx0 = x0 + c;
    if ((cols & 3) == 0 && gx0 + 3 < cols) *reinterpret_cast<uint32_t*>(q0 + gx0) = packed;   // quant strides are multiples of 4
    else
      for (int k = 0; k < 4; ++k)
        if (gx0 + k < cols) q0[gx0 + k] = (uint8_t)(packed >> (8 * k));
    // [OCV] DepthNormalPyramid::pyrDown: level l is the NN decimation src(2^l y, 2^l x) of the level-0 map
#pragma unroll
    for (int k = 0; k < 4; k += 2) {                          // odd columns never survive a decimation
      const int gx = gx0 + k;
      if (gx >= cols) continue;
      const uint8_t v = (uint8_t)(packed >> (8 * k));
#pragma unroll
      for (int l = 1; l < LM_MAX_LEVELS; ++l) {
        const int mask = (1 << l) - 1;
        if (l >= P.n_levels || (gy & mask) || (gx & mask)) break;
        const int lr = rows >> l, lc = cols >> l;
        if ((gy >> l) < lr && (gx >> l) < lc) P.quant[l][(size_t)frame * P.quant_stride[l] + (size_t)(gy >> l) * lc + (gx >> l)] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_dn_fused(const DnParams P) {
  const int frame = blockIdx.z;
  if (frame >= P.ctl->ft.n_frames) return;
  __shared__ __align__(16) uint32_t smem[D_COUNT_WORDS];
  static_assert(sizeof(uint32_t) * D_COUNT_WORDS >= (D_TH + 4) * (D_TW + 8), "the network path's byte tile fits");
  const bool fast = P.difference_threshold <= 200;
  if (P.lut[8000]) {   // one-hot table (checked by the host at upload): median by counting
    if (fast) dn_tile_count<true>(P, blockIdx.x, blockIdx.y, frame, smem);
    else dn_tile_count<false>(P, blockIdx.x, blockIdx.y, frame, smem);
  } else {
    if (fast) dn_tile<true>(P, blockIdx.x, blockIdx.y, frame, smem);
    else dn_tile<false>(P, blockIdx.x, blockIdx.y, frame, smem);
  }
}

// ---------------------------------------------------------------------------------------------- spread -> LM
// One CTA = R rows of T-cells x SP_CW cells of one (level, modality); R = 4 when the staging fits (T <= 8), else 2 or 1.
// Everything is done on 32-bit words (four pixels): masked quantisation -> separable OR (T-1 funnel shifts along x, T-1
// word ORs along y; with R cell rows the halo of the OR is (R+1)T-1 pixel rows for R*T) -> one LUT word per spread byte
// (8 response nibbles) -> 4x4 byte transposes (PRMT) so that every store is a full word of one (orientation, T^2 phase)
// row of the linear memories.  Nibble planes are written directly (an 8x8 nibble transpose: one word = 8 consecutive
// cells) when the level's rows are word-aligned: flat for the coarsest level, column-blocked for refinement levels, where
// the R rows of a 16-column block are 8R contiguous bytes (a full 32-byte sector at R = 4).
constexpr int SP_CW = 32;  // grid cells per block along x

__host__ __device__ inline int sp_nwo(int T) { return SP_CW * T / 4; }
__host__ __device__ inline int sp_nwq(int T) { return (SP_CW * T + T - 1 + 3) / 4 + 1; }
__host__ __device__ inline size_t sp_smem(int T, int R) {
  const int IH = (R + 1) * T - 1;
  return 1024 + (size_t)4 * (IH * sp_nwq(T) + IH * sp_nwo(T) + R * T * sp_nwo(T));
}
__host__ __device__ inline int sp_rows_per_block(int T) { return sp_smem(T, 4) <= 48 * 1024 ? 4 : (sp_smem(T, 2) <= 48 * 1024 ? 2 : 1); }

// 4x4 byte transpose: in[j] = bytes (k = 0..3) of item j  ->  out[k] = bytes (j = 0..3)
__device__ __forceinline__ void transpose4x4(const uint32_t (&in)[4], uint32_t (&out)[4]) {
  const uint32_t t0 = __byte_perm(in[0], in[1], 0x5140), t1 = __byte_perm(in[2], in[3], 0x5140);
  const uint32_t t2 = __byte_perm(in[0], in[1], 0x7362), t3 = __byte_perm(in[2], in[3], 0x7362);
  out[0] = __byte_perm(t0, t1, 0x5410); out[1] = __byte_perm(t0, t1, 0x7632);
  out[2] = __byte_perm(t2, t3, 0x5410); out[3] = __byte_perm(t2, t3, 0x7632);
}

template <int TT>
__device__ __forceinline__ void spread_tile(const SpreadParams& P, const SpreadEntry& E, uint8_t* smem, int frame) {
  const int T = TT ? TT : E.T;
  const int R = sp_rows_per_block(T);
  const int W = E.W, H = E.H, rows = E.rows, cols = E.cols;
  const int OH = R * T, IH = OH + T - 1, NWO = sp_nwo(T), NWQ = sp_nwq(T);
  uint32_t* s_resp = reinterpret_cast<uint32_t*>(smem);
  uint32_t* sq = s_resp + 256;
  uint32_t* sh = sq + IH * NWQ;
  uint32_t* sp = sh + IH * NWO;
  const int b = blockIdx.x - E.block_begin;
  const int c0 = (b % E.blocks_x) * SP_CW, a0 = (b / E.blocks_x) * R;   // first cell column / cell row of this block
  const int nr = min(R, H - a0);                                        // cell rows of this block inside the image
  const int px0 = c0 * T, py0 = a0 * T;
  const int tid = threadIdx.x;
  const uint8_t* __restrict__ qraw = E.qraw + (size_t)frame * E.qraw_stride;
  uint8_t* __restrict__ quantized = E.quantized + (size_t)frame * E.quantized_stride;
  const uint8_t* __restrict__ mask0 = E.mask0;
  uint8_t* tap_spread = frame == 0 ? E.spread : nullptr;
  uint8_t* tap_response = frame == 0 ? E.response : nullptr;
  s_resp[tid] = P.resp_all[tid];
  // ---- masked quantisation, four pixels per word; pixels outside the image are 0 ([OCV] spread stays in bounds)
  const bool fast = mask0 == nullptr && (cols & 3) == 0;
  for (int i = tid; i < IH * NWQ; i += 256) {
    const int r = i / NWQ, wi = i - r * NWQ;
    const int gy = py0 + r, gx = px0 + 4 * wi;
    uint32_t v = 0;
    const bool own = r < OH && wi < NWO;  // inside the block's own pixels: write Detector::match's quantized image
    if (fast) {
      if (gy < rows && gx < cols) {
        v = __ldg(reinterpret_cast<const uint32_t*>(qraw + (size_t)gy * cols + gx));
        if (own) *reinterpret_cast<uint32_t*>(quantized + (size_t)gy * cols + gx) = v;
      }
    } else if (gy < rows) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = gx + k;
        if (x < cols) {
          uint32_t q = qraw[(size_t)gy * cols + x];
          if (mask0 && !mask0[(size_t)(gy << E.level) * E.mask_cols0 + (x << E.level)]) q = 0;
          if (own) quantized[(size_t)gy * cols + x] = (uint8_t)q;
          v |= q << (8 * k);
        }
      }
    }
    sq[i] = v;
  }
  __syncthreads();
  // ---- OR over T pixels along x: byte x of the result needs bytes x .. x+T-1
  for (int i = tid; i < IH * NWO; i += 256) {
    const int r = i / NWO, wi = i - r * NWO;
    const uint32_t* q = sq + r * NWQ + wi;
    uint32_t v = q[0];
    if (TT) {
      uint32_t w[(TT + 2) / 4 + 2];
#pragma unroll
      for (int k = 0; k < (TT + 2) / 4 + 2; ++k) w[k] = q[k];
#pragma unroll
      for (int c = 1; c < TT; ++c) v |= __funnelshift_r(w[c >> 2], w[(c >> 2) + 1], 8 * (c & 3));
    } else {
      for (int c = 1; c < T; ++c) v |= __funnelshift_r(q[c >> 2], q[(c >> 2) + 1], 8 * (c & 3));
    }
    sh[i] = v;
  }
  __syncthreads();
  // ---- OR over T rows
  unsigned int n_bits = 0;
  for (int i = tid; i < OH * NWO; i += 256) {
    const int r = i / NWO, wi = i - r * NWO;
    uint32_t v = 0;
    if (TT) {
#pragma unroll
      for (int k = 0; k < TT; ++k) v |= sh[(r + k) * NWO + wi];
    } else {
      for (int k = 0; k < T; ++k) v |= sh[(r + k) * NWO + wi];
    }
    sp[i] = v;
    n_bits += __popc(v);
    if (tap_spread) {
      const int gy = py0 + r;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gx = px0 + 4 * wi + k;
        if (gy < rows && gx < cols) tap_spread[(size_t)gy * cols + gx] = (uint8_t)(v >> (8 * k));
      }
    }
  }
  if (E.count_bits) {  // a performance hint for the coarse kernel, not a result: see SpreadEntry::count_bits
    n_bits = __reduce_add_sync(0xffffffffu, n_bits);
    if ((tid & 31) == 0 && n_bits) atomicAdd(&P.ctl->mod_bits[frame][E.modality], n_bits);
  }
  __syncthreads();
  // ---- responses -> linear memories
  const uint8_t* spb = reinterpret_cast<const uint8_t*>(sp);
  const int row_bytes = NWO * 4;
  const size_t WH = (size_t)W * H;
  const int ncell = min(SP_CW, W - c0);
  if (E.lm_nib != nullptr) {  // nibble-packed planes: one word = 8 consecutive cells of one (orientation, phase) row
    const size_t nib_stride = (size_t)E.nib_plane / 2;
    const int Hh = E.tiled_Hh;  // refinement levels: column-blocked layout (lm_kernels.cuh tiled_nibble_index)
    for (int it = tid; it < T * T * R * (SP_CW / 8); it += 256) {
      // consecutive threads write consecutive words: flat planes -- the four words of a cell row; column-blocked planes --
      // the two words of a block row, then the block's next rows (8R contiguous bytes), then the second block
      int b8, ra, g;
      if (Hh) { b8 = (it & 1) | (((it / (2 * R)) & 1) << 1); ra = (it >> 1) % R; g = it / (4 * R); }
      else { b8 = it & 3; ra = (it >> 2) % R; g = it / (4 * R); }
      if (b8 * 8 >= ncell || ra >= nr) continue;
      const int a = a0 + ra;
      const int rs = g / T, cs = g - rs * T;
      const uint8_t* p = spb + (ra * T + rs) * row_bytes + cs + T * (b8 * 8);
      uint32_t ev[4], od[4];  // pixel pair (2p, 2p+1): even / odd orientations, one byte per orientation pair
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const uint32_t r0 = s_resp[p[T * (2 * pr)]], r1 = s_resp[p[T * (2 * pr + 1)]];
        ev[pr] = (r0 & 0x0f0f0f0fu) | ((r1 & 0x0f0f0f0fu) << 4);
        od[pr] = ((r0 >> 4) & 0x0f0f0f0fu) | (r1 & 0xf0f0f0f0u);
      }
      uint32_t oe[4], oo[4];
      transpose4x4(ev, oe);  // oe[k]: orientation 2k, nibbles = cells 0..7
      transpose4x4(od, oo);  // oo[k]: orientation 2k+1
      // nibble index inside the plane (multiple of 8)
      const size_t n0 = Hh ? tiled_nibble_index(W, Hh, g, a, c0 + b8 * 8) : (size_t)g * WH + (size_t)a * W + c0 + b8 * 8;
      uint8_t* dst = E.lm_nib + (size_t)frame * E.lm_nib_stride + n0 / 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k) * nib_stride) = oe[k];
        *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k + 1) * nib_stride) = oo[k];
      }
      if (Hh && a < 16 && g > 0) {  // rows 0..15 of a phase are also the rows H..H+15 below the previous phase
        uint8_t* halo = dst - ((size_t)W * Hh - (size_t)H * 16) / 2;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          *reinterpret_cast<uint32_t*>(halo + (size_t)(2 * k) * nib_stride) = oe[k];
          *reinterpret_cast<uint32_t*>(halo + (size_t)(2 * k + 1) * nib_stride) = oo[k];
        }
      }
    }
  }
  uint8_t* __restrict__ lm = E.lm ? E.lm + (size_t)frame * E.lm_stride : nullptr;
  if (lm != nullptr) {
    if ((W & 3) == 0) {
      for (int it = tid; it < T * T * R * (SP_CW / 4); it += 256) {
        const int b4 = it & (SP_CW / 4 - 1), ra = (it / (SP_CW / 4)) % R, g = it / ((SP_CW / 4) * R);
        if (b4 * 4 >= ncell || ra >= nr) continue;
        const int rs = g / T, cs = g - rs * T;
        const uint8_t* p = spb + (ra * T + rs) * row_bytes + cs + T * (b4 * 4);
        uint32_t ev[4], od[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t r = s_resp[p[T * j]];
          ev[j] = r & 0x0f0f0f0fu;
          od[j] = (r >> 4) & 0x0f0f0f0fu;
        }
        uint32_t oe[4], oo[4];
        transpose4x4(ev, oe);
        transpose4x4(od, oo);
        uint8_t* dst = lm + (size_t)g * WH + (size_t)(a0 + ra) * W + c0 + b4 * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k) * E.plane_stride) = oe[k];
          *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k + 1) * E.plane_stride) = oo[k];
        }
      }
    } else {
      for (int it = tid; it < T * T * R * SP_CW; it += 256) {
        const int bb = it & (SP_CW - 1), ra = (it / SP_CW) % R, g = it / (SP_CW * R);
        if (bb >= ncell || ra >= nr) continue;
        const int rs = g / T, cs = g - rs * T;
        const uint32_t r0 = s_resp[spb[(ra * T + rs) * row_bytes + cs + T * bb]];
        const size_t o = (size_t)g * WH + (size_t)(a0 + ra) * W + c0 + bb;
#pragma unroll
        for (int ori = 0; ori < 8; ++ori) lm[ori * E.plane_stride + o] = (uint8_t)((r0 >> (4 * ori)) & 15);
      }
    }
  }
  if (tap_response) {
    for (int i = tid; i < OH * NWO * 4; i += 256) {
      const int r = i / (NWO * 4), x = i - r * (NWO * 4);
      const int gy = py0 + r, gx = px0 + x;
      if (gy < rows && gx < cols) {
        const uint32_t r0 = s_resp[spb[r * row_bytes + x]];
#pragma unroll
        for (int ori = 0; ori < 8; ++ori)
          tap_response[(size_t)ori * rows * cols + (size_t)gy * cols + gx] = (uint8_t)((r0 >> (4 * ori)) & 15);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_spread_all(const SpreadParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int frame = blockIdx.y;
  if (frame >= P.ctl->ft.n_frames) return;
  int ei = 0;
  while (ei + 1 < P.n && (int)blockIdx.x >= P.e[ei + 1].block_begin) ++ei;
  const SpreadEntry& E = P.e[ei];
  if (E.T == 5) spread_tile<5>(P, E, smem, frame);        // the reference trainer's T pyramid {5, 8}
  else if (E.T == 8) spread_tile<8>(P, E, smem, frame);
  else spread_tile<0>(P, E, smem, frame);
}

size_t spread_all_smem(int T) { return sp_smem(T, sp_rows_per_block(T)); }

}  // namespace

// ================================================================================================ launchers
int cg_fused_blocks(int rows, int cols, int* blocks_x) {
  *blocks_x = (cols + C_TW - 1) / C_TW;
  return *blocks_x * ((rows + C_TH - 1) / C_TH);
}
void launch_cg_fused(const CgParams& p, int total_blocks, int n_frames, cudaStream_t s) {
  k_cg_fused<<<dim3(total_blocks, n_frames), C_THREADS, 0, s>>>(p);
}
void launch_dn_fused(const DnParams& p, int n_frames, cudaStream_t s) {
  dim3 grid((p.cols + D_TW - 1) / D_TW, (p.rows + D_TH - 1) / D_TH, n_frames);
  k_dn_fused<<<grid, 256, 0, s>>>(p);
}
void launch_pyrdown_fast(const BatchCtl* ctl, int modality, const uint8_t* src, size_t src_stride, int rows, int cols,
                         uint8_t* dst, size_t dst_stride, int n_frames, cudaStream_t s) {
  dim3 grid((cols / 2 + Y_TW - 1) / Y_TW, (rows / 2 + Y_TH - 1) / Y_TH, n_frames);
  k_pyrdown_fast<<<grid, 256, 0, s>>>(ctl, modality, src, src_stride, rows, cols, dst, dst_stride);
}
int spread_all_blocks(int T, int W, int H, int* blocks_x) {
  const int R = sp_rows_per_block(T);
  *blocks_x = (W + SP_CW - 1) / SP_CW;
  return *blocks_x * ((H + R - 1) / R);
}
bool launch_spread_all(const SpreadParams& p, int total_blocks, int n_frames, cudaStream_t s) {
  size_t smem = 0;
  for (int i = 0; i < p.n; ++i) smem = std::max(smem, spread_all_smem(p.e[i].T));
  if (smem > 48 * 1024) return false;  // T <= 16 (checked by the host) needs 41 KB at one cell row per block
  k_spread_all<<<dim3(total_blocks, n_frames), 256, smem, s>>>(p);
  return true;
}

}  // namespace lmk
