// lm_frontend_fused.cu -- the production front end: three launches per frame instead of fourteen.
//
//   k_cg_fused    [OCV] quantizedOrientations + hysteresisGradient for every pyramid level in one grid:
//                 GaussianBlur 7x7 -> Sobel 3x3 -> max-magnitude channel -> fastAtan2 -> 16-bin rounding -> 3x3 vote,
//                 all staged through shared memory (the per-level source of levels >= 1 comes from k_pyrdown_u8c3).
//   k_dn_fused    [OCV] quantizedNormals incl. medianBlur(5) and DepthNormalPyramid::pyrDown: plane fit + LUT, a
//                 99-exchange median-of-25 network on packed u16x2 (VIMNMX.U16x2), NN decimation to all levels.
//   k_spread_all  [OCV] quantize(mask) + spread + computeResponseMaps + linearize for every (level, modality).
//
// The front end is O(pixels) and tiny next to a B200 (7.7 MB of algorithmic traffic per 640x480 frame), so it is bound
// by launch count and dependent-phase latency, not by bandwidth: fusing removes the global round trips between stages.
// Arithmetic is identical to the stage-by-stage kernels in lm_frontend.cu (kept as the A/B reference).
#include <float.h>

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
constexpr int C_TW = 64, C_TH = 8;

__global__ void __launch_bounds__(256) k_cg_fused(const CgParams P) {
  __shared__ uint8_t s_in[C_TH + 10][(C_TW + 10) * 3];   // source, halo 5, replicate-extended
  __shared__ uint16_t s_h[C_TH + 10][(C_TW + 4) * 3];    // horizontal blur pass
  __shared__ uint8_t s_sm[C_TH + 4][(C_TW + 4) * 3];     // smoothed, halo 2
  __shared__ uint8_t s_q[C_TH + 2][C_TW + 2];            // unfiltered quantisation, halo 1
  __shared__ float s_mag[C_TH][C_TW];
  int lvl = 0;
  while (lvl + 1 < P.n_levels && (int)blockIdx.x >= P.lv[lvl + 1].block_begin) ++lvl;
  const uint8_t* __restrict__ src = P.lv[lvl].src;
  const int rows = P.lv[lvl].rows, cols = P.lv[lvl].cols;
  const int b = blockIdx.x - P.lv[lvl].block_begin;
  const int x0 = (b % P.lv[lvl].blocks_x) * C_TW, y0 = (b / P.lv[lvl].blocks_x) * C_TH;
  const int tid = threadIdx.x;

  constexpr int IN_W = (C_TW + 10) * 3;
  for (int i = tid; i < (C_TH + 10) * IN_W; i += 256) {
    int r = i / IN_W, rem = i - r * IN_W;
    int cx = rem / 3, c = rem - cx * 3;
    int gy = clampi(y0 - 5 + r, 0, rows - 1), gx = clampi(x0 - 5 + cx, 0, cols - 1);
    s_in[r][rem] = src[((size_t)gy * cols + gx) * 3 + c];
  }
  __syncthreads();
  constexpr int H_W = (C_TW + 4) * 3;
  for (int i = tid; i < (C_TH + 10) * H_W; i += 256) {
    int r = i / H_W, e = i - r * H_W;
    const uint8_t* p = &s_in[r][e];
    s_h[r][e] = (uint16_t)(8 * (p[0] + p[18]) + 28 * (p[3] + p[15]) + 56 * (p[6] + p[12]) + 72 * p[9]);
  }
  __syncthreads();
  for (int i = tid; i < (C_TH + 4) * H_W; i += 256) {
    int r = i / H_W, e = i - r * H_W;
    int s = 8 * ((int)s_h[r][e] + s_h[r + 6][e]) + 28 * ((int)s_h[r + 1][e] + s_h[r + 5][e]) +
            56 * ((int)s_h[r + 2][e] + s_h[r + 4][e]) + 72 * (int)s_h[r + 3][e];
    s_sm[r][e] = (uint8_t)((s + 32768) >> 16);
  }
  __syncthreads();
  // Sobel on the smoothed image with BORDER_REPLICATE: out-of-image neighbours read the smoothed value at the clamped
  // coordinate (which is inside this tile whenever the tile touches the border).
  for (int i = tid; i < (C_TH + 2) * (C_TW + 2); i += 256) {
    int r = i / (C_TW + 2), x = i - r * (C_TW + 2);
    int gy = y0 - 1 + r, gx = x0 - 1 + x;
    uint8_t q8 = 0;
    if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
      const int ra = clampi(gy - 1, 0, rows - 1) - (y0 - 2), rb = gy - (y0 - 2), rd = clampi(gy + 1, 0, rows - 1) - (y0 - 2);
      const int ca = (clampi(gx - 1, 0, cols - 1) - (x0 - 2)) * 3, cb = (gx - (x0 - 2)) * 3,
                cd = (clampi(gx + 1, 0, cols - 1) - (x0 - 2)) * 3;
      int m[3], dxs[3], dys[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int aa = s_sm[ra][ca + c], ab = s_sm[ra][cb + c], ad = s_sm[ra][cd + c];
        int ba = s_sm[rb][ca + c], bd = s_sm[rb][cd + c];
        int da = s_sm[rd][ca + c], db = s_sm[rd][cb + c], dd = s_sm[rd][cd + c];
        int dx = (ad + 2 * bd + dd) - (aa + 2 * ba + da);
        int dy = (da + 2 * db + dd) - (aa + 2 * ab + ad);
        dxs[c] = dx; dys[c] = dy; m[c] = dx * dx + dy * dy;
      }
      int bm, bdx, bdy;
      if (m[0] >= m[1] && m[0] >= m[2]) { bm = m[0]; bdx = dxs[0]; bdy = dys[0]; }
      else if (m[1] >= m[0] && m[1] >= m[2]) { bm = m[1]; bdx = dxs[1]; bdy = dys[1]; }
      else { bm = m[2]; bdx = dxs[2]; bdy = dys[2]; }
      float angle = fast_atan2_deg((float)bdy, (float)bdx);
      int q = clampi(__float2int_rn(__fmul_rn(angle, (float)(16.0 / 360.0))), 0, 255);
      const bool border = gy == 0 || gy == rows - 1 || gx == 0 || gx == cols - 1;
      q8 = border ? 0 : (uint8_t)(q & 7);
      if (r >= 1 && r <= C_TH && x >= 1 && x <= C_TW) {
        s_mag[r - 1][x - 1] = (float)bm;
        P.lv[lvl].mag[(size_t)gy * cols + gx] = (float)bm;
      }
    }
    s_q[r][x] = q8;
  }
  __syncthreads();
  for (int i = tid; i < C_TH * C_TW; i += 256) {
    int r = i / C_TW, x = i - r * C_TW;
    int gy = y0 + r, gx = x0 + x;
    if (gy >= rows || gx >= cols) continue;
    uint8_t out = 0;
    if (gy >= 1 && gy < rows - 1 && gx >= 1 && gx < cols - 1 && s_mag[r][x] > P.thr_sq) {
      unsigned hist = 0;
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) hist += 1u << (4 * s_q[r + j][x + k]);
      int max_votes = 0, index = -1;
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) {
        int v = (hist >> (4 * bb)) & 15;
        if (max_votes < v) { index = bb; max_votes = v; }
      }
      if (max_votes >= 5) out = (uint8_t)(1 << index);
    }
    P.lv[lvl].quant[(size_t)gy * cols + gx] = out;
  }
}

// ---------------------------------------------------------------------------------------------- DepthNormal
__device__ __forceinline__ uint8_t dn_normal_at(const uint16_t* __restrict__ depth, int rows, int cols, int y, int x,
                                                int distance_threshold, int difference_threshold,
                                                const uint8_t* __restrict__ lut) {
  const int r = 5;
  if (!(y >= r && y < rows - r - 1 && x >= r && x < cols - r - 1)) return 0;
  long long d = depth[(size_t)y * cols + x];
  if (!(d < distance_threshold)) return 0;
  long long A0 = 0, A1 = 0, A3 = 0, b0 = 0, b1 = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int kk = k < 4 ? k : k + 1;
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
  float s = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  if (!(s > 0)) return 0;
  float inv = __fdiv_rn(1.0f, s);
  nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
  int v1 = __float2int_rz(__fadd_rn(__fmul_rn(nx, 10.0f), 10.0f));
  int v2 = __float2int_rz(__fadd_rn(__fmul_rn(ny, 10.0f), 10.0f));
  int v3 = __float2int_rz(__fadd_rn(__fmul_rn(nz, 20.0f), 20.0f));
  int flat = (v3 * 20 + v2) * 20 + v1;
  return (flat >= 0 && flat < 8000) ? lut[flat] : (uint8_t)0;
}

constexpr int D_TW = 64, D_TH = 8;

__device__ __forceinline__ void cswap_u16x2(uint32_t& a, uint32_t& b) {
  uint32_t lo, hi;
  asm("min.u16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
  asm("max.u16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
  a = lo; b = hi;
}

__global__ void __launch_bounds__(256) k_dn_fused(const DnParams P) {
  __shared__ uint8_t s_raw[D_TH + 4][D_TW + 4 + 4];
  const int rows = P.rows, cols = P.cols;
  const int x0 = blockIdx.x * D_TW, y0 = blockIdx.y * D_TH;
  const int tid = threadIdx.x;
  // raw quantised normals for the tile + halo 2; medianBlur's BORDER_REPLICATE = value at the clamped coordinate
  for (int i = tid; i < (D_TH + 4) * (D_TW + 4); i += 256) {
    int r = i / (D_TW + 4), c = i - r * (D_TW + 4);
    int gy = clampi(y0 - 2 + r, 0, rows - 1), gx = clampi(x0 - 2 + c, 0, cols - 1);
    s_raw[r][c] = dn_normal_at(P.depth, rows, cols, gy, gx, P.distance_threshold, P.difference_threshold, P.lut);
  }
  __syncthreads();
  // median of 25 for two horizontally adjacent pixels at once (one per 16-bit lane)
  const int r = tid >> 5, c = (tid & 31) * 2;
  uint32_t p[25];
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    uint32_t b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = s_raw[r + dy][c + k];
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) p[dy * 5 + dx] = b[dx] | (b[dx + 1] << 16);
  }
#define LM_CSWAP(a, b) cswap_u16x2(p[a], p[b]);
  LM_MEDIAN25_NET(LM_CSWAP)
#undef LM_CSWAP
  const uint32_t med = p[12];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int gy = y0 + r, gx = x0 + c + k;
    if (gy >= rows || gx >= cols) continue;
    const uint8_t v = (uint8_t)(k ? (med >> 16) : (med & 0xffffu));
    P.quant[0][(size_t)gy * cols + gx] = v;
    // [OCV] DepthNormalPyramid::pyrDown: level l is the NN decimation src(2^l y, 2^l x) of the level-0 map
    for (int l = 1; l < P.n_levels; ++l) {
      const int mask = (1 << l) - 1;
      if ((gy & mask) || (gx & mask)) break;
      const int lr = rows >> l, lc = cols >> l;
      if ((gy >> l) < lr && (gx >> l) < lc) P.quant[l][(size_t)(gy >> l) * lc + (gx >> l)] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------- spread -> LM
// One CTA = one row of T-cells x SP_CW cells of one (level, modality).  Everything is done on 32-bit words (four
// pixels): masked quantisation -> separable OR (T-1 funnel shifts along x, T-1 word ORs along y) -> one LUT word per
// spread byte (8 response nibbles) -> 4x4 byte transposes (PRMT) so that every store is a full word of one
// (orientation, T^2 phase) row of the linear memories.  The coarsest level is written nibble-packed directly (an 8x8
// nibble transpose: one word = 8 consecutive cells) when its rows are word-aligned.
constexpr int SP_CW = 32;  // grid cells per block along x

__host__ __device__ inline int sp_nwo(int T) { return SP_CW * T / 4; }
__host__ __device__ inline int sp_nwq(int T) { return (SP_CW * T + T - 1 + 3) / 4 + 1; }

// 4x4 byte transpose: in[j] = bytes (k = 0..3) of item j  ->  out[k] = bytes (j = 0..3)
__device__ __forceinline__ void transpose4x4(const uint32_t (&in)[4], uint32_t (&out)[4]) {
  const uint32_t t0 = __byte_perm(in[0], in[1], 0x5140), t1 = __byte_perm(in[2], in[3], 0x5140);
  const uint32_t t2 = __byte_perm(in[0], in[1], 0x7362), t3 = __byte_perm(in[2], in[3], 0x7362);
  out[0] = __byte_perm(t0, t1, 0x5410); out[1] = __byte_perm(t0, t1, 0x7632);
  out[2] = __byte_perm(t2, t3, 0x5410); out[3] = __byte_perm(t2, t3, 0x7632);
}

template <int TT>
__device__ __forceinline__ void spread_tile(const SpreadParams& P, const SpreadEntry& E, uint8_t* smem) {
  const int T = TT ? TT : E.T;
  const int W = E.W, H = E.H, rows = E.rows, cols = E.cols;
  const int IH = 2 * T - 1, NWO = sp_nwo(T), NWQ = sp_nwq(T);
  uint32_t* s_resp = reinterpret_cast<uint32_t*>(smem);
  uint32_t* sq = s_resp + 256;
  uint32_t* sh = sq + IH * NWQ;
  uint32_t* sp = sh + IH * NWO;
  const int b = blockIdx.x - E.block_begin;
  const int c0 = (b % E.blocks_x) * SP_CW, a = b / E.blocks_x;
  const int px0 = c0 * T, py0 = a * T;
  const int tid = threadIdx.x;
  const uint8_t* __restrict__ qraw = E.qraw;
  const uint8_t* __restrict__ mask0 = E.mask0;
  s_resp[tid] = P.resp_all[tid];
  // ---- masked quantisation, four pixels per word; pixels outside the image are 0 ([OCV] spread stays in bounds)
  const bool fast = mask0 == nullptr && (cols & 3) == 0;
  for (int i = tid; i < IH * NWQ; i += 256) {
    const int r = i / NWQ, wi = i - r * NWQ;
    const int gy = py0 + r, gx = px0 + 4 * wi;
    uint32_t v = 0;
    const bool own = r < T && wi < NWO;  // inside the block's own T x (SP_CW*T) pixels: write Detector::match's quantized image
    if (fast) {
      if (gy < rows && gx < cols) {
        v = __ldg(reinterpret_cast<const uint32_t*>(qraw + (size_t)gy * cols + gx));
        if (own) *reinterpret_cast<uint32_t*>(E.quantized + (size_t)gy * cols + gx) = v;
      }
    } else if (gy < rows) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = gx + k;
        if (x < cols) {
          uint32_t q = qraw[(size_t)gy * cols + x];
          if (mask0 && !mask0[(size_t)(gy << E.level) * E.mask_cols0 + (x << E.level)]) q = 0;
          if (own) E.quantized[(size_t)gy * cols + x] = (uint8_t)q;
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
  for (int i = tid; i < T * NWO; i += 256) {
    const int r = i / NWO, wi = i - r * NWO;
    uint32_t v = 0;
    if (TT) {
#pragma unroll
      for (int k = 0; k < TT; ++k) v |= sh[(r + k) * NWO + wi];
    } else {
      for (int k = 0; k < T; ++k) v |= sh[(r + k) * NWO + wi];
    }
    sp[i] = v;
    if (E.spread) {
      const int gy = py0 + r;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int gx = px0 + 4 * wi + k;
        if (gy < rows && gx < cols) E.spread[(size_t)gy * cols + gx] = (uint8_t)(v >> (8 * k));
      }
    }
  }
  __syncthreads();
  // ---- responses -> linear memories
  const uint8_t* spb = reinterpret_cast<const uint8_t*>(sp);
  const int row_bytes = NWO * 4;
  const size_t WH = (size_t)W * H;
  const int ncell = min(SP_CW, W - c0);
  if (E.lm_nib != nullptr) {  // coarsest level, nibble-packed: one word = 8 consecutive cells of one (orientation, phase) row
    const size_t nib_stride = (size_t)E.plane_stride / 2;
    for (int it = tid; it < T * T * (SP_CW / 8); it += 256) {
      const int b8 = it & (SP_CW / 8 - 1), g = it / (SP_CW / 8);
      if (b8 * 8 >= ncell) continue;
      const int rs = g / T, cs = g - rs * T;
      const uint8_t* p = spb + rs * row_bytes + cs + T * (b8 * 8);
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
      const size_t n0 = (size_t)g * WH + (size_t)a * W + c0 + b8 * 8;  // nibble index inside the plane (multiple of 8)
      uint8_t* dst = E.lm_nib + n0 / 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k) * nib_stride) = oe[k];
        *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k + 1) * nib_stride) = oo[k];
      }
    }
  }
  uint8_t* __restrict__ lm = E.lm;
  if (lm != nullptr) {
    if ((W & 3) == 0) {
      for (int it = tid; it < T * T * (SP_CW / 4); it += 256) {
        const int b4 = it & (SP_CW / 4 - 1), g = it / (SP_CW / 4);
        if (b4 * 4 >= ncell) continue;
        const int rs = g / T, cs = g - rs * T;
        const uint8_t* p = spb + rs * row_bytes + cs + T * (b4 * 4);
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
        uint8_t* dst = lm + (size_t)g * WH + (size_t)a * W + c0 + b4 * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k) * E.plane_stride) = oe[k];
          *reinterpret_cast<uint32_t*>(dst + (size_t)(2 * k + 1) * E.plane_stride) = oo[k];
        }
      }
    } else {
      for (int it = tid; it < T * T * SP_CW; it += 256) {
        const int bb = it & (SP_CW - 1), g = it / SP_CW;
        if (bb >= ncell) continue;
        const int rs = g / T, cs = g - rs * T;
        const uint32_t r0 = s_resp[spb[rs * row_bytes + cs + T * bb]];
        const size_t o = (size_t)g * WH + (size_t)a * W + c0 + bb;
#pragma unroll
        for (int ori = 0; ori < 8; ++ori) lm[ori * E.plane_stride + o] = (uint8_t)((r0 >> (4 * ori)) & 15);
      }
    }
  }
  if (E.response) {
    for (int i = tid; i < T * NWO * 4; i += 256) {
      const int r = i / (NWO * 4), x = i - r * (NWO * 4);
      const int gy = py0 + r, gx = px0 + x;
      if (gy < rows && gx < cols) {
        const uint32_t r0 = s_resp[spb[r * row_bytes + x]];
#pragma unroll
        for (int ori = 0; ori < 8; ++ori)
          E.response[(size_t)ori * rows * cols + (size_t)gy * cols + gx] = (uint8_t)((r0 >> (4 * ori)) & 15);
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_spread_all(const SpreadParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  int ei = 0;
  while (ei + 1 < P.n && (int)blockIdx.x >= P.e[ei + 1].block_begin) ++ei;
  const SpreadEntry& E = P.e[ei];
  if (E.T == 5) spread_tile<5>(P, E, smem);        // the reference trainer's T pyramid {5, 8}
  else if (E.T == 8) spread_tile<8>(P, E, smem);
  else spread_tile<0>(P, E, smem);
}

size_t spread_all_smem(int T) {
  const int IH = 2 * T - 1;
  return 1024 + (size_t)4 * (IH * sp_nwq(T) + IH * sp_nwo(T) + T * sp_nwo(T));
}

}  // namespace

// ================================================================================================ launchers
int cg_fused_blocks(int rows, int cols, int* blocks_x) {
  *blocks_x = (cols + C_TW - 1) / C_TW;
  return *blocks_x * ((rows + C_TH - 1) / C_TH);
}
void launch_cg_fused(const CgParams& p, int total_blocks, cudaStream_t s) {
  k_cg_fused<<<total_blocks, 256, 0, s>>>(p);
}
void launch_dn_fused(const DnParams& p, cudaStream_t s) {
  dim3 grid((p.cols + D_TW - 1) / D_TW, (p.rows + D_TH - 1) / D_TH);
  k_dn_fused<<<grid, 256, 0, s>>>(p);
}
int spread_all_blocks(int W, int H, int* blocks_x) {
  *blocks_x = (W + SP_CW - 1) / SP_CW;
  return *blocks_x * H;
}
bool launch_spread_all(const SpreadParams& p, int total_blocks, int max_T, cudaStream_t s) {
  size_t smem = spread_all_smem(max_T);
  if (smem > 48 * 1024) return false;  // T <= 16 (checked by the host) needs 41 KB
  k_spread_all<<<total_blocks, 256, smem, s>>>(p);
  return true;
}

}  // namespace lmk
