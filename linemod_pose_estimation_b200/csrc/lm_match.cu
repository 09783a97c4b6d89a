// lm_match.cu -- the template-matching kernels of the LINEMOD hot path (sm_100a).
//
// k_similarity_coarse_rec63 / k_similarity_coarse_rec restate [OCV] similarity + addSimilarities + the coarse scan of Detector::matchClass
// (OpenCV 2.4.x objdetect/linemod.cpp, reached from /root/reference/src/rgbdDetector.cpp:33); k_refine_nib restates
// [OCV] similarityLocal and matchClass's refinement loop.  Spec: SURVEY.md App. A.7-A.9, quirks App. D.  Both take a
// CHUNK of frames per launch (frame table + per-frame strides, lm_kernels.cuh); k_begin_chunk installs the table.
//
// The work is a byte gather-accumulate: S_t[j] = sum_f LM[a_f + j].  The linear memories of a frame (nibble-packed,
// 0.3 MB per modality at the coarsest level of 640x480) are shared by every template and stay L2/L1 resident; HBM only
// sees the template records.  Hence no tensor cores: the binding resources are instruction issue, L1 wavefronts and load
// latency.  Responses are <= 4 and a template has <= 63 features per modality, so byte lanes packed in a 32-bit
// register never carry into each other: plain integer adds are bit-identical to the reference's _mm_add_epi8 (and to
// __vaddus4) at a quarter of the instruction count.
#include <string.h>

#include <algorithm>

#include "lm_kernels.cuh"

namespace lmk {

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint4 ldg128(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint2 ldg64(const uint8_t* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ uint32_t ldg32(const uint8_t* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }

// Responses are <= 4, so the linear memories are kept packed two positions per byte: position p of the reference's flat
// byte plane is nibble p.  That halves the bytes every window load moves through L1/L2.  Up to three features are summed
// in the nibble domain (3 x 4 = 12 < 16: no carry between positions) before the even / odd nibbles are spread into the
// u8 accumulators the reference uses, so the sums are bit-identical to _mm_add_epi8 over bytes.
template <int WORDS>
__device__ __forceinline__ void nib_flush(const uint32_t (&nib)[WORDS], uint32_t (&acc_e)[WORDS], uint32_t (&acc_o)[WORDS]) {
#pragma unroll
  for (int k = 0; k < WORDS; ++k) {
    acc_e[k] += nib[k] & 0x0f0f0f0fu;
    acc_o[k] += (nib[k] >> 4) & 0x0f0f0f0fu;
  }
}

// ------------------------------------------------------------------------------------------------ production coarse
// similarity_coarse_body (k_similarity_coarse_rec63 / k_similarity_coarse_rec): nibble-packed linear memories + self-contained
// tile records + exact early termination.
//
//  * The host plan stores one record per (template, 1024-position pass) tile: a 48-byte header followed by the feature
//    words (aligned chunk byte offset of lane 0's window | nibble shift).  A warp prefetches the record of its next tile
//    (and draws the one after from the dispenser) while it scores the current one, so the tile bookkeeping adds no
//    dependent global-load latency between tiles.
//  * Early termination ([OCV] matchClass keeps a position only if raw > raw_threshold): a response is at most 4, so
//    after `done` of the tile's n features a position can still pass only if partial + 4 * (n - done) > raw_threshold.
//    The warp stops as soon as none of its 1024 positions can.  Tiles that survive are summed to the end, so every
//    reported raw score is the complete sum: the candidate list is bit-identical to the exhaustive scan.  At the
//    reference's thresholds (92 / 94) random templates stop after a third of their features; trained ones gather 65 %
//    (most of a frame scores high for most of a template: DepthNormal on flat background).
//    Disabled (prune = 0) for the parity tap, which wants every position's full sum.
constexpr int kRecHdrWords = 12;                      // TileRec header, see lm_kernels.cuh
constexpr int kRecMaxWords = kRecHdrWords + 256;      // header + LM_MAX_MODALITIES * 64 feature words
constexpr int kRecWarpPos = 32 * 32;

// ---- bulk asynchronous copy (TMA's linear form) + mbarrier, the mechanism that stages a warp's next tile record in shared
// memory while the current tile is scored: one elected lane arms the barrier with the byte count and issues
// cp.async.bulk (SASS: UBLKCP); the copy engine writes shared memory and completes the transaction on the barrier; the warp
// waits on the barrier's phase.  No registers hold the record in flight and no LDG / STS instructions move it.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LM_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LM_DONE;\n"
      "bra LM_WAIT;\n"
      "LM_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int Q>
__device__ __forceinline__ void rec_window(const uint8_t* __restrict__ lmn, uint32_t v, uint32_t lane_byte,
                                           uint32_t (&n)[4]) {
  const uint8_t* p = lmn + (v & ~15u);   // lmn already points at this lane's first chunk (plane base + 16 * lane)
  const uint32_t sh = v << 2;  // funnel shifts use the low five bits: 4 * (v & 7)
  uint32_t w[8];
  const uint4 c = ldg128(p);
  w[0] = c.x; w[1] = c.y; w[2] = c.z; w[3] = c.w;
  if (Q == 0) {
    w[4] = ldg32(p + 16);
  } else if (Q == 1) {
    const uint2 e = ldg64(p + 16);
    w[4] = e.x; w[5] = e.y;
  } else {
    const uint4 e = ldg128(p + 16);
    w[4] = e.x; w[5] = e.y; w[6] = e.z; w[7] = e.w;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) n[k] = __funnelshift_r(w[Q + k], w[Q + k + 1], sh);
}

// The five words w[Q .. Q+4] a window of class Q needs (the rest of the second vector load is dropped at once).
template <int Q>
__device__ __forceinline__ void rec_load(const uint8_t* __restrict__ lmn, uint32_t v, uint32_t lane_byte, uint32_t (&x)[5]) {
  const uint8_t* p = lmn + (v & ~15u);   // lmn already points at this lane's first chunk (plane base + 16 * lane)
  const uint4 c = ldg128(p);
  if (Q == 0) {
    x[0] = c.x; x[1] = c.y; x[2] = c.z; x[3] = c.w; x[4] = ldg32(p + 16);
  } else if (Q == 1) {
    const uint2 e = ldg64(p + 16);
    x[0] = c.y; x[1] = c.z; x[2] = c.w; x[3] = e.x; x[4] = e.y;
  } else {
    const uint4 e = ldg128(p + 16);
    if (Q == 2) { x[0] = c.z; x[1] = c.w; x[2] = e.x; x[3] = e.y; x[4] = e.z; }
    else { x[0] = c.w; x[1] = e.x; x[2] = e.y; x[3] = e.z; x[4] = e.w; }
  }
}

// `zero` is a run-time zero the compiler cannot see through.  OR-ing (last loaded word & zero) into the first word makes
// every shift of the group depend on the LAST load, so all twelve loads of a group are in flight before the first result
// is consumed; left alone, ptxas recycles the destination registers of the first loads for the later ones and turns one
// round trip to L2 per group into two or three -- on the kernel's critical path, the tiles that survive every check.
// (Measured alternatives that lost: rounds of nine features with predicated loads, an L1 prefetch of the next class.)
template <int Q>
__device__ __forceinline__ void rec_group(const uint8_t* __restrict__ lmn, const uint32_t* fw, int n, uint32_t lane_byte,
                                          uint32_t (&acc_e)[4], uint32_t (&acc_o)[4], uint32_t zero) {
  int f = 0;
  for (; f + 6 <= n; f += 6) {
    uint32_t x[6][5], a[6][4];
#pragma unroll
    for (int i = 0; i < 6; ++i) rec_load<Q>(lmn, fw[f + i], lane_byte, x[i]);
    x[0][0] |= x[5][4] & zero;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const uint32_t sh = fw[f + i] << 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) a[i][k] = __funnelshift_r(x[i][k], x[i][k + 1], sh);
    }
    uint32_t s0[4], s1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { s0[k] = a[0][k] + a[1][k] + a[2][k]; s1[k] = a[3][k] + a[4][k] + a[5][k]; }
    nib_flush<4>(s0, acc_e, acc_o);
    nib_flush<4>(s1, acc_e, acc_o);
  }
  if (f + 3 <= n) {
    uint32_t a[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i) rec_window<Q>(lmn, fw[f + i], lane_byte, a[i]);
    uint32_t s0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) s0[k] = a[0][k] + a[1][k] + a[2][k];
    nib_flush<4>(s0, acc_e, acc_o);
    f += 3;
  }
  if (f < n) {  // one or two left
    uint32_t a0[4], a1[4];
    rec_window<Q>(lmn, fw[f], lane_byte, a0);
    if (f + 1 < n) {
      rec_window<Q>(lmn, fw[f + 1], lane_byte, a1);
#pragma unroll
      for (int k = 0; k < 4; ++k) a0[k] += a1[k];
    }
    nib_flush<4>(a0, acc_e, acc_o);
  }
}
// (Code size matters here: the SM's instruction cache holds 32 KB and the warps of a CTA are at different places of the
// loop.  A variant with one batch template per tail length 1..5 -- one round trip fewer for n = 4, 5, 10, 11 -- grew the hot
// code from about 21 KB to 32 KB and measured 40 % SLOWER, profiles/r02_experiments.md.)

// u8 sums of at most 63 features -> added to the u16 totals ([OCV] addSimilarities widening), accumulators cleared.
__device__ __forceinline__ void rec_widen(uint32_t (&acc_e)[4], uint32_t (&acc_o)[4], uint32_t (&tot)[4][4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    tot[k][0] += acc_e[k] & 0x00ff00ffu;
    tot[k][1] += (acc_e[k] >> 8) & 0x00ff00ffu;
    tot[k][2] += acc_o[k] & 0x00ff00ffu;
    tot[k][3] += (acc_o[k] >> 8) & 0x00ff00ffu;
    acc_e[k] = 0; acc_o[k] = 0;
  }
}

// Can any position of the warp's tile still exceed raw_threshold, `remaining` features (<= 4 each) to go?
__device__ __forceinline__ bool rec_alive(const uint32_t (&tot)[4][4], int thr, int remaining, bool active) {
  const int need = thr - 4 * remaining;  // a position passes only if partial > need
  if (need < 0) return true;             // warp-uniform
  uint32_t mx = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int r = 0; r < 4; ++r) mx = __vmaxu2(mx, tot[k][r]);
  const int best = (int)max(mx & 0xffffu, mx >> 16);
  return __any_sync(kFull, active && best > need);
}

// NARROW tiles (at most 63 features in all: the reference's two-modality detector has 31 + 31 at its coarsest level) never
// leave the u8 accumulators: 63 * 4 < 256.  Largest of the lane's 32 sums: bytes 0, 2 and 1, 3 of every register as u16 pairs.
__device__ __forceinline__ uint32_t rec_max_narrow(const uint32_t (&acc_e)[4], const uint32_t (&acc_o)[4]) {
  uint32_t mx = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    mx = __vmaxu2(mx, __vmaxu2(acc_e[k] & 0x00ff00ffu, __byte_perm(acc_e[k], 0u, 0x4341)));
    mx = __vmaxu2(mx, __vmaxu2(acc_o[k] & 0x00ff00ffu, __byte_perm(acc_o[k], 0u, 0x4341)));
  }
  return max(mx & 0xffffu, mx >> 16);
}
__device__ __forceinline__ bool rec_alive_narrow(const uint32_t (&acc_e)[4], const uint32_t (&acc_o)[4], int thr, int remaining,
                                                 bool active) {
  const int need = thr - 4 * remaining;  // a position passes only if partial > need
  if (need < 0) return true;             // warp-uniform
  return __any_sync(kFull, active && (int)rec_max_narrow(acc_e, acc_o) > need);
}

// One kernel body, two instantiations: NARROW (every tile of the request has at most 63 features: u8 sums only, no u16
// totals -- sixteen registers fewer, three CTAs per SM) and the general one (u8 sums widened into u16 totals every half
// modality, two CTAs per SM).
template <bool NARROW>
__device__ __forceinline__ void similarity_coarse_body(const CoarseParams& P) {
  constexpr bool BULK = true;   // (the register-staged alternative measured 4-6 % slower: profiles/r02_experiments.md)
  __shared__ __align__(16) uint32_t s_rec[8][2][kRecMaxWords];
  __shared__ __align__(8) unsigned long long s_bar[8][2];
  __shared__ unsigned long long s_bytes;
  __shared__ uint32_t s_done, s_expected;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* sr = s_rec[warp][0];
  int buf = 0;
  uint32_t phases = 0;   // bit b: parity of the next completion of barrier b
  if (BULK && lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cudaGridDependencySynchronize();            // linear memories (previous kernel in the stream) are complete
  cudaTriggerProgrammaticLaunchCompletion();  // let k_refine's blocks be scheduled as this grid drains
  BatchCtl* ctl = P.ctl;
  const uint32_t* __restrict__ recs = P.recs;
  const int rec_words = P.rec_words, M = P.M, prune = P.prune;
  // Virtual tiles of the chunk.  The first n_full records are FULL tiles, one (frame, tile) per warp, frame-major:
  // v = frame * n_full + tile.  The rest are SHARED tiles -- passes of at most 128 positions (the tail of a template whose
  // span exceeds one 1 024-position pass): one warp scores the tile for EIGHT frames at once, four lanes per frame, instead
  // of eight warps with four busy lanes each; v = n_full * F + frame group * n_shared + (tile - n_full).
  const uint32_t n_tiles = (uint32_t)P.n_tiles, n_full = (uint32_t)P.n_full, n_shared = n_tiles - n_full;
  const uint32_t n_frames = (uint32_t)ctl->ft.n_frames;
  const uint32_t v_full = n_full * n_frames;
  const uint32_t n_virtual = v_full + n_shared * ((n_frames + 7u) >> 3);
  // Tile hand-out: the first two tiles of every warp are static (tile = global warp index, then + number of warps) --
  // thousands of warps drawing from one counter at kernel start serialise on that address -- and only the rest comes from
  // the atomic dispenser (tiles are sorted heaviest first, so the dynamic part balances the light tail).
  const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
  uint32_t cur = gwarp;
  if (threadIdx.x == 0) {
    s_bytes = 0; s_done = 0;
    const uint32_t first_warp = blockIdx.x * (blockDim.x >> 5);  // warps of this CTA that have a first tile
    s_expected = n_virtual > first_warp ? min(n_virtual - first_warp, blockDim.x >> 5) : 0u;
  }
  __syncthreads();
  if (cur >= n_virtual) return;
  // virtual tile -> (record, first frame, shared?)
  auto decode = [&](uint32_t v, uint32_t& tile, uint32_t& frame0) -> bool {
    if (v < v_full) { frame0 = v / n_full; tile = v - frame0 * n_full; return false; }
    const uint32_t u = v - v_full, fg = u / n_shared;
    tile = n_full + (u - fg * n_shared); frame0 = fg * 8u;
    return true;
  };
  uint32_t cur_tile, cur_frame;
  bool cur_shared = decode(cur, cur_tile, cur_frame);
  const uint32_t rec_bytes = (uint32_t)rec_words * 4u;   // a multiple of 16 (records are padded to four words)
  if (BULK) {
    if (lane == 0) {
      mbar_expect_tx(&s_bar[warp][0], rec_bytes);
      bulk_g2s(sr, recs + (size_t)cur_tile * rec_words, rec_bytes, &s_bar[warp][0]);
    }
    mbar_wait(&s_bar[warp][0], 0u);
    phases = 1u;
  }
  uint32_t nxt = gwarp + n_warps;
  const bool do_prune = (prune & 1) != 0;
  const uint32_t zero = n_tiles >> 31;  // n_tiles > 0: zero, but only at run time (see rec_group)
  unsigned long long bytes = 0;  // (feature, position) pairs actually gathered by this warp
  uint32_t thr_key = 0xffffffffu;
  int thr_val = 0;
  for (;;) {
    __syncwarp();
    // prefetch the next tile's record into registers and draw the tile after it
    const bool has_next = nxt < n_virtual;
    uint32_t nxt_tile = 0, nxt_frame = 0;
    const bool nxt_shared = has_next ? decode(nxt, nxt_tile, nxt_frame) : false;
    // the other buffer was last read during the previous tile (every lane is past the __syncwarp above): refill it
    if (has_next && lane == 0) {
      mbar_expect_tx(&s_bar[warp][buf ^ 1], rec_bytes);
      bulk_g2s(s_rec[warp][buf ^ 1], recs + (size_t)nxt_tile * rec_words, rec_bytes, &s_bar[warp][buf ^ 1]);
    }
    // the ticket of the tile after next: issued now, read at the end of this tile (the atomic's round trip -- long when
    // thousands of warps draw at once -- overlaps the scoring instead of stalling it)
    uint32_t ticket = 0;
    if (has_next && lane == 0) ticket = atomicAdd(&ctl->next_tile, 1u);

    // this lane's chunk of every window.  The empty asm pins the pointer in a register pair: left alone, ptxas recomputes
    // frame * stride + lane offset + base for every feature (six address instructions per window instead of three).
    // a full tile: 32 lanes x 32 positions of frame cur_frame; a shared tile: lanes 4f .. 4f+3 take frame cur_frame + f
    const uint32_t lane_frame = cur_shared ? cur_frame + ((uint32_t)lane >> 2) : cur_frame;
    const uint32_t lane_slot = cur_shared ? ((uint32_t)lane & 3u) : (uint32_t)lane;
    const int first = (int)lane_slot * 32;  // first position of this lane within the pass
    unsigned long long lane_base = (unsigned long long)P.lmn + (unsigned long long)lane_frame * P.lmn_stride + lane_slot * 16u;
    asm volatile("" : "+l"(lane_base));
    const uint8_t* __restrict__ lmn = reinterpret_cast<const uint8_t*>(lane_base);
    // bit 8 of `prune`: sum the modalities in reverse order; bit 9: decide per frame from the front end's counters
    const bool mod_reversed = (prune & 0x200) ? (ctl->mod_bits[cur_frame][M - 1] < ctl->mod_bits[cur_frame][0]) : (prune & 0x100) != 0;
    const uint32_t item = sr[0], tg = sr[1], nfq = sr[2], order = sr[6];
    const int n_feat = (int)sr[3], j0 = (int)sr[4], rem = (int)sr[5];
    const bool active = first < rem && lane_frame < n_frames;
    // [OCV] matchClass: raw_threshold = (int)(2*nf + (threshold / 100.f) * (2*nf) + 0.5f), same f32 roundings; consecutive
    // tiles nearly always share the query and the feature count, so the division is redone only when they change
    if (nfq != thr_key) {
      const float threshold = P.thr.v[nfq >> 28];
      const float two_nf = (float)(2 * (int)(nfq & 0x0fffffffu));
      thr_val = __float2int_rz(__fadd_rn(__fadd_rn(two_nf, __fmul_rn(__fdiv_rn(threshold, 100.f), two_nf)), 0.5f));
      thr_key = nfq;
    }
    const int thr = thr_val;
    // u16 x 2 per register: [k][r], r = (i & 1) * 2 + ((i >> 1) & 1) for position 8k + i (general tiles only)
    uint32_t tot[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tot[k][0] = tot[k][1] = tot[k][2] = tot[k][3] = 0;
    uint32_t acc_e[4] = {0, 0, 0, 0}, acc_o[4] = {0, 0, 0, 0};  // u8 x 4: position 8k + i is byte i / 2 of (i & 1 ? acc_o : acc_e)[k]
    int done = 0;
    bool alive = true;
    for (int mi = 0; mi < M && alive; ++mi) {
      // The order in which the modalities are summed is free (the pruning bound is exact for any order and survivors are
      // summed completely): start with the modality the front end found more discriminative on this frame.
      const int m = mod_reversed ? M - 1 - mi : mi;
      const uint32_t* fw = sr + kRecHdrWords;
      for (int k = 0; k < m; ++k) fw += __dp4a(sr[8 + k], 0x01010101u, 0u);  // features of the modalities before m
      const uint32_t c4 = sr[8 + m];  // 4 class sizes, one word
      const int n0 = c4 & 255, n1 = (c4 >> 8) & 255, n2 = (c4 >> 16) & 255, n3 = c4 >> 24;
      if (active) {
        rec_group<0>(lmn, fw, n0, 0u, acc_e, acc_o, zero);
        rec_group<1>(lmn, fw + n0, n1, 0u, acc_e, acc_o, zero);
      }
      fw += n0 + n1; done += n0 + n1;
      if constexpr (NARROW) {
        if (do_prune && !rec_alive_narrow(acc_e, acc_o, thr, n_feat - done, active)) { alive = false; break; }
      } else {
        rec_widen(acc_e, acc_o, tot);
        if (do_prune && !rec_alive(tot, thr, n_feat - done, active)) { alive = false; break; }
      }
      if (active) {
        rec_group<2>(lmn, fw, n2, 0u, acc_e, acc_o, zero);
        rec_group<3>(lmn, fw + n2, n3, 0u, acc_e, acc_o, zero);
      }
      done += n2 + n3;
      if constexpr (NARROW) {
        if (do_prune && !rec_alive_narrow(acc_e, acc_o, thr, n_feat - done, active)) alive = false;
      } else {
        rec_widen(acc_e, acc_o, tot);
        if (do_prune && !rec_alive(tot, thr, n_feat - done, active)) alive = false;
      }
    }
    bytes += (unsigned long long)done * (unsigned)rem * (cur_shared ? min(8u, n_frames - cur_frame) : 1u);
    if (alive && active) {
      bool hit = thr < 0;
      if (!hit) {
        if constexpr (NARROW) {
          hit = (int)rec_max_narrow(acc_e, acc_o) > thr;
        } else {
          const uint32_t thr2 = (uint32_t)min(thr, 0xffff) * 0x00010001u;
          uint32_t any = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int r = 0; r < 4; ++r) any |= __vcmpgtu2(tot[k][r], thr2);
          hit = any != 0;
        }
      }
      if (P.dump != nullptr || hit) {  // rare: [OCV] matchClass raster scan, "raw_score > raw_threshold"
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int p = first + 8 * k + i;
            int raw;
            if constexpr (NARROW) {
              raw = (int)((((i & 1) ? acc_o[k] : acc_e[k]) >> (8 * (i >> 1))) & 0xffu);
            } else {
              const uint32_t src = tot[k][(i & 1) * 2 + ((i >> 1) & 1)];
              raw = (int)((i & 4) ? (src >> 16) : (src & 0xffffu));
            }
            if (p < rem) {
              if (P.dump != nullptr) P.dump[(size_t)item * P.dump_stride + j0 + p] = (uint16_t)raw;
              if (raw > thr) {
                uint32_t idx = atomicAdd(&ctl->n_cands, 1u);
                if (idx < P.cand_cap) {
                  Cand c;
                  c.tglob = tg; c.pos = (uint32_t)(j0 + p) | (lane_frame << 24);
                  c.raw_nf = (uint32_t)raw | ((nfq & 0xffffu) << 16); c.order = order;
                  P.cand[idx] = c;
                }
              }
            }
          }
      }
    }
    if (!has_next) break;
    __syncwarp();
    buf ^= 1;
    sr = s_rec[warp][buf];
    mbar_wait(&s_bar[warp][buf], (phases >> buf) & 1u);
    phases ^= 1u << buf;
    cur_frame = nxt_frame; cur_shared = nxt_shared;
    nxt = __shfl_sync(kFull, ticket, 0) + 2u * n_warps;
  }
  // gathered-bytes statistic: summed per CTA in shared memory, one global atomic per CTA by the warp that finishes last
  // (one atomic per warp on a single address costs microseconds at the end of a launch of thousands of warps)
  if (lane == 0 && P.touched != nullptr) {
    atomicAdd(&s_bytes, bytes);
    __threadfence_block();
    if (atomicAdd(&s_done, 1u) + 1u == s_expected) atomicAdd(P.touched, s_bytes);
  }
}

__global__ void __launch_bounds__(256, 2) k_similarity_coarse_rec(const CoarseParams P) { similarity_coarse_body<false>(P); }
__global__ void __launch_bounds__(256, 3) k_similarity_coarse_rec63(const CoarseParams P) { similarity_coarse_body<true>(P); }
// (A/B: the u8-only body at two CTAs per SM, to tell the gain of the leaner body from that of the third CTA)
__global__ void __launch_bounds__(256, 2) k_similarity_coarse_rec63_2cta(const CoarseParams P) { similarity_coarse_body<true>(P); }

// Byte linear memories -> nibble-packed copy (two positions per byte), 16 bytes in / 8 bytes out per thread; blockIdx.y = frame.
__global__ void __launch_bounds__(256) k_pack_nibbles(const uint8_t* __restrict__ src0, size_t src_stride,
                                                      uint8_t* __restrict__ dst0, size_t dst_stride, size_t n16,
                                                      const BatchCtl* __restrict__ ctl) {
  const int frame = blockIdx.y;
  if (frame >= ctl->ft.n_frames) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n16) return;
  const uint4 v = reinterpret_cast<const uint4*>(src0 + (size_t)frame * src_stride)[i];
  // bytes b0 b1 b2 b3 of a word -> nibbles b0 | b1<<4 | b2<<8 | b3<<12
  auto squeeze = [](uint32_t w) -> uint32_t {
    uint32_t t = (w | (w >> 4)) & 0x00ff00ffu;  // b0|b1<<4 in byte 0, b2|b3<<4 in byte 2
    return (t | (t >> 8)) & 0xffffu;
  };
  uint2 o;
  o.x = squeeze(v.x) | (squeeze(v.y) << 16);
  o.y = squeeze(v.z) | (squeeze(v.w) << 16);
  reinterpret_cast<uint2*>(dst0 + (size_t)frame * dst_stride)[i] = o;
}

// Everything that differs between two replays of a lane's CUDA graph, installed by one tiny launch in front of it: the
// frame table arrives as a kernel parameter (copied at launch time: no host buffer to keep alive), the coarse kernel's
// dispenser / counters and the (statistics + header) prefix of the lane's result blocks are zeroed.
__global__ void __launch_bounds__(256) k_begin_chunk(const FrameTable ft, BatchCtl* ctl, uint8_t* results, size_t result_stride,
                                                     int n_blocks) {
  const int tid = threadIdx.x;
  constexpr int kWords = (int)(sizeof(BatchCtl) / 4), kFtWords = (int)(sizeof(FrameTable) / 4);
  const uint32_t* in = reinterpret_cast<const uint32_t*>(&ft);
  uint32_t* out = reinterpret_cast<uint32_t*>(ctl);
  for (int i = tid; i < kWords; i += 256) out[i] = i < kFtWords ? in[i] : 0u;
  for (int i = tid; i < n_blocks * 8; i += 256)   // 16 B statistics + ResultHeader = 8 words per block
    reinterpret_cast<uint32_t*>(results + (size_t)(i >> 3) * result_stride)[i & 7] = 0u;
}

constexpr int kRefineWarps = 8;
constexpr int kRefineMaxFeat = LM_MAX_MODALITIES * 64;
constexpr size_t kResultStatsBytes = 16;  // statistics in front of every frame's ResultHeader

// [OCV] matchClass drops a refined candidate when best_score * 100.f / (4 * nf) < threshold (f32, two roundings).  The
// predicate is monotone in the integer score, so there is a smallest passing score: found here with the very same f32
// operations, it lets the warp-per-candidate kernel stop a candidate exactly as soon as no position of its 16 x 16 window
// can reach it any more (a response is at most 4 per remaining feature).
// Called by a whole (converged) warp: the lanes test the 32 scores from an estimate two below the threshold at once and the
// first passing one is the answer; the sequential search remains for the cases the estimate does not settle.
__device__ __forceinline__ int min_passing_score(float threshold, int nf) {
  const float den = (float)(4 * nf);
  const int s0 = max(0, __float2int_rd(threshold * den * 0.01f) - 2);
  const int cap = 4 * nf + 1;  // scores never exceed 4 * nf: `cap` means "cannot pass"
  {
    const int t = s0 + (int)(threadIdx.x & 31u);
    const bool pass = t >= cap || !(__fdiv_rn(__fmul_rn((float)t, 100.f), den) < threshold);
    const unsigned m = __ballot_sync(kFull, pass);
    // lane 0 failing (or s0 = 0) means no smaller score passes: the predicate is monotone in the score
    if (m != 0u && ((m & 1u) == 0u || s0 == 0)) return min(cap, s0 + __ffs((int)m) - 1);
  }
  int s = s0;
  while (s < cap && __fdiv_rn(__fmul_rn((float)s, 100.f), den) < threshold) ++s;
  while (s > 0 && !(__fdiv_rn(__fmul_rn((float)(s - 1), 100.f), den) < threshold)) --s;
  return s;
}

// Where a feature's 16 x 16 window starts, for one lane of the address phase.  Flat planes: the nibble index of the
// window's first position (rows W apart).  Column-blocked planes (RefineLevel::Hh != 0): two words -- byte offsets of the
// window's two 16-column chunks at row 0 (rows 8 bytes apart; offsets are multiples of 8 below 2^31) with the nibble shift
// (0..15) in the first one's spare bits: w0 = offset0 | (shift & 7) | (shift >> 3) << 31, w1 = offset1.  Features outside
// the image (and the padding slots of the warp path, pk = kNoFeature) point at the plane's zero run.
constexpr uint32_t kNoFeature = 0u;   // packed feature (x, y) = (-4096, -4096): outside every image
template <bool TILED>
__device__ __forceinline__ void refine_feature_address(const RefineLevel& L, uint32_t pk, int offset_x, int offset_y, uint32_t WH,
                                                       uint32_t zero_run, uint32_t* s_addr, int i) {
  const uint32_t T = (uint32_t)L.T, W = (uint32_t)L.W;
  const int fx = (int)(pk & 0x1fffu) - 4096 + offset_x, fy = (int)((pk >> 13) & 0x1fffu) - 4096 + offset_y;
  const bool inside = fx >= 0 && fy >= 0 && fx < L.cols && fy < L.rows;  // "Discard feature if out of bounds"
  const uint32_t label = pk >> 26;
  // fx / T, fx % T (and fy) without integer divisions: coordinates are < 8192 and T <= 16, so the high word of the product
  // with ceil(2^32 / T) is the exact quotient; coordinates outside the image are never used
  const uint32_t ux = (uint32_t)fx, uy = (uint32_t)fy;
  const uint32_t col = T == 1u ? ux : __umulhi(ux, L.inv_T), row = T == 1u ? uy : __umulhi(uy, L.inv_T);
  const uint32_t phase = (uy - row * T) * T + (ux - col * T);
  if (!TILED) {
    const uint32_t addr = label * (uint32_t)L.plane_stride + phase * WH + row * W + col;
    s_addr[i] = inside ? addr : zero_run;
  } else {
    const uint32_t Hh = (uint32_t)L.Hh, block_bytes = Hh * 8u, phase_bytes = W * Hh / 2u;
    const uint32_t cb = col >> 4, sh = col & 15u;
    const uint32_t phase0 = label * (uint32_t)(L.plane_stride / 2) + phase * phase_bytes;
    uint32_t b0 = phase0 + cb * block_bytes + row * 8u;
    // second chunk: the next column block, or -- past the last one -- column block 0 one row down (the flat order's successor)
    uint32_t b1 = (cb + 1u < (W >> 4)) ? b0 + block_bytes : phase0 + (row + 1u) * 8u;
    uint32_t s = sh;
    if (!inside) { b0 = b1 = zero_run; s = 0; }
    s_addr[2 * i] = b0 | (s & 7u) | ((s >> 3) << 31);
    s_addr[2 * i + 1] = b1;
  }
}

// Refinement on nibble-packed planes, ONE BLOCK PER CANDIDATE (few candidates: lowest latency).  A patch row is 16 positions =
// 8 bytes at an arbitrary nibble offset: lane r (< 16) and lane r + 16 load the two aligned 8-byte chunks the row spans
// -- one LDG.64 per feature and lane -- lane r takes the second chunk from its partner with two shuffles, realigns with two
// funnel shifts and sums up to three features in the nibble domain before spreading even / odd nibbles into byte
// accumulators.  One level of one candidate; returns through s_state (x, y, alive, best_score).
template <bool TILED>
__device__ __forceinline__ void refine_block_level(const RefineParams& P, const RefineLevel& L, const RefineTpl* rtp, uint32_t frame,
                                                   float threshold, int x, int y, uint32_t (*s_part)[16][4], uint32_t* s_addr,
                                                   int* s_state) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int prow = lane & 15, half = lane >> 4;
  const int T = L.T, W = L.W;
  const int off = T / 2 + (T % 2 - 1);
  const int offset_x = (x / T - 8) * T, offset_y = (y / T - 8) * T;
  const uint32_t WH = (uint32_t)W * (uint32_t)(L.rows / T);
  // the tail of every plane is zero (App. D-2 padding; 128 zero bytes close a column-blocked plane)
  const uint32_t zero_run = TILED ? (uint32_t)(L.plane_stride / 2) - 128u : (uint32_t)L.plane_stride - 32u;
  int n_all = 0;
  for (int m = 0; m < P.M; ++m) n_all += rtp->cnt[m];
  const uint32_t* fp = L.feats + rtp->feat_begin;
  for (int i = threadIdx.x; i < n_all; i += kRefineWarps * 32)
    refine_feature_address<TILED>(L, fp[i], offset_x, offset_y, WH, zero_run, s_addr, i);
  __syncthreads();
  const uint32_t row_off = (uint32_t)(prow * W);
  uint32_t tot_e[2] = {0, 0}, tot_o[2] = {0, 0};  // u8 x 4: even / odd columns 0..7 and 8..15 of row prow (<= 16 * 4 per warp)
  int begin = 0;
  for (int m = 0; m < P.M; ++m) {
    const uint8_t* lmm = L.lmn + (size_t)frame * L.frame_stride + (size_t)m * 4 * L.plane_stride + (TILED ? prow * 8 : 0);
    const int n = rtp->cnt[m];  // <= 63: at most 8 features per warp
    uint2 w[8];
    uint32_t sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int f = warp + kRefineWarps * k;
      if (TILED) {
        const uint2 a = f < n ? reinterpret_cast<const uint2*>(s_addr)[begin + f] : make_uint2(zero_run, zero_run);
        sh[k] = (a.x & 7u) | ((a.x >> 31) << 3);
        w[k] = ldg64(lmm + (half ? a.y : (a.x & 0x7ffffff8u)));
      } else {
        const uint32_t base = f < n ? s_addr[begin + f] : zero_run;
        const uint32_t nidx = base + (base == zero_run ? 0u : row_off);   // nibble index of this row's first position
        sh[k] = nidx & 15u;
        w[k] = ldg64(lmm + (size_t)((nidx >> 4) + (uint32_t)half) * 8u);
      }
    }
    uint32_t nib0 = 0, nib1 = 0;
    int in_group = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t w2 = __shfl_down_sync(kFull, w[k].x, 16), w3 = __shfl_down_sync(kFull, w[k].y, 16);
      const bool hi = sh[k] >= 8u;                 // window starts in the second word of the first chunk
      const uint32_t a = hi ? w[k].y : w[k].x, b = hi ? w2 : w[k].y, cc = hi ? w3 : w2;
      const uint32_t bits = (sh[k] & 7u) * 4u;
      nib0 += __funnelshift_r(a, b, bits);
      nib1 += __funnelshift_r(b, cc, bits);
      if (++in_group == 3 || k == 7) {             // <= 3 features per nibble sum (3 * 4 < 16)
        tot_e[0] += nib0 & 0x0f0f0f0fu; tot_o[0] += (nib0 >> 4) & 0x0f0f0f0fu;
        tot_e[1] += nib1 & 0x0f0f0f0fu; tot_o[1] += (nib1 >> 4) & 0x0f0f0f0fu;
        nib0 = nib1 = 0; in_group = 0;
      }
    }
    begin += n;
  }
  if (half == 0) {
    s_part[warp][prow][0] = tot_e[0]; s_part[warp][prow][1] = tot_o[0];
    s_part[warp][prow][2] = tot_e[1]; s_part[warp][prow][3] = tot_o[1];
  }
  __syncthreads();
  if (warp == 0) {
    uint32_t best_key = 0;
    if (half == 0) {
      // u16 totals over the 8 warps: [j][0] = bytes 0, 2 of word j, [j][1] = bytes 1, 3
      uint32_t sum[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) { sum[j][0] = 0; sum[j][1] = 0; }
#pragma unroll
      for (int w2 = 0; w2 < kRefineWarps; ++w2)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t v = s_part[w2][prow][j];
          sum[j][0] += v & 0x00ff00ffu;
          sum[j][1] += (v >> 8) & 0x00ff00ffu;
        }
      // word j: j = 0 even columns 0..7, 1 odd columns 0..7, 2 even columns 8..15, 3 odd columns 8..15;
      // byte b of a word is column 8 * (j / 2) + 2 * b + (j & 1)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t src = sum[j][b & 1];
          const uint32_t sc = (b & 2) ? (src >> 16) : (src & 0xffffu);
          const int col = 8 * (j >> 1) + 2 * b + (j & 1);
          best_key = max(best_key, (sc << 8) | (uint32_t)(255 - (prow * 16 + col)));
        }
    }
    best_key = __reduce_max_sync(kFull, best_key);  // first maximum in raster order
    if (lane == 0) {
      const int best_score = (int)(best_key >> 8);
      int best_r = -1, best_c = -1;
      if (best_score > 0) {
        int idx = 255 - (int)(best_key & 0xffu);
        best_r = idx >> 4; best_c = idx & 15;
      }
      const int nfl = (int)rtp->nf;
      float sim = __fdiv_rn(__fmul_rn((float)best_score, 100.f), (float)(4 * nfl));
      s_state[0] = (x / T - 8 + best_c) * T + off;
      s_state[1] = (y / T - 8 + best_r) * T + off;
      s_state[2] = (sim < threshold) ? 0 : 1;  // [OCV] remove_if(MatchPredicate(threshold))
      s_state[3] = best_score;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void refine_nib_block(const RefineParams& P, const CoarseTpl* __restrict__ ctpl,
                                                 const Cand* __restrict__ cand, uint32_t n_cands, uint8_t* results,
                                                 size_t result_stride, uint32_t out_cap, uint32_t* smem) {
  uint32_t (*s_part)[16][4] = reinterpret_cast<uint32_t (*)[16][4]>(smem);                       // [kRefineWarps][16][4]
  uint32_t* s_addr = smem + kRefineWarps * 16 * 4;  // [2 * kRefineMaxFeat] window address words of each feature
  int* s_state = reinterpret_cast<int*>(s_addr + 2 * kRefineMaxFeat);  // x, y, alive, best_score
  for (uint32_t ci = blockIdx.x; ci < n_cands; ci += gridDim.x) {
    const Cand c = cand[ci];
    const uint32_t order = c.order;
    const uint32_t frame = c.pos >> 24, cpos = c.pos & 0x00ffffffu;
    ResultHeader* hdr = reinterpret_cast<ResultHeader*>(results + (size_t)frame * result_stride + kResultStatsBytes);
    lm_raw_match* __restrict__ out = reinterpret_cast<lm_raw_match*>(hdr + 1);
    const float threshold = P.threshold[order >> 28];
    const int cT = P.coarse_T;
    const int coff = cT / 2 + (cT % 2 - 1);
    int x = (int)(cpos % (uint32_t)P.coarse_W) * cT + coff;
    int y = (int)(cpos / (uint32_t)P.coarse_W) * cT + coff;
    uint32_t score = c.raw_nf & 0xffffu, nf = c.raw_nf >> 16;
    bool alive = true;
    for (int l = P.levels - 2; l >= 0 && alive; --l) {
      const RefineLevel& L = P.level[l];
      const RefineTpl* rtp = L.tpl + c.tglob;
      const int border = 8 * L.T;
      const int max_x = L.cols - rtp->width - border, max_y = L.rows - rtp->height - border;
      x = x * 2 + 1; y = y * 2 + 1;
      x = max(x, border); y = max(y, border);
      x = min(x, max_x); y = min(y, max_y);
      if (L.Hh) refine_block_level<true>(P, L, rtp, frame, threshold, x, y, s_part, s_addr, s_state);
      else refine_block_level<false>(P, L, rtp, frame, threshold, x, y, s_part, s_addr, s_state);
      x = s_state[0]; y = s_state[1]; alive = s_state[2] != 0; score = (uint32_t)s_state[3]; nf = rtp->nf;
      __syncthreads();  // s_state / s_part / s_addr are rewritten by the next level
    }
    if (threadIdx.x == 0) atomicAdd(&hdr->n_cands, 1u);
    if (alive && threadIdx.x == 0) {
      uint32_t idx = atomicAdd(&hdr->count, 1u);
      if (idx < out_cap) {
        lm_raw_match r;
        r.order_key = order; r.coarse_pos = cpos; r.x = x; r.y = y; r.score = score; r.nf = nf;
        r.template_id = ctpl[c.tglob].template_id; r.class_index = ctpl[c.tglob].class_index;
        out[idx] = r;
      } else hdr->overflow = 1;
    }
  }
}

// Refinement on nibble-packed planes, ONE WARP PER CANDIDATE: no block barriers, candidates of a frame are refined by up to
// 148 x 32 warps at once.  A patch row is 16 positions = 8 bytes at an arbitrary nibble offset.  Lanes r and r + 16 share
// patch row r and split the features between them (even / odd feature of a pair): every lane fetches the two aligned
// 8-byte chunks its window spans itself (2 LDG.64), realigns with two funnel shifts and sums pairs of features in the
// nibble domain before spreading even / odd nibbles into byte accumulators (widened to u16 after each modality: <= 63
// features x 4); the two halves' partial sums meet (one shuffle per register) only at the pruning checks and at the end.
// The warp first turns the template's features into window addresses in its slice of shared memory (one lane per
// feature).  On column-blocked planes (TILED) the 16 rows of a chunk are 128 contiguous bytes: a warp-wide load touches
// two to four cache lines instead of 32.  One level of one candidate; returns false when the candidate is dropped.
__device__ __forceinline__ bool refine_warp_level_flat(const RefineParams& P, const RefineLevel& L, const RefineTpl* rtp, uint32_t frame,
                                                  float threshold, const BatchCtl* ctl, uint32_t* s_addr, int& x, int& y,
                                                  uint32_t& score, uint32_t& nf) {
  const int lane = threadIdx.x & 31;
  const int prow = lane & 15, half = lane >> 4;
  const int T = L.T, W = L.W;
  const int off = T / 2 + (T % 2 - 1);
  const int offset_x = (x / T - 8) * T, offset_y = (y / T - 8) * T;
  const uint32_t WH = (uint32_t)W * (uint32_t)(L.rows / T);
  const uint32_t zero_run = (uint32_t)L.plane_stride - 32u;   // the tail of every plane is zero (App. D-2 padding)
  int n_all = 0;
  for (int m = 0; m < P.M; ++m) n_all += rtp->cnt[m];
  const uint32_t* fp = L.feats + rtp->feat_begin;
  __syncwarp();
  for (int i = lane; i < n_all; i += 32) refine_feature_address<false>(L, fp[i], offset_x, offset_y, WH, zero_run, s_addr, i);
  __syncwarp();
  const uint32_t row_off = (uint32_t)(prow * W);
  // u16 totals of row prow: [j][h], word j: 0 even columns 0..7, 1 odd 0..7, 2 even 8..15, 3 odd 8..15; h: bytes (0,2) / (1,3)
  uint32_t tot[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) tot[j][0] = tot[j][1] = 0;
  int begin = 0;
  // exact early termination (see min_passing_score): every 16 features the warp checks whether any position can still pass
  const int need = P.prune ? min_passing_score(threshold, (int)rtp->nf) : 0;
  const bool mod_reversed = P.M > 1 && (P.mod_order == 2 ? ctl->mod_bits[frame][P.M - 1] < ctl->mod_bits[frame][0] : P.mod_order == 1);
  int remaining = n_all;
  bool hopeless = false;
  auto best_so_far = [&](const uint32_t (&acc)[4], bool with_acc) -> int {
    uint32_t mx = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t t0 = tot[j][0] + (with_acc ? (acc[j] & 0x00ff00ffu) : 0u);
      uint32_t t1 = tot[j][1] + (with_acc ? ((acc[j] >> 8) & 0x00ff00ffu) : 0u);
      t0 += __shfl_xor_sync(kFull, t0, 16);
      t1 += __shfl_xor_sync(kFull, t1, 16);
      mx = __vmaxu2(mx, __vmaxu2(t0, t1));
    }
    return (int)__reduce_max_sync(kFull, max(mx & 0xffffu, mx >> 16));
  };
  for (int mi = 0; mi < P.M && !hopeless; ++mi) {
    // the sum does not depend on the order of the modalities; template order unless the host asks for the reverse
    const int m = mod_reversed ? P.M - 1 - mi : mi;
    begin = 0;
    for (int k = 0; k < m; ++k) begin += rtp->cnt[k];
    const uint8_t* lmm = L.lmn + (size_t)frame * L.frame_stride + (size_t)m * 4 * L.plane_stride;
    const int n = rtp->cnt[m];  // <= 63 features, 32 per half: the u8 sums below cannot overflow
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int f0 = 0; f0 < n; f0 += 8) {
      if ((f0 & 8) == 0 && f0 > 0 && need > 0) {   // after 16, 32, 48 features of this modality (and see below at its end)
        if (best_so_far(acc, true) + 4 * remaining < need) { hopeless = true; break; }
      }
      remaining -= min(8, n - f0);
      uint2 c0[4], c1[4];
      uint32_t sh[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int fi = f0 + 2 * k + half;
        const uint32_t base = fi < n ? s_addr[begin + fi] : zero_run;
        const uint32_t nidx = base + (base == zero_run ? 0u : row_off);   // nibble index of this row's first position
        sh[k] = nidx & 15u;
        const uint8_t* p = lmm + (size_t)(nidx >> 4) * 8u;
        c0[k] = ldg64(p);
        c1[k] = ldg64(p + 8);
      }
      uint32_t nib0 = 0, nib1 = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool hi = sh[k] >= 8u;                 // window starts in the second word of the first chunk
        const uint32_t a = hi ? c0[k].y : c0[k].x, b = hi ? c1[k].x : c0[k].y, cc = hi ? c1[k].y : c1[k].x;
        const uint32_t bits = (sh[k] & 7u) * 4u;
        nib0 += __funnelshift_r(a, b, bits);
        nib1 += __funnelshift_r(b, cc, bits);
        if (k == 1 || k == 3) {                      // <= 3 features per nibble sum (3 * 4 < 16); two here
          acc[0] += nib0 & 0x0f0f0f0fu; acc[1] += (nib0 >> 4) & 0x0f0f0f0fu;
          acc[2] += nib1 & 0x0f0f0f0fu; acc[3] += (nib1 >> 4) & 0x0f0f0f0fu;
          nib0 = nib1 = 0;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // [OCV] similarityLocal totals are u16: widen this modality's u8 sums
      tot[j][0] += acc[j] & 0x00ff00ffu;
      tot[j][1] += (acc[j] >> 8) & 0x00ff00ffu;
    }
    if (need > 0 && !hopeless && mi + 1 < P.M) {   // between modalities
      const uint32_t none[4] = {0, 0, 0, 0};
      if (best_so_far(none, false) + 4 * remaining < need) hopeless = true;
    }
  }
  if (hopeless) return false;   // [OCV] would finish the sum and drop the candidate: sim < threshold
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // the two halves of a row meet
    tot[j][0] += __shfl_xor_sync(kFull, tot[j][0], 16);
    tot[j][1] += __shfl_xor_sync(kFull, tot[j][1], 16);
  }
  // first maximum in raster order: key = score << 8 | (255 - raster index); byte b of word j is column
  // 8 * (j / 2) + 2 * b + (j & 1)
  uint32_t best_key = 0;
  if (half == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t src = tot[j][b & 1];
        const uint32_t sc = (b & 2) ? (src >> 16) : (src & 0xffffu);
        const int col = 8 * (j >> 1) + 2 * b + (j & 1);
        best_key = max(best_key, (sc << 8) | (uint32_t)(255 - (prow * 16 + col)));
      }
  }
  best_key = __reduce_max_sync(kFull, best_key);
  const int best_score = (int)(best_key >> 8);
  int best_r = -1, best_c = -1;
  if (best_score > 0) {
    const int idx = 255 - (int)(best_key & 0xffu);
    best_r = idx >> 4; best_c = idx & 15;
  }
  const int nfl = (int)rtp->nf;
  const float sim = __fdiv_rn(__fmul_rn((float)best_score, 100.f), (float)(4 * nfl));
  x = (x / T - 8 + best_c) * T + off;
  y = (y / T - 8 + best_r) * T + off;
  score = (uint32_t)best_score; nf = rtp->nf;
  return !(sim < threshold);  // [OCV] remove_if(MatchPredicate(threshold))
}

// The same level on COLUMN-BLOCKED planes (RefineLevel::Hh != 0; lm_kernels.cuh tiled_nibble_index): the 16 rows of a window
// chunk are 128 contiguous bytes, so a warp-wide load touches two to four cache lines instead of 32, and -- W being a
// multiple of 16 -- a feature's shift is the same in every row: the address phase leaves (offset0 | shift, offset1) per
// feature, a modality's slots padded to a multiple of 16 with zero-run entries, and a step of the sum is eight features per
// lane (sixteen per warp) without a bounds test: LDS.64, two LDG.64, three selects, two funnel shifts, nibble sums of
// 3 + 3 + 2 features between the even / odd splits.
__device__ __forceinline__ bool refine_warp_level_tiled(const RefineParams& P, const RefineLevel& L, const RefineTpl* rtp, uint32_t frame,
                                                        float threshold, const BatchCtl* ctl, uint32_t* s_addr, int& x, int& y,
                                                        uint32_t& score, uint32_t& nf) {
  const int lane = threadIdx.x & 31;
  const int prow = lane & 15, half = lane >> 4;
  const int T = L.T;
  const int off = T / 2 + (T % 2 - 1);
  const int offset_x = (x / T - 8) * T, offset_y = (y / T - 8) * T;
  const uint32_t zero_run = (uint32_t)(L.plane_stride / 2) - 128u;   // 128 zero bytes close every plane
  const uint32_t* fp = L.feats + rtp->feat_begin;
  int n_all = 0;
  __syncwarp();
  {   // address phase: modality m's features go to slots seg .. seg + cnt[m], the rest of its 16-slot groups is padding
    int seg = 0, first = 0;
    for (int m = 0; m < P.M; ++m) {
      const int n = rtp->cnt[m], padded = (n + 15) & ~15;
      for (int i = lane; i < padded; i += 32)
        refine_feature_address<true>(L, i < n ? fp[first + i] : kNoFeature, offset_x, offset_y, 0u, zero_run, s_addr, seg + i);
      seg += padded; first += n; n_all += n;
    }
  }
  __syncwarp();
  const uint2* s_pair = reinterpret_cast<const uint2*>(s_addr);
  // u16 totals of row prow: [j][h], word j: 0 even columns 0..7, 1 odd 0..7, 2 even 8..15, 3 odd 8..15; h: bytes (0,2) / (1,3)
  uint32_t tot[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) tot[j][0] = tot[j][1] = 0;
  // exact early termination (see min_passing_score): every 16 features the warp checks whether any position can still pass
  const int need = P.prune ? min_passing_score(threshold, (int)rtp->nf) : 0;
  const bool mod_reversed = P.M > 1 && (P.mod_order == 2 ? ctl->mod_bits[frame][P.M - 1] < ctl->mod_bits[frame][0] : P.mod_order == 1);
  int remaining = n_all;
  bool hopeless = false;
  auto best_so_far = [&](const uint32_t (&acc)[4], bool with_acc) -> int {
    uint32_t mx = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t t0 = tot[j][0] + (with_acc ? (acc[j] & 0x00ff00ffu) : 0u);
      uint32_t t1 = tot[j][1] + (with_acc ? ((acc[j] >> 8) & 0x00ff00ffu) : 0u);
      t0 += __shfl_xor_sync(kFull, t0, 16);
      t1 += __shfl_xor_sync(kFull, t1, 16);
      mx = __vmaxu2(mx, __vmaxu2(t0, t1));
    }
    return (int)__reduce_max_sync(kFull, max(mx & 0xffffu, mx >> 16));
  };
  for (int mi = 0; mi < P.M && !hopeless; ++mi) {
    // the sum does not depend on the order of the modalities; template order unless the host asks for the reverse
    const int m = mod_reversed ? P.M - 1 - mi : mi;
    int seg = 0;
    for (int k = 0; k < m; ++k) seg += (rtp->cnt[k] + 15) & ~15;
    // this lane's row of every window chunk of modality m; pinned in a register pair (see the coarse kernel)
    unsigned long long lane_base = (unsigned long long)L.lmn + (unsigned long long)frame * L.frame_stride +
                                   (unsigned long long)m * 4ull * L.plane_stride + (unsigned long long)(prow * 8);
    asm volatile("" : "+l"(lane_base));
    const uint8_t* lmm = reinterpret_cast<const uint8_t*>(lane_base);
    const int n = rtp->cnt[m];  // <= 63 features, 32 per half: the u8 sums below cannot overflow
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int f0 = 0; f0 < n; f0 += 16) {
      if (f0 > 0 && need > 0) {   // after 16, 32, 48 features of this modality (and see below at its end)
        if (best_so_far(acc, true) + 4 * remaining < need) { hopeless = true; break; }
      }
      remaining -= min(16, n - f0);
      const uint2* sp = s_pair + seg + f0 + half;
      uint32_t nib0 = 0, nib1 = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {   // two batches of four features per lane: eight loads in flight each
        uint2 c0[4], c1[4];
        uint32_t key[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint2 a = sp[2 * (4 * h + k)];
          key[k] = a.x;
          c0[k] = ldg64(lmm + (a.x & 0x7ffffff8u));
          c1[k] = ldg64(lmm + a.y);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool hi = (int)key[k] < 0;               // window starts in the second word of the first chunk
          const uint32_t a = hi ? c0[k].y : c0[k].x, b = hi ? c1[k].x : c0[k].y, cc = hi ? c1[k].y : c1[k].x;
          const uint32_t bits = key[k] << 2;             // funnel shifts use the low five bits: 4 * (shift & 7)
          nib0 += __funnelshift_r(a, b, bits);
          nib1 += __funnelshift_r(b, cc, bits);
          const int j = 4 * h + k;
          if (j == 2 || j == 5 || j == 7) {              // <= 3 features per nibble sum (3 * 4 < 16): 3 + 3 + 2
            acc[0] += nib0 & 0x0f0f0f0fu; acc[1] += (nib0 >> 4) & 0x0f0f0f0fu;
            acc[2] += nib1 & 0x0f0f0f0fu; acc[3] += (nib1 >> 4) & 0x0f0f0f0fu;
            nib0 = nib1 = 0;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // [OCV] similarityLocal totals are u16: widen this modality's u8 sums
      tot[j][0] += acc[j] & 0x00ff00ffu;
      tot[j][1] += (acc[j] >> 8) & 0x00ff00ffu;
    }
    if (need > 0 && !hopeless && mi + 1 < P.M) {   // between modalities
      const uint32_t none[4] = {0, 0, 0, 0};
      if (best_so_far(none, false) + 4 * remaining < need) hopeless = true;
    }
  }
  if (hopeless) return false;   // [OCV] would finish the sum and drop the candidate: sim < threshold
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // the two halves of a row meet
    tot[j][0] += __shfl_xor_sync(kFull, tot[j][0], 16);
    tot[j][1] += __shfl_xor_sync(kFull, tot[j][1], 16);
  }
  // first maximum in raster order: key = score << 8 | (255 - raster index); byte b of word j is column
  // 8 * (j / 2) + 2 * b + (j & 1)
  uint32_t best_key = 0;
  if (half == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t src = tot[j][b & 1];
        const uint32_t sc = (b & 2) ? (src >> 16) : (src & 0xffffu);
        const int col = 8 * (j >> 1) + 2 * b + (j & 1);
        best_key = max(best_key, (sc << 8) | (uint32_t)(255 - (prow * 16 + col)));
      }
  }
  best_key = __reduce_max_sync(kFull, best_key);
  const int best_score = (int)(best_key >> 8);
  int best_r = -1, best_c = -1;
  if (best_score > 0) {
    const int idx = 255 - (int)(best_key & 0xffu);
    best_r = idx >> 4; best_c = idx & 15;
  }
  const int nfl = (int)rtp->nf;
  const float sim = __fdiv_rn(__fmul_rn((float)best_score, 100.f), (float)(4 * nfl));
  x = (x / T - 8 + best_c) * T + off;
  y = (y / T - 8 + best_r) * T + off;
  score = (uint32_t)best_score; nf = rtp->nf;
  return !(sim < threshold);  // [OCV] remove_if(MatchPredicate(threshold))
}

__device__ __forceinline__ void refine_nib_warp(const RefineParams& P, const CoarseTpl* __restrict__ ctpl,
                                                const Cand* __restrict__ cand, uint32_t n_cands, const BatchCtl* ctl, uint8_t* results,
                                                size_t result_stride, uint32_t out_cap, uint32_t* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* s_addr = smem + warp * 2 * kRefineMaxFeat;  // window address words of each feature (this warp's slice)
  // consecutive candidates go to different CTAs (SMs): a handful of candidates must not queue on one SM's L1
  for (uint32_t ci = (uint32_t)warp * gridDim.x + blockIdx.x; ci < n_cands; ci += gridDim.x * kRefineWarps) {
    const Cand c = cand[ci];
    const uint32_t order = c.order;
    const uint32_t frame = c.pos >> 24, cpos = c.pos & 0x00ffffffu;
    ResultHeader* hdr = reinterpret_cast<ResultHeader*>(results + (size_t)frame * result_stride + kResultStatsBytes);
    lm_raw_match* __restrict__ out = reinterpret_cast<lm_raw_match*>(hdr + 1);
    const float threshold = P.threshold[order >> 28];
    const int cT = P.coarse_T;
    const int coff = cT / 2 + (cT % 2 - 1);
    int x = (int)(cpos % (uint32_t)P.coarse_W) * cT + coff;
    int y = (int)(cpos / (uint32_t)P.coarse_W) * cT + coff;
    uint32_t score = c.raw_nf & 0xffffu, nf = c.raw_nf >> 16;
    bool alive = true;
    for (int l = P.levels - 2; l >= 0 && alive; --l) {
      const RefineLevel& L = P.level[l];
      const RefineTpl* rtp = L.tpl + c.tglob;
      const int border = 8 * L.T;
      const int max_x = L.cols - rtp->width - border, max_y = L.rows - rtp->height - border;
      x = x * 2 + 1; y = y * 2 + 1;
      x = max(x, border); y = max(y, border);
      x = min(x, max_x); y = min(y, max_y);
      alive = L.Hh ? refine_warp_level_tiled(P, L, rtp, frame, threshold, ctl, s_addr, x, y, score, nf)
                   : refine_warp_level_flat(P, L, rtp, frame, threshold, ctl, s_addr, x, y, score, nf);
    }
    if (lane == 0) atomicAdd(&hdr->n_cands, 1u);
    if (alive && lane == 0) {
      uint32_t idx = atomicAdd(&hdr->count, 1u);
      if (idx < out_cap) {
        lm_raw_match r;
        r.order_key = order; r.coarse_pos = cpos; r.x = x; r.y = y; r.score = score; r.nf = nf;
        r.template_id = ctpl[c.tglob].template_id; r.class_index = ctpl[c.tglob].class_index;
        out[idx] = r;
      } else hdr->overflow = 1;
    }
  }
}

// Few candidates (the usual case at the reference's thresholds): a block per candidate finishes each one soonest.  Many
// candidates (loose thresholds, BASELINE config 3): a warp per candidate keeps every SM's L1 busy without block barriers.
__global__ void __launch_bounds__(kRefineWarps * 32) k_refine_nib(const RefineParams P, const CoarseTpl* __restrict__ ctpl,
                                                                 const Cand* __restrict__ cand, uint32_t cand_cap,
                                                                 BatchCtl* ctl, uint8_t* results, size_t result_stride,
                                                                 uint32_t out_cap) {
  __shared__ __align__(8) uint32_t smem[kRefineWarps * 2 * kRefineMaxFeat];
  static_assert(kRefineWarps * 2 * kRefineMaxFeat >= kRefineWarps * 16 * 4 + 2 * kRefineMaxFeat + 4, "block path fits the warp path's smem");
  cudaGridDependencySynchronize();  // programmatic dependent launch: the coarse kernel's candidates are complete from here on
  const uint32_t found = ctl->n_cands;
  const uint32_t n_cands = min(found, cand_cap);
  if (blockIdx.x == 0 && found > cand_cap) {  // candidate list truncated: every frame of the chunk is incomplete
    if ((int)threadIdx.x < ctl->ft.n_frames)
      reinterpret_cast<ResultHeader*>(results + (size_t)threadIdx.x * result_stride + kResultStatsBytes)->overflow = 1;
    if (threadIdx.x == 0) ctl->overflow = 1;
  }
  if (n_cands <= 2u * gridDim.x) refine_nib_block(P, ctpl, cand, n_cands, results, result_stride, out_cap, smem);
  else refine_nib_warp(P, ctpl, cand, n_cands, ctl, results, result_stride, out_cap, smem);
}

}  // namespace

int coarse_positions_per_pass() { return kRecWarpPos; }
int coarse_record_header_words() { return kRecHdrWords; }
int coarse_record_max_words() { return kRecMaxWords; }

// Launch configuration with programmatic stream serialization: the grid may be scheduled while its predecessor in the
// stream drains; the kernel itself waits (cudaGridDependencySynchronize) before it reads the predecessor's results.
static cudaLaunchAttribute g_pdl_attr;
static thread_local bool g_pdl_enabled = true;  // off while a stream capture records the launches (graph edges instead)
static cudaLaunchConfig_t pdl_config(int blocks, int threads, cudaStream_t s) {
  g_pdl_attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  g_pdl_attr.val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks, 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cfg.attrs = &g_pdl_attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cfg;
}

void set_programmatic_launch(bool enabled) { g_pdl_enabled = enabled; }

// Upper bound on the production coarse kernel's persistent grid (0 = every resident CTA slot).  With several frames in
// flight a smaller grid per frame lets the coarse kernels of consecutive frames share the SMs instead of queueing.
static int g_coarse_grid_limit = 0;
void set_coarse_grid_limit(int blocks) { g_coarse_grid_limit = blocks; }
// A/B switch: 0 sends requests whose tiles all have <= 63 features through the general kernel too
static int g_coarse_narrow = 1;   // 1: u8-only kernel, three CTAs per SM; 2: the same body at two CTAs per SM (A/B)
void set_coarse_narrow(int mode) { g_coarse_narrow = mode; }

template <class K>
static int resident_ctas(K kernel) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 2;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return per_sm * sms;
}

void launch_similarity_coarse(const CoarseParams& p, int max_frames, cudaStream_t s) {
  if (p.n_tiles <= 0) return;
  static const int persistent = resident_ctas(k_similarity_coarse_rec), persistent63 = resident_ctas(k_similarity_coarse_rec63),
                   persistent63_2 = resident_ctas(k_similarity_coarse_rec63_2cta);
  const int narrow = p.max_feat <= 63 ? g_coarse_narrow : 0;   // u8 sums cannot overflow: 63 * 4 < 256
  const long long blocks = ((long long)p.n_tiles * max_frames + 7) / 8;
  int grid = (int)std::min<long long>(blocks, narrow == 1 ? persistent63 : (narrow == 2 ? persistent63_2 : persistent));
  if (g_coarse_grid_limit > 0) grid = min(grid, g_coarse_grid_limit);
  CoarseParams q = p;
  if (q.dump != nullptr) q.prune = 0;
  cudaLaunchConfig_t cfg = pdl_config(grid, 256, s);
  if (narrow == 1) cudaLaunchKernelEx(&cfg, k_similarity_coarse_rec63, q);
  else if (narrow == 2) cudaLaunchKernelEx(&cfg, k_similarity_coarse_rec63_2cta, q);
  else cudaLaunchKernelEx(&cfg, k_similarity_coarse_rec, q);
}

void launch_pack_nibbles(const uint8_t* lm_bytes, size_t bytes_stride, uint8_t* lm_nibbles, size_t nib_stride, size_t n_bytes,
                         const BatchCtl* ctl, int n_frames, cudaStream_t s) {
  const size_t n16 = n_bytes / 16;
  if (n16 == 0) return;
  k_pack_nibbles<<<dim3((unsigned)((n16 + 255) / 256), n_frames), 256, 0, s>>>(lm_bytes, bytes_stride, lm_nibbles, nib_stride, n16, ctl);
}

void launch_begin_chunk(const FrameTable& ft, BatchCtl* ctl, uint8_t* results, size_t result_stride, int n_blocks,
                        cudaStream_t s) {
  k_begin_chunk<<<1, 256, 0, s>>>(ft, ctl, results, result_stride, n_blocks);
}

void launch_refine(const RefineParams& p, const CoarseTpl* ctpl, const Cand* cand, uint32_t cand_cap, BatchCtl* ctl,
                   uint8_t* results, size_t result_stride, uint32_t out_cap, cudaStream_t s) {
  cudaLaunchConfig_t cfg = pdl_config(148 * 4, kRefineWarps * 32, s);  // 4 resident CTAs per SM (64 registers)
  cudaLaunchKernelEx(&cfg, k_refine_nib, p, ctpl, cand, cand_cap, ctl, results, result_stride, out_cap);
}

}  // namespace lmk
