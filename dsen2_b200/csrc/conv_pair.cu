// DSen2 convolution layers as CTA-pair (cta_group::2) tcgen05 implicit GEMMs with RESIDENT weights.
//
// One tile = 16 rows x 8 pixels of one patch (M = 128); a CTA pair computes two horizontally adjacent
// tiles with one M = 256 MMA stream issued by the leader CTA.  Per tile and 64-channel k-block the
// producer makes ONE TMA load of the 18 x 10 pixel halo box ([y][x][64 ch], 128-byte swizzled rows);
// the nine taps are nine smem descriptors into that box (start shifted by (dy*10 + dx) rows, 8-row
// groups 10*128 B apart), so activations cross L2 -> SMEM once (x1.41 halo) instead of nine times.
// TMA out-of-bounds zero fill is the Conv2D(padding='same') border of the PATCH (DSen2Net.py:10,12,29,35):
// the patch index is its own tensor dimension, neighbours are never read.
// The layer's weights stay in SMEM for the whole launch: in pair mode each CTA holds N/2 rows of every
// [tap][k-block] slab (144 KB for 128->128), loaded once.
//
// Three layer shapes share the kernel (template parameters):
//   RES   F -> F (128: resident weights; 256: streamed through a ring), 9 taps, N = F
//         epilogues RELU | RESIDUALQ (x + 0.1*conv on the fp16 + 8 bit trunk) | RESIDUAL32 / MASK (training step)
//   HEAD  16-channel input (32-byte rows) -> F, 9 taps, A = hi | lo k-blocks, three products per tap (hi*W_hi, hi*W_lo,
//         lo*W_hi) into ONE accumulator: fp32-equivalent first layer.  (The training step uses the older form on three
//         pre-gathered horizontal taps x 16 ch, B = [W_hi ; W_lo] stacked along N and summed in the epilogue.)
//   TAIL  128 -> cout <= 16, 9 taps, A = x_hi | x_lo (4 k-blocks), B = [W_hi ; W_lo] (N = 32), epilogue adds the
//         global skip (DSen2Net.py:38,41), scales, and writes NCHW predictions or the stitched HWC canvas
//         (patches.py:374-405 ownership rule, supres.py:29).
//
// Warp roles (320 threads): warps 0-7 epilogue (two warps per TMEM lane quarter, splitting the channels), warp 8 TMA
// producer, warp 9 TMEM alloc + MMA issue (leader CTA); two TMEM accumulator buffers.
#include <stdlib.h>

#include "common.cuh"
#include "tiling.cuh"

namespace dsen2 {

// Warp roles.  The warp scheduler prefers the HIGHEST warp id among eligible warps of a sub-partition, so the
// two latency-critical single-thread roles get the highest ids: the MMA issuer must never queue behind the
// instruction-heavy epilogue warps that share its sub-partition (measured: tensor pipe 59 % -> see profiles/).
// (PairCfg::PRODUCER_WARP / MMA_WARP = the two warps after the epilogue warps)
static constexpr int kBoxH = 18;
// (An L2 prefetch of the next tile's activation boxes was measured and dropped: 559-560 ms per full tile without it
// against 565-567 ms with distance 1 on the same box.  The trunk lines the residual epilogues read ARE prefetched, one
// tile ahead.)

enum { kEpiRelu = 0, kEpiTail = 2, kEpiResidual32 = 3, kEpiMask = 4, kEpiResidualQ = 5, kEpiResidualQLast = 6,
       kEpiHeadQ = 7 };
// How a split-precision layer (operands v = hi + lo, fp16 each) lays out its products:
//   kSplitStack  B = [W_hi ; W_lo] stacked along N, A = hi then lo k-blocks; the epilogue adds the two column halves
//                (all four products; the tail, where N is tiny and the MMAs are bound by the A-operand reads)
//   kSplit3      ONE accumulator of N = CH columns: hi*W_hi + hi*W_lo + lo*W_hi (lo*W_lo ~ 2^-22 of the result is dropped);
//                three quarters of the MMAs and no accumulator-half sum in the epilogue (the first layer)
enum { kSplitNone = 0, kSplitStack = 1, kSplit3 = 2 };

struct PairParams {
  int n, H, W;
  int tiles_x, tiles_y;    // tiles per patch; tile column t covers pixels 8 * (tx0 + t) ..
  int tx0;                 // first tile column (the stitching tail skips the columns that lie inside the patch border)
  uint32_t num_tiles;
  const float* bias;
  const __half* res_hi;
  float res_scale;
  __half* out_hi;
  __half* out_lo;
  float* x32;              // fp32 trunk, tile-row-major (n, H, ceil(W/8), C/4, 8, 4): RESIDUAL32 in/out, RELU (head) optional out
  uint8_t* xq;             // low bytes of the fp16+8 trunk, tile-row-major (n, H, ceil(W/8), C/16, 8, 16): RESIDUALQ in/out, head out
  // TAIL
  const __half* skip_hi;   // prepared input (n,H,W,64): centre-tap channels 16..31 hold the network inputs
  const __half* skip_lo;
  int skip_ch0;            // first skip band inside the 16-channel group
  int skip_pitch, skip_off;   // channels per pixel of the prepared input and offset of the centre tap (64 / 16, or 16 / 0)
  int cout_real;
  float out_mul;
  float* out_f32;
  int tail_mode;           // 0: NCHW (n,cout,H,W) predictions; 1: stitched HWC canvas
  int first_patch, img_h, img_w, border, grid_ny, grid_nx;
};

template <int NTOT_, int SPLIT_, int NMAPS_, int KPM_, int NTAPS_, int KSTEPS_, int STAGES_, int EPI_, int WSTAGES_ = 0,
          int ROWB_ = 128, bool XT_ = false>
struct PairCfg {
  // XT: the epilogue moves its x_hi tiles through the async proxy -- every epilogue warp owns 4 KB of shared memory holding
  // its 4 rows x 8 pixels x 64 channels in the layout TMA boxes of the NHWC tensor produce; TMA loads bring the residual,
  // TMA stores write the result back in place, instead of LSU loads / stores through a transposing 1 KB buffer
  static constexpr bool XT = XT_;
  static constexpr int ROWB = ROWB_;            // bytes per pixel row of a k-block in smem = swizzle span (128: 64 ch, 32: 16 ch)
  static constexpr int KCH = ROWB_ / 2;         // channels per k-block
  static_assert(KSTEPS_ * 32 <= ROWB_, "the K = 16 MMAs of a k-block stay inside its row");
  static constexpr int EPI_WARPS = 8;           // two per TMEM lane quarter, half of the channels each
  static constexpr int PRODUCER_WARP = EPI_WARPS;
  static constexpr int MMA_WARP = EPI_WARPS + 1;
  static constexpr int THREADS = (EPI_WARPS + 2) * 32;
  static constexpr int WSTAGES = WSTAGES_;      // 0: the layer's weights are resident in smem; > 0: ring of streamed slabs
  static constexpr bool RESIDENT = WSTAGES_ == 0;
  static constexpr int NTOT = NTOT_;            // UMMA N over the pair
  static constexpr bool SPLIT = SPLIT_ == kSplitStack;   // B = [W_hi ; W_lo]: epilogue sums the two column halves
  static constexpr bool SPLIT3 = SPLIT_ == kSplit3;      // B slabs [tap][W_hi | W_lo], one accumulator (see kSplit3)
  static_assert(SPLIT_ != kSplit3 || (NMAPS_ == 2 && WSTAGES_ == 0), "three-product split: hi and lo A tensors, resident weights");
  static constexpr int NMAPS = NMAPS_;          // A tensors (1, or 2 = hi then lo)
  static constexpr int KPM = KPM_;              // 64-channel k-blocks per A tensor
  static constexpr int KB = NMAPS_ * KPM_;      // pipeline stages consumed per tile
  static constexpr int NTAPS = NTAPS_;          // 9 (3x3) or 3 (vertical; horizontal taps pre-gathered)
  static constexpr int KSTEPS = KSTEPS_;        // K = 16 MMAs per tap and k-block
  static constexpr int STAGES = STAGES_;
  static constexpr int EPI = EPI_;
  static constexpr int CH = SPLIT ? NTOT_ / 2 : NTOT_;    // output channels
  static constexpr int BOXW = NTAPS_ == 9 ? 10 : 8;
  static constexpr int BOX_BYTES = kBoxH * BOXW * ROWB_;
  static constexpr int STAGE_BYTES = (BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int SLAB_ROWS = NTOT_ / 2;
  static constexpr int SLAB_BYTES = SLAB_ROWS * ROWB_;
  static constexpr int WGROUPS = NTAPS_ * (SPLIT3 ? 2 : 1);   // slab groups of the weight tensor (its outermost dimension)
  static constexpr int NSLABS = WGROUPS * KPM_;
  static constexpr int W_BYTES = (WSTAGES_ == 0 ? NSLABS : WSTAGES_) * SLAB_BYTES;
  static constexpr int TMEM_COLS = (2 * NTOT_ < 32) ? 32 : 2 * NTOT_;
  static constexpr int BAR_BYTES = 2048;        // barriers + tmem pointer (first 512 B) + bias
  static constexpr int STG_BYTES = EPI_WARPS * (XT_ ? 4096 : 1024);   // per-warp x_hi tiles / transpose buffers of the epilogue
  static constexpr int SMEM_BYTES = W_BYTES + STAGES_ * STAGE_BYTES + BAR_BYTES + STG_BYTES + 1024 /*align slack*/;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
  static_assert(W_BYTES % 1024 == 0, "weight slabs must keep the stages 1024-byte aligned");
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two <= 512");
  static_assert(CH * 4 + 512 <= BAR_BYTES, "bias does not fit next to the barriers");
  static_assert((2 * STAGES_ + 6 + 2 * WSTAGES_ + 2 * EPI_WARPS) * 8 <= 512, "too many barriers");
};

// tile index -> (patch, tile row, tile column); 32-bit arithmetic (64-bit divisions cost ~100 instructions each)
struct TileXY { int b, ty, tx; };
__device__ __forceinline__ TileXY decode_tile(uint32_t tile, uint32_t tiles_x, uint32_t tiles_y) {
  const uint32_t row = tile / tiles_x;
  TileXY t;
  t.tx = (int)(tile - row * tiles_x);
  t.b = (int)(row / tiles_y);
  t.ty = (int)(row - (uint32_t)t.b * tiles_y);
  return t;
}

// the persistent loops step through the tile list with a fixed stride: advance (patch, row, column) with carries
// instead of two divisions per tile
__device__ __forceinline__ void advance_tile(TileXY& t, const TileXY& step, int tiles_x, int tiles_y) {
  t.tx += step.tx;
  const int cx = t.tx >= tiles_x;
  t.tx -= cx ? tiles_x : 0;
  t.ty += step.ty + cx;
  const int cy = t.ty >= tiles_y;
  t.ty -= cy ? tiles_y : 0;
  t.b += step.b + cy;
}

__device__ __forceinline__ int stitch_tile_of(int y, int size, int S, int n) {
  return (size % S != 0 && y >= size - S) ? n - 1 : y / S;
}

// ---- epilogue I/O --------------------------------------------------------------------------
// After tcgen05.ld (32x32b) thread t of a warp owns pixel t of the warp's 32 tile rows and 64 consecutive
// channels of it (128 B).  Written straight to NHWC memory that is 32 different cache lines per
// instruction (L1TEX wavefront bound).  Instead each warp transposes through a 1 KB smem buffer, 8 pixels
// (= the 8 pixels of one image row of the tile) at a time, so that 8 consecutive lanes move the 128
// contiguous bytes of one pixel: 4 full lines per instruction.
__device__ __forceinline__ uint32_t stg_off(int r, int chunk) { return (uint32_t)(r * 128 + ((chunk ^ r) & 7) * 16); }

struct EpiGeom {
  long long base;     // element offset of (image row of round 0, pixel 0 of the tile, first channel of this warp)
  long long row_pitch;   // elements per image row
  int ch;             // channels per pixel (pixel pitch)
  int rows_valid;     // rounds (image rows) of this warp that are inside the patch
  int px_valid;       // pixels of the tile row that are inside the patch
};

// registers (thread = pixel, v[q] = channels 8q..8q+7) -> global, coalesced.  `stg` is a shared-space address.
__device__ __forceinline__ void staged_store(uint32_t stg, const uint4 (&v)[8], __half* out, const EpiGeom& g,
                                             int lane) {
  const int k_own = lane >> 3, r_own = lane & 7, rr = lane >> 3, chunk = lane & 7;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __syncwarp();
    if (k_own == k) {
#pragma unroll
      for (int q = 0; q < 8; ++q) sts128(stg + stg_off(r_own, q), v[q]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 4 * j + rr;
      const uint4 val = lds128(stg + stg_off(r, chunk));
      if (k < g.rows_valid && r < g.px_valid)
        *reinterpret_cast<uint4*>(out + g.base + k * g.row_pitch + (long long)r * g.ch + chunk * 8) = val;
    }
  }
}

// global -> registers (lane-coalesced layout: gl[k*2+j] = 16 B chunk `lane&7` of pixel 4j + lane>>3 of round k)
// (plain loads: the RESIDUAL epilogue updates the trunk in place, the tensor is not read-only for the kernel)
__device__ __forceinline__ void coalesced_load(uint4 (&gl)[8], const __half* in, const EpiGeom& g, int lane) {
  const int rr = lane >> 3, chunk = lane & 7;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 4 * j + rr;
      gl[k * 2 + j] = (k < g.rows_valid && r < g.px_valid)
                          ? *reinterpret_cast<const uint4*>(in + g.base + k * g.row_pitch + (long long)r * g.ch + chunk * 8)
                          : make_uint4(0, 0, 0, 0);
    }
}

// one 128-byte line per lane: lane (k = lane>>3, r = lane&7) owns pixel r of round k
__device__ __forceinline__ void prefetch_rows(const __half* in, const EpiGeom& g, int lane) {
  const int k = lane >> 3, r = lane & 7;
  if (k < g.rows_valid && r < g.px_valid) prefetch_l2(in + g.base + k * g.row_pitch + (long long)r * g.ch);
}

// lane-coalesced registers -> thread = pixel registers
__device__ __forceinline__ void staged_gather(uint32_t stg, const uint4 (&gl)[8], uint4 (&v)[8], int lane) {
  const int k_own = lane >> 3, r_own = lane & 7, rr = lane >> 3, chunk = lane & 7;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) sts128(stg + stg_off(4 * j + rr, chunk), gl[k * 2 + j]);
    __syncwarp();
    if (k_own == k) {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = lds128(stg + stg_off(r_own, q));
    }
  }
}

// ---- fp16 + 8 bit residual trunk ("Q" trunk) ---------------------------------------------------
// The trunk value x of a resblock chain is kept as the fp16 tensor the next convolution's TMA reads anyway (h) plus ONE
// extra byte per element: with s = bits(x) + 0x1010, h = fp16 of s truncated to 10 mantissa bits (x rounded to nearest
// fp16, ties away from zero) and lo = bits 12..5 of s, stored biased by -128 as a signed byte.  Then
//   bits(x) ~ bits(float(h)) + (lo << 5)          (|error| <= 16 fp32 ulps: 19 significant bits, unbiased)
// which is pure integer arithmetic on both sides.  For |x| < 2^-14 h is an fp16 subnormal (its truncation point moves up):
// the code then only keeps x to an absolute 2^-24, and lo is forced to 0 where h is zero.
// The fp32 trunk costs 1536 B/pixel per resblock of HBM traffic (read t, read + write fp32, write the fp16 copy);
// this one 1024 (read t, read + write fp16, read + write the bytes).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t cvt_rz_f16x2(uint32_t hi_bits, uint32_t lo_bits) {
  uint32_t d;
  asm("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_bits)), "f"(__uint_as_float(lo_bits)));
  return d;
}
// byte k of w, sign-extended (prmt replicates the sign of the selected byte when bit 3 of a selector nibble is set)
template <int K>
__device__ __forceinline__ int sext_byte(uint32_t w) {
  return (int)prmt(w, 0u, 0x8880u + 0x1111u * K);
}
__device__ __forceinline__ float q_decode(float hf, int lo) {
  return __uint_as_float(__float_as_uint(hf) + (uint32_t)(lo << 5));
}
// four trunk values -> two packed fp16 pairs + one word of low bytes
__device__ __forceinline__ void q_encode4(float x0, float x1, float x2, float x3, uint32_t& h01, uint32_t& h23, uint32_t& lo) {
  const uint32_t s0 = __float_as_uint(x0) + 0x1010u, s1 = __float_as_uint(x1) + 0x1010u;
  const uint32_t s2 = __float_as_uint(x2) + 0x1010u, s3 = __float_as_uint(x3) + 0x1010u;
  h01 = cvt_rz_f16x2(s1, s0);
  h23 = cvt_rz_f16x2(s3, s2);
  const uint32_t p01 = prmt(s0, s1, 0x5410u) >> 5, p23 = prmt(s2, s3, 0x5410u) >> 5;   // bits 12..5 land in bytes 0 and 2
  // |x| < 2^-24 gives h = +-0: there is no binade for the byte to refine (a negative one would borrow through the sign
  // bit on decode), so it must be 0.  0xffff per zero half -> 0xff per byte.
  const uint32_t z01 = __heq2_mask(*reinterpret_cast<const __half2*>(&h01), __half2half2(__ushort_as_half(0)));
  const uint32_t z23 = __heq2_mask(*reinterpret_cast<const __half2*>(&h23), __half2half2(__ushort_as_half(0)));
  lo = (prmt(p01, p23, 0x6420u) ^ 0x80808080u) & ~prmt(z01, z23, 0x6420u);
}

template <class Cfg>
__device__ __forceinline__ EpiGeom epi_geom(const PairParams& p, const TileXY& t, int wq, int chan0) {
  const int yw = t.ty * 16 + wq * 4;            // first image row of this warp's 4 rounds
  EpiGeom g;
  g.ch = Cfg::CH;
  g.row_pitch = (long long)p.W * Cfg::CH;
  g.base = (((long long)t.b * p.H + yw) * p.W + t.tx * 8) * Cfg::CH + chan0;
  g.rows_valid = (t.b < p.n) ? max(0, min(4, p.H - yw)) : 0;
  g.px_valid = max(0, min(8, p.W - t.tx * 8));
  return g;
}

// ---- resblock output on the fp16 + 8 bit trunk: one pass over CPT channels of this thread's pixel ----------------
// x <- x + scale * (conv + bias) (DSen2Net.py:13,15) on the trunk held as x_hi (NHWC fp16, updated in place: this
// kernel's TMA reads t, never x_hi) + one byte per element (see q_encode4).  The LAST block of the chain writes
// x_hi = fp16(x) and x_lo = fp16(x - x_hi) instead: the tail's split operand (and nothing reads the bytes again).
// `full` / `empty`: accumulator barriers to wait on before / arrive on after the TMEM reads of this pass (either may
// be null when the pass is not the first / last one that touches the accumulator).
template <class Cfg, int CPT>
__device__ __forceinline__ void q_epilogue_pass(const PairParams& p, const TileXY& tc, const TileXY& tn, bool prefetch, int wq,
                                                int chan0, int lane, int row, bool valid, uint32_t stg, uint32_t taddr,
                                                uint32_t s_bias_addr, uint64_t* full, uint32_t full_phase, uint64_t* empty) {
  constexpr bool LAST = Cfg::EPI == kEpiResidualQLast;
  static_assert(CPT == 64 && !Cfg::SPLIT, "Q-trunk epilogue: 64 channels per thread");
  const int y = tc.ty * 16 + (row >> 3);
  const EpiGeom g = epi_geom<Cfg>(p, tc, wq, chan0);
  uint4 vh[CPT / 8];
  uint4 lq[CPT / 16];
  // low bytes: (n, H, W/8, C/16, 8 px, 16 ch) -- thread = pixel reads 16 B, a tile row of one chunk is one 128 B line
  uint8_t* const qp = p.xq + ((((long long)tc.b * p.H + y) * p.tiles_x + tc.tx) * (Cfg::CH / 16) + chan0 / 16) * 128 +
                      (row & 7) * 16;
  {
    uint4 gh[CPT / 8];
    coalesced_load(gh, p.out_hi, g, lane);
    if (valid) {
#pragma unroll
      for (int q = 0; q < CPT / 16; ++q) lq[q] = *reinterpret_cast<const uint4*>(qp + q * 128);
    } else {
#pragma unroll
      for (int q = 0; q < CPT / 16; ++q) lq[q] = make_uint4(0, 0, 0, 0);
    }
    if (prefetch) {                                          // a later tile's trunk lines -> L2
      prefetch_rows(p.out_hi, epi_geom<Cfg>(p, tn, wq, chan0), lane);
      constexpr int QL = CPT / 16;                           // byte-chunk lines per tile row of this pass
      const int ny = tn.ty * 16 + wq * 4 + lane / QL;        // 4 rows x QL lines per warp
      if (lane < 4 * QL && tn.b < p.n && ny < p.H)
        prefetch_l2(p.xq + ((((long long)tn.b * p.H + ny) * p.tiles_x + tn.tx) * (Cfg::CH / 16) + chan0 / 16 + (lane % QL)) * 128);
    }
    staged_gather(stg, gh, vh, lane);
  }
  uint4 vl[LAST ? CPT / 8 : 1];
  if (full != nullptr) {
    mbar_wait(full, full_phase);
    tc_fence_after();
  }
#pragma unroll
  for (int chunk = 0; chunk < CPT / 32; ++chunk) {
    const int c0 = chan0 + chunk * 32;
    uint32_t r[32];
    tmem_ld_32x32(taddr + c0, r);
    tmem_ld_wait();
    uint32_t* hw = reinterpret_cast<uint32_t*>(vh) + chunk * 16;
    uint32_t* qw = reinterpret_cast<uint32_t*>(lq) + chunk * 8;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bq = lds_f4(s_bias_addr + (uint32_t)(c0 + j) * 4);   // broadcast LDS.128
      const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&hw[j >> 1]));
      const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&hw[(j >> 1) + 1]));
      const uint32_t w = qw[j >> 2];
      const float x0 = fmaf(__uint_as_float(r[j]) + bq.x, p.res_scale, q_decode(f01.x, sext_byte<0>(w)));
      const float x1 = fmaf(__uint_as_float(r[j + 1]) + bq.y, p.res_scale, q_decode(f01.y, sext_byte<1>(w)));
      const float x2 = fmaf(__uint_as_float(r[j + 2]) + bq.z, p.res_scale, q_decode(f23.x, sext_byte<2>(w)));
      const float x3 = fmaf(__uint_as_float(r[j + 3]) + bq.w, p.res_scale, q_decode(f23.y, sext_byte<3>(w)));
      if constexpr (LAST) {
        const __half2 h0 = __floats2half2_rn(x0, x1), h1 = __floats2half2_rn(x2, x3);
        const float2 g0 = __half22float2(h0), g1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(x0 - g0.x, x1 - g0.y), l1 = __floats2half2_rn(x2 - g1.x, x3 - g1.y);
        uint32_t* lw = reinterpret_cast<uint32_t*>(vl) + chunk * 16;
        hw[j >> 1] = *reinterpret_cast<const uint32_t*>(&h0);
        hw[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        lw[j >> 1] = *reinterpret_cast<const uint32_t*>(&l0);
        lw[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&l1);
      } else {
        q_encode4(x0, x1, x2, x3, hw[j >> 1], hw[(j >> 1) + 1], qw[j >> 2]);
      }
    }
  }
  if (empty != nullptr) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_leader(empty);
  }
  if constexpr (!LAST) {
    if (valid) {
#pragma unroll
      for (int q = 0; q < CPT / 16; ++q) *reinterpret_cast<uint4*>(qp + q * 128) = lq[q];
    }
  }
  staged_store(stg, vh, p.out_hi, g, lane);
  if constexpr (LAST) staged_store(stg, vl, p.out_lo, g, lane);
}

// ---- x_hi tiles through the async proxy (PairCfg::XT) ------------------------------------------------------------
// An epilogue warp's 32 pixels (4 image rows x 8 pixels of the tile) x 64 channels are one TMA box {64 ch, 8 px, 4 rows, 1}
// of the NHWC fp16 tensor = 4 KB of shared memory in the 128-byte-swizzled layout: pixel r (= lane) at r * 128, its 16-byte
// chunk q at ((q ^ (r & 7)) * 16) -- a quarter warp touches eight different chunks, so thread = pixel LDS.128 / STS.128
// are conflict-free.  Out-of-patch rows / pixels are zero-filled by the load and clipped by the store.
__device__ __forceinline__ uint32_t xt_off(int lane, int q) { return (uint32_t)(lane * 128 + ((q ^ (lane & 7)) << 4)); }

// registers (thread = pixel, 64 channels) -> the warp's tile -> global, one TMA store issued by lane 0
__device__ __forceinline__ void xt_store(uint32_t xt, const uint4 (&v)[8], const CUtensorMap* m, int c0, int x0, int y0, int b,
                                         int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q) sts128(xt + xt_off(lane, q), v[q]);
  fence_proxy_async_smem();                         // generic-proxy writes -> visible to the TMA engine
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)), "r"(xt), "r"(c0), "r"(x0), "r"(y0), "r"(b)
                 : "memory");
    tma_store_commit();
  }
}
// the TMA engine has read the warp's tile out of shared memory (all but the N most recent stores): it may be overwritten
template <int N>
__device__ __forceinline__ void xt_release(int lane) {
  if (lane == 0) tma_store_wait_read<N>();
  __syncwarp();
}

// Half tiles: 32 pixels x 32 channels = TMA box {32 ch, 8 px, 4 rows, 1}, 2 KB, 64-byte rows in the 64-byte-swizzled layout
// (16-byte chunk q of pixel r at r * 64 + ((q ^ ((r >> 1) & 3)) * 16): conflict-free thread = pixel access again).
__device__ __forceinline__ uint32_t xh_off(int lane, int q) { return (uint32_t)(lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)); }
__device__ __forceinline__ void xh_fetch(uint32_t xh, const CUtensorMap* m, uint64_t* bar, int c0, int x0, int y0, int b) {
  mbar_expect_tx(bar, 2048);
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(xh), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(x0), "r"(y0), "r"(b)
      : "memory");
}
__device__ __forceinline__ void xh_store(uint32_t xh, const uint4 (&v)[4], const CUtensorMap* m, int c0, int x0, int y0, int b,
                                         int lane) {
#pragma unroll
  for (int q = 0; q < 4; ++q) sts128(xh + xh_off(lane, q), v[q]);
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)), "r"(xh), "r"(c0), "r"(x0), "r"(y0), "r"(b)
                 : "memory");
    tma_store_commit();
  }
}

// RESIDUALQ / RESIDUALQ-last epilogue of one tile for one warp (128 features: a thread owns 64 channels), x_hi in and out
// through the async proxy.  Same arithmetic as q_epilogue_pass.  The warp's 4 KB of shared memory are two half tiles A, B
// (channels 0..31 / 32..63 of its 64 = the two 32-column TMEM reads); each is filled by a TMA load of the residual, updated
// in place by its threads and written back by a TMA store (in place in global memory too: this kernel's activation TMA
// reads t, never x_hi).  The halves ping-pong so that nobody waits for the TMA engine: while half A of tile i is being
// updated, B's store of tile i-1 is awaited and B(i) fetched; while B(i) is updated, A's store is awaited and A(i+1)
// fetched.  (With ONE 4 KB tile per warp -- store, wait, load, wait -- 40 % of the epilogue warps' samples sat on those
// two waits, profiles/r02_whole_batch_xt1_ncu.txt.)
template <class Cfg>
__device__ __forceinline__ void q_epilogue_xt(const PairParams& p, const CUtensorMap* tm_xhi, const CUtensorMap* tm_xlo,
                                              const TileXY& tc, const TileXY& tn, bool first, bool has_next, int wq, int chan0,
                                              int lane, int row, bool valid, uint32_t xt, uint64_t* xbar, uint32_t xphase,
                                              uint32_t taddr, uint32_t s_bias_addr, uint64_t* full, uint32_t full_phase,
                                              uint64_t* empty) {
  constexpr bool LAST = Cfg::EPI == kEpiResidualQLast;
  static_assert(Cfg::CH == 128 && !Cfg::SPLIT, "XT Q-trunk epilogue: 128 features, 64 channels per thread");
  const int y = tc.ty * 16 + (row >> 3);
  const int gx = tc.tx * 8, gy = tc.ty * 16 + wq * 4;
  uint4 lq[4];
  uint8_t* const qp = p.xq + ((((long long)tc.b * p.H + y) * p.tiles_x + tc.tx) * (Cfg::CH / 16) + chan0 / 16) * 128 +
                      (row & 7) * 16;
  if (valid) {
#pragma unroll
    for (int q = 0; q < 4; ++q) lq[q] = *reinterpret_cast<const uint4*>(qp + q * 128);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) lq[q] = make_uint4(0, 0, 0, 0);
  }
  if (has_next) {                                            // the next tile's byte lines -> L2 (4 rows x 4 lines per warp)
    const int ny = tn.ty * 16 + wq * 4 + lane / 4;
    if (lane < 16 && tn.b < p.n && ny < p.H)
      prefetch_l2(p.xq + ((((long long)tn.b * p.H + ny) * p.tiles_x + tn.tx) * (Cfg::CH / 16) + chan0 / 16 + (lane % 4)) * 128);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t xh = xt + (uint32_t)(h * 2048);
    const int c0 = chan0 + h * 32;
    uint4 vh[4], vl[LAST ? 4 : 1];
    mbar_wait(xbar + h, xphase);                             // this half of the residual has landed
#pragma unroll
    for (int q = 0; q < 4; ++q) vh[q] = lds128(xh + xh_off(lane, q));
    if (h == 0) {
      mbar_wait(full, full_phase);
      tc_fence_after();
    }
    uint32_t r[32];
    tmem_ld_32x32(taddr + c0, r);
    tmem_ld_wait();
    uint32_t* hw = reinterpret_cast<uint32_t*>(vh);
    uint32_t* qw = reinterpret_cast<uint32_t*>(lq) + h * 8;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j == 16 && lane == 0) {
        // Half-way through this half's arithmetic the OTHER half's last store (issued ~half a tile period ago) has been
        // read out of shared memory: fetch its next residual -- B of this tile while A is being updated, A of the next
        // tile while B is.  (The first tile's B came with the prologue.)
        if (h == 0 ? !first : has_next) {
          tma_store_wait_read<0>();
          if (h == 0) xh_fetch(xt + 2048, tm_xhi, xbar + 1, chan0 + 32, gx, gy, tc.b);
          else xh_fetch(xt, tm_xhi, xbar, chan0, tn.tx * 8, tn.ty * 16 + wq * 4, tn.b);
        }
      }
      const float4 bq = lds_f4(s_bias_addr + (uint32_t)(c0 + j) * 4);   // broadcast LDS.128
      const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&hw[j >> 1]));
      const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&hw[(j >> 1) + 1]));
      const uint32_t w = qw[j >> 2];
      const float x0 = fmaf(__uint_as_float(r[j]) + bq.x, p.res_scale, q_decode(f01.x, sext_byte<0>(w)));
      const float x1 = fmaf(__uint_as_float(r[j + 1]) + bq.y, p.res_scale, q_decode(f01.y, sext_byte<1>(w)));
      const float x2 = fmaf(__uint_as_float(r[j + 2]) + bq.z, p.res_scale, q_decode(f23.x, sext_byte<2>(w)));
      const float x3 = fmaf(__uint_as_float(r[j + 3]) + bq.w, p.res_scale, q_decode(f23.y, sext_byte<3>(w)));
      if constexpr (LAST) {
        const __half2 h0 = __floats2half2_rn(x0, x1), h1 = __floats2half2_rn(x2, x3);
        const float2 g0 = __half22float2(h0), g1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn(x0 - g0.x, x1 - g0.y), l1 = __floats2half2_rn(x2 - g1.x, x3 - g1.y);
        uint32_t* lw = reinterpret_cast<uint32_t*>(vl);
        hw[j >> 1] = *reinterpret_cast<const uint32_t*>(&h0);
        hw[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        lw[j >> 1] = *reinterpret_cast<const uint32_t*>(&l0);
        lw[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&l1);
      } else {
        q_encode4(x0, x1, x2, x3, hw[j >> 1], hw[(j >> 1) + 1], qw[j >> 2]);
      }
    }
    if (h == 1) {                                            // the accumulator has been read: hand the TMEM buffer back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(empty);
    }
    xh_store(xh, vh, tm_xhi, c0, gx, gy, tc.b, lane);
    if constexpr (LAST) {                                    // x_lo = fp16(x - x_hi): the last layer's split operand
      xt_release<0>(lane);
      xh_store(xh, vl, tm_xlo, c0, gx, gy, tc.b, lane);
    }
  }
  if constexpr (!LAST) {
    if (valid) {
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(qp + q * 128) = lq[q];
    }
  }
}

template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg::THREADS, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
                 const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x0,
                 const __grid_constant__ CUtensorMap tm_x1, const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + Cfg::W_BYTES;
  uint8_t* bar_base = s_a + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);      // leader's copies are the live ones
  uint64_t* empty = full + Cfg::STAGES;                        // per CTA (multicast commit)
  uint64_t* wfull = empty + Cfg::STAGES;                       // leader
  uint64_t* tmem_full = wfull + 1;                             // per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;                        // leader; 2 * EPI_WARPS arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* wr_full = tmem_empty + 3;                          // streamed-weight ring: leader's copies are live
  uint64_t* wr_empty = wr_full + Cfg::WSTAGES;                 // per CTA (multicast commit)
  uint64_t* xfull = wr_empty + Cfg::WSTAGES;                   // per epilogue warp and half: its x_hi half tile has landed (XT)
  float* s_bias = reinterpret_cast<float*>(bar_base + 512);
  const uint32_t s_bias_addr = smem_u32(s_bias);
  uint8_t* s_stg = bar_base + Cfg::BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = blockIdx.x >> 1;
  const uint32_t npairs = gridDim.x >> 1;
  const uint32_t pair_tiles = (p.num_tiles + 1) >> 1;

  if (warp == Cfg::PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tm_a0);
    if (Cfg::NMAPS == 2) tma_prefetch_desc(&tm_a1);
    tma_prefetch_desc(&tm_w);
    if (Cfg::XT) {
      tma_prefetch_desc(&tm_x0);
      if (Cfg::EPI == kEpiResidualQLast) tma_prefetch_desc(&tm_x1);
      for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(&xfull[i], 1);
    }
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(wfull, 1);
    for (int i = 0; i < Cfg::WSTAGES; ++i) {
      mbar_init(&wr_full[i], 1);
      mbar_init(&wr_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * Cfg::EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == Cfg::MMA_WARP) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_ptr);
  for (int i = threadIdx.x; i < Cfg::CH; i += Cfg::THREADS) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == Cfg::PRODUCER_WARP) {
    // ================================ TMA producer (both CTAs) ================================
    if (elect_one()) {
      if constexpr (Cfg::RESIDENT) {
        if (rank == 0) mbar_expect_tx(wfull, 2 * Cfg::W_BYTES);
        for (int s = 0; s < Cfg::NSLABS; ++s)
          tma_load_3d_pair(s_w + s * Cfg::SLAB_BYTES, &tm_w, wfull, (s % Cfg::KPM) * Cfg::KCH, (int)rank * Cfg::SLAB_ROWS,
                           s / Cfg::KPM);
      }
      int stage = 0, ws = 0;
      uint32_t phase = 0, wphase = 0;
      for (uint32_t pt = pair; pt < pair_tiles; pt += npairs) {
        // tile may equal num_tiles (odd count): then b == n and the whole box is zero fill
        const TileXY t = decode_tile(2 * pt + rank, p.tiles_x, p.tiles_y);
        const int b = t.b;
        const int bx = (t.tx + p.tx0) * 8 - (Cfg::NTAPS == 9 ? 1 : 0), by = t.ty * 16 - 1;
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::BOX_BYTES);
          const CUtensorMap* m = (Cfg::NMAPS == 2 && kb >= Cfg::KPM) ? &tm_a1 : &tm_a0;
          tma_load_4d_pair(s_a + stage * Cfg::STAGE_BYTES, m, &full[stage], (kb % Cfg::KPM) * Cfg::KCH, bx, by, b);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          if constexpr (!Cfg::RESIDENT) {            // this k-block's nine weight slabs, in the order the MMAs use them
#pragma unroll 1
            for (int tap = 0; tap < Cfg::NTAPS; ++tap) {
              mbar_wait(&wr_empty[ws], wphase ^ 1);
              if (rank == 0) mbar_expect_tx(&wr_full[ws], 2 * Cfg::SLAB_BYTES);
              tma_load_3d_pair(s_w + ws * Cfg::SLAB_BYTES, &tm_w, &wr_full[ws], (kb % Cfg::KPM) * Cfg::KCH,
                               (int)rank * Cfg::SLAB_ROWS, tap);
              if (++ws == Cfg::WSTAGES) { ws = 0; wphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == Cfg::MMA_WARP) {
    // ================================ MMA issuer (leader CTA) ==================================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(256, Cfg::NTOT);
      if constexpr (Cfg::RESIDENT) {
        mbar_wait(wfull, 0);
        tc_fence_after();
      }
      int stage = 0, ws = 0;
      uint32_t phase = 0, wphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (uint32_t pt = pair; pt < pair_tiles; pt += npairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::NTOT);
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(s_a + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = smem_u32(s_w) + (uint32_t)((kb % Cfg::KPM) * Cfg::SLAB_BYTES);
#pragma unroll
          for (int tap = 0; tap < Cfg::NTAPS; ++tap) {
            const int dy = Cfg::NTAPS == 9 ? tap / 3 : tap;
            const int dx = Cfg::NTAPS == 9 ? tap % 3 : 0;
            const uint32_t a0 = sa + (uint32_t)((dy * Cfg::BOXW + dx) * Cfg::ROWB);
            if constexpr (Cfg::SPLIT3) {
              // hi k-blocks meet W_hi and W_lo, lo k-blocks W_hi only; everything lands in the one accumulator
              const int nb = kb < Cfg::KPM ? 2 : 1;
              for (int hl = 0; hl < nb; ++hl) {
                const uint32_t b3 = sb + (uint32_t)((tap * 2 + hl) * Cfg::KPM * Cfg::SLAB_BYTES);
#pragma unroll
                for (int k = 0; k < Cfg::KSTEPS; ++k)
                  umma_f16_ss_pair(d_tmem, umma_desc_k_sbo<Cfg::ROWB>(a0 + k * 32, Cfg::BOXW * Cfg::ROWB),
                                   umma_desc_k_sbo<Cfg::ROWB>(b3 + k * 32, 8 * Cfg::ROWB), idesc,
                                   (uint32_t)((kb | tap | hl | k) != 0));
              }
              continue;
            }
            uint32_t b0 = sb + (uint32_t)(tap * Cfg::KPM * Cfg::SLAB_BYTES);
            if constexpr (!Cfg::RESIDENT) {
              mbar_wait(&wr_full[ws], wphase);
              tc_fence_after();
              b0 = smem_u32(s_w) + (uint32_t)(ws * Cfg::SLAB_BYTES);
            }
#pragma unroll
            for (int k = 0; k < Cfg::KSTEPS; ++k)
              umma_f16_ss_pair(d_tmem, umma_desc_k_sbo<Cfg::ROWB>(a0 + k * 32, Cfg::BOXW * Cfg::ROWB),
                               umma_desc_k_sbo<Cfg::ROWB>(b0 + k * 32, 8 * Cfg::ROWB), idesc,
                               (uint32_t)((kb | tap | k) != 0));
            if constexpr (!Cfg::RESIDENT) {
              umma_commit_pair(&wr_empty[ws]);                   // frees the weight slot in both CTAs
              if (++ws == Cfg::WSTAGES) { ws = 0; wphase ^= 1; }
            }
          }
          umma_commit_pair(&empty[stage]);                       // frees the slot in both CTAs
          if (kb == Cfg::KB - 1) umma_commit_pair(&tmem_full[acc]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (both CTAs) =====================================
    const int wq = warp & 3;                       // TMEM lane quarter of this warp
    const int half = warp >> 2;                    // which half of the channels (epilogue warps are 0..7)
    const int row = wq * 32 + lane;                // tile row = group * 8 + pixel
    int acc = 0;
    uint32_t acc_phase = 0;
    // current tile and this CTA's next one (whose epilogue operands are prefetched into L2 / loaded by TMA)
    const TileXY tstep = decode_tile(2 * npairs, p.tiles_x, p.tiles_y);
    TileXY tci = decode_tile(2 * pair + rank, p.tiles_x, p.tiles_y);
    TileXY tni = decode_tile(2 * (pair + npairs) + rank, p.tiles_x, p.tiles_y);     // this CTA's next tile
    const uint32_t xt = smem_u32(s_stg) + (uint32_t)(warp * 4096);                  // XT: this warp's x_hi tile
    uint32_t xphase = 0;
    if constexpr (Cfg::XT && (Cfg::EPI == kEpiResidualQ || Cfg::EPI == kEpiResidualQLast)) {
      if (lane == 0 && pair < pair_tiles) {        // both halves of the first tile's residual
        const int c0 = half * (Cfg::CH / 2), gx = (tci.tx + p.tx0) * 8, gy = tci.ty * 16 + wq * 4;
        xh_fetch(xt, &tm_x0, &xfull[2 * warp], c0, gx, gy, tci.b);
        xh_fetch(xt + 2048, &tm_x0, &xfull[2 * warp + 1], c0 + 32, gx, gy, tci.b);
      }
      __syncwarp();
    }
    for (uint32_t pt = pair; pt < pair_tiles;
         pt += npairs, advance_tile(tci, tstep, p.tiles_x, p.tiles_y), advance_tile(tni, tstep, p.tiles_x, p.tiles_y)) {
      const TileXY tc = {tci.b, tci.ty, tci.tx + p.tx0}, tn = {tni.b, tni.ty, tni.tx + p.tx0};   // in patch tile coordinates
      const int b = tc.b;
      const int y = tc.ty * 16 + (row >> 3);
      const int x = tc.tx * 8 + (row & 7);
      const bool valid = (b < p.n) && (y < p.H) && (x < p.W);
      const long long pix = ((long long)b * p.H + y) * p.W + x;
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * Cfg::NTOT);
      if constexpr (Cfg::EPI == kEpiTail) {
        // ------------------------------------------------------------------ tail: skip + scale + stitch
        float skip[16];
        bool write = valid && half == 0;
        long long obase = 0, ostride = 0;
        if (write) {
          if (p.tail_mode == 0) {
            ostride = (long long)p.H * p.W;
            obase = (long long)b * p.cout_real * ostride + (long long)y * p.W + x;
          } else {
            const int S = p.H - 2 * p.border;
            const int patch = p.first_patch + b;
            const int pty = patch / p.grid_nx, ptx = patch - pty * p.grid_nx;
            const int gy = min(pty * S, p.img_h - S) + y - p.border;
            const int gx = min(ptx * S, p.img_w - S) + x - p.border;
            write = y >= p.border && y < p.H - p.border && x >= p.border && x < p.W - p.border &&
                    stitch_tile_of(gy, p.img_h, S, p.grid_ny) == pty && stitch_tile_of(gx, p.img_w, S, p.grid_nx) == ptx;
            ostride = 1;
            obase = ((long long)gy * p.img_w + gx) * p.cout_real;
          }
        }
        if (write) {
          const __half* sh = p.skip_hi + pix * p.skip_pitch + p.skip_off + p.skip_ch0;
          const __half* sl = p.skip_lo + pix * p.skip_pitch + p.skip_off + p.skip_ch0;
#pragma unroll
          for (int c = 0; c < 16; ++c)
            skip[c] = (c < p.cout_real) ? __half2float(__ldg(sh + c)) + __half2float(__ldg(sl + c)) : 0.f;
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (half == 0) {
          uint32_t r[32];
          tmem_ld_32x32(taddr, r);
          tmem_ld_wait();
          if (write) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < p.cout_real) {
                const float v = (__uint_as_float(r[c]) + __uint_as_float(r[16 + c])) + s_bias[c];
                p.out_f32[obase + c * ostride] = (v + skip[c]) * p.out_mul;   // Add (DSen2Net.py:38), then x SCALE
              }
          }
        }
      } else if constexpr (Cfg::EPI == kEpiResidual32) {
        // ------------------------------------------------------------------ resblock output, fp32 trunk
        // x <- x + scale * (conv + bias) on the fp32 trunk (DSen2Net.py:13,15), which lives in a tile-row-major
        // layout (n, H, W/8, C/4, 8, 4) so that thread = pixel access is coalesced: the 8 pixels of a tile row are
        // 8 x 16 contiguous bytes for every 4-channel chunk, and the 32 chunks of that tile row are 4 KB contiguous
        // (DRAM page locality).  The fp16 NHWC copy the next convolution's TMA
        // reads goes through the staged (transposing) store; x_lo is only produced when the tail needs it.
        constexpr int CPT = Cfg::CH / 2;            // a thread owns CH / 2 channels of its pixel: one pass of 64 (128 features) or two
        static_assert(CPT % 64 == 0 && !Cfg::SPLIT, "fp32-trunk epilogue: 64 channels per thread and pass");
        const uint32_t stg = smem_u32(s_stg) + (uint32_t)(warp * 1024);
        // tile-row-major trunk: (n, H, W/8, C/4, 8 px, 4 ch) -- the C/4 chunks of a tile row are contiguous
        constexpr long long cpitch = 32;                                // floats between 4-channel chunks
#pragma unroll 1
        for (int sc = 0; sc < CPT / 64; ++sc) {
          const int cb = half * CPT + sc * 64;      // first channel of this pass
          const EpiGeom g = epi_geom<Cfg>(p, tc, wq, cb);
          float* xp = p.x32 + ((((long long)b * p.H + y) * p.tiles_x + tc.tx) * (Cfg::CH / 4) + cb / 4) * cpitch + (row & 7) * 4;
          float4 xr[16];
          if (valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) xr[q] = *reinterpret_cast<const float4*>(xp + q * cpitch);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) xr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (pt + npairs < pair_tiles) {   // the next tile's trunk lines -> L2 (64 lines per warp and pass, 2 per lane)
            const int ny = tn.ty * 16 + wq * 4 + (lane >> 3);
            if (tn.b < p.n && ny < p.H) {
              const float* np = p.x32 + ((((long long)tn.b * p.H + ny) * p.tiles_x + tn.tx) * (Cfg::CH / 4) + cb / 4 +
                                         (lane & 7) * 2) * cpitch;
              prefetch_l2(np);
              prefetch_l2(np + cpitch);
            }
          }
          if (sc == 0) {
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
          }
          uint4 vh[8];
#pragma unroll
          for (int chunk = 0; chunk < 2; ++chunk) {
            const int c0 = cb + chunk * 32;
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0, r);
            tmem_ld_wait();
            uint32_t* hw = reinterpret_cast<uint32_t*>(vh) + chunk * 16;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bq = lds_f4(s_bias_addr + (uint32_t)(c0 + j) * 4);   // broadcast LDS.128
              float4& xv = xr[chunk * 8 + (j >> 2)];
              xv.x = fmaf(__uint_as_float(r[j]) + bq.x, p.res_scale, xv.x);
              xv.y = fmaf(__uint_as_float(r[j + 1]) + bq.y, p.res_scale, xv.y);
              xv.z = fmaf(__uint_as_float(r[j + 2]) + bq.z, p.res_scale, xv.z);
              xv.w = fmaf(__uint_as_float(r[j + 3]) + bq.w, p.res_scale, xv.w);
              const __half2 h0 = __floats2half2_rn(xv.x, xv.y), h1 = __floats2half2_rn(xv.z, xv.w);
              hw[j >> 1] = *reinterpret_cast<const uint32_t*>(&h0);
              hw[(j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            }
          }
          if (sc == CPT / 64 - 1) {   // the accumulator has been read: hand the TMEM buffer back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
          }
          if (valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) *reinterpret_cast<float4*>(xp + q * cpitch) = xr[q];
          }
          staged_store(stg, vh, p.out_hi, g, lane);
          if (p.out_lo != nullptr) {                  // last resblock only: the tail's split operand needs x - fp16(x)
            uint32_t* lw = reinterpret_cast<uint32_t*>(vh);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float4 xv = xr[q];
              const float2 f0 = __half22float2(__floats2half2_rn(xv.x, xv.y)), f1 = __half22float2(__floats2half2_rn(xv.z, xv.w));
              const __half2 l0 = __floats2half2_rn(xv.x - f0.x, xv.y - f0.y), l1 = __floats2half2_rn(xv.z - f1.x, xv.w - f1.y);
              lw[2 * q] = *reinterpret_cast<const uint32_t*>(&l0);
              lw[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&l1);
            }
            staged_store(stg, vh, p.out_lo, g, lane);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      } else if constexpr (Cfg::EPI == kEpiResidualQ || Cfg::EPI == kEpiResidualQLast) {
        // ------------------------------------------------------------------ resblock output, fp16 + 8 bit trunk
        // a thread owns CH / 2 channels of its pixel: one pass of 64 (128 features) or two (256)
        constexpr int CPT = Cfg::CH / 2;
        const bool pf = pt + npairs < pair_tiles;
        if constexpr (Cfg::XT) {
          q_epilogue_xt<Cfg>(p, &tm_x0, &tm_x1, tc, tn, pt == pair, pf, wq, half * CPT, lane, row, valid, xt, &xfull[2 * warp],
                             xphase, taddr, s_bias_addr, &tmem_full[acc], acc_phase, &tmem_empty[acc]);
          xphase ^= 1;
        } else {
#pragma unroll 1
          for (int sc = 0; sc < CPT / 64; ++sc)
            q_epilogue_pass<Cfg, 64>(p, tc, tn, pf, wq, half * CPT + sc * 64, lane, row, valid,
                                     smem_u32(s_stg) + (uint32_t)(warp * 1024), taddr, s_bias_addr,
                                     sc == 0 ? &tmem_full[acc] : nullptr, acc_phase,
                                     sc == CPT / 64 - 1 ? &tmem_empty[acc] : nullptr);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      } else if constexpr (Cfg::EPI == kEpiHeadQ) {
        // ------------------------------------------------------------------ first layer: seeds the fp16 + 8 bit trunk
        // x = relu(conv + bias) (DSen2Net.py:29) -> x_hi (NHWC fp16, staged store) + one byte per element (see q_encode4)
        constexpr int CPT = Cfg::CH / 2;
        static_assert(CPT % 64 == 0 && !Cfg::SPLIT, "Q-trunk head: 64 channels per thread and pass, one accumulator");
        const uint32_t stg = smem_u32(s_stg) + (uint32_t)(warp * 1024);
#pragma unroll 1
        for (int sc = 0; sc < CPT / 64; ++sc) {
          const int cb = half * CPT + sc * 64;
          const EpiGeom g = epi_geom<Cfg>(p, tc, wq, cb);
          uint4 vh[8], lq[4];
          if (sc == 0) {
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
          }
#pragma unroll
          for (int chunk = 0; chunk < 2; ++chunk) {
            const int c0 = cb + chunk * 32;
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0, r);
            tmem_ld_wait();
            uint32_t* hw = reinterpret_cast<uint32_t*>(vh) + chunk * 16;
            uint32_t* qw = reinterpret_cast<uint32_t*>(lq) + chunk * 8;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bq = lds_f4(s_bias_addr + (uint32_t)(c0 + j) * 4);   // broadcast LDS.128
              q_encode4(fmaxf(__uint_as_float(r[j]) + bq.x, 0.f), fmaxf(__uint_as_float(r[j + 1]) + bq.y, 0.f),
                        fmaxf(__uint_as_float(r[j + 2]) + bq.z, 0.f), fmaxf(__uint_as_float(r[j + 3]) + bq.w, 0.f),
                        hw[j >> 1], hw[(j >> 1) + 1], qw[j >> 2]);
            }
          }
          if (sc == CPT / 64 - 1) {   // the accumulator has been read: hand the TMEM buffer back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
          }
          if (valid) {
            uint8_t* const qp = p.xq + ((((long long)b * p.H + y) * p.tiles_x + tc.tx) * (Cfg::CH / 16) + cb / 16) * 128 +
                                (row & 7) * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(qp + q * 128) = lq[q];
          }
          if constexpr (Cfg::XT) {
            xt_release<0>(lane);                                                    // the previous store has left the tile
            xt_store(xt, vh, &tm_x0, cb, tc.tx * 8, tc.ty * 16 + wq * 4, b, lane);
          } else {
            staged_store(stg, vh, p.out_hi, g, lane);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      } else {
        // ------------------------------------------------------------------ trunk layers
        // Each thread owns one pixel and CPT = CH/2 channels, processed as CPT/64 "super-chunks" of 64 channels
        // (128 B per pixel: the unit the staged transposing load / store moves).  128 features: one; 256: two.
        constexpr int CPT = Cfg::CH / 2;
        static_assert(CPT % 64 == 0, "the staged epilogue moves 64 channels (128 B) per pixel, warp and super-chunk");
        const uint32_t stg = smem_u32(s_stg) + (uint32_t)(warp * 1024);
        const bool want_lo = p.out_lo != nullptr;
#pragma unroll 1
        for (int sc = 0; sc < CPT / 64; ++sc) {
          const int cb = half * CPT + sc * 64;        // first channel of this super-chunk
          const EpiGeom g = epi_geom<Cfg>(p, tc, wq, cb);
          uint4 vh[8], vl[8];                         // residual in, then outputs (thread = pixel layout)
          if (Cfg::EPI == kEpiMask) {                 // ReLU backward: the forward activation decides which gradients pass
            uint4 gm[8];
            coalesced_load(gm, p.res_hi, g, lane);
            if (pt + npairs < pair_tiles) {
              const EpiGeom gn = epi_geom<Cfg>(p, tn, wq, cb);
              prefetch_rows(p.res_hi, gn, lane);
            }
            staged_gather(stg, gm, vh, lane);
          }
          if (sc == 0) {
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
          }
#pragma unroll
          for (int chunk = 0; chunk < 2; ++chunk) {
            const int c0 = cb + chunk * 32;
            uint32_t r[32];
            tmem_ld_32x32(taddr + c0, r);
            if (Cfg::SPLIT) {
              uint32_t r2[32];
              tmem_ld_32x32(taddr + Cfg::NTOT / 2 + c0, r2);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            } else {
              tmem_ld_wait();
            }
            uint32_t* hw = reinterpret_cast<uint32_t*>(vh) + chunk * 16;
            uint32_t* lw = reinterpret_cast<uint32_t*>(vl) + chunk * 16;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bq = lds_f4(s_bias_addr + (uint32_t)(c0 + j) * 4);   // broadcast LDS.128
              const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
              if (Cfg::EPI == kEpiRelu && p.x32 != nullptr && valid) {   // first layer: seed the fp32 trunk (tile-row-major)
                const float4 xv = make_float4(fmaxf(__uint_as_float(r[j]) + bq.x, 0.f), fmaxf(__uint_as_float(r[j + 1]) + bq.y, 0.f),
                                              fmaxf(__uint_as_float(r[j + 2]) + bq.z, 0.f), fmaxf(__uint_as_float(r[j + 3]) + bq.w, 0.f));
                *reinterpret_cast<float4*>(p.x32 + ((((long long)b * p.H + y) * p.tiles_x + tc.tx) * (Cfg::CH / 4) + ((c0 + j) >> 2)) * 32 +
                                           (row & 7) * 4) = xv;
              }
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const int jj = j + 2 * h2;
                float v0 = __uint_as_float(r[jj]) + bb[2 * h2];
                float v1 = __uint_as_float(r[jj + 1]) + bb[2 * h2 + 1];
                if (Cfg::EPI == kEpiMask) {
                  const float2 m = __half22float2(*reinterpret_cast<const __half2*>(&hw[jj >> 1]));
                  v0 = m.x > 0.f ? v0 : 0.f;               // d relu(z) / dz = [z > 0]  (relu(z) > 0  <=>  z > 0)
                  v1 = m.y > 0.f ? v1 : 0.f;
                } else {
                  v0 = fmaxf(v0, 0.f);
                  v1 = fmaxf(v1, 0.f);
                }
                const __half2 h = __floats2half2_rn(v0, v1);
                hw[jj >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                if (want_lo) {
                  const float2 hf = __half22float2(h);
                  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                  lw[jj >> 1] = *reinterpret_cast<const uint32_t*>(&l);
                }
              }
            }
          }
          if (sc == CPT / 64 - 1) {   // the accumulator has been read: hand the TMEM buffer back before the last stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
          }
          staged_store(stg, vh, p.out_hi, g, lane);
          if (want_lo) staged_store(stg, vl, p.out_lo, g, lane);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if constexpr (Cfg::XT) {
    if (warp < Cfg::EPI_WARPS && lane == 0) tma_store_wait_all<0>();   // the last tile's store has left shared memory
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // nobody leaves while the peer may still signal it or read its smem
  tc_fence_after();
  if (warp == Cfg::MMA_WARP) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------ //
// host side
// ------------------------------------------------------------------------------------------ //
using CfgRelu = PairCfg<128, kSplitNone, 1, 2, 9, 4, 3, kEpiRelu>;
using CfgResidual32 = PairCfg<128, kSplitNone, 1, 2, 9, 4, 3, kEpiResidual32>;
using CfgMask = PairCfg<128, kSplitNone, 1, 2, 9, 4, 3, kEpiMask>;
// fp16 + 8 bit trunk update: epilogue-bound, not load-bound -- two activation stages are enough (measured: 3 -> 2 stages costs
// the RELU layer 9 % and this one nothing), which frees the shared memory for the eight 4 KB x_hi tiles of the XT epilogue
using CfgResidualQ = PairCfg<128, kSplitNone, 1, 2, 9, 4, 2, kEpiResidualQ, 0, 128, true>;
using CfgResidualQLast = PairCfg<128, kSplitNone, 1, 2, 9, 4, 2, kEpiResidualQLast, 0, 128, true>;
// VDSen2 trunk (256 -> 256): 1.18 MB of weights per layer cannot be resident -- ring of 8 streamed [tap][k-block] slabs
using CfgRelu256 = PairCfg<256, kSplitNone, 1, 4, 9, 4, 3, kEpiRelu, 8>;
using CfgResidualQ256 = PairCfg<256, kSplitNone, 1, 4, 9, 4, 3, kEpiResidualQ, 8>;
using CfgResidualQLast256 = PairCfg<256, kSplitNone, 1, 4, 9, 4, 3, kEpiResidualQLast, 8>;
// training step of the 256-feature network: fp32 trunk update and ReLU backward with the streamed weight ring
using CfgResidual32_256 = PairCfg<256, kSplitNone, 1, 4, 9, 4, 3, kEpiResidual32, 8>;
using CfgMask256 = PairCfg<256, kSplitNone, 1, 4, 9, 4, 3, kEpiMask, 8>;
using CfgHead = PairCfg<256, kSplitStack, 2, 1, 3, 3, 6, kEpiRelu>;
// first layer on the un-gathered 16-channel input: nine taps through shifted descriptors into a 32-byte-row halo box,
// three products per tap into one accumulator (27 MMAs of N = F per tile pair)
using CfgHead16 = PairCfg<128, kSplit3, 2, 1, 9, 1, 6, kEpiHeadQ, 0, 32, true>;
using CfgHead16_256 = PairCfg<256, kSplit3, 2, 1, 9, 1, 6, kEpiHeadQ, 0, 32, true>;
// the same first layer seeding an fp32 trunk (training step of the 256-feature network)
using CfgHead16Relu256 = PairCfg<256, kSplit3, 2, 1, 9, 1, 6, kEpiRelu, 0, 32>;
using CfgTail = PairCfg<32, kSplitStack, 2, 2, 9, 4, 6, kEpiTail>;

// x0 / x1: tensor maps of the epilogue's x_hi / x_lo tiles (PairCfg::XT kernels only)
template <class Cfg>
static int launch_pair(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const PairParams& p,
                       int sms, cudaStream_t stream, const char* what, const CUtensorMap* x0 = nullptr,
                       const CUtensorMap* x1 = nullptr) {
  DSEN2_REQUIRE(!Cfg::XT || x0 != nullptr, DSEN2_E_BADARG, "%s: internal error, epilogue tensor map missing", what);
  static bool configured[64] = {};
  if (needs_config(configured)) {
    DSEN2_CUDA(cudaFuncSetAttribute(conv_pair_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::SMEM_BYTES));
  }
  const long long pair_tiles = ((long long)p.num_tiles + 1) / 2;
  const long long max_pairs = sms / 2;
  const int pairs = (int)(pair_tiles < max_pairs ? pair_tiles : max_pairs);
  conv_pair_kernel<Cfg><<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(a0, a1, w, x0 ? *x0 : a0, x1 ? *x1 : (x0 ? *x0 : a0), p);
  return check_launch(what);
}

template <class Cfg>
static int make_maps(CUtensorMap* a0, CUtensorMap* a1, CUtensorMap* w, const void* d_a0, const void* d_a1,
                     const void* d_w, int n, int H, int W) {
  const int ca = Cfg::KPM * Cfg::KCH;
  const uint64_t dims[4] = {(uint64_t)ca, (uint64_t)W, (uint64_t)H, (uint64_t)n};
  const uint32_t box[4] = {(uint32_t)Cfg::KCH, (uint32_t)Cfg::BOXW, (uint32_t)kBoxH, 1};
  int rc = make_tmap_f16_sw(a0, d_a0, 4, dims, box, Cfg::ROWB);
  if (rc) return rc;
  rc = make_tmap_f16_sw(a1, d_a1 ? d_a1 : d_a0, 4, dims, box, Cfg::ROWB);
  if (rc) return rc;
  const uint64_t wd[3] = {(uint64_t)ca, (uint64_t)Cfg::NTOT, (uint64_t)Cfg::WGROUPS};
  const uint32_t wb[3] = {(uint32_t)Cfg::KCH, (uint32_t)Cfg::SLAB_ROWS, 1};
  return make_tmap_f16_sw(w, d_w, 3, wd, wb, Cfg::ROWB);
}

// epilogue tile map of an NHWC fp16 tensor (n, H, W, C): box = box_ch (64 or 32) channels x 8 pixels x 4 rows (one warp's
// share of a tile, or half of it), swizzle span = the box row
static int make_xt_map(CUtensorMap* m, const void* d_x, int n, int H, int W, int C, int box_ch) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)n};
  const uint32_t box[4] = {(uint32_t)box_ch, 8, 4, 1};
  return make_tmap_f16_sw(m, d_x, 4, dims, box, box_ch * 2);
}

static int fill_tiles(PairParams& p, int n, int H, int W) {
  p.n = n; p.H = H; p.W = W;
  p.tiles_x = ceil_div(W, 8);
  p.tiles_y = ceil_div(H, 16);
  const long long tiles = (long long)n * p.tiles_x * p.tiles_y;
  DSEN2_REQUIRE(tiles < (1LL << 30), DSEN2_E_BADARG, "batch too large: %lld tiles of 16x8 pixels (limit 2^30)", tiles);
  p.num_tiles = (uint32_t)tiles;
  return 0;
}

}  // namespace dsen2

using namespace dsen2;

// first convolution of a resBlock: relu(conv3x3(x) + b) -> NHWC fp16 (DSen2Net.py:10-11), F = 128 or 256
extern "C" int dsen2_conv_relu(const void* d_in, const void* d_w, const float* d_bias, int n, int H, int W, int feature_size,
                               void* d_out, void* stream) {
  DSEN2_REQUIRE(d_in && d_w && d_bias && d_out, DSEN2_E_BADARG, "dsen2_conv_relu: null pointer");
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0, DSEN2_E_BADARG, "dsen2_conv_relu: bad shape (n %d, %dx%d)", n, H, W);
  DSEN2_REQUIRE(feature_size == 128 || feature_size == 256, DSEN2_E_BADARG,
                "dsen2_conv_relu: feature size must be 128 or 256 (got %d)", feature_size);
  DSEN2_REQUIRE(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_w % 16) == 0 && ((uintptr_t)d_out % 16) == 0, DSEN2_E_ALIGN,
                "dsen2_conv_relu: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  PairParams p{};
  if ((rc = fill_tiles(p, n, H, W)) != 0) return rc;
  p.bias = d_bias;
  p.out_hi = (__half*)d_out;
  CUtensorMap a0, a1, w;
  if (feature_size == 256) {
    rc = make_maps<CfgRelu256>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
    if (rc) return rc;
    return launch_pair<CfgRelu256>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<relu,256>");
  }
  rc = make_maps<CfgRelu>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
  if (rc) return rc;
  return launch_pair<CfgRelu>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<relu>");
}

extern "C" int dsen2_conv_relu_bwd(const void* d_in, const void* d_w, const float* d_bias, const void* d_fwd_act, int n,
                                   int H, int W, int feature_size, void* d_out, void* stream) {
  DSEN2_REQUIRE(d_in && d_w && d_bias && d_fwd_act && d_out, DSEN2_E_BADARG, "dsen2_conv_relu_bwd: null pointer");
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0, DSEN2_E_BADARG, "dsen2_conv_relu_bwd: bad shape");
  DSEN2_REQUIRE(feature_size == 128 || feature_size == 256, DSEN2_E_BADARG,
                "dsen2_conv_relu_bwd: feature size must be 128 or 256 (got %d)", feature_size);
  DSEN2_REQUIRE(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_w % 16) == 0 && ((uintptr_t)d_fwd_act % 16) == 0 &&
                    ((uintptr_t)d_out % 16) == 0,
                DSEN2_E_ALIGN, "dsen2_conv_relu_bwd: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  PairParams p{};
  if ((rc = fill_tiles(p, n, H, W)) != 0) return rc;
  p.bias = d_bias;
  p.res_hi = (const __half*)d_fwd_act;
  p.out_hi = (__half*)d_out;
  CUtensorMap a0, a1, w;
  if (feature_size == 256) {
    rc = make_maps<CfgMask256>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
    if (rc) return rc;
    return launch_pair<CfgMask256>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<relu_bwd,256>");
  }
  rc = make_maps<CfgMask>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
  if (rc) return rc;
  return launch_pair<CfgMask>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<relu_bwd>");
}

extern "C" int dsen2_conv_res32(const void* d_in, const void* d_w, const float* d_bias, int n, int H, int W, int feature_size,
                                float res_scale, float* d_trunk32, void* d_out_hi, void* d_out_lo, void* stream) {
  DSEN2_REQUIRE(d_in && d_w && d_bias && d_trunk32 && d_out_hi, DSEN2_E_BADARG, "dsen2_conv_res32: null pointer");
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0, DSEN2_E_BADARG, "dsen2_conv_res32: bad shape");
  DSEN2_REQUIRE(feature_size == 128 || feature_size == 256, DSEN2_E_BADARG,
                "dsen2_conv_res32: feature size must be 128 or 256 (got %d)", feature_size);
  DSEN2_REQUIRE(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_w % 16) == 0 && ((uintptr_t)d_trunk32 % 16) == 0 &&
                    ((uintptr_t)d_out_hi % 16) == 0 && ((uintptr_t)d_out_lo % 16) == 0,
                DSEN2_E_ALIGN, "dsen2_conv_res32: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  PairParams p{};
  if ((rc = fill_tiles(p, n, H, W)) != 0) return rc;
  p.bias = d_bias;
  p.res_scale = res_scale;
  p.x32 = d_trunk32;
  p.out_hi = (__half*)d_out_hi; p.out_lo = (__half*)d_out_lo;
  CUtensorMap a0, a1, w;
  if (feature_size == 256) {
    rc = make_maps<CfgResidual32_256>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
    if (rc) return rc;
    return launch_pair<CfgResidual32_256>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<residual32,256>");
  }
  rc = make_maps<CfgResidual32>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
  if (rc) return rc;
  return launch_pair<CfgResidual32>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<residual32>");
}

template <class CfgH>
static int head_common(const char* name, const void* d_xin_hi, const void* d_xin_lo, const void* d_w, const float* d_bias,
                       int n, int H, int W, void* d_out_hi, void* d_out_lo, float* d_trunk32, void* d_trunk_lo8,
                       void* stream) {
  DSEN2_REQUIRE(d_xin_hi && d_xin_lo && d_w && d_bias && d_out_hi, DSEN2_E_BADARG, "%s: null pointer", name);
  DSEN2_REQUIRE(((uintptr_t)d_trunk32 % 16) == 0 && ((uintptr_t)d_trunk_lo8 % 16) == 0, DSEN2_E_ALIGN,
                "%s: trunk must be 16-byte aligned", name);
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0, DSEN2_E_BADARG, "%s: bad shape", name);
  DSEN2_REQUIRE(((uintptr_t)d_xin_hi % 16) == 0 && ((uintptr_t)d_xin_lo % 16) == 0 && ((uintptr_t)d_w % 16) == 0 &&
                    ((uintptr_t)d_out_hi % 16) == 0 && ((uintptr_t)d_out_lo % 16) == 0,
                DSEN2_E_ALIGN, "%s: pointers must be 16-byte aligned", name);
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  PairParams p{};
  if ((rc = fill_tiles(p, n, H, W)) != 0) return rc;
  p.bias = d_bias;
  p.out_hi = (__half*)d_out_hi; p.out_lo = (__half*)d_out_lo;
  p.x32 = d_trunk32;
  p.xq = (uint8_t*)d_trunk_lo8;
  CUtensorMap a0, a1, w;
  rc = make_maps<CfgH>(&a0, &a1, &w, d_xin_hi, d_xin_lo, d_w, n, H, W);
  if (rc) return rc;
  CUtensorMap x0;
  if (CfgH::XT && (rc = make_xt_map(&x0, d_out_hi, n, H, W, CfgH::CH, 64)) != 0) return rc;
  return launch_pair<CfgH>(a0, a1, w, p, sms, (cudaStream_t)stream, name, CfgH::XT ? &x0 : nullptr);
}

extern "C" int dsen2_conv_head(const void* d_xin_hi, const void* d_xin_lo, const void* d_w, const float* d_bias,
                               int n, int H, int W, int feature_size, void* d_out_hi, void* d_out_lo,
                               float* d_trunk32, void* stream) {
  DSEN2_REQUIRE(feature_size == 128, DSEN2_E_BADARG, "dsen2_conv_head: feature_size 128 only (got %d)", feature_size);
  return head_common<CfgHead>("dsen2_conv_head", d_xin_hi, d_xin_lo, d_w, d_bias, n, H, W, d_out_hi, d_out_lo, d_trunk32,
                              nullptr, stream);
}

extern "C" int dsen2_conv_head16_q(const void* d_xin_hi, const void* d_xin_lo, const void* d_w, const float* d_bias,
                                   int n, int H, int W, int feature_size, void* d_x_hi, void* d_trunk_lo8, void* stream) {
  DSEN2_REQUIRE(d_trunk_lo8, DSEN2_E_BADARG, "dsen2_conv_head16_q: null pointer");
  DSEN2_REQUIRE(feature_size == 128 || feature_size == 256, DSEN2_E_BADARG,
                "dsen2_conv_head16_q: feature_size must be 128 or 256 (got %d)", feature_size);
  if (feature_size == 256)
    return head_common<CfgHead16_256>("dsen2_conv_head16_q", d_xin_hi, d_xin_lo, d_w, d_bias, n, H, W, d_x_hi, nullptr,
                                      nullptr, d_trunk_lo8, stream);
  return head_common<CfgHead16>("dsen2_conv_head16_q", d_xin_hi, d_xin_lo, d_w, d_bias, n, H, W, d_x_hi, nullptr, nullptr,
                                d_trunk_lo8, stream);
}

// The training step's first layer of the 256-feature network: the same three-product convolution on the 16-channel prepared
// input, relu -> NHWC fp16 + the seed of the fp32 trunk (DSen2Net.py:29).
extern "C" int dsen2_conv_head16_relu(const void* d_xin_hi, const void* d_xin_lo, const void* d_w, const float* d_bias,
                                      int n, int H, int W, int feature_size, void* d_out_hi, float* d_trunk32, void* stream) {
  DSEN2_REQUIRE(feature_size == 256, DSEN2_E_BADARG, "dsen2_conv_head16_relu: feature_size 256 only (got %d)", feature_size);
  return head_common<CfgHead16Relu256>("dsen2_conv_head16_relu", d_xin_hi, d_xin_lo, d_w, d_bias, n, H, W, d_out_hi, nullptr,
                                       d_trunk32, nullptr, stream);
}

static int resq_common(int features, const void* d_in, const void* d_w, const float* d_bias, int n, int H, int W,
                       float res_scale, void* d_x_hi, void* d_trunk_lo8, void* d_out_lo, void* stream) {
  DSEN2_REQUIRE(d_in && d_w && d_bias && d_x_hi && d_trunk_lo8, DSEN2_E_BADARG, "dsen2_conv_resq: null pointer");
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0, DSEN2_E_BADARG, "dsen2_conv_resq: bad shape");
  DSEN2_REQUIRE(d_in != d_x_hi, DSEN2_E_BADARG, "dsen2_conv_resq: the convolution input must not alias x_hi (updated in place)");
  DSEN2_REQUIRE(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_w % 16) == 0 && ((uintptr_t)d_trunk_lo8 % 16) == 0 &&
                    ((uintptr_t)d_x_hi % 16) == 0 && ((uintptr_t)d_out_lo % 16) == 0,
                DSEN2_E_ALIGN, "dsen2_conv_resq: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  PairParams p{};
  if ((rc = fill_tiles(p, n, H, W)) != 0) return rc;
  p.bias = d_bias;
  p.res_scale = res_scale;
  p.xq = (uint8_t*)d_trunk_lo8;
  p.out_hi = (__half*)d_x_hi; p.out_lo = (__half*)d_out_lo;
  CUtensorMap a0, a1, w;
  if (features == 256) {
    rc = make_maps<CfgResidualQ256>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
    if (rc) return rc;
    if (d_out_lo)
      return launch_pair<CfgResidualQLast256>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<residualq,last,256>");
    return launch_pair<CfgResidualQ256>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<residualq,256>");
  }
  rc = make_maps<CfgResidualQ>(&a0, &a1, &w, d_in, nullptr, d_w, n, H, W);
  if (rc) return rc;
  CUtensorMap x0, x1;
  if ((rc = make_xt_map(&x0, d_x_hi, n, H, W, 128, 32)) != 0) return rc;
  if (d_out_lo) {
    if ((rc = make_xt_map(&x1, d_out_lo, n, H, W, 128, 32)) != 0) return rc;
    return launch_pair<CfgResidualQLast>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<residualq,last>", &x0, &x1);
  }
  return launch_pair<CfgResidualQ>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<residualq>", &x0);
}

extern "C" int dsen2_conv_resq(const void* d_in, const void* d_w, const float* d_bias, int n, int H, int W,
                               float res_scale, void* d_x_hi, void* d_trunk_lo8, void* d_out_lo, void* stream) {
  return resq_common(128, d_in, d_w, d_bias, n, H, W, res_scale, d_x_hi, d_trunk_lo8, d_out_lo, stream);
}

extern "C" int dsen2_conv_resq256(const void* d_in, const void* d_w, const float* d_bias, int n, int H, int W,
                                  float res_scale, void* d_x_hi, void* d_trunk_lo8, void* d_out_lo, void* stream) {
  return resq_common(256, d_in, d_w, d_bias, n, H, W, res_scale, d_x_hi, d_trunk_lo8, d_out_lo, stream);
}

// Last layer on the 64-channel prepared input (training step, networks without resblocks): pixels on the M side, B =
// [W_hi ; W_lo] stacked along N.  The inference path uses the swapped form in conv_tail.cu (dsen2_conv_tail16[_stitch]).
extern "C" int dsen2_conv_tail(const void* d_x_hi, const void* d_x_lo, const void* d_w, const float* d_bias,
                               const void* d_xin_hi, const void* d_xin_lo, int skip_ch0, int cout, int n, int H, int W,
                               float* d_pred_nchw, void* stream) {
  DSEN2_REQUIRE(d_x_hi && d_x_lo && d_w && d_bias && d_xin_hi && d_xin_lo && d_pred_nchw, DSEN2_E_BADARG,
                "dsen2_conv_tail: null pointer");
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0 && cout > 0 && cout <= 16 && skip_ch0 >= 0 && skip_ch0 + cout <= 16,
                DSEN2_E_BADARG, "dsen2_conv_tail: bad shape (cout %d, skip_ch0 %d)", cout, skip_ch0);
  DSEN2_REQUIRE(((uintptr_t)d_x_hi % 16) == 0 && ((uintptr_t)d_x_lo % 16) == 0 && ((uintptr_t)d_w % 16) == 0,
                DSEN2_E_ALIGN, "dsen2_conv_tail: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  PairParams p{};
  if ((rc = fill_tiles(p, n, H, W)) != 0) return rc;
  p.tail_mode = 0;
  p.out_mul = 1.0f;
  p.bias = d_bias;
  p.skip_hi = (const __half*)d_xin_hi; p.skip_lo = (const __half*)d_xin_lo; p.skip_ch0 = skip_ch0;
  p.skip_pitch = 64; p.skip_off = 16;
  p.cout_real = cout; p.out_f32 = d_pred_nchw;
  CUtensorMap a0, a1, w;
  rc = make_maps<CfgTail>(&a0, &a1, &w, d_x_hi, d_x_lo, d_w, n, H, W);
  if (rc) return rc;
  return launch_pair<CfgTail>(a0, a1, w, p, sms, (cudaStream_t)stream, "conv_pair<tail>");
}
