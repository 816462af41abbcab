// Training step of DSen2 (training/supres_train.py:137-144,218-230): the pieces that are not convolutions.
//
//   backward data path   = the CTA-pair convolution kernels of conv_pair.cu on flipped / transposed weights
//                          (dsen2_pack_dgrad_weights), with the ReLU-backward mask or the fp32 "gradient trunk"
//                          accumulate epilogue -- no new convolution code;
//   weight gradients     = dsen2_wgrad_nhwc: a tcgen05 GEMM over the pixel dimension straight from the NHWC tensors,
//                              dW[tap][ci][co] = sum_px X[px + off(tap)][ci] * dY[px][co]
//                          (MN-major operands, see wgrad_direct_kernel);
//   bias gradients       = column sums of dY, taken inside the weight-gradient kernel from the tiles it stages;
//   the rest             = layout changes, MAE loss + gradient, Keras-2 Nadam.
// Gradients flow in fp16 with a power-of-two loss scale chosen by the host so that d(pred) = +-2^-4 exactly.
#include <stdlib.h>

#include "common.cuh"
#include "tiling.cuh"

namespace dsen2 {

// ------------------------------------------------------------------------------------------ //
// layout changes
// ------------------------------------------------------------------------------------------ //
// up to three NCHW fp32 inputs (n,c_i,H,W), at most 16 channels in total, concatenated along channels -> the first 16
// channels of NHWC fp16 (n,H,W,cpad).  One thread per pixel: plane reads are coalesced across the warp, each thread
// writes one 32-byte sector.  (The remaining channels are zeroed by a memset in the entry point.)
__global__ void nchw_to_nhwc_f16_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1,
                                        const float* __restrict__ x2, int c2, long long hw, int cpad, long long npix,
                                        __half* __restrict__ out) {
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < npix;
       pix += (long long)gridDim.x * blockDim.x) {
    const long long b = pix / hw, r = pix - b * hw;
    __align__(16) __half v[16];
#pragma unroll
    for (int ch = 0; ch < 16; ++ch) {
      float f = 0.f;
      if (ch < c0) f = x0[(b * c0 + ch) * hw + r];
      else if (ch < c0 + c1) f = x1[(b * c1 + (ch - c0)) * hw + r];
      else if (ch < c0 + c1 + c2) f = x2[(b * c2 + (ch - c0 - c1)) * hw + r];
      v[ch] = __float2half_rn(f);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + pix * cpad);
    dst[0] = reinterpret_cast<const uint4*>(v)[0];
    dst[1] = reinterpret_cast<const uint4*>(v)[1];
  }
}

// out = in where the forward activation is positive, else 0 (ReLU backward on NHWC fp16 tensors)
__global__ void relu_mask_kernel(const uint4* __restrict__ in, const uint4* __restrict__ act, long long total16,
                                 uint4* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total16;
       idx += (long long)gridDim.x * blockDim.x) {
    uint4 v = in[idx];
    const uint4 a = act[idx];
    __half* vh = reinterpret_cast<__half*>(&v);
    const __half* ah = reinterpret_cast<const __half*>(&a);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(__half2float(ah[j]) > 0.f)) vh[j] = __float2half_rn(0.f);
    out[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------ //
// loss (mean_absolute_error, metric mean_squared_error; supres_train.py:144)
// ------------------------------------------------------------------------------------------ //
// dpred = gscale * sign(pred - y); sums[0] += sum |pred - y|, sums[1] += sum (pred - y)^2  (double accumulators)
__global__ void mae_grad_kernel(const float* __restrict__ pred, const float* __restrict__ y, long long total, float gscale,
                                float* __restrict__ dpred, double* __restrict__ sums) {
  double a = 0.0, q = 0.0;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const float d = pred[idx] - y[idx];
    a += fabsf(d);
    q += (double)d * d;
    dpred[idx] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
  }
  __shared__ double ra[32], rq[32];
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) { ra[threadIdx.x >> 5] = a; rq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sq = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { sa += ra[i]; sq += rq[i]; }
    atomicAdd(&sums[0], sa);
    atomicAdd(&sums[1], sq);
  }
}

// ------------------------------------------------------------------------------------------ //
// Keras-2 Nadam (supres_train.py:137-141); the schedule scalars are computed on the host
// ------------------------------------------------------------------------------------------ //
struct NadamHp { float gmul, lr, beta1, beta2, eps, mu_t, mu_next, sched_new, sched_next, bias2; };

__device__ __forceinline__ void nadam_element(float& p, float g, float& m, float& v, const NadamHp& h) {
  const float gr = g * h.gmul;
  const float g_prime = gr / (1.f - h.sched_new);
  const float m_t = h.beta1 * m + (1.f - h.beta1) * gr;
  const float m_prime = m_t / (1.f - h.sched_next);
  const float v_t = h.beta2 * v + (1.f - h.beta2) * gr * gr;
  const float v_prime = v_t / h.bias2;
  const float m_bar = (1.f - h.mu_t) * g_prime + h.mu_next * m_prime;
  p = p - h.lr * m_bar / (sqrtf(v_prime) + h.eps);
  m = m_t;
  v = v_t;
}

// 28 bytes of traffic per parameter (VDSen2: 37.8 M parameters, 1.06 GB per step): 128-bit accesses, scalar tail
__device__ __forceinline__ void nadam_sweep(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                            float* __restrict__ v, long long total, const NadamHp& h) {
  const long long t4 = total >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < t4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    nadam_element(pp.x, gg.x, mm.x, vv.x, h);
    nadam_element(pp.y, gg.y, mm.y, vv.y, h);
    nadam_element(pp.z, gg.z, mm.z, vv.z, h);
    nadam_element(pp.w, gg.w, mm.w, vv.w, h);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  const long long idx = (t4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx < total) nadam_element(p[idx], g[idx], m[idx], v[idx], h);
}

__global__ void nadam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long total, NadamHp h) {
  nadam_sweep(p, g, m, v, total, h);
}

// the same update with the step-dependent scalars read from device memory (CUDA-graph replay)
__global__ void nadam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long total, const float* __restrict__ hp) {
  const NadamHp h = {hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7], hp[8], hp[9]};
  nadam_sweep(p, g, m, v, total, h);
}

// dgrad weights: the backward-data convolution is a forward convolution of dY with the taps flipped and the channel
// roles swapped: packed[t][i][o] = scale * hwio[8 - t][i][o]   (rows = former input channels, k = former outputs)
__global__ void pack_dgrad_weights_kernel(const float* __restrict__ hwio, int cin, int cout, int rows_pad, int k_pad,
                                          float scale, long long total, __half* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % k_pad);
    const int i = (int)((idx / k_pad) % rows_pad);
    const int t = (int)(idx / ((long long)k_pad * rows_pad));
    float v = 0.f;
    if (i < cin && o < cout) v = scale * hwio[((long long)(8 - t) * cin + i) * cout + o];
    out[idx] = __float2half_rn(v);
  }
}

// All F -> F layers of the network in ONE launch (they are a regular stride apart in the flat Keras-order parameter vector):
// forward operand fwd[l][t][o][i] = f16(w[t][i][o]) and backward-data operand bwd[l][t][i][o] = f16(s_l * w[8-t][i][o]),
// s_l = scale_second for the second convolution of a resBlock (odd l), 1 otherwise.  Same arithmetic as
// pack_weights_kernel / pack_dgrad_weights_kernel, 2 * layers - 1 launches fewer per training step.
// One block = one 32 x 32 tile of one (layer, tap) matrix w[i][o]: read once (coalesced along o), written straight to bwd
// (same orientation) and, transposed through shared memory, to fwd (coalesced along i; 256 features: 37.7 M elements per
// step).
__global__ void __launch_bounds__(256)
pack_trunk_layers_kernel(const float* __restrict__ params, long long layer_stride, int F, float scale_second,
                         __half* __restrict__ fwd, __half* __restrict__ bwd) {
  __shared__ float tile[32][33];
  const int l = blockIdx.z, t = blockIdx.y;
  const int tiles = F / 32;
  const int i0 = (int)(blockIdx.x / tiles) * 32, o0 = (int)(blockIdx.x % tiles) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* w = params + (long long)l * layer_stride + (long long)t * F * F;          // w[t][i][o]
  const float sc = (l & 1) ? scale_second : 1.0f;
  __half* f = fwd + ((long long)l * 9 + t) * F * F;
  __half* b = bwd + ((long long)l * 9 + (8 - t)) * F * F;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const float v = w[(long long)(i0 + r) * F + o0 + tx];
    tile[r][tx] = v;
    b[(long long)(i0 + r) * F + o0 + tx] = __float2half_rn(sc * v);
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) f[(long long)(o0 + r) * F + i0 + tx] = __float2half_rn(tile[tx][r]);
}

struct TileXY3 { int b, ty, tx; };
__device__ __forceinline__ TileXY3 decode_tile3(uint32_t tile, uint32_t tiles_x, uint32_t tiles_y) {
  const uint32_t row = tile / tiles_x;
  TileXY3 t;
  t.tx = (int)(tile - row * tiles_x);
  t.b = (int)(row / tiles_y);
  t.ty = (int)(row - (uint32_t)t.b * tiles_y);
  return t;
}

// ------------------------------------------------------------------------------------------ //
// dW[tap][ci][co] = sum_px X[px + off(tap)][ci] * dY[px][co]: the reduction runs over PIXELS, and in NHWC a pixel
// is a 128-byte row of 64 channels -- exactly an MN-major (M = ci or N = co contiguous, K = pixel) UMMA operand.
// So the TMA tiles of the forward convolution serve as they are: X as the halo box of a 16x8-pixel tile (the tap
// offset is a shift of the descriptor start, zero padding is TMA out-of-bounds fill), dY as the plain tile.
// One CTA owns ONE vertical tap (three accumulators = its three horizontal taps, 384 TMEM columns) and a slice
// of the tiles; at the end the accumulators go TMEM -> smem -> cp.reduce.async.bulk (fp32 add) into dW.
static constexpr int kWdThreads = 192;          // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 drain
static constexpr int kWdXBox = 16 * 10 * 128;   // 16 rows x 10 pixels x 64 channels
static constexpr int kWdYBox = 16 * 8 * 128;
static constexpr int kWdStage = 2 * kWdXBox + 2 * kWdYBox;   // both channel halves of X and of dY: 73,728 B
static constexpr int kWdStages = 3;

__global__ void __launch_bounds__(kWdThreads, 1)
wgrad_direct_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy, int tiles_x,
                    int tiles_y, int num_tiles, int tiles_per_split, int C, float scale, float* __restrict__ dw,
                    float* __restrict__ db) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWdStages * kWdStage);
  uint64_t* empty = full + kWdStages;
  uint64_t* done = empty + kWdStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dy = blockIdx.y;                                   // vertical tap 0..2
  // 256 features: blockIdx.z = (block of 128 input channels, block of 128 output channels) of the (C, C) gradient
  const int ci0 = (int)(blockIdx.z / (C / 128)) * 128, co0 = (int)(blockIdx.z % (C / 128)) * 128;
  // bias gradient db[co] += scale * sum_px dY[px][co]: the four drain warps (idle until the last MMA) sum the dY tiles the
  // CTA stages anyway -- no separate pass over dY.  The 3 * C/128 CTAs that stage the same dY tile (vertical taps x
  // input-channel blocks) take every (3 * C/128)-th pixel row each: the MMAs keep the shared-memory port busy, and one CTA
  // reading whole tiles became the straggler of the grid (measured: slower than the separate kernel).
  const bool colsum = db != nullptr;
  const int nshare = 3 * (C / 128), share = dy * (C / 128) + ci0 / 128;
  const int t0 = blockIdx.x * tiles_per_split;
  const int t1 = min(num_tiles, t0 + tiles_per_split);
  const int nt = t1 - t0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_dy);
    for (int i = 0; i < kWdStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], colsum ? 5 : 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nt; ++i) {
        const TileXY3 t = decode_tile3((uint32_t)(t0 + i), (uint32_t)tiles_x, (uint32_t)tiles_y);
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* s = smem + stage * kWdStage;
        mbar_expect_tx(&full[stage], kWdStage);
        tma_load_4d(s, &tm_x, &full[stage], ci0, t.tx * 8 - 1, t.ty * 16 + dy - 1, t.b);
        tma_load_4d(s + kWdXBox, &tm_x, &full[stage], ci0 + 64, t.tx * 8 - 1, t.ty * 16 + dy - 1, t.b);
        tma_load_4d(s + 2 * kWdXBox, &tm_dy, &full[stage], co0, t.tx * 8, t.ty * 16, t.b);
        tma_load_4d(s + 2 * kWdXBox + kWdYBox, &tm_dy, &full[stage], co0 + 64, t.tx * 8, t.ty * 16, t.b);
        if (++stage == kWdStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16_mn(128, 128);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nt; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + stage * kWdStage), sy = sx + 2 * kWdXBox;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
          for (int j = 0; j < 8; ++j)            // 16 pixels (two tile rows) per MMA
            umma_f16_ss(tmem_base + dx * 128, umma_desc_mn_sw128(sx + dx * 128 + j * 2 * 1280, kWdXBox, 1280),
                        umma_desc_mn_sw128(sy + j * 2 * 1024, kWdYBox, 1024), idesc, (uint32_t)((i | j) != 0));
        }
        umma_commit(&empty[stage]);
        if (i == nt - 1) umma_commit(done);
        if (++stage == kWdStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // drain: accumulator dx (128 ci x 128 co fp32) -> smem (row-major, 64 KB, reusing the pipeline stages) -> bulk add
    const int wq = warp & 3;
    const int row = wq * 32 + lane;              // ci
    if (colsum) {
      // thread = one channel pair (4 bytes of the 128-byte pixel row; a warp reads one whole row per instruction) and one
      // half of this CTA's share of the tile's 128 pixels; the tile is SWIZZLE_128B: 16-byte chunk index ^ (pixel row & 7)
      const int t = (int)threadIdx.x - 64, pair = t & 63, ph = t >> 6;
      const uint32_t off = (uint32_t)((pair >> 5) * kWdYBox + 2 * kWdXBox + (pair & 3) * 4);
      const int chunk = (pair & 31) >> 2;
      float sx = 0.f, sy = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nt; ++i) {
        mbar_wait(&full[stage], phase);
        const uint8_t* base = smem + stage * kWdStage + off;
#pragma unroll 4
        for (int r = ph * 64 + share; r < ph * 64 + 64; r += nshare) {
          const __half2 h = *reinterpret_cast<const __half2*>(base + r * 128 + ((chunk ^ (r & 7)) << 4));
          const float2 f = __half22float2(h);
          sx += f.x;
          sy += f.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kWdStages) { stage = 0; phase ^= 1; }
      }
      float* d = db + co0 + (pair >> 5) * 64 + (pair & 31) * 2;
      atomicAdd(d, sx * scale);
      atomicAdd(d + 1, sy * scale);
      // the drain below overwrites the pipeline stages: wait until all four warps have read their last dY tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    mbar_wait(done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int dx = 0; dx < 3; ++dx) {
      float* sacc = reinterpret_cast<float*>(smem + dx * 65536);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + dx * 128 + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          // 16-byte chunk index rotated by the row: conflict-free smem stores; the bulk copy below is per row
          const int ch = ((c0 + j) >> 2);
          *reinterpret_cast<float4*>(sacc + row * 128 + (((ch + row) & 31) << 2)) =
              make_float4(__uint_as_float(r[j]) * scale, __uint_as_float(r[j + 1]) * scale, __uint_as_float(r[j + 2]) * scale,
                          __uint_as_float(r[j + 3]) * scale);
        }
      }
    }
    fence_proxy_async_smem();
    // each thread adds its own row of the three accumulators; the rotation is undone by splitting the row in two runs
    for (int dx = 0; dx < 3; ++dx) {
      const float* srow = reinterpret_cast<const float*>(smem + dx * 65536) + row * 128;
      float* grow = dw + ((long long)(dy * 3 + dx) * C + ci0 + row) * C + co0;
      const int rot = (row & 31) << 2;           // element offset where column 0 of this row sits
      bulk_reduce_add_f32(grow, srow + rot, (uint32_t)((128 - rot) * 4));
      if (rot) bulk_reduce_add_f32(grow + (128 - rot), srow, (uint32_t)(rot * 4));
    }
    tma_store_commit();
    tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace dsen2

using namespace dsen2;

extern "C" int dsen2_wgrad_nhwc(const void* d_x, const void* d_dy, int n, int H, int W, int channels, float scale, float* d_dw,
                                float* d_db, void* stream) {
  DSEN2_REQUIRE(d_x && d_dy && d_dw, DSEN2_E_BADARG, "dsen2_wgrad_nhwc: null pointer");
  DSEN2_REQUIRE(n > 0 && H > 0 && W > 0, DSEN2_E_BADARG, "dsen2_wgrad_nhwc: bad shape");
  DSEN2_REQUIRE(channels == 128 || channels == 256, DSEN2_E_BADARG, "dsen2_wgrad_nhwc: 128 or 256 channels (got %d)", channels);
  DSEN2_REQUIRE(((uintptr_t)d_x % 16) == 0 && ((uintptr_t)d_dy % 16) == 0 && ((uintptr_t)d_dw % 16) == 0, DSEN2_E_ALIGN,
                "dsen2_wgrad_nhwc: pointers must be 16-byte aligned");
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  const int tiles_x = ceil_div(W, 8), tiles_y = ceil_div(H, 16);
  const long long tiles = (long long)n * tiles_x * tiles_y;
  DSEN2_REQUIRE(tiles < (1LL << 30), DSEN2_E_BADARG, "dsen2_wgrad_nhwc: batch too large");
  const int blocks = (channels / 128) * (channels / 128);       // 128 x 128 blocks of the (C, C) gradient
  int splits = sms / (3 * blocks);
  if (splits > tiles) splits = (int)tiles;
  const int per = (int)((tiles + splits - 1) / splits);
  splits = (int)((tiles + per - 1) / per);
  CUtensorMap tx, ty;
  const uint64_t dims[4] = {(uint64_t)channels, (uint64_t)W, (uint64_t)H, (uint64_t)n};
  const uint32_t bx[4] = {64, 10, 16, 1};
  rc = make_tmap_f16_sw(&tx, d_x, 4, dims, bx, 128);
  if (rc) return rc;
  const uint32_t by[4] = {64, 8, 16, 1};
  rc = make_tmap_f16_sw(&ty, d_dy, 4, dims, by, 128);
  if (rc) return rc;
  constexpr int SMEM = kWdStages * kWdStage + 256 + 1024;
  static_assert(kWdStages * kWdStage >= 3 * 65536, "the drain reuses the pipeline stages for three 64 KB accumulators");
  static bool configured[64] = {};
  if (needs_config(configured)) {
    DSEN2_CUDA(cudaFuncSetAttribute(wgrad_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  }
  wgrad_direct_kernel<<<dim3(splits, 3, blocks), kWdThreads, SMEM, (cudaStream_t)stream>>>(tx, ty, tiles_x, tiles_y, (int)tiles,
                                                                                         per, channels, scale, d_dw, d_db);
  return check_launch("wgrad_direct_kernel");
}

extern "C" int dsen2_nchw_to_nhwc_f16(const float* d_x0, int c0, const float* d_x1, int c1, const float* d_x2, int c2,
                                      int n, int H, int W, int cpad, void* d_out, void* stream) {
  DSEN2_REQUIRE(d_x0 && d_out && (c1 == 0 || d_x1) && (c2 == 0 || d_x2) && c0 > 0 && c1 >= 0 && c2 >= 0 &&
                    c0 + c1 + c2 <= 16 && cpad >= 16 && cpad % 8 == 0 && n > 0 && H > 0 && W > 0,
                DSEN2_E_BADARG, "dsen2_nchw_to_nhwc_f16: bad arguments (at most 16 channels, cpad a multiple of 8 >= 16)");
  DSEN2_REQUIRE(((uintptr_t)d_out % 16) == 0, DSEN2_E_ALIGN, "dsen2_nchw_to_nhwc_f16: output must be 16-byte aligned");
  const long long npix = (long long)n * H * W;
  if (cpad > 16) DSEN2_CUDA(cudaMemsetAsync(d_out, 0, (size_t)npix * cpad * 2, (cudaStream_t)stream));
  nchw_to_nhwc_f16_kernel<<<grid_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(d_x0, c0, d_x1, c1, d_x2, c2, (long long)H * W,
                                                                                 cpad, npix, (__half*)d_out);
  return check_launch("nchw_to_nhwc_f16");
}

extern "C" int dsen2_relu_mask(const void* d_in, const void* d_act, long long total, void* d_out, void* stream) {
  DSEN2_REQUIRE(d_in && d_act && d_out && total > 0 && total % 8 == 0, DSEN2_E_BADARG, "dsen2_relu_mask: bad arguments");
  relu_mask_kernel<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)d_in, (const uint4*)d_act, total / 8,
                                                                              (uint4*)d_out);
  return check_launch("relu_mask");
}

extern "C" int dsen2_mae_grad(const float* d_pred, const float* d_y, long long total, float gscale, float* d_dpred,
                              double* d_sums, void* stream) {
  DSEN2_REQUIRE(d_pred && d_y && d_dpred && d_sums && total > 0, DSEN2_E_BADARG, "dsen2_mae_grad: bad arguments");
  mae_grad_kernel<<<grid_for(total, 256, 4), 256, 0, (cudaStream_t)stream>>>(d_pred, d_y, total, gscale, d_dpred, d_sums);
  return check_launch("mae_grad");
}

static long long nadam_blocks(long long total) {
  const long long want = ((total >> 2) + 255) / 256, cap = (long long)sm_count() * 8;
  return want < 1 ? 1 : (want < cap ? want : cap);
}

extern "C" int dsen2_nadam_step(float* d_p, const float* d_g, float* d_m, float* d_v, long long total, float grad_mul,
                                float lr, float beta1, float beta2, float eps, float mu_t, float mu_next,
                                float sched_new, float sched_next, float bias2, void* stream) {
  DSEN2_REQUIRE(d_p && d_g && d_m && d_v && total > 0, DSEN2_E_BADARG, "dsen2_nadam_step: bad arguments");
  DSEN2_REQUIRE(((uintptr_t)d_p % 16) == 0 && ((uintptr_t)d_g % 16) == 0 && ((uintptr_t)d_m % 16) == 0 && ((uintptr_t)d_v % 16) == 0,
                DSEN2_E_ALIGN, "dsen2_nadam_step: pointers must be 16-byte aligned");
  const NadamHp h = {grad_mul, lr, beta1, beta2, eps, mu_t, mu_next, sched_new, sched_next, bias2};
  nadam_kernel<<<(unsigned)nadam_blocks(total), 256, 0, (cudaStream_t)stream>>>(d_p, d_g, d_m, d_v, total, h);
  return check_launch("nadam_step");
}

extern "C" int dsen2_nadam_step_dev(float* d_p, const float* d_g, float* d_m, float* d_v, long long total,
                                    const float* d_hp, void* stream) {
  DSEN2_REQUIRE(d_p && d_g && d_m && d_v && d_hp && total > 0, DSEN2_E_BADARG, "dsen2_nadam_step_dev: bad arguments");
  DSEN2_REQUIRE(((uintptr_t)d_p % 16) == 0 && ((uintptr_t)d_g % 16) == 0 && ((uintptr_t)d_m % 16) == 0 && ((uintptr_t)d_v % 16) == 0,
                DSEN2_E_ALIGN, "dsen2_nadam_step_dev: pointers must be 16-byte aligned");
  nadam_dev_kernel<<<(unsigned)nadam_blocks(total), 256, 0, (cudaStream_t)stream>>>(d_p, d_g, d_m, d_v, total, d_hp);
  return check_launch("nadam_step_dev");
}

extern "C" int dsen2_pack_dgrad_weights(const float* d_hwio, int cin, int cout, int rows_pad, int k_pad, float scale,
                                        void* d_packed, void* stream) {
  DSEN2_REQUIRE(d_hwio && d_packed && cin > 0 && cout > 0 && rows_pad >= cin && k_pad >= cout && k_pad % 64 == 0,
                DSEN2_E_BADARG, "dsen2_pack_dgrad_weights: bad arguments");
  const long long total = 9LL * rows_pad * k_pad;
  pack_dgrad_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d_hwio, cin, cout, rows_pad, k_pad, scale,
                                                                                   total, (__half*)d_packed);
  return check_launch("pack_dgrad_weights");
}


extern "C" int dsen2_pack_trunk_layers(const float* d_first_kernel, long long layer_stride, int num_layers, int feature_size,
                                       float scale_second, void* d_fwd, void* d_bwd, void* stream) {
  DSEN2_REQUIRE(d_first_kernel && d_fwd && d_bwd, DSEN2_E_BADARG, "dsen2_pack_trunk_layers: null pointer");
  DSEN2_REQUIRE(num_layers >= 0 && num_layers < 65536 && feature_size > 0 && feature_size % 64 == 0 &&
                    layer_stride >= 9LL * feature_size * feature_size,
                DSEN2_E_BADARG, "dsen2_pack_trunk_layers: bad arguments");
  if (num_layers == 0) return 0;
  const int tiles = feature_size / 32;
  const dim3 grid((unsigned)(tiles * tiles), 9, (unsigned)num_layers);
  pack_trunk_layers_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_first_kernel, layer_stride, feature_size, scale_second,
                                                                   (__half*)d_fwd, (__half*)d_bwd);
  return check_launch("pack_trunk_layers");
}
