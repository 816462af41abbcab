// 3x3 / 1x1 convolution over NHWC fp16 patches as a tcgen05 implicit GEMM (sm_100a).
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[tap, cout, cin]
//
//   M = 128 pixels of one patch (a BW x BH box, BW*BH = 128), N = cout_pad, K = taps * cin_pad.
//   A tiles come straight from the NHWC activation tensor through a 4-D TMA tensor map
//   {C, W, H, n}: the box for tap (dy,dx) is loaded at (x0+dx-1, y0+dy-1) and TMA's out-of-bounds
//   zero fill IS the Conv2D(padding='same') zero border of the patch (DSen2Net.py:10,12,29,35) --
//   neighbouring patches are never touched because the patch index is its own tensor dimension.
//   B tiles come from the packed [tap][cout_pad][cin_pad] weights.  Both land 128B-swizzled,
//   K-major, and feed tcgen05.mma.cta_group::1.kind::f16 with fp32 accumulators in TMEM.
//
//   Warp roles (256 threads, persistent over tiles): warp 0 = TMA producer, warp 1 = MMA issuer,
//   warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> global).  Two accumulator
//   buffers in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
#include "common.cuh"

namespace dsen2 {

static constexpr int kTileM = 128;
static constexpr int kBlockK = 64;                 // fp16 elements per 128-byte swizzle row
static constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KB
static constexpr int kThreads = 256;

struct ConvParams {
  int n, H, W;
  int taps;        // 1 or 9
  int kpc;         // cin_pad / 64 (k-blocks per tap)
  int bw, bh;      // tile box, bw*bh == 128
  int tiles_x, tiles_y;
  long long num_tiles;
  int cout_pad;    // row pitch (channels) of the NHWC outputs / residuals
  int cout_real;   // TAIL only
  const float* bias;
  const __half* res_hi;
  const __half* res_lo;
  float res_scale;
  __half* out_hi;
  __half* out_lo;
  const float* skip;
  float* out_f32;
};

template <int N>
struct ConvCfg {
  static constexpr int kBBytes = N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + (kBBytes < 1024 ? 1024 : kBBytes);
  static constexpr int kStages = (N == 256) ? 4 : (N == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * N < 32) ? 32 : 2 * N;   // two accumulators
  static constexpr int kBarBytes = 2048;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024 /*align slack*/;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int N, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                    const ConvParams p) {
  using Cfg = ConvCfg<N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_bias = reinterpret_cast<float*>(bar_base + 512);   // N <= 256 floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb_per_tile = p.taps * p.kpc;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);   // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
  for (int i = threadIdx.x; i < N; i += kThreads) s_bias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % p.tiles_x);
        const int ty = (int)((tile / p.tiles_x) % p.tiles_y);
        const int b = (int)(tile / ((long long)p.tiles_x * p.tiles_y));
        const int x0 = tx * p.bw, y0 = ty * p.bh;
        for (int kb = 0; kb < kb_per_tile; ++kb) {
          const int tap = kb / p.kpc, kc = kb - tap * p.kpc;
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full[stage], kABytes + Cfg::kBBytes);
          tma_load_4d(sa, &tm_in, &full[stage], kc * kBlockK, x0 + dx, y0 + dy, b);
          tma_load_3d(sb, &tm_w, &full[stage], kc * kBlockK, 0, tap);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
        for (int kb = 0; kb < kb_per_tile; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_f16_ss(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc,
                        (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty[stage]);            // frees the smem slot once these MMAs have read it
          if (kb == kb_per_tile - 1) umma_commit(&tmem_full[acc]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int wq = warp & 3;                     // TMEM lane quarter this warp may read
    const int row = wq * 32 + lane;              // pixel row of the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = (int)(tile % p.tiles_x);
      const int ty = (int)((tile / p.tiles_x) % p.tiles_y);
      const int b = (int)(tile / ((long long)p.tiles_x * p.tiles_y));
      const int y = ty * p.bh + row / p.bw;
      const int x = tx * p.bw + row % p.bw;
      const bool valid = (y < p.H) && (x < p.W);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * N);
      if (EPI == DSEN2_EPI_TAIL_NCHW) {
        uint32_t r[16];
        tmem_ld_32x16(taddr, r);
        tmem_ld_wait();
        if (valid) {
          const long long plane = (long long)p.H * p.W;
          const long long o = (long long)b * p.cout_real * plane + (long long)y * p.W + x;
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < p.cout_real)
              p.out_f32[o + c * plane] = (__uint_as_float(r[c]) + s_bias[c]) + __ldg(p.skip + o + c * plane);
        }
      } else {
        const long long pix = ((long long)b * p.H + y) * p.W + x;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c0, r);
          tmem_ld_wait();
          if (valid) {
            const long long off = pix * p.cout_pad + c0;
            uint4 hi4[4], lo4[4];
            if (EPI == DSEN2_EPI_RESIDUAL) {
              const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + off);
              const uint4* rl = reinterpret_cast<const uint4*>(p.res_lo + off);
#pragma unroll
              for (int q = 0; q < 4; ++q) { hi4[q] = rh[q]; lo4[q] = rl[q]; }
            }
            uint32_t* hw = reinterpret_cast<uint32_t*>(hi4);
            uint32_t* lw = reinterpret_cast<uint32_t*>(lo4);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float v0 = __uint_as_float(r[j]) + s_bias[c0 + j];
              float v1 = __uint_as_float(r[j + 1]) + s_bias[c0 + j + 1];
              if (EPI == DSEN2_EPI_RESIDUAL) {
                const float2 xh = __half22float2(*reinterpret_cast<const __half2*>(&hw[j >> 1]));
                const float2 xl = __half22float2(*reinterpret_cast<const __half2*>(&lw[j >> 1]));
                v0 = (xh.x + xl.x) + v0 * p.res_scale;   // Lambda(x*scale) then Add (DSen2Net.py:13,15)
                v1 = (xh.y + xl.y) + v1 * p.res_scale;
              } else {
                v0 = fmaxf(v0, 0.f);
                v1 = fmaxf(v1, 0.f);
              }
              const __half2 h = __floats2half2_rn(v0, v1);
              const float2 hf = __half22float2(h);
              hw[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
              lw[j >> 1] = pack_h2(v0 - hf.x, v1 - hf.y);
            }
            uint4* oh = reinterpret_cast<uint4*>(p.out_hi + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) oh[q] = hi4[q];
            if (p.out_lo != nullptr) {
              uint4* ol = reinterpret_cast<uint4*>(p.out_lo + off);
#pragma unroll
              for (int q = 0; q < 4; ++q) ol[q] = lo4[q];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------------ //
// host side
// ------------------------------------------------------------------------------------------ //
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_tmap_f16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box) {
  return make_tmap_f16_sw(m, ptr, rank, dims, box, 128);
}

int make_tmap_f16_sw(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  DSEN2_REQUIRE(enc != nullptr, DSEN2_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  uint64_t stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                       : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSEN2_REQUIRE(r == CUDA_SUCCESS, DSEN2_E_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

int device_sm_count_and_check(int* sms) {
  static int cached_sms = 0, cached_major = 0;
  if (cached_sms == 0) {
    int dev = 0;
    DSEN2_CUDA(cudaGetDevice(&dev));
    DSEN2_CUDA(cudaDeviceGetAttribute(&cached_major, cudaDevAttrComputeCapabilityMajor, dev));
    DSEN2_CUDA(cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  DSEN2_REQUIRE(cached_major == 10, DSEN2_E_NOTSM100, "dsen2_b200 needs an sm_100 device (found compute capability %d.x)",
                cached_major);
  *sms = cached_sms;
  return 0;
}

static void pick_box(int H, int W, int* bw, int* bh) {
  long long best = -1;
  for (int w = 128; w >= 8; w >>= 1) {
    const int h = kTileM / w;
    const long long tiles = (long long)ceil_div(W, w) * ceil_div(H, h);
    if (best < 0 || tiles < best) { best = tiles; *bw = w; *bh = h; }
  }
}

template <int N, int EPI>
static int launch_conv(const CUtensorMap& tm_in, const CUtensorMap& tm_w, const ConvParams& p, int sms,
                       cudaStream_t stream) {
  using Cfg = ConvCfg<N>;
  static bool configured[64] = {};
  if (needs_config(configured)) {
    DSEN2_CUDA(cudaFuncSetAttribute(conv_tcgen05_kernel<N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kSmemBytes));
  }
  const int grid = (int)(p.num_tiles < sms ? p.num_tiles : sms);
  conv_tcgen05_kernel<N, EPI><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(tm_in, tm_w, p);
  return check_launch("conv_tcgen05_kernel");
}

}  // namespace dsen2

using namespace dsen2;

static bool g_force_v1 = false;
extern "C" int dsen2_debug_force_v1(int on) {   // tests / A-B timing: route 128-feature layers to the single-CTA kernel
  g_force_v1 = on != 0;
  return 0;
}

extern "C" int dsen2_conv3x3(const void* d_in, const void* d_w, const float* d_bias, int n, int H, int W, int cin_pad,
                             int cout_pad, int taps, int epilogue, const void* d_res_hi, const void* d_res_lo,
                             float res_scale, void* d_out_hi, void* d_out_lo, const float* d_skip_f32,
                             float* d_out_f32, int cout_real, void* stream) {
  DSEN2_REQUIRE(d_in && d_w && d_bias, DSEN2_E_BADARG, "dsen2_conv3x3: null input/weights/bias");
  DSEN2_REQUIRE(n >= 0 && H > 0 && W > 0 && cin_pad > 0 && cin_pad % 64 == 0 && (taps == 1 || taps == 9),
                DSEN2_E_BADARG, "dsen2_conv3x3: bad shape (n %d, %dx%d, cin_pad %d, taps %d)", n, H, W, cin_pad, taps);
  DSEN2_REQUIRE(((uintptr_t)d_in % 16) == 0 && ((uintptr_t)d_w % 16) == 0, DSEN2_E_ALIGN,
                "dsen2_conv3x3: input / weights must be 16-byte aligned");
  if (n == 0) return 0;
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;

  ConvParams p{};
  p.n = n; p.H = H; p.W = W; p.taps = taps; p.kpc = cin_pad / kBlockK;
  pick_box(H, W, &p.bw, &p.bh);
  p.tiles_x = ceil_div(W, p.bw);
  p.tiles_y = ceil_div(H, p.bh);
  p.num_tiles = (long long)n * p.tiles_x * p.tiles_y;
  p.cout_pad = cout_pad; p.cout_real = cout_real; p.bias = d_bias;
  p.res_hi = (const __half*)d_res_hi; p.res_lo = (const __half*)d_res_lo; p.res_scale = res_scale;
  p.out_hi = (__half*)d_out_hi; p.out_lo = (__half*)d_out_lo; p.skip = d_skip_f32; p.out_f32 = d_out_f32;

  CUtensorMap tm_in, tm_w;
  {
    const uint64_t dims[4] = {(uint64_t)cin_pad, (uint64_t)W, (uint64_t)H, (uint64_t)n};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    rc = make_tmap_f16(&tm_in, d_in, 4, dims, box);
    if (rc) return rc;
    const uint64_t wd[3] = {(uint64_t)cin_pad, (uint64_t)cout_pad, (uint64_t)taps};
    const uint32_t wb[3] = {(uint32_t)kBlockK, (uint32_t)cout_pad, 1};
    rc = make_tmap_f16(&tm_w, d_w, 3, wd, wb);
    if (rc) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (epilogue == DSEN2_EPI_TAIL_NCHW) {
    DSEN2_REQUIRE(cout_pad == 16 && cout_real > 0 && cout_real <= 16 && d_skip_f32 && d_out_f32, DSEN2_E_BADARG,
                  "dsen2_conv3x3: TAIL epilogue needs cout_pad 16, 0 < cout_real <= 16, skip and out pointers");
    return launch_conv<16, DSEN2_EPI_TAIL_NCHW>(tm_in, tm_w, p, sms, s);
  }
  DSEN2_REQUIRE(cout_pad == 128 || cout_pad == 256, DSEN2_E_BADARG,
                "dsen2_conv3x3: feature size must be 128 or 256 (got %d)", cout_pad);
  DSEN2_REQUIRE(d_out_hi && ((uintptr_t)d_out_hi % 16) == 0 && ((uintptr_t)d_out_lo % 16) == 0, DSEN2_E_ALIGN,
                "dsen2_conv3x3: out_hi must be non-null and outputs 16-byte aligned");
  if ((cout_pad == 128 || cout_pad == 256) && cin_pad == cout_pad && taps == 9 &&
      (epilogue == DSEN2_EPI_RELU || epilogue == DSEN2_EPI_RESIDUAL) && !g_force_v1) {
    // trunk layers: CTA-pair kernel with halo-box activations; weights resident (128) or streamed (256) (conv_pair.cu)
    if (epilogue == DSEN2_EPI_RESIDUAL) {
      DSEN2_REQUIRE(d_res_hi && d_res_lo && d_out_lo, DSEN2_E_BADARG,
                    "dsen2_conv3x3: RESIDUAL epilogue needs res_hi, res_lo and out_lo");
      DSEN2_REQUIRE(((uintptr_t)d_res_hi % 16) == 0 && ((uintptr_t)d_res_lo % 16) == 0, DSEN2_E_ALIGN,
                    "dsen2_conv3x3: residual pointers must be 16-byte aligned");
    }
    DSEN2_REQUIRE(d_out_hi && ((uintptr_t)d_out_hi % 16) == 0 && ((uintptr_t)d_out_lo % 16) == 0, DSEN2_E_ALIGN,
                  "dsen2_conv3x3: out_hi must be non-null and outputs 16-byte aligned");
    return conv_pair_res(d_in, d_w, d_bias, n, H, W, cout_pad, epilogue, d_res_hi, d_res_lo, res_scale, d_out_hi, d_out_lo, s);
  }
  if (epilogue == DSEN2_EPI_RELU) {
    return cout_pad == 128 ? launch_conv<128, DSEN2_EPI_RELU>(tm_in, tm_w, p, sms, s)
                           : launch_conv<256, DSEN2_EPI_RELU>(tm_in, tm_w, p, sms, s);
  }
  if (epilogue == DSEN2_EPI_RESIDUAL) {
    DSEN2_REQUIRE(d_res_hi && d_res_lo && d_out_lo, DSEN2_E_BADARG,
                  "dsen2_conv3x3: RESIDUAL epilogue needs res_hi, res_lo and out_lo");
    DSEN2_REQUIRE(((uintptr_t)d_res_hi % 16) == 0 && ((uintptr_t)d_res_lo % 16) == 0, DSEN2_E_ALIGN,
                  "dsen2_conv3x3: residual pointers must be 16-byte aligned");
    return cout_pad == 128 ? launch_conv<128, DSEN2_EPI_RESIDUAL>(tm_in, tm_w, p, sms, s)
                           : launch_conv<256, DSEN2_EPI_RESIDUAL>(tm_in, tm_w, p, sms, s);
  }
  DSEN2_REQUIRE(false, DSEN2_E_BADARG, "dsen2_conv3x3: unknown epilogue %d", epilogue);
}

// ------------------------------------------------------------------------------------------ //
// whole-network forward (model.predict on one batch)
// ------------------------------------------------------------------------------------------ //
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t dsen2_s2model_workspace_bytes(int n, int P, int in_channels, int feature_size) {
  if (n <= 0 || P <= 0 || in_channels <= 0 || feature_size <= 0) return 0;
  const size_t pix = (size_t)n * P * P;
  // x_in hi/lo (64 ch; 16 are used when there are resblocks) + trunk hi/lo + resblock intermediate + low bytes of the
  // fp16 + 8 bit trunk
  return 2 * align_up(pix * 64 * 2, 1024) + 3 * align_up(pix * feature_size * 2, 1024) +
         align_up((size_t)n * P * ((P + 7) / 8 * 8) * feature_size, 1024) + 1024;
}

extern "C" int dsen2_s2model_forward(const float* const* d_x, const int* channels, int n_inputs, int n, int P,
                                     int num_layers, int feature_size, const void* const* d_weights,
                                     const float* const* d_bias, void* d_workspace, size_t workspace_bytes,
                                     float* d_out_f32, void* stream) {
  DSEN2_REQUIRE(d_x && channels && d_weights && d_bias && d_workspace && d_out_f32, DSEN2_E_BADARG,
                "dsen2_s2model_forward: null pointer");
  DSEN2_REQUIRE(n_inputs == 2 || n_inputs == 3, DSEN2_E_BADARG, "dsen2_s2model_forward: s2model takes 2 or 3 inputs");
  DSEN2_REQUIRE(n >= 0 && P > 0 && num_layers >= 0 && (feature_size == 128 || feature_size == 256), DSEN2_E_BADARG,
                "dsen2_s2model_forward: bad sizes (n %d, P %d, layers %d, features %d)", n, P, num_layers, feature_size);
  DSEN2_REQUIRE(num_layers > 0 || feature_size == 128, DSEN2_E_BADARG,
                "dsen2_s2model_forward: a network without resblocks is only served for feature_size 128");
  if (n == 0) return 0;
  int ctot = 0;
  for (int i = 0; i < n_inputs; ++i) {
    DSEN2_REQUIRE(d_x[i] && channels[i] > 0, DSEN2_E_BADARG, "dsen2_s2model_forward: bad input %d", i);
    ctot += channels[i];
  }
  const int cout_real = channels[n_inputs - 1];   // DSen2Net.py:35
  DSEN2_REQUIRE(cout_real <= 16 && ctot <= 16, DSEN2_E_BADARG, "dsen2_s2model_forward: at most 16 input / output bands");
  DSEN2_REQUIRE(workspace_bytes >= dsen2_s2model_workspace_bytes(n, P, ctot, feature_size), DSEN2_E_BADARG,
                "dsen2_s2model_forward: workspace too small");
  const size_t pix = (size_t)n * P * P;
  const int F = feature_size;
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up((size_t)(uintptr_t)d_workspace, 1024));
  int rc;

  // prepared input -> split-precision first layer -> CTA-pair trunk (fp16 + 8 bit) -> split-precision last layer + global skip
  void* xin_hi = ws;
  ws += align_up(pix * 64 * 2, 1024);
  void* xin_lo = ws;
  ws += align_up(pix * 64 * 2, 1024);
  void* x_hi = ws;
  ws += align_up(pix * F * 2, 1024);
  void* x_lo = ws;
  ws += align_up(pix * F * 2, 1024);
  void* t = ws;
  ws += align_up(pix * F * 2, 1024);
  void* xq = ws;               // low bytes of the fp16 + 8 bit trunk, tile-row-major (dsen2_conv_resq)
  const float* x2 = n_inputs == 3 ? d_x[2] : nullptr;
  const int c2 = n_inputs == 3 ? channels[2] : 0;
  if (num_layers == 0) {
    // no resblocks: the 64-channel form (three horizontal taps pre-gathered, dsen2_pack_head_weights), whose first layer
    // writes the x_hi + x_lo pair the last layer reads
    rc = dsen2_prep_from_patches(d_x[0], channels[0], d_x[1], channels[1], x2, c2, n, P, xin_hi, xin_lo, stream);
    if (rc) return rc;
    rc = dsen2_conv_head(xin_hi, xin_lo, d_weights[0], d_bias[0], n, P, P, 128, x_hi, x_lo, nullptr, stream);
    if (rc) return rc;
    return dsen2_conv_tail(x_hi, x_lo, d_weights[1], d_bias[1], xin_hi, xin_lo, ctot - cout_real, cout_real, n, P, P,
                           d_out_f32, stream);
  }
  // un-gathered 16-channel input, nine-tap first layer (weights from dsen2_pack_head16_weights)
  rc = dsen2_prep16_from_patches(d_x[0], channels[0], d_x[1], channels[1], x2, c2, n, P, xin_hi, xin_lo, stream);
  if (rc) return rc;
  rc = dsen2_conv_head16_q(xin_hi, xin_lo, d_weights[0], d_bias[0], n, P, P, F, x_hi, xq, stream);
  if (rc) return rc;
  for (int l = 0; l < num_layers; ++l) {
    rc = dsen2_conv3x3(x_hi, d_weights[1 + 2 * l], d_bias[1 + 2 * l], n, P, P, F, F, 9, DSEN2_EPI_RELU, nullptr, nullptr,
                       0.f, t, nullptr, nullptr, nullptr, 0, stream);
    if (rc) return rc;
    // the fp16 x_lo is only needed by the last layer: the last resblock writes it
    rc = (F == 256 ? dsen2_conv_resq256 : dsen2_conv_resq)(t, d_weights[2 + 2 * l], d_bias[2 + 2 * l], n, P, P, 0.1f, x_hi,
                                                           xq, l == num_layers - 1 ? x_lo : nullptr, stream);
    if (rc) return rc;
  }
  return dsen2_conv_tail16(x_hi, x_lo, d_weights[2 * num_layers + 1], d_bias[2 * num_layers + 1], xin_hi, xin_lo,
                           ctot - cout_real, cout_real, F, n, P, P, d_out_f32, stream);
}
