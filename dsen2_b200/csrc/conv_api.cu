// Host side shared by the convolution entry points: TMA tensor-map construction, device check, and the whole-network
// forward `dsen2_s2model_forward` (model.predict on one batch).  The kernels live in conv_pair.cu.
#include "common.cuh"

namespace dsen2 {

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

}  // namespace dsen2

using namespace dsen2;

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
    rc = dsen2_conv_relu(x_hi, d_weights[1 + 2 * l], d_bias[1 + 2 * l], n, P, P, F, t, stream);
    if (rc) return rc;
    // the fp16 x_lo is only needed by the last layer: the last resblock writes it
    rc = (F == 256 ? dsen2_conv_resq256 : dsen2_conv_resq)(t, d_weights[2 + 2 * l], d_bias[2 + 2 * l], n, P, P, 0.1f, x_hi,
                                                           xq, l == num_layers - 1 ? x_lo : nullptr, stream);
    if (rc) return rc;
  }
  return dsen2_conv_tail16(x_hi, x_lo, d_weights[2 * num_layers + 1], d_bias[2 * num_layers + 1], xin_hi, xin_lo,
                           ctot - cout_real, cout_real, F, n, P, P, d_out_f32, stream);
}
