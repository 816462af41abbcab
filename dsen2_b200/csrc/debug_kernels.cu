// Self-test kernels (tests only): one UMMA tile with a row-shifted A descriptor.  Answers the
// question the halo-reuse conv design depends on: can a 128B-swizzled K-major operand start at an
// arbitrary 128-byte row inside a TMA-written block (swizzle taken from absolute smem address bits)?
#include "common.cuh"

namespace dsen2 {

static constexpr int kDbgRows = 160;   // A block rows held in smem (>= 128 + max shift)

__global__ void __launch_bounds__(128, 1)
umma_rowshift_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int shift_rows,
                     int base_offset_mode, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                              // kDbgRows x 128 B
  uint8_t* sb = smem + kDbgRows * 128;             // 128 x 128 B (kDbgRows*128 is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 128 * 128);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], kDbgRows * 128 + 128 * 128);
    tma_load_2d(sa, &tm_a, &bars[0], 0, 0);
    tma_load_2d(sb, &tm_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sa) + shift_rows * 128;
    const uint32_t bo = base_offset_mode ? ((a0 >> 7) & 7) : 0;
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16_ss(tmem_base, umma_desc_sw128(a0 + k * 32, bo), umma_desc_sw128(smem_u32(sb) + k * 32), idesc, k != 0);
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[row * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

}  // namespace dsen2

using namespace dsen2;

extern "C" int dsen2_debug_umma_rowshift(const void* d_a_f16, int rows, const void* d_b_f16, int shift_rows,
                                         int base_offset_mode, float* d_out, void* stream) {
  DSEN2_REQUIRE(d_a_f16 && d_b_f16 && d_out, DSEN2_E_BADARG, "dsen2_debug_umma_rowshift: null pointer");
  DSEN2_REQUIRE(rows >= 128 && shift_rows >= 0 && shift_rows + 128 <= kDbgRows, DSEN2_E_BADARG,
                "dsen2_debug_umma_rowshift: need rows >= 128 and shift_rows + 128 <= %d", kDbgRows);
  int sms = 0;
  int rc = device_sm_count_and_check(&sms);
  if (rc) return rc;
  CUtensorMap tm_a, tm_b;
  const uint64_t da[2] = {64, (uint64_t)rows};
  const uint32_t ba[2] = {64, (uint32_t)kDbgRows};   // rows beyond `rows` are zero-filled
  rc = make_tmap_f16(&tm_a, d_a_f16, 2, da, ba);
  if (rc) return rc;
  const uint64_t db[2] = {64, 128};
  const uint32_t bb[2] = {64, 128};
  rc = make_tmap_f16(&tm_b, d_b_f16, 2, db, bb);
  if (rc) return rc;
  const int smem_bytes = kDbgRows * 128 + 128 * 128 + 64 + 1024;
  umma_rowshift_kernel<<<1, 128, smem_bytes, (cudaStream_t)stream>>>(tm_a, tm_b, shift_rows, base_offset_mode, d_out);
  return check_launch("umma_rowshift_kernel");
}
