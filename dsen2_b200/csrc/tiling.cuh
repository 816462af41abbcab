// Tiling / resampling index arithmetic shared by the patch, prep and stitch kernels.
#pragma once
#include "common.cuh"

namespace dsen2 {

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

inline dim3 grid_for(long long work_items, int block, int max_waves = 16) {
  long long blocks = (work_items + block - 1) / block;
  long long cap = (long long)sm_count() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return dim3((unsigned)blocks);
}

// ------------------------------------------------------------------------------------------ //
// tiling arithmetic shared by host and device (patches.py:32-53)
// ------------------------------------------------------------------------------------------ //
struct Tiling {
  int k_i, k_j;    // full strides per axis
  int n_i, n_j;    // filled crop starts per axis
  int stride;      // patch_lr - 2*border_lr
  int last_i, last_j;  // clamped start (padded lr coords) of the extra crop, if any
};

__host__ __device__ inline Tiling make_tiling(int grid_h, int grid_w, int patch_lr, int border_lr) {
  Tiling t;
  t.stride = patch_lr - 2 * border_lr;
  t.k_i = grid_h / t.stride;
  t.k_j = grid_w / t.stride;
  t.n_i = t.k_i + (grid_h % t.stride != 0);
  t.n_j = t.k_j + (grid_w % t.stride != 0);
  t.last_i = grid_h + 2 * border_lr - patch_lr;
  t.last_j = grid_w + 2 * border_lr - patch_lr;
  return t;
}

// ------------------------------------------------------------------------------------------ //
// IEEE division by the two constants of the path -- SCALE = 2000 (supres.py:11,23-24) and the 30000 of interp_patches
// (patches.py:15) -- without the division sequence (MUFU.RCP + 4 FFMA + FCHK + slow-path branch, ~10 issue slots each;
// the preparation kernel does 34 of them per pixel).  With rc = RN(1 / c):
//     q0 = RN(x * rc);   r = x - c * q0  (exact in one FMA);   q = RN(q0 + r * rc)
// is the correctly rounded quotient for every x whose quotient is a normal number; that is verified EXHAUSTIVELY, all 2^32
// bit patterns of x against __fdiv_rn, for exactly these two constants (dsen2_debug_divconst_mismatches,
// tests/test_gpu_patches.py) -- any other divisor takes __fdiv_rn.  Outside [kDivLo, kDivHi] (quotients that are denormal,
// or x so large that c * q0 overflows) the slow path is kept.
// ------------------------------------------------------------------------------------------ //
struct DivC {
  float c, rc;
  int fast;
};
static constexpr float kDivLo = 1e-30f, kDivHi = 1e30f;
inline DivC make_divc(float c) {
  DivC d;
  d.c = c;
  d.rc = 1.0f / c;
  d.fast = (c == 2000.0f || c == 30000.0f) ? 1 : 0;
  return d;
}
#ifdef __CUDACC__
__device__ __forceinline__ float div_fast(float x, float c, float rc) {
  const float q0 = x * rc;
  const float r = fmaf(-c, q0, x);
  return fmaf(r, rc, q0);
}
__device__ __forceinline__ float div_c(float x, const DivC& d) {
  const float ax = fabsf(x);
  if (d.fast && ax >= kDivLo && ax <= kDivHi) return div_fast(x, d.c, d.rc);
  return __fdiv_rn(x, d.c);          // zeros (the FMA pair would turn -0 into +0), denormal quotients, huge x, other divisors
}
#endif

__device__ __forceinline__ int sym_index(int j, int n) {  // numpy pad(mode='symmetric')
  if (j < 0) j = -j - 1;
  if (j >= n) j = 2 * n - 1 - j;
  return j;
}

__device__ __forceinline__ void bilin_tap(int o, int s, int n, int& a0, int& a1, float& f) {
  const int t = 2 * o + 1 - s;               // u = t / (2 s)
  int i0 = (t >= 0) ? t / (2 * s) : -((-t + 2 * s - 1) / (2 * s));
  f = (float)(t - i0 * 2 * s) / (float)(2 * s);
  int i1 = i0 + 1;
  if (i0 < 0) i0 = -i0;                      // mirror: -1 -> 1
  if (i1 > n - 1) i1 = 2 * (n - 1) - i1;     // mirror:  n -> n-2
  if (i1 < 0) i1 = 0;                        // n == 1
  a0 = i0; a1 = i1;
}

__device__ __forceinline__ int tile_of(int y, int size, int S, int n) {
  return (size % S != 0 && y >= size - S) ? n - 1 : y / S;
}


}  // namespace dsen2
