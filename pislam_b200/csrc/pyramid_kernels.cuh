// Pyramid-preparation helpers -- not on the ORB hot path: the reference expects
// its caller to blur and build the pyramid (README.md:28-31) and ships these three building blocks:
//   gaussian5x5    Gaussian.h:48-72     separable [1 4 6 4 1]/16 from rounding halving adds, reflect-101 borders,
//                                       vertical pass first (semantics: test/GaussianTest.cpp:159-215)
//   bilinear7_8    Bilinear.h:42-150    8x8 -> 7x7 blocks, weights /256, rounding shifts
//   bilinear13_16  Bilinear.h:165-321   16x16 -> 13x13 blocks
// Byte-in/byte-out streaming kernels: four pixels per thread for the blur (per-byte rounding average in SWAR),
// one output pixel per thread for the block resamplers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pb {

// per-byte rounding average (a + b + 1) >> 1 -- NEON vrhadd.u8
__device__ __forceinline__ uint32_t rhadd4(uint32_t a, uint32_t b) { return __vavgu4(a, b); }
__device__ __forceinline__ uint32_t gauss5_4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) {
  uint32_t x = rhadd4(a, e);
  x = rhadd4(x, c);
  x = rhadd4(x, c);
  return rhadd4(x, rhadd4(b, d));
}
__device__ __forceinline__ int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// vertical pass: tmp[y][x..x+3] from rows y-2..y+2 of img (reflect-101); tmp has pitch tpitch (multiple of 4)
__global__ void gaussian_vertical_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ tmp, int vstep,
                                         int tpitch, int width, int height) {
  const int xw = blockIdx.x * blockDim.x + threadIdx.x;  // word column
  const int y = blockIdx.y;
  if (4 * xw >= width) return;
  uint32_t r[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint8_t* p = img + (size_t)reflect101(y + k - 2, height) * vstep + 4 * xw;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (4 * xw + b < width) v |= (uint32_t)p[b] << (8 * b);
    r[k] = v;
  }
  *reinterpret_cast<uint32_t*>(tmp + (size_t)y * tpitch + 4 * xw) = gauss5_4(r[0], r[1], r[2], r[3], r[4]);
}

// horizontal pass: out[y][x..x+3] from tmp[y][x-2..x+5] (reflect-101 at the row ends)
__global__ void gaussian_horizontal_kernel(const uint8_t* __restrict__ tmp, uint8_t* __restrict__ out, int vstep,
                                           int tpitch, int width, int height) {
  const int xw = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int x0 = 4 * xw;
  if (x0 >= width) return;
  const uint8_t* row = tmp + (size_t)y * tpitch;
  uint32_t t[5] = {0, 0, 0, 0, 0};  // t[k] byte b = tmp[y][x0 + b + k - 2]
  if (x0 >= 4 && x0 + 8 <= width) {   // interior: three aligned words and funnel shifts
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row + x0);
    const uint32_t l = w[-1], c = w[0], rr = w[1];
    t[0] = __funnelshift_r(l, c, 16);
    t[1] = __funnelshift_r(l, c, 24);
    t[2] = c;
    t[3] = __funnelshift_r(c, rr, 8);
    t[4] = __funnelshift_r(c, rr, 16);
  } else {
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (x0 + b < width) t[k] |= (uint32_t)row[reflect101(x0 + b + k - 2, width)] << (8 * b);
  }
  const uint32_t v = gauss5_4(t[0], t[1], t[2], t[3], t[4]);
  uint8_t* o = out + (size_t)y * vstep + x0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (x0 + b < width) o[b] = (uint8_t)(v >> (8 * b));
}

__device__ __forceinline__ int rshr8(int a) { return (a + 128) >> 8; }  // VRSHR #8 on a value < 2^16

// one thread per output pixel of the O x O blocks; MAP13 inserts the two skipped source rows/columns
template <int B, int O, bool MAP13>
__global__ void bilinear_blocks_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int vstep, int nbx,
                                       int nby, int zero_col) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  if (oy >= O * nby) return;
  if (ox == O * nbx && zero_col) {  // bilinear7_8 stores 8 bytes per block row: lane 7 is always 0
    out[(size_t)oy * vstep + ox] = 0;
    return;
  }
  if (ox >= O * nbx) return;
  const int bj = ox / O, x = ox - bj * O, bi = oy / O, y = oy - bi * O;
  int fx, fxr, fy, fyr, sx = x, sy = y;
  if (MAP13) {
    const int f[13] = {226, 167, 108, 49, 246, 187, 128, 69, 10, 207, 138, 89, 30};
    fx = f[x]; fxr = f[12 - x]; fy = f[y]; fyr = f[12 - y];
    sx = x + (x > 3) + (x > 8);
    sy = y + (y > 3) + (y > 8);
  } else {
    const int f[7] = {238, 201, 165, 128, 91, 55, 18};
    fx = f[x]; fxr = f[6 - x]; fy = f[y]; fyr = f[6 - y];
  }
  const uint8_t* p = img + (size_t)(bi * B + sy) * vstep + (bj * B + sx);
  const int h0 = rshr8(p[0] * fx + p[1] * fxr);
  const int h1 = rshr8(p[vstep] * fx + p[vstep + 1] * fxr);
  out[(size_t)oy * vstep + ox] = (uint8_t)rshr8(h0 * fy + h1 * fyr);
}

}  // namespace pb
