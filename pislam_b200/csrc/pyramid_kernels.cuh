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

// --------------------------------------------------------------------------------------------------
// The x1/1.2 pyramid builder in front of the detector: base frames in, the reference's
// vertically stacked pyramid layout out (README.md:56-83), for a whole batch in one launch.
// NOT reference code -- the reference leaves the pyramid to its caller (README.md:28-31).  Definition
// (oracle/pislam_oracle.h:po_resize_bilinear): every level is a pixel-centre-aligned bilinear resize of the base
// frame, taps and 8-bit weights from per-level tables the host builds once per geometry.
//   grid (tiles of 128 x 32 output pixels over all levels, nframes), block (32, 8): one thread = four adjacent
//   output pixels (one 32-bit store) in each of four rows, with the same horizontal taps; level 0 (identity) is a
//   straight copy.
// --------------------------------------------------------------------------------------------------
constexpr int PYR_TW = 128, PYR_BLOCK_ROWS = 8, PYR_ROWS_PER_THREAD = 4, PYR_TH = PYR_BLOCK_ROWS * PYR_ROWS_PER_THREAD,
              PYR_MAX_LEVELS = 16;

struct PyramidParams {
  const uint8_t* base;  // frame 0
  unsigned long long base_stride;
  int base_pitch, W, H;
  uint8_t* stacks;      // frame 0
  unsigned long long stack_stride;
  int vstep, nlevels;
  int w[PYR_MAX_LEVELS], h[PYR_MAX_LEVELS], row_off[PYR_MAX_LEVELS];
  int tiles_x[PYR_MAX_LEVELS], tile0[PYR_MAX_LEVELS + 1];  // tiles per level row; first tile index of each level
  int xt0[PYR_MAX_LEVELS], yt0[PYR_MAX_LEVELS];             // first tap-table entry of each level
  const uint32_t* xtaps;  // i0 | (i1 - i0) << 15 | f << 16 ... see pack_tap
  const uint32_t* ytaps;
  int zero_padding;       // also write zeros right of each level (columns w .. vstep)
};

// tap entry: bits 0..14 first source index, bit 15 = second tap is the next index (0 at the last row/column),
// bits 16..23 weight of the second tap (0..255)
__host__ __device__ __forceinline__ uint32_t pack_tap(int i0, int i1, int f) {
  return (uint32_t)i0 | ((uint32_t)(i1 - i0) << 15) | ((uint32_t)f << 16);
}

__global__ void __launch_bounds__(256) pyramid_build_kernel(const __grid_constant__ PyramidParams P) {
  // tile -> (level, tile row, tile column): a handful of uniform instructions per thread, amortised over the
  // PYR_ROWS_PER_THREAD rows every thread produces (per 4 pixels they were 40 % of the kernel's instructions)
  int l = 0;
  const int t = blockIdx.x;
#pragma unroll 1
  while (l + 1 < P.nlevels && t >= P.tile0[l + 1]) ++l;
  const int tl = t - P.tile0[l];
  const int ty = tl / P.tiles_x[l], tx = tl - ty * P.tiles_x[l];
  const int w = P.w[l], h = P.h[l];
  const int x = tx * PYR_TW + 4 * threadIdx.x;
  const bool identity = w == P.W && h == P.H;
  const bool pad = P.zero_padding && tx == P.tiles_x[l] - 1;  // the last tile of a level row also clears the padding
  const uint8_t* src = P.base + (size_t)blockIdx.y * P.base_stride;
  uint8_t* lvl = P.stacks + (size_t)blockIdx.y * P.stack_stride + (size_t)P.row_off[l] * P.vstep;
  // this thread's four horizontal taps (the table is padded to a multiple of 4 entries per level)
  int x0[4] = {0, 0, 0, 0}, dx[4] = {0, 0, 0, 0}, fx[4] = {0, 0, 0, 0};
  if (!identity && x < w) {
    const uint4 xe4 = __ldg(reinterpret_cast<const uint4*>(P.xtaps + P.xt0[l] + x));
    const uint32_t xe[4] = {xe4.x, xe4.y, xe4.z, xe4.w};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      x0[b] = (int)(xe[b] & 0x7fffu);
      dx[b] = (int)((xe[b] >> 15) & 1u);
      fx[b] = (int)((xe[b] >> 16) & 0xffu);
    }
  }
#pragma unroll 1
  for (int k = 0; k < PYR_ROWS_PER_THREAD; ++k) {
    const int y = ty * PYR_TH + threadIdx.y + k * PYR_BLOCK_ROWS;
    if (y >= h) break;
    uint8_t* orow = lvl + (size_t)y * P.vstep;
    // four pixels at column xs of the output row (bytes past the pitch are dropped)
    auto store4 = [&](int xs, uint32_t v) {
      uint8_t* o = orow + xs;
      if (xs + 4 <= P.vstep && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = v;
      else
        for (int b = 0; b < 4 && xs + b < P.vstep; ++b) o[b] = (uint8_t)(v >> (8 * b));
    };
    if (pad)
      for (int xz = x + PYR_TW; xz < P.vstep; xz += PYR_TW) store4(xz, 0u);
    if (x >= w) {
      if (pad && x < P.vstep) store4(x, 0u);
      continue;
    }
    uint32_t v = 0;
    if (identity) {
      const uint8_t* p = src + (size_t)y * P.base_pitch + x;
      if (x + 4 <= w && (reinterpret_cast<uintptr_t>(p) & 3) == 0) v = __ldg(reinterpret_cast<const uint32_t*>(p));
      else
        for (int b = 0; b < 4 && x + b < w; ++b) v |= (uint32_t)__ldg(p + b) << (8 * b);
    } else {
      const uint32_t ye = __ldg(P.ytaps + P.yt0[l] + y);
      const int fy = (int)((ye >> 16) & 0xffu);
      const uint8_t* r0 = src + (size_t)(ye & 0x7fffu) * P.base_pitch;
      const uint8_t* r1 = r0 + ((ye >> 15) & 1u) * P.base_pitch;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int a00 = __ldg(r0 + x0[b]), a01 = __ldg(r0 + x0[b] + dx[b]), a10 = __ldg(r1 + x0[b]),
                  a11 = __ldg(r1 + x0[b] + dx[b]);
        const int top = (a00 << 8) + (a01 - a00) * fx[b];  // a00*(256-fx) + a01*fx
        const int bot = (a10 << 8) + (a11 - a10) * fx[b];
        const int o = ((top << 8) + (bot - top) * fy + (1 << 15)) >> 16;
        v |= (uint32_t)o << (8 * b);
      }
      if (x + 4 > w) v &= 0xffffffffu >> (8 * (x + 4 - w));  // padded table entries: columns >= w stay zero
    }
    if (x + 4 > w && !P.zero_padding) {  // do not touch the bytes right of the level
      for (int b = 0; x + b < w; ++b) orow[x + b] = (uint8_t)(v >> (8 * b));
      continue;
    }
    store4(x, v);
  }
}

// batched forms of the two blur passes (blockIdx.z = frame)
__global__ void gaussian_vertical_batch_kernel(const uint8_t* __restrict__ img, unsigned long long img_stride,
                                               uint8_t* __restrict__ tmp, unsigned long long tmp_stride, int vstep,
                                               int tpitch, int width, int height) {
  const int xw = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (4 * xw >= width) return;
  img += (size_t)blockIdx.z * img_stride;
  tmp += (size_t)blockIdx.z * tmp_stride;
  uint32_t r[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint8_t* p = img + (size_t)reflect101(y + k - 2, height) * vstep + 4 * xw;
    uint32_t v = 0;
    if (4 * xw + 4 <= width && (reinterpret_cast<uintptr_t>(p) & 3) == 0) v = __ldg(reinterpret_cast<const uint32_t*>(p));
    else
      for (int b = 0; b < 4; ++b)
        if (4 * xw + b < width) v |= (uint32_t)p[b] << (8 * b);
    r[k] = v;
  }
  *reinterpret_cast<uint32_t*>(tmp + (size_t)y * tpitch + 4 * xw) = gauss5_4(r[0], r[1], r[2], r[3], r[4]);
}

__global__ void gaussian_horizontal_batch_kernel(const uint8_t* __restrict__ tmp, unsigned long long tmp_stride,
                                                 uint8_t* __restrict__ out, unsigned long long out_stride, int vstep,
                                                 int tpitch, int width, int height) {
  const int xw = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int x0 = 4 * xw;
  if (x0 >= width) return;
  const uint8_t* row = tmp + (size_t)blockIdx.z * tmp_stride + (size_t)y * tpitch;
  uint32_t t[5] = {0, 0, 0, 0, 0};
  if (x0 >= 4 && x0 + 8 <= width) {
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
  uint8_t* o = out + (size_t)blockIdx.z * out_stride + (size_t)y * vstep + x0;
  if (x0 + 4 <= width && (reinterpret_cast<uintptr_t>(o) & 3) == 0) *reinterpret_cast<uint32_t*>(o) = v;
  else
    for (int b = 0; b < 4; ++b)
      if (x0 + b < width) o[b] = (uint8_t)(v >> (8 * b));
}

}  // namespace pb
