// Host build of the device-side arithmetic (pislam_b200/csrc/device_ops.cuh) for tests/test_device_arith.py.
//
// Test infrastructure only: g++ compiles the very expressions the CUDA kernels use (the SWAR pre-tests, the
// packed FAST-9 variants, the Harris code) through scalar shims of the handful of CUDA intrinsics they call, so
// that a bit trick can be checked against a plain per-pixel model on the CPU before it costs GPU time.
// Nothing in the product links this file.
#include <cmath>
#include <cstdint>
#include <cstring>

// ---- shims of the CUDA intrinsics used by device_ops.cuh (semantics from the CUDA math API reference) ----
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (s & 31u));
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)((v << (s & 31u)) >> 32);
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t src = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) r |= ((uint32_t)(src >> (8 * ((s >> (4 * k)) & 7u))) & 0xffu) << (8 * k);
  return r;
}
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) {
    const int x = (a >> (8 * k)) & 0xff, y = (b >> (8 * k)) & 0xff;
    r |= (uint32_t)(x > y ? x - y : y - x) << (8 * k);
  }
  return r;
}
static inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float __uint_as_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline int __float2int_rz(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return -2147483647 - 1;
  return (int)f;
}

#include "../pislam_b200/csrc/device_ops.cuh"

namespace {
constexpr int PW = 40;  // words per staged row of the tile kernel (PIX_W / 4)

// plain per-pixel models
inline bool ring_model(const uint8_t* c, int pitch, int t) {
  static const int ox[16] = {-1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2};
  static const int oy[16] = {-3, -3, -3, -2, -1, 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2};
  const int v = c[0];
  for (int pol = 0; pol < 2; ++pol)
    for (int s = 0; s < 16; ++s) {
      int k = 0;
      for (; k < 9; ++k) {
        const int p = c[oy[(s + k) & 15] * pitch + ox[(s + k) & 15]];
        if (!(pol ? (v - p > t) : (p - v > t))) break;
      }
      if (k == 9) return true;
    }
  return false;
}
inline bool gt(const uint8_t* c, int pitch, int dx, int dy, int t) {
  const int d = (int)c[dy * pitch + dx] - (int)c[0];
  return (d < 0 ? -d : d) > t;
}
inline bool ge(const uint8_t* c, int pitch, int dx, int dy, int t) {
  const int d = (int)c[dy * pitch + dx] - (int)c[0];
  return (d < 0 ? -d : d) >= t;
}

template <int R>
int swar_check(const uint8_t* tile, int rows, int t) {
  // every word column 1..PW-2 and every row group that fits: the device masks against the per-pixel model.
  // `exact` must be contained in the mask; the mask must be contained in the tie-tolerant model (>= t) for t < 128
  // and equal the exact model for t >= 128.
  int bad = 0;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(tile);
  for (int r0 = 0; r0 + R + 6 <= rows; ++r0)
    for (int wx = 1; wx < PW - 1; ++wx) {
      const uint32_t* c = words + r0 * PW + wx;
      uint32_t pre, dia;
      bool any;
      if (t < 128) {
        pre = pb::pretest_rows<true, PW, R>(c, t);
        dia = pb::diag_rows<true, PW, R>(c, t);
        any = pb::ns_any<true, PW, R>(c, t);
      } else {
        pre = pb::pretest_rows<false, PW, R>(c, t);
        dia = pb::diag_rows<false, PW, R>(c, t);
        any = pb::ns_any<false, PW, R>(c, t);
      }
      bool any_model = false;
      for (int i = 0; i < R; ++i)
        for (int k = 0; k < 4; ++k) {
          const uint8_t* p = tile + (r0 + i + 3) * PW * 4 + 4 * wx + k;
          const int pitch = PW * 4;
          const bool ns = gt(p, pitch, 0, -3, t) || gt(p, pitch, 0, 3, t);
          const bool pre_x = ns && (gt(p, pitch, 3, 0, t) || gt(p, pitch, -3, 0, t));
          const bool pre_l = (ge(p, pitch, 0, -3, t) || ge(p, pitch, 0, 3, t)) && (ge(p, pitch, 3, 0, t) || ge(p, pitch, -3, 0, t));
          const bool dia_x = (gt(p, pitch, 2, -2, t) || gt(p, pitch, -2, 2, t)) && (gt(p, pitch, 2, 2, t) || gt(p, pitch, -2, -2, t));
          const bool dia_l = (ge(p, pitch, 2, -2, t) || ge(p, pitch, -2, 2, t)) && (ge(p, pitch, 2, 2, t) || ge(p, pitch, -2, -2, t));
          const bool pre_d = (pre >> (8 * k + i)) & 1u, dia_d = (dia >> (8 * k + i)) & 1u;
          any_model |= ns;
          if (pre_x && !pre_d) ++bad;
          if (dia_x && !dia_d) ++bad;
          if (t < 128 ? (pre_d && !pre_l) : (pre_d != pre_x)) ++bad;
          if (t < 128 ? (dia_d && !dia_l) : (dia_d != dia_x)) ++bad;
          // a FAST-9 corner passes both device tests
          if (ring_model(p, pitch, t) && !(pre_d && dia_d)) ++bad;
        }
      if (any_model && !any) ++bad;
      if (pre != 0u && !any) ++bad;  // the cascade's first stage never drops a task the compass test would keep
      // bits outside the R rows stay clear
      const uint32_t rowmask = ((1u << R) - 1u) * 0x01010101u;
      if ((pre | dia) & ~rowmask) ++bad;
    }
  return bad;
}
}  // namespace

extern "C" {

// FAST-9 on every pixel with a full ring inside [3, w-3) x [3, h-3): model, serial-mask and PRMT variants; returns the
// number of pixels on which a device variant differs from the model
int dah_fast9(const uint8_t* img, int pitch, int w, int h, int t, int* ncorners) {
  int bad = 0, n = 0;
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* c = img + y * pitch + x;
      const bool m = ring_model(c, pitch, t);
      n += m;
      if (pb::fast9_at(c, pitch, t) != m) ++bad;
      if (pb::fast9_prmt_at(c, pitch, t) != m) ++bad;
    }
  *ncorners = n;
  return bad;
}

// SWAR pre-tests on a tile of `rows` rows of PW words, rows-per-task 5 and 6
int dah_swar(const uint8_t* tile, int rows, int t) { return swar_check<5>(tile, rows, t) + swar_check<6>(tile, rows, t); }

// Harris code at n points (the expression the tile kernel's phase C evaluates)
void dah_harris(const uint8_t* img, int pitch, const int* xs, const int* ys, int n, int32_t threshold, uint8_t* out) {
  for (int i = 0; i < n; ++i) out[i] = pb::harris_code_at(img + ys[i] * pitch + xs[i], pitch, threshold);
}

uint32_t dah_prmt(uint32_t a, uint32_t b, uint32_t sel) { return pb::prmt(a, b, sel); }

}  // extern "C"
