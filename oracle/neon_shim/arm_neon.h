// TEST INFRASTRUCTURE ONLY -- never linked into libpislam_b200.so.
//
// A scalar, x86-hosted stand-in for the subset of <arm_neon.h> that the
// reference hot path uses (Fast.h, Harris.h, Orb.h under /root/reference/include).
// It exists so that the *reference's own headers* can be compiled here and act
// as the oracle-of-the-oracle (see oracle/Makefile, oracle/ref_driver.cpp).
//
// Every NEON type is a small value class (`nv<T,N>`) that supports what the
// reference does with NEON types besides calling intrinsics: brace
// initialisation, operator[] on lanes, and the arithmetic / bitwise / shift
// operators GCC accepts on ARM vector types.  Lane semantics follow the ARM
// Architecture Reference Manual (wrap-around integer lanes, floor halving
// add/sub, VSHL by a signed per-lane count, truncating float->int conversion
// with NaN -> 0, and the 8-bit FPRecipEstimate table function).
#ifndef PISLAM_ORACLE_NEON_SHIM_H_
#define PISLAM_ORACLE_NEON_SHIM_H_

#include <cmath>
#include <cstdint>
#include <cstring>

template <typename T, int N>
struct nv {
  T v[N];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};

#define NV_BINOP(op)                                                         \
  template <typename T, int N>                                               \
  static inline nv<T, N> operator op(const nv<T, N>& a, const nv<T, N>& b) { \
    nv<T, N> r;                                                              \
    for (int i = 0; i < N; ++i) r.v[i] = (T)(a.v[i] op b.v[i]);              \
    return r;                                                                \
  }                                                                          \
  template <typename T, int N, typename S>                                   \
  static inline nv<T, N> operator op(const nv<T, N>& a, S b) {               \
    nv<T, N> r;                                                              \
    for (int i = 0; i < N; ++i) r.v[i] = (T)(a.v[i] op (T)b);                \
    return r;                                                                \
  }
NV_BINOP(+)
NV_BINOP(-)
NV_BINOP(*)
NV_BINOP(&)
NV_BINOP(|)
NV_BINOP(^)
#undef NV_BINOP

template <typename T, int N, typename S>
static inline nv<T, N>& operator+=(nv<T, N>& a, S b) { a = a + b; return a; }
template <typename T, int N, typename S>
static inline nv<T, N>& operator-=(nv<T, N>& a, S b) { a = a - b; return a; }

// `vec << vec` on ARM is VSHL: per-lane count taken as a SIGNED byte, negative
// counts shift right, |count| >= lane width gives 0.
template <typename T, int N>
static inline nv<T, N> operator<<(const nv<T, N>& a, const nv<T, N>& b) {
  nv<T, N> r;
  const int bits = 8 * (int)sizeof(T);
  for (int i = 0; i < N; ++i) {
    int s = (int8_t)(b.v[i] & 0xff);
    uint64_t x = (uint64_t)a.v[i] & ((bits == 64) ? ~0ull : ((1ull << bits) - 1));
    if (s >= bits || s <= -bits) r.v[i] = 0;
    else if (s >= 0) r.v[i] = (T)(x << s);
    else r.v[i] = (T)(x >> (-s));
  }
  return r;
}

typedef nv<uint8_t, 16> uint8x16_t;
typedef nv<uint8_t, 8> uint8x8_t;
typedef nv<int8_t, 8> int8x8_t;
typedef nv<uint16_t, 8> uint16x8_t;
typedef nv<uint16_t, 4> uint16x4_t;
typedef nv<int16_t, 8> int16x8_t;
typedef nv<uint32_t, 4> uint32x4_t;
typedef nv<int32_t, 4> int32x4_t;
typedef nv<uint32_t, 2> uint32x2_t;
typedef nv<int32_t, 2> int32x2_t;
typedef nv<uint64_t, 1> uint64x1_t;
typedef nv<int64_t, 1> int64x1_t;
typedef nv<float, 4> float32x4_t;
typedef nv<float, 2> float32x2_t;

template <typename D, typename S>
static inline D nv_cast(const S& s) {
  static_assert(sizeof(D) == sizeof(S), "reinterpret size");
  D d;
  std::memcpy(&d, &s, sizeof(D));
  return d;
}
#define NV_REINTERPRET(name, D, S) \
  static inline D name(const S& a) { return nv_cast<D>(a); }
NV_REINTERPRET(vreinterpret_s8_u8, int8x8_t, uint8x8_t)
NV_REINTERPRET(vreinterpret_u8_s8, uint8x8_t, int8x8_t)
NV_REINTERPRET(vreinterpret_s8_u64, int8x8_t, uint64x1_t)
NV_REINTERPRET(vreinterpret_u64_s8, uint64x1_t, int8x8_t)
NV_REINTERPRET(vreinterpret_u8_u64, uint8x8_t, uint64x1_t)
NV_REINTERPRET(vreinterpret_u64_u8, uint64x1_t, uint8x8_t)
NV_REINTERPRET(vreinterpretq_u16_s16, uint16x8_t, int16x8_t)
NV_REINTERPRET(vreinterpretq_s16_u16, int16x8_t, uint16x8_t)
NV_REINTERPRET(vreinterpret_u32_u64, uint32x2_t, uint64x1_t)
NV_REINTERPRET(vreinterpret_s32_s64, int32x2_t, int64x1_t)
NV_REINTERPRET(vreinterpret_u32_s32, uint32x2_t, int32x2_t)
NV_REINTERPRET(vreinterpret_s32_u32, int32x2_t, uint32x2_t)
NV_REINTERPRET(vreinterpret_u32_f32, uint32x2_t, float32x2_t)
NV_REINTERPRET(vreinterpret_u32_u8, uint32x2_t, uint8x8_t)
NV_REINTERPRET(vreinterpret_u16_u64, uint16x4_t, uint64x1_t)
NV_REINTERPRET(vreinterpret_u64_u16, uint64x1_t, uint16x4_t)
#undef NV_REINTERPRET

// ---- loads / stores / dup / lane ------------------------------------------
static inline uint8x16_t vld1q_u8(const uint8_t* p) { uint8x16_t r; std::memcpy(r.v, p, 16); return r; }
static inline uint8x8_t vld1_u8(const uint8_t* p) { uint8x8_t r; std::memcpy(r.v, p, 8); return r; }
static inline int32x4_t vld1q_s32(const int32_t* p) { int32x4_t r; std::memcpy(r.v, p, 16); return r; }
static inline void vst1q_u8(uint8_t* p, const uint8x16_t& a) { std::memcpy(p, a.v, 16); }
static inline uint8x16_t vdupq_n_u8(int x) { uint8x16_t r; for (int i = 0; i < 16; ++i) r.v[i] = (uint8_t)x; return r; }
static inline uint8x8_t vdup_n_u8(int x) { uint8x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint8_t)x; return r; }
static inline float32x4_t vdupq_n_f32(float x) { float32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = x; return r; }
static inline uint8x8_t vset_lane_u8(uint8_t x, uint8x8_t a, int lane) { a.v[lane] = x; return a; }

// ---- 8-bit x16 (Fast.h) -----------------------------------------------------
static inline uint8x16_t vqaddq_u8(const uint8x16_t& a, const uint8x16_t& b) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) { int s = a.v[i] + b.v[i]; r.v[i] = (uint8_t)(s > 255 ? 255 : s); } return r; }
static inline uint8x16_t vqsubq_u8(const uint8x16_t& a, const uint8x16_t& b) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) { int s = a.v[i] - b.v[i]; r.v[i] = (uint8_t)(s < 0 ? 0 : s); } return r; }
static inline uint8x16_t vcgeq_u8(const uint8x16_t& a, const uint8x16_t& b) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) r.v[i] = a.v[i] >= b.v[i] ? 0xff : 0; return r; }
static inline uint8x16_t vcleq_u8(const uint8x16_t& a, const uint8x16_t& b) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) r.v[i] = a.v[i] <= b.v[i] ? 0xff : 0; return r; }
static inline uint8x16_t vbslq_u8(const uint8x16_t& m, const uint8x16_t& a, const uint8x16_t& b) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) r.v[i] = (uint8_t)((m.v[i] & a.v[i]) | (~m.v[i] & b.v[i])); return r; }
static inline uint8x16_t vtstq_u8(const uint8x16_t& a, const uint8x16_t& b) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) r.v[i] = (a.v[i] & b.v[i]) ? 0xff : 0; return r; }
static inline uint8x16_t vclzq_u8(const uint8x16_t& a) {
  uint8x16_t r;
  for (int i = 0; i < 16; ++i) { int n = 0; for (int b = 7; b >= 0 && !((a.v[i] >> b) & 1); --b) ++n; r.v[i] = (uint8_t)n; }
  return r; }
// stands in for the `vceq.u8 q, q, #0` inline asm of Fast.h:140,144 (see oracle/Makefile)
static inline uint8x16_t nv_vceqzq_u8(const uint8x16_t& a) {
  uint8x16_t r; for (int i = 0; i < 16; ++i) r.v[i] = a.v[i] == 0 ? 0xff : 0; return r; }

// ---- 8-bit x8 (Harris.h, Orb.h) ----------------------------------------------
static inline uint8x8_t vhsub_u8(const uint8x8_t& a, const uint8x8_t& b) {
  uint8x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint8_t)(((int)a.v[i] - (int)b.v[i]) >> 1); return r; }
static inline int8x8_t vhadd_s8(const int8x8_t& a, const int8x8_t& b) {
  int8x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (int8_t)(((int)a.v[i] + (int)b.v[i]) >> 1); return r; }
static inline uint8x8_t vand_u8(const uint8x8_t& a, const uint8x8_t& b) { return a & b; }
static inline uint8x8_t vcle_u8(const uint8x8_t& a, const uint8x8_t& b) {
  uint8x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = a.v[i] <= b.v[i] ? 0xff : 0; return r; }
static inline int16x8_t vmull_s8(const int8x8_t& a, const int8x8_t& b) {
  int16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (int16_t)((int)a.v[i] * (int)b.v[i]); return r; }
static inline int16x8_t vmlal_s8(const int16x8_t& c, const int8x8_t& a, const int8x8_t& b) {
  int16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (int16_t)((int)c.v[i] + (int)a.v[i] * (int)b.v[i]); return r; }
static inline uint16x8_t vmull_u8(const uint8x8_t& a, const uint8x8_t& b) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint16_t)((unsigned)a.v[i] * (unsigned)b.v[i]); return r; }
static inline uint16x8_t vmlal_u8(const uint16x8_t& c, const uint8x8_t& a, const uint8x8_t& b) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint16_t)((unsigned)c.v[i] + (unsigned)a.v[i] * (unsigned)b.v[i]); return r; }
static inline uint16x8_t vmlsl_u8(const uint16x8_t& c, const uint8x8_t& a, const uint8x8_t& b) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint16_t)((unsigned)c.v[i] - (unsigned)a.v[i] * (unsigned)b.v[i]); return r; }
static inline uint16x8_t vsubl_u8(const uint8x8_t& a, const uint8x8_t& b) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint16_t)((unsigned)a.v[i] - (unsigned)b.v[i]); return r; }

// ---- pairwise widening adds ---------------------------------------------------
static inline uint32x4_t vpaddlq_u16(const uint16x8_t& a) {
  uint32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (uint32_t)a.v[2 * i] + (uint32_t)a.v[2 * i + 1]; return r; }
static inline int32x4_t vpaddlq_s16(const int16x8_t& a) {
  int32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (int32_t)a.v[2 * i] + (int32_t)a.v[2 * i + 1]; return r; }
static inline uint32x4_t vpadalq_u16(const uint32x4_t& c, const uint16x8_t& a) {
  uint32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = c.v[i] + (uint32_t)a.v[2 * i] + (uint32_t)a.v[2 * i + 1]; return r; }
static inline int32x4_t vpadalq_s16(const int32x4_t& c, const int16x8_t& a) {
  int32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (int32_t)((uint32_t)c.v[i] + (uint32_t)((int32_t)a.v[2 * i] + (int32_t)a.v[2 * i + 1])); return r; }
static inline uint64x1_t vpaddl_u32(const uint32x2_t& a) { uint64x1_t r; r.v[0] = (uint64_t)a.v[0] + (uint64_t)a.v[1]; return r; }
static inline int64x1_t vpaddl_s32(const int32x2_t& a) { int64x1_t r; r.v[0] = (int64_t)a.v[0] + (int64_t)a.v[1]; return r; }

// ---- 32-bit x2 / x4 -------------------------------------------------------------
static inline uint32x2_t vget_low_u32(const uint32x4_t& a) { uint32x2_t r = {{a.v[0], a.v[1]}}; return r; }
static inline uint32x2_t vget_high_u32(const uint32x4_t& a) { uint32x2_t r = {{a.v[2], a.v[3]}}; return r; }
static inline int32x2_t vget_low_s32(const int32x4_t& a) { int32x2_t r = {{a.v[0], a.v[1]}}; return r; }
static inline int32x2_t vget_high_s32(const int32x4_t& a) { int32x2_t r = {{a.v[2], a.v[3]}}; return r; }
static inline uint32x2_t vadd_u32(const uint32x2_t& a, const uint32x2_t& b) { return a + b; }
static inline int32x2_t vadd_s32(const int32x2_t& a, const int32x2_t& b) {
  int32x2_t r; for (int i = 0; i < 2; ++i) r.v[i] = (int32_t)((uint32_t)a.v[i] + (uint32_t)b.v[i]); return r; }
static inline int32x2_t vsub_s32(const int32x2_t& a, const int32x2_t& b) {
  int32x2_t r; for (int i = 0; i < 2; ++i) r.v[i] = (int32_t)((uint32_t)a.v[i] - (uint32_t)b.v[i]); return r; }
static inline uint32x2_t vmul_u32(const uint32x2_t& a, const uint32x2_t& b) { return a * b; }
static inline int32x2_t vmls_s32(const int32x2_t& c, const int32x2_t& a, const int32x2_t& b) {
  int32x2_t r; for (int i = 0; i < 2; ++i) r.v[i] = (int32_t)((uint32_t)c.v[i] - (uint32_t)a.v[i] * (uint32_t)b.v[i]); return r; }
static inline uint32x4_t vsubq_u32(const uint32x4_t& a, const uint32x4_t& b) { return a - b; }
#define vshr_n_u64(a, n) nv_shr_u64((a), (n))
#define vshr_n_u32(a, n) nv_shr_u32((a), (n))
#define vshr_n_s32(a, n) nv_shr_s32((a), (n))
static inline uint64x1_t nv_shr_u64(const uint64x1_t& a, int n) { uint64x1_t r; r.v[0] = a.v[0] >> n; return r; }
static inline uint32x2_t nv_shr_u32(const uint32x2_t& a, int n) { uint32x2_t r; for (int i = 0; i < 2; ++i) r.v[i] = a.v[i] >> n; return r; }
static inline int32x2_t nv_shr_s32(const int32x2_t& a, int n) { int32x2_t r; for (int i = 0; i < 2; ++i) r.v[i] = a.v[i] >> n; return r; }

// ---- float ------------------------------------------------------------------------
static inline float32x2_t vcvt_f32_s32(const int32x2_t& a) { float32x2_t r; for (int i = 0; i < 2; ++i) r.v[i] = (float)a.v[i]; return r; }
static inline float32x4_t vcvtq_f32_s32(const int32x4_t& a) { float32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (float)a.v[i]; return r; }
static inline float32x4_t vabsq_f32(const float32x4_t& a) { float32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = std::fabs(a.v[i]); return r; }
static inline float32x4_t vmaxq_f32(const float32x4_t& a, const float32x4_t& b) {
  float32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (a.v[i] != a.v[i] || b.v[i] != b.v[i]) ? NAN : (a.v[i] > b.v[i] ? a.v[i] : b.v[i]); return r; }
static inline float32x4_t vminq_f32(const float32x4_t& a, const float32x4_t& b) {
  float32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = (a.v[i] != a.v[i] || b.v[i] != b.v[i]) ? NAN : (a.v[i] < b.v[i] ? a.v[i] : b.v[i]); return r; }
static inline float32x4_t vmulq_f32(const float32x4_t& a, const float32x4_t& b) { return a * b; }
// VCVT.S32.F32: round toward zero, saturate, NaN -> 0.
static inline int32x4_t vcvtq_s32_f32(const float32x4_t& a) {
  int32x4_t r;
  for (int i = 0; i < 4; ++i) {
    float f = a.v[i];
    if (f != f) r.v[i] = 0;
    else if (f >= 2147483648.0f) r.v[i] = INT32_MAX;
    else if (f <= -2147483648.0f) r.v[i] = INT32_MIN;
    else r.v[i] = (int32_t)f;
  }
  return r;
}
// VRECPE.F32 = FPRecipEstimate (ARM ARM, AArch32 Advanced SIMD: denormals flush
// to zero, 8-bit estimate from the 9-bit scaled mantissa).
static inline float nv_recpe(float x) {
  uint32_t u; std::memcpy(&u, &x, 4);
  uint32_t sign = u & 0x80000000u;
  uint32_t exp = (u >> 23) & 0xff;
  uint32_t frac = u & 0x7fffffu;
  uint32_t out;
  if (exp == 0xff) out = frac ? 0x7fc00000u : sign;              // NaN -> default NaN, Inf -> 0
  else if (exp == 0) out = sign | 0x7f800000u;                   // zero / flushed denormal -> Inf
  else if (exp >= 253) out = sign;                               // |x| >= 2^126 underflows to 0
  else {
    uint32_t q = 256 + (frac >> 15);                             // floor(scaled * 512), scaled in [0.5, 1)
    uint32_t s = (uint32_t)((256.0 * 1024.0) / (2.0 * q + 1.0) + 0.5);  // floor(256 / ((q + .5) / 512) + .5)
    out = sign | ((253 - exp) << 23) | ((s - 256) << 15);
  }
  float r; std::memcpy(&r, &out, 4);
  return r;
}
static inline float32x4_t vrecpeq_f32(const float32x4_t& a) { float32x4_t r; for (int i = 0; i < 4; ++i) r.v[i] = nv_recpe(a.v[i]); return r; }

// ---- Bilinear.h (pyramid helpers: 7/8 and 13/16 down-scaling) ------------------------------------
static inline void vst1_u8(uint8_t* p, const uint8x8_t& a) { std::memcpy(p, a.v, 8); }
#define vst1_lane_u8(p, a, lane) ((p)[0] = (a).v[(lane)])
#define vst1_lane_u32(p, a, lane) nv_st1_lane_u32((p), (a), (lane))
static inline void nv_st1_lane_u32(uint32_t* p, const uint32x2_t& a, int lane) { std::memcpy(p, &a.v[lane], 4); }
static inline uint8x8_t vget_low_u8(const uint8x16_t& a) { uint8x8_t r; std::memcpy(r.v, a.v, 8); return r; }
static inline uint8x8_t vget_high_u8(const uint8x16_t& a) { uint8x8_t r; std::memcpy(r.v, a.v + 8, 8); return r; }
static inline uint16x4_t vget_low_u16(const uint16x8_t& a) { uint16x4_t r; std::memcpy(r.v, a.v, 8); return r; }
static inline uint16x4_t vget_high_u16(const uint16x8_t& a) { uint16x4_t r; std::memcpy(r.v, a.v + 4, 8); return r; }
static inline uint16x8_t vcombine_u16(const uint16x4_t& lo, const uint16x4_t& hi) {
  uint16x8_t r; std::memcpy(r.v, lo.v, 8); std::memcpy(r.v + 4, hi.v, 8); return r; }
static inline uint16x8_t vaddq_u16(const uint16x8_t& a, const uint16x8_t& b) { return a + b; }
// VEXT: lanes n.. of a followed by lanes 0..n-1 of b
#define vextq_u16(a, b, n) nv_extq_u16((a), (b), (n))
static inline uint16x8_t nv_extq_u16(const uint16x8_t& a, const uint16x8_t& b, int n) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (i + n < 8) ? a.v[i + n] : b.v[i + n - 8]; return r; }
// VRSHR: rounding shift right, the rounding add does not overflow the lane (computed wider)
#define vrshrq_n_u16(a, n) nv_rshrq_u16((a), (n))
static inline uint16x8_t nv_rshrq_u16(const uint16x8_t& a, int n) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint16_t)(((uint32_t)a.v[i] + (1u << (n - 1))) >> n); return r; }
#define vmulq_lane_u16(a, v, lane) nv_mulq_lane_u16((a), (v), (lane))
static inline uint16x8_t nv_mulq_lane_u16(const uint16x8_t& a, const uint16x4_t& v, int lane) {
  uint16x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint16_t)((uint32_t)a.v[i] * (uint32_t)v.v[lane]); return r; }
// VRADDHN: (a + b + 0x80) >> 8 on wrap-around 16-bit sums, narrowed to 8 bits
static inline uint8x8_t vraddhn_u16(const uint16x8_t& a, const uint16x8_t& b) {
  uint8x8_t r; for (int i = 0; i < 8; ++i) r.v[i] = (uint8_t)(((uint16_t)(a.v[i] + b.v[i] + 0x80u)) >> 8); return r; }
// VSLI: shift b left by n and insert into a, keeping the low n bits of a
#define vsli_n_u64(a, b, n) nv_sli_u64((a), (b), (n))
static inline uint64x1_t nv_sli_u64(const uint64x1_t& a, const uint64x1_t& b, int n) {
  uint64x1_t r; r.v[0] = (b.v[0] << n) | (a.v[0] & ((1ull << n) - 1ull)); return r; }

#endif  // PISLAM_ORACLE_NEON_SHIM_H_
