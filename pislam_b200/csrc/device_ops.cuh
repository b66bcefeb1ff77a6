// Device-side arithmetic of the ORB front-end (sm_100a).  Every function states the rule the
// reference implements and cites the reference lines it must match
// bit-for-bit; none of them mirrors the reference's NEON instruction sequences.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pb {

// ---------------------------------------------------------------------------------------------
// FAST-9 (reference Fast.h:54-158).  Ring clockwise from (-1,-3) (Fast.h:66-128).
// ---------------------------------------------------------------------------------------------
// >= 9 consecutive set bits in a 16-bit circular word
__device__ __forceinline__ bool arc9(uint32_t m) {
  uint32_t r = m | (m << 16);
  r &= r >> 1;
  r &= r >> 2;
  r &= r >> 4;
  r &= r >> 1;
  return (r & 0xffffu) != 0;
}

// `c` points at the centre pixel inside a buffer with row pitch `pitch` (shared or global).
// Saturating thresholds (Fast.h:63-64) reduce to plain integer compares: p > min(c+t,255) <=>
// p - c > t, p < max(c-t,0) <=> c - p > t.
template <typename Ptr>
__device__ __forceinline__ bool fast9_at(Ptr c, int pitch, int t) {
  const int v = c[0];
  const int kb = v + t, kd = t - v;  // bright: kb - p < 0 ; dark: p + kd < 0
  uint32_t mb = 0, md = 0;
#define PB_RING(dx, dy)                                              \
  {                                                                  \
    const int p = c[(dy) * pitch + (dx)];                            \
    mb = __funnelshift_l((uint32_t)(kb - p), mb, 1);                 \
    md = __funnelshift_l((uint32_t)(p + kd), md, 1);                 \
  }
  PB_RING(-1, -3) PB_RING(0, -3) PB_RING(1, -3) PB_RING(2, -2)
  PB_RING(3, -1) PB_RING(3, 0) PB_RING(3, 1) PB_RING(2, 2)
  PB_RING(1, 3) PB_RING(0, 3) PB_RING(-1, 3) PB_RING(-2, 2)
  PB_RING(-3, 1) PB_RING(-3, 0) PB_RING(-3, -1) PB_RING(-2, -2)
#undef PB_RING
  return arc9(mb & 0xffffu) || arc9(md & 0xffffu);
}

// The same test with both polarities packed into one word per ring pixel (the tile kernel's phase B):
// u = 65536*(c+t-p) + (p+t-c) = p*0xFFFF0001 + K has the "brighter" sign in bit 31 and the "darker" sign in bit 15
// (the borrow of a negative low half can only turn c+t-p = 0 into -1 when p < c-t, impossible with t >= 0), so one
// IMAD replaces two subtractions and one shift + one logic op collect both ring masks: bright in the high half,
// dark in the low half, ring order reversed (irrelevant for a circular run).  The 9-run test then folds both
// halves at once with rotations inside the 16-bit halves.
__device__ __forceinline__ uint32_t rotl16x2(uint32_t x, int k) {  // rotate both halves left by k (1..15)
  const uint32_t hm = (0xffffu << k) & 0xffffu;
  const uint32_t hm2 = hm | (hm << 16);
  return ((x << k) & hm2) | ((x >> (16 - k)) & ~hm2);
}
template <typename Ptr>
__device__ __forceinline__ bool fast9_packed_at(Ptr c, int pitch, int t) {
  const uint32_t v = c[0];
  const uint32_t K = ((v + (uint32_t)t) << 16) + ((uint32_t)t - v);
  uint32_t m = 0;
#define PB_RING(dx, dy)                                               \
  {                                                                   \
    const uint32_t u = (uint32_t)c[(dy) * pitch + (dx)] * 0xffff0001u + K; \
    m = (m >> 1) | (u & 0x80008000u);                                 \
  }
  PB_RING(-1, -3) PB_RING(0, -3) PB_RING(1, -3) PB_RING(2, -2)
  PB_RING(3, -1) PB_RING(3, 0) PB_RING(3, 1) PB_RING(2, 2)
  PB_RING(1, 3) PB_RING(0, 3) PB_RING(-1, 3) PB_RING(-2, 2)
  PB_RING(-3, 1) PB_RING(-3, 0) PB_RING(-3, -1) PB_RING(-2, -2)
#undef PB_RING
  uint32_t r = m & rotl16x2(m, 1);   // runs of 2
  r &= rotl16x2(r, 2);               // runs of 4
  r &= rotl16x2(r, 4);               // runs of 8
  r &= __byte_perm(m, 0u, 0x2301);   // ... preceded by a ninth (rotation by 8 = byte swap inside each half)
  return r != 0u;
}

// ---------------------------------------------------------------------------------------------
// Harris code (reference Harris.h:37-69, 80-248).  P(r,j) = pixel (x-3+j, y-3+r), r,j = 0..7.
// Central differences with floor halving, [1 2 1]/4 smoothing by two floor-halving adds, 6x6
// structure tensor over [x-2,x+3]x[y-2,y+3], k = 1/16, 8-bit log code.  The reference's 16-bit
// intermediate lanes never wrap (|dx*dy + dx'*dy'| <= 32512), so plain 32-bit integer
// accumulation is exact; the final products wrap mod 2^32 like the reference's 32-bit lanes.
// ---------------------------------------------------------------------------------------------
// Hides a pixel's 0..255 range from the optimiser: with it, NVVM narrows the gradient arithmetic to 16 bits
// and pays a sign-extending PRMT per operation (~25% more instructions in this function).
__device__ __forceinline__ int opaque(int v) {
  asm("" : "+r"(v));
  return v;
}

template <typename Ptr>
__device__ __forceinline__ uint8_t harris_code_at(Ptr c, int pitch, int32_t threshold) {
  // e(r,j) = (P(r,j+2) - P(r,j)) >> 1 for 8 rows; dx(n,j) = (((e(n,j)+e(n+2,j))>>1) + e(n+1,j)) >> 1
  // d(n,j) = (P(n+2,j) - P(n,j)) >> 1 ; dy(n,j) = (d(n,j+1) + ((d(n,j)+d(n,j+2))>>1)) >> 1
  // Walk rows keeping a 3-row window of e and of raw pixels.
  int P0[8], P1[8], P2[8];
  int e0[6], e1[6], e2[6];
  uint32_t sxx = 0, syy = 0;
  int32_t sxy = 0;
  Ptr base = c - 3 * pitch - 3;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    P0[j] = opaque(base[j]);
    P1[j] = opaque(base[pitch + j]);
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    e0[j] = (P0[j + 2] - P0[j]) >> 1;
    e1[j] = (P1[j + 2] - P1[j]) >> 1;
  }
#pragma unroll
  for (int n = 0; n < 6; ++n) {
#pragma unroll
    for (int j = 0; j < 8; ++j) P2[j] = opaque(base[(n + 2) * pitch + j]);
    int d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = (P2[j] - P0[j]) >> 1;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      e2[j] = (P2[j + 2] - P2[j]) >> 1;
      const int dx = (((e0[j] + e2[j]) >> 1) + e1[j]) >> 1;
      const int dy = (d[j + 1] + ((d[j] + d[j + 2]) >> 1)) >> 1;
      sxx += (uint32_t)(dx * dx);
      syy += (uint32_t)(dy * dy);
      sxy += dx * dy;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      P0[j] = P1[j];
      P1[j] = P2[j];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      e0[j] = e1[j];
      e1[j] = e2[j];
    }
  }
  const uint32_t Ixx = sxx >> 4, Iyy = syy >> 4;  // Harris.h:243-244
  const int32_t Ixy = sxy >> 4;                   // Harris.h:245 (arithmetic)
  uint32_t tr = Ixx + Iyy;
  tr = (tr * tr) >> 4;  // Harris.h:41-44
  const uint32_t det = Ixx * Iyy - (uint32_t)Ixy * (uint32_t)Ixy;
  const int32_t score = (int32_t)(det - tr);
  if (threshold < score) return (uint8_t)((__float_as_uint(__int2float_rn(score)) >> 20) & 0xffu);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Keypoint packing (reference Util.h:27-45)
// ---------------------------------------------------------------------------------------------
__device__ __host__ __forceinline__ uint32_t encode_fast(uint32_t score, uint32_t x, uint32_t y) {
  return (score << 24) | (x << 12) | y;
}
__device__ __host__ __forceinline__ uint32_t decode_x(uint32_t e) { return (e >> 12) & 0xfffu; }
__device__ __host__ __forceinline__ uint32_t decode_y(uint32_t e) { return e & 0xfffu; }

// ---------------------------------------------------------------------------------------------
// Orientation (reference Orb.h:80-308 moments, Orb.h:310-387 angle bin)
// ---------------------------------------------------------------------------------------------
// half-width of the moment disc per |dy| (Orb.h:118-126 masks/weights + row macros :163-286)
__host__ __device__ __forceinline__ int disc_halfwidth(int ady) {
  // {15,15,15,15,15,15,14,14,13,13,12,11,10,9,7,5} packed 4 bits each, |dy| = 0 in the low nibble
  return (int)((0x579abcddeeffffffull >> (4 * ady)) & 15u);
}

// dot products of four unsigned pixel bytes with signed / unsigned byte weights (IDP.4A)
__device__ __forceinline__ int32_t dp4a_us(uint32_t pix, uint32_t w_s8, int32_t acc) {
  int32_t r;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(pix), "r"(w_s8), "r"(acc));
  return r;
}
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t pix, uint32_t w_u8, uint32_t acc) {
  uint32_t r;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(pix), "r"(w_u8), "r"(acc));
  return r;
}

// ARM FPRecipEstimate (vrecpeq_f32, Orb.h:329): 8 fractional bits, flush-to-zero.  The estimate depends on the
// top 8 fraction bits only: q = 256 + (frac >> 15), s = floor(2^19 / (2q+1) + 1/2) in [256, 511].
__device__ __forceinline__ uint32_t recpe_mantissa(uint32_t top8) {
  const uint32_t q2 = 2u * (256u + top8) + 1u;  // 2q+1
  return (524288u + q2) / (2u * q2);            // floor(2^18/(2q+1) + 1/2)
}

// `lut`, when given, holds recpe_mantissa(0..255) (built once per CTA: the division is ~20 instructions)
__device__ __forceinline__ float recpe(float x, const uint16_t* lut = nullptr) {
  const uint32_t u = __float_as_uint(x);
  const uint32_t sign = u & 0x80000000u, e = (u >> 23) & 0xffu, frac = u & 0x7fffffu;
  uint32_t o;
  if (e == 0xff) o = frac ? 0x7fc00000u : sign;
  else if (e == 0) o = sign | 0x7f800000u;
  else if (e >= 253) o = sign;
  else {
    const uint32_t s = lut ? (uint32_t)lut[frac >> 15] : recpe_mantissa(frac >> 15);
    o = sign | ((253u - e) << 23) | ((s - 256u) << 15);
  }
  return __uint_as_float(o);
}

// Every fp32 operation individually rounded (no FMA contraction): Orb.h:338-350.
__device__ __forceinline__ uint32_t angle_bin(int32_t m10, int32_t m01, const uint16_t* recpe_lut = nullptr) {
  const float ax = fabsf(__int2float_rn(m10)), ay = fabsf(__int2float_rn(m01));
  const float zmax = fmaxf(ax, ay), zmin = fminf(ax, ay);
  const float z = __fmul_rn(zmin, recpe(zmax, recpe_lut));  // flat patch: 0 * inf = NaN
  const float c0 = (float)(256 * 14.999998), c1 = (float)(256 * 4.723436), c2 = (float)(256 * 1.266240);
  const float t1 = __fadd_rn(c1, __fmul_rn(c2, z));
  const float t3 = __fmul_rn(__fsub_rn(z, 1.0f), t1);
  const float t = __fmul_rn(z, __fsub_rn(c0, t3));
  int32_t a = (t != t) ? 0 : __float2int_rz(t);  // vcvtq_s32_f32: NaN -> 0, truncate, saturate
  const uint32_t absx = m10 < 0 ? 0u - (uint32_t)m10 : (uint32_t)m10;
  const uint32_t absy = m01 < 0 ? 0u - (uint32_t)m01 : (uint32_t)m01;
  if (absx > absy) {  // Orb.h:357-375
    if ((m10 ^ m01) < 0) a = -a;
    if (m10 < 0) a += 256 * 60;
    else if (a < 0) a += 256 * 120;
  } else {
    if ((m10 ^ m01) >= 0) a = -a;
    a += (m01 >= 0) ? 256 * 30 : 256 * 90;
  }
  a >>= 10;
  if (!(0 <= a && a < 30)) a = 0;
  return (uint32_t)a;
}

}  // namespace pb
