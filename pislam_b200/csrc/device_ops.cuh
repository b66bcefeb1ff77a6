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

// The tile kernel's phase B tests both polarities at once: u = 65536*(c+t-p) + (p+t-c) = p*0xFFFF0001 + K has the
// "brighter" sign in bit 31 and the "darker" sign in bit 15 (the borrow of a negative low half can only turn
// c+t-p = 0 into -1 when p < c-t, impossible with t >= 0), so one IMAD replaces two subtractions; the 9-run test
// then folds both 16-bit masks at once with rotations inside the halves.
__device__ __forceinline__ uint32_t rotl16x2(uint32_t x, int k) {  // rotate both halves left by k (1..15)
  const uint32_t hm = (0xffffu << k) & 0xffffu;
  const uint32_t hm2 = hm | (hm << 16);
  return ((x << k) & hm2) | ((x >> (16 - k)) & ~hm2);
}

// Byte permute with the selector's sign-replicate mode (PTX prmt.b32: nibble bit 3 set = the result byte is the
// selected byte's msb replicated eight times).  The host branch is the model tests/test_device_arith.py compiles.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
#else
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) {
    const uint32_t n = (sel >> (4 * k)) & 0xfu;
    uint32_t byte = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
    if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
    r |= byte << (8 * k);
  }
  return r;
#endif
}

// The sign bits are collected two ring pixels at a time: one PRMT turns the four signs of u[i], u[i+8] (bytes 1
// and 3 of each word) into four 0x00/0xff bytes, one LOP3 keeps bit i of each byte.  Result: dark mask in the low
// half, bright mask in the high half, ring pixel i at bit i -- one ALU instruction per ring pixel instead of the
// shift + or of a serial mask (the ALU pipe is the busiest pipe of the tile kernel; 0.2875 -> 0.2817 ms).
template <typename Ptr>
__device__ __forceinline__ bool fast9_prmt_at(Ptr c, int pitch, int t) {
  const uint32_t v = c[0];
  const uint32_t K = ((v + (uint32_t)t) << 16) + ((uint32_t)t - v);
  uint32_t u[16];
  int i = 0;
#define PB_RING(dx, dy) u[i++] = (uint32_t)c[(dy) * pitch + (dx)] * 0xffff0001u + K;
  PB_RING(-1, -3) PB_RING(0, -3) PB_RING(1, -3) PB_RING(2, -2)
  PB_RING(3, -1) PB_RING(3, 0) PB_RING(3, 1) PB_RING(2, 2)
  PB_RING(1, 3) PB_RING(0, 3) PB_RING(-1, 3) PB_RING(-2, 2)
  PB_RING(-3, 1) PB_RING(-3, 0) PB_RING(-3, -1) PB_RING(-2, -2)
#undef PB_RING
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) m |= prmt(u[k], u[k + 8], 0xfbd9u) & (0x01010101u << k);
  uint32_t r = m & rotl16x2(m, 1);   // runs of 2
  r &= rotl16x2(r, 2);               // runs of 4
  r &= rotl16x2(r, 4);               // runs of 8
  r &= __byte_perm(m, 0u, 0x2301);   // ... preceded by a ninth
  return r != 0u;
}

// ---------------------------------------------------------------------------------------------
// SWAR pre-tests of the tile kernel: four pixels per 32-bit word, flags in bit 7 of every byte.
// A 9-arc of the 16-pixel ring always contains two neighbouring compass pixels (ring 1,5,9,13) and two
// neighbouring diagonal pixels (ring 3,7,11,15), so a corner needs |p-c| > t on (N or S), on (E or W), on
// (NE or SW) and on (SE or NW).  The tests only have to be SUPERSETS of the corners: the exact test follows.
// ---------------------------------------------------------------------------------------------
// bit 7 of every byte: (byte of d) > t, for a uniform t in 0..255; k7 = (0x7f - (t & 0x7f)) * 0x01010101.
// With s = low 7 bits + k7 (bit 7 of s: low7 > t7): t < 128 needs d | s, t >= 128 needs d & s; `m_or` is
// all-ones / zero accordingly, so both cases are one three-input logic op.
__device__ __forceinline__ uint32_t bytes_gt(uint32_t d, uint32_t k7, uint32_t m_or) {
  const uint32_t s = (d & 0x7f7f7f7fu) + k7;
  return (d & s) | ((d | s) & m_or);
}
// Superset form of bytes_gt for t < 128, K = (0x7f - t) * 0x01010101: bit 7 of d | (d + K).  Bytes >= 128 pass
// through d itself; for bytes < 128 the byte sum d + (127 - t) + carry-in stays below 256, so bit 7 says
// d + carry-in > t.  A carry from the byte below (that byte is >= 129 + t) can only turn a tie d == t into a
// pass, never a pass into a fail: fine for a pre-test whose survivors are tested exactly (1 extra candidate in
// 200 000 on the synthetic set), and it saves the per-byte masking.
__device__ __forceinline__ uint32_t bytes_gt_loose(uint32_t d, uint32_t K) { return d | (d + K); }

template <bool T_LT_128>
__device__ __forceinline__ uint32_t swar_gt(uint32_t a, uint32_t b, uint32_t K) {
  const uint32_t d = __vabsdiffu4(a, b);
  return T_LT_128 ? bytes_gt_loose(d, K) : bytes_gt(d, K, 0u);
}
__device__ __forceinline__ uint32_t swar_k(int t) { return (0x7fu - (uint32_t)(t & 0x7f)) * 0x01010101u; }

// Compass test of R rows of one word column `c` (row pitch PW words): rows 0..R+5 of the column and the words
// left/right of rows 3..R+2.  Bit 8k+i of the result: pixel k of row i+3 passes (N or S) and (E or W).
// |I[q+3]-I[q]| > t serves as the S test of row q and the N test of row q+3.
template <bool T_LT_128, int PW, int R>
__device__ __forceinline__ uint32_t pretest_rows(const uint32_t* c, int t) {
  static_assert(R >= 1 && R <= 8, "one flag bit per row inside each byte");
  const uint32_t K = swar_k(t);
  uint32_t I[R + 6];
#pragma unroll
  for (int q = 0; q < R + 6; ++q) I[q] = c[q * PW];
  uint32_t V[R + 3];
#pragma unroll
  for (int q = 0; q < R + 3; ++q) V[q] = swar_gt<T_LT_128>(I[q + 3], I[q], K);
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const uint32_t C = I[i + 3];
    const uint32_t* cr = c + (i + 3) * PW;
    const uint32_t e = swar_gt<T_LT_128>(__funnelshift_r(C, cr[1], 24), C, K);   // pixels x+3 .. x+6 against x .. x+3
    const uint32_t w = swar_gt<T_LT_128>(__funnelshift_r(cr[-1], C, 8), C, K);   // pixels x-3 .. x
    const uint32_t m = (V[i] | V[i + 3]) & (e | w);
    acc |= (m >> (7 - i)) & (0x01010101u << i);
  }
  return acc;
}

// Does ANY pixel of the R rows pass (N or S)?  The cheap first stage of the cascade: one compare per row and
// word, no horizontal neighbours.  (The OR over rows q of V[q] covers V[i] | V[i+3] of every row i.)
template <bool T_LT_128, int PW, int R>
__device__ __forceinline__ bool ns_any(const uint32_t* c, int t) {
  const uint32_t K = swar_k(t);
  uint32_t I[R + 6];
#pragma unroll
  for (int q = 0; q < R + 6; ++q) I[q] = c[q * PW];
  uint32_t any = 0;
#pragma unroll
  for (int q = 0; q < R + 3; ++q) any |= swar_gt<T_LT_128>(I[q + 3], I[q], K);
  return (any & 0x80808080u) != 0u;
}

// Diagonal test in the layout of pretest_rows: bit 8k+i = pixel k of row i+3 passes (NE or SW) and (SE or NW),
// the ring pixels (+-2, +-2).  Rows 1..R+4 of the word column and of its two neighbours.
template <bool T_LT_128, int PW, int R>
__device__ __forceinline__ uint32_t diag_rows(const uint32_t* c, int t) {
  const uint32_t K = swar_k(t);
  uint32_t E2[R + 4], W2[R + 4];  // rows 1..R+4 shifted by +2 / -2 pixels
#pragma unroll
  for (int q = 0; q < R + 4; ++q) {
    const uint32_t* cr = c + (q + 1) * PW;
    E2[q] = __funnelshift_r(cr[0], cr[1], 16);   // pixels x+2 .. x+5
    W2[q] = __funnelshift_r(cr[-1], cr[0], 16);  // pixels x-2 .. x+1
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const uint32_t C = c[(i + 3) * PW];
    // row i+1 is two rows above the centre row i+3 (index i of E2/W2), row i+5 two rows below (index i+4)
    const uint32_t a = swar_gt<T_LT_128>(E2[i], C, K) | swar_gt<T_LT_128>(W2[i + 4], C, K);      // NE | SW
    const uint32_t b = swar_gt<T_LT_128>(E2[i + 4], C, K) | swar_gt<T_LT_128>(W2[i], C, K);      // SE | NW
    const uint32_t m = a & b;
    acc |= (m >> (7 - i)) & (0x01010101u << i);
  }
  return acc;
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
