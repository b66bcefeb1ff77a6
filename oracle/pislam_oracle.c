/* TEST INFRASTRUCTURE ONLY -- see pislam_oracle.h.
 *
 * Plain-C scalar restatement of the reference ORB front-end.  Each function states the
 * *rule* the reference's NEON / bit-twiddling code implements rather than
 * mirroring its instruction sequence; tests/test_oracle.py pins every one of them bit-for-bit
 * against the reference's own headers (oracle/_ref) and against tests/golden/.
 *
 * Build: see oracle/Makefile (-ffp-contract=off is REQUIRED: the angle polynomial and the
 * pattern rotation are sequences of individually rounded fp32 operations).
 *
 * Images are addressed as linear memory, p[y * vstep + x], exactly like the reference's
 * `img[y][x]` on a `uint8_t (*)[vstep]`: an x beyond the row pitch lands in the next row.
 */
#include "pislam_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "brief_pattern.h"

/* ------------------------------------------------------------------ Util.h:27-45 */
uint32_t po_encode_fast(uint32_t score, uint32_t x, uint32_t y) { return (score << 24) | (x << 12) | y; }
uint32_t po_rencode_fast_score(uint32_t score, uint32_t e) { return (score << 24) | (e & 0xffffffu); }
uint32_t po_decode_fast_x(uint32_t e) { return (e >> 12) & 0xfffu; }
uint32_t po_decode_fast_y(uint32_t e) { return e & 0xfffu; }
uint32_t po_decode_fast_score(uint32_t e) { return e >> 24; }

/* ------------------------------------------------------------------ Fast.h:54-158
 * Ring in clockwise order starting at (-1,-3): Fast.h:66-128 (d0 MSB..LSB, then d1 MSB..LSB). */
static const int8_t RING[16][2] = {
    {-1, -3}, {0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},   {2, 2},
    {1, 3},   {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}};

/* 1 iff the 16-bit circular word has >= 9 consecutive set bits */
static int arc9(unsigned m) {
  unsigned r = m | (m << 16); /* unrolled ring */
  r &= r >> 1;                /* runs >= 2 */
  r &= r >> 2;                /* runs >= 4 */
  r &= r >> 4;                /* runs >= 8 */
  r &= r >> 1;                /* runs >= 9 */
  return (r & 0xffffu) != 0;
}

static uint8_t fast9_pixel(const uint8_t* c, long vstep, int t) {
  int v = *c;
  int hi = v + t > 255 ? 255 : v + t; /* vqaddq_u8, Fast.h:63 */
  int lo = v - t < 0 ? 0 : v - t;     /* vqsubq_u8, Fast.h:64 */
  unsigned brighter = 0, darker = 0;
  for (int i = 0; i < 16; ++i) {
    int p = c[RING[i][1] * vstep + RING[i][0]];
    brighter |= (unsigned)(p > hi) << i;
    darker |= (unsigned)(p < lo) << i;
  }
  /* Fast.h:130-147 picks ONE of the two polarities (dark unless some opposite pair is
   * entirely "not darker") and looks for 9 contiguous; a 9-arc of one polarity forces that
   * choice, so this equals the plain either-polarity FAST-9 test. */
  return (arc9(brighter) || arc9(darker)) ? 0xff : 0x00;
}

void po_fast_detect(const uint8_t* img, uint8_t* out, int vstep, int border, int width, int height,
                    int threshold) {
  int t = threshold & 0xff; /* vdupq_n_u8 keeps the low byte, Fast.h:58 */
  for (int y = border; y < height - border; ++y) {
    /* 16-pixel blocks: the last one overruns width-border by up to 15 columns, Fast.h:61 */
    for (int x0 = border; x0 < width - border; x0 += 16)
      for (int x = x0; x < x0 + 16; ++x)
        out[(long)y * vstep + x] = fast9_pixel(img + (long)y * vstep + x, vstep, t);
    if (width % 16 != 0) { /* Fast.h:153-156 */
      out[(long)y * vstep + width] = 0;
      out[(long)y * vstep + width + 1] = 0;
    }
  }
}

/* ------------------------------------------------------------------ Harris.h:37-69, 80-248 */
static inline int asr1(int v) { return v >> 1; } /* arithmetic: floor halving (vhsub/vhadd) */

uint8_t po_harris_score_sobel(const uint8_t* img, int vstep, int x, int y, int32_t threshold) {
  int P[8][8];
  for (int r = 0; r < 8; ++r)
    for (int j = 0; j < 8; ++j) P[r][j] = img[(long)(y - 3 + r) * vstep + (x - 3 + j)];

  int dy[6][6], dx[6][6];
  /* vertical central difference, then [1 2 1]/4 smoothing along x (Harris.h:123-135) */
  for (int n = 0; n < 6; ++n) {
    int d[8];
    for (int j = 0; j < 8; ++j) d[j] = asr1(P[n + 2][j] - P[n][j]);
    for (int j = 0; j < 6; ++j) dy[n][j] = asr1(d[j + 1] + asr1(d[j] + d[j + 2]));
  }
  /* horizontal central difference, then [1 2 1]/4 smoothing along y (Harris.h:139-162) */
  {
    int e[8][6];
    for (int r = 0; r < 8; ++r)
      for (int j = 0; j < 6; ++j) e[r][j] = asr1(P[r][j + 2] - P[r][j]);
    for (int n = 0; n < 6; ++n)
      for (int j = 0; j < 6; ++j) dx[n][j] = asr1(asr1(e[n][j] + e[n + 2][j]) + e[n + 1][j]);
  }
  /* structure tensor: row pairs are summed in 16-bit lanes first (Harris.h:166-213):
   * xx/yy as uint16, xy as int16 (wrap-around), then widened */
  uint32_t sxx = 0, syy = 0;
  int32_t sxy = 0;
  for (int n = 0; n < 6; n += 2)
    for (int j = 0; j < 6; ++j) {
      uint16_t xx = (uint16_t)(dx[n][j] * dx[n][j] + dx[n + 1][j] * dx[n + 1][j]);
      uint16_t yy = (uint16_t)(dy[n][j] * dy[n][j] + dy[n + 1][j] * dy[n + 1][j]);
      int16_t xy = (int16_t)(uint16_t)(dx[n][j] * dy[n][j] + dx[n + 1][j] * dy[n + 1][j]);
      sxx += xx;
      syy += yy;
      sxy += xy;
    }
  uint32_t Ixx = sxx >> 4, Iyy = syy >> 4; /* Harris.h:243-244 */
  int32_t Ixy = sxy >> 4;                  /* Harris.h:245, arithmetic */

  /* harrisEval, k = 1/16, everything mod 2^32 (Harris.h:41-57) */
  uint32_t tr = Ixx + Iyy;
  tr = (tr * tr) >> 4;
  uint32_t det = Ixx * Iyy - (uint32_t)Ixy * (uint32_t)Ixy;
  int32_t score = (int32_t)(det - tr);
  if (threshold < score) {
    float f = (float)score; /* vcvt_f32_s32: round to nearest even */
    uint32_t bits;
    memcpy(&bits, &f, 4);
    return (uint8_t)((bits >> 20) & 0xff); /* 5 exponent + 3 mantissa bits, Harris.h:58-66 */
  }
  return 0;
}

void po_fast_score_harris(const uint8_t* img, uint8_t* out, int vstep, int border, int width,
                          int height, int32_t threshold) {
  for (int y = border; y < height - border; ++y)
    for (int x = border; x < width - border; ++x) /* stops at width-border: Fast.h:172 */
      if (out[(long)y * vstep + x])
        out[(long)y * vstep + x] = po_harris_score_sobel(img, vstep, x, y, threshold);
}

/* ------------------------------------------------------------------ Fast.h:196-355
 * Position-independent form of the decision tree at Fast.h:264-310: a pixel survives iff it
 * is >= each of the 4 neighbours that precede it in raster order and > each of the 4 that
 * follow.  Within a 2x2 cell at most one pixel can satisfy this. */
static int nms_survives(const uint8_t* o, long vstep) {
  int s = *o;
  if (!s) return 0;
  return s >= o[-vstep - 1] && s >= o[-vstep] && s >= o[-vstep + 1] && s >= o[-1] && s > o[1] &&
         s > o[vstep - 1] && s > o[vstep] && s > o[vstep + 1];
}

static int cmp_u32(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return (x > y) - (x < y);
}

int po_fast_extract(const uint8_t* out, int vstep, int border, int width, int height,
                    int logBucketSize, int bucketLimit, uint32_t* kp, int cap) {
  int n = 0;
  const int bucketSize = 1 << logBucketSize;
  if (width - 2 * border <= 0) return 0;
  const int numBuckets = (width - 2 * border - 1) / bucketSize + 1; /* Fast.h:201 */
  uint32_t* band = NULL; /* survivors of the current band of bucketSize rows, per bucket */
  int* bandCount = NULL;
  int bandCap = 0;
  if (logBucketSize) {
    /* a bucket holds at most (bucketSize/2)^2 cells, one survivor each */
    bandCap = (bucketSize / 2) * (bucketSize / 2);
    if (bandCap < 1) bandCap = 1;
    band = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)numBuckets * bandCap);
    bandCount = (int*)calloc((size_t)numBuckets, sizeof(int));
  }
  int status = 0;
  for (int y = border; y < height - border && status == 0; y += 2) {
    for (int x = border; x < width - border; x += 2) {
      uint32_t enc = 0;
      int found = 0;
      for (int k = 0; k < 4 && !found; ++k) {
        int px = x + (k & 1), py = y + (k >> 1);
        const uint8_t* o = out + (long)py * vstep + px;
        if (nms_survives(o, vstep)) {
          enc = po_encode_fast(*o, (uint32_t)px, (uint32_t)py);
          found = 1;
        }
      }
      if (!found) continue;
      if (!logBucketSize) { /* Fast.h:319-320 */
        if (n >= cap) { status = -2; break; }
        kp[n++] = enc;
      } else {
        int b = (x - border) / bucketSize; /* Fast.h:316, the CELL's x */
        band[(size_t)b * bandCap + bandCount[b]++] = enc;
      }
    }
    /* flush when the next cell row starts a new band, or after the last row
     * (Fast.h:211-226, 345-352): buckets in index order, each keeping its bucketLimit
     * largest encoded values (Fast.h:321-341), emitted ascending */
    if (logBucketSize && status == 0) {
      int ynext = y + 2;
      if (ynext >= height - border || (ynext - border) % bucketSize == 0) {
        for (int b = 0; b < numBuckets && status == 0; ++b) {
          int c = bandCount[b];
          uint32_t* v = band + (size_t)b * bandCap;
          qsort(v, (size_t)c, sizeof(uint32_t), cmp_u32);
          int first = c > bucketLimit ? c - bucketLimit : 0;
          for (int i = first; i < c; ++i) {
            if (n >= cap) { status = -2; break; }
            kp[n++] = v[i];
          }
          bandCount[b] = 0;
        }
      }
    }
  }
  free(band);
  free(bandCount);
  return status ? status : n;
}

/* ------------------------------------------------------------------ Orb.h:80-308
 * Half-width of the intensity-centroid disc per |dy| (from the masks/weights at Orb.h:118-126
 * and the row macros at Orb.h:163-286).  NOT OpenCV's umax table. */
static const int8_t DISC_U[16] = {15, 15, 15, 15, 15, 15, 14, 14, 13, 13, 12, 11, 10, 9, 7, 5};

static void centroid_moments(const uint8_t* img, long vstep, int x, int y, int32_t* m10, int32_t* m01) {
  int32_t sx = 0, sy = 0;
  for (int dy = -15; dy <= 15; ++dy) {
    int u = DISC_U[dy < 0 ? -dy : dy];
    const uint8_t* row = img + (long)(y + dy) * vstep + x;
    for (int dx = -u; dx <= u; ++dx) {
      int v = row[dx];
      sx += dx * v;
      sy += dy * v;
    }
  }
  *m10 = sx;
  *m01 = sy;
}

int po_orb_centroids(const uint8_t* img, int vstep, const uint32_t* kp, int n, int32_t* centroids) {
  int size = (2 * n + 7) & ~7; /* Orb.h:113-114 */
  memset(centroids, 0, sizeof(int32_t) * (size_t)size);
  for (int i = 0; i < n; ++i) {
    int32_t mx, my;
    centroid_moments(img, vstep, (int)po_decode_fast_x(kp[i]), (int)po_decode_fast_y(kp[i]), &mx, &my);
    int base = (i >> 2) * 8 + (i & 3); /* groups of [x0 x1 x2 x3 y0 y1 y2 y3], Orb.h:298-304 */
    centroids[base] = mx;
    centroids[base + 4] = my;
  }
  return size;
}

/* ------------------------------------------------------------------ Orb.h:310-387 */
float po_recpe(float x) { /* ARM FPRecipEstimate (VRECPE.F32), flush-to-zero */
  uint32_t u, o;
  memcpy(&u, &x, 4);
  uint32_t sign = u & 0x80000000u, e = (u >> 23) & 0xff, frac = u & 0x7fffffu;
  if (e == 0xff) o = frac ? 0x7fc00000u : sign;
  else if (e == 0) o = sign | 0x7f800000u;
  else if (e >= 253) o = sign;
  else {
    uint32_t q = 256 + (frac >> 15);               /* 9-bit scaled operand in [256, 512) */
    uint32_t s = (524288u + (2 * q + 1)) / (2 * (2 * q + 1)); /* floor(2^19/(2q+1) + 1/2) */
    o = sign | ((253 - e) << 23) | ((s - 256) << 15);
  }
  float r;
  memcpy(&r, &o, 4);
  return r;
}

uint8_t po_angle_bin(int32_t m10, int32_t m01) {
  float ax = fabsf((float)m10), ay = fabsf((float)m01);
  float zmax = ax > ay ? ax : ay, zmin = ax < ay ? ax : ay;
  float z = zmin * po_recpe(zmax); /* 0 * inf = NaN for the zero-moment patch */
  const float c0 = (float)(256 * 14.999998), c1 = (float)(256 * 4.723436), c2 = (float)(256 * 1.266240);
  /* every operation individually rounded to fp32 (Orb.h:347) */
  float t0 = c2 * z;
  float t1 = c1 + t0;
  float t2 = z - 1.0f;
  float t3 = t2 * t1;
  float t4 = c0 - t3;
  float t = z * t4;
  int32_t a;
  if (t != t) a = 0; /* vcvtq_s32_f32: NaN -> 0, truncation, saturation */
  else if (t >= 2147483648.0f) a = INT32_MAX;
  else if (t <= -2147483648.0f) a = INT32_MIN;
  else a = (int32_t)t;
  int64_t absx = m10 < 0 ? -(int64_t)m10 : m10, absy = m01 < 0 ? -(int64_t)m01 : m01;
  if (absx > absy) { /* Orb.h:357-375 */
    if ((m10 ^ m01) < 0) a = -a;
    if (m10 < 0) a += 256 * 60;
    else if (a < 0) a += 256 * 120;
  } else {
    if ((m10 ^ m01) >= 0) a = -a;
    a += (m01 >= 0) ? 256 * 30 : 256 * 90;
  }
  a >>= 10;
  if (!(0 <= a && a < 30)) a = 0;
  return (uint8_t)a;
}

int po_atan2(const int32_t* xys, int nxys, uint8_t* angles) {
  int n = 0;
  for (int g = 0; g + 8 <= nxys; g += 8)
    for (int i = 0; i < 4; ++i) angles[n++] = po_angle_bin(xys[g + i], xys[g + 4 + i]);
  return n;
}

/* ------------------------------------------------------------------ Brief.h:28-53 */
static int8_t g_table[30][256][4];
static pthread_once_t g_table_once = PTHREAD_ONCE_INIT;

static int rot_clamp(float v) {
  int r = (int)roundf(v); /* exact half-away-from-zero */
  return r < -15 ? -15 : (r > 15 ? 15 : r);
}

static void build_table(void) {
  for (int rot = 0; rot < 30; ++rot) {
    float c, s;
    memcpy(&c, &PISLAM_BRIEF_COS_BITS[rot], 4);
    memcpy(&s, &PISLAM_BRIEF_SIN_BITS[rot], 4);
    for (int k = 0; k < 256; ++k)
      for (int e = 0; e < 2; ++e) {
        float dx = (float)PISLAM_BRIEF_BASE[k][2 * e], dy = (float)PISLAM_BRIEF_BASE[k][2 * e + 1];
        float cx = c * dx, sy = s * dy, sx = s * dx, cy = c * dy; /* each rounded to fp32 */
        g_table[rot][k][2 * e] = (int8_t)rot_clamp(cx - sy);
        g_table[rot][k][2 * e + 1] = (int8_t)rot_clamp(sx + cy);
      }
  }
}

const int8_t* po_brief_table(void) {
  pthread_once(&g_table_once, build_table);
  return &g_table[0][0][0];
}

void po_brief_describe(const uint8_t* img, int vstep, int x, int y, int rot, int words, uint32_t* desc) {
  if (rot < 0 || rot >= 30) return; /* Brief.h:641-732: no case -> untouched */
  po_brief_table();
  const uint8_t* base = img + (long)y * vstep + x;
  for (int w = 0; w < words; ++w) {
    uint32_t bits = 0;
    for (int k = 0; k < 32; ++k) {
      const int8_t* p = g_table[rot][32 * w + k];
      if (base[(long)p[1] * vstep + p[0]] < base[(long)p[3] * vstep + p[2]]) bits |= 1u << k;
    }
    desc[w] = bits;
  }
}

void po_orb_compute(const uint8_t* img, int vstep, const uint32_t* kp, int n, int words, uint32_t* desc) {
  memset(desc, 0, sizeof(uint32_t) * (size_t)n * words); /* vector::resize value-initialises, Orb.h:423 */
  for (int i = 0; i < n; ++i) {
    int x = (int)po_decode_fast_x(kp[i]), y = (int)po_decode_fast_y(kp[i]);
    int32_t mx, my;
    centroid_moments(img, vstep, x, y, &mx, &my);
    po_brief_describe(img, vstep, x, y, po_angle_bin(mx, my), words, desc + (size_t)i * words);
  }
}

/* ------------------------------------------------------------------ demo/demo.cpp:77-101 */
int po_extract_pyramid(const uint8_t* stack, int vstep, int nlevels, const int* widths,
                       const int* heights, const int* row_offsets, int border, int fast_threshold,
                       int32_t harris_threshold, int logBucketSize, int bucketLimit, int words,
                       int stacked_y, uint8_t* scratch_out, uint32_t* kp, uint8_t* kp_level,
                       uint32_t* desc, int cap) {
  int n = 0;
  long rows = 0;
  for (int l = 0; l < nlevels; ++l)
    if (row_offsets[l] + heights[l] > rows) rows = row_offsets[l] + heights[l];
  memset(scratch_out, 0, (size_t)rows * vstep);
  for (int l = 0; l < nlevels; ++l) {
    const uint8_t* img = stack + (long)row_offsets[l] * vstep;
    uint8_t* out = scratch_out + (long)row_offsets[l] * vstep;
    po_fast_detect(img, out, vstep, border, widths[l], heights[l], fast_threshold);
    po_fast_score_harris(img, out, vstep, border, widths[l], heights[l], harris_threshold);
    int m = po_fast_extract(out, vstep, border, widths[l], heights[l], logBucketSize, bucketLimit,
                            kp + n, cap - n);
    if (m < 0) return m;
    if (desc) /* descriptors are position-only, so per-level == whole-stack orbCompute */
      po_orb_compute(img, vstep, kp + n, m, words, desc + (size_t)n * words);
    for (int i = 0; i < m; ++i) {
      if (kp_level) kp_level[n + i] = (uint8_t)l;
      if (stacked_y) /* demo.cpp:91-97 */
        kp[n + i] = po_encode_fast(po_decode_fast_score(kp[n + i]), po_decode_fast_x(kp[n + i]),
                                   po_decode_fast_y(kp[n + i]) + (uint32_t)row_offsets[l]);
    }
    n += m;
  }
  return n;
}

/* ------------------------------------------------------------------ pyramid helpers
 * Gaussian.h:48-72 (the separable [1 4 6 4 1]/16 kernel built from rounding halving adds) is hand-written
 * ARM assembly and cannot be built here; this follows the scalar model the reference's own test checks it
 * against, test/GaussianTest.cpp:32,159-215: vertical pass, then horizontal pass, reflect-101 borders. */
static inline int rhadd(int a, int b) { return (a >> 1) + (b >> 1) + ((a | b) & 1); }
static inline int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
static inline int gauss5(int a, int b, int c, int d, int e) {
  int x = rhadd(a, e);
  x = rhadd(x, c);
  x = rhadd(x, c);
  return rhadd(x, rhadd(b, d));
}

void po_gaussian5x5(const uint8_t* img, uint8_t* out, int vstep, int width, int height) {
  uint8_t* tmp = (uint8_t*)malloc((size_t)width * height);
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      int r[5];
      for (int k = 0; k < 5; ++k) r[k] = img[(long)reflect101(y + k - 2, height) * vstep + x];
      tmp[(long)y * width + x] = (uint8_t)gauss5(r[0], r[1], r[2], r[3], r[4]);
    }
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      int r[5];
      for (int k = 0; k < 5; ++k) r[k] = tmp[(long)y * width + reflect101(x + k - 2, width)];
      out[(long)y * vstep + x] = (uint8_t)gauss5(r[0], r[1], r[2], r[3], r[4]);
    }
  free(tmp);
}

/* Bilinear.h:42-150 (7/8) and :165-321 (13/16); scalar models test/BilinearTest.cpp:35,171-233.
 * Blocks of B x B input pixels (B = 8 / 16, the image is padded to a multiple of B) give O x O outputs
 * (O = 7 / 13) with fixed-point weights /256 and rounding shifts; block (bi,bj) lands at (O*bi, O*bj).
 * The 7/8 variant stores 8 bytes per block row, so one extra zero byte follows the last block of a row
 * (Bilinear.h:139-145: lane 7 of every row is 0). */
static inline int rshr8(int a) { return (a >> 8) + ((a >> 7) & 1); }
static int map13(int i) { return i + (i > 3) + (i > 8); } /* BilinearTest.cpp:196-204 */

static void bilinear_blocks(const uint8_t* img, uint8_t* out, int vstep, int width, int height, int B, int O,
                            const int* f, int use_map13) {
  for (int i = 0, oi = 0; i < height; i += B, oi += O)
    for (int j = 0, oj = 0; j < width; j += B, oj += O)
      for (int y = 0; y < O; ++y)
        for (int x = 0; x < O; ++x) {
          const int yy = i + (use_map13 ? map13(y) : y), xx = j + (use_map13 ? map13(x) : x);
          const uint8_t* p = img + (long)yy * vstep + xx;
          const int h0 = rshr8(p[0] * f[x] + p[1] * f[O - 1 - x]);
          const int h1 = rshr8(p[vstep] * f[x] + p[vstep + 1] * f[O - 1 - x]);
          out[(long)(oi + y) * vstep + (oj + x)] = (uint8_t)rshr8(h0 * f[y] + h1 * f[O - 1 - y]);
        }
}

void po_bilinear7_8(const uint8_t* img, uint8_t* out, int vstep, int width, int height) {
  static const int f[7] = {238, 201, 165, 128, 91, 55, 18};
  if (width <= 0 || height <= 0) return;
  uint8_t* src = (uint8_t*)malloc((size_t)((height + 7) / 8 * 8 + 1) * vstep); /* img and out may alias */
  memcpy(src, img, (size_t)((height + 7) / 8 * 8) * vstep);
  bilinear_blocks(src, out, vstep, width, height, 8, 7, f, 0);
  const int nby = (height + 7) / 8, nbx = (width + 7) / 8;
  for (int y = 0; y < 7 * nby; ++y) out[(long)y * vstep + 7 * nbx] = 0;
  free(src);
}

void po_bilinear13_16(const uint8_t* img, uint8_t* out, int vstep, int width, int height) {
  static const int f[13] = {226, 167, 108, 49, 246, 187, 128, 69, 10, 207, 138, 89, 30};
  if (width <= 0 || height <= 0) return;
  uint8_t* src = (uint8_t*)malloc((size_t)((height + 15) / 16 * 16 + 1) * vstep);
  memcpy(src, img, (size_t)((height + 15) / 16 * 16) * vstep);
  bilinear_blocks(src, out, vstep, width, height, 16, 13, f, 1);
  free(src);
}

/* ------------------------------------------------------------------ pyramid builder
 * Not reference code: the reference leaves the x1/1.2 pyramid to its caller (README.md:28-31).  Restates
 * pislam_b200/pyramid.py:_resize_bilinear (the definition this repo's synthetic pyramids are built with). */
static void resize_taps(int n_out, int n_in, int i, int* i0, int* i1, int* f) {
  long pos = ((long)(2 * i + 1) * n_in * 128) / n_out - 128; /* 8.8 fixed point, floor division of a positive value */
  const long hi = (long)(n_in - 1) * 256;
  if (pos < 0) pos = 0;
  if (pos > hi) pos = hi;
  *i0 = (int)(pos >> 8);
  *f = (int)(pos & 255);
  *i1 = *i0 + 1 < n_in ? *i0 + 1 : n_in - 1;
}

void po_resize_bilinear(const uint8_t* src, int src_pitch, int sw, int sh, uint8_t* dst, int dst_pitch, int dw,
                        int dh) {
  for (int y = 0; y < dh; ++y) {
    int y0, y1, fy;
    resize_taps(dh, sh, y, &y0, &y1, &fy);
    const uint8_t* r0 = src + (long)y0 * src_pitch;
    const uint8_t* r1 = src + (long)y1 * src_pitch;
    for (int x = 0; x < dw; ++x) {
      int x0, x1, fx;
      resize_taps(dw, sw, x, &x0, &x1, &fx);
      const int top = r0[x0] * (256 - fx) + r0[x1] * fx;
      const int bot = r1[x0] * (256 - fx) + r1[x1] * fx;
      dst[(long)y * dst_pitch + x] = (uint8_t)((top * (256 - fy) + bot * fy + (1 << 15)) >> 16);
    }
  }
}

void po_build_pyramid(const uint8_t* base, int base_pitch, int blur, int vstep, int nlevels, const int* widths,
                      const int* heights, const int* row_offsets, uint8_t* stack) {
  const int W = widths[0], H = heights[0];
  uint8_t* tmp = NULL;
  if (blur) {
    tmp = (uint8_t*)malloc((size_t)W * H);
    /* po_gaussian5x5 takes one pitch for input and output: blur through a pitch-W copy */
    uint8_t* in = (uint8_t*)malloc((size_t)W * H);
    for (int y = 0; y < H; ++y) memcpy(in + (long)y * W, base + (long)y * base_pitch, W);
    po_gaussian5x5(in, tmp, W, W, H);
    free(in);
    base = tmp;
    base_pitch = W;
  }
  for (int l = 0; l < nlevels; ++l) {
    uint8_t* lvl = stack + (long)row_offsets[l] * vstep;
    po_resize_bilinear(base, base_pitch, W, H, lvl, vstep, widths[l], heights[l]);
    for (int y = 0; y < heights[l]; ++y) memset(lvl + (long)y * vstep + widths[l], 0, vstep - widths[l]);
  }
  free(tmp);
}

/* ------------------------------------------------------------------ all-host-cores baseline */
typedef struct {
  const uint8_t* stacks;
  long frame_stride;
  int nframes, vstep, nlevels;
  const int *widths, *heights, *row_offsets;
  int border, fast_threshold;
  int32_t harris_threshold;
  int logBucketSize, bucketLimit, words, cap;
  int* counts;
  uint32_t *kp, *desc;
  int next; /* work counter */
  pthread_mutex_t mu;
  int status;
} batch_t;

static void* batch_worker(void* arg) {
  batch_t* b = (batch_t*)arg;
  long rows = 0;
  for (int l = 0; l < b->nlevels; ++l)
    if (b->row_offsets[l] + b->heights[l] > rows) rows = b->row_offsets[l] + b->heights[l];
  uint8_t* scratch = (uint8_t*)malloc((size_t)rows * b->vstep);
  uint32_t* kp_local = b->kp ? NULL : (uint32_t*)malloc(sizeof(uint32_t) * (size_t)b->cap);
  uint32_t* desc_local = b->desc ? NULL : (uint32_t*)malloc(sizeof(uint32_t) * (size_t)b->cap * b->words);
  for (;;) {
    pthread_mutex_lock(&b->mu);
    int f = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (f >= b->nframes) break;
    uint32_t* kp = b->kp ? b->kp + (size_t)f * b->cap : kp_local;
    uint32_t* desc = b->desc ? b->desc + (size_t)f * b->cap * b->words : desc_local;
    int n = po_extract_pyramid(b->stacks + f * b->frame_stride, b->vstep, b->nlevels, b->widths,
                               b->heights, b->row_offsets, b->border, b->fast_threshold,
                               b->harris_threshold, b->logBucketSize, b->bucketLimit, b->words, 0,
                               scratch, kp, NULL, desc, b->cap);
    b->counts[f] = n;
    if (n < 0) b->status = n;
  }
  free(scratch);
  free(kp_local);
  free(desc_local);
  return NULL;
}

int po_extract_batch(const uint8_t* stacks, long frame_stride, int nframes, int vstep, int nlevels,
                     const int* widths, const int* heights, const int* row_offsets, int border,
                     int fast_threshold, int32_t harris_threshold, int logBucketSize,
                     int bucketLimit, int words, int cap, int nthreads, int* counts, uint32_t* kp,
                     uint32_t* desc) {
  batch_t b = {stacks, frame_stride, nframes, vstep, nlevels, widths, heights, row_offsets, border,
               fast_threshold, harris_threshold, logBucketSize, bucketLimit, words, cap, counts,
               kp, desc, 0, PTHREAD_MUTEX_INITIALIZER, 0};
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, batch_worker, &b);
  for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
  return b.status;
}
