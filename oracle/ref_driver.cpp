// TEST INFRASTRUCTURE ONLY -- never linked into libpislam_b200.so.
//
// C-ABI wrapper around the UNMODIFIED reference templates (compiled from
// /root/reference/include where they lie; nothing is copied into this repo):
//   fastDetect        Fast.h:54-158    (via the build-time substitution of the two
//                                       `vceq.u8` inline-asm lines, see Makefile)
//   fastScoreHarris   Fast.h:166-180 -> Harris.h:80-248
//   fastExtract       Fast.h:196-355
//   orbCentroids      Orb.h:80-308
//   atan2             Orb.h:310-387
//   orbCompute        Orb.h:396-441 -> Brief.h:57-733
//   briefDescribe     Brief.h:637-733
//   bilinear7_8 / bilinear13_16   Bilinear.h:42-150, 165-321 (pyramid helpers, not on the hot path)
// The reference takes vstep/border/bucket/words as template arguments; this
// driver instantiates a fixed menu of them and dispatches at run time. A
// combination outside the menu returns -1.
//
// Built by oracle/Makefile into oracle/_ref/libpislam_ref.so (git-ignored).
#include <cstdint>
#include <cstring>
#include <vector>

#include "arm_neon.h"   // oracle/neon_shim
#include "Fast.h"       // oracle/_ref/Fast.h (generated), includes reference Harris.h/Util.h
#include "Orb.h"        // /root/reference/include/Orb.h, includes Brief.h
#include "Bilinear.h"   // /root/reference/include/Bilinear.h (pyramid helper)

#define REF_VSTEPS(X) X(640) X(1920) X(3840) X(128) X(96)
#define REF_BORDERS(X, V) X(V, 16) X(V, 9) X(V, 4)
#define REF_BUCKETS(X, V, B) X(V, B, 0, 5) X(V, B, 4, 3) X(V, B, 4, 5) X(V, B, 3, 2) X(V, B, 5, 4)

extern "C" {

int ref_fast_detect(const uint8_t* img, uint8_t* out, int vstep, int border, int width, int height,
                    int threshold) {
#define CASE(V, B)                                                                             \
  if (vstep == V && border == B) {                                                             \
    pislam::fastDetect<V, B>(width, height, (uint8_t(*)[V])img, (uint8_t(*)[V])out, threshold); \
    return 0;                                                                                  \
  }
#define PER_V(V) REF_BORDERS(CASE, V)
  REF_VSTEPS(PER_V)
#undef PER_V
#undef CASE
  return -1;
}

int ref_fast_score_harris(const uint8_t* img, uint8_t* out, int vstep, int border, int width,
                          int height, int32_t threshold) {
#define CASE(V, B)                                                                                  \
  if (vstep == V && border == B) {                                                                  \
    pislam::fastScoreHarris<V, B>(width, height, (uint8_t(*)[V])img, threshold, (uint8_t(*)[V])out); \
    return 0;                                                                                       \
  }
#define PER_V(V) REF_BORDERS(CASE, V)
  REF_VSTEPS(PER_V)
#undef PER_V
#undef CASE
  return -1;
}

int ref_harris_score_sobel(const uint8_t* img, int vstep, int x, int y, int32_t threshold) {
#define CASE(V) \
  if (vstep == V) return pislam::harrisScoreSobel<V>((uint8_t(*)[V])img, x, y, threshold);
  REF_VSTEPS(CASE)
#undef CASE
  return -1;
}

// returns number of keypoints appended (written to kp[0..n)), or -1 / -2 (cap too small)
int ref_fast_extract(const uint8_t* out, int vstep, int border, int width, int height,
                     int logBucketSize, int bucketLimit, uint32_t* kp, int cap) {
  std::vector<uint32_t> results;
  bool done = false;
#define CASE(V, B, L, K)                                                                     \
  if (!done && vstep == V && border == B && logBucketSize == L && bucketLimit == K) {        \
    pislam::fastExtract<V, B, L, K>(width, height, (uint8_t(*)[V])out, results);             \
    done = true;                                                                             \
  }
#define PER_B(V, B) REF_BUCKETS(CASE, V, B)
#define PER_V(V) REF_BORDERS(PER_B, V)
  REF_VSTEPS(PER_V)
#undef PER_V
#undef PER_B
#undef CASE
  if (!done) return -1;
  if ((int)results.size() > cap) return -2;
  if (!results.empty()) std::memcpy(kp, results.data(), results.size() * sizeof(uint32_t));
  return (int)results.size();
}

// centroids must hold (2n+7)&~7 int32 (reference layout, Orb.h:113-114,298-304)
int ref_orb_centroids(const uint8_t* img, int vstep, const uint32_t* kp, int n, int32_t* centroids) {
  std::vector<uint32_t> points(kp, kp + n);
#define CASE(V)                                                                      \
  if (vstep == V) {                                                                  \
    std::vector<int32_t> c = pislam::orbCentroids<V>((uint8_t(*)[V])img, points);    \
    if (!c.empty()) std::memcpy(centroids, c.data(), c.size() * sizeof(int32_t));    \
    return (int)c.size();                                                            \
  }
  REF_VSTEPS(CASE)
#undef CASE
  return -1;
}

int ref_atan2(const int32_t* xys, int nxys, uint8_t* angles) {
  std::vector<int32_t> v(xys, xys + nxys);
  std::vector<uint8_t> a = pislam::atan2(v);
  if (!a.empty()) std::memcpy(angles, a.data(), a.size());
  return (int)a.size();
}

int ref_orb_compute(const uint8_t* img, int vstep, const uint32_t* kp, int n, int words,
                    uint32_t* desc) {
  std::vector<uint32_t> points(kp, kp + n);
  std::vector<uint32_t> d;
#define CASE(V, W)                                                    \
  if (vstep == V && words == W) {                                     \
    pislam::orbCompute<V, W>((uint8_t(*)[V])img, points, d);          \
    if (!d.empty()) std::memcpy(desc, d.data(), d.size() * 4);        \
    return (int)d.size();                                             \
  }
  CASE(640, 8) CASE(1920, 8) CASE(3840, 8) CASE(128, 8) CASE(96, 8)
  CASE(128, 1) CASE(128, 4) CASE(128, 5)
#undef CASE
  return -1;
}

int ref_brief_describe(const uint8_t* img, int vstep, int x, int y, int rot, int words,
                       uint32_t* desc) {
#define CASE(V, W)                                                              \
  if (vstep == V && words == W) {                                               \
    pislam::briefDescribe<V, W>((uint8_t(*)[V])img, x, y, rot, desc);           \
    return 0;                                                                   \
  }
  CASE(640, 8) CASE(1920, 8) CASE(3840, 8) CASE(128, 8) CASE(96, 8)
  CASE(128, 1) CASE(128, 4) CASE(128, 5)
#undef CASE
  return -1;
}

int ref_bilinear7_8(uint8_t* img, uint8_t* out, int vstep, int width, int height) {
#define CASE(V) \
  if (vstep == V) { pislam::bilinear7_8<V>(width, height, (uint8_t(*)[V])img, (uint8_t(*)[V])out); return 0; }
  REF_VSTEPS(CASE)
#undef CASE
  return -1;
}

int ref_bilinear13_16(uint8_t* img, uint8_t* out, int vstep, int width, int height) {
#define CASE(V) \
  if (vstep == V) { pislam::bilinear13_16<V>(width, height, (uint8_t(*)[V])img, (uint8_t(*)[V])out); return 0; }
  REF_VSTEPS(CASE)
#undef CASE
  return -1;
}

uint32_t ref_encode_fast(uint32_t score, uint32_t x, uint32_t y) { return pislam::encodeFast(score, x, y); }
uint32_t ref_rencode_fast_score(uint32_t score, uint32_t e) { return pislam::rencodeFastScore(score, e); }
uint32_t ref_decode_fast_x(uint32_t e) { return pislam::decodeFastX(e); }
uint32_t ref_decode_fast_y(uint32_t e) { return pislam::decodeFastY(e); }
uint32_t ref_decode_fast_score(uint32_t e) { return pislam::decodeFastScore(e); }

}  // extern "C"
