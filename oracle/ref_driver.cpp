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
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
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

// The reference's caller loop (demo/demo.cpp:77-101, README.md:56-83) over a batch of stacked pyramids, one whole
// pyramid per task on `nthreads` std::threads -- the reference itself is single-threaded (nthreads = 1 is its
// own shape; more threads = "all host cores" of bench.py's CPU arm).  No Python between the stages.
// `out` is zeroed per frame (Fast.h:42-44 requires it); orbCompute runs per level on level-local keypoints, which
// yields the same descriptors as the demo's single call on stack coordinates and stays valid for stacks taller than
// the 12-bit y range (1080p, 4K).  counts[f] = keypoints of frame f; kp / desc (optional) are [nframes][cap] and
// [nframes][cap*8].  Returns 0, -1 for a (vstep, bucket) combination outside the instantiated menu, -2 if cap is
// too small.
}  // extern "C"

template <int V, int L, int K>
static int ref_frame(const uint8_t* stack, uint8_t* out, long rows, int nlevels, const int* widths,
                     const int* heights, const int* row_offsets, int fast_t, int32_t harris_t,
                     std::vector<uint32_t>& points, std::vector<uint32_t>& desc) {
  std::memset(out, 0, (size_t)rows * V);
  points.clear();
  desc.clear();
  for (int l = 0; l < nlevels; ++l) {
    uint8_t(*img)[V] = (uint8_t(*)[V])(stack + (size_t)row_offsets[l] * V);
    uint8_t(*o)[V] = (uint8_t(*)[V])(out + (size_t)row_offsets[l] * V);
    std::vector<uint32_t> lvl;
    pislam::fastDetect<V, 16>(widths[l], heights[l], img, o, fast_t);
    pislam::fastScoreHarris<V, 16>(widths[l], heights[l], img, harris_t, o);
    pislam::fastExtract<V, 16, L, K>(widths[l], heights[l], o, lvl);
    pislam::orbCompute<V, 8>(img, lvl, desc);
    points.insert(points.end(), lvl.begin(), lvl.end());
  }
  return (int)points.size();
}

extern "C" {

int ref_extract_batch(const uint8_t* stacks, long frame_stride, int nframes, int vstep, int nlevels,
                      const int* widths, const int* heights, const int* row_offsets, int fast_t,
                      int32_t harris_t, int logBucketSize, int bucketLimit, int cap, int nthreads,
                      int* counts, uint32_t* kp, uint32_t* desc) {
  typedef int (*FrameFn)(const uint8_t*, uint8_t*, long, int, const int*, const int*, const int*, int, int32_t,
                         std::vector<uint32_t>&, std::vector<uint32_t>&);
  FrameFn fn = nullptr;
#define CASE(V, B, L, K) \
  if (vstep == V && logBucketSize == L && bucketLimit == K) fn = ref_frame<V, L, K>;
#define PER_V(V) REF_BUCKETS(CASE, V, 16)
  PER_V(640) PER_V(1920) PER_V(3840)
#undef PER_V
#undef CASE
  if (!fn) return -1;
  long rows = 0;
  for (int l = 0; l < nlevels; ++l)
    if (row_offsets[l] + heights[l] > rows) rows = row_offsets[l] + heights[l];
  std::atomic<int> next(0), status(0);
  auto work = [&]() {
    std::vector<uint8_t> out((size_t)rows * vstep + 64);
    std::vector<uint32_t> points, d;
    for (;;) {
      const int f = next.fetch_add(1);
      if (f >= nframes) break;
      const int n = fn(stacks + (size_t)f * frame_stride, out.data(), rows, nlevels, widths, heights, row_offsets,
                       fast_t, harris_t, points, d);
      counts[f] = n;
      if (n > cap) {
        status = -2;
        continue;
      }
      if (kp) std::memcpy(kp + (size_t)f * cap, points.data(), (size_t)n * 4);
      if (desc) std::memcpy(desc + (size_t)f * cap * 8, d.data(), (size_t)n * 32);
    }
  };
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return status;
}

uint32_t ref_encode_fast(uint32_t score, uint32_t x, uint32_t y) { return pislam::encodeFast(score, x, y); }
uint32_t ref_rencode_fast_score(uint32_t score, uint32_t e) { return pislam::rencodeFastScore(score, e); }
uint32_t ref_decode_fast_x(uint32_t e) { return pislam::decodeFastX(e); }
uint32_t ref_decode_fast_y(uint32_t e) { return pislam::decodeFastY(e); }
uint32_t ref_decode_fast_score(uint32_t e) { return pislam::decodeFastScore(e); }

}  // extern "C"
