// The reference's caller loop (reference demo/demo.cpp:77-101, README.md:56-83) written against the drop-in
// headers in include/pislam/: per level fastDetect -> fastScoreHarris -> fastExtract, then one orbCompute over
// the stacked pyramid -- and the same frame through the fused C-ABI call for comparison.
//
//   demo <stack.raw>      640 x 2210 uint8, the VGA 8-level pyramid stacked at pitch 640
//
// Prints keypoint counts and 64-bit FNV-1a hashes of the keypoint and descriptor arrays (the hashes
// SURVEY.md section 8c pins for the reference's fixture); tests/test_cpp_dropin.py checks them.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pislam/Fast.h"
#include "pislam/Harris.h"
#include "pislam/Orb.h"
#include "pislam/Util.h"

#define IMG_W 640
#define IMG_H 480
static const int kLevels = 8;
static const int kWidths[kLevels] = {640, 533, 444, 370, 309, 257, 214, 179};
static const int kHeights[kLevels] = {480, 400, 333, 278, 231, 193, 161, 134};

static uint64_t fnv1a(const void* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <stack.raw (640x2210 uint8)>\n", argv[0]);
    return 2;
  }
  int rows = 0;
  for (int l = 0; l < kLevels; ++l) rows += kHeights[l];
  std::vector<uint8_t> stack((size_t)rows * IMG_W), scores((size_t)rows * IMG_W, 0);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(stack.data(), 1, stack.size(), f) != stack.size()) {
    fprintf(stderr, "cannot read %zu bytes from %s\n", stack.size(), argv[1]);
    return 2;
  }
  fclose(f);
  uint8_t(*img)[IMG_W] = reinterpret_cast<uint8_t(*)[IMG_W]>(stack.data());
  uint8_t(*out)[IMG_W] = reinterpret_cast<uint8_t(*)[IMG_W]>(scores.data());

  // the loop runs a few times: the first pass pays for the CUDA context and the library's first allocations, the
  // last one is what a caller that processes a stream of frames sees per frame
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  std::vector<uint32_t> points, descriptors;
  double ms = 0.0, ms_first = 0.0;
  for (int rep = 0; rep < (reps < 1 ? 1 : reps); ++rep) {
    std::fill(scores.begin(), scores.end(), 0);
    points.clear();
    descriptors.clear();
    const auto t0 = std::chrono::steady_clock::now();
    int row = 0;
    for (int l = 0; l < kLevels; ++l) {
      const int w = kWidths[l], h = kHeights[l];
      const size_t first = points.size();
      pislam::fastDetect<IMG_W, 16>(w, h, img + row, out + row, 20);
      pislam::fastScoreHarris<IMG_W, 16>(w, h, img + row, 1 << 15, out + row);
      pislam::fastExtract<IMG_W, 16>(w, h, out + row, points);
      for (size_t i = first; i < points.size(); ++i) {  // level-local y -> stack y, as the reference demo does
        const uint32_t p = points[i];
        points[i] = pislam::encodeFast(pislam::decodeFastScore(p), pislam::decodeFastX(p), pislam::decodeFastY(p) + row);
      }
      row += h;
    }
    pislam::orbCompute<IMG_W, 8>(img, points, descriptors);
    ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (rep == 0) ms_first = ms;
  }
  printf("stages: %zu features, keypoints %016llx, descriptors %016llx, %.2f ms (first pass %.2f ms)\n", points.size(),
         (unsigned long long)fnv1a(points.data(), points.size() * 4),
         (unsigned long long)fnv1a(descriptors.data(), descriptors.size() * 4), ms, ms_first);

  // pislam::harrisScoreSobel point by point: a keypoint's score byte IS its Harris code (255 = unscored overrun flag)
  {
    int checked = 0, same = 0;
    for (size_t i = 0; i < points.size() && checked < 16; i += points.size() / 16 + 1) {
      const uint32_t p = points[i];
      if (pislam::decodeFastScore(p) == 255) continue;
      ++checked;
      same += pislam::harrisScoreSobel<IMG_W>(img, (int)pislam::decodeFastX(p), (int)pislam::decodeFastY(p), 1 << 15) ==
              pislam::decodeFastScore(p);
    }
    printf("harrisScoreSobel: %d of %d sampled keypoint scores reproduced\n", same, checked);
  }

  // the same frame through the fused whole-pyramid call
  pislam_pyramid_desc d;
  memset(&d, 0, sizeof(d));
  d.vstep = IMG_W;
  d.nlevels = kLevels;
  int row = 0;
  for (int l = 0; l < kLevels; ++l) {
    d.widths[l] = kWidths[l];
    d.heights[l] = kHeights[l];
    d.row_offsets[l] = row;
    row += kHeights[l];
  }
  d.border = 16;
  d.fast_threshold = 20;
  d.harris_threshold = 1 << 15;
  d.log_bucket_size = 0;
  d.bucket_limit = 5;
  d.words = 8;
  const int cap = 8192;
  std::vector<uint32_t> kp(cap), desc((size_t)cap * 8);
  std::vector<uint8_t> level(cap);
  int n = 0;
  const auto t1 = std::chrono::steady_clock::now();
  pislam::b200::check(
      pislam_b200_extract_pyramid(pislam::b200::context(), stack.data(), &d, kp.data(), level.data(), desc.data(), cap, &n),
      "extract_pyramid");
  const double ms2 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
  for (int i = 0; i < n; ++i)
    kp[i] = pislam::encodeFast(pislam::decodeFastScore(kp[i]), pislam::decodeFastX(kp[i]),
                               pislam::decodeFastY(kp[i]) + d.row_offsets[level[i]]);
  printf("fused:  %d features, keypoints %016llx, descriptors %016llx, %.2f ms\n", n,
         (unsigned long long)fnv1a(kp.data(), (size_t)n * 4), (unsigned long long)fnv1a(desc.data(), (size_t)n * 32), ms2);
  return 0;
}
