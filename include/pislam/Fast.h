/* pislam::fastDetect / fastScoreHarris / fastExtract -- drop-in for reference include/Fast.h
 * (signatures :54-56, :166-168, :196-198), each stage executed by hand-written sm_100a CUDA behind
 * the C ABI (include/pislam_b200.h).  Host buffers end up exactly as the reference leaves them:
 * only the bytes the reference writes are modified, keypoints are appended in the reference's order.
 *
 * For throughput use the fused whole-pyramid entry point (pislam_b200_extract_batch) instead of these
 * per-stage calls: every call here is a host->device->host round trip of the level.
 */
#ifndef PISLAM_B200_FAST_H_
#define PISLAM_B200_FAST_H_
#include <cstdint>
#include <vector>

#include "Context.h"
#include "Harris.h"
#include "Util.h"

namespace pislam {

template <int vstep, int border>
void fastDetect(const int width, const int height, uint8_t img[][vstep], uint8_t out[][vstep], int threshold) {
  b200::check(pislam_b200_fast_detect(b200::context(), &img[0][0], &out[0][0], vstep, border, width, height, threshold),
              "fastDetect");
}

template <int vstep, int border>
void fastScoreHarris(int width, int height, uint8_t img[][vstep], int32_t threshold, uint8_t out[][vstep]) {
  b200::check(
      pislam_b200_fast_score_harris(b200::context(), &img[0][0], &out[0][0], vstep, border, width, height, threshold),
      "fastScoreHarris");
}

/* Appends to `results` and, like the reference (Fast.h:197-198,354), returns a copy of the whole vector. */
template <int vstep, int border, int logBucketSize = 0, int bucketLimit = 5>
std::vector<uint32_t> fastExtract(const int width, const int height, uint8_t out[][vstep],
                                  std::vector<uint32_t>& results) {
  const int iw = width - 2 * border, ih = height - 2 * border;
  const int cap = (iw > 0 && ih > 0) ? ((iw + 1) / 2) * ((ih + 1) / 2) : 0; /* one keypoint per 2x2 cell at most */
  std::vector<uint32_t> found(cap > 0 ? cap : 1);
  int n = 0;
  b200::check(pislam_b200_fast_extract(b200::context(), &out[0][0], vstep, border, width, height, logBucketSize,
                                       bucketLimit, found.data(), cap, &n),
              "fastExtract");
  results.insert(results.end(), found.begin(), found.begin() + n);
  return results;
}

}  // namespace pislam
#endif
