/* pislam::orbCentroids / atan2 / orbCompute -- drop-in for reference include/Orb.h
 * (signatures :80-82, :310, :396-398).  The image has no height in these signatures; the rows that are
 * uploaded are the ones the keypoints can touch (max y + 16). */
#ifndef PISLAM_B200_ORB_H_
#define PISLAM_B200_ORB_H_
#include <algorithm>
#include <cstdint>
#include <vector>

#include "Brief.h"
#include "Context.h"
#include "Util.h"

namespace pislam {

namespace b200 {
inline int rowsTouched(const std::vector<uint32_t>& points) {
  uint32_t ymax = 0;
  for (uint32_t p : points) ymax = std::max(ymax, decodeFastY(p));
  return (int)ymax + 16;
}
}  // namespace b200

/* m10 / m01 of every keypoint in groups of [x0 x1 x2 x3 y0 y1 y2 y3], size (2n+7)&~7 (Orb.h:113-114,298-304) */
template <int vstep>
std::vector<int32_t> orbCentroids(uint8_t img[][vstep], const std::vector<uint32_t>& points) {
  std::vector<int32_t> centroids((2 * points.size() + 7) & ~(size_t)7);
  if (points.empty()) return centroids;
  b200::check(pislam_b200_orb_centroids(b200::context(), &img[0][0], vstep, b200::rowsTouched(points), points.data(),
                                        (int)points.size(), centroids.data()),
              "orbCentroids");
  return centroids;
}

/* angle bin 0..29 per keypoint from the interleaved centroid layout (Orb.h:310-387); like the reference this
 * name shadows ::atan2 inside namespace pislam */
inline std::vector<uint8_t> atan2(const std::vector<int32_t>& xys) {
  std::vector<uint8_t> angles(xys.size() / 2);
  if (xys.empty()) return angles;
  b200::check(pislam_b200_atan2(b200::context(), xys.data(), (int)xys.size(), angles.data()), "atan2");
  return angles;
}

/* appends points.size()*words uint32 to `descriptors`; descriptor i belongs to points[i] (Orb.h:419-424) */
template <int vstep, int words>
void orbCompute(uint8_t img[][vstep], const std::vector<uint32_t>& points, std::vector<uint32_t>& descriptors) {
  static_assert(words >= 1 && words <= 8, "1..8 descriptor words");
  const size_t old = descriptors.size();
  descriptors.resize(old + points.size() * words);
  if (points.empty()) return;
  b200::check(pislam_b200_orb_compute(b200::context(), &img[0][0], vstep, b200::rowsTouched(points), points.data(),
                                      (int)points.size(), words, descriptors.data() + old, nullptr),
              "orbCompute");
}

}  // namespace pislam
#endif
