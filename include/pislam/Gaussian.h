/* pislam::gaussian5x5 -- drop-in for reference include/Gaussian.h:48-50 (signature): separable
 * [1 4 6 4 1]/16 blur built from rounding halving adds, reflect-101 borders, vertical pass first
 * (the semantics the reference's own test model pins, test/GaussianTest.cpp:159-215).
 * A pyramid-preparation helper, not part of the ORB hot path. */
#ifndef PISLAM_B200_GAUSSIAN_H_
#define PISLAM_B200_GAUSSIAN_H_
#include <cstdint>

#include "Context.h"

namespace pislam {

/* img and out may be the same array (in-place blur), as in the reference */
template <int vstep>
void gaussian5x5(const int width, const int height, uint8_t img[][vstep], uint8_t out[][vstep]) {
  b200::check(pislam_b200_gaussian5x5(b200::context(), &img[0][0], &out[0][0], vstep, width, height), "gaussian5x5");
}

}  // namespace pislam
#endif
