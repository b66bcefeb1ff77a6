/* pislam::bilinear7_8 / bilinear13_16 -- drop-in for reference include/Bilinear.h:42-44, :165-167 (signatures):
 * fixed-point bilinear down-scaling by 7/8 (8x8 -> 7x7 blocks) and 13/16 (16x16 -> 13x13 blocks).
 * The image must be padded to a multiple of 8 / 16 rows and columns; output dimensions round down
 * (39x29 -> 34x25 and 31x23).  img and out may be the same array.  Pyramid-preparation helpers, not part
 * of the ORB hot path. */
#ifndef PISLAM_B200_BILINEAR_H_
#define PISLAM_B200_BILINEAR_H_
#include <cstdint>

#include "Context.h"

namespace pislam {

template <int vstep>
void bilinear7_8(const int width, const int height, uint8_t img[][vstep], uint8_t out[][vstep]) {
  b200::check(pislam_b200_bilinear7_8(b200::context(), &img[0][0], &out[0][0], vstep, width, height), "bilinear7_8");
}

template <int vstep>
void bilinear13_16(const int width, const int height, uint8_t img[][vstep], uint8_t out[][vstep]) {
  b200::check(pislam_b200_bilinear13_16(b200::context(), &img[0][0], &out[0][0], vstep, width, height),
              "bilinear13_16");
}

}  // namespace pislam
#endif
