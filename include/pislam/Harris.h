/* pislam::harrisScoreSobel -- drop-in for reference include/Harris.h:80-82 (signature) / :80-248 (behaviour),
 * evaluated on the GPU through pislam_b200_harris_score_sobel. */
#ifndef PISLAM_B200_HARRIS_H_
#define PISLAM_B200_HARRIS_H_
#include <cstdint>

#include "Context.h"

namespace pislam {

template <int vstep>
uint8_t harrisScoreSobel(uint8_t img[][vstep], int x, int y, int32_t threshold) {
  uint8_t score = 0;
  /* the 8x8 window covers rows y-3 .. y+4 and columns x-3 .. x+4: only those eight rows travel to the device (the
   * reference addresses img[y][x] linearly, so a window that crosses the pitch needs the ninth row too) */
  if (y >= 3) {
    int y_local = 3;
    b200::check(pislam_b200_harris_score_sobel(b200::context(), &img[y - 3][0], vstep, x + 4 < vstep ? 8 : 9, &x,
                                               &y_local, 1, threshold, &score),
                "harrisScoreSobel");
  } else {
    b200::check(pislam_b200_harris_score_sobel(b200::context(), &img[0][0], vstep, y + 5, &x, &y, 1, threshold, &score),
                "harrisScoreSobel");
  }
  return score;
}

}  // namespace pislam
#endif
