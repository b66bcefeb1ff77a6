/* pislam::briefDescribe -- drop-in for reference include/Brief.h:637-639 (signature) / :28-53, :57-733
 * (behaviour: 256 rotated test pairs, 30 angle bins, coordinates clamped to +-15). */
#ifndef PISLAM_B200_BRIEF_H_
#define PISLAM_B200_BRIEF_H_
#include <cstdint>

#include "Context.h"

namespace pislam {

/* `rot` outside 0..29 leaves `descriptor` untouched, like the reference's switch (Brief.h:641-732). */
template <int vstep, int words>
void briefDescribe(uint8_t img[][vstep], int x, int y, int rot, uint32_t descriptor[words]) {
  static_assert(words >= 1 && words <= 8, "1..8 descriptor words");
  b200::check(pislam_b200_brief_describe(b200::context(), &img[0][0], vstep, y + 16, x, y, rot, words, descriptor),
              "briefDescribe");
}

}  // namespace pislam
#endif
