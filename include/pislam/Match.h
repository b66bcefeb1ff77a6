/* pislam::matchHamming -- descriptor matching on top of orbCompute's output.
 * NOT a drop-in for reference code: the reference never released its matcher (README.md:125-139: "feature
 * matching ... using flann", a demo tracing "best matches within a 20 pixel radius").  The definition is stated in
 * include/pislam_b200.h (pislam_b200_match_hamming): popcount-of-XOR distance, optional window on same-level
 * keypoints within `radius` level-local pixels, smallest index wins ties. */
#ifndef PISLAM_B200_MATCH_H_
#define PISLAM_B200_MATCH_H_
#include <cstdint>
#include <vector>

#include "Context.h"

namespace pislam {

struct Match {
  int32_t index;     /* index into the train set, -1 when no candidate qualifies */
  int32_t distance;  /* Hamming distance of the best candidate (PISLAM_MATCH_NONE when index == -1) */
  int32_t second;    /* distance of the runner-up (a tie with the best counts), PISLAM_MATCH_NONE if there is none */
};

/* queryDescriptors / trainDescriptors: `words` uint32 per keypoint, as orbCompute<vstep,words> appends them.
 * radius <= 0: exhaustive; radius > 0: keypoints (encodeFast values with level-local y) and levels are required. */
template <int words>
std::vector<Match> matchHamming(const std::vector<uint32_t>& queryDescriptors, const std::vector<uint32_t>& trainDescriptors,
                                int radius = 0, const std::vector<uint32_t>* queryPoints = nullptr,
                                const std::vector<uint8_t>* queryLevels = nullptr,
                                const std::vector<uint32_t>* trainPoints = nullptr,
                                const std::vector<uint8_t>* trainLevels = nullptr) {
  static_assert(words >= 1 && words <= 8, "1..8 descriptor words");
  static_assert(sizeof(Match) == 3 * sizeof(int32_t), "Match is three packed int32");
  const int nq = (int)(queryDescriptors.size() / words), nt = (int)(trainDescriptors.size() / words);
  std::vector<Match> out((size_t)nq);
  if (nq == 0) return out;
  b200::check(pislam_b200_match_hamming(b200::context(), queryDescriptors.data(),
                                        queryPoints ? queryPoints->data() : nullptr,
                                        queryLevels ? queryLevels->data() : nullptr, nq, trainDescriptors.data(),
                                        trainPoints ? trainPoints->data() : nullptr,
                                        trainLevels ? trainLevels->data() : nullptr, nt, words, radius,
                                        reinterpret_cast<int32_t*>(out.data())),
              "matchHamming");
  return out;
}

}  // namespace pislam
#endif /* PISLAM_B200_MATCH_H_ */
