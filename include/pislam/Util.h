/* pislam::encodeFast & friends -- same packing as reference include/Util.h:27-45:
 * score in bits 31..24, x in bits 23..12, y in bits 11..0. Pure host arithmetic. */
#ifndef PISLAM_B200_UTIL_H_
#define PISLAM_B200_UTIL_H_
#include <cstdint>

namespace pislam {

inline uint32_t encodeFast(uint32_t score, uint32_t x, uint32_t y) { return (score << 24) | (x << 12) | y; }
inline uint32_t rencodeFastScore(uint32_t score, uint32_t encoded) { return (score << 24) | (encoded & 0x00ffffffu); }
inline uint32_t decodeFastX(uint32_t encoded) { return (encoded >> 12) & 0xfffu; }
inline uint32_t decodeFastY(uint32_t encoded) { return encoded & 0xfffu; }
inline uint32_t decodeFastScore(uint32_t encoded) { return encoded >> 24; }

}  // namespace pislam
#endif
