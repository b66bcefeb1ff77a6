// Descriptor matcher (sm_100a), SURVEY.md section 8f row N2.
//
// The reference never released a matcher: README.md:125-139 only mentions "feature matching ... under 20 ms per
// frame using flann" with table-driven bit counts, and a demo that traces "best matches within a 20 pixel
// radius".  The definition below is therefore OURS (stated in include/pislam_b200.h and restated in numpy in
// tests/oracle_lib.py): exhaustive Hamming distance between rBRIEF descriptors, optionally windowed to train
// keypoints of the same pyramid level within a pixel radius, best and second-best distance per query.
//
//   match_kernel<RECORDS>   one thread per query descriptor (8 words in registers); the train set streams through
//                           shared memory in tiles of 128 entries; POPC of the XOR, branch-free best/second update.
//                           RECORDS = 1: queries / train sets are frames of a batch of keypoint records (device),
//                           grid.y = frame pair; RECORDS = 0: plain arrays, one pair.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_ops.cuh"
#include "front_kernels.cuh"

namespace pb {

constexpr int MATCH_NT = 128;           // threads per CTA = queries per CTA = train entries per shared tile
constexpr int MATCH_NONE = 0x7fffffff;  // distance reported when no train keypoint qualifies

struct MatchParams {
  int words;   // descriptor words, 1..8
  int radius;  // <= 0: every train descriptor is a candidate; > 0: same level and dx^2 + dy^2 <= radius^2
  // RECORDS = 0
  const uint32_t* q_desc;
  const uint32_t* q_kp;  // may be null when radius <= 0
  const uint8_t* q_level;
  int nq;
  const uint32_t* t_desc;
  const uint32_t* t_kp;
  const uint8_t* t_level;
  int nt;
  // RECORDS = 1
  const uint8_t* records;
  RecordLayout rec;
  int cap;
  const int32_t* pairs;  // [npairs][2]: query frame, train frame
  // output: [pair][nq or cap][3] = best train index (-1: none), best distance, second-best distance
  int32_t* out;
};

template <int RECORDS>
__global__ void __launch_bounds__(MATCH_NT) match_kernel(const __grid_constant__ MatchParams P) {
  __shared__ uint32_t s_desc[MATCH_NT][9];  // pitch 9 words: lanes reading one entry broadcast, writers conflict-free
  __shared__ uint32_t s_kp[MATCH_NT];
  __shared__ uint8_t s_level[MATCH_NT];
  const int tid = threadIdx.x;
  const uint32_t *q_desc = P.q_desc, *q_kp = P.q_kp, *t_desc = P.t_desc, *t_kp = P.t_kp;
  const uint8_t *q_level = P.q_level, *t_level = P.t_level;
  int nq = P.nq, nt = P.nt;
  int32_t* out = P.out;
  if (RECORDS) {
    const int fa = P.pairs[2 * blockIdx.y], fb = P.pairs[2 * blockIdx.y + 1];
    const uint8_t* ra = P.records + (size_t)fa * P.rec.bytes;
    const uint8_t* rb = P.records + (size_t)fb * P.rec.bytes;
    nq = (int)min(*reinterpret_cast<const uint32_t*>(ra), (uint32_t)P.cap);
    nt = (int)min(*reinterpret_cast<const uint32_t*>(rb), (uint32_t)P.cap);
    q_kp = reinterpret_cast<const uint32_t*>(ra + P.rec.kp_off);
    t_kp = reinterpret_cast<const uint32_t*>(rb + P.rec.kp_off);
    q_desc = reinterpret_cast<const uint32_t*>(ra + P.rec.desc_off);
    t_desc = reinterpret_cast<const uint32_t*>(rb + P.rec.desc_off);
    q_level = ra + P.rec.level_off;
    t_level = rb + P.rec.level_off;
    out += (size_t)blockIdx.y * P.cap * 3;
  }
  const int q = blockIdx.x * MATCH_NT + tid;
  if (blockIdx.x * MATCH_NT >= nq) return;  // whole CTA beyond the query set (uniform)
  const bool have_q = q < nq;
  const bool windowed = P.radius > 0;
  uint32_t d[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = (have_q && k < P.words) ? q_desc[(size_t)q * P.words + k] : 0u;
  int qx = 0, qy = 0, ql = 0;
  if (have_q && windowed) {
    const uint32_t e = q_kp[q];
    qx = (int)decode_x(e);
    qy = (int)decode_y(e);
    ql = q_level[q];
  }
  const int r2 = P.radius * P.radius;
  int best = MATCH_NONE, second = MATCH_NONE, bidx = -1;
  for (int t0 = 0; t0 < nt; t0 += MATCH_NT) {
    __syncthreads();  // the previous tile has been consumed
    const int j = t0 + tid;
    if (j < nt) {
#pragma unroll
      for (int k = 0; k < 8; ++k) s_desc[tid][k] = k < P.words ? t_desc[(size_t)j * P.words + k] : 0u;
      if (windowed) {
        s_kp[tid] = t_kp[j];
        s_level[tid] = t_level[j];
      }
    }
    __syncthreads();
    const int m = min(MATCH_NT, nt - t0);
    for (int i = 0; i < m; ++i) {
      int dist = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) dist += __popc(d[k] ^ s_desc[i][k]);
      bool ok = true;
      if (windowed) {
        const uint32_t e = s_kp[i];
        const int dx = (int)decode_x(e) - qx, dy = (int)decode_y(e) - qy;
        ok = s_level[i] == ql && dx * dx + dy * dy <= r2;
      }
      // ascending train index + strict "<": the smallest index wins a tie; a tie with the best becomes the second
      if (ok) {
        if (dist < best) {
          second = best;
          best = dist;
          bidx = t0 + i;
        } else if (dist < second) {
          second = dist;
        }
      }
    }
  }
  if (have_q) {
    out[3 * q] = bidx;
    out[3 * q + 1] = best;
    out[3 * q + 2] = second;
  }
}

}  // namespace pb
