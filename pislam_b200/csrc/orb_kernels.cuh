// Descriptor kernels (sm_100a): intensity-centroid orientation + rotated BRIEF, one warp per keypoint.
//
// Lane r owns patch row dy = r-15: it loads the row as aligned 32-bit words, funnel-shifts them so that
// byte 0 is column x-15, and reduces the moments with IDP.4A against a per-|dy| weight table (the disc
// of Orb.h:118-126).  The aligned rows are then parked in shared memory (row pitch 36 B, conflict-free)
// and each lane evaluates 8 of the 256 pattern pairs of the keypoint's angle bin -> one descriptor byte.
// Rules: SURVEY.md section 8a rows O1/O2/O3/B1; nothing here mirrors the reference's NEON code.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_ops.cuh"
#include "front_kernels.cuh"

namespace pb {

constexpr int ORB_WARPS = 8;
constexpr int PATCH_PITCH_W = 9;  // words per patch row in shared memory (36 bytes)

struct OrbParams {
  const uint8_t* img;  // frame 0
  unsigned long long frame_stride;
  unsigned long long frame_bytes;  // readable bytes per frame (bounds the word loads)
  int vstep;
  int words;  // descriptor words per keypoint, 1..8
  // keypoint source: records (fused path) or a plain array (literal orbCompute)
  uint8_t* records;
  RecordLayout rec;
  int row_off[MAX_LEVELS];
  int nframes;         // fused: frames in this launch; literal: 1
  int frame0;          // fused TMA path: index of the launch's first frame in the tensor map (records/img point at it)
  int cap;             // fused: record capacity (slots per frame)
  const uint32_t* kp;  // literal: n keypoints with y relative to img
  int n;
  uint32_t* desc;      // literal: [n][words]
  uint8_t* angles;     // literal, optional
  const uint32_t* brief;    // [30][256] (off1 << 16) | off0, byte offsets into the 31 x 36-byte patch
  const uint32_t* brief48;  // same for the 48-byte row pitch of the TMA patch buffers
};

// 4 unaligned-safe bytes at p as a little-endian word; bytes at or past `end` read as zero
__device__ __forceinline__ uint32_t load_word_bounded(const uint8_t* p, const uint8_t* end) {
  if (p + 4 <= end) return *reinterpret_cast<const uint32_t*>(p);
  uint32_t v = 0;
  for (int k = 0; k < 4; ++k)
    if (p + k < end) v |= (uint32_t)p[k] << (8 * k);
  return v;
}

// Shared per-CTA tables: s_wx[|dy|][k] = signed byte weights dx (0 outside the disc) of columns 4k..4k+3,
// s_w1 = 0/1 membership, row pitch 9 words so that the 16 |dy| rows fall into distinct banks.
__device__ __forceinline__ void build_disc_tables(uint32_t* s_wx, uint32_t* s_w1) {
  for (int i = threadIdx.x; i < 16 * 8; i += blockDim.x) {
    const int ady = i >> 3, k = i & 7;
    const int u = disc_halfwidth(ady);
    uint32_t wx = 0, w1 = 0;
    for (int b = 0; b < 4; ++b) {
      const int dx = 4 * k + b - 15;
      if (dx >= -u && dx <= u) {
        wx |= (uint32_t)(uint8_t)(int8_t)dx << (8 * b);
        w1 |= 1u << (8 * b);
      }
    }
    s_wx[ady * 9 + k] = wx;
    s_w1[ady * 9 + k] = w1;
  }
}

// Moments of the warp's patch from the aligned row words w[0..7] (byte j of the row = column x-15+j): lane r owns
// row r-15, lane 31 contributes nothing.
__device__ __forceinline__ void moments_from_rows(const uint32_t w[8], const uint32_t* s_wx, const uint32_t* s_w1,
                                                  int32_t& m10, int32_t& m01) {
  const int lane = threadIdx.x & 31;
  m10 = 0;
  m01 = 0;
  if (lane < 31) {
    const int dy = lane - 15;
    const int ady = dy < 0 ? -dy : dy;
    uint32_t rowsum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      m10 = dp4a_us(w[k], s_wx[ady * 9 + k], m10);
      rowsum = dp4a_uu(w[k], s_w1[ady * 9 + k], rowsum);
    }
    m01 = dy * (int32_t)rowsum;
  }
  m10 = __reduce_add_sync(0xffffffffu, m10);
  m01 = __reduce_add_sync(0xffffffffu, m01);
}

// q[0..11] hold the 48 bytes of this lane's patch row starting at its 16-byte chunk; column x-15 is byte `ob`
// (any value per lane).  Produces w[0..7] (byte j = column x-15+j) and this warp's moments.
__device__ __forceinline__ void moments_from_chunks(uint32_t q[13], uint32_t ob, const uint32_t* s_wx,
                                                    const uint32_t* s_w1, uint32_t w[8], int32_t& m10, int32_t& m01) {
  q[12] = 0u;
  // drop (ob >> 2) leading words, then shift by the remaining (ob & 3) bytes
  if (ob & 8) {
#pragma unroll
    for (int k = 0; k < 11; ++k) q[k] = q[k + 2];
  }
  if (ob & 4) {
#pragma unroll
    for (int k = 0; k < 9; ++k) q[k] = q[k + 1];
  }
  const uint32_t sh = (ob & 3) * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) w[k] = __funnelshift_r(q[k], q[k + 1], sh);
  moments_from_rows(w, s_wx, s_w1, m10, m01);
}

// Loads patch row (lane-15) of the keypoint at (x, y) from global memory (linear addressing, bounded by `end`)
// into w[0..7] and returns the warp's moments.
__device__ __forceinline__ void row_moments(const uint8_t* img, const uint8_t* end, int vstep, int x, int y,
                                            const uint32_t* s_wx, const uint32_t* s_w1, uint32_t w[8], int32_t& m10,
                                            int32_t& m01) {
  const int lane = threadIdx.x & 31;
  uint32_t q[13];
#pragma unroll
  for (int k = 0; k < 13; ++k) q[k] = 0u;
  uint32_t ob = 0;
  if (lane < 31) {
    const uint8_t* a = img + (size_t)(y + lane - 15) * vstep + (x - 15);
    const uintptr_t ai = reinterpret_cast<uintptr_t>(a);
    ob = (uint32_t)(ai & 15);  // byte offset of column x-15 inside its 16-byte chunk
    const uint8_t* a16 = reinterpret_cast<const uint8_t*>(ai & ~(uintptr_t)15);
    // three 16-byte loads cover bytes ob .. ob+30 (<= 45): one L1 wavefront per row and load instead of
    // one per row and 4-byte word
    if (a16 + 48 <= end) {
      const uint4* a4 = reinterpret_cast<const uint4*>(a16);
      const uint4 c0 = __ldg(a4), c1 = __ldg(a4 + 1), c2 = __ldg(a4 + 2);
      q[0] = c0.x; q[1] = c0.y; q[2] = c0.z; q[3] = c0.w;
      q[4] = c1.x; q[5] = c1.y; q[6] = c1.z; q[7] = c1.w;
      q[8] = c2.x; q[9] = c2.y; q[10] = c2.z; q[11] = c2.w;
    } else {  // the very end of a buffer
#pragma unroll
      for (int k = 0; k < 12; ++k) q[k] = load_word_bounded(a16 + 4 * k, end);
    }
  }
  moments_from_chunks(q, ob, s_wx, s_w1, w, m10, m01);
}

// One descriptor byte per lane: pairs 8*lane .. 8*lane+7 of angle bin `rot` (Brief.h:28-53, 71-631)
__device__ __forceinline__ uint32_t brief_byte(const uint8_t* patch, const uint32_t* __restrict__ brief, int rot) {
  // `patch` points at the byte of (dx,dy) = (-15,-15); the table holds offsets for the patch's row pitch
  const int lane = threadIdx.x & 31;
  const uint4* t = reinterpret_cast<const uint4*>(brief + rot * 256 + lane * 8);
  const uint4 ta = __ldg(t), tb = __ldg(t + 1);
  const uint32_t e[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) bits |= (uint32_t)(patch[e[j] & 0xffffu] < patch[e[j] >> 16]) << j;
  return bits;
}

// FUSED = 1: keypoints come from the frame records (level-local y + row_off[level]);
// FUSED = 0: the literal orbCompute on a keypoint array (one frame).
// Persistent: every warp strides over the (frame, slot) space, so the disc tables are built once per CTA and
// no CTA is launched for the unused tail of a record.
template <int FUSED>
__global__ void __launch_bounds__(ORB_WARPS * 32) orb_kernel(const __grid_constant__ OrbParams P) {
  __shared__ uint32_t s_wx[16 * 9], s_w1[16 * 9];
  __shared__ uint32_t s_patch[ORB_WARPS][32 * PATCH_PITCH_W];
  build_disc_tables(s_wx, s_w1);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * ORB_WARPS;
  const int cap = FUSED ? P.cap : P.n;
  const long long nslots = (long long)P.nframes * cap;
  uint32_t* patch = s_patch[warp];
  for (long long slot = (long long)blockIdx.x * ORB_WARPS + warp; slot < nslots; slot += nwarps) {
    const int frame = (int)(slot / cap);
    const int i = (int)(slot - (long long)frame * cap);
    const uint8_t* img = P.img + (size_t)frame * P.frame_stride;
    const uint8_t* end = img + P.frame_bytes;
    int x, y;
    uint32_t* desc;
    if (FUSED) {
      uint8_t* rec = P.records + (size_t)frame * P.rec.bytes;
      const int n = (int)__ldg(reinterpret_cast<const uint32_t*>(rec));
      if (i >= n) continue;
      const uint32_t e = __ldg(reinterpret_cast<const uint32_t*>(rec + P.rec.kp_off) + i);
      x = (int)decode_x(e);
      y = (int)decode_y(e) + P.row_off[rec[P.rec.level_off + i]];
      desc = reinterpret_cast<uint32_t*>(rec + P.rec.desc_off) + (size_t)i * P.words;
    } else {
      const uint32_t e = P.kp[i];
      x = (int)decode_x(e);
      y = (int)decode_y(e);
      desc = P.desc ? P.desc + (size_t)i * P.words : nullptr;
    }
    uint32_t w[8];
    int32_t m10, m01;
    row_moments(img, end, P.vstep, x, y, s_wx, s_w1, w, m10, m01);
    const int rot = (int)angle_bin(m10, m01);
    if (!FUSED && P.angles && lane == 0) P.angles[i] = (uint8_t)rot;
    if (desc == nullptr) continue;
    __syncwarp();  // the previous keypoint's pattern reads are done
#pragma unroll
    for (int k = 0; k < 8; ++k) patch[lane * PATCH_PITCH_W + k] = w[k];
    __syncwarp();
    const uint32_t b = brief_byte(reinterpret_cast<const uint8_t*>(patch), P.brief, rot);
    if (lane < 4 * P.words) reinterpret_cast<uint8_t*>(desc)[lane] = (uint8_t)b;
  }
}

// --------------------------------------------------------------------------------------------------
// Fused-path descriptor kernel with TMA patch prefetch.  Every warp owns two 31 x 48-byte patch buffers
// and two mbarriers: lane 0 issues the box load of the NEXT keypoint (box origin = column (x-15) & ~15, a
// TMA box must start on a 16-byte column) before the warp waits for the CURRENT one, so the L2 latency of
// a patch hides behind the previous keypoint's arithmetic and no load instruction touches the pixels.
// The pattern is evaluated straight out of the TMA buffer (table offsets for a 48-byte pitch + ob).
// --------------------------------------------------------------------------------------------------
constexpr int OT_W = 48, OT_H = 31, OT_BYTES = OT_W * OT_H, OT_STRIDE = 1536;
constexpr int ORB_FRAMES_PER_THREAD = 4;
constexpr int ORB_MAX_FRAMES = ORB_WARPS * 32 * ORB_FRAMES_PER_THREAD;  // frames per launch (the host splits larger batches)

struct OrbKp {
  int frame, i, x, y;
};

__global__ void __launch_bounds__(ORB_WARPS * 32) orb_tma_kernel(const __grid_constant__ OrbParams P,
                                                                 const __grid_constant__ CUtensorMap tmap) {
  __shared__ __align__(128) uint8_t s_buf[ORB_WARPS][2][OT_STRIDE];
  __shared__ __align__(8) uint64_t s_bar[ORB_WARPS][2];
  __shared__ uint32_t s_wx[16 * 9], s_w1[16 * 9];
  __shared__ uint16_t s_recpe[256];
  __shared__ uint32_t s_start[ORB_MAX_FRAMES + 1], s_wsum[ORB_WARPS];
  build_disc_tables(s_wx, s_w1);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_recpe[i] = (uint16_t)recpe_mantissa((uint32_t)i);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    mbar_init(&s_bar[warp][0], 1);
    mbar_init(&s_bar[warp][1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // Keypoints are numbered across the launch's frames (exclusive scan of the record counts, once per CTA); CTA c
  // owns the contiguous range [K c / G, K (c+1) / G) and its warps interleave inside it.  No empty record slot is
  // ever visited, a warp stays in one frame (and its neighbourhood of the image) for a long time, and every CTA
  // gets the same number of keypoints.
  {
    const int f0 = (int)threadIdx.x * ORB_FRAMES_PER_THREAD;
    uint32_t c[ORB_FRAMES_PER_THREAD], sum = 0;
#pragma unroll
    for (int j = 0; j < ORB_FRAMES_PER_THREAD; ++j) {
      c[j] = f0 + j < P.nframes
                 ? min(__ldg(reinterpret_cast<const uint32_t*>(P.records + (size_t)(f0 + j) * P.rec.bytes)), (uint32_t)P.cap)
                 : 0u;
      sum += c[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t run = incl - sum;
    for (int w = 0; w < warp; ++w) run += s_wsum[w];
    if (threadIdx.x == 0) s_start[0] = 0u;
#pragma unroll
    for (int j = 0; j < ORB_FRAMES_PER_THREAD; ++j) {
      run += c[j];
      if (f0 + j < P.nframes) s_start[f0 + j + 1] = run;
    }
  }
  __syncthreads();
  const uint32_t K = s_start[P.nframes];
  const uint32_t k_end = (uint32_t)((unsigned long long)K * (blockIdx.x + 1) / gridDim.x);
  uint32_t kk = (uint32_t)((unsigned long long)K * blockIdx.x / gridDim.x) + (uint32_t)warp;
  int frame = 0;
  if (kk < k_end) {  // frame of the warp's first keypoint: the last f with s_start[f] <= kk
    int lo = 0, hi = P.nframes;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_start[mid] <= kk) lo = mid;
      else hi = mid;
    }
    frame = lo;
  }

  // this warp's next keypoint (all lanes compute the same values)
  auto advance = [&](OrbKp& k) -> bool {
    if (kk >= k_end) return false;
    while (kk >= s_start[frame + 1]) ++frame;
    const int idx = (int)(kk - s_start[frame]);
    kk += ORB_WARPS;
    const uint8_t* rec = P.records + (size_t)frame * P.rec.bytes;
    const uint32_t e = __ldg(reinterpret_cast<const uint32_t*>(rec + P.rec.kp_off) + idx);
    k.frame = frame;
    k.i = idx;
    k.x = (int)decode_x(e);
    k.y = (int)decode_y(e) + P.row_off[__ldg(rec + P.rec.level_off + idx)];
    return true;
  };
  auto issue = [&](const OrbKp& k, int b) {
    if (lane == 0) {
      mbar_expect_tx(&s_bar[warp][b], OT_BYTES);
      tma_load_3d(s_buf[warp][b], &tmap, (k.x - 15) & ~15, k.y - 15, P.frame0 + k.frame, &s_bar[warp][b]);
    }
  };

  OrbKp cur, nxt;
  bool have = advance(cur);
  if (have) issue(cur, 0);
  int b = 0;
  uint32_t phase = 0;  // bit b = parity to wait for on buffer b
  while (have) {
    const bool have_next = advance(nxt);
    if (have_next) issue(nxt, b ^ 1);
    mbar_wait(&s_bar[warp][b], (phase >> b) & 1u);
    phase ^= 1u << b;

    const uint8_t* buf = s_buf[warp][b];
    const uint32_t ob = (uint32_t)((cur.x - 15) & 15);
    uint32_t q[13];
    {
      const uint4* r4 = reinterpret_cast<const uint4*>(buf + (lane < 31 ? lane : 0) * OT_W);
      const uint4 c0 = r4[0], c1 = r4[1], c2 = r4[2];
      q[0] = c0.x; q[1] = c0.y; q[2] = c0.z; q[3] = c0.w;
      q[4] = c1.x; q[5] = c1.y; q[6] = c1.z; q[7] = c1.w;
      q[8] = c2.x; q[9] = c2.y; q[10] = c2.z; q[11] = c2.w;
    }
    uint32_t w[8];
    int32_t m10, m01;
    moments_from_chunks(q, ob, s_wx, s_w1, w, m10, m01);
    const int rot = (int)angle_bin(m10, m01, s_recpe);
    const uint32_t bits = brief_byte(buf + ob, P.brief48, rot);
    uint8_t* rec = P.records + (size_t)cur.frame * P.rec.bytes;
    if (lane < 4 * P.words) (rec + P.rec.desc_off)[(uint32_t)(cur.i * P.words * 4 + lane)] = (uint8_t)bits;
    __syncwarp();  // every lane is done with buf before lane 0 lets the next TMA overwrite it
    cur = nxt;
    have = have_next;
    b ^= 1;
  }
}

// orbCentroids (Orb.h:80-308): moments only, written in the reference's [x0 x1 x2 x3 y0 y1 y2 y3] layout
__global__ void __launch_bounds__(ORB_WARPS * 32)
    orb_centroids_kernel(const uint8_t* img, const uint8_t* end, int vstep, const uint32_t* __restrict__ kp, int n,
                         int32_t* __restrict__ centroids) {
  __shared__ uint32_t s_wx[16 * 9], s_w1[16 * 9];
  build_disc_tables(s_wx, s_w1);
  __syncthreads();
  const int i = blockIdx.x * ORB_WARPS + (threadIdx.x >> 5);
  if (i >= n) return;
  uint32_t w[8];
  int32_t m10, m01;
  row_moments(img, end, vstep, (int)decode_x(kp[i]), (int)decode_y(kp[i]), s_wx, s_w1, w, m10, m01);
  if ((threadIdx.x & 31) == 0) {
    const int base = (i >> 2) * 8 + (i & 3);
    centroids[base] = m10;
    centroids[base + 4] = m01;
  }
}

// pislam::atan2 (Orb.h:310-387) on the interleaved centroid layout: group g holds x0..x3 y0..y3
__global__ void atan2_kernel(const int32_t* __restrict__ xys, int ngroups, uint8_t* __restrict__ angles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ngroups * 4) return;
  const int g = i >> 2, k = i & 3;
  angles[i] = (uint8_t)angle_bin(xys[g * 8 + k], xys[g * 8 + 4 + k]);
}

// briefDescribe (Brief.h:637-733) for one point and a caller-given rot; one warp.
__global__ void brief_describe_kernel(const uint8_t* img, const uint8_t* end, int vstep, int x, int y, int rot,
                                      int words, const uint32_t* __restrict__ brief, uint32_t* __restrict__ desc) {
  __shared__ uint32_t s_patch[32 * PATCH_PITCH_W];
  const int lane = threadIdx.x & 31;
  if (rot < 0 || rot >= 30) return;  // Brief.h:641-732: no case -> descriptor untouched
  uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < 31) {
    const uint8_t* a = img + (size_t)(y + lane - 15) * vstep + (x - 15);
    for (int k = 0; k < 8; ++k) {
      uint32_t v = 0;
      for (int b = 0; b < 4; ++b) {
        const uint8_t* p = a + 4 * k + b;
        if (4 * k + b < 31 && p < end) v |= (uint32_t)*p << (8 * b);
      }
      w[k] = v;
    }
  }
  for (int k = 0; k < 8; ++k) s_patch[lane * PATCH_PITCH_W + k] = w[k];
  __syncwarp();
  const uint32_t b = brief_byte(reinterpret_cast<const uint8_t*>(s_patch), brief, rot);
  if (lane < 4 * words) reinterpret_cast<uint8_t*>(desc)[lane] = (uint8_t)b;
}

}  // namespace pb
