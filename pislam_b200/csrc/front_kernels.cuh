// Detector front-end kernels (sm_100a): FAST-9 -> Harris code -> 3x3 non-max -> ordered keypoint list.
//
//   front_tile_kernel   one CTA per 128x32 tile of one pyramid level of one frame; the whole batch
//                       (all levels of all frames) is ONE launch over a tile table.
//                         MODE_TMA       pixel tile (+halo) staged by one 3-D TMA box load
//                         MODE_LDG       same, staged with plain loads (vstep not a multiple of 16)
//                         MODE_SCOREMAP  the literal fastExtract stage: the score tile is read from a
//                                        caller-supplied `out` map instead of being computed
//   order_kernel        one CTA per frame: turns the per-(segment,tile) pieces the tiles wrote into the
//                       reference's keypoint order (Fast.h:210,228 / buckets :211-227,326-352)
//   fast_detect_kernel, fast_score_harris_kernel
//                       the literal per-level stages of the template API (Fast.h:54-158, :166-180)
//
// Nothing here mirrors the reference's NEON code: the rules are the ones stated in SURVEY.md section 8a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "device_ops.cuh"

namespace pb {

constexpr int TW = 128;             // tile width  (output pixels)
constexpr int TH = 32;              // tile height (output pixels)
constexpr int NT = 256;             // threads per CTA
constexpr int PIX_W = TW + 32;      // staged pixels: origin (x0-8-dxo, y0-4); a TMA box must start on a
                                    // 16-byte boundary of the row, so dxo = (x0-8) & 15 there, 0 otherwise
constexpr int PIX_H = TH + 10;
constexpr int SC_W = TW + 8;        // score tile: origin (x0-4, y0-1)
constexpr int SC_H = TH + 2;
constexpr int SC_WW = SC_W / 4;     // score-tile words per row
constexpr int NWORD = SC_WW * SC_H; // FAST work items (4 pixels each)
constexpr int CELLS_X = TW / 2, CELLS_Y = TH / 2;
constexpr int NCELL = CELLS_X * CELLS_Y;  // 1024 = 4 per thread
static_assert(NCELL == 4 * NT, "four cells per thread");
constexpr int MAX_LEVELS = 16;

enum { MODE_TMA = 0, MODE_LDG = 1, MODE_SCOREMAP = 2 };

struct LevelInfo {
  int w, h, row_off;  // level size and first stack row
  int xend;           // border + 16*ceil((w-2*border)/16): end of the columns fastDetect writes
  int ntile_cols;
  int piece_base;     // index of the level's first piece
};

struct TileDesc {
  int level, x0, y0, piece0;  // piece of segment s = piece0 + s*ntile_cols
};

struct FrontParams {
  const uint8_t* src;  // MODE_LDG: stack of frame 0; MODE_SCOREMAP: `out` map of frame 0 (level row 0)
  unsigned long long frame_stride;
  int vstep, nlevels;
  LevelInfo lv[MAX_LEVELS];
  int border, fast_t;
  int32_t harris_t;
  int log_bucket, bucket_limit;
  int cap;      // staging / record capacity per frame
  int npieces;  // per frame
  const TileDesc* tiles;
  uint32_t* stage;        // [nframes][cap]
  uint32_t* stage_count;  // [nframes]
  uint32_t* piece_cnt;    // [nframes][npieces]
  uint32_t* piece_off;    // [nframes][npieces]
};

// --------------------------------------------------------------------------------------------------
// TMA + mbarrier (PTX)
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a TMA that never lands traps (aborting the launch) after ~2 s instead of hanging the GPU
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
#pragma unroll 1
  for (;;) {
    if (mbar_try_wait(bar, parity)) return;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

// --------------------------------------------------------------------------------------------------
// SWAR helpers: four pixels per 32-bit word
// --------------------------------------------------------------------------------------------------
// bit 7 of every byte: (byte of d) > t, for a uniform t in 0..255; k7 = (0x7f - (t & 0x7f)) * 0x01010101
__device__ __forceinline__ uint32_t bytes_gt(uint32_t d, uint32_t k7, bool t_lt_128) {
  const uint32_t s = (d & 0x7f7f7f7fu) + k7;
  return t_lt_128 ? (d | s) : (d & s);
}

// Warp-aggregated append of one 16-bit item per participating lane to a shared list.
__device__ __forceinline__ void warp_push(bool want, uint16_t item, uint16_t* list, uint32_t* counter) {
  const uint32_t b = __ballot_sync(0xffffffffu, want);
  if (b == 0) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(b) - 1;
  uint32_t pos = 0;
  if (lane == leader) pos = atomicAdd(counter, (uint32_t)__popc(b));
  pos = __shfl_sync(0xffffffffu, pos, leader);
  if (want) list[pos + __popc(b & ((1u << lane) - 1u))] = item;
}

// Non-max rule of Fast.h:264-310 in position-independent form: a pixel survives iff it is >= the four
// neighbours that precede it in raster order and > the four that follow (SURVEY.md spec F3).
__device__ __forceinline__ bool nms_survives(const uint8_t* q, int pitch) {
  const int s = q[0];
  if (s == 0) return false;
  return s >= q[-pitch - 1] && s >= q[-pitch] && s >= q[-pitch + 1] && s >= q[-1] && s > q[1] &&
         s > q[pitch - 1] && s > q[pitch] && s > q[pitch + 1];
}

// --------------------------------------------------------------------------------------------------
// The fused tile kernel
// --------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(NT) front_tile_kernel(const __grid_constant__ FrontParams P,
                                                        const __grid_constant__ CUtensorMap tmap) {
  __shared__ __align__(128) uint8_t s_pix[PIX_H * PIX_W];
  __shared__ __align__(16) uint8_t s_score[SC_H * SC_W];
  __shared__ uint16_t s_cand[NWORD];       // (work item << 4) | 4-bit candidate mask
  __shared__ uint16_t s_hlist[SC_H * SC_W];  // score-tile indices of FAST corners to be Harris-scored
  __shared__ uint32_t s_cell[NCELL];
  __shared__ uint32_t s_cell2[NCELL];
  __shared__ uint32_t s_tprefix[NT + 1];
  __shared__ uint32_t s_warp[NT / 32];
  __shared__ uint32_t s_ncand, s_nh, s_base;
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int frame = blockIdx.y;
  const TileDesc td = P.tiles[blockIdx.x];
  const LevelInfo L = P.lv[td.level];
  const int B = P.border;
  const int x0 = td.x0, y0 = td.y0;
  const int xs_end = L.w - B;  // end of the Harris-scored columns and of the cell origins
  const int ys_end = L.h - B;
  // column of the staged tile that holds level column x0-8 (a multiple of 4: the host checks border % 4)
  const int dxo = (MODE == MODE_TMA) ? ((x0 - 8) & 15) : 0;

  if (MODE == MODE_TMA) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&s_bar, PIX_H * PIX_W);
      tma_load_3d(s_pix, &tmap, x0 - 8 - dxo, L.row_off + y0 - 4, frame, &s_bar);
    }
  }
  if (tid == 0) {
    s_ncand = 0;
    s_nh = 0;
  }

  if (MODE == MODE_SCOREMAP) {
    // literal fastExtract: the 3x3 neighbourhoods of all cell pixels, straight from the caller's map
    const uint8_t* out = P.src + (size_t)frame * P.frame_stride;
    const int xc_end = B + 2 * ((xs_end - B + 1) / 2), yc_end = B + 2 * ((ys_end - B + 1) / 2);
    for (int i = tid; i < SC_H * SC_W; i += NT) {
      const int ry = i / SC_W, sx = i - ry * SC_W;
      const int x = x0 - 4 + sx, y = y0 - 1 + ry;
      uint8_t v = 0;
      if (x >= B - 1 && x <= xc_end && y >= B - 1 && y <= yc_end) v = out[(size_t)y * P.vstep + x];
      s_score[i] = v;
    }
  } else {
    for (int i = tid; i < SC_H * SC_W / 4; i += NT) reinterpret_cast<uint32_t*>(s_score)[i] = 0u;
    if (MODE == MODE_LDG) {
      // linear addressing like the reference's img[y][x]; positions outside the frame read as zero
      const uint8_t* img = P.src + (size_t)frame * P.frame_stride;
      const long long lo = -(long long)L.row_off * P.vstep;
      const long long hi = (long long)P.frame_stride - (long long)L.row_off * P.vstep;
      const uint8_t* lvl = img + (size_t)L.row_off * P.vstep;
      for (int i = tid; i < PIX_H * PIX_W; i += NT) {
        const int py = i / PIX_W, px = i - py * PIX_W;
        const int x = x0 - 8 + px;
        const long long a = (long long)(y0 - 4 + py) * P.vstep + x;
        s_pix[i] = (x >= 0 && x < P.vstep && a >= lo && a < hi) ? lvl[a] : (uint8_t)0;
      }
    }
  }
  __syncthreads();

  if (MODE != MODE_SCOREMAP) {
    if (MODE == MODE_TMA) mbar_wait(&s_bar, 0);

    // ---- phase A: SWAR pre-test on the four compass ring pixels ---------------------------------
    // A 9-arc always contains two neighbouring compass points (ring 1,5,9,13), so a corner needs
    // |p-c| > t on (N or S) and on (E or W).  Exact FAST-9 is decided in phase B on the survivors.
    const int t = P.fast_t & 0xff;  // Fast.h:58 keeps the low byte
    const uint32_t k7 = (0x7fu - (uint32_t)(t & 0x7f)) * 0x01010101u;
    const bool t_lt_128 = t < 128;
    const int xlo = max(B, x0 - 1), xhi = min(L.xend, x0 + TW + 1);  // columns whose flag matters here
    const uint32_t* pw = reinterpret_cast<const uint32_t*>(s_pix);
    for (int base = 0; base < NWORD; base += NT) {
      const int id = base + tid;
      uint32_t nib = 0;
      if (id < NWORD) {
        const int ry = id / SC_WW, wx = id - ry * SC_WW;
        const int y = y0 - 1 + ry, x = x0 - 4 + 4 * wx;
        const int nlo = min(max(xlo - x, 0), 4), nhi = min(max(xhi - x, 0), 4);
        if (y >= B && y < ys_end && nlo < nhi) {
          const uint32_t* c = pw + (ry + 3) * (PIX_W / 4) + (dxo >> 2) + 1 + wx;
          const uint32_t C = c[0];
          const uint32_t E = __funnelshift_r(C, c[1], 24);   // pixels x+3 .. x+6
          const uint32_t W = __funnelshift_r(c[-1], C, 8);   // pixels x-3 .. x
          const uint32_t N = c[-3 * (PIX_W / 4)], S = c[3 * (PIX_W / 4)];
          const uint32_t ns = bytes_gt(__vabsdiffu4(N, C), k7, t_lt_128) | bytes_gt(__vabsdiffu4(S, C), k7, t_lt_128);
          const uint32_t ew = bytes_gt(__vabsdiffu4(E, C), k7, t_lt_128) | bytes_gt(__vabsdiffu4(W, C), k7, t_lt_128);
          uint32_t m = ns & ew & 0x80808080u;
          m &= (0x80808080u << (8 * nlo)) & (0x80808080u >> (8 * (4 - nhi)));
          nib = (((m >> 7) * 0x01020408u) >> 24) & 15u;
        }
      }
      warp_push(nib != 0, (uint16_t)((id << 4) | nib), s_cand, &s_ncand);
    }
    __syncthreads();

    // ---- phase B: exact FAST-9 on the candidates (Fast.h:54-158) ---------------------------------
    const int ncand = (int)s_ncand;
    for (int base = 0; base < ncand; base += NT) {
      const int i = base + tid;
      const uint32_t e = i < ncand ? s_cand[i] : 0u;
      const int id = (int)(e >> 4);
      const int ry = id / SC_WW, wx = id - ry * SC_WW;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const int sx = 4 * wx + k;
        bool corner = false;
        if ((e >> k) & 1u) corner = fast9_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + sx, PIX_W, t);
        // columns >= w-border are never scored (Fast.h:172): they keep the raw 0xff flag
        const bool scored = corner && (x0 - 4 + sx) < xs_end;
        if (corner && !scored) s_score[ry * SC_W + sx] = 0xff;
        warp_push(scored, (uint16_t)(ry * SC_W + sx), s_hlist, &s_nh);
      }
    }
    __syncthreads();

    // ---- phase C: Harris code of every flagged pixel (Fast.h:166-180, Harris.h:80-248) ------------
    const int nh = (int)s_nh;
    for (int i = tid; i < nh; i += NT) {
      const int idx = s_hlist[i];
      const int ry = idx / SC_W, sx = idx - ry * SC_W;
      s_score[idx] = harris_code_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + sx, PIX_W, P.harris_t);
    }
    __syncthreads();
  }

  // ---- phase D: non-max suppression, one 2x2 cell per slot (Fast.h:228-310) ------------------------
  // slot order == output order inside a segment: row-major cells (plain) or band / bucket /
  // position-in-bucket (bucketed, then re-ranked by encoded value in phase D2)
  const int lb = P.log_bucket;
  const int cb_log = lb > 0 ? lb - 1 : 0;  // log2(cells per bucket side)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + NT * j;
    const int cx = c & (CELLS_X - 1), cy = c >> 6;
    int slot = c;
    if (lb > 0) {
      const int cbm = (1 << cb_log) - 1;
      slot = ((cy >> cb_log) << (6 + cb_log)) + ((cx >> cb_log) << (2 * cb_log)) + ((cy & cbm) << cb_log) + (cx & cbm);
    }
    uint32_t val = 0;
    const int x = x0 + 2 * cx, y = y0 + 2 * cy;
    if (x < xs_end && y < ys_end) {
      const uint8_t* q = s_score + (1 + 2 * cy) * SC_W + 4 + 2 * cx;
      const uint32_t top = *reinterpret_cast<const uint16_t*>(q);
      const uint32_t bot = *reinterpret_cast<const uint16_t*>(q + SC_W);
      if (top | bot) {
        if (nms_survives(q, SC_W)) val = encode_fast(q[0], x, y);
        else if (nms_survives(q + 1, SC_W)) val = encode_fast(q[1], x + 1, y);
        else if (nms_survives(q + SC_W, SC_W)) val = encode_fast(q[SC_W], x, y + 1);
        else if (nms_survives(q + SC_W + 1, SC_W)) val = encode_fast(q[SC_W + 1], x + 1, y + 1);
      }
    }
    s_cell[slot] = val;
    s_cell2[tid + NT * j] = 0u;
  }
  __syncthreads();

  const uint32_t* cells = s_cell;
  if (lb > 0) {
    // ---- phase D2: per bucket keep the bucket_limit largest encoded values, ascending (Fast.h:321-352)
    const int bslots = 1 << (2 * cb_log);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      const int s = tid + NT * j;
      const uint32_t v = s_cell[s];
      if (v != 0u) {
        const int b0 = s & ~(bslots - 1);
        int cnt = 0, rank = 0;
        for (int k = 0; k < bslots; ++k) {
          const uint32_t o = s_cell[b0 + k];
          cnt += o != 0u;
          rank += o > v;
        }
        const int kept = min(cnt, P.bucket_limit);
        if (rank < kept) s_cell2[b0 + (kept - 1 - rank)] = v;
      }
    }
    __syncthreads();
    cells = s_cell2;
  }

  // ---- phase E: ordered compaction; one staging segment per tile, one piece per (segment, tile) -----
  uint32_t v4[4];
  uint32_t mine = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v4[j] = cells[4 * tid + j];
    mine += v4[j] != 0u;
  }
  uint32_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= d) incl += o;
  }
  if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w)
    if (w < (tid >> 5)) wbase += s_warp[w];
  const uint32_t excl = wbase + incl - mine;
  s_tprefix[tid] = excl;
  if (tid == NT - 1) {
    const uint32_t total = excl + mine;
    s_tprefix[NT] = total;
    s_base = total ? atomicAdd(P.stage_count + frame, total) : 0u;
  }
  __syncthreads();
  const uint32_t base = s_base;
  uint32_t* stage = P.stage + (size_t)frame * P.cap;
  uint32_t pos = base + excl;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (v4[j] != 0u) {
      if (pos < (uint32_t)P.cap) stage[pos] = v4[j];
      ++pos;
    }
  // a segment is one cell row (64 slots) or one bucket band (64 << cb_log slots)
  const int seg_threads = 16 << cb_log;  // threads (4 slots each) per segment
  const int nseg = NT / seg_threads;
  if (tid < nseg) {
    const uint32_t a = s_tprefix[tid * seg_threads], b = s_tprefix[(tid + 1) * seg_threads];
    const size_t pi = (size_t)frame * P.npieces + td.piece0 + tid * L.ntile_cols;
    P.piece_cnt[pi] = b - a;
    P.piece_off[pi] = base + a;
  }
}

// --------------------------------------------------------------------------------------------------
// Keypoint records
// --------------------------------------------------------------------------------------------------
constexpr int REC_HEADER_WORDS = 32;  // [0]=n stored, [1]=n found, [2..2+nlevels]=first index of each level

struct RecordLayout {
  unsigned long long bytes, kp_off, desc_off, level_off;
  __host__ __device__ static RecordLayout make(int cap, int words) {
    RecordLayout r;
    r.kp_off = REC_HEADER_WORDS * 4ull;
    r.desc_off = r.kp_off + 4ull * cap;
    r.level_off = r.desc_off + 4ull * cap * words;
    r.bytes = (r.level_off + (unsigned long long)cap + 15ull) & ~15ull;
    return r;
  }
};

struct OrderParams {
  int nlevels, npieces, cap;
  int piece_base[MAX_LEVELS + 1];
  const uint32_t* stage;
  uint32_t* stage_count;
  const uint32_t* piece_cnt;
  const uint32_t* piece_off;
  uint8_t* records;  // [nframes][rec.bytes]
  RecordLayout rec;
};

constexpr int ORDER_NT = 1024;

__global__ void __launch_bounds__(ORDER_NT) order_kernel(const __grid_constant__ OrderParams P) {
  __shared__ uint32_t s_warp[ORDER_NT / 32];
  __shared__ uint32_t s_running;
  const int tid = threadIdx.x, frame = blockIdx.x;
  const uint32_t* pcnt = P.piece_cnt + (size_t)frame * P.npieces;
  const uint32_t* poff = P.piece_off + (size_t)frame * P.npieces;
  const uint32_t* stage = P.stage + (size_t)frame * P.cap;
  uint8_t* rec = P.records + (size_t)frame * P.rec.bytes;
  uint32_t* header = reinterpret_cast<uint32_t*>(rec);
  uint32_t* kp = reinterpret_cast<uint32_t*>(rec + P.rec.kp_off);
  uint8_t* lvl = rec + P.rec.level_off;
  const uint32_t cap = (uint32_t)P.cap;
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < P.npieces; base += ORDER_NT) {
    const int p = base + tid;
    const uint32_t c = p < P.npieces ? pcnt[p] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += o;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    uint32_t wbase = s_running;
    for (int w = 0; w < (tid >> 5); ++w) wbase += s_warp[w];
    const uint32_t excl = wbase + incl - c;
    if (p < P.npieces) {
      int level = 0;
      for (int l = 0; l < P.nlevels; ++l) {
        if (P.piece_base[l] == p) header[2 + l] = min(excl, cap);
        if (P.piece_base[l] <= p) level = l;
      }
      const uint32_t off = poff[p];
      for (uint32_t j = 0; j < c; ++j) {
        const uint32_t d = excl + j;
        if (d < cap && off + j < cap) {
          kp[d] = stage[off + j];
          lvl[d] = (uint8_t)level;
        }
      }
    }
    __syncthreads();
    if (tid == ORDER_NT - 1) s_running = excl + c;
    __syncthreads();
  }
  if (tid == 0) {
    const uint32_t total = s_running;
    const uint32_t staged = P.stage_count[frame];
    // staging overflow loses entries: report it as "found more than stored"
    const uint32_t n = (staged > cap) ? 0u : min(total, cap);
    header[0] = n;
    header[1] = max(total, staged);
    for (int l = 0; l <= P.nlevels; ++l)
      if (l == P.nlevels || P.piece_base[l] >= P.npieces) header[2 + l] = n;
    P.stage_count[frame] = 0u;  // ready for the next batch
  }
}

// --------------------------------------------------------------------------------------------------
// Literal stages on global memory, addressed linearly like the reference's img[y][x] on a
// uint8_t(*)[vstep] (an x past the pitch lands in the next row).
// --------------------------------------------------------------------------------------------------
// fastDetect (Fast.h:54-158): rows [border, h-border), 16-pixel blocks from `border` while < w-border.
__global__ void fast_detect_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int vstep, int border,
                                   int width, int height, int threshold) {
  const int xend = border + 16 * ((width - 2 * border + 15) / 16);
  const int x = border + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = border + blockIdx.y;
  if (y >= height - border) return;
  if (x < xend) {
    const size_t a = (size_t)y * vstep + x;
    out[a] = fast9_at(img + a, vstep, threshold & 0xff) ? 0xff : 0x00;
  }
}
// Fast.h:153-156 runs after the row's blocks: out[y][w] = out[y][w+1] = 0 when w % 16 != 0
__global__ void fast_detect_tail_kernel(uint8_t* __restrict__ out, int vstep, int border, int width, int height) {
  const int y = border + blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= height - border) return;
  out[(size_t)y * vstep + width] = 0;
  out[(size_t)y * vstep + width + 1] = 0;
}

// fastScoreHarris (Fast.h:166-180)
__global__ void fast_score_harris_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int vstep,
                                         int border, int width, int height, int32_t threshold) {
  const int x = border + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = border + blockIdx.y;
  if (y >= height - border || x >= width - border) return;
  const size_t a = (size_t)y * vstep + x;
  if (out[a] != 0) out[a] = harris_code_at(img + a, vstep, threshold);
}

// harrisScoreSobel at a list of points (Harris.h:80-248)
__global__ void harris_points_kernel(const uint8_t* __restrict__ img, int vstep, const int* __restrict__ xs,
                                     const int* __restrict__ ys, int n, int32_t threshold,
                                     uint8_t* __restrict__ scores) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) scores[i] = harris_code_at(img + (size_t)ys[i] * vstep + xs[i], vstep, threshold);
}

}  // namespace pb
