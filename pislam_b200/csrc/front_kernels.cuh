// Detector front-end kernels (sm_100a): FAST-9 -> Harris code -> 3x3 non-max -> ordered keypoint list.
//
//   front_tile_kernel   one CTA per 128x64 tile of one pyramid level of one frame; the whole batch
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
constexpr int NT = 128;             // threads per CTA
constexpr int PIX_W = TW + 32;      // staged pixels: origin (x0-8-dxo, y0-4); a TMA box must start on a
                                    // 16-byte boundary of the row, so dxo = (x0-8) & 15 there, 0 otherwise
constexpr int SC_W = TW + 8;        // score tile: origin (x0-4, y0-1)
constexpr int SC_WW = SC_W / 4;     // score-tile words per row
constexpr int CELLS_X = TW / 2;
constexpr int A_COLS = SC_WW;       // phase A: 34 word columns
constexpr int A_ROWS = 6;           // phase A: rows per task (the 66 score rows of a 64-row tile are 11 groups of 6)

// Everything that depends on the tile height.  Two heights are compiled: 64 for batches (per-tile set-up, the
// one-warp tail phases and the halo rows are amortised over twice the pixels: 0.310 vs 0.347 ms per 256 VGA
// frames) and 32 for calls with few tiles (a single VGA pyramid is 119 tiles of 64 rows for 148 SMs).
template <int TH_>
struct TileCfg {
  static_assert(TH_ == 32 || TH_ == 64, "tile heights 32 and 64 are instantiated");
  static constexpr int TH = TH_;              // tile height (output pixels)
  static constexpr int PIX_H = TH + 10;       // staged rows
  static constexpr int SC_H = TH + 2;         // score-tile rows
  static constexpr int CELLS_Y = TH / 2;
  static constexpr int NCELL = CELLS_X * CELLS_Y;  // 2048 / 1024 2x2 cells per tile
  static constexpr int A_GROUPS = (SC_H + A_ROWS - 1) / A_ROWS;  // 11 / 6 row groups
  static constexpr int A_TASKS = A_COLS * A_GROUPS;              // 374 / 204 (word column, row group) tasks
  // rows the last (partly invalid) row group of phase A reads beyond the staged tile: kept inside s_pix by padding
  static constexpr int PIX_PAD_ROWS = A_GROUPS * A_ROWS + 6 > PIX_H ? A_GROUPS * A_ROWS + 6 - PIX_H : 0;
  // Work-list capacities.  Natural images need ~130 candidates / ~50 corners per 4096 pixels; the lists are sized
  // for ~10x that and anything beyond is handled EXACTLY but slowly (inline, then a dense non-max scan), so that
  // adversarial tiles (noise, checkerboards) stay bit-exact while the common case keeps shared memory small.
  static constexpr int CAP_A = 2048;   // list A: FAST candidates, later non-zero scores (one per corner <= CAP_B)
  static constexpr int CAP_B = NCELL;  // list B: FAST corners, later NMS survivors (at most one per cell)
  // resident CTAs per SM the register budget is set for: 30 KB / 72 registers (TH 64; 6 CTAs measured +7 %),
  // 18 KB / 48 registers (TH 32; 8: +3 %, 9: +0.7 %, 11-12: +4 % with spills)
  static constexpr int MIN_CTAS = TH == 64 ? 7 : 10;
  static_assert(NCELL % 1024 == 0 && NCELL / 32 <= 64, "phase E: one or two bitmap words per lane of one warp");
  static_assert(CAP_A >= CAP_B && CAP_B >= NCELL, "list A: one entry per corner of list B; list B: one survivor per cell");
  static_assert(NCELL / 32 + A_COLS + A_GROUPS + 1 <= NT, "set-up roles fit the CTA");
};
constexpr int TH_BATCH = 64, TH_SMALL = 32;
constexpr int MAX_LEVELS = 16;

enum { MODE_TMA = 0, MODE_LDG = 1, MODE_SCOREMAP = 2 };

struct LevelInfo {
  int w, h, row_off;  // level size and first stack row
  int xend;           // border + 16*ceil((w-2*border)/16): end of the columns fastDetect writes
  int ntile_cols;
  int piece_base;     // index of the level's first piece
  int spill;          // fastDetect's last 16-pixel block runs past the row pitch far enough (xend > vstep + border - 1)
                      // for its flags to land on column border-1 of the NEXT row, which fastExtract reads
};

struct __align__(16) TileDesc {  // one 16-byte load per CTA
  int level, x0, y0, piece0;  // piece of segment s = piece0 + s*ntile_cols
};

struct FrontParams {
  const uint8_t* src;  // MODE_LDG: stack of frame 0; MODE_SCOREMAP: `out` map of frame 0 (level row 0)
  unsigned long long frame_stride;
  unsigned long long frame_bytes;  // readable bytes per frame (rows * vstep <= frame_stride): bounds MODE_LDG's reads
  int vstep, nlevels;
  LevelInfo lv[MAX_LEVELS];
  int border, fast_t;
  int32_t harris_t;
  int log_bucket, bucket_limit;
  int cap;      // record capacity per frame; also the size of the staging overflow region
  int npieces;  // per frame
  const uint8_t* zeros;  // >= 16 KB of zeros, 16-byte aligned: source of the score tile's bulk clear (MODE_TMA)
  int prefetch_dist;  // CTAs resident on the whole GPU (0: no L2 prefetch of a later CTA's tile)
  int tile_slot;   // staging entries every tile owns (a tile with more survivors reserves in the overflow region)
  int stage_cap;   // staging entries per frame = ntiles * tile_slot + cap
  const TileDesc* tiles;
  uint32_t* stage;        // [nframes][stage_cap]
  uint32_t* stage_count;  // [nframes] entries used in the overflow region
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
// One lane of the (fully converged) warp adds `v` to the shared counter at `addr`; every lane gets the old value.
// elect.sync keeps ptxas from wrapping the atomic into its generic warp-aggregation sequence (17 instructions
// per call site, 3.6 % of the kernel's instructions before), REDUX broadcasts the result.
__device__ __forceinline__ uint32_t warp_atom_shared_add(uint32_t addr, uint32_t v) {
  uint32_t old = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
      : "+r"(old)
      : "r"(addr), "r"(v)
      : "memory");
  return __reduce_max_sync(0xffffffffu, old);
}

// Warp-aggregated reservation in a shared list: every lane of the warp calls it (uniform loops only).
// Returns the lane's position (meaningless when !want).
__device__ __forceinline__ uint32_t warp_reserve(bool want, uint32_t* counter) {
  const uint32_t b = __ballot_sync(0xffffffffu, want);
  if (b == 0) return 0u;
  const int lane = threadIdx.x & 31;
  const uint32_t pos = warp_atom_shared_add(smem_u32(counter), (uint32_t)__popc(b));
  return pos + __popc(b & ((1u << lane) - 1u));
}
// append of one 16-bit item per participating lane
__device__ __forceinline__ void warp_push(bool want, uint16_t item, uint16_t* list, uint32_t* counter) {
  const uint32_t pos = warp_reserve(want, counter);
  if (want) list[pos] = item;
}

// if (bits & bit) { *(uint16_t*)addr = v; addr += 2; }  as two predicated instructions
__device__ __forceinline__ void push_if_bit(uint32_t& addr, uint32_t bits, uint32_t bit, uint16_t v) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 m;\n\t"
      "and.b32 m, %2, %3;\n\t"
      "setp.ne.u32 p, m, 0;\n\t"
      "@p st.shared.u16 [%0], %1;\n\t"
      "@p add.u32 %0, %0, 2;\n\t}"
      : "+r"(addr)
      : "h"(v), "r"(bits), "r"(bit));
}

// Non-max rule of Fast.h:264-310 in position-independent form: a pixel survives iff it is >= the four
// neighbours that precede it in raster order and > the four that follow (SURVEY.md spec F3).
// Branch-free: the nine loads are independent (a chain of early exits costs a shared-memory round trip per test).
__device__ __forceinline__ bool nms_survives(const uint8_t* q, int pitch) {
  const int s = q[0];
  const int a0 = q[-pitch - 1], a1 = q[-pitch], a2 = q[-pitch + 1], a3 = q[-1];
  const int b0 = q[1], b1 = q[pitch - 1], b2 = q[pitch], b3 = q[pitch + 1];
  const int before = max(max(a0, a1), max(a2, a3)), after = max(max(b0, b1), max(b2, b3));
  return (s != 0) & (s >= before) & (s > after);
}

// --------------------------------------------------------------------------------------------------
// The fused tile kernel
// --------------------------------------------------------------------------------------------------
// Work lists keep every phase dense: A1 ((N or S) on every task of 6 rows x 4 pixels) -> surviving tasks -> A2
// (compass + diagonal SWAR tests) -> candidates -> B (exact FAST-9) -> corners -> C (Harris code) -> non-zero
// scores -> D (non-max) -> survivors -> D2 (bucket top-k) -> bitmap over output slots -> E (rank by popcount
// prefix = the reference's order).
static_assert(A_ROWS >= 5 && A_ROWS <= 8, "one flag bit per row and pixel inside a 32-bit task mask");

// Slow path of a corner that did not fit into list B: score it in place (rarely executed, kept out of line)
// List entries are "rc" codes: (score-tile row << 8) | score-tile column (rows <= 65, columns <= 135), so that
// neither the pixel address nor the score index needs a division.
__device__ __forceinline__ int rc_row(uint32_t rc) { return (int)(rc >> 8) & 0x7f; }
__device__ __forceinline__ int rc_col(uint32_t rc) { return (int)(rc & 0xffu); }

__device__ __noinline__ void score_corner_inline(const uint8_t* s_pix, uint8_t* s_score, uint32_t rc, int dxo,
                                                 bool scored, int32_t harris_t) {
  const int ry = rc_row(rc), sx = rc_col(rc);
  const uint32_t code = scored ? harris_code_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + sx, PIX_W, harris_t) : 0xffu;
  if (code) s_score[ry * SC_W + sx] = (uint8_t)code;
}
// Slow path of a candidate that did not fit into list A: exact FAST-9 and score in place
__device__ __noinline__ void test_candidate_inline(const uint8_t* s_pix, uint8_t* s_score, uint32_t rc, int dxo, int t,
                                                   bool scored, int32_t harris_t) {
  const int ry = rc_row(rc), sx = rc_col(rc);
  if (fast9_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + sx, PIX_W, t))
    score_corner_inline(s_pix, s_score, rc, dxo, scored, harris_t);
}

constexpr unsigned TAIL_MAX = 32;  // non-zero scores up to which one warp runs phases D/D2/E alone (64 / 128: no gain)
template <int MODE, int TH_>
__global__ void __launch_bounds__(NT, TileCfg<TH_>::MIN_CTAS) front_tile_kernel(const __grid_constant__ FrontParams P,
                                                                              const __grid_constant__ CUtensorMap tmap) {
  using Cfg = TileCfg<TH_>;
  constexpr int TH = Cfg::TH, PIX_H = Cfg::PIX_H, SC_H = Cfg::SC_H, NCELL = Cfg::NCELL, A_GROUPS = Cfg::A_GROUPS,
                A_TASKS = Cfg::A_TASKS, PIX_PAD_ROWS = Cfg::PIX_PAD_ROWS, CAP_A = Cfg::CAP_A, CAP_B = Cfg::CAP_B;
  __shared__ __align__(128) uint8_t s_pix[(PIX_H + PIX_PAD_ROWS) * PIX_W];
  __shared__ __align__(16) uint8_t s_score[SC_H * SC_W];
  __shared__ __align__(16) uint16_t s_listA[CAP_A];  // FAST candidates (A->B), then non-zero scores (C->D)
  __shared__ __align__(8) uint16_t s_listB[CAP_B];   // FAST corners (B->C), then NMS survivors (D->D2); hosts the
                                                     // phase-A task list before that
  __shared__ uint32_t s_occ[NCELL / 32], s_bm[NCELL / 32];
  __shared__ uint32_t s_vmcol[A_COLS], s_vmrow[A_GROUPS + 1];  // phase-A validity masks per word column / row group
  __shared__ uint32_t s_nA, s_nC, s_nS;
  __shared__ uint32_t s_nW[NT / 32];  // entries per warp segment: of the task list (A1 -> A2), then of list B (B -> C)
  __shared__ uint32_t s_dense;  // a list overflowed: scores were written in place, phase D scans every cell
  __shared__ __align__(8) uint64_t s_bar;
  // the slot arrays reuse buffers that are dead by then: the pixels after phase C, the scores after phase D
  static_assert(sizeof(s_pix) >= NCELL * 4 && sizeof(s_score) >= NCELL * 4, "slot arrays alias s_pix / s_score");
  uint32_t* const s_cell = reinterpret_cast<uint32_t*>(s_pix);
  uint32_t* const s_cell2 = reinterpret_cast<uint32_t*>(s_score);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int frame = blockIdx.y;
  const TileDesc td = P.tiles[blockIdx.x];
  const LevelInfo L = P.lv[td.level];
  const int B = P.border;
  const int x0 = td.x0, y0 = td.y0;
  const int xs_end = L.w - B;  // end of the Harris-scored columns and of the cell origins
  const int ys_end = L.h - B;
  const int lb = P.log_bucket;
  const int cb_log = lb > 0 ? lb - 1 : 0;  // log2(cells per bucket side)
  // column of the staged tile that holds level column x0-8 (a multiple of 4: the host checks border % 4)
  const int dxo = (MODE == MODE_TMA) ? ((x0 - 8) & 15) : 0;

  if (MODE == MODE_TMA) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&s_bar, PIX_H * PIX_W + SC_H * SC_W);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_score)),
          "l"(P.zeros), "r"(SC_H * SC_W), "r"(smem_u32(&s_bar))
          : "memory");
      tma_load_3d(s_pix, &tmap, x0 - 8 - dxo, L.row_off + y0 - 4, frame, &s_bar);
    }
  }
  if (tid == 0) {
    s_nA = 0;
    s_nC = 0;
    s_nS = 0;
    s_dense = 0;
  }
  if (tid < NCELL / 32) {
    s_bm[tid] = 0u;
    s_occ[tid] = 0u;
  } else if (MODE != MODE_SCOREMAP && tid < NCELL / 32 + A_COLS) {
    // phase-A validity masks: columns of word column wx whose flag matters (B <= x < xend, inside the tile + 1)
    const int wx = tid - NCELL / 32;
    const int x = x0 - 4 + 4 * wx;
    const int nlo = min(max(max(B, x0 - 1) - x, 0), 4), nhi = min(max(min(L.xend, x0 + TW + 1) - x, 0), 4);
    s_vmcol[wx] = nlo < nhi ? ((0x01010101u << (8 * nlo)) & (0x01010101u >> (8 * (4 - nhi)))) * ((1u << A_ROWS) - 1u) : 0u;
  } else if (MODE != MODE_SCOREMAP && tid < NCELL / 32 + A_COLS + A_GROUPS + 1) {
    // ... and rows of row group g (B <= y < h-B, inside the score tile)
    const int g = tid - (NCELL / 32 + A_COLS);
    const int r0 = g * A_ROWS;
    const int y = y0 - 1 + r0;
    const int ilo = min(max(B - y, 0), A_ROWS), ihi = max(min(min(SC_H - r0, ys_end - y), A_ROWS), 0);
    s_vmrow[g] = ilo < ihi ? (((1u << ihi) - 1u) & ~((1u << ilo) - 1u)) * 0x01010101u : 0u;
  }

  if (MODE == MODE_SCOREMAP) {
    __syncthreads();
    // literal fastExtract: the 3x3 neighbourhoods of all cell pixels, straight from the caller's map
    const uint8_t* out = P.src + (size_t)frame * P.frame_stride;
    const int xc_end = B + 2 * ((xs_end - B + 1) / 2), yc_end = B + 2 * ((ys_end - B + 1) / 2);
    for (int base = 0; base < SC_H * SC_W; base += NT) {
      const int i = base + tid;
      bool cand = false;
      if (i < SC_H * SC_W) {
        const int ry = i / SC_W, sx = i - ry * SC_W;
        const int x = x0 - 4 + sx, y = y0 - 1 + ry;
        uint8_t v = 0;
        if (x >= B - 1 && x <= xc_end && y >= B - 1 && y <= yc_end) v = out[(size_t)y * P.vstep + x];
        s_score[i] = v;
        const int cxp = sx - 4, cyp = ry - 1;
        cand = v != 0 && cxp >= 0 && cxp < TW && cyp >= 0 && cyp < TH && x0 + (cxp & ~1) < xs_end &&
               y0 + (cyp & ~1) < ys_end;
      }
      const uint32_t pos = warp_reserve(cand, &s_nC);
      if (cand) {
        if (pos < CAP_A) s_listA[pos] = (uint16_t)(((i / SC_W) << 8) | (i % SC_W));
        else s_dense = 1u;
      }
    }
  } else {
    static_assert((SC_H * SC_W) % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (MODE != MODE_TMA)  // (the TMA mode clears the score tile with a bulk copy from a zero page, issued above)
      for (int i = tid; i < SC_H * SC_W / 16; i += NT) reinterpret_cast<uint4*>(s_score)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (MODE == MODE_LDG) {
      // linear addressing like the reference's img[y][x] on a uint8_t(*)[vstep]: a column past the pitch is the
      // start of the next row (the overrun blocks of Fast.h:36-40,61 read there when w is close to vstep);
      // positions outside the frame read as zero
      const uint8_t* img = P.src + (size_t)frame * P.frame_stride;
      const long long lo = -(long long)L.row_off * P.vstep;
      const long long hi = (long long)P.frame_bytes - (long long)L.row_off * P.vstep;
      const uint8_t* lvl = img + (size_t)L.row_off * P.vstep;
      for (int i = tid; i < PIX_H * PIX_W; i += NT) {
        const int py = i / PIX_W, px = i - py * PIX_W;
        const int x = x0 - 8 + px;
        const long long a = (long long)(y0 - 4 + py) * P.vstep + x;
        s_pix[i] = (a >= lo && a < hi) ? lvl[a] : (uint8_t)0;
      }
    }
    __syncthreads();
    if (MODE == MODE_TMA) mbar_wait(&s_bar, 0);

    // ---- phase A: SWAR pre-tests, a cascade over (word column, 6-row group) tasks --------------------
    // A 9-arc always contains two neighbouring compass points (ring 1,5,9,13) and two neighbouring diagonal points
    // (ring 3,7,11,15), so a corner needs |p-c| > t on (N or S), (E or W), (NE or SW) and (SE or NW).  The tests
    // only have to be supersets of the corners (phase B is exact), which buys the carry-tolerant compare of
    // bytes_gt_loose.  |I[q+3]-I[q]| > t serves as the S test of row q and the N test of row q+3.
    const int t = P.fast_t & 0xff;  // Fast.h:58 keeps the low byte
    const uint32_t listA_s = smem_u32(s_listA), nA_s = smem_u32(&s_nA);
    // ---- phase A1: which tasks have a pixel passing (N or S) at all?  (all pixels; one in four does) ----
    // Lane = word column, so no task decode and conflict-free loads; warp w takes row groups w, w+4, ...; the
    // two word columns beyond the 32 lanes are one extra round of the last warp.  Non-empty tasks go to the
    // warp's own segment of the task list (register counter, no atomic).
    constexpr int NWARP = NT / 32, SEG_T = 256, XCOLS = A_COLS - 32;
    constexpr int MAIN_ROUNDS = (A_GROUPS + NWARP - 1) / NWARP;  // of warp 0
    static_assert(XCOLS == 2 && XCOLS * A_GROUPS <= 32, "the columns beyond the 32 lanes fit one warp round");
    static_assert((MAIN_ROUNDS + 1) * 32 <= SEG_T && NWARP * SEG_T <= CAP_B, "task segments fit list B");
    static_assert(NWARP == 4, "four task segments");
    const int wp = tid >> 5;
    const uint32_t* const pixw = reinterpret_cast<const uint32_t*>(s_pix) + (dxo >> 2) + 1;
    uint16_t* const my_task = s_listB + wp * SEG_T;  // list B is dead until phase B
    uint32_t cntT = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    {
      // one (row group, word column) task per lane: emits its code when a pixel of the task passes (N or S)
      auto probe = [&](bool have, int g, int wx, uint32_t vmc) {
        bool any = false;
        if (have && (vmc & s_vmrow[g]) != 0u) {
          const uint32_t* c = pixw + g * (A_ROWS * (PIX_W / 4)) + wx;
          any = t < 128 ? ns_any<true, PIX_W / 4, A_ROWS>(c, t) : ns_any<false, PIX_W / 4, A_ROWS>(c, t);
        }
        const uint32_t nz = __ballot_sync(0xffffffffu, any);
        if (any) my_task[cntT + __popc(nz & lt_mask)] = (uint16_t)((g << 8) | wx);
        cntT += (uint32_t)__popc(nz);
      };
      const uint32_t vmc = s_vmcol[lane];
#pragma unroll 1
      for (int g = wp; g < A_GROUPS; g += NWARP) probe(true, g, lane, vmc);
      if (wp == NWARP - 1) {
        const bool have = lane < XCOLS * A_GROUPS;
        const int wx = 32 + (lane & 1);
        probe(have, lane >> 1, wx, have ? s_vmcol[wx] : 0u);
      }
    }
    if (lane == 0) s_nW[wp] = cntT;
    __syncthreads();
    const int nT0 = (int)s_nW[0], nT1 = nT0 + (int)s_nW[1], nT2 = nT1 + (int)s_nW[2], nT = nT2 + (int)s_nW[3];
    for (int base = 0; base < nT; base += NT) {
      if (base + (tid & ~31) >= nT) break;  // whole warps without a task
      const int ti = base + tid;
      const int si = ti < nT0 ? ti : (ti < nT1 ? SEG_T + ti - nT0 : (ti < nT2 ? 2 * SEG_T + ti - nT1 : 3 * SEG_T + ti - nT2));
      // ---- phase A2: compass and diagonal tests of the surviving tasks, one per lane of the concatenated
      // segments (dense lanes; 4 % of the pixels come out as candidates, 2.6 per corner) -> list A
      uint32_t acc = 0;  // bit 8k+i: pixel k of row r0+i is a candidate
      int idx0 = 0;
      if (ti < nT) {
        const uint32_t code = s_listB[si];
        const int g = (int)(code >> 8), wx = (int)(code & 0xffu);
        const uint32_t* c = pixw + g * (A_ROWS * (PIX_W / 4)) + wx;
        const uint32_t vm = s_vmcol[wx] & s_vmrow[g];  // bit 8k+i: the flag of pixel k, row r0+i matters here
        if (t < 128)
          acc = pretest_rows<true, PIX_W / 4, A_ROWS>(c, t) & diag_rows<true, PIX_W / 4, A_ROWS>(c, t) & vm;
        else
          acc = pretest_rows<false, PIX_W / 4, A_ROWS>(c, t) & diag_rows<false, PIX_W / 4, A_ROWS>(c, t) & vm;
        idx0 = ((g * A_ROWS) << 8) | (4 * wx);
      }
      const uint32_t cnt = (uint32_t)__popc(acc);
      // one list reservation per warp for all rows
      constexpr uint32_t amask = 0xffffffffu;
      uint32_t incl = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(amask, incl, d);
        if (lane >= d) incl += o;
      }
      const uint32_t pos0 = warp_atom_shared_add(nA_s, __shfl_sync(0xffffffffu, incl, 31));
      uint32_t pos = pos0 + incl - cnt;
      if (pos + cnt <= CAP_A) {
        // 24 predicated stores (a loop over the set bits runs with ~3 of 32 lanes busy); explicit shared
        // addresses: the compiler otherwise rebuilds the shared window base inside every predicated block
        uint32_t a = listA_s + 2u * pos;
#pragma unroll
        for (int i = 0; i < A_ROWS; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) push_if_bit(a, acc, 1u << (8 * k + i), (uint16_t)(idx0 + i * 256 + k));
      } else {  // list A is (nearly) full, adversarial tile: what does not fit is decided and scored right here
        while (acc) {
          const int b = __ffs(acc) - 1;
          acc &= acc - 1;
          const uint32_t rc = (uint32_t)(idx0 + ((b & 7) << 8) + (b >> 3));
          if (pos < CAP_A) {
            s_listA[pos] = (uint16_t)rc;
          } else {
            test_candidate_inline(s_pix, s_score, rc, dxo, t, (x0 - 4 + rc_col(rc)) < xs_end, P.harris_t);
            s_dense = 1u;
          }
          ++pos;
        }
      }
    }
    __syncthreads();

    // ---- phase B: exact FAST-9 on the candidates (Fast.h:54-158) ---------------------------------
    const int nA = min((int)s_nA, CAP_A);
    // corners go to the warp's own quarter of list B, counted in a register (no shared atomic per round)
    constexpr int SEG_B = CAP_B / (NT / 32);
    uint16_t* const my_listB = s_listB + (tid >> 5) * SEG_B;
    uint32_t cntB = 0;
    for (int base = 0; base < nA; base += NT) {
      const int i = base + tid;
      const uint32_t rc = i < nA ? s_listA[i] : 0u;
      const int ry = rc_row(rc), sx = rc_col(rc);
      const bool corner = i < nA && fast9_prmt_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + sx, PIX_W, t);
      // columns >= w-border are never scored (Fast.h:172): they keep the raw 0xff flag
      const bool scored = (x0 - 4 + sx) < xs_end;
      const uint32_t bal = __ballot_sync(0xffffffffu, corner);
      const uint32_t pos = cntB + __popc(bal & ((1u << lane) - 1u));
      cntB += (uint32_t)__popc(bal);
      if (corner) {
        if (pos < (uint32_t)SEG_B) {
          my_listB[pos] = (uint16_t)(rc | (scored ? 0u : 0x8000u));
        } else {  // list B is full: score in place
          score_corner_inline(s_pix, s_score, rc, dxo, scored, P.harris_t);
          s_dense = 1u;
        }
      }
    }
    if (lane == 0) s_nW[tid >> 5] = min(cntB, (uint32_t)SEG_B);
    __syncthreads();

    // ---- phase C: Harris code of every flagged pixel (Fast.h:166-180, Harris.h:80-248) ------------
    const int nB0 = (int)s_nW[0], nB1 = nB0 + (int)s_nW[1], nB2 = nB1 + (int)s_nW[2], nB = nB2 + (int)s_nW[3];
    static_assert(NT / 32 == 4, "four list-B segments");
    for (int base = 0; base < nB; base += NT) {
      const int i = base + tid;
      // corner i of the concatenated segments
      const int si = i < nB0 ? i : (i < nB1 ? SEG_B + i - nB0 : (i < nB2 ? 2 * SEG_B + i - nB1 : 3 * SEG_B + i - nB2));
      const uint32_t e = i < nB ? s_listB[si] : 0u;
      const uint32_t rc = e & 0x7fffu;
      const int ry = rc_row(rc), sx = rc_col(rc);
      uint32_t code = 0;
      if (i < nB) {
        code = (e & 0x8000u) ? 0xffu
                             : harris_code_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + sx, PIX_W, P.harris_t);
        if (code) s_score[ry * SC_W + sx] = (uint8_t)code;
      }
      // non-max candidates: non-zero scores on the tile's own cells
      const int cxp = sx - 4, cyp = ry - 1;
      const bool cand = code != 0 && cxp >= 0 && cxp < TW && cyp >= 0 && cyp < TH && x0 + (cxp & ~1) < xs_end &&
                        y0 + (cyp & ~1) < ys_end;
      warp_push(cand, (uint16_t)rc, s_listA, &s_nC);
    }
    // The reference's fastDetect writes whole 16-pixel blocks with linear addressing: when the last block of row
    // y-1 runs past the pitch, its flag for "column vstep + c" is stored at out[y][c] -- and is in fact the FAST
    // flag of pixel (c, y).  Columns c < border are never rewritten or scored, so for tiny borders (<= 7) the raw
    // 0xff at column border-1 takes part in the non-max test of the first cell column (Fast.h:36-40,61,231-245).
    if (MODE == MODE_LDG && L.spill && x0 == B) {
      for (int ry = tid; ry < SC_H; ry += NT) {
        const int y = y0 - 1 + ry;
        if (y >= B + 1 && y <= ys_end && fast9_at(s_pix + (ry + 3) * PIX_W + dxo + 4 + 3, PIX_W, t))
          s_score[ry * SC_W + 3] = 0xff;
      }
    }
  }
  __syncthreads();

  // ---- phase D: non-max suppression of the non-zero scores (Fast.h:228-310) -------------------------
  // slot = position in the reference's output order inside the tile: row-major 2x2 cells (plain) or
  // band / bucket / position-in-bucket (bucketed; re-ranked by encoded value in phase D2)
  // records a surviving pixel (score-tile index idx) in its slot; returns the slot
  auto emit = [&](int ry, int sx) -> int {
    const int idx = ry * SC_W + sx;
    const int cx = (sx - 4) >> 1, cy = (ry - 1) >> 1;
    int slot = cy * CELLS_X + cx;
    if (lb > 0) {
      const int cbm = (1 << cb_log) - 1;
      slot = ((cy >> cb_log) << (6 + cb_log)) + ((cx >> cb_log) << (2 * cb_log)) + ((cy & cbm) << cb_log) + (cx & cbm);
    }
    s_cell[slot] = encode_fast(s_score[idx], x0 - 4 + sx, y0 - 1 + ry);
    atomicOr((lb == 0 ? s_bm : s_occ) + (slot >> 5), 1u << (slot & 31));
    return slot;
  };
  // A typical tile has a few dozen non-zero scores and keeps a handful of keypoints: one warp finishes it with
  // warp-level synchronisation instead of three more block barriers; the other warps retire.
  const bool warp_tail = !s_dense && s_nC <= TAIL_MAX;
  if (warp_tail && tid >= 32) return;
  const int tail_nt = warp_tail ? 32 : NT;  // threads that take part in phases D and D2
  auto tail_sync = [&]() {
    if (warp_tail) __syncwarp();
    else __syncthreads();
  };
  if (!s_dense) {
    const int nC = (int)s_nC;
    for (int base = 0; base < nC; base += tail_nt) {
      const int i = base + tid;
      const uint32_t rc = i < nC ? s_listA[i] : (uint32_t)((1 << 8) | 4);  // idle lanes probe a harmless interior cell
      const int ry = rc_row(rc), sx = rc_col(rc);
      const bool surv = i < nC && nms_survives(s_score + ry * SC_W + sx, SC_W);
      const int slot = surv ? emit(ry, sx) : 0;
      if (lb > 0) {
        const uint32_t pos = warp_reserve(surv, &s_nS);
        if (surv) s_listB[pos] = (uint16_t)slot;  // at most one survivor per cell: pos < CAP_B
      }
    }
  } else {
    // a work list overflowed on this tile: visit every 2x2 cell (the first survivor in raster order, if any)
    for (int base = 0; base < NCELL; base += NT) {
      const int c = base + tid;
      const int cx = c & (CELLS_X - 1), cy = c >> 6;
      int idx = -1;
      if (x0 + 2 * cx < xs_end && y0 + 2 * cy < ys_end) {
        const int i0 = (1 + 2 * cy) * SC_W + 4 + 2 * cx;
        if (nms_survives(s_score + i0, SC_W)) idx = i0;
        else if (nms_survives(s_score + i0 + 1, SC_W)) idx = i0 + 1;
        else if (nms_survives(s_score + i0 + SC_W, SC_W)) idx = i0 + SC_W;
        else if (nms_survives(s_score + i0 + SC_W + 1, SC_W)) idx = i0 + SC_W + 1;
      }
      const bool surv = idx >= 0;
      const int slot = surv ? emit(idx / SC_W, idx % SC_W) : 0;
      if (lb > 0) {
        const uint32_t pos = warp_reserve(surv, &s_nS);
        if (surv) s_listB[pos] = (uint16_t)slot;
      }
    }
  }
  tail_sync();

  const uint32_t* cells = s_cell;
  if (lb > 0) {
    // ---- phase D2: per bucket keep the bucket_limit largest encoded values, ascending (Fast.h:321-352)
    const int bslots = 1 << (2 * cb_log);
    const int nS = (int)s_nS;
    for (int i = tid; i < nS; i += tail_nt) {
      const int s = s_listB[i];
      const uint32_t v = s_cell[s];
      const int b0 = s & ~(bslots - 1);
      // the bucket's occupied slots: cnt of them, rank = how many hold a larger encoded value
      int cnt = 0, rank = 0;
      const uint32_t wmask = bslots >= 32 ? 0xffffffffu : (((1u << bslots) - 1u) << (b0 & 31));
      for (int wd = b0 >> 5; wd <= (b0 + bslots - 1) >> 5; ++wd) {
        uint32_t bits = s_occ[wd] & wmask;
        cnt += __popc(bits);
        while (bits) {
          const int j = __ffs(bits) - 1;
          bits &= bits - 1;
          rank += s_cell[(wd << 5) + j] > v;
        }
      }
      const int kept = min(cnt, P.bucket_limit);
      if (rank < kept) {
        const int f = b0 + (kept - 1 - rank);
        s_cell2[f] = v;
        atomicOr(&s_bm[f >> 5], 1u << (f & 31));
      }
    }
    tail_sync();
    cells = s_cell2;
  }

  // ---- phase E: rank by slot = popcount prefix over the slot bitmap; one staging segment per tile ----
  // One warp: lane w owns WPL consecutive bitmap words (a tile keeps a handful of keypoints; 128 threads sweeping
  // all slots cost more than the few set bits).
  if (tid < 32) {
    constexpr int WPL = NCELL / 1024;  // bitmap words per lane (consecutive, so the lane order is the slot order)
    uint32_t w[WPL], pc = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      w[k] = s_bm[tid * WPL + k];
      pc += (uint32_t)__popc(w[k]);
    }
    uint32_t incl = pc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    // every tile owns tile_slot staging entries, so the usual tile needs no reservation at all (a global atomic with a
    // return value is a ~1 us round trip at the very end of the CTA's life); a fuller tile reserves in the overflow region
    uint32_t base = blockIdx.x * (uint32_t)P.tile_slot;
    if (total > (uint32_t)P.tile_slot) {
      if (lane == 31) base = gridDim.x * (uint32_t)P.tile_slot + atomicAdd(P.stage_count + frame, total);
      base = __shfl_sync(0xffffffffu, base, 31);
    }
    const uint32_t excl = incl - pc;
    uint32_t* stage = P.stage + (size_t)frame * P.stage_cap;
    uint32_t pos = base + excl;
#pragma unroll
    for (int k = 0; k < WPL; ++k)
      for (uint32_t bits = w[k]; bits; bits &= bits - 1u) {
        if (pos < (uint32_t)P.stage_cap) stage[pos] = cells[(tid * WPL + k) * 32 + (__ffs(bits) - 1)];
        ++pos;
      }
    // a segment is one cell row (64 slots = 2 bitmap words) or one bucket band (64 << cb_log slots): always whole lanes
    const int seg_shift = 1 + cb_log - (WPL == 2 ? 1 : 0);  // log2 of the lanes per segment
    const int seg_lanes = 1 << seg_shift;
    const int nseg = 32 >> seg_shift;
    const uint32_t a = __shfl_sync(0xffffffffu, excl, (tid * seg_lanes) & 31);
    const uint32_t bnext = __shfl_sync(0xffffffffu, excl, ((tid + 1) * seg_lanes) & 31);
    const uint32_t b = (tid + 1) * seg_lanes >= 32 ? total : bnext;
    if (tid < nseg) {
      const size_t pi = (size_t)frame * P.npieces + td.piece0 + tid * L.ntile_cols;
      P.piece_cnt[pi] = b - a;
      P.piece_off[pi] = base + a;
    }
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
  int stage_cap;   // staging entries per frame
  int over_cap;    // size of the staging overflow region (= cap)
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
  const uint32_t* stage = P.stage + (size_t)frame * P.stage_cap;
  uint8_t* rec = P.records + (size_t)frame * P.rec.bytes;
  uint32_t* header = reinterpret_cast<uint32_t*>(rec);
  uint32_t* kp = reinterpret_cast<uint32_t*>(rec + P.rec.kp_off);
  uint8_t* lvl = rec + P.rec.level_off;
  const uint32_t cap = (uint32_t)P.cap;
  if (tid == 0) s_running = 0;
  if (tid < REC_HEADER_WORDS) header[tid] = 0u;  // unused header words are defined (zero)
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
        if (d < cap && off + j < (uint32_t)P.stage_cap) {
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
    // an overflowing overflow region loses entries (its tiles alone hold more than cap keypoints): report it as
    // "found more than stored"
    const uint32_t n = (staged > (uint32_t)P.over_cap) ? 0u : min(total, cap);
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
