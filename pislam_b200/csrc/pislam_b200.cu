// libpislam_b200.so -- host side of the C ABI declared in include/pislam_b200.h.
//
// Everything here is plumbing around the sm_100a kernels in front_kernels.cuh / orb_kernels.cuh:
// device buffers, the tile / piece tables of a pyramid geometry, the TMA tensor map, stream-ordered
// launches, and (optionally) one ncclAllGather of the fixed-capacity frame records.  There is no CPU
// implementation of any stage: without a CUDA device pislam_b200_create fails.
#include "../../include/pislam_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "brief_pattern.h"
#include "front_kernels.cuh"
#include "orb_kernels.cuh"
#include "match_kernels.cuh"
#include "pyramid_kernels.cuh"
#include "p2p_gather.cuh"

using namespace pb;

#define CK(call)                                  \
  do {                                            \
    cudaError_t e_ = (call);                      \
    if (e_ != cudaSuccess) return -(int)e_;       \
  } while (0)
#define CKS(call)                                 \
  do {                                            \
    int s_ = (call);                              \
    if (s_ != PISLAM_OK) return s_;               \
  } while (0)

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n, bool zero_new = false, cudaStream_t st = nullptr) {
    if (n <= cap) return cudaSuccess;
    if (p) {
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return e;
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    const size_t want = ((n + (n >> 2) + 4095) & ~(size_t)4095) + 256;  // slack: see load_word_bounded
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    if (zero_new) e = cudaMemsetAsync(p, 0, want, st);
    return e;
  }
  // for buffers used from several streams: growing waits for the whole device
  cudaError_t reserve_any(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) {
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) return e;
      cudaFree(p);
      p = nullptr;
      cap = 0;
    }
    const size_t want = ((n + (n >> 2) + 4095) & ~(size_t)4095) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

// ---- NCCL through dlopen: the library has no link-time dependency on it --------------------------
struct NcclUid {
  char internal[128];
};
typedef void* NcclComm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    GetUniqueId = reinterpret_cast<int (*)(NcclUid*)>(dlsym(lib, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUid, int)>(dlsym(lib, "ncclCommInitRank"));
    CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(lib, "ncclCommDestroy"));
    AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(
        dlsym(lib, "ncclAllGather"));
    return GetUniqueId && CommInitRank && CommDestroy && AllGather;
  }
};
NcclApi g_nccl;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// ---- rotated rBRIEF pattern (reference Brief.h:28-50) ---------------------------------------------
// theta = rot*pi/15 and its cosf/sinf are folded at compile time in the reference; their fp32 bit
// patterns are in brief_pattern.h.  Each product and sum is an individually rounded fp32 operation and
// the rounding is an exact half-away-from-zero roundf, then each coordinate is clamped to [-15, 15].
int rot_clamp(float v) {
  const int r = (int)roundf(v);
  return r < -15 ? -15 : (r > 15 ? 15 : r);
}
void build_brief_table(int8_t* t /* [30][256][4] */) {
  for (int rot = 0; rot < 30; ++rot) {
    float c, s;
    memcpy(&c, &PISLAM_BRIEF_COS_BITS[rot], 4);
    memcpy(&s, &PISLAM_BRIEF_SIN_BITS[rot], 4);
    for (int k = 0; k < 256; ++k)
      for (int e = 0; e < 2; ++e) {
        const float dx = (float)PISLAM_BRIEF_BASE[k][2 * e], dy = (float)PISLAM_BRIEF_BASE[k][2 * e + 1];
        volatile float cx = c * dx, sy = s * dy, sx = s * dx, cy = c * dy;
        volatile float rx = cx - sy, ry = sx + cy;
        t[(rot * 256 + k) * 4 + 2 * e] = (int8_t)rot_clamp(rx);
        t[(rot * 256 + k) * 4 + 2 * e + 1] = (int8_t)rot_clamp(ry);
      }
  }
}

constexpr int kTileSlot = 32;  // staging entries every tile owns (FrontParams::tile_slot)

struct Geometry {
  pislam_pyramid_desc d;
  int cap = 0;
  bool valid = false;
  int th = TH_BATCH;  // tile height the tables were built for
  DevBuf tiles;       // the geometry's tile table on the device
  unsigned long long last_use = 0;
  int ntiles = 0, npieces = 0;
  LevelInfo lv[MAX_LEVELS];
  int piece_base[MAX_LEVELS + 1];
  long long rows = 0;  // stack rows covered by the levels
};

bool same_desc(const pislam_pyramid_desc& a, const pislam_pyramid_desc& b) {
  if (a.vstep != b.vstep || a.nlevels != b.nlevels || a.border != b.border || a.log_bucket_size != b.log_bucket_size)
    return false;
  for (int l = 0; l < a.nlevels; ++l)
    if (a.widths[l] != b.widths[l] || a.heights[l] != b.heights[l] || a.row_offsets[l] != b.row_offsets[l])
      return false;
  return true;
}

}  // namespace

struct pislam_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long launches = 0;
  int num_sms = 148;
  int orb_tma_ctas_per_sm = 0;  // occupancy of orb_tma_kernel (queried once)
  int front_resident_ctas = 0;  // CTAs of front_tile_kernel<TMA, 64 rows> resident on the whole GPU (L2 prefetch distance)
  bool use_tma = true;
  EncodeTiledFn encode = nullptr;
  uint32_t* d_brief = nullptr;    // pattern offsets for the 36-byte patch pitch
  uint32_t* d_brief48 = nullptr;  // ... for the 48-byte pitch of the TMA patch buffers
  uint8_t* d_zeros = nullptr;     // 16 KB of zeros: source of the score tile's bulk clear (front_tile_kernel, MODE_TMA)
  // literal stages
  DevBuf img, out, kp, desc, aux0, aux1;
  // fused path
  // small LRU of tile / piece tables keyed by (tile height, full descriptor): a caller alternating a few
  // geometries (levels of different cameras, literal fastExtract per level) re-uses them without any upload
  static const int kGeoCache = 8;
  Geometry geos[kGeoCache];
  Geometry* geo = nullptr;  // geometry of the current call
  unsigned long long geo_clock = 0;
  long geo_uploads = 0;
  DevBuf stage, stage_count, piece_cnt, piece_off, stacks, records, hdr;
  // pyramid builder (N1): tap tables per geometry (LRU), staging for base frames and the blur passes
  struct PyrTables {
    pislam_pyramid_desc d;
    bool valid = false;
    unsigned long long last_use = 0;
    DevBuf taps;
    PyramidParams proto;  // everything but the pointers / strides of a call
  };
  static const int kPyrCache = 4;
  PyrTables pyr[kPyrCache];
  DevBuf bases, blur_tmp, blur_out;
  // double-buffered device staging of the host-source pipeline (pipeline_enqueue)
  struct Staging {
    DevBuf stacks, bases, records;
    cudaEvent_t src_free = nullptr, rec_free = nullptr;
  } staging[2];
  int stage_turn = 0;
  // multi-GPU: the all-gather runs on its own stream so that it overlaps the next batch's kernels
  NcclComm comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_extract = nullptr;
  // sub-batches: the descriptor kernel of sub-batch k runs on its own (high-priority) stream under the detector of
  // sub-batch k+1, while the sub-batch's stacks are still in L2
  cudaStream_t orb_stream = nullptr, flag_stream = nullptr;
  cudaEvent_t ev_sub = nullptr, ev_orb = nullptr, ev_ready = nullptr;
  int sub_frames = 0;  // frames per sub-batch (0: the whole call in one piece)
  static const int kGatherSlots = 4;
  struct GatherSlot {
    void* ptr = nullptr;  // gather buffer last used with this event
    cudaEvent_t done = nullptr;
  } gathers[kGatherSlots];  // the last few gather buffers (round robin): a caller may rotate up to that many
  int gather_turn = 0;
  // size-aware exchange over NVLink peer mappings (p2p_gather.cuh): flag block + registered buffers ("windows")
  struct IpcMap {
    cudaIpcMemHandle_t handle;
    int rank;
    void* base;
  };
  struct PeerWindow {
    uint8_t* local = nullptr;
    size_t bytes = 0;
    uint8_t* peer[P2P_MAX_RANKS] = {};
    uint32_t epoch = 0;
  };
  bool p2p_ok = false;
  int push_ctas = 32;  // CTAs of p2p_push_kernel (it runs under the next batch's detector: few CTAs, see p2p_gather.cuh)
  uint32_t* d_flags = nullptr;  // this rank's flag block (written by the peers)
  uint32_t* peer_flags[P2P_MAX_RANKS] = {};
  int* d_p2p_timeout = nullptr;
  std::vector<IpcMap> ipc_maps;
  std::vector<PeerWindow> windows;
  long p2p_bytes_pushed = 0;  // host-side estimate is impossible (sizes live on the device): see pislam_b200_comm_stats
  // host-buffer batches are pipelined in chunks: H2D (copy_in) | kernels (stream) | D2H (copy_out)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  static const int kMaxChunks = 32;
  cudaEvent_t ev_in[kMaxChunks] = {}, ev_k[kMaxChunks] = {};
  int pipeline_chunks = 4;
  bool pipeline_chunks_forced = false;  // PISLAM_B200_CHUNKS given: use it for every host source
  size_t l2_chunk_bytes = 48u << 20;  // device-resident base frames: stack bytes built per chunk (stays in the 126 MB L2)
  bool trim_copies = true;
  void* h_rec = nullptr;  // pinned staging for single-frame results
  size_t h_rec_cap = 0;
  // optional per-kernel timing (bench.py roofline): events around each fused-path launch
  bool profiling = false;
  struct ProfSpan {
    cudaEvent_t a, b;
    int which;  // 0 front_tile_kernel, 1 order_kernel, 2 orb_kernel, 3 pyramid_build_kernel, 4 record exchange
  };
  std::vector<ProfSpan> spans;
};

namespace {

int check_launch(pislam_ctx* c) {
  ++c->launches;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? PISLAM_OK : -(int)e;
}

// RAII-less span helpers: no-ops unless profiling is on
void span_begin(pislam_ctx* c, int which, cudaStream_t st = nullptr) {
  if (!c->profiling) return;
  pislam_ctx::ProfSpan s;
  s.which = which;
  if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
  cudaEventRecord(s.a, st ? st : c->stream);
  c->spans.push_back(s);
}
void span_end(pislam_ctx* c, cudaStream_t st = nullptr) {
  if (!c->profiling || c->spans.empty()) return;
  cudaEventRecord(c->spans.back().b, st ? st : c->stream);
}

int set_device(pislam_ctx* c) {
  CK(cudaSetDevice(c->device));
  return PISLAM_OK;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// The reference addresses img[y][x] linearly on a uint8_t(*)[vstep]; a window is readable iff its first and last
// byte lie inside the rows the caller handed over.  lo / hi: window extent around (x, y), e.g. -15 / +15.
bool window_inside(long long x, long long y, int lo, int hi, int vstep, int rows) {
  const long long first = (y + lo) * vstep + x + lo, last = (y + hi) * vstep + x + hi;
  return first >= 0 && last < (long long)rows * vstep;
}
bool keypoints_inside(const uint32_t* kp, int n, int vstep, int rows) {
  for (int i = 0; i < n; ++i)
    if (!window_inside(decode_x(kp[i]), decode_y(kp[i]), -15, 15, vstep, rows)) return false;
  return true;
}

int validate_desc(const pislam_pyramid_desc* d, int cap, bool fused) {
  if (!d || d->vstep <= 0 || d->nlevels < 1 || d->nlevels > PISLAM_MAX_LEVELS || cap < 1) return PISLAM_ERR_INVALID_ARG;
  if (d->words < 0 || d->words > 8 || d->bucket_limit < 1 || d->log_bucket_size < 0) return PISLAM_ERR_INVALID_ARG;
  for (int l = 0; l < d->nlevels; ++l)
    if (d->widths[l] < 1 || d->heights[l] < 1 || d->widths[l] > d->vstep || d->row_offsets[l] < 0)
      return PISLAM_ERR_INVALID_ARG;
  // 12-bit keypoint coordinates (Util.h:27-29); the fused tiles need >= 5 border pixels so that no ring,
  // Harris window or 3x3 neighbourhood ever leaves the level's own columns, and buckets must tile 128x32.
  for (int l = 0; l < d->nlevels; ++l)
    if (d->widths[l] > 4096 || d->heights[l] > 4096) return PISLAM_ERR_UNSUPPORTED;
  if (fused && d->border < 5) return PISLAM_ERR_UNSUPPORTED;
  if (d->words > 0 && d->border < 16) return PISLAM_ERR_UNSUPPORTED;
  if (d->log_bucket_size > 5) return PISLAM_ERR_UNSUPPORTED;
  return PISLAM_OK;
}

// Tile height for a call: 64-row tiles unless they would leave most of the SMs' CTA slots empty
int pick_tile_height(const pislam_ctx* c, const pislam_pyramid_desc* d, int nframes) {
  long long tiles64 = 0;
  for (int l = 0; l < d->nlevels; ++l) {
    const int iw = d->widths[l] - 2 * d->border, ih = d->heights[l] - 2 * d->border;
    if (iw > 0 && ih > 0) tiles64 += (long long)((iw + TW - 1) / TW) * ((ih + TH_BATCH - 1) / TH_BATCH);
  }
  static const int forced = getenv("PISLAM_B200_TILE_H") ? atoi(getenv("PISLAM_B200_TILE_H")) : 0;  // timing experiments
  if (forced == TH_SMALL || forced == TH_BATCH) return forced;
  return tiles64 * nframes < 4LL * c->num_sms ? TH_SMALL : TH_BATCH;
}

// Tile + piece tables of a geometry; cached (LRU) per tile height and descriptor.  A hit costs no upload and no
// synchronisation; a miss uploads the tile table stream-ordered (the pageable source is consumed before
// cudaMemcpyAsync returns, and the evicted table's earlier launches precede the copy on the stream).
int prepare_geometry(pislam_ctx* c, const pislam_pyramid_desc* d, int cap, int nframes) {
  const int TH = pick_tile_height(c, d, nframes);
  Geometry* hit = nullptr;
  Geometry* victim = nullptr;  // a free entry, else the least recently used one
  for (Geometry& e : c->geos) {
    if (e.valid && e.th == TH && same_desc(e.d, *d)) {
      hit = &e;
      break;
    }
    if (!victim || (victim->valid && (!e.valid || e.last_use < victim->last_use))) victim = &e;
  }
  if (hit) {
    hit->d = *d;  // thresholds / bucket_limit / words may differ: they are launch parameters, not table content
    hit->cap = cap;
    hit->last_use = ++c->geo_clock;
    c->geo = hit;
    return PISLAM_OK;
  }
  Geometry& g = *victim;
  g.valid = false;
  g.th = TH;
  const int B = d->border;
  const int cb_log = d->log_bucket_size > 0 ? d->log_bucket_size - 1 : 0;
  const int nseg = (TH / 2) >> cb_log;  // segments per tile: cell rows, or bucket bands
  std::vector<TileDesc> tiles;
  int piece = 0;
  long long rows = 0;
  for (int l = 0; l < d->nlevels; ++l) {
    LevelInfo& L = g.lv[l];
    L.w = d->widths[l];
    L.h = d->heights[l];
    L.row_off = d->row_offsets[l];
    const int iw = L.w - 2 * B, ih = L.h - 2 * B;
    L.xend = B + 16 * ((iw + 15) / 16);
    L.spill = L.xend > d->vstep + B - 1;
    L.ntile_cols = (iw > 0 && ih > 0) ? (iw + TW - 1) / TW : 0;
    const int ntile_rows = (iw > 0 && ih > 0) ? (ih + TH - 1) / TH : 0;
    L.piece_base = piece;
    g.piece_base[l] = piece;
    for (int tr = 0; tr < ntile_rows; ++tr)
      for (int tc = 0; tc < L.ntile_cols; ++tc) {
        TileDesc t;
        t.level = l;
        t.x0 = B + tc * TW;
        t.y0 = B + tr * TH;
        t.piece0 = piece + tr * nseg * L.ntile_cols + tc;
        tiles.push_back(t);
      }
    piece += ntile_rows * nseg * L.ntile_cols;
    rows = std::max<long long>(rows, (long long)L.row_off + L.h);
  }
  g.piece_base[d->nlevels] = piece;
  g.ntiles = (int)tiles.size();
  g.npieces = piece;
  g.rows = rows;
  g.d = *d;
  g.cap = cap;
  if (g.ntiles > 0) {
    CK(g.tiles.reserve(tiles.size() * sizeof(TileDesc), false, c->stream));
    CK(cudaMemcpyAsync(g.tiles.p, tiles.data(), tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, c->stream));
    ++c->geo_uploads;
  }
  g.valid = true;
  g.last_use = ++c->geo_clock;
  c->geo = &g;
  return PISLAM_OK;
}

int make_tensor_map(pislam_ctx* c, CUtensorMap* m, const uint8_t* base, int vstep, long long rows, int nframes,
                    size_t frame_stride, int box_w, int box_h) {
  const cuuint64_t dims[3] = {(cuuint64_t)vstep, (cuuint64_t)rows, (cuuint64_t)nframes};
  const cuuint64_t strides[2] = {(cuuint64_t)vstep, (cuuint64_t)(nframes > 1 ? frame_stride : (size_t)rows * vstep)};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PISLAM_OK : PISLAM_ERR_UNSUPPORTED;
}

void fill_front_params(const pislam_ctx* c, FrontParams& P, const uint8_t* src, size_t frame_stride) {
  const Geometry& g = *c->geo;
  memset(&P, 0, sizeof(P));
  P.src = src;
  P.frame_stride = frame_stride;
  P.frame_bytes = std::min<unsigned long long>(frame_stride, (unsigned long long)g.rows * g.d.vstep);
  P.vstep = g.d.vstep;
  P.nlevels = g.d.nlevels;
  for (int l = 0; l < g.d.nlevels; ++l) P.lv[l] = g.lv[l];
  P.border = g.d.border;
  P.fast_t = g.d.fast_threshold;
  P.harris_t = g.d.harris_threshold;
  P.log_bucket = g.d.log_bucket_size;
  P.bucket_limit = g.d.bucket_limit;
  P.cap = g.cap;
  P.npieces = g.npieces;
  P.zeros = c->d_zeros;
  P.prefetch_dist = c->front_resident_ctas;
  P.tile_slot = kTileSlot;
  P.stage_cap = g.ntiles * kTileSlot + g.cap;
  P.tiles = g.tiles.as<TileDesc>();
  P.stage = c->stage.as<uint32_t>();
  P.stage_count = c->stage_count.as<uint32_t>();
  P.piece_cnt = c->piece_cnt.as<uint32_t>();
  P.piece_off = c->piece_off.as<uint32_t>();
}

int reserve_front_scratch(pislam_ctx* c, int nframes) {
  const Geometry& g = *c->geo;
  CK(c->stage.reserve((size_t)nframes * ((size_t)g.ntiles * kTileSlot + g.cap) * 4, false, c->stream));
  CK(c->stage_count.reserve((size_t)nframes * 4, true, c->stream));
  CK(c->piece_cnt.reserve((size_t)nframes * std::max(1, g.npieces) * 4, false, c->stream));
  CK(c->piece_off.reserve((size_t)nframes * std::max(1, g.npieces) * 4, false, c->stream));
  return PISLAM_OK;
}

int launch_order(pislam_ctx* c, int nframes, int words, uint8_t* d_records) {
  const Geometry& g = *c->geo;
  OrderParams O;
  memset(&O, 0, sizeof(O));
  O.nlevels = g.d.nlevels;
  O.npieces = g.npieces;
  O.cap = g.cap;
  O.stage_cap = g.ntiles * kTileSlot + g.cap;
  O.over_cap = g.cap;
  for (int l = 0; l <= g.d.nlevels; ++l) O.piece_base[l] = g.piece_base[l];
  O.stage = c->stage.as<uint32_t>();
  O.stage_count = c->stage_count.as<uint32_t>();
  O.piece_cnt = c->piece_cnt.as<uint32_t>();
  O.piece_off = c->piece_off.as<uint32_t>();
  O.records = d_records;
  O.rec = RecordLayout::make(g.cap, words);
  span_begin(c, 1);
  order_kernel<<<nframes, ORDER_NT, 0, c->stream>>>(O);
  span_end(c);
  return check_launch(c);
}

// all device pointers; enqueue only
int enqueue_extract_one(pislam_ctx* c, const uint8_t* d_stacks, size_t frame_stride, int nframes,
                        const pislam_pyramid_desc* d, int cap, uint8_t* d_records, cudaStream_t orb_st) {
  if (nframes < 0 || nframes > 65535) return PISLAM_ERR_INVALID_ARG;
  CKS(validate_desc(d, cap, true));
  if (nframes == 0) return PISLAM_OK;  // an empty shard: nothing to launch
  if (!d_stacks || !d_records) return PISLAM_ERR_INVALID_ARG;
  CKS(prepare_geometry(c, d, cap, nframes));
  const Geometry& g = *c->geo;
  if (nframes > 1 && frame_stride < (size_t)g.rows * d->vstep) return PISLAM_ERR_INVALID_ARG;
  if (nframes == 1) frame_stride = (size_t)g.rows * d->vstep;
  CKS(reserve_front_scratch(c, nframes));
  FrontParams P;
  fill_front_params(c, P, d_stacks, frame_stride);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  const bool tma_ok = c->use_tma && c->encode && d->vstep % 16 == 0 && frame_stride % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(d_stacks) & 15) == 0 && d->vstep >= 16 && d->border % 4 == 0;
  if (g.ntiles > 0) {
    const dim3 grid(g.ntiles, nframes);
    span_begin(c, 0);
    // occupancy experiments: extra (unused) dynamic shared memory per CTA
    static const int pad = getenv("PISLAM_B200_DEBUG_SMEM_PAD") ? atoi(getenv("PISLAM_B200_DEBUG_SMEM_PAD")) : 0;
    const int box_h = g.th == TH_BATCH ? TileCfg<TH_BATCH>::PIX_H : TileCfg<TH_SMALL>::PIX_H;
    const bool tma = tma_ok && make_tensor_map(c, &tmap, d_stacks, d->vstep, g.rows, nframes, frame_stride, PIX_W, box_h) == PISLAM_OK;
    if (tma && c->front_resident_ctas == 0) {
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, front_tile_kernel<MODE_TMA, TH_BATCH>, NT, 0) != cudaSuccess || occ < 1)
        occ = 1;
      c->front_resident_ctas = occ * c->num_sms;
      if (const char* pd = getenv("PISLAM_B200_PREFETCH_DIST")) c->front_resident_ctas = atoi(pd);
      P.prefetch_dist = c->front_resident_ctas;
    }
    if (g.th == TH_BATCH) {
      if (tma) front_tile_kernel<MODE_TMA, TH_BATCH><<<grid, NT, pad, c->stream>>>(P, tmap);
      else front_tile_kernel<MODE_LDG, TH_BATCH><<<grid, NT, 0, c->stream>>>(P, tmap);
    } else {
      if (tma) front_tile_kernel<MODE_TMA, TH_SMALL><<<grid, NT, pad, c->stream>>>(P, tmap);
      else front_tile_kernel<MODE_LDG, TH_SMALL><<<grid, NT, 0, c->stream>>>(P, tmap);
    }
    span_end(c);
    CKS(check_launch(c));
  }
  CKS(launch_order(c, nframes, d->words, d_records));
  if (d->words > 0) {
    OrbParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.img = d_stacks;
    Q.frame_stride = frame_stride;
    Q.frame_bytes = (size_t)g.rows * d->vstep;
    Q.vstep = d->vstep;
    Q.words = d->words;
    Q.records = d_records;
    Q.rec = RecordLayout::make(cap, d->words);
    for (int l = 0; l < d->nlevels; ++l) Q.row_off[l] = d->row_offsets[l];
    Q.brief = c->d_brief;
    Q.brief48 = c->d_brief48;
    Q.nframes = nframes;
    Q.cap = cap;
    const long long want = ((long long)nframes * cap + ORB_WARPS - 1) / ORB_WARPS;
    CUtensorMap omap;
    memset(&omap, 0, sizeof(omap));
    // the patch prefetch needs 16-byte-aligned rows and descriptors' border >= 16 keeps every patch column
    // inside its level row (no wrap into the next row, which a box load would zero-fill instead)
    const bool orb_tma = tma_ok && d->vstep >= OT_W &&
                         make_tensor_map(c, &omap, d_stacks, d->vstep, g.rows, nframes, frame_stride, OT_W, OT_H) == PISLAM_OK;
    if (orb_st != c->stream) {  // the keypoints of this sub-batch are final once order_kernel has run
      CK(cudaEventRecord(c->ev_sub, c->stream));
      CK(cudaStreamWaitEvent(orb_st, c->ev_sub, 0));
    }
    span_begin(c, 2, orb_st);
    if (orb_tma) {
      // every CTA of the grid is resident (the kernel splits the keypoints evenly over gridDim.x CTAs)
      if (c->orb_tma_ctas_per_sm == 0) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, orb_tma_kernel, ORB_WARPS * 32, 0) != cudaSuccess || occ < 1)
          occ = 1;
        c->orb_tma_ctas_per_sm = occ;
      }
      for (int f0 = 0; f0 < nframes; f0 += ORB_MAX_FRAMES) {
        OrbParams R = Q;
        R.nframes = std::min(ORB_MAX_FRAMES, nframes - f0);
        R.frame0 = f0;
        R.records = d_records + (size_t)f0 * Q.rec.bytes;
        const long long slots = ((long long)R.nframes * cap + ORB_WARPS - 1) / ORB_WARPS;
        const dim3 grid((unsigned)std::min<long long>(slots, (long long)c->num_sms * c->orb_tma_ctas_per_sm));
        orb_tma_kernel<<<grid, ORB_WARPS * 32, 0, orb_st>>>(R, omap);
        if (f0 + ORB_MAX_FRAMES < nframes) ++c->launches;
      }
    } else {
      const dim3 grid((unsigned)std::min<long long>(want, (long long)c->num_sms * 8));
      orb_kernel<1><<<grid, ORB_WARPS * 32, 0, orb_st>>>(Q);
    }
    span_end(c, orb_st);
    CKS(check_launch(c));
  }
  return PISLAM_OK;
}

// all device pointers; enqueue only.  With sub_frames > 0 the call is cut into sub-batches: detector + ordering of
// sub-batch k+1 on the ctx stream, the descriptor kernel of sub-batch k on orb_stream (it reads the sub-batch's
// stacks while they are still in L2 and fills the detector's idle issue slots); the ctx stream is re-joined at the end.
int enqueue_extract(pislam_ctx* c, const uint8_t* d_stacks, size_t frame_stride, int nframes,
                    const pislam_pyramid_desc* d, int cap, uint8_t* d_records) {
  const int sub = c->sub_frames;
  if (sub <= 0 || nframes <= sub || !d || d->words <= 0)
    return enqueue_extract_one(c, d_stacks, frame_stride, nframes, d, cap, d_records, c->stream);
  const size_t rec_bytes = RecordLayout::make(cap, d->words).bytes;
  // earlier work on the ctx stream (e.g. a consumer of the records) precedes the descriptor kernels too
  for (int f0 = 0; f0 < nframes; f0 += sub) {
    const int nf = std::min(sub, nframes - f0);
    CKS(enqueue_extract_one(c, d_stacks + (size_t)f0 * frame_stride, frame_stride, nf, d, cap,
                            d_records + (size_t)f0 * rec_bytes, c->orb_stream));
  }
  CK(cudaEventRecord(c->ev_orb, c->orb_stream));
  CK(cudaStreamWaitEvent(c->stream, c->ev_orb, 0));
  return PISLAM_OK;
}

// ---- pyramid builder ----------------------------------------------------------------
// taps of oracle/pislam_oracle.h:po_resize_bilinear for one axis
void resize_taps(int n_out, int n_in, std::vector<uint32_t>& t) {
  for (int i = 0; i < n_out; ++i) {
    long long pos = ((long long)(2 * i + 1) * n_in * 128) / n_out - 128;
    const long long hi = (long long)(n_in - 1) * 256;
    pos = std::min(std::max(pos, 0LL), hi);
    const int i0 = (int)(pos >> 8), f = (int)(pos & 255), i1 = std::min(i0 + 1, n_in - 1);
    t.push_back(pack_tap(i0, i1, f));
  }
  while (t.size() % 4) t.push_back(t.back());  // 16-byte loads of four taps never leave the table
}

int prepare_pyramid_tables(pislam_ctx* c, const pislam_pyramid_desc* d, pislam_ctx::PyrTables** out) {
  pislam_ctx::PyrTables* victim = nullptr;
  for (pislam_ctx::PyrTables& e : c->pyr) {
    if (e.valid && same_desc(e.d, *d)) {
      e.last_use = ++c->geo_clock;
      *out = &e;
      return PISLAM_OK;
    }
    if (!victim || (victim->valid && (!e.valid || e.last_use < victim->last_use))) victim = &e;
  }
  pislam_ctx::PyrTables& e = *victim;
  e.valid = false;
  PyramidParams& P = e.proto;
  memset(&P, 0, sizeof(P));
  P.W = d->widths[0];
  P.H = d->heights[0];
  P.vstep = d->vstep;
  P.nlevels = d->nlevels;
  std::vector<uint32_t> xt, yt;
  int tile = 0;
  for (int l = 0; l < d->nlevels; ++l) {
    P.w[l] = d->widths[l];
    P.h[l] = d->heights[l];
    P.row_off[l] = d->row_offsets[l];
    P.tiles_x[l] = (d->widths[l] + PYR_TW - 1) / PYR_TW;
    P.tile0[l] = tile;
    tile += P.tiles_x[l] * ((d->heights[l] + PYR_TH - 1) / PYR_TH);
    P.xt0[l] = (int)xt.size();
    resize_taps(d->widths[l], P.W, xt);
    P.yt0[l] = (int)yt.size();
    resize_taps(d->heights[l], P.H, yt);
  }
  P.tile0[d->nlevels] = tile;
  CK(e.taps.reserve((xt.size() + yt.size()) * 4, false, c->stream));
  CK(cudaMemcpyAsync(e.taps.p, xt.data(), xt.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(e.taps.as<uint32_t>() + xt.size(), yt.data(), yt.size() * 4, cudaMemcpyHostToDevice, c->stream));
  P.xtaps = e.taps.as<uint32_t>();
  P.ytaps = e.taps.as<uint32_t>() + xt.size();
  e.d = *d;
  e.valid = true;
  e.last_use = ++c->geo_clock;
  ++c->geo_uploads;
  *out = &e;
  return PISLAM_OK;
}

int validate_frames(const pislam_pyramid_desc* d, int frame_pitch, size_t frame_stride, int nframes, int blur) {
  if (!d || nframes < 0 || d->nlevels < 1) return PISLAM_ERR_INVALID_ARG;
  const int W = d->widths[0], H = d->heights[0];
  if (frame_pitch < W || (nframes > 1 && frame_stride < (size_t)frame_pitch * (H - 1) + W)) return PISLAM_ERR_INVALID_ARG;
  if (blur && (W < 3 || H < 3)) return PISLAM_ERR_INVALID_ARG;  // reflect-101 borders of the 5x5 blur
  for (int l = 0; l < d->nlevels; ++l)
    if (d->widths[l] > 32767 || d->heights[l] > 32767) return PISLAM_ERR_UNSUPPORTED;  // 15-bit tap indices
  return PISLAM_OK;
}

// device pointers; enqueue only: [blur passes] + one pyramid_build_kernel launch for the whole batch
int enqueue_pyramid(pislam_ctx* c, const uint8_t* d_frames, int frame_pitch, size_t frame_stride, int nframes, int blur,
                    const pislam_pyramid_desc* d, uint8_t* d_stacks, size_t stack_stride, bool zero_padding) {
  if (nframes == 0) return PISLAM_OK;
  if (nframes > 65535) return PISLAM_ERR_INVALID_ARG;
  pislam_ctx::PyrTables* T = nullptr;
  CKS(prepare_pyramid_tables(c, d, &T));
  PyramidParams P = T->proto;
  P.base = d_frames;
  P.base_pitch = frame_pitch;
  P.base_stride = frame_stride;
  if (blur) {
    // Gaussian.h:48-72 semantics (vertical pass, then horizontal) on every base frame; the levels are resized from
    // the blurred frame (pitch = width rounded up to 4)
    const int W = P.W, H = P.H, tpitch = (W + 3) & ~3;
    const size_t fbytes = (size_t)tpitch * H;
    CK(c->blur_tmp.reserve(fbytes * nframes + 16, false, c->stream));
    CK(c->blur_out.reserve(fbytes * nframes + 16, false, c->stream));
    const dim3 grid(((W + 3) / 4 + 127) / 128, H, nframes);
    gaussian_vertical_batch_kernel<<<grid, 128, 0, c->stream>>>(d_frames, frame_stride, c->blur_tmp.as<uint8_t>(), fbytes,
                                                                frame_pitch, tpitch, W, H);
    CKS(check_launch(c));
    gaussian_horizontal_batch_kernel<<<grid, 128, 0, c->stream>>>(c->blur_tmp.as<uint8_t>(), fbytes,
                                                                  c->blur_out.as<uint8_t>(), fbytes, tpitch, tpitch, W, H);
    CKS(check_launch(c));
    P.base = c->blur_out.as<uint8_t>();
    P.base_pitch = tpitch;
    P.base_stride = fbytes;
  }
  P.stacks = d_stacks;
  P.stack_stride = stack_stride;
  P.zero_padding = zero_padding ? 1 : 0;
  span_begin(c, 3);
  pyramid_build_kernel<<<dim3(P.tile0[P.nlevels], nframes), dim3(32, PYR_BLOCK_ROWS), 0, c->stream>>>(P);
  span_end(c);
  return check_launch(c);
}

// H2D of base frames [f0, f0+nf) into a tight device layout (pitch dp, stride dp*H)
int copy_bases_h2d(cudaStream_t st, const uint8_t* host, int pitch, size_t stride, uint8_t* dev, int dp, int W, int H,
                   int f0, int nf) {
  const size_t dstride = (size_t)dp * H;
  if (pitch == dp && stride == dstride) {
    CK(cudaMemcpyAsync(dev + f0 * dstride, host + f0 * stride, dstride * nf, cudaMemcpyHostToDevice, st));
  } else if (stride == (size_t)pitch * H) {  // frames are consecutive rows of one tall image
    CK(cudaMemcpy2DAsync(dev + f0 * dstride, dp, host + f0 * stride, pitch, W, (size_t)H * nf, cudaMemcpyHostToDevice, st));
  } else {
    for (int f = f0; f < f0 + nf; ++f)
      CK(cudaMemcpy2DAsync(dev + f * dstride, dp, host + f * stride, pitch, W, H, cudaMemcpyHostToDevice, st));
  }
  return PISLAM_OK;
}

// H2D of frames [f0, f0+nf) of a host batch into the same layout on the device.  Only the level pixels travel
// (one pitched 3-D copy per level): with border >= 5 the fused kernels never read a column >= w of a level,
// and the padding right of each level is ~1/3 of a stacked pyramid.
int copy_frames_h2d(pislam_ctx* c, cudaStream_t st, const uint8_t* host, uint8_t* dev, size_t frame_stride,
                    size_t frame_bytes, int f0, int nf, const pislam_pyramid_desc* d) {
  if (c->trim_copies && frame_stride % (size_t)d->vstep == 0) {
    for (int l = 0; l < d->nlevels; ++l) {
      cudaMemcpy3DParms p;
      memset(&p, 0, sizeof(p));
      p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(host), d->vstep, d->vstep, frame_stride / d->vstep);
      p.dstPtr = make_cudaPitchedPtr(dev, d->vstep, d->vstep, frame_stride / d->vstep);
      p.srcPos = make_cudaPos(0, d->row_offsets[l], f0);
      p.dstPos = p.srcPos;
      p.extent = make_cudaExtent(d->widths[l], d->heights[l], nf);
      p.kind = cudaMemcpyHostToDevice;
      CK(cudaMemcpy3DAsync(&p, st));
    }
    return PISLAM_OK;
  }
  const size_t bytes = (size_t)(nf - 1) * frame_stride + frame_bytes;
  CK(cudaMemcpyAsync(dev + (size_t)f0 * frame_stride, host + (size_t)f0 * frame_stride, bytes, cudaMemcpyHostToDevice, st));
  return PISLAM_OK;
}

// ---- size-aware exchange over peer mappings (p2p_gather.cuh) ---------------------------------------------------
typedef CUresult (*MemGetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

// all-gather of a small host blob through the communicator (setup paths only): all[r*bytes ..] = rank r's blob
int comm_allgather_host(pislam_ctx* c, const void* mine, size_t bytes, std::vector<uint8_t>& all) {
  all.assign(bytes * (size_t)c->nranks, 0);
  CK(c->hdr.reserve(bytes * (size_t)c->nranks, false, c->stream));
  uint8_t* d = c->hdr.as<uint8_t>();
  CK(cudaMemcpyAsync(d + bytes * (size_t)c->rank, mine, bytes, cudaMemcpyHostToDevice, c->stream));
  const int r = g_nccl.AllGather(d + bytes * (size_t)c->rank, d, bytes, /*ncclInt8*/ 0, c->comm, c->stream);
  if (r != 0) return PISLAM_ERR_NCCL - r;
  CK(cudaMemcpyAsync(all.data(), d, all.size(), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

// maps a peer's allocation (once per handle and rank)
int ipc_open(pislam_ctx* c, const cudaIpcMemHandle_t& h, int rank, void** base) {
  for (const pislam_ctx::IpcMap& m : c->ipc_maps)
    if (m.rank == rank && memcmp(&m.handle, &h, sizeof(h)) == 0) {
      *base = m.base;
      return PISLAM_OK;
    }
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return -(int)e;
  }
  pislam_ctx::IpcMap m;
  m.handle = h;
  m.rank = rank;
  m.base = p;
  c->ipc_maps.push_back(m);
  *base = p;
  return PISLAM_OK;
}

struct IpcBlob {
  cudaIpcMemHandle_t handle;
  unsigned long long offset, bytes;
  int ok;
};

// base allocation + offset of a device pointer (IPC handles describe whole allocations)
bool ipc_describe(const void* ptr, size_t bytes, IpcBlob* b) {
  memset(b, 0, sizeof(*b));
  static MemGetAddressRangeFn get_range = nullptr;
  if (!get_range) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return false;
    }
    get_range = reinterpret_cast<MemGetAddressRangeFn>(fn);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) return false;
  if (cudaIpcGetMemHandle(&b->handle, reinterpret_cast<void*>((uintptr_t)base)) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  b->offset = (unsigned long long)((uintptr_t)ptr - (uintptr_t)base);
  b->bytes = bytes;
  b->ok = 1;
  return true;
}

// the flag block every peer writes its READY / DONE epochs into; collective (called from comm_init)
int p2p_setup(pislam_ctx* c) {
  c->p2p_ok = false;
  if (c->nranks < 2 || c->nranks > P2P_MAX_RANKS) return PISLAM_OK;
  const size_t fbytes = (size_t)P2P_MAX_WINDOWS * 2 * P2P_MAX_RANKS * 4;
  IpcBlob mine;
  memset(&mine, 0, sizeof(mine));
  if (getenv("PISLAM_B200_NO_P2P") == nullptr && cudaMalloc(&c->d_flags, fbytes) == cudaSuccess &&
      cudaMalloc(&c->d_p2p_timeout, 4) == cudaSuccess) {
    cudaMemset(c->d_flags, 0, fbytes);
    cudaMemset(c->d_p2p_timeout, 0, 4);
    ipc_describe(c->d_flags, fbytes, &mine);
  } else {
    cudaGetLastError();
  }
  std::vector<uint8_t> all;
  CKS(comm_allgather_host(c, &mine, sizeof(mine), all));
  bool ok = true;
  for (int r = 0; r < c->nranks; ++r) ok = ok && reinterpret_cast<const IpcBlob*>(all.data() + r * sizeof(IpcBlob))->ok;
  int opened = ok ? 1 : 0;
  if (ok)
    for (int r = 0; r < c->nranks && opened; ++r) {
      const IpcBlob* b = reinterpret_cast<const IpcBlob*>(all.data() + r * sizeof(IpcBlob));
      if (r == c->rank) {
        c->peer_flags[r] = c->d_flags;
        continue;
      }
      void* base = nullptr;
      if (ipc_open(c, b->handle, r, &base) != PISLAM_OK) opened = 0;
      else c->peer_flags[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(base) + b->offset);
    }
  // every rank must agree, or some would wait in NCCL and others on flags
  CKS(comm_allgather_host(c, &opened, sizeof(opened), all));
  for (int r = 0; r < c->nranks; ++r) opened = opened && *reinterpret_cast<const int*>(all.data() + r * sizeof(int));
  c->p2p_ok = opened != 0;
  return PISLAM_OK;
}

void p2p_teardown(pislam_ctx* c) {
  for (const pislam_ctx::IpcMap& m : c->ipc_maps) cudaIpcCloseMemHandle(m.base);
  c->ipc_maps.clear();
  c->windows.clear();
  if (c->d_flags) cudaFree(c->d_flags);
  if (c->d_p2p_timeout) cudaFree(c->d_p2p_timeout);
  c->d_flags = nullptr;
  c->d_p2p_timeout = nullptr;
  c->p2p_ok = false;
  cudaGetLastError();
}

// after a sync: PISLAM_ERR_CAPACITY if any frame found more keypoints than it stored
int check_capacity(pislam_ctx* c, const uint8_t* d_records, int nframes, size_t rec_bytes) {
  std::vector<uint32_t> h((size_t)nframes * 2);
  CK(cudaMemcpy2DAsync(h.data(), 8, d_records, rec_bytes, 8, nframes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  for (int f = 0; f < nframes; ++f)
    if (h[2 * f + 1] > h[2 * f]) return PISLAM_ERR_CAPACITY;
  return PISLAM_OK;
}

}  // namespace

// ======================================================================================================
extern "C" {

int pislam_b200_create(pislam_ctx** out, int device, size_t max_stack_bytes, int max_keypoints) {
  if (!out) return PISLAM_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return PISLAM_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return PISLAM_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
  pislam_ctx* c = new (std::nothrow) pislam_ctx();
  if (!c) return PISLAM_ERR_INVALID_ARG;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  int st = PISLAM_OK;
  do {
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking)) != cudaSuccess) { st = -(int)e; break; }
    {
      // the all-gather of batch k runs under the kernels of batch k+1, whose grids keep every CTA slot busy:
      // the gather's CTAs have to win the slots that free up, so its stream gets the highest priority
      int prio_lo = 0, prio_hi = 0;
      if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) prio_hi = 0;
      if ((e = cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) { st = -(int)e; break; }
    }
    if ((e = cudaEventCreateWithFlags(&c->ev_extract, cudaEventDisableTiming)) != cudaSuccess) { st = -(int)e; break; }
    {
      int prio_lo = 0, prio_hi = 0;
      if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) prio_hi = 0;
      if ((e = cudaStreamCreateWithPriority(&c->orb_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) { st = -(int)e; break; }
      if ((e = cudaStreamCreateWithPriority(&c->flag_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) { st = -(int)e; break; }
      if ((e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming)) != cudaSuccess) { st = -(int)e; break; }
    }
    for (int i = 0; i < 2 && st == PISLAM_OK; ++i) {
      if ((e = cudaEventCreateWithFlags(&c->staging[i].src_free, cudaEventDisableTiming)) != cudaSuccess) st = -(int)e;
      else if ((e = cudaEventCreateWithFlags(&c->staging[i].rec_free, cudaEventDisableTiming)) != cudaSuccess) st = -(int)e;
    }
    if (st != PISLAM_OK) break;
    if ((e = cudaEventCreateWithFlags(&c->ev_sub, cudaEventDisableTiming)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaEventCreateWithFlags(&c->ev_orb, cudaEventDisableTiming)) != cudaSuccess) { st = -(int)e; break; }
    if (const char* sf = getenv("PISLAM_B200_SUB_FRAMES")) c->sub_frames = std::max(0, atoi(sf));
    for (int i = 0; i < pislam_ctx::kGatherSlots && st == PISLAM_OK; ++i)
      if ((e = cudaEventCreateWithFlags(&c->gathers[i].done, cudaEventDisableTiming)) != cudaSuccess) st = -(int)e;
    if (st != PISLAM_OK) break;
    for (int i = 0; i < pislam_ctx::kMaxChunks && st == PISLAM_OK; ++i) {
      if ((e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming)) != cudaSuccess) st = -(int)e;
      else if ((e = cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming)) != cudaSuccess) st = -(int)e;
    }
    if (st != PISLAM_OK) break;
    if (const char* ch = getenv("PISLAM_B200_CHUNKS")) {
      c->pipeline_chunks = std::max(1, std::min(atoi(ch), (int)pislam_ctx::kMaxChunks));
      c->pipeline_chunks_forced = true;
    }
    if (const char* tr = getenv("PISLAM_B200_TRIM")) c->trim_copies = tr[0] != '0';
    if (const char* l2 = getenv("PISLAM_B200_L2_CHUNK_MB")) c->l2_chunk_bytes = (size_t)std::max(1, atoi(l2)) << 20;
    const char* no_tma = getenv("PISLAM_B200_NO_TMA");
    c->use_tma = !(no_tma && no_tma[0] == '1');
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      c->encode = reinterpret_cast<EncodeTiledFn>(fn);
    else
      cudaGetLastError();
    // rBRIEF table as byte offsets into the 31-row x 36-byte shared-memory patch
    std::vector<int8_t> t(30 * 256 * 4);
    build_brief_table(t.data());
    std::vector<uint32_t> offs(30 * 256);
    for (int i = 0; i < 30 * 256; ++i) {
      const uint32_t o0 = (uint32_t)((t[4 * i + 1] + 15) * (PATCH_PITCH_W * 4) + (t[4 * i + 0] + 15));
      const uint32_t o1 = (uint32_t)((t[4 * i + 3] + 15) * (PATCH_PITCH_W * 4) + (t[4 * i + 2] + 15));
      offs[i] = o0 | (o1 << 16);
    }
    if ((e = cudaMalloc(&c->d_brief, offs.size() * 4)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaMemcpy(c->d_brief, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
      st = -(int)e;
      break;
    }
    for (int i = 0; i < 30 * 256; ++i) {
      const uint32_t o0 = (uint32_t)((t[4 * i + 1] + 15) * OT_W + (t[4 * i + 0] + 15));
      const uint32_t o1 = (uint32_t)((t[4 * i + 3] + 15) * OT_W + (t[4 * i + 2] + 15));
      offs[i] = o0 | (o1 << 16);
    }
    if ((e = cudaMalloc(&c->d_zeros, 16384)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaMemset(c->d_zeros, 0, 16384)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaMalloc(&c->d_brief48, offs.size() * 4)) != cudaSuccess) { st = -(int)e; break; }
    if ((e = cudaMemcpy(c->d_brief48, offs.data(), offs.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
      st = -(int)e;
      break;
    }
    if (max_stack_bytes && (e = c->stacks.reserve(max_stack_bytes)) != cudaSuccess) { st = -(int)e; break; }
    (void)max_keypoints;
  } while (0);
  if (st != PISLAM_OK) {
    pislam_b200_destroy(c);
    return st;
  }
  *out = c;
  return PISLAM_OK;
}

int pislam_b200_destroy(pislam_ctx* c) {
  if (!c) return PISLAM_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_in) cudaStreamSynchronize(c->copy_in);
  if (c->copy_out) cudaStreamSynchronize(c->copy_out);
  if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
  p2p_teardown(c);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  DevBuf* bufs[] = {&c->img, &c->out, &c->kp, &c->desc, &c->aux0, &c->aux1, &c->stage, &c->stage_count,
                    &c->piece_cnt, &c->piece_off, &c->stacks, &c->records, &c->hdr};
  for (DevBuf* b : bufs) b->release();
  for (Geometry& e : c->geos) e.tiles.release();
  for (pislam_ctx::Staging& st_ : c->staging) {
    st_.stacks.release();
    st_.bases.release();
    st_.records.release();
    if (st_.src_free) cudaEventDestroy(st_.src_free);
    if (st_.rec_free) cudaEventDestroy(st_.rec_free);
  }
  for (pislam_ctx::PyrTables& e : c->pyr) e.taps.release();
  c->bases.release();
  c->blur_tmp.release();
  c->blur_out.release();
  if (c->d_brief) cudaFree(c->d_brief);
  if (c->h_rec) cudaFreeHost(c->h_rec);
  if (c->d_brief48) cudaFree(c->d_brief48);
  if (c->d_zeros) cudaFree(c->d_zeros);
  for (int i = 0; i < pislam_ctx::kMaxChunks; ++i) {
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
  }
  if (c->ev_extract) cudaEventDestroy(c->ev_extract);
  if (c->ev_sub) cudaEventDestroy(c->ev_sub);
  if (c->ev_orb) cudaEventDestroy(c->ev_orb);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->flag_stream) {
    cudaStreamSynchronize(c->flag_stream);
    cudaStreamDestroy(c->flag_stream);
  }
  if (c->orb_stream) {
    cudaStreamSynchronize(c->orb_stream);
    cudaStreamDestroy(c->orb_stream);
  }
  for (int i = 0; i < pislam_ctx::kGatherSlots; ++i)
    if (c->gathers[i].done) cudaEventDestroy(c->gathers[i].done);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return PISLAM_OK;
}

const char* pislam_b200_error_string(int status) {
  switch (status) {
    case PISLAM_OK: return "ok";
    case PISLAM_ERR_INVALID_ARG: return "invalid argument";
    case PISLAM_ERR_UNSUPPORTED: return "geometry or parameters outside the supported domain";
    case PISLAM_ERR_CAPACITY: return "keypoint capacity exceeded";
    case PISLAM_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case PISLAM_ERR_PEER_TIMEOUT: return "a peer rank never reached the record exchange (10 s)";
    default: break;
  }
  if (status <= PISLAM_ERR_NCCL) return "NCCL error";
  if (status < 0) return cudaGetErrorString((cudaError_t)(-status));
  return "unknown status";
}

void* pislam_b200_stream(pislam_ctx* c) { return c ? (void*)c->stream : nullptr; }

int pislam_b200_synchronize(pislam_ctx* c) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->comm_stream));
  CK(cudaStreamSynchronize(c->copy_in));
  CK(cudaStreamSynchronize(c->copy_out));
  CK(cudaStreamSynchronize(c->flag_stream));
  if (c->d_p2p_timeout) {
    int t = 0;
    CK(cudaMemcpy(&t, c->d_p2p_timeout, 4, cudaMemcpyDeviceToHost));
    if (t) return PISLAM_ERR_PEER_TIMEOUT;
  }
  return PISLAM_OK;
}

long pislam_b200_launch_count(pislam_ctx* c) { return c ? c->launches : 0; }
long pislam_b200_geometry_uploads(pislam_ctx* c) { return c ? c->geo_uploads : 0; }

int pislam_b200_set_tma(pislam_ctx* c, int enable) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  c->use_tma = enable != 0;
  return PISLAM_OK;
}

int pislam_b200_set_sub_batch(pislam_ctx* c, int frames) {
  if (!c || frames < 0) return PISLAM_ERR_INVALID_ARG;
  c->sub_frames = frames;
  return PISLAM_OK;
}

int pislam_b200_set_profiling(pislam_ctx* c, int enable) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  c->profiling = enable != 0;
  return PISLAM_OK;
}

int pislam_b200_kernel_times(pislam_ctx* c, double ms[5], long count[5]) {
  if (!c || !ms || !count) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->comm_stream));
  for (int k = 0; k < 5; ++k) {
    ms[k] = 0.0;
    count[k] = 0;
  }
  for (auto& s : c->spans) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) {
      ms[s.which] += t;
      ++count[s.which];
    } else {
      cudaGetLastError();
    }
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  c->spans.clear();
  return PISLAM_OK;
}

// ---- literal stages -----------------------------------------------------------------------------------
int pislam_b200_fast_detect(pislam_ctx* c, const uint8_t* img, uint8_t* out, int vstep, int border, int width,
                            int height, int threshold) {
  if (!c || !img || !out || vstep < 1 || width < 1 || height < 1 || border < 3 || width > vstep)
    return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  const int iw = width - 2 * border, ih = height - 2 * border;
  if (iw <= 0 || ih <= 0) return PISLAM_OK;  // both reference loops are empty
  const size_t bytes = (size_t)height * vstep;
  CK(c->img.reserve(bytes, false, c->stream));
  CK(c->out.reserve(bytes, false, c->stream));
  CK(cudaMemcpyAsync(c->img.p, img, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->out.p, out, bytes, cudaMemcpyHostToDevice, c->stream));
  const int xcount = 16 * ((iw + 15) / 16);
  fast_detect_kernel<<<dim3((xcount + 127) / 128, ih), 128, 0, c->stream>>>(c->img.as<uint8_t>(), c->out.as<uint8_t>(),
                                                                            vstep, border, width, height, threshold);
  CKS(check_launch(c));
  if (width % 16 != 0) {
    fast_detect_tail_kernel<<<(ih + 127) / 128, 128, 0, c->stream>>>(c->out.as<uint8_t>(), vstep, border, width, height);
    CKS(check_launch(c));
  }
  CK(cudaMemcpyAsync(out, c->out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_fast_score_harris(pislam_ctx* c, const uint8_t* img, uint8_t* out, int vstep, int border, int width,
                                  int height, int32_t threshold) {
  if (!c || !img || !out || vstep < 1 || width < 1 || height < 1 || border < 3 || width > vstep)
    return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  const int iw = width - 2 * border, ih = height - 2 * border;
  if (iw <= 0 || ih <= 0) return PISLAM_OK;
  const size_t bytes = (size_t)height * vstep;
  CK(c->img.reserve(bytes, false, c->stream));
  CK(c->out.reserve(bytes, false, c->stream));
  CK(cudaMemcpyAsync(c->img.p, img, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->out.p, out, bytes, cudaMemcpyHostToDevice, c->stream));
  fast_score_harris_kernel<<<dim3((iw + 127) / 128, ih), 128, 0, c->stream>>>(
      c->img.as<uint8_t>(), c->out.as<uint8_t>(), vstep, border, width, height, threshold);
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(out, c->out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_fast_extract(pislam_ctx* c, const uint8_t* out, int vstep, int border, int width, int height,
                             int logBucketSize, int bucketLimit, uint32_t* kp, int cap, int* n) {
  if (!c || !out || !n || cap < 0 || (cap > 0 && !kp) || vstep < 1 || border < 2 || width > vstep)
    return PISLAM_ERR_INVALID_ARG;
  *n = 0;
  CKS(set_device(c));
  pislam_pyramid_desc d;
  memset(&d, 0, sizeof(d));
  d.vstep = vstep;
  d.nlevels = 1;
  d.widths[0] = width;
  d.heights[0] = height;
  d.row_offsets[0] = 0;
  d.border = border;
  d.log_bucket_size = logBucketSize;
  d.bucket_limit = bucketLimit;
  d.words = 0;
  // capacity: one keypoint per 2x2 cell at most
  const int iw = width - 2 * border, ih = height - 2 * border;
  if (iw <= 0 || ih <= 0) return PISLAM_OK;
  const int maxkp = ((iw + 1) / 2) * ((ih + 1) / 2);
  const int dcap = std::max(1, std::min(maxkp, std::max(cap, 1)));
  CKS(validate_desc(&d, dcap, false));
  CKS(prepare_geometry(c, &d, dcap, 1));
  const size_t bytes = (size_t)height * vstep;
  CK(c->out.reserve(bytes, false, c->stream));
  CK(cudaMemcpyAsync(c->out.p, out, bytes, cudaMemcpyHostToDevice, c->stream));
  CKS(reserve_front_scratch(c, 1));
  const RecordLayout rec = RecordLayout::make(dcap, 0);
  CK(c->records.reserve(rec.bytes, false, c->stream));
  FrontParams P;
  fill_front_params(c, P, c->out.as<uint8_t>(), bytes);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (c->geo->th == TH_BATCH)
    front_tile_kernel<MODE_SCOREMAP, TH_BATCH><<<dim3(c->geo->ntiles, 1), NT, 0, c->stream>>>(P, tmap);
  else
    front_tile_kernel<MODE_SCOREMAP, TH_SMALL><<<dim3(c->geo->ntiles, 1), NT, 0, c->stream>>>(P, tmap);
  CKS(check_launch(c));
  CKS(launch_order(c, 1, 0, c->records.as<uint8_t>()));
  uint32_t header[2];
  CK(cudaMemcpyAsync(header, c->records.p, 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *n = (int)header[1];
  if (header[1] > (uint32_t)cap || header[1] > header[0]) return PISLAM_ERR_CAPACITY;
  if (header[0])
    CK(cudaMemcpy(kp, c->records.as<uint8_t>() + rec.kp_off, (size_t)header[0] * 4, cudaMemcpyDeviceToHost));
  return PISLAM_OK;
}

int pislam_b200_harris_score_sobel(pislam_ctx* c, const uint8_t* img, int vstep, int rows, const int* xs,
                                   const int* ys, int npts, int32_t threshold, uint8_t* scores) {
  if (!c || !img || vstep < 1 || rows < 1 || npts < 0 || (npts && (!xs || !ys || !scores))) return PISLAM_ERR_INVALID_ARG;
  if (npts == 0) return PISLAM_OK;
  for (int i = 0; i < npts; ++i)  // the 8x8 window rows y-3..y+4, columns x-3..x+4 (Harris.h:103-111)
    if (!window_inside(xs[i], ys[i], -3, 4, vstep, rows)) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  const size_t bytes = (size_t)rows * vstep;
  CK(c->img.reserve(bytes, false, c->stream));
  CK(c->aux0.reserve((size_t)npts * 8, false, c->stream));
  CK(c->aux1.reserve((size_t)npts, false, c->stream));
  int* dx = c->aux0.as<int>();
  int* dy = dx + npts;
  CK(cudaMemcpyAsync(c->img.p, img, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(dx, xs, (size_t)npts * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(dy, ys, (size_t)npts * 4, cudaMemcpyHostToDevice, c->stream));
  harris_points_kernel<<<(npts + 127) / 128, 128, 0, c->stream>>>(c->img.as<uint8_t>(), vstep, dx, dy, npts, threshold,
                                                                  c->aux1.as<uint8_t>());
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(scores, c->aux1.p, (size_t)npts, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_orb_centroids(pislam_ctx* c, const uint8_t* img, int vstep, int rows, const uint32_t* kp, int n,
                              int32_t* centroids) {
  if (!c || !img || vstep < 1 || rows < 1 || n < 0 || (n && (!kp || !centroids))) return PISLAM_ERR_INVALID_ARG;
  if (n == 0) return PISLAM_OK;
  if (!keypoints_inside(kp, n, vstep, rows)) return PISLAM_ERR_INVALID_ARG;  // 31x31 patch outside the image
  CKS(set_device(c));
  const size_t bytes = (size_t)rows * vstep;
  const int size = (2 * n + 7) & ~7;  // Orb.h:113-114
  CK(c->img.reserve(bytes, false, c->stream));
  CK(c->kp.reserve((size_t)n * 4, false, c->stream));
  CK(c->aux0.reserve((size_t)size * 4, false, c->stream));
  CK(cudaMemcpyAsync(c->img.p, img, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->kp.p, kp, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(c->aux0.p, 0, (size_t)size * 4, c->stream));
  orb_centroids_kernel<<<(n + ORB_WARPS - 1) / ORB_WARPS, ORB_WARPS * 32, 0, c->stream>>>(
      c->img.as<uint8_t>(), c->img.as<uint8_t>() + bytes, vstep, c->kp.as<uint32_t>(), n, c->aux0.as<int32_t>());
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(centroids, c->aux0.p, (size_t)size * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_atan2(pislam_ctx* c, const int32_t* xys, int nxys, uint8_t* angles) {
  if (!c || nxys < 0 || (nxys % 8) != 0 || (nxys && (!xys || !angles))) return PISLAM_ERR_INVALID_ARG;
  if (nxys == 0) return PISLAM_OK;
  CKS(set_device(c));
  CK(c->aux0.reserve((size_t)nxys * 4, false, c->stream));
  CK(c->aux1.reserve((size_t)nxys / 2, false, c->stream));
  CK(cudaMemcpyAsync(c->aux0.p, xys, (size_t)nxys * 4, cudaMemcpyHostToDevice, c->stream));
  atan2_kernel<<<(nxys / 2 + 127) / 128, 128, 0, c->stream>>>(c->aux0.as<int32_t>(), nxys / 8, c->aux1.as<uint8_t>());
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(angles, c->aux1.p, (size_t)nxys / 2, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_brief_describe(pislam_ctx* c, const uint8_t* img, int vstep, int rows, int x, int y, int rot,
                               int words, uint32_t* desc) {
  if (!c || !img || !desc || vstep < 1 || rows < 1 || words < 1 || words > 8) return PISLAM_ERR_INVALID_ARG;
  if (rot < 0 || rot >= 30) return PISLAM_OK;  // Brief.h:641-732: descriptor untouched
  if (!window_inside(x, y, -15, 15, vstep, rows)) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  const size_t bytes = (size_t)rows * vstep;
  CK(c->img.reserve(bytes, false, c->stream));
  CK(c->desc.reserve(32, false, c->stream));
  CK(cudaMemcpyAsync(c->img.p, img, bytes, cudaMemcpyHostToDevice, c->stream));
  brief_describe_kernel<<<1, 32, 0, c->stream>>>(c->img.as<uint8_t>(), c->img.as<uint8_t>() + bytes, vstep, x, y, rot,
                                                 words, c->d_brief, c->desc.as<uint32_t>());
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(desc, c->desc.p, (size_t)words * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_orb_compute(pislam_ctx* c, const uint8_t* img, int vstep, int rows, const uint32_t* kp, int n,
                            int words, uint32_t* desc, uint8_t* angles) {
  if (!c || !img || vstep < 1 || rows < 1 || n < 0 || words < 1 || words > 8 || (n && (!kp || !desc)))
    return PISLAM_ERR_INVALID_ARG;
  if (n == 0) return PISLAM_OK;
  if (!keypoints_inside(kp, n, vstep, rows)) return PISLAM_ERR_INVALID_ARG;  // 31x31 patch outside the image
  CKS(set_device(c));
  const size_t bytes = (size_t)rows * vstep;
  CK(c->img.reserve(bytes, false, c->stream));
  CK(c->kp.reserve((size_t)n * 4, false, c->stream));
  CK(c->desc.reserve((size_t)n * words * 4, false, c->stream));
  CK(c->aux1.reserve((size_t)n, false, c->stream));
  CK(cudaMemcpyAsync(c->img.p, img, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->kp.p, kp, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
  OrbParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.img = c->img.as<uint8_t>();
  Q.frame_stride = bytes;
  Q.frame_bytes = bytes;
  Q.vstep = vstep;
  Q.words = words;
  Q.kp = c->kp.as<uint32_t>();
  Q.n = n;
  Q.desc = c->desc.as<uint32_t>();
  Q.angles = c->aux1.as<uint8_t>();
  Q.brief = c->d_brief;
  Q.nframes = 1;
  Q.cap = n;
  orb_kernel<0><<<dim3(std::min((n + ORB_WARPS - 1) / ORB_WARPS, c->num_sms * 8)), ORB_WARPS * 32, 0, c->stream>>>(Q);
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(desc, c->desc.p, (size_t)n * words * 4, cudaMemcpyDeviceToHost, c->stream));
  if (angles) CK(cudaMemcpyAsync(angles, c->aux1.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

// ---- pyramid helpers (not on the hot path) ----------------------------------------------------------
namespace {
// img/out round-trip like the literal stages: only the bytes the reference writes change in `out`
int pyramid_helper(pislam_ctx* c, const uint8_t* img, uint8_t* out, int vstep, int width, int height, int kind) {
  if (!c || !img || !out || vstep < 1 || width < 1 || height < 1 || width > vstep) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  const int B = kind == 1 ? 8 : (kind == 2 ? 16 : 1);
  const int rows_in = (height + B - 1) / B * B;  // the reference reads whole blocks (padded input)
  if (kind == 0 && (width < 3 || height < 3)) return PISLAM_ERR_INVALID_ARG;
  if (kind != 0 && (width + B - 1) / B * B > vstep) return PISLAM_ERR_INVALID_ARG;
  const int out_rows = kind == 0 ? height : (kind == 1 ? 7 : 13) * ((height + B - 1) / B);  // rows the reference writes
  const size_t in_bytes = (size_t)rows_in * vstep, out_bytes = (size_t)out_rows * vstep;
  CK(c->img.reserve(in_bytes + 16, false, c->stream));
  CK(c->out.reserve(out_bytes + 16, false, c->stream));
  CK(cudaMemcpyAsync(c->img.p, img, in_bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->out.p, out, out_bytes, cudaMemcpyHostToDevice, c->stream));
  if (kind == 0) {
    const int tpitch = (width + 3) & ~3;
    CK(c->aux0.reserve((size_t)tpitch * height + 16, false, c->stream));
    const dim3 grid(((width + 3) / 4 + 127) / 128, height);
    gaussian_vertical_kernel<<<grid, 128, 0, c->stream>>>(c->img.as<uint8_t>(), c->aux0.as<uint8_t>(), vstep, tpitch,
                                                          width, height);
    CKS(check_launch(c));
    gaussian_horizontal_kernel<<<grid, 128, 0, c->stream>>>(c->aux0.as<uint8_t>(), c->out.as<uint8_t>(), vstep, tpitch,
                                                            width, height);
    CKS(check_launch(c));
  } else {
    const int nbx = (width + B - 1) / B, nby = (height + B - 1) / B;
    const int O = kind == 1 ? 7 : 13;
    const dim3 grid((O * nbx + 1 + 127) / 128, O * nby);
    if (kind == 1)
      bilinear_blocks_kernel<8, 7, false><<<grid, 128, 0, c->stream>>>(c->img.as<uint8_t>(), c->out.as<uint8_t>(), vstep,
                                                                       nbx, nby, 1);
    else
      bilinear_blocks_kernel<16, 13, true><<<grid, 128, 0, c->stream>>>(c->img.as<uint8_t>(), c->out.as<uint8_t>(), vstep,
                                                                        nbx, nby, 0);
    CKS(check_launch(c));
  }
  CK(cudaMemcpyAsync(out, c->out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}
}  // namespace

int pislam_b200_gaussian5x5(pislam_ctx* c, const uint8_t* img, uint8_t* out, int vstep, int width, int height) {
  return pyramid_helper(c, img, out, vstep, width, height, 0);
}
int pislam_b200_bilinear7_8(pislam_ctx* c, const uint8_t* img, uint8_t* out, int vstep, int width, int height) {
  return pyramid_helper(c, img, out, vstep, width, height, 1);
}
int pislam_b200_bilinear13_16(pislam_ctx* c, const uint8_t* img, uint8_t* out, int vstep, int width, int height) {
  return pyramid_helper(c, img, out, vstep, width, height, 2);
}

// ---- fused path ---------------------------------------------------------------------------------------
size_t pislam_b200_record_bytes(int cap, int words) { return RecordLayout::make(cap, words).bytes; }
size_t pislam_b200_record_kp_offset(int cap, int words) { return RecordLayout::make(cap, words).kp_off; }
size_t pislam_b200_record_desc_offset(int cap, int words) { return RecordLayout::make(cap, words).desc_off; }
size_t pislam_b200_record_level_offset(int cap, int words) { return RecordLayout::make(cap, words).level_off; }

namespace {
// What a fused call starts from: stacked pyramids (the reference's input, README.md:56-83) or base frames that the
// pyramid builder expands on the device first.  Host or device memory.
struct FrameSource {
  const uint8_t* p;
  bool base;      // base frames (pitch / blur apply) instead of stacks
  int pitch;      // base frames: row pitch
  size_t stride;  // bytes between frames
  int blur;       // base frames: Gaussian.h:48-72 blur before resizing
};

// The whole fused path for a batch, blocking.  Host sources are pipelined in chunks over three streams:
// copy_in (H2D of chunk i+1) || stream (kernels of chunk i) || copy_out (D2H of the records of chunk i-1).
// Enqueues the whole fused path for a batch; nothing here waits for the device.  Host sources are pipelined in
// chunks over three streams: copy_in (H2D of chunk i+1) || stream (kernels of chunk i) || copy_out (D2H of the
// records of chunk i-1).  The device staging (stacks / base frames / records of host destinations) is double-buffered
// per call, so a caller that issues calls back to back (pislam_b200_extract_*_async) keeps the copy engine busy
// across calls: the H2D of call k+1 runs under the kernels and the D2H of call k.
int pipeline_enqueue(pislam_ctx* c, const FrameSource& src, int nframes, const pislam_pyramid_desc* d, int cap,
                     void* records, uint8_t** d_records_out, bool* rec_on_host_out) {
  if (!c || !src.p || !records || nframes < 1) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  CKS(validate_desc(d, cap, true));
  long long rows = 0;
  for (int l = 0; l < d->nlevels; ++l) rows = std::max<long long>(rows, (long long)d->row_offsets[l] + d->heights[l]);
  const size_t frame_bytes = (size_t)rows * d->vstep;
  size_t stack_stride = src.base ? frame_bytes : (nframes == 1 ? frame_bytes : src.stride);
  if (src.base) CKS(validate_frames(d, src.pitch, src.stride, nframes, src.blur));
  else if (stack_stride < frame_bytes) return PISLAM_ERR_INVALID_ARG;
  const size_t rec_bytes = RecordLayout::make(cap, d->words).bytes;
  const bool src_on_host = !is_device_ptr(src.p);
  const bool rec_on_host = !is_device_ptr(records);
  *rec_on_host_out = rec_on_host;
  // staging set of this call; the other set may still be in use by the previous call's kernels and copies
  const int b = c->stage_turn;
  c->stage_turn ^= 1;
  pislam_ctx::Staging& S = c->staging[b];
  CK(cudaStreamWaitEvent(c->stream, S.rec_free, 0));   // the D2H of the call before last has left S.records
  CK(cudaStreamWaitEvent(c->copy_in, S.src_free, 0));  // ... and its kernels have read S.stacks / S.bases
  uint8_t* d_records = static_cast<uint8_t*>(records);
  if (rec_on_host) {
    CK(S.records.reserve_any(rec_bytes * nframes));
    d_records = S.records.as<uint8_t>();
  }
  *d_records_out = d_records;
  if (!src.base && !src_on_host) {
    CKS(enqueue_extract(c, src.p, stack_stride, nframes, d, cap, d_records));
    if (rec_on_host) {
      CK(cudaEventRecord(c->ev_k[0], c->stream));
      CK(cudaStreamWaitEvent(c->copy_out, c->ev_k[0], 0));
      CK(cudaMemcpyAsync(records, d_records, rec_bytes * nframes, cudaMemcpyDeviceToHost, c->copy_out));
    }
  } else {
    // chunks: a host source is cut for the copy/compute overlap; device-resident base frames are cut so that a
    // chunk's freshly built stacks are still in L2 when the detector and the descriptor kernel read them
    int per = nframes;
    if (src_on_host) {
      // stacks travel as one pitched copy per level and chunk (few, large chunks: 4 measured best); base frames as one
      // linear copy per chunk, so smaller chunks (~5 MB, at most 16) shorten the un-overlapped head and tail of the
      // pipeline (profiles/r02_e2e_chunks.txt: 130 -> 148 Gpix/s from 4 to 16 chunks, worse again at 32)
      int nchunks = c->pipeline_chunks;
      if (src.base && !c->pipeline_chunks_forced) {
        const size_t bytes = (size_t)d->widths[0] * d->heights[0] * nframes;
        nchunks = (int)std::min<size_t>(16, std::max<size_t>(1, bytes / (5u << 20)));
      }
      nchunks = std::max(1, std::min(nchunks, nframes));
      per = (nframes + nchunks - 1) / nchunks;
    } else {
      per = (int)std::max<long long>(1, std::min<long long>(nframes, (long long)c->l2_chunk_bytes / (long long)frame_bytes));
    }
    uint8_t* d_stacks = nullptr;
    const uint8_t* d_bases = src.p;
    int bp = src.pitch;
    size_t bstride = src.stride;
    if (src.base) {
      // the stacks of ALL chunks stay allocated (the descriptor kernel of a chunk may still run while the next
      // chunk's pyramid is being built); padding columns are never read by the detector, so they are not cleared
      CK(S.stacks.reserve_any(frame_bytes * nframes));
      d_stacks = S.stacks.as<uint8_t>();
      if (src_on_host) {
        bp = (d->widths[0] + 15) & ~15;
        bstride = (size_t)bp * d->heights[0];
        CK(S.bases.reserve_any(bstride * nframes));
        d_bases = S.bases.as<uint8_t>();
      }
    } else {
      const size_t total = (size_t)(nframes - 1) * stack_stride + frame_bytes;
      CK(S.stacks.reserve_any(total));
      d_stacks = S.stacks.as<uint8_t>();
    }
    int i = 0;
    for (int f0 = 0; f0 < nframes; f0 += per, ++i) {
      const int nf = std::min(per, nframes - f0);
      const int e = i % pislam_ctx::kMaxChunks;
      if (src_on_host) {
        if (src.base)
          CKS(copy_bases_h2d(c->copy_in, src.p, src.pitch, src.stride, S.bases.as<uint8_t>(), bp, d->widths[0],
                             d->heights[0], f0, nf));
        else
          CKS(copy_frames_h2d(c, c->copy_in, src.p, d_stacks, stack_stride, frame_bytes, f0, nf, d));
        CK(cudaEventRecord(c->ev_in[e], c->copy_in));
        CK(cudaStreamWaitEvent(c->stream, c->ev_in[e], 0));
      }
      if (src.base)
        CKS(enqueue_pyramid(c, d_bases + (size_t)f0 * bstride, bp, bstride, nf, src.blur, d,
                            d_stacks + (size_t)f0 * stack_stride, stack_stride, false));
      CKS(enqueue_extract(c, d_stacks + (size_t)f0 * stack_stride, stack_stride, nf, d, cap,
                          d_records + (size_t)f0 * rec_bytes));
      if (rec_on_host) {
        CK(cudaEventRecord(c->ev_k[e], c->stream));
        CK(cudaStreamWaitEvent(c->copy_out, c->ev_k[e], 0));
        CK(cudaMemcpyAsync(static_cast<uint8_t*>(records) + (size_t)f0 * rec_bytes, d_records + (size_t)f0 * rec_bytes,
                           rec_bytes * nf, cudaMemcpyDeviceToHost, c->copy_out));
      }
    }
  }
  // the staging set is free again once this call's kernels have run and its records have left the device
  CK(cudaEventRecord(S.src_free, c->stream));
  CK(cudaEventRecord(S.rec_free, rec_on_host ? c->copy_out : c->stream));
  return PISLAM_OK;
}

// The blocking form: enqueue, wait, report capacity overflow
int extract_pipeline(pislam_ctx* c, const FrameSource& src, int nframes, const pislam_pyramid_desc* d, int cap,
                     void* records) {
  uint8_t* d_records = nullptr;
  bool rec_on_host = false;
  CKS(pipeline_enqueue(c, src, nframes, d, cap, records, &d_records, &rec_on_host));
  CK(cudaStreamSynchronize(c->stream));
  if (rec_on_host) CK(cudaStreamSynchronize(c->copy_out));
  const size_t rec_bytes = RecordLayout::make(cap, d->words).bytes;
  if (rec_on_host) {
    const uint8_t* r = static_cast<const uint8_t*>(records);
    for (int f = 0; f < nframes; ++f) {
      const uint32_t* h = reinterpret_cast<const uint32_t*>(r + (size_t)f * rec_bytes);
      if (h[1] > h[0]) return PISLAM_ERR_CAPACITY;
    }
    return PISLAM_OK;
  }
  return check_capacity(c, d_records, nframes, rec_bytes);
}
}  // namespace

int pislam_b200_extract_batch_async(pislam_ctx* c, const uint8_t* stacks, size_t frame_stride, int nframes,
                                    const pislam_pyramid_desc* d, int cap, void* records) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  if (stacks && records && is_device_ptr(stacks) && is_device_ptr(records))
    return enqueue_extract(c, stacks, frame_stride, nframes, d, cap, static_cast<uint8_t*>(records));
  const FrameSource src = {stacks, false, 0, frame_stride, 0};
  uint8_t* d_records = nullptr;
  bool rec_on_host = false;
  return pipeline_enqueue(c, src, nframes, d, cap, records, &d_records, &rec_on_host);
}

int pislam_b200_extract_frames_async(pislam_ctx* c, const uint8_t* frames, int frame_pitch, size_t frame_stride,
                                     int nframes, int blur, const pislam_pyramid_desc* d, int cap, void* records) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  const FrameSource src = {frames, true, frame_pitch, frame_stride, blur};
  uint8_t* d_records = nullptr;
  bool rec_on_host = false;
  return pipeline_enqueue(c, src, nframes, d, cap, records, &d_records, &rec_on_host);
}

int pislam_b200_extract_batch(pislam_ctx* c, const uint8_t* stacks, size_t frame_stride, int nframes,
                              const pislam_pyramid_desc* d, int cap, void* records) {
  const FrameSource src = {stacks, false, 0, frame_stride, 0};
  return extract_pipeline(c, src, nframes, d, cap, records);
}

int pislam_b200_extract_frames(pislam_ctx* c, const uint8_t* frames, int frame_pitch, size_t frame_stride, int nframes,
                               int blur, const pislam_pyramid_desc* d, int cap, void* records) {
  const FrameSource src = {frames, true, frame_pitch, frame_stride, blur};
  return extract_pipeline(c, src, nframes, d, cap, records);
}

int pislam_b200_build_pyramid(pislam_ctx* c, const uint8_t* frames, int frame_pitch, size_t frame_stride, int nframes,
                              int blur, const pislam_pyramid_desc* d, uint8_t* stacks, size_t stack_stride) {
  if (!c || !frames || !stacks || nframes < 1) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  CKS(validate_desc(d, 1, false));
  CKS(validate_frames(d, frame_pitch, frame_stride, nframes, blur));
  long long rows = 0;
  for (int l = 0; l < d->nlevels; ++l) rows = std::max<long long>(rows, (long long)d->row_offsets[l] + d->heights[l]);
  const size_t frame_bytes = (size_t)rows * d->vstep;
  if (nframes == 1) stack_stride = frame_bytes;
  if (stack_stride < frame_bytes) return PISLAM_ERR_INVALID_ARG;
  const uint8_t* d_frames = frames;
  int bp = frame_pitch;
  size_t bstride = frame_stride;
  if (!is_device_ptr(frames)) {
    bp = (d->widths[0] + 15) & ~15;
    bstride = (size_t)bp * d->heights[0];
    CK(c->bases.reserve(bstride * nframes, false, c->stream));
    CKS(copy_bases_h2d(c->stream, frames, frame_pitch, frame_stride, c->bases.as<uint8_t>(), bp, d->widths[0],
                       d->heights[0], 0, nframes));
    d_frames = c->bases.as<uint8_t>();
  }
  const bool out_on_host = !is_device_ptr(stacks);
  uint8_t* d_stacks = stacks;
  const size_t total = (size_t)(nframes - 1) * stack_stride + frame_bytes;
  if (out_on_host) {
    CK(c->stacks.reserve(total, false, c->stream));
    d_stacks = c->stacks.as<uint8_t>();
    // rows of the stack that belong to no level come back as the caller left them
    CK(cudaMemcpyAsync(d_stacks, stacks, total, cudaMemcpyHostToDevice, c->stream));
  }
  CKS(enqueue_pyramid(c, d_frames, bp, bstride, nframes, blur, d, d_stacks, stack_stride, true));
  if (out_on_host) {
    CK(cudaMemcpyAsync(stacks, d_stacks, total, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return PISLAM_OK;
}

int pislam_b200_extract_pyramid(pislam_ctx* c, const uint8_t* stack, const pislam_pyramid_desc* d, uint32_t* kp,
                                uint8_t* kp_level, uint32_t* descriptors, int cap, int* n) {
  if (!c || !stack || !kp || !n || cap < 1) return PISLAM_ERR_INVALID_ARG;
  *n = 0;
  CKS(set_device(c));
  CKS(validate_desc(d, cap, true));
  const RecordLayout rec = RecordLayout::make(cap, d->words);
  CK(c->records.reserve(rec.bytes, false, c->stream));
  // run with the ctx record buffer as a DEVICE destination, then scatter to the caller's arrays
  long long rows = 0;
  for (int l = 0; l < d->nlevels; ++l) rows = std::max<long long>(rows, (long long)d->row_offsets[l] + d->heights[l]);
  const size_t frame_bytes = (size_t)rows * d->vstep;
  const uint8_t* d_stack = stack;
  if (!is_device_ptr(stack)) {
    CK(c->stacks.reserve(frame_bytes, false, c->stream));
    CK(cudaMemcpyAsync(c->stacks.p, stack, frame_bytes, cudaMemcpyHostToDevice, c->stream));
    d_stack = c->stacks.as<uint8_t>();
  }
  uint8_t* d_rec = c->records.as<uint8_t>();
  CKS(enqueue_extract(c, d_stack, frame_bytes, 1, d, cap, d_rec));
  // one D2H of the whole (small) record into pinned staging, one sync, then scatter on the host
  if (c->h_rec_cap < rec.bytes) {
    if (c->h_rec) cudaFreeHost(c->h_rec);
    c->h_rec = nullptr;
    c->h_rec_cap = 0;
    CK(cudaHostAlloc(&c->h_rec, rec.bytes, cudaHostAllocDefault));
    c->h_rec_cap = rec.bytes;
  }
  const uint8_t* h = static_cast<const uint8_t*>(c->h_rec);
  const size_t small = 1u << 20;  // large capacities: fetch the header first, then only the used part
  if (rec.bytes <= small) {
    CK(cudaMemcpyAsync(c->h_rec, d_rec, rec.bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  } else {
    CK(cudaMemcpyAsync(c->h_rec, d_rec, REC_HEADER_WORDS * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  const uint32_t* header = reinterpret_cast<const uint32_t*>(h);
  *n = (int)header[1];
  if (header[1] > header[0]) return PISLAM_ERR_CAPACITY;
  const size_t m = header[0];
  if (m) {
    if (rec.bytes > small) {
      uint8_t* hw = static_cast<uint8_t*>(c->h_rec);
      CK(cudaMemcpyAsync(hw + rec.kp_off, d_rec + rec.kp_off, m * 4, cudaMemcpyDeviceToHost, c->stream));
      if (kp_level) CK(cudaMemcpyAsync(hw + rec.level_off, d_rec + rec.level_off, m, cudaMemcpyDeviceToHost, c->stream));
      if (descriptors && d->words > 0)
        CK(cudaMemcpyAsync(hw + rec.desc_off, d_rec + rec.desc_off, m * 4 * d->words, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    memcpy(kp, h + rec.kp_off, m * 4);
    if (kp_level) memcpy(kp_level, h + rec.level_off, m);
    if (descriptors && d->words > 0) memcpy(descriptors, h + rec.desc_off, m * 4 * d->words);
  }
  return PISLAM_OK;
}

// ---- multi-GPU ----------------------------------------------------------------------------------------
int pislam_b200_comm_unique_id(void* id128) {
  if (!id128) return PISLAM_ERR_INVALID_ARG;
  if (!g_nccl.load()) return PISLAM_ERR_NCCL;
  NcclUid id;
  const int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return PISLAM_ERR_NCCL - r;
  memcpy(id128, &id, sizeof(id));
  return PISLAM_OK;
}

int pislam_b200_comm_init(pislam_ctx* c, int nranks, int rank, const void* id128) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return PISLAM_ERR_INVALID_ARG;
  if (!g_nccl.load()) return PISLAM_ERR_NCCL;
  CKS(set_device(c));
  NcclUid id;
  memcpy(&id, id128, sizeof(id));
  if (c->comm) CKS(pislam_b200_comm_destroy(c));  // a second init replaces the communicator (no leak)
  const int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) {
    c->comm = nullptr;
    return PISLAM_ERR_NCCL - r;
  }
  c->nranks = nranks;
  c->rank = rank;
  return p2p_setup(c);  // collective: the flag blocks of the size-aware exchange (falls back to NCCL when IPC is unavailable)
}

int pislam_b200_comm_destroy(pislam_ctx* c) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  if (c->comm) {
    CKS(set_device(c));
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamSynchronize(c->flag_stream);
    p2p_teardown(c);
    g_nccl.CommDestroy(c->comm);
    c->comm = nullptr;
  }
  c->nranks = 1;
  c->rank = 0;
  return PISLAM_OK;
}

int pislam_b200_comm_register(pislam_ctx* c, void* d_buf, size_t bytes) {
  if (!c || !d_buf || bytes == 0 || (reinterpret_cast<uintptr_t>(d_buf) & 15) != 0) return PISLAM_ERR_INVALID_ARG;
  if (!c->comm) return PISLAM_ERR_NCCL;
  CKS(set_device(c));
  // collective even when this rank cannot take part: every rank learns the common verdict
  IpcBlob mine;
  memset(&mine, 0, sizeof(mine));
  if (c->p2p_ok && (int)c->windows.size() < P2P_MAX_WINDOWS && is_device_ptr(d_buf)) ipc_describe(d_buf, bytes, &mine);
  std::vector<uint8_t> all;
  CKS(comm_allgather_host(c, &mine, sizeof(mine), all));
  int ok = 1;
  for (int r = 0; r < c->nranks; ++r) {
    const IpcBlob* b = reinterpret_cast<const IpcBlob*>(all.data() + r * sizeof(IpcBlob));
    ok = ok && b->ok && b->bytes == bytes;
  }
  pislam_ctx::PeerWindow w;
  w.local = static_cast<uint8_t*>(d_buf);
  w.bytes = bytes;
  for (int r = 0; r < c->nranks && ok; ++r) {
    const IpcBlob* b = reinterpret_cast<const IpcBlob*>(all.data() + r * sizeof(IpcBlob));
    if (r == c->rank) {
      w.peer[r] = w.local;
      continue;
    }
    void* base = nullptr;
    if (ipc_open(c, b->handle, r, &base) != PISLAM_OK) ok = 0;
    else w.peer[r] = static_cast<uint8_t*>(base) + b->offset;
  }
  CKS(comm_allgather_host(c, &ok, sizeof(ok), all));
  for (int r = 0; r < c->nranks; ++r) ok = ok && *reinterpret_cast<const int*>(all.data() + r * sizeof(int));
  if (!ok) return PISLAM_ERR_UNSUPPORTED;  // on every rank alike: the buffer keeps working through ncclAllGather
  c->windows.push_back(w);
  return PISLAM_OK;
}

int pislam_b200_set_exchange_ctas(pislam_ctx* c, int ctas) {
  if (!c || ctas < 1) return PISLAM_ERR_INVALID_ARG;
  c->push_ctas = ctas;
  return PISLAM_OK;
}

int pislam_b200_comm_p2p_enabled(pislam_ctx* c) { return (c && c->p2p_ok) ? (int)c->windows.size() : 0; }

namespace {
int extract_allgather_impl(pislam_ctx* c, const FrameSource& src, int nframes_local, int slots_per_rank,
                           const pislam_pyramid_desc* d, int cap, void* d_all_records) {
  const uint8_t* d_stacks = src.p;
  const size_t frame_stride = src.stride;
  if (!c || !d_all_records || !d || slots_per_rank < 1 || nframes_local < 0 || nframes_local > slots_per_rank ||
      (nframes_local > 0 && !d_stacks))
    return PISLAM_ERR_INVALID_ARG;
  if (!c->comm) return PISLAM_ERR_NCCL;
  CKS(set_device(c));
  // everything that can fail on this rank alone is checked BEFORE any work is enqueued: once the kernels are
  // in flight the exchange is issued unconditionally, so no peer is left waiting
  CKS(validate_desc(d, cap, true));
  const RecordLayout rec = RecordLayout::make(cap, d->words);
  const size_t rec_bytes = rec.bytes;
  const size_t slot = rec_bytes * (size_t)slots_per_rank;
  // the kernels write this rank's records straight into its slot of the gather buffer (in-place gather)
  uint8_t* all = static_cast<uint8_t*>(d_all_records);
  uint8_t* mine = all + slot * (size_t)c->rank;
  // a buffer registered with pislam_b200_comm_register takes the size-aware exchange over peer mappings
  int win = -1;
  for (size_t i = 0; i < c->windows.size(); ++i)
    if (all >= c->windows[i].local && all + slot * (size_t)c->nranks <= c->windows[i].local + c->windows[i].bytes)
      win = (int)i;
  // a gather that is still filling this buffer must finish before the kernels write into it again; with two
  // alternating buffers the previous gather (other buffer) keeps running under this batch's kernels
  pislam_ctx::GatherSlot& g = c->gathers[c->gather_turn];
  for (int i = 0; i < pislam_ctx::kGatherSlots; ++i)
    if (c->gathers[i].ptr == d_all_records) CK(cudaStreamWaitEvent(c->stream, c->gathers[i].done, 0));
  // a buffer beyond the tracked ones evicts the oldest entry: from then on nothing remembers that entry's gather, so
  // order this call (and through the stream everything later) behind it
  if (g.ptr && g.ptr != d_all_records) CK(cudaStreamWaitEvent(c->stream, g.done, 0));
  P2pPeers peers;
  memset(&peers, 0, sizeof(peers));
  uint32_t epoch = 0;
  if (win >= 0) {
    pislam_ctx::PeerWindow& w = c->windows[win];
    epoch = ++w.epoch;
    for (int r = 0; r < c->nranks; ++r) {
      peers.buf[r] = w.peer[r] + (all - w.local);
      peers.flags[r] = c->peer_flags[r];
    }
    // READY(e): from here on (everything enqueued on the ctx stream so far has run) the peers may overwrite their
    // slots of this rank's buffer.  The flag goes out on its own stream: on the ctx stream it would sit between the
    // descriptor kernel of one batch and the detector of the next
    CK(cudaEventRecord(c->ev_ready, c->stream));
    CK(cudaStreamWaitEvent(c->flag_stream, c->ev_ready, 0));
    p2p_flag_kernel<<<1, 32, 0, c->flag_stream>>>(peers, c->nranks, c->rank, win, P2P_READY, epoch);
    CKS(check_launch(c));
  }
  int st = PISLAM_OK;
  if (nframes_local > 0 && (src.base || !is_device_ptr(d_stacks))) {
    // host stacks, or base frames (host or device): the chunked pipeline of the blocking calls, enqueue only, with
    // this rank's slot of the gather buffer as its (device) destination
    uint8_t* d_rec = nullptr;
    bool rec_on_host = false;
    st = pipeline_enqueue(c, src, nframes_local, d, cap, mine, &d_rec, &rec_on_host);
  } else {
    st = enqueue_extract(c, d_stacks, frame_stride, nframes_local, d, cap, mine);
  }
  // a ragged shard leaves trailing slots unused: their headers read "no keypoints" (n = n_total = 0, level starts 0)
  if (st == PISLAM_OK && nframes_local < slots_per_rank) {
    cudaError_t e = cudaMemset2DAsync(mine + rec_bytes * (size_t)nframes_local, rec_bytes, 0, REC_HEADER_WORDS * 4,
                                      (size_t)(slots_per_rank - nframes_local), c->stream);
    if (e != cudaSuccess) st = -(int)e;
  }
  cudaError_t e = cudaEventRecord(c->ev_extract, c->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(c->comm_stream, c->ev_extract, 0);
  if (e != cudaSuccess && st == PISLAM_OK) st = -(int)e;
  static const bool no_exchange = getenv("PISLAM_B200_DEBUG_NO_EXCHANGE") != nullptr;  // timing experiments only
  span_begin(c, 4, c->comm_stream);
  if (no_exchange) {
  } else if (win >= 0) {
    P2pPushParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.peers = peers;
    Q.nranks = c->nranks;
    Q.rank = c->rank;
    Q.slots_per_rank = slots_per_rank;
    Q.slot_offset = slot * (size_t)c->rank;
    Q.rec = rec;
    Q.words = d->words;
    Q.cap = cap;
    p2p_wait_kernel<<<1, 32, 0, c->comm_stream>>>(c->d_flags, c->nranks, c->rank, win, P2P_READY, epoch, c->d_p2p_timeout);
    p2p_push_kernel<<<std::min(c->push_ctas, slots_per_rank * (c->nranks - 1)), 256, 0, c->comm_stream>>>(Q);
    p2p_flag_kernel<<<1, 32, 0, c->comm_stream>>>(peers, c->nranks, c->rank, win, P2P_DONE, epoch);
    p2p_wait_kernel<<<1, 32, 0, c->comm_stream>>>(c->d_flags, c->nranks, c->rank, win, P2P_DONE, epoch, c->d_p2p_timeout);
    c->launches += 3;
    const int ls = check_launch(c);
    if (ls != PISLAM_OK && st == PISLAM_OK) st = ls;
  } else {
    const int r = g_nccl.AllGather(mine, d_all_records, slot, /*ncclInt8*/ 0, c->comm, c->comm_stream);
    if (r != 0) return PISLAM_ERR_NCCL - r;
  }
  span_end(c, c->comm_stream);
  c->gather_turn = (c->gather_turn + 1) % pislam_ctx::kGatherSlots;
  g.ptr = d_all_records;
  CK(cudaEventRecord(g.done, c->comm_stream));
  return st;
}
}  // namespace

int pislam_b200_extract_allgather(pislam_ctx* c, const uint8_t* stacks, size_t frame_stride, int nframes_local,
                                  int slots_per_rank, const pislam_pyramid_desc* d, int cap, void* d_all_records) {
  const FrameSource src = {stacks, false, 0, frame_stride, 0};
  return extract_allgather_impl(c, src, nframes_local, slots_per_rank, d, cap, d_all_records);
}

int pislam_b200_extract_frames_allgather(pislam_ctx* c, const uint8_t* frames, int frame_pitch, size_t frame_stride,
                                         int nframes_local, int blur, int slots_per_rank, const pislam_pyramid_desc* d,
                                         int cap, void* d_all_records) {
  const FrameSource src = {frames, true, frame_pitch, frame_stride, blur};
  return extract_allgather_impl(c, src, nframes_local, slots_per_rank, d, cap, d_all_records);
}

int pislam_b200_join(pislam_ctx* c) {
  if (!c) return PISLAM_ERR_INVALID_ARG;
  CKS(set_device(c));
  for (int i = 0; i < pislam_ctx::kGatherSlots; ++i)
    if (c->gathers[i].ptr) CK(cudaStreamWaitEvent(c->stream, c->gathers[i].done, 0));
  // ... and for the record copies of the issue-only host pipelines (pislam_b200_extract_*_async)
  for (int i = 0; i < 2; ++i) CK(cudaStreamWaitEvent(c->stream, c->staging[i].rec_free, 0));
  return PISLAM_OK;
}

int pislam_b200_allgather(pislam_ctx* c, void* d_all, size_t bytes_per_rank) {
  if (!c || !d_all) return PISLAM_ERR_INVALID_ARG;
  if (!c->comm) return PISLAM_ERR_NCCL;
  CKS(set_device(c));
  const uint8_t* mine = static_cast<const uint8_t*>(d_all) + bytes_per_rank * (size_t)c->rank;
  const int r = g_nccl.AllGather(mine, d_all, bytes_per_rank, /*ncclInt8*/ 0, c->comm, c->stream);
  if (r != 0) return PISLAM_ERR_NCCL - r;
  return PISLAM_OK;
}

// ---- descriptor matching (definition in include/pislam_b200.h) ------------------------
int pislam_b200_match_hamming(pislam_ctx* c, const uint32_t* q_desc, const uint32_t* q_kp, const uint8_t* q_level,
                              int nq, const uint32_t* t_desc, const uint32_t* t_kp, const uint8_t* t_level, int nt,
                              int words, int radius, int32_t* matches) {
  if (!c || nq < 0 || nt < 0 || words < 1 || words > 8 || (nq && (!q_desc || !matches)) || (nt && !t_desc))
    return PISLAM_ERR_INVALID_ARG;
  if (radius > 0 && ((nq && (!q_kp || !q_level)) || (nt && (!t_kp || !t_level)))) return PISLAM_ERR_INVALID_ARG;
  if (radius > 5792) return PISLAM_ERR_INVALID_ARG;  // radius^2 and the 12-bit coordinate differences stay in int32
  if (nq == 0) return PISLAM_OK;
  CKS(set_device(c));
  // one staging buffer: query desc | train desc | query kp | train kp | query level | train level | out
  const size_t qd = (size_t)nq * words * 4, td = (size_t)nt * words * 4, qk = (size_t)nq * 4, tk = (size_t)nt * 4;
  const size_t lv_q = ((size_t)nq + 3) & ~(size_t)3, lv_t = ((size_t)nt + 3) & ~(size_t)3, ob = (size_t)nq * 12;
  CK(c->aux0.reserve(qd + td + qk + tk + lv_q + lv_t + ob + 64, false, c->stream));
  uint8_t* base = c->aux0.as<uint8_t>();
  uint8_t *d_qd = base, *d_td = d_qd + qd, *d_qk = d_td + td, *d_tk = d_qk + qk, *d_ql = d_tk + tk, *d_tl = d_ql + lv_q,
          *d_out = d_tl + lv_t;
  CK(cudaMemcpyAsync(d_qd, q_desc, qd, cudaMemcpyHostToDevice, c->stream));
  if (nt) CK(cudaMemcpyAsync(d_td, t_desc, td, cudaMemcpyHostToDevice, c->stream));
  if (radius > 0) {
    CK(cudaMemcpyAsync(d_qk, q_kp, qk, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d_ql, q_level, (size_t)nq, cudaMemcpyHostToDevice, c->stream));
    if (nt) {
      CK(cudaMemcpyAsync(d_tk, t_kp, tk, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(d_tl, t_level, (size_t)nt, cudaMemcpyHostToDevice, c->stream));
    }
  }
  MatchParams M;
  memset(&M, 0, sizeof(M));
  M.words = words;
  M.radius = radius;
  M.q_desc = reinterpret_cast<const uint32_t*>(d_qd);
  M.q_kp = reinterpret_cast<const uint32_t*>(d_qk);
  M.q_level = d_ql;
  M.nq = nq;
  M.t_desc = reinterpret_cast<const uint32_t*>(d_td);
  M.t_kp = reinterpret_cast<const uint32_t*>(d_tk);
  M.t_level = d_tl;
  M.nt = nt;
  M.out = reinterpret_cast<int32_t*>(d_out);
  match_kernel<0><<<dim3((nq + MATCH_NT - 1) / MATCH_NT, 1), MATCH_NT, 0, c->stream>>>(M);
  CKS(check_launch(c));
  CK(cudaMemcpyAsync(matches, d_out, ob, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

int pislam_b200_match_records(pislam_ctx* c, const uint8_t* records, int nframes, int cap, int words,
                              const int32_t* pairs, int npairs, int radius, int32_t* matches) {
  if (!c || !records || nframes < 1 || cap < 1 || words < 1 || words > 8 || npairs < 0 || (npairs && (!pairs || !matches)))
    return PISLAM_ERR_INVALID_ARG;
  if (radius > 5792 || npairs > 65535) return PISLAM_ERR_INVALID_ARG;
  for (int i = 0; i < 2 * npairs; ++i)
    if (pairs[i] < 0 || pairs[i] >= nframes) return PISLAM_ERR_INVALID_ARG;
  if (npairs == 0) return PISLAM_OK;
  CKS(set_device(c));
  const RecordLayout rec = RecordLayout::make(cap, words);
  const bool rec_on_host = !is_device_ptr(records), out_on_host = !is_device_ptr(matches);
  const uint8_t* d_rec = records;
  if (rec_on_host) {
    CK(c->records.reserve((size_t)nframes * rec.bytes, false, c->stream));
    CK(cudaMemcpyAsync(c->records.p, records, (size_t)nframes * rec.bytes, cudaMemcpyHostToDevice, c->stream));
    d_rec = c->records.as<uint8_t>();
  }
  const size_t out_bytes = (size_t)npairs * cap * 12, pair_bytes = (size_t)npairs * 8;
  CK(c->aux1.reserve(pair_bytes + (out_on_host ? out_bytes : 0) + 64, false, c->stream));
  int32_t* d_pairs = c->aux1.as<int32_t>();
  int32_t* d_out = out_on_host ? reinterpret_cast<int32_t*>(c->aux1.as<uint8_t>() + ((pair_bytes + 15) & ~(size_t)15)) : matches;
  CK(cudaMemcpyAsync(d_pairs, pairs, pair_bytes, cudaMemcpyHostToDevice, c->stream));
  if (out_on_host) CK(cudaMemcpyAsync(d_out, matches, out_bytes, cudaMemcpyHostToDevice, c->stream));  // untouched rows stay
  MatchParams M;
  memset(&M, 0, sizeof(M));
  M.words = words;
  M.radius = radius;
  M.records = d_rec;
  M.rec = rec;
  M.cap = cap;
  M.pairs = d_pairs;
  M.out = d_out;
  match_kernel<1><<<dim3((cap + MATCH_NT - 1) / MATCH_NT, npairs), MATCH_NT, 0, c->stream>>>(M);
  CKS(check_launch(c));
  if (out_on_host) CK(cudaMemcpyAsync(matches, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  if (out_on_host || rec_on_host) CK(cudaStreamSynchronize(c->stream));
  return PISLAM_OK;
}

// ---- introspection ------------------------------------------------------------------------------------
int pislam_b200_brief_table(int8_t* table30x256x4) {
  if (!table30x256x4) return PISLAM_ERR_INVALID_ARG;
  build_brief_table(table30x256x4);
  return PISLAM_OK;
}

}  // extern "C"
