/* pislam_b200 -- C ABI of the B200 (sm_100a) ORB front-end.
 *
 * This is the whole drop-in boundary: plain pointers and sizes, no C++/torch types.  The
 * reference (0xfaded/pislam) is a header-only C++ template library without any FFI; the
 * `pislam::` template wrappers in include/pislam/{Fast,Harris,Orb,Brief,Util}.h forward to the
 * entry points below with their template arguments (vstep, border, logBucketSize, bucketLimit,
 * words) as run-time integers.  Every entry point cites the reference function it replaces.
 *
 * Conventions
 *   - status return: 0 = ok; -(cudaError_t) for a CUDA failure; PISLAM_ERR_* otherwise.
 *     Nothing throws.  There is NO CPU fallback: without a usable GPU pislam_b200_create fails.
 *   - one pislam_ctx per host thread and GPU; it owns one CUDA stream, device staging buffers
 *     and scratch.  Calls on one ctx are serialised by the caller (the reference is
 *     single-threaded and re-entrant; use one ctx per thread for concurrency).
 *   - images are row-major uint8 with a row pitch of `vstep` bytes; `img` / `out` point at the
 *     first row of the LEVEL being processed, like the reference's `uint8_t img[][vstep]`.
 *   - "literal stage" calls take HOST pointers, do H2D + kernel + D2H and leave the host
 *     buffers exactly as the reference would.  The fused calls accept host OR device pointers
 *     (detected with cudaPointerGetAttributes) for every buffer.
 */
#ifndef PISLAM_B200_H_
#define PISLAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PISLAM_OK 0
#define PISLAM_ERR_INVALID_ARG (-10001) /* null pointer, negative size, words outside 1..8 ... */
#define PISLAM_ERR_UNSUPPORTED (-10002) /* geometry outside the fused kernels' domain, see DESIGN.md */
#define PISLAM_ERR_CAPACITY (-10003)    /* more keypoints than `cap`; *n receives the required count */
#define PISLAM_ERR_NO_DEVICE (-10004)   /* no CUDA device / wrong architecture */
#define PISLAM_ERR_PEER_TIMEOUT (-10005) /* a peer rank never reached the record exchange */
#define PISLAM_ERR_NCCL (-11000)        /* -11000 - ncclResult_t */

#define PISLAM_MAX_LEVELS 16

typedef struct pislam_ctx pislam_ctx;

/* ---- context ------------------------------------------------------------------------------ */
/* max_stack_bytes / max_keypoints are sizing hints (buffers grow on demand). */
int pislam_b200_create(pislam_ctx** ctx, int device, size_t max_stack_bytes, int max_keypoints);
int pislam_b200_destroy(pislam_ctx* ctx);
const char* pislam_b200_error_string(int status);
/* The CUDA stream (cudaStream_t) all work of this ctx is issued on. */
void* pislam_b200_stream(pislam_ctx* ctx);
/* Blocks until everything this ctx has issued has finished (kernels, copies of the host pipelines, exchanges). */
int pislam_b200_synchronize(pislam_ctx* ctx);
/* Number of kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
long pislam_b200_launch_count(pislam_ctx* ctx);
/* Number of tile-table uploads so far: the fused calls cache the tile / piece tables of the last few pyramid
 * geometries per ctx (LRU), so a caller alternating a handful of geometries uploads each of them once. */
long pislam_b200_geometry_uploads(pislam_ctx* ctx);
/* Tile staging of the fused kernel: 1 = TMA box loads (default), 0 = plain loads.  The TMA path is also
 * skipped automatically when vstep / frame_stride / the base address are not multiples of 16 bytes. */
int pislam_b200_set_tma(pislam_ctx* ctx, int enable);
/* Sub-batching of the fused calls: frames per sub-batch (0 = off).  The descriptor kernel of sub-batch k then runs
 * on a second stream under the detector of sub-batch k+1, while the sub-batch's stacks are still in the L2. */
int pislam_b200_set_sub_batch(pislam_ctx* ctx, int frames);
/* Per-kernel timing of the fused path (bench.py's roofline): while enabled, CUDA events are recorded on
 * the ctx stream around every launch of [0] front_tile_kernel, [1] order_kernel, [2] orb_kernel,
 * [3] pyramid_build_kernel, [4] the record exchange of pislam_b200_extract_allgather (on its own stream: from the
 * moment the records are ready to the moment every peer's records have arrived).
 * pislam_b200_kernel_times synchronises, returns the summed durations (ms) and launch counts since the
 * last call, and clears them. */
int pislam_b200_set_profiling(pislam_ctx* ctx, int enable);
int pislam_b200_kernel_times(pislam_ctx* ctx, double ms[5], long count[5]);

/* ---- literal stages (host buffers) ---------------------------------------------------------- */
/* pislam::fastDetect<vstep,border>(width,height,img,out,threshold)        reference Fast.h:54-158 */
int pislam_b200_fast_detect(pislam_ctx* ctx, const uint8_t* img, uint8_t* out, int vstep, int border,
                            int width, int height, int threshold);
/* pislam::fastScoreHarris<vstep,border>(width,height,img,threshold,out)   reference Fast.h:166-180 */
int pislam_b200_fast_score_harris(pislam_ctx* ctx, const uint8_t* img, uint8_t* out, int vstep, int border,
                                  int width, int height, int32_t threshold);
/* pislam::fastExtract<vstep,border,logBucketSize,bucketLimit>(width,height,out,results)
 * reference Fast.h:196-355.  Writes kp[0..*n) (the entries the reference would append). */
int pislam_b200_fast_extract(pislam_ctx* ctx, const uint8_t* out, int vstep, int border, int width,
                             int height, int logBucketSize, int bucketLimit, uint32_t* kp, int cap, int* n);
/* pislam::harrisScoreSobel<vstep>(img,x,y,threshold)                      reference Harris.h:80-248
 * `img` points at row 0 of a buffer of `rows` rows; scores npts points (xs[i], ys[i]). */
int pislam_b200_harris_score_sobel(pislam_ctx* ctx, const uint8_t* img, int vstep, int rows, const int* xs,
                                   const int* ys, int npts, int32_t threshold, uint8_t* scores);
/* pislam::orbCentroids<vstep>(img,points)                                 reference Orb.h:80-308
 * centroids receives (2n+7)&~7 int32 in the reference's [x0 x1 x2 x3 y0 y1 y2 y3] layout. */
int pislam_b200_orb_centroids(pislam_ctx* ctx, const uint8_t* img, int vstep, int rows, const uint32_t* kp,
                              int n, int32_t* centroids);
/* pislam::atan2(xys)                                                      reference Orb.h:310-387
 * nxys must be a multiple of 8; angles receives nxys/2 bins. */
int pislam_b200_atan2(pislam_ctx* ctx, const int32_t* xys, int nxys, uint8_t* angles);
/* pislam::briefDescribe<vstep,words>(img,x,y,rot,descriptor)              reference Brief.h:637-733
 * rot outside 0..29 leaves `desc` untouched, like the reference's switch. */
int pislam_b200_brief_describe(pislam_ctx* ctx, const uint8_t* img, int vstep, int rows, int x, int y,
                               int rot, int words, uint32_t* desc);
/* pislam::orbCompute<vstep,words>(img,points,descriptors)                 reference Orb.h:396-441
 * desc receives n*words uint32 (what the reference appends); angles (may be NULL) the n bins. */
int pislam_b200_orb_compute(pislam_ctx* ctx, const uint8_t* img, int vstep, int rows, const uint32_t* kp,
                            int n, int words, uint32_t* desc, uint8_t* angles);

/* ---- pyramid-preparation helpers (not on the ORB hot path; host buffers) -----------------------
 * pislam::gaussian5x5<vstep>(width,height,img,out)      reference Gaussian.h:48-72 (semantics of its test
 *                                                       model test/GaussianTest.cpp:159-215); width,height >= 3
 * pislam::bilinear7_8<vstep>(width,height,img,out)      reference Bilinear.h:42-150
 * pislam::bilinear13_16<vstep>(width,height,img,out)    reference Bilinear.h:165-321
 * img and out may be the same buffer.  The resamplers read whole 8x8 / 16x16 blocks: like the reference
 * they need the image padded to a multiple of 8 / 16 rows and columns (inside the pitch). */
int pislam_b200_gaussian5x5(pislam_ctx* ctx, const uint8_t* img, uint8_t* out, int vstep, int width, int height);
int pislam_b200_bilinear7_8(pislam_ctx* ctx, const uint8_t* img, uint8_t* out, int vstep, int width, int height);
int pislam_b200_bilinear13_16(pislam_ctx* ctx, const uint8_t* img, uint8_t* out, int vstep, int width, int height);

/* ---- fused whole-pyramid path (the caller loop of reference demo/demo.cpp:77-101) ----------- */
typedef struct pislam_pyramid_desc {
  int vstep;   /* row pitch of the stack in bytes */
  int nlevels; /* <= PISLAM_MAX_LEVELS */
  int widths[PISLAM_MAX_LEVELS];
  int heights[PISLAM_MAX_LEVELS];
  int row_offsets[PISLAM_MAX_LEVELS]; /* first stack row of each level */
  int border;                         /* >= 5 (>= 16 when descriptors are wanted) */
  int fast_threshold;
  int32_t harris_threshold;
  int log_bucket_size; /* 0 = no bucketing; fused path supports 0..5 */
  int bucket_limit;
  int words; /* 1..8 descriptor words; 0 = keypoints only */
} pislam_pyramid_desc;

/* Per-frame result record, fixed capacity so that a batch is one contiguous buffer (and one
 * NCCL all-gather): see pislam_b200_record_bytes / the offsets below.
 *   uint32 header[32]: [0]=n (keypoints stored, <= cap) [1]=n_total (found) [2..2+nlevels]=index of
 *                      each level's first keypoint (level_start[nlevels] == n)
 *   uint32 kp[cap]           encodeFast(score,x,y) with LEVEL-LOCAL y        (Util.h:27-29)
 *   uint32 desc[cap][words]
 *   uint8  level[cap]        (padded to 16 bytes)
 */
size_t pislam_b200_record_bytes(int cap, int words);
size_t pislam_b200_record_kp_offset(int cap, int words);
size_t pislam_b200_record_desc_offset(int cap, int words);
size_t pislam_b200_record_level_offset(int cap, int words);

/* fastDetect + fastScoreHarris + fastExtract per level, then orbCompute, for `nframes`
 * independent pyramids of identical geometry; frame f starts at stacks + f*frame_stride.
 * `stacks` and `records` may each be host or device memory; records receives
 * nframes * pislam_b200_record_bytes(cap, words) bytes.  Keypoint order and content are the
 * reference's (with `out` zero-initialised).  Returns PISLAM_ERR_CAPACITY if any frame found
 * more than `cap` keypoints (header[1] of that frame then holds the count to size `cap` for; its
 * stored keypoints are unspecified). */
int pislam_b200_extract_batch(pislam_ctx* ctx, const uint8_t* stacks, size_t frame_stride, int nframes,
                              const pislam_pyramid_desc* desc, int cap, void* records);

/* Same, issue only: returns after enqueueing (nothing waits for the device); pislam_b200_synchronize waits for
 * everything, records on the host included.  Capacity overflow is reported by header[1] > header[0] of a frame.
 * Device stacks and device records: the kernels are enqueued on the ctx stream.  Host buffers (pinned memory, or
 * the copies are not asynchronous): the chunked H2D | kernels | D2H pipeline of pislam_b200_extract_batch with
 * double-buffered device staging, so consecutive calls overlap -- the H2D of call k+1 runs under the kernels and
 * the D2H of call k, and the copy engine never idles between calls.  The caller must not reuse a host records
 * buffer before the call that fills it has completed. */
int pislam_b200_extract_batch_async(pislam_ctx* ctx, const uint8_t* stacks, size_t frame_stride,
                                    int nframes, const pislam_pyramid_desc* desc, int cap, void* records);

/* Single pyramid convenience wrapper over extract_batch that scatters the record into the
 * caller's arrays (host pointers; kp_level / desc may be NULL). */
int pislam_b200_extract_pyramid(pislam_ctx* ctx, const uint8_t* stack, const pislam_pyramid_desc* desc,
                                uint32_t* kp, uint8_t* kp_level, uint32_t* descriptors, int cap, int* n);

/* ---- pyramid builder in front of the path --------------------------
 * The reference consumes a pre-built stacked pyramid and leaves its construction to the caller
 * ("externally computing the image pyramid", reference README.md:28-31) -- on a GPU, it suggests.  These two
 * entry points do that on the device, so that only the BASE frame crosses PCIe (1/3 of the stack's level
 * pixels for 8 levels at x1/1.2).  NOT a replacement of reference code: the reference ships no x1/1.2
 * resize, so the definition is this library's own and its parity is unpinned by the reference
 * (oracle/pislam_oracle.h: po_resize_bilinear / po_build_pyramid; the CUDA path is bit-exact to those).
 *   level l = pixel-centre-aligned bilinear resize of the base frame (widths[0] x heights[0]) to
 *   widths[l] x heights[l]: 8.8 fixed-point positions, 8-bit weights, one rounding; level 0 is the base
 *   frame itself.  blur != 0 first applies pislam::gaussian5x5 (reference Gaussian.h:48-72) to the base
 *   frame -- the reference expects pre-blurred input (README.md:28-31).
 * frames: nframes base frames, row pitch frame_pitch, frame f at frames + f*frame_stride; host or device. */
/* base frames -> stacked pyramids in the layout `desc` describes (rows of the stack that belong to no level are
 * left as they are; columns right of each level are zero-filled).  stacks: host or device. */
int pislam_b200_build_pyramid(pislam_ctx* ctx, const uint8_t* frames, int frame_pitch, size_t frame_stride,
                              int nframes, int blur, const pislam_pyramid_desc* desc, uint8_t* stacks,
                              size_t stack_stride);
/* base frames -> records: pislam_b200_build_pyramid into device memory followed by the fused path of
 * pislam_b200_extract_batch, chunk by chunk (host frames: H2D of chunk i+1 under the kernels of chunk i under
 * the D2H of the records of chunk i-1).  Same records, status and capacity behaviour as extract_batch on the
 * stacks build_pyramid would produce. */
int pislam_b200_extract_frames(pislam_ctx* ctx, const uint8_t* frames, int frame_pitch, size_t frame_stride,
                               int nframes, int blur, const pislam_pyramid_desc* desc, int cap, void* records);

/* issue-only form of pislam_b200_extract_frames (see pislam_b200_extract_batch_async) */
int pislam_b200_extract_frames_async(pislam_ctx* ctx, const uint8_t* frames, int frame_pitch, size_t frame_stride,
                                     int nframes, int blur, const pislam_pyramid_desc* desc, int cap, void* records);

/* ---- multi-GPU: frames shard one-per-rank, results meet in ONE all-gather --------------------- */
/* 128-byte ncclUniqueId; create on rank 0, distribute with any host-side channel. */
int pislam_b200_comm_unique_id(void* id128);
int pislam_b200_comm_init(pislam_ctx* ctx, int nranks, int rank, const void* id128);
int pislam_b200_comm_destroy(pislam_ctx* ctx);
/* extract_batch on this rank's `nframes_local` frames (0 <= nframes_local <= slots_per_rank), then a single
 * ncclAllGather of the fixed-capacity records: d_all_records (device) receives
 * nranks * slots_per_rank * record_bytes, rank-major; the kernels write this rank's records straight into
 * its slot (in-place gather).  `slots_per_rank` must be the same on every rank (ceil(total frames / nranks)
 * for a contiguous split: trailing ranks may own fewer frames, or none); the unused slots of a rank arrive
 * with an all-zero header (n = 0).  Arguments are validated before anything is enqueued, and once the kernels
 * are enqueued the collective is always issued, so a failing rank never leaves its peers waiting.
 * The gather is issued on a second stream behind the kernels, so with two alternating gather
 * buffers it overlaps the next batch's kernels; a buffer is never rewritten before its previous
 * gather has finished.  Enqueue only: pislam_b200_synchronize waits for kernels and gathers. */
int pislam_b200_extract_allgather(pislam_ctx* ctx, const uint8_t* stacks, size_t frame_stride,
                                  int nframes_local, int slots_per_rank, const pislam_pyramid_desc* desc,
                                  int cap, void* d_all_records);
/* `stacks` may also be HOST memory (pinned): the rank's frames then go through the chunked H2D | kernels pipeline
 * of pislam_b200_extract_batch_async straight into its slot, and the exchange follows on the second stream.
 * The same from BASE frames (host or device; see pislam_b200_extract_frames): */
int pislam_b200_extract_frames_allgather(pislam_ctx* ctx, const uint8_t* frames, int frame_pitch, size_t frame_stride,
                                         int nframes_local, int blur, int slots_per_rank,
                                         const pislam_pyramid_desc* desc, int cap, void* d_all_records);

/* Registers a device buffer for the SIZE-AWARE record exchange (collective: every rank calls it, in the same
 * order, with a buffer of the same size; 16-byte aligned).  The buffer's allocation is mapped into every peer
 * (CUDA IPC over NVLink), and pislam_b200_extract_allgather on a gather buffer that lies inside a registered
 * buffer then moves only the LIVE part of every record -- header, n keypoints, n descriptors, n level bytes --
 * with a copy kernel that stores straight into the peers' buffers, instead of ncclAllGather on the
 * fixed-capacity records (~2.4x the bytes at 850 of 2048 slots).  Everything pislam_b200_extract_allgather
 * promises stays as it is (in-place slots, second stream, alternating buffers); the slack of a record beyond
 * its n entries is simply not transported.  Returns PISLAM_ERR_UNSUPPORTED -- on every rank alike -- when a
 * peer mapping is not possible (no IPC, more than 16 ranks, more than 8 buffers): the buffer then keeps working
 * through ncclAllGather.  Registrations live until pislam_b200_comm_destroy. */
int pislam_b200_comm_register(pislam_ctx* ctx, void* d_buf, size_t bytes);
/* CTAs of the exchange's copy kernel (default 32: it runs under the next batch's detector and should leave it the SMs). */
int pislam_b200_set_exchange_ctas(pislam_ctx* ctx, int ctas);
/* Number of buffers registered for the size-aware exchange (0: every gather goes through ncclAllGather). */
int pislam_b200_comm_p2p_enabled(pislam_ctx* ctx);

/* Makes the ctx stream wait (on the device) for the gathers issued so far by pislam_b200_extract_allgather and
 * for the record copies of the issue-only host pipelines (pislam_b200_extract_*_async), so that later work --
 * or an event recorded -- on the ctx stream is ordered after them. */
int pislam_b200_join(pislam_ctx* ctx);
/* In-place ncclAllGather on the ctx stream of a rank-major buffer whose slot `rank` (bytes_per_rank
 * bytes) this rank has already filled, e.g. with pislam_b200_extract_batch on device records. */
int pislam_b200_allgather(pislam_ctx* ctx, void* d_all, size_t bytes_per_rank);

/* ---- descriptor matching --------------------------------------------------
 * NOT a replacement of reference code: the reference never released its matcher (README.md:125-139
 * mentions flann-based matching with table-driven bit counts and a demo tracing "best matches within a
 * 20 pixel radius").  Definition (restated in numpy in tests/oracle_lib.py:match_hamming):
 *   distance(q, t) = popcount of the XOR of the `words` descriptor words;
 *   candidates of query q: every train descriptor (radius <= 0), or the train keypoints on the SAME pyramid
 *   level with (xq-xt)^2 + (yq-yt)^2 <= radius^2 in level-local pixels (radius > 0);
 *   result [3] per query: index of the closest candidate (smallest index wins ties; -1 if none), its
 *   distance, and the distance of the runner-up (a tie with the best counts), PISLAM_MATCH_NONE when absent.
 * kp/level may be null when radius <= 0.  Host pointers; blocking. */
#define PISLAM_MATCH_NONE 0x7fffffff
int pislam_b200_match_hamming(pislam_ctx* ctx, const uint32_t* q_desc, const uint32_t* q_kp, const uint8_t* q_level,
                              int nq, const uint32_t* t_desc, const uint32_t* t_kp, const uint8_t* t_level, int nt,
                              int words, int radius, int32_t* matches /* [nq][3] */);
/* The same for frame pairs of a batch of records (pislam_b200_extract_batch output, host or device):
 * pairs[2*p] = query frame, pairs[2*p+1] = train frame (host array); matches[p][cap][3] (host or device,
 * rows >= n of the query frame are left untouched).  One launch for all pairs; blocking only when a buffer
 * is on the host. */
int pislam_b200_match_records(pislam_ctx* ctx, const uint8_t* records, int nframes, int cap, int words,
                              const int32_t* pairs, int npairs, int radius, int32_t* matches);

/* ---- introspection for tests ----------------------------------------------------------------- */
/* The rotated + clamped rBRIEF pattern the device uses, int8[30][256][4] = (x0,y0,x1,y1)
 * (reference Brief.h:28-50). */
int pislam_b200_brief_table(int8_t* table30x256x4);

#ifdef __cplusplus
}
#endif
#endif /* PISLAM_B200_H_ */
