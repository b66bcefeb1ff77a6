// Size-aware record exchange over NVLink peer mappings (sm_100a): the path's only exchange step.
//
// The reference has no collective at all (its caller loop, demo/demo.cpp:77-101, runs on one core); frames shard
// one batch per GPU and the per-frame records meet on every GPU.  ncclAllGather has to move the fixed-capacity
// records (2048 slots per frame, ~850 used); this exchange moves only what is live:
//
//   p2p_push_kernel   a few dozen persistent CTAs loop over (record, peer) pairs: header + n keypoints + n descriptors
//                     + n level bytes of a local record go straight into the same slot of the peer's gather buffer
//                     (16-byte stores over the peer mapping of a buffer registered with pislam_b200_comm_register)
//   p2p_flag_kernel   one thread per peer: publishes an epoch in the peer's flag block (system-scope release)
//   p2p_wait_kernel   one thread per peer: waits for that peer's epoch in the local flag block (bounded spin)
//
// Protocol per call and registered buffer (epoch e, identical on every rank because the calls are collective):
//   flag stream:  (after everything enqueued on the ctx stream so far)  flag READY(e) to every peer
//   ctx stream :  detector / descriptor kernels write the local slot
//   comm stream:  wait READY(e) from every peer (their buffer may be overwritten)  ->  push  ->  flag DONE(e)
//                 ->  wait DONE(e) from every peer (their pushes have landed here)
// No kernel spins while holding more than one CTA, and a peer that never arrives trips a ~10 s timeout that is
// reported through the ctx (PISLAM_ERR_PEER_TIMEOUT) instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "front_kernels.cuh"

namespace pb {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_MAX_WINDOWS = 8;
enum { P2P_READY = 0, P2P_DONE = 1 };
// flag block of one rank: [window][kind][source rank]
__host__ __device__ __forceinline__ int p2p_flag_index(int window, int kind, int src) {
  return (window * 2 + kind) * P2P_MAX_RANKS + src;
}

struct P2pPeers {
  uint8_t* buf[P2P_MAX_RANKS];     // the registered buffer as mapped on this rank (buf[rank] = the local one)
  uint32_t* flags[P2P_MAX_RANKS];  // every rank's flag block as mapped on this rank
};

__global__ void p2p_flag_kernel(const __grid_constant__ P2pPeers peers, int nranks, int rank, int window, int kind,
                                uint32_t epoch) {
  const int p = threadIdx.x;
  if (p >= nranks || p == rank) return;
  __threadfence_system();  // everything this GPU wrote before (earlier kernels of the stream) precedes the flag
  volatile uint32_t* f = peers.flags[p] + p2p_flag_index(window, kind, rank);
  *f = epoch;
}

__global__ void p2p_wait_kernel(const uint32_t* my_flags, int nranks, int rank, int window, int kind, uint32_t epoch,
                                int* timeout_flag) {
  const int p = threadIdx.x;
  if (p >= nranks || p == rank) return;
  const volatile uint32_t* f = my_flags + p2p_flag_index(window, kind, p);
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((int32_t)(*f - epoch) < 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) {  // a peer never arrived: give up, report, do not hang
      *timeout_flag = 1;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

struct P2pPushParams {
  P2pPeers peers;
  int nranks, rank;
  int slots_per_rank;
  unsigned long long slot_offset;  // byte offset of this rank's first record inside the gather buffer
  RecordLayout rec;
  int words, cap;
};

// copies `bytes` (rounded up to whole 16-byte / 4-byte units; the slack stays inside the record's own field)
__device__ __forceinline__ void p2p_copy(uint8_t* dst, const uint8_t* src, unsigned bytes, bool vec16) {
  if (vec16) {
    const unsigned n = (bytes + 15u) >> 4;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else {
    const unsigned n = (bytes + 3u) >> 2;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
  }
}

// A small persistent grid: the exchange runs on its own stream UNDER the next batch's detector, whose grid wants
// every CTA slot; one CTA per (record, peer) pair (1792 CTAs of 256 threads for 256 records and 7 peers) takes all of
// them for ~80 us and costs the step exactly that (profiles/r02_probe_8gpu_exchange.txt: 0.498 ms per step against
// 0.419 without any exchange).  A few dozen CTAs that loop over the pairs keep enough 16-byte stores in flight for
// the NVLink rate the exchange needs (56 MB in one step time) and leave the SMs to the detector.
__global__ void __launch_bounds__(256) p2p_push_kernel(const __grid_constant__ P2pPushParams P) {
  const int npairs = P.slots_per_rank * (P.nranks - 1);
  // every field starts on a 16-byte boundary when cap is a multiple of 4 (kp 4*cap bytes, descriptors 4*cap*words)
  const bool vec16 = (P.cap & 3) == 0;
  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    // consecutive CTAs serve different peers, so every NVLink port is busy from the start
    const int pi = pair % (P.nranks - 1), slot = pair / (P.nranks - 1);
    const int peer = (P.rank + 1 + pi) % P.nranks;
    const unsigned long long off = P.slot_offset + (unsigned long long)slot * P.rec.bytes;
    const uint8_t* src = P.peers.buf[P.rank] + off;
    uint8_t* dst = P.peers.buf[peer] + off;
    const uint32_t n = min(*reinterpret_cast<const uint32_t*>(src), (uint32_t)P.cap);
    p2p_copy(dst, src, REC_HEADER_WORDS * 4, true);
    if (n == 0) continue;
    p2p_copy(dst + P.rec.kp_off, src + P.rec.kp_off, 4u * n, vec16);
    if (P.words > 0) p2p_copy(dst + P.rec.desc_off, src + P.rec.desc_off, 4u * n * (unsigned)P.words, vec16);
    p2p_copy(dst + P.rec.level_off, src + P.rec.level_off, n, vec16);
  }
}

}  // namespace pb
