// Probe: which box start coordinates does a uint8 3-D tiled TMA load accept on sm_100a?
// usage: tma_probe <x> <y> [boxw]   -> prints checksum of the loaded box vs expected (zero fill outside)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cstring>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int bw, int bh, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t s[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(s);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bw * bh) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(d), "l"(&tmap), "r"(x), "r"(y), "r"(0), "r"(b) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  for (int i = 0; i < 1000000 && !done; ++i)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(b), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = done ? s[i] : 0xEE;
}
int main(int argc, char** argv) {
  int x = atoi(argv[1]), y = atoi(argv[2]);
  int bw = argc > 3 ? atoi(argv[3]) : 144, bh = 42;
  const int W = 320, H = 500;
  std::vector<uint8_t> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *o;
  cudaMalloc(&d, W * H); cudaMalloc(&o, bw * bh);
  cudaMemcpy(d, h.data(), W * H, cudaMemcpyHostToDevice);
  void* fn; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  CUtensorMap m;
  cuuint64_t dims[3] = {W, H, 1}, strides[2] = {W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("x=%d y=%d bw=%d encode failed %d\n", x, y, bw, (int)r); return 0; }
  k<<<1, 128, bw * bh>>>(m, x, y, bw, bh, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("x=%d y=%d bw=%d launch error: %s\n", x, y, bw, cudaGetErrorString(e)); return 0; }
  std::vector<uint8_t> g(bw * bh);
  cudaMemcpy(g.data(), o, bw * bh, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
    int xx = x + i, yy = y + j;
    uint8_t want = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? h[yy * W + xx] : 0;
    bad += g[j * bw + i] != want;
  }
  printf("x=%d y=%d bw=%d ok, mismatches=%d\n", x, y, bw, bad);
  return 0;
}
