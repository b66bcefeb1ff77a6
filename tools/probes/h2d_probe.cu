// H2D copy-rate probe for the level-only upload of stacked pyramids (pinned host memory).
//   nvcc -O2 -o h2d_probe h2d_probe.cu && ./h2d_probe
// Compares: one linear copy of the whole stacks; per-level pitched 3-D copies of exactly w bytes per row (what
// libpislam_b200 does); the same with the row width rounded up to 64 / 128 bytes; level 0 as one linear copy per frame.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const int W[8] = {640, 533, 444, 370, 309, 257, 214, 179}, H[8] = {480, 400, 333, 278, 231, 193, 161, 134};
  const int vstep = 640, F = 64;
  int off[8], rows = 0; size_t pix = 0;
  for (int l = 0; l < 8; ++l) { off[l] = rows; rows += H[l]; pix += (size_t)W[l] * H[l]; }
  const size_t fb = (size_t)rows * vstep;
  uint8_t *h, *d;
  CK(cudaMallocHost(&h, fb * F)); CK(cudaMalloc(&d, fb * F)); memset(h, 1, fb * F);
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](const char* name, int mode, int round) -> int {
    size_t useful = pix * F, moved = 0; float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      moved = 0;
      CK(cudaEventRecord(a, st));
      if (mode == 0) { CK(cudaMemcpyAsync(d, h, fb * F, cudaMemcpyHostToDevice, st)); moved = fb * F; }
      else for (int l = 0; l < 8; ++l) {
        if (mode == 2 && l == 0) {  // level 0 is contiguous per frame
          for (int f = 0; f < F; ++f) CK(cudaMemcpyAsync(d + f * fb, h + f * fb, (size_t)W[0] * H[0], cudaMemcpyHostToDevice, st));
          moved += (size_t)W[0] * H[0] * F; continue;
        }
        int w = std::min(vstep, (W[l] + round - 1) / round * round);
        cudaMemcpy3DParms p; memset(&p, 0, sizeof(p));
        p.srcPtr = make_cudaPitchedPtr(h, vstep, vstep, rows); p.dstPtr = make_cudaPitchedPtr(d, vstep, vstep, rows);
        p.srcPos = make_cudaPos(0, off[l], 0); p.dstPos = p.srcPos; p.extent = make_cudaExtent(w, H[l], F);
        p.kind = cudaMemcpyHostToDevice; CK(cudaMemcpy3DAsync(&p, st)); moved += (size_t)w * H[l] * F;
      }
      CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b));
      float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms);
    }
    printf("%-44s %7.3f ms  moved %6.1f MB at %5.1f GB/s  -> %6.1f Gpix/s of level pixels\n", name, best, moved / 1e6, moved / best / 1e6, useful / best / 1e6);
    return 0;
  };
  run("linear, whole stacks (with padding)", 0, 1);
  run("per-level 3-D copies, exact widths", 1, 1);
  run("per-level 3-D copies, widths up to 32", 1, 32);
  run("per-level 3-D copies, widths up to 64", 1, 64);
  run("per-level 3-D copies, widths up to 128", 1, 128);
  run("level 0 linear per frame + exact 3-D", 2, 1);
  return 0;
}
