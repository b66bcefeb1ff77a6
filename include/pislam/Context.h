/* pislam:: drop-in headers over libpislam_b200.so -- shared plumbing.
 *
 * The reference (0xfaded/pislam) is header-only and stateless; here every template forwards its template
 * arguments as run-time integers to the C ABI in include/pislam_b200.h.  Each host thread lazily gets its
 * own pislam_ctx (one CUDA stream on device $PISLAM_B200_DEVICE, default 0), which keeps the reference's
 * "re-entrant, safe to call concurrently on disjoint buffers" property.  There is no CPU fallback: if the
 * context cannot be created, or a call fails, a std::runtime_error carrying the C-ABI status is thrown.
 */
#ifndef PISLAM_B200_CONTEXT_H_
#define PISLAM_B200_CONTEXT_H_

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../pislam_b200.h"

namespace pislam {
namespace b200 {

inline void check(int status, const char* what) {
  if (status != PISLAM_OK)
    throw std::runtime_error(std::string("pislam_b200: ") + what + ": " + pislam_b200_error_string(status) + " (" +
                             std::to_string(status) + ")");
}

struct ThreadContext {
  pislam_ctx* ctx = nullptr;
  ThreadContext() {
    const char* dev = std::getenv("PISLAM_B200_DEVICE");
    check(pislam_b200_create(&ctx, dev ? std::atoi(dev) : 0, 0, 0), "create");
  }
  ~ThreadContext() { pislam_b200_destroy(ctx); }
  ThreadContext(const ThreadContext&) = delete;
  ThreadContext& operator=(const ThreadContext&) = delete;
};

/* the calling thread's context */
inline pislam_ctx* context() {
  static thread_local ThreadContext tc;
  return tc.ctx;
}

}  // namespace b200
}  // namespace pislam
#endif
