// C-ABI plumbing: error string, launch counter, version.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace zest {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace zest

extern "C" const char* zest_last_error(void) { return zest::g_err; }

// Peer / device copy on the caller's stream (the frame driver's CUDA-IPC transport: `src` may be another GPU's memory mapped
// through cudaIpcOpenMemHandle; unified addressing resolves the direction, the copy engines move the bytes).
extern "C" int zest_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
  ZEST_CHECK_ARG(dst && src && bytes >= 0, "zest_memcpy_async: bad arguments");
  if (bytes == 0) return ZEST_OK;
  ZEST_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return ZEST_OK;
}
extern "C" int zest_version(void) { return 100; }
extern "C" int64_t zest_launch_count(void) { return (int64_t)zest::g_launches.load(); }
