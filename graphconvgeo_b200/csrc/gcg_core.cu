// Error string, version and launch accounting for libgcg.so.
#include <atomic>
#include <string.h>

#include "gcg_common.cuh"

namespace gcg {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace gcg

extern "C" int gcg_version(void) { return 100; }
extern "C" const char* gcg_last_error(void) { return gcg::g_err; }
extern "C" int64_t gcg_launch_count(void) { return gcg::g_launches.load(); }
extern "C" void gcg_launch_count_reset(void) { gcg::g_launches.store(0); }
