// Library-level entry points: version, thread-local error string, launch counter.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace etpgt {
namespace {
thread_local char g_error[512] = "";
// process-wide: autograd runs backward on its own worker threads
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace etpgt

extern "C" {
int etpgt_version(void) { return 101; }
const char* etpgt_last_error(void) { return etpgt::g_error; }
int64_t etpgt_launch_count(void) { return etpgt::g_launches.load(std::memory_order_relaxed); }
void etpgt_reset_launch_count(void) { etpgt::g_launches.store(0, std::memory_order_relaxed); }
}
