// Error reporting / bookkeeping shared by all entry points.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "sq_common.cuh"

namespace sq {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SQ_PDL");           // programmatic dependent launch: on by default, SQ_PDL=0 turns it off
    v = (e && !atoi(e)) ? 0 : 1;
  }
  return v == 1;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace sq

extern "C" const char* sq_last_error(void) { return sq::g_err; }
extern "C" int sq_version(void) { return 100; }
extern "C" uint64_t sq_launch_count(void) { return sq::g_launches.load(std::memory_order_relaxed); }
