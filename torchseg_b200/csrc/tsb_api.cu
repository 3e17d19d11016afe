// libtsb: library-level entry points (error string, version, launch counter).
#include <stdarg.h>
#include <atomic>

#include "tsb_common.cuh"

static thread_local char g_err[512] = "";
std::atomic<long long> g_tsb_launches{0};

void tsb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void tsb_count_launch(int n) { g_tsb_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* tsb_last_error(void) { return g_err; }
extern "C" int tsb_version(void) { return 100; }
extern "C" long long tsb_launch_count(void) { return g_tsb_launches.load(); }
