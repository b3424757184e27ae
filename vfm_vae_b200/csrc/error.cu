// Error string, launch counter and version for libvfmops.
#include "common.cuh"
#include <atomic>

namespace vfm {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace vfm

extern "C" const char* vfm_last_error(void) { return vfm::g_err; }
extern "C" int vfm_abi_version(void) { return VFM_ABI_VERSION; }
extern "C" uint64_t vfm_launch_count(void) { return vfm::g_launches.load(std::memory_order_relaxed); }
