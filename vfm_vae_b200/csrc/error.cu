// Error string, launch counter, version and the optional per-kernel CUDA-event timing registry of libvfmops.
#include "common.cuh"
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <stdlib.h>
#include <string.h>
#include <vector>

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

// ---- timing registry: when enabled, every KernelTimer scope records two events on the launching stream ----
struct TimingRecord { std::string name; cudaEvent_t e0, e1; double flops, bytes; };
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
static std::vector<TimingRecord> g_records;

bool timing_enabled() { return g_timing.load(std::memory_order_relaxed) != 0; }

static bool timing_detail() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VFM_TIMING_DETAIL"); v = (e && e[0] && e[0] != '0') ? 1 : 0; }
    return v == 1;
}

KernelTimer::KernelTimer(const char* name, cudaStream_t stream, double flops, double bytes, const char* tag_fmt, ...)
    : stream_(stream), flops_(flops), bytes_(bytes), e0_(nullptr), on_(timing_enabled()) {
    if (!on_) return;
    int n = snprintf(name_, sizeof(name_), "%s", name);
    if (tag_fmt && timing_detail() && n > 0 && n < (int)sizeof(name_) - 2) {
        name_[n++] = ':';
        va_list ap;
        va_start(ap, tag_fmt);
        vsnprintf(name_ + n, sizeof(name_) - n, tag_fmt, ap);
        va_end(ap);
    }
    if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventRecord(e0_, stream_) != cudaSuccess) { on_ = false; }
}
KernelTimer::~KernelTimer() {
    if (!on_) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) != cudaSuccess) return;
    cudaEventRecord(e1, stream_);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_records.push_back(TimingRecord{name_, e0_, e1, flops_, bytes_});
}
}  // namespace vfm

extern "C" const char* vfm_last_error(void) { return vfm::g_err; }
extern "C" int vfm_abi_version(void) { return VFM_ABI_VERSION; }
extern "C" uint64_t vfm_launch_count(void) { return vfm::g_launches.load(std::memory_order_relaxed); }

extern "C" void vfm_timing_enable(int on) {
    using namespace vfm;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (on) {
        for (auto& r : g_records) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_records.clear();
    }
    g_timing.store(on ? 1 : 0);
}

extern "C" int vfm_timing_report(vfm_kernel_stat* out, int max_entries) {
    using namespace vfm;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    std::map<std::string, vfm_kernel_stat> agg;
    for (auto& r : g_records) {
        if (cudaEventSynchronize(r.e1) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        auto it = agg.find(r.name);
        if (it == agg.end()) {
            vfm_kernel_stat s;
            memset(&s, 0, sizeof(s));
            strncpy(s.name, r.name.c_str(), sizeof(s.name) - 1);
            it = agg.emplace(r.name, s).first;
        }
        it->second.launches += 1;
        it->second.total_ms += ms;
        it->second.flops += r.flops;
        it->second.bytes += r.bytes;
    }
    int n = 0;
    for (auto& kv : agg) {
        if (out && n < max_entries) out[n] = kv.second;
        n++;
    }
    return n;
}
