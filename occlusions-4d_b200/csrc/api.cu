// Library-wide bits of the C ABI: per-thread error string, version probes, launch counter
// and the opt-in per-family kernel timer bench.py uses for its roofline line.
#include "o4d_common.cuh"
#include <atomic>
#include <mutex>
#include <vector>

namespace o4d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- diagnostics (not on the data path; the only process-wide state in the library) ----
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec {
    int family;
    double flops;
    cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

ProfScope::ProfScope(int family, double flops, cudaStream_t st) : rec_(-1), st_(st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r;
    r.family = family;
    r.flops = flops;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    rec_ = (int)g_prof.size() - 1;
}

ProfScope::~ProfScope() {
    if (rec_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (rec_ < (int)g_prof.size()) cudaEventRecord(g_prof[rec_].b, st_);
}

}  // namespace o4d

extern "C" const char* o4d_last_error(void) { return o4d::g_err; }
extern "C" int o4d_abi_version(void) { return 2; }
extern "C" uint64_t o4d_launch_count(void) { return o4d::g_launches.load(); }

extern "C" void o4d_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(o4d::g_prof_mu);
    for (auto& r : o4d::g_prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    o4d::g_prof.clear();
    o4d::g_prof_on.store(on ? 1 : 0);
}

extern "C" int o4d_profile_read(int n_families, double* ms_out, double* flops_out, int64_t* count_out) {
    std::lock_guard<std::mutex> lk(o4d::g_prof_mu);
    for (int f = 0; f < n_families; ++f) {
        ms_out[f] = 0.0;
        flops_out[f] = 0.0;
        count_out[f] = 0;
    }
    for (auto& r : o4d::g_prof) {
        if (r.family < 0 || r.family >= n_families) continue;
        if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        ms_out[r.family] += ms;
        flops_out[r.family] += r.flops;
        count_out[r.family] += 1;
    }
    return 0;
}
