// Library-wide bits of the C ABI: per-thread error string, version probes.
#include "o4d_common.cuh"

namespace o4d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace o4d

extern "C" const char* o4d_last_error(void) { return o4d::g_err; }
extern "C" int o4d_abi_version(void) { return 1; }
