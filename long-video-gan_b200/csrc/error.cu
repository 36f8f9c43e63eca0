// Thread-local error string + build info for the C ABI.
#include "common.cuh"
#include <stdarg.h>

namespace lvg {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace lvg

extern "C" int lvg_abi_version(void) { return LVG_ABI_VERSION; }
extern "C" const char* lvg_last_error(void) { return lvg::g_err; }
extern "C" const char* lvg_build_info(void) {
    return "liblvg_ops sm_100a nvcc " __DATE__;
}
