// Entry points whose kernels are not built yet report LVG_UNSUPPORTED so that
// callers compose the general kernels instead. Each one moves to its own file
// as the kernel lands.
#include "common.cuh"

using namespace lvg;

extern "C" int lvg_upfirdn2d_sep(const void*, const float*, const float*, void*, int, const int64_t*, const int64_t*,
                                 const int64_t*, const int64_t*, int, int, int, int, int, int, int, int, int, float, void*)
{
    set_error("upfirdn2d_sep: no tiled kernel for this configuration");
    return LVG_UNSUPPORTED;
}

extern "C" int lvg_filtered_lrelu(const void*, const float*, const float*, const void*, const uint8_t*, void*, uint8_t*, int,
                                  const int64_t*, const int64_t*, const int64_t*, const int64_t*, int, int, int, int, int, int,
                                  int, int, int, int, int, int, float, float, float, int, int, void*)
{
    set_error("filtered_lrelu: no fused kernel for this configuration");
    return LVG_UNSUPPORTED;
}

extern "C" int lvg_filtered_lrelu_supported(int, int, int, int, int, int, int) { return LVG_UNSUPPORTED; }

extern "C" int lvg_conv2d_fprop(const void*, const void*, void*, int, int, int, int, int, int, int, int, int, int, int, int,
                                void*, int64_t, void*)
{
    set_error("conv2d_fprop: no kernel for this configuration");
    return LVG_UNSUPPORTED;
}

extern "C" int64_t lvg_conv2d_fprop_workspace(int, int, int, int, int, int, int, int, int, int, int, int) { return -1; }
