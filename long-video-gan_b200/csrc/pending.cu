// Entry points whose kernels are not built yet report LVG_UNSUPPORTED so that
// callers compose the general kernels instead. Each one moves to its own file
// as the kernel lands.
#include "common.cuh"

using namespace lvg;

extern "C" int lvg_conv2d_fprop(const void*, const void*, void*, int, int, int, int, int, int, int, int, int, int, int, int,
                                void*, int64_t, void*)
{
    set_error("conv2d_fprop: no kernel for this configuration");
    return LVG_UNSUPPORTED;
}

extern "C" int64_t lvg_conv2d_fprop_workspace(int, int, int, int, int, int, int, int, int, int, int, int) { return -1; }
