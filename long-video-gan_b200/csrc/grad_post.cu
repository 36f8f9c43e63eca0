// In-place scale + NaN/Inf sanitising of the all-reduced flat gradient buffer
// (the tail of utils.sync_grads, utils.py:116-124): one read and one write per element.
#include "common.cuh"

namespace lvg {
namespace {

__device__ __forceinline__ float sanitize(float v, float scale, float limit)
{
    v *= scale;
    if (v != v) return 0.f;                          // NaN -> 0
    if (isinf(v)) return v > 0.f ? limit : -limit;   // +-inf -> +-limit; finite values pass (torch.nan_to_num)
    return v;
}

__global__ void __launch_bounds__(256) grad_post_kernel(float* __restrict__ g, int64_t n, float scale, float limit)
{
    const int64_t n4 = aligned16(g) ? n / 4 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = reinterpret_cast<float4*>(g)[i];
        v.x = sanitize(v.x, scale, limit); v.y = sanitize(v.y, scale, limit);
        v.z = sanitize(v.z, scale, limit); v.w = sanitize(v.w, scale, limit);
        reinterpret_cast<float4*>(g)[i] = v;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        g[i] = sanitize(g[i], scale, limit);
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int lvg_grad_postprocess(float* g, int64_t n, float scale, float limit, void* stream)
{
    LVG_REQUIRE(g != nullptr || n == 0, "grad_postprocess: buffer must not be NULL");
    LVG_REQUIRE(n >= 0 && limit >= 0.f, "grad_postprocess: bad arguments");
    if (n == 0) return LVG_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8 * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    grad_post_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, scale, limit);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}
