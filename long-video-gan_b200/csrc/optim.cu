// Fused optimiser step over flat fp32 buffers (the step overheads around the operator path).
//
// The reference ends every update with `utils.sync_grads` (a scale + nan_to_num pass, utils.py:116-124),
// `torch.optim.Adam.step()` over ~300 parameter tensors (model/video_gan_lres.py:84-85,128,174) and, for the generator,
// `tensor_ema.lerp_(tensor, 1 - ema_beta)` over every parameter and buffer (video_gan_lres.py:208-214): three to four
// sweeps over the parameter-sized buffers, hundreds of launches in the per-tensor form. With parameters, gradients
// and moments kept as flat buffers (lvg_dist/flat_optim.py) the whole tail of an update is ONE streaming kernel:
//   g' = nan_to_num(g * grad_scale)                       (optional: grad_limit > 0)
//   m  = lerp(m, g', 1 - beta1);  v = v * beta2 + (1 - beta2) * g' * g'
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)      (torch.optim.Adam, no amsgrad / weight decay / maximize)
//   p_ema = lerp(p_ema, p, 1 - ema_beta)                   (optional)
// HBM-bound: 16 B read + 12 B written per element (+ 8 with the EMA).
#include "common.cuh"

namespace lvg {
namespace {

struct AdamArgs {
    float lr_over_bc1, beta1, beta2, eps, inv_sqrt_bc2, grad_scale, grad_limit, ema_w;
};

// torch.lerp: weight < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
__device__ __forceinline__ float lerp_torch(float a, float b, float w)
{
    const float d = b - a;
    return w < 0.5f ? fmaf(w, d, a) : b - d * (1.f - w);
}

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float* pe, const AdamArgs& a)
{
    if (a.grad_limit > 0.f) {
        g *= a.grad_scale;
        if (g != g) g = 0.f;
        else if (isinf(g)) g = g > 0.f ? a.grad_limit : -a.grad_limit;
    }
    m = lerp_torch(m, g, 1.f - a.beta1);
    v = fmaf(v, a.beta2, (1.f - a.beta2) * g * g);
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p = p - a.lr_over_bc1 * (m / denom);
    if (pe) *pe = lerp_torch(*pe, p, a.ema_w);
}

template <bool EMA>
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ pe, int64_t n, AdamArgs a,
                                                        int write_grad)
{
    const bool vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!EMA || aligned16(pe));
    const int64_t n4 = vec ? n / 4 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        float4 E = EMA ? reinterpret_cast<float4*>(pe)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        adam_one(P.x, G.x, M.x, V.x, EMA ? &E.x : nullptr, a);
        adam_one(P.y, G.y, M.y, V.y, EMA ? &E.y : nullptr, a);
        adam_one(P.z, G.z, M.z, V.z, EMA ? &E.z : nullptr, a);
        adam_one(P.w, G.w, M.w, V.w, EMA ? &E.w : nullptr, a);
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        if (write_grad) reinterpret_cast<float4*>(g)[i] = G;
        if (EMA) reinterpret_cast<float4*>(pe)[i] = E;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float P = p[i], G = g[i], M = m[i], V = v[i], E = EMA ? pe[i] : 0.f;
        adam_one(P, G, M, V, EMA ? &E : nullptr, a);
        p[i] = P; m[i] = M; v[i] = V;
        if (write_grad) g[i] = G;
        if (EMA) pe[i] = E;
    }
}

__global__ void __launch_bounds__(256) lerp_kernel(float* __restrict__ a, const float* __restrict__ b, int64_t n, float w)
{
    const bool vec = aligned16(a) && aligned16(b);
    const int64_t n4 = vec ? n / 4 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 A = reinterpret_cast<float4*>(a)[i];
        const float4 B = reinterpret_cast<const float4*>(b)[i];
        A.x = lerp_torch(A.x, B.x, w); A.y = lerp_torch(A.y, B.y, w); A.z = lerp_torch(A.z, B.z, w); A.w = lerp_torch(A.w, B.w, w);
        reinterpret_cast<float4*>(a)[i] = A;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = lerp_torch(a[i], b[i], w);
}

unsigned stream_grid(int64_t n)
{
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8 * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int lvg_adam_step(float* p, float* g, float* m, float* v, float* p_ema, int64_t n, float lr, float beta1, float beta2,
                             float eps, int64_t step, float grad_scale, float grad_limit, int write_grad, float ema_beta,
                             void* stream)
{
    LVG_REQUIRE(n >= 0 && (n == 0 || (p && g && m && v)), "adam_step: p, g, m, v must not be NULL");
    LVG_REQUIRE(step >= 1, "adam_step: step counts from 1");
    LVG_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_step: bad hyper-parameters");
    if (n == 0) return LVG_OK;
    // bias corrections in double like the Python scalars of torch.optim.Adam (1 - beta ** step)
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    AdamArgs a;
    a.lr_over_bc1 = (float)((double)lr / bc1);
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    a.grad_scale = grad_scale; a.grad_limit = grad_limit;
    a.ema_w = 1.f - ema_beta;
    cudaStream_t s = (cudaStream_t)stream;
    if (p_ema) adam_step_kernel<true><<<stream_grid(n), 256, 0, s>>>(p, g, m, v, p_ema, n, a, write_grad);
    else       adam_step_kernel<false><<<stream_grid(n), 256, 0, s>>>(p, g, m, v, nullptr, n, a, write_grad);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

extern "C" int lvg_lerp(float* a, const float* b, int64_t n, float weight, void* stream)
{
    LVG_REQUIRE(n >= 0 && (n == 0 || (a && b)), "lerp: buffers must not be NULL");
    if (n == 0) return LVG_OK;
    lerp_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(a, b, n, weight);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}
