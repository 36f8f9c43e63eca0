// bias_act: y = clamp(act(x + b) * gain) and its first / second derivatives.
//
// Semantics follow the reference kernel (torch_utils/ops/bias_act.cu:23-147):
// grad=0 evaluates the activation, grad=1 scales an incoming gradient by
// act'(.) expressed through the saved output yref (or the saved input xref for
// swish), grad=2 is the second derivative; clamping saturates in the forward
// and zeroes gradients where the forward output was saturated.
//
// B200 design: a pure streaming op (1 read + 1 write forward, 2 reads + 1 write
// backward), so the kernel is organised around bytes in flight: 128-bit
// loads/stores, four independent packs per thread issued before any use
// (64 KB in flight per SM at 4 CTAs/SM), grid sized in whole waves of the SM
// count, one bias lookup per 16-byte pack instead of a div+mod per element.
// The bias-gradient reduction can be fused into the backward pass
// (lvg_bias_act_grad_db) which removes the separate full read of dx.

#include "common.cuh"

namespace lvg {
namespace {

struct BiasActParams {
    const void* x;
    const void* b;
    const void* xref;
    const void* yref;
    const void* dy;
    void* y;
    float* db;        // fused bias-gradient accumulators (grad=1 only) or NULL
    int64_t n;
    int64_t size_b;
    int64_t step_b;
    uint64_t magic_step, magic_size;   // ceil(2^64 / d): exact quotients for operands < 2^32 (0 = divide for real)
    int grad;
    float alpha, gain, clamp;
};

// n / d through a multiply-high when the host could prepare a magic number (n, d < 2^32)
__device__ __forceinline__ int64_t fast_div(int64_t n, int64_t d, uint64_t magic)
{
    if (d == 1) return n;
    return magic ? (int64_t)__umul64hi((uint64_t)n, magic) : n / d;
}

// How the bias index of a 16-byte pack is obtained.
enum BiasMode {
    BIAS_NONE = 0,
    BIAS_PER_PACK = 1,   // step_b is a multiple of the pack width: one index per pack
    BIAS_PACKED = 2,     // step_b == 1 and size_b multiple of the pack width: load a bias pack
    BIAS_PER_ELEM = 3,   // anything else
};

__device__ __forceinline__ int64_t bias_row(int64_t elem, const BiasActParams& p) { return fast_div(elem, p.step_b, p.magic_step); }
__device__ __forceinline__ int64_t bias_index(int64_t elem, const BiasActParams& p)
{
    const int64_t row = bias_row(elem, p);
    return row - fast_div(row, p.size_b, p.magic_size) * p.size_b;
}

template <class S> __device__ __forceinline__ S fexp(S v);
template <> __device__ __forceinline__ float fexp<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double fexp<double>(double v) { return exp(v); }
template <class S> __device__ __forceinline__ S flog(S v);
template <> __device__ __forceinline__ float flog<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double flog<double>(double v) { return log(v); }
template <class S> __device__ __forceinline__ S ftanh(S v);
template <> __device__ __forceinline__ float ftanh<float>(float v) { return tanhf(v); }
template <> __device__ __forceinline__ double ftanh<double>(double v) { return tanh(v); }

// One element. `v` is x (grad 0) or the incoming gradient (grad>0); xr = xref + b;
// yr = yref; dyv = dy (grad 2) or 1.
template <class S, int A>
__device__ __forceinline__ S bias_act_elem(S v, S xr, S yr, S dyv, int G, S alpha, S gain, S inv_gain, S clamp)
{
    const S one = (S)1, two = (S)2;
    const S kExpRange = (S)80, kHalfExpRange = (S)40;
    const S kSeluScale = (S)1.0507009873554804934193349852946;
    const S kSeluAlpha = (S)1.6732632423543772848170429916717;
    const S r = yr * inv_gain;   // activation output before gain (inv_gain = 1/gain, 0 when gain is 0)
    S out = (S)0;

    if (A == LVG_ACT_LINEAR) {
        out = (G <= 1) ? v : (S)0;
    } else if (A == LVG_ACT_RELU) {
        if (G == 0) out = v > (S)0 ? v : (S)0;
        else if (G == 1) out = r > (S)0 ? v : (S)0;
    } else if (A == LVG_ACT_LRELU) {
        if (G == 0) out = v > (S)0 ? v : v * alpha;
        else if (G == 1) out = r > (S)0 ? v : v * alpha;
    } else if (A == LVG_ACT_TANH) {
        if (G == 0) out = ftanh(v);
        else if (G == 1) out = v * (one - r * r);
        else out = v * (one - r * r) * (-two * r);
    } else if (A == LVG_ACT_SIGMOID) {
        if (G == 0) out = (v < -kExpRange) ? (S)0 : one / (fexp(-v) + one);
        else if (G == 1) out = v * r * (one - r);
        else out = v * r * (one - r) * (one - two * r);
    } else if (A == LVG_ACT_ELU) {
        if (G == 0) out = (v >= (S)0) ? v : fexp(v) - one;
        else if (G == 1) out = (r >= (S)0) ? v : v * (r + one);
        else out = (r >= (S)0) ? (S)0 : v * (r + one);
    } else if (A == LVG_ACT_SELU) {
        if (G == 0) out = (v >= (S)0) ? kSeluScale * v : (kSeluScale * kSeluAlpha) * (fexp(v) - one);
        else if (G == 1) out = (r >= (S)0) ? v * kSeluScale : v * (r + kSeluScale * kSeluAlpha);
        else out = (r >= (S)0) ? (S)0 : v * (r + kSeluScale * kSeluAlpha);
    } else if (A == LVG_ACT_SOFTPLUS) {
        if (G == 0) out = (v > kExpRange) ? v : flog(fexp(v) + one);
        else if (G == 1) out = v * (one - fexp(-r));
        else { S c = fexp(-r); out = v * c * (one - c); }
    } else if (A == LVG_ACT_SWISH) {
        if (G == 0) {
            out = (v < -kExpRange) ? (S)0 : v / (fexp(-v) + one);
        } else {
            S c = fexp(xr);
            S d = c + one;
            if (G == 1) out = (xr > kHalfExpRange) ? v : v * c * (xr + d) / (d * d);
            else        out = (xr > kHalfExpRange) ? (S)0 : v * c * (xr * (two - d) + two * d) / (d * d * d);
            // swish keeps x, not y: rebuild the forward output for the clamp mask
            yr = (xr < -kExpRange) ? (S)0 : xr / (fexp(-xr) + one) * gain;
        }
    }

    out *= gain * dyv;

    if (clamp >= (S)0) {
        if (G == 0) out = (out > -clamp && out < clamp) ? out : (out >= (S)0 ? clamp : -clamp);
        else        out = (yr > -clamp && yr < clamp) ? out : (S)0;
    }
    return out;
}

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <class T>
__device__ __forceinline__ void fetch_bias(typename Acc<T>::type (&bias)[VecOf<T>::N], int64_t& bidx,
                                           const T* __restrict__ pb, int bmode, int64_t e0, const BiasActParams& p)
{
    typedef typename Acc<T>::type S;
    constexpr int N = VecOf<T>::N;
    bidx = 0;
    if (bmode == BIAS_PER_PACK) {
        bidx = bias_index(e0, p);
        S bv = pb ? to_acc(pb[bidx]) : (S)0;
#pragma unroll
        for (int k = 0; k < N; k++) bias[k] = bv;
    } else if (bmode == BIAS_PACKED) {
        bidx = e0 - fast_div(e0, p.size_b, p.magic_size) * p.size_b;
        Pack<T> vb;
        if (pb) vb = load_pack(pb + bidx);
#pragma unroll
        for (int k = 0; k < N; k++) bias[k] = pb ? to_acc(vb.v[k]) : (S)0;
    } else if (bmode == BIAS_PER_ELEM) {
#pragma unroll
        for (int k = 0; k < N; k++) bias[k] = pb ? to_acc(pb[bias_index(e0 + k, p)]) : (S)0;
    } else {
#pragma unroll
        for (int k = 0; k < N; k++) bias[k] = (S)0;
    }
}

// Vector kernel: n_pack packs of VecOf<T>::N elements, all operands 16-byte aligned.
// G is a template parameter so that grad 0 carries no reference operands at all.
template <class T, int A, int G, bool FUSE_DB>
__global__ void __launch_bounds__(kThreads, (G == 2 ? 2 : 4)) bias_act_vec_kernel(BiasActParams p, int64_t n_pack, int bmode)
{
    typedef typename Acc<T>::type S;
    constexpr int N = VecOf<T>::N;
    constexpr bool kUseX = (G > 0) && (A == LVG_ACT_SWISH);       // saved input
    constexpr bool kUseY = (G > 0) && (A != LVG_ACT_SWISH);       // saved output
    constexpr bool kUseDy = (G == 2);
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const S inv_gain = gain != (S)0 ? (S)1 / gain : (S)0;
    const T* __restrict__ px = (const T*)p.x;
    const T* __restrict__ pb = (const T*)p.b;
    const T* __restrict__ pxr = kUseX ? (const T*)p.xref : nullptr;
    const T* __restrict__ pyr = kUseY ? (const T*)p.yref : nullptr;
    const T* __restrict__ pdy = kUseDy ? (const T*)p.dy : nullptr;
    T* __restrict__ py = (T*)p.y;

    // Fused bias gradient (FUSE_DB): the tiles are walked in RUNS of kRun consecutive tiles (128 KB of each operand),
    // runs interleaved over the CTAs like single tiles are in the forward pass -- concurrently resident CTAs stream
    // one contiguous window of memory, which HBM rewards (a fully contiguous per-CTA range measured 0.75 of the copy
    // rate, single interleaved tiles with one atomic per warp and tile 0.5: ~100 consecutive tiles share a bias row
    // and their atomics serialise on one address). Inside a run a warp stays inside one channel ("row" = run of
    // step_b elements sharing a bias index) for many packs: each lane keeps a running sum for the warp's current
    // row; only when the row changes or the run ends the warp reduces by shuffle and issues ONE global atomic.
    // No shared memory, no block barriers in the streaming loop.
    constexpr int kRun = FUSE_DB ? 8 : 1;
    const int64_t tile = (int64_t)kThreads * kUnroll;
    const int64_t n_tiles = (n_pack + tile - 1) / tile;
    float run_sum = 0.f;           // this lane's share of the warp's current row
    int64_t run_row = -1, run_idx = 0;
    const unsigned full = 0xffffffffu;
    auto warp_flush = [&]() {
        float s = run_sum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
        if ((threadIdx.x & 31) == 0 && run_row >= 0) atomicAdd(p.db + run_idx, s);
        run_sum = 0.f;
    };

    for (int64_t run = blockIdx.x; run * kRun < n_tiles; run += gridDim.x) {
    const int64_t t_end = (run + 1) * kRun < n_tiles ? (run + 1) * kRun : n_tiles;
    for (int64_t t = run * kRun; t < t_end; t++) {
        const int64_t base = t * tile;
        Pack<T> vx[kUnroll], vref[kUnroll], vdy[kUnroll];
        const T* __restrict__ pref = kUseX ? pxr : pyr;
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int64_t pk = base + (int64_t)u * kThreads + threadIdx.x;
            if (pk < n_pack) {
                vx[u] = load_pack(px + pk * N);
                if ((kUseX || kUseY) && pref) vref[u] = load_pack(pref + pk * N);
                if (kUseDy && pdy) vdy[u] = load_pack(pdy + pk * N);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const int64_t pk = base + (int64_t)u * kThreads + threadIdx.x;
            int64_t db_row = -1, db_idx = 0;      // fused bias gradient of this pack: row, bias index, value
            float db_val = 0.f;
            if (pk < n_pack) {
                const int64_t e0 = pk * N;
                S bias[N];
                int64_t bidx;
                fetch_bias<T>(bias, bidx, pb, bmode, e0, p);

                S fx[N], fref[N], fdy[N], fo[N];
                unpack<T>(vx[u], fx);
                if ((kUseX || kUseY) && pref) unpack<T>(vref[u], fref);
                if (kUseDy && pdy) unpack<T>(vdy[u], fdy);
#pragma unroll
                for (int k = 0; k < N; k++) {
                    S v = fx[k];
                    S xr = (kUseX && pref) ? fref[k] : (S)0;
                    S yr = (kUseY && pref) ? fref[k] : (S)0;
                    S dyv = (kUseDy && pdy) ? fdy[k] : (S)1;
                    if (G == 0) v += bias[k]; else xr += bias[k];
                    fo[k] = bias_act_elem<S, A>(v, xr, yr, dyv, G, alpha, gain, inv_gain, clamp);
                }
                const Pack<T> out = pack<T>(fo);
                store_pack(py + e0, out);
                if (FUSE_DB) {
                    // accumulate what was actually stored, like dx.sum() would see it
                    S fs[N];
                    unpack<T>(out, fs);
                    if (bmode == BIAS_PER_PACK) {
                        S s = (S)0;
#pragma unroll
                        for (int k = 0; k < N; k++) s += fs[k];
                        db_row = bias_row(e0, p); db_idx = bidx; db_val = (float)s;
                    } else {
#pragma unroll
                        for (int k = 0; k < N; k++) atomicAdd(p.db + bias_index(e0 + k, p), (float)fs[k]);
                    }
                }
            }
            if (FUSE_DB && bmode == BIAS_PER_PACK) {
                // lane 0 holds the smallest pack index: if it is past the end, every lane is
                const int64_t row0 = __shfl_sync(full, db_row, 0);
                const int64_t idx0 = __shfl_sync(full, db_idx, 0);
                const bool uniform = __all_sync(full, db_row < 0 || db_row == row0);
                if (uniform) {
                    if (row0 >= 0) {
                        if (row0 != run_row) {          // warp-uniform branch
                            warp_flush();
                            run_row = row0; run_idx = idx0;
                        }
                        run_sum += db_val;              // lanes past the end add 0
                    }
                } else if (db_row >= 0) {
                    atomicAdd(p.db + db_idx, db_val);   // a row boundary runs through this warp's 32 packs
                }
            }
        }
    }
    if (FUSE_DB && bmode == BIAS_PER_PACK) { warp_flush(); run_row = -1; }
    }
}

// Scalar kernel: fp64, unaligned buffers, and the < one-pack tail of the vector kernel.
template <class T, int A, bool FUSE_DB>
__global__ void __launch_bounds__(kThreads) bias_act_scalar_kernel(BiasActParams p, int64_t first)
{
    typedef typename Acc<T>::type S;
    const int G = p.grad;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const S inv_gain = gain != (S)0 ? (S)1 / gain : (S)0;
    for (int64_t i = first + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * kThreads) {
        S v = to_acc(((const T*)p.x)[i]);
        int64_t bidx = (p.b || FUSE_DB) ? bias_index(i, p) : 0;
        S bias = p.b ? to_acc(((const T*)p.b)[bidx]) : (S)0;
        S xr = p.xref ? to_acc(((const T*)p.xref)[i]) : (S)0;
        S yr = p.yref ? to_acc(((const T*)p.yref)[i]) : (S)0;
        S dyv = p.dy ? to_acc(((const T*)p.dy)[i]) : (S)1;
        if (G == 0) v += bias; else xr += bias;
        T o = from_acc<T>(bias_act_elem<S, A>(v, xr, yr, dyv, G, alpha, gain, inv_gain, clamp));
        ((T*)p.y)[i] = o;
        if (FUSE_DB) atomicAdd(p.db + bidx, (float)to_acc(o));
    }
}

template <class T> struct HasVecPath { static constexpr bool value = true; };
template <> struct HasVecPath<double> { static constexpr bool value = false; };   // fp64 is a side path: scalar only

template <class T, int A, bool FUSE_DB>
int launch_typed(const BiasActParams& p, cudaStream_t stream)
{
    const int sms = num_sms();
    int64_t first = 0;

    if constexpr (HasVecPath<T>::value) {
        constexpr int N = VecOf<T>::N;
        const bool vec_ok = aligned16(p.x) && aligned16(p.y) && (!p.xref || aligned16(p.xref)) &&
                            (!p.yref || aligned16(p.yref)) && (!p.dy || aligned16(p.dy));
        const int64_t n_pack = vec_ok ? p.n / N : 0;
        if (n_pack > 0) {
            int mode = BIAS_NONE;
            if (p.b || FUSE_DB) {   // the fused db reduction needs the index even without bias values
                if (p.step_b % N == 0) mode = BIAS_PER_PACK;
                else if (p.step_b == 1 && p.size_b % N == 0 && (!p.b || aligned16(p.b))) mode = BIAS_PACKED;
                else mode = BIAS_PER_ELEM;
            }
            const int64_t tile = (int64_t)kThreads * kUnroll;
            int64_t blocks = (n_pack + tile - 1) / tile;
            // whole waves of 4 CTAs per SM; beyond 8 waves the grid-stride loop takes over
            int64_t cap = (int64_t)sms * 4 * 8;
            if (FUSE_DB) {          // runs of 8 tiles (kRun in the kernel), one resident wave so that every CTA gets ~the same number
                blocks = (blocks + 7) / 8;
                cap = (int64_t)sms * 4;
            }
            if (blocks > cap) blocks = cap;
            void (*k)(BiasActParams, int64_t, int) = nullptr;
            if (FUSE_DB)          k = bias_act_vec_kernel<T, A, 1, FUSE_DB>;
            else if (p.grad == 0) k = bias_act_vec_kernel<T, A, 0, false>;
            else if (p.grad == 1) k = bias_act_vec_kernel<T, A, 1, false>;
            else                  k = bias_act_vec_kernel<T, A, 2, false>;
            k<<<(unsigned)blocks, kThreads, 0, stream>>>(p, n_pack, mode);
            LVG_LAUNCH_CHECK();
        }
        first = n_pack * N;
    }
    if (first < p.n) {
        int64_t rest = p.n - first;
        int64_t blocks = (rest + kThreads - 1) / kThreads;
        const int64_t cap = (int64_t)sms * 8 * 4;
        if (blocks > cap) blocks = cap;
        bias_act_scalar_kernel<T, A, FUSE_DB><<<(unsigned)blocks, kThreads, 0, stream>>>(p, first);
        LVG_LAUNCH_CHECK();
    }
    return LVG_OK;
}

template <class T, bool FUSE_DB>
int launch_act(int act, const BiasActParams& p, cudaStream_t s)
{
    switch (act) {
        case LVG_ACT_LINEAR:   return launch_typed<T, LVG_ACT_LINEAR, FUSE_DB>(p, s);
        case LVG_ACT_RELU:     return launch_typed<T, LVG_ACT_RELU, FUSE_DB>(p, s);
        case LVG_ACT_LRELU:    return launch_typed<T, LVG_ACT_LRELU, FUSE_DB>(p, s);
        case LVG_ACT_TANH:     return launch_typed<T, LVG_ACT_TANH, FUSE_DB>(p, s);
        case LVG_ACT_SIGMOID:  return launch_typed<T, LVG_ACT_SIGMOID, FUSE_DB>(p, s);
        case LVG_ACT_ELU:      return launch_typed<T, LVG_ACT_ELU, FUSE_DB>(p, s);
        case LVG_ACT_SELU:     return launch_typed<T, LVG_ACT_SELU, FUSE_DB>(p, s);
        case LVG_ACT_SOFTPLUS: return launch_typed<T, LVG_ACT_SOFTPLUS, FUSE_DB>(p, s);
        case LVG_ACT_SWISH:    return launch_typed<T, LVG_ACT_SWISH, FUSE_DB>(p, s);
    }
    set_error("bias_act: unknown activation code %d", act);
    return LVG_ERR_ARG;
}

uint64_t magic_for(int64_t d, int64_t n)
{
    if (d <= 1 || d >= (1ll << 32) || n >= (1ll << 32)) return 0;
    return (uint64_t)(~0ull / (uint64_t)d) + 1;      // ceil(2^64 / d) for d that is not a power of two; exact enough otherwise too
}

int check_common(const void* x, const void* y, int dtype, int64_t n, const void* b, int64_t size_b, int64_t step_b, int act)
{
    LVG_REQUIRE(x && y, "bias_act: x and y must not be NULL");
    LVG_REQUIRE(dtype == LVG_F32 || dtype == LVG_F16 || dtype == LVG_F64, "bias_act: unsupported dtype %d", dtype);
    LVG_REQUIRE(n >= 0, "bias_act: negative element count");
    LVG_REQUIRE(!b || (size_b >= 1 && step_b >= 1), "bias_act: bias needs size_b >= 1 and step_b >= 1");
    LVG_REQUIRE(act >= LVG_ACT_LINEAR && act <= LVG_ACT_SWISH, "bias_act: unknown activation code %d", act);
    return LVG_OK;
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int lvg_bias_act(const void* x, const void* b, const void* xref, const void* yref,
                            const void* dy, void* y, int dtype, int64_t n, int64_t size_b,
                            int64_t step_b, int grad, int act, float alpha, float gain,
                            float clamp, void* stream)
{
    int rc = check_common(x, y, dtype, n, b, size_b, step_b, act);
    if (rc) return rc;
    LVG_REQUIRE(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2 (got %d)", grad);
    if (n == 0) return LVG_OK;
    BiasActParams p = {x, b, xref, yref, dy, y, nullptr, n, b ? size_b : 1, b ? step_b : 1, 0, 0, grad, alpha, gain, clamp};
    p.magic_step = magic_for(p.step_b, n); p.magic_size = magic_for(p.size_b, n);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LVG_F32) return launch_act<float, false>(act, p, s);
    if (dtype == LVG_F16) return launch_act<__half, false>(act, p, s);
    return launch_act<double, false>(act, p, s);
}

extern "C" int lvg_bias_act_grad_db(const void* dy_in, const void* b, const void* xref,
                                    const void* yref, void* dx, float* db_f32, int dtype,
                                    int64_t n, int64_t size_b, int64_t step_b, int act,
                                    float alpha, float gain, float clamp, void* stream)
{
    int rc = check_common(dy_in, dx, dtype, n, b, size_b, step_b, act);
    if (rc) return rc;
    LVG_REQUIRE(db_f32 != nullptr, "bias_act_grad_db: db buffer must not be NULL");
    LVG_REQUIRE(size_b >= 1 && step_b >= 1, "bias_act_grad_db: needs size_b >= 1 and step_b >= 1");
    LVG_REQUIRE(dtype == LVG_F32 || dtype == LVG_F16, "bias_act_grad_db: fp32/fp16 only");
    if (n == 0) return LVG_OK;
    // The fused reduction pays off when a 16-byte pack lies within one channel (bias along an outer
    // dimension). With the bias along the contiguous dimension (fully connected layers) every element
    // of a pack belongs to a different channel: leave that reduction to a separate pass.
    if (step_b % (dtype == LVG_F16 ? 8 : 4) != 0) {
        set_error("bias_act_grad_db: bias runs along the contiguous dimension; use bias_act(grad=1) + a reduction");
        return LVG_UNSUPPORTED;
    }
    // b may be NULL here (bias values are only needed by swish); the index math still applies.
    BiasActParams p = {dy_in, b, xref, yref, nullptr, dx, db_f32, n, size_b, step_b, 0, 0, 1, alpha, gain, clamp};
    p.magic_step = magic_for(p.step_b, n); p.magic_size = magic_for(p.size_b, n);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LVG_F32) return launch_act<float, true>(act, p, s);
    return launch_act<__half, true>(act, p, s);
}
