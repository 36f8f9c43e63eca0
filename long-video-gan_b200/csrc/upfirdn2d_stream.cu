// Register-streaming separable upfirdn2d for the 2x resampling signatures of the low-res networks
// (U2-U5 -- upsample2d / downsample2d with the 4-tap filter of
// model/generator_lres.py:86-128 and model/discriminator_lres.py:42-86, and their adjoints).
//
// The tiled kernel (upfirdn2d_tiled.cu) stages the input tile and the x-pass result in shared memory;
// for 4-tap filters that costs far more instructions than arithmetic (zero fill, scatter, two barriers,
// per-CTA index setup). Here nothing goes through shared memory:
//
//   * a thread owns a strip of NI consecutive input columns of one (n, c) plane (NI = 4, or 8 when the
//     x axis down-samples) and walks down the rows;
//   * per input row it loads its strip with 128-bit loads, gets the one-sample halos from the
//     neighbouring lanes with two shuffles (zero at the row ends = the zero padding of the operator),
//     and applies the x filter in registers -> NO outputs of that row;
//   * the y filter runs over a sliding window of the last x-filtered rows held in registers, and every
//     finished output row leaves as 128-bit stores.
//
// HBM sees x once and y once, there are no barriers, and the per-output instruction count is ~10x
// below the tiled kernel's. Rows of the next step are loaded before the current ones are filtered
// (register double buffer) so every thread keeps 2-4 vector loads in flight.
//
// Supported (anything else returns LVG_UNSUPPORTED and the caller uses the tiled kernel):
//   fp32 / fp16, dense NCHW input and output, 4-tap filters, per axis one of
//     ID     no filter, out = in
//     UP2    up = 2, pad0 = 2, out size = 2 * in     (upsample2d of the networks; adjoint of DOWN2)
//     DOWN2  down = 2, pad0 = 1, out size = in / 2   (downsample2d of the networks; adjoint of UP2)
//   and row lengths that split into 1, 2, 4, ... 32 strips (x filtered) or any number of strips (x = ID).

#include "common.cuh"

namespace lvg {
namespace {

enum { K_ID = 0, K_UP2 = 1, K_DOWN2 = 2, K_UP2N = 3 };   // UP2N: UP2 along x with 2-sample strips (x axis only)
constexpr int kF = 4;           // taps per filtered axis
constexpr int kThreads = 256;

template <int KIND> struct Geo;
template <> struct Geo<K_ID>    { static constexpr int NI = 4, NO = 4; };
template <> struct Geo<K_UP2>   { static constexpr int NI = 4, NO = 8; };
template <> struct Geo<K_DOWN2> { static constexpr int NI = 8, NO = 4; };
template <> struct Geo<K_UP2N>  { static constexpr int NI = 2, NO = 4; };   // one 16-byte store per row: lanes write contiguously

struct StreamParams {
    const void* x;
    void* y;
    const float* fx;
    const float* fy;
    int64_t fsx, fsy;
    int flip;
    float gain;
    int64_t planes;
    int ih, iw, oh, ow;
    int strips;          // threads per row
    int seg_rows;        // output rows per y segment (even)
};

// N consecutive elements -> fp32 registers (N = 4 or 8), 16-byte accesses where the type allows
template <class T, int N> __device__ __forceinline__ void load_strip(const T* p, float (&v)[N]);
template <> __device__ __forceinline__ void load_strip<float, 2>(const float* p, float (&v)[2])
{
    const float2 a = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = a.x; v[1] = a.y;
}
template <> __device__ __forceinline__ void load_strip<__half, 2>(const __half* p, float (&v)[2])
{
    const unsigned a = __ldg(reinterpret_cast<const unsigned*>(p));
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&a));
    v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void load_strip<float, 4>(const float* p, float (&v)[4])
{
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load_strip<float, 8>(const float* p, float (&v)[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load_strip<__half, 4>(const __half* p, float (&v)[4])
{
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
}
template <> __device__ __forceinline__ void load_strip<__half, 8>(const __half* p, float (&v)[8])
{
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; k++) { const float2 t = __half22float2(h[k]); v[2 * k] = t.x; v[2 * k + 1] = t.y; }
}

template <class T, int N> __device__ __forceinline__ void store_strip(T* p, const float (&v)[N]);
template <> __device__ __forceinline__ void store_strip<float, 4>(float* p, const float (&v)[4])
{
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store_strip<float, 8>(float* p, const float (&v)[8])
{
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store_strip<__half, 4>(__half* p, const float (&v)[4])
{
    uint2 a;
    *reinterpret_cast<__half2*>(&a.x) = __floats2half2_rn(v[0], v[1]);
    *reinterpret_cast<__half2*>(&a.y) = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = a;
}
template <> __device__ __forceinline__ void store_strip<__half, 8>(__half* p, const float (&v)[8])
{
    uint4 a;
    __half2* h = reinterpret_cast<__half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; k++) h[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
    *reinterpret_cast<uint4*>(p) = a;
}

// x filter of one row strip. `a` = the thread's NI inputs, hl / hr = the sample left / right of the strip
// (0 beyond the row), g = taps oriented for correlation.
//   UP2   (pad0 2): out[2q] = g0 in[q-1] + g2 in[q],   out[2q+1] = g1 in[q] + g3 in[q+1]
//   DOWN2 (pad0 1): out[c]  = sum_t g[t] in[2c - 1 + t]
template <int KIND>
__device__ __forceinline__ void x_filter(const float (&a)[Geo<KIND>::NI], float hl, float hr, const float (&g)[kF], float (&o)[Geo<KIND>::NO])
{
    constexpr int NI = Geo<KIND>::NI;
    if constexpr (KIND == K_ID) {
#pragma unroll
        for (int c = 0; c < NI; c++) o[c] = a[c];
    } else {
        float e[NI + 2];
        e[0] = hl;
#pragma unroll
        for (int c = 0; c < NI; c++) e[c + 1] = a[c];
        e[NI + 1] = hr;
        if constexpr (KIND == K_UP2 || KIND == K_UP2N) {
#pragma unroll
            for (int q = 0; q < NI; q++) {
                o[2 * q] = fmaf(g[2], e[q + 1], g[0] * e[q]);
                o[2 * q + 1] = fmaf(g[3], e[q + 2], g[1] * e[q + 1]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < NI / 2; c++) {
                float acc = g[0] * e[2 * c];
#pragma unroll
                for (int t = 1; t < kF; t++) acc = fmaf(g[t], e[2 * c + t], acc);
                o[c] = acc;
            }
        }
    }
}

template <class T, int KX, int KY>
__global__ void __launch_bounds__(kThreads) upfirdn2d_stream_kernel(StreamParams p)
{
    constexpr int NI = Geo<KX>::NI, NO = Geo<KX>::NO;
    // thread -> (plane, strip); blockIdx.y -> segment of output rows
    const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t plane = gid / p.strips;
    const int strip = (int)(gid - plane * p.strips);
    const bool active = plane < p.planes;
    const int o_b = blockIdx.y * p.seg_rows;
    const int o_e = min(o_b + p.seg_rows, p.oh);

    float gx[kF], gy[kF];
#pragma unroll
    for (int t = 0; t < kF; t++) {
        gx[t] = (KX == K_ID) ? 0.f : __ldg(p.fx + (p.flip ? t : kF - 1 - t) * p.fsx);
        gy[t] = (KY == K_ID) ? 0.f : __ldg(p.fy + (p.flip ? t : kF - 1 - t) * p.fsy);
    }
    // the gain rides on the taps of one filtered axis
    if (KY != K_ID) {
#pragma unroll
        for (int t = 0; t < kF; t++) gy[t] *= p.gain;
    } else {
#pragma unroll
        for (int t = 0; t < kF; t++) gx[t] *= p.gain;
    }

    const T* xp = (const T*)p.x + (active ? plane : 0) * (int64_t)p.ih * p.iw + strip * NI;
    T* yp = (T*)p.y + (active ? plane : 0) * (int64_t)p.oh * p.ow + strip * NO;
    const bool first = strip == 0, last = strip == p.strips - 1;

    // raw input row i of this strip (zeros outside the image)
    auto load_row = [&](int i, float (&raw)[NI]) {
        if (active && (unsigned)i < (unsigned)p.ih) {
            load_strip<T, NI>(xp + (int64_t)i * p.iw, raw);
        } else {
#pragma unroll
            for (int c = 0; c < NI; c++) raw[c] = 0.f;
        }
    };
    // raw row -> x-filtered row (all lanes of the warp take part in the halo exchange)
    auto x_row = [&](const float (&raw)[NI], float (&m)[NO]) {
        float hl = 0.f, hr = 0.f;
        if constexpr (KX != K_ID) {
            hl = __shfl_up_sync(0xffffffffu, raw[NI - 1], 1);
            hr = __shfl_down_sync(0xffffffffu, raw[0], 1);
            if (first) hl = 0.f;
            if (last) hr = 0.f;
        }
        x_filter<KX>(raw, hl, hr, gx, m);
    };
    auto emit = [&](int o, const float (&v)[NO]) {
        if (active && o >= o_b && o < o_e) store_strip<T, NO>(yp + (int64_t)o * p.ow, v);
    };

    if constexpr (KY == K_UP2) {
        // input row i completes output rows 2i - 1 = g1 m[i-1] + g3 m[i] and 2i = g0 m[i-1] + g2 m[i]
        const int i_lo = o_b / 2, i_hi = o_e / 2;          // o_b, o_e even
        float prev[NO], r0[NI], r1[NI], n0[NI], n1[NI];
        load_row(i_lo - 1, r0);
        load_row(i_lo, n0);
        load_row(i_lo + 1, n1);
        x_row(r0, prev);
        auto step = [&](int i, const float (&raw)[NI]) {
            float cur[NO], out[NO];
            x_row(raw, cur);
#pragma unroll
            for (int c = 0; c < NO; c++) out[c] = fmaf(gy[3], cur[c], gy[1] * prev[c]);
            emit(2 * i - 1, out);
#pragma unroll
            for (int c = 0; c < NO; c++) out[c] = fmaf(gy[2], cur[c], gy[0] * prev[c]);
            emit(2 * i, out);
#pragma unroll
            for (int c = 0; c < NO; c++) prev[c] = cur[c];
        };
        for (int i = i_lo; i <= i_hi; i += 2) {            // two input rows per trip, the next two already in flight
#pragma unroll
            for (int c = 0; c < NI; c++) { r0[c] = n0[c]; r1[c] = n1[c]; }
            if (i + 2 <= i_hi) load_row(i + 2, n0);
            if (i + 3 <= i_hi) load_row(i + 3, n1);
            step(i, r0);
            if (i + 1 <= i_hi) step(i + 1, r1);
        }
    } else if constexpr (KY == K_DOWN2) {
        // out[o] = sum_t g[t] m[2o - 1 + t]; per step two new rows enter the window of four
        float w0[NO], w1[NO], ra[NI], rb[NI], na[NI], nb[NI];
        load_row(2 * o_b - 1, ra);
        load_row(2 * o_b, rb);
        load_row(2 * o_b + 1, na);
        load_row(2 * o_b + 2, nb);
        x_row(ra, w0);
        x_row(rb, w1);
        for (int o = o_b; o < o_e; o++) {
#pragma unroll
            for (int c = 0; c < NI; c++) { ra[c] = na[c]; rb[c] = nb[c]; }
            if (o + 1 < o_e) {
                load_row(2 * o + 3, na);
                load_row(2 * o + 4, nb);
            }
            float w2[NO], w3[NO], out[NO];
            x_row(ra, w2);
            x_row(rb, w3);
#pragma unroll
            for (int c = 0; c < NO; c++) out[c] = fmaf(gy[3], w3[c], fmaf(gy[2], w2[c], fmaf(gy[1], w1[c], gy[0] * w0[c])));
            emit(o, out);
#pragma unroll
            for (int c = 0; c < NO; c++) { w0[c] = w2[c]; w1[c] = w3[c]; }
        }
    } else {
        // x-only filter: rows are independent
        float raw[NI], nxt[NI];
        load_row(o_b, nxt);
        for (int o = o_b; o < o_e; o++) {
#pragma unroll
            for (int c = 0; c < NI; c++) raw[c] = nxt[c];
            if (o + 1 < o_e) load_row(o + 1, nxt);
            float m[NO];
            x_row(raw, m);
            emit(o, m);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// 12-tap 2x resampling along the contiguous axis only (the Kaiser filters of TemporalKaiserDownsample,
// model/generator_lres.py:219-262, applied to [N, C, L, 1] tensors that reach this file as [N*C, 1, L] rows, and
// their adjoints). Rows are short and there are many of them: a thread produces one 16-byte (DOWN2) or 32-byte (UP2)
// piece of an output row straight from 3-6 aligned vector loads of the input row; neighbouring threads re-read the
// halo through L1. No shared memory, no shuffles, every vector is either entirely inside the row or entirely padding.
//   DOWN2 (pad0 5): out[o] = sum_t g[t] in[2 o - 5 + t]
//   UP2   (pad0 6): out[2 q] = sum_k g[2 k] in[q + k - 3],  out[2 q + 1] = sum_k g[2 k + 1] in[q + k - 2]
constexpr int kF12 = 12;

struct Row12Params {
    const void* x;
    void* y;
    const float* f;
    int64_t fs;
    int flip;
    float gain;
    int64_t rows;        // planes * height
    int iw, ow, strips;
};

template <class T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]) { load_strip<T, 4>(p, v); }

template <class T, bool UP>
__global__ void __launch_bounds__(kThreads) upfirdn2d_row12_kernel(Row12Params p)
{
    const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t row = gid / p.strips;
    const int s = (int)(gid - row * p.strips);
    if (row >= p.rows) return;
    float g[kF12];
#pragma unroll
    for (int t = 0; t < kF12; t++) g[t] = __ldg(p.f + (p.flip ? t : kF12 - 1 - t) * p.fs) * p.gain;
    const T* xr = (const T*)p.x + row * p.iw;
    T* yr = (T*)p.y + row * p.ow;
    constexpr int NV = UP ? 3 : 6;                       // aligned 4-element vectors that cover the taps of this thread
    const int i0 = UP ? 4 * s - 4 : 8 * s - 8;           // input index of e[0]
    float e[4 * NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        const int i = i0 + 4 * v;
        if (i >= 0 && i < p.iw) load4<T>(xr + i, q);      // iw % 4 == 0: a vector never straddles the row end
#pragma unroll
        for (int k = 0; k < 4; k++) e[4 * v + k] = q[k];
    }
    if (UP) {
        float o[8];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < kF12 / 2; k++) {
                a0 = fmaf(g[2 * k], e[c + k + 1], a0);
                a1 = fmaf(g[2 * k + 1], e[c + k + 2], a1);
            }
            o[2 * c] = a0; o[2 * c + 1] = a1;
        }
        store_strip<T, 8>(yr + 8 * s, o);
    } else {
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float a = 0.f;
#pragma unroll
            for (int t = 0; t < kF12; t++) a = fmaf(g[t], e[2 * c + t + 3], a);
            o[c] = a;
        }
        store_strip<T, 4>(yr + 4 * s, o);
    }
}

template <class T>
int launch_row12(Row12Params& p, bool up, cudaStream_t s)
{
    const int64_t threads = p.rows * p.strips;
    const int64_t blocks = (threads + kThreads - 1) / kThreads;
    if (blocks > INT32_MAX) return LVG_UNSUPPORTED;
    if (up) upfirdn2d_row12_kernel<T, true><<<(unsigned)blocks, kThreads, 0, s>>>(p);
    else    upfirdn2d_row12_kernel<T, false><<<(unsigned)blocks, kThreads, 0, s>>>(p);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

int classify(bool has_filter, int taps, int up, int down, int pad0, int in, int out)
{
    if (!has_filter) return (up == 1 && down == 1 && pad0 == 0 && out == in) ? K_ID : -1;
    if (taps != kF) return -1;
    if (up == 2 && down == 1 && pad0 == 2 && out == 2 * in) return K_UP2;
    if (up == 1 && down == 2 && pad0 == 1 && in % 2 == 0 && out == in / 2) return K_DOWN2;
    return -1;
}

template <class T, int KX, int KY>
int launch(StreamParams& p, cudaStream_t s)
{
    constexpr int NI = Geo<KX>::NI;
    if (p.iw % NI != 0) return LVG_UNSUPPORTED;
    p.strips = p.iw / NI;
    if (KX != K_ID && (p.strips > 32 || (p.strips & (p.strips - 1)) != 0)) return LVG_UNSUPPORTED;   // halo exchange stays inside a warp
    const int64_t threads = p.planes * p.strips;
    // split the rows into segments until the grid fills the machine (each segment recomputes <= 2 halo rows)
    const int64_t want = (int64_t)num_sms() * 2048;
    int nseg = (int)((want + threads - 1) / threads);
    const int max_seg = p.oh / 8 > 0 ? p.oh / 8 : 1;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    p.seg_rows = ((p.oh + nseg - 1) / nseg + 1) & ~1;
    nseg = (p.oh + p.seg_rows - 1) / p.seg_rows;
    const int64_t blocks = (threads + kThreads - 1) / kThreads;
    if (blocks > INT32_MAX || nseg > 65535) return LVG_UNSUPPORTED;
    upfirdn2d_stream_kernel<T, KX, KY><<<dim3((unsigned)blocks, (unsigned)nseg), kThreads, 0, s>>>(p);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

template <class T>
int dispatch(int kx, int ky, StreamParams& p, cudaStream_t s)
{
    if (kx == K_UP2 && ky == K_UP2) return p.iw <= 64 ? launch<T, K_UP2N, K_UP2>(p, s) : launch<T, K_UP2, K_UP2>(p, s);
    if (kx == K_DOWN2 && ky == K_DOWN2) return launch<T, K_DOWN2, K_DOWN2>(p, s);
    if (kx == K_ID && ky == K_UP2) return launch<T, K_ID, K_UP2>(p, s);
    if (kx == K_ID && ky == K_DOWN2) return launch<T, K_ID, K_DOWN2>(p, s);
    if (kx == K_UP2 && ky == K_ID) return p.iw <= 64 ? launch<T, K_UP2N, K_ID>(p, s) : launch<T, K_UP2, K_ID>(p, s);
    if (kx == K_DOWN2 && ky == K_ID) return launch<T, K_DOWN2, K_ID>(p, s);
    return LVG_UNSUPPORTED;
}

}  // namespace

// same contract as upfirdn2d_tiled(); LVG_UNSUPPORTED = not one of the streamed signatures
int upfirdn2d_stream(const void* x, const float* fx, int64_t fsx, const float* fy, int64_t fsy, void* y, int dtype,
                     const int64_t* xsh, const int64_t* xst, const int64_t* ysh, const int64_t* yst,
                     int fw, int fh, int upx, int upy, int downx, int downy, int padx0, int pady0,
                     int flip, float gain, cudaStream_t s)
{
    static const bool enabled = [] { const char* e = getenv("LVG_UPFIRDN_STREAM"); return !(e && e[0] == '0'); }();
    if (!enabled) return LVG_UNSUPPORTED;
    if (dtype != LVG_F32 && dtype != LVG_F16) return LVG_UNSUPPORTED;
    const int64_t n = xsh[0], c = xsh[1], ih = xsh[2], iw = xsh[3], oh = ysh[2], ow = ysh[3];
    if (n * c < 1 || ih < 1 || iw < 1 || ih > (1 << 24) || iw > (1 << 24) || oh > (1 << 24) || ow > (1 << 24)) return LVG_UNSUPPORTED;
    // dense NCHW on both sides
    // (the stride of a size-1 dimension is arbitrary: [N, C, L, 1] tensors arrive here as the view [N, C, 1, L])
    if (xst[3] != 1 || (ih > 1 && xst[2] != iw) || (c > 1 && xst[1] != ih * iw) || (n > 1 && xst[0] != c * ih * iw)) return LVG_UNSUPPORTED;
    if (yst[3] != 1 || (oh > 1 && yst[2] != ow) || (c > 1 && yst[1] != oh * ow) || (n > 1 && yst[0] != c * oh * ow)) return LVG_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(y)) return LVG_UNSUPPORTED;
    // 12 taps along x only (temporal Kaiser resampling on the transposed [N*C, 1, L] view, or any [.., H, W] with a [1, 12] filter)
    if (fx != nullptr && fy == nullptr && fw == kF12 && fh == 1 && upy == 1 && downy == 1 && pady0 == 0 && oh == ih) {
        const bool up = upx == 2 && downx == 1 && padx0 == 6 && ow == 2 * iw;
        const bool down = upx == 1 && downx == 2 && padx0 == 5 && iw % 2 == 0 && ow == iw / 2;
        const int vec = dtype == LVG_F16 ? 8 : 4;       // the 8-byte fp16 loads of the strip helpers need 4-element alignment, stores up to 8
        if ((up || down) && iw % 4 == 0 && ow % 4 == 0 && (dtype == LVG_F32 || (up ? ow % vec == 0 : true))) {
            Row12Params q;
            q.x = x; q.y = y; q.f = fx; q.fs = fsx; q.flip = flip ? 1 : 0; q.gain = gain;
            q.rows = n * c * ih; q.iw = (int)iw; q.ow = (int)ow; q.strips = up ? (int)iw / 4 : (int)ow / 4;
            return dtype == LVG_F32 ? launch_row12<float>(q, up, s) : launch_row12<__half>(q, up, s);
        }
    }
    const int kx = classify(fx != nullptr, fw, upx, downx, padx0, (int)iw, (int)ow);
    const int ky = classify(fy != nullptr, fh, upy, downy, pady0, (int)ih, (int)oh);
    if (kx < 0 || ky < 0 || (kx == K_ID && ky == K_ID)) return LVG_UNSUPPORTED;
    if (ow % 4 != 0) return LVG_UNSUPPORTED;      // vector stores of whole strips
    StreamParams p;
    p.x = x; p.y = y; p.fx = fx; p.fy = fy; p.fsx = fsx; p.fsy = fsy;
    p.flip = flip ? 1 : 0; p.gain = gain;
    p.planes = n * c; p.ih = (int)ih; p.iw = (int)iw; p.oh = (int)oh; p.ow = (int)ow;
    return dtype == LVG_F32 ? dispatch<float>(kx, ky, p, s) : dispatch<__half>(kx, ky, p, s);
}

}  // namespace lvg
