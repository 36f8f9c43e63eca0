// In-place gain * leaky-ReLU * clamp over an up-sampled tensor, with the 2-bit
// sign/clamp codes written (forward) or consumed (backward). This is the
// activation stage of the generic (non-fused) filtered_lrelu path; semantics as
// torch_utils/ops/filtered_lrelu.cu:1105-1211.
//
// Sign tensor: uint8 [N][C][s_h][s_wbytes]; sample (x, y) lives in byte x>>2 of
// row y at bit position 2*(x&3); bit0 = "was negative", bit1 = "was clamped"
// (the clamp code replaces the negative code).
//
// One thread owns one sign byte = 4 horizontally adjacent samples, so sign
// bytes are produced without any cross-lane traffic and stored coalesced.

#include "common.cuh"

namespace lvg {
namespace {

struct ActParams {
    void* x;
    const uint8_t* si;
    uint8_t* so;
    int64_t xs[4];
    int n, c, h, w;
    int s_h, s_wbytes;
    int sx, sy;
    float gain, slope, clamp;
};

enum { SIGN_NONE = 0, SIGN_WRITE = 1, SIGN_READ = 2 };

template <class T, int MODE>
__global__ void __launch_bounds__(256) filtered_lrelu_act_kernel(ActParams p, int64_t total, int wq, int rows)
{
    typedef typename Acc<T>::type S;
    T* __restrict__ x = (T*)p.x;
    const S gain = (S)p.gain, slope = (S)p.slope, clamp = (S)p.clamp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        // idx -> (plane, row, quad)
        const int q = (int)(idx % wq);
        int64_t r = idx / wq;
        const int yy = (int)(r % rows);
        const int64_t plane = r / rows;
        const int cc = (int)(plane % p.c);
        const int nn = (int)(plane / p.c);
        T* row = x + (int64_t)nn * p.xs[0] + (int64_t)cc * p.xs[1] + (int64_t)yy * p.xs[2];
        const int x0 = q * 4;
        unsigned sbyte = 0;

#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = x0 + k;
            if (xx < p.w && yy < p.h) {
                T* pv = row + (int64_t)xx * p.xs[3];
                S v = to_acc(*pv) * gain;
                if (MODE == SIGN_READ) {
                    const int64_t ssx = (int64_t)xx + p.sx;
                    const int64_t ssy = (int64_t)yy + p.sy;
                    if (ssx >= 0 && ssx < (int64_t)p.s_wbytes * 4 && ssy >= 0 && ssy < p.s_h) {
                        const unsigned s = p.si[(plane * p.s_h + ssy) * p.s_wbytes + (ssx >> 2)] >> ((ssx & 3) << 1);
                        if (s & 1u) v *= slope;
                        if (s & 2u) v = (S)0;
                    }
                } else {
                    unsigned code = 0;
                    if (v < (S)0) { v *= slope; code = 1; }
                    if (fabs(v) > clamp) { v = v < (S)0 ? -clamp : clamp; code = 2; }
                    sbyte |= code << (2 * k);
                }
                *pv = from_acc<T>(v);
            }
        }
        if (MODE == SIGN_WRITE) {
            if (q < p.s_wbytes && yy < p.s_h)
                p.so[(plane * p.s_h + yy) * p.s_wbytes + q] = (uint8_t)sbyte;
        }
    }
}

template <class T>
int launch_act(const ActParams& p, int mode, cudaStream_t s)
{
    // cover the larger of the data domain and (when writing) the sign domain
    int wq = (p.w + 3) / 4;
    int rows = p.h;
    if (mode == SIGN_WRITE) {
        if (p.s_wbytes > wq) wq = p.s_wbytes;
        if (p.s_h > rows) rows = p.s_h;
    }
    const int64_t total = (int64_t)p.n * p.c * rows * wq;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (mode == SIGN_WRITE)     filtered_lrelu_act_kernel<T, SIGN_WRITE><<<(unsigned)blocks, 256, 0, s>>>(p, total, wq, rows);
    else if (mode == SIGN_READ) filtered_lrelu_act_kernel<T, SIGN_READ><<<(unsigned)blocks, 256, 0, s>>>(p, total, wq, rows);
    else                        filtered_lrelu_act_kernel<T, SIGN_NONE><<<(unsigned)blocks, 256, 0, s>>>(p, total, wq, rows);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int lvg_filtered_lrelu_act(void* x, const uint8_t* si, uint8_t* so, int dtype,
                                      const int64_t x_shape[4], const int64_t x_stride[4],
                                      int s_h, int s_wbytes, int sx, int sy, float gain,
                                      float slope, float clamp, int write_signs, void* stream)
{
    LVG_REQUIRE(x != nullptr, "filtered_lrelu_act: x must not be NULL");
    LVG_REQUIRE(dtype == LVG_F32 || dtype == LVG_F16 || dtype == LVG_F64, "filtered_lrelu_act: unsupported dtype %d", dtype);
    for (int i = 0; i < 4; i++)
        LVG_REQUIRE(x_shape[i] >= 1 && x_shape[i] <= INT32_MAX, "filtered_lrelu_act: x dimension %d out of range", i);
    LVG_REQUIRE(!(write_signs && si), "filtered_lrelu_act: cannot both read and write signs");
    LVG_REQUIRE(!write_signs || so, "filtered_lrelu_act: write_signs needs an output sign buffer");
    LVG_REQUIRE(!(write_signs || si) || (s_h >= 1 && s_wbytes >= 1), "filtered_lrelu_act: bad sign tensor shape");

    ActParams p;
    p.x = x; p.si = si; p.so = so;
    for (int i = 0; i < 4; i++) p.xs[i] = x_stride[i];
    p.n = (int)x_shape[0]; p.c = (int)x_shape[1]; p.h = (int)x_shape[2]; p.w = (int)x_shape[3];
    p.s_h = s_h; p.s_wbytes = s_wbytes; p.sx = sx; p.sy = sy;
    p.gain = gain; p.slope = slope; p.clamp = clamp;
    const int mode = write_signs ? SIGN_WRITE : (si ? SIGN_READ : SIGN_NONE);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LVG_F32) return launch_act<float>(p, mode, s);
    if (dtype == LVG_F16) return launch_act<__half>(p, mode, s);
    return launch_act<double>(p, mode, s);
}
