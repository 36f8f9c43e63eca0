// upfirdn2d: zero-insert upsample -> pad/crop -> FIR -> decimate, per (n, c) plane.
//
// Definition used throughout (equivalent to the reference kernels,
// torch_utils/ops/upfirdn2d.cu:29-200, and to _upfirdn2d_ref, upfirdn2d.py:167-211):
//
//   out[o] = gain * sum_t g[t] * u[o*down + t - pad0],   u[j] = x[j/up] if up | j (0 <= j < in*up) else 0
//   g = f            if flip   (correlation)
//   g = reversed f   otherwise (true convolution)
//
// so only the taps t == (pad0 - o*down) mod up contribute (polyphase form).
//
// This file holds the fully general kernel (any strides, any up/down/taps,
// fp16/fp32/fp64, 64-bit indexing). The tiled single-launch separable kernel
// for the shapes the networks use lives in upfirdn2d_tiled.cu.

#include "common.cuh"

namespace lvg {

struct UpfirdnParams {
    const void* x;
    const float* f;
    void* y;
    int64_t xs[4];   // x strides  N C H W (elements)
    int64_t ys[4];   // y strides
    int64_t fsx, fsy;
    int n, c, ih, iw, oh, ow;
    int fw, fh;
    int upx, upy, downx, downy, padx0, pady0;
    int flip;
    int c_minor;     // 1: walk channels fastest (channels-last storage)
    float gain;
};

namespace {

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_any_kernel(UpfirdnParams p, int64_t total)
{
    typedef typename Acc<T>::type S;
    const T* __restrict__ x = (const T*)p.x;
    T* __restrict__ y = (T*)p.y;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int ox, oy, c, n;
        int64_t r = idx;
        if (p.c_minor) {
            c = (int)(r % p.c); r /= p.c;
            ox = (int)(r % p.ow); r /= p.ow;
            oy = (int)(r % p.oh); n = (int)(r / p.oh);
        } else {
            ox = (int)(r % p.ow); r /= p.ow;
            oy = (int)(r % p.oh); r /= p.oh;
            c = (int)(r % p.c); n = (int)(r / p.c);
        }
        // first contributing tap and the input sample it meets, per axis
        const int bx = ox * p.downx - p.padx0;
        const int by = oy * p.downy - p.pady0;
        int tx0 = posmod(-bx, p.upx);
        int ty0 = posmod(-by, p.upy);
        int ix0 = (bx + tx0) / p.upx;     // exact division
        int iy0 = (by + ty0) / p.upy;
        // skip taps that fall before the first input sample
        if (ix0 < 0) { tx0 += -ix0 * p.upx; ix0 = 0; }
        if (iy0 < 0) { ty0 += -iy0 * p.upy; iy0 = 0; }

        const T* xp = x + (int64_t)n * p.xs[0] + (int64_t)c * p.xs[1];
        S acc = (S)0;
        for (int ty = ty0, iy = iy0; ty < p.fh && iy < p.ih; ty += p.upy, iy++) {
            const int fy = p.flip ? ty : p.fh - 1 - ty;
            const float* frow = p.f + fy * p.fsy;
            const T* xrow = xp + (int64_t)iy * p.xs[2];
            for (int tx = tx0, ix = ix0; tx < p.fw && ix < p.iw; tx += p.upx, ix++) {
                const int fx = p.flip ? tx : p.fw - 1 - tx;
                acc += to_acc(xrow[(int64_t)ix * p.xs[3]]) * (S)frow[fx * p.fsx];
            }
        }
        acc *= (S)p.gain;
        y[(int64_t)n * p.ys[0] + (int64_t)c * p.ys[1] + (int64_t)oy * p.ys[2] + (int64_t)ox * p.ys[3]] = from_acc<T>(acc);
    }
}

}  // namespace

int upfirdn2d_tiled(const void* x, const float* fx, int64_t fsx, const float* fy, int64_t fsy, void* y, int dtype,
                    const int64_t* xsh, const int64_t* xst, const int64_t* ysh, const int64_t* yst,
                    int fw, int fh, int upx, int upy, int downx, int downy, int padx0, int pady0,
                    int flip, float gain, cudaStream_t s);

int upfirdn2d_check(const void* x, const void* y, int dtype, const int64_t* xsh, const int64_t* ysh,
                    int fw, int fh, int upx, int upy, int downx, int downy)
{
    LVG_REQUIRE(x && y, "upfirdn2d: x and y must not be NULL");
    LVG_REQUIRE(dtype == LVG_F32 || dtype == LVG_F16 || dtype == LVG_F64, "upfirdn2d: unsupported dtype %d", dtype);
    LVG_REQUIRE(fw >= 1 && fh >= 1, "upfirdn2d: filter must be at least 1x1");
    LVG_REQUIRE(upx >= 1 && upy >= 1, "upfirdn2d: upsampling factor must be at least 1");
    LVG_REQUIRE(downx >= 1 && downy >= 1, "upfirdn2d: downsampling factor must be at least 1");
    for (int i = 0; i < 4; i++) {
        LVG_REQUIRE(xsh[i] >= 1 && xsh[i] <= INT32_MAX, "upfirdn2d: x dimension %d out of range", i);
        LVG_REQUIRE(ysh[i] >= 1 && ysh[i] <= INT32_MAX, "upfirdn2d: output must be at least 1x1 (dimension %d)", i);
    }
    LVG_REQUIRE(xsh[0] == ysh[0] && xsh[1] == ysh[1], "upfirdn2d: x and y disagree on batch/channels");
    return LVG_OK;
}

}  // namespace lvg

using namespace lvg;

extern "C" int lvg_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                             const int64_t x_shape[4], const int64_t x_stride[4],
                             const int64_t y_shape[4], const int64_t y_stride[4],
                             int fw, int fh, int64_t f_stride_x, int64_t f_stride_y,
                             int upx, int upy, int downx, int downy, int padx0, int pady0,
                             int flip, float gain, void* stream)
{
    int rc = upfirdn2d_check(x, y, dtype, x_shape, y_shape, fw, fh, upx, upy, downx, downy);
    if (rc) return rc;
    LVG_REQUIRE(f != nullptr, "upfirdn2d: f must not be NULL");

    // a filter spanning one axis only ([k, 1] temporal filters, [1, k]) is a 1-D pass: tiled kernel
    if (fw == 1 && fh > 1 && upx == 1 && downx == 1 && x_shape[3] == 1 && y_shape[3] == 1) {
        // [N, C, L, 1] tensors (the temporal embedding pyramid): the filtered axis is the contiguous one.
        // Same op on the transposed view [N, C, 1, L] with the filter along x, so lanes run along L.
        const int64_t xsh[4] = {x_shape[0], x_shape[1], 1, x_shape[2]}, xst[4] = {x_stride[0], x_stride[1], x_stride[3], x_stride[2]};
        const int64_t ysh[4] = {y_shape[0], y_shape[1], 1, y_shape[2]}, yst[4] = {y_stride[0], y_stride[1], y_stride[3], y_stride[2]};
        rc = upfirdn2d_tiled(x, f, f_stride_y, nullptr, 1, y, dtype, xsh, xst, ysh, yst, fh, 1, upy, 1, downy, 1, pady0, 0,
                             flip, gain, (cudaStream_t)stream);
        if (rc != LVG_UNSUPPORTED) return rc;
    }
    if ((fw == 1) != (fh == 1) && ((fw == 1 && upx == 1 && downx == 1) || (fh == 1 && upy == 1 && downy == 1))) {
        rc = upfirdn2d_tiled(x, fw == 1 ? nullptr : f, f_stride_x, fh == 1 ? nullptr : f, f_stride_y, y, dtype,
                             x_shape, x_stride, y_shape, y_stride, fw, fh, upx, upy, downx, downy, padx0, pady0,
                             flip, gain, (cudaStream_t)stream);
        if (rc != LVG_UNSUPPORTED) return rc;
    }

    UpfirdnParams p;
    p.x = x; p.f = f; p.y = y;
    for (int i = 0; i < 4; i++) { p.xs[i] = x_stride[i]; p.ys[i] = y_stride[i]; }
    p.fsx = f_stride_x; p.fsy = f_stride_y;
    p.n = (int)x_shape[0]; p.c = (int)x_shape[1]; p.ih = (int)x_shape[2]; p.iw = (int)x_shape[3];
    p.oh = (int)y_shape[2]; p.ow = (int)y_shape[3];
    p.fw = fw; p.fh = fh;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0;
    p.flip = flip ? 1 : 0;
    p.c_minor = (p.c > 1 && y_stride[1] == 1) ? 1 : 0;
    p.gain = gain;

    const int64_t total = (int64_t)p.n * p.c * p.oh * p.ow;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8 * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LVG_F32)      upfirdn2d_any_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(p, total);
    else if (dtype == LVG_F16) upfirdn2d_any_kernel<__half><<<(unsigned)blocks, 256, 0, s>>>(p, total);
    else                       upfirdn2d_any_kernel<double><<<(unsigned)blocks, 256, 0, s>>>(p, total);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}
