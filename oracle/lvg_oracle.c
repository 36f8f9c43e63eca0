/*
 * lvg_oracle.c -- CPU restatement of the torch_utils.ops operator path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (long-video-gan_b200/) may
 * link, import or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Every function restates, from the operator's definition, what the reference
 * computes (file:line relative to NVlabs/long-video-gan):
 *   orc_bias_act        torch_utils/ops/bias_act.py:91-120 (forward) and
 *                       bias_act.cu:55-142 (gradient formulas, clamp masking)
 *   orc_upfirdn2d       torch_utils/ops/upfirdn2d.py:167-211
 *   orc_upfirdn2d_adj   the exact adjoint (transpose) of orc_upfirdn2d, written
 *                       as a scatter -- checks the padding algebra of
 *                       upfirdn2d.py:250-269 independently
 *   orc_filtered_lrelu  torch_utils/ops/filtered_lrelu.py:121-153 plus the sign
 *                       coding of filtered_lrelu.cu:494-519 / :567-579
 *   orc_conv2d          torch.nn.functional.conv2d (cross-correlation), grouped
 *
 * Storage is float32 (fp16 test inputs are widened exactly), arithmetic is
 * float64 so the oracle sits within half an fp32 ulp of the real-number result.
 * Pinned against the reference's own _ref implementations by
 * oracle/pin_against_reference.py -> tests/golden/.
 *
 * Parallelised with OpenMP over planes/rows so it can also serve as the
 * "all host cores" CPU baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* bias_act                                                                  */

static double act_eval(int act, double x, double alpha)
{
    switch (act) {
        case 1: return x;                                   /* linear   */
        case 2: return x > 0 ? x : 0;                       /* relu     */
        case 3: return x > 0 ? x : x * alpha;               /* lrelu    */
        case 4: return tanh(x);                             /* tanh     */
        case 5: return 1.0 / (1.0 + exp(-x));               /* sigmoid  */
        case 6: return x >= 0 ? x : expm1(x);               /* elu      */
        case 7: { const double s = 1.0507009873554804934193349852946, a = 1.6732632423543772848170429916717;
                  return x >= 0 ? s * x : s * a * expm1(x); } /* selu   */
        case 8: return x > 80 ? x : log1p(exp(x));          /* softplus */
        case 9: return x / (1.0 + exp(-x));                 /* swish    */
    }
    return NAN;
}

/* first derivative of the activation, written in the variable the reference keeps:
 * r = act(x) (everything except swish) or x itself (swish) */
static double act_d1(int act, double r, double xr, double alpha)
{
    const double s = 1.0507009873554804934193349852946, a = 1.6732632423543772848170429916717;
    switch (act) {
        case 1: return 1;
        case 2: return r > 0 ? 1 : 0;
        case 3: return r > 0 ? 1 : alpha;
        case 4: return 1 - r * r;
        case 5: return r * (1 - r);
        case 6: return r >= 0 ? 1 : r + 1;
        case 7: return r >= 0 ? s : r + s * a;
        case 8: return 1 - exp(-r);
        case 9: { double sg = 1.0 / (1.0 + exp(-xr)); return sg * (1 + xr * (1 - sg)); }
    }
    return NAN;
}

/* second derivative (zero for the piecewise-linear activations) */
static double act_d2(int act, double r, double xr)
{
    const double s = 1.0507009873554804934193349852946, a = 1.6732632423543772848170429916717;
    switch (act) {
        case 1: case 2: case 3: return 0;
        case 4: return (1 - r * r) * (-2 * r);
        case 5: return r * (1 - r) * (1 - 2 * r);
        case 6: return r >= 0 ? 0 : r + 1;
        case 7: return r >= 0 ? 0 : r + s * a;
        case 8: { double c = exp(-r); return c * (1 - c); }
        case 9: { double sg = 1.0 / (1.0 + exp(-xr)); double d = sg * (1 - sg); return d * (2 + xr * (1 - 2 * sg)); }
    }
    return NAN;
}

/*
 * grad 0: y = clamp(act(x + b) * gain)
 * grad 1: y = x * gain * act'(.)        masked where the forward output yref left (-clamp, clamp)
 * grad 2: y = x * dy * gain * act''(.)  same mask
 * NULL operands are absent (treated as 0; dy as 1), exactly like the plugin's empty tensors.
 * For swish the forward output used by the mask is rebuilt from xref (+b).
 */
void orc_bias_act(const float* x, const float* b, const float* xref, const float* yref, const float* dy,
                  float* y, int64_t n, int64_t size_b, int64_t step_b, int grad, int act,
                  double alpha, double gain, double clamp)
{
    /* the plugin interface carries these as float32 (bias_act.cpp:32); compare/scale with the same values */
    alpha = (double)(float)alpha; gain = (double)(float)gain; clamp = (double)(float)clamp;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const double bias = b ? (double)b[(i / step_b) % size_b] : 0.0;
        double out;
        if (grad == 0) {
            out = act_eval(act, (double)x[i] + bias, alpha) * gain;
            if (clamp >= 0) out = out > clamp ? clamp : (out < -clamp ? -clamp : out);
        } else {
            const double g = (double)x[i];
            const double xr = (xref ? (double)xref[i] : 0.0) + bias;
            double yr = yref ? (double)yref[i] : 0.0;
            const double r = gain != 0 ? yr / gain : 0.0;
            const double d = grad == 1 ? act_d1(act, r, xr, alpha) : act_d2(act, r, xr);
            out = g * d * gain * ((grad == 2 && dy) ? (double)dy[i] : 1.0);
            if (act == 9) yr = act_eval(9, xr, alpha) * gain;
            if (clamp >= 0 && !(yr > -clamp && yr < clamp)) out = 0;
        }
        y[i] = (float)out;
    }
}

/* ------------------------------------------------------------------------ */
/* upfirdn2d on dense [planes][h][w] arrays                                  */

static inline int64_t floordiv64(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

/*
 * out[oy][ox] = gain * sum_{ty,tx} g[ty][tx] * u[oy*downy + ty - pady0][ox*downx + tx - padx0]
 * u = zero-stuffed x (u[j] = x[j/up] when up divides j and the index is inside the image),
 * g = f when flip (correlation), g = f mirrored in both axes otherwise (convolution).
 */
void orc_upfirdn2d(const float* x, const float* f, float* y, int64_t planes, int ih, int iw, int oh, int ow,
                   int fw, int fh, int upx, int upy, int downx, int downy, int padx0, int pady0,
                   int flip, double gain)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t p = 0; p < planes; p++) {
        for (int oy = 0; oy < oh; oy++) {
            const float* xp = x + p * (int64_t)ih * iw;
            for (int ox = 0; ox < ow; ox++) {
                double acc = 0;
                for (int ty = 0; ty < fh; ty++) {
                    const int64_t uy = (int64_t)oy * downy + ty - pady0;
                    if (uy < 0 || uy % upy != 0) continue;
                    const int64_t iy = uy / upy;
                    if (iy >= ih) continue;
                    for (int tx = 0; tx < fw; tx++) {
                        const int64_t ux = (int64_t)ox * downx + tx - padx0;
                        if (ux < 0 || ux % upx != 0) continue;
                        const int64_t ix = ux / upx;
                        if (ix >= iw) continue;
                        const float g = flip ? f[ty * fw + tx] : f[(fh - 1 - ty) * fw + (fw - 1 - tx)];
                        acc += (double)g * (double)xp[iy * iw + ix];
                    }
                }
                y[(p * oh + oy) * (int64_t)ow + ox] = (float)(acc * gain);
            }
        }
    }
}

/* adjoint of the map above: dx = A^T dy, accumulated as a scatter per plane */
void orc_upfirdn2d_adj(const float* dy, const float* f, float* dx, int64_t planes, int ih, int iw, int oh, int ow,
                       int fw, int fh, int upx, int upy, int downx, int downy, int padx0, int pady0,
                       int flip, double gain)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < planes; p++) {
        double* acc = (double*)calloc((size_t)ih * iw, sizeof(double));
        for (int oy = 0; oy < oh; oy++)
            for (int ox = 0; ox < ow; ox++) {
                const double d = (double)dy[(p * oh + oy) * (int64_t)ow + ox] * gain;
                for (int ty = 0; ty < fh; ty++) {
                    const int64_t uy = (int64_t)oy * downy + ty - pady0;
                    if (uy < 0 || uy % upy != 0 || uy / upy >= ih) continue;
                    for (int tx = 0; tx < fw; tx++) {
                        const int64_t ux = (int64_t)ox * downx + tx - padx0;
                        if (ux < 0 || ux % upx != 0 || ux / upx >= iw) continue;
                        const float g = flip ? f[ty * fw + tx] : f[(fh - 1 - ty) * fw + (fw - 1 - tx)];
                        acc[(uy / upy) * iw + ux / upx] += (double)g * d;
                    }
                }
            }
        for (int64_t i = 0; i < (int64_t)ih * iw; i++) dx[p * (int64_t)ih * iw + i] = (float)acc[i];
        free(acc);
    }
}

/* ------------------------------------------------------------------------ */
/* filtered_lrelu                                                            */

/*
 * y = downfir( act( upfir(x + b) ) ), full 2-D filters fu[fuh][fuw], fd[fdh][fdw]
 * (separable filters are passed as their outer product by the wrapper).
 *   stage 1: t = upfirdn2d(x + b[c], fu, up, pad (px0, py0), gain up^2)   size uh x uw
 *   stage 2: v = t * gain; write mode / plain: negative -> *slope (code 1), |v| > clamp -> +-clamp (code 2)
 *            read mode: code looked up at (ux + sx, uy + sy); bit0 -> *slope, bit1 -> 0; outside the
 *            sign tensor the value passes unchanged
 *   stage 3: y = upfirdn2d(v, fd, down)
 * rounding points: `stage_round` != 0 rounds t and v to float between stages (what the composed
 * _ref path does in fp32); 0 keeps float64 throughout (what a fused kernel approximates).
 * signs (optional): uint8 [planes][s_h][s_wbytes], 2 bits per sample, 4 samples per byte.
 * channels = C so that plane p uses bias b[p % C].
 */
void orc_filtered_lrelu(const float* x, const float* fu, const float* fd, const float* b,
                        const uint8_t* si, float* y, uint8_t* so,
                        int64_t planes, int channels, int ih, int iw, int oh, int ow,
                        int fuw, int fuh, int fdw, int fdh, int up, int down, int px0, int py0,
                        int s_h, int s_wbytes, int sx, int sy,
                        double gain, double slope, double clamp, int flip, int stage_round)
{
    gain = (double)(float)gain; slope = (double)(float)slope; clamp = (double)(float)clamp;   /* float32 in the plugin interface */
    const int uw = ow * down - (down - 1) + (fdw - 1);   /* up-sampled extent actually consumed */
    const int uh = oh * down - (down - 1) + (fdh - 1);
#pragma omp parallel for schedule(dynamic)
    for (int64_t p = 0; p < planes; p++) {
        const float* xp = x + p * (int64_t)ih * iw;
        const double bias = b ? (double)b[p % channels] : 0.0;
        double* t = (double*)malloc(sizeof(double) * (size_t)uh * uw);
        for (int uy = 0; uy < uh; uy++)
            for (int ux = 0; ux < uw; ux++) {
                double acc = 0;
                for (int ty = 0; ty < fuh; ty++) {
                    const int64_t zy = (int64_t)uy + ty - py0;
                    if (zy < 0 || zy % up != 0 || zy / up >= ih) continue;
                    for (int tx = 0; tx < fuw; tx++) {
                        const int64_t zx = (int64_t)ux + tx - px0;
                        if (zx < 0 || zx % up != 0 || zx / up >= iw) continue;
                        const float g = flip ? fu[ty * fuw + tx] : fu[(fuh - 1 - ty) * fuw + (fuw - 1 - tx)];
                        acc += (double)g * ((double)xp[(zy / up) * iw + zx / up] + bias);
                    }
                }
                acc *= (double)up * up;
                if (stage_round) acc = (double)(float)acc;
                double v = acc * gain;
                unsigned code = 0;
                if (si) {
                    const int64_t qx = (int64_t)ux + sx, qy = (int64_t)uy + sy;
                    if (qx >= 0 && qx < (int64_t)s_wbytes * 4 && qy >= 0 && qy < s_h) {
                        code = (si[(p * s_h + qy) * s_wbytes + (qx >> 2)] >> ((qx & 3) * 2)) & 3u;
                        if (code & 1u) v *= slope;
                        if (code & 2u) v = 0;
                    }
                } else {
                    if (v < 0) { v *= slope; code = 1; }
                    if (fabs(v) > clamp) { v = v < 0 ? -clamp : clamp; code = 2; }
                    if (so && ux < s_wbytes * 4 && uy < s_h) {
                        uint8_t* q = so + (p * s_h + uy) * s_wbytes + (ux >> 2);
                        *q = (uint8_t)((*q & ~(3u << ((ux & 3) * 2))) | (code << ((ux & 3) * 2)));
                    }
                }
                if (stage_round) v = (double)(float)v;
                t[uy * uw + ux] = v;
            }
        for (int oy = 0; oy < oh; oy++)
            for (int ox = 0; ox < ow; ox++) {
                double acc = 0;
                for (int ty = 0; ty < fdh; ty++)
                    for (int tx = 0; tx < fdw; tx++) {
                        const float g = flip ? fd[ty * fdw + tx] : fd[(fdh - 1 - ty) * fdw + (fdw - 1 - tx)];
                        acc += (double)g * t[(oy * down + ty) * uw + ox * down + tx];
                    }
                y[(p * oh + oy) * (int64_t)ow + ox] = (float)acc;
            }
        free(t);
    }
}

/* ------------------------------------------------------------------------ */
/* grouped conv2d (cross-correlation), NCHW, zero padding                    */

void orc_conv2d(const float* x, const float* w, float* y, int n, int groups, int cin, int cout,
                int h, int wd, int kh, int kw, int stride, int pad_h, int pad_w)
{
    const int oh = (h + 2 * pad_h - kh) / stride + 1;
    const int ow = (wd + 2 * pad_w - kw) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t ng = 0; ng < (int64_t)n * groups; ng++) {
        for (int co = 0; co < cout; co++) {
            const int nn = (int)(ng / groups), g = (int)(ng % groups);
            const float* xg = x + ((int64_t)nn * groups + g) * cin * h * wd;
            const float* wg = w + ((int64_t)g * cout + co) * cin * kh * kw;
            float* yg = y + (((int64_t)nn * groups + g) * cout + co) * oh * ow;
            for (int oy = 0; oy < oh; oy++)
                for (int ox = 0; ox < ow; ox++) {
                    double acc = 0;
                    for (int ci = 0; ci < cin; ci++)
                        for (int ky = 0; ky < kh; ky++) {
                            const int iy = oy * stride + ky - pad_h;
                            if (iy < 0 || iy >= h) continue;
                            for (int kx = 0; kx < kw; kx++) {
                                const int ix = ox * stride + kx - pad_w;
                                if (ix < 0 || ix >= wd) continue;
                                acc += (double)wg[(ci * kh + ky) * kw + kx] * (double)xg[((int64_t)ci * h + iy) * wd + ix];
                            }
                        }
                    yg[oy * ow + ox] = (float)acc;
                }
        }
    }
}
