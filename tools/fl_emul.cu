// CPU emulation of the fused filtered_lrelu kernel (csrc/filtered_lrelu_v3.cuh compiled as host code): the stage
// functions run thread by thread over a heap "shared memory", so index arithmetic, alignment of the vector accesses
// and the sign-tensor plumbing are checked against the operator's definition without a GPU.
//   nvcc -O1 -std=c++17 -DFLV3_HOST_EMU -o /tmp/fl_emul tools/fl_emul.cu && /tmp/fl_emul
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../long-video-gan_b200/csrc/filtered_lrelu_v3.cuh"

using namespace lvg::flv3;

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

struct Case {
    int up, fu, down, fd, ih, iw, px0, px1, py0, py1;
    float gain, slope, clamp;
    int flip;
    float xscale;
};

// the operator by definition, double precision; mode as the kernel; returns y, writes/reads codes (unpacked, one per sample)
static void naive(const Case& c, int C, const std::vector<float>& x, const std::vector<float>& b, const std::vector<float>& fu,
                  const std::vector<float>& fd, int oh, int ow, int uh, int uw, int sgh, int sgw, int mode, std::vector<uint8_t>& codes, int sx, int sy,
                  std::vector<double>& y, std::vector<double>& vact)
{
    // uh x uw = consumed up-sampled extent; sgh x sgw = extent of the (unpacked) sign tensor
    std::vector<double> t((size_t)C * uh * uw);
    auto gu = [&](int s) { return (double)(c.flip ? fu[s] : fu[c.fu - 1 - s]); };
    auto gd = [&](int s) { return (double)(c.flip ? fd[s] : fd[c.fd - 1 - s]); };
    for (int ch = 0; ch < C; ch++) {
        // up x then up y
        std::vector<double> ux((size_t)c.ih * uw);
        for (int iy = 0; iy < c.ih; iy++)
            for (int U = 0; U < uw; U++) {
                double a = 0;
                for (int s = 0; s < c.fu; s++) {
                    const int pz = U + s - c.px0;
                    if (pz % c.up != 0 || pz < 0) continue;
                    const int m = pz / c.up;
                    if (m >= c.iw) continue;
                    a += gu(s) * ((double)x[((size_t)ch * c.ih + iy) * c.iw + m] + (double)b[ch]);
                }
                ux[(size_t)iy * uw + U] = a * c.up;
            }
        for (int V = 0; V < uh; V++)
            for (int U = 0; U < uw; U++) {
                double a = 0;
                for (int s = 0; s < c.fu; s++) {
                    const int pz = V + s - c.py0;
                    if (pz % c.up != 0 || pz < 0) continue;
                    const int m = pz / c.up;
                    if (m >= c.ih) continue;
                    a += gu(s) * ux[(size_t)m * uw + U];
                }
                double v = a * c.up * c.gain;
                const size_t ci = ((size_t)ch * uh + V) * uw + U;
                if (mode == SIGN_READ) {
                    const int qx = U + sx, qy = V + sy;
                    uint8_t code = 0;
                    if (qx >= 0 && qy >= 0 && qy < sgh && qx < sgw) code = codes[((size_t)ch * sgh + qy) * sgw + qx];
                    if (code & 1) v *= c.slope;
                    if (code & 2) v = 0;
                } else {
                    uint8_t code = 0;
                    if (v < 0) { v *= c.slope; code = 1; }
                    if (fabs(v) > c.clamp) { v = v < 0 ? -c.clamp : c.clamp; code = 2; }
                    if (mode == SIGN_WRITE) codes[ci] = code;
                }
                vact[ci] = v;
                t[ci] = v;
            }
        std::vector<double> dx((size_t)uh * ow);
        for (int V = 0; V < uh; V++)
            for (int o = 0; o < ow; o++) {
                double a = 0;
                for (int s = 0; s < c.fd; s++) a += gd(s) * t[((size_t)ch * uh + V) * uw + o * c.down + s];
                dx[(size_t)V * ow + o] = a;
            }
        for (int o2 = 0; o2 < oh; o2++)
            for (int o = 0; o < ow; o++) {
                double a = 0;
                for (int s = 0; s < c.fd; s++) a += gd(s) * dx[(size_t)(o2 * c.down + s) * ow + o];
                y[((size_t)ch * oh + o2) * ow + o] = a;
            }
    }
}

template <class T, class G, int MODE>
static void run_emul(FlParams p)
{
    fill_launch_constants<G>(p);
    const size_t bytes = G::smem_bytes(SIGN_READ);
    std::vector<float> smem(bytes / 4 + 16);
    float* base = (float*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    const unsigned blocks = (unsigned)(p.n * p.c * p.tiles_x * p.tiles_y);
    for (unsigned bid = 0; bid < blocks; bid++) {
        for (size_t i = 0; i < bytes / 4; i++) base[i] = NAN;          // uninitialised shared memory must never reach an output
        const Tile t = make_tile<G>(p, bid);
        const Smem s = carve<G>(base);
        for (int tid = 0; tid < kThreads; tid++) stage0<G, MODE>(p, t, s, tid);
        for (int tid = 0; tid < kThreads; tid++) stage1<T, G>(p, t, s, tid);
        for (int tid = 0; tid < kThreads; tid++) stage2<G>(t, s, tid);
        for (int tid = 0; tid < kThreads; tid++) stage3<G, MODE>(p, t, s, tid);
        if (MODE == SIGN_WRITE) for (int tid = 0; tid < kThreads; tid++) stage3_fixup<G>(p, t, tid);
        for (int tid = 0; tid < kThreads; tid++) stage4<G>(t, s, tid);
        for (int tid = 0; tid < kThreads; tid++) stage5<T, G>(p, t, s, tid);
    }
}

template <class G>
static int run_case(const Case& c, const char* name)
{
    const int C = 2;
    const int ow = (c.iw * c.up + c.px0 + c.px1 - (c.fu - 1) - (c.fd - 1) + (c.down - 1)) / c.down;
    const int oh = (c.ih * c.up + c.py0 + c.py1 - (c.fu - 1) - (c.fd - 1) + (c.down - 1)) / c.down;
    const int sw = ow * c.down - (c.down - 1) + c.fd - 1, sh = oh * c.down - (c.down - 1) + c.fd - 1;
    const int s_wb = (sw + 15) / 16 * 4;
    std::vector<float> x((size_t)C * c.ih * c.iw), b(C), fu(c.fu), fd(c.fd);
    for (auto& v : x) v = frand() * c.xscale;
    for (auto& v : b) v = frand();
    for (auto& v : fu) v = frand() / 3;
    for (auto& v : fd) v = frand() / 3;
    std::vector<float> y((size_t)C * oh * ow, NAN);
    std::vector<uint8_t> so((size_t)C * sh * s_wb, 0xAA);

    FlParams p;
    memset(&p, 0, sizeof p);
    p.x = x.data(); p.fu = fu.data(); p.fd = fd.data(); p.b = b.data(); p.y = y.data(); p.so = so.data();
    p.xs[0] = (int64_t)C * c.ih * c.iw; p.xs[1] = (int64_t)c.ih * c.iw; p.xs[2] = c.iw; p.xs[3] = 1;
    p.ys[0] = (int64_t)C * oh * ow; p.ys[1] = (int64_t)oh * ow; p.ys[2] = ow; p.ys[3] = 1;
    p.n = 1; p.c = C; p.ih = c.ih; p.iw = c.iw; p.oh = oh; p.ow = ow;
    p.px0 = c.px0; p.py0 = c.py0; p.s_h = sh; p.s_wb = s_wb; p.sx = p.sy = 0; p.sw_active = sw;
    p.gain = c.gain; p.slope = c.slope; p.clamp = c.clamp; p.flip = c.flip;

    int bad = 0;
    // ---- write mode
    run_emul<float, G, SIGN_WRITE>(p);
    std::vector<uint8_t> codes((size_t)C * sh * sw);
    std::vector<double> yr((size_t)C * oh * ow), vact((size_t)C * sh * sw);
    naive(c, C, x, b, fu, fd, oh, ow, sh, sw, sh, sw, SIGN_WRITE, codes, 0, 0, yr, vact);
    double ymax = 0, emax = 0;
    for (size_t i = 0; i < yr.size(); i++) { ymax = std::max(ymax, fabs(yr[i])); emax = std::max(emax, fabs(yr[i] - (double)y[i])); if (!(y[i] == y[i])) emax = 1e30; }
    size_t flips = 0, padbad = 0;
    for (int ch = 0; ch < C; ch++)
        for (int V = 0; V < sh; V++)
            for (int bb = 0; bb < s_wb; bb++) {
                const uint8_t byte = so[((size_t)ch * sh + V) * s_wb + bb];
                for (int k = 0; k < 4; k++) {
                    const int U = bb * 4 + k;
                    const int got = (byte >> (2 * k)) & 3;
                    if (U >= sw) { if (got != 0) padbad++; continue; }
                    const size_t ci = ((size_t)ch * sh + V) * sw + U;
                    if (got != codes[ci]) {
                        const double v = vact[ci];
                        const bool near = fabs(v) < 1e-4 * (1 + c.clamp * 0) || fabs(fabs(v) - c.clamp) < 1e-3 * c.clamp || fabs(fabs(v) / c.slope - c.clamp) < 1e-3 * c.clamp;
                        if (!near) flips++;
                    }
                }
            }
    printf("%-28s write: out %dx%d max|y| %.3g  err %.3g  sign mismatches %zu  pad bytes bad %zu\n", name, oh, ow, ymax, emax, flips, padbad);
    if (emax > 2e-5 * ymax || flips || padbad) bad++;

    // ---- plain mode
    std::fill(y.begin(), y.end(), NAN);
    run_emul<float, G, SIGN_NONE>(p);
    emax = 0;
    for (size_t i = 0; i < yr.size(); i++) { emax = std::max(emax, fabs(yr[i] - (double)y[i])); if (!(y[i] == y[i])) emax = 1e30; }
    printf("%-28s plain: err %.3g\n", name, emax);
    if (emax > 2e-5 * ymax) bad++;

    // ---- read mode: random sign tensor of a different size, offsets sx, sy (incl. negative / beyond the tensor)
    for (int trial = 0; trial < 3; trial++) {
        const int s2h = sh + (trial == 1 ? -7 : 5), s2wb = s_wb + (trial == 1 ? -4 : 4);    // rows of the sign tensor are multiples of 4 bytes (the kernel loads words)
        std::vector<uint8_t> si((size_t)C * s2h * s2wb);
        for (auto& v : si) v = (uint8_t)(rand() & 0xFF);
        const int sx = trial == 0 ? 3 : (trial == 1 ? -5 : 6), sy = trial == 0 ? 2 : (trial == 1 ? -3 : 0);
        std::vector<uint8_t> cu((size_t)C * s2h * s2wb * 4);
        for (int ch = 0; ch < C; ch++)
            for (int V = 0; V < s2h; V++)
                for (int q = 0; q < s2wb * 4; q++)
                    cu[((size_t)ch * s2h + V) * s2wb * 4 + q] = (si[((size_t)ch * s2h + V) * s2wb + q / 4] >> (2 * (q & 3))) & 3;
        FlParams pr = p;
        pr.si = si.data(); pr.so = nullptr; pr.s_h = s2h; pr.s_wb = s2wb; pr.sx = sx; pr.sy = sy; pr.clamp = INFINITY;
        std::fill(y.begin(), y.end(), NAN);
        run_emul<float, G, SIGN_READ>(pr);
        Case cr = c; cr.clamp = INFINITY;
        // naive with read semantics: its codes vector is [C][s2h][s2wb*4]
        std::vector<double> yr2((size_t)C * oh * ow), va2((size_t)C * sh * sw);
        naive(cr, C, x, b, fu, fd, oh, ow, sh, sw, s2h, s2wb * 4, SIGN_READ, cu, sx, sy, yr2, va2);
        (void)va2;
        emax = 0; double ym = 0;
        for (size_t i = 0; i < yr2.size(); i++) { ym = std::max(ym, fabs(yr2[i])); emax = std::max(emax, fabs(yr2[i] - (double)y[i])); if (!(y[i] == y[i])) emax = 1e30; }
        printf("%-28s read(sx %d, sy %d): err %.3g (max|y| %.3g)\n", name, sx, sy, emax, ym);
        if (emax > 2e-5 * ym) bad++;
    }
    return bad;
}

int main()
{
    srand(1);
    int bad = 0;
    typedef Geom<2, 12, 2, 12, 56, 24, 6, 2, 4, 4> U2D2;
    typedef Geom<4, 24, 2, 12, 56, 24, 2, 2, 4, 4> U4D2;
    typedef Geom<2, 12, 4, 24, 31, 16, 8, 6, 4, 2> U2D4;
    printf("U2D2 smem %zu B, U4D2 %zu B, U2D4 %zu B\n", U2D2::smem_bytes(SIGN_READ), U4D2::smem_bytes(SIGN_READ), U2D4::smem_bytes(SIGN_READ));
    {
        Case c = {2, 12, 2, 12, 31, 38, 9, 8, 9, 8, 1.4142135f, 0.2f, 256.f, 0, 1.f};
        bad += run_case<U2D2>(c, "U2D2 31x38 pad 9,8");
        Case c2 = {2, 12, 2, 12, 94, 150, 9, 8, 9, 8, 1.4142135f, 0.2f, 3.f, 1, 10.f};
        bad += run_case<U2D2>(c2, "U2D2 94x150 clamp 3 flip");
        Case c3 = {2, 12, 2, 12, 70, 139, -11, -12, -11, -12, 1.f, 0.3f, 256.f, 0, 1.f};
        bad += run_case<U2D2>(c3, "U2D2 70x139 pad -11,-12");
        Case c4 = {2, 12, 2, 12, 9, 7, 10, 11, 12, 9, 1.f, 0.3f, 256.f, 0, 1.f};
        bad += run_case<U2D2>(c4, "U2D2 9x7 pad 10,11,12,9");
    }
    {
        Case c = {4, 24, 2, 12, 31, 38, -6, -9, -6, -9, 1.4142135f, 0.2f, 256.f, 0, 1.f};
        bad += run_case<U4D2>(c, "U4D2 31x38 pad -6,-9");
        Case c2 = {4, 24, 2, 12, 40, 54, 17, 14, 15, 16, 1.4142135f, 0.2f, 2.f, 1, 10.f};
        bad += run_case<U4D2>(c2, "U4D2 40x54 clamp 2 flip");
    }
    {
        Case c = {2, 12, 4, 24, 38, 52, 40, 29, 40, 29, 1.4142135f * 0.25f, 0.2f, 256.f, 1, 1.f};
        bad += run_case<U2D4>(c, "U2D4 38x52 pad 40,29");
        Case c2 = {2, 12, 4, 24, 77, 41, 23, 24, 25, 22, 1.f, 0.2f, 1.f, 0, 10.f};
        bad += run_case<U2D4>(c2, "U2D4 77x41 clamp 1");
    }
    printf(bad ? "FAILED (%d)\n" : "all ok\n", bad);
    return bad ? 1 : 0;
}
