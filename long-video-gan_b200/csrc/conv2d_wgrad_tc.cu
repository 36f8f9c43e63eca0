// Weight gradient of the grouped 2-D convolution as an implicit GEMM on tcgen05 tensor cores.
//
//   dW[g][co][ci][ky][kx] = sum_n sum_{oy,ox} dy[n][g][co][oy][ox] * x[n][g][ci][oy + ky - pad][ox + kx - pad]
//
// (the third leg of conv2d_gradfix: torch_utils/ops/conv2d_gradfix.py:119-141 hands it to
// aten::convolution_backward / cuDNN). For the per-sample-weight ("modulated") convolutions of the
// super-res generator every sample is its own group, so dW is as large as all the weights of the
// batch (318 MB for 64 groups of 512 x 539 x 3 x 3) and cuDNN's grouped wgrad is far off the tensor-core rate.
//
// GEMM view per (group g, 128-channel tile of co, NT-channel tile of ci, filter row ky):
//   D_kx[co][ci] += A[co][pix] * B_kx[ci][pix]     M = 128 (co), N = NT (ci), K = output pixels
// for the KW taps kx of that filter row at once: KW accumulators of NT fp32 columns each live in TMEM
// (3 x 144 = 432 of the 512 columns for Cin = 539). The K dimension runs over the output pixels in
// stages of 64 pixels of one output row (4 UMMA k-steps of 16).
//
// Operand staging, both K-major (the pixel axis is contiguous in NCHW memory, so a core-matrix row of
// 8 pixels is one 16-byte piece of a row), no swizzle:
//   A  dy rows:   [k-step][k8: 2][m8: 16][8 rows x 16 B]
//   B  x rows, one copy per tap kx, shifted by kx pixels (a one-pixel shift cannot be expressed in a
//      descriptor, so the shift happens while the registers are written to shared memory):
//                 [kx][k-step][k8: 2][n8: NT/8][8 rows x 16 B]
// Pixels are fetched as 4-byte pairs into registers one stage ahead (software prefetch), written to
// one of two stage buffers, and one thread issues 4 x KW MMAs per stage and commits them to that
// buffer's mbarrier; the buffer is reused two stages later. One CTA per SM (142-156 KB of shared
// memory, 512 TMEM columns). Epilogue: TMEM -> registers -> fp16 -> a [co][ci][kx] tile in shared
// memory -> global, so that a warp's stores cover runs of KW values per (co, ci) instead of 2-byte
// stores 18 bytes apart per lane.

#include "common.cuh"
#include "tcgen05.cuh"

namespace lvg {
namespace {

using namespace tc;

constexpr int kProducers = 256;               // warps 0-7 stage the operands (and run the epilogue)
constexpr int kThreads = kProducers + 32;    // warp 8 issues the MMAs
constexpr int kBM = 128;
constexpr int kStagePx = 64;                 // pixels per stage = 4 k-steps
constexpr int kAStage = 4 * kBM * 16 * 2;    // 16 KB
constexpr int kMaxA = 4;                     // A items per thread: 128 rows x 8 chunks / 256
constexpr int kMaxB = 5;                     // B items per thread: NT <= 160 rows x 8 chunks / 256
constexpr int kNP = 6;                       // pixel pairs fetched per B item: 8 + KW - 1 (+1 for an odd start) <= 12 pixels

struct WgradParams {
    const __half* x;
    const __half* dy;
    __half* dw;
    int n, groups, cin, cout;
    int h, w, ho, wo;
    int kh, kw, pad_h, pad_w;
    int nt;                 // ci per n-tile (multiple of 16, <= 160 = 32 * kMaxB)
    int x_pair_ok, dy_pair_ok;
};

// 8 consecutive fp16 starting `v` pixels into the fetched pairs (v = 0..3)
template <int V>
__device__ __forceinline__ uint4 pick8(const uint32_t (&q)[kNP])
{
    static_assert(V >= 0 && V <= 3, "shift out of range");
    if constexpr (V % 2 == 0) return make_uint4(q[V / 2], q[V / 2 + 1], q[V / 2 + 2], q[V / 2 + 3]);
    else return make_uint4(__byte_perm(q[V / 2], q[V / 2 + 1], 0x5432), __byte_perm(q[V / 2 + 1], q[V / 2 + 2], 0x5432),
                           __byte_perm(q[V / 2 + 2], q[V / 2 + 3], 0x5432), __byte_perm(q[V / 2 + 3], q[V / 2 + 4], 0x5432));
}

// pixel pair (e, e + 1) of a row; elements outside [lo, hi) read as zero
__device__ __forceinline__ uint32_t load_pair(const __half* row, int e, int lo, int hi, bool paired)
{
    const bool v0 = e >= lo && e < hi, v1 = e + 1 >= lo && e + 1 < hi;
    uint32_t v = 0;
    if (paired && v0 && v1) return __ldg(reinterpret_cast<const unsigned int*>(row + e));
    if (v0) v = __ldg(reinterpret_cast<const unsigned short*>(row + e));
    if (v1) v |= (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(row + e + 1)) << 16;
    return v;
}

template <int KH, int KW, int SH>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_tc_kernel(WgradParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full_bar[2];      // stage buffer written (all producer threads arrive)
    __shared__ uint64_t empty_bar[2];     // the MMAs reading the buffer have completed (tcgen05.commit arrives)
    __shared__ uint32_t tmem_base_slot;

    const int NT = p.nt;
    const int b_kstep = NT * 32;                 // bytes of one k-step of one copy: NT rows x 16 pixels
    const int b_copy = 4 * b_kstep;
    const int stage_bytes = kAStage + KW * b_copy;
    const int nti = blockIdx.x / KH, ky = blockIdx.x - nti * KH;
    const int mti = blockIdx.y, g = blockIdx.z;
    const int ci0 = nti * NT, co0 = mti * kBM;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], kProducers);
        mbar_init(&full_bar[1], kProducers);
        mbar_init(&empty_bar[0], 1);
        mbar_init(&empty_bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    // instruction descriptor: D = f32, A = B = f16, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

    // ---- staging items. Item = (row, 8-pixel chunk); thread t handles chunk c = (t / 8) % 8 of rows
    // rsub + 32 * s (rsub = (t / 64) * 8 + t % 8), so 8 consecutive lanes fill one 128-byte core matrix, a warp's
    // 4-byte loads touch 8 rows x 64 contiguous bytes (a lane-per-row mapping doubled the sector requests and
    // measured 28 % slower), and everything that depends on the chunk is computed once per stage and thread.
    const int c = (threadIdx.x / 8) % 8, c8 = 8 * c;
    const int rsub = (threadIdx.x / 64) * 8 + threadIdx.x % 8;
    const int a_soff0 = (c / 2) * 4096 + (c % 2) * 2048 + (rsub / 8) * 128 + (rsub % 8) * 16;      // + 512 per item
    const int b_soff0 = (c / 2) * b_kstep + (c % 2) * (NT * 16) + (rsub / 8) * 128 + (rsub % 8) * 16;
    const int a_rows = min(kMaxA, max(0, (p.cout - co0 - rsub + 31) / 32));                         // items with a real co row
    const int b_items = min(kMaxB, max(0, (NT - rsub + 31) / 32));                                  // items inside the n-tile
    const int b_rows = min(b_items, max(0, (p.cin - ci0 - rsub + 31) / 32));                        // ... with a real ci row
    const int a_rstride = 32 * p.ho * p.wo, b_rstride = 32 * p.h * p.w;
    const int a_goff0 = (co0 + rsub) * p.ho * p.wo + c8, b_goff0 = (ci0 + rsub) * p.h * p.w + c8;

    const int xstages = (p.wo + kStagePx - 1) / kStagePx;
    const int n_stages = p.n * p.ho * xstages;
    constexpr int sh = SH;               // = pad_w & 1: the fetched window starts at an even pixel, ox0 - pad_w - sh
    const bool x_paired = p.x_pair_ok != 0, dy_paired = p.dy_pair_ok != 0;

    // Two register sets: while the set of stage s is written to shared memory and multiplied, the loads of stage
    // s + 1 (other set) have been in flight for a whole stage already, and the loads of stage s + 2 start right after.
    struct Regs { uint32_t pa[kMaxA][4], pb[kMaxB][kNP]; int ks; };
    Regs R0, R1;
    int pf_inst = 0, pf_oy = 0, pf_xs = 0;       // the next stage to fetch

    auto prefetch = [&](Regs& R) {
        uint32_t (&pa)[kMaxA][4] = R.pa;
        uint32_t (&pb)[kMaxB][kNP] = R.pb;
        const int ox0 = pf_xs * kStagePx;
        const int left = p.wo - ox0;
        R.ks = ((left < kStagePx ? left : kStagePx) + 15) / 16;
        // A: dy, pixels [0, hi_a) of this thread's chunk are inside the row
        const __half* rowa = p.dy + ((int64_t)(pf_inst * p.groups + g) * p.cout) * p.ho * p.wo + (int64_t)pf_oy * p.wo + ox0 + a_goff0;
        const int hi_a = left - c8;
#pragma unroll
        for (int s = 0; s < kMaxA; s++) {
            const __half* row = rowa + s * a_rstride;
            const int hi = s < a_rows ? hi_a : 0;
            if (dy_paired) {         // even row length, aligned base: a pair is never half valid
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t v = 0;
                    if (2 * q < hi) v = __ldg(reinterpret_cast<const unsigned int*>(row) + q);     // hi is even here
                    pa[s][q] = v;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; q++) pa[s][q] = load_pair(row, 2 * q, 0, hi, false);
            }
        }
        // B: x, window elements [lo, hi_b) are inside the image row (element 0 = image column ix0)
        const int iy = pf_oy + ky - p.pad_h;
        const bool rowok = iy >= 0 && iy < p.h;
        const int ix0 = ox0 - p.pad_w - sh + c8;
        const __half* rowb = p.x + ((int64_t)(pf_inst * p.groups + g) * p.cin) * p.h * p.w + (int64_t)iy * p.w + (ox0 - p.pad_w - sh) + b_goff0;
        const int lo_b = rowok ? -ix0 : 1 << 20;
        int hi_b = p.w - ix0;
        if (left < kStagePx) {                                        // ragged last stage of a row: zero what meets no valid
            const int need = (sh + KW - 1 + left - c8 + 1) & ~1;     // output pixel (keeps stray NaNs out of the sums); even,
            if (need < hi_b) hi_b = need;                             // so that whole pixel pairs stay valid
        }
#pragma unroll
        for (int s = 0; s < kMaxB; s++) {
            if (s < b_items) {
                const __half* row = rowb + s * b_rstride;
                const int lo = s < b_rows ? lo_b : 1 << 20;
                if (x_paired) {      // lo and the row end are even
#pragma unroll
                    for (int q = 0; q < kNP; q++) {
                        uint32_t v = 0;
                        if (2 * q >= lo && 2 * q < hi_b) v = __ldg(reinterpret_cast<const unsigned int*>(row) + q);
                        pb[s][q] = v;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < kNP; q++) pb[s][q] = load_pair(row, 2 * q, lo, hi_b, false);
                }
            }
        }
        // advance to the following stage
        if (++pf_xs == xstages) { pf_xs = 0; if (++pf_oy == p.ho) { pf_oy = 0; ++pf_inst; } }
    };
    auto commit = [&](unsigned char* buf, const Regs& R) {
        const uint32_t (&pa)[kMaxA][4] = R.pa;
        const uint32_t (&pb)[kMaxB][kNP] = R.pb;
#pragma unroll
        for (int s = 0; s < kMaxA; s++)
            *reinterpret_cast<uint4*>(buf + a_soff0 + s * 512) = make_uint4(pa[s][0], pa[s][1], pa[s][2], pa[s][3]);
        unsigned char* bb = buf + kAStage + b_soff0;
#pragma unroll
        for (int s = 0; s < kMaxB; s++) {
            if (s < b_items) {
                *reinterpret_cast<uint4*>(bb + s * 512) = pick8<SH>(pb[s]);
                if constexpr (KW > 1) *reinterpret_cast<uint4*>(bb + b_copy + s * 512) = pick8<SH + 1>(pb[s]);
                if constexpr (KW > 2) *reinterpret_cast<uint4*>(bb + 2 * b_copy + s * 512) = pick8<SH + 2>(pb[s]);
            }
        }
    };

    // Warp-specialised pipeline over the stages, two buffers, no block-wide barrier inside the loop:
    //   producers (warps 0-7): wait empty[b] (MMAs of stage s - 2 done) -> registers of stage s to buffer b ->
    //                          fence -> arrive full[b] -> start the loads of stage s + 2
    //   MMA warp (warp 8)    : wait full[b] -> 4 x KW MMAs of stage s -> tcgen05.commit -> empty[b]
    if (warp < kProducers / 32) {
        auto stage = [&](int s, Regs& R) {
            const int b = s & 1;
            unsigned char* buf = smem + (size_t)b * stage_bytes;
            if (s >= 2) mbar_wait(&empty_bar[b], (uint32_t)(((s >> 1) - 1) & 1));
            commit(buf, R);
            fence_proxy_async();
            mbar_arrive(&full_bar[b]);
            if (s + 2 < n_stages) prefetch(R);
        };
        prefetch(R0);
        if (n_stages > 1) prefetch(R1);
        for (int s = 0; s < n_stages; s += 2) {
            stage(s, R0);
            if (s + 1 < n_stages) stage(s + 1, R1);
        }
    } else if (lane == 0) {
        // descriptors advance by plain adds on the 16-byte-unit start-address field (shared memory < 256 KB)
        const uint64_t adesc0 = make_desc(smem_u32(smem), 2048, 128);
        const uint64_t bdesc0 = make_desc(smem_u32(smem + kAStage), (uint32_t)NT * 16, 128);
        const uint32_t stage16 = (uint32_t)stage_bytes >> 4, bk16 = (uint32_t)b_kstep >> 4, bc16 = (uint32_t)b_copy >> 4;
        int xs = 0;
        for (int s = 0; s < n_stages; s++) {
            const int b = s & 1;
            const int left = p.wo - xs * kStagePx;
            const int ksteps = ((left < kStagePx ? left : kStagePx) + 15) / 16;
            if (++xs == xstages) xs = 0;
            mbar_wait(&full_bar[b], (uint32_t)((s >> 1) & 1));
            tc_fence_after();
            const uint64_t ad = adesc0 + (uint64_t)(b * stage16), bd = bdesc0 + (uint64_t)(b * stage16);
            for (int k = 0; k < ksteps; k++) {
#pragma unroll
                for (int kx = 0; kx < KW; kx++)
                    umma_f16(tmem_d + (uint32_t)(kx * NT), ad + (uint64_t)(k * 256), bd + (uint64_t)(kx * bc16 + k * bk16), idesc,
                             (s > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[b]);
        }
    }
    if (warp == kProducers / 32) __syncwarp();        // lanes 1-31 of the MMA warp sleep here instead of polling a barrier

    // ---- epilogue: all MMAs done when the last commit has arrived
    {
        const int last = n_stages - 1;
        mbar_wait(&empty_bar[last & 1], (uint32_t)((last >> 1) & 1));
        if (n_stages >= 2) mbar_wait(&empty_bar[(last - 1) & 1], (uint32_t)(((last - 1) >> 1) & 1));
    }
    tc_fence_after();
    __syncthreads();                                   // stage buffers are free: reuse them as the output tile
    __half* tile = reinterpret_cast<__half*>(smem);
    const int row_pitch = NT * KW + 2;                 // halves; (pitch / 2) odd -> the per-row 2-byte stores spread over the banks
    {
        const int q = warp % 4;
        const int r = q * 32 + lane;                   // co row = TMEM lane
        const int ncols = warp < kProducers / 32 ? KW * NT : 0;
        for (int n0 = (warp / 4) * 32; n0 < ncols; n0 += 64) {
            uint32_t acc[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, acc);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int col = n0 + j;
                if (col < ncols) {
                    const int kx = col / NT, ci = col - kx * NT;
                    tile[r * row_pitch + ci * KW + kx] = __float2half_rn(__uint_as_float(acc[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    {
        // global: dw[((g * cout + co) * cin + ci) * KH * KW + ky * KW + kx]; lanes run over j = ci * KW + kx
        const int ci_n = min(NT, p.cin - ci0);
        const int per_row = ci_n * KW;
        for (int r = warp; r < kBM; r += kThreads / 32) {
            const int co = co0 + r;
            if (co >= p.cout) break;
            __half* dst = p.dw + (((int64_t)g * p.cout + co) * p.cin + ci0) * (KH * KW) + ky * KW;
            const __half* src = tile + r * row_pitch;
            for (int j = lane; j < per_row; j += 32) {
                const int ci = j / KW, kx = j - ci * KW;
                dst[ci * (KH * KW) + kx] = src[j];
            }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 512);
}

bool wgrad_supported(int dtype, int kh, int kw, int stride)
{
    return dtype == LVG_F16 && stride == 1 && ((kh == 3 && kw == 3) || (kh == 1 && kw == 1));
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int lvg_conv2d_wgrad(const void* x, const void* dy, void* dw, int dtype, int n, int groups, int cin, int cout,
                                int h, int wd, int kh, int kw, int stride, int pad_h, int pad_w, void* stream)
{
    LVG_REQUIRE(x && dy && dw, "conv2d_wgrad: x, dy, dw must not be NULL");
    if (!wgrad_supported(dtype, kh, kw, stride) || n < 1 || pad_h < 0 || pad_w < 0) {
        set_error("conv2d_wgrad: outside the tensor-core kernel's envelope (fp16, stride 1, 3x3 or 1x1)");
        return LVG_UNSUPPORTED;
    }
    WgradParams p;
    p.x = (const __half*)x; p.dy = (const __half*)dy; p.dw = (__half*)dw;
    p.n = n; p.groups = groups; p.cin = cin; p.cout = cout; p.h = h; p.w = wd;
    p.kh = kh; p.kw = kw; p.pad_h = pad_h; p.pad_w = pad_w;
    p.ho = h + 2 * pad_h - kh + 1;
    p.wo = wd + 2 * pad_w - kw + 1;
    LVG_REQUIRE(p.ho >= 1 && p.wo >= 1, "conv2d_wgrad: empty output");
    if ((int64_t)cin * h * wd >= (1ll << 31) || (int64_t)cout * p.ho * p.wo >= (1ll << 31)) {
        set_error("conv2d_wgrad: a group's activations exceed 32-bit offsets");
        return LVG_UNSUPPORTED;
    }
    const int nt_max = 160;        // kMaxB B items per thread cover rows rsub + 32 s < 160 (a 256-row tile for 1x1 kernels left rows 160.. unstaged)
    const int ntiles = (cin + nt_max - 1) / nt_max;
    p.nt = (((cin + ntiles - 1) / ntiles) + 15) / 16 * 16;
    const int mt = (cout + kBM - 1) / kBM;
    p.x_pair_ok = ((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (wd & 1) == 0) ? 1 : 0;
    p.dy_pair_ok = ((reinterpret_cast<uintptr_t>(dy) & 3) == 0 && (p.wo & 1) == 0) ? 1 : 0;
    LVG_REQUIRE(groups <= 65535 && mt <= 65535, "conv2d_wgrad: too many groups / channel tiles for one launch");
    const size_t stage = (size_t)kAStage + (size_t)kw * 4 * p.nt * 32;
    size_t smem = 2 * stage;
    const size_t tile_bytes = (size_t)kBM * (p.nt * kw + 2) * 2;
    if (tile_bytes > smem) smem = tile_bytes;
    if (smem < 120 * 1024) smem = 120 * 1024;        // one CTA per SM: each allocates all 512 TMEM columns
    LVG_REQUIRE(smem <= 227 * 1024, "conv2d_wgrad: tile does not fit shared memory");
    dim3 grid((unsigned)(ntiles * kh), (unsigned)mt, (unsigned)groups);
    cudaStream_t s = (cudaStream_t)stream;
    void (*k)(WgradParams) = kh == 3 ? ((pad_w & 1) ? conv_wgrad_tc_kernel<3, 3, 1> : conv_wgrad_tc_kernel<3, 3, 0>)
                                     : ((pad_w & 1) ? conv_wgrad_tc_kernel<1, 1, 1> : conv_wgrad_tc_kernel<1, 1, 0>);
    LVG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kThreads, smem, s>>>(p);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}
