// Grouped 1-D / 2-D / 3-D convolution (cross-correlation) as a TMA-fed implicit GEMM on tcgen05 tensor cores.
//
// One engine for every dense contraction on the path:
//   conv2d_gradfix.conv2d / conv_transpose2d          (torch_utils/ops/conv2d_gradfix.py:37-45; modulated convolutions of
//                                                      model/generator_sres.py:63-65, discriminator convs of conv2d_resample.py:29-41)
//   F.conv3d of the low-res networks                  (model/generator_lres.py:119,578, model/discriminator_lres.py:172)
//   F.conv1d of the low-res discriminator epilogue    (model/discriminator_lres.py:108-127)
// fp16 tensors run fp16 x fp16 -> fp32; fp32 tensors (the reference keeps TF32 off, train_lres.py:269) are split into
// bf16 hi + lo halves and accumulate hi*hi + lo*hi + hi*lo in fp32 (relative error ~2^-16, three tensor-core products
// instead of one SIMT fp32 pass).
//
// Data layout. Activations are first re-tiled (conv_pack_act_kernel, one streaming pass) from NC(T)HW to channel blocks
// of 8:   X8[instance][block][t][h][w][8]   (16 bytes per pixel and block; instance = sample x group).
// In this layout
//   * a TMA tensor map (8, W, H, T, instance*block) addresses any halo tile with hardware zero fill -- no alignment
//     constraints from odd row pitches (the fp16 NCHW rows of the super-res layers are 4 mod 8 elements long);
//   * the tile lands in shared memory as [block][row][pixel][16 B], which IS the canonical K-major no-swizzle operand
//     layout with the pixel index running linearly at 16 bytes: a filter tap (ky, kx) is nothing but a descriptor start
//     address advanced by (ky * tile_width + kx) * 16 bytes. No shifted copies, no register staging.
//   * N of one MMA runs linearly over the tile INCLUDING its halo columns; the accumulator columns that straddle two rows
//     are computed and dropped in the epilogue (kw - 1 of every tile_width columns).
// GEMM view per instance:  D[co][pix] = sum_{kt} sum_{ci} sum_{ky,kx} W[co][ci][kt][ky][kx] * X[ci][t + kt][y + ky][x + kx]
//   M = 128 output channels, N = TH x (WT + kw - 1) pixels (<= 256 TMEM columns, 2 CTAs per SM),
//   K = 16 channels per MMA; the K loop runs over (kt, 16-channel step), every stage issues kh*kw MMAs (one per tap).
// Weights are re-tiled per call into 128 x 16 K-major images per (m-tile, k-step, tap) (conv_pack_w_kernel) and arrive
// with one cp.async.bulk per stage.
//
// Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane, owns TMEM), warps 2-5 = epilogue
// (TMEM -> registers -> [bias, lrelu, gain, clamp] -> NC(T)HW global). Stages hand over through full/empty mbarriers.

#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace lvg {
// conv_pointwise.cu: streaming fp32 kernels for 1x1x1 convolutions with few channels (HBM-bound; the engine would re-tile and pad)
bool pw_supported(int dtype, int groups, int cin, int cout, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride, int64_t P, int wgrad);
int pw_conv(const float* x, const float* w, float* y, int n, int cin, int cout, int64_t P, int64_t w_sco, int64_t w_sci, cudaStream_t s);
int64_t pw_wgrad_workspace(int cin, int cout);
int pw_wgrad(const float* x, const float* dy, float* dw, int n, int cin, int cout, int64_t P, void* workspace, int64_t workspace_bytes, cudaStream_t s);
}

namespace lvg {
namespace {

using namespace tc;

constexpr int kBM = 128;
constexpr int kATile = kBM * 16 * 2;          // one 128 x 16 weight image: 4096 bytes
constexpr int kThreads = 192;
constexpr int kEpiWarps = 8;                          // conv_igemm_kernel: two epilogue warps per TMEM lane quadrant (they alternate 32-column blocks)
constexpr int kIgemmThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxStages = 6;

struct IgemmParams {
    const unsigned char* wp;     // packed weights [wgroups][mt][kc][taps_all][4096]
    void* y;
    const float* bias;           // per output channel (group-local index g*cout + co), or nullptr
    int act;                     // 0 = none, 1 = (x + b) * gain clamped, 2 = lrelu(x + b, alpha) * gain clamped
    float alpha, gain, clamp;    // clamp < 0: none
    int out_f32;
    int bf16;                    // operands are bf16 (split fp32) instead of fp16
    int wgroups, cout, mt, kc;   // kc = 16-channel steps of the packed K axis
    int nblk;                    // channel blocks per instance in X8
    int nimg;                    // operand images per k-step: 1 (fp16) or 2 (split: hi and lo halves of both operands)
    int lo_blk;                  // split: block offset of the lo halves inside an instance of X8
    int to, ho, wo;              // output extent
    int kt, kh, kw, pad_t, pad_h, pad_w;
    int tt, th, wt, wtb, thb;    // tile frames / rows / cols, box cols / rows
    int frame_px;                // thb * wtb: accumulator columns from one frame of the tile to the next
    int ncols;                   // accumulator columns (multiple of 16, <= 512)
    int n0;                      // columns of the first MMA of a tap (the second one takes ncols - n0; 0 = single MMA)
    int epi_warps;               // epilogue warps that take part (4 or 8)
    int nbuf;                    // accumulator buffers in TMEM: 2 when ncols <= 256 (epilogue of tile i overlaps tile i + 1)
    int tiles_x, tiles_y, tiles_t;
    int64_t total_tiles;         // tiles_x * tiles_y * tiles_t * mt * instances
    int ks;                      // k-steps per stage (> 1 only when kt == 1)
    int stages;
    int a_resident;              // the ring length is a multiple of the stages per tile and every tile of the launch uses the same weights:
                                 // slot s always holds the same weight images -- they are fetched for the first `stages` iterations only
    int a_stage, b_step, b_bytes, b_box, stage_bytes;    // bytes: A per stage, B stride per k-step / per pair of blocks, TMA payload of a pair, whole stage
    int64_t y_cs;                // output channel stride (= to*hos*wos)
    int ostride, hos, wos;       // output decimation (strided convolution): only rows / columns divisible by ostride are stored
};

// ------------------------------------------------------------------------------------------------ re-tiling passes

__device__ __forceinline__ unsigned short bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(unsigned short b) { return __uint_as_float((uint32_t)b << 16); }

// NC(T)HW -> X8. One thread = one pixel of one channel block: 8 strided reads (coalesced across the warp), one or two
// 16-byte writes. SPLIT: fp32 in, bf16 hi blocks [0, cblk) and lo blocks [cblk, 2 cblk) out.
// Dilation (gradients of strided convolutions): input pixel (t, i, j) of an ih x iw image lands at (t, i*dil, j*dil) of
// the oh x ow output image, which the caller has zeroed.
struct PackGeom { int64_t thw_in, thw_out; int ih, iw, oh, ow, dil; };

template <class TIn, bool SPLIT>
__global__ void __launch_bounds__(256) conv_pack_act_kernel(const TIn* __restrict__ x, uint4* __restrict__ y, int64_t inst, int c,
                                                             int cblk, PackGeom gm)
{
    const int64_t thw = gm.thw_in;
    const int64_t total = inst * cblk * thw;
    const int nblk = SPLIT ? 2 * cblk : cblk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i % thw;
        const int64_t r = i / thw;
        const int blk = (int)(r % cblk);
        const int64_t in = r / cblk;
        const int c0 = blk * 8;
        const TIn* src = x + ((int64_t)in * c + c0) * thw + p;
        int64_t po = p;
        if (gm.dil > 1) {
            const int j = (int)(p % gm.iw);
            const int64_t q = p / gm.iw;
            const int ii = (int)(q % gm.ih);
            const int64_t t = q / gm.ih;
            po = (t * gm.oh + (int64_t)ii * gm.dil) * gm.ow + (int64_t)j * gm.dil;
        }
        if constexpr (!SPLIT) {
            alignas(16) unsigned short v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (c0 + j < c) ? __half_as_ushort(__ldg(reinterpret_cast<const __half*>(src) + (int64_t)j * thw)) : (unsigned short)0;
            y[((int64_t)in * nblk + blk) * gm.thw_out + po] = *reinterpret_cast<const uint4*>(v);
        } else {
            alignas(16) unsigned short hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float f = (c0 + j < c) ? __ldg(reinterpret_cast<const float*>(src) + (int64_t)j * thw) : 0.f;
                hi[j] = bf16_bits(f);
                lo[j] = bf16_bits(f - bf16_val(hi[j]));
            }
            y[((int64_t)in * nblk + blk) * gm.thw_out + po] = *reinterpret_cast<const uint4*>(hi);
            y[((int64_t)in * nblk + cblk + blk) * gm.thw_out + po] = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

// launches the re-tiling of one activation tensor (zeroing the target first when it is dilated)
int pack_act(const void* x, void* x8, int split, int64_t inst, int c, int cblk, int t, int ih, int iw, int oh, int ow, int dil, cudaStream_t s)
{
    PackGeom gm;
    gm.thw_in = (int64_t)t * ih * iw; gm.thw_out = (int64_t)t * oh * ow; gm.ih = ih; gm.iw = iw; gm.oh = oh; gm.ow = ow; gm.dil = dil;
    if (dil > 1) LVG_CUDA(cudaMemsetAsync(x8, 0, (size_t)(inst * (split ? 2 : 1) * cblk * gm.thw_out * 16), s));
    const int64_t total = inst * cblk * gm.thw_in;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 64;
    if (blocks > cap) blocks = cap;
    if (split) conv_pack_act_kernel<float, true><<<(unsigned)blocks, 256, 0, s>>>((const float*)x, (uint4*)x8, inst, c, cblk, gm);
    else conv_pack_act_kernel<__half, false><<<(unsigned)blocks, 256, 0, s>>>((const __half*)x, (uint4*)x8, inst, c, cblk, gm);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

// weights -> tile images.  Element (m, k, tap) of the logical A matrix of group g sits at
//   w[g * gstride + m * sm + k * sk + (flip ? taps-1-tap : tap)]
// fprop: m = co, k = ci (sm = cin*taps, sk = taps); dgrad: m = ci, k = co (sm = taps, sk = cin*taps), taps mirrored.
// Image of one 128 x 16 tile (K-major, no swizzle): byte offset(m, k) = (k/8)*2048 + (m/8)*128 + (m%8)*16 + (k%8)*2.
// SPLIT: every (k-step, tap) gets TWO images, the bf16 hi and lo halves of the fp32 weights; the main loop pairs them with
// the hi / lo activation blocks: hi*hi + hi*lo + lo*hi (each operand half is fetched once per k-step).
// One CTA = one (group, m-tile, k-step). The 128 x 16 x taps source elements form contiguous RUNS in memory (fprop: one
// run of 16*taps elements per output channel; dgrad: one run of 128*taps elements per k): a warp copies a run into shared
// memory with aligned 4-byte loads (no per-element index arithmetic), then every thread assembles one 16-byte image row
// from 8 shared-memory reads and consecutive threads write consecutive rows (512 contiguous bytes per warp).
// Rows of the 128-row weight image (= TMEM lanes of the accumulator). An epilogue warp can only read the 32 lanes of its
// quadrant, so an m-tile with fewer than 128 output channels spreads them evenly over the four quadrants (`per` channels at
// the start of each) instead of filling quadrant after quadrant: all epilogue warps share the store work of a 32- or
// 64-channel layer. Channels >= 4 * per (zero rows) take the remaining rows in order.
__host__ __device__ __forceinline__ int m_rows_per_quadrant(int channels_left)
{
    const int cv = channels_left < kBM ? (channels_left > 0 ? channels_left : 0) : kBM;
    return cv >= kBM ? 32 : (cv + 3) / 4 > 0 ? (cv + 3) / 4 : 1;
}
__host__ __device__ __forceinline__ int m_row_of_channel(int ch, int per)
{
    if (per >= 32) return ch;
    if (ch < 4 * per) return (ch / per) * 32 + ch % per;
    const int r = ch - 4 * per;
    return (r / (32 - per)) * 32 + per + r % (32 - per);
}

template <class TIn, bool SPLIT>
__global__ void __launch_bounds__(256) conv_pack_w_kernel(const TIn* __restrict__ w, unsigned char* __restrict__ wp, int m_total, int k_total,
                                                           int kpad, int taps, int64_t gstride, int64_t sm, int64_t sk, int flip, int mt, int kc,
                                                           int rows_per_pass, int rows_per_cta)
{
    extern __shared__ uint32_t sw32[];               // [runs][pitch] words, then the runs' element offsets
    constexpr int ES = (int)sizeof(TIn);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int kci = blockIdx.x % kc;
    const int mti = (blockIdx.x / kc) % mt;
    const int g = blockIdx.x / (kc * mt);
    const int k0 = kci * 16;
    (void)kpad;
    constexpr int NIMG = SPLIT ? 2 : 1;
    const TIn* wg = w + (int64_t)g * gstride;
    const bool mrows = sk < sm;                      // fprop layout: runs along (k, tap) per m; else runs along (m, tap) per k
    const int R = rows_per_pass;
    const int run_el = mrows ? 16 * taps : R * taps;
    const int pitch = ((run_el * ES + 2 + 3) / 4) | 1;
    const int max_runs = mrows ? R : 16;
    int* s_off = reinterpret_cast<int*>(sw32 + (size_t)max_runs * pitch);
    unsigned char* dst0 = wp + ((((int64_t)g * mt + mti) * kc + kci) * taps) * (int64_t)(NIMG * kATile);
    const int kvalid = max(0, min(16, k_total - k0));
    const int per = m_rows_per_quadrant(m_total - mti * kBM);
    // blockIdx.y = chunk of rows_per_cta image rows (small layers have only a handful of (group, m-tile, k-step) blocks: the
    // row chunks spread their latency-bound staging over more SMs)
    const int row_end = min(kBM, ((int)blockIdx.y + 1) * rows_per_cta);
    for (int r0 = (int)blockIdx.y * rows_per_cta; r0 < row_end; r0 += R) {
        const int Rn = min(R, row_end - r0);
        const int m0 = mti * kBM + r0;
        const int mvalid = max(0, min(Rn, m_total - m0));
        const int len = mrows ? kvalid * taps : mvalid * taps;             // valid elements of a run
        const int nruns = len > 0 ? (mrows ? mvalid : kvalid) : 0;
        for (int run = warp; run < nruns; run += 8) {
            const int64_t e0 = mrows ? (int64_t)(m0 + run) * sm + (int64_t)k0 * sk : (int64_t)(k0 + run) * sk + (int64_t)m0 * sm;
            const unsigned char* gb = reinterpret_cast<const unsigned char*>(wg + e0);
            const int a = (int)(reinterpret_cast<uintptr_t>(gb) & 3);       // 0, or 2 for an odd fp16 element offset
            const uint32_t* gw = reinterpret_cast<const uint32_t*>(gb - a);
            const int nbytes = len * ES + a;
            const int nwords = (nbytes + 3) / 4;
            // whole words strictly inside the run: four independent loads in flight per lane
            const int w_lo = a != 0 ? 1 : 0, w_hi = (nbytes & 3) != 0 ? nwords - 1 : nwords;
            int wi = w_lo + lane;
            for (; wi + 96 < w_hi; wi += 128) {
                const uint32_t v0 = __ldg(gw + wi), v1 = __ldg(gw + wi + 32), v2 = __ldg(gw + wi + 64), v3 = __ldg(gw + wi + 96);
                uint32_t* d = sw32 + run * pitch + wi;
                d[0] = v0; d[32] = v1; d[64] = v2; d[96] = v3;
            }
            for (; wi < w_hi; wi += 32) sw32[run * pitch + wi] = __ldg(gw + wi);
            if (lane == 0 && a != 0) sw32[run * pitch] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(gb)) << 16;              // do not touch bytes before the run
            if (lane == 1 && w_hi < nwords && !(a != 0 && nwords == 1))
                sw32[run * pitch + nwords - 1] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(gw + nwords - 1));              // ... or after it
            if (lane == 0) s_off[run] = a / ES;
        }
        __syncthreads();
        const bool pow2 = Rn == kBM;
        for (int o = threadIdx.x; o < taps * 2 * Rn; o += 256) {
            int mrow, k8, tap;
            if (pow2) { mrow = o & (kBM - 1); k8 = (o >> 7) & 1; tap = o >> 8; }
            else { mrow = o % Rn; k8 = (o / Rn) % 2; tap = o / (2 * Rn); }
            const int wtap = flip ? taps - 1 - tap : tap;
            alignas(16) unsigned short v[8], vlo[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = k8 * 8 + j;
                const int run = mrows ? mrow : k;
                const int idx = mrows ? k * taps + wtap : mrow * taps + wtap;
                float f = 0.f;
                unsigned short hbits = 0;
                if (mrow < mvalid && k < kvalid) {
                    if constexpr (ES == 2) hbits = reinterpret_cast<const unsigned short*>(sw32 + run * pitch)[s_off[run] + idx];
                    else f = __uint_as_float(sw32[run * pitch + idx]);
                }
                if constexpr (SPLIT) {
                    if constexpr (ES == 2) f = __half2float(__ushort_as_half(hbits));
                    const unsigned short h = bf16_bits(f);
                    v[j] = h;
                    vlo[j] = bf16_bits(f - bf16_val(h));
                } else {
                    v[j] = ES == 2 ? hbits : __half_as_ushort(__float2half_rn(f));
                }
            }
            unsigned char* dimg = dst0 + (size_t)tap * (NIMG * kATile) + k8 * 2048 + m_row_of_channel(r0 + mrow, per) * 16;
            *reinterpret_cast<uint4*>(dimg) = *reinterpret_cast<const uint4*>(v);
            if constexpr (SPLIT) *reinterpret_cast<uint4*>(dimg + kATile) = *reinterpret_cast<const uint4*>(vlo);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ main kernel

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
                 : "memory");
}

// X8 seen as 8-byte elements: (2*W, H, T, instance*block). A pixel of a block is two elements, so a box row is one
// contiguous run of 16*width bytes (with a 16-byte innermost dimension TMA moves one pixel per request).
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}

struct TileCoord { int ox0, oy0, t0, mti, inst; };

__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int64_t L)
{
    TileCoord c;
    int64_t r = L;
    c.ox0 = (int)(r % p.tiles_x) * p.wt; r /= p.tiles_x;
    c.oy0 = (int)(r % p.tiles_y) * p.th; r /= p.tiles_y;
    c.t0 = (int)(r % p.tiles_t) * p.tt; r /= p.tiles_t;
    c.mti = (int)(r % p.mt);
    c.inst = (int)(r / p.mt);
    return c;
}

// Persistent: CTA b works on tiles b, b + gridDim.x, ... (pixel tile fastest, so that concurrently running CTAs share the
// weight tiles of one (instance, m-tile) in L2). The operand ring runs across tile boundaries; with <= 256 accumulator
// columns two TMEM buffers alternate, so the epilogue of tile i overlaps the main loop of tile i + 1.
__global__ void __launch_bounds__(kIgemmThreads, 1) conv_igemm_kernel(const __grid_constant__ CUtensorMap tmx, const IgemmParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int taps2 = p.kh * p.kw;
    const int kchunks = (p.kc + p.ks - 1) / p.ks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], (uint32_t)p.epi_warps); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int64_t L = blockIdx.x; L < p.total_tiles; L += gridDim.x) {
                const TileCoord c = decode_tile(p, L);
                const unsigned char* wpg = p.wp + (((int64_t)(c.inst % p.wgroups) * p.mt + c.mti) * p.kc) * (int64_t)(p.kt * taps2) * (p.nimg * kATile);
                const int blk0 = c.inst * p.nblk;
                for (int kt = 0; kt < p.kt; kt++) {
                    for (int kcix = 0; kcix < kchunks; kcix++, it++) {
                        const int s = it % p.stages;
                        if (it >= p.stages) mbar_wait(&empty_bar[s], (uint32_t)((it / p.stages - 1) & 1));
                        unsigned char* st = smem + (size_t)s * p.stage_bytes;
                        const int k0 = kcix * p.ks;
                        const int nks = min(p.ks, p.kc - k0);
                        const bool load_a = !p.a_resident || it < p.stages;
                        const uint32_t a_bytes = load_a ? (uint32_t)(nks * taps2 * p.nimg * kATile) : 0u;
                        mbar_expect_tx(&full_bar[s], a_bytes + (uint32_t)(nks * p.nimg * p.b_box));
                        // A: kt == 1 -> the nks steps' tiles are contiguous; kt > 1 -> ks == 1, the taps of this kt are contiguous
                        if (load_a) bulk_copy_g2s(st, wpg + ((int64_t)k0 * p.kt + kt) * (int64_t)taps2 * (p.nimg * kATile), a_bytes, &full_bar[s]);
                        for (int j = 0; j < nks; j++) {
                            const int kb = (k0 + j) * 2;
                            for (int im = 0; im < p.nimg; im++)
                                tma_load_4d(st + p.a_stage + (size_t)j * p.b_step + (size_t)im * p.b_bytes, &tmx, 2 * (c.ox0 - p.pad_w), c.oy0 - p.pad_h,
                                            c.t0 + kt - p.pad_t, blk0 + im * p.lo_blk + kb, &full_bar[s]);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptors: D = f32, A and B K-major, fp16 or bf16 operands, N >> 3, M >> 4
            const uint32_t ibase = (1u << 4) | (p.bf16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(kBM >> 4) << 24);
            const int na = p.n0 > 0 ? p.n0 : p.ncols, nb = p.n0 > 0 ? p.ncols - p.n0 : 0;
            const uint32_t idesc_a = ibase | ((uint32_t)(na >> 3) << 17), idesc_b = ibase | ((uint32_t)(nb >> 3) << 17);
            const uint32_t blk_bytes = (uint32_t)p.b_box / 2;
            int it = 0, i = 0;
            for (int64_t L = blockIdx.x; L < p.total_tiles; L += gridDim.x, i++) {
                const int buf = i % p.nbuf;
                if (i >= p.nbuf) mbar_wait(&acc_empty[buf], (uint32_t)((i / p.nbuf - 1) & 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256;
                bool first = true;
                for (int kt = 0; kt < p.kt; kt++) {
                    for (int kcix = 0; kcix < kchunks; kcix++, it++) {
                        const int s = it % p.stages;
                        mbar_wait(&full_bar[s], (uint32_t)((it / p.stages) & 1));
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
                        const int nks = min(p.ks, p.kc - kcix * p.ks);
                        // descriptors as (lo, hi) words, stepped with 32-bit adds on the low word (address >> 4): the next tap's weight
                        // image(s) + nimg * 4096 bytes, the next tap column + 16 bytes, the next tap row + wtb * 16 bytes
                        const uint32_t a_hi = desc_hi(128), b_hi = desc_hi(128);
                        const uint32_t a_tap = (uint32_t)(p.nimg * kATile) >> 4, b_lo_img = (uint32_t)p.b_bytes >> 4, b_nb = (uint32_t)na;
                        // wide tiles (two MMAs per tap, columns [0, na) and [na, ncols)): all taps of one half, then the other --
                        // consecutive MMAs keep accumulating into the same tensor-memory columns (switching them costs ~60 clocks)
                        for (int j = 0; j < nks; j++) {
                            const uint32_t a_base = desc_lo(st + (uint32_t)(j * taps2 * p.nimg) * kATile, 2048);
                            const uint32_t b_base = desc_lo(st + (uint32_t)p.a_stage + (uint32_t)j * (uint32_t)p.b_step, blk_bytes);
                            for (int half = 0; half < (nb > 0 ? 2 : 1); half++) {
                                const uint32_t tm = tmem_d + (half ? (uint32_t)na : 0u);
                                const uint32_t id = half ? idesc_b : idesc_a;
                                uint32_t a_lo = a_base, b_row = b_base + (half ? b_nb : 0u);
                                uint32_t acc = first ? 0u : 1u;
                                for (int ky = 0; ky < p.kh; ky++, b_row += (uint32_t)p.wtb) {
                                    uint32_t b_lo = b_row;
                                    for (int kx = 0; kx < p.kw; kx++, b_lo++, a_lo += a_tap) {
                                        // fp16: one product; split: hi*hi, hi*lo, lo*hi (A image hi, hi, lo; B image hi, lo, hi)
                                        umma_f16_w(tm, a_lo, a_hi, b_lo, b_hi, id, acc);
                                        acc = 1u;
                                        if (p.nimg == 2) {
                                            umma_f16_w(tm, a_lo, a_hi, b_lo + b_lo_img, b_hi, id, 1u);
                                            umma_f16_w(tm, a_lo + (kATile >> 4), a_hi, b_lo, b_hi, id, 1u);
                                        }
                                    }
                                }
                            }
                            first = false;
                        }
                        umma_commit(&empty_bar[s]);
                    }
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp - 2 < p.epi_warps) {
        // ---- epilogue warps: TMEM lane quadrant = warp % 4 (a warp can only read its own 32 lanes = 32 output channels of the
        // tile); the two warps of a quadrant alternate the 32-column blocks. Each warp has a private 32 x 33 transposition
        // buffer so that a store instruction covers 32 consecutive pixels of one channel. With short K loops (1x3x3 / 1x1x1
        // layers, few channels) this epilogue, not the MMA loop, bounds the kernel (ncu on the 32 -> 64 layer of the low-res
        // discriminator: tensor pipe 30 % active with ONE latency-bound warp per quadrant storing channel by channel), hence
        // two warps per quadrant (p.epi_warps = 8) for short K loops and eight independent shared-memory reads in flight
        // before their stores. Long K loops keep four warps: the second set's transposition buffers would cost a pipeline stage.
        const int q = warp % 4, half = (warp - 2) / 4, nhalf = p.epi_warps / 4;
        float* sT = reinterpret_cast<float*>(smem + (size_t)p.stages * p.stage_bytes + 512) + (warp - 2) * (32 * 33);
        int i = 0;
        for (int64_t L = blockIdx.x; L < p.total_tiles; L += gridDim.x, i++) {
            const TileCoord c = decode_tile(p, L);
            const int buf = i % p.nbuf;
            mbar_wait(&acc_full[buf], (uint32_t)((i / p.nbuf) & 1));
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * 256;
            const int per = m_rows_per_quadrant(p.cout - c.mti * kBM);     // channels per lane quadrant (conv_pack_w_kernel's row order)
            const int m0 = c.mti * kBM + q * per;
            const int rows_ok = min(per, p.cout - m0);               // channels of this warp that exist (lanes 0 .. rows_ok - 1)
            const int m_lane = m0 + lane;                            // this lane's channel while the values are in registers
            const float b = (p.bias != nullptr && lane < rows_ok) ? __ldg(p.bias + (int64_t)(c.inst % p.wgroups) * p.cout + m_lane) : 0.f;
            const int64_t ch0 = ((int64_t)c.inst * p.cout + m0) * p.y_cs;
            for (int n0 = 32 * half; n0 < p.ncols && rows_ok > 0; n0 += 32 * nhalf) {
                uint32_t acc[32];
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, acc);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    float v = __uint_as_float(acc[j]);
                    if (p.act) {
                        v += b;
                        if (p.act == 2) v = v < 0.f ? v * p.alpha : v;
                        v *= p.gain;
                        if (p.clamp >= 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
                    }
                    sT[j * 33 + lane] = v;
                }
                __syncwarp();
                // lane = accumulator column n0 + lane -> (frame, row, col) of the tile
                const int n = n0 + lane;
                const int f = n / p.frame_px, rem = n - f * p.frame_px;
                const int r = rem / p.wtb, cc = rem - r * p.wtb;
                const int ot = c.t0 + f, oy = c.oy0 + r, ox = c.ox0 + cc;
                bool ok = n < p.ncols && f < p.tt && r < p.th && cc < p.wt && ot < p.to && oy < p.ho && ox < p.wo;
                int64_t off;
                if (p.ostride == 1) {
                    off = ch0 + ((int64_t)ot * p.ho + oy) * p.wo + ox;
                } else {
                    ok = ok && (oy % p.ostride == 0) && (ox % p.ostride == 0);
                    off = ch0 + ((int64_t)ot * p.hos + oy / p.ostride) * p.wos + ox / p.ostride;
                }
                if (ok) {
                    const float* src = sT + lane * 33;
                    if (p.out_f32) {
                        float* y = reinterpret_cast<float*>(p.y) + off;
                        for (int row0 = 0; row0 < rows_ok; row0 += 8, y += 8 * p.y_cs) {
                            float v8[8];
#pragma unroll
                            for (int k = 0; k < 8; k++) v8[k] = src[row0 + k];
#pragma unroll
                            for (int k = 0; k < 8; k++) if (row0 + k < rows_ok) y[(int64_t)k * p.y_cs] = v8[k];
                        }
                    } else {
                        __half* y = reinterpret_cast<__half*>(p.y) + off;
                        for (int row0 = 0; row0 < rows_ok; row0 += 8, y += 8 * p.y_cs) {
                            float v8[8];
#pragma unroll
                            for (int k = 0; k < 8; k++) v8[k] = src[row0 + k];
#pragma unroll
                            for (int k = 0; k < 8; k++) if (row0 + k < rows_ok) y[(int64_t)k * p.y_cs] = __float2half_rn(v8[k]);
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// tensor map over a channel-block-of-8 tensor [blocks][t][h (pitch_hw / pitch_w rows)][w][8 x 2 bytes], seen as 8-byte
// elements: (2 w, h, t, blocks); box = box_w pixels x box_h rows x box_t frames x box_blk blocks
int encode_map(CUtensorMap* tm, void* base, int w, int h, int t, int64_t blocks, int64_t pitch_w, int64_t pitch_hw, int64_t pitch_thw, int box_w,
               int box_h, int box_t, int box_blk)
{
    EncodeTiledFn enc = encode_fn();
    LVG_REQUIRE(enc != nullptr, "convnd: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {(cuuint64_t)w * 2, (cuuint64_t)h, (cuuint64_t)t, (cuuint64_t)blocks};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch_w * 16, (cuuint64_t)pitch_hw * 16, (cuuint64_t)pitch_thw * 16};
    const cuuint32_t box[4] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, (cuuint32_t)box_t, (cuuint32_t)box_blk};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    LVG_REQUIRE(box_w <= 128 && box_h <= 256 && box_t <= 256 && box_blk <= 256, "convnd: TMA box too large");
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LVG_REQUIRE(r == CUDA_SUCCESS, "convnd: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return LVG_OK;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

inline int env_flag(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

struct Geometry {
    int cpad, cblk, nblk, nimg, kc, mt;
    int64_t act_bytes, w_bytes;
};

// K-side geometry for `ck` contraction channels and `cm` output channels per group
Geometry geometry(int split, int64_t inst, int groups, int ck, int cm, int64_t thw, int taps)
{
    Geometry g;
    g.cpad = round_up(ck, 16);
    g.cblk = g.cpad / 8;
    g.nblk = split ? 2 * g.cblk : g.cblk;
    g.nimg = split ? 2 : 1;
    g.kc = g.cpad / 16;
    g.mt = (cm + kBM - 1) / kBM;
    g.act_bytes = inst * g.nblk * thw * 16;
    g.w_bytes = (int64_t)groups * g.mt * g.kc * taps * g.nimg * kATile;
    return g;
}

// shared driver of fprop and dgrad: `x` has `ck` channels per group, `y` gets `cm`; logical A[m][k][tap] = w[g*gs + m*sm + k*sk + tap']
// `dil` > 1: x is given as an xin_h x xin_w image that is placed on every dil-th pixel of the h x wd grid (input gradient of
// a strided convolution); `ostride` > 1: only every ostride-th output row / column is stored (strided forward convolution).
// lvg_convnd_plan: run_igemm stops after its tile selection (pure host arithmetic, no CUDA call) and hands the parameters out
thread_local IgemmParams* g_plan_only = nullptr;

int run_igemm(const void* x, const void* w, void* y, int dtype, int n, int groups, int ck, int cm, int t, int h, int wd, int kt, int kh,
              int kw, int pad_t, int pad_h, int pad_w, int64_t w_gs, int64_t w_sm, int64_t w_sk, int flip, const float* bias, int act,
              float alpha, float gain, float clamp, int xin_h, int xin_w, int dil, int ostride, const unsigned char* x8_pre, void* workspace,
              int64_t workspace_bytes, cudaStream_t s)
{
    // `x8_pre` != nullptr: x is already re-tiled there (lvg_convnd_backward shares dy8 with the weight gradient); the workspace
    // then only holds the packed weights
    const int split = dtype == LVG_F32 ? 1 : 0;
    const int taps = kt * kh * kw;
    const int64_t inst = (int64_t)n * groups;
    const int64_t thw = (int64_t)t * h * wd;
    const Geometry g = geometry(split, inst, groups, ck, cm, thw, taps);
    const bool plan_only = g_plan_only != nullptr;
    LVG_REQUIRE(plan_only || (workspace && workspace_bytes >= (x8_pre ? 0 : g.act_bytes) + g.w_bytes + 256), "convnd: workspace too small");
    LVG_REQUIRE(plan_only || aligned16(workspace), "convnd: workspace must be 16-byte aligned");
    LVG_REQUIRE(inst * g.nblk < (1ll << 31), "convnd: too many channel blocks for a tensor map");
    EncodeTiledFn enc = plan_only ? nullptr : encode_fn();
    LVG_REQUIRE(plan_only || enc != nullptr, "convnd: cuTensorMapEncodeTiled is not available from this driver");

    unsigned char* wp = reinterpret_cast<unsigned char*>(workspace);
    unsigned char* x8 = x8_pre ? const_cast<unsigned char*>(x8_pre) : wp + ((g.w_bytes + 127) / 128) * 128;

    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.wp = wp; p.y = y; p.bias = bias; p.act = act; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
    p.out_f32 = split; p.bf16 = split;
    p.wgroups = groups; p.cout = cm; p.mt = g.mt; p.kc = g.kc; p.nblk = g.nblk;
    p.nimg = g.nimg; p.lo_blk = g.cblk;
    p.to = t + 2 * pad_t - kt + 1; p.ho = h + 2 * pad_h - kh + 1; p.wo = wd + 2 * pad_w - kw + 1;
    LVG_REQUIRE(p.to >= 1 && p.ho >= 1 && p.wo >= 1, "convnd: empty output");
    p.kt = kt; p.kh = kh; p.kw = kw; p.pad_t = pad_t; p.pad_h = pad_h; p.pad_w = pad_w;
    // tile: whole rows (and, for small frames, several frames) up to 512 accumulator columns; wide images are cut into
    // column tiles. A tile of <= 256 columns leaves room for two accumulator buffers (epilogue overlap).
    // (short K loops -- up to ~160 MMAs per tile -- are epilogue-bound: they take <= 256 columns and alternate two accumulator
    // buffers; measured on the sres discriminator shapes: 0.173 vs 0.200 ms at 256 -> 256 channels 64x64)
    const char* cb_env = getenv("LVG_CONV_COLS");
    const int taps2 = kh * kw;
    int epi_bytes = 4 * 32 * 33 * 4 + 512;              // four epilogue warps; eight when that costs neither tile size nor the second stage (below)
    int smem_budget = 224 * 1024 - epi_bytes;
    p.ostride = ostride;
    p.hos = (p.ho - 1) / ostride + 1; p.wos = (p.wo - 1) / ostride + 1;
    p.ks = (kt == 1) ? (taps2 == 1 ? 4 : (taps2 <= 3 ? 2 : 1)) : 1;
    if (p.ks > g.kc) p.ks = g.kc;
    p.a_stage = p.ks * taps2 * g.nimg * kATile;
    // the column budget shrinks until two stages fit shared memory (split precision doubles both operands of a stage)
    // (fp32 / split precision: 256 columns whatever the K loop, since the MMA issue path costs 4 instructions per MMA: two
    // alternating accumulators -- the epilogue of tile i under the main loop of tile i + 1 -- beat the better weight-tile
    // amortisation of 512-column tiles on every layer class of a low-res step, 72.7 -> 69.8 ms forward, 55.6 -> 54.2 ms input
    // gradients; the small-frame layers also get more tiles than SMs: 512 channels at 3x4 pixels 0.30 -> 0.22 ms. fp16 layers
    // (one product per tap against the same weight bytes per tap: the weight tile weighs 1.5x more) keep 512 columns for long
    // K loops, the configuration their super-res numbers were measured with. LVG_CONV_COLS overrides both.)
    for (int col_budget = cb_env ? atoi(cb_env) : ((split || g.kc * kt * taps2 <= 160) ? 256 : 512);; col_budget -= 64) {
        LVG_REQUIRE(col_budget >= 64, "convnd: no tile fits shared memory");
        const int max_wt = 128 - (kw - 1);                    // a TMA box row is at most 256 8-byte elements
        p.tiles_x = (p.wo + max_wt - 1) / max_wt;
        p.wt = (p.wo + p.tiles_x - 1) / p.tiles_x;
        p.wtb = p.wt + kw - 1;
        p.th = col_budget / p.wtb;
        if (p.th > p.ho) p.th = p.ho;
        if (p.th < 1) p.th = 1;
        p.tiles_y = (p.ho + p.th - 1) / p.th;
        p.th = (p.ho + p.tiles_y - 1) / p.tiles_y;           // balance the row tiles
        if (ostride > 1) {                                     // tile origins on the output lattice
            if (p.tiles_x > 1) { p.wt = round_up(p.wt, ostride); p.wtb = p.wt + kw - 1; p.tiles_x = (p.wo + p.wt - 1) / p.wt; if (p.th * p.wtb > col_budget) p.th = col_budget / p.wtb; }
            if (p.th < p.ho) { p.th = p.th / ostride * ostride; if (p.th < ostride) p.th = ostride; }
            p.tiles_y = (p.ho + p.th - 1) / p.th;
        }
        p.thb = p.th + kh - 1;
        p.frame_px = p.thb * p.wtb;
        p.tt = 1;
        if (p.tiles_y == 1 && p.tiles_x == 1) {                // whole frames: take as many as fit
            while (p.tt < p.to && p.tt * p.frame_px + p.th * p.wtb <= col_budget && p.tt < 64) p.tt++;
        }
        p.tiles_t = (p.to + p.tt - 1) / p.tt;
        p.tt = (p.to + p.tiles_t - 1) / p.tiles_t;
        p.ncols = round_up((p.tt - 1) * p.frame_px + p.th * p.wtb, 16);
        // TMA writes the two 8-channel blocks of a k-step densely: block 1 starts tt*thb*wtb*16 bytes after block 0 (= LBO);
        // every pair of blocks starts at a 128-byte multiple (TMA destination alignment)
        p.b_box = 2 * p.tt * p.frame_px * 16;                  // one pair of blocks as TMA writes it
        p.b_bytes = round_up(p.b_box, 128);
        p.b_step = g.nimg * p.b_bytes;
        p.stage_bytes = round_up(p.a_stage + p.ks * p.b_step + 512, 128);      // + slack: the last taps read a few pixels past the tile
        if (p.ncols <= 512 && 2 * p.stage_bytes <= smem_budget) break;
        if (p.th == 1 && p.tt == 1 && col_budget <= p.wtb) { LVG_REQUIRE(false, "convnd: a one-row tile does not fit shared memory"); }
    }
    LVG_REQUIRE(p.th >= 1 && p.ncols <= 512 && p.ncols >= 16, "convnd: tile geometry");
    // epilogue warps: eight when their extra transposition buffers cost no pipeline stage (every layer gains: the fp16
    // 539 -> 512 layer 1.29 -> 1.17 ms), or when the K loop is short anyway (epilogue-bound layers: 32 -> 32 1x3x3 of the
    // low-res discriminator 1.81 -> 1.22 ms); else four (the fp32 512-channel 3x3x3 layer loses a stage: 3.9 -> 6.5 ms with eight)
    auto stages_for = [&](int budget) { int st = 2; while (st < kMaxStages && (st + 1) * p.stage_bytes <= budget) st++; return st; };
    const int budget8 = 224 * 1024 - (kEpiWarps * 32 * 33 * 4 + 512);
    p.epi_warps = 4;
    if (2 * p.stage_bytes <= budget8 && (stages_for(budget8) == stages_for(smem_budget) || g.kc * kt * taps2 <= 160)) {
        p.epi_warps = kEpiWarps;
        epi_bytes = kEpiWarps * 32 * 33 * 4 + 512;
        smem_budget = budget8;
    }
    p.nbuf = p.ncols <= 256 ? 2 : 1;
    p.n0 = p.ncols <= 256 ? 0 : round_up(p.ncols / 2, 16);
    p.stages = stages_for(smem_budget);
    // Short K loops (few input channels, 1x1 / 1x3x3 kernels): when the ring can be cut to a multiple of the stages one tile
    // takes, slot s sees the same (kt, k-chunk) on every tile -- with one weight set for the whole launch (no groups, one
    // m-tile) the weight images stay where the first pass put them and only the activation tiles stream (the weight images
    // of a 32-channel 3x3 layer are 72 KB per k-step against 21 KB of activations: the re-fetch per 256-pixel tile was the
    // L2 -> SM traffic of these layers).
    {
        const int period = kt * ((g.kc + p.ks - 1) / p.ks);
        p.a_resident = 0;
        if (groups == 1 && g.mt == 1 && period <= p.stages && env_flag("LVG_CONV_RESIDENT_W", 1)) {
            p.stages = p.stages / period * period;
            p.a_resident = 1;
        }
    }
    p.y_cs = (int64_t)p.to * p.hos * p.wos;
    p.total_tiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_t * g.mt * inst;

    if (plan_only) { *g_plan_only = p; return LVG_OK; }

    // re-tile the operands
    {
        const int rc = x8_pre ? LVG_OK : pack_act(x, x8, split, inst, ck, g.cblk, t, xin_h, xin_w, h, wd, dil, s);
        if (rc) return rc;
        const int64_t wblocks = (int64_t)groups * g.mt * g.kc;
        LVG_REQUIRE(wblocks < (1ll << 31), "convnd: too many weight tiles");
        // rows of m per pass: <= ~64 KB of staging (16 * rows * taps elements either way)
        const int es = split ? 4 : 2;
        int rpp = (int)((64 * 1024) / (16 * taps * es + 16)) / 8 * 8;
        if (rpp > kBM) rpp = kBM;
        if (rpp < 8) rpp = 8;
        const bool mrows = w_sk < w_sm;
        const int run_el = mrows ? 16 * taps : rpp * taps;
        const int pitch = ((run_el * es + 2 + 3) / 4) | 1;
        const int max_runs = mrows ? rpp : 16;
        const size_t wsm = ((size_t)max_runs * pitch + max_runs + 4) * 4;
        // row chunks per 128-row image: enough CTAs for two per SM (1, 2, 4 or 8 chunks of 128 / chunks rows)
        int chunks = 1;
        {
            const char* e = getenv("LVG_PACKW_CHUNKS");       // experiments
            if (e && atoi(e) >= 1) chunks = atoi(e) >= 8 ? 8 : (atoi(e) >= 4 ? 4 : (atoi(e) >= 2 ? 2 : 1));
            else while (chunks < 8 && wblocks * chunks < 2 * (int64_t)num_sms()) chunks *= 2;
        }
        const int rows_per_cta = kBM / chunks;
        const dim3 wgrid((unsigned)wblocks, (unsigned)chunks);
        if (split) {
            LVG_CUDA(cudaFuncSetAttribute(conv_pack_w_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
            conv_pack_w_kernel<float, true><<<wgrid, 256, wsm, s>>>((const float*)w, wp, cm, ck, g.cpad, taps, w_gs, w_sm, w_sk, flip, g.mt, g.kc, rpp, rows_per_cta);
        } else {
            LVG_CUDA(cudaFuncSetAttribute(conv_pack_w_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
            conv_pack_w_kernel<__half, false><<<wgrid, 256, wsm, s>>>((const __half*)w, wp, cm, ck, g.cpad, taps, w_gs, w_sm, w_sk, flip, g.mt, g.kc, rpp, rows_per_cta);
        }
        LVG_LAUNCH_CHECK();
    }

    // tensor map over X8 as 8-byte elements: (2 W, H, T, instance * block)
    CUtensorMap tm;
    {
        const int rc = encode_map(&tm, x8, wd, h, t, inst * g.nblk, wd, (int64_t)h * wd, thw, p.wtb, p.thb, p.tt, 2);
        if (rc) return rc;
    }
    const size_t smem = (size_t)p.stages * p.stage_bytes + epi_bytes + 128;
    LVG_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* cta_env = getenv("LVG_CONV_CTAS");          // experiments: fewer persistent CTAs than SMs
    const int max_ctas = cta_env ? atoi(cta_env) : num_sms();
    int64_t ctas = p.total_tiles < max_ctas ? p.total_tiles : max_ctas;
    conv_igemm_kernel<<<(unsigned)ctas, kIgemmThreads, smem, s>>>(tm, p);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

bool nd_supported(int dtype, int kt, int kh, int kw)
{
    return (dtype == LVG_F16 || dtype == LVG_F32) && kt >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 9 && kt <= 7;
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int64_t lvg_convnd_workspace(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw,
                                        int pad_t, int pad_h, int pad_w)
{
    if (!nd_supported(dtype, kt, kh, kw) || n < 1 || groups < 1) return -1;
    const int split = dtype == LVG_F32;
    const int taps = kt * kh * kw;
    const int64_t inst = (int64_t)n * groups;
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    if (to < 1 || ho < 1 || wo < 1) return -1;
    const Geometry a = geometry(split, inst, groups, cin, cout, (int64_t)t * h * wd, taps);        // fprop
    const Geometry b = geometry(split, inst, groups, cout, cin, (int64_t)to * ho * wo, taps);      // dgrad
    const int64_t fa = a.act_bytes + a.w_bytes, fb = b.act_bytes + b.w_bytes;
    return (fa > fb ? fa : fb) + 1024;
}

extern "C" int lvg_convnd_fprop(const void* x, const void* w, void* y, int dtype, int n, int groups, int cin, int cout, int t, int h, int wd,
                                int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride, const float* bias, int act, float alpha,
                                float gain, float clamp, void* workspace, int64_t workspace_bytes, void* stream)
{
    LVG_REQUIRE(x && w && y, "convnd_fprop: x, w, y must not be NULL");
    if (!bias && act == 0 && gain == 1.f && clamp < 0.f &&
        pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, stride, (int64_t)t * h * wd, 0))
        return pw_conv((const float*)x, (const float*)w, (float*)y, n, cin, cout, (int64_t)t * h * wd, cin, 1, (cudaStream_t)stream);
    if (!nd_supported(dtype, kt, kh, kw) || n < 1 || pad_t < 0 || pad_h < 0 || pad_w < 0 || stride < 1 || stride > 4) {
        set_error("convnd_fprop: outside the tensor-core kernel's envelope");
        return LVG_UNSUPPORTED;
    }
    const int taps = kt * kh * kw;
    return run_igemm(x, w, y, dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, (int64_t)cout * cin * taps,
                     (int64_t)cin * taps, taps, 0, bias, act, alpha, gain, clamp, h, wd, 1, stride, nullptr, workspace, workspace_bytes,
                     (cudaStream_t)stream);
}

extern "C" int lvg_convnd_dgrad(const void* dy, const void* w, void* dx, int dtype, int n, int groups, int cin, int cout, int t, int h, int wd,
                                int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride, void* workspace, int64_t workspace_bytes,
                                void* stream)
{
    LVG_REQUIRE(dy && w && dx, "convnd_dgrad: dy, w, dx must not be NULL");
    if (pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, stride, (int64_t)t * h * wd, 0))   // dx = W^T dy
        return pw_conv((const float*)dy, (const float*)w, (float*)dx, n, cout, cin, (int64_t)t * h * wd, 1, cin, (cudaStream_t)stream);
    if (!nd_supported(dtype, kt, kh, kw) || n < 1 || pad_t < 0 || pad_h < 0 || pad_w < 0 || pad_t > kt - 1 || pad_h > kh - 1 || pad_w > kw - 1 ||
        stride < 1 || stride > 4) {
        set_error("convnd_dgrad: outside the tensor-core kernel's envelope");
        return LVG_UNSUPPORTED;
    }
    // dx = correlation of dy (to x ho x wo, cout channels; for a strided convolution: dy spread over every stride-th pixel of
    // that grid) with the channel-transposed, mirrored weights, padding k-1-pad
    const int taps = kt * kh * kw;
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    const int hos = (ho - 1) / stride + 1, wos = (wo - 1) / stride + 1;
    return run_igemm(dy, w, dx, dtype, n, groups, cout, cin, to, ho, wo, kt, kh, kw, kt - 1 - pad_t, kh - 1 - pad_h, kw - 1 - pad_w,
                     (int64_t)cout * cin * taps, taps, (int64_t)cin * taps, 1, nullptr, 0, 0.f, 1.f, -1.f, hos, wos, stride, 1, nullptr, workspace,
                     workspace_bytes, (cudaStream_t)stream);
}

// =================================================================================================
// Weight gradient:  dW[g][co][ci][kt][ky][kx] = sum over samples and output pixels of dy[co][pix] * x[ci][pix + tap]
//
// GEMM view per CTA (group g, 128-channel tile of co, NT-channel tile of ci, tap row (kt, ky)):
//   D_kx[co][ci] += A[co][k] * B_kx[ci][k],   K = output pixels, the kw taps of the row as kw accumulators in TMEM.
// Both operands come from the SAME channel-block-of-8 tensors the forward kernels use, now as MN-major operands: a pixel
// is a 16-byte row (8 channels) and the pixel index runs linearly at 16 bytes over a stage tile of RH rows x PS columns,
// PS = (segment width + kw - 1) rounded up to 16. The dy tile is loaded through a tensor map whose W extent is clipped to
// the column segment, so TMA zero-fills its last PS - width columns -- those are the K positions where the x tile holds
// the row's halo; with equal pitches on both sides a tap (kx) is a start-address shift of kx * 16 bytes and the tap row
// (kt, ky) a shifted TMA box. One MMA covers 16 consecutive pixels of the linear index.
// Few groups (discriminators, low-res networks): the stages are cut into `nsplit` ranges over the grid, fp32 partial
// sums go to a workspace and conv_wgrad_reduce_kernel folds them. fp32 tensors: hi/lo bf16 halves of BOTH operands are
// staged, three MMAs per k-step (hi*hi + lo*hi + hi*lo).

namespace lvg {
namespace {

struct WgradV2Params {
    void* dw;                    // final output when nsplit == 1, else fp32 partials [nsplit][...]
    int out_f32;                 // element type of `dw` (partials are always fp32)
    int bf16, split;             // operand format; split = hi/lo pairs staged
    int n, groups, cin, cout;
    int to, ho;                  // output rows / frames iterated
    int kt, kh, kw, pad_t, pad_h, pad_w;
    int nt;                      // ci per n-tile (multiple of 16)
    int nblk_a, nblk_b;          // channel blocks per instance in dy8 / x8 (incl. the lo half in split mode)
    int lo_a, lo_b;              // block offset of the lo half
    int nseg, seg_w[4], seg_x0[4], ps[4];
    int rh;                      // dy rows per stage
    int khc;                     // tap rows (ky) per CTA: 1, or kh (folded: one dy tile and one x tile of rh + kh - 1 rows serve all ky)
    int ablk;                    // channel blocks of one dy image in a stage (16, or fewer for cout < 128: the rest of the 128 MMA rows is never stored)
    int nsplit;
    int a_bytes, b_bytes, stage_bytes, stages;      // per stage: one A (B) operand image; a stage holds split+1 of each
    int64_t split_stride;        // elements between fp32 partials
    int tap_major;               // MMA order inside a stage: tap outermost, K steps innermost (else K step outermost, taps innermost)
    int mrows;                   // M of the MMA: 128, or 64 (cout <= 64): accumulator row 16 j + i then sits in TMEM lane 32 j + i
    int tail_bytes;              // shared memory behind the stage ring that the MMAs may read (kx-shifted last rows; the 128 - 8 * ablk rows without data)
};

struct WgradMaps { CUtensorMap a[4]; CUtensorMap b; };     // dy8 clipped to each column segment; x8

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_v2_kernel(const __grid_constant__ WgradMaps maps, const WgradV2Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], acc_bar;
    __shared__ uint32_t tmem_slot;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    int bx = blockIdx.x;
    const int sp = bx % p.nsplit; bx /= p.nsplit;
    int ky0 = 0;
    if (p.khc == 1) { ky0 = bx % p.kh; bx /= p.kh; }
    const int kt = bx % p.kt;
    const int nti = bx / p.kt;
    const int mti = blockIdx.y, g = blockIdx.z;
    const int NT = p.nt;
    const int nop = p.split ? 2 : 1;                              // operand images per stage and side

    // stages of this CTA: (sample, frame, segment, row block), range [s0, s1)
    const int rblocks = (p.ho + p.rh - 1) / p.rh;
    const int per_plane = p.nseg * rblocks;
    const int total = p.n * p.to * per_plane;
    const int s0 = (int)((int64_t)total * sp / p.nsplit), s1 = (int)((int64_t)total * (sp + 1) / p.nsplit);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        fence_barrier_init();
    }
    // zero the slack behind the stage buffers (the kx-shifted reads of the last block run 32 bytes past its end; the dy
    // values they meet are zero, and zero * garbage must not become NaN)
    // (uninitialised shared memory may hold NaN patterns: clear all of it once)
    for (int i = threadIdx.x; i < (p.stages * p.stage_bytes + p.tail_bytes) / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int s = s0; s < s1; s++) {
                const int it = s - s0, slot = it % p.stages;
                if (it >= p.stages) mbar_wait(&empty_bar[slot], (uint32_t)((it / p.stages - 1) & 1));
                int r = s;
                const int rb = r % rblocks; r /= rblocks;
                const int seg = r % p.nseg; r /= p.nseg;
                const int t = r % p.to;
                const int n = r / p.to;
                const int inst = n * p.groups + g;
                unsigned char* st = smem + (size_t)slot * p.stage_bytes;
                mbar_expect_tx(&full_bar[slot], (uint32_t)(nop * (p.a_bytes + p.b_bytes)));
                const int oy0 = rb * p.rh;
                for (int o = 0; o < nop; o++) {
                    tma_load_4d(st + (size_t)o * p.a_bytes, &maps.a[seg], 0, oy0, t, inst * p.nblk_a + o * p.lo_a + mti * 16, &full_bar[slot]);
                    tma_load_4d(st + (size_t)nop * p.a_bytes + (size_t)o * p.b_bytes, &maps.b, 2 * (p.seg_x0[seg] - p.pad_w), oy0 + ky0 - p.pad_h,
                                t + kt - p.pad_t, inst * p.nblk_b + o * p.lo_b + nti * (NT / 8), &full_bar[slot]);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            // instruction descriptor: D = f32, A and B MN-major (bits 15, 16), N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | (p.bf16 ? ((1u << 7) | (1u << 10)) : 0u) | (1u << 15) | (1u << 16) | ((uint32_t)(NT >> 3) << 17) |
                                   ((uint32_t)(p.mrows >> 4) << 24);
            bool first = true;
            for (int s = s0; s < s1; s++) {
                const int it = s - s0, slot = it % p.stages;
                const int seg = (s / rblocks) % p.nseg;
                const int rb = s % rblocks;
                const int rows = min(p.rh, p.ho - rb * p.rh);
                const int ksteps = (rows * p.ps[seg] + 15) / 16;                 // (a trailing half step reads the zero-filled next row)
                const uint32_t blk_a = (uint32_t)(p.rh * p.ps[seg] * 16);                    // bytes of one channel block of the dy tile
                const uint32_t blk_b = (uint32_t)((p.rh + p.khc - 1) * p.ps[seg] * 16);      // ... of the x tile (kh - 1 halo rows when folded)
                const uint32_t ky_step = (uint32_t)(p.ps[seg] * 16);                          // one tap row down = one tile row further
                mbar_wait(&full_bar[slot], (uint32_t)((it / p.stages) & 1));
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + (size_t)slot * p.stage_bytes);
                const uint32_t b0 = a0 + (uint32_t)(nop * p.a_bytes);
                // descriptors as (lo, hi) words: a K step is +16 on the low word (256 bytes >> 4), a tap column +1, a tap row
                // + ky_step >> 4 -- the loop below is all this thread does, and its instruction count per MMA bounds the kernel
                const uint32_t a_hi = desc_hi(blk_a), b_hi = desc_hi(blk_b);
                const uint32_t ky16 = ky_step >> 4;
                if (p.tap_major) {
                    // K steps innermost: consecutive MMAs accumulate into the SAME tensor-memory columns (one tap's accumulator)
                    uint32_t tcol = tmem_d;
                    for (int kyi = 0; kyi < p.khc; kyi++) {
                        for (int kx = 0; kx < p.kw; kx++, tcol += (uint32_t)NT) {
                            uint32_t acc = first ? 0u : 1u;
                            for (int term = 0; term < (p.split ? 3 : 1); term++) {
                                uint32_t a_lo = desc_lo(a0 + (term == 1 ? (uint32_t)p.a_bytes : 0u), 128);     // hi*hi, lo*hi, hi*lo
                                uint32_t b_lo = desc_lo(b0 + (term == 2 ? (uint32_t)p.b_bytes : 0u), 128) + (uint32_t)kyi * ky16 + (uint32_t)kx;
                                for (int k = 0; k < ksteps; k++, a_lo += 16, b_lo += 16) { umma_f16_w(tcol, a_lo, a_hi, b_lo, b_hi, idesc, acc); acc = 1u; }
                            }
                        }
                    }
                    first = false;
                } else {
                    for (int term = 0; term < (p.split ? 3 : 1); term++) {
                        uint32_t a_lo = desc_lo(a0 + (term == 1 ? (uint32_t)p.a_bytes : 0u), 128);     // hi*hi, lo*hi, hi*lo
                        uint32_t b_lo = desc_lo(b0 + (term == 2 ? (uint32_t)p.b_bytes : 0u), 128);
                        uint32_t acc = first ? 0u : 1u;
                        for (int k = 0; k < ksteps; k++, a_lo += 16, b_lo += 16) {
                            uint32_t b_row = b_lo, tcol = tmem_d;
                            for (int kyi = 0; kyi < p.khc; kyi++, b_row += ky16) {
                                uint32_t b_tap = b_row;
                                for (int kx = 0; kx < p.kw; kx++, b_tap++, tcol += (uint32_t)NT) umma_f16_w(tcol, a_lo, a_hi, b_tap, b_hi, idesc, acc);
                            }
                            acc = 1u;
                        }
                        first = false;
                    }
                }
                umma_commit(&empty_bar[slot]);
            }
            umma_commit(&acc_bar);
        }
    }
    // ---- epilogue, 32 input channels at a time: TMEM -> shared [co][ci * kw + kx] (fp32) -> global rows of kw-runs
    float* tile = reinterpret_cast<float*>(smem);
    const int row_pitch = 32 * p.kw + 1;
    if (warp >= 2) {
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
    }
    __syncthreads();                       // every stage buffer is free from here on
    {
        const int taps = p.kt * p.kh * p.kw;
        const int ci0 = nti * NT, co0 = mti * kBM;
        const bool any = s1 > s0;
        for (int kc0 = 0; kc0 < p.khc * ((NT + 31) / 32); kc0++) {
            const int kyi = kc0 / ((NT + 31) / 32), c0 = (kc0 - kyi * ((NT + 31) / 32)) * 32;      // tap row of this CTA, first of 32 input channels
            const int tap0 = (kt * p.kh + ky0 + kyi) * p.kw;
            if (warp >= 2) {
                const int q = warp % 4;
                const bool m64 = p.mrows == 64;
                const int r = m64 ? q * 16 + (lane & 15) : q * 32 + lane;        // accumulator row held by this lane (M = 64: lanes 16-31 hold none)
                const bool holds = !m64 || lane < 16;
                for (int kx = 0; kx < p.kw; kx++) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((kyi * p.kw + kx) * NT + c0), acc);
                    if (holds) {
#pragma unroll
                        for (int j = 0; j < 32; j++) tile[r * row_pitch + j * p.kw + kx] = any ? __uint_as_float(acc[j]) : 0.f;
                    }
                }
            }
            __syncthreads();
            const int ci_n = min(32, min(NT - c0, p.cin - ci0 - c0));
            const int per_row = ci_n * p.kw;
            if (per_row > 0) {
                for (int r = warp; r < p.mrows; r += kThreads / 32) {
                    const int co = co0 + r;
                    if (co >= p.cout) break;
                    const int64_t base = (((int64_t)g * p.cout + co) * p.cin + ci0 + c0) * taps + tap0;
                    const float* src = tile + r * row_pitch;
                    if (p.nsplit > 1) {
                        float* dst = reinterpret_cast<float*>(p.dw) + (int64_t)sp * p.split_stride + base;
                        for (int j = lane; j < per_row; j += 32) { const int ci = j / p.kw, kx = j - ci * p.kw; dst[ci * taps + kx] = src[j]; }
                    } else if (p.out_f32) {
                        float* dst = reinterpret_cast<float*>(p.dw) + base;
                        for (int j = lane; j < per_row; j += 32) { const int ci = j / p.kw, kx = j - ci * p.kw; dst[ci * taps + kx] = src[j]; }
                    } else {
                        __half* dst = reinterpret_cast<__half*>(p.dw) + base;
                        for (int j = lane; j < per_row; j += 32) { const int ci = j / p.kw, kx = j - ci * p.kw; dst[ci * taps + kx] = __float2half_rn(src[j]); }
                    }
                }
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_d, 512);
}

template <class TOut>
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ part, TOut* __restrict__ out, int64_t n, int nsplit, int64_t stride)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nsplit; k++) s += part[(int64_t)k * stride + i];
        if constexpr (sizeof(TOut) == 2) out[i] = __float2half_rn(s);
        else out[i] = s;
    }
}

struct WgradPlan {
    int split, cpad_a, cpad_b, nt, ntiles, mt, nsplit;
    int ablk, khc, mrows;
    int nseg, seg_w[4], seg_x0[4], ps, rh, stages;
    int a_stage, b_stage, stage_bytes, tail_bytes;
    size_t smem;
    int64_t a_bytes, b_bytes, part_bytes, dw_elems;
};

// channel padding of the dy8 operand of the weight gradient. cout < 128: only the channel blocks that exist (the same
// tensor the input gradient reads, so one re-tiling pass serves both; the remaining rows of the 128-row MMA read whatever
// follows in shared memory and are never stored -- rows of D depend on the same rows of A only). Otherwise whole m-tiles.
inline int wgrad_cpad_a(int cout) { return (cout < kBM && env_flag("LVG_WGRAD_COMPACT", 1)) ? round_up(cout, 16) : round_up(cout, kBM); }

WgradPlan wgrad_plan(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int to, int ho, int wo, int kt, int kh, int kw)
{
    WgradPlan q;
    q.split = dtype == LVG_F32;
    q.cpad_a = wgrad_cpad_a(cout);
    q.ablk = q.cpad_a < kBM ? q.cpad_a / 8 : 16;
    // M of the MMA. This kernel's MN-major MMAs are paced by their shared-memory operand reads (~64 clocks for 128 dy rows +
    // N / 2 for the x columns, measured); with at most 64 output channels M = 64 halves the dy part. Accumulator rows
    // 16 j .. 16 j + 15 then sit in TMEM lanes 32 j .. 32 j + 15 (probed on B200: profiles/r02_probe_m64_tmem_lanes.txt).
    q.mrows = (q.cpad_a <= 64 && env_flag("LVG_WGRAD_M64", 1)) ? 64 : kBM;
    q.cpad_b = round_up(cin, 16);
    // Few input channels (<= 32): ONE CTA takes all kh tap rows -- the x tile carries kh - 1 halo rows and a tap row is a
    // start-address shift of one tile row, like kx is a shift of one pixel -- so dy and x are fetched once per (kt, n-tile)
    // instead of once per (kt, ky). The kh * kw accumulators of NT columns each must fit the 512 TMEM columns, so NT <= 48:
    // measured (B200, profiles/r02_lres_conv_table.txt), an MN-major MMA of this kernel costs ~(64 + N / 2) clocks whatever
    // the issue rate -- 32 input channels: 2.56 vs 3.07 ms folded (same MMA count, a third of the operand traffic); 64 input
    // channels would need two n-tiles of 32 = twice the MMAs: 6.5 vs 4.2 ms, hence the limit (LVG_WGRAD_FOLD_CIN).
    q.khc = (kh > 1 && cin <= env_flag("LVG_WGRAD_FOLD_CIN", 32) && env_flag("LVG_WGRAD_FOLD", 1)) ? kh : 1;
    int nt_cap = (512 / (q.khc * kw)) / 16 * 16;
    if (nt_cap > 256) nt_cap = 256;
    if (q.split && nt_cap > 128) nt_cap = 128;
    // column segments (a TMA box row is at most 128 pixels incl. the kw - 1 halo) and the common tile pitch
    q.nseg = (wo + 128 - kw) / (129 - kw);
    if (q.nseg < 1) q.nseg = 1;
    const int w0 = (wo + q.nseg - 1) / q.nseg;
    for (int j = 0; j < 4; j++) { q.seg_x0[j] = j * w0; q.seg_w[j] = j < q.nseg ? (wo - j * w0 < w0 ? wo - j * w0 : w0) : 0; }
    q.ps = round_up(w0 + kw - 1, 8);                           // rows x pitch must be a multiple of 16 pixels (one MMA K step)
    const int nop = q.split ? 2 : 1;
    // bytes of a stage with `rows` dy rows and an n-tile of `nt` channels
    auto stage_of = [&](int rows, int nt) { return nop * (q.ablk * rows + (nt / 8) * (rows + q.khc - 1)) * q.ps * 16; };
    {   // a pitch of 8 mod 16 needs row pairs: take it only when two stages of two rows fit with a useful n-tile
        if (q.ps % 16 != 0 && 2 * stage_of(2, 64 < nt_cap ? 64 : nt_cap) > 200 * 1024) q.ps = round_up(q.ps, 16);
    }
    {   // the smallest stage (1 or 2 rows) must leave room for two stages
        const int min_rows = q.ps % 16 != 0 ? 2 : 1;
        while (nt_cap > 16 && 2 * stage_of(min_rows, nt_cap) > 200 * 1024) nt_cap -= 16;
    }
    q.ntiles = (q.cpad_b + nt_cap - 1) / nt_cap;
    q.nt = round_up((q.cpad_b + q.ntiles - 1) / q.ntiles, 16);
    q.cpad_b = q.nt * q.ntiles;
    q.mt = (q.cpad_a + kBM - 1) / kBM;
    const int64_t inst = (int64_t)n * groups;
    q.a_bytes = inst * nop * (q.cpad_a / 8) * (int64_t)to * ho * wo * 16;
    q.b_bytes = inst * nop * (q.cpad_b / 8) * (int64_t)t * h * wd * 16;
    q.dw_elems = (int64_t)groups * cout * cin * kt * kh * kw;
    // rows per stage: about 80 KB of operands per stage
    q.rh = 1;
    while (q.rh < ho && q.rh < 255 - q.khc && stage_of(q.rh + 1, q.nt) <= 80 * 1024) q.rh++;
    if (q.ps % 16 != 0) q.rh = q.rh >= 2 ? q.rh / 2 * 2 : 2;   // even row count (rows past the image are zero-filled)
    // behind the ring: the kx-shifted reads of the last block (32 bytes) and, with fewer dy blocks than the MMA has rows / 8,
    // the rows beyond them (read from the lo image's start: (nop - 1) * a_stage + mrows / 8 blocks). Two stages + tail must fit.
    for (;;) {
        q.a_stage = q.ablk * q.rh * q.ps * 16;
        q.b_stage = (q.nt / 8) * (q.rh + q.khc - 1) * q.ps * 16;
        q.stage_bytes = round_up(nop * (q.a_stage + q.b_stage), 128);
        const int over = (nop - 1) * q.a_stage + (q.mrows / 8) * q.rh * q.ps * 16 - q.stage_bytes;
        q.tail_bytes = round_up(256 + (over > 0 ? over : 0), 128);
        const int step = q.ps % 16 != 0 ? 2 : 1;
        if (2 * q.stage_bytes + q.tail_bytes <= 220 * 1024 || q.rh <= step) break;
        q.rh -= step;
    }
    q.stages = 2;
    while (q.stages < kMaxStages && (q.stages + 1) * q.stage_bytes + q.tail_bytes <= 220 * 1024) q.stages++;
    const size_t tile_bytes = (size_t)kBM * (32 * kw + 1) * 4;
    q.smem = (size_t)q.stages * q.stage_bytes + q.tail_bytes;
    if (tile_bytes > q.smem) q.smem = tile_bytes;
    q.smem += 128;
    // Split the pixel range over `nsplit` CTAs per output tile. One CTA per SM is resident, so the kernel runs in waves of
    // num_sms CTAs: choose the split that minimises waves x (stages per CTA + a fixed per-CTA cost of ~4 stages: clearing
    // shared memory, the TMEM -> global epilogue) -- e.g. 3 output tiles: 49 splits = 147 CTAs = one wave of 470 stages
    // instead of 64 splits = 192 CTAs = two waves of 360. Partial sums are capped at 256 MB.
    const int64_t ctas = (int64_t)q.ntiles * (kh / q.khc) * kt * q.mt * groups;
    const int64_t stages = (int64_t)n * to * q.nseg * ((ho + q.rh - 1) / q.rh);
    const int sms = 148;
    int64_t cap = 160;
    if (cap > stages) cap = stages;
    while (cap > 1 && cap * q.dw_elems * 4 > (256ll << 20)) cap--;
    int64_t ns = 1;
    double best = 1e300;
    for (int64_t k = 1; k <= cap; k++) {
        const int64_t waves = (ctas * k + sms - 1) / sms;
        const double tm = (double)waves * ((double)((stages + k - 1) / k) + 4.0 * q.khc);
        if (tm < best * 0.999) { best = tm; ns = k; }
    }
    {
        const char* e = getenv("LVG_WGRAD_NSPLIT");       // experiments
        if (e && atoi(e) >= 1) ns = atoi(e) < stages ? atoi(e) : stages;
    }
    q.nsplit = (int)ns;
    q.part_bytes = q.nsplit > 1 ? q.nsplit * q.dw_elems * 4 : 0;
    return q;
}

// the weight-gradient launch; `dy8_pre` != nullptr: dy is already re-tiled (the tensor the input gradient of the same call read)
int run_wgrad(const void* x, const void* dy, void* dw, int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh,
              int kw, int pad_t, int pad_h, int pad_w, int stride, const unsigned char* dy8_pre, void* workspace, int64_t workspace_bytes,
              cudaStream_t s)
{
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    const WgradPlan q = wgrad_plan(dtype, n, groups, cin, cout, t, h, wd, to, ho, wo, kt, kh, kw);
    const int64_t a_room = dy8_pre ? 0 : ((q.a_bytes + 255) / 256) * 256;
    LVG_REQUIRE(workspace && workspace_bytes >= a_room + q.b_bytes + q.part_bytes + 768, "convnd_wgrad: workspace too small");
    LVG_REQUIRE(aligned16(workspace), "convnd_wgrad: workspace must be 16-byte aligned");
    LVG_REQUIRE(groups <= 65535 && q.mt <= 65535, "convnd_wgrad: too many groups / channel tiles");
    const int64_t inst = (int64_t)n * groups;
    unsigned char* x8 = reinterpret_cast<unsigned char*>(workspace) + a_room;
    unsigned char* dy8 = dy8_pre ? const_cast<unsigned char*>(dy8_pre) : reinterpret_cast<unsigned char*>(workspace);
    float* part = reinterpret_cast<float*>(x8 + ((q.b_bytes + 255) / 256) * 256);
    const int64_t thw_a = (int64_t)to * ho * wo, thw_b = (int64_t)t * h * wd;
    {
        // dy of a strided convolution is spread over every stride-th pixel of the stride-1 output grid (zeros between)
        int rc = dy8_pre ? LVG_OK : pack_act(dy, dy8, q.split, inst, cout, q.cpad_a / 8, to, (ho - 1) / stride + 1, (wo - 1) / stride + 1, ho, wo, stride, s);
        if (rc) return rc;
        rc = pack_act(x, x8, q.split, inst, cin, q.cpad_b / 8, t, h, wd, h, wd, 1, s);
        if (rc) return rc;
    }
    WgradV2Params p;
    memset(&p, 0, sizeof(p));
    p.out_f32 = q.split; p.bf16 = q.split; p.split = q.split;
    p.n = n; p.groups = groups; p.cin = cin; p.cout = cout; p.to = to; p.ho = ho;
    p.kt = kt; p.kh = kh; p.kw = kw; p.pad_t = pad_t; p.pad_h = pad_h; p.pad_w = pad_w;
    p.nt = q.nt;
    p.nblk_a = (q.split ? 2 : 1) * (q.cpad_a / 8); p.nblk_b = (q.split ? 2 : 1) * (q.cpad_b / 8);
    p.lo_a = q.cpad_a / 8; p.lo_b = q.cpad_b / 8;
    p.nseg = q.nseg;
    for (int j = 0; j < 4; j++) { p.seg_w[j] = q.seg_w[j]; p.seg_x0[j] = q.seg_x0[j]; p.ps[j] = q.ps; }
    p.rh = q.rh; p.khc = q.khc; p.ablk = q.ablk;
    p.mrows = q.mrows;
    // MMA order inside a stage. Switching the accumulator (another tap's TMEM columns) between two MMAs costs ~60 clocks on
    // B200 whatever M and N are (measured: the 32 -> 32 layer of the low-res discriminator 2.63 -> 1.66 ms with the K steps
    // innermost); with a single K step per stage there is nothing to keep together and the K-outermost order pipelines better
    // (512 channels at 3x4 pixels: 0.36 vs 0.46 ms).
    {
        const int e = env_flag("LVG_WGRAD_TAP_MAJOR", -1);
        p.tap_major = e >= 0 ? e : (q.rh * q.ps / 16 >= 2 ? 1 : 0);
    }
    p.a_bytes = q.a_stage; p.b_bytes = q.b_stage; p.stage_bytes = q.stage_bytes; p.stages = q.stages; p.tail_bytes = q.tail_bytes;
    const size_t smem = q.smem;
    LVG_REQUIRE(smem <= 227 * 1024, "convnd_wgrad: stage does not fit shared memory (%zu bytes)", smem);
    LVG_REQUIRE(q.khc * kw * q.nt <= 512, "convnd_wgrad: accumulators exceed tensor memory");
    p.nsplit = q.nsplit;
    p.split_stride = q.dw_elems;
    p.dw = q.nsplit > 1 ? (void*)part : dw;
    // NOTE the per-segment tile pitch: a stage tile of segment j is [block][rows][ps[j]][16 B] -- the box width IS the pitch
    WgradMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int j = 0; j < p.nseg; j++) {
        const int rc = encode_map(&maps.a[j], dy8 + (size_t)p.seg_x0[j] * 16, p.seg_w[j], ho, to, inst * p.nblk_a, wo, (int64_t)ho * wo, thw_a,
                                  p.ps[j], p.rh, 1, q.ablk);
        if (rc) return rc;
    }
    {
        const int rc = encode_map(&maps.b, x8, wd, h, t, inst * p.nblk_b, wd, (int64_t)h * wd, thw_b, p.ps[0], p.rh + q.khc - 1, 1, q.nt / 8);
        if (rc) return rc;
    }
    LVG_CUDA(cudaFuncSetAttribute(conv_wgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(q.ntiles * kt * (kh / q.khc) * q.nsplit), (unsigned)q.mt, (unsigned)groups);
    conv_wgrad_v2_kernel<<<grid, kThreads, smem, s>>>(maps, p);
    LVG_LAUNCH_CHECK();
    if (q.nsplit > 1) {
        int64_t blocks = (q.dw_elems + 255) / 256;
        const int64_t cap = (int64_t)num_sms() * 32;
        if (blocks > cap) blocks = cap;
        if (q.split) conv_wgrad_reduce_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(part, (float*)dw, q.dw_elems, q.nsplit, q.dw_elems);
        else conv_wgrad_reduce_kernel<__half><<<(unsigned)blocks, 256, 0, s>>>(part, (__half*)dw, q.dw_elems, q.nsplit, q.dw_elems);
        LVG_LAUNCH_CHECK();
    }
    return LVG_OK;
}

bool wgrad_in_envelope(int dtype, int n, int groups, int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride)
{
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    return nd_supported(dtype, kt, kh, kw) && n >= 1 && groups >= 1 && kw <= 3 && pad_t >= 0 && pad_h >= 0 && pad_w >= 0 && to >= 1 && ho >= 1 &&
           wo >= 1 && wo <= 4 * (128 - kw + 1) && stride >= 1 && stride <= 4;
}

}  // namespace
}  // namespace lvg

extern "C" int64_t lvg_convnd_wgrad_workspace(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw,
                                              int pad_t, int pad_h, int pad_w)
{
    if (!wgrad_in_envelope(dtype, n, groups, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, 1)) return -1;
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    if (pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, 1, (int64_t)t * h * wd, 1)) return pw_wgrad_workspace(cin, cout);
    const WgradPlan q = wgrad_plan(dtype, n, groups, cin, cout, t, h, wd, to, ho, wo, kt, kh, kw);
    return q.a_bytes + q.b_bytes + q.part_bytes + 1024;
}

extern "C" int lvg_convnd_wgrad(const void* x, const void* dy, void* dw, int dtype, int n, int groups, int cin, int cout, int t, int h, int wd,
                                int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride, void* workspace, int64_t workspace_bytes,
                                void* stream)
{
    LVG_REQUIRE(x && dy && dw, "convnd_wgrad: x, dy, dw must not be NULL");
    if (pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, stride, (int64_t)t * h * wd, 1)) {
        LVG_REQUIRE(workspace, "convnd_wgrad: workspace must not be NULL");
        return pw_wgrad((const float*)x, (const float*)dy, (float*)dw, n, cin, cout, (int64_t)t * h * wd, workspace, workspace_bytes, (cudaStream_t)stream);
    }
    if (!wgrad_in_envelope(dtype, n, groups, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride)) {
        set_error("convnd_wgrad: outside the tensor-core kernel's envelope");
        return LVG_UNSUPPORTED;
    }
    return run_wgrad(x, dy, dw, dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride, nullptr, workspace, workspace_bytes,
                     (cudaStream_t)stream);
}

// the tiling lvg_convnd_fprop (mode 0) / lvg_convnd_dgrad (mode 1) would launch with, as ints (host arithmetic only;
// tests/test_igemm_emul.py replays the forward kernel's addressing with it on the CPU)
extern "C" int lvg_convnd_plan(int mode, int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw, int pad_t,
                               int pad_h, int pad_w, int stride, int* out, int out_len)
{
    LVG_REQUIRE(out && out_len >= 48, "convnd_plan: out must hold 48 ints");
    LVG_REQUIRE(mode == 0 || mode == 1, "convnd_plan: mode 0 (forward) or 1 (input gradient)");
    const int taps = kt * kh * kw;
    if (!nd_supported(dtype, kt, kh, kw) || n < 1 || groups < 1 || pad_t < 0 || pad_h < 0 || pad_w < 0 || stride < 1 || stride > 4 ||
        (mode == 1 && (pad_t > kt - 1 || pad_h > kh - 1 || pad_w > kw - 1))) {
        set_error("convnd_plan: outside the tensor-core kernel's envelope");
        return LVG_UNSUPPORTED;
    }
    for (int i = 0; i < out_len; i++) out[i] = 0;
    if (pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, stride, (int64_t)t * h * wd, 0)) { out[47] = 1; return LVG_OK; }
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    void* dummy = reinterpret_cast<void*>(256);          // never dereferenced in plan-only mode
    g_plan_only = &p;
    int rc;
    if (mode == 0) {
        rc = run_igemm(dummy, dummy, dummy, dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, (int64_t)cout * cin * taps,
                       (int64_t)cin * taps, taps, 0, nullptr, 0, 0.f, 1.f, -1.f, h, wd, 1, stride, nullptr, nullptr, 0, nullptr);
    } else {
        const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
        if (to < 1 || ho < 1 || wo < 1) { g_plan_only = nullptr; set_error("convnd_plan: empty output"); return LVG_UNSUPPORTED; }
        const int hos = (ho - 1) / stride + 1, wos = (wo - 1) / stride + 1;
        rc = run_igemm(dummy, dummy, dummy, dtype, n, groups, cout, cin, to, ho, wo, kt, kh, kw, kt - 1 - pad_t, kh - 1 - pad_h, kw - 1 - pad_w,
                       (int64_t)cout * cin * taps, taps, (int64_t)cin * taps, 1, nullptr, 0, 0.f, 1.f, -1.f, hos, wos, stride, 1, nullptr, nullptr, 0,
                       nullptr);
    }
    g_plan_only = nullptr;
    if (rc) return rc;
    const int v[48] = {p.wgroups, p.cout, p.mt, p.kc, p.nblk, p.nimg, p.lo_blk, p.to, p.ho, p.wo, p.kt, p.kh, p.kw, p.pad_t, p.pad_h, p.pad_w,
                       p.tt, p.th, p.wt, p.wtb, p.thb, p.frame_px, p.ncols, p.n0, p.epi_warps, p.nbuf, p.tiles_x, p.tiles_y, p.tiles_t,
                       (int)p.total_tiles, p.ks, p.stages, p.a_resident, p.a_stage, p.b_step, p.b_bytes, p.b_box, p.stage_bytes, p.ostride, p.hos, p.wos,
                       0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 48; i++) out[i] = v[i];
    return LVG_OK;
}

// the tiling lvg_convnd_wgrad would launch with (host arithmetic only; tests/test_wgrad_emul.py replays it on the CPU, tools print it)
extern "C" int lvg_convnd_wgrad_plan(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw, int pad_t,
                                     int pad_h, int pad_w, int* out, int out_len)
{
    LVG_REQUIRE(out && out_len >= 32, "convnd_wgrad_plan: out must hold 32 ints");
    if (!wgrad_in_envelope(dtype, n, groups, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, 1)) {
        set_error("convnd_wgrad_plan: outside the tensor-core kernel's envelope");
        return LVG_UNSUPPORTED;
    }
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    const WgradPlan q = wgrad_plan(dtype, n, groups, cin, cout, t, h, wd, to, ho, wo, kt, kh, kw);
    const int v[32] = {q.split, q.cpad_a, q.cpad_b, q.nt, q.ntiles, q.mt, q.nsplit, q.ablk, q.khc, q.nseg, q.ps, q.rh, q.stages, q.a_stage, q.b_stage,
                       q.stage_bytes, q.tail_bytes, (int)q.smem, q.seg_w[0], q.seg_w[1], q.seg_w[2], q.seg_w[3], q.seg_x0[0], q.seg_x0[1], q.seg_x0[2],
                       q.seg_x0[3], pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, 1, (int64_t)t * h * wd, 1) ? 1 : 0, q.mrows, 0, 0, 0, 0};
    for (int i = 0; i < 32; i++) out[i] = v[i];
    return LVG_OK;
}

// ---- both gradients of one convolution call. dy is re-tiled ONCE: the input gradient (forward kernel on dy) and the weight
// gradient (dy as the M-side operand) read the same channel-block tensor whenever their paddings agree (cout < 128 or a
// multiple of 128: every layer of the networks); otherwise, and for the streaming 1x1x1 kernels, the two entry points above run
// one after the other. Workspace layout of the shared case: [dy8][packed weights of the input gradient][x8][partial sums].
namespace lvg {
namespace {
bool backward_shares_dy8(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h,
                         int pad_w, int stride)
{
    const int64_t P = (int64_t)t * h * wd;
    if (!env_flag("LVG_CONV_SHARED_DY8", 1)) return false;
    if (pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, stride, P, 0) ||
        pw_supported(dtype, groups, cin, cout, kt, kh, kw, pad_t, pad_h, pad_w, stride, P, 1))
        return false;
    return wgrad_cpad_a(cout) == round_up(cout, 16);
}
}  // namespace
}  // namespace lvg

extern "C" int64_t lvg_convnd_backward_workspace(int dtype, int n, int groups, int cin, int cout, int t, int h, int wd, int kt, int kh, int kw,
                                                 int pad_t, int pad_h, int pad_w)
{
    const int64_t a = lvg_convnd_workspace(dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w);
    const int64_t b = lvg_convnd_wgrad_workspace(dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w);
    if (a < 0 || b < 0) return -1;
    // (the shared layout never needs more than the two separate ones together)
    return a + b + 1024;
}

extern "C" int lvg_convnd_backward(const void* x, const void* dy, const void* w, void* dx, void* dw, int dtype, int n, int groups, int cin, int cout,
                                   int t, int h, int wd, int kt, int kh, int kw, int pad_t, int pad_h, int pad_w, int stride, void* workspace,
                                   int64_t workspace_bytes, void* stream)
{
    LVG_REQUIRE(x && dy && w && dx && dw, "convnd_backward: x, dy, w, dx, dw must not be NULL");
    LVG_REQUIRE(workspace && aligned16(workspace), "convnd_backward: workspace must be 16-byte aligned and not NULL");
    const bool tc_ok = wgrad_in_envelope(dtype, n, groups, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride) && pad_t <= kt - 1 && pad_h <= kh - 1 &&
                       pad_w <= kw - 1;
    if (!tc_ok || !backward_shares_dy8(dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride)) {
        const int rc = lvg_convnd_dgrad(dy, w, dx, dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride, workspace,
                                        workspace_bytes, stream);
        if (rc) return rc;
        return lvg_convnd_wgrad(x, dy, dw, dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride, workspace, workspace_bytes,
                                stream);
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int split = dtype == LVG_F32 ? 1 : 0;
    const int taps = kt * kh * kw;
    const int to = t + 2 * pad_t - kt + 1, ho = h + 2 * pad_h - kh + 1, wo = wd + 2 * pad_w - kw + 1;
    const int hos = (ho - 1) / stride + 1, wos = (wo - 1) / stride + 1;
    const int64_t inst = (int64_t)n * groups;
    const Geometry gd = geometry(split, inst, groups, cout, cin, (int64_t)to * ho * wo, taps);      // the input gradient's operands
    const int64_t dy8_room = ((gd.act_bytes + 255) / 256) * 256;
    const int64_t wp_room = ((gd.w_bytes + 255) / 256) * 256 + 256;
    LVG_REQUIRE(workspace_bytes >= dy8_room + wp_room, "convnd_backward: workspace too small");
    unsigned char* dy8 = reinterpret_cast<unsigned char*>(workspace);
    unsigned char* rest = dy8 + dy8_room;
    int rc = pack_act(dy, dy8, split, inst, cout, gd.cblk, to, hos, wos, ho, wo, stride, s);
    if (rc) return rc;
    rc = run_igemm(dy, w, dx, dtype, n, groups, cout, cin, to, ho, wo, kt, kh, kw, kt - 1 - pad_t, kh - 1 - pad_h, kw - 1 - pad_w,
                   (int64_t)cout * cin * taps, taps, (int64_t)cin * taps, 1, nullptr, 0, 0.f, 1.f, -1.f, hos, wos, stride, 1, dy8, rest, wp_room, s);
    if (rc) return rc;
    // (stream order: the weight gradient's re-tiling of x overwrites the packed weights only after the input gradient read them)
    return run_wgrad(x, dy, dw, dtype, n, groups, cin, cout, t, h, wd, kt, kh, kw, pad_t, pad_h, pad_w, stride, dy8, rest,
                     workspace_bytes - dy8_room, s);
}
