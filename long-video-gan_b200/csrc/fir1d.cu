// Depthwise long FIR along the last axis:  y[n][g][t] = sum_k w[g][k] * x[n][g][t + k]   (cross-correlation, no padding)
//
// This is BlurredNoise.blur of the low-res generator (model/generator_lres.py:378-387): F.conv1d(noise [N*8, 128, L + 4999],
// filters [128, 1, 5000], groups = 128) -- 128 low-pass filters of 3 ... 5000 taps, right-aligned in a 5000-tap buffer, applied
// to (copies of) the same noise rows; the reference hands it to cuDNN's grouped convolution. Temporal stencil with register
// sliding windows: a thread owns 8 consecutive outputs of one (n, g) row; per 4 taps it loads 4 new samples and 4 taps from
// shared memory as two 128-bit words and issues 32 FMAs. Leading zero taps of a filter (most filters are much shorter than
// the buffer) are skipped: fir_first_tap_kernel finds the first non-zero tap per filter.
#include "common.cuh"

namespace lvg {
namespace {

constexpr int kThreads = 128;
constexpr int kOut = 8;                       // outputs per thread
constexpr int kTile = kThreads * kOut;        // outputs per CTA pass
constexpr int kChunk = 512;                   // taps staged per pass

__global__ void __launch_bounds__(256) fir_first_tap_kernel(const float* __restrict__ w, int* __restrict__ first, int k)
{
    __shared__ int best;
    if (threadIdx.x == 0) best = k;
    __syncthreads();
    const float* f = w + (int64_t)blockIdx.x * k;
    int mine = k;
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        if (f[i] != 0.f && i < mine) mine = i;
    atomicMin(&best, mine);
    __syncthreads();
    if (threadIdx.x == 0) first[blockIdx.x] = best & ~3;       // multiple of 4: the tap loop runs in groups of four
}

__global__ void __launch_bounds__(kThreads) fir1d_depthwise_kernel(const float* __restrict__ x, const float* __restrict__ w, const int* __restrict__ first,
                                                                   float* __restrict__ y, int g_total, int lin, int k, int lout)
{
    __shared__ __align__(16) float sw[kChunk];
    __shared__ __align__(16) float sx[kTile + kChunk + 8];
    const int g = blockIdx.y, n = blockIdx.z;
    const int t_tile = blockIdx.x * kTile;
    const float* xr = x + ((int64_t)n * g_total + g) * lin;
    const float* wr = w + (int64_t)g * k;
    float acc[kOut];
#pragma unroll
    for (int j = 0; j < kOut; j++) acc[j] = 0.f;
    const int t0 = threadIdx.x * kOut;
    for (int k0 = first[g]; k0 < k; k0 += kChunk) {
        const int kc = min(kChunk, k - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kChunk; i += kThreads) sw[i] = i < kc ? wr[k0 + i] : 0.f;
        for (int i = threadIdx.x; i < kTile + kChunk + 8; i += kThreads) {
            const int64_t p = (int64_t)t_tile + k0 + i;
            sx[i] = p < lin ? xr[p] : 0.f;
        }
        __syncthreads();
        // window: samples t0 + kk .. t0 + kk + 7 (+4 new ones per group of four taps)
        float win[kOut + 4];
        {
            const float4 a = *reinterpret_cast<const float4*>(sx + t0), b = *reinterpret_cast<const float4*>(sx + t0 + 4);
            win[0] = a.x; win[1] = a.y; win[2] = a.z; win[3] = a.w; win[4] = b.x; win[5] = b.y; win[6] = b.z; win[7] = b.w;
        }
        const int kc4 = (kc + 3) & ~3;
        for (int kk = 0; kk < kc4; kk += 4) {
            const float4 f = *reinterpret_cast<const float4*>(sw + kk);
            const float4 nx = *reinterpret_cast<const float4*>(sx + t0 + kk + kOut);
            win[8] = nx.x; win[9] = nx.y; win[10] = nx.z; win[11] = nx.w;
#pragma unroll
            for (int j = 0; j < kOut; j++) {
                acc[j] = fmaf(f.x, win[j], acc[j]);
                acc[j] = fmaf(f.y, win[j + 1], acc[j]);
                acc[j] = fmaf(f.z, win[j + 2], acc[j]);
                acc[j] = fmaf(f.w, win[j + 3], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < kOut; j++) win[j] = win[j + 4];
        }
    }
    float* yr = y + ((int64_t)n * g_total + g) * lout + t_tile + t0;
#pragma unroll
    for (int j = 0; j < kOut; j++)
        if (t_tile + t0 + j < lout) yr[j] = acc[j];
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int64_t lvg_fir1d_depthwise_workspace(int groups) { return groups < 1 ? -1 : (int64_t)groups * 4 + 16; }

extern "C" int lvg_fir1d_depthwise(const float* x, const float* w, float* y, int n, int groups, int lin, int k, void* workspace,
                                   int64_t workspace_bytes, void* stream)
{
    LVG_REQUIRE(x && w && y, "fir1d_depthwise: x, w, y must not be NULL");
    LVG_REQUIRE(n >= 1 && groups >= 1 && k >= 1 && lin >= k, "fir1d_depthwise: need lin >= k >= 1");
    LVG_REQUIRE(workspace && workspace_bytes >= (int64_t)groups * 4, "fir1d_depthwise: workspace too small");
    LVG_REQUIRE(groups <= 65535 && n <= 65535, "fir1d_depthwise: too many rows for one launch");
    const int lout = lin - k + 1;
    cudaStream_t s = (cudaStream_t)stream;
    int* first = reinterpret_cast<int*>(workspace);
    fir_first_tap_kernel<<<groups, 256, 0, s>>>(w, first, k);
    LVG_LAUNCH_CHECK();
    dim3 grid((unsigned)((lout + kTile - 1) / kTile), (unsigned)groups, (unsigned)n);
    fir1d_depthwise_kernel<<<grid, kThreads, 0, s>>>(x, w, first, y, groups, lin, k, lout);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}
