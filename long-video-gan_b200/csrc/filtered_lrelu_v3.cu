// Launch side of the vectorised fused filtered_lrelu kernel (filtered_lrelu_v3.cuh).
#include "common.cuh"
#include "filtered_lrelu_v3.cuh"

namespace lvg {
namespace flv3 {

template <class T, class G, int MODE>
__global__ void __launch_bounds__(kThreads, G::ctas(MODE)) filtered_lrelu_v3_kernel(FlParams p)
{
    extern __shared__ __align__(16) float smem[];
    const Tile t = make_tile<G>(p, blockIdx.x);
    const Smem s = carve<G>(smem);
    const int tid = threadIdx.x;
    stage0<G, MODE>(p, t, s, tid);
    stage1<T, G>(p, t, s, tid);
    __syncthreads();
    stage2<G>(t, s, tid);
    __syncthreads();
    stage3<G, MODE>(p, t, s, tid);
    __syncthreads();
    if (MODE == SIGN_WRITE) stage3_fixup<G>(p, t, tid);
    stage4<G>(t, s, tid);
    __syncthreads();
    stage5<T, G>(p, t, s, tid);
}

template <class T, class G>
static int launch_cfg(FlParams& p, int mode, cudaStream_t s)
{
    fill_launch_constants<G>(p);
    const int64_t tiles = (int64_t)p.tiles_x * p.tiles_y;
    const int64_t blocks = (int64_t)p.n * p.c * tiles;
    // one-multiply block-index decomposition: exact while dividend * divisor < 2^32
    LVG_REQUIRE(blocks <= INT32_MAX && blocks * tiles < (1ll << 32) && (int64_t)p.n * p.c * p.c < (1ll << 32),
                "filtered_lrelu: grid too large");
    const size_t smem = G::smem_bytes(mode);
    void (*k)(FlParams) = nullptr;
    if (mode == SIGN_WRITE)     k = filtered_lrelu_v3_kernel<T, G, SIGN_WRITE>;
    else if (mode == SIGN_READ) k = filtered_lrelu_v3_kernel<T, G, SIGN_READ>;
    else                        k = filtered_lrelu_v3_kernel<T, G, SIGN_NONE>;
    LVG_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)blocks, kThreads, smem, s>>>(p);
    LVG_LAUNCH_CHECK();
    return LVG_OK;
}

// cfg: 1 = up 2 (12 taps) / down 2 (12), 2 = up 4 (24) / down 2 (12), 3 = up 2 (12) / down 4 (24)
template <class T>
int launch(int cfg, FlParams& p, int mode, cudaStream_t s)
{
    switch (cfg) {
        case 1: return launch_cfg<T, Geom<2, 12, 2, 12, 56, 24, 6, 2, 4, 4>>(p, mode, s);
        case 2: return launch_cfg<T, Geom<4, 24, 2, 12, 56, 24, 2, 2, 4, 4>>(p, mode, s);
        case 3: return launch_cfg<T, Geom<2, 12, 4, 24, 31, 16, 8, 6, 4, 2>>(p, mode, s);
        default: break;
    }
    return LVG_UNSUPPORTED;
}

template int launch<float>(int, FlParams&, int, cudaStream_t);
template int launch<__half>(int, FlParams&, int, cudaStream_t);

}  // namespace flv3
}  // namespace lvg
