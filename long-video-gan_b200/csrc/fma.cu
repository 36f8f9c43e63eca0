// fma: out = a * b + c with broadcasting (torch_utils/ops/fma.py:15-25; the
// reference routes to torch.addcmul). Operands are described on the common
// broadcast shape; stride 0 marks a broadcast axis. Dense same-shape operands
// take the 128-bit vector path.

#include "common.cuh"

namespace lvg {
namespace {

struct FmaParams {
    const void *a, *b, *c;
    void* out;
    int rank;
    int64_t shape[6], as[6], bs[6], cs[6];
    int64_t n;
};

template <class T>
__global__ void __launch_bounds__(256) fma_dense_kernel(const T* __restrict__ a, const T* __restrict__ b,
                                                        const T* __restrict__ c, T* __restrict__ out, int64_t n_pack)
{
    typedef typename Acc<T>::type S;
    constexpr int N = VecOf<T>::N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pack; i += (int64_t)gridDim.x * blockDim.x) {
        Pack<T> va = load_pack(a + i * N), vb = load_pack(b + i * N), vc = load_pack(c + i * N), vo;
#pragma unroll
        for (int k = 0; k < N; k++) vo.v[k] = from_acc<T>(to_acc(va.v[k]) * to_acc(vb.v[k]) + to_acc(vc.v[k]));
        store_pack(out + i * N, vo);
    }
}

template <class T>
__global__ void __launch_bounds__(256) fma_strided_kernel(FmaParams p, int64_t first)
{
    typedef typename Acc<T>::type S;
    for (int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i, ao = 0, bo = 0, co = 0;
        for (int d = p.rank - 1; d >= 0; d--) {
            const int64_t k = r % p.shape[d];
            r /= p.shape[d];
            ao += k * p.as[d]; bo += k * p.bs[d]; co += k * p.cs[d];
        }
        S v = to_acc(((const T*)p.a)[ao]) * to_acc(((const T*)p.b)[bo]) + to_acc(((const T*)p.c)[co]);
        ((T*)p.out)[i] = from_acc<T>(v);
    }
}

template <class T>
int launch_fma(const FmaParams& p, bool dense, cudaStream_t s)
{
    constexpr int N = VecOf<T>::N;
    int64_t first = 0;
    const int64_t cap = (int64_t)num_sms() * 8 * 8;
    if (dense && aligned16(p.a) && aligned16(p.b) && aligned16(p.c) && aligned16(p.out) && p.n >= N) {
        const int64_t n_pack = p.n / N;
        int64_t blocks = (n_pack + 255) / 256;
        if (blocks > cap) blocks = cap;
        fma_dense_kernel<T><<<(unsigned)blocks, 256, 0, s>>>((const T*)p.a, (const T*)p.b, (const T*)p.c, (T*)p.out, n_pack);
        LVG_LAUNCH_CHECK();
        first = n_pack * N;
    }
    if (first < p.n) {
        int64_t blocks = (p.n - first + 255) / 256;
        if (blocks > cap) blocks = cap;
        fma_strided_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(p, first);
        LVG_LAUNCH_CHECK();
    }
    return LVG_OK;
}

}  // namespace
}  // namespace lvg

using namespace lvg;

extern "C" int lvg_fma(const void* a, const void* b, const void* c, void* out, int dtype,
                       int rank, const int64_t shape[6], const int64_t a_stride[6],
                       const int64_t b_stride[6], const int64_t c_stride[6], void* stream)
{
    LVG_REQUIRE(a && b && c && out, "fma: operands must not be NULL");
    LVG_REQUIRE(rank >= 0 && rank <= 6, "fma: rank must be in [0, 6] (got %d)", rank);
    LVG_REQUIRE(dtype == LVG_F32 || dtype == LVG_F16 || dtype == LVG_F64, "fma: unsupported dtype %d", dtype);
    FmaParams p;
    p.a = a; p.b = b; p.c = c; p.out = out; p.rank = rank; p.n = 1;
    bool dense = true;
    int64_t expect = 1;
    for (int d = rank - 1; d >= 0; d--) {
        LVG_REQUIRE(shape[d] >= 0, "fma: negative dimension");
        p.shape[d] = shape[d]; p.as[d] = a_stride[d]; p.bs[d] = b_stride[d]; p.cs[d] = c_stride[d];
        if (shape[d] != 1 && (a_stride[d] != expect || b_stride[d] != expect || c_stride[d] != expect)) dense = false;
        expect *= shape[d];
        p.n *= shape[d];
    }
    for (int d = rank; d < 6; d++) { p.shape[d] = 1; p.as[d] = p.bs[d] = p.cs[d] = 0; }
    if (p.n == 0) return LVG_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == LVG_F32) return launch_fma<float>(p, dense, s);
    if (dtype == LVG_F16) return launch_fma<__half>(p, dense, s);
    return launch_fma<double>(p, dense, s);
}
