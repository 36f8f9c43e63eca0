// Micro-benchmark: FFMA vs FFMA2 (f32x2) issue throughput per SM on sm_100a. Build & run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_rate tools/ubench/fma_rate.cu && /tmp/fma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b)
{
    float2 acc[8];
    float2 g[4];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
#pragma unroll
    for (int i = 0; i < 4; i++) g[i] = make_float2(a + i, a + i);
    float2 v = make_float2(b, b * 1.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) {            // scalar FFMA, 3 distinct registers, 2 per pair
                    acc[i].x = fmaf(g[r].x, v.x, acc[i].x);
                    acc[i].y = fmaf(g[r].y, v.y, acc[i].y);
                } else {                    // packed
                    acc[i] = __ffma2_rn(g[r], v, acc[i]);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int bps : {1, 2, 4, 8}) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            auto launch = [&]() { if (mode == 0) k<0><<<148 * bps, 256>>>(out, iters, 1.0001f, 0.5f); else k<1><<<148 * bps, 256>>>(out, iters, 1.0001f, 0.5f); };
            launch();
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = (double)148 * bps * 256 * iters * 4 * 8 * 2;
            printf("%s blocks/SM=%d: %.3f ms  %.1f TFMA/s  (%.1f TFLOP/s)\n", mode == 0 ? "FFMA " : "FFMA2", bps, ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
