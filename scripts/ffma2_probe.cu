// Throughput probe: FFMA vs FFMA2 (packed f32x2, sm_100) incl. broadcast / swapped operands.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float y, float z) {
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = make_float2((threadIdx.x + i) * 1e-3f, (threadIdx.x - i) * 1e-3f);
    float2 yy = make_float2(y, y * 0.5f), zz = make_float2(z, -z);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, y, z); a[i].y = fmaf(a[i].y, y, z); }
                if (MODE == 1) a[i] = __ffma2_rn(a[i], yy, zz);
                if (MODE == 2) a[i] = __ffma2_rn(make_float2(y, y), make_float2(a[i].y, a[i].x), zz);           // bcast A, swapped B
                if (MODE == 3) a[i] = __ffma2_rn(make_float2(y, y), make_float2(-a[(i + 1) & 15].y, a[(i + 1) & 15].x), a[i]); // complex-style
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name) {
    float* d; cudaMalloc(&d, 4);
    int iters = 4000, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, 100, 0.999f, 1e-4f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, 0.999f, 1e-4f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 2 * 16 * 8 * (double)iters * grid * 256;
    printf("%-40s %8.3f ms  %7.2f TFLOP/s\n", name, ms, flops / (ms * 1e-3) / 1e12);
}
int main() {
    run<0>("FFMA scalar (2 per float2)");
    run<1>("FFMA2 plain");
    run<2>("FFMA2 bcast A + swapped B");
    run<3>("FFMA2 complex-style (neg/swap, 3 srcs)");
    return 0;
}
