// Accuracy probe: CUDA sincosf vs quanonet_b200's sincos_half vs double, on the GPU.
#include <cstdio>
#include <cmath>
#include "../quanonet_b200/csrc/hea_common.cuh"
__global__ void probe(int n, float lo, float hi, double* stats) {
    // stats: [0] max ulp s (cuda) [1] max ulp c (cuda) [2] sum norm err (cuda) [3..5] same for ours, [6] sum |norm err| cuda [7] ours
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = lo + (hi - lo) * ((float)i + 0.5f) / n;
    float h = 0.5f * t;
    double sd, cd; sincos((double)h, &sd, &cd);
    float s1, c1; sincosf(h, &s1, &c1);
    float s2, c2; qon::sincos_half(t, s2, c2);
    auto ulp = [](float v, double ref) { float r = (float)ref; int e; frexpf(r, &e); double u = ldexp(1.0, e - 24); return fabs((double)v - ref) / u; };
    double u;
    u = ulp(s1, sd); atomicMax((unsigned long long*)&stats[0], __double_as_longlong(u));
    u = ulp(c1, cd); atomicMax((unsigned long long*)&stats[1], __double_as_longlong(u));
    double n1 = (double)s1 * s1 + (double)c1 * c1 - 1.0; atomicAdd(&stats[2], n1); atomicAdd(&stats[6], fabs(n1));
    u = ulp(s2, sd); atomicMax((unsigned long long*)&stats[3], __double_as_longlong(u));
    u = ulp(c2, cd); atomicMax((unsigned long long*)&stats[4], __double_as_longlong(u));
    double n2 = (double)s2 * s2 + (double)c2 * c2 - 1.0; atomicAdd(&stats[5], n2); atomicAdd(&stats[7], fabs(n2));
}
int main() {
    double* st; cudaMallocManaged(&st, 8 * sizeof(double));
    float ranges[][2] = {{-3.1415927f, 3.1415927f}, {-20.f, 20.f}, {-2000.f, 2000.f}, {60000.f, 70000.f}};
    for (auto& r : ranges) {
        for (int i = 0; i < 8; ++i) st[i] = 0;
        int n = 1 << 22;
        probe<<<(n + 255) / 256, 256>>>(n, r[0], r[1], st);
        cudaDeviceSynchronize();
        printf("theta in [%g,%g]: cuda sincosf max ulp s %.2f c %.2f mean(c2+s2-1) %.3e mean|.| %.3e | ours max ulp s %.2f c %.2f mean %.3e mean|.| %.3e\n",
               r[0], r[1], st[0], st[1], st[2] / n, st[6] / n, st[3], st[4], st[5] / n, st[7] / n);
    }
    return 0;
}
