#include <cstdio>
#include <cuda_runtime.h>
// MUFU.EX2 throughput: W warps per block, 1 block per SM x occupancy; each thread runs ILP independent chains.
template <int ILP>
__global__ void k(float* out, int iters, float seed) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}
int main() {
    float* d; cudaMalloc(&d, 4);
    int dev_clk; cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps : {4, 8, 16, 32}) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int iters = 4000;
        k<16><<<sms, warps * 32>>>(d, 10, -0.5f);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        k<16><<<sms, warps * 32>>>(d, iters, -0.5f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double ops = (double)sms * warps * 32 * 16.0 * iters;
        printf("warps/SM %2d: %.3f ms, %.2f ex2/ns total, %.2f per SM per ns (x1.965GHz -> %.2f /clk/SM)\n", warps, ms, ops / ms * 1e-6,
               ops / ms * 1e-6 / sms, ops / ms * 1e-6 / sms / 1.965);
    }
    printf("clockRate attr %d kHz, SMs %d\n", dev_clk, sms);
    return 0;
}
