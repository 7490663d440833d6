// Micro-benchmark: sustained FP32 FMA rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on this GPU.
// Gives the real denominator for the extraction kernel's FP32 roofline.   nvcc -arch=sm_100a -O3 -o fp32_peak fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    if (MODE >= 2) { a += threadIdx.x * 1e-9f; b += threadIdx.x * 1e-9f; }   // per-thread operands: 3-register forms
    if (MODE == 0 || MODE == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    } else {
        unsigned long long p[8], pa, pb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[4] = {"FFMA  scalar, uniform operands", "FFMA2 f32x2, uniform operands", "FFMA  scalar, 3 registers",
                            "FFMA2 f32x2, 3 register pairs"};
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
            else if (mode == 1) k<1><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
            else if (mode == 2) k<2><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
            else k<3><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flop = 2.0 * 16 * (double)iters * blocks * threads;
            printf("%-32s rep %d: %.3f ms  %.2f TFLOP/s (%d SMs)\n", names[mode], rep, ms, flop / ms * 1e-9, sms);
        }
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
