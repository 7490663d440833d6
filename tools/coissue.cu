// Micro-benchmark: does a packed FP32 instruction (FFMA2, two cycles on the 32-lane FMA pipe of a scheduler) also hold
// the scheduler's DISPATCH port for two cycles, or can another pipe's instruction be dispatched in its shadow?
// The answer decides the instruction-mix cap of extract_kernel: FMA-pipe time only (packed x 2 + scalar FP)
// or dispatch time (packed x 2 + everything else).
//
// Per loop trip a warp executes 8 independent FFMA2 plus K other instructions (integer adds on the ALU pipe, or
// shared-memory loads), K = 0 / 4 / 8 / 16, with 2 or 4 warps per scheduler.  Reported: cycles per trip per scheduler
// divided by the warps sharing it.  Shadow dispatch would keep K <= 8 at 16 cycles; a blocked port gives 16 + K.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coissue coissue.cu && ./coissue
#include <cuda_runtime.h>
#include <cstdio>

template <int K, int KIND>   // KIND 0: IMNMX (ALU pipe; adds would be merged three at a time), 1: LDS, 2: scalar FMUL (same pipe as FFMA2: control)
__global__ void __launch_bounds__(512) k(float* out, int iters, float a, float b, long long* cycles) {
    __shared__ float sh[1024];
    sh[threadIdx.x] = (float)threadIdx.x;
    sh[threadIdx.x + 512] = 1.0f;
    __syncthreads();
    unsigned long long p[8], pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(threadIdx.x * 1e-3f + i), "f"(1.0f * i));
    int x[16];
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x + i; y[i] = 1.0f + i; }
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(sh) + 4u * (threadIdx.x & 31);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
#pragma unroll
            for (int j = 0; j < K / 8; ++j) {
                const int q = (i * (K / 8) + j) & 15;
                if (KIND == 0) asm volatile("min.s32 %0, %0, %1;" : "+r"(x[q]) : "r"(it));
                else if (KIND == 1) asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(y[q]) : "r"(saddr + 128u * q) : "memory");
                else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(y[q]) : "f"(a));
            }
            if (K == 4 && (i & 1)) {
                const int q = i & 15;
                if (KIND == 0) asm volatile("min.s32 %0, %0, %1;" : "+r"(x[q]) : "r"(it));
                else if (KIND == 1) asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(y[q]) : "r"(saddr + 128u * q) : "memory");
                else asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(y[q]) : "f"(a));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo_, hi_; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo_), "=f"(hi_) : "l"(p[i])); s += lo_ + hi_; }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += (float)x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int K, int KIND>
static void run(const char* name, int threads, float* out, long long* cyc) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 1 << 13;
    for (int rep = 0; rep < 2; ++rep) k<K, KIND><<<sms, threads>>>(out, iters, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const int warps_per_sched = threads / 32 / 4;
    printf("%-6s K=%2d  %d warps/scheduler: %6.2f cycles per trip per warp (8 FFMA2 = 16 FMA-pipe cycles; blocked port: %d, shadow dispatch: %d)\n",
           name, K, warps_per_sched, (double)c / iters / warps_per_sched, 16 + K, K > 8 ? 8 + K : 16);
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 512 * 2);
    cudaMalloc(&cyc, sizeof(long long));
    for (int threads : {256, 512}) {
        run<0, 0>("none", threads, out, cyc);
        run<4, 0>("ALU", threads, out, cyc);
        run<8, 0>("ALU", threads, out, cyc);
        run<16, 0>("ALU", threads, out, cyc);
        run<4, 1>("LDS", threads, out, cyc);
        run<8, 1>("LDS", threads, out, cyc);
        run<16, 1>("LDS", threads, out, cyc);
        run<8, 2>("FMUL", threads, out, cyc);
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
