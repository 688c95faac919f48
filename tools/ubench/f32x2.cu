// Microbenchmark: issue rate of FADD/FFMA vs the packed FADD2/FFMA2 (f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run: ./f32x2
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE> __global__ void k(float *out, int iters, float s)
{
    float a[8]; u64 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    const u64 s2 = ((u64)__float_as_uint(s) << 32) | __float_as_uint(s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma1(a[i], s, a[(i + 1) & 7]);
            if (MODE == 1) p[i] = fma2(p[i], s2, p[(i + 1) & 7]);
            if (MODE == 2) a[i] = add1(a[i], a[(i + 1) & 7]);
            if (MODE == 3) p[i] = add2(p[i], p[(i + 1) & 7]);
        }
    }
    float r = 0; u64 q = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r += a[i]; q ^= p[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)(q & 0xff);
}
template <int MODE> void run(const char *name, float *d, int warps_per_smsp)
{
    const int iters = 4096, threads = 128 * warps_per_smsp, blocks = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 16, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int dev_clock; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
    const double inst_per_smsp = (double)iters * 8 * warps_per_smsp;
    printf("%-6s warps/SMSP %d: %.3f ms  -> %.3f warp-instr/cycle/SMSP at %.0f MHz nominal\n", name, warps_per_smsp, ms,
           inst_per_smsp / (ms * 1e-3 * dev_clock * 1e3), dev_clock / 1e3);
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 1024 * sizeof(float));
    for (int w : {1, 2, 4}) { run<0>("FFMA", d, w); run<1>("FFMA2", d, w); run<2>("FADD", d, w); run<3>("FADD2", d, w); }
    return 0;
}
