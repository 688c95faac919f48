// Does a packed (f32x2) cubic reproduce the scalar one bit for bit?  As written naively: NO (measured on B200, CUDA 12.9) --
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite the explicit rounding modifiers, also when the product is
// written as fma(a, b, -0) with a literal zero and also with -Xptxas -fmad=false, so 30 % of random cubics come out one ulp
// off (kernel `k`).  Variant Z (kernel `kz`): every product is fma(a, b, z) with z = -0.0f passed as a KERNEL ARGUMENT.  ptxas
// cannot fold an addend it does not know, nothing is left to contract (SASS: 11 FFMA2 with a broadcast .F32 addend + 11 FADD2
// per pair of outputs), and a * b + (-0) == RN(a * b) for every a, b: 0 mismatches in 2^20 pairs.  That variant is
// interp_cubic2 (af_device.cuh), used by the general-ratio resampler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o cubic2_check cubic2_check.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../audio-flow-rs_b200/csrc/af_device.cuh"
using namespace af;
__device__ __forceinline__ f2 mul2_exact(f2 a, f2 b) { return fma2(a, b, mk2(-0.0f, -0.0f)); }
__device__ __forceinline__ f2 interp_cubic2(f2 x, f2 y0, f2 y1, f2 y2, f2 y3)
{
    const float c13 = 1.0f / 3.0f, c16 = 1.0f / 6.0f;
    const f2 h = mk2(0.5f, 0.5f), s6 = mk2(c16, c16);
    const f2 a1 = sub2(add2(sub2(mul2_exact(mk2(-c13, -c13), y0), mul2_exact(h, y1)), y2), mul2_exact(s6, y3));
    const f2 a2 = sub2(mul2_exact(h, add2(y0, y2)), y1);
    const f2 a3 = add2(mul2_exact(h, sub2(y1, y2)), mul2_exact(s6, sub2(y3, y0)));
    const f2 x2 = mul2_exact(x, x);
    const f2 x3 = mul2_exact(x2, x);
    return add2(add2(add2(y1, mul2_exact(a1, x)), mul2_exact(a2, x2)), mul2_exact(a3, x3));
}
// Variant Z: every product is fma(a, b, z) with z = -0.0f that only the host knows (a kernel argument): ptxas cannot fold
// it into a multiply, so there is nothing to contract with the addition that follows -- and x * y + (-0) == x * y bit for
// bit for every x, y (also for zero products of either sign).
__device__ __forceinline__ f2 interp_cubic2z(f2 x, f2 y0, f2 y1, f2 y2, f2 y3, f2 z)
{
    const float c13 = 1.0f / 3.0f, c16 = 1.0f / 6.0f;
    const f2 h = mk2(0.5f, 0.5f), s6 = mk2(c16, c16);
    const f2 a1 = sub2(add2(sub2(fma2(mk2(-c13, -c13), y0, z), fma2(h, y1, z)), y2), fma2(s6, y3, z));
    const f2 a2 = sub2(fma2(h, add2(y0, y2), z), y1);
    const f2 a3 = add2(fma2(h, sub2(y1, y2), z), fma2(s6, sub2(y3, y0), z));
    const f2 x2 = fma2(x, x, z);
    const f2 x3 = fma2(x2, x, z);
    return add2(add2(add2(y1, fma2(a1, x, z)), fma2(a2, x2, z)), fma2(a3, x3, z));
}
__global__ void kz(const float *in, int n, int *bad, float *ex, float zneg)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + 10 * (size_t)i;
    float a = interp_cubic(p[0], p[1], p[2], p[3], p[4]), b = interp_cubic(p[5], p[6], p[7], p[8], p[9]);
    f2 r = interp_cubic2z(mk2(p[0], p[5]), mk2(p[1], p[6]), mk2(p[2], p[7]), mk2(p[3], p[8]), mk2(p[4], p[9]), mk2(zneg, zneg));
    if (__float_as_uint(a) != __float_as_uint(r.x) || __float_as_uint(b) != __float_as_uint(r.y)) {
        int s = atomicAdd(bad, 1);
        if (s < 4) { for (int j = 0; j < 10; ++j) ex[14 * s + j] = p[j]; ex[14 * s + 10] = a; ex[14 * s + 11] = r.x; ex[14 * s + 12] = b; ex[14 * s + 13] = r.y; }
    }
}
__global__ void k(const float *in, int n, int *bad, float *ex)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + 10 * (size_t)i;
    float a = interp_cubic(p[0], p[1], p[2], p[3], p[4]), b = interp_cubic(p[5], p[6], p[7], p[8], p[9]);
    f2 r = interp_cubic2(mk2(p[0], p[5]), mk2(p[1], p[6]), mk2(p[2], p[7]), mk2(p[3], p[8]), mk2(p[4], p[9]));
    if (__float_as_uint(a) != __float_as_uint(r.x) || __float_as_uint(b) != __float_as_uint(r.y)) {
        int s = atomicAdd(bad, 1);
        if (s < 4) { for (int j = 0; j < 10; ++j) ex[14 * s + j] = p[j]; ex[14 * s + 10] = a; ex[14 * s + 11] = r.x; ex[14 * s + 12] = b; ex[14 * s + 13] = r.y; }
    }
}
int main()
{
    const int n = 1 << 20;
    float *h = new float[10 * (size_t)n];
    uint64_t s = 12345;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) & 0xffffff) / 16777216.0f; };
    for (int i = 0; i < n; ++i)
        for (int half = 0; half < 2; ++half) {
            float *p = h + 10 * (size_t)i + 5 * half;
            const int mode = i & 3;
            p[0] = mode == 0 ? rnd() : (mode == 1 ? rnd() * 1e-11f : (mode == 2 ? 1.0f - rnd() * 1e-7f : rnd() * 1e-4f));
            const float amp = (i & 4) ? 1e-3f : 0.3f;
            for (int j = 1; j < 5; ++j) p[j] = (i & 8) && j == 1 ? 0.0f : (rnd() * 2.0f - 1.0f) * amp;
        }
    float *d; int *bad; float *ex;
    cudaMalloc(&d, sizeof(float) * 10 * (size_t)n); cudaMalloc(&bad, 4); cudaMalloc(&ex, sizeof(float) * 56);
    cudaMemcpy(d, h, sizeof(float) * 10 * (size_t)n, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 4);
    k<<<n / 256, 256>>>(d, n, bad, ex);
    int hb = 0; float hex[56];
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hex, ex, sizeof(hex), cudaMemcpyDeviceToHost);
    printf("mismatching pairs: %d of %d (%s)\n", hb, n, cudaGetErrorString(cudaGetLastError()));
    for (int e = 0; e < (hb < 4 ? hb : 4); ++e) {
        printf(" inputs:"); for (int j = 0; j < 10; ++j) printf(" %.9g", hex[14 * e + j]);
        printf("\n  scalar %.9g packed %.9g | scalar %.9g packed %.9g\n", hex[14 * e + 10], hex[14 * e + 11], hex[14 * e + 12], hex[14 * e + 13]);
    }
    // variant Z (opaque -0.0f addend)
    cudaMemset(bad, 0, 4);
    kz<<<n / 256, 256>>>(d, n, bad, ex, -0.0f);
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hex, ex, sizeof(hex), cudaMemcpyDeviceToHost);
    printf("variant Z (fma with an opaque -0): mismatching pairs: %d of %d (%s)\n", hb, n, cudaGetErrorString(cudaGetLastError()));
    for (int e = 0; e < (hb < 4 ? hb : 4); ++e) {
        printf(" inputs:"); for (int j = 0; j < 10; ++j) printf(" %.9g", hex[14 * e + j]);
        printf("\n  scalar %.9g packed %.9g | scalar %.9g packed %.9g\n", hex[14 * e + 10], hex[14 * e + 11], hex[14 * e + 12], hex[14 * e + 13]);
    }
    return 0;
}
