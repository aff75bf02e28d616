// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue/throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<const unsigned long long*>(&b);
    unsigned long long rd = *reinterpret_cast<unsigned long long*>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    d = *reinterpret_cast<float2*>(&rd);
}

template <int MODE>
__global__ void k(float* out, int iters, float a0) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 a = make_float2(a0, a0), b = make_float2(0.999f, 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(acc[i].x, b.x, a.x); acc[i].y = fmaf(acc[i].y, b.y, a.y); }
            else if (MODE == 1) { float2 t = acc[i]; unsigned long long rt = *reinterpret_cast<unsigned long long*>(&t), rb = *reinterpret_cast<const unsigned long long*>(&b), ra = *reinterpret_cast<const unsigned long long*>(&a);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(rt) : "l"(rb), "l"(ra)); acc[i] = *reinterpret_cast<float2*>(&rt); }
            else { acc[i].x = fmaf(acc[i].x, 0.999f, 1.25f); acc[i].y = fmaf(acc[i].y, 0.999f, 1.25f); }   // immediates
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode) {
        for (int warps = 4; warps <= 32; warps *= 2) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            dim3 grid(148), block(warps * 32);
            auto run = [&] { if (mode == 0) k<0><<<grid, block>>>(out, iters, 0.5f); else if (mode == 1) k<1><<<grid, block>>>(out, iters, 0.5f); else k<2><<<grid, block>>>(out, iters, 0.5f); };
            run(); cudaDeviceSynchronize();
            cudaEventRecord(a); run(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double fma = double(148) * warps * 32 * iters * 32.0;   // scalar FMAs
            printf("mode %d (%s) warps/SM %2d: %.3f ms  %.1f TFMA/s (%.1f TFLOP/s)\n", mode, mode == 0 ? "FFMA reg" : mode == 1 ? "FFMA2" : "FFMA imm", warps, ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
