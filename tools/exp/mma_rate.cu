// Microbenchmark: throughput of the legacy-path mma.sync.m16n8k8 TF32 on sm_100a (is the J multiply of the Wigner chain
// affordable on it?) and of cvt.rna.tf32.f32.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(512, 1) mma_kernel(float* out, int iters) {
    float d[CHAINS][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int c = 0; c < CHAINS; ++c)
        for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int c = 0; c < CHAINS; ++c)
        for (int i = 0; i < 4; ++i) s += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the chained form the Wigner kernel would use: accumulator of one MMA becomes (after a tf32 split) the A operand of the next
__global__ void __launch_bounds__(512, 1) chain_kernel(float* out, int iters) {
    float x[2][4];
    unsigned b[2] = {__float_as_uint(0.25f), __float_as_uint(-0.125f)};
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 4; ++i) x[c][i] = 1.0f + threadIdx.x * 1e-3f + i + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            unsigned hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi[i]) : "f"(x[c][i]));
                lo[i] = __float_as_uint(x[c][i] - __uint_as_float(hi[i]));
            }
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(b[0]), "r"(b[1]));
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(b[1]), "r"(b[0]));
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(b[0]), "r"(b[1]));
#pragma unroll
            for (int i = 0; i < 4; ++i) x[c][i] = d[i];
        }
    }
    float s = 0.f;
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 4; ++i) s += x[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    const int iters = 20000;
    const double warps = 148.0 * 16;
    float ms;
    ms = time_ms([&] { mma_kernel<1><<<148, 512>>>(out, iters); });
    printf("m16n8k8 tf32, 16 warps/SM, 1 dependent chain : %.3f ms  %.1f TFLOP/s  %.2f cycles/mma/warp@1.9GHz\n", ms, warps * iters * 2048.0 / ms / 1e9,
           ms * 1e-3 * 1.9e9 / iters);
    ms = time_ms([&] { mma_kernel<4><<<148, 512>>>(out, iters); });
    printf("m16n8k8 tf32, 16 warps/SM, 4 independent     : %.3f ms  %.1f TFLOP/s\n", ms, warps * iters * 4 * 2048.0 / ms / 1e9);
    ms = time_ms([&] { mma_kernel<8><<<148, 512>>>(out, iters); });
    printf("m16n8k8 tf32, 16 warps/SM, 8 independent     : %.3f ms  %.1f TFLOP/s\n", ms, warps * iters * 8 * 2048.0 / ms / 1e9);
    ms = time_ms([&] { chain_kernel<<<148, 512>>>(out, iters); });
    printf("3xTF32 chained (split + 3 mma) x2 per iter   : %.3f ms  %.1f TFLOP/s of mma, %.1f G chained 16x8x8 stages/s\n", ms,
           warps * iters * 6 * 2048.0 / ms / 1e9, warps * iters * 2 / ms / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
