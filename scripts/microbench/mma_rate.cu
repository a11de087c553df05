// Microbenchmark: legacy mma.sync throughput on sm_100a (TF32 m16n8k8, BF16 m16n8k16) and FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__global__ void k_tf32(float* out, int iters) {
    float c[16][4];
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f810000u, 0x3f820000u};
    unsigned b[2] = {0x3f800000u, 0x3f830000u + threadIdx.x};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_bf16(float* out, int iters) {
    float c[16][4];
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f813f80u, 0x3f823f80u};
    unsigned b[2] = {0x3f803f80u, 0x3f833f80u + threadIdx.x};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, int iters, float x, float y) {
    float c[64];
    for (int i = 0; i < 64; ++i) c[i] = threadIdx.x * 1e-3f + i;
    float a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = x + i; b[i] = y - i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) c[i * 8 + j] = fmaf(a[i], b[j], c[i * 8 + j]);
    }
    float s = 0.f;
    for (int i = 0; i < 64; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    int sms = 148; cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, 64 << 20);
    const int iters = 4096;
    for (int warps : {4, 8, 16}) for (int cps : {1, 2}) {
        dim3 grid(sms * cps), block(warps * 32);
        float ms = timeit([&] { k_tf32<<<grid, block>>>(out, iters); });
        double mac = (double)grid.x * warps * iters * 16 * (16 * 8 * 8);
        printf("tf32 m16n8k8  warps/CTA=%2d CTA/SM=%d : %8.1f TFLOP/s  %7.1f MAC/clk/SM (at 1.965GHz)\n", warps, cps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / 1.965e9);
        ms = timeit([&] { k_bf16<<<grid, block>>>(out, iters); });
        mac = (double)grid.x * warps * iters * 16 * (16 * 8 * 16);
        printf("bf16 m16n8k16 warps/CTA=%2d CTA/SM=%d : %8.1f TFLOP/s  %7.1f MAC/clk/SM\n", warps, cps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / 1.965e9);
        ms = timeit([&] { k_ffma<<<grid, block>>>(out, iters, 1.0001f, 0.9999f); });
        mac = (double)grid.x * warps * 32 * iters * 64;
        printf("ffma          warps/CTA=%2d CTA/SM=%d : %8.1f TFLOP/s  %7.1f FMA/clk/SM\n", warps, cps, 2 * mac / ms / 1e9, mac / (ms * 1e-3) / sms / 1.965e9);
    }
    return 0;
}
