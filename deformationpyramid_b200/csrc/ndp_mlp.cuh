// Register-tiled fp32 GEMM micro-kernels on [points][128] shared-memory tiles (pitch NDP_PITCH).
// 256 threads, thread (tr = tid>>4, tc = tid&15) owns an 8x8 patch:
//     rows  {tr*4 .. tr*4+3, 64+tr*4 .. 64+tr*4+3}
//     cols  {tc*4 .. tc*4+3, 64+tc*4 .. 64+tc*4+3}
// so that every shared-memory access is a conflict-free 128-bit load: the 16 lanes that differ in
// tc read 256 contiguous bytes, the two tr values of a warp hit rows 4 apart = 16 banks apart.
// FP32 FFMA on the CUDA cores: results must stay within 1e-4 relative of the reference's fp32
// path (BASELINE.json north_star), which rules out single-pass TF32/BF16 tensor-core products.
#pragma once
#include "ndp_common.cuh"

__device__ __forceinline__ int ndp_row8(int t, int r) { return (r < 4) ? (t * 4 + r) : (64 + t * 4 + (r - 4)); }

__device__ __forceinline__ void ndp_acc_zero(float (&acc)[8][8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
}

__device__ __forceinline__ void ndp_fma_row(float (&acc)[8], float a, const float4& b0, const float4& b1) {
    acc[0] = fmaf(a, b0.x, acc[0]); acc[1] = fmaf(a, b0.y, acc[1]);
    acc[2] = fmaf(a, b0.z, acc[2]); acc[3] = fmaf(a, b0.w, acc[3]);
    acc[4] = fmaf(a, b1.x, acc[4]); acc[5] = fmaf(a, b1.y, acc[5]);
    acc[6] = fmaf(a, b1.z, acc[6]); acc[7] = fmaf(a, b1.w, acc[7]);
}

// acc[r][c] += sum_k A[row(r)][k] * B[k][col(c)]     A: smem [128][PITCH], B: smem [K][128] (k-major)
__device__ __forceinline__ void ndp_gemm_nn(const float* __restrict__ A, const float* __restrict__ B,
                                            float (&acc)[8][8], int tr, int tc) {
#pragma unroll 1
    for (int k0 = 0; k0 < NDP_W; k0 += 4) {
        float4 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *(const float4*)(A + ndp_row8(tr, r) * NDP_PITCH + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 b0 = *(const float4*)(B + (k0 + kk) * NDP_W + tc * 4);
            const float4 b1 = *(const float4*)(B + (k0 + kk) * NDP_W + 64 + tc * 4);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float av = (kk == 0) ? a[r].x : (kk == 1) ? a[r].y : (kk == 2) ? a[r].z : a[r].w;
                ndp_fma_row(acc[r], av, b0, b1);
            }
        }
    }
}

// acc[r][c] += sum_p A[p][row(r)] * B[p][col(c)]     A, B: smem [128][PITCH] (reduction over points)
__device__ __forceinline__ void ndp_gemm_tn(const float* __restrict__ A, const float* __restrict__ B,
                                            float (&acc)[8][8], int tr, int tc) {
#pragma unroll 2
    for (int p = 0; p < NDP_TP; ++p) {
        const float4 a0 = *(const float4*)(A + p * NDP_PITCH + tr * 4);
        const float4 a1 = *(const float4*)(A + p * NDP_PITCH + 64 + tr * 4);
        const float4 b0 = *(const float4*)(B + p * NDP_PITCH + tc * 4);
        const float4 b1 = *(const float4*)(B + p * NDP_PITCH + 64 + tc * 4);
        ndp_fma_row(acc[0], a0.x, b0, b1); ndp_fma_row(acc[1], a0.y, b0, b1);
        ndp_fma_row(acc[2], a0.z, b0, b1); ndp_fma_row(acc[3], a0.w, b0, b1);
        ndp_fma_row(acc[4], a1.x, b0, b1); ndp_fma_row(acc[5], a1.y, b0, b1);
        ndp_fma_row(acc[6], a1.z, b0, b1); ndp_fma_row(acc[7], a1.w, b0, b1);
    }
}

// Coalesced load of a [TP][128] tile of a global [n][128] array into smem (rows >= n zero-filled).
__device__ __forceinline__ void ndp_load_tile(float* __restrict__ dst, const float* __restrict__ src,
                                              int tile, int n, int tid) {
#pragma unroll 4
    for (int i = 0; i < (NDP_TP * NDP_W / 4) / NDP_THREADS; ++i) {
        const int idx = tid + i * NDP_THREADS;
        const int row = idx >> 5, c4 = idx & 31;
        const int gp = tile * NDP_TP + row;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (gp < n) v = __ldg((const float4*)(src + (long long)gp * NDP_W) + c4);
        *(float4*)(dst + row * NDP_PITCH + c4 * 4) = v;
    }
}
