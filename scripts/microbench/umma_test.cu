// Bring-up test for the tcgen05 (UMMA) path: 128x128x128 GEMM with bf16x3 operand splitting
// (6 products, fp32 accumulation in TMEM) on the no-swizzle core-matrix shared-memory layout,
// in the three operand-major combinations the warp-field kernels need.
//   mode 0: D[m][n] = sum_k A[m][k] * B[n][k]   (A K-major, B K-major)      forward:  h W^T
//   mode 1: D[m][n] = sum_k A[k][m] * B[k][n]   (A MN-major, B MN-major)    dW = delta^T h
//   mode 2: D[m][n] = sum_k A[m][k] * B[k][n]   (A K-major, B MN-major)     dH = delta W
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_test umma_test.cu && ./umma_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define IMG_BYTES 32768          // one 128x128 bf16 image
#define RS 2048                  // byte stride between 8-row groups
#define CS 128                   // byte stride between 8-column chunks (one 8x16B core matrix)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(lbo >> 4) << 16;
    d |= (unsigned long long)(sbo >> 4) << 32;
    d |= 1ull << 46;                                   // descriptor version (Blackwell)
    return d;                                          // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ int mbar_wait_bounded(unsigned long long* bar, unsigned parity) {
    unsigned addr = smem_u32(bar);
    for (int spin = 0; spin < (1 << 22); ++spin) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return 1;
    }
    return 0;
}

// element (r, c) of a 128x128 bf16 image in the core-matrix layout
__device__ __forceinline__ unsigned img_off(int r, int c) { return (r >> 3) * RS + (c >> 3) * CS + (r & 7) * 16 + (c & 7) * 2; }

__global__ void __launch_bounds__(128) umma_gemm(const float* A, const float* B, float* D, int mode, int reps, int* err) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* imgA = smem;                         // 3 images
    unsigned char* imgB = smem + 3 * IMG_BYTES;         // 3 images
    __shared__ unsigned tmem_base_s;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // split the fp32 operands into three bf16 images each (row r = tid of the STORED matrix [128][128])
    for (int c = 0; c < 128; ++c) {
        for (int which = 0; which < 2; ++which) {
            float x = (which ? B : A)[tid * 128 + c];
            unsigned char* img = which ? imgB : imgA;
            __nv_bfloat16 b1 = __float2bfloat16_rn(x);
            float r1 = x - __bfloat162float(b1);
            __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
            float r2 = r1 - __bfloat162float(b2);
            __nv_bfloat16 b3 = __float2bfloat16_rn(r2);
            *(__nv_bfloat16*)(img + 0 * IMG_BYTES + img_off(tid, c)) = b1;
            *(__nv_bfloat16*)(img + 1 * IMG_BYTES + img_off(tid, c)) = b2;
            *(__nv_bfloat16*)(img + 2 * IMG_BYTES + img_off(tid, c)) = b3;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_base_s;

    const int a_mn = (mode == 1), b_mn = (mode >= 1);
    // instruction descriptor: D=f32, A=B=bf16, M=128, N=128
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) | (16u << 17) | (8u << 24);
    if (tid == 0) {
        for (int rep = 0; rep < reps; ++rep) {
            int first = 1;
            // products (i, j) with i + j <= 2 : a1b1, a1b2, a2b1, a1b3, a2b2, a3b1
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j + i < 3; ++j)
                    for (int ks = 0; ks < 8; ++ks) {
                        unsigned aaddr = smem_u32(imgA + i * IMG_BYTES), baddr = smem_u32(imgB + j * IMG_BYTES);
                        unsigned long long da = a_mn ? make_desc(aaddr + ks * 2 * RS, /*lbo=*/RS, /*sbo=*/CS)
                                                     : make_desc(aaddr + ks * 2 * CS, /*lbo=*/CS, /*sbo=*/RS);
                        unsigned long long db = b_mn ? make_desc(baddr + ks * 2 * RS, RS, CS)
                                                     : make_desc(baddr + ks * 2 * CS, CS, RS);
                        umma_bf16(tmem, da, db, idesc, first ? 0u : 1u);
                        first = 0;
                    }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (!mbar_wait_bounded(&bar, 0)) { if (lane == 0) atomicAdd(err, 1); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // TMEM -> registers: warp w owns lanes 32w..32w+31; 4 chunks of 32 columns
    for (int c0 = 0; c0 < 128; c0 += 32) {
        unsigned v[32];
        unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(size_t)blockIdx.x * 128 * 128 + (warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
    std::vector<float> A(128 * 128), B(128 * 128), D(128 * 128);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
    float *dA, *dB, *dD; int* derr;
    const int blocks = 148;
    cudaMalloc(&dA, 65536); cudaMalloc(&dB, 65536); cudaMalloc(&dD, 65536 * blocks); cudaMalloc(&derr, 4);
    cudaMemcpy(dA, A.data(), 65536, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), 65536, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(umma_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * IMG_BYTES);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(derr, 0, 4); cudaMemset(dD, 0, 65536);
        umma_gemm<<<1, 128, 6 * IMG_BYTES>>>(dA, dB, dD, mode, 1, derr);
        cudaError_t e = cudaDeviceSynchronize();
        int herr = 0; cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D.data(), dD, 65536, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                double r = 0;
                for (int k = 0; k < 128; ++k) {
                    double a = (mode == 1) ? A[k * 128 + m] : A[m * 128 + k];
                    double b = (mode == 0) ? B[n * 128 + k] : B[k * 128 + n];
                    r += a * b;
                }
                maxerr = fmax(maxerr, fabs(r - D[m * 128 + n])); maxref = fmax(maxref, fabs(r));
            }
        printf("mode %d: cuda=%s timeouts=%d max_abs_err=%.3e max_ref=%.3e rel=%.3e\n", mode, cudaGetErrorString(e), herr, maxerr, maxref, maxerr / maxref);
        if (e != cudaSuccess) return 1;
    }
    // throughput: 148 CTAs x reps GEMMs (6 products each)
    const int reps = 200;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    umma_gemm<<<blocks, 128, 6 * IMG_BYTES>>>(dA, dB, dD, 0, reps, derr); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    umma_gemm<<<blocks, 128, 6 * IMG_BYTES>>>(dA, dB, dD, 0, reps, derr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mac = (double)blocks * reps * 6 * 128.0 * 128 * 128;
    printf("throughput: %.3f ms for %d CTAs x %d GEMM(6 products): %.1f dense bf16 TFLOP/s, %.2f us per fp32-accurate 128^3 GEMM per SM (incl. fixed overhead %s)\n",
           ms, blocks, reps, 2 * mac / ms / 1e9, ms * 1e3 / reps, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
