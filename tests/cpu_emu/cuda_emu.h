// TEST INFRASTRUCTURE ONLY -- never built into, loaded by, or shipped with the product library.
//
// A minimal CUDA-on-CPU execution shim: it lets the *same* kernel sources under
// deformationpyramid_b200/csrc/ be compiled with g++ (-DNDP_EMU) and executed with one OS thread
// per CUDA thread, one block at a time, so that tiling / indexing / barrier logic can be
// checked against the oracle in the GPU-less build container (tests/test_emu_*.py).  The
// product path has no CPU fallback: deformationpyramid_b200/_lib.py loads only the nvcc-built
// libndp_b200.so and raises if it or a CUDA device is missing.
#pragma once
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <chrono>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace ndp_emu {

struct Barrier {
    std::mutex mu;
    std::condition_variable cv;
    int count = 0, waiting = 0;
    uint64_t gen = 0;
    void reset(int n) { count = n; waiting = 0; }
    const char* name = "barrier";
    void stuck(int have) {   // watchdog: a protocol error becomes a diagnosable abort instead of a hung test
        fprintf(stderr, "[emu] DEADLOCK at %s: %d of %d threads arrived (block %u %u %u, thread %u)\n", name, have, count,
                0u, 0u, 0u, 0u);
        fflush(stderr);
        abort();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        uint64_t g = gen;
        if (++waiting == count) { waiting = 0; ++gen; cv.notify_all(); }
        else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g; })) stuck(waiting);
    }
    void wait_n(int n) {      // named barrier: every participant passes the same thread count
        std::unique_lock<std::mutex> lk(mu);
        count = n;
        uint64_t g = gen;
        if (++waiting == count) { waiting = 0; ++gen; cv.notify_all(); }
        else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g; })) stuck(waiting);
    }
};

struct Ctx {
    uint3 tid{0, 0, 0}, bid{0, 0, 0};
    dim3 bdim, gdim;
    int lane = 0, warp = 0;
};
extern thread_local Ctx ctx;
extern Barrier block_barrier;
extern Barrier named_barriers[16];
extern std::vector<Barrier*> warp_barriers;
extern std::vector<uint64_t> warp_slots;   // [warp][32]
extern unsigned char* dyn_smem_ptr;

inline unsigned char* dyn_smem() { return dyn_smem_ptr; }

void prepare(unsigned nthreads, size_t smem);

template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
    unsigned nthreads = block.x * block.y * block.z;
    prepare(nthreads, smem);
    if (getenv("NDP_EMU_TRACE")) { fprintf(stderr, "[emu] launch %p grid %u %u %u block %u\n", (void*)kernel, grid.x, grid.y, grid.z, nthreads); fflush(stderr); }
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) {
        pool.emplace_back([=]() {
            ctx.bdim = block; ctx.gdim = grid;
            ctx.tid.x = t % block.x; ctx.tid.y = (t / block.x) % block.y; ctx.tid.z = t / (block.x * block.y);
            ctx.lane = t & 31; ctx.warp = t >> 5;
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        ctx.bid = uint3{bx, by, bz};
                        kernel(args...);
                        block_barrier.wait();   // blocks run one after another
                    }
        });
    }
    for (auto& th : pool) th.join();
}

uint64_t warp_exchange(uint64_t v, int src_lane);
unsigned warp_ballot(int pred);

}  // namespace ndp_emu

#define threadIdx (ndp_emu::ctx.tid)
#define blockIdx (ndp_emu::ctx.bid)
#define blockDim (ndp_emu::ctx.bdim)
#define gridDim (ndp_emu::ctx.gdim)

static inline void __syncthreads() { ndp_emu::block_barrier.wait(); }
static inline void ndp_emu_named_sync(int id, int n) { ndp_emu::named_barriers[id & 15].wait_n(n); }
static inline void __syncwarp(unsigned = 0xffffffffu) { ndp_emu::warp_barriers[ndp_emu::ctx.warp]->wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// --- math intrinsics (compile with -ffp-contract=off so that nothing else is contracted)
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __float2ll_rn(float f) { return llrintf(f); }
static inline long long __double2ll_rn(double f) { return llrint(f); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
template <class T> static inline T __ldg(const T* p) { return *p; }

// --- atomics
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicCAS(int* p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

// --- warp collectives
static inline float __shfl_xor_sync(unsigned, float v, int m) {
    uint64_t r = ndp_emu::warp_exchange(__float_as_uint(v), ndp_emu::ctx.lane ^ m);
    return __uint_as_float((unsigned)r);
}
static inline double __shfl_xor_sync(unsigned, double v, int m) {
    uint64_t u; memcpy(&u, &v, 8);
    u = ndp_emu::warp_exchange(u, ndp_emu::ctx.lane ^ m);
    double d; memcpy(&d, &u, 8); return d;
}
static inline int __shfl_xor_sync(unsigned, int v, int m) {
    return (int)ndp_emu::warp_exchange((uint64_t)(unsigned)v, ndp_emu::ctx.lane ^ m);
}
static inline float __shfl_sync(unsigned, float v, int src) {
    uint64_t r = ndp_emu::warp_exchange(__float_as_uint(v), src);
    return __uint_as_float((unsigned)r);
}
static inline int __shfl_sync(unsigned, int v, int src) {
    return (int)ndp_emu::warp_exchange((uint64_t)(unsigned)v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) { return ndp_emu::warp_ballot(pred); }
static inline int __any_sync(unsigned, int pred) { return ndp_emu::warp_ballot(pred) != 0; }

// --- runtime API subset ("device" memory is host memory)
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
typedef double* cudaEvent_t;   // an event is a wall-clock timestamp in the emulation
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new double(0.0); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new double(0.0); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); *e = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(*b - *a); return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
