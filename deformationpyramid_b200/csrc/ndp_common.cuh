// Shared definitions of the sm_100a kernels: platform shim (CUDA vs the test-only CPU emulation),
// bulk-copy (TMA) + mbarrier helpers, parameter-block layout, tile constants.
#pragma once
#include <stdint.h>

#ifdef NDP_EMU
#include "cuda_emu.h"
#define NDP_LAUNCH(kernel, grid, block, smem, stream, ...) ndp_emu::launch(kernel, grid, block, smem, __VA_ARGS__)
#define NDP_LAUNCH_PRIO(cls, kernel, grid, block, smem, stream, ...) ndp_emu::launch(kernel, grid, block, smem, __VA_ARGS__)
#define NDP_DYN_SMEM(name) unsigned char* name = ndp_emu::dyn_smem()
#else
#include <cuda_runtime.h>
#define NDP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define NDP_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
// launch with a per-launch scheduling priority (cudaLaunchAttributePriority; lower number = served first when an SM frees up)
extern int ndp_launch_prio[2];      // [0] tensor-core kernels, [1] everything else; 0 / 0 = plain launches (ndp_cabi.cu)
#define NDP_LAUNCH_PRIO(cls, kernel, grid, block, smem, strm_, ...) do { \
        if (ndp_launch_prio[0] == ndp_launch_prio[1]) { kernel<<<grid, block, smem, strm_>>>(__VA_ARGS__); } else { \
            cudaLaunchConfig_t lc_ = {}; lc_.gridDim = grid; lc_.blockDim = block; lc_.dynamicSmemBytes = smem; lc_.stream = strm_; \
            cudaLaunchAttribute la_[1]; la_[0].id = cudaLaunchAttributePriority; la_[0].val.priority = ndp_launch_prio[cls]; \
            lc_.attrs = la_; lc_.numAttrs = 1; cudaLaunchKernelEx(&lc_, kernel, __VA_ARGS__); } } while (0)
#endif

#include "ndp_math.cuh"

// ---- packed fp32 pairs (Blackwell FADD2 / FMUL2 / FFMA2): two values per instruction ----------------
// Each component is the same IEEE round-to-nearest operation as the scalar __fadd_rn / __fmul_rn / __fmaf_rn.
#ifdef NDP_EMU
struct NdpF2 { float a, b; };
static inline NdpF2 ndp_f2_make(float a, float b) { return NdpF2{a, b}; }
static inline NdpF2 ndp_f2_bcast(float v) { return NdpF2{v, v}; }
static inline void ndp_f2_get(NdpF2 v, float& a, float& b) { a = v.a; b = v.b; }
static inline NdpF2 ndp_f2_add(NdpF2 x, NdpF2 y) { return NdpF2{__fadd_rn(x.a, y.a), __fadd_rn(x.b, y.b)}; }
static inline NdpF2 ndp_f2_sub(NdpF2 x, NdpF2 y) { return NdpF2{__fsub_rn(x.a, y.a), __fsub_rn(x.b, y.b)}; }
static inline NdpF2 ndp_f2_mul(NdpF2 x, NdpF2 y) { return NdpF2{__fmul_rn(x.a, y.a), __fmul_rn(x.b, y.b)}; }
static inline NdpF2 ndp_f2_fma(NdpF2 x, NdpF2 y, NdpF2 z) { return NdpF2{__fmaf_rn(x.a, y.a, z.a), __fmaf_rn(x.b, y.b, z.b)}; }
#else
typedef unsigned long long NdpF2;
__device__ __forceinline__ NdpF2 ndp_f2_make(float a, float b) { NdpF2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ NdpF2 ndp_f2_bcast(float v) { return ndp_f2_make(v, v); }
__device__ __forceinline__ void ndp_f2_get(NdpF2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ NdpF2 ndp_f2_add(NdpF2 x, NdpF2 y) { NdpF2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y)); return r; }
__device__ __forceinline__ NdpF2 ndp_f2_sub(NdpF2 x, NdpF2 y) { NdpF2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y)); return r; }
__device__ __forceinline__ NdpF2 ndp_f2_mul(NdpF2 x, NdpF2 y) { NdpF2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y)); return r; }
__device__ __forceinline__ NdpF2 ndp_f2_fma(NdpF2 x, NdpF2 y, NdpF2 z) { NdpF2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z)); return r; }
#endif

// ------------------------------------------------------------------------------------------------
// Tile constants.  The MLP kernels are specialised for the reference's hidden width 128
// (config/NDP.yaml:26, shape_transfer.py:43); other widths are rejected at the C-ABI.
// ------------------------------------------------------------------------------------------------
#define NDP_W 128          // hidden width
#define NDP_PITCH 132      // smem row pitch (floats) of a [points][128] tile: 4-bank skew per row
#define NDP_TP 128         // points per CTA tile
#define NDP_THREADS 256    // threads per CTA of the MLP kernels
#define NDP_MAX_HIDDEN 8   // depth-1 <= 8
#define NDP_ZPITCH 12      // saved head vector pitch (floats)

// Flat parameter block of one pyramid level, in nn.Module.parameters() order of the reference
// NDPLayer (model/nets.py:75-103): input.0.{weight[W,6],bias[W]}, mlp.pts_linears.l.{weight[W,W],
// bias[W]}, rot_brach.{weight[R,W],bias[R]}, s_branch (Sim3), trn_branch.{weight[3,W],bias[3]},
// nr_branch (nonrigidity).  The "pack" block holds the transposed copies the forward kernel
// streams through shared memory: WT_in[6][W], WT_l[W][W] (k-major).
struct NdpLayout {
    int hidden;                 // L = depth - 1
    int motion, rot, nonrigid;
    int head_dim;               // rows of the head matrix = NdpHeadIdx.dim
    int off_w_in, off_b_in;
    int off_w[NDP_MAX_HIDDEN], off_b[NDP_MAX_HIDDEN];
    int head_w[NDP_MAX_HEAD], head_b[NDP_MAX_HEAD];   // per head row: weight row / bias offsets
    int param_count;
    int pack_in, pack_w[NDP_MAX_HIDDEN], pack_count;
    int pack_img;               // float offset of the fp16 hi/lo images of the hidden weights (tensor-core kernels)
    float freq, mu;
};

static inline NdpLayout ndp_make_layout(int depth, int motion, int rot, int nonrigid, float freq, float mu) {
    NdpLayout L;
    L.hidden = depth - 1; L.motion = motion; L.rot = rot; L.nonrigid = nonrigid ? 1 : 0;
    L.freq = freq; L.mu = mu;
    int o = 0;
    L.off_w_in = o; o += NDP_W * 6;
    L.off_b_in = o; o += NDP_W;
    for (int l = 0; l < NDP_MAX_HIDDEN; ++l) { L.off_w[l] = 0; L.off_b[l] = 0; L.pack_w[l] = 0; }
    for (int l = 0; l < L.hidden; ++l) {
        L.off_w[l] = o; o += NDP_W * NDP_W;
        L.off_b[l] = o; o += NDP_W;
    }
    NdpHeadIdx h = ndp_head_idx(motion, rot, nonrigid);
    L.head_dim = h.dim;
    for (int r = 0; r < NDP_MAX_HEAD; ++r) { L.head_w[r] = 0; L.head_b[r] = 0; }
    int row = 0;
    int R = ndp_rot_dim(motion, rot);
    if (R > 0) {
        for (int r = 0; r < R; ++r) { L.head_w[row + r] = o + r * NDP_W; L.head_b[row + r] = o + R * NDP_W + r; }
        o += R * NDP_W + R; row += R;
        if (motion == NDP_MOTION_SIM3) { L.head_w[row] = o; L.head_b[row] = o + NDP_W; o += NDP_W + 1; row += 1; }
    }
    for (int r = 0; r < 3; ++r) { L.head_w[row + r] = o + r * NDP_W; L.head_b[row + r] = o + 3 * NDP_W + r; }
    o += 3 * NDP_W + 3; row += 3;
    if (nonrigid) { L.head_w[row] = o; L.head_b[row] = o + NDP_W; o += NDP_W + 1; row += 1; }
    L.param_count = o;
    int p = 0;
    L.pack_in = p; p += 6 * NDP_W;
    for (int l = 0; l < L.hidden; ++l) { L.pack_w[l] = p; p += NDP_W * NDP_W; }
    L.pack_img = p; p += L.hidden * (2 * 128 * 128 * 2 / 4);   // hi / lo fp16 images of 128x128 per hidden layer
    L.pack_count = p;
    return L;
}

// ------------------------------------------------------------------------------------------------
// Bulk asynchronous copy global -> shared (TMA, non-tensor form: cp.async.bulk, SASS UBLKCP) with
// mbarrier transaction-count completion.  Sizes and addresses must be multiples of 16 bytes.
// ------------------------------------------------------------------------------------------------
#ifdef NDP_EMU
struct NdpMbar { volatile long long pending; volatile unsigned phase; unsigned count; volatile unsigned arrived; unsigned pad; };
static inline void ndp_mbar_init(NdpMbar* b, int count) { b->pending = 0; b->phase = 0; b->count = (unsigned)count; b->arrived = 0; }
// plain arrival (no transaction bytes): the phase completes when `count` arrivals have been made
static inline void ndp_mbar_arrive(NdpMbar* b) {
    const unsigned have = __atomic_add_fetch((unsigned*)&b->arrived, 1u, __ATOMIC_SEQ_CST);
    if (have == b->count) { __atomic_store_n((unsigned*)&b->arrived, 0u, __ATOMIC_SEQ_CST); __atomic_fetch_add((unsigned*)&b->phase, 1u, __ATOMIC_SEQ_CST); }
}
static inline void ndp_mbar_expect_tx(NdpMbar* b, unsigned bytes) { __atomic_fetch_add((long long*)&b->pending, (long long)bytes, __ATOMIC_SEQ_CST); }
static inline void ndp_bulk_g2s(void* dst, const void* src, unsigned bytes, NdpMbar* b) {
    memcpy(dst, src, bytes);
    long long left = __atomic_sub_fetch((long long*)&b->pending, (long long)bytes, __ATOMIC_SEQ_CST);
    if (left == 0) __atomic_fetch_add((unsigned*)&b->phase, 1u, __ATOMIC_SEQ_CST);
}
static inline void ndp_mbar_wait(NdpMbar* b, unsigned parity) {
    int us = 1;   // sleep instead of spinning: the emulated MMA runs in ONE of the block's OS threads
    long long slept = 0;
    while ((__atomic_load_n((unsigned*)&b->phase, __ATOMIC_SEQ_CST) & 1u) == parity) {
        std::this_thread::sleep_for(std::chrono::microseconds(us));
        slept += us;
        if (us < 2000) us *= 2;
        if (slept > 120LL * 1000 * 1000) { fprintf(stderr, "[emu] DEADLOCK in mbarrier wait (parity %u, thread %u)\n", parity, threadIdx.x); abort(); }
    }
}
static inline void ndp_fence_proxy_async() {}
#else
struct __align__(8) NdpMbar { unsigned long long v; };
__device__ __forceinline__ unsigned ndp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ndp_mbar_init(NdpMbar* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ndp_smem_u32(b)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ndp_mbar_expect_tx(NdpMbar* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ndp_smem_u32(b)), "r"(bytes) : "memory");
}
// plain arrival (release at CTA scope): the phase completes when the barrier's arrival count is reached
__device__ __forceinline__ void ndp_mbar_arrive(NdpMbar* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ndp_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void ndp_bulk_g2s(void* dst, const void* src, unsigned bytes, NdpMbar* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ndp_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ndp_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void ndp_mbar_wait(NdpMbar* b, unsigned parity) {
    // try_wait suspends the thread in hardware for a bounded time; the loop is bounded too, so a
    // protocol error surfaces as a trapped kernel (CUDA error) instead of a hung GPU
    const unsigned addr = ndp_smem_u32(b);
    for (unsigned spin = 0; spin < (1u << 24); ++spin) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void ndp_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// Align a pointer into the dynamic shared-memory array WITHOUT a round trip through an integer (which would lose the
// address space: every access through the result would compile to generic LD / ST instead of LDS / STS).
#ifdef NDP_EMU
#define NDP_SMEM_ALIGN(ptr, A) ((unsigned char*)(((uintptr_t)(ptr) + ((A) - 1)) & ~(uintptr_t)((A) - 1)))
#else
#define NDP_SMEM_ALIGN(ptr, A) ((ptr) + (((unsigned)(A) - (ndp_smem_u32(ptr) & ((unsigned)(A) - 1u))) & ((unsigned)(A) - 1u)))
#endif

// Warp-uniform helpers.  tcgen05.mma takes its operands from UNIFORM registers: issued from a branch the
// compiler cannot prove single-threaded (e.g. `tid == 0`) every MMA is wrapped in a per-thread
// "waterfall" loop (VOTEU / ELECT / 7 x R2UR / UTCHMMA / BRA, ~100 cycles).  Branching on elect.sync
// of a warp whose index came through a shuffle broadcast lets ptxas keep everything in uniform registers.
#ifdef NDP_EMU
static inline bool ndp_elect_one() { return (threadIdx.x & 31) == 0; }
static inline int ndp_warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
#else
__device__ __forceinline__ bool ndp_elect_one() {      // exactly one lane of the (converged) warp; always the same lane
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int ndp_warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }
#endif

// Named barrier over `n` threads (a multiple of 32) of the CTA: bar.sync id, n  (id 1..15; 0 is __syncthreads)
#ifdef NDP_EMU
static inline void ndp_group_sync(int id, int n) { ndp_emu_named_sync(id, n); }
#else
__device__ __forceinline__ void ndp_group_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#endif

// Stage `bytes` (multiple of 16) with bulk copies of at most 32 KiB each; one thread calls this.
__device__ __forceinline__ void ndp_stage_bulk(void* dst, const void* src, unsigned bytes, NdpMbar* bar) {
    ndp_fence_proxy_async();
    ndp_mbar_expect_tx(bar, bytes);
    const unsigned CH = 32768u;
    for (unsigned o = 0; o < bytes; o += CH) {
        unsigned n = bytes - o < CH ? bytes - o : CH;
        ndp_bulk_g2s((char*)dst + o, (const char*)src + o, n, bar);
    }
}

// Per-pair optimisation state (device resident; model/registration.py:179-180, 225-232).
struct NdpPairState {
    int stopped;         // 1 once the early-stop rule fired for the current level
    int steps;           // Adam steps taken in the current level
    int break_counter;   // cumulative per level
    int evals;           // forward+loss evaluations in the current level
    double loss_prev;    // initial 1e6
    float last_loss;
    float pad;
};
