// Kernel (1), tensor-core version: one pyramid level of the warp field over TWO tiles of 128 points
// per CTA.
//
// Reference: model/nets.py:111-140 (NDPLayer.forward), :164-177 (posenc), :295-304 (MLP),
//            :144-161 (get_Rotation), model/rigid_body.py.
//
// Every contraction of the layer runs on the 5th-generation tensor cores (tcgen05.mma issued by ONE
// thread per tile, fp32 accumulators in TMEM, operands in shared memory as fp16 hi/lo image sets --
// ndp_tc.cuh: 2-way split, three partial products => fp32-level accuracy):
//     input layer   h_0 = relu([e | 0] [W_in | 0]^T + b_in)     M = 128 points, N = 128, K = 16
//     hidden layers h_l+1 = relu(h_l W_l^T + b_l)               M = 128, N = 128, K = 128
//     heads         z = h_L W_h^T                               M = 128, N = 16,  K = 128
// A CTA is two independent groups of 8 warps, one 128-point tile each, that share the weight images
// (one 64 KB buffer, refilled by TMA bulk copies as soon as BOTH groups' MMAs of the previous layer
// have retired): while one group's MMAs run, the other group's warps do the TMEM -> bias + ReLU ->
// re-split epilogue that writes the next layer's A operand in place, so the tensor pipe and the
// FP32 pipes overlap and the weight traffic per tile halves.  Saved activations leave by TMA bulk
// stores straight from the operand images.  Only the per-point rotation / warp composition
// (rigid_body.py) stays scalar.
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

// optional phase timestamps of CTA (0,0), group 0 (debug aid, read back through ndp_debug_phase_times)
#ifndef NDP_EMU
__device__ unsigned long long ndp_dbg_fwd[64];
#define NDP_T(i) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ndp_dbg_fwd[i] = t_; } } while (0)
#else
#define NDP_T(i) do {} while (0)
#endif

#define NDP_FWD_TC_THREADS 512
#define NDP_FWD_MAX_ROUNDS 4               // tile pairs a CTA processes one after the other (amortises the set-up)
#define NDP_GROUP 256                      // threads per tile group
#define NDP_IMG16 NDP_IMG_BYTES(16)        // [128][16] fp16 image: 4096 bytes
#define NDP_HWIMG (2 * NDP_IMG_RS(128))    // [16][128] fp16 image: 4096 bytes

struct FwdTcSmem {
    unsigned char A[2][NDP_SET128];        // per group: activation hi/lo images (operand A, K-major); first the posenc image
    unsigned char B[NDP_SET128];           // weight hi/lo images of the current hidden layer (operand B, K-major)
    unsigned char WIN[2 * NDP_IMG16];      // [128 outputs][16]: cols 0..5 = W_in, rest 0 (operand B of the input layer)
    unsigned char HW[2 * NDP_HWIMG];       // [16 head rows][128]: head weights, rows >= head_dim 0 (operand B of the heads)
    float xs[2][NDP_TP * 4];
    float hb[16];
    NdpMbar bar_w, bar_mma[2];
    int bcount[NDP_FWD_MAX_ROUNDS * NDP_MAX_HIDDEN];   // per (round, hidden layer): groups whose MMAs have retired
    unsigned tmem_slot, pad[3];
};
size_t ndp_fwd_tc_smem_bytes() { return sizeof(FwdTcSmem) + 1024; }

// ---- per-point rotation + warp composition (nets.py:119-137); outputs, saved head vector, and the
//      float4 copy + 32-point boxes for the culled NN search.  Called by all threads of a tile group;
//      threads gt < 128 own one point each and read their row of the head accumulator from TMEM.
__device__ __forceinline__ void ndp_fwd_point_tail(const NdpFwdArgs& a, const NdpLayout& L, int pair, int tile, int n, int gt,
                                                   int HD, unsigned tlane, const float* xs, const float* hb) {
            if (gt < NDP_TP) {
                const int gp = tile * NDP_TP + gt;
                const float INF = __int_as_float(0x7f800000);
                float y[3] = {INF, INF, INF};
                float zr[16];
                ndp_tmem_ld16(tlane, zr);
                if (gp < n) {
                    float z[NDP_MAX_HEAD], nu = 0.0f;
    #pragma unroll
                    for (int r = 0; r < NDP_MAX_HEAD; ++r) z[r] = (r < HD) ? L.mu * (zr[r] + hb[r]) : 0.0f;
                    ndp_point_forward(L.motion, L.rot, L.nonrigid, z, xs + gt * 4, y, &nu);
                    if (a.y_add) {
                        const float* ya = a.y_add + (long long)pair * a.y_add_stride;
                        y[0] += ya[0]; y[1] += ya[1]; y[2] += ya[2];
                    }
                    float* yp = a.y + (long long)pair * a.y_stride + (long long)gp * 3;
                    yp[0] = y[0]; yp[1] = y[1]; yp[2] = y[2];
                    if (a.nu && L.nonrigid) a.nu[(long long)pair * a.nu_stride + gp] = nu;
                    if (a.zsave) {
                        float4* zp = (float4*)(a.zsave + (long long)pair * a.z_stride + (long long)gp * NDP_ZPITCH);
                        zp[0] = make_float4(z[0], z[1], z[2], z[3]);
                        zp[1] = make_float4(z[4], z[5], z[6], z[7]);
                        zp[2] = make_float4(z[8], z[9], z[10], z[11]);
                    }
                }
                if (a.y4) {
                    const int o = (gp < n) ? a.orig[(long long)pair * a.orig_stride + gp] : 0x7fffffff;
                    a.y4[(long long)pair * a.y4_stride + gp] = make_float4(y[0], y[1], y[2], __int_as_float(o));
                    float l0 = y[0], l1 = y[1], l2 = y[2];
                    float h0 = (gp < n) ? y[0] : -INF, h1 = (gp < n) ? y[1] : -INF, h2 = (gp < n) ? y[2] : -INF;
                    for (int s = 16; s > 0; s >>= 1) {
                        l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, s)); l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, s));
                        l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, s)); h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, s));
                        h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, s)); h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, s));
                    }
                    if ((gt & 31) == 0 && gp < n) {
                        float* bx = a.ybox + ((long long)pair * a.box_stride + (gp >> 5)) * 8;
                        bx[0] = l0; bx[1] = l1; bx[2] = l2; bx[3] = 0.0f; bx[4] = h0; bx[5] = h1; bx[6] = h2; bx[7] = 0.0f;
                    }
                }
            }
}

__global__ void __launch_bounds__(NDP_FWD_TC_THREADS, 1) ndp_warp_fwd_tc_kernel(NdpFwdArgs a) {
    NDP_DYN_SMEM(smem_raw);
    FwdTcSmem& S = *(FwdTcSmem*)NDP_SMEM_ALIGN(smem_raw, 1024);

    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0;
    const int g = tid >> 8, gt = tid & (NDP_GROUP - 1);          // tile group, thread within the group
    const int rounds = a.rounds;                                 // CTA = tiles [2 rounds bx, 2 rounds (bx + 1)): round r, group g -> tile
    const int tile_first = blockIdx.x * 2 * rounds;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile_first * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    int tile = tile_first + g;
    bool active = tile * NDP_TP < n;
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const unsigned char* wimg = (const unsigned char*)(a.pack + (long long)pair * a.pack_stride + L.pack_img);
    const int LH = L.hidden, HD = L.head_dim;
    const int warp = tid >> 5, p = gt & (NDP_TP - 1), half = gt >> 7;
    const bool ldw = (ndp_warp_uniform(warp) & 7) == 0;   // the group's issuing warp: one elected lane launches MMAs / bulk copies
#define NDP_LEADER (ldw && ndp_elect_one())
    const int RS = NDP_IMG_RS(128), CS = NDP_IMG_CS, RS16 = NDP_IMG_RS(16);
    unsigned char* const gact_pair = a.act ? (unsigned char*)a.act + ((long long)pair * a.act_stride) * 4 : nullptr;
    unsigned char* A = S.A[g];
    float* xs = S.xs[g];

    NDP_T(0);
    if (warp == 0) ndp_tmem_alloc_warp(&S.tmem_slot, 256);
    if (tid == 0) { ndp_mbar_init(&S.bar_w, 1); ndp_mbar_init(&S.bar_mma[0], 1); ndp_mbar_init(&S.bar_mma[1], 1); }
    if (tid < NDP_FWD_MAX_ROUNDS * NDP_MAX_HIDDEN) S.bcount[tid] = 0;
    if (tid < 16) S.hb[tid] = (tid < HD) ? __ldg(params + L.head_b[tid]) : 0.0f;
    if (tid < 256) {            // input-layer weight image: row o = tid / 2, 8-column chunk tid & 1
        const int o = tid >> 1, c8 = tid & 1;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c8 == 0 && j < 6) ? __ldg(params + L.off_w_in + o * 6 + j) : 0.0f;
        ndp_store_chunk2(S.WIN, NDP_IMG16, ndp_img_off(o, c8 * 8, RS16), v);
    } else {                    // head weight image: row r = (tid - 256) / 16, 8-column chunk (tid - 256) & 15
        const int r = (tid - 256) >> 4, c8 = (tid - 256) & 15;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (r < HD) ? __ldg(params + L.head_w[r] + c8 * 8 + j) : 0.0f;   // head rows are not 16-byte aligned
        ndp_store_chunk2(S.HW, NDP_HWIMG, ndp_img_off(r, c8 * 8, RS), v);
    }
    auto build_posenc = [&](int tile) {   // points + positional encoding image (nets.py:164-177), in the head of A
      if (gt < NDP_TP) {
        const int gp = tile * NDP_TP + gt;
        float px = 0.0f, py = 0.0f, pz = 0.0f;
        if (gp < n) {
            const float* xp = a.x + (long long)pair * a.x_stride + (long long)gp * 3;
            px = __ldg(xp); py = __ldg(xp + 1); pz = __ldg(xp + 2);
        }
        xs[gt * 4 + 0] = px; xs[gt * 4 + 1] = py; xs[gt * 4 + 2] = pz;
        float e0[8], e1[8], s, c;
        sincosf(px * L.freq, &s, &c); e0[0] = s; e0[1] = c;
        sincosf(py * L.freq, &s, &c); e0[2] = s; e0[3] = c;
        sincosf(pz * L.freq, &s, &c); e0[4] = s; e0[5] = c;
        e0[6] = 0.0f; e0[7] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) e1[j] = 0.0f;
        ndp_store_chunk2(A, NDP_IMG16, ndp_img_off(gt, 0, RS16), e0);
        ndp_store_chunk2(A, NDP_IMG16, ndp_img_off(gt, 8, RS16), e1);
      }
    };
    build_posenc(tile);
    ndp_tc_fence_before();
    ndp_fence_proxy_async();
    __syncthreads();
    ndp_tc_fence_after();
    NDP_T(1);
    const unsigned tmem = S.tmem_slot + (unsigned)g * 128u;
    const unsigned tlane = tmem + ((unsigned)((warp & 3) * 32) << 16);
    if (tid == 0 && LH > 0) ndp_stage_bulk(S.B, wimg, NDP_SET128, &S.bar_w);

    unsigned mph = 0;
    for (int r = 0; r < rounds && active; ++r) {
        const int nactive = ((tile_first + 2 * r + 1) * NDP_TP < n) ? 2 : 1;                 // groups with a tile in this round
        const bool more = r + 1 < rounds && (tile_first + 2 * (r + 1)) * NDP_TP < n;         // the CTA has another round
        unsigned char* gact = gact_pair ? gact_pair + (long long)tile * (LH + 1) * NDP_SET128 : nullptr;
        if (r > 0) {            // this group's next tile: A is free (top-activation store drained by the leader below)
            build_posenc(tile);
            ndp_fence_proxy_async();
            ndp_group_sync(1 + g, NDP_GROUP);
        }
        // ---- stage s = 0: input layer; s = 1..LH: hidden layer s - 1.  Each stage: MMAs by the group's
        //      thread 0, then the epilogue h = relu(acc + bias) re-split into the A images (in place).
        for (int s = 0; s <= LH; ++s) {
            if (NDP_LEADER) {
                if (s == 0) {
                    ndp_umma_gemm3(tmem, ndp_umma_desc(A, CS, RS16), NDP_IMG16, 0, ndp_umma_desc(S.WIN, CS, RS16), NDP_IMG16, 0, 1,
                                   ndp_idesc_f16(128, 128, 0, 0), false);
                } else {
                    if (gact) {      // save h_{s-1} for the backward pass straight from the operand images
                        ndp_bulk_s2g(gact + (long long)(s - 1) * NDP_SET128, A, NDP_IMG128);
                        ndp_bulk_s2g(gact + (long long)(s - 1) * NDP_SET128 + NDP_IMG128, A + NDP_IMG128, NDP_IMG128);
                        ndp_bulk_commit();
                    }
                    ndp_mbar_wait(&S.bar_w, (unsigned)((r * LH + s - 1) & 1));
                    ndp_tc_fence_after();
                    ndp_umma_gemm3(tmem, ndp_umma_desc(A, CS, RS), NDP_IMG128, 2 * CS, ndp_umma_desc(S.B, CS, RS), NDP_IMG128, 2 * CS, 8,
                                   ndp_idesc_f16(128, 128, 0, 0), false);
                }
                ndp_umma_commit(&S.bar_mma[g]);
                NDP_T(8 + 4 * s);
            }
            ndp_mbar_wait(&S.bar_mma[g], mph); mph ^= 1;
            ndp_tc_fence_after();
            NDP_T(9 + 4 * s);
            if (s > 0 && NDP_LEADER) {
                if (gact) ndp_bulk_wait_read0();                  // the store has finished reading A
                // the last group to retire hidden layer s - 1 refills the weight buffer with the next layer
                // (the first layer again when the CTA has another round)
                if ((s < LH || more) && atomicAdd(&S.bcount[r * LH + s - 1], 1) == nactive - 1)
                    ndp_stage_bulk(S.B, wimg + (long long)(s < LH ? s : 0) * NDP_SET128, NDP_SET128, &S.bar_w);
            }
            if (gact && s > 0) ndp_group_sync(1 + g, NDP_GROUP);  // A may be overwritten only after wait_read0
            NDP_T(10 + 4 * s);
            const float* bias = params + (s == 0 ? L.off_b_in : L.off_b[s - 1]);
#pragma unroll 1
            for (int c32 = 0; c32 < 2; ++c32) {
                float v[32];
                const int col0 = half * 64 + c32 * 32;
                ndp_tmem_ld32(tlane + col0, v);
#pragma unroll
                for (int s8 = 0; s8 < 4; ++s8) {
                    float u[8];
                    const float4 b0 = __ldg((const float4*)(bias + col0 + s8 * 8)), b1 = __ldg((const float4*)(bias + col0 + s8 * 8 + 4));
                    u[0] = ndp_relu_img(v[s8 * 8 + 0] + b0.x); u[1] = ndp_relu_img(v[s8 * 8 + 1] + b0.y);
                    u[2] = ndp_relu_img(v[s8 * 8 + 2] + b0.z); u[3] = ndp_relu_img(v[s8 * 8 + 3] + b0.w);
                    u[4] = ndp_relu_img(v[s8 * 8 + 4] + b1.x); u[5] = ndp_relu_img(v[s8 * 8 + 5] + b1.y);
                    u[6] = ndp_relu_img(v[s8 * 8 + 6] + b1.z); u[7] = ndp_relu_img(v[s8 * 8 + 7] + b1.w);
                    ndp_store_chunk2(A, NDP_IMG128, ndp_img_off(p, col0 + s8 * 8, RS), u);
                }
            }
            ndp_tc_fence_before();
            ndp_fence_proxy_async();
            ndp_group_sync(1 + g, NDP_GROUP);
            NDP_T(11 + 4 * s);
        }

        // ---- heads on the tensor core: z_raw[128 points][16] = h_L W_h^T; top activation saved meanwhile
        if (NDP_LEADER) {
            ndp_tc_fence_after();
            if (gact) {
                ndp_bulk_s2g(gact + (long long)LH * NDP_SET128, A, NDP_IMG128);
                ndp_bulk_s2g(gact + (long long)LH * NDP_SET128 + NDP_IMG128, A + NDP_IMG128, NDP_IMG128);
                ndp_bulk_commit();
            }
            ndp_umma_gemm3(tmem, ndp_umma_desc(A, CS, RS), NDP_IMG128, 2 * CS, ndp_umma_desc(S.HW, CS, RS), NDP_HWIMG, 2 * CS, 8,
                           ndp_idesc_f16(128, 16, 0, 0), false);
            ndp_umma_commit(&S.bar_mma[g]);
        }
        ndp_mbar_wait(&S.bar_mma[g], mph); mph ^= 1;
        ndp_tc_fence_after();
        NDP_T(60);

        ndp_fwd_point_tail(a, L, pair, tile, n, gt, HD, tlane, xs, S.hb);
        NDP_T(61);
        if (gact && NDP_LEADER) ndp_bulk_wait0();     // smem must outlive the bulk stores (and A is rebuilt next round)
        NDP_T(62);
        tile += 2;
        active = more && tile * NDP_TP < n;
        if (active) { ndp_tc_fence_before(); ndp_group_sync(1 + g, NDP_GROUP); ndp_tc_fence_after(); }   // head results read, stores drained
    }
    ndp_tc_fence_before();
    __syncthreads();
    NDP_T(63);
    if (warp == 0) ndp_tmem_dealloc(S.tmem_slot, 256);
}

// =================================================================================================
// Version 2 (depth <= 3, i.e. at most two hidden layers -- the reference's configuration): the A
// operand (activations) lives in TENSOR MEMORY instead of shared memory.
//   TMEM per group (256 columns): [0,128) fp32 accumulator, [128,192) hi image, [192,256) lo image of
//   the A operand (fp16 pairs packed in 32-bit cells, K along the columns: 8 columns per 16-deep step).
// The epilogue writes the next layer's operand with tcgen05.st and the saved activations straight to
// HBM in the image layout, so shared memory holds nothing per tile and BOTH hidden weight sets stay
// resident (2 x 64 KB, one TMA load per CTA): the two tile groups no longer share a refilled buffer,
// run independently and settle in anti-phase -- one group's MMAs under the other's epilogue.
// =================================================================================================
struct FwdTc2Smem {
    unsigned char W[2][NDP_SET128];        // hidden weight hi/lo images, resident (operand B, K-major)
    unsigned char WIN[2 * NDP_IMG16];      // [128 outputs][16]: cols 0..5 = W_in, rest 0 (operand B of the input layer)
    unsigned char HW[2 * NDP_HWIMG];       // [16 head rows][128]: head weights, rows >= head_dim 0 (operand B of the heads)
    float xs[2][NDP_TP * 4];
    float hb[16];
    float bias[3][NDP_W];                  // b_in, b_0, b_1: every thread needs 64 of them per epilogue (broadcast LDS instead of LDG + address arithmetic)
    NdpMbar bar_w[2], bar_mma[2];          // one arrival barrier per resident weight set (the first is needed first)
    int pipe_lock;                         // the group that holds it issues its MMA batch alone (see ndp_pipe_acquire)
    unsigned tmem_slot, pad[2];
};
// The two groups start together; if both issued their batches at the same time the tensor pipe would
// interleave them and both would finish late and together (lock-step for the whole kernel).  Issuing one
// batch at a time makes the first group's accumulator ready half a batch earlier, which puts the groups
// in anti-phase: one group's MMAs run under the other group's epilogue.
__device__ __forceinline__ void ndp_pipe_acquire(int* lock) { while (atomicCAS(lock, 0, 1) != 0) {} }
__device__ __forceinline__ void ndp_pipe_release(int* lock) { __threadfence_block(); atomicExch(lock, 0); }
size_t ndp_fwd_tc2_smem_bytes() { return sizeof(FwdTc2Smem) + 1024; }
#define TM2_ACC 0u
#define TM2_AOP 128u
#define TM2_LO 64u

template <bool SAVE>       // SAVE: the activations are saved as images (a.act), from which the non-recomputing backward reads relu'
__global__ void __launch_bounds__(NDP_FWD_TC_THREADS, 1) ndp_warp_fwd_tc2_kernel(NdpFwdArgs a) {
    NDP_DYN_SMEM(smem_raw);
    FwdTc2Smem& S = *(FwdTc2Smem*)NDP_SMEM_ALIGN(smem_raw, 1024);

    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0;
    const int g = tid >> 8, gt = tid & (NDP_GROUP - 1);          // tile group, thread within the group
    const int rounds = a.rounds;                                 // CTA = tiles [2 rounds bx, 2 rounds (bx + 1)): round r, group g -> tile
    const int tile_first = blockIdx.x * 2 * rounds;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile_first * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const unsigned char* wimg = (const unsigned char*)(a.pack + (long long)pair * a.pack_stride + L.pack_img);
    const int LH = L.hidden, HD = L.head_dim;
    const int warp = tid >> 5, p = gt & (NDP_TP - 1), half = gt >> 7;
    const bool ldw = (ndp_warp_uniform(warp) & 7) == 0;   // the group's issuing warp: one elected lane launches the MMAs
    const int RS = NDP_IMG_RS(128), CS = NDP_IMG_CS, RS16 = NDP_IMG_RS(16);
    unsigned char* const gact_pair = (SAVE && a.act) ? (unsigned char*)a.act + ((long long)pair * a.act_stride) * 4 : nullptr;
    float* xs = S.xs[g];

    NDP_T(0);
#ifndef NDP_EMU
    unsigned long long t_cta0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cta0));
#endif
    if (warp == 0) ndp_tmem_alloc_warp(&S.tmem_slot, 512);
    if (tid == 0) {
        ndp_mbar_init(&S.bar_w[0], 1); ndp_mbar_init(&S.bar_w[1], 1); ndp_mbar_init(&S.bar_mma[0], 1); ndp_mbar_init(&S.bar_mma[1], 1);
        S.pipe_lock = 0;
        // the first hidden weight set now; the second one after the input layer's MMAs (every CTA of the wave
        // starts at the same time: asking for 128 KB at once doubles the burst the first layer waits behind)
        if (LH > 0) ndp_stage_bulk(S.W[0], wimg, NDP_SET128, &S.bar_w[0]);
        NDP_T(2);
    }
    if (tid < 16) S.hb[tid] = (tid < HD) ? __ldg(params + L.head_b[tid]) : 0.0f;
    if (tid < 3 * NDP_W) {
        const int l = tid >> 7, o = tid & (NDP_W - 1);
        S.bias[l][o] = (l == 0) ? __ldg(params + L.off_b_in + o) : (l - 1 < LH ? __ldg(params + L.off_b[l - 1] + o) : 0.0f);
    }
    if (tid < 256) {            // input-layer weight image: row o = tid / 2, 8-column chunk tid & 1
        const int o = tid >> 1, c8 = tid & 1;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c8 == 0 && j < 6) ? __ldg(params + L.off_w_in + o * 6 + j) : 0.0f;
        ndp_store_chunk2(S.WIN, NDP_IMG16, ndp_img_off(o, c8 * 8, RS16), v);
    } else {                    // head weight image: row r = (tid - 256) / 16, 8-column chunk (tid - 256) & 15
        const int r = (tid - 256) >> 4, c8 = (tid - 256) & 15;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (r < HD) ? __ldg(params + L.head_w[r] + c8 * 8 + j) : 0.0f;   // head rows are not 16-byte aligned
        ndp_store_chunk2(S.HW, NDP_HWIMG, ndp_img_off(r, c8 * 8, RS), v);
    }
    // this point's positional encoding (nets.py:164-177) as packed fp16 hi / lo pairs -> the first 8 columns
    // (K = 16) of the group's TMEM operand images; also the point itself for the warp composition
    auto stage_points = [&](int tile, unsigned tl) {
        if (gt < NDP_TP) {
            unsigned ehi[8], elo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { ehi[j] = 0u; elo[j] = 0u; }
            const int gp = tile * NDP_TP + gt;
            float px = 0.0f, py = 0.0f, pz = 0.0f;
            if (gp < n) {
                const float* xp = a.x + (long long)pair * a.x_stride + (long long)gp * 3;
                px = __ldg(xp); py = __ldg(xp + 1); pz = __ldg(xp + 2);
            }
            xs[gt * 4 + 0] = px; xs[gt * 4 + 1] = py; xs[gt * 4 + 2] = pz;
            float sn, cs;
            sincosf(px * L.freq, &sn, &cs); ndp_split2_pair(sn, cs, ehi[0], elo[0]);
            sincosf(py * L.freq, &sn, &cs); ndp_split2_pair(sn, cs, ehi[1], elo[1]);
            sincosf(pz * L.freq, &sn, &cs); ndp_split2_pair(sn, cs, ehi[2], elo[2]);
            ndp_tmem_st<8>(tl + TM2_AOP, ehi);
            ndp_tmem_st<8>(tl + TM2_AOP + TM2_LO, elo);
            ndp_tmem_wait_st();
        }
    };
    ndp_tc_fence_before();
    ndp_fence_proxy_async();
    __syncthreads();
    ndp_tc_fence_after();
    NDP_T(1);
    const unsigned tmem = S.tmem_slot + (unsigned)g * 256u;
    const unsigned tlane = tmem + ((unsigned)((warp & 3) * 32) << 16);

    unsigned mph = 0;
    for (int r = 0; r < rounds; ++r) {
        const int tile = tile_first + 2 * r + g;
        if (tile * NDP_TP >= n) break;                  // group-uniform: this group has no further tile
        unsigned char* gact = gact_pair ? gact_pair + (long long)tile * (LH + 1) * NDP_SET128 : nullptr;
        stage_points(tile, tlane);
        ndp_tc_fence_before();
        ndp_group_sync(1 + g, NDP_GROUP);
        // ---- stage s = 0: input layer; s = 1..LH: hidden layer s - 1.  MMAs by the group's elected thread,
        //      then the epilogue h = relu(acc + bias) re-split into the TMEM operand images (and saved to HBM).
        for (int s = 0; s <= LH; ++s) {
            if (ldw && ndp_elect_one()) {
                if (s > 0 && r == 0) ndp_mbar_wait(&S.bar_w[s - 1], 0u);  // this layer's weights have landed (once per CTA)
                NDP_T(40 + 2 * s);
                ndp_pipe_acquire(&S.pipe_lock);
                NDP_T(41 + 2 * s);
                ndp_tc_fence_after();
                if (s == 0) {
                    ndp_umma_gemm3_ta(tmem + TM2_ACC, tmem + TM2_AOP, TM2_LO, ndp_umma_desc(S.WIN, CS, RS16), NDP_IMG16, 0, 1,
                                      ndp_idesc_f16(128, 128, 0, 0));
                } else {
                    ndp_umma_gemm3_ta(tmem + TM2_ACC, tmem + TM2_AOP, TM2_LO, ndp_umma_desc(S.W[s - 1], CS, RS), NDP_IMG128, 2 * CS, 8,
                                      ndp_idesc_f16(128, 128, 0, 0));
                }
                ndp_umma_commit(&S.bar_mma[g]);
                ndp_pipe_release(&S.pipe_lock);
                if (s == 0 && r == 0 && g == 0 && LH > 1) ndp_stage_bulk(S.W[1], wimg + NDP_SET128, NDP_SET128, &S.bar_w[1]);
                NDP_T(8 + 4 * s);
            }
            if (ldw) ndp_mbar_wait(&S.bar_mma[g], mph);      // one warp polls, the other seven sleep at the named barrier
            mph ^= 1;
            ndp_group_sync(1 + g, NDP_GROUP);
            ndp_tc_fence_after();
            NDP_T(9 + 4 * s);
            const float* bias = S.bias[s];
            unsigned char* gimg = gact ? gact + (long long)s * NDP_SET128 : nullptr;
#pragma unroll 1
            for (int c32 = 0; c32 < 2; ++c32) {
                float v[32];
                unsigned hi[16], lo[16];
                const int col0 = half * 64 + c32 * 32;
                ndp_tmem_ld32(tlane + TM2_ACC + col0, v);
#pragma unroll
                for (int s8 = 0; s8 < 4; ++s8) {
                    const float4 b0 = *(const float4*)(bias + col0 + s8 * 8), b1 = *(const float4*)(bias + col0 + s8 * 8 + 4);
                    if (!SAVE) {         // nothing reads relu' back from an image (the recomputing backward): plain max, packed adds
                        ndp_bias_relu_split2(v[s8 * 8 + 0], v[s8 * 8 + 1], b0.x, b0.y, hi[s8 * 4 + 0], lo[s8 * 4 + 0]);
                        ndp_bias_relu_split2(v[s8 * 8 + 2], v[s8 * 8 + 3], b0.z, b0.w, hi[s8 * 4 + 1], lo[s8 * 4 + 1]);
                        ndp_bias_relu_split2(v[s8 * 8 + 4], v[s8 * 8 + 5], b1.x, b1.y, hi[s8 * 4 + 2], lo[s8 * 4 + 2]);
                        ndp_bias_relu_split2(v[s8 * 8 + 6], v[s8 * 8 + 7], b1.z, b1.w, hi[s8 * 4 + 3], lo[s8 * 4 + 3]);
                        continue;
                    }
                    ndp_split2_pair(ndp_relu_img(v[s8 * 8 + 0] + b0.x), ndp_relu_img(v[s8 * 8 + 1] + b0.y), hi[s8 * 4 + 0], lo[s8 * 4 + 0]);
                    ndp_split2_pair(ndp_relu_img(v[s8 * 8 + 2] + b0.z), ndp_relu_img(v[s8 * 8 + 3] + b0.w), hi[s8 * 4 + 1], lo[s8 * 4 + 1]);
                    ndp_split2_pair(ndp_relu_img(v[s8 * 8 + 4] + b1.x), ndp_relu_img(v[s8 * 8 + 5] + b1.y), hi[s8 * 4 + 2], lo[s8 * 4 + 2]);
                    ndp_split2_pair(ndp_relu_img(v[s8 * 8 + 6] + b1.z), ndp_relu_img(v[s8 * 8 + 7] + b1.w), hi[s8 * 4 + 3], lo[s8 * 4 + 3]);
                    if (gimg) {          // saved activation h_s in the image layout the backward kernel loads by TMA
                        const unsigned off = ndp_img_off(p, col0 + s8 * 8, RS);
                        *(uint4*)(gimg + off) = make_uint4(hi[s8 * 4 + 0], hi[s8 * 4 + 1], hi[s8 * 4 + 2], hi[s8 * 4 + 3]);
                        *(uint4*)(gimg + NDP_IMG128 + off) = make_uint4(lo[s8 * 4 + 0], lo[s8 * 4 + 1], lo[s8 * 4 + 2], lo[s8 * 4 + 3]);
                    }
                }
                ndp_tmem_st<16>(tlane + TM2_AOP + (unsigned)(col0 >> 1), hi);
                ndp_tmem_st<16>(tlane + TM2_AOP + TM2_LO + (unsigned)(col0 >> 1), lo);
            }
            ndp_tmem_wait_st();
            ndp_tc_fence_before();
            ndp_group_sync(1 + g, NDP_GROUP);
            NDP_T(11 + 4 * s);
        }

        // ---- heads on the tensor core: z_raw[128 points][16] = h_L W_h^T
        if (ldw && ndp_elect_one()) {
            ndp_tc_fence_after();
            ndp_umma_gemm3_ta(tmem + TM2_ACC, tmem + TM2_AOP, TM2_LO, ndp_umma_desc(S.HW, CS, RS), NDP_HWIMG, 2 * CS, 8,
                              ndp_idesc_f16(128, 16, 0, 0));
            ndp_umma_commit(&S.bar_mma[g]);
        }
        if (ldw) ndp_mbar_wait(&S.bar_mma[g], mph);      // one warp polls, the other seven sleep at the named barrier
        mph ^= 1;
        ndp_group_sync(1 + g, NDP_GROUP);
        ndp_tc_fence_after();
        NDP_T(60);
        ndp_fwd_point_tail(a, L, pair, tile, n, gt, HD, tlane + TM2_ACC, xs, S.hb);
        NDP_T(61);
        ndp_tc_fence_before();
        ndp_group_sync(1 + g, NDP_GROUP);               // head results read, xs consumed: the next round may overwrite them
        ndp_tc_fence_after();
    }
    if (tid == 0) for (int l = 0; l < LH; ++l) ndp_mbar_wait(&S.bar_w[l], 0u);   // the weight copies must not outlive the CTA's shared memory
    ndp_tc_fence_before();
    __syncthreads();
    NDP_T(63);
#ifndef NDP_EMU
    if (tid == 0 && blockIdx.x == 0) {   // self-contained duration of this CTA (valid with several launches in flight)
        unsigned long long t1_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_));
        ndp_dbg_fwd[62] = t1_ - t_cta0;
    }
    if (tid == 0) {
        unsigned long long t1_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_));
        atomicAdd(&ndp_dbg_fwd[58], t1_ - t_cta0); atomicAdd(&ndp_dbg_fwd[59], 1ull);
    }
#endif
    if (warp == 0) ndp_tmem_dealloc(S.tmem_slot, 512);
}

void ndp_launch_fwd_tc(const NdpFwdArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    const int tiles = (a.n + NDP_TP - 1) / NDP_TP;
    int rounds = a.rounds > 0 ? a.rounds : 1;    // measured: more rounds only help at large batches and cost at small ones
    if (a.tc_version != 1 && a.lay.hidden <= 2) {      // A operand in TMEM, both hidden weight sets resident
        if (rounds > 8) rounds = 8;
        NdpFwdArgs b2 = a;
        b2.rounds = rounds;
        if (a.act) NDP_LAUNCH_PRIO(0, ndp_warp_fwd_tc2_kernel<true>, dim3((tiles + 2 * rounds - 1) / (2 * rounds), a.npairs), dim3(NDP_FWD_TC_THREADS),
                                   ndp_fwd_tc2_smem_bytes(), s, b2);
        else NDP_LAUNCH_PRIO(0, ndp_warp_fwd_tc2_kernel<false>, dim3((tiles + 2 * rounds - 1) / (2 * rounds), a.npairs), dim3(NDP_FWD_TC_THREADS),
                             ndp_fwd_tc2_smem_bytes(), s, b2);
        return;
    }
    if (rounds > NDP_FWD_MAX_ROUNDS) rounds = NDP_FWD_MAX_ROUNDS;
    NdpFwdArgs b = a;
    b.rounds = rounds;
    dim3 grid((tiles + 2 * rounds - 1) / (2 * rounds), a.npairs);
    NDP_LAUNCH(ndp_warp_fwd_tc_kernel, grid, dim3(NDP_FWD_TC_THREADS), ndp_fwd_tc_smem_bytes(), s, b);
}

int ndp_fwd_tc_init() {
    int e = (int)cudaFuncSetAttribute(ndp_warp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)ndp_fwd_tc_smem_bytes());
    if (e == 0) e = (int)cudaFuncSetAttribute(ndp_warp_fwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)ndp_fwd_tc2_smem_bytes());
    if (e == 0) e = (int)cudaFuncSetAttribute(ndp_warp_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)ndp_fwd_tc2_smem_bytes());
    return e;
}

#ifndef NDP_EMU
int ndp_debug_copy_fwd(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, ndp_dbg_fwd, sizeof(unsigned long long) * 64); }
#else
int ndp_debug_copy_fwd(unsigned long long* out) { for (int i = 0; i < 64; ++i) out[i] = 0; return 0; }
#endif
