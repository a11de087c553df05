// Kernel (1), tensor-core version: one pyramid level of the warp field over a tile of 128 points.
//
// Reference: model/nets.py:111-140 (NDPLayer.forward), :164-177 (posenc), :295-304 (MLP),
//            :144-161 (get_Rotation), model/rigid_body.py.
//
// The hidden 128x128 layers (the only GEMM-shaped work: M = 128 points fills one UMMA tile exactly)
// run on the 5th-generation tensor cores: tcgen05.mma issued by ONE thread, operands in shared
// memory as fp16 hi/lo image sets (ndp_tc.cuh: 2-way fp16 split, three partial products, fp32
// accumulation in TMEM => fp32-level accuracy), accumulator [128 lanes x 128 columns] in TMEM, read
// back with tcgen05.ld by 8 warps (warp w: lanes 32(w%4).., column half w/4) for the
// bias + ReLU + re-split epilogue that writes the next layer's A operand in place.
// Weight image sets (64 KB/layer, maintained by the Adam kernel) arrive by TMA bulk copies
// (cp.async.bulk + mbarrier) issued as soon as the previous layer's MMAs have retired; the saved
// activations leave by TMA bulk stores straight from the operand images.  Input layer (K = 6), heads
// (K = 128, N <= 11) and the per-point rotation / warp composition stay on the FP32 pipes.
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

// optional phase timestamps of CTA (0,0) (debug aid, read back through ndp_debug_phase_times)
#ifndef NDP_EMU
__device__ unsigned long long ndp_dbg_fwd[64];
#define NDP_T(i) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ndp_dbg_fwd[i] = t_; } } while (0)
#else
#define NDP_T(i) do {} while (0)
#endif

struct FwdTcSmem {
    unsigned char A[NDP_SET128];              // activation hi/lo images (operand A, K-major)
    unsigned char B[NDP_SET128];              // weight hi/lo images of the current layer (operand B, K-major)
    float win[NDP_W * 6];
    float bin[NDP_W];
    float hw[NDP_MAX_HEAD * NDP_W];
    float hb[16];
    float es[NDP_TP * 8];
    float xs[NDP_TP * 4];
    float zpart[2 * NDP_TP * NDP_ZPITCH];
    NdpMbar bar_w, bar_mma;
    unsigned tmem_slot, pad[3];
};
size_t ndp_fwd_tc_smem_bytes() { return sizeof(FwdTcSmem) + 1024; }

__device__ __forceinline__ void ndp_head_accum(float (&hacc)[NDP_MAX_HEAD], const float* hw, int HD, const float (&v)[8], int col0) {
#pragma unroll
    for (int r = 0; r < NDP_MAX_HEAD; ++r) {
        if (r < HD) {
            const float4 w0 = *(const float4*)(hw + r * NDP_W + col0), w1 = *(const float4*)(hw + r * NDP_W + col0 + 4);
            float s = hacc[r];
            s = fmaf(v[0], w0.x, s); s = fmaf(v[1], w0.y, s); s = fmaf(v[2], w0.z, s); s = fmaf(v[3], w0.w, s);
            s = fmaf(v[4], w1.x, s); s = fmaf(v[5], w1.y, s); s = fmaf(v[6], w1.z, s); s = fmaf(v[7], w1.w, s);
            hacc[r] = s;
        }
    }
}

__global__ void __launch_bounds__(NDP_THREADS, 1) ndp_warp_fwd_tc_kernel(NdpFwdArgs a) {
    NDP_DYN_SMEM(smem_raw);
    FwdTcSmem& S = *(FwdTcSmem*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    const int tid = threadIdx.x, pair = blockIdx.y, tile = blockIdx.x;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const unsigned char* wimg = (const unsigned char*)(a.pack + (long long)pair * a.pack_stride + L.pack_img);
    const int LH = L.hidden, HD = L.head_dim;
    const int warp = tid >> 5, p = tid & (NDP_TP - 1), half = tid >> 7;
    unsigned char* gact = a.act ? (unsigned char*)a.act + ((long long)pair * a.act_stride) * 4 +
                                      (long long)tile * (LH + 1) * NDP_SET128 : nullptr;

    NDP_T(0);
    if (warp == 0) ndp_tmem_alloc_warp(&S.tmem_slot, 128);
    if (tid == 0) { ndp_mbar_init(&S.bar_w, 1); ndp_mbar_init(&S.bar_mma, 1); }
    // stage the small fp32 operands
    for (int i = tid; i < NDP_W * 6; i += NDP_THREADS) S.win[i] = __ldg(params + L.off_w_in + i);
    if (tid < NDP_W) S.bin[tid] = __ldg(params + L.off_b_in + tid);
    for (int i = tid; i < HD * NDP_W; i += NDP_THREADS) S.hw[i] = __ldg(params + L.head_w[i >> 7] + (i & 127));
    if (tid < HD) S.hb[tid] = __ldg(params + L.head_b[tid]);
    if (tid < NDP_TP) {        // points + positional encoding (nets.py:164-177)
        const int gp = tile * NDP_TP + tid;
        float px = 0.0f, py = 0.0f, pz = 0.0f;
        if (gp < n) {
            const float* xp = a.x + (long long)pair * a.x_stride + (long long)gp * 3;
            px = __ldg(xp); py = __ldg(xp + 1); pz = __ldg(xp + 2);
        }
        S.xs[tid * 4 + 0] = px; S.xs[tid * 4 + 1] = py; S.xs[tid * 4 + 2] = pz;
        float s, c;
        float* e = S.es + tid * 8;
        sincosf(px * L.freq, &s, &c); e[0] = s; e[1] = c;
        sincosf(py * L.freq, &s, &c); e[2] = s; e[3] = c;
        sincosf(pz * L.freq, &s, &c); e[4] = s; e[5] = c;
    }
    ndp_tc_fence_before();
    __syncthreads();
    ndp_tc_fence_after();
    NDP_T(1);
    const unsigned tmem = S.tmem_slot;
    if (tid == 0 && LH > 0) ndp_stage_bulk(S.B, wimg, NDP_SET128, &S.bar_w);

    float hacc[NDP_MAX_HEAD];
#pragma unroll
    for (int r = 0; r < NDP_MAX_HEAD; ++r) hacc[r] = 0.0f;

    // ---- input layer (K = 6) on the FP32 pipes: h0 = relu(W_in e + b_in)  (nets.py:75,114)
    {
        float e[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) e[c] = S.es[p * 8 + c];
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
            const int o0 = half * 64 + ch * 8;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* w = S.win + (o0 + j) * 6;
                float s = S.bin[o0 + j];
#pragma unroll
                for (int c = 0; c < 6; ++c) s = fmaf(e[c], w[c], s);
                v[j] = ndp_relu_img(s);
            }
            ndp_store_chunk2(S.A, NDP_IMG128, ndp_img_off(p, o0, NDP_IMG_RS(128)), v);
            if (LH == 0) ndp_head_accum(hacc, S.hw, HD, v, o0);
        }
    }
    ndp_fence_proxy_async();
    __syncthreads();
    NDP_T(2);

    const unsigned idesc = ndp_idesc_f16(128, 128, 0, 0);
    for (int l = 0; l < LH; ++l) {
        // ---- h_{l+1} = relu(W_l h_l + b_l): three fp16 partial products into TMEM, one issuing thread
        if (tid == 0) {
            if (gact) {      // save h_l for the backward pass straight from the operand image
                for (int i = 0; i < 2; ++i) ndp_bulk_s2g(gact + (long long)l * NDP_SET128 + i * NDP_IMG128, S.A + i * NDP_IMG128, NDP_IMG128);
                ndp_bulk_commit();
            }
            ndp_mbar_wait(&S.bar_w, (unsigned)(l & 1));
            NDP_T(8 + 4 * l);
            ndp_tc_fence_after();
            ndp_umma_gemm3(tmem, ndp_umma_desc(S.A, NDP_IMG_CS, NDP_IMG_RS(128)), NDP_IMG128, 2 * NDP_IMG_CS,
                           ndp_umma_desc(S.B, NDP_IMG_CS, NDP_IMG_RS(128)), NDP_IMG128, 2 * NDP_IMG_CS, 8, idesc, false);
            ndp_umma_commit(&S.bar_mma);
        }
        ndp_mbar_wait(&S.bar_mma, (unsigned)(l & 1));
        ndp_tc_fence_after();
        NDP_T(9 + 4 * l);
        if (tid == 0) {
            if (gact) ndp_bulk_wait_read0();                      // the store has finished reading A
            if (l + 1 < LH) ndp_stage_bulk(S.B, wimg + (long long)(l + 1) * NDP_SET128, NDP_SET128, &S.bar_w);
        }
        __syncthreads();
        NDP_T(10 + 4 * l);
        // ---- epilogue: TMEM -> registers, bias + ReLU, re-split into the A images (in place)
        const float* bias = params + L.off_b[l];
#pragma unroll 1
        for (int c32 = 0; c32 < 2; ++c32) {
            float v[32];
            const int col0 = half * 64 + c32 * 32;
            ndp_tmem_ld32(tmem + ((unsigned)((warp & 3) * 32) << 16) + col0, v);
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
                float u[8];
                const float4 b0 = __ldg((const float4*)(bias + col0 + s8 * 8)), b1 = __ldg((const float4*)(bias + col0 + s8 * 8 + 4));
                u[0] = ndp_relu_img(v[s8 * 8 + 0] + b0.x); u[1] = ndp_relu_img(v[s8 * 8 + 1] + b0.y);
                u[2] = ndp_relu_img(v[s8 * 8 + 2] + b0.z); u[3] = ndp_relu_img(v[s8 * 8 + 3] + b0.w);
                u[4] = ndp_relu_img(v[s8 * 8 + 4] + b1.x); u[5] = ndp_relu_img(v[s8 * 8 + 5] + b1.y);
                u[6] = ndp_relu_img(v[s8 * 8 + 6] + b1.z); u[7] = ndp_relu_img(v[s8 * 8 + 7] + b1.w);
                ndp_store_chunk2(S.A, NDP_IMG128, ndp_img_off(p, col0 + s8 * 8, NDP_IMG_RS(128)), u);
                if (l == LH - 1) ndp_head_accum(hacc, S.hw, HD, u, col0 + s8 * 8);
            }
        }
        ndp_tc_fence_before();
        ndp_fence_proxy_async();
        __syncthreads();
        NDP_T(11 + 4 * l);
    }
    if (tid == 0 && gact) {      // top activation
        for (int i = 0; i < 2; ++i) ndp_bulk_s2g(gact + (long long)LH * NDP_SET128 + i * NDP_IMG128, S.A + i * NDP_IMG128, NDP_IMG128);
        ndp_bulk_commit();
    }

    // ---- heads: z = mlp_scale * (W_h h + b_h), the two column halves combined through smem
#pragma unroll
    for (int r = 0; r < NDP_MAX_HEAD; ++r) S.zpart[(half * NDP_TP + p) * NDP_ZPITCH + r] = hacc[r];
    __syncthreads();

    // ---- per-point rotation + warp composition (nets.py:119-137)
    if (tid < NDP_TP) {
        const int gp = tile * NDP_TP + tid;
        const float INF = __int_as_float(0x7f800000);
        float y[3] = {INF, INF, INF};
        if (gp < n) {
            float z[NDP_MAX_HEAD], nu = 0.0f;
#pragma unroll
            for (int r = 0; r < NDP_MAX_HEAD; ++r)
                z[r] = (r < HD) ? L.mu * (S.zpart[tid * NDP_ZPITCH + r] + S.zpart[(NDP_TP + tid) * NDP_ZPITCH + r] + S.hb[r]) : 0.0f;
            ndp_point_forward(L.motion, L.rot, L.nonrigid, z, S.xs + tid * 4, y, &nu);
            if (a.y_add) {
                const float* ya = a.y_add + (long long)pair * a.y_add_stride;
                y[0] += ya[0]; y[1] += ya[1]; y[2] += ya[2];
            }
            float* yp = a.y + (long long)pair * a.y_stride + (long long)gp * 3;
            yp[0] = y[0]; yp[1] = y[1]; yp[2] = y[2];
            if (a.nu && L.nonrigid) a.nu[(long long)pair * a.nu_stride + gp] = nu;
            if (a.zsave) {
                float* zp = a.zsave + (long long)pair * a.z_stride + (long long)gp * NDP_ZPITCH;
#pragma unroll
                for (int r = 0; r < NDP_ZPITCH; ++r) zp[r] = z[r];
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
            if ((tid & 31) == 0 && gp < n) {
                float* bx = a.ybox + ((long long)pair * a.box_stride + (gp >> 5)) * 8;
                bx[0] = l0; bx[1] = l1; bx[2] = l2; bx[3] = 0.0f; bx[4] = h0; bx[5] = h1; bx[6] = h2; bx[7] = 0.0f;
            }
        }
    }
    NDP_T(61);
    if (tid == 0 && gact) ndp_bulk_wait0();     // smem must outlive the bulk stores
    NDP_T(62);
    ndp_tc_fence_before();
    __syncthreads();
    NDP_T(63);
    if (warp == 0) ndp_tmem_dealloc(tmem, 128);
}

void ndp_launch_fwd_tc(const NdpFwdArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    dim3 grid((a.n + NDP_TP - 1) / NDP_TP, a.npairs);
    NDP_LAUNCH(ndp_warp_fwd_tc_kernel, grid, dim3(NDP_THREADS), ndp_fwd_tc_smem_bytes(), s, a);
}

int ndp_fwd_tc_init() {
    return (int)cudaFuncSetAttribute(ndp_warp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ndp_fwd_tc_smem_bytes());
}

#ifndef NDP_EMU
int ndp_debug_copy_fwd(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, ndp_dbg_fwd, sizeof(unsigned long long) * 64); }
#else
int ndp_debug_copy_fwd(unsigned long long* out) { for (int i = 0; i < 64; ++i) out[i] = 0; return 0; }
#endif
