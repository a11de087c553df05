// Kernel (3a), tensor-core version: backward of kernel (1) over a tile of 128 points -> this
// tile's partial parameter gradients (reduced in fixed order by kernel (3b)).
//
// Replaces the autograd backward of NDPLayer.forward (model/nets.py:111-140) that the reference
// runs in loss.backward() (model/registration.py:236).
//
// Every reduction over the tile's 128 points is a tcgen05 GEMM with K = points (operands as fp16
// hi/lo image sets, fp32 accumulation in TMEM, see ndp_tc.cuh), issued by one thread:
//     head grads   dW_h  = hg^T h_L          A = hg image   (MN-major)  B = h_L   (MN-major)
//     weight grads dW_l  = delta^T h_l       A = delta      (MN-major)  B = h_l   (MN-major)
//     bias grads   db_l  = delta^T 1         A = delta      (MN-major)  B = E[:,6] = 1
//     input layer  dW_in = delta_0^T e       A = delta_0    (MN-major)  B = E[:,0:6] = posenc
//     back-prop    delta_l = (delta W_l).relu'   A = delta  (K-major)   B = W_l   (MN-major)
// The SAME delta image is the MN-major operand of the dW products and the K-major operand of the
// back-propagation product, and the SAME weight image serves forward and backward (the core-matrix
// layout is both canonical UMMA layouts at once), so nothing is ever transposed.  h_l image sets
// come back from HBM by TMA bulk copies exactly as the forward kernel's bulk stores wrote them;
// the weight image of the layer replaces h_l in shared memory as soon as the dW MMAs retire,
// overlapping with the TMEM -> HBM epilogue of dW.  No atomics: one partial row per tile.
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

// optional phase timestamps of CTA (0,0) (debug aid, read back through ndp_debug_phase_times)
#ifndef NDP_EMU
__device__ unsigned long long ndp_dbg_bwd[64];
#define NDP_T(i) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ndp_dbg_bwd[i] = t_; } } while (0)
#else
#define NDP_T(i) do {} while (0)
#endif

#define NDP_IMG16 NDP_IMG_BYTES(16)     // [128][16] fp16 image: 4096 bytes
struct BwdTcSmem {
    unsigned char D[NDP_SET128];        // delta hi/lo images (scaled by the tile's power of two)
    unsigned char X[NDP_SET128];        // h_l images, then W_l images
    unsigned char HG[2 * NDP_IMG16];    // [128 points][16]: mlp_scale * dL/dz (head gradients); must precede E:
    unsigned char E[2 * NDP_IMG16];     // [128 points][16]: cols 0..5 positional encoding, col 6 = 1
    float hw[NDP_MAX_HEAD * NDP_W];     // head weights (fp32), rows >= head_dim zero
    float xs[NDP_TP * 4];
    float gxs[NDP_TP * 4];
    float red[4];                       // per-warp max |head gradient| of the tile
    NdpMbar bar_x, bar_mma;
    unsigned tmem_slot, pad[3];
};
size_t ndp_bwd_tc_smem_bytes() { return sizeof(BwdTcSmem) + 128; }

#define TM_DW 0u
#define TM_DH 128u
#define TM_SM 256u

__global__ void __launch_bounds__(NDP_THREADS, 1) ndp_warp_bwd_tc_kernel(NdpBwdArgs a) {
    NDP_DYN_SMEM(smem_raw);
    BwdTcSmem& S = *(BwdTcSmem*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);

    const int tid = threadIdx.x, pair = blockIdx.y, tile = blockIdx.x;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const unsigned char* wimg = (const unsigned char*)(a.pack + (long long)pair * a.pack_stride + L.pack_img);
    const int LH = L.hidden, HD = L.head_dim;
    const unsigned char* gact = (const unsigned char*)a.act + ((long long)pair * a.act_stride) * 4 +
                                (long long)tile * (LH + 1) * NDP_SET128;
    float* part = a.partials + (long long)pair * a.partials_stride + (long long)tile * a.partial_pitch;
    const int warp = tid >> 5, lane = tid & 31, p = tid & (NDP_TP - 1), half = tid >> 7;
    const int RS = NDP_IMG_RS(128), CS = NDP_IMG_CS, RS16 = NDP_IMG_RS(16);
    unsigned xph = 0, mph = 0;
    NDP_T(0);

    if (warp == 0) ndp_tmem_alloc_warp(&S.tmem_slot, 512);
    if (tid == 0) { ndp_mbar_init(&S.bar_x, 1); ndp_mbar_init(&S.bar_mma, 1); }
    for (int i = tid; i < NDP_MAX_HEAD * NDP_W; i += NDP_THREADS)
        S.hw[i] = ((i >> 7) < HD) ? __ldg(params + L.head_w[i >> 7] + (i & 127)) : 0.0f;
    ndp_tc_fence_before();
    __syncthreads();
    ndp_tc_fence_after();
    NDP_T(1);
    const unsigned tmem = S.tmem_slot;
    const unsigned tlane = tmem + ((unsigned)((warp & 3) * 32) << 16);
    if (tid == 0) ndp_stage_bulk(S.X, gact + (long long)LH * NDP_SET128, NDP_SET128, &S.bar_x);      // h_L

    // ---- per point: dL/dy -> dL/dz (heads) and the direct part of dL/dx; E and head-gradient images
    float hgv[16], e0[8];
#pragma unroll
    for (int r = 0; r < 16; ++r) hgv[r] = 0.0f;
    if (tid < NDP_TP) {
        const int gp = tile * NDP_TP + tid;
        float gz[NDP_MAX_HEAD];
#pragma unroll
        for (int r = 0; r < NDP_MAX_HEAD; ++r) gz[r] = 0.0f;
        float x[3] = {0.0f, 0.0f, 0.0f}, gxd[3] = {0.0f, 0.0f, 0.0f};
        if (gp < n) {
            const float* xp = a.x + (long long)pair * a.x_stride + (long long)gp * 3;
            x[0] = __ldg(xp); x[1] = __ldg(xp + 1); x[2] = __ldg(xp + 2);
            const float* gp_ = a.gy + (long long)pair * a.gy_stride + (long long)gp * 3;
            float gy[3] = {gp_[0], gp_[1], gp_[2]};
            if (a.gacc) {   // scattered Chamfer term, 2^-40 fixed point (order independent)
                unsigned long long* ga = a.gacc + (long long)pair * a.gacc_stride + (long long)gp * 3;
                const int m = a.mcounts ? a.mcounts[pair] : a.m;
                const double sc = 9.094947017729282e-13 / (double)m;
                const long long a0 = (long long)ga[0], a1 = (long long)ga[1], a2 = (long long)ga[2];
                const bool poison = (a0 >= (1LL << 60)) || (a0 <= -(1LL << 60));
                gy[0] += poison ? __int_as_float(0x7fc00000) : (float)((double)a0 * sc);
                gy[1] += (float)((double)a1 * sc);
                gy[2] += (float)((double)a2 * sc);
                ga[0] = 0ull; ga[1] = 0ull; ga[2] = 0ull;
            }
            float z[NDP_MAX_HEAD];
            const float4* zp = (const float4*)(a.zsave + (long long)pair * a.z_stride + (long long)gp * NDP_ZPITCH);
            const float4 z0 = zp[0], z1 = zp[1], z2 = zp[2];
            z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
            z[8] = z2.x; z[9] = z2.y; z[10] = z2.z; z[11] = z2.w;
            const float gnu = (a.gnu && L.nonrigid) ? a.gnu[(long long)pair * a.gnu_stride + gp] : 0.0f;
            ndp_point_backward(L.motion, L.rot, L.nonrigid, z, x, gy, gnu, gz, gxd);
        }
        float mx = 0.0f;
#pragma unroll
        for (int r = 0; r < NDP_MAX_HEAD; ++r) { hgv[r] = L.mu * gz[r]; mx = fmaxf(mx, fabsf(hgv[r])); }
        for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        if (lane == 0) S.red[warp] = mx;
        S.xs[tid * 4 + 0] = x[0]; S.xs[tid * 4 + 1] = x[1]; S.xs[tid * 4 + 2] = x[2];
        S.gxs[tid * 4 + 0] = gxd[0]; S.gxs[tid * 4 + 1] = gxd[1]; S.gxs[tid * 4 + 2] = gxd[2];
        float s, c;
        sincosf(x[0] * L.freq, &s, &c); e0[0] = s; e0[1] = c;
        sincosf(x[1] * L.freq, &s, &c); e0[2] = s; e0[3] = c;
        sincosf(x[2] * L.freq, &s, &c); e0[4] = s; e0[5] = c;
        e0[6] = 1.0f; e0[7] = 0.0f;
    }
    __syncthreads();
    // the tile's delta scale: an exact power of two that brings the largest head gradient into [1, 2)
    // (fp16 operand range, see ndp_tc.cuh); undone when the gradients leave TMEM
    float dscale, dinv;
    ndp_pow2_scale(fmaxf(fmaxf(S.red[0], S.red[1]), fmaxf(S.red[2], S.red[3])), dscale, dinv);
    if (tid < NDP_TP) {
        float v0[8], v1[8], e1[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { v0[r] = hgv[r] * dscale; v1[r] = hgv[8 + r] * dscale; e1[r] = 0.0f; }
        ndp_store_chunk2(S.HG, NDP_IMG16, ndp_img_off(tid, 0, RS16), v0);
        ndp_store_chunk2(S.HG, NDP_IMG16, ndp_img_off(tid, 8, RS16), v1);
        ndp_store_chunk2(S.E, NDP_IMG16, ndp_img_off(tid, 0, RS16), e0);
        ndp_store_chunk2(S.E, NDP_IMG16, ndp_img_off(tid, 8, RS16), e1);
    }
    ndp_fence_proxy_async();
    __syncthreads();
    NDP_T(2);

    const unsigned id_nn = ndp_idesc_f16(128, 128, 1, 1), id_sm = ndp_idesc_f16(128, 16, 1, 1), id_kn = ndp_idesc_f16(128, 128, 0, 1);
    const NdpUmmaDesc dD_mn = ndp_umma_desc(S.D, RS, CS), dD_k = ndp_umma_desc(S.D, CS, RS), dX_mn = ndp_umma_desc(S.X, RS, CS);
    const NdpUmmaDesc dE = ndp_umma_desc(S.E, RS16, CS), dHG = ndp_umma_desc(S.HG, RS16, CS);
    // ---- head gradients: dW_h = hg^T h_L, db_h = hg^T 1.  The [128][16] head-gradient image is read as
    //      a 128-row MN-major operand: rows >= 16 of the result are other rows' data and are never read.
    ndp_mbar_wait(&S.bar_x, xph); xph ^= 1;
    NDP_T(3);
    if (tid == 0) {
        ndp_tc_fence_after();
        ndp_umma_gemm3(tmem + TM_DH, dHG, NDP_IMG16, 2 * RS16, dX_mn, NDP_IMG128, 2 * RS, 8, id_nn, false);
        ndp_umma_gemm_a2(tmem + TM_SM, dHG, NDP_IMG16, 2 * RS16, dE, 2 * RS16, 8, id_sm);
        ndp_umma_commit(&S.bar_mma);
    }
    // ---- delta at the top activation straight into the delta image while the head GEMMs run:
    //      (W_h^T hg) . relu'(h_L)
    {
        // this row's (scaled) head gradients, re-assembled from the two fp16 parts of the image
        float g[16];
        {
            float t8[8];
            ndp_load_chunk2(S.HG, NDP_IMG16, ndp_img_off(p, 0, RS16), t8);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = t8[k];
            ndp_load_chunk2(S.HG, NDP_IMG16, ndp_img_off(p, 8, RS16), t8);
#pragma unroll
            for (int k = 0; k < 8; ++k) g[8 + k] = t8[k];
        }
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
            const int o0 = half * 64 + ch * 8;
            const unsigned pm = ndp_pos_mask8(*(const uint4*)(S.X + ndp_img_off(p, o0, RS)));
            float u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = 0.0f;
#pragma unroll
            for (int r = 0; r < NDP_MAX_HEAD; ++r) {     // rows >= head_dim of hw are zero
                const float4 wa = *(const float4*)(S.hw + r * NDP_W + o0), wb = *(const float4*)(S.hw + r * NDP_W + o0 + 4);
                u[0] = fmaf(g[r], wa.x, u[0]); u[1] = fmaf(g[r], wa.y, u[1]); u[2] = fmaf(g[r], wa.z, u[2]); u[3] = fmaf(g[r], wa.w, u[3]);
                u[4] = fmaf(g[r], wb.x, u[4]); u[5] = fmaf(g[r], wb.y, u[5]); u[6] = fmaf(g[r], wb.z, u[6]); u[7] = fmaf(g[r], wb.w, u[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = ((pm >> j) & 1u) ? u[j] : 0.0f;
            ndp_store_chunk2(S.D, NDP_IMG128, ndp_img_off(p, o0, RS), u);
        }
    }
    NDP_T(4);
    ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
    ndp_tc_fence_after();
    NDP_T(5);
    ndp_fence_proxy_async();
    __syncthreads();            // every thread is done with h_L (relu' masks) and delta_top is complete
    if (tid == 0 && LH > 0) ndp_stage_bulk(S.X, gact + (long long)(LH - 1) * NDP_SET128, NDP_SET128, &S.bar_x);
    if ((warp & 3) == 0) {      // TMEM lanes 0..31 hold the head rows
#pragma unroll 1
        for (int c32 = 0; c32 < 2; ++c32) {
            float v[32];
            const int col0 = half * 64 + c32 * 32;
            ndp_tmem_ld32(tlane + TM_DH + col0, v);
            if (lane < HD) {
                float* dst = part + L.head_w[lane] + col0;
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[j] = v[j] * dinv;
            }
        }
        if (half == 0) {
            float v[32];
            ndp_tmem_ld32(tlane + TM_SM, v);
            if (lane < HD) part[L.head_b[lane]] = v[6] * dinv;
        }
    }
    ndp_tc_fence_before();
    __syncthreads();
    ndp_tc_fence_after();
    NDP_T(6);

    for (int l = LH - 1; l >= 0; --l) {
        const int tb = 8 + 8 * (LH - 1 - l);
        ndp_mbar_wait(&S.bar_x, xph); xph ^= 1;        // h_l (requested before the previous epilogue)
        NDP_T(tb + 0);
        if (tid == 0) {
            ndp_tc_fence_after();
            // dW_l[o][i] = sum_p delta[p][o] h_l[p][i];  db_l[o] = sum_p delta[p][o] (ones column of E)
            ndp_umma_gemm3(tmem + TM_DW, dD_mn, NDP_IMG128, 2 * RS, dX_mn, NDP_IMG128, 2 * RS, 8, id_nn, false);
            ndp_umma_gemm_a2(tmem + TM_SM, dD_mn, NDP_IMG128, 2 * RS, dE, 2 * RS16, 8, id_sm);
            ndp_umma_commit(&S.bar_mma);
        }
        // relu' mask of h_l for this thread's row / column half, before W_l replaces h_l
        unsigned mask[2] = {0u, 0u};
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
            mask[ch >> 2] |= ndp_pos_mask8(*(const uint4*)(S.X + ndp_img_off(p, half * 64 + ch * 8, RS))) << ((ch & 3) * 8);
        NDP_T(tb + 1);
        ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
        ndp_tc_fence_after();
        __syncthreads();        // all masks extracted: h_l may be overwritten
        NDP_T(tb + 2);
        if (tid == 0) ndp_stage_bulk(S.X, wimg + (long long)l * NDP_SET128, NDP_SET128, &S.bar_x);     // W_l over h_l
        // dW epilogue: TMEM -> this tile's partial row (thread = output row o, 64 input columns)
        {
            const int o = (warp & 3) * 32 + lane;
#pragma unroll 1
            for (int c32 = 0; c32 < 2; ++c32) {
                float v[32];
                const int col0 = half * 64 + c32 * 32;
                ndp_tmem_ld32(tlane + TM_DW + col0, v);
                float* dst = part + L.off_w[l] + o * NDP_W + col0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *(float4*)(dst + j) = make_float4(v[j] * dinv, v[j + 1] * dinv, v[j + 2] * dinv, v[j + 3] * dinv);
            }
            if (half == 0) {
                float v[32];
                ndp_tmem_ld32(tlane + TM_SM, v);
                part[L.off_b[l] + o] = v[6] * dinv;
            }
        }
        NDP_T(tb + 3);
        ndp_mbar_wait(&S.bar_x, xph); xph ^= 1;
        NDP_T(tb + 4);
        if (tid == 0) {
            ndp_tc_fence_after();
            // delta_l[p][i] = sum_o delta[p][o] W_l[o][i]
            ndp_umma_gemm3(tmem + TM_DH, dD_k, NDP_IMG128, 2 * CS, dX_mn, NDP_IMG128, 2 * RS, 8, id_kn, false);
            ndp_umma_commit(&S.bar_mma);
        }
        ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
        ndp_tc_fence_after();
        NDP_T(tb + 5);
        if (tid == 0 && l > 0) ndp_stage_bulk(S.X, gact + (long long)(l - 1) * NDP_SET128, NDP_SET128, &S.bar_x);   // h_{l-1} over W_l
        // dH epilogue: relu' mask, re-split, delta image updated in place
#pragma unroll 1
        for (int c32 = 0; c32 < 2; ++c32) {
            float v[32];
            const int col0 = half * 64 + c32 * 32;
            ndp_tmem_ld32(tlane + TM_DH + col0, v);
            const unsigned mk = mask[c32];
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
                float u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = ((mk >> (s8 * 8 + j)) & 1u) ? v[s8 * 8 + j] : 0.0f;
                ndp_store_chunk2(S.D, NDP_IMG128, ndp_img_off(p, col0 + s8 * 8, RS), u);
            }
        }
        ndp_tc_fence_before();
        ndp_fence_proxy_async();
        __syncthreads();
        ndp_tc_fence_after();
        NDP_T(tb + 6);
    }

    // ---- input layer: [dW_in | db_in][o][0..6] = sum_p delta_0[p][o] E[p][0..6]
    if (tid == 0) {
        ndp_umma_gemm3(tmem + TM_SM, dD_mn, NDP_IMG128, 2 * RS, dE, NDP_IMG16, 2 * RS16, 8, id_sm, false);
        ndp_umma_commit(&S.bar_mma);
    }
    ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
    ndp_tc_fence_after();
    if (half == 0) {
        float v[32];
        const int o = (warp & 3) * 32 + lane;
        ndp_tmem_ld32(tlane + TM_SM, v);
        float* dst = part + L.off_w_in + o * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c) dst[c] = v[c] * dinv;
        part[L.off_b_in + o] = v[6] * dinv;
    }
    if (a.gx && tid < NDP_TP) {     // optional dL/dx: direct part + path through the positional encoding
        const int gp = tile * NDP_TP + tid;
        if (gp < n) {
            float de[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            const float* wi = params + L.off_w_in;
            for (int ch = 0; ch < 16; ++ch) {
                float d8[8];
                ndp_load_chunk2(S.D, NDP_IMG128, ndp_img_off(tid, ch * 8, RS), d8);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = d8[j] * dinv;
                    const int o = ch * 8 + j;
#pragma unroll
                    for (int c = 0; c < 6; ++c) de[c] = fmaf(d, __ldg(wi + o * 6 + c), de[c]);
                }
            }
            float* gxp = a.gx + (long long)pair * a.gx_stride + (long long)gp * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float s, c;
                sincosf(S.xs[tid * 4 + d] * L.freq, &s, &c);
                gxp[d] = S.gxs[tid * 4 + d] + L.freq * (c * de[2 * d] - s * de[2 * d + 1]);
            }
        }
    }
    NDP_T(62);
    ndp_tc_fence_before();
    __syncthreads();
    NDP_T(63);
    if (warp == 0) ndp_tmem_dealloc(tmem, 512);
}

void ndp_launch_bwd_tc(const NdpBwdArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    dim3 grid((a.n + NDP_TP - 1) / NDP_TP, a.npairs);
    NDP_LAUNCH(ndp_warp_bwd_tc_kernel, grid, dim3(NDP_THREADS), ndp_bwd_tc_smem_bytes(), s, a);
}

int ndp_bwd_tc_init() {
    return (int)cudaFuncSetAttribute(ndp_warp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ndp_bwd_tc_smem_bytes());
}

#ifndef NDP_EMU
int ndp_debug_copy_bwd(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, ndp_dbg_bwd, sizeof(unsigned long long) * 64); }
#else
int ndp_debug_copy_bwd(unsigned long long* out) { for (int i = 0; i < 64; ++i) out[i] = 0; return 0; }
#endif
