// Kernel (1): one pyramid level of the warp field over a tile of 128 points --
// positional encoding -> input layer -> (depth-1) hidden 128x128 layers -> rotation / scale /
// translation / nonrigidity heads -> SE(3) / Sim(3) / scene-flow composition.
//
// Reference: model/nets.py:111-140 (NDPLayer.forward), :164-177 (posenc), :295-304 (MLP),
//            :144-161 (get_Rotation), model/rigid_body.py.
//
// Data flow per CTA (256 threads, 1 CTA/SM, ~208 KB shared memory):
//   * the two next hidden-layer weight matrices (transposed copies, 64 KB each) are staged into a
//     shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier) while the previous layer
//     computes;
//   * the [128 points][128 features] activation tile lives in shared memory and is updated in place;
//     each thread accumulates an 8x8 output patch in registers (ndp_gemm_nn);
//   * activations and the head vector are written to HBM once for the backward kernel (coalesced
//     128-bit stores, [point][feature] layout).
#include "ndp_kernels.h"
#include "ndp_mlp.cuh"

#define FWD_SMEM_FLOATS (NDP_TP * NDP_PITCH + 2 * NDP_W * NDP_W + NDP_TP * NDP_ZPITCH + NDP_TP * 4 + \
                         NDP_MAX_HEAD * NDP_W + 16)
size_t ndp_fwd_smem_bytes() { return FWD_SMEM_FLOATS * sizeof(float) + 64; }

__global__ void __launch_bounds__(NDP_THREADS, 1) ndp_warp_fwd_kernel(NdpFwdArgs a) {
    NDP_DYN_SMEM(smem);
    float* act = (float*)smem;                          // [TP][PITCH]
    float* wbuf = act + NDP_TP * NDP_PITCH;             // 2 x [128][128] (k-major: WT[k][o])
    float* zs = wbuf + 2 * NDP_W * NDP_W;               // [TP][ZPITCH] head outputs
    float* xs = zs + NDP_TP * NDP_ZPITCH;               // [TP][4] input points
    float* hw = xs + NDP_TP * 4;                        // [head_dim][128] head weights
    float* hb = hw + NDP_MAX_HEAD * NDP_W;              // [head_dim] head biases
    NdpMbar* bar = (NdpMbar*)(hb + 16);                 // 2 barriers

    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0, tile = blockIdx.x;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const float* pack = a.pack + (long long)pair * a.pack_stride;
    const int LH = L.hidden, HD = L.head_dim;
    const unsigned WBYTES = NDP_W * NDP_W * sizeof(float);

    if (tid == 0) { ndp_mbar_init(&bar[0], 1); ndp_mbar_init(&bar[1], 1); }
    __syncthreads();
    if (tid == 0) {
        if (LH > 0) ndp_stage_bulk(wbuf, pack + L.pack_w[0], WBYTES, &bar[0]);
        if (LH > 1) ndp_stage_bulk(wbuf + NDP_W * NDP_W, pack + L.pack_w[1], WBYTES, &bar[1]);
    }
    // head weights / biases -> smem (the canonical block is not 16-byte aligned per row)
    for (int i = tid; i < HD * NDP_W; i += NDP_THREADS) hw[i] = __ldg(params + L.head_w[i >> 7] + (i & 127));
    if (tid < HD) hb[tid] = __ldg(params + L.head_b[tid]);
    // input points + positional encoding (nets.py:164-177; accurate sin/cos, f = 2^(m+k0))
    if (tid < NDP_TP) {
        const int gp = tile * NDP_TP + tid;
        float px = 0.0f, py = 0.0f, pz = 0.0f;
        if (gp < n) {
            const float* xp = a.x + (long long)pair * a.x_stride + (long long)gp * 3;
            px = __ldg(xp); py = __ldg(xp + 1); pz = __ldg(xp + 2);
        }
        xs[tid * 4 + 0] = px; xs[tid * 4 + 1] = py; xs[tid * 4 + 2] = pz;
        float s, c;
        float* e = act + tid * NDP_PITCH;
        sincosf(px * L.freq, &s, &c); e[0] = s; e[1] = c;
        sincosf(py * L.freq, &s, &c); e[2] = s; e[3] = c;
        sincosf(pz * L.freq, &s, &c); e[4] = s; e[5] = c;
    }
    __syncthreads();

    const int tr = tid >> 4, tc = tid & 15;
    float acc[8][8];
    float* act_out = a.act ? a.act + (long long)pair * a.act_stride : nullptr;

    // ---- input layer: h0 = relu(W_in e + b_in), K = 6 (nets.py:75,114)
    {
        const float* wt = pack + L.pack_in;            // WT_in[6][128]
        const float4 bi0 = __ldg((const float4*)(params + L.off_b_in) + tc);
        const float4 bi1 = __ldg((const float4*)(params + L.off_b_in + 64) + tc);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            acc[r][0] = bi0.x; acc[r][1] = bi0.y; acc[r][2] = bi0.z; acc[r][3] = bi0.w;
            acc[r][4] = bi1.x; acc[r][5] = bi1.y; acc[r][6] = bi1.z; acc[r][7] = bi1.w;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float4 b0 = __ldg((const float4*)(wt + k * NDP_W) + tc);
            const float4 b1 = __ldg((const float4*)(wt + k * NDP_W + 64) + tc);
#pragma unroll
            for (int r = 0; r < 8; ++r) ndp_fma_row(acc[r], act[ndp_row8(tr, r) * NDP_PITCH + k], b0, b1);
        }
    }
    __syncthreads();   // every thread has read its encoding columns
    for (int l = 0;; ++l) {
        // ---- ReLU, write the tile back in place, save it for the backward pass
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = ndp_row8(tr, r);
            float4 v0 = make_float4(fmaxf(acc[r][0], 0.0f), fmaxf(acc[r][1], 0.0f), fmaxf(acc[r][2], 0.0f), fmaxf(acc[r][3], 0.0f));
            float4 v1 = make_float4(fmaxf(acc[r][4], 0.0f), fmaxf(acc[r][5], 0.0f), fmaxf(acc[r][6], 0.0f), fmaxf(acc[r][7], 0.0f));
            *(float4*)(act + row * NDP_PITCH + tc * 4) = v0;
            *(float4*)(act + row * NDP_PITCH + 64 + tc * 4) = v1;
            const int gp = tile * NDP_TP + row;
            if (act_out && gp < n) {
                float* dst = act_out + (long long)l * a.act_layer_stride + (long long)gp * NDP_W;
                *(float4*)(dst + tc * 4) = v0;
                *(float4*)(dst + 64 + tc * 4) = v1;
            }
        }
        __syncthreads();
        if (l == LH) break;
        // ---- hidden layer l: h_{l+1} = relu(W_l h_l + b_l)   (nets.py:300-304)
        const float* wt = wbuf + (l & 1) * NDP_W * NDP_W;
        {
            const float4 bi0 = __ldg((const float4*)(params + L.off_b[l]) + tc);
            const float4 bi1 = __ldg((const float4*)(params + L.off_b[l] + 64) + tc);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                acc[r][0] = bi0.x; acc[r][1] = bi0.y; acc[r][2] = bi0.z; acc[r][3] = bi0.w;
                acc[r][4] = bi1.x; acc[r][5] = bi1.y; acc[r][6] = bi1.z; acc[r][7] = bi1.w;
            }
        }
        ndp_mbar_wait(&bar[l & 1], (unsigned)((l >> 1) & 1));
        ndp_gemm_nn(act, wt, acc, tr, tc);
        __syncthreads();   // all reads of act and of this ring slot are done
        if (tid == 0 && l + 2 < LH)
            ndp_stage_bulk(wbuf + (l & 1) * NDP_W * NDP_W, pack + L.pack_w[l + 2], WBYTES, &bar[l & 1]);
    }

    // ---- heads: z = mlp_scale * (W_h h + b_h)   (nets.py:117,125,133,146)
    {
        const int p = tid & (NDP_TP - 1), grp = tid >> 7;
        float hacc[NDP_MAX_HEAD / 2];
#pragma unroll
        for (int i = 0; i < NDP_MAX_HEAD / 2; ++i) hacc[i] = 0.0f;
#pragma unroll 2
        for (int k0 = 0; k0 < NDP_W; k0 += 4) {
            const float4 h = *(const float4*)(act + p * NDP_PITCH + k0);
#pragma unroll
            for (int i = 0; i < NDP_MAX_HEAD / 2; ++i) {
                const int r = grp + 2 * i;
                if (r < HD) {
                    const float4 w = *(const float4*)(hw + r * NDP_W + k0);
                    hacc[i] = fmaf(h.x, w.x, hacc[i]); hacc[i] = fmaf(h.y, w.y, hacc[i]);
                    hacc[i] = fmaf(h.z, w.z, hacc[i]); hacc[i] = fmaf(h.w, w.w, hacc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NDP_MAX_HEAD / 2; ++i) {
            const int r = grp + 2 * i;
            if (r < HD) zs[p * NDP_ZPITCH + r] = L.mu * (hacc[i] + hb[r]);
        }
    }
    __syncthreads();

    // ---- per-point rotation + warp composition (nets.py:119-137)
    if (tid < NDP_TP) {
        const int gp = tile * NDP_TP + tid;
        const float INF = __int_as_float(0x7f800000);
        float y[3] = {INF, INF, INF};
        if (gp < n) {
            float z[NDP_MAX_HEAD], nu = 0.0f;
#pragma unroll
            for (int r = 0; r < NDP_MAX_HEAD; ++r) z[r] = (r < HD) ? zs[tid * NDP_ZPITCH + r] : 0.0f;
            ndp_point_forward(L.motion, L.rot, L.nonrigid, z, xs + tid * 4, y, &nu);
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
            // float4 copy (x, y, z, original sample index) and the box of this warp's 32 outputs for
            // the culled NN search; slots past the cloud hold +inf and are excluded from the box
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
}

void ndp_launch_fwd(const NdpFwdArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    dim3 grid((a.n + NDP_TP - 1) / NDP_TP, a.npairs);
    NDP_LAUNCH(ndp_warp_fwd_kernel, grid, dim3(NDP_THREADS), ndp_fwd_smem_bytes(), s, a);
}

int ndp_fwd_init() {
    return (int)cudaFuncSetAttribute(ndp_warp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ndp_fwd_smem_bytes());
}
