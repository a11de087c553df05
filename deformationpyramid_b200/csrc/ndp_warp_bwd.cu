// Kernel (3a): backward of kernel (1) over a tile of 128 points -> this tile's contribution to
// every parameter gradient of the level ("partials"), reduced in fixed order by kernel (3b).
//
// Replaces the autograd graph the reference builds for NDPLayer.forward (model/nets.py:111-140)
// and walks in loss.backward() (model/registration.py:236): closed-form rotation / warp backward
// per point (ndp_math.cuh), ReLU-MLP back-propagation, dW = sum_n delta_n h_n^T.
//
// Per CTA (256 threads, ~212 KB smem): delta tile [128][128] and activation tile [128][128] in
// shared memory; the canonical W_l (64 KB) of the layer being back-propagated through is staged
// by a TMA bulk copy issued one layer ahead; dW (reduction over the tile's points) and
// dH = delta W are the register-tiled GEMMs of ndp_mlp.cuh.  No atomics: each tile writes its own
// partial row, so the parameter gradients are bit-reproducible run to run.
#include "ndp_kernels.h"
#include "ndp_mlp.cuh"

#define BWD_SMEM_FLOATS (2 * NDP_TP * NDP_PITCH + NDP_W * NDP_W + NDP_MAX_HEAD * NDP_W + \
                         NDP_TP * NDP_ZPITCH + NDP_TP * 4 + NDP_TP * 4 + 16)
size_t ndp_bwd_smem_bytes() { return BWD_SMEM_FLOATS * sizeof(float) + 64; }

__global__ void __launch_bounds__(NDP_THREADS, 1) ndp_warp_bwd_kernel(NdpBwdArgs a) {
    NDP_DYN_SMEM(smem);
    float* dbuf = (float*)smem;                         // [TP][PITCH] delta (grad wrt pre-activation)
    float* hbuf = dbuf + NDP_TP * NDP_PITCH;            // [TP][PITCH] activation of the layer below
    float* wbuf = hbuf + NDP_TP * NDP_PITCH;            // [128][128] canonical W_l[o][i]
    float* hw = wbuf + NDP_W * NDP_W;                   // [head_dim][128]
    float* hg = hw + NDP_MAX_HEAD * NDP_W;              // [TP][ZPITCH] mlp_scale * dL/dz
    float* xs = hg + NDP_TP * NDP_ZPITCH;               // [TP][4]
    float* gxs = xs + NDP_TP * 4;                       // [TP][4] direct part of dL/dx
    NdpMbar* bar = (NdpMbar*)(gxs + NDP_TP * 4);

    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0, tile = blockIdx.x;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const float* actg = a.act + (long long)pair * a.act_stride;
    float* part = a.partials + (long long)pair * a.partials_stride + (long long)tile * a.partial_pitch;
    const int LH = L.hidden, HD = L.head_dim;
    const unsigned WBYTES = NDP_W * NDP_W * sizeof(float);

    if (tid == 0) ndp_mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0 && LH > 0) ndp_stage_bulk(wbuf, params + L.off_w[LH - 1], WBYTES, bar);
    for (int i = tid; i < HD * NDP_W; i += NDP_THREADS) hw[i] = __ldg(params + L.head_w[i >> 7] + (i & 127));
    ndp_load_tile(hbuf, actg + (long long)LH * a.act_layer_stride, tile, n, tid);

    // ---- per point: dL/dy -> dL/dz (heads) and the direct part of dL/dx
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
            if (a.gacc) {
                // scattered Chamfer term, accumulated in 2^-40 fixed point (order independent)
                unsigned long long* ga = a.gacc + (long long)pair * a.gacc_stride + (long long)gp * 3;
                const int m = a.mcounts ? a.mcounts[pair] : a.m;
                const double sc = 9.094947017729282e-13 / (double)m;    // 2^-40 / m
                const long long a0 = (long long)ga[0], a1 = (long long)ga[1], a2 = (long long)ga[2];
                const bool poison = (a0 >= (1LL << 60)) || (a0 <= -(1LL << 60));
                gy[0] += poison ? __int_as_float(0x7fc00000) : (float)((double)a0 * sc);
                gy[1] += (float)((double)a1 * sc);
                gy[2] += (float)((double)a2 * sc);
                ga[0] = 0ull; ga[1] = 0ull; ga[2] = 0ull;
            }
            float z[NDP_MAX_HEAD];
            const float* zp = a.zsave + (long long)pair * a.z_stride + (long long)gp * NDP_ZPITCH;
#pragma unroll
            for (int r = 0; r < NDP_MAX_HEAD; ++r) z[r] = zp[r];
            const float gnu = (a.gnu && L.nonrigid) ? a.gnu[(long long)pair * a.gnu_stride + gp] : 0.0f;
            ndp_point_backward(L.motion, L.rot, L.nonrigid, z, x, gy, gnu, gz, gxd);
        }
#pragma unroll
        for (int r = 0; r < NDP_MAX_HEAD; ++r) hg[tid * NDP_ZPITCH + r] = L.mu * gz[r];
        xs[tid * 4 + 0] = x[0]; xs[tid * 4 + 1] = x[1]; xs[tid * 4 + 2] = x[2];
        gxs[tid * 4 + 0] = gxd[0]; gxs[tid * 4 + 1] = gxd[1]; gxs[tid * 4 + 2] = gxd[2];
    }
    __syncthreads();

    // ---- head parameter gradients: dW_h[r][k] = sum_p hg[p][r] h[p][k], db_h[r] = sum_p hg[p][r]
    {
        const int k = tid & (NDP_W - 1), grp = tid >> 7;
        float hacc[NDP_MAX_HEAD / 2];
#pragma unroll
        for (int i = 0; i < NDP_MAX_HEAD / 2; ++i) hacc[i] = 0.0f;
#pragma unroll 2
        for (int p = 0; p < NDP_TP; ++p) {
            const float hv = hbuf[p * NDP_PITCH + k];
#pragma unroll
            for (int i = 0; i < NDP_MAX_HEAD / 2; ++i) hacc[i] = fmaf(hg[p * NDP_ZPITCH + grp + 2 * i], hv, hacc[i]);
        }
#pragma unroll
        for (int i = 0; i < NDP_MAX_HEAD / 2; ++i) {
            const int r = grp + 2 * i;
            if (r < HD) part[L.head_w[r] + k] = hacc[i];
        }
        if (tid < HD) {
            float s = 0.0f;
            for (int p = 0; p < NDP_TP; ++p) s += hg[p * NDP_ZPITCH + tid];
            part[L.head_b[tid]] = s;
        }
    }
    // ---- delta at the top activation: (W_h^T hg) masked by relu'
    {
        const int o = tid & (NDP_W - 1), half = tid >> 7;
        float wh[NDP_MAX_HEAD];
#pragma unroll
        for (int r = 0; r < NDP_MAX_HEAD; ++r) wh[r] = (r < HD) ? hw[r * NDP_W + o] : 0.0f;
        for (int p = half * 64; p < half * 64 + 64; ++p) {
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < NDP_MAX_HEAD; ++r) s = fmaf(hg[p * NDP_ZPITCH + r], wh[r], s);
            dbuf[p * NDP_PITCH + o] = (hbuf[p * NDP_PITCH + o] > 0.0f) ? s : 0.0f;
        }
    }
    __syncthreads();

    const int tr = tid >> 4, tc = tid & 15;
    float acc[8][8];
    for (int l = LH - 1; l >= 0; --l) {
        // db_l = column sums of delta_{l+1}
        if (tid < NDP_W) {
            float s = 0.0f;
            for (int p = 0; p < NDP_TP; ++p) s += dbuf[p * NDP_PITCH + tid];
            part[L.off_b[l] + tid] = s;
        }
        // activation below this layer
        ndp_load_tile(hbuf, actg + (long long)l * a.act_layer_stride, tile, n, tid);
        __syncthreads();
        // dW_l[o][i] = sum_p delta[p][o] h_l[p][i]
        ndp_acc_zero(acc);
        ndp_gemm_tn(dbuf, hbuf, acc, tr, tc);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float* dst = part + L.off_w[l] + ndp_row8(tr, r) * NDP_W;
            *(float4*)(dst + tc * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            *(float4*)(dst + 64 + tc * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
        // delta_l = (delta_{l+1} W_l) masked by relu'(h_l)
        ndp_acc_zero(acc);
        ndp_mbar_wait(bar, (unsigned)((LH - 1 - l) & 1));
        ndp_gemm_nn(dbuf, wbuf, acc, tr, tc);
        __syncthreads();   // all reads of dbuf / wbuf done
        if (tid == 0 && l > 0) ndp_stage_bulk(wbuf, params + L.off_w[l - 1], WBYTES, bar);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = ndp_row8(tr, r);
            const float4 h0 = *(const float4*)(hbuf + row * NDP_PITCH + tc * 4);
            const float4 h1 = *(const float4*)(hbuf + row * NDP_PITCH + 64 + tc * 4);
            *(float4*)(dbuf + row * NDP_PITCH + tc * 4) =
                make_float4(h0.x > 0.0f ? acc[r][0] : 0.0f, h0.y > 0.0f ? acc[r][1] : 0.0f,
                            h0.z > 0.0f ? acc[r][2] : 0.0f, h0.w > 0.0f ? acc[r][3] : 0.0f);
            *(float4*)(dbuf + row * NDP_PITCH + 64 + tc * 4) =
                make_float4(h1.x > 0.0f ? acc[r][4] : 0.0f, h1.y > 0.0f ? acc[r][5] : 0.0f,
                            h1.z > 0.0f ? acc[r][6] : 0.0f, h1.w > 0.0f ? acc[r][7] : 0.0f);
        }
        __syncthreads();
    }

    // ---- input layer: db_in, dW_in[o][c] = sum_p delta0[p][o] e[p][c]; optional dL/dx
    if (tid < NDP_W) {
        float s = 0.0f;
        for (int p = 0; p < NDP_TP; ++p) s += dbuf[p * NDP_PITCH + tid];
        part[L.off_b_in + tid] = s;
    }
    if (tid < NDP_TP) {   // recompute the encoding into hbuf[:, 0..5] (h_0 is no longer needed)
        float s, c;
        float* e = hbuf + tid * NDP_PITCH;
        sincosf(xs[tid * 4 + 0] * L.freq, &s, &c); e[0] = s; e[1] = c;
        sincosf(xs[tid * 4 + 1] * L.freq, &s, &c); e[2] = s; e[3] = c;
        sincosf(xs[tid * 4 + 2] * L.freq, &s, &c); e[4] = s; e[5] = c;
    }
    __syncthreads();
    {
        const int o = tid & (NDP_W - 1), half = tid >> 7;
        float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
        for (int p = 0; p < NDP_TP; ++p) {
            const float d = dbuf[p * NDP_PITCH + o];
            const float* e = hbuf + p * NDP_PITCH + half * 3;
            w0 = fmaf(d, e[0], w0); w1 = fmaf(d, e[1], w1); w2 = fmaf(d, e[2], w2);
        }
        float* dst = part + L.off_w_in + o * 6 + half * 3;
        dst[0] = w0; dst[1] = w1; dst[2] = w2;
    }
    if (a.gx && tid < NDP_TP) {
        const int gp = tile * NDP_TP + tid;
        if (gp < n) {
            float de[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            const float* wi = params + L.off_w_in;
            for (int o = 0; o < NDP_W; ++o) {
                const float d = dbuf[tid * NDP_PITCH + o];
#pragma unroll
                for (int c = 0; c < 6; ++c) de[c] = fmaf(d, __ldg(wi + o * 6 + c), de[c]);
            }
            const float* e = hbuf + tid * NDP_PITCH;
            float* gxp = a.gx + (long long)pair * a.gx_stride + (long long)gp * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d)   // d sin(f x)/dx = f cos, d cos(f x)/dx = -f sin
                gxp[d] = gxs[tid * 4 + d] + L.freq * (e[2 * d + 1] * de[2 * d] - e[2 * d] * de[2 * d + 1]);
        }
    }
}

void ndp_launch_bwd(const NdpBwdArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    dim3 grid((a.n + NDP_TP - 1) / NDP_TP, a.npairs);
    NDP_LAUNCH(ndp_warp_bwd_kernel, grid, dim3(NDP_THREADS), ndp_bwd_smem_bytes(), s, a);
}

int ndp_bwd_init() {
    return (int)cudaFuncSetAttribute(ndp_warp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ndp_bwd_smem_bytes());
}
