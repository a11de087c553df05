// Kernel (3b): fixed-order reduction of the per-tile gradient partials, fused with the Adam update
// and with the refresh of the transposed weight copies the forward kernel streams by TMA.
//
// Reference: torch.optim.Adam(params, lr) as constructed at model/registration.py:176 and stepped
// at :237 [upstream torch; defaults betas=(0.9,0.999), eps=1e-8, no weight decay, no amsgrad]:
//   m <- m + (1-b1)(g - m);  v <- b2 v + (1-b2) g^2
//   p <- p - (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// A fresh optimiser per level (registration.py:176) => t restarts at 1 and m = v = 0.
// Also: small helper kernels of the per-pair driver (means, gather+centre, state reset, pack).
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

__device__ __forceinline__ void ndp_pack_store(const NdpLayout& L, float* pack, int idx, float val, int fp32_copies) {
    // canonical index -> transposed copy (W_in[o][c] -> WT_in[c][o]; W_l[o][i] -> WT_l[i][o])
    if (idx < L.off_b_in) {
        const int o = idx / 6, c = idx - o * 6;
        if (fp32_copies) pack[L.pack_in + c * NDP_W + o] = val;
        return;
    }
    for (int l = 0; l < L.hidden; ++l) {
        const int rel = idx - L.off_w[l];
        if (rel >= 0 && rel < NDP_W * NDP_W) {
            const int o = rel >> 7, i = rel & 127;
            if (fp32_copies) pack[L.pack_w[l] + i * NDP_W + o] = val;
            // fp16 hi/lo images of W_l (row o, column i) for the tensor-core kernels
            unsigned h1, h2;
            ndp_split2(val, h1, h2);
            unsigned char* img = (unsigned char*)(pack + L.pack_img) + (long long)l * NDP_SET128 + ndp_img_off(o, i, NDP_IMG_RS(128));
            *(unsigned short*)img = (unsigned short)h1;
            *(unsigned short*)(img + NDP_IMG128) = (unsigned short)h2;
            return;
        }
    }
}

__global__ void __launch_bounds__(256) ndp_reduce_adam_kernel(NdpAdamArgs a) {
    __shared__ float s_fac[2];              // step_size, sqrt(bias_correction2): one fp64 evaluation per CTA
    const int pair = blockIdx.y + a.pair0;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    if (a.do_adam && threadIdx.x == 0) {
        // torch evaluates the scalar factors in Python floats (fp64) and rounds them to fp32 once
        const int step = a.state ? a.state[pair].evals : a.fixed_step;
        const double bc1 = 1.0 - pow(a.beta1, (double)step);
        const double bc2 = 1.0 - pow(a.beta2, (double)step);
        s_fac[0] = (float)(a.lr / bc1);
        s_fac[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= L.param_count) return;
    const int n = a.counts ? a.counts[pair] : a.n;
    const int tiles = n > 0 ? ((n + NDP_TP - 1) / NDP_TP + a.tiles_per_row - 1) / a.tiles_per_row : 1;   // partial rows
    const float* part = a.partials + (long long)pair * a.partials_stride + idx;
    float g = 0.0f;
    int t = 0;
    for (; t + 8 <= tiles; t += 8) {        // ascending tile order; the eight loads are issued together
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = part[(long long)(t + k) * a.partial_pitch];
#pragma unroll
        for (int k = 0; k < 8; ++k) g += v[k];
    }
    for (; t < tiles; ++t) g += part[(long long)t * a.partial_pitch];
    if (a.grads_out) a.grads_out[(long long)pair * a.grads_stride + idx] = g;
    if (!a.do_adam) return;
    float* p = a.params + (long long)pair * a.params_stride + idx;
    float* mp = a.m + (long long)pair * a.mv_stride + idx;
    float* vp = a.v + (long long)pair * a.mv_stride + idx;
    const float step_size = s_fac[0], bc2_sqrt = s_fac[1];
    const float w1 = (float)(1.0 - a.beta1), b2 = (float)a.beta2, w2 = (float)(1.0 - a.beta2);
    float m = *mp, v = *vp;
    m = m + w1 * (g - m);                 // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + w2 * (g * g);            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / bc2_sqrt + (float)a.eps;
    const float pv = *p - step_size * (m / denom);
    *mp = m; *vp = v; *p = pv;
    if (a.pack) ndp_pack_store(L, a.pack + (long long)pair * a.pack_stride, idx, pv, a.pack_fp32);
}

void ndp_launch_adam(const NdpAdamArgs& a, cudaStream_t s) {
    if (a.npairs <= 0) return;
    dim3 grid((a.lay.param_count + 255) / 256, a.npairs);
    NDP_LAUNCH_PRIO(1, ndp_reduce_adam_kernel, grid, dim3(256), 0, s, a);
}

__global__ void __launch_bounds__(256) ndp_pack_kernel(NdpPackArgs a) {
    const int pair = blockIdx.y;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const NdpLayout& L = a.lay;
    if (idx >= L.pack_img + L.hidden * NDP_W * NDP_W) return;
    const float* params = a.params + (long long)pair * a.params_stride;
    float* pack = a.pack + (long long)pair * a.pack_stride;
    if (idx >= L.pack_img) {        // one thread per hidden weight: its two fp16 image entries
        const int rel = idx - L.pack_img;
        const int l = rel / (NDP_W * NDP_W), r2 = rel - l * NDP_W * NDP_W;
        const int o = r2 >> 7, i = r2 & 127;
        unsigned h1, h2;
        ndp_split2(params[L.off_w[l] + o * NDP_W + i], h1, h2);
        unsigned char* img = (unsigned char*)(pack + L.pack_img) + (long long)l * NDP_SET128 + ndp_img_off(o, i, NDP_IMG_RS(128));
        *(unsigned short*)img = (unsigned short)h1;
        *(unsigned short*)(img + NDP_IMG128) = (unsigned short)h2;
        return;
    }
    float v;
    if (idx < 6 * NDP_W) {
        const int c = idx >> 7, o = idx & 127;
        v = params[L.off_w_in + o * 6 + c];
    } else {
        const int rel = idx - 6 * NDP_W;
        const int l = rel / (NDP_W * NDP_W), r2 = rel - l * NDP_W * NDP_W;
        const int i = r2 >> 7, o = r2 & 127;
        v = params[L.off_w[l] + o * NDP_W + i];
    }
    pack[idx] = v;
}

void ndp_launch_pack(const NdpPackArgs& a, cudaStream_t s) {
    if (a.npairs <= 0) return;
    dim3 grid((a.lay.pack_img + a.lay.hidden * NDP_W * NDP_W + 255) / 256, a.npairs);
    NDP_LAUNCH(ndp_pack_kernel, grid, dim3(256), 0, s, a);
}

// ---- cloud means (registration.py:150-151): one CTA per (pair, cloud); fp64 tree => deterministic
__global__ void __launch_bounds__(256) ndp_means_kernel(NdpCenterArgs a) {
    __shared__ double red[3][256];
    const int pair = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
    const int n = which ? (a.ntcounts ? a.ntcounts[pair] : a.nt) : (a.nscounts ? a.nscounts[pair] : a.ns);
    const float* P = which ? a.tgt + (long long)pair * a.tgt_stride : a.src + (long long)pair * a.src_stride;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = tid; i < n; i += 256) { s0 += P[(long long)i * 3]; s1 += P[(long long)i * 3 + 1]; s2 += P[(long long)i * 3 + 2]; }
    red[0][tid] = s0; red[1][tid] = s1; red[2][tid] = s2;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) { red[0][tid] += red[0][tid + off]; red[1][tid] += red[1][tid + off]; red[2][tid] += red[2][tid + off]; }
        __syncthreads();
    }
    if (tid < 3) a.means[((long long)pair * 2 + which) * 3 + tid] = (float)(red[tid][0] / (double)(n > 0 ? n : 1));
}

void ndp_launch_means(const NdpCenterArgs& a, cudaStream_t s) {
    if (a.npairs <= 0) return;
    NDP_LAUNCH(ndp_means_kernel, dim3(a.npairs, 2), dim3(256), 0, s, a);
}

__global__ void __launch_bounds__(256) ndp_gather_center_kernel(NdpGatherArgs a) {
    const int pair = blockIdx.y;
    const int n = a.counts ? a.counts[pair] : a.n;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int src = a.idx ? a.idx[(long long)pair * a.idx_stride + i] : i;
    const float* in = a.in + (long long)pair * a.in_stride + (long long)src * 3;
    const float* mu = a.means + ((long long)pair * 2 + a.which) * 3;
    float* out = a.out + (long long)pair * a.out_stride + (long long)i * 3;
    out[0] = in[0] - mu[0]; out[1] = in[1] - mu[1]; out[2] = in[2] - mu[2];
}

void ndp_launch_gather_center(const NdpGatherArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    dim3 grid((a.n + 255) / 256, a.npairs);
    NDP_LAUNCH(ndp_gather_center_kernel, grid, dim3(256), 0, s, a);
}

__global__ void ndp_state_reset_kernel(NdpPairState* st, int npairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    st[i].stopped = 0; st[i].steps = 0; st[i].break_counter = 0; st[i].evals = 0;
    st[i].loss_prev = 1e6;                                        // registration.py:179-180
    st[i].last_loss = 0.0f; st[i].pad = 0.0f;
}

void ndp_launch_state_reset(NdpPairState* state, int npairs, cudaStream_t s) {
    if (npairs <= 0) return;
    NDP_LAUNCH(ndp_state_reset_kernel, dim3((npairs + 127) / 128), dim3(128), 0, s, state, npairs);
}

