// Kernel (3a), tensor-core version: backward of kernel (1) over a tile of 128 points -> this
// tile's partial parameter gradients (reduced in fixed order by kernel (3b)).
//
// Replaces the autograd backward of NDPLayer.forward (model/nets.py:111-140) that the reference
// runs in loss.backward() (model/registration.py:236).
//
// Every contraction is a tcgen05 GEMM (operands as fp16 hi/lo image sets, fp32 accumulation in
// TMEM, see ndp_tc.cuh) issued by one thread; reductions over the tile's points have K = points:
//     delta at top  delta_L = (hg W_h) . relu'(h_L)     A = hg image   (K-major)   B = W_h   (MN-major)
//     head grads    dW_h^T = h_L^T hg                   A = h_L        (MN-major)  B = hg    (MN-major)
//                   db_h   = hg^T 1                     A = hg         (MN-major)  B = E[:,6] = 1
//     weight grads  dW_l  = delta^T h_l                 A = delta      (MN-major)  B = h_l   (MN-major)
//     bias grads    db_l  = delta^T 1                   A = delta      (MN-major)  B = E[:,6] = 1
//     back-prop     delta_l = (delta W_l) . relu'(h_l)  A = delta      (K-major)   B = W_l   (MN-major)
//     input layer   dW_in = delta_0^T e                 A = delta_0    (MN-major)  B = E[:,0:6] = posenc
// The SAME delta image is the MN-major operand of the dW products and the K-major operand of the
// back-propagation product, and the SAME weight image serves forward and backward (the core-matrix
// layout is both canonical UMMA layouts at once), so nothing is ever transposed.  h_l and W_l image
// sets arrive by TMA bulk copies (exactly as the forward kernel's bulk stores / the Adam kernel
// wrote them) into two dedicated buffers, both requested as soon as the previous layer's MMAs have
// retired, so that per layer ONE batch of MMAs (dW_l, db_l, delta_l) is issued back to back; the
// dW accumulators are double buffered in TMEM and drained to HBM by all 16 warps while the next
// layer's MMAs run.  Deltas are carried multiplied by the tile's power-of-two scale (fp16 range).
// No atomics: one partial row per tile.
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

// optional phase timestamps of CTA (0,0) (debug aid, read back through ndp_debug_phase_times)
#ifndef NDP_EMU
__device__ unsigned long long ndp_dbg_bwd[64];
#define NDP_T(i) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ndp_dbg_bwd[i] = t_; } } while (0)
#else
#define NDP_T(i) do {} while (0)
#endif

#define NDP_BWD_TC_THREADS 544              // 16 worker warps + 1 warp whose lane 0 issues every MMA / TMA copy
#define NDP_IMG16 NDP_IMG_BYTES(16)        // [128][16] fp16 image: 4096 bytes
#define NDP_HWIMG (2 * NDP_IMG_RS(128))    // [16][128] fp16 image: 4096 bytes
struct BwdTcSmem {
    unsigned char D[NDP_SET128];        // delta hi/lo images (scaled by the tile's power of two)
    unsigned char H[NDP_SET128];        // h_l images
    unsigned char W[NDP_SET128];        // W_l images
    unsigned char HG[2 * NDP_IMG16];    // [128 points][16]: scaled mlp_scale * dL/dz (head gradients)
    unsigned char E[2 * NDP_IMG16];     // [128 points][16]: cols 0..5 positional encoding, rest 0
    unsigned char HW[2 * NDP_HWIMG];    // [16 head rows][128]: head weights, rows >= head_dim zero
    float dbred[4][NDP_W];              // per lane-quarter column sums of delta (bias gradients)
    float dbacc[NDP_MAX_HIDDEN + 1][NDP_W];   // bias gradients accumulated over the CTA's tiles (slot LH = input layer)
    NdpMbar bar_h, bar_w, bar_mma;
    unsigned tmem_slot, pad[3];
};
size_t ndp_bwd_tc_smem_bytes() { return sizeof(BwdTcSmem) + 128; }

// TMEM columns
#define TM_DH 0u                                 // delta_l accumulator                  [128 points][128]
#define TM_DW(b) (128u + 128u * (unsigned)(b))   // dW_l^T accumulators, double buffered [128 i][128 o]
#define TM_HDW 384u                              // dW_h^T                               [128 i][16 head rows]
#define TM_DIN 400u                              // delta_0^T E: cols 0..5 = dW_in       [128 o][16]

// Column sums over the 32 lanes of a warp of v[0..31] (one row per lane): lane c returns the sum of
// column c.  Halving butterfly: 31 shuffles, fixed order (deterministic).
__device__ __forceinline__ float ndp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const float send = upper ? v[j] : v[j + off];
            const float keep = upper ? v[j + off] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// ---- kernel (3a-0): per point dL/dy -> dL/dz (head gradients), positional encoding, tile maximum ----
// One CTA of 128 threads per tile, full occupancy (the per-point chain of dependent global loads,
// fp64 fixed-point conversion and rotation backward is latency bound and would otherwise sit at the
// head of every tensor-core CTA).  Record per tile (NDP_HGREC floats): hg[128][16], e[128][8]
// (column 7 of e: row 0 = the tile's max |hg|, rows 1.. = db_h = sum_p hg[p]).
__global__ void __launch_bounds__(NDP_TP) ndp_head_grad_kernel(NdpBwdArgs a) {
    __shared__ float red[4];
    __shared__ float hsum[4][NDP_MAX_HEAD];
    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0, tile = blockIdx.x;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const NdpLayout& L = a.lay;
    const int HD = L.head_dim;
    const int warp = tid >> 5, lane = tid & 31;
    float* rec = a.hgbuf + (long long)pair * a.hgbuf_stride + (long long)tile * NDP_HGREC;
    float hgv[16], e0[8];
#pragma unroll
    for (int r = 0; r < 16; ++r) hgv[r] = 0.0f;
    {
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
        if (lane == 0) red[warp] = mx;
        if (a.gx && gp < n) {   // direct part of dL/dx; the tensor-core kernel adds the path through the encoding
            float* gxp = a.gx + (long long)pair * a.gx_stride + (long long)gp * 3;
            gxp[0] = gxd[0]; gxp[1] = gxd[1]; gxp[2] = gxd[2];
        }
        float s, c;
        sincosf(x[0] * L.freq, &s, &c); e0[0] = s; e0[1] = c;
        sincosf(x[1] * L.freq, &s, &c); e0[2] = s; e0[3] = c;
        sincosf(x[2] * L.freq, &s, &c); e0[4] = s; e0[5] = c;
        e0[6] = 0.0f; e0[7] = 0.0f;
    }

    }
    // db_h[r] = sum over the tile's points (fixed order: lanes by butterfly, then warps 0..3)
#pragma unroll
    for (int r = 0; r < NDP_MAX_HEAD; ++r) {
        float sv = hgv[r];
        for (int s = 16; s > 0; s >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, s);
        if (lane == 0) hsum[warp][r] = sv;
    }
    float4* hp = (float4*)(rec + tid * 16);
    hp[0] = make_float4(hgv[0], hgv[1], hgv[2], hgv[3]);   hp[1] = make_float4(hgv[4], hgv[5], hgv[6], hgv[7]);
    hp[2] = make_float4(hgv[8], hgv[9], hgv[10], hgv[11]); hp[3] = make_float4(hgv[12], hgv[13], hgv[14], hgv[15]);
    __syncthreads();
    // column 7 of the e rows carries the tile's scalars: row 0 = max |hg|, row 1 + r = db_h[r]
    if (tid == 0) e0[7] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    else if (tid <= HD) e0[7] = (hsum[0][tid - 1] + hsum[1][tid - 1]) + (hsum[2][tid - 1] + hsum[3][tid - 1]);
    float4* ep = (float4*)(rec + NDP_TP * 16 + tid * 8);
    ep[0] = make_float4(e0[0], e0[1], e0[2], e0[3]); ep[1] = make_float4(e0[4], e0[5], e0[6], e0[7]);
}

__global__ void __launch_bounds__(NDP_BWD_TC_THREADS, 1) ndp_warp_bwd_tc_kernel(NdpBwdArgs a) {
    NDP_DYN_SMEM(smem_raw);
    BwdTcSmem& S = *(BwdTcSmem*)NDP_SMEM_ALIGN(smem_raw, 128);

    // A CTA owns a.tpc consecutive tiles of one pair and accumulates their gradients in TMEM / smem:
    // one partial row per CTA (fixed grouping => deterministic), drained once.
    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0, tile0 = blockIdx.x * a.tpc;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile0 * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const int tiles_all = (n + NDP_TP - 1) / NDP_TP;
    const int ntl = tiles_all - tile0 < a.tpc ? tiles_all - tile0 : a.tpc;     // tiles of this CTA
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const unsigned char* wimg = (const unsigned char*)(a.pack + (long long)pair * a.pack_stride + L.pack_img);
    const int LH = L.hidden, HD = L.head_dim;
    const unsigned char* gact0 = (const unsigned char*)a.act + ((long long)pair * a.act_stride) * 4 +
                                 (long long)tile0 * (LH + 1) * NDP_SET128;
    const long long gact_tile = (long long)(LH + 1) * NDP_SET128;
    float* part = a.partials + (long long)pair * a.partials_stride + (long long)blockIdx.x * a.partial_pitch;
    // thread -> TMEM lane quarter q (hardware: warp % 4), row p = 32 q + lane, column quarter cq
    const int warp = tid >> 5, lane = tid & 31, q = warp & 3, cq = warp >> 2, p = q * 32 + lane;
    const bool wk = tid < 512;                     // workers (epilogues, drains)
    const bool issw = ndp_warp_uniform(warp) == 16; // the issuing warp: one elected lane launches every MMA / TMA copy
#define iss (issw && ndp_elect_one())
    const int RS = NDP_IMG_RS(128), CS = NDP_IMG_CS, RS16 = NDP_IMG_RS(16);
    unsigned hph = 0, wph = 0, mph = 0;
    NDP_T(0);

    if (warp == 0) ndp_tmem_alloc_warp(&S.tmem_slot, 512);
    if (iss) {                  // barriers, and the first operands requested before anything else
        ndp_mbar_init(&S.bar_h, 1); ndp_mbar_init(&S.bar_w, 1); ndp_mbar_init(&S.bar_mma, 1);
        ndp_stage_bulk(S.H, gact0 + (long long)LH * NDP_SET128, NDP_SET128, &S.bar_h);                // h_L of the first tile
        if (LH > 0) ndp_stage_bulk(S.W, wimg + (long long)(LH - 1) * NDP_SET128, NDP_SET128, &S.bar_w);   // W_{L-1}
    }
    // the CTA's delta scale: an exact power of two that brings the largest head gradient of its tiles
    // into [1, 2) (fp16 operand range, see ndp_tc.cuh); undone when the gradients leave TMEM
    const float* rec0 = a.hgbuf + (long long)pair * a.hgbuf_stride + (long long)tile0 * NDP_HGREC;
    float dscale, dinv;
    {
        float mx = 0.0f;
        for (int t = 0; t < ntl; ++t) mx = fmaxf(mx, rec0[(long long)t * NDP_HGREC + NDP_TP * 16 + 7]);
        ndp_pow2_scale(mx, dscale, dinv);
    }

  for (int t = 0; t < ntl; ++t) {
    const int tile = tile0 + t;
    const bool last = t == ntl - 1, acc = t > 0;
    const unsigned char* gact = gact0 + (long long)t * gact_tile;
    const float* rec = rec0 + (long long)t * NDP_HGREC;
    if (!wk) {
    } else if (tid >= 256) {    // head weight image (first tile only; the guard is warp uniform):
      if (t == 0) {             // row r = (tid - 256) / 16, 8-column chunk (tid - 256) & 15
        const int r = (tid - 256) >> 4, c8 = (tid - 256) & 15;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (r < HD) ? __ldg(params + L.head_w[r] + c8 * 8 + j) : 0.0f;   // head rows are not 16-byte aligned
        ndp_store_chunk2(S.HW, NDP_HWIMG, ndp_img_off(r, c8 * 8, RS), v);
      }
    } else {                    // head-gradient / encoding images from the record of ndp_head_grad_kernel
        const int pt = tid >> 1, c8 = tid & 1;
        const float4 h0 = *(const float4*)(rec + pt * 16 + c8 * 8), h1 = *(const float4*)(rec + pt * 16 + c8 * 8 + 4);
        float v[8] = {h0.x * dscale, h0.y * dscale, h0.z * dscale, h0.w * dscale, h1.x * dscale, h1.y * dscale, h1.z * dscale, h1.w * dscale};
        ndp_store_chunk2(S.HG, NDP_IMG16, ndp_img_off(pt, c8 * 8, RS16), v);
        float e[8];
        if (c8 == 0) {
            const float4 e0 = *(const float4*)(rec + NDP_TP * 16 + pt * 8), e1 = *(const float4*)(rec + NDP_TP * 16 + pt * 8 + 4);
            e[0] = e0.x; e[1] = e0.y; e[2] = e0.z; e[3] = e0.w; e[4] = e1.x; e[5] = e1.y; e[6] = 0.0f; e[7] = 0.0f;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) e[j] = 0.0f;
        }
        ndp_store_chunk2(S.E, NDP_IMG16, ndp_img_off(pt, c8 * 8, RS16), e);
    }
    ndp_tc_fence_before();
    ndp_fence_proxy_async();
    __syncthreads();
    ndp_tc_fence_after();
    NDP_T(1);
    const unsigned tmem = S.tmem_slot;
    const unsigned tlane = tmem + ((unsigned)(q * 32) << 16);

    const unsigned id_nn = ndp_idesc_f16(128, 128, 1, 1), id_sm = ndp_idesc_f16(128, 16, 1, 1), id_kn = ndp_idesc_f16(128, 128, 0, 1);
    const NdpUmmaDesc dD_mn = ndp_umma_desc(S.D, RS, CS), dD_k = ndp_umma_desc(S.D, CS, RS);
    const NdpUmmaDesc dH_mn = ndp_umma_desc(S.H, RS, CS), dW_mn = ndp_umma_desc(S.W, RS, CS);
    const NdpUmmaDesc dE = ndp_umma_desc(S.E, RS16, CS), dHG_mn = ndp_umma_desc(S.HG, RS16, CS), dHG_k = ndp_umma_desc(S.HG, CS, RS16);
    const NdpUmmaDesc dHW_mn = ndp_umma_desc(S.HW, RS, CS);

    // The issuing thread launches the MMAs of hidden layer l: delta_l (needs W_l), then dW_l^T (needs h_l)
    auto issue_layer = [&](int l, int b) {
        ndp_mbar_wait(&S.bar_w, wph);                        // W_l
        ndp_tc_fence_after();
        // delta_l[p][i] = sum_o delta[p][o] W_l[o][i]  (runs while h_l is still landing)
        ndp_umma_gemm3(tmem + TM_DH, dD_k, NDP_IMG128, 2 * CS, dW_mn, NDP_IMG128, 2 * RS, 8, id_kn, false);
        ndp_mbar_wait(&S.bar_h, hph);                        // h_l
        ndp_tc_fence_after();
        // dW_l^T[i][o] = sum_p h_l[p][i] delta[p][o]  (lanes = i: the drain is coalesced)
        ndp_umma_gemm3(tmem + TM_DW(b), dH_mn, NDP_IMG128, 2 * RS, dD_mn, NDP_IMG128, 2 * RS, 8, id_nn, acc);
        ndp_umma_commit(&S.bar_mma);
        (void)l;
    };

    // ---- batch 0: raw delta at the top activation (hg W_h), head weight gradients
    if (iss) {
        ndp_umma_gemm3(tmem + TM_DH, dHG_k, NDP_IMG16, 0, dHW_mn, NDP_HWIMG, 0, 1, id_kn, false);
        ndp_mbar_wait(&S.bar_h, hph);                        // h_L has landed
        ndp_tc_fence_after();
        ndp_umma_gemm3(tmem + TM_HDW, dH_mn, NDP_IMG128, 2 * RS, dHG_mn, NDP_IMG16, 2 * RS16, 8, id_sm, acc);
        ndp_umma_commit(&S.bar_mma);
    }
    hph ^= 1;
    __syncthreads();            // the issuer has seen h_L land => visible to everybody
    NDP_T(2);
    ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
    ndp_tc_fence_after();
    NDP_T(3);
    float v[32];
    const int col0 = cq * 32;
    {   // delta_L = raw . relu'(h_L), re-split into the delta images; h_{L-1} replaces h_L meanwhile
        unsigned mask = 0u;
        if (wk) {
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8)
                mask |= ndp_pos_mask8(*(const uint4*)(S.H + ndp_img_off(p, col0 + s8 * 8, RS))) << (s8 * 8);
        }
        __syncthreads();        // all masks extracted, head MMAs retired: h_L is free
        if (iss && LH > 0) ndp_stage_bulk(S.H, gact + (long long)(LH - 1) * NDP_SET128, NDP_SET128, &S.bar_h);
        if (iss && LH == 0 && !last) ndp_stage_bulk(S.H, gact + gact_tile, NDP_SET128, &S.bar_h);   // next tile's h_L
        if (wk) {
            ndp_tmem_ld32(tlane + TM_DH + col0, v);
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
                float u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[s8 * 8 + j] = ((mask >> (s8 * 8 + j)) & 1u) ? v[s8 * 8 + j] : 0.0f; u[j] = v[s8 * 8 + j]; }
                ndp_store_chunk2(S.D, NDP_IMG128, ndp_img_off(p, col0 + s8 * 8, RS), u);
            }
        }
    }
    ndp_tc_fence_before();
    ndp_fence_proxy_async();
    __syncthreads();            // delta_L complete
    ndp_tc_fence_after();
    NDP_T(4);
    if (iss && LH > 0) issue_layer(LH - 1, 0);               // the tensor pipe works while the workers drain
    if (LH > 0) { hph ^= 1; wph ^= 1; }
    // bias gradient of the layer below = column sums of delta (per lane quarter; combined after the next sync)
    if (wk) S.dbred[q][col0 + lane] = ndp_colsum32(v, lane);
    // head weight gradients leave TMEM: dW_h^T[i][r] (lanes = input feature i)
    if (cq == 0 && last) {
        float w[16];
        ndp_tmem_ld16(tlane + TM_HDW, w);
#pragma unroll
        for (int r = 0; r < NDP_MAX_HEAD; ++r)
            if (r < HD) part[L.head_w[r] + p] = w[r] * dinv;
    }

    for (int l = LH - 1; l >= 0; --l) {
        const int b = (LH - 1 - l) & 1, tb = 8 + 8 * (LH - 1 - l);
        NDP_T(tb + 1);
        __syncthreads();        // (A) the issuer has seen h_l land => visible to everybody; dbred complete
        if (tid < NDP_W) {      // db_l[o] = sum_p delta_{l+1}[p][o], quarters in fixed order, tiles in order
            const float val = (S.dbred[0][tid] + S.dbred[1][tid]) + (S.dbred[2][tid] + S.dbred[3][tid]);
            const float tot = acc ? S.dbacc[l][tid] + val : val;
            S.dbacc[l][tid] = tot;
            if (last) part[L.off_b[l] + tid] = tot * dinv;
        }
        // relu' mask of h_l for this thread's row / column quarter, before h_{l-1} replaces h_l
        unsigned mask = 0u;
        if (wk) {
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8)
                mask |= ndp_pos_mask8(*(const uint4*)(S.H + ndp_img_off(p, col0 + s8 * 8, RS))) << (s8 * 8);
        }
        ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
        ndp_tc_fence_after();
        NDP_T(tb + 2);
        __syncthreads();        // (B) all masks extracted, MMAs retired: h_l, W_l and the delta images are free
        if (iss && l > 0) {
            ndp_stage_bulk(S.H, gact + (long long)(l - 1) * NDP_SET128, NDP_SET128, &S.bar_h);
            ndp_stage_bulk(S.W, wimg + (long long)(l - 1) * NDP_SET128, NDP_SET128, &S.bar_w);
        }
        if (iss && l == 0 && !last) {       // both buffers are free for the rest of this tile: prefetch the next one
            ndp_stage_bulk(S.H, gact + gact_tile + (long long)LH * NDP_SET128, NDP_SET128, &S.bar_h);
            ndp_stage_bulk(S.W, wimg + (long long)(LH - 1) * NDP_SET128, NDP_SET128, &S.bar_w);
        }
        if (wk) {   // delta_l = raw . relu'(h_l), re-split, delta images updated in place
            ndp_tmem_ld32(tlane + TM_DH + col0, v);
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
                float u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[s8 * 8 + j] = ((mask >> (s8 * 8 + j)) & 1u) ? v[s8 * 8 + j] : 0.0f; u[j] = v[s8 * 8 + j]; }
                ndp_store_chunk2(S.D, NDP_IMG128, ndp_img_off(p, col0 + s8 * 8, RS), u);
            }
        }
        ndp_tc_fence_before();
        ndp_fence_proxy_async();
        __syncthreads();        // (C) delta_l complete
        ndp_tc_fence_after();
        NDP_T(tb + 3);
        if (iss) {
            if (l > 0) issue_layer(l - 1, b ^ 1);
            else {
                // input layer: dW_in[o][0..5] = sum_p delta_0[p][o] E[p][0..5]
                ndp_umma_gemm3(tmem + TM_DIN, dD_mn, NDP_IMG128, 2 * RS, dE, NDP_IMG16, 2 * RS16, 8, id_sm, acc);
                ndp_umma_commit(&S.bar_mma);
            }
        }
        if (l > 0) { hph ^= 1; wph ^= 1; }
        if (wk) S.dbred[q][col0 + lane] = ndp_colsum32(v, lane);     // -> db_{l-1} (db_in for l == 0)
        // dW_l^T leaves TMEM while the next batch of MMAs runs: lane = input feature i, column = output o,
        // so every store instruction of a warp writes 128 contiguous bytes of the canonical [o][i] block
        if (wk && last) {
            float w[32];
            ndp_tmem_ld32(tlane + TM_DW(b) + col0, w);
            float* dst = part + L.off_w[l] + col0 * NDP_W + p;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * NDP_W] = w[j] * dinv;
        }
        NDP_T(tb + 4);
    }

    if (LH == 0 && iss) {
        ndp_umma_gemm3(tmem + TM_DIN, dD_mn, NDP_IMG128, 2 * RS, dE, NDP_IMG16, 2 * RS16, 8, id_sm, acc);
        ndp_umma_commit(&S.bar_mma);
    }
    __syncthreads();            // dbred of delta_0 complete
    if (tid < NDP_W) {
        const float val = (S.dbred[0][tid] + S.dbred[1][tid]) + (S.dbred[2][tid] + S.dbred[3][tid]);
        const float tot = acc ? S.dbacc[LH][tid] + val : val;
        S.dbacc[LH][tid] = tot;
        if (last) part[L.off_b_in + tid] = tot * dinv;
    }
    ndp_mbar_wait(&S.bar_mma, mph); mph ^= 1;
    ndp_tc_fence_after();
    if (cq == 0 && last) {
        float w[16];
        ndp_tmem_ld16(tlane + TM_DIN, w);
        float* dst = part + L.off_w_in + p * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c) dst[c] = w[c] * dinv;
    }
    if (a.gx && tid < NDP_TP) {     // optional dL/dx: + the path through the positional encoding
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
            const float* xp = a.x + (long long)pair * a.x_stride + (long long)gp * 3;
            float* gxp = a.gx + (long long)pair * a.gx_stride + (long long)gp * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float sn, cs;
                sincosf(__ldg(xp + d) * L.freq, &sn, &cs);
                gxp[d] += L.freq * (cs * de[2 * d] - sn * de[2 * d + 1]);
            }
        }
    }
    NDP_T(62);
    ndp_tc_fence_before();
    __syncthreads();            // end of the tile: images, dbred and the DH / DIN accumulators are free again
    ndp_tc_fence_after();
  }
    if (tid < HD) {             // db_h: per-tile sums of the pre-kernel, tiles in order
        float sv = 0.0f;
        for (int t = 0; t < ntl; ++t) sv += rec0[(long long)t * NDP_HGREC + NDP_TP * 16 + (1 + tid) * 8 + 7];
        part[L.head_b[tid]] = sv;
    }
    NDP_T(63);
    if (warp == 0) ndp_tmem_dealloc(S.tmem_slot, 512);
#undef iss
}

// Tiles whose gradients one CTA accumulates (= tiles per partial row).  `forced` > 0 (ndp_solver_cfg::
// tiles_per_bwd_cta / ndp_set_layer_tuning) wins; otherwise a function of the cloud size only -- never of
// the batch: the summation grouping must not depend on what a pair is batched with.  The old kernel keeps
// per-layer-parity dW accumulators, so it accumulates over tiles only for at most two hidden layers.
int ndp_bwd_tc_tiles_per_cta(int hidden, int n, int forced) {
    if (hidden > 2) return 1;
    if (forced > 0) return forced > 16 ? 16 : forced;
    // at least 16 CTAs per pair, at most 8 tiles per CTA.  8192 points -> 4.
    const int tiles = (n + NDP_TP - 1) / NDP_TP;
    int tpc = 1;
    while (tpc < 8 && tiles / (2 * tpc) >= 16) tpc *= 2;
    return tpc;
}
// hidden == 2 (depth 3, the reference's configuration): the recomputing kernel of ndp_warp_bwd_rc.cu, which
// needs no saved activations; every other depth: the kernel above, fed by the forward kernel's saved images.
bool ndp_tc_recompute(int hidden) { return hidden == 2; }
void ndp_launch_bwd_rc_main(const NdpBwdArgs& b, int grid_x, cudaStream_t s);
void ndp_launch_bwd_tc(const NdpBwdArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    const int tiles = (a.n + NDP_TP - 1) / NDP_TP;
    NDP_LAUNCH_PRIO(1, ndp_head_grad_kernel, dim3(tiles, a.npairs), dim3(NDP_TP), 0, s, a);
    NdpBwdArgs b = a;
    b.tpc = ndp_bwd_tc_tiles_per_cta(a.lay.hidden, a.n, a.tpc);
    const int gx = (tiles + b.tpc - 1) / b.tpc;
    if (ndp_tc_recompute(a.lay.hidden)) { ndp_launch_bwd_rc_main(b, gx, s); return; }
    NDP_LAUNCH(ndp_warp_bwd_tc_kernel, dim3(gx, a.npairs), dim3(NDP_BWD_TC_THREADS), ndp_bwd_tc_smem_bytes(), s, b);
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
