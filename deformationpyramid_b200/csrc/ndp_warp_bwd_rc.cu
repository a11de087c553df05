// Kernel (3a), recomputing tensor-core version (depth 3 = two hidden layers, the reference's
// configuration): backward of kernel (1) WITHOUT saved activations.
//
// Replaces the autograd backward of NDPLayer.forward (model/nets.py:111-140) that the reference
// runs in loss.backward() (model/registration.py:236).
//
// Why recompute: streaming the saved activations (12.6 MB per 8192-point pair and iteration, written
// by the forward kernel and read back here) plus one 64 KB weight image set per layer and tile costs
// 320 KB of L2 -> shared-memory traffic per 128 points -- more than the tensor pipe's time for those
// points leaves room for -- and ties up 128 KB of shared memory per tile, so only ONE dependency chain
// fits per SM.  Here the activations are rebuilt from x (12 bytes per point) on the tensor cores
// (+53 % MMAs), nothing but x, z and dL/dy is read per point, and the work is laid out so that TWO
// independent chains (64-point half tiles) share an SM: one chain's epilogue runs under the other
// chain's MMAs.
//
// Everything is TRANSPOSED relative to kernel (1): TMEM lanes / image rows are FEATURES, columns are
// POINTS.  An activation / delta image is [128 features][64 points] (fp16 hi + lo, core-matrix layout
// of ndp_tc.cuh with row pitch NDP_RS64), which makes
//   * the bias a per-thread scalar and the bias gradient a per-thread sum (no shuffles),
//   * every epilogue store a 16-byte vector (8 consecutive points of one feature),
//   * ONE image serve as the MN-major B operand of the layer products (N = points) and as the K-major
//     A / B operand of the weight-gradient products (K = points).
// Per half tile (chain), with X, Y the chain's two image buffers and WB the shared weight buffer:
//   h0  -> X    CUDA cores: relu(W_in e + b_in), 6 FMAs per element (also rebuilt later, see below)
//   F1: acc = W_0 X           -> Y = h1 = relu(acc + b_0)                 A = W_0 (K-major), B = X (MN-major)
//   F2: acc = W_1 Y           -> X = h2 = relu(acc + b_1)
//   B2: acc = W_h^T hg^T,  dW_h^T += X hg          -> X = delta2 = acc . relu'(h2)   (in place)
//   B1: acc = W_1^T X,     dW_1^T += Y X^T         -> Y = delta1 = acc . relu'(h1),  X = h0 again
//   B0: dW_0^T += X Y^T,   acc = W_0^T Y           -> X = delta0 = acc . relu'(h0)
//   Bin: dW_in += X E
// h0 is rebuilt because only two buffers per chain fit next to the weight buffer (2 x 2 x 32 KB + 64 KB).
// relu' of every activation stays in a register of the thread that produced it (the same thread handles the
// same (feature, points) cell in every epilogue).  X and Y swap roles from tile to tile and the encoding
// image is double buffered, so the chain never waits for the tile's last products (Bin).
// The weight buffer holds W_0 during B0 / Bin / F1 and W_1 during F2 / B2 / B1: two TMA loads of 64 KB per
// 128 points (the old kernel: two weight sets + three activation sets).  Gradients accumulate in TMEM
// over all tiles of the CTA (dW_1^T, dW_0^T: 128 columns each, dW_h^T, dW_in: 16 each) and leave once.
// A 17th warp's elected lane issues every MMA and TMA copy from a static schedule, alternating between
// the chains.  The 16 worker warps are NOT split between the chains: all of them do chain 0's epilogue
// (16 of its 64 columns each) while chain 1's MMAs run, then chain 1's while chain 0's next MMAs run, so an
// epilogue takes half as long as with eight warps and the tensor pipe and the FP32 pipes stay busy
// together.  Workers and issuer meet through mbarriers (operands ready: 16 warp arrivals; MMAs retired:
// tcgen05.commit; one warp polls, the others sleep at a named barrier).
// Deltas are carried multiplied by the CTA's power-of-two scale (fp16 range), as in ndp_warp_bwd_tc.cu.
// No atomics: one partial row per CTA, fixed summation order => bit-reproducible.
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

#ifndef NDP_EMU
__device__ unsigned long long ndp_dbg_rc[64];
__device__ unsigned long long ndp_dbg_rc_sum[2];      // sum of CTA durations (ns), CTAs: since the last read
#define NDP_TR(i) do { if ((tid & 255) == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ndp_dbg_rc[(i) + 24 * (tid >> 8)] = t_; } } while (0)
#define NDP_TI(i) do { if (blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); ndp_dbg_rc[i] = t_; } } while (0)
#else
#define NDP_TR(i) do {} while (0)
#define NDP_TI(i) do {} while (0)
#endif

#define NDP_RC_THREADS 544                 // 2 chains x 8 warps + the issuing warp
#define NDP_HP 64                          // points per half tile (one chain)
#define NDP_RS64 NDP_IMG_RS(64)            // 1024: bytes between 8-feature row groups of a [128][64] image
#define NDP_IMG64 NDP_IMG_BYTES(64)        // 16384
#define NDP_SET64 (2 * NDP_IMG64)          // 32768: hi + lo
#define NDP_RS16 NDP_IMG_RS(16)            // 256
#define NDP_IMG16F NDP_IMG_BYTES(16)       // [128 features][16]: 4096
#define NDP_IMG16H (8 * NDP_RS16)          // [64 points][16]: 2048

struct BwdRcSmem {
    unsigned char WB[NDP_SET128];          // hidden weight hi/lo images of the current step (TMA)
    unsigned char X[2][NDP_SET64];         // per chain: h0 -> h2 -> delta2 -> h0 -> delta0
    unsigned char Y[2][NDP_SET64];         // per chain: h1 -> delta1
    unsigned char HG[2][2 * NDP_IMG16H];   // per chain [64 points][16]: scaled mlp_scale * dL/dz; at the very end: bias-gradient scratch
    unsigned char E[2][2 * NDP_IMG16H];    // per chain [64 points][16]: cols 0..5 positional encoding, rest 0
    unsigned char HWT[2 * NDP_IMG16F];     // [128 features][16 head rows]: head weights transposed, rows >= head_dim zero
    float ef[2][6][NDP_HP];                // per chain: fp32 positional encoding, one row per component (h0 on the CUDA cores)
    float win[7][NDP_W];                   // input layer: rows 0..5 = W_in[:, k], row 6 = b_in (fp32; registers are too scarce to hold them)
    NdpMbar bar_w, bar_wfree, bar_ready[2][2], bar_mma[2], bar_fin;   // ready: [chain][signal parity]
    unsigned tmem_slot, pad[3];
};
size_t ndp_bwd_rc_smem_bytes() { return sizeof(BwdRcSmem) + 128; }

// TMEM columns
#define RC_ACC(c) (64u * (unsigned)(c))    // per chain: layer accumulator [128 features][64 points]
#define RC_DW1 128u                        // dW_1^T [128 i][128 o]
#define RC_DW0 256u                        // dW_0^T
#define RC_DWH 384u                        // dW_h^T [128 i][16 head rows]
#define RC_DWIN 400u                       // dW_in  [128 o][16] (cols 0..5)

// 544 threads = 17 warps are allocated as 20 (groups of four): 65 536 / 640 = 102 -> 96 registers per thread is the ceiling.
__global__ void __launch_bounds__(NDP_RC_THREADS, 1) ndp_warp_bwd_rc_kernel(NdpBwdArgs a) {
    NDP_DYN_SMEM(smem_raw);
    // aligned by pointer arithmetic on the shared array itself: a round trip through an integer would lose the address
    // space and turn every shared-memory access of the kernel into a generic LD / ST
    BwdRcSmem& S = *(BwdRcSmem*)NDP_SMEM_ALIGN(smem_raw, 128);

    const int tid = threadIdx.x, pair = blockIdx.y + a.pair0, tile0 = blockIdx.x * a.tpc;
    const int n = a.counts ? a.counts[pair] : a.n;
    if (tile0 * NDP_TP >= n) return;
    if (a.state && a.state[pair].stopped) return;
    const int tiles_all = (n + NDP_TP - 1) / NDP_TP;
    const int ntl = tiles_all - tile0 < a.tpc ? tiles_all - tile0 : a.tpc;     // tiles of this CTA
    const NdpLayout& L = a.lay;
    const float* params = a.params + (long long)pair * a.params_stride;
    const unsigned char* wimg = (const unsigned char*)(a.pack + (long long)pair * a.pack_stride + L.pack_img);
    const int HD = L.head_dim;
    float* part = a.partials + (long long)pair * a.partials_stride + (long long)blockIdx.x * a.partial_pitch;
    const int warp = tid >> 5, lane = tid & 31;
    const bool issw = ndp_warp_uniform(warp) == 16;
    // worker thread: TMEM lane quarter q (hardware: warp % 4), feature f = its TMEM lane / image row, column group cg:
    // points [16 cg, 16 cg + 16) of EACH chain's 64.  c / ct: the thread's role in the per-tile staging of the records
    const int c = (warp >> 3) & 1, q = warp & 3, cg = (warp >> 2) & 3, f = q * 32 + lane, ct = tid & 255;
    const int RS = NDP_IMG_RS(128), CS = NDP_IMG_CS;
    NDP_TR(0);
#ifndef NDP_EMU
    unsigned long long t_cta0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cta0));
#endif

    if (warp == 0) ndp_tmem_alloc_warp(&S.tmem_slot, 512);
    if (issw && ndp_elect_one()) {
        ndp_mbar_init(&S.bar_w, 1); ndp_mbar_init(&S.bar_wfree, 1);
        ndp_mbar_init(&S.bar_ready[0][0], 16); ndp_mbar_init(&S.bar_ready[0][1], 16);
        ndp_mbar_init(&S.bar_ready[1][0], 16); ndp_mbar_init(&S.bar_ready[1][1], 16);
        ndp_mbar_init(&S.bar_mma[0], 1); ndp_mbar_init(&S.bar_mma[1], 1); ndp_mbar_init(&S.bar_fin, 1);
        ndp_stage_bulk(S.WB, wimg, NDP_SET128, &S.bar_w);                      // W_0 for the first F1
    }
    // h0 role of a worker thread (gen_h0 below): 8 consecutive points (group hp8) of TWO features fA, fB = fA + 64.  All lanes
    // of a warp share hp8 (their encoding loads are one broadcast per warp), a warp's lanes hold 32 consecutive features
    // (conflict-free loads of the input-layer weights from S.win) and a quarter warp's 16-byte image stores fill all 32 banks.
    const int hp8 = (tid >> 6) & 7, fA = tid & 63, fB = fA + 64;
    float b_0 = 0.0f, b_1 = 0.0f;
    if (!issw) {
        // transposed head weight image: row i = ct & 127, 8-column chunk ct >> 7 (head rows 8 chunk .. 8 chunk + 7)
        const int i = ct & 127, c8 = ct >> 7;
        if (c == 0) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int r = c8 * 8 + j; v[j] = (r < HD) ? __ldg(params + L.head_w[r < NDP_MAX_HEAD ? r : 0] + i) : 0.0f; }
            ndp_store_chunk2(S.HWT, NDP_IMG16F, ndp_img_off(i, c8 * 8, NDP_RS16), v);
        }
        {   // columns 8..15 of the encoding images stay zero: written once
            float z8[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (ct < 64) ndp_store_chunk2(S.E[c], NDP_IMG16H, ndp_img_off(ct, 8, NDP_RS16), z8);
        }
        for (int idx = tid; idx < 7 * NDP_W; idx += 512) {
            const int k = idx >> 7, o = idx & (NDP_W - 1);
            S.win[k][o] = k < 6 ? __ldg(params + L.off_w_in + o * 6 + k) : __ldg(params + L.off_b_in + o);
        }
        b_0 = __ldg(params + L.off_b[0] + f); b_1 = __ldg(params + L.off_b[1] + f);
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
    ndp_tc_fence_before();
    ndp_fence_proxy_async();
    __syncthreads();
    ndp_tc_fence_after();
    const unsigned tmem = S.tmem_slot;
    NDP_TR(1);

    if (issw) {
      if (ndp_elect_one()) {
        // ================================================================ the issuing thread
        // The chains do not wait after their last signal of a tile (-> Bin), so a warp may make its NEXT arrival before the
        // other warps have made this one: consecutive signals therefore alternate between two barriers per chain (a warp
        // can run at most one signal ahead: its next wait needs the issuer to have consumed both).
        unsigned rph[2][2] = {{0u, 0u}, {0u, 0u}}, rk[2] = {0u, 0u}, wph = 0u, fph = 0u;
        const unsigned id_f = ndp_idesc_f16(128, 64, 0, 1);      // A K-major (weights), B MN-major (image, N = points)
        const unsigned id_h = ndp_idesc_f16(128, 64, 0, 0);      // A K-major (HWT), B K-major (HG rows = points)
        const unsigned id_b = ndp_idesc_f16(128, 64, 1, 1);      // A MN-major (weights transposed), B MN-major
        const unsigned id_w = ndp_idesc_f16(128, 128, 0, 0);     // both K-major over the points
        const unsigned id_s = ndp_idesc_f16(128, 16, 0, 1);      // A K-major over the points, B = [points][16] MN-major
        const NdpUmmaDesc dW_k = ndp_umma_desc(S.WB, CS, RS), dW_mn = ndp_umma_desc(S.WB, RS, CS);
        const NdpUmmaDesc dHWT = ndp_umma_desc(S.HWT, CS, NDP_RS16);
#define RC_READY(cc) do { const unsigned k_ = rk[cc] & 1u; ndp_mbar_wait(&S.bar_ready[cc][k_], rph[cc][k_]); rph[cc][k_] ^= 1u; rk[cc] += 1u; \
                          ndp_tc_fence_after(); } while (0)
#define RC_WAIT_W() do { ndp_mbar_wait(&S.bar_w, wph); wph ^= 1u; ndp_tc_fence_after(); } while (0)
#define RC_RELOAD_W(layer) do { ndp_mbar_wait(&S.bar_wfree, fph); fph ^= 1u; \
                                ndp_stage_bulk(S.WB, wimg + (long long)(layer) * NDP_SET128, NDP_SET128, &S.bar_w); } while (0)
        for (int t = 0; t < ntl; ++t) {
            // X and Y swap roles from tile to tile: A holds h0 -> h2 -> delta2 -> h0 -> delta0, B holds h1 -> delta1
            unsigned char* const A0 = (t & 1) ? S.Y[0] : S.X[0];
            unsigned char* const A1 = (t & 1) ? S.Y[1] : S.X[1];
            unsigned char* const B0 = (t & 1) ? S.X[0] : S.Y[0];
            unsigned char* const B1 = (t & 1) ? S.X[1] : S.Y[1];
            const bool first = t == 0;
            // ---- F1: acc = W_0 h0
            for (int cc = 0; cc < 2; ++cc) {
                RC_READY(cc);
                if (first && cc == 0) RC_WAIT_W();
                NDP_TI(40 + cc * 2);
                ndp_umma_gemm3_ar(tmem + RC_ACC(cc), dW_k, NDP_IMG128, 2 * CS, ndp_umma_desc(cc ? A1 : A0, NDP_RS64, CS), NDP_IMG64, 2 * NDP_RS64, 8, id_f, false);
                ndp_umma_commit(&S.bar_mma[cc]);
                NDP_TI(41 + cc * 2);
            }
            ndp_umma_commit(&S.bar_wfree);
            RC_RELOAD_W(1);
            NDP_TI(44);
            // ---- F2: acc = W_1 h1
            for (int cc = 0; cc < 2; ++cc) {
                RC_READY(cc);
                NDP_TI(45 + cc * 3);
                if (cc == 0) RC_WAIT_W();
                NDP_TI(46 + cc * 3);
                ndp_umma_gemm3_ar(tmem + RC_ACC(cc), dW_k, NDP_IMG128, 2 * CS, ndp_umma_desc(cc ? B1 : B0, NDP_RS64, CS), NDP_IMG64, 2 * NDP_RS64, 8, id_f, false);
                ndp_umma_commit(&S.bar_mma[cc]);
                NDP_TI(47 + cc * 3);
            }
            // ---- B2: acc = W_h^T hg^T;  dW_h^T += h2 hg
            for (int cc = 0; cc < 2; ++cc) {
                RC_READY(cc);
                const bool acc = !(first && cc == 0);
                ndp_umma_gemm3_ar(tmem + RC_ACC(cc), dHWT, NDP_IMG16F, 0, ndp_umma_desc(S.HG[cc], CS, NDP_RS16), NDP_IMG16H, 0, 1, id_h, false);
                ndp_umma_gemm3_ar(tmem + RC_DWH, ndp_umma_desc(cc ? A1 : A0, CS, NDP_RS64), NDP_IMG64, 2 * CS,
                               ndp_umma_desc(S.HG[cc], NDP_RS16, CS), NDP_IMG16H, 2 * NDP_RS16, 4, id_s, acc);
                ndp_umma_commit(&S.bar_mma[cc]);
            }
            // ---- B1: acc = W_1^T delta2;  dW_1^T += h1 delta2^T
            for (int cc = 0; cc < 2; ++cc) {
                RC_READY(cc);
                const bool acc = !(first && cc == 0);
                NDP_TI(51 + cc * 3);
                ndp_umma_gemm3_ar(tmem + RC_ACC(cc), dW_mn, NDP_IMG128, 2 * RS, ndp_umma_desc(cc ? A1 : A0, NDP_RS64, CS), NDP_IMG64, 2 * NDP_RS64, 8, id_b, false);
                if (cc == 1) ndp_umma_commit(&S.bar_wfree);       // W_1 is free once both chains' products have retired
                NDP_TI(52 + cc * 3);
                ndp_umma_gemm3_ar(tmem + RC_DW1, ndp_umma_desc(cc ? B1 : B0, CS, NDP_RS64), NDP_IMG64, 2 * CS,
                               ndp_umma_desc(cc ? A1 : A0, CS, NDP_RS64), NDP_IMG64, 2 * CS, 4, id_w, acc);
                ndp_umma_commit(&S.bar_mma[cc]);
                NDP_TI(53 + cc * 3);
            }
            RC_RELOAD_W(0);
            NDP_TI(57);
            // ---- B0: dW_0^T += h0 delta1^T (needs no weights: runs while W_0 lands);  acc = W_0^T delta1
            for (int cc = 0; cc < 2; ++cc) {
                RC_READY(cc);
                const bool acc = !(first && cc == 0);
                NDP_TI(58 + cc * 2);
                ndp_umma_gemm3_ar(tmem + RC_DW0, ndp_umma_desc(cc ? A1 : A0, CS, NDP_RS64), NDP_IMG64, 2 * CS,
                               ndp_umma_desc(cc ? B1 : B0, CS, NDP_RS64), NDP_IMG64, 2 * CS, 4, id_w, acc);
                if (cc == 0) RC_WAIT_W();
                ndp_umma_gemm3_ar(tmem + RC_ACC(cc), dW_mn, NDP_IMG128, 2 * RS, ndp_umma_desc(cc ? B1 : B0, NDP_RS64, CS), NDP_IMG64, 2 * NDP_RS64, 8, id_b, false);
                ndp_umma_commit(&S.bar_mma[cc]);
                NDP_TI(59 + cc * 2);
            }
            // ---- Bin: dW_in += delta0 E.  Nobody waits for these: the next commit (or the final one) covers them
            for (int cc = 0; cc < 2; ++cc) {
                RC_READY(cc);
                const bool acc = !(first && cc == 0);
                ndp_umma_gemm3_ar(tmem + RC_DWIN, ndp_umma_desc(cc ? A1 : A0, CS, NDP_RS64), NDP_IMG64, 2 * CS,
                               ndp_umma_desc(S.E[cc], NDP_RS16, CS), NDP_IMG16H, 2 * NDP_RS16, 4, id_s, acc);
            }
        }
        ndp_umma_commit(&S.bar_fin);
#undef RC_READY
#undef RC_WAIT_W
#undef RC_RELOAD_W
      }
    } else {
        // ================================================================ the 16 worker warps
        unsigned mph[2] = {0u, 0u};
        const int col0 = 16 * cg;
        const unsigned tlq = tmem + ((unsigned)(q * 32) << 16);
        float dbs0 = 0.0f, dbs1 = 0.0f, dbs2 = 0.0f;            // sum over points of delta0 / delta1 / delta2 (this thread's feature, its columns, both chains)

        // "my part of chain cc's next operands is in shared memory and I have finished reading its accumulator": one
        // arrival per warp; consecutive signals of a chain alternate between its two barriers (see the issuer)
        unsigned sk[2] = {0u, 0u};
        auto signal_ready = [&](int cc) {
            ndp_fence_proxy_async();
            ndp_tc_fence_before();
            __syncwarp();
            if (lane == 0) ndp_mbar_arrive(&S.bar_ready[cc][sk[cc] & 1u]);
            sk[cc] += 1u;
        };
        // one warp polls the barrier, the other fifteen sleep at a named barrier (a polling warp costs issue slots and
        // shared-memory bandwidth every ~100 cycles)
        auto wait_mma = [&](int cc) {
            if (warp == 0) ndp_mbar_wait(&S.bar_mma[cc], mph[cc]);
            mph[cc] ^= 1u;
            ndp_group_sync(1, 512);
            ndp_tc_fence_after();
        };
        // h0 = relu(W_in e + b_in) (nets.py:114, 164-177) for this thread's two features and its 8 points of chain cc -> dst
        // image.  The encodings are the same for every lane, and a broadcast LDS.128 still costs four cycles of the SM's
        // shared-memory pipe (the resource this kernel is bound by, next to the MMAs' operand fetch): two features x eight
        // points per thread need 12 such loads per warp where one feature x sixteen points needed 32.  Two points per
        // FFMA2, same accumulation order as a scalar chain.  relu' is NOT kept: the backward epilogue that needs it
        // (delta0) reads it back from the image ("hi > 0"), which ndp_relu_img makes exact.
        auto gen_h0 = [&](unsigned char* dst, int cc) {
            NdpF2 aA[4], aB[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { aA[j] = ndp_f2_bcast(S.win[6][fA]); aB[j] = ndp_f2_bcast(S.win[6][fB]); }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const float4 e0 = *(const float4*)&S.ef[cc][k][8 * hp8], e1 = *(const float4*)&S.ef[cc][k][8 * hp8 + 4];
                const NdpF2 p[4] = {ndp_f2_make(e0.x, e0.y), ndp_f2_make(e0.z, e0.w), ndp_f2_make(e1.x, e1.y), ndp_f2_make(e1.z, e1.w)};
                const NdpF2 wa = ndp_f2_bcast(S.win[k][fA]), wb = ndp_f2_bcast(S.win[k][fB]);
#pragma unroll
                for (int j = 0; j < 4; ++j) { aA[j] = ndp_f2_fma(wa, p[j], aA[j]); aB[j] = ndp_f2_fma(wb, p[j], aB[j]); }
            }
            float u[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { ndp_f2_get(aA[j], u[2 * j], u[2 * j + 1]); u[2 * j] = ndp_relu_img(u[2 * j]); u[2 * j + 1] = ndp_relu_img(u[2 * j + 1]); }
            ndp_store_chunk2(dst, NDP_IMG64, ndp_img_off(fA, 8 * hp8, NDP_RS64), u);
#pragma unroll
            for (int j = 0; j < 4; ++j) { ndp_f2_get(aB[j], u[2 * j], u[2 * j + 1]); u[2 * j] = ndp_relu_img(u[2 * j]); u[2 * j + 1] = ndp_relu_img(u[2 * j + 1]); }
            ndp_store_chunk2(dst, NDP_IMG64, ndp_img_off(fB, 8 * hp8, NDP_RS64), u);
        };
        // relu'(h0) of this thread's epilogue cell (feature f, points col0 .. col0 + 15 of chain cc) from the h0 image
        auto h0_mask = [&](const unsigned char* img) -> unsigned {
            return ndp_pos_mask8(ndp_lds128(img + ndp_img_off(f, col0, NDP_RS64))) |
                   (ndp_pos_mask8(ndp_lds128(img + ndp_img_off(f, col0 + 8, NDP_RS64))) << 8);
        };
        // forward epilogue of chain cc: dst = relu(acc + bias) re-split into the image; returns relu' as a bit mask
        auto epi_fwd = [&](unsigned char* dst, int cc, float bias) -> unsigned {
            float v[16];
            ndp_tmem_ld16(tlq + RC_ACC(cc) + (unsigned)col0, v);
            unsigned mask = 0u;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float u[8];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    float x0, x1;
                    ndp_f2_get(ndp_f2_add(ndp_f2_make(v[8 * k + j], v[8 * k + j + 1]), ndp_f2_bcast(bias)), x0, x1);     // FADD2
                    mask |= (x0 > 0.0f ? 1u : 0u) << (8 * k + j);
                    mask |= (x1 > 0.0f ? 1u : 0u) << (8 * k + j + 1);
                    u[j] = fmaxf(x0, 0.0f); u[j + 1] = fmaxf(x1, 0.0f);
                }
                ndp_store_chunk2(dst, NDP_IMG64, ndp_img_off(f, col0 + 8 * k, NDP_RS64), u);
            }
            return mask;
        };
        // backward epilogue of chain cc: buf <- delta = acc . relu'(h) (mask from the epilogue that produced h); returns the sum
        auto epi_bwd = [&](unsigned char* buf, int cc, unsigned mask) -> float {
            float v[16];
            ndp_tmem_ld16(tlq + RC_ACC(cc) + (unsigned)col0, v);
            NdpF2 sum2 = ndp_f2_bcast(0.0f);            // even / odd columns (fixed order)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) u[j] = ((mask >> (8 * k + j)) & 1u) ? v[8 * k + j] : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j += 2) sum2 = ndp_f2_add(sum2, ndp_f2_make(u[j], u[j + 1]));
                ndp_store_chunk2(buf, NDP_IMG64, ndp_img_off(f, col0 + 8 * k, NDP_RS64), u);
            }
            float s0, s1;
            ndp_f2_get(sum2, s0, s1);
            return s0 + s1;
        };
        // this thread's share of the record of ndp_head_grad_kernel for one tile (staging role: chain c = tid / 256,
        // ct = tid % 256): ct < 128: 8 head gradients of point ct / 2 of half c; 128 <= ct < 192: the encoding of point
        // ct - 128 of half c.  Loaded one tile ahead (global latency off the critical path).
        float4 r0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), r1 = r0;
        auto load_rec = [&](int t) {
            const float* rec = rec0 + (long long)t * NDP_HGREC;
            if (ct < 128) {
                const float* hp = rec + (long long)c * NDP_HP * 16 + (ct >> 1) * 16 + (ct & 1) * 8;
                r0 = *(const float4*)hp; r1 = *(const float4*)(hp + 4);
            } else if (ct < 192) {
                const float* ep = rec + NDP_TP * 16 + (long long)c * NDP_HP * 8 + (ct - 128) * 8;
                r0 = *(const float4*)ep; r1 = *(const float4*)(ep + 4);
            }
        };
        load_rec(0);

        for (int t = 0; t < ntl; ++t) {
            const int tile = tile0 + t;
            // X and Y are adjacent members: the role swap is an OFFSET on one shared-memory object, so the compiler keeps the
            // address space (a select between two pointers turns the epilogues' STS / LDS into generic ST / LD)
            const int swp = (t & 1) * (int)sizeof(S.X);
            unsigned char* const A[2] = {S.X[0] + swp, S.X[1] + swp};     // h0 -> h2 -> delta2 -> h0 -> delta0
            unsigned char* const B[2] = {S.Y[0] - swp, S.Y[1] - swp};     // h1 -> delta1
            NDP_TR(2);
            // head-gradient images and fp32 encodings of both half tiles (HG: its readers, B2 of the previous tile, retired
            // long ago; A = the previous tile's delta1, dead since B0).  The encoding IMAGES are still being read by the
            // previous tile's Bin products: they are rewritten after the first wait for F1 below, whose commit covers them.
            if (ct < 128) {
                float v[8] = {r0.x * dscale, r0.y * dscale, r0.z * dscale, r0.w * dscale, r1.x * dscale, r1.y * dscale, r1.z * dscale, r1.w * dscale};
                ndp_store_chunk2(S.HG[c], NDP_IMG16H, ndp_img_off(ct >> 1, (ct & 1) * 8, NDP_RS16), v);
            } else if (ct < 192) {
                const int pt = ct - 128;
                S.ef[c][0][pt] = r0.x; S.ef[c][1][pt] = r0.y; S.ef[c][2][pt] = r0.z; S.ef[c][3][pt] = r0.w;
                S.ef[c][4][pt] = r1.x; S.ef[c][5][pt] = r1.y;
            }
            ndp_group_sync(1, 512);         // ef complete
            NDP_TR(13);
            unsigned m1[2], m2[2];
            gen_h0(A[0], 0); NDP_TR(14); signal_ready(0); NDP_TR(15);
            gen_h0(A[1], 1); signal_ready(1);       // -> F1
            NDP_TR(3);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                wait_mma(cc);
                if (cc == 0) {
                    NDP_TR(4);
                    if (ct >= 128 && ct < 192) {
                        float e[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, 0.0f, 0.0f};
                        ndp_store_chunk2(S.E[c], NDP_IMG16H, ndp_img_off(ct - 128, 0, NDP_RS16), e);
                    }
                    if (t + 1 < ntl) load_rec(t + 1);
                }
                m1[cc] = epi_fwd(B[cc], cc, b_0);
                signal_ready(cc);           // -> F2
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) { wait_mma(cc); if (cc == 0) NDP_TR(5); m2[cc] = epi_fwd(A[cc], cc, b_1); signal_ready(cc); }       // -> B2
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) { wait_mma(cc); if (cc == 0) NDP_TR(6); dbs2 += epi_bwd(A[cc], cc, m2[cc]); signal_ready(cc); }    // -> B1
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                wait_mma(cc); if (cc == 0) NDP_TR(7);
                dbs1 += epi_bwd(B[cc], cc, m1[cc]);
                gen_h0(A[cc], cc);              // delta2 is dead (B1 retired): h0 again, for dW_0 and relu'(h0)
                signal_ready(cc);           // -> B0
            }
#pragma unroll
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) { wait_mma(cc); if (cc == 0) NDP_TR(8); dbs0 += epi_bwd(A[cc], cc, h0_mask(A[cc]));                 signal_ready(cc); }    // -> Bin
            NDP_TR(9);
            if (a.gx) {     // optional dL/dx: + the path through the positional encoding (ndp_head_grad_kernel wrote the direct part)
                ndp_group_sync(1, 512);     // delta0 of both chains complete
                const int p = ct >> 2, oq = ct & 3;
                float de[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
                const float* wi = params + L.off_w_in;
                for (int o = 32 * oq; o < 32 * oq + 32; ++o) {
                    const unsigned off = ndp_img_off(o, p, NDP_RS64);
                    const float d = (ndp_f16_to_f32(*(const unsigned short*)(A[c] + off)) + ndp_f16_to_f32(*(const unsigned short*)(A[c] + NDP_IMG64 + off))) * dinv;
#pragma unroll
                    for (int k = 0; k < 6; ++k) de[k] = fmaf(d, __ldg(wi + o * 6 + k), de[k]);
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    de[k] += __shfl_xor_sync(0xffffffffu, de[k], 1);
                    de[k] += __shfl_xor_sync(0xffffffffu, de[k], 2);
                }
                const int gp = tile * NDP_TP + c * NDP_HP + p;
                if (oq == 0 && gp < n) {
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
        }
        if (warp == 0) ndp_mbar_wait(&S.bar_fin, 0u);      // every product of the CTA has retired
        ndp_group_sync(1, 512);
        ndp_tc_fence_after();
        NDP_TR(10);
        // bias gradients: this thread's sums, combined below in a fixed order
        ndp_tc_fence_before();
        float* dbred = (float*)S.HG[0];     // [column group][layer][feature] = 6 KB <= 8 KB (HG is free: every product has retired)
        dbred[(cg * 3 + 0) * NDP_W + f] = dbs0;
        dbred[(cg * 3 + 1) * NDP_W + f] = dbs1;
        dbred[(cg * 3 + 2) * NDP_W + f] = dbs2;
    }
    ndp_tc_fence_before();
    __syncthreads();            // every MMA of the CTA has retired (each chain waited for its last commit), dbred complete
    ndp_tc_fence_after();
    NDP_TR(11);
    if (!issw) {
        const unsigned tlane = tmem + ((unsigned)(q * 32) << 16);
        const int cq = warp >> 2;                               // 0..3: 32-column group of the 128-column accumulators
        float w[32];
        // dW_l^T leaves TMEM: lane = input feature i, column = output o, so every store instruction of a warp
        // writes 128 contiguous bytes of the canonical [o][i] block
        ndp_tmem_ld32(tlane + RC_DW1 + 32u * (unsigned)cq, w);
        {
            float* dst = part + L.off_w[1] + (32 * cq) * NDP_W + f;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * NDP_W] = w[j] * dinv;
        }
        ndp_tmem_ld32(tlane + RC_DW0 + 32u * (unsigned)cq, w);
        {
            float* dst = part + L.off_w[0] + (32 * cq) * NDP_W + f;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j * NDP_W] = w[j] * dinv;
        }
        if (cq == 0) {          // head weight gradients dW_h^T[i][r]
            float h[16];
            ndp_tmem_ld16(tlane + RC_DWH, h);
#pragma unroll
            for (int r = 0; r < NDP_MAX_HEAD; ++r)
                if (r < HD) part[L.head_w[r] + f] = h[r] * dinv;
        } else if (cq == 1) {   // input layer dW_in[o][0..5]
            float h[16];
            ndp_tmem_ld16(tlane + RC_DWIN, h);
            float* dst = part + L.off_w_in + f * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) dst[k] = h[k] * dinv;
        } else if (cq == 2) {   // bias gradients: the four column groups in fixed order
            const float* dbred = (const float*)S.HG[0];
            float tot[3];
#pragma unroll
            for (int l = 0; l < 3; ++l)
                tot[l] = (dbred[(0 * 3 + l) * NDP_W + f] + dbred[(1 * 3 + l) * NDP_W + f]) +
                         (dbred[(2 * 3 + l) * NDP_W + f] + dbred[(3 * 3 + l) * NDP_W + f]);
            part[L.off_b_in + f] = tot[0] * dinv;
            part[L.off_b[0] + f] = tot[1] * dinv;
            part[L.off_b[1] + f] = tot[2] * dinv;
        }
    }
    if (tid < HD) {             // db_h: per-tile sums of the pre-kernel, tiles in order
        float sv = 0.0f;
        for (int t = 0; t < ntl; ++t) sv += rec0[(long long)t * NDP_HGREC + NDP_TP * 16 + (1 + tid) * 8 + 7];
        part[L.head_b[tid]] = sv;
    }
    ndp_tc_fence_before();
    __syncthreads();
    NDP_TR(12);
#ifndef NDP_EMU
    if (tid == 0 && blockIdx.x == 0) {   // self-contained duration of this CTA (valid with several launches in flight) + its SM
        unsigned long long t1_; unsigned sm_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
        ndp_dbg_rc[62] = t1_ - t_cta0; ndp_dbg_rc[63] = sm_;
    }
    if (tid == 0) {                      // running sum / count of CTA durations over every launch since the last read
        unsigned long long t1_;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_));
        atomicAdd(&ndp_dbg_rc_sum[0], t1_ - t_cta0); atomicAdd(&ndp_dbg_rc_sum[1], 1ull);
    }
#endif
    if (warp == 0) ndp_tmem_dealloc(S.tmem_slot, 512);
}

void ndp_launch_bwd_rc_main(const NdpBwdArgs& b, int grid_x, cudaStream_t s) {
    NDP_LAUNCH_PRIO(0, ndp_warp_bwd_rc_kernel, dim3(grid_x, b.npairs), dim3(NDP_RC_THREADS), ndp_bwd_rc_smem_bytes(), s, b);
}

int ndp_bwd_rc_init() {
    return (int)cudaFuncSetAttribute(ndp_warp_bwd_rc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)ndp_bwd_rc_smem_bytes());
}

#ifndef NDP_EMU
int ndp_debug_copy_rc(unsigned long long* out) {
    int e = (int)cudaMemcpyFromSymbol(out, ndp_dbg_rc, sizeof(unsigned long long) * 64);
    unsigned long long s2[2] = {0, 0}, z[2] = {0, 0};
    if (e == 0) e = (int)cudaMemcpyFromSymbol(s2, ndp_dbg_rc_sum, sizeof(s2));
    if (e == 0) e = (int)cudaMemcpyToSymbol(ndp_dbg_rc_sum, z, sizeof(z));
    out[60] = s2[0]; out[61] = s2[1];
    return e;
}
#else
int ndp_debug_copy_rc(unsigned long long* out) { for (int i = 0; i < 64; ++i) out[i] = 0; return 0; }
#endif
