// Kernel (2), fused-driver variant: EXACT nearest neighbour with spatial culling.
//
// Same contract as the brute-force search of ndp_chamfer.cu (same fp32 distance expression,
// lexicographic (distance, original index) minimum => lowest index wins ties, identical results),
// but most of the N x M distance evaluations are skipped:
//   * once per pair both sampled clouds are put in Morton (Z-curve) order (30-bit keys, bitonic
//     sort); the order is kept for all levels and iterations (the warp field is smooth, so a block
//     of 32 consecutive source points stays compact while it deforms);
//   * every block of 32 points carries an axis-aligned bounding box (the target's once, the warped
//     source's re-computed by the forward kernel every iteration);
//   * a warp owns 32 consecutive queries; its running minima are seeded with the distance to each
//     query's nearest neighbour of the PREVIOUS iteration (a real candidate => a valid upper bound);
//   * a target block is scanned only if its box can contain a point at distance <= the current
//     minimum of some lane.  Box distances are evaluated with the same monotone fp32 expression
//     (componentwise gaps -> fma chain), so in floating point lb(box) <= d(q, t) for every t in the
//     box and culling never discards a candidate, not even an exact tie.
//   * inside a block the points are ordered by ORIGINAL sample index, so the first minimum of a scan in block order is
//     the one the reference's tie rule picks: a block is scanned with a plain "d < best" (distance + position, no index
//     in the loop) and only its winner is merged into the running (distance, original index) key;
//   * the scan is bound by the shared-memory pipe, not by the FP32 pipes: every candidate is broadcast to the 32 lanes
//     and an LDS.128 occupies that pipe for four cycles whether or not the lanes read the same address.  The staged
//     block therefore holds coordinates only (12 bytes per candidate, negated and pair-interleaved for the packed
//     FADD2 / FMUL2 / FFMA2), and the index is fetched once per block.
// Reference semantics: pytorch3d knn_points(K=1) as called at model/loss.py:177-178.
#include "ndp_kernels.h"

#ifdef NDP_EMU
static inline int __ffs(int v) { return __builtin_ffs(v); }
#endif

#ifdef NDP_EMU
static inline int ndp_atomic_add_release(int* p, int v) { return atomicAdd(p, v); }
#else
__device__ __forceinline__ int ndp_atomic_add_release(int* p, int v) {
    int old;
    asm volatile("atom.add.release.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
#endif
__device__ __forceinline__ float ndp_sqdist3(float qx, float qy, float qz, float tx, float ty, float tz) {
    const float dx = __fsub_rn(qx, tx), dy = __fsub_rn(qy, ty), dz = __fsub_rn(qz, tz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
// gap between [alo,ahi] and [blo,bhi] along one axis (0 when they overlap); NaN-free for finite boxes
__device__ __forceinline__ float ndp_gap(float alo, float ahi, float blo, float bhi) {
    return fmaxf(0.0f, fmaxf(__fsub_rn(blo, ahi), __fsub_rn(alo, bhi)));
}
__device__ __forceinline__ float ndp_sq3(float gx, float gy, float gz) {
    return __fmaf_rn(gz, gz, __fmaf_rn(gy, gy, __fmul_rn(gx, gx)));
}
// squared distances of one query to a staged candidate pair (nx = (-x0, -x1), ...): each component is the same IEEE
// operation sequence as ndp_sqdist3 (q - t == q + (-t) exactly), so the results are bit-identical to it
__device__ __forceinline__ void ndp_sqdist3_pair(NdpF2 qx, NdpF2 qy, NdpF2 qz, NdpF2 nx, NdpF2 ny, NdpF2 nz, float& d0, float& d1) {
    const NdpF2 dx = ndp_f2_add(qx, nx), dy = ndp_f2_add(qy, ny), dz = ndp_f2_add(qz, nz);
    ndp_f2_get(ndp_f2_fma(dz, dz, ndp_f2_fma(dy, dy, ndp_f2_mul(dx, dx))), d0, d1);
}

// ---- cloud bounds (one CTA per (pair, cloud)); deterministic min/max tree -----------------------
__global__ void __launch_bounds__(256) ndp_bounds_kernel(NdpSortArgs a) {
    __shared__ float lo[3][256], hi[3][256];
    const int pair = blockIdx.x, which = blockIdx.y, tid = threadIdx.x;
    const int n = which ? (a.mcounts ? a.mcounts[pair] : a.m) : (a.ncounts ? a.ncounts[pair] : a.n);
    const float* P = (which ? a.tgt : a.src) + (long long)pair * a.cloud_stride;
    const float INF = __int_as_float(0x7f800000);
    float l0 = INF, l1 = INF, l2 = INF, h0 = -INF, h1 = -INF, h2 = -INF;
    for (int i = tid; i < n; i += 256) {
        const float x = P[(long long)i * 3], y = P[(long long)i * 3 + 1], z = P[(long long)i * 3 + 2];
        l0 = fminf(l0, x); l1 = fminf(l1, y); l2 = fminf(l2, z);
        h0 = fmaxf(h0, x); h1 = fmaxf(h1, y); h2 = fmaxf(h2, z);
    }
    lo[0][tid] = l0; lo[1][tid] = l1; lo[2][tid] = l2; hi[0][tid] = h0; hi[1][tid] = h1; hi[2][tid] = h2;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off)
            for (int c = 0; c < 3; ++c) {
                lo[c][tid] = fminf(lo[c][tid], lo[c][tid + off]);
                hi[c][tid] = fmaxf(hi[c][tid], hi[c][tid + off]);
            }
        __syncthreads();
    }
    if (tid < 3) {
        float* b = a.bounds + ((long long)pair * 2 + which) * 6;
        b[tid] = lo[tid][0]; b[3 + tid] = hi[tid][0];
    }
}

__device__ __forceinline__ unsigned ndp_expand10(unsigned v) {   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

// key = morton30 << 32 | sample index; slots >= n get the maximum key (sorted to the end)
__global__ void __launch_bounds__(256) ndp_morton_keys_kernel(NdpSortArgs a) {
    const int pair = blockIdx.y, which = blockIdx.z;
    const int n = which ? (a.mcounts ? a.mcounts[pair] : a.m) : (a.ncounts ? a.ncounts[pair] : a.n);
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= a.npad) return;
    unsigned long long* keys = a.keys + ((long long)pair * 2 + which) * a.npad;
    if (i >= n) { keys[i] = ~0ull; return; }
    const float* P = (which ? a.tgt : a.src) + (long long)pair * a.cloud_stride + (long long)i * 3;
    const float* b = a.bounds + ((long long)pair * 2 + which) * 6;
    unsigned code = 0;
    for (int c = 0; c < 3; ++c) {
        const float ext = b[3 + c] - b[c];
        float u = ext > 0.0f ? (P[c] - b[c]) / ext : 0.0f;
        u = fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f);
        unsigned q = (u == u) ? (unsigned)u : 0u;     // NaN coordinates sort first
        code |= ndp_expand10(q) << c;
    }
    keys[i] = ((unsigned long long)code << 32) | (unsigned)i;
}

__global__ void __launch_bounds__(256) ndp_bitonic_step_kernel(unsigned long long* keys, int npad, int k, int j) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= npad) return;
    unsigned long long* K = keys + (long long)blockIdx.y * npad;
    const int l = i ^ j;
    if (l > i) {
        const unsigned long long x = K[i], y = K[l];
        const bool up = (i & k) == 0;
        if ((x > y) == up) { K[i] = y; K[l] = x; }
    }
}

// sorted [n][3] copy (+ float4 copy carrying the original sample index) and the 32-point block boxes
__global__ void __launch_bounds__(256) ndp_apply_order_kernel(NdpSortArgs a) {
    const int pair = blockIdx.y, which = blockIdx.z;
    const int n = which ? (a.mcounts ? a.mcounts[pair] : a.m) : (a.ncounts ? a.ncounts[pair] : a.n);
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int npts = which ? a.m : a.n;
    if (i >= ((npts + 31) / 32) * 32) return;                 // whole warps stay together for the shuffles
    const unsigned long long* keys = a.keys + ((long long)pair * 2 + which) * a.npad;
    const float* P = (which ? a.tgt : a.src) + (long long)pair * a.cloud_stride;
    float* out = (which ? a.tgt_sorted : a.src_sorted) + (long long)pair * a.cloud_stride;
    const float INF = __int_as_float(0x7f800000);
    float x = INF, y = INF, z = INF;
    int o = 0x7fffffff;
    if (i < n) o = (int)(unsigned)(keys[i] & 0xffffffffull);
    // inside the 32-point block: ascending original index (bitonic network over the warp; padded slots stay last)
    for (int k = 2; k <= 32; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int other = __shfl_xor_sync(0xffffffffu, o, j);
            const bool up = (lane & k) == 0, lower = (lane & j) == 0;
            o = (lower == up) ? (o < other ? o : other) : (o < other ? other : o);
        }
    if (i < n) {
        x = P[(long long)o * 3]; y = P[(long long)o * 3 + 1]; z = P[(long long)o * 3 + 2];
        out[(long long)i * 3] = x; out[(long long)i * 3 + 1] = y; out[(long long)i * 3 + 2] = z;
        (which ? a.tgt_orig : a.src_orig)[(long long)pair * a.orig_stride + i] = o;
        (which ? a.tgt_inv : a.src_inv)[(long long)pair * a.orig_stride + o] = i;
    }
    if (which) a.tgt4[(long long)pair * a.p4_stride + i] = make_float4(x, y, z, __int_as_float(o));
    // block box of the (static) target; padded slots (+inf) are excluded
    float l0 = x, l1 = y, l2 = z, h0 = (i < n) ? x : -INF, h1 = (i < n) ? y : -INF, h2 = (i < n) ? z : -INF;
    for (int s = 16; s > 0; s >>= 1) {
        l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, s)); l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, s));
        l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, s)); h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, s));
        h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, s)); h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, s));
    }
    if (which && lane == 0) {
        float* bx = a.tgt_box + ((long long)pair * a.box_stride + (i >> 5)) * 8;
        bx[0] = l0; bx[1] = l1; bx[2] = l2; bx[3] = 0.0f; bx[4] = h0; bx[5] = h1; bx[6] = h2; bx[7] = 0.0f;
    }
}

int ndp_launch_sort(const NdpSortArgs& a, cudaStream_t s) {
    if (a.npairs <= 0) return 0;
    NDP_LAUNCH(ndp_bounds_kernel, dim3(a.npairs, 2), dim3(256), 0, s, a);
    NDP_LAUNCH(ndp_morton_keys_kernel, dim3((a.npad + 255) / 256, a.npairs, 2), dim3(256), 0, s, a);
    int launches = 2;
    for (int k = 2; k <= a.npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            NDP_LAUNCH(ndp_bitonic_step_kernel, dim3((a.npad + 255) / 256, a.npairs * 2), dim3(256), 0, s, a.keys, a.npad, k, j);
            ++launches;
        }
    const int nmax = a.n > a.m ? a.n : a.m;
    NDP_LAUNCH(ndp_apply_order_kernel, dim3((((nmax + 31) / 32) * 32 + 255) / 256, a.npairs, 2), dim3(256), 0, s, a);
    return launches + 1;
}

// ---- the culled search ---------------------------------------------------------------------------
// One warp = 32 consecutive (Morton-ordered) queries.
// With FUSE the kernel also does the whole Chamfer epilogue of ndp_chamfer.cu (truncation, L1 sums, direct
// and scattered gradient terms, loss + early-stop rule in the last CTA of the pair): one launch less per
// iteration and no second pass over the (d2, idx) arrays.
#ifndef NDP_PN_WARPS
#define NDP_PN_WARPS 4
#endif
struct NdpFuseArgs {             // the Chamfer epilogue's arguments (subset of NdpChamferArgs), by value
    float trunc;
    float* gx; long long gx_stride;
    unsigned long long* gacc; long long gacc_stride;
    double* blocksums; int blocks_pitch;
    int* counters; float* loss_out; NdpPairState* state;
    float* loss_hist; long long hist_stride; int hist_cap;
    int max_break_count; double break_ratio;
};

template <bool FUSE>
__global__ void __launch_bounds__(NDP_PN_WARPS * 32, 48 / NDP_PN_WARPS) ndp_nn_pruned_kernel(NdpPrunedArgs a, NdpFuseArgs fz) {
    __shared__ __align__(16) float stage[NDP_PN_WARPS][16][6];   // per warp: 16 candidate pairs (-x0, -x1, -y0, -y1, -z0, -z1)
    __shared__ int stage_o[NDP_PN_WARPS][32];                    // their original sample indices
    __shared__ double wsum[NDP_PN_WARPS];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int dir = blockIdx.z & 1, pair = (blockIdx.z >> 1) + a.pair0;
    if (a.state && a.state[pair].stopped) return;
    const int n = a.ncounts ? a.ncounts[pair] : a.n;
    const int m = a.mcounts ? a.mcounts[pair] : a.m;
    const int nq = dir ? m : n, nt = dir ? n : m;
    if ((int)blockIdx.x * NDP_PN_WARPS * 32 >= nq) return;    // CTA-uniform: no query block of this direction here
    const int qblk = blockIdx.x * NDP_PN_WARPS + w;            // one 32-query block per warp
    const bool wlive = qblk * 32 < nq;                         // warp-uniform
    const float4* Q4 = (dir ? a.y4 : a.x4) + (long long)pair * a.p4_stride;
    const float4* T4 = (dir ? a.x4 : a.y4) + (long long)pair * a.p4_stride;
    const float* tbox = (dir ? a.xbox : a.ybox) + (long long)pair * a.box_stride * 8;
    const float INF = __int_as_float(0x7f800000);
    double lsum = 0.0;                                         // this lane's L1 term (FUSE)

    if (wlive) {
    const float* qbox = (dir ? a.ybox : a.xbox) + ((long long)pair * a.box_stride + qblk) * 8;
    int* prev = (dir ? a.prev_y : a.prev_x) + (long long)pair * a.prev_stride;
    const int ntblk = (nt + 31) >> 5;
    const int q = qblk * 32 + lane;
    const bool live = q < nq;
    const float4 qq = Q4[live ? q : (nq - 1)];
    // seed: nearest neighbour of the previous iteration (or a position-based guess)
    int js = live ? prev[q] : 0;
    if (js < 0 || js >= nt) js = (int)(((long long)q * nt) / nq);
    if (js >= nt) js = nt - 1;
    const float4 ts = T4[js];
    float best = ndp_sqdist3(qq.x, qq.y, qq.z, ts.x, ts.y, ts.z);
    if (!(best == best)) best = INF;                           // NaN seed: fall back to a full scan
    // (distance, original index) packed into one 64-bit key: for non-negative floats the bit pattern
    // is monotone, so "d < best || (d == best && o < bo)" is ONE unsigned compare (NaN sorts last)
    unsigned long long bestk = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)__float_as_int(ts.w);
    float wmax = live ? best : 0.0f;                           // dead lanes never widen the search
    for (int s = 16; s > 0; s >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, s));
    const float4 qlo = *(const float4*)qbox, qhi = *(const float4*)(qbox + 4);
    int scanned = 0;
    const NdpF2 qx2 = ndp_f2_bcast(qq.x), qy2 = ndp_f2_bcast(qq.y), qz2 = ndp_f2_bcast(qq.z);

    // Coarse test of all target blocks first, 32 rounds of 32 blocks at most per batch (lane = block; the loads of the rounds
    // are independent, two rounds in flight); lane r keeps round r's ballot.  Then the survivors are walked.
    for (int sbase = 0; sbase < ntblk && !(a.dbg & 1); sbase += 1024) {
    const int nrounds = (ntblk - sbase + 31) >> 5 < 32 ? (ntblk - sbase + 31) >> 5 : 32;
    int mreg = 0;
#pragma unroll 2
    for (int r = 0; r < nrounds; ++r) {
        const int blk = sbase + r * 32 + lane;
        bool need = false;
        if (blk < ntblk) {
            const float4 blo = *(const float4*)(tbox + (long long)blk * 8), bhi = *(const float4*)(tbox + (long long)blk * 8 + 4);
            const float lbw = ndp_sq3(ndp_gap(qlo.x, qhi.x, blo.x, bhi.x), ndp_gap(qlo.y, qhi.y, blo.y, bhi.y),
                                      ndp_gap(qlo.z, qhi.z, blo.z, bhi.z));
            need = lbw <= wmax;
        }
        const int m = (int)__ballot_sync(0xffffffffu, need);
        if (lane == r) mreg = m;
    }
    for (int r = 0; r < nrounds; ++r) {
        const int base = sbase + r * 32;
        unsigned mask = (unsigned)__shfl_sync(0xffffffffu, mreg, r);
        while (mask) {
            const int b = __ffs((int)mask) - 1;
            mask &= mask - 1;
            const int tb = base + b;
            // the box again, now as a broadcast load (the boxes of a pair are 8 KB per direction: L1 hits).  Measured
            // alternatives, all slower at full occupancy: box shuffled from the lane that tested it (+11 %), next
            // block's points prefetched one ahead (+35 % with speculative loads)
            const float4 blo = *(const float4*)(tbox + (long long)tb * 8), bhi = *(const float4*)(tbox + (long long)tb * 8 + 4);
            const float lbq = ndp_sq3(ndp_gap(qq.x, qq.x, blo.x, bhi.x), ndp_gap(qq.y, qq.y, blo.y, bhi.y),
                                      ndp_gap(qq.z, qq.z, blo.z, bhi.z));
            if (!__any_sync(0xffffffffu, live && lbq <= best)) continue;
            ++scanned;
            const int tj = tb * 32 + lane;
            {   // stage the block: negated, pair-interleaved coordinates (two candidates per packed FP32 instruction)
                const float4 t = (tj < nt) ? T4[tj] : make_float4(INF, INF, INF, __int_as_float(0x7fffffff));
                float* sp = &stage[w][lane >> 1][lane & 1];
                sp[0] = -t.x; sp[2] = -t.y; sp[4] = -t.z;
                stage_o[w][lane] = __float_as_int(t.w);
            }
            __syncwarp();
            // block-local minimum and its position; strict '<' in block order = lowest original index among equal distances
            // (two running minima, even / odd candidates: the compare-select chain is the longest dependency of the scan)
            float bd = INF, bd1 = INF;
            int bp = 0, bp1 = 1;
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const float4 g = *(const float4*)&stage[w][p][0];
                const float2 z = *(const float2*)&stage[w][p][4];
                float d0, d1;
                ndp_sqdist3_pair(qx2, qy2, qz2, ndp_f2_make(g.x, g.y), ndp_f2_make(g.z, g.w), ndp_f2_make(z.x, z.y), d0, d1);
                if (d0 < bd) { bd = d0; bp = 2 * p; }
                if (d1 < bd1) { bd1 = d1; bp1 = 2 * p + 1; }
            }
            if (bd1 < bd || (bd1 == bd && bp1 < bp)) { bd = bd1; bp = bp1; }
            if (bd < INF) {     // (candidates at infinite distance never replace the seed; NaN distances never compare below)
                const unsigned long long k = ((unsigned long long)__float_as_uint(bd) << 32) | (unsigned)stage_o[w][bp];
                bestk = k < bestk ? k : bestk;
                best = __uint_as_float((unsigned)(bestk >> 32));
            }
            __syncwarp();
        }
    }
    }   // sbase
    if (a.stats && lane == 0) {
        atomicAdd(a.stats, (unsigned long long)scanned * 1024ull);
        atomicAdd(a.stats + 1, 1ull);
        atomicMax(a.stats + 3, (unsigned long long)scanned);                     // most blocks any warp scanned
    }
    if (live) {
        const int bo = (int)(unsigned)(bestk & 0xffffffffull);
        const int bj = (dir ? a.inv_x : a.inv_y)[(long long)pair * a.inv_stride + bo];   // sorted position of the winner
        // NaN query: reference semantics are (NaN, index 0)
        if (!(qq.x == qq.x) || !(qq.y == qq.y) || !(qq.z == qq.z)) best = __int_as_float(0x7fc00000);
        a.part[(long long)pair * a.part_pair_stride + (long long)dir * a.qpitch + q] = make_float2(best, __int_as_float(bj));
        prev[q] = bj;
        if (FUSE && !(a.dbg & 2)) {
            // Chamfer epilogue (ndp_chamfer.cu: ndp_chamfer_reduce_kernel), same expressions on the same coordinates
            if (dir == 0) {     // x -> y: point q owns its gradient slot
                float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
                if (!(best >= fz.trunc)) {
                    const float4 t = T4[bj];
                    const float sd = sqrtf(best);
                    const float inv = 1.0f / ((float)n * sd);
                    g0 = (qq.x - t.x) * inv; g1 = (qq.y - t.y) * inv; g2 = (qq.z - t.z) * inv;
                    lsum = (double)sd;
                }
                float* gp = fz.gx + (long long)pair * fz.gx_stride + (long long)q * 3;
                gp[0] = g0; gp[1] = g1; gp[2] = g2;
            } else if (!(best >= fz.trunc)) {   // y -> x: scatter onto the nearest source point (2^-40 fixed point)
                const float sd = sqrtf(best);
                lsum = (double)sd;
                unsigned long long* acc = fz.gacc + (long long)pair * fz.gacc_stride + (long long)bj * 3;
                if (sd > 0.0f && sd < INF) {
                    const float4 t = T4[bj];            // the source point x_k
                    const float inv = 1.0f / sd;
                    const float SC = 1099511627776.0f;   // 2^40
                    atomicAdd(acc + 0, (unsigned long long)__float2ll_rn((t.x - qq.x) * inv * SC));
                    atomicAdd(acc + 1, (unsigned long long)__float2ll_rn((t.y - qq.y) * inv * SC));
                    atomicAdd(acc + 2, (unsigned long long)__float2ll_rn((t.z - qq.z) * inv * SC));
                } else {
                    atomicAdd(acc + 0, 1ull << 62);       // 0/0 or NaN: poison marker -> NaN gradient, as the reference
                }
            }
        }
    }
    }   // wlive

    if (FUSE) {
        // L1 sums: lanes by a fixed butterfly, warps in order, CTAs in index order by the pair's last CTA (deterministic)
        for (int s = 16; s > 0; s >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
        if (lane == 0) wsum[w] = lsum;
        __syncthreads();
        const int nbx = (n + NDP_PN_WARPS * 32 - 1) / (NDP_PN_WARPS * 32), nby = (m + NDP_PN_WARPS * 32 - 1) / (NDP_PN_WARPS * 32);
        double* bs = fz.blocksums + ((long long)pair * fz.blocks_pitch) * 2;
        if (threadIdx.x == 0) {
            double cs = wsum[0];
#pragma unroll
            for (int k = 1; k < NDP_PN_WARPS; ++k) cs += wsum[k];
            bs[(long long)blockIdx.x * 2 + dir] = cs;
            // publish the block sum with a RELEASE on the ticket instead of __threadfence(): the full fence also invalidates
            // the SM's L1 (CCTL.IVALL), i.e. the boxes and tiles the other resident CTAs of the search keep re-reading
            const int ticket = ndp_atomic_add_release(fz.counters + pair, 1);
            is_last = (ticket == nbx + nby - 1);
        }
        __syncthreads();
        if (is_last && threadIdx.x == 0) {
            __threadfence();
            double tx = 0.0, ty = 0.0;
            for (int b = 0; b < nbx; ++b) tx += ((volatile double*)bs)[b * 2];
            for (int b = 0; b < nby; ++b) ty += ((volatile double*)bs)[b * 2 + 1];
            const float loss = (float)tx / (float)n + (float)ty / (float)m;
            fz.loss_out[pair] = loss;
            fz.counters[pair] = 0;
            if (fz.state) {
                NdpPairState* st = fz.state + pair;
                if (fz.loss_hist && st->evals < fz.hist_cap) fz.loss_hist[(long long)pair * fz.hist_stride + st->evals] = loss;
                st->evals += 1;
                st->last_loss = loss;
                const double l = (double)loss;                       // registration.py:225-232
                if (l < 1e-4) {
                    st->stopped = 1;
                } else {
                    if (fabs(st->loss_prev - l) < st->loss_prev * fz.break_ratio) st->break_counter += 1;
                    if (st->break_counter >= fz.max_break_count) st->stopped = 1;
                    else st->loss_prev = l;
                }
            }
        }
    }
}

void ndp_launch_nn_pruned(const NdpPrunedArgs& a, cudaStream_t s) {
    if (a.npairs <= 0) return;
    const int nmax = a.n > a.m ? a.n : a.m;
    if (nmax <= 0) return;
    dim3 grid(((nmax + 31) / 32 + NDP_PN_WARPS - 1) / NDP_PN_WARPS, 1, 2 * a.npairs);
    NdpPrunedArgs b = a;
    b.fuse = nullptr;
    NdpFuseArgs fz = {};
    if (a.fuse) {
        const NdpChamferArgs& c = *a.fuse;
        fz.trunc = c.trunc; fz.gx = c.gx; fz.gx_stride = c.gx_stride; fz.gacc = c.gacc; fz.gacc_stride = c.gacc_stride;
        fz.blocksums = c.blocksums; fz.blocks_pitch = c.blocks_pitch; fz.counters = c.counters; fz.loss_out = c.loss_out;
        fz.state = c.state; fz.loss_hist = c.loss_hist; fz.hist_stride = c.hist_stride; fz.hist_cap = c.hist_cap;
        fz.max_break_count = c.max_break_count; fz.break_ratio = c.break_ratio;
        NDP_LAUNCH_PRIO(1, ndp_nn_pruned_kernel<true>, grid, dim3(NDP_PN_WARPS * 32), 0, s, b, fz);
    } else {
        NDP_LAUNCH(ndp_nn_pruned_kernel<false>, grid, dim3(NDP_PN_WARPS * 32), 0, s, b, fz);
    }
}

// ---- export of the last search in sample order (ndp_solver_last_nn) --------------------------------
__global__ void __launch_bounds__(256) ndp_nn_export_kernel(NdpNnExportArgs a) {
    const int q = blockIdx.x * 256 + threadIdx.x, dir = blockIdx.y;
    const int nq = dir ? a.m : a.n, nt = dir ? a.n : a.m;
    if (q >= nq) return;
    const int chunks = (nt + a.chunk_targets - 1) / a.chunk_targets;
    const float2* part = a.part + (long long)dir * a.chunks * a.qpitch;
    float2 best = part[q];
    for (int c = 1; c < chunks; ++c) {             // brute-force mode: target chunks in ascending order, strict '<'
        const float2 p = part[(long long)c * a.qpitch + q];
        if (p.x < best.x) best = p;
    }
    const int* oq = dir ? a.orig_t : a.orig_s;
    const int* ot = dir ? a.orig_s : a.orig_t;
    const int i = oq ? oq[q] : q;
    const int j = __float_as_int(best.y);
    (dir ? a.idx_y : a.idx_x)[i] = ot ? ot[j] : j;
    (dir ? a.d2_y : a.d2_x)[i] = best.x;
    const float* cin = dir ? a.target : a.warped;
    float* cout = dir ? a.target_out : a.warped_out;
    if (cout) {
        cout[(long long)i * 3] = cin[(long long)q * 3];
        cout[(long long)i * 3 + 1] = cin[(long long)q * 3 + 1];
        cout[(long long)i * 3 + 2] = cin[(long long)q * 3 + 2];
    }
}

void ndp_launch_nn_export(const NdpNnExportArgs& a, cudaStream_t s) {
    const int nmax = a.n > a.m ? a.n : a.m;
    if (nmax <= 0) return;
    NDP_LAUNCH(ndp_nn_export_kernel, dim3((nmax + 255) / 256, 2), dim3(256), 0, s, a);
}
