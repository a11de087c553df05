// Kernel (2): brute-force nearest neighbour between the warped source and the target, both
// directions, and the truncated-L1 Chamfer epilogue (loss, point-wise gradient, early stop).
//
// Reference: model/loss.py:94-258 (compute_truncated_chamfer_distance), which calls
// pytorch3d.ops.knn.knn_points(K=1) twice (loss.py:177-178) [upstream, un-vendored].
// Parity contract for the search (bit-exact indices, see oracle/knn_oracle.c):
//   d(i,j) = fma(dz,dz, fma(dy,dy, dx*dx)), dx = q_i - t_j, fp32, round-to-nearest;
//   ascending j, strict '<' => lowest index wins ties; candidate j = 0 is always taken first.
//
// ndp_nn_kernel: grid (query tiles, target chunks, 2*pairs).  Each thread keeps NDP_NN_Q query
// points and their running (min, argmin) in registers; the target chunk streams through shared
// memory in SoA tiles of 512 points read back as 128-bit broadcast loads.  Target chunks are
// combined in ascending order by the epilogue, which keeps the tie rule exact without atomics.
#include "ndp_kernels.h"

__device__ __forceinline__ float ndp_sqdist(float qx, float qy, float qz, float tx, float ty, float tz) {
    const float dx = __fsub_rn(qx, tx), dy = __fsub_rn(qy, ty), dz = __fsub_rn(qz, tz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ int ndp_nn_chunks_for(int nt, int chunk_targets) {
    return (nt + chunk_targets - 1) / chunk_targets;
}

__global__ void __launch_bounds__(NDP_NN_THREADS) ndp_nn_kernel(NdpNnArgs a) {
    __shared__ __align__(16) float sx[NDP_NN_TS];
    __shared__ __align__(16) float sy[NDP_NN_TS];
    __shared__ __align__(16) float sz[NDP_NN_TS];

    const int tid = threadIdx.x;
    const int dir = blockIdx.z & 1, pair = (blockIdx.z >> 1) + a.pair0, chunk = blockIdx.y, qt = blockIdx.x;
    if (a.state && a.state[pair].stopped) return;
    const int n = a.ncounts ? a.ncounts[pair] : a.n;
    const int m = a.mcounts ? a.mcounts[pair] : a.m;
    const int nq = dir ? m : n, nt = dir ? n : m;
    const float* Q = dir ? a.y + (long long)pair * a.y_stride : a.x + (long long)pair * a.x_stride;
    const float* T = dir ? a.x + (long long)pair * a.x_stride : a.y + (long long)pair * a.y_stride;
    if (qt * NDP_NN_QT >= nq) return;
    const int t0 = chunk * a.chunk_targets;
    if (t0 >= nt) return;
    const int t1 = (t0 + a.chunk_targets < nt) ? t0 + a.chunk_targets : nt;

    float qx[NDP_NN_Q], qy[NDP_NN_Q], qz[NDP_NN_Q], best[NDP_NN_Q];
    int bi[NDP_NN_Q];
    const float INF = __int_as_float(0x7f800000);
#pragma unroll
    for (int u = 0; u < NDP_NN_Q; ++u) {
        const int q = qt * NDP_NN_QT + tid + u * NDP_NN_THREADS;
        qx[u] = qy[u] = qz[u] = 0.0f;
        if (q < nq) { qx[u] = __ldg(Q + (long long)q * 3); qy[u] = __ldg(Q + (long long)q * 3 + 1); qz[u] = __ldg(Q + (long long)q * 3 + 2); }
        best[u] = INF; bi[u] = 0x7fffffff;
        if (chunk == 0) {   // the first candidate is always accepted (NaN sticks), as in pytorch3d
            best[u] = ndp_sqdist(qx[u], qy[u], qz[u], __ldg(T), __ldg(T + 1), __ldg(T + 2));
            bi[u] = 0;
        }
    }

    for (int base = t0; base < t1; base += NDP_NN_TS) {
        __syncthreads();
        // coalesced AoS read -> SoA tile; slots past the chunk end hold +inf (never selected)
#pragma unroll
        for (int i = 0; i < 3 * NDP_NN_TS / NDP_NN_THREADS; ++i) {
            const int f = tid + i * NDP_NN_THREADS;
            const int j = f / 3, c = f - 3 * j;
            const float v = (base + j < t1) ? __ldg(T + (long long)base * 3 + f) : INF;
            float* dst = (c == 0) ? sx : (c == 1) ? sy : sz;
            dst[j] = v;
        }
        __syncthreads();
        const int cnt = (t1 - base < NDP_NN_TS) ? t1 - base : NDP_NN_TS;
        const int n4 = (cnt + 3) >> 2;
#pragma unroll 2
        for (int j4 = 0; j4 < n4; ++j4) {
            const float4 tx = *(const float4*)(sx + j4 * 4);
            const float4 ty = *(const float4*)(sy + j4 * 4);
            const float4 tz = *(const float4*)(sz + j4 * 4);
            const int jb = base + j4 * 4;
#pragma unroll
            for (int u = 0; u < NDP_NN_Q; ++u) {
                float d;
                d = ndp_sqdist(qx[u], qy[u], qz[u], tx.x, ty.x, tz.x); if (d < best[u]) { best[u] = d; bi[u] = jb; }
                d = ndp_sqdist(qx[u], qy[u], qz[u], tx.y, ty.y, tz.y); if (d < best[u]) { best[u] = d; bi[u] = jb + 1; }
                d = ndp_sqdist(qx[u], qy[u], qz[u], tx.z, ty.z, tz.z); if (d < best[u]) { best[u] = d; bi[u] = jb + 2; }
                d = ndp_sqdist(qx[u], qy[u], qz[u], tx.w, ty.w, tz.w); if (d < best[u]) { best[u] = d; bi[u] = jb + 3; }
            }
        }
    }
    float2* out = a.part + (long long)pair * a.part_pair_stride + ((long long)dir * a.chunks + chunk) * a.qpitch;
#pragma unroll
    for (int u = 0; u < NDP_NN_Q; ++u) {
        const int q = qt * NDP_NN_QT + tid + u * NDP_NN_THREADS;
        if (q < nq) out[q] = make_float2(best[u], __int_as_float(bi[u]));
    }
}

void ndp_launch_nn(const NdpNnArgs& a, cudaStream_t s) {
    if (a.npairs <= 0) return;
    const int nmax = a.n > a.m ? a.n : a.m;
    if (nmax <= 0) return;
    dim3 grid((nmax + NDP_NN_QT - 1) / NDP_NN_QT, a.chunks, 2 * a.npairs);
    NDP_LAUNCH(ndp_nn_kernel, grid, dim3(NDP_NN_THREADS), 0, s, a);
}

// ------------------------------------------------------------------------------------------------
// Epilogue: combine target chunks (ascending, strict '<'), truncation mask, L1 sums, dL/dx.
//   loss = (1/n) sum_i sqrt(d2x_i) + (1/m) sum_j sqrt(d2y_j)                  loss.py:227-255
//   dL/dx_i = (x_i - y_nn(i)) / (n |.|)  +  sum_{j: nn(j) = i} (x_i - y_j) / (m |.|)
// The second (scattered) term is accumulated in 2^-40 fixed point with 64-bit integer atomics:
// integer addition is associative, so the gradient is bit-reproducible (pytorch3d's float
// atomicAdd backward is not).  Entries with squared distance >= trunc contribute neither loss nor
// gradient (loss.py:185-188); the divisors stay the full lengths (loss.py:233-235).
// The last CTA of a pair to finish sums the per-CTA partial sums in index order and applies the
// early-stop rule of model/registration.py:225-232 in double precision on the fp32 loss.
// ------------------------------------------------------------------------------------------------
#define NDP_CR_THREADS 256

__device__ __forceinline__ void ndp_combine(const float2* part, int chunks, int qpitch, int q, float& d, int& idx) {
    const float2 p0 = part[q];
    d = p0.x; idx = __float_as_int(p0.y);
    for (int c = 1; c < chunks; ++c) {
        const float2 p = part[(long long)c * qpitch + q];
        if (p.x < d) { d = p.x; idx = __float_as_int(p.y); }
    }
}

__global__ void __launch_bounds__(NDP_CR_THREADS) ndp_chamfer_reduce_kernel(NdpChamferArgs a) {
    __shared__ double red[2][NDP_CR_THREADS];
    __shared__ int is_last;
    const int tid = threadIdx.x, pair = blockIdx.y + a.nn.pair0;
    if (a.state && a.state[pair].stopped) return;
    const NdpNnArgs& g = a.nn;
    const int n = g.ncounts ? g.ncounts[pair] : g.n;
    const int m = g.mcounts ? g.mcounts[pair] : g.m;
    const int nmax = n > m ? n : m;
    const int nblocks = (nmax + NDP_CR_THREADS - 1) / NDP_CR_THREADS;
    if ((int)blockIdx.x >= nblocks) return;
    const float* X = g.x + (long long)pair * g.x_stride;
    const float* Y = g.y + (long long)pair * g.y_stride;
    const float2* part0 = g.part + (long long)pair * g.part_pair_stride;
    const float2* part1 = part0 + (long long)g.chunks * g.qpitch;
    const int i = blockIdx.x * NDP_CR_THREADS + tid;
    double sx = 0.0, sy = 0.0;

    if (a.paired) {
        if (i < n) {   // landmark term: d/dx_i mean_k |x_k - y_k|^2 = 2 (x_i - y_i) / n   (registration.py:200-203)
            const float dx = X[(long long)i * 3] - Y[(long long)i * 3], dy = X[(long long)i * 3 + 1] - Y[(long long)i * 3 + 1],
                        dz = X[(long long)i * 3 + 2] - Y[(long long)i * 3 + 2];
            const float sc = 2.0f / (float)n;
            float* gp = a.gx + (long long)pair * a.gx_stride + (long long)i * 3;
            gp[0] = dx * sc; gp[1] = dy * sc; gp[2] = dz * sc;
            sx = (double)(dx * dx + dy * dy + dz * dz);
        }
    } else {
    if (i < n) {   // direction x -> y : point i owns its gradient slot
        float d; int j;
        ndp_combine(part0, ndp_nn_chunks_for(m, g.chunk_targets), g.qpitch, i, d, j);
        if (a.d2x) { a.d2x[(long long)pair * a.nx_stride + i] = d; a.idxx[(long long)pair * a.nx_stride + i] = j; }
        float gx0 = 0.0f, gx1 = 0.0f, gx2 = 0.0f;
        if (!(d >= a.trunc)) {
            const float s = sqrtf(d);
            const float inv = 1.0f / ((float)n * s);
            gx0 = (X[(long long)i * 3] - Y[(long long)j * 3]) * inv;
            gx1 = (X[(long long)i * 3 + 1] - Y[(long long)j * 3 + 1]) * inv;
            gx2 = (X[(long long)i * 3 + 2] - Y[(long long)j * 3 + 2]) * inv;
            sx = (double)s;
        }
        float* gp = a.gx + (long long)pair * a.gx_stride + (long long)i * 3;
        gp[0] = gx0; gp[1] = gx1; gp[2] = gx2;
    }
    if (i < m) {   // direction y -> x : scatter onto the nearest source point
        float d; int k;
        ndp_combine(part1, ndp_nn_chunks_for(n, g.chunk_targets), g.qpitch, i, d, k);
        if (a.d2y) { a.d2y[(long long)pair * a.ny_stride + i] = d; a.idxy[(long long)pair * a.ny_stride + i] = k; }
        if (!(d >= a.trunc)) {
            const float s = sqrtf(d);
            sy = (double)s;
            unsigned long long* acc = a.gacc + (long long)pair * a.gacc_stride + (long long)k * 3;
            if (s > 0.0f && s < __int_as_float(0x7f800000)) {
                const float inv = 1.0f / s;
                const float SC = 1099511627776.0f;   // 2^40
                atomicAdd(acc + 0, (unsigned long long)__float2ll_rn((X[(long long)k * 3] - Y[(long long)i * 3]) * inv * SC));
                atomicAdd(acc + 1, (unsigned long long)__float2ll_rn((X[(long long)k * 3 + 1] - Y[(long long)i * 3 + 1]) * inv * SC));
                atomicAdd(acc + 2, (unsigned long long)__float2ll_rn((X[(long long)k * 3 + 2] - Y[(long long)i * 3 + 2]) * inv * SC));
            } else {
                atomicAdd(acc + 0, 1ull << 62);   // 0/0 or NaN: poison marker -> NaN gradient, as the reference
            }
        }
    }
    }   // !paired
    // block reduction of the two L1 sums (fixed tree => deterministic)
    red[0][tid] = sx; red[1][tid] = sy;
    __syncthreads();
    for (int off = NDP_CR_THREADS / 2; off > 0; off >>= 1) {
        if (tid < off) { red[0][tid] += red[0][tid + off]; red[1][tid] += red[1][tid + off]; }
        __syncthreads();
    }
    double* bs = a.blocksums + ((long long)pair * a.blocks_pitch) * 2;
    if (tid == 0) {
        bs[blockIdx.x * 2 + 0] = red[0][0];
        bs[blockIdx.x * 2 + 1] = red[1][0];
        __threadfence();
        const int ticket = atomicAdd(a.counters + pair, 1);
        is_last = (ticket == nblocks - 1);
    }
    __syncthreads();
    if (is_last && tid == 0) {
        __threadfence();
        double tx = 0.0, ty = 0.0;
        for (int b = 0; b < nblocks; ++b) { tx += ((volatile double*)bs)[b * 2]; ty += ((volatile double*)bs)[b * 2 + 1]; }
        const float loss = a.paired ? (float)tx / (float)n : (float)tx / (float)n + (float)ty / (float)m;
        a.loss_out[pair] = loss;
        a.counters[pair] = 0;
        if (a.state) {
            NdpPairState* st = a.state + pair;
            if (a.loss_hist && st->evals < a.hist_cap) a.loss_hist[(long long)pair * a.hist_stride + st->evals] = loss;
            st->evals += 1;
            st->last_loss = loss;
            const double l = (double)loss;                       // registration.py:225-232
            if (l < 1e-4) {
                st->stopped = 1;
            } else {
                if (fabs(st->loss_prev - l) < st->loss_prev * a.break_ratio) st->break_counter += 1;
                if (st->break_counter >= a.max_break_count) st->stopped = 1;
                else st->loss_prev = l;
            }
        }
    }
}

void ndp_launch_chamfer_reduce(const NdpChamferArgs& a, cudaStream_t s) {
    if (a.nn.npairs <= 0) return;
    const int nmax = a.nn.n > a.nn.m ? a.nn.n : a.nn.m;
    if (nmax <= 0) return;
    dim3 grid((nmax + NDP_CR_THREADS - 1) / NDP_CR_THREADS, a.nn.npairs);
    NDP_LAUNCH(ndp_chamfer_reduce_kernel, grid, dim3(NDP_CR_THREADS), 0, s, a);
}

// gx[i] = scale * (gx[i] + fixed-point accumulator), accumulator cleared (standalone Chamfer API)
__global__ void ndp_grad_finalize_kernel(NdpGradFinalizeArgs a) {
    const int pair = blockIdx.y;
    const int n = a.ncounts ? a.ncounts[pair] : a.n;
    const int m = a.mcounts ? a.mcounts[pair] : a.m;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* g = a.gx + (long long)pair * a.gx_stride + (long long)i * 3;
    unsigned long long* acc = a.gacc + (long long)pair * a.gacc_stride + (long long)i * 3;
    const double sc = 9.094947017729282e-13 / (double)m;
    const long long a0 = (long long)acc[0], a1 = (long long)acc[1], a2 = (long long)acc[2];
    const bool poison = (a0 >= (1LL << 60)) || (a0 <= -(1LL << 60));
    g[0] = a.scale * (g[0] + (poison ? __int_as_float(0x7fc00000) : (float)((double)a0 * sc)));
    g[1] = a.scale * (g[1] + (float)((double)a1 * sc));
    g[2] = a.scale * (g[2] + (float)((double)a2 * sc));
    acc[0] = 0ull; acc[1] = 0ull; acc[2] = 0ull;
}

void ndp_launch_grad_finalize(const NdpGradFinalizeArgs& a, cudaStream_t s) {
    if (a.npairs <= 0 || a.n <= 0) return;
    dim3 grid((a.n + 255) / 256, a.npairs);
    NDP_LAUNCH(ndp_grad_finalize_kernel, grid, dim3(256), 0, s, a);
}
