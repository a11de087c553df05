/*
 * ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product path (deformationpyramid_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement of the K=1 brute-force nearest-neighbour search that the
 * reference's Chamfer loss delegates to pytorch3d:
 *     model/loss.py:177-178   x_nn = knn_points(x, y, ..., K=1); y_nn = knn_points(y, x, ..., K=1)
 *     model/loss.py:180-181   cham_x = x_nn.dists[..., 0]
 * pytorch3d is a third-party dependency that is NOT vendored under /root/reference and
 * whose version the reference does not pin (README.md:15-16 only says "pytorch3d").  Its
 * published CPU algorithm (pytorch3d/csrc/knn/knn_cpu.cpp, KNearestNeighborIdxCpu) is
 * restated here for K=1:
 *   - for every query point p1[i] scan p2[j] for j = 0..m-1 in ascending order,
 *   - squared L2 distance accumulated dimension by dimension in fp32 from the direct
 *     differences (no |a|^2+|b|^2-2ab expansion),
 *   - the candidate replaces the incumbent only on strict '<'  => lowest index wins ties,
 *     and the very first candidate (j = 0) is always taken (the queue is not yet full), so a
 *     NaN distance at j = 0 sticks,
 *   - outputs: squared distance (fp32) and index (int64).
 *
 * The rounding of the 3-term sum is not pinned by the C++ source (it depends on whether the
 * compiler contracts mul+add into fma), so the oracle DEFINES it, in two flavours:
 *   mode 0 ("fma", the parity contract of the CUDA kernel):
 *        d = fmaf(dz, dz, fmaf(dy, dy, dx * dx))
 *   mode 1 ("sep", what a baseline x86-64 build of pytorch3d evaluates):
 *        d = (dx*dx + dy*dy) + dz*dz     each operation rounded separately
 * Build with -ffp-contract=off so the compiler adds no contraction of its own.
 *
 * threads > 1 parallelises over queries with OpenMP (pytorch3d's CPU kNN is single-threaded;
 * the parallel variant exists so that the CPU baseline is not flattered by that limitation).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float sqdist_fma(const float *a, const float *b) {
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

static inline float sqdist_sep(const float *a, const float *b) {
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    float s = dx * dx;
    float t = dy * dy;
    s = s + t;
    t = dz * dz;
    return s + t;
}

/* p1: [n,3], p2: [m,3] row-major fp32.  d2: [n] fp32, idx: [n] int64.  m must be >= 1. */
void ndp_oracle_knn1(const float *p1, int64_t n, const float *p2, int64_t m,
                     float *d2, int64_t *idx, int mode, int threads) {
    if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t i = 0; i < n; ++i) {
        const float *q = p1 + 3 * i;
        /* first candidate is always accepted (queue not full yet) */
        float best = mode ? sqdist_sep(q, p2) : sqdist_fma(q, p2);
        int64_t bi = 0;
        if (mode) {
            for (int64_t j = 1; j < m; ++j) {
                float d = sqdist_sep(q, p2 + 3 * j);
                if (d < best) { best = d; bi = j; }
            }
        } else {
            for (int64_t j = 1; j < m; ++j) {
                float d = sqdist_fma(q, p2 + 3 * j);
                if (d < best) { best = d; bi = j; }
            }
        }
        d2[i] = best;
        idx[i] = bi;
    }
}

/* Count how many of the n queries get a different index under mode 0 and mode 1
 * (documents how often FMA contraction flips a near-tie). */
int64_t ndp_oracle_knn1_mode_disagreements(const float *p1, int64_t n, const float *p2, int64_t m) {
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *q = p1 + 3 * i;
        float b0 = sqdist_fma(q, p2), b1 = sqdist_sep(q, p2);
        int64_t i0 = 0, i1 = 0;
        for (int64_t j = 1; j < m; ++j) {
            float d0 = sqdist_fma(q, p2 + 3 * j);
            float d1 = sqdist_sep(q, p2 + 3 * j);
            if (d0 < b0) { b0 = d0; i0 = j; }
            if (d1 < b1) { b1 = d1; i1 = j; }
        }
        cnt += (i0 != i1);
    }
    return cnt;
}

int ndp_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
