// TEST INFRASTRUCTURE ONLY -- host-side entry points to the per-point math of
// deformationpyramid_b200/csrc/ndp_math.cuh (compiled as plain C++), so that the closed-form
// rotation / warp derivatives can be compared with torch autograd on the CPU.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include "ndp_math.cuh"

extern "C" void ndp_hook_point_forward(int motion, int rot, int nonrigid, const float* z, const float* x,
                                       long long n, float* y, float* nu) {
    const int zd = NDP_MAX_HEAD;
    for (long long i = 0; i < n; ++i) {
        float v = 0.0f;
        ndp_point_forward(motion, rot, nonrigid, z + i * zd, x + i * 3, y + i * 3, &v);
        if (nu) nu[i] = v;
    }
}

extern "C" void ndp_hook_point_backward(int motion, int rot, int nonrigid, const float* z, const float* x,
                                        const float* gy, const float* gnu, long long n, float* gz, float* gx) {
    const int zd = NDP_MAX_HEAD;
    for (long long i = 0; i < n; ++i) {
        float g[NDP_MAX_HEAD];
        for (int r = 0; r < zd; ++r) g[r] = 0.0f;
        ndp_point_backward(motion, rot, nonrigid, z + i * zd, x + i * 3, gy + i * 3, gnu ? gnu[i] : 0.0f, g, gx + i * 3);
        for (int r = 0; r < zd; ++r) gz[i * zd + r] = g[r];
    }
}

extern "C" int ndp_hook_head_dim(int motion, int rot, int nonrigid) { return ndp_head_idx(motion, rot, nonrigid).dim; }
