// Per-point warp math of one pyramid level: head outputs -> rotation -> warp, and its closed-form
// backward.  Compiles as device code (nvcc) and as host code (g++, for tests/test_point_math.py),
// so the derivatives are checked against torch autograd without a GPU.
//
// Reference (all paths relative to /root/reference):
//   model/nets.py:111-140   NDPLayer.forward      (SE3 / Sim3 / sflow composition, nonrigidity blend)
//   model/nets.py:144-161   NDPLayer.get_Rotation (euler / axis_angle / quaternion / 6D)
//   model/rigid_body.py:5-16, 19-56, 58-85, 89-95, 113-119
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define NDP_HD __host__ __device__ __forceinline__
#else
#define NDP_HD inline
#endif

enum { NDP_MOTION_SE3 = 0, NDP_MOTION_SIM3 = 1, NDP_MOTION_SFLOW = 2 };
enum { NDP_ROT_AXIS_ANGLE = 0, NDP_ROT_EULER = 1, NDP_ROT_QUATERNION = 2, NDP_ROT_6D = 3 };

#define NDP_MAX_HEAD 12   // rot (<=6) + scale (1) + translation (3) + nonrigidity (1) = 11, padded

NDP_HD int ndp_rot_dim(int motion, int rot) {
    if (motion == NDP_MOTION_SFLOW) return 0;
    return rot == NDP_ROT_QUATERNION ? 4 : (rot == NDP_ROT_6D ? 6 : 3);
}

// Head vector z[] (already scaled by mlp_scale, nets.py:117,125,133,146) is laid out in the order
// of NDPLayer's parameters (nets.py:82-103): [rot(R) | scale(1, Sim3 only) | trn(3) | nonrigid(1)].
struct NdpHeadIdx { int rot, s, t, nr, dim; };
NDP_HD NdpHeadIdx ndp_head_idx(int motion, int rot, int nonrigid) {
    NdpHeadIdx h;
    int R = ndp_rot_dim(motion, rot);
    h.rot = 0;
    h.s = (motion == NDP_MOTION_SIM3) ? R : -1;
    h.t = R + (motion == NDP_MOTION_SIM3 ? 1 : 0);
    h.nr = nonrigid ? h.t + 3 : -1;
    h.dim = h.t + 3 + (nonrigid ? 1 : 0);
    return h;
}

// ---------------------------------------------------------------- rotation forward: a -> R (row-major 3x3)
NDP_HD void ndp_rot_forward(int rot, const float* a, float* R) {
    if (rot == NDP_ROT_AXIS_ANGLE) {
        // nets.py:150-153, rigid_body.py:113-119: theta=|a|, w=a/theta, R = I + sin K + (1-cos) K K
        float th = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        float w0 = a[0] / th, w1 = a[1] / th, w2 = a[2] / th;
        float s, c;
        sincosf(th, &s, &c);
        float v = 1.0f - c;
        R[0] = 1.0f - v * (w1 * w1 + w2 * w2); R[1] = -s * w2 + v * (w0 * w1);         R[2] = s * w1 + v * (w0 * w2);
        R[3] = s * w2 + v * (w0 * w1);         R[4] = 1.0f - v * (w0 * w0 + w2 * w2);  R[5] = -s * w0 + v * (w1 * w2);
        R[6] = -s * w1 + v * (w0 * w2);        R[7] = s * w0 + v * (w1 * w2);          R[8] = 1.0f - v * (w0 * w0 + w1 * w1);
    } else if (rot == NDP_ROT_EULER) {
        // rigid_body.py:19-56: R = Rx(a0) Ry(a1) Rz(a2)
        float sx, cx, sy, cy, sz, cz;
        sincosf(a[0], &sx, &cx); sincosf(a[1], &sy, &cy); sincosf(a[2], &sz, &cz);
        R[0] = cy * cz;                 R[1] = -cy * sz;                R[2] = sy;
        R[3] = sx * sy * cz + cx * sz;  R[4] = -sx * sy * sz + cx * cz; R[5] = -sx * cy;
        R[6] = -cx * sy * cz + sx * sz; R[7] = cx * sy * sz + sx * cz;  R[8] = cx * cy;
    } else if (rot == NDP_ROT_QUATERNION) {
        // nets.py:154-157: q = a / copysign(|a|, a0); rigid_body.py:62-85 with two_s = 2/(q.q)
        float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
        float den = (a[0] < 0.0f) ? -n : n;
        float r = a[0] / den, i = a[1] / den, j = a[2] / den, k = a[3] / den;
        float ts = 2.0f / (r * r + i * i + j * j + k * k);
        R[0] = 1.0f - ts * (j * j + k * k); R[1] = ts * (i * j - k * r);        R[2] = ts * (i * k + j * r);
        R[3] = ts * (i * j + k * r);        R[4] = 1.0f - ts * (i * i + k * k); R[5] = ts * (j * k - i * r);
        R[6] = ts * (i * k - j * r);        R[7] = ts * (j * k + i * r);        R[8] = 1.0f - ts * (i * i + j * j);
    } else {
        // rigid_body.py:5-16 (rows b1, b2, b3; F.normalize clamps the norm at 1e-12)
        float n1 = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
        float b10 = a[0] / n1, b11 = a[1] / n1, b12 = a[2] / n1;
        float d = b10 * a[3] + b11 * a[4] + b12 * a[5];
        float u0 = a[3] - d * b10, u1 = a[4] - d * b11, u2 = a[5] - d * b12;
        float n2 = fmaxf(sqrtf(u0 * u0 + u1 * u1 + u2 * u2), 1e-12f);
        float b20 = u0 / n2, b21 = u1 / n2, b22 = u2 / n2;
        R[0] = b10; R[1] = b11; R[2] = b12;
        R[3] = b20; R[4] = b21; R[5] = b22;
        R[6] = b11 * b22 - b12 * b21; R[7] = b12 * b20 - b10 * b22; R[8] = b10 * b21 - b11 * b20;
    }
}

// ---------------------------------------------------------------- rotation backward: G = dL/dR -> ga = dL/da
NDP_HD void ndp_rot_backward(int rot, const float* a, const float* G, float* ga) {
    if (rot == NDP_ROT_AXIS_ANGLE) {
        float th = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        float w0 = a[0] / th, w1 = a[1] / th, w2 = a[2] / th;
        float s, c;
        sincosf(th, &s, &c);
        float v = 1.0f - c;
        // K and K^2
        float K[9] = {0.0f, -w2, w1, w2, 0.0f, -w0, -w1, w0, 0.0f};
        float K2[9] = {-(w1 * w1 + w2 * w2), w0 * w1, w0 * w2,
                       w0 * w1, -(w0 * w0 + w2 * w2), w1 * w2,
                       w0 * w2, w1 * w2, -(w0 * w0 + w1 * w1)};
        float gth = 0.0f;
#pragma unroll
        for (int e = 0; e < 9; ++e) gth += G[e] * (c * K[e] + s * K2[e]);
        // dL/dK = s G + v (G K^T + K^T G)
        float gK[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float t = 0.0f;
#pragma unroll
                for (int m = 0; m < 3; ++m) t += G[r * 3 + m] * K[q * 3 + m] + K[m * 3 + r] * G[m * 3 + q];
                gK[r * 3 + q] = s * G[r * 3 + q] + v * t;
            }
        float gw0 = gK[7] - gK[5], gw1 = gK[2] - gK[6], gw2 = gK[3] - gK[1];
        float gww = gw0 * w0 + gw1 * w1 + gw2 * w2;
        ga[0] = (gw0 - gww * w0) / th + gth * w0;
        ga[1] = (gw1 - gww * w1) / th + gth * w1;
        ga[2] = (gw2 - gww * w2) / th + gth * w2;
    } else if (rot == NDP_ROT_EULER) {
        float sx, cx, sy, cy, sz, cz;
        sincosf(a[0], &sx, &cx); sincosf(a[1], &sy, &cy); sincosf(a[2], &sz, &cz);
        // dR/da0: rows 1,2 of Rx' (Ry Rz);  Rx' = [[0,0,0],[0,-sx,-cx],[0,cx,-sx]]
        // A = Ry Rz
        float A[9] = {cy * cz, -cy * sz, sy, sz, cz, 0.0f, -sy * cz, sy * sz, cy};
        float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            g0 += G[3 + q] * (-sx * A[3 + q] - cx * A[6 + q]) + G[6 + q] * (cx * A[3 + q] - sx * A[6 + q]);
        }
        // dR/da1 = Rx Ry' Rz ; Ry' = [[-sy,0,cy],[0,0,0],[-cy,0,-sy]];  B = Ry' Rz
        float B[9] = {-sy * cz, sy * sz, cy, 0.0f, 0.0f, 0.0f, -cy * cz, cy * sz, -sy};
        // Rx B: row0 = B row0; row1 = cx*Brow1 - sx*Brow2 ; row2 = sx*Brow1 + cx*Brow2
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            g1 += G[q] * B[q] + G[3 + q] * (-sx * B[6 + q]) + G[6 + q] * (cx * B[6 + q]);
        }
        // dR/da2 = (Rx Ry) Rz' ; Rz' = [[-sz,-cz,0],[cz,-sz,0],[0,0,0]];  C = Rx Ry
        float C[9] = {cy, 0.0f, sy, sx * sy, cx, -sx * cy, -cx * sy, sx, cx * cy};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float d0 = C[r * 3 + 0] * (-sz) + C[r * 3 + 1] * cz;
            float d1 = C[r * 3 + 0] * (-cz) + C[r * 3 + 1] * (-sz);
            g2 += G[r * 3 + 0] * d0 + G[r * 3 + 1] * d1;
        }
        ga[0] = g0; ga[1] = g1; ga[2] = g2;
    } else if (rot == NDP_ROT_QUATERNION) {
        float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
        float den = (a[0] < 0.0f) ? -n : n;
        float r = a[0] / den, i = a[1] / den, j = a[2] / den, k = a[3] / den;
        float nn = r * r + i * i + j * j + k * k;
        float ts = 2.0f / nn;
        // R = I + ts * P(q)
        float P[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r,
                      i * j + k * r, -(i * i + k * k), j * k - i * r,
                      i * k - j * r, j * k + i * r, -(i * i + j * j)};
        float gts = 0.0f;
#pragma unroll
        for (int e = 0; e < 9; ++e) gts += G[e] * P[e];
        float gPr = -k * G[1] + j * G[2] + k * G[3] - i * G[5] - j * G[6] + i * G[7];
        float gPi = j * (G[1] + G[3]) + k * (G[2] + G[6]) - 2.0f * i * (G[4] + G[8]) + r * (G[7] - G[5]);
        float gPj = -2.0f * j * (G[0] + G[8]) + i * (G[1] + G[3]) + r * (G[2] - G[6]) + k * (G[5] + G[7]);
        float gPk = -2.0f * k * (G[0] + G[4]) + r * (G[3] - G[1]) + i * (G[2] + G[6]) + j * (G[5] + G[7]);
        float c2 = ts * ts * gts;              // d ts / d q_c = -ts^2 q_c
        float gq0 = ts * gPr - c2 * r, gq1 = ts * gPi - c2 * i, gq2 = ts * gPj - c2 * j, gq3 = ts * gPk - c2 * k;
        float gqq = gq0 * r + gq1 * i + gq2 * j + gq3 * k;
        ga[0] = (gq0 - gqq * r) / den; ga[1] = (gq1 - gqq * i) / den;
        ga[2] = (gq2 - gqq * j) / den; ga[3] = (gq3 - gqq * k) / den;
    } else {
        float m1 = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        float n1 = fmaxf(m1, 1e-12f);
        float b1[3] = {a[0] / n1, a[1] / n1, a[2] / n1};
        float d = b1[0] * a[3] + b1[1] * a[4] + b1[2] * a[5];
        float u[3] = {a[3] - d * b1[0], a[4] - d * b1[1], a[5] - d * b1[2]};
        float m2 = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        float n2 = fmaxf(m2, 1e-12f);
        float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
        float gb1[3] = {G[0], G[1], G[2]}, gb2[3] = {G[3], G[4], G[5]};
        const float* gb3 = G + 6;
        // b3 = b1 x b2:  gb1 += b2 x gb3 ; gb2 += gb3 x b1
        gb1[0] += b2[1] * gb3[2] - b2[2] * gb3[1];
        gb1[1] += b2[2] * gb3[0] - b2[0] * gb3[2];
        gb1[2] += b2[0] * gb3[1] - b2[1] * gb3[0];
        gb2[0] += gb3[1] * b1[2] - gb3[2] * b1[1];
        gb2[1] += gb3[2] * b1[0] - gb3[0] * b1[2];
        gb2[2] += gb3[0] * b1[1] - gb3[1] * b1[0];
        // b2 = u / max(|u|, eps)
        float gu[3];
        if (m2 > 1e-12f) {
            float t = gb2[0] * b2[0] + gb2[1] * b2[1] + gb2[2] * b2[2];
#pragma unroll
            for (int e = 0; e < 3; ++e) gu[e] = (gb2[e] - t * b2[e]) / n2;
        } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) gu[e] = gb2[e] / n2;
        }
        // u = a2 - d b1 ; d = b1 . a2
        float gd = -(gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2]);
        float ga2[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            ga2[e] = gu[e] + gd * b1[e];
            gb1[e] += -d * gu[e] + gd * a[3 + e];
        }
        if (m1 > 1e-12f) {
            float t = gb1[0] * b1[0] + gb1[1] * b1[1] + gb1[2] * b1[2];
#pragma unroll
            for (int e = 0; e < 3; ++e) ga[e] = (gb1[e] - t * b1[e]) / n1;
        } else {
#pragma unroll
            for (int e = 0; e < 3; ++e) ga[e] = gb1[e] / n1;
        }
        ga[3] = ga2[0]; ga[4] = ga2[1]; ga[5] = ga2[2];
    }
}

NDP_HD float ndp_sigmoid(float z) { return 1.0f / (1.0f + expf(-z)); }

// ---------------------------------------------------------------- point forward
// z: scaled head outputs (layout NdpHeadIdx), x: input point.  y: warped point, *nu: nonrigidity
// (written only when nonrigid).  nets.py:117-137.
NDP_HD void ndp_point_forward(int motion, int rot, int nonrigid, const float* z, const float* x,
                              float* y, float* nu) {
    NdpHeadIdx h = ndp_head_idx(motion, rot, nonrigid);
    const float* t = z + h.t;
    if (motion == NDP_MOTION_SFLOW) {
        y[0] = x[0] + t[0]; y[1] = x[1] + t[1]; y[2] = x[2] + t[2];
    } else {
        float R[9];
        ndp_rot_forward(rot, z + h.rot, R);
        float r0 = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
        float r1 = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
        float r2 = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
        if (motion == NDP_MOTION_SIM3) {
            float s = z[h.s] + 1.0f;                      // nets.py:125
            r0 *= s; r1 *= s; r2 *= s;
        }
        y[0] = r0 + t[0]; y[1] = r1 + t[1]; y[2] = r2 + t[2];
    }
    if (nonrigid) {
        float v = ndp_sigmoid(z[h.nr]);                   // nets.py:133
        y[0] = x[0] + v * (y[0] - x[0]);                  // nets.py:134
        y[1] = x[1] + v * (y[1] - x[1]);
        y[2] = x[2] + v * (y[2] - x[2]);
        *nu = v;
    }
}

// ---------------------------------------------------------------- point backward
// gy = dL/dy' (the level output), gnu = dL/dnu coming from outside the layer (e.g. the BCE
// regulariser, registration.py:216-220).  Outputs gz[h.dim] = dL/dz and gx = the DIRECT part of
// dL/dx (through R x, the identity paths), i.e. everything except the path through the
// positional encoding, which the caller adds from the MLP input gradient.
NDP_HD void ndp_point_backward(int motion, int rot, int nonrigid, const float* z, const float* x,
                               const float* gy_in, float gnu, float* gz, float* gx) {
    NdpHeadIdx h = ndp_head_idx(motion, rot, nonrigid);
    float gy[3] = {gy_in[0], gy_in[1], gy_in[2]};
    gx[0] = gx[1] = gx[2] = 0.0f;
    float R[9];
    float rx[3] = {0.0f, 0.0f, 0.0f};
    float s = 1.0f;
    if (motion != NDP_MOTION_SFLOW) {
        ndp_rot_forward(rot, z + h.rot, R);
        rx[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
        rx[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
        rx[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
        if (motion == NDP_MOTION_SIM3) s = z[h.s] + 1.0f;
    }
    if (nonrigid) {
        // y' = x + nu (y - x): recompute the rigid-part output y
        const float* t = z + h.t;
        float y[3];
        if (motion == NDP_MOTION_SFLOW) {
            y[0] = x[0] + t[0]; y[1] = x[1] + t[1]; y[2] = x[2] + t[2];
        } else {
            y[0] = s * rx[0] + t[0]; y[1] = s * rx[1] + t[1]; y[2] = s * rx[2] + t[2];
        }
        float v = ndp_sigmoid(z[h.nr]);
        float gv = gnu + gy[0] * (y[0] - x[0]) + gy[1] * (y[1] - x[1]) + gy[2] * (y[2] - x[2]);
        gz[h.nr] = gv * v * (1.0f - v);
#pragma unroll
        for (int e = 0; e < 3; ++e) { gx[e] += (1.0f - v) * gy[e]; gy[e] *= v; }
    }
    gz[h.t] = gy[0]; gz[h.t + 1] = gy[1]; gz[h.t + 2] = gy[2];
    if (motion == NDP_MOTION_SFLOW) {
        gx[0] += gy[0]; gx[1] += gy[1]; gx[2] += gy[2];
        return;
    }
    if (motion == NDP_MOTION_SIM3) gz[h.s] = gy[0] * rx[0] + gy[1] * rx[1] + gy[2] * rx[2];
    float G[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) G[r * 3 + q] = s * gy[r] * x[q];
    ndp_rot_backward(rot, z + h.rot, G, gz + h.rot);
    gx[0] += s * (R[0] * gy[0] + R[3] * gy[1] + R[6] * gy[2]);
    gx[1] += s * (R[1] * gy[0] + R[4] * gy[1] + R[7] * gy[2]);
    gx[2] += s * (R[2] * gy[0] + R[5] * gy[1] + R[8] * gy[2]);
}
