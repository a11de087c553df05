// tcgen05 (5th-generation tensor core) + TMEM primitives for the warp-field kernels, and the
// shared-memory "image" layout they use.
//
// Precision scheme: every fp32 operand x is split into two fp16 terms x ~= x1 + x2 (11 + 11
// significand bits, |x - x1 - x2| <= 2^-22 |x|, absolute floor 2^-25) and the product is accumulated
// in fp32 in TMEM from the three partial products x1y1, x1y2, x2y1 (the dropped x2y2 is O(2^-22)):
// ~fp32-level accuracy (the fp32 pipes' own rounding of a K = 128 dot product is of the same order)
// at a third of the dense fp16 issue rate, which keeps the 1e-4 parity contract that single-pass
// TF32/BF16 would break.  fp16 has a narrow exponent range, so
//   * activations / weights are converted with saturation (|x| <= 65504; the network's
//     activations are O(1)), and a positive activation is never flushed to zero (relu' is read back
//     from the image: see ndp_relu_img);
//   * back-propagated deltas (O(1e-8)) are multiplied by an exact per-tile power of two chosen so
//     that the tile's largest head gradient lies in [1, 2); the factor is undone in the epilogues.
//
// Image layout: a [128][C] fp16 matrix is stored as 8x8 "core matrices" of 8 rows x 16 bytes,
//     byte offset(r, c) = (r/8) * RS + (c/8) * 128 + (r%8) * 16 + (c%8) * 2,   RS = (C/8) * 128
// which is BOTH canonical no-swizzle UMMA layouts at once: as a K-major operand (K = c) with
// LBO = 128, SBO = RS, and as an MN-major operand (MN = c, K = r) with SBO = 128, LBO = RS.  One
// image of the weights therefore serves h W^T (forward) and delta W (backward), one image of delta
// serves delta W and delta^T h, without any transposed copy.
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#pragma once
#include "ndp_common.cuh"

#define NDP_IMG_CS 128                                   // bytes between 8-column chunks
#define NDP_IMG_RS(C) (((C) / 8) * 128)                  // bytes between 8-row groups
#define NDP_IMG_BYTES(C) (16 * NDP_IMG_RS(C))            // 128 rows
#define NDP_IMG128 NDP_IMG_BYTES(128)                    // 32768
#define NDP_SET128 (2 * NDP_IMG128)                      // 65536: hi / lo images of a [128][128] operand

__device__ __forceinline__ unsigned ndp_img_off(int r, int c, int rs) {
    return (unsigned)((r >> 3) * rs + (c >> 3) * NDP_IMG_CS + (r & 7) * 16 + (c & 7) * 2);
}

// relu that keeps every positive value representable in the fp16 hi image (>= 2^-24, the smallest
// fp16 subnormal), so that relu'(h) can be recovered from the saved image as "hi > 0"
__device__ __forceinline__ float ndp_relu_img(float v) { return v > 0.0f ? fmaxf(v, 5.9604644775390625e-08f) : 0.0f; }

#ifdef NDP_EMU
static inline unsigned ndp_f16_rn_sat(float x) {
    if (x == x) x = fminf(fmaxf(x, -65504.0f), 65504.0f);
    _Float16 h = (_Float16)x;
    unsigned short u; memcpy(&u, &h, 2);
    return u;
}
static inline float ndp_f16_to_f32(unsigned h) { unsigned short u = (unsigned short)h; _Float16 f; memcpy(&f, &u, 2); return (float)f; }
static inline unsigned ndp_pack2_f16(float lo, float hi) { return (ndp_f16_rn_sat(hi) << 16) | ndp_f16_rn_sat(lo); }
static inline void ndp_unpack2_f16(unsigned w, float& lo, float& hi) { lo = ndp_f16_to_f32(w & 0xffffu); hi = ndp_f16_to_f32(w >> 16); }
#else
// two fp32 -> packed f16x2 (round to nearest even, saturating to +-65504): ONE instruction
__device__ __forceinline__ unsigned ndp_pack2_f16(float lo, float hi) {
    unsigned d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ void ndp_unpack2_f16(unsigned w, float& lo, float& hi) {
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(w));
}
__device__ __forceinline__ float ndp_f16_to_f32(unsigned h) { float lo, hi; ndp_unpack2_f16(h & 0xffffu, lo, hi); return lo; }
#endif

// 2-way split of a pair of values into two packed f16x2 words (hi and lo parts)
__device__ __forceinline__ void ndp_split2_pair(float x0, float x1, unsigned& w1, unsigned& w2) {
    w1 = ndp_pack2_f16(x0, x1);
    float f0, f1, r0, r1;
    ndp_unpack2_f16(w1, f0, f1);
    ndp_f2_get(ndp_f2_sub(ndp_f2_make(x0, x1), ndp_f2_make(f0, f1)), r0, r1);     // one FADD2 for both residuals
    w2 = ndp_pack2_f16(r0, r1);
}
// relu(v + b) of a value pair, split into its packed hi / lo f16x2 words; the adds run on the packed FP32 pipe
// (FADD2: one instruction per pair).  Plain max(x, 0): for kernels that do not read relu' back from the image.
__device__ __forceinline__ void ndp_bias_relu_split2(float v0, float v1, float b0, float b1, unsigned& w1, unsigned& w2) {
    float x0, x1;
    ndp_f2_get(ndp_f2_add(ndp_f2_make(v0, v1), ndp_f2_make(b0, b1)), x0, x1);
    x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f);
    w1 = ndp_pack2_f16(x0, x1);
    float f0, f1, r0, r1;
    ndp_unpack2_f16(w1, f0, f1);
    ndp_f2_get(ndp_f2_sub(ndp_f2_make(x0, x1), ndp_f2_make(f0, f1)), r0, r1);
    w2 = ndp_pack2_f16(r0, r1);
}
// one value -> its two fp16 bit patterns
__device__ __forceinline__ void ndp_split2(float x, unsigned& h1, unsigned& h2) {
    unsigned w1, w2;
    ndp_split2_pair(x, 0.0f, w1, w2);
    h1 = w1 & 0xffffu; h2 = w2 & 0xffffu;
}

// 16-byte shared-memory store / load that cannot degrade to a generic ST / LD when the compiler loses the address space of
// a pointer (buffers selected at run time): explicit st.shared / ld.shared on the 32-bit shared address
#ifdef NDP_EMU
static inline void ndp_sts128(void* p, const uint4& v) { *(uint4*)p = v; }
static inline uint4 ndp_lds128(const void* p) { return *(const uint4*)p; }
#else
__device__ __forceinline__ void ndp_sts128(void* p, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ndp_smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ndp_lds128(const void* p) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ndp_smem_u32(p)) : "memory");
    return v;
}
#endif
// split 8 consecutive values and store them as one 16-byte chunk in each of the two images
__device__ __forceinline__ void ndp_store_chunk2(unsigned char* set, unsigned img_bytes, unsigned off, const float (&v)[8]) {
    uint4 p0, p1;
    ndp_split2_pair(v[0], v[1], p0.x, p1.x);
    ndp_split2_pair(v[2], v[3], p0.y, p1.y);
    ndp_split2_pair(v[4], v[5], p0.z, p1.z);
    ndp_split2_pair(v[6], v[7], p0.w, p1.w);
    ndp_sts128(set + off, p0);
    ndp_sts128(set + img_bytes + off, p1);
}
// the 8 values of a 16-byte chunk re-assembled from the two images
__device__ __forceinline__ void ndp_load_chunk2(const unsigned char* set, unsigned img_bytes, unsigned off, float (&v)[8]) {
    const uint4 q0 = *(const uint4*)(set + off), q1 = *(const uint4*)(set + img_bytes + off);
    const unsigned w0[4] = {q0.x, q0.y, q0.z, q0.w}, w1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float a0, a1, b0, b1;
        ndp_unpack2_f16(w0[k], a0, a1);
        ndp_unpack2_f16(w1[k], b0, b1);
        v[2 * k] = a0 + b0; v[2 * k + 1] = a1 + b1;
    }
}

// bit j of the result: element j (0..7) of a 16-byte fp16 chunk is > 0  (relu' of a saved activation)
__device__ __forceinline__ unsigned ndp_pos_mask8(const uint4& q) {
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned lo = w[k] & 0xffffu, hi = w[k] >> 16;
        m |= ((lo - 1u) < 0x7fffu ? 1u : 0u) << (2 * k);          // 0x0001 .. 0x7fff: positive, non-zero
        m |= ((hi - 1u) < 0x7fffu ? 1u : 0u) << (2 * k + 1);
    }
    return m;
}

// exact power-of-two scale that brings m (finite, > 0) into [1, 2); 1 for m == 0 / non-finite
__device__ __forceinline__ void ndp_pow2_scale(float m, float& scale, float& inv_scale) {
    int e = (int)((__float_as_uint(m) >> 23) & 0xffu) - 127;
    if (!(m > 0.0f) || e == 128) e = 0;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    scale = __uint_as_float((unsigned)(127 - e) << 23);
    inv_scale = __uint_as_float((unsigned)(127 + e) << 23);
}

// instruction descriptor: D = f32, A = B = f16 (format code 0), M x N tile, operand majors (0 = K-major, 1 = MN-major)
__device__ __forceinline__ unsigned ndp_idesc_f16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
           ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

#ifdef NDP_EMU
// ---------------------------------------------------------------- CPU emulation (tests only)
struct NdpUmmaDesc { const unsigned char* p; unsigned lbo, sbo; };
namespace ndp_emu { extern float tmem[128][512]; }
static inline NdpUmmaDesc ndp_umma_desc(const void* p, unsigned lbo, unsigned sbo) { return NdpUmmaDesc{(const unsigned char*)p, lbo, sbo}; }
static inline NdpUmmaDesc ndp_umma_desc_adv(NdpUmmaDesc d, unsigned bytes) { d.p += bytes; return d; }
static inline unsigned ndp_tmem_alloc(unsigned* slot, int) { *slot = 0; return 0; }
static inline void ndp_tmem_dealloc(unsigned, int) {}
static inline void ndp_tc_fence_before() {}
static inline void ndp_tc_fence_after() {}
static inline void ndp_umma_f16(unsigned tmem_d, NdpUmmaDesc da, NdpUmmaDesc db, unsigned idesc, unsigned acc) {
    const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
    const int a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
    const int col0 = (int)(tmem_d & 0xffff), lane0 = (int)(tmem_d >> 16);
    auto ld = [](NdpUmmaDesc d, int mn, int k, int is_mn) -> float {
        const unsigned off = is_mn ? (unsigned)((mn >> 3) * d.sbo + (mn & 7) * 2 + (k >> 3) * d.lbo + (k & 7) * 16)
                                   : (unsigned)((mn >> 3) * d.sbo + (mn & 7) * 16 + (k >> 3) * d.lbo + (k & 7) * 2);
        unsigned short h; memcpy(&h, d.p + off, 2);
        return ndp_f16_to_f32(h);
    };
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = acc ? ndp_emu::tmem[lane0 + m][col0 + n] : 0.0f;
            for (int k = 0; k < 16; ++k) s += ld(da, m, k, a_mn) * ld(db, n, k, b_mn);
            ndp_emu::tmem[lane0 + m][col0 + n] = s;
        }
}
static inline void ndp_umma_f16_akeep(unsigned d, NdpUmmaDesc da, NdpUmmaDesc db, unsigned idesc, unsigned acc) { ndp_umma_f16(d, da, db, idesc, acc); }
static inline void ndp_umma_f16_areuse(unsigned d, NdpUmmaDesc da, NdpUmmaDesc db, unsigned idesc, unsigned acc) { ndp_umma_f16(d, da, db, idesc, acc); }
static inline void ndp_umma_commit(NdpMbar* b) { __atomic_fetch_add((unsigned*)&b->phase, 1u, __ATOMIC_SEQ_CST); }
static inline void ndp_tmem_ld32(unsigned taddr, float (&v)[32]) {
    const int col0 = (int)(taddr & 0xffff), lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31);
    for (int j = 0; j < 32; ++j) v[j] = ndp_emu::tmem[lane][col0 + j];
}
static inline void ndp_tmem_ld16(unsigned taddr, float (&v)[16]) {
    const int col0 = (int)(taddr & 0xffff), lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31);
    for (int j = 0; j < 16; ++j) v[j] = ndp_emu::tmem[lane][col0 + j];
}
template <int N> static inline void ndp_tmem_st(unsigned taddr, const unsigned (&r)[N]) {
    const int col0 = (int)(taddr & 0xffff), lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31);
    for (int j = 0; j < N; ++j) memcpy(&ndp_emu::tmem[lane][col0 + j], &r[j], 4);
}
static inline void ndp_tmem_wait_st() {}
// A operand in TMEM: element (m, k) of the 16-deep step = half (k & 1) of the 32-bit cell [lane m][column k / 2]
static inline void ndp_umma_f16_ta(unsigned tmem_d, unsigned tmem_a, NdpUmmaDesc db, unsigned idesc, unsigned acc) {
    const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
    const int b_mn = (idesc >> 16) & 1;
    const int col0 = (int)(tmem_d & 0xffff), lane0 = (int)(tmem_d >> 16), acol0 = (int)(tmem_a & 0xffff), alane0 = (int)(tmem_a >> 16);
    auto ldb = [](NdpUmmaDesc d, int mn, int k, int is_mn) -> float {
        const unsigned off = is_mn ? (unsigned)((mn >> 3) * d.sbo + (mn & 7) * 2 + (k >> 3) * d.lbo + (k & 7) * 16)
                                   : (unsigned)((mn >> 3) * d.sbo + (mn & 7) * 16 + (k >> 3) * d.lbo + (k & 7) * 2);
        unsigned short h; memcpy(&h, d.p + off, 2);
        return ndp_f16_to_f32(h);
    };
    for (int m = 0; m < M; ++m) {
        float av[16];
        for (int k = 0; k < 16; ++k) {
            unsigned cell; memcpy(&cell, &ndp_emu::tmem[alane0 + m][acol0 + (k >> 1)], 4);
            av[k] = ndp_f16_to_f32((k & 1) ? (cell >> 16) : (cell & 0xffffu));
        }
        for (int n = 0; n < N; ++n) {
            float sacc = acc ? ndp_emu::tmem[lane0 + m][col0 + n] : 0.0f;
            for (int k = 0; k < 16; ++k) sacc += av[k] * ldb(db, n, k, b_mn);
            ndp_emu::tmem[lane0 + m][col0 + n] = sacc;
        }
    }
}
static inline void ndp_bulk_s2g(void* g, const void* s, unsigned bytes) { memcpy(g, s, bytes); }
static inline void ndp_bulk_commit() {}
static inline void ndp_bulk_wait_read0() {}
static inline void ndp_bulk_wait0() {}
#else
// ---------------------------------------------------------------- sm_100a
typedef unsigned long long NdpUmmaDesc;
__device__ __forceinline__ NdpUmmaDesc ndp_umma_desc(const void* p, unsigned lbo, unsigned sbo) {
    const unsigned saddr = ndp_smem_u32(p);
    unsigned long long d = (unsigned long long)((saddr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(lbo >> 4) << 16;      // leading-dimension byte offset
    d |= (unsigned long long)(sbo >> 4) << 32;      // stride-dimension byte offset
    d |= 1ull << 46;                                // descriptor version (Blackwell); no swizzle, base offset 0
    return d;
}
__device__ __forceinline__ NdpUmmaDesc ndp_umma_desc_adv(NdpUmmaDesc d, unsigned bytes) { return d + (bytes >> 4); }
// warp-collective: allocate ncols (power of two >= 32) TMEM columns, base address written to *slot
__device__ __forceinline__ void ndp_tmem_alloc_warp(unsigned* slot, int ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ndp_smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void ndp_tmem_dealloc(unsigned taddr, int ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void ndp_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ndp_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ndp_umma_f16(unsigned tmem_d, NdpUmmaDesc da, NdpUmmaDesc db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// the same with the A operand kept in / taken from the tensor core's collector buffer: KEEP on the first of two
// consecutive MMAs with the SAME A descriptor, REUSE on the second -- the second one does not fetch A from shared memory
// (SASS UTCHMMA ... .A_KEEP / .A_REUSE)
__device__ __forceinline__ void ndp_umma_f16_akeep(unsigned tmem_d, NdpUmmaDesc da, NdpUmmaDesc db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void ndp_umma_f16_areuse(unsigned tmem_d, NdpUmmaDesc da, NdpUmmaDesc db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// arrive on the mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void ndp_umma_commit(NdpMbar* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ndp_smem_u32(b)) : "memory");
}
// this warp's 32 TMEM lanes x 32 consecutive columns -> 32 registers per thread (blocking)
__device__ __forceinline__ void ndp_tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void ndp_tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
// this warp's 32 TMEM lanes x N consecutive columns <- N registers per thread (N = 8 or 16)
template <int N> __device__ __forceinline__ void ndp_tmem_st(unsigned taddr, const unsigned (&r)[N]);
template <> __device__ __forceinline__ void ndp_tmem_st<8>(unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <> __device__ __forceinline__ void ndp_tmem_st<16>(unsigned taddr, const unsigned (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ndp_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand in TMEM (M = 128 lanes, 16-bit elements packed two per 32-bit column, K along the columns)
__device__ __forceinline__ void ndp_umma_f16_ta(unsigned tmem_d, unsigned tmem_a, NdpUmmaDesc db, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// bulk asynchronous copy shared -> global (TMA store, SASS UBLKCP) in the calling thread's bulk group
__device__ __forceinline__ void ndp_bulk_s2g(void* g, const void* s, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(ndp_smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ndp_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ndp_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ndp_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

#ifdef NDP_EMU
static inline void ndp_tmem_alloc_warp(unsigned* slot, int n) { ndp_tmem_alloc(slot, n); }
#endif

// One fp32-accurate product D[128 x N] (+)= A . B over K = 16 * ksteps from hi/lo image sets:
// the three fp16 partial products, issued by ONE thread (call from a branch on ndp_elect_one(), see ndp_common.cuh).  a_step / b_step: descriptor byte advance
// per 16-deep k-step; a_img / b_img: byte distance between the hi and lo images.
// ndp_umma_gemm_a2: only the hi image of B (for operands that are exact in fp16, e.g. a column of ones).
#ifdef NDP_EMU
static inline
#else
static __device__ __forceinline__
#endif
void ndp_umma_gemm_a2(unsigned tmem_d, NdpUmmaDesc a0, unsigned a_img, unsigned a_step, NdpUmmaDesc b0, unsigned b_step,
                      int ksteps, unsigned idesc) {
    unsigned acc = 0u;
#pragma unroll 1
    for (int i = 1; i >= 0; --i) {
        NdpUmmaDesc da = ndp_umma_desc_adv(a0, i * a_img), db = b0;
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
            ndp_umma_f16(tmem_d, da, db, idesc, acc);
            acc = 1u;
            da = ndp_umma_desc_adv(da, a_step);
            db = ndp_umma_desc_adv(db, b_step);
        }
    }
}

#ifdef NDP_EMU
static inline
#else
static __device__ __forceinline__
#endif
void ndp_umma_gemm3(unsigned tmem_d, NdpUmmaDesc a0, unsigned a_img, unsigned a_step,
                    NdpUmmaDesc b0, unsigned b_img, unsigned b_step, int ksteps,
                    unsigned idesc, bool accumulate) {
    unsigned acc = accumulate ? 1u : 0u;
    // small cross terms first, the hi x hi product last
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
        const int i = (t == 0) ? 1 : 0, j = (t == 1) ? 1 : 0;
        NdpUmmaDesc da = ndp_umma_desc_adv(a0, i * a_img), db = ndp_umma_desc_adv(b0, j * b_img);
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks) {
            ndp_umma_f16(tmem_d, da, db, idesc, acc);
            acc = 1u;
            da = ndp_umma_desc_adv(da, a_step);
            db = ndp_umma_desc_adv(db, b_step);
        }
    }
}

// The same product issued k-step by k-step so that the two partial products that share the A hi chunk are adjacent:
// hi x lo (A kept in the collector), hi x hi (A reused: no second fetch of the 4 KB chunk from shared memory), lo x hi.
// For kernels bound by the shared-memory pipe (the recomputing backward): a third of the A-side operand traffic less.
#ifdef NDP_EMU
static inline
#else
static __device__ __forceinline__
#endif
void ndp_umma_gemm3_ar(unsigned tmem_d, NdpUmmaDesc a0, unsigned a_img, unsigned a_step,
                       NdpUmmaDesc b0, unsigned b_img, unsigned b_step, int ksteps,
                       unsigned idesc, bool accumulate) {
    unsigned acc = accumulate ? 1u : 0u;
    NdpUmmaDesc da = a0, db = b0;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
        ndp_umma_f16_akeep(tmem_d, da, ndp_umma_desc_adv(db, b_img), idesc, acc);
        ndp_umma_f16_areuse(tmem_d, da, db, idesc, 1u);
        ndp_umma_f16(tmem_d, ndp_umma_desc_adv(da, a_img), db, idesc, 1u);
        acc = 1u;
        da = ndp_umma_desc_adv(da, a_step);
        db = ndp_umma_desc_adv(db, b_step);
    }
}

// A operand in TMEM with explicit geometry: the hi words of k-step j sit at ta0 + j * ta_step (8 columns = 16 fp16), the lo
// words ta_img columns further; B from shared memory.  hi x lo, hi x hi, lo x hi per k-step.
#ifdef NDP_EMU
static inline
#else
static __device__ __forceinline__
#endif
void ndp_umma_gemm3_tak(unsigned tmem_d, unsigned ta0, unsigned ta_img, unsigned ta_step, NdpUmmaDesc b0, unsigned b_img,
                        unsigned b_step, int ksteps, unsigned idesc, bool accumulate) {
    unsigned acc = accumulate ? 1u : 0u, ta = ta0;
    NdpUmmaDesc db = b0;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ++ks) {
        ndp_umma_f16_ta(tmem_d, ta, ndp_umma_desc_adv(db, b_img), idesc, acc);
        ndp_umma_f16_ta(tmem_d, ta, db, idesc, 1u);
        ndp_umma_f16_ta(tmem_d, ta + ta_img, db, idesc, 1u);
        acc = 1u;
        ta += ta_step;
        db = ndp_umma_desc_adv(db, b_step);
    }
}

// Same with the A operand in TMEM: ta0 = TMEM address of the hi image (8 columns per 16-deep k-step),
// ta_img = column distance to the lo image.
#ifdef NDP_EMU
static inline
#else
static __device__ __forceinline__
#endif
void ndp_umma_gemm3_ta(unsigned tmem_d, unsigned ta0, unsigned ta_img, NdpUmmaDesc b0, unsigned b_img, unsigned b_step,
                       int ksteps, unsigned idesc) {
    unsigned acc = 0u;
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
        const int i = (t == 0) ? 1 : 0, j = (t == 1) ? 1 : 0;
        unsigned ta = ta0 + (unsigned)i * ta_img;
        NdpUmmaDesc db = ndp_umma_desc_adv(b0, j * b_img);
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ++ks) {
            ndp_umma_f16_ta(tmem_d, ta, db, idesc, acc);
            acc = 1u;
            ta += 8u;
            db = ndp_umma_desc_adv(db, b_step);
        }
    }
}
