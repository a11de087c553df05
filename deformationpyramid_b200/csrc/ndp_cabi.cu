// C ABI of libndp_b200.so (include/ndp_b200.h): argument checking, workspace carving and the
// host-side orchestration of the kernels.  No torch types, no CPU fallback.
#include "../../include/ndp_b200.h"
#include "ndp_kernels.h"
#include "ndp_tc.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

int ndp_launch_prio[2] = {0, 0};   // per-launch scheduling priorities: [0] tensor-core kernels, [1] the others (NDP_LAUNCH_PRIO)
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (cudaError_t)(call);                                             \
        if (e_ != cudaSuccess)                                                            \
            return fail(NDP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
    } while (0)

extern "C" const char* ndp_last_error(void) { return g_err.c_str(); }
extern "C" int32_t ndp_version(void) { return 100; }

// 0: hidden layers on the tensor cores (tcgen05, fp16 hi/lo split, fp32 accumulate) -- default;
// 1: everything on the FP32 pipes.  Process-wide; solvers capture it at creation.
static int g_mlp_mode = 0;
extern "C" int ndp_set_mlp_mode(int32_t mode) {
    if (mode != 0 && mode != 1) return fail(NDP_E_INVALID, "mlp mode must be 0 (tensor cores) or 1 (fp32 pipes)");
    g_mlp_mode = mode;
    return NDP_OK;
}
extern "C" int32_t ndp_get_mlp_mode(void) { return g_mlp_mode; }
// Work grouping of the standalone layer calls (ndp_layer_forward / ndp_layer_backward); solvers take theirs
// from ndp_solver_cfg.  0 = automatic.
static int g_layer_tpc = 0, g_layer_rounds = 0;
extern "C" int ndp_set_layer_tuning(int32_t tiles_per_bwd_cta, int32_t fwd_rounds) {
    if (tiles_per_bwd_cta < 0 || tiles_per_bwd_cta > 16 || fwd_rounds < 0 || fwd_rounds > 8)
        return fail(NDP_E_INVALID, "tiles_per_bwd_cta must be in [0, 16] and fwd_rounds in [0, 8]");
    g_layer_tpc = tiles_per_bwd_cta; g_layer_rounds = fwd_rounds;
    return NDP_OK;
}

static int init_once() {
    static std::once_flag once;
    static int rc = 0;
    std::call_once(once, [] {
        int e = ndp_fwd_init();
        if (e == 0) e = ndp_bwd_init();
        if (e == 0) e = ndp_fwd_tc_init();
        if (e == 0) e = ndp_bwd_tc_init();
        if (e == 0) e = ndp_bwd_rc_init();
        rc = e;
    });
    if (rc != 0) return fail(NDP_E_CUDA, std::string("kernel init: ") + cudaGetErrorString((cudaError_t)rc));
    return NDP_OK;
}

static int check_cfg(const ndp_layer_cfg* c) {
    if (!c) return fail(NDP_E_INVALID, "cfg is NULL");
    if (c->width != NDP_W) return fail(NDP_E_INVALID, "this build supports width 128 only");
    if (c->depth < 1 || c->depth > NDP_MAX_HIDDEN + 1) return fail(NDP_E_INVALID, "depth must be in [1, 9]");
    if (c->motion < 0 || c->motion > 2) return fail(NDP_E_INVALID, "motion must be SE3, Sim3 or sflow");
    if (c->rot_format < 0 || c->rot_format > 3) return fail(NDP_E_INVALID, "unknown rotation format");
    return NDP_OK;
}
static NdpLayout layout_of(const ndp_layer_cfg* c) {
    return ndp_make_layout(c->depth, c->motion, c->rot_format, c->nonrigidity, c->freq, c->mlp_scale);
}
static bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static long long pad4(long long v) { return (v + 3) / 4 * 4; }
static long long pad256(long long v) { return (v + 255) / 256 * 256; }

extern "C" int64_t ndp_param_count(const ndp_layer_cfg* c) { return check_cfg(c) ? -1 : layout_of(c).param_count; }
extern "C" int64_t ndp_pack_count(const ndp_layer_cfg* c) { return check_cfg(c) ? -1 : layout_of(c).pack_count; }
static long long act_floats(int depth, long long n, int mode) {   // saved activations of one pair
    const long long tiles = (n + NDP_TP - 1) / NDP_TP;
    if (mode == 0 && ndp_tc_recompute(depth - 1)) return 0;    // the backward kernel rebuilds them from x
    return mode == 0 ? tiles * depth * (long long)(NDP_SET128 / 4) : (long long)depth * n * NDP_W;
}
extern "C" int64_t ndp_saved_floats(const ndp_layer_cfg* c, int64_t n) {
    if (check_cfg(c) || n < 0) return -1;
    return act_floats(c->depth, n, g_mlp_mode) + n * NDP_ZPITCH;
}
extern "C" int64_t ndp_backward_workspace_bytes(const ndp_layer_cfg* c, int64_t n) {
    if (check_cfg(c)) return -1;
    const long long tiles = n > 0 ? (n + NDP_TP - 1) / NDP_TP : 1;
    return tiles * (pad4(layout_of(c).param_count) + NDP_HGREC) * (long long)sizeof(float);
}

// ---- nearest-neighbour launch plan -------------------------------------------------------------
struct NnPlan { int chunk_targets, chunks, qpitch, blocks; };
static NnPlan nn_plan(long long n, long long m) {
    NnPlan p;
    const long long big = n > m ? n : m;
    long long k = (big + 16LL * NDP_NN_TS - 1) / (16LL * NDP_NN_TS);
    if (k < 1) k = 1;
    p.chunk_targets = (int)(k * NDP_NN_TS);
    p.chunks = (int)((big + p.chunk_targets - 1) / p.chunk_targets);
    if (p.chunks < 1) p.chunks = 1;
    p.qpitch = (int)pad4(big);
    p.blocks = (int)((big + 31) / 32);          // per-CTA partial sums: 256 points per CTA of the stand-alone epilogue, 128 of the fused search
    if (p.blocks < 1) p.blocks = 1;
    return p;
}
struct ChamferWs { long long part, gacc, sums, counter, total; };
static ChamferWs chamfer_ws(long long n, long long m) {
    NnPlan p = nn_plan(n, m);
    ChamferWs w;
    long long o = 0;
    w.part = o; o += pad256(2LL * p.chunks * p.qpitch * (long long)sizeof(float2));
    w.gacc = o; o += pad256(n * 3 * 8);
    w.sums = o; o += pad256((long long)p.blocks * 2 * 8);
    w.counter = o; o += 256;
    w.total = o;
    return w;
}
extern "C" int64_t ndp_chamfer_workspace_bytes(int64_t n, int64_t m) {
    if (n < 0 || m < 0) return -1;
    return chamfer_ws(n, m).total;
}

extern "C" int ndp_pack_params(const ndp_layer_cfg* c, const float* params, float* pack, void* stream) {
    if (int e = check_cfg(c)) return e;
    if (!params || !pack) return fail(NDP_E_INVALID, "NULL buffer");
    NdpPackArgs a; a.lay = layout_of(c); a.params = params; a.params_stride = 0; a.pack = pack; a.pack_stride = 0; a.npairs = 1;
    ndp_launch_pack(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return NDP_OK;
}

extern "C" int ndp_layer_forward(const ndp_layer_cfg* c, const float* params, const float* pack, const float* x,
                                 int64_t n, float* y, float* nu, float* saved, void* stream) {
    if (int e = check_cfg(c)) return e;
    if (int e = init_once()) return e;
    if (n < 0 || n > 0x7fffffff / 4) return fail(NDP_E_INVALID, "bad point count");
    if (n == 0) return NDP_OK;
    if (!params || !pack || !x || !y) return fail(NDP_E_INVALID, "NULL buffer");
    if (!aligned16(params) || !aligned16(pack) || (saved && !aligned16(saved)))
        return fail(NDP_E_INVALID, "params, pack and saved must be 16-byte aligned");
    NdpFwdArgs a;
    a.lay = layout_of(c);
    a.params = params; a.params_stride = 0; a.pack = pack; a.pack_stride = 0;
    a.x = x; a.x_stride = 0; a.y = y; a.y_stride = 0; a.nu = nu; a.nu_stride = 0;
    const long long actf = act_floats(c->depth, n, g_mlp_mode);
    a.act = (saved && actf > 0) ? saved : nullptr; a.act_stride = 0; a.act_layer_stride = n * NDP_W;
    a.zsave = saved ? saved + actf : nullptr; a.z_stride = 0;
    a.rounds = g_layer_rounds;
    a.y_add = nullptr; a.y_add_stride = 0; a.y4 = nullptr; a.y4_stride = 0; a.orig = nullptr; a.orig_stride = 0;
    a.ybox = nullptr; a.box_stride = 0; a.n = (int)n; a.counts = nullptr; a.state = nullptr; a.npairs = 1;
    if (g_mlp_mode == 0) ndp_launch_fwd_tc(a, (cudaStream_t)stream); else ndp_launch_fwd(a, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return NDP_OK;
}

extern "C" int ndp_layer_backward(const ndp_layer_cfg* c, const float* params, const float* pack, const float* x, int64_t n,
                                  const float* saved, const float* grad_y, const float* grad_nu,
                                  float* grad_params, float* grad_x, void* workspace, void* stream) {
    if (int e = check_cfg(c)) return e;
    if (int e = init_once()) return e;
    if (n <= 0 || n > 0x7fffffff / 4) return fail(NDP_E_INVALID, "bad point count");
    if (!params || !pack || !x || !saved || !grad_y || !grad_params || !workspace) return fail(NDP_E_INVALID, "NULL buffer");
    if (!aligned16(params) || !aligned16(pack) || !aligned16(saved) || !aligned16(workspace))
        return fail(NDP_E_INVALID, "params, saved and workspace must be 16-byte aligned");
    NdpLayout L = layout_of(c);
    NdpBwdArgs b;
    b.lay = L; b.params = params; b.params_stride = 0; b.pack = pack; b.pack_stride = 0; b.x = x; b.x_stride = 0;
    b.act = saved; b.act_stride = 0; b.act_layer_stride = n * NDP_W;
    b.zsave = saved + act_floats(c->depth, n, g_mlp_mode); b.z_stride = 0;
    b.gy = grad_y; b.gy_stride = 0; b.gacc = nullptr; b.gacc_stride = 0; b.m = 1; b.mcounts = nullptr;
    b.gnu = grad_nu; b.gnu_stride = 0;
    b.partials = (float*)workspace; b.partials_stride = 0; b.partial_pitch = (int)pad4(L.param_count);
    b.gx = grad_x; b.gx_stride = 0; b.n = (int)n; b.counts = nullptr; b.state = nullptr; b.npairs = 1;
    b.hgbuf = (float*)workspace + ((n + NDP_TP - 1) / NDP_TP) * (long long)b.partial_pitch; b.hgbuf_stride = 0;
    b.tpc = g_layer_tpc;
    if (g_mlp_mode == 0) ndp_launch_bwd_tc(b, (cudaStream_t)stream); else ndp_launch_bwd(b, (cudaStream_t)stream);
    NdpAdamArgs r;
    r.lay = L; r.params = nullptr; r.params_stride = 0; r.pack = nullptr; r.pack_stride = 0;
    r.m = nullptr; r.v = nullptr; r.mv_stride = 0;
    r.partials = (const float*)workspace; r.partials_stride = 0; r.partial_pitch = b.partial_pitch;
    r.n = (int)n; r.counts = nullptr; r.grads_out = grad_params; r.grads_stride = 0; r.state = nullptr;
    r.fixed_step = 0; r.lr = r.beta1 = r.beta2 = r.eps = 0.0; r.do_adam = 0; r.npairs = 1;
    r.tiles_per_row = g_mlp_mode == 0 ? ndp_bwd_tc_tiles_per_cta(L.hidden, (int)n, g_layer_tpc) : 1;
    ndp_launch_adam(r, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return NDP_OK;
}

extern "C" int ndp_chamfer(const float* x, int64_t n, const float* y, int64_t m, float trunc, float grad_scale,
                           float* loss, float* grad_x, float* d2_x, int64_t* idx_x, float* d2_y, int64_t* idx_y,
                           void* workspace, void* stream) {
    if (n < 1 || m < 1 || n > (1 << 27) || m > (1 << 27)) return fail(NDP_E_INVALID, "clouds must hold 1..2^27 points");
    if (!x || !y || !loss || !grad_x || !workspace) return fail(NDP_E_INVALID, "NULL buffer");
    if ((d2_x == nullptr) != (idx_x == nullptr) || (d2_y == nullptr) != (idx_y == nullptr))
        return fail(NDP_E_INVALID, "d2/idx outputs must be given in pairs");
    if (!aligned16(workspace)) return fail(NDP_E_INVALID, "workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    NnPlan p = nn_plan(n, m);
    ChamferWs w = chamfer_ws(n, m);
    char* ws = (char*)workspace;
    CK(cudaMemsetAsync(ws + w.gacc, 0, (size_t)(w.total - w.gacc), st));   // accumulators, sums, counter
    NdpChamferArgs a;
    a.nn.x = x; a.nn.x_stride = 0; a.nn.n = (int)n; a.nn.ncounts = nullptr;
    a.nn.y = y; a.nn.y_stride = 0; a.nn.m = (int)m; a.nn.mcounts = nullptr;
    a.nn.part = (float2*)(ws + w.part); a.nn.part_pair_stride = 0;
    a.nn.qpitch = p.qpitch; a.nn.chunks = p.chunks; a.nn.chunk_targets = p.chunk_targets;
    a.nn.state = nullptr; a.nn.npairs = 1;
    a.trunc = trunc; a.gx = grad_x; a.gx_stride = 0;
    a.gacc = (unsigned long long*)(ws + w.gacc); a.gacc_stride = 0;
    a.d2x = d2_x; a.idxx = (long long*)idx_x; a.nx_stride = 0;
    a.d2y = d2_y; a.idxy = (long long*)idx_y; a.ny_stride = 0;
    a.blocksums = (double*)(ws + w.sums); a.blocks_pitch = p.blocks;
    a.counters = (int*)(ws + w.counter); a.loss_out = loss; a.state = nullptr;
    a.loss_hist = nullptr; a.hist_stride = 0; a.hist_cap = 0; a.max_break_count = 0; a.break_ratio = 0.0;
    ndp_launch_nn(a.nn, st);
    ndp_launch_chamfer_reduce(a, st);
    NdpGradFinalizeArgs f;
    f.gx = grad_x; f.gx_stride = 0; f.gacc = a.gacc; f.gacc_stride = 0;
    f.n = (int)n; f.ncounts = nullptr; f.m = (int)m; f.mcounts = nullptr; f.scale = grad_scale; f.npairs = 1;
    ndp_launch_grad_finalize(f, st);
    CK(cudaGetLastError());
    return NDP_OK;
}

extern "C" int ndp_adam_step(const ndp_layer_cfg* c, float* params, const float* grads, float* exp_avg,
                             float* exp_avg_sq, int64_t count, int32_t step, double lr, double beta1, double beta2,
                             double eps, float* pack, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq) return fail(NDP_E_INVALID, "NULL buffer");
    if (step < 1) return fail(NDP_E_INVALID, "step is 1-based");
    NdpAdamArgs r;
    if (pack) {
        if (int e = check_cfg(c)) return e;
        r.lay = layout_of(c);
        if (count != r.lay.param_count) return fail(NDP_E_INVALID, "count does not match cfg");
    } else {
        r.lay = ndp_make_layout(1, NDP_MOTION_SFLOW, 0, 0, 1.0f, 1.0f);
        r.lay.hidden = 0; r.lay.off_b_in = 0;
        if (count < 0 || count > 0x7fffffff) return fail(NDP_E_INVALID, "bad count");
        r.lay.param_count = (int)count;
    }
    r.params = params; r.params_stride = 0; r.pack = pack; r.pack_stride = 0;
    r.m = exp_avg; r.v = exp_avg_sq; r.mv_stride = 0;
    r.partials = grads; r.partials_stride = 0; r.partial_pitch = 0; r.n = 0; r.counts = nullptr;
    r.grads_out = nullptr; r.grads_stride = 0; r.state = nullptr; r.fixed_step = step;
    r.lr = lr; r.beta1 = beta1; r.beta2 = beta2; r.eps = eps; r.do_adam = 1; r.npairs = 1;
    ndp_launch_adam(r, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return NDP_OK;
}

// =================================================================================================
// Fused per-pair driver
// =================================================================================================
#define NDP_MAX_STREAMS 8
// The solver's buffers, streams and events belong to the device that was current at creation: make it
// current for the duration of a call (and restore the caller's), whatever the calling thread had selected.
struct DeviceGuard {
    int prev = -1; bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct ndp_solver {
    ndp_solver_cfg cfg;
    std::vector<NdpLayout> lay;
    int P, Ppad, packn, tiles, B, S, NS, NT;
    NnPlan plan;
    // device
    float *src_raw = nullptr, *tgt_raw = nullptr, *src_c = nullptr, *wbuf[2] = {nullptr, nullptr};
    float *means = nullptr, *smp[2] = {nullptr, nullptr}, *tsmp = nullptr;
    int *perm_s = nullptr, *perm_t = nullptr, *ncount = nullptr, *mcount = nullptr, *nscount = nullptr, *ntcount = nullptr;
    float *params = nullptr, *pack = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    float *act = nullptr, *zsave = nullptr, *gx = nullptr, *partials = nullptr, *hgbuf = nullptr, *loss = nullptr, *loss_hist = nullptr;
    unsigned long long* gacc = nullptr;
    float2* nnpart = nullptr;
    // culled NN search (nn_mode 0): Morton order, float4 copies, block boxes, previous-NN seeds
    unsigned long long* keys = nullptr;
    float *sraw = nullptr, *traw = nullptr, *bounds = nullptr, *xbox = nullptr, *tbox = nullptr;
    float4 *x4 = nullptr, *t4 = nullptr;
    int *orig_s = nullptr, *orig_t = nullptr, *inv_s = nullptr, *inv_t = nullptr, *prev_x = nullptr, *prev_y = nullptr;
    int npad = 0, S128 = 0, nboxes = 0, mlp_mode = 0;
    int device = 0;                     // the CUDA device the solver's buffers, streams and events live on
    int tpc = 0, fwd_rounds = 0;        // work grouping of the tensor-core kernels (0 = automatic)
    unsigned long long* nnstats = nullptr;   // [4] device counters of the culled search (profile_every > 0)
    int last_npairs = 0, last_cur = 0;  // the last register call: pairs, and which sample buffer holds the last warped samples
    long long act_pair = 0;
    double* blocksums = nullptr;
    int* counters = nullptr;
    NdpPairState* state = nullptr;
    // host (pinned)
    NdpPairState* h_state = nullptr;
    // early-stop polling that does not drain the pipeline: a snapshot of the pair states is copied on its own stream behind
    // events of the stream groups and looked at one poll interval later (the device-side stop is authoritative: kernels of
    // stopped pairs return at once, the host merely stops launching)
    NdpPairState* h_poll = nullptr;
    cudaStream_t st_poll = nullptr;
    cudaEvent_t ev_poll_g[NDP_MAX_STREAMS] = {}, ev_poll_done = nullptr;
    int* h_counts = nullptr;
    long long launches = 0;
    std::vector<void*> allocs;
    // sampled per-kernel timing (profile_every > 0): 6 events bracket the 5 kernels of an iteration
    std::vector<cudaEvent_t> events;
    int ev_used = 0;
    double prof_ms[5] = {0, 0, 0, 0, 0};
    long long prof_samples = 0;
    int prof_pairs = 0;                 // pairs per profiled launch
    // the batch is split into two halves that run on two streams, so that one half's small kernels
    // (NN search, Chamfer epilogue, Adam) fill the SM time the other half's tensor-core CTAs leave idle
    int dbg_nn = 0;                     // NDP_DEBUG_NN, see NdpPrunedArgs::dbg
    int dbg_skip = 0;                   // NDP_DEBUG_SKIP (read once at creation; measurement aid only, results are garbage): bit 0 forward, 1 NN search, 2 backward, 3 Adam
    int nstreams = 1;                   // stream groups (1..NDP_MAX_STREAMS), env NDP_SOLVER_STREAMS, default 4
    cudaStream_t st_extra[NDP_MAX_STREAMS - 1] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[NDP_MAX_STREAMS - 1] = {};
};

static int prof_flush(ndp_solver* s) {   // call after a stream synchronisation
    for (int i = 0; i + 6 <= s->ev_used; i += 6) {
        for (int k = 0; k < 5; ++k) {
            float ms = 0.0f;
            CK(cudaEventElapsedTime(&ms, s->events[i + k], s->events[i + k + 1]));
            s->prof_ms[k] += ms;
        }
        s->prof_samples += 1;
    }
    s->ev_used = 0;
    return NDP_OK;
}

template <class T> static int dalloc(ndp_solver* s, T** p, long long count) {
    void* q = nullptr;
    if (cudaMalloc(&q, (size_t)(count > 0 ? count : 1) * sizeof(T)) != cudaSuccess) return fail(NDP_E_NOMEM, "cudaMalloc failed");
    s->allocs.push_back(q);
    *p = (T*)q;
    return NDP_OK;
}

extern "C" void ndp_solver_destroy(ndp_solver* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    for (cudaEvent_t e : s->events) cudaEventDestroy(e);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    for (int i = 0; i < NDP_MAX_STREAMS - 1; ++i) {
        if (s->ev_join[i]) cudaEventDestroy(s->ev_join[i]);
        if (s->st_extra[i]) cudaStreamDestroy(s->st_extra[i]);
    }
    if (s->h_state) cudaFreeHost(s->h_state);
    if (s->h_poll) cudaFreeHost(s->h_poll);
    if (s->st_poll) cudaStreamDestroy(s->st_poll);
    if (s->ev_poll_done) cudaEventDestroy(s->ev_poll_done);
    for (int i = 0; i < NDP_MAX_STREAMS; ++i)
        if (s->ev_poll_g[i]) cudaEventDestroy(s->ev_poll_g[i]);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    delete s;
}

extern "C" int ndp_solver_create(const ndp_solver_cfg* c, ndp_solver** out) {
    if (!c || !out) return fail(NDP_E_INVALID, "NULL argument");
    if (int e = init_once()) return e;
    ndp_layer_cfg lc{c->width, c->depth, c->motion, c->rot_format, 0, 1.0f, 0.001f};
    if (int e = check_cfg(&lc)) return e;
    if (c->max_pairs < 1 || c->samples < 1 || c->levels < 1 || c->iters < 0 || c->max_src_points < 1 || c->max_tgt_points < 1)
        return fail(NDP_E_INVALID, "bad solver sizes");
    ndp_solver* s = new ndp_solver();
    s->cfg = *c;
    for (int l = 0; l < c->levels; ++l)   // Deformation_Pyramid.__init__, nets.py:20-30: level i has m = i+1
        s->lay.push_back(ndp_make_layout(c->depth, c->motion, c->rot_format, 0, ldexpf(1.0f, l + 1 + c->k0), 0.001f));
    s->P = s->lay[0].param_count; s->Ppad = (int)pad4(s->P); s->packn = s->lay[0].pack_count;
    s->B = c->max_pairs; s->S = c->samples; s->NS = c->max_src_points; s->NT = c->max_tgt_points;
    s->tiles = (s->S + NDP_TP - 1) / NDP_TP;
    if (c->mlp_mode < 0 || c->mlp_mode > 2 || c->tiles_per_bwd_cta < 0 || c->tiles_per_bwd_cta > 16 || c->fwd_rounds < 0 ||
        c->fwd_rounds > 8 || c->streams < 0 || c->streams > NDP_MAX_STREAMS) {
        delete s;
        return fail(NDP_E_INVALID, "mlp_mode must be 0..2, tiles_per_bwd_cta 0..16, fwd_rounds 0..8, streams 0..8");
    }
    if (c->nn_mode < 0 || c->nn_mode > 2) { delete s; return fail(NDP_E_INVALID, "nn_mode must be 0 (culled search), 1 (brute force) or 2 (paired samples)"); }
    s->mlp_mode = c->mlp_mode == 0 ? g_mlp_mode : c->mlp_mode - 1;
    s->tpc = c->tiles_per_bwd_cta; s->fwd_rounds = c->fwd_rounds;
    if (cudaGetDevice(&s->device) != cudaSuccess) { delete s; return fail(NDP_E_CUDA, "cudaGetDevice failed"); }
    s->act_pair = act_floats(c->depth, s->S, s->mlp_mode);
    s->plan = nn_plan(s->S, s->S);
    if (c->nn_mode == 0) { s->plan.chunks = 1; s->plan.chunk_targets = 1 << 30; }
    s->S128 = (s->S + NDP_TP - 1) / NDP_TP * NDP_TP; s->nboxes = s->S128 / 32;
    s->npad = 1; while (s->npad < s->S) s->npad <<= 1;
    const long long B = s->B, S = s->S;
    int e = NDP_OK;
#define DA(ptr, count) if (!e) e = dalloc(s, &s->ptr, (count))
    DA(src_raw, B * s->NS * 3); DA(tgt_raw, B * s->NT * 3); DA(src_c, B * s->NS * 3);
    DA(wbuf[0], B * s->NS * 3); DA(wbuf[1], B * s->NS * 3);
    DA(means, B * 6); DA(smp[0], B * S * 3); DA(smp[1], B * S * 3); DA(tsmp, B * S * 3);
    DA(perm_s, B * S); DA(perm_t, B * S); DA(ncount, B); DA(mcount, B); DA(nscount, B); DA(ntcount, B);
    DA(params, B * c->levels * s->Ppad); DA(pack, B * s->packn); DA(adam_m, B * s->Ppad); DA(adam_v, B * s->Ppad);
    DA(act, B * s->act_pair); DA(zsave, B * S * NDP_ZPITCH); DA(gx, B * S * 3); DA(gacc, B * S * 3);
    DA(partials, B * s->tiles * s->Ppad); DA(hgbuf, B * s->tiles * (long long)NDP_HGREC); DA(loss, B);
    DA(nnpart, B * 2 * s->plan.chunks * s->plan.qpitch); DA(blocksums, B * s->plan.blocks * 2); DA(counters, B);
    DA(state, B);
    if (c->nn_mode == 0) {
        DA(keys, B * 2 * s->npad); DA(sraw, B * S * 3); DA(traw, B * S * 3); DA(bounds, B * 12);
        DA(xbox, B * s->nboxes * 8); DA(tbox, B * s->nboxes * 8); DA(x4, B * s->S128); DA(t4, B * s->S128);
        DA(orig_s, B * S); DA(orig_t, B * S); DA(inv_s, B * S); DA(inv_t, B * S); DA(prev_x, B * S); DA(prev_y, B * S);
    }
    if (c->record_loss) DA(loss_hist, B * c->levels * (long long)c->iters);
    if (c->profile_every > 0 && c->nn_mode == 0) DA(nnstats, 4);
#undef DA
    if (!e && cudaMallocHost((void**)&s->h_state, sizeof(NdpPairState) * B) != cudaSuccess) e = fail(NDP_E_NOMEM, "cudaMallocHost failed");
    if (!e && cudaMallocHost((void**)&s->h_poll, sizeof(NdpPairState) * B) != cudaSuccess) e = fail(NDP_E_NOMEM, "cudaMallocHost failed");
    if (!e && (cudaStreamCreateWithFlags(&s->st_poll, cudaStreamNonBlocking) != cudaSuccess ||
               cudaEventCreateWithFlags(&s->ev_poll_done, cudaEventDisableTiming) != cudaSuccess))
        e = fail(NDP_E_CUDA, "cannot create the polling stream");
    for (int i = 0; !e && i < NDP_MAX_STREAMS; ++i)
        if (cudaEventCreateWithFlags(&s->ev_poll_g[i], cudaEventDisableTiming) != cudaSuccess) e = fail(NDP_E_CUDA, "cudaEventCreate failed");
    if (!e && cudaMallocHost((void**)&s->h_counts, sizeof(int) * B * 4) != cudaSuccess) e = fail(NDP_E_NOMEM, "cudaMallocHost failed");
    if (!e && cudaMemset(s->counters, 0, sizeof(int) * B) != cudaSuccess) e = fail(NDP_E_CUDA, "cudaMemset failed");
    if (!e) {
        if (const char* dv = getenv("NDP_DEBUG_SKIP")) s->dbg_skip = atoi(dv);
        if (const char* dv = getenv("NDP_DEBUG_NN")) s->dbg_nn = atoi(dv);
        if (const char* dv = getenv("NDP_DEBUG_PRIO")) { ndp_launch_prio[0] = atoi(dv); const char* c2 = strchr(dv, ':'); ndp_launch_prio[1] = c2 ? atoi(c2 + 1) : 0; }
        const int want = c->streams > 0 ? c->streams : 4;
        s->nstreams = (int)(B < want ? B : want);
        if (s->nstreams > 1 && cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess)
            e = fail(NDP_E_CUDA, "event creation failed");
        for (int i = 0; !e && i < s->nstreams - 1; ++i)
            if (cudaStreamCreateWithFlags(&s->st_extra[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming) != cudaSuccess)
                e = fail(NDP_E_CUDA, "stream / event creation failed");
    }
    if (!e && cudaMemset(s->gacc, 0, 8 * B * S * 3) != cudaSuccess) e = fail(NDP_E_CUDA, "cudaMemset failed");
    if (!e && s->nnstats && cudaMemset(s->nnstats, 0, 32) != cudaSuccess) e = fail(NDP_E_CUDA, "cudaMemset failed");
    if (e) { ndp_solver_destroy(s); return e; }
    *out = s;
    return NDP_OK;
}

extern "C" int64_t ndp_solver_params_per_pair(const ndp_solver* s) { return s ? (int64_t)s->cfg.levels * s->P : -1; }
extern "C" int64_t ndp_solver_launch_count(const ndp_solver* s) { return s ? s->launches : -1; }
extern "C" int32_t ndp_solver_profiled_pairs(const ndp_solver* s) { return s ? s->prof_pairs : -1; }
extern "C" int ndp_solver_nn_stats(const ndp_solver* s, int64_t* pair_evals, int64_t* query_blocks, int64_t* max_blocks) {
    if (!s || !pair_evals || !query_blocks) return fail(NDP_E_INVALID, "NULL argument");
    if (!s->nnstats) return fail(NDP_E_INVALID, "the solver was created without profile_every > 0 (or runs the brute-force search)");
    DeviceGuard guard(s->device);
    unsigned long long h[4] = {0, 0, 0, 0};
    CK(cudaMemcpy(h, s->nnstats, 32, cudaMemcpyDeviceToHost));
    *pair_evals = (int64_t)h[0]; *query_blocks = (int64_t)h[1];
    if (max_blocks) *max_blocks = (int64_t)h[3];
    return NDP_OK;
}
extern "C" int ndp_solver_profile(const ndp_solver* s, double* ms, int64_t* samples) {
    if (!s || !ms || !samples) return fail(NDP_E_INVALID, "NULL argument");
    for (int k = 0; k < 5; ++k) ms[k] = s->prof_ms[k];
    *samples = s->prof_samples;
    return NDP_OK;
}

// Optimise all levels for the npairs pairs whose raw clouds / perms / params are already in the
// solver's device buffers; leaves the warped full clouds in s->wbuf[final] and returns which.
static int solver_run(ndp_solver* s, int npairs, bool have_perm_s, bool have_perm_t, int32_t* iters_out,
                      float* loss_out, cudaStream_t st, int* final_buf) {
    const ndp_solver_cfg& c = s->cfg;
    const long long S = s->S;
    // centring (registration.py:150-153) and sub-sampling (:156-159)
    NdpCenterArgs ce;
    ce.src = s->src_raw; ce.src_stride = (long long)s->NS * 3; ce.ns = s->NS; ce.nscounts = s->nscount;
    ce.tgt = s->tgt_raw; ce.tgt_stride = (long long)s->NT * 3; ce.nt = s->NT; ce.ntcounts = s->ntcount;
    ce.means = s->means; ce.npairs = npairs;
    ndp_launch_means(ce, st);
    NdpGatherArgs g;
    g.means = s->means; g.npairs = npairs;
    g.in = s->src_raw; g.in_stride = (long long)s->NS * 3; g.idx = nullptr; g.idx_stride = 0; g.which = 0;
    g.out = s->src_c; g.out_stride = (long long)s->NS * 3; g.n = s->NS; g.counts = s->nscount;
    ndp_launch_gather_center(g, st);
    const bool culled = c.nn_mode == 0, paired = c.nn_mode == 2;
    if (paired)
        for (int p = 0; p < npairs; ++p)
            if (s->h_counts[p] != s->h_counts[s->B + p]) return fail(NDP_E_INVALID, "paired samples (nn_mode 2): src_samples must equal tgt_samples");
    g.idx = have_perm_s ? s->perm_s : nullptr; g.idx_stride = S;
    g.out = culled ? s->sraw : s->smp[0]; g.out_stride = S * 3; g.n = s->S; g.counts = s->ncount;
    ndp_launch_gather_center(g, st);
    g.in = s->tgt_raw; g.in_stride = (long long)s->NT * 3; g.idx = have_perm_t ? s->perm_t : nullptr; g.which = 1;
    g.out = culled ? s->traw : s->tsmp; g.counts = s->mcount;
    ndp_launch_gather_center(g, st);
    s->launches += 4;
    if (culled) {
        // Morton order of both sampled clouds, kept for every level and iteration (ndp_spatial.cu)
        NdpSortArgs so;
        so.src = s->sraw; so.tgt = s->traw; so.cloud_stride = S * 3; so.n = s->S; so.ncounts = s->ncount;
        so.m = s->S; so.mcounts = s->mcount; so.bounds = s->bounds; so.keys = s->keys; so.npad = s->npad;
        so.src_sorted = s->smp[0]; so.tgt_sorted = s->tsmp; so.src_orig = s->orig_s; so.tgt_orig = s->orig_t; so.src_inv = s->inv_s; so.tgt_inv = s->inv_t;
        so.orig_stride = S; so.tgt4 = s->t4; so.p4_stride = s->S128; so.tgt_box = s->tbox; so.box_stride = s->nboxes;
        so.npairs = npairs;
        s->launches += ndp_launch_sort(so, st);
        CK(cudaMemsetAsync(s->prev_x, 0xff, sizeof(int) * (size_t)npairs * S, st));
        CK(cudaMemsetAsync(s->prev_y, 0xff, sizeof(int) * (size_t)npairs * S, st));
    }

    int cur = 0;
    const int poll = (c.max_break_count > c.iters) ? 64 : 8;
    for (int level = 0; level < c.levels; ++level) {
        const NdpLayout& L = s->lay[level];
        float* lvl_params = s->params + (long long)level * s->Ppad;
        const long long pstride = (long long)c.levels * s->Ppad;
        ndp_launch_state_reset(s->state, npairs, st);
        CK(cudaMemsetAsync(s->adam_m, 0, sizeof(float) * (size_t)npairs * s->Ppad, st));
        CK(cudaMemsetAsync(s->adam_v, 0, sizeof(float) * (size_t)npairs * s->Ppad, st));
        // a pair that stops early has scattered its last Chamfer gradient without a backward pass
        // consuming it: clear the fixed-point accumulators before every level
        CK(cudaMemsetAsync(s->gacc, 0, sizeof(unsigned long long) * (size_t)npairs * S * 3, st));
        NdpPackArgs pk; pk.lay = L; pk.params = lvl_params; pk.params_stride = pstride; pk.pack = s->pack; pk.pack_stride = s->packn; pk.npairs = npairs;
        ndp_launch_pack(pk, st);
        s->launches += 2;

        NdpFwdArgs f;
        f.lay = L; f.params = lvl_params; f.params_stride = pstride; f.pack = s->pack; f.pack_stride = s->packn;
        f.x = s->smp[cur]; f.x_stride = S * 3; f.y = s->smp[cur ^ 1]; f.y_stride = S * 3; f.nu = nullptr; f.nu_stride = 0;
        f.act = s->act; f.act_stride = s->act_pair; f.act_layer_stride = S * NDP_W;
        f.zsave = s->zsave; f.z_stride = S * NDP_ZPITCH; f.y_add = nullptr; f.y_add_stride = 0;
        f.y4 = culled ? s->x4 : nullptr; f.y4_stride = s->S128; f.orig = s->orig_s; f.orig_stride = S;
        f.ybox = culled ? s->xbox : nullptr; f.box_stride = s->nboxes;
        f.n = s->S; f.counts = s->ncount; f.state = s->state; f.npairs = npairs;
        f.rounds = s->fwd_rounds;
        if (s->act_pair == 0) { f.act = nullptr; f.act_stride = 0; }     // nothing is saved: the backward kernel recomputes

        NdpChamferArgs ch;
        ch.nn.x = s->smp[cur ^ 1]; ch.nn.x_stride = S * 3; ch.nn.n = s->S; ch.nn.ncounts = s->ncount;
        ch.nn.y = s->tsmp; ch.nn.y_stride = S * 3; ch.nn.m = s->S; ch.nn.mcounts = s->mcount;
        ch.nn.part = s->nnpart; ch.nn.part_pair_stride = 2LL * s->plan.chunks * s->plan.qpitch;
        ch.nn.qpitch = s->plan.qpitch; ch.nn.chunks = s->plan.chunks; ch.nn.chunk_targets = s->plan.chunk_targets;
        ch.nn.state = s->state; ch.nn.npairs = npairs;
        NdpPrunedArgs pn;
        pn.x4 = s->x4; pn.y4 = s->t4; pn.p4_stride = s->S128; pn.xbox = s->xbox; pn.ybox = s->tbox; pn.box_stride = s->nboxes;
        pn.prev_x = s->prev_x; pn.prev_y = s->prev_y; pn.prev_stride = S; pn.inv_x = s->inv_s; pn.inv_y = s->inv_t; pn.inv_stride = S; pn.n = s->S; pn.ncounts = s->ncount;
        pn.m = s->S; pn.mcounts = s->mcount; pn.part = s->nnpart; pn.part_pair_stride = ch.nn.part_pair_stride;
        pn.qpitch = s->plan.qpitch; pn.state = s->state; pn.npairs = npairs; pn.stats = s->nnstats; pn.dbg = s->dbg_nn;
        ch.trunc = c.trunc; ch.gx = s->gx; ch.gx_stride = S * 3; ch.gacc = s->gacc; ch.gacc_stride = S * 3;
        ch.d2x = nullptr; ch.idxx = nullptr; ch.nx_stride = 0; ch.d2y = nullptr; ch.idxy = nullptr; ch.ny_stride = 0;
        ch.blocksums = s->blocksums; ch.blocks_pitch = s->plan.blocks; ch.counters = s->counters; ch.loss_out = s->loss;
        ch.state = s->state;
        ch.loss_hist = s->loss_hist ? s->loss_hist + (long long)level * c.iters : nullptr;
        ch.hist_stride = (long long)c.levels * c.iters; ch.hist_cap = c.iters;
        ch.max_break_count = c.max_break_count; ch.break_ratio = (double)c.break_threshold_ratio;
        ch.paired = paired ? 1 : 0;

        NdpBwdArgs b;
        b.lay = L; b.params = lvl_params; b.params_stride = pstride; b.pack = s->pack; b.pack_stride = s->packn;
        b.x = s->smp[cur]; b.x_stride = S * 3;
        b.act = s->act; b.act_stride = f.act_stride; b.act_layer_stride = f.act_layer_stride;
        b.zsave = s->zsave; b.z_stride = f.z_stride; b.gy = s->gx; b.gy_stride = S * 3;
        b.gacc = s->gacc; b.gacc_stride = S * 3; b.m = s->S; b.mcounts = s->mcount; b.gnu = nullptr; b.gnu_stride = 0;
        b.partials = s->partials; b.partials_stride = (long long)s->tiles * s->Ppad; b.partial_pitch = s->Ppad;
        b.gx = nullptr; b.gx_stride = 0; b.n = s->S; b.counts = s->ncount; b.state = s->state; b.npairs = npairs;
        b.hgbuf = s->hgbuf; b.hgbuf_stride = (long long)s->tiles * NDP_HGREC;
        b.tpc = s->tpc;

        NdpAdamArgs ad;
        ad.lay = L; ad.params = lvl_params; ad.params_stride = pstride; ad.pack = s->pack; ad.pack_stride = s->packn;
        ad.m = s->adam_m; ad.v = s->adam_v; ad.mv_stride = s->Ppad;
        ad.partials = s->partials; ad.partials_stride = b.partials_stride; ad.partial_pitch = s->Ppad;
        ad.n = s->S; ad.counts = s->ncount; ad.grads_out = nullptr; ad.grads_stride = 0; ad.state = s->state;
        ad.fixed_step = 0; ad.lr = c.lr; ad.beta1 = 0.9; ad.beta2 = 0.999; ad.eps = 1e-8; ad.do_adam = 1; ad.npairs = npairs;
        ad.tiles_per_row = s->mlp_mode == 0 ? ndp_bwd_tc_tiles_per_cta(L.hidden, s->S, s->tpc) : 1;
        ad.pack_fp32 = s->mlp_mode == 0 ? 0 : 1;

        // the batch is split into stream groups (contiguous pair ranges, sizes differ by at most one)
        const int ng = npairs < s->nstreams ? npairs : s->nstreams;
        int gfirst[NDP_MAX_STREAMS], gcount[NDP_MAX_STREAMS];
        cudaStream_t gs[NDP_MAX_STREAMS];
        for (int g = 0, at = 0; g < ng; ++g) {
            gcount[g] = npairs / ng + (g < npairs % ng ? 1 : 0);
            gfirst[g] = at; at += gcount[g];
            gs[g] = g == 0 ? st : s->st_extra[g - 1];
        }
        s->prof_pairs = gcount[0];
        if (ng > 1) {           // the level's set-up (state reset, moments, pack) precedes every group
            CK(cudaEventRecord(s->ev_fork, st));
            for (int g = 1; g < ng; ++g) CK(cudaStreamWaitEvent(gs[g], s->ev_fork, 0));
        }
        auto join = [&]() -> int {
            for (int g = 1; g < ng; ++g) {
                CK(cudaEventRecord(s->ev_join[g - 1], gs[g]));
                CK(cudaStreamWaitEvent(st, s->ev_join[g - 1], 0));
            }
            return NDP_OK;
        };
        bool poll_pending = false;
        for (int it = 0; it < c.iters; ++it) {
            const bool prof = c.profile_every > 0 && (it % c.profile_every) == c.profile_every / 2;
            cudaEvent_t* ev = nullptr;
            if (prof) {
                while ((int)s->events.size() < s->ev_used + 6) {
                    cudaEvent_t e;
                    CK(cudaEventCreate(&e));
                    s->events.push_back(e);
                }
                ev = s->events.data() + s->ev_used;
                s->ev_used += 6;
            }
            for (int g = 0; g < ng; ++g) {
                cudaStream_t q = gs[g];
                const bool pg = prof && g == 0;               // sampled timing: the first half's launches, on their stream
                f.pair0 = pn.pair0 = ch.nn.pair0 = b.pair0 = ad.pair0 = gfirst[g];
                f.npairs = pn.npairs = ch.nn.npairs = b.npairs = ad.npairs = gcount[g];
                if (pg) CK(cudaEventRecord(ev[0], q));
                if (s->dbg_skip & 1) {} else
                if (s->mlp_mode == 0) ndp_launch_fwd_tc(f, q); else ndp_launch_fwd(f, q);
                if (pg) CK(cudaEventRecord(ev[1], q));
                if (s->dbg_skip & 2) {} else
                if (culled) { pn.fuse = &ch; ndp_launch_nn_pruned(pn, q); } else if (!paired) ndp_launch_nn(ch.nn, q);   // culled: search + Chamfer epilogue in one launch
                if (pg) CK(cudaEventRecord(ev[2], q));
                if (!culled) ndp_launch_chamfer_reduce(ch, q);
                if (pg) CK(cudaEventRecord(ev[3], q));
                if (s->dbg_skip & 4) {} else
                if (s->mlp_mode == 0) ndp_launch_bwd_tc(b, q); else ndp_launch_bwd(b, q);
                if (pg) CK(cudaEventRecord(ev[4], q));
                if (!(s->dbg_skip & 8)) ndp_launch_adam(ad, q);
                if (pg) CK(cudaEventRecord(ev[5], q));
                s->launches += ((s->mlp_mode == 0) ? 6 : 5) - ((culled || paired) ? 1 : 0);
            }
            if ((it + 1) % poll == 0 && it + 1 < c.iters) {
                if (poll_pending) {         // the snapshot taken one interval ago (complete by now in all but pathological cases)
                    CK(cudaEventSynchronize(s->ev_poll_done));
                    bool all = true;
                    for (int p = 0; p < npairs; ++p) all = all && s->h_poll[p].stopped;
                    if (all) break;
                }
                for (int g = 0; g < ng; ++g) {
                    CK(cudaEventRecord(s->ev_poll_g[g], gs[g]));
                    CK(cudaStreamWaitEvent(s->st_poll, s->ev_poll_g[g], 0));
                }
                CK(cudaMemcpyAsync(s->h_poll, s->state, sizeof(NdpPairState) * npairs, cudaMemcpyDeviceToHost, s->st_poll));
                CK(cudaEventRecord(s->ev_poll_done, s->st_poll));
                poll_pending = true;
            }
        }
        if (poll_pending) CK(cudaStreamSynchronize(s->st_poll));
        if (int e = join()) return e;
        CK(cudaMemcpyAsync(s->h_state, s->state, sizeof(NdpPairState) * npairs, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (int e = prof_flush(s)) return e;
        for (int p = 0; p < npairs; ++p) {
            if (iters_out) iters_out[(long long)p * c.levels + level] = s->h_state[p].evals - (s->h_state[p].stopped ? 1 : 0);
            if (loss_out) loss_out[(long long)p * c.levels + level] = s->h_state[p].last_loss;
        }
        cur ^= 1;   // the level's output feeds the next level (registration.py:249)
    }
    s->last_npairs = npairs; s->last_cur = cur;

    // final warp of the full, centred source cloud through every level (registration.py:254-259)
    const float* xin = s->src_c;
    int wb = 0;
    for (int level = 0; level < c.levels; ++level) {
        NdpPackArgs pk; pk.lay = s->lay[level]; pk.params = s->params + (long long)level * s->Ppad;
        pk.params_stride = (long long)c.levels * s->Ppad; pk.pack = s->pack; pk.pack_stride = s->packn; pk.npairs = npairs;
        ndp_launch_pack(pk, st);
        NdpFwdArgs f;
        f.lay = s->lay[level]; f.params = pk.params; f.params_stride = pk.params_stride; f.pack = s->pack; f.pack_stride = s->packn;
        f.x = xin; f.x_stride = (long long)s->NS * 3; f.y = s->wbuf[wb]; f.y_stride = (long long)s->NS * 3;
        f.nu = nullptr; f.nu_stride = 0; f.act = nullptr; f.act_stride = 0; f.act_layer_stride = 0; f.zsave = nullptr; f.z_stride = 0;
        const bool last = level == c.levels - 1;
        f.y_add = last ? s->means + 3 : nullptr; f.y_add_stride = 6;
        f.y4 = nullptr; f.y4_stride = 0; f.orig = nullptr; f.orig_stride = 0; f.ybox = nullptr; f.box_stride = 0;
        f.n = s->NS; f.counts = s->nscount; f.state = nullptr; f.npairs = npairs;
        f.rounds = s->fwd_rounds;
        if (s->mlp_mode == 0) ndp_launch_fwd_tc(f, st); else ndp_launch_fwd(f, st);
        s->launches += 2;
        xin = s->wbuf[wb];
        *final_buf = wb;
        wb ^= 1;
    }
    CK(cudaGetLastError());
    return NDP_OK;
}

static int solver_counts(ndp_solver* s, int npairs, const int32_t* ns, const int32_t* nt, const int32_t* nss,
                         const int32_t* nts, cudaStream_t st) {
    if (npairs < 1 || npairs > s->B) return fail(NDP_E_INVALID, "npairs exceeds max_pairs");
    for (int p = 0; p < npairs; ++p) {
        if (ns[p] < 1 || ns[p] > s->NS || nt[p] < 1 || nt[p] > s->NT) return fail(NDP_E_INVALID, "cloud size out of range");
        const int ds = ns[p] < s->S ? ns[p] : s->S, dt = nt[p] < s->S ? nt[p] : s->S;       // src[: samples]
        const int cs = nss ? nss[p] : ds, ct = nts ? nts[p] : dt;
        if (cs < 1 || cs > ds || ct < 1 || ct > dt)
            return fail(NDP_E_INVALID, "sample counts must be in [1, min(samples, cloud size)]");
        s->h_counts[p] = cs;
        s->h_counts[s->B + p] = ct;
        s->h_counts[2 * s->B + p] = ns[p];
        s->h_counts[3 * s->B + p] = nt[p];
    }
    CK(cudaMemcpyAsync(s->ncount, s->h_counts, sizeof(int) * npairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->mcount, s->h_counts + s->B, sizeof(int) * npairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->nscount, s->h_counts + 2 * s->B, sizeof(int) * npairs, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->ntcount, s->h_counts + 3 * s->B, sizeof(int) * npairs, cudaMemcpyHostToDevice, st));
    return NDP_OK;
}

static int solver_register(ndp_solver* s, int32_t npairs, const float* const* src, const int32_t* ns,
                           const float* const* tgt, const int32_t* nt, const int32_t* const* src_perm,
                           const int32_t* const* tgt_perm, const int32_t* src_samples, const int32_t* tgt_samples,
                           float* params_host, float* const* params_dev,
                           int params_out, float* const* warped, int32_t* iters_out, float* loss_out,
                           cudaStream_t st, cudaMemcpyKind in_kind, cudaMemcpyKind out_kind) {
    if (!s || !src || !ns || !tgt || !nt || !warped) return fail(NDP_E_INVALID, "NULL argument");
    if (!params_host && !params_dev) return fail(NDP_E_INVALID, "initial weights are required");
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(NDP_E_CUDA, "cannot select the solver's device");
    if (int e = solver_counts(s, npairs, ns, nt, src_samples, tgt_samples, st)) return e;
    const ndp_solver_cfg& c = s->cfg;
    bool hps = src_perm != nullptr, hpt = tgt_perm != nullptr;
    for (int p = 0; p < npairs; ++p) {
        if (hps != (src_perm && src_perm[p]) || hpt != (tgt_perm && tgt_perm[p]))
            return fail(NDP_E_INVALID, "permutations must be given for all pairs or none");
    }
    for (int p = 0; p < npairs; ++p) {
        CK(cudaMemcpyAsync(s->src_raw + (long long)p * s->NS * 3, src[p], sizeof(float) * 3 * ns[p], in_kind, st));
        CK(cudaMemcpyAsync(s->tgt_raw + (long long)p * s->NT * 3, tgt[p], sizeof(float) * 3 * nt[p], in_kind, st));
        if (hps) CK(cudaMemcpyAsync(s->perm_s + (long long)p * s->S, src_perm[p], sizeof(int) * s->h_counts[p], in_kind, st));
        if (hpt) CK(cudaMemcpyAsync(s->perm_t + (long long)p * s->S, tgt_perm[p], sizeof(int) * s->h_counts[s->B + p], in_kind, st));
        for (int l = 0; l < c.levels; ++l) {
            const float* from = params_host ? params_host + ((long long)p * c.levels + l) * s->P
                                            : params_dev[p] + (long long)l * s->P;
            CK(cudaMemcpyAsync(s->params + ((long long)p * c.levels + l) * s->Ppad, from, sizeof(float) * s->P, in_kind, st));
        }
    }
    int fb = 0;
    if (int e = solver_run(s, npairs, hps, hpt, iters_out, loss_out, st, &fb)) return e;
    for (int p = 0; p < npairs; ++p) {
        CK(cudaMemcpyAsync(warped[p], s->wbuf[fb] + (long long)p * s->NS * 3, sizeof(float) * 3 * ns[p], out_kind, st));
        if (params_out || params_dev) {
            for (int l = 0; l < c.levels; ++l) {
                float* to = params_host ? params_host + ((long long)p * c.levels + l) * s->P
                                        : params_dev[p] + (long long)l * s->P;
                CK(cudaMemcpyAsync(to, s->params + ((long long)p * c.levels + l) * s->Ppad, sizeof(float) * s->P, out_kind, st));
            }
        }
    }
    CK(cudaStreamSynchronize(st));
    return NDP_OK;
}

extern "C" int ndp_solver_register_host(ndp_solver* s, int32_t npairs, const float* const* src, const int32_t* ns,
                                        const float* const* tgt, const int32_t* nt, const int32_t* const* src_perm,
                                        const int32_t* const* tgt_perm, const int32_t* src_samples,
                                        const int32_t* tgt_samples, float* params, int32_t params_out,
                                        float* const* warped, int32_t* iters_out, float* loss_out, void* stream) {
    return solver_register(s, npairs, src, ns, tgt, nt, src_perm, tgt_perm, src_samples, tgt_samples, params, nullptr,
                           params_out, warped, iters_out, loss_out, (cudaStream_t)stream, cudaMemcpyHostToDevice,
                           cudaMemcpyDeviceToHost);
}

extern "C" int ndp_solver_register_device(ndp_solver* s, int32_t npairs, const float* const* src, const int32_t* ns,
                                          const float* const* tgt, const int32_t* nt, const int32_t* const* src_perm,
                                          const int32_t* const* tgt_perm, const int32_t* src_samples,
                                          const int32_t* tgt_samples, float* const* params, float* const* warped,
                                          int32_t* iters_out, float* loss_out, void* stream) {
    return solver_register(s, npairs, src, ns, tgt, nt, src_perm, tgt_perm, src_samples, tgt_samples, nullptr, params, 1,
                           warped, iters_out, loss_out, (cudaStream_t)stream, cudaMemcpyDeviceToDevice,
                           cudaMemcpyDeviceToDevice);
}

// Nearest neighbours of the LAST loss evaluation of the last register call (last level), in the sample
// index space of that call: idx_x[i] = index (into the target samples) of the nearest target sample of
// warped source sample i, d2_x[i] its squared distance; idx_y / d2_y the converse; warped_samples /
// target_samples = the two clouds the search ran on.  Host buffers of `samples` entries (x 3 floats for a cloud).
extern "C" int ndp_solver_last_nn(ndp_solver* s, int32_t pair, int64_t* idx_x, float* d2_x, int64_t* idx_y, float* d2_y,
                                  float* warped_samples, float* target_samples, void* stream) {
    if (!s || pair < 0 || pair >= s->last_npairs) return fail(NDP_E_INVALID, "no such pair in the last register call");
    if ((idx_x == nullptr) != (d2_x == nullptr) || (idx_y == nullptr) != (d2_y == nullptr))
        return fail(NDP_E_INVALID, "d2/idx outputs must be given in pairs");
    DeviceGuard guard(s->device);
    if (!guard.ok) return fail(NDP_E_CUDA, "cannot select the solver's device");
    cudaStream_t st = (cudaStream_t)stream;
    const long long S = s->S;
    const int n = s->h_counts[pair], m = s->h_counts[s->B + pair];
    if (s->cfg.nn_mode == 2) return fail(NDP_E_INVALID, "paired samples (nn_mode 2): there is no nearest-neighbour search to report");
    const bool culled = s->cfg.nn_mode == 0;
    NdpNnExportArgs e;
    e.part = s->nnpart + (long long)pair * 2LL * s->plan.chunks * s->plan.qpitch;
    e.qpitch = s->plan.qpitch; e.chunks = s->plan.chunks; e.chunk_targets = s->plan.chunk_targets;
    e.n = n; e.m = m;
    e.orig_s = culled ? s->orig_s + pair * S : nullptr; e.orig_t = culled ? s->orig_t + pair * S : nullptr;
    e.warped = s->smp[s->last_cur] + pair * S * 3;
    e.target = s->tsmp + pair * S * 3;
    long long* d_idx = nullptr; float* d_d2 = nullptr; float* d_w = nullptr;
    if (cudaMalloc((void**)&d_idx, sizeof(long long) * 2 * S) != cudaSuccess ||
        cudaMalloc((void**)&d_d2, sizeof(float) * 2 * S) != cudaSuccess ||
        cudaMalloc((void**)&d_w, sizeof(float) * 6 * S) != cudaSuccess) {
        cudaFree(d_idx); cudaFree(d_d2); cudaFree(d_w);
        return fail(NDP_E_NOMEM, "cudaMalloc failed");
    }
    e.idx_x = d_idx; e.idx_y = d_idx + S; e.d2_x = d_d2; e.d2_y = d_d2 + S; e.warped_out = d_w; e.target_out = d_w + 3 * S;
    ndp_launch_nn_export(e, st);
    if (idx_x) {
        CK(cudaMemcpyAsync(idx_x, d_idx, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(d2_x, d_d2, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    }
    if (idx_y) {
        CK(cudaMemcpyAsync(idx_y, d_idx + S, sizeof(long long) * m, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(d2_y, d_d2 + S, sizeof(float) * m, cudaMemcpyDeviceToHost, st));
    }
    if (warped_samples) CK(cudaMemcpyAsync(warped_samples, d_w, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    if (target_samples) CK(cudaMemcpyAsync(target_samples, d_w + 3 * S, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, st));
    const cudaError_t se = cudaStreamSynchronize(st);
    cudaFree(d_idx); cudaFree(d_d2); cudaFree(d_w);
    CK(se);
    CK(cudaGetLastError());
    return NDP_OK;
}

extern "C" int ndp_solver_losses(ndp_solver* s, int32_t pair, float* out, void* stream) {
    if (!s || !out || pair < 0 || pair >= s->B) return fail(NDP_E_INVALID, "bad argument");
    if (!s->loss_hist) return fail(NDP_E_INVALID, "solver was created with record_loss = 0");
    const long long n = (long long)s->cfg.levels * s->cfg.iters;
    CK(cudaMemcpyAsync(out, s->loss_hist + pair * n, sizeof(float) * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return NDP_OK;
}

// Debug aid (not part of the public header): globaltimer stamps of CTA (0,0) of the last
// tensor-core forward (which = 0) / backward (which = 1) launch; see NDP_T in the kernels.
size_t ndp_bwd_rc_smem_bytes();
size_t ndp_fwd_tc2_smem_bytes();
// Debug aid (not part of the public header): dynamic shared memory of the tensor-core kernels, for the
// "fits the 227 KB opt-in limit" check that runs without a GPU.
extern "C" long long ndp_debug_smem_bytes(int which) {
    switch (which) {
        case 0: return (long long)ndp_fwd_tc_smem_bytes();
        case 1: return (long long)ndp_fwd_tc2_smem_bytes();
        case 2: return (long long)ndp_bwd_tc_smem_bytes();
        case 3: return (long long)ndp_bwd_rc_smem_bytes();
        case 4: return (long long)ndp_fwd_smem_bytes();
        case 5: return (long long)ndp_bwd_smem_bytes();
    }
    return -1;
}
int ndp_debug_copy_fwd(unsigned long long* out);
int ndp_debug_copy_bwd(unsigned long long* out);
int ndp_debug_copy_rc(unsigned long long* out);
extern "C" int ndp_debug_phase_times(int which, unsigned long long* out) {
    cudaDeviceSynchronize();
    return which == 2 ? ndp_debug_copy_rc(out) : (which ? ndp_debug_copy_bwd(out) : ndp_debug_copy_fwd(out));
}
