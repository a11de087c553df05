// Internal launcher interface between the C-ABI (ndp_cabi.cu / ndp_solver.cu) and the kernels.
// Every kernel is batched over point-cloud PAIRS (blockIdx.y or .z = pair): each pair has its own
// clouds, parameter block, Adam state and early-stop state, addressed by uniform strides.
#pragma once
#include "ndp_common.cuh"

// ---- kernel (1): positional encoding + MLP + heads + SE(3)/Sim(3)/flow warp ---------------------
struct NdpFwdArgs {
    NdpLayout lay;
    const float* params; long long params_stride;   // canonical flat block (heads, biases)
    const float* pack;   long long pack_stride;     // transposed hidden/input weights (TMA source)
    const float* x;      long long x_stride;        // [pair][n][3]
    float* y;            long long y_stride;        // [pair][n][3]
    float* nu;           long long nu_stride;       // [pair][n] or null
    float* act;          long long act_stride;      // saved activations [pair][L+1][n_alloc][128] or null
    long long act_layer_stride;
    float* zsave;        long long z_stride;        // saved head vectors [pair][n_alloc][12] or null
    const float* y_add; int y_add_stride;            // [pair][stride] 3 floats added to the output, or null
    // optional extras for the culled NN search (ndp_spatial.cu): float4 copy of the output carrying
    // the original sample index, and the bounding box of every block of 32 consecutive outputs
    float4* y4;          long long y4_stride;       // [pair][n_pad32] or null
    const int* orig;     long long orig_stride;     // [pair][n]
    float* ybox;         long long box_stride;      // [pair][box_stride boxes][8] or null
    int n; const int* counts;                        // points per pair (counts overrides n when non-null)
    const NdpPairState* state;                       // pairs with state.stopped are skipped, or null
    int npairs;
    int pair0 = 0;                                   // first pair of this launch (the driver splits a batch over stream groups)
    int rounds = 0;                                  // tensor-core version: tile pairs per CTA; 0 = default (1)
    int tc_version = 0;                              // 0 = default (2: operand in TMEM when depth <= 3), 1 = shared-memory operand
};
void ndp_launch_fwd(const NdpFwdArgs& a, cudaStream_t s);      // FP32-pipe version (act: fp32 [L+1][n][128])
// tensor-core version: act = [tile][L+1] fp16 hi/lo image sets (65536 bytes each), act_stride in floats per pair
void ndp_launch_fwd_tc(const NdpFwdArgs& a, cudaStream_t s);

// ---- kernel (3a): backward of kernel (1) -> per-tile parameter-gradient partials -----------------
struct NdpBwdArgs {
    NdpLayout lay;
    const float* params; long long params_stride;
    const float* pack;   long long pack_stride;       // tensor-core version only (weight image sets)
    const float* x;      long long x_stride;
    const float* act;    long long act_stride; long long act_layer_stride;
    const float* zsave;  long long z_stride;
    const float* gy;     long long gy_stride;        // dL/dy [pair][n][3]
    unsigned long long* gacc; long long gacc_stride; // optional fixed-point addend [pair][n][3]; consumed+zeroed
    int m; const int* mcounts;                        // gacc scale = 2^-40 / m
    const float* gnu;    long long gnu_stride;        // dL/dnu [pair][n] or null
    float* partials;     long long partials_stride; int partial_pitch;   // [pair][tile][pitch]
    float* gx;           long long gx_stride;         // dL/dx [pair][n][3] or null
    float* hgbuf;        long long hgbuf_stride;      // tensor-core version: per-tile head-gradient records [pair][tile][NDP_HGREC]
    int n; const int* counts;
    const NdpPairState* state;
    int npairs;
    int pair0 = 0;
    int tpc = 0;                                       // tensor-core versions: tiles per CTA = tiles per partial row; 0 = by cloud size
};
int ndp_bwd_tc_tiles_per_cta(int hidden, int n, int forced);
bool ndp_tc_recompute(int hidden);                     // the tensor-core backward rebuilds the activations (no saved images needed)
#define NDP_HGREC (NDP_TP * 24)    // floats per tile: hg[128][16], e[128][8] (e[0][7] = tile max |hg|)
void ndp_launch_bwd(const NdpBwdArgs& a, cudaStream_t s);
void ndp_launch_bwd_tc(const NdpBwdArgs& a, cudaStream_t s);   // needs `pack` (weight image sets)

// ---- kernel (3b): fixed-order reduction of the partials + Adam + transposed-copy refresh ---------
struct NdpAdamArgs {
    NdpLayout lay;
    float* params;  long long params_stride;
    float* pack;    long long pack_stride;           // may be null
    float* m; float* v; long long mv_stride;         // Adam moments (null when do_adam == 0)
    const float* partials; long long partials_stride; int partial_pitch;
    int n; const int* counts;                         // partial rows per pair = ceil(ceil(n / NDP_TP) / tiles_per_row); n == 0 -> 1 row
    int tiles_per_row = 1;
    float* grads_out; long long grads_stride;         // reduced gradient, or null
    const NdpPairState* state;                        // step = state.evals; stopped pairs skipped
    int fixed_step;                                   // used when state == null
    double lr, beta1, beta2, eps;                     // Python floats in torch.optim.Adam => double here
    int do_adam;
    int npairs;
    int pair0 = 0;
    int pack_fp32 = 1;                                // 0: refresh only the fp16 hi/lo images (the tensor-core kernels read nothing else of the pack)
};
void ndp_launch_adam(const NdpAdamArgs& a, cudaStream_t s);

struct NdpPackArgs {
    NdpLayout lay;
    const float* params; long long params_stride;
    float* pack; long long pack_stride;
    int npairs;
};
void ndp_launch_pack(const NdpPackArgs& a, cudaStream_t s);

// ---- kernel (2): brute-force nearest neighbour, both directions, + Chamfer epilogue --------------
#define NDP_NN_THREADS 128
#define NDP_NN_Q 4
#define NDP_NN_QT (NDP_NN_THREADS * NDP_NN_Q)   // queries per CTA
#define NDP_NN_TS 512                            // targets per shared-memory tile
struct NdpNnArgs {
    const float* x; long long x_stride; int n; const int* ncounts;   // warped source [pair][n][3]
    const float* y; long long y_stride; int m; const int* mcounts;   // target        [pair][m][3]
    float2* part; long long part_pair_stride;   // [pair][dir][chunk][qpitch] (d2, idx bits)
    int qpitch; int chunks; int chunk_targets;  // chunk_targets is a multiple of NDP_NN_TS
    const NdpPairState* state;
    int npairs;
    int pair0 = 0;
};
void ndp_launch_nn(const NdpNnArgs& a, cudaStream_t s);

struct NdpChamferArgs {
    NdpNnArgs nn;
    float trunc;
    float* gx; long long gx_stride;                   // direct term of dL/dx [pair][n][3]
    unsigned long long* gacc; long long gacc_stride;  // scattered term, fixed point [pair][n][3]
    float* d2x; long long* idxx; long long nx_stride; // optional NN outputs (raw squared distances)
    float* d2y; long long* idxy; long long ny_stride;
    double* blocksums; int blocks_pitch;              // [pair][blocks_pitch][2]
    int* counters;                                    // [pair], zero before the first launch
    float* loss_out;                                  // [pair]
    NdpPairState* state;                              // early-stop update, or null
    float* loss_hist; long long hist_stride; int hist_cap;   // optional loss curve: loss_hist[pair*stride + eval], eval < cap
    int max_break_count; double break_ratio;
    int paired = 0;                                   // 1: no search -- source sample i is matched to target sample i and the loss is
                                                      // mean_i |x_i - y_i|^2 (the landmark term of LNDP, model/registration.py:200-203)
};
void ndp_launch_chamfer_reduce(const NdpChamferArgs& a, cudaStream_t s);

struct NdpGradFinalizeArgs {                          // gx += fixed-point gacc (standalone Chamfer API)
    float* gx; long long gx_stride;
    unsigned long long* gacc; long long gacc_stride;
    int n; const int* ncounts; int m; const int* mcounts;
    float scale;                                      // upstream dL/dloss
    int npairs;
};
void ndp_launch_grad_finalize(const NdpGradFinalizeArgs& a, cudaStream_t s);

// ---- kernel (2), fused-driver variant: Morton ordering + exact culled NN search (ndp_spatial.cu) ---
struct NdpSortArgs {
    const float* src; const float* tgt; long long cloud_stride;      // centred samples [pair][S][3]
    int n; const int* ncounts; int m; const int* mcounts;
    float* bounds;                                                    // [pair][2][6]
    unsigned long long* keys; int npad;                               // [pair][2][npad], npad = pow2 >= max(n, m)
    float* src_sorted; float* tgt_sorted;                             // [pair][S][3]
    int* src_orig; int* tgt_orig; long long orig_stride;              // sorted position -> sample index
    int* src_inv; int* tgt_inv;                                       // sample index -> sorted position (same stride)
    float4* tgt4; long long p4_stride;                                // [pair][S_pad32]
    float* tgt_box; long long box_stride;                             // [pair][box_stride][8]
    int npairs;
};
int ndp_launch_sort(const NdpSortArgs& a, cudaStream_t s);           // returns the number of launches

struct NdpPrunedArgs {
    const float4* x4; const float4* y4; long long p4_stride;         // warped source / target, sorted, (x,y,z,orig)
    const float* xbox; const float* ybox; long long box_stride;
    int* prev_x; int* prev_y; long long prev_stride;                  // previous NN (sorted index), -1 = none
    const int* inv_x; const int* inv_y; long long inv_stride;         // sample index -> sorted position of the source / target cloud
    int n; const int* ncounts; int m; const int* mcounts;
    float2* part; long long part_pair_stride; int qpitch;             // [pair][dir][qpitch] (d2, sorted idx bits)
    const NdpPairState* state;
    int npairs;
    int pair0 = 0;
    unsigned long long* stats = nullptr;                              // optional [4]: distance evaluations issued (32 lanes x 32 targets per scanned block), 32-query blocks searched, -, most blocks one warp scanned
    int dbg = 0;                                                      // measurement aid (NDP_DEBUG_NN): 1 no candidate walk, 2 no per-query epilogue, 4 no CTA epilogue
    const struct NdpChamferArgs* fuse = nullptr;                      // host pointer, read at launch: when set the kernel also does the whole Chamfer epilogue
};
void ndp_launch_nn_pruned(const NdpPrunedArgs& a, cudaStream_t s);

// Nearest-neighbour results of one pair's last search, exported in the SAMPLE index space of the register call
// (the culled search works on Morton-sorted clouds): ndp_solver_last_nn.
struct NdpNnExportArgs {
    const float2* part; int qpitch; int chunks; int chunk_targets;   // [dir][chunk][qpitch] (d2, index bits)
    int n, m;                                                        // source / target samples
    const int* orig_s; const int* orig_t;                            // sorted position -> sample index (null: already sample order)
    const float* warped; const float* target;                        // [n][3] warped source / [m][3] target samples, in search order
    long long* idx_x; float* d2_x; long long* idx_y; float* d2_y;    // [n] / [m]
    float* warped_out; float* target_out;                            // [n][3] / [m][3] in sample order
};
void ndp_launch_nn_export(const NdpNnExportArgs& a, cudaStream_t s);

// ---- small helpers of the per-pair driver -------------------------------------------------------
struct NdpCenterArgs {                                // registration.py:150-159
    const float* src; long long src_stride; int ns; const int* nscounts;   // full clouds
    const float* tgt; long long tgt_stride; int nt; const int* ntcounts;
    float* means;                                     // [pair][2][3]  (src mean, tgt mean)
    int npairs;
};
void ndp_launch_means(const NdpCenterArgs& a, cudaStream_t s);

struct NdpGatherArgs {                                // out[i] = in[idx[i]] - mean   (idx null: identity)
    const float* in; long long in_stride;
    const int* idx; long long idx_stride;
    const float* means; int which;                    // means[pair][which][3]
    float* out; long long out_stride;
    int n; const int* counts;
    int npairs;
};
void ndp_launch_gather_center(const NdpGatherArgs& a, cudaStream_t s);

void ndp_launch_state_reset(NdpPairState* state, int npairs, cudaStream_t s);

size_t ndp_fwd_smem_bytes();
size_t ndp_bwd_smem_bytes();
int ndp_fwd_init();   // opt-in dynamic shared memory size; returns cudaError_t
int ndp_bwd_init();
int ndp_fwd_tc_init();
int ndp_bwd_tc_init();
int ndp_bwd_rc_init();
size_t ndp_fwd_tc_smem_bytes();
size_t ndp_bwd_tc_smem_bytes();
