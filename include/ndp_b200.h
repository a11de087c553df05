/*
 * ndp_b200.h -- C ABI of libndp_b200.so: the B200 (sm_100a) implementation of the per-pair
 * Neural-Deformation-Pyramid hot path of rabbityl/DeformationPyramid.
 *
 * Plain pointers and sizes only (no torch types).  Unless a function says "host", every buffer
 * is a DEVICE pointer owned by the caller, fp32, contiguous, 16-byte aligned; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous on `stream`
 * unless stated otherwise.  Every function returns 0 on success or a negative NDP_E_* code;
 * ndp_last_error() gives the message of the calling thread's last failure.  NaNs are propagated,
 * not trapped (as in the reference).  There is no CPU fallback.
 *
 * Each entry point names the reference interface (path:line under rabbityl/DeformationPyramid)
 * it replaces; INTEGRATION.md shows the ctypes binding the reference side would add.
 */
#ifndef NDP_B200_H
#define NDP_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NDP_OK 0
#define NDP_E_INVALID -1      /* bad argument (shape, enum, alignment, unsupported width) */
#define NDP_E_CUDA -2         /* CUDA runtime error                                        */
#define NDP_E_NOMEM -3

enum { NDP_SE3 = 0, NDP_SIM3 = 1, NDP_SFLOW = 2 };                         /* nets.py:17        */
enum { NDP_AXIS_ANGLE = 0, NDP_EULER = 1, NDP_QUATERNION = 2, NDP_6D = 3 }; /* nets.py:84-90    */

/* One pyramid level = the constructor arguments of NDPLayer (model/nets.py:67). */
typedef struct ndp_layer_cfg {
    int32_t width;        /* hidden width; this build supports 128 (config/NDP.yaml:26)          */
    int32_t depth;        /* input layer + (depth-1) hidden layers, 1 <= depth <= 9             */
    int32_t motion;       /* NDP_SE3 | NDP_SIM3 | NDP_SFLOW                                      */
    int32_t rot_format;   /* NDP_AXIS_ANGLE | NDP_EULER | NDP_QUATERNION | NDP_6D                */
    int32_t nonrigidity;  /* 1: nr_branch present (nets.py:98-101)                               */
    float freq;           /* 2^(m + k0), nets.py:168                                             */
    float mlp_scale;      /* 0.001, nets.py:107                                                  */
} ndp_layer_cfg;

const char* ndp_last_error(void);
int32_t ndp_version(void);

/* Floats in the flat parameter block of a level, laid out in NDPLayer.parameters() order
 * (nets.py:75-103): input.0.weight[W,6], input.0.bias[W], mlp.pts_linears.l.weight[W,W], .bias[W],
 * rot_brach.weight[R,W], .bias[R], (s_branch), trn_branch.weight[3,W], .bias[3], (nr_branch).  */
int64_t ndp_param_count(const ndp_layer_cfg* cfg);
/* Floats in the kernel-layout block of transposed weight copies ("pack"). */
int64_t ndp_pack_count(const ndp_layer_cfg* cfg);
/* Floats the forward pass saves for the backward pass of n points: the head vectors (12 per point) and,
 * except on the tensor-core path at depth 3 (whose backward kernel REBUILDS the activations from x), the
 * hidden activations. */
int64_t ndp_saved_floats(const ndp_layer_cfg* cfg, int64_t n);
/* Where the layer's contractions run: 0 = tensor cores (tcgen05 / TMEM, every fp32 operand split into
 * an fp16 hi and an fp16 lo term, three partial products, fp32 accumulation: fp32-level accuracy) --
 * the default; 1 = FP32 pipes.  Process-wide; a solver captures the mode at creation. */
int ndp_set_mlp_mode(int32_t mode);
int32_t ndp_get_mlp_mode(void);
/* Work grouping of the tensor-core kernels behind the STANDALONE layer calls below (solvers take theirs
 * from ndp_solver_cfg): tiles (of 128 points) whose gradients one backward CTA accumulates = tiles per
 * partial row of the reduction (0 = a function of n only: 4 at 8192 points), and tile pairs one forward
 * CTA processes (0 = 1).  Both only regroup work; the gradient's summation order follows the first.
 * Process-wide. */
int ndp_set_layer_tuning(int32_t tiles_per_bwd_cta, int32_t fwd_rounds);
/* Bytes of scratch ndp_layer_backward needs for n points. */
int64_t ndp_backward_workspace_bytes(const ndp_layer_cfg* cfg, int64_t n);
/* Bytes of scratch ndp_chamfer needs for clouds of n and m points. */
int64_t ndp_chamfer_workspace_bytes(int64_t n, int64_t m);

/* Refresh the transposed copies after `params` changed (e.g. after an external optimiser step). */
int ndp_pack_params(const ndp_layer_cfg* cfg, const float* params, float* pack, void* stream);

/* Kernel (1).  Replaces NDPLayer.forward (model/nets.py:111-140): x[n,3] -> y[n,3] and, when
 * cfg->nonrigidity, nu[n].  `saved` (ndp_saved_floats(cfg, n) floats) may be NULL when no
 * backward pass follows (inference, registration.py:254-255).                                  */
int ndp_layer_forward(const ndp_layer_cfg* cfg, const float* params, const float* pack,
                      const float* x, int64_t n, float* y, float* nu, float* saved, void* stream);

/* Kernel (3a)+(3b, reduction only).  Replaces the autograd backward of NDPLayer.forward
 * (loss.backward(), model/registration.py:236): given dL/dy[n,3] (and dL/dnu[n] or NULL) writes
 * dL/dparams (flat, parameters() order) and, if grad_x != NULL, dL/dx[n,3].                    */
int ndp_layer_backward(const ndp_layer_cfg* cfg, const float* params, const float* pack, const float* x, int64_t n,
                       const float* saved, const float* grad_y, const float* grad_nu,
                       float* grad_params, float* grad_x, void* workspace, void* stream);

/* Kernel (2).  Replaces compute_truncated_chamfer_distance (model/loss.py:94-258) for one pair,
 * including the two pytorch3d knn_points(K=1) calls it makes (loss.py:177-178).
 *   x[n,3] (differentiable cloud), y[m,3]; trunc compares against SQUARED distances (loss.py:185).
 *   loss[1]; grad_x[n,3] = grad_scale * dloss/dx;  optional NN outputs (NULL to skip):
 *   d2_x[n], idx_x[n] (int64) nearest y of every x; d2_y[m], idx_y[m] nearest x of every y.     */
int ndp_chamfer(const float* x, int64_t n, const float* y, int64_t m, float trunc, float grad_scale,
                float* loss, float* grad_x, float* d2_x, int64_t* idx_x, float* d2_y,
                int64_t* idx_y, void* workspace, void* stream);

/* Kernel (3b).  Replaces torch.optim.Adam.step() as used at model/registration.py:176,237
 * (lr/betas/eps are doubles because torch evaluates them as Python floats; `step` is 1-based).  If `pack` != NULL the
 * transposed copies are refreshed in the same launch (cfg may be NULL when pack is NULL).      */
int ndp_adam_step(const ndp_layer_cfg* cfg, float* params, const float* grads, float* exp_avg,
                  float* exp_avg_sq, int64_t count, int32_t step, double lr, double beta1,
                  double beta2, double eps, float* pack, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused per-pair driver.  Replaces Registration.optimize_deformation_pyramid
 * (model/registration.py:126-262) for the NDP Chamfer objective, batched over independent pairs:
 * centring (:150-153), sub-sampling by caller-supplied permutations (:156-159), level loop
 * (:170), per-level Adam (:176), iteration loop with the early-stop rule evaluated on the device
 * (:184-237), level hand-off (:249), final full-cloud warp through all levels (:254-259).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ndp_solver_cfg {
    int32_t max_pairs;          /* pairs optimised concurrently by one call                       */
    int32_t max_src_points;     /* capacity: full source cloud (<= 30000 in 4DMatch)              */
    int32_t max_tgt_points;
    int32_t samples;            /* config.samples (NDP.yaml:19)                                   */
    int32_t levels;             /* config.m       (NDP.yaml:23)                                   */
    int32_t k0;                 /* config.k0      (NDP.yaml:24)                                   */
    int32_t depth, width;       /* NDP.yaml:25-26                                                 */
    int32_t motion, rot_format; /* NDP.yaml:31-32                                                 */
    int32_t iters;              /* per-level cap  (NDP.yaml:8)                                    */
    int32_t max_break_count;    /* NDP.yaml:10                                                    */
    float break_threshold_ratio;/* NDP.yaml:11                                                    */
    double lr;                  /* NDP.yaml:9 (Python float in torch.optim.Adam => double)        */
    float trunc;                /* 1e9 on the NDP path (registration.py:212)                      */
    int32_t record_loss;        /* 1: keep the per-iteration loss curve (see ndp_solver_losses)   */
    int32_t profile_every;      /* k > 0: bracket the kernels of every k-th iteration with CUDA
                                   events on `stream` (see ndp_solver_profile); 0: off            */
    int32_t nn_mode;            /* 0: exact culled search (Morton blocks + boxes + temporal seeds),
                                   1: plain brute force; identical results.
                                   2: PAIRED samples, no search: source sample i is matched to target
                                   sample i and the loss is mean_i |x'_i - y_i|^2 -- the landmark objective
                                   of LNDP with w_cd = 0 (model/registration.py:200-203, config/LNDP.yaml);
                                   src_samples must equal tgt_samples                             */
    /* ---- execution profile (0 = default everywhere).  These regroup work; none changes the arithmetic
     * of a single product, but tiles_per_bwd_cta sets the summation grouping of the gradients, so results
     * are bit-reproducible for a given value.                                                       */
    int32_t mlp_mode;           /* 0: process default (ndp_set_mlp_mode), 1: tensor cores, 2: FP32 pipes */
    int32_t tiles_per_bwd_cta;  /* 1..16 tiles of 128 samples per backward CTA; 0: by `samples` (4 at 8192) */
    int32_t fwd_rounds;         /* 1..8 tile pairs per forward CTA; 0: 1                             */
    int32_t streams;            /* 1..8 stream groups the batch is split into; 0: 4                  */
} ndp_solver_cfg;

typedef struct ndp_solver ndp_solver;

int ndp_solver_create(const ndp_solver_cfg* cfg, ndp_solver** out);
void ndp_solver_destroy(ndp_solver* s);
/* Floats of one pair's initial weights: sum over levels of ndp_param_count (levels are stored
 * back to back, level 0 first; every level has the same layout on this path).                  */
int64_t ndp_solver_params_per_pair(const ndp_solver* s);

/* Register `npairs` pairs whose clouds and initial weights live in HOST memory (pinned for
 * asynchronous copies).  src[p] -> ns[p] x 3 floats, tgt[p] -> nt[p] x 3; src_perm[p] / tgt_perm[p]
 * hold src_samples[p] / tgt_samples[p] int32 indices (the head of torch.randperm, registration.py:156-159)
 * or NULL for the identity; src_samples / tgt_samples (npairs entries each, or NULL = min(samples, n))
 * are the numbers of points optimised per pair, 1 <= count <= min(samples, n) -- the length of the
 * permutation arrays is explicit, never inferred; params -> npairs x params_per_pair floats (updated in place with the
 * optimised weights when params_out != 0); warped[p] receives ns[p] x 3 floats (the return value
 * of Registration.register()).  iters_out / loss_out (npairs x levels, may be NULL) receive the
 * Adam steps taken and the last loss per level.  Synchronises `stream` before returning.       */
int ndp_solver_register_host(ndp_solver* s, int32_t npairs, const float* const* src,
                             const int32_t* ns, const float* const* tgt, const int32_t* nt,
                             const int32_t* const* src_perm, const int32_t* const* tgt_perm,
                             const int32_t* src_samples, const int32_t* tgt_samples,
                             float* params, int32_t params_out, float* const* warped,
                             int32_t* iters_out, float* loss_out, void* stream);

/* Same with DEVICE buffers (src/tgt/params/warped/perms are host arrays of device pointers).
 * iters_out / loss_out are host arrays; the call synchronises `stream` before returning.       */
int ndp_solver_register_device(ndp_solver* s, int32_t npairs, const float* const* src,
                               const int32_t* ns, const float* const* tgt, const int32_t* nt,
                               const int32_t* const* src_perm, const int32_t* const* tgt_perm,
                               const int32_t* src_samples, const int32_t* tgt_samples,
                               float* const* params, float* const* warped, int32_t* iters_out,
                               float* loss_out, void* stream);

/* Nearest neighbours of the LAST loss evaluation of the last register call (last level) for pair `pair`,
 * i.e. the two pytorch3d knn_points(K=1) results of model/loss.py:177-181 that the reference computes but
 * does not return.  Index space = the sample order of that call (position i = src[src_perm[i]]):
 * idx_x[i] / d2_x[i] = nearest target sample of warped source sample i and its squared distance (n_src
 * samples), idx_y / d2_y the converse (n_tgt samples); warped_samples (n_src x 3) / target_samples (n_tgt x 3)
 * = the two clouds the search ran on (centred, sub-sampled, the source warped by the level's last weights).
 * HOST buffers of `samples` entries; any output (indices and distances in pairs) may be NULL.  Bit-exact with
 * the reference contract (ascending scan, strict '<', fma distance): oracle/knn_oracle.c.  Synchronises.  */
int ndp_solver_last_nn(ndp_solver* s, int32_t pair, int64_t* idx_x, float* d2_x, int64_t* idx_y, float* d2_y,
                       float* warped_samples, float* target_samples, void* stream);

/* Loss curve of the last register call (record_loss = 1): copies levels x iters floats of pair
 * `pair` to the host buffer `out`; entries past the evaluations done are left untouched.       */
int ndp_solver_losses(ndp_solver* s, int32_t pair, float* out, void* stream);
/* Number of kernels launched by this solver since creation (bench.py's gpu_launches). */
int64_t ndp_solver_launch_count(const ndp_solver* s);
/* Sampled device time per kernel since creation (profile_every > 0): ms[5] = accumulated
 * milliseconds of {warp forward, NN search, Chamfer epilogue, warp backward, reduce+Adam} over
 * *samples sampled iterations (each sample is one launch of each kernel over
 * ndp_solver_profiled_pairs() pairs).                                                          */
int ndp_solver_profile(const ndp_solver* s, double* ms, int64_t* samples);
/* The driver splits a batch into contiguous stream groups (ndp_solver_cfg::streams, default 4)
 * that run on `stream` and on internal streams, so that one group's small kernels fill
 * the SM time the other groups' tensor-core CTAs leave idle.  The sampled launches of
 * ndp_solver_profile are the FIRST group's: this many pairs each.                            */
int32_t ndp_solver_profiled_pairs(const ndp_solver* s);
/* Work actually done by the culled search since creation (profile_every > 0, nn_mode 0): distance evaluations
 * issued (every scanned 32-target block costs 32 x 32 of them per warp of 32 queries) and 32-query blocks
 * searched; the brute-force equivalent is (queries x targets).  max_blocks (may be NULL): the most 32-target
 * blocks any warp of 32 queries scanned in one search (the tail that bounds an isolated launch).
 * Synchronises the device.                                                                               */
int ndp_solver_nn_stats(const ndp_solver* s, int64_t* pair_evals, int64_t* query_blocks, int64_t* max_blocks);

#ifdef __cplusplus
}
#endif
#endif /* NDP_B200_H */
