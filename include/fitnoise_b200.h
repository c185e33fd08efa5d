/* fitnoise_b200 -- C ABI of the B200-native Fitnoise hot path.
 *
 * The reference (pfh/fitnoise 2.2) is pure Python and has NO foreign-function interface; this is
 * the boundary a maintainer would bind from Python with ctypes (see INTEGRATION.md for the stub)
 * to replace the per-gene Python loops of the reference.  Each entry point names the reference
 * code it stands in for (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C, no torch/numpy types; every array argument is a HOST pointer unless its name ends
 *     in _dev; matrices are row-major float64 (numpy C order, env.py:57-65); missing = NaN.
 *   - every function returns 0 on success, a negative code on failure; fn_last_error() gives the
 *     message for the calling thread.  There is no CPU fallback: without a CUDA device fn_create
 *     fails.
 *   - a handle owns one GPU's shard of the genes (device memory, stream, pinned result buffer).
 *     One host thread per handle.  Multi-GPU: one handle per process/GPU, genes sharded by the
 *     caller.  After fn_comm_mailbox / fn_comm_connect the per-evaluation sums of all shards are
 *     added inside the evaluation kernel itself (peer stores over NVLink) and fn_nll_grad returns
 *     the totals over all ranks; without them the caller adds the shards' sums (e.g. an NCCL
 *     all-reduce of the P+3 doubles fn_nll_grad_dev leaves in device memory).
 */
#ifndef FITNOISE_B200_H
#define FITNOISE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fn_handle fn_handle;

/* Variance family of a noise model: sigma2_j = theta_a(j) * U_j + theta_b(j) * T_j.
 * Stands in for the plug-in interface Model._get_dist / _unpack (fitnoise.py:198-205, 410-428,
 * 522-535, 573-578, 599-614, 659-678, 705-724, 757-776, 807-832, 863-882).
 * theta is ordered as the reference's _initial: [df if dist==FN_DIST_T], theta_a block, theta_b block. */
enum { FN_KIND_SCALE1 = 0, FN_KIND_SCALE_PS = 1, FN_KIND_PATSEQ = 2 };
enum { FN_DIST_NORMAL = 0, FN_DIST_T = 1, FN_DIST_T_FIXED = 2 };
typedef struct fn_family {
    int32_t kind;       /* FN_KIND_*: SCALE1 sigma2 = theta/w (Model_normal/_t/_independent);
                           SCALE_PS sigma2_j = theta_j/w_j (Model_*_per_sample);
                           PATSEQ sigma2_j = theta_a/count_j + theta_b*tail^2 (Model_*_patseq_*) */
    int32_t na, nb;     /* entries in the theta_a / theta_b blocks: 0, 1 or m */
    int32_t tail_own;   /* PATSEQ v1: tail = the gene's own y_j; else the count-weighted average tail */
    int32_t v3;         /* PATSEQ v3: sigma2_j = (theta_a/count_j + theta_b) * avgtail^2 */
    int32_t dist;       /* FN_DIST_* : Mvnormal, Mvt with df = theta[0], Mvt with fixed df */
    int32_t nf;         /* unknown batch-effect factors (Model_*_factors, fitnoise.py:439-518), 0..4; SCALE1 only */
    double fixed_df;    /* Model_independent: 1e-10 (fitnoise.py:577) */
} fn_family;

/* summary of fn_upload's preparation pass */
typedef struct fn_prep_info {
    double row_variance_mean;   /* mean over genes of var(y_row) over finite entries (fitnoise.py:546,627) */
    double cst_sum;             /* theta-independent part of sum_g ln det S_g over eligible genes */
    int64_t n_eligible;         /* genes entering the REML sum: k > c and full-rank retained design (fitnoise.py:28) */
    int64_t n_bad_counts;       /* PATSEQ: observations with count == 0 but finite y (fitnoise.py:647) */
    int64_t n_variance_rows;    /* genes with >= 1 finite value (the rows averaged in row_variance_mean) */
} fn_prep_info;

const char* fn_last_error(void);
int fn_device_count(void);
const char* fn_version(void);
/* hash of the sources the library was built from (the binding refuses a stale binary) */
const char* fn_source_hash(void);

/* One handle per GPU.  `stream` is a cudaStream_t to launch on (e.g. torch's current stream) or
 * NULL for a stream owned by the handle. */
int fn_create(int device, void* stream, fn_handle** out);
int fn_destroy(fn_handle* h);
int fn_synchronize(fn_handle* h);

/* Upload a shard of genes and prepare it: replaces Model.fit_noise's set-up and the per-gene prep
 * loop of _fit_noise (fitnoise.py:211-233, 23-35) and Model_*._configured (fitnoise.py:538-552 ...
 * 843-861).  y: n x m.  aux: n x m precision WEIGHTS (SCALE kinds; w = 0 or NaN drops the
 * observation) or read COUNTS (PATSEQ), or NULL.  controls: n bytes (non-zero = control gene,
 * fitted with design_ctrl) or NULL.  Designs are m x c row-major, c <= 16 (c + n_factors <= 16), full column rank. */
int fn_upload(fn_handle* h, int64_t n, int32_t m, const double* y, const double* aux,
              const fn_family* family, const uint8_t* controls,
              int32_t c_noise, const double* design_noise,
              int32_t c_ctrl, const double* design_ctrl, fn_prep_info* info);
/* Same with y/aux/controls already in device memory (n_pad = n rounded up to 1024 rows is NOT
 * required of the caller; the data is copied device-to-device into the handle's padded buffers). */
int fn_upload_dev(fn_handle* h, int64_t n, int32_t m, const double* y_dev, const double* aux_dev,
                  const fn_family* family, const uint8_t* controls_dev,
                  int32_t c_noise, const double* design_noise,
                  int32_t c_ctrl, const double* design_ctrl, fn_prep_info* info);

/* One evaluation of the REML objective: sum over this handle's genes of score_row (fitnoise.py:
 * 44-50) and its gradient (theano.gradient.grad, fitnoise.py:93; accumulation :98-106,125-136).
 * theta has P = (dist==T) + na + nb entries.  value includes the theta-independent terms.
 * n_used / n_bad (may be NULL): genes that contributed / that produced a non-finite term. */
int fn_nll_grad(fn_handle* h, const double* theta, int32_t P, double* value, double* grad,
                int64_t* n_used, int64_t* n_bad);
/* Resident evaluation: between fn_resident_begin and fn_resident_end ONE kernel stays on the GPU and
 * fn_nll_grad hands it each theta through a command block in mapped pinned memory -- no kernel
 * launch and no launch latency per evaluation (what the optimiser loop of fitnoise.py:143-158 needs:
 * tens of evaluations of a few microseconds each).  Results are identical to the one-launch form.
 * Any other entry point ends the resident phase by itself; the kernel also leaves on its own after
 * FN_RESIDENT_IDLE_MS (default 20) without a command and is relaunched transparently by the next
 * fn_nll_grad.  Not for factor models.  While resident the handle's GPU runs nothing else. */
int fn_resident_begin(fn_handle* h);
int fn_resident_end(fn_handle* h);
/* returns 1 while resident; launches = resident kernels started, evaluations = commands executed */
int64_t fn_resident_stats(fn_handle* h, int64_t* launches, int64_t* evaluations);

/* Benchmarking aid: sweeps a buffer twice the size of the L2 so that the next evaluation reads its
 * input from DRAM (bench.py uses it between timed steps when a shard would stay L2-resident). */
int fn_l2_flush(fn_handle* h);

/* Asynchronous form for multi-GPU: leaves [value, grad[0..P-1], n_used, n_bad] (P+3 doubles) in
 * device memory at out_dev on the handle's stream, for the caller to allreduce.  No host sync. */
int fn_nll_grad_dev(fn_handle* h, const double* theta, int32_t P, double* out_dev);

/* Factor models (family.nf > 0; Model_normal_factors / Model_t_factors / Model_independent_factors,
 * fitnoise.py:439-518, 558-565, 592-594): covariance diag(sigma2) + F F' with F (m x nf, row-major)
 * shared by all genes.  theta holds only [df], [variance]; the caller maps its packed parameters
 * to F (F = Q2 . pack, fitnoise.py:447-458) and the returned d/dF back (and adds the penalty
 * _model_cost, :470-478).  fn_nll_grad_factors also makes `factors` current for fn_per_gene,
 * fn_noise_stats, fn_coef and fn_weights; fn_set_factors does only that. */
int fn_nll_grad_factors(fn_handle* h, const double* theta, int32_t P, const double* factors,
                        double* value, double* grad, double* grad_factors, int64_t* n_used, int64_t* n_bad);
int fn_set_factors(fn_handle* h, const double* factors);

/* Per-gene score_row, Mahalanobis distance d2 and residual df p at theta (NaN / 0 for genes the
 * REML prep skips).  Parity/diagnostic entry; any output pointer may be NULL. */
int fn_per_gene(fn_handle* h, const double* theta, int32_t P, double* score, double* d2, int32_t* p);

/* Tail of Model.fit_noise (fitnoise.py:257-267): noise p-values and row averages. */
int fn_noise_stats(fn_handle* h, const double* theta, int32_t P, double* noise_p, double* averages);

/* Model.fit_coef (fitnoise.py:294-347): GLS coefficients and their posterior at the fitted theta.
 * design: m x c.  coef: n x c; covar: n x c x c; dfpost: n (NaN for Mvnormal).  Rows the reference
 * leaves as None are NaN.  Results also stay on the device for fn_test. Outputs may be NULL. */
int fn_coef(fn_handle* h, const double* theta, int32_t P, int32_t c, const double* design,
            double* coef, double* covar, double* dfpost);

/* Model.test (fitnoise.py:368-396): contrasts (c x kc) of the last fn_coef, their p-values.
 * contrasted: n x kc; ccov: n x kc x kc or NULL; pval: n. */
int fn_test(fn_handle* h, int32_t kc, const double* contrasts, double* contrasted, double* ccov, double* pval);

/* fn_test followed by Benjamini-Hochberg over this handle's genes without p leaving the device
 * (Model.test computes q_values right after p_values, fitnoise.py:395). */
int fn_test_bh(fn_handle* h, int32_t kc, const double* contrasts, double* contrasted, double* pval, double* qval,
               uint32_t* order /* n genes by ascending p (stable, NaN last): the top-table order; or NULL */);

/* Copy the device-resident results of the last fn_coef to the host (any pointer may be NULL): lets a
 * caller take `coef` at once and the c x c covariances only if somebody asks for coef_dists. */
int fn_coef_fetch(fn_handle* h, double* coef, double* covar, double* dfpost);

/* p_adjust.fdr (p_adjust.py:3-21): Benjamini-Hochberg q-values; NaN p counts as 1.0. */
int fn_bh(fn_handle* h, int64_t n, const double* p, double* q);
/* ... and the top-table order: genes by ascending p, ties in input order, NaN last (what the legacy
 * R test.fit sorts its table by, backyard/old_fitnoise.R:822-823). */
int fn_bh_ranked(fn_handle* h, int64_t n, const double* p, double* q, uint32_t* order);
/* device-resident form: p_dev, q_dev are device pointers of n doubles */
int fn_bh_dev(fn_handle* h, int64_t n, const double* p_dev, double* q_dev);

/* Device-resident results.  Every output pointer of fn_noise_stats / fn_coef / fn_test / fn_test_bh
 * may be NULL: the results then stay in device memory, where fn_result_dev finds them (pointer,
 * rows = this handle's genes, columns) -- a multi-GPU caller all-gathers them device-to-device
 * (NCCL on the same stream) instead of bouncing every array through the host.  fn_fetch copies
 * device memory to the host on the handle's stream and waits. */
enum { FN_RES_SCORE = 0, FN_RES_D2 = 1, FN_RES_PDF = 2 /* int32 */, FN_RES_NOISE_P = 3, FN_RES_AVERAGES = 4,
       FN_RES_COEF = 5, FN_RES_COVAR = 6, FN_RES_DFPOST = 7, FN_RES_CONTRASTED = 8, FN_RES_P = 9,
       FN_RES_Q = 10, FN_RES_ORDER = 11 /* uint32 */ };
int fn_result_dev(fn_handle* h, int32_t which, void** ptr, int64_t* rows, int32_t* cols);
int fn_fetch(fn_handle* h, void* dst_host, const void* src_dev, int64_t bytes);
/* Benjamini-Hochberg over device-resident p (e.g. all ranks' p after a gather); *order_dev: genes by
 * ascending p, in the handle's memory, valid until the next BH call. */
int fn_bh_ranked_dev(fn_handle* h, int64_t n, const double* p_dev, double* q_dev, uint32_t** order_dev);
/* Pinned host memory from a process-wide pool (result arrays are DMA targets; pinning is slow, so
 * freed blocks are reused). */
int fn_pinned_alloc(int64_t bytes, void** out);
int fn_pinned_free(void* p);

/* Model.get_weights (fitnoise.py:399-404): 1/sigma2, n x m. */
int fn_weights(fn_handle* h, const double* theta, int32_t P, double* weights);

/* Fused multi-GPU all-reduce (one process per GPU on one NVLink/NVSwitch box).  After these two
 * calls fn_nll_grad returns the sums over ALL ranks: the last CTA of each rank's kernel stores its
 * sums into every peer's mailbox (device memory mapped with CUDA IPC, plain stores over NVLink),
 * waits for the peers' flags and adds the slots in rank order -- compute, collective and hand-over
 * to the host in one kernel.  Every rank must then call fn_nll_grad the same number of times.
 *   fn_comm_mailbox  allocates this rank's mailbox for `world` senders of up to max_values doubles
 *                    and returns its 64-byte CUDA IPC handle for the caller to all-gather;
 *   fn_comm_connect  takes the world x 64 bytes of gathered handles (rank order). */
int fn_comm_mailbox(fn_handle* h, int32_t world, int32_t max_values, void* ipc_handle_out);
int fn_comm_connect(fn_handle* h, int32_t rank, int32_t world, const void* ipc_handles);
int fn_comm_disconnect(fn_handle* h);

/* fitnoise.transform (fittransform.py:65-219): moderated log transform of a count matrix,
 * y = ln(P_j(x)) / (order ln 2) with a per-sample monic polynomial P_j of degree `order`.
 *   fn_transform_upload  x: n x m counts, design: m x c (fittransform.py:76-85)
 *   fn_transform_sums    one pass over x at parameters a (order x m, row i = coefficient of
 *                        x^(order-1-i), Transform_polynomial._unpack_transform :180-186): returns
 *                        [Syy, Stt, s[m], A[order][m], B[order][m], D[order][m]] from which the caller
 *                        forms the objective ln(var_h1) - ln(var_h0) (:103-121) and its gradient
 *                        (theano.gradient.grad, :137) -- raw sums so that shards of genes add up.
 *   fn_transform_apply   y (n x m) = transform(x) - col_mean  (:161-163) */
int fn_transform_upload(fn_handle* h, int64_t n, int32_t m, const double* x, int32_t c, const double* design);
int fn_transform_sums(fn_handle* h, int32_t order, const double* a, double* sums);
int fn_transform_apply(fn_handle* h, int32_t order, const double* a, const double* col_mean, double* y);

/* Tuning knobs for experiments (0 = automatic): lanes per gene, threads per CTA, pipeline stages,
 * CTAs per SM. */
int fn_set_tuning(fn_handle* h, int32_t lpg, int32_t threads, int32_t stages, int32_t ctas_per_sm);
/* How many kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t fn_launch_count(fn_handle* h);
/* Average device time (ms) of the last fn_nll_grad kernel, measured with CUDA events on the
 * handle's stream when timing is enabled with fn_set_timing(h, 1). */
int fn_set_timing(fn_handle* h, int32_t on);
/* Kernel-only duration of the nll+grad kernel: `reps` launches queued back to back between two CUDA
 * events (no idle gap, so launch latency is not counted); this handle's genes only (no exchange). */
int fn_time_nll_grad(fn_handle* h, const double* theta, int32_t P, int32_t reps, double* ms_per_launch);
double fn_last_kernel_ms(fn_handle* h);

#ifdef __cplusplus
}
#endif
#endif
