/*
 * cbfssm_b200 -- C ABI of the B200-native CBF-SSM sampled-ELBO hot path.
 *
 * The reference (silvanmelchior/CBF-SSM) has no FFI: the path is a TensorFlow-1.8
 * graph executed by sess.run.  Each entry point below names the reference graph
 * section it replaces (paths relative to the reference checkout).  A maintainer of
 * the reference binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is named *_host;
 *   - the caller owns every buffer, including the workspace; the library never
 *     allocates or frees device memory.  Process-wide state it does keep: (1) the
 *     register-resident kernels (compiled-in small M) read their packed GP operands from a
 *     per-instantiation __constant__ image that every call refreshes on its stream, so two
 *     calls for the same (dims, M) must not run concurrently on different streams of one
 *     process; (2) the float64 path holds one cuBLAS handle per host thread; (3) the
 *     thread-local measurement aids below (off by default);
 *   - all work is enqueued on the cudaStream_t passed in (as void*), no internal
 *     synchronisation (except cbf_timing_read and cbf_measure_fp32_peak, which say so); one
 *     host thread per GPU; calls on one stream are re-entrant in the sense that they keep
 *     nothing between calls except the caller's workspace;
 *   - return value: 0 ok, negative = CBF_ERR_* below, positive = cudaError_t;
 *     cbf_last_error_string() gives a thread-local description;
 *   - pointers must be 16-byte aligned.
 *   - particles are flattened n = b*S + s (the [B,S] order of cbfssm/model/cbfssm.py:134,149,209).
 *     A call handles the contiguous particle range [n_offset, n_offset+n_local) so
 *     that ranks of a data-parallel job can shard (b,s); draw tensors are LOCAL
 *     ([..., n_local]), u / y are the full [B,T,d] minibatch.
 */
#ifndef CBFSSM_B200_H
#define CBFSSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBF_ABI_VERSION 2

#if defined(__GNUC__)
#define CBF_API __attribute__((visibility("default")))
#else
#define CBF_API
#endif

#define CBF_ERR_INVALID_SHAPE   (-1)
#define CBF_ERR_UNSUPPORTED_DIMS (-2)  /* (dim_x,dim_u,dim_y) not compiled in */
#define CBF_ERR_UNSUPPORTED_M   (-3)   /* resident parameter set does not fit one SM */
#define CBF_ERR_ALIGNMENT       (-4)
#define CBF_ERR_NULL            (-5)

typedef struct cbf_shape {
  int32_t B;          /* minibatch size (sequences)                base_model.py:30 */
  int32_t S;          /* config['samples'] particles per sequence  cbfssm.py:70     */
  int32_t T;          /* sequence length                           base_model.py:31 */
  int32_t M;          /* config['ind_pnt_num']                     cbfssm.py:32     */
  int32_t dx, du, dy; /* config['dim_x'], ds.dim_u, ds.dim_y        cbfssm.py:26-28  */
  int32_t R;          /* config['recog_len']                       cbfssm.py:120    */
  int32_t condition;  /* the `condition` placeholder               base_model.py:23 */
  int32_t n_offset;   /* first global particle of this shard                        */
  int32_t n_local;    /* particles in this shard (B*S when unsharded)               */
  float   k_factor;   /* config['k_factor']                        cbfssm.py:191    */
  int32_t flags;      /* CBF_FLAG_* (0 = default kernel selection)                  */
} cbf_shape;

/* Use the cooperative shared-memory kernels even when a register-resident
 * instantiation for this (dims, M) is compiled in (parity tests cover both). */
#define CBF_FLAG_FORCE_COOPERATIVE 1
/* Do not use the tcgen05 tensor-core forward kernels (selected by default for 16 <= M <= 128). */
#define CBF_FLAG_NO_TENSOR_CORES 2
/* The register-resident kernels (compiled-in small M) and the tensor-core kernels (16 <= M <= 128) are the
 * default at every particle count; the cooperative kernels serve every other M.  These flags select a
 * specialised path explicitly (they fail if it does not exist for the shape). */
#define CBF_FLAG_FORCE_REGISTER 4
#define CBF_FLAG_FORCE_TENSOR_CORES 8
/* CBFSSMHALF (cbfssm/model/cbfssmhalf.py): no backward-message GP, x_0 supplied by the caller's
 * recognition model, only the first dy state dims are conditioned.  Use the *_half entry points. */
#define CBF_FLAG_HALF_MODEL 32
/* Float64 batched path (any M, any dims with dx <= 16, dx + du <= 31): per time step all particles are one
 * batch -- kernel matrix, one DGEMM against P, moments, step arithmetic -- as the reference graph itself is
 * laid out (cbfssm.py:107-111,176-179), entirely in float64 between stored states.  Selected automatically for
 * M > 128 (P no longer fits an SM) and for dims without a compiled instantiation; set the flag for inducing
 * sets whose cond(K_zz) is beyond float32's reach (crowded inducing points), where the float32 paths are
 * accuracy-limited.  Needs cbf_gp.state. */
#define CBF_FLAG_FP64 128
/* Prediction only (cbfssm/outputs/outputs.py:68-71,128-130: pred_mean / pred_var with condition = False): with
 * condition == 0 the rollout reads y2[t] only for t < R (cbfssm.py:227-228), so cbf_elbo_forward runs just the
 * message chain(s) that write those steps -- <= 2R message steps instead of ~2T.  terms[2] (entropy) then covers
 * only those steps and y_tilde beyond t = R-1 is not produced; cbf_elbo_backward refuses the flag.  No effect
 * when condition != 0. */
#define CBF_FLAG_PREDICT_ONLY 256


/* Kernel-level operands of one sparse GP (gp_tf.py:103-130), float32, produced by
 * cbf_gp_prologue (or by the caller).  Dout = dx for gp_f, dx-dy for gp_b. */
typedef struct cbf_gp {
  const float *Z;      /* [M, dx+du]  zeta_pos                                  */
  const float *ell;    /* [dx+du]     kern.lengthscales (constrained)           */
  const float *sig2;   /* [1]         kern.variance (constrained)               */
  const float *P;      /* [M, M]      (K_zz + 1e-8 I)^-1, symmetric             */
  const float *alpha;  /* [M, Dout]   P @ zeta_mean                             */
  const float *S;      /* [M, Dout]   zeta_var (constrained)                    */
  const double *state; /* the float64 state buffer cbf_gp_prologue filled (holds P, alpha, S, Z/ell,
                          ell, sig2 in float64); read by the float64 path only, may be NULL otherwise */
} cbf_gp;

/* Offsets (in doubles) into the flat kernel-level gradient vector written by
 * cbf_elbo_backward: d loss / d {P, alpha, S, Z, ell, sig2} per GP, then var_x, var_y. */
typedef struct cbf_grad_layout {
  int64_t f_P, f_alpha, f_S, f_Z, f_ell, f_sig2;
  int64_t b_P, b_alpha, b_S, b_Z, b_ell, b_sig2;
  int64_t var_x, var_y;
  int64_t total;
} cbf_grad_layout;

CBF_API int         cbf_abi_version(void);
CBF_API const char *cbf_last_error_string(void);

/* 0: not supported; 1: cooperative / tensor-core kernels; 2: register-resident kernels compiled in
 * for exactly this (M, dims) (used by default); 3: only the float64 batched path takes the shape
 * (M > 128, or dims without a compiled instantiation). */
CBF_API int cbf_supported(int32_t M, int32_t dx, int32_t du, int32_t dy);

/* Bytes of caller-allocated device workspace one forward+backward needs. */
CBF_API int cbf_workspace_bytes(const cbf_shape *shape, size_t *bytes_out);

CBF_API int cbf_grad_layout_get(const cbf_shape *shape, cbf_grad_layout *out);

/* Replaces the three tf.while_loop executions and the loss reductions:
 *   backward message, both runs   cbfssm.py:84-158
 *   forward conditional rollout   cbfssm.py:160-237
 *   log-likelihood sum            cbfssm.py:245-251
 * Inputs: u [B,T,du], y [B,T,dy] (base_model.py:29); var_x [dx], var_y [dx]
 * constrained (cbfssm.py:52,54); draws eps_b [2,T,n_local], z_b [2,T,n_local]
 * (read only at resample steps), eps_f [T-1,n_local].
 * Output: terms[0..2] = this shard's loglik, kl_x, entropy (float64); particle
 * states stay in `workspace` for cbf_elbo_backward / cbf_export_states. */
CBF_API int cbf_elbo_forward(const cbf_shape *shape, const cbf_gp *gp_f, const cbf_gp *gp_b,
                     const float *var_x, const float *var_y,
                     const float *u, const float *y,
                     const float *eps_b, const float *z_b, const float *eps_f,
                     double *terms, void *workspace, void *stream);

/* Replaces tf.gradients through the three loops (cbfssm.py:274-275) down to the
 * kernel-level operands.  term_weights_host[3] = d loss / d (loglik, kl_x, entropy),
 * i.e. (-l1/S, +l1/S, -l2/S) for cbfssm.py:257-262.  Must follow cbf_elbo_forward on
 * the same workspace and inputs.  grad_flat: cbf_grad_layout.total doubles. */
CBF_API int cbf_elbo_backward(const cbf_shape *shape, const cbf_gp *gp_f, const cbf_gp *gp_b,
                      const float *var_x, const float *var_y,
                      const float *u, const float *y,
                      const float *eps_b, const float *z_b, const float *eps_f,
                      const double *term_weights_host,
                      double *grad_flat, void *workspace, void *stream);

/* CBFSSMHALF variants (shape->flags must contain CBF_FLAG_HALF_MODEL; sequence-aligned shards only).
 * Replace the forward while_loop + loss of cbfssmhalf.py:97-193 and tf.gradients through it.
 * x0 [B, dx] float32: output of the recognition model (cbfssmhalf.py:64-95), tiled over the particles
 * inside; var_y: dx floats of which the first dy are used; terms[2] (entropy) is 0.
 * cbf_elbo_backward_half additionally returns x0_bar [n_local/S, dx] float64 = d loss / d x0 of this
 * shard's sequences; the gp_b block of grad_flat is zero. */
CBF_API int cbf_elbo_forward_half(const cbf_shape *shape, const cbf_gp *gp_f,
                          const float *var_x, const float *var_y,
                          const float *u, const float *y, const float *x0, const float *eps_f,
                          double *terms, void *workspace, void *stream);
CBF_API int cbf_elbo_backward_half(const cbf_shape *shape, const cbf_gp *gp_f,
                           const float *var_x, const float *var_y,
                           const float *u, const float *y, const float *x0, const float *eps_f,
                           const double *term_weights_host,
                           double *grad_flat, double *x0_bar, void *workspace, void *stream);

/* x_final [B?,T,S,dx] / y_tilde in the reference layout (cbfssm.py:97,181) for this
 * shard: out tensors are [n_local/S, T, S, dx] when the shard is sequence-aligned,
 * generally [T-major gather] n_local particles: layout [nb, T, S, dx] with
 * nb*S == n_local required. Either pointer may be NULL. */
CBF_API int cbf_export_states(const cbf_shape *shape, const float *y,
                      float *x_final, float *y_tilde, const void *workspace, void *stream);

/* Per-sequence partial sums of the forward states over THIS shard's particles, for tf.nn.moments over the
 * particle axis (cbfssm.py:267,269) when the particles of a sequence are split over ranks (any particle range):
 * sums [B, T, dx, 2] float64 = (sum_s x, sum_s x^2); sequences without a local particle get zeros.  The caller
 * all-reduces `sums` and forms mean = s1/S, var = s2/S - mean^2 (+ var_y for pred_var, cbfssm.py:268). */
CBF_API int cbf_state_sums(const cbf_shape *shape, double *sums, const void *workspace, void *stream);

/* tf.nn.moments(axes=[2]) of cbfssm.py:267-269 over the particle axis of a
 * [nb, T, S, d] tensor: mean and population variance (+ add_var[j] if non-NULL). */
CBF_API int cbf_moments(const float *x, int32_t nb, int32_t T, int32_t S, int32_t d, int32_t d_keep,
                const float *add_var, float *mean, float *var, void *stream);

/* Parameter-only prologue in float64 (gp_tf.py:22-31,104-130,163-172):
 * raw tensors -> constrained/derived operands (float32 copies for the rollout,
 * float64 kept in `state` for the adjoint) and KL(q(u)||p(u)).
 * state: cbf_gp_prologue_state_doubles(M,Din,Dout) doubles. */
CBF_API int64_t cbf_gp_prologue_state_doubles(int32_t M, int32_t Din, int32_t Dout);
CBF_API int cbf_gp_prologue(int32_t M, int32_t Din, int32_t Dout,
                    const double *zeta_pos, const double *zeta_mean, const double *zeta_var_unc,
                    const double *variance_unc, const double *lengthscales_unc,
                    float *Z32, float *ell32, float *sig232, float *P32, float *alpha32, float *S32,
                    double *kl_out, double *state, void *stream);
/* Adjoint: kernel-level gradients (pointers into grad_flat) + kl_weight * dKL ->
 * gradients w.r.t. the five raw tensors (float64). */
CBF_API int cbf_gp_prologue_backward(int32_t M, int32_t Din, int32_t Dout,
                             const double *gP, const double *galpha, const double *gS,
                             const double *gZ, const double *gell, const double *gsig2,
                             double kl_weight, double *state /* scratch part is overwritten */,
                             double *g_zeta_pos, double *g_zeta_mean, double *g_zeta_var_unc,
                             double *g_variance_unc, double *g_lengthscales_unc, void *stream);

/* Constrain / chain the two noise vectors (cbfssm.py:51-54). */
CBF_API int cbf_noise_forward(int32_t dx, const double *var_x_unc, const double *var_y_unc,
                      float *var_x32, float *var_y32, void *stream);
CBF_API int cbf_noise_backward(int32_t dx, const double *var_x_unc, const double *var_y_unc,
                       const double *g_var_x, const double *g_var_y,
                       double *g_var_x_unc, double *g_var_y_unc, void *stream);

/* TF-1.8 AdamOptimizer update on a flat float64 vector (cbfssm.py:274):
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps). step is 1-based. */
CBF_API int cbf_adam_step(int64_t n, double *theta, const double *grad, double *m, double *v,
                  int64_t step, double lr, double beta1, double beta2, double eps, void *stream);

/* Standard-normal draws for a throughput run (replaces tf.random_normal,
 * cbfssm.py:134,149,209): Philox4x32-10 + Box-Muller, counter = element index. */
CBF_API int cbf_fill_normal(float *out, int64_t n, uint64_t seed, uint64_t stream_id, void *stream);

/* Measurement aid for bench.py (the only thread-local state the library keeps, off by
 * default): when enabled, the four rollout kernels (0 backward-message forward,
 * 1 forward rollout, 2 forward-rollout reverse, 3 backward-message reverse, and on the tensor
 * path 4 / 5 the outer-product accumulation of the forward / message GP) are bracketed by CUDA
 * events on the caller's stream.  cbf_timing_read synchronises on
 * those events and returns the summed milliseconds and launch counts per kernel since
 * the last read. */
CBF_API int cbf_timing_enable(int enable);
CBF_API int cbf_timing_read(double *ms_sum_host /*[8]*/, int64_t *count_host /*[8]*/);

/* Measurement aid for bench.py's roofline denominator: runs a packed-FP32-FMA (fma.rn.f32x2) kernel that
 * fills every SM and returns the measured TFLOP/s of this GPU (best of 3 after one warm-up; synchronises).
 * scratch: at least 8 * 256 * (number of SMs) floats.  No reference counterpart. */
CBF_API int cbf_measure_fp32_peak(float *scratch, int iters, double *tflops_host, void *stream);

/* Number of kernels this library has launched on the calling host thread since the last reset
 * (measurement aid: bench.py's gpu_launches; there is no reference counterpart). */
CBF_API int cbf_launches_read(int64_t *count_host, int reset);

#ifdef __cplusplus
}
#endif
#endif /* CBFSSM_B200_H */
