/*
 * psvi_b200.h -- C ABI of libpsvi_b200.so: the B200-native (sm_100a) PSVI hot path.
 *
 * The reference (souravc83/Blackbox-Coresets-VI) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b), so every entry point below cites the reference *Python* interface it replaces.
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); nothing synchronises inside;
 *   - no allocation inside: the caller owns every buffer; *_workspace_bytes() say how large scratch must be;
 *   - return 0 on success, negative on error (PSVI_ERR_*); psvi_last_error() gives a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns PSVI_ERR_CUDA.
 *
 * Layouts ("TL" = theta layout): for an MLP with L VILinear layers of sizes dims[0..L]
 *   (psvi/models/neural_net.py:267-297, psvi/inference/psvi_classes.py:694-717) a per-sample weight vector is, per
 *   layer l=1..L, W_l[dims[l]][dims[l-1]] row-major followed by b_l[dims[l]];  P = sum_l dims[l]*(dims[l-1]+1).
 *   mu[P], rho[P] are the variational means / pre-softplus scales in TL (sigma = softplus(rho), neural_net.py:129-131).
 *   eps[...][S][P] are standard-normal draws in TL, one [S][P] slab per forward (weight draw then bias draw per
 *   layer, neural_net.py:155-162).  Labels are int32 class ids (the reference casts with .long(), Q11).
 */
#ifndef PSVI_B200_H
#define PSVI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSVI_MAX_LAYERS 6

#define PSVI_OK 0
#define PSVI_ERR_INVALID (-1)     /* bad argument / null pointer                       */
#define PSVI_ERR_UNSUPPORTED (-2) /* shape does not fit this kernel family             */
#define PSVI_ERR_CUDA (-3)        /* CUDA runtime error (message in psvi_last_error)   */

/* f(v) of the coreset weights a = N * f(v):  psvi_classes.py:111 (identity), :1358-1360 (softmax), :1486-1488. */
#define PSVI_VMODE_IDENTITY 0
#define PSVI_VMODE_SOFTMAX 1
#define PSVI_VMODE_EXPALPHA_SOFTMAX 2

/* inner optimiser arithmetic */
#define PSVI_ADAM_ROBUST_HIGHER 0 /* psvi/robust_higher/optim.py:303-367 (DifferentiableAdam, +1e-8 and v==0 mask) */
#define PSVI_ADAM_TORCH 1         /* torch.optim.Adam (baselines.py:1013)                                           */
#define PSVI_ADAM_HYPERGRAD 2     /* psvi/hypergrad/diff_optimizers.py:184-213 (u += 1e-12, sqrt(u/bc2)+eps)        */

/* noise source */
#define PSVI_NOISE_EXTERNAL 0 /* caller supplies eps (exact-noise mode, used for parity)                   */
#define PSVI_NOISE_PHILOX 1   /* in-kernel Philox4x32-10 + Box-Muller keyed by (seed, domain, slab, s, idx) */

/* phases of psvi_mf_nested_step */
#define PSVI_PHASE_UNROLL 1  /* T differentiable Adam steps on inner_elbo + outer psvi_elbo forward/backward */
#define PSVI_PHASE_REVERSE 2 /* reverse sweep through the T steps -> hypergradients                            */

typedef struct psvi_mf_model {
  int32_t n_layers;                   /* L >= 1: number of VILinear layers (1 = logistic_regression)  */
  int32_t dims[PSVI_MAX_LAYERS + 1];  /* dims[0]=D ... dims[L]=C                                       */
  int32_t mc_samples;                 /* S (psvi_classes.py:107)                                       */
} psvi_mf_model;

typedef struct psvi_noise {
  int32_t mode;          /* PSVI_NOISE_*                                                            */
  const float* eps;      /* EXTERNAL: [n_slabs][S][P] (TL); ignored for PHILOX                      */
  uint64_t seed;         /* PHILOX key                                                              */
  uint32_t domain;       /* PHILOX: stream id (callers bump it per outer step / evaluate call)      */
} psvi_noise;

/* ---- error string ------------------------------------------------------------------------------------------- */
const char* psvi_last_error(void);
/* number of SMs of the current device, or negative error: lets callers check a GPU is really there */
int psvi_device_sm_count(void);

/* ---- sizes -------------------------------------------------------------------------------------------------- */
/* P for a model (host arithmetic only). */
int64_t psvi_mf_num_theta(const psvi_mf_model* model);
/* bytes of trajectory scratch psvi_mf_nested_step / psvi_mf_unroll need for T steps: T*8*P floats */
size_t psvi_mf_traj_bytes(const psvi_mf_model* model, int32_t T);
/* floats in the phase-boundary buffer `gout` (what multi-GPU callers all-reduce between the two phases):
 * [2P: dLoss/d(mu_T,rho_T)] [M*D: direct dLoss/du] [M: direct dLoss/da] [S: data_nll_s] [4: loss, sum_s w_s e_s,
 * mean_s lw_s, sum_s beta_s] -- the all-reduced part ends here -- [S: w_s] [S: beta_s] [S: dLoss/dp_s] (diagnostics) */
int64_t psvi_mf_gout_floats(const psvi_mf_model* model, int32_t M);

/* ---- the bilevel step: replaces PSVI.nested_step + innerloop_ctx + DifferentiableAdam -----------------------
 * reference: psvi/inference/psvi_classes.py:541-600 (nested_step), :445-486 (psvi_elbo), :488-511 (inner_elbo);
 *            psvi/robust_higher/__init__.py:28-95, optim.py:152-257,299-367.
 * Computes, on device, T unrolled Adam steps on inner_elbo over the pseudo-data (u,z) weighted by a = N f(v), the outer
 * objective psvi_elbo on (u | xb), and its hypergradient wrt u and v by a hand-written reverse sweep (no autograd).
 * noise slabs: t = 0..T-1 inner steps, slab T = outer forward.
 *   mu, rho   [P]  in: phi_0 ; out (after PHASE_UNROLL): phi_T  (the copy-back of psvi_classes.py:596-599)
 *   u [M][D], z [M] int32, v [M], xb [B][D], yb [B] int32      (B = rows of the minibatch held by THIS rank)
 *   n_total_rows: B summed over ranks (= the reference's Nx);  pseudo_scale: 1/world_size (1 on a single GPU)
 *   traj: scratch of psvi_mf_traj_bytes();  gout: psvi_mf_gout_floats() floats
 *   u_grad [M][D], v_grad [M], alpha_grad [1] (nullable unless vmode 2)   (PHASE_REVERSE outputs)
 *   loss_out [1]: psvi_elbo value (this rank's share when sharded: shares sum to the loss)
 *   inner_losses [T] (nullable): inner_elbo value at every inner step (psvi_classes.py:551-554)
 */
int psvi_mf_nested_step(const psvi_mf_model* model, const psvi_noise* noise,
                        float* mu, float* rho,
                        const float* u, const int32_t* z, const float* v, int32_t M,
                        const float* xb, const int32_t* yb, int32_t B, int32_t n_total_rows,
                        float N, int32_t vmode, float alpha,
                        int32_t T, float lr, float pseudo_scale, int32_t phase_mask,
                        float* traj, float* gout,
                        float* u_grad, float* v_grad, float* alpha_grad,
                        float* loss_out, float* inner_losses, void* stream);

/* ---- plain (non-differentiated) inner optimisation: replaces run_mfvi_subset's training loop and hyper_step's
 * inner loop.  reference: psvi/inference/baselines.py:1019-1032;  psvi_classes.py:622-654 + hypergrad/diff_optimizers.py
 * Runs T Adam steps on  sum_s sum_m a_m nll[s,m] + KL  with a_m = row_weights[m] (nullable -> N f(v) from v).
 *   adam_m, adam_v [2P] in/out moments ([mu part | rho part] in TL), step0 = number of steps already taken.
 *   losses [T] nullable.  noise slabs t = 0..T-1. */
int psvi_mf_unroll(const psvi_mf_model* model, const psvi_noise* noise,
                   float* mu, float* rho, float* adam_m, float* adam_v, int32_t step0,
                   const float* x, const int32_t* y, const float* row_weights, const float* v, int32_t M,
                   float N, int32_t vmode, float alpha,
                   int32_t T, float lr, int32_t adam_mode, float* losses, void* stream);

/* ---- objective values and first/second-order products at a fixed phi (building blocks; also what the hyper trainer
 * and the parity tests call).
 * psvi_mf_outer_grad: psvi_elbo value and gradient wrt (mu, rho), u, v.   reference psvi_classes.py:445-486 (+ autograd)
 *   gout as in psvi_mf_gout_floats(); u_grad/v_grad receive the *direct* partials. noise slab 0. */
int psvi_mf_outer_grad(const psvi_mf_model* model, const psvi_noise* noise,
                       const float* mu, const float* rho,
                       const float* u, const int32_t* z, const float* v, int32_t M,
                       const float* xb, const int32_t* yb, int32_t B, int32_t n_total_rows,
                       float N, int32_t vmode, float alpha, float pseudo_scale,
                       float* gout, float* u_grad, float* v_grad, float* alpha_grad, float* loss_out, void* stream);

/* psvi_mf_inner_grad: inner_elbo value and gradient wrt (mu, rho).        reference psvi_classes.py:488-511 (+ autograd)
 *   grad [2P] = [d/dmu | d/drho]; value [1]. noise slab 0. */
int psvi_mf_inner_grad(const psvi_mf_model* model, const psvi_noise* noise,
                       const float* mu, const float* rho,
                       const float* u, const int32_t* z, const float* v, int32_t M,
                       float N, int32_t vmode, float alpha,
                       float* grad, float* value, void* stream);

/* psvi_mf_inner_hvp: second-order products of inner_elbo along direction gdot [2P]:
 *   h_phi [2P] = (d2 L / dphi dphi) gdot,  h_u [M][D] = (d2 L / du dphi) gdot,  h_v [M] = (d2 L / dv dphi) gdot.
 * reference: what torch.autograd's double backward yields inside nested_step (optim.py:224-229) and inside
 * hypergrad.CG_normaleq / jvp (psvi/hypergrad/hypergradients.py:199-244,308-311). noise slab 0. */
int psvi_mf_inner_hvp(const psvi_mf_model* model, const psvi_noise* noise,
                      const float* mu, const float* rho,
                      const float* u, const int32_t* z, const float* v, int32_t M,
                      float N, int32_t vmode, float alpha,
                      const float* gdot, float* h_phi, float* h_u, float* h_v, float* h_alpha, void* stream);

/* ---- predictive pass: replaces PSVI.evaluate (importance-weighted, psvi_classes.py:1031-1108) and the test loop of
 * run_mfvi_subset (mean-of-logits, baselines.py:1035-1043).
 *   xt [n_rows][D], yt [n_rows] int32: the rows held by THIS rank; batch = data_minibatch (rows per noise slab);
 *   first_slab: global index of this rank's first batch (so a sharded pass consumes the same slabs as one rank would).
 *   mode 0: importance-weighted (correction=True), 1: plain mean over samples of softmax (correction=False),
 *        2: softmax of mean logits (mfvi baselines; u/z/v ignored).
 *   out [8] (accumulated into with atomics-free two-stage reduction; caller zeroes it):
 *        [0] sum nll, [1] #correct, [2] #rows, [3] IW entropy (last slab), [4] normalised ESS (last slab), [5..7] spare
 *   scratch: psvi_mf_eval_scratch_bytes() bytes. */
size_t psvi_mf_eval_scratch_bytes(const psvi_mf_model* model, int32_t n_rows, int32_t batch);
int psvi_mf_evaluate(const psvi_mf_model* model, const psvi_noise* noise,
                     const float* mu, const float* rho,
                     const float* u, const int32_t* z, const float* v, int32_t M,
                     const float* xt, const int32_t* yt, int32_t n_rows, int32_t batch, int32_t first_slab,
                     float N, int32_t vmode, float alpha, int32_t mode,
                     float* out, void* scratch, void* stream);

/* ---- module-level forward: replaces nn.Sequential(VILinear, ReLU, ..., VILinear)(x)  (neural_net.py:155-179,267-297)
 *   x [n_rows][D] -> logits [S][n_rows][C]; theta_out [S][P] (nullable) receives the sampled weights in TL (the
 *   reference's _cached_weight/_cached_bias), nkl_out [S] (nullable) sum_layers sampled_nkl() (:110-115), kl_out [1]
 *   (nullable) sum_layers kl() (:101-108).  noise slab 0. */
int psvi_mf_forward(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                    const float* x, int32_t n_rows, float* logits, float* theta_out, float* nkl_out, float* kl_out,
                    void* stream);

/* ---- streaming ("medium regime") variants for models whose parameter vector does not fit the shared-memory-resident
 * engine (psvi_mf_unroll / psvi_mf_evaluate return PSVI_ERR_UNSUPPORTED), e.g. the baselines' fn with two 100-unit hidden
 * layers (experiments_utils.py:346-371: P = 10 602).  Same arithmetic, parameters in global memory, one CTA per MC
 * sample; `workspace` of psvi_mf_stream_workspace_bytes(model, n_rows) bytes (n_rows = 0 for training only).
 *   psvi_mf_unroll_stream: row weights either an array [M] or the scalar `row_weight_scalar`; adam_m/adam_v required.
 *   psvi_mf_evaluate_stream: a_weights [M] = N f(v) (mode 0); out [8] as psvi_mf_evaluate (zeroed inside). */
size_t psvi_mf_stream_workspace_bytes(const psvi_mf_model* model, int32_t n_rows);
int psvi_mf_unroll_stream(const psvi_mf_model* model, const psvi_noise* noise, float* mu, float* rho, float* adam_m,
                          float* adam_v, int32_t step0, const float* x, const int32_t* y, const float* row_weights,
                          float row_weight_scalar, int32_t M, int32_t T, float lr, int32_t adam_mode, float* losses,
                          void* workspace, void* stream);
int psvi_mf_evaluate_stream(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                            const float* u, const int32_t* z, const float* a_weights, int32_t M, const float* xt,
                            const int32_t* yt, int32_t n_rows, int32_t batch, int32_t first_slab, int32_t mode,
                            float* out, void* workspace, void* stream);

/* ---- building blocks of the streaming path for ANY variational family: the per-sample network pass on externally
 * supplied weights, and the predictive kernel on externally supplied weights.  Used with the mean-field sampling for
 * medium-size MLPs and with the full-covariance layer (fn2) below.
 * psvi_net_pass: theta [S][P] (TL); thetad [S][P] nullable.  thetad == NULL, tbar == NULL: forward only -> nll [S][R];
 *   thetad == NULL: gradient pass with per-sample row weights cw [S][R] -> tbar [S][P] (d/dtheta_s), xbar [S][R][D]
 *   (nullable, d/dx per sample), nll;  thetad != NULL: dual (Hessian-vector) pass, SURVEY Appendix A.6 -> tbar = A_theta,
 *   tdbar = A_thetadot, xbar = A_x, acbar [S][R] = adjoint of the row weights.
 *   Replaces VILinear(MultivariateNormal).forward + Categorical.log_prob + autograd (neural_net.py:176-179,485-491). */
int psvi_net_pass(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const int32_t* y,
                  const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                  float* logits /* [S][R][C], nullable */, void* stream);

/* The same per-sample pass with a GAUSSIAN likelihood of precision tau on a one-output network (the regressors: reference
 * psvi/inference/psvi_classes.py:1986 `gaussian_fn(scale = 1 / sqrt(tau))`, :2034-2057 psvi_elbo / inner_elbo of PSVI_regressor):
 *   nll[s][r] = tau / 2 (o_s(x_r) - y_r)^2 + 1/2 log(2 pi / tau),   y [R] float targets.
 * ybar [S][R] (nullable): gradient pass: d(sum_r cw[s][r] nll[s][r]) / dy_r = -cw tau (o - y); dual pass: its directional derivative
 * along thetad, -cw tau odot (the mixed term the hypergradient on learnable targets z needs, psvi_classes.py:2064-2087).
 * outputs [S][R] (nullable, forward mode): the network outputs. */
int psvi_net_pass_gaussian(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const float* y,
                           const float* cw, int32_t R, float tau, float* nll, float* tbar, float* tdbar, float* xbar,
                           float* acbar, float* ybar, float* outputs, void* stream);

/* ... and with a BERNOULLI likelihood on one logit (sparse-BBVI: reference psvi/inference/utils.py:85-141, `dist.Bernoulli(logits =
 * net(x).squeeze(-1)).log_prob(y)`):  nll[s][r] = softplus(o_s(x_r)) - y_r o_s(x_r),  y [R] float labels in {0, 1}. */
int psvi_net_pass_bernoulli(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const float* y,
                            const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                            float* outputs, void* stream);
/* psvi_net_predict: predictive metrics of rows xt with sampled weights theta [S][P]; mode 0 needs log_weights [S]
 *   (softmax-ed inside); out [8] as psvi_mf_evaluate; workspace of psvi_mf_stream_workspace_bytes(model, n_rows). */
int psvi_net_predict(const psvi_mf_model* model, const float* theta, const float* log_weights, int32_t mode,
                     const float* xt, const int32_t* yt, int32_t n_rows, float* out, void* workspace, void* stream);

/* ---- full-covariance layer (fn2): MultivariateNormalVIMixin, psvi/models/neural_net.py:408-491.  L = scale_tril stays
 * PACKED: dg[n] (diagonal), off[(n-1)(n-2)/2] in torch.tril_indices(n-1, n-1, -1) order (k = r(r-1)/2 + c; last row has
 * no off-diagonals, Q6).
 * psvi_fc_matvec: out[s][i] = base[i] + dg[i] eps[s][i] + sum_{c<i} off[k(i,c)] eps[s][c]   (rsample :467-472; base may
 *   be NULL; with (mdot, Ldot) it yields the tangent sample);  eps / out are row-major with leading dims ld_eps / ld_out.
 * psvi_fc_outer: g_base[i] = sum_s A[s][i]; g_dg[i] = sum_s A[s][i] eps[s][i]; g_off[k(r,c)] = sum_s A[s][r] eps[s][c]
 *   (what autograd accumulates into mean / _sd / _corr through scale_tril, :452-461). */
int psvi_fc_matvec(int32_t n, int32_t S, const float* base, const float* dg, const float* off, const float* eps,
                   int32_t ld_eps, float* out, int32_t ld_out, void* stream);
int psvi_fc_outer(int32_t n, int32_t S, const float* A, int32_t ld_a, const float* eps, int32_t ld_eps, float* g_base,
                  float* g_dg, float* g_off, void* stream);
/* The same two operations on the layer's parameter block phi = [mean | _sd | _corr] (2n + (n-1)(n-2)/2 floats, the
 * parameters_to_vector order of VILinearMultivariateNormal) with the transforms folded in (diag = softplus(_sd)) -- what
 * rsample / kl / sampled_nkl and autograd compute through scale_tril (neural_net.py:435-472):
 *   psvi_fc_sample:       phidot == NULL: theta[s] = mean + L eps[s];  else the tangent meandot + Ldot eps[s],
 *                         Ldot: diag sigmoid(_sd) _sddot, off _corrdot
 *   psvi_fc_reparam_grad: g [2n + c] = d/dphi of (sum_s <A[s], theta[s]> + kl_coef KL + nkl_coef sum log diag)
 *   psvi_fc_reparam_hvp:  h [2n + c] from A_theta, A_thetadot along phidot (SURVEY A.6 carried over to the Cholesky family) */
int psvi_fc_sample(int32_t n, int32_t S, const float* phi, const float* phidot, const float* eps, int32_t ld_eps, float* out,
                   int32_t ld_out, void* stream);
int psvi_fc_reparam_grad(int32_t n, int32_t S, const float* phi, const float* A, int32_t ld_a, const float* eps, int32_t ld_eps,
                         float kl_coef, float nkl_coef, float* g, void* stream);
int psvi_fc_reparam_hvp(int32_t n, int32_t S, const float* phi, const float* phidot, const float* A_t, const float* A_td,
                        int32_t ld_a, const float* eps, int32_t ld_eps, float* h, void* stream);

/* ---- tensor-core full-data predictive pass for the single-layer model (logistic_regression): the HBM-bound member of
 * the predictive kernels.  Same quantities as psvi_mf_evaluate (PSVI.evaluate,
 * psvi_classes.py:1031-1108) for ONE noise slab over all n_rows (the reference with data_minibatch >= n_rows), computed
 * with bf16 operands / fp32 accumulation: TMA-streamed row tiles, tcgen05.mma into TMEM, fused softmax / mixture / NLL.
 *   xt_bf16 [n_rows][D] bf16 row-major (16-byte aligned; D a multiple of 64, <= 256), yt [n_rows] int32,
 *   needs C <= 16 and S <= 16;  mode 0: importance weighted, 1: uniform mean of softmax.
 *   out [8] as psvi_mf_evaluate;  scratch: psvi_lr_predictive_tc_scratch_bytes() bytes. */
size_t psvi_lr_predictive_tc_scratch_bytes(const psvi_mf_model* model);
int psvi_lr_predictive_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                          const float* u, const int32_t* z, const float* v, int32_t M, const void* xt_bf16,
                          const int32_t* yt, int64_t n_rows, int32_t slab, float N, int32_t vmode, float alpha,
                          int32_t mode, float* out, void* scratch, void* stream);
/* fp32 -> bf16 (round to nearest even) copy of a row matrix, for callers that keep fp32 masters. */
int psvi_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ---- tensor-core sampled-GEMM forward for `fn` with ONE hidden layer in the large regime (kernels
 * A/D "large"; BASELINE config 5: D=256, H=1024, S=64): VILinear -> ReLU -> VILinear -> Categorical log-lik
 * (neural_net.py:155-179,267-297; psvi_classes.py:488-511,1031-1108) with bf16 operands / fp32 accumulation.  Sampled
 * first-layer weights stream from L2 by TMA, both GEMMs run on tcgen05 (the hidden activations stay in TMEM), softmax /
 * NLL / importance-weighted mixture are fused into the epilogue.
 *   needs n_layers == 2, D a multiple of 64 (<= 256), H a multiple of 128, C <= 16, S <= 64.
 * psvi_fn_predictive_tc: same contract as psvi_lr_predictive_tc (one noise slab over all n_rows; out [8]).
 * psvi_fn_nll_tc: wsum_out[s] = sum_r row_weights[r] * nll[s, r] (row_weights nullable -> 1), nkl_out[s] (nullable) =
 *   sum_layers sampled_nkl(), nll_out [S][n_rows] (nullable) -- the per-sample terms of inner_elbo / psvi_elbo.
 *   scratch: psvi_fn_tc_scratch_bytes(model, max rows of any call, M) bytes. */
size_t psvi_fn_tc_scratch_bytes(const psvi_mf_model* model, int64_t max_rows, int32_t M);
int psvi_fn_predictive_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                          const float* u, const int32_t* z, const float* v, int32_t M, const void* xt_bf16,
                          const int32_t* yt, int64_t n_rows, int32_t slab, float N, int32_t vmode, float alpha,
                          int32_t mode, float* out, void* scratch, void* stream);
int psvi_fn_nll_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                   const void* x_bf16, const int32_t* labels, const float* row_weights, int64_t n_rows, int32_t slab,
                   float* wsum_out, float* nkl_out, float* nll_out, void* scratch, void* stream);

/* ---- large-regime per-sample network pass for `fn` with ONE hidden layer (BASELINE config 5: D=256, H=1024, S=64,
 * M=1000): same contract as psvi_net_pass (forward / gradient / dual Hessian-vector pass on sampled weights theta [S][P],
 * reference neural_net.py:155-179 + autograd), for models whose per-sample weights fit no CTA.  Every matrix product is a
 * batched TMA + tcgen05 GEMM with fp32 accumulation (csrc/psvi_fn_large.cu).  precision 0: bf16 operands / activations
 * (full tensor rate, ~1e-2 relative error: values, first-order training, prediction); precision 1: "tf32x3" -- every operand
 * is an fp32 (hi, lo) pair of TF32 numbers and every K step issues three kind::tf32 MMAs (fp32-class accuracy, which the
 * unrolled hypergradient of nested_step needs: its reverse sweep through Adam divides by |g_i|); precision 2: "bf16x3" -- the
 * same split with BF16 numbers and three kind::f16 MMAs (16 mantissa bits per operand, twice the MMA rate and half the operand
 * bytes of tf32x3; hypergradient cosine >= 0.999 against the fp64 oracle: opt-in).
 *   needs n_layers == 2, D a multiple of 64, H a multiple of 128, C <= 16, S <= 64;  x [R][D] fp32, y [R] int32,
 *   cw [S][R] (nullable -> 1);  outputs as psvi_net_pass (logits [S][R][C]);  workspace: psvi_fnl_workspace_bytes(). */
size_t psvi_fnl_workspace_bytes(const psvi_mf_model* model, int32_t R, int32_t precision);
int psvi_fnl_pass(const psvi_mf_model* model, int32_t precision, const float* theta, const float* thetad, const float* x,
                  const int32_t* y, const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                  float* logits, void* workspace, void* stream);

/* ---- PSVI.evaluate over MANY test batches on the tensor path in ONE call (reference psvi_classes.py:1038-1092: a fresh
 * noise slab per test batch): loops the slabs on the host side of the library -- psvi_lr_predictive_tc for the single-layer model,
 * psvi_fn_predictive_tc for fn -- and accumulates out[0..2] (nll sum, correct, rows) on the device; out[3..4] are the
 * importance-weight diagnostics of the LAST slab (Q12).  scratch: the per-slab scratch of the underlying call for `batch` rows
 * + 512 bytes. */
int psvi_predictive_tc_slabs(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                             const float* u, const int32_t* z, const float* v, int32_t M, const void* xt_bf16,
                             const int32_t* yt, int64_t n_rows, int32_t batch, int32_t first_slab, float N, int32_t vmode,
                             float alpha, int32_t mode, float* out, void* scratch, void* stream);

/* ---- full-data DATA-TERM gradient of the outer objective on the tensor path (the part of the hot path that shards).
 * reference: data_nll = N / Nx * all_nlls[:, Nu:].sum(-1) inside psvi_elbo (psvi/inference/psvi_classes.py:477,484-486)
 * and what autograd back-propagates through it; SURVEY.md section 7 step 7 / section 8e.
 * For externally sampled weights theta [S][P] (theta layout) and a shard of R rows x_bf16 [R][D] (bf16), y [R] int32:
 *     dsum [S]    = sum_r nll[s, r]
 *     tbar [S][P] = coef[s] * sum_r d nll[s, r] / d theta_s        (overwritten)
 * coef [S] is the caller's per-sample coefficient (w_s N / B: the importance weights depend on the pseudo-data only, so the
 * rows enter linearly and ranks add up: ONE all-reduce of the reparameterised gradient + dsum).  Two tcgen05 passes over
 * the rows (forward + seeds, then the backward with the W1 adjoint slices accumulated in tensor memory across row tiles);
 * bf16 operands, fp32 accumulation.  Same shape limits as psvi_fn_predictive_tc (D multiple of 64 <= 256, H multiple of
 * 128, C <= 16, S <= 64).  scratch: psvi_fn_data_grad_tc_scratch_bytes(model, R). */
size_t psvi_fn_data_grad_tc_scratch_bytes(const psvi_mf_model* model, int64_t n_rows);
int psvi_fn_data_grad_tc(const psvi_mf_model* model, const float* theta, const void* x_bf16, const int32_t* y, int64_t n_rows,
                         const float* coef, float* dsum, float* tbar, void* scratch, void* stream);

/* ---- convolutional family (lenet): per-sample network pass on externally supplied weights, same contract as
 * psvi_net_pass.  Replaces VIConv2d.forward (grouped conv over samples), BatchMaxPool2d, nn.Flatten and the three VILinear
 * layers of make_lenet + Categorical.log_prob + autograd (psvi/models/neural_net.py:194-255,334-359).
 *   theta / thetad [S][P] with P = psvi_lenet_num_theta() = 61 706 in TL (per layer weight then bias, module order);
 *   x [R][784] (1 x 28 x 28 images, shared by all samples), y [R] int32, cw [S][R] per-sample row weights (nullable -> 1);
 *   tbar == NULL: forward only -> nll [S][R] (+ logits [S][R][10]);  thetad == NULL: gradient pass -> tbar [S][P], xbar
 *   [S][R][784] (nullable);  thetad != NULL: dual (Hessian-vector) pass -> tbar = A_theta, tdbar = A_thetadot, xbar = A_x,
 *   acbar [S][R] = adjoint of the row weights.   workspace: psvi_lenet_workspace_bytes(S, R) bytes.
 * psvi_logits_predict: predictive metrics (PSVI.evaluate, psvi_classes.py:1072-1092; baselines.py:1039-1043) from per-sample
 *   logits [S][R][C]; mode / out [8] as psvi_mf_evaluate (mode 0 needs log_weights [S], softmax-ed inside). */
int64_t psvi_lenet_num_theta(void);
size_t psvi_lenet_workspace_bytes(int32_t S, int32_t R);
int psvi_lenet_pass(int32_t S, const float* theta, const float* thetad, const float* x, const int32_t* y, const float* cw,
                    int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar, float* logits,
                    void* workspace, void* stream);
int psvi_logits_predict(const float* logits, const float* log_weights, int32_t mode, const int32_t* yt, int32_t S, int32_t R,
                        int32_t C, float* out, float* probs_out /* [R][C] mixture per row, nullable; yt nullable then */,
                        void* stream);

/* ---- the mean-field family as fused maps over [S][P] slabs (theta layout), used by the streaming engine.  Replace
 * VIMixin.rsample / kl / sampled_nkl (psvi/models/neural_net.py:101-115,155-162) and what autograd accumulates into
 * (weight, bias, _weight_sd, _bias_sd) through theta = mu + softplus(rho) eps.
 * `mask` [P] (nullable = ones): 1 for parameters that enter the KL / sampled-nkl sums (0 for conv layers, quirk Q5).
 *   psvi_mf_sample:       theta[s] = mu + softplus(rho) eps[s]
 *   psvi_mf_tangent:      thetad[s] = mud + sigmoid(rho) rhod eps[s]
 *   psvi_mf_reparam_grad: g [2P] = [sum_s tb + mask kl_coef mu | sigmoid(rho)(sum_s tb eps + mask (kl_coef (sg - 1/sg) + nkl_coef/sg))]
 *                         with tb = tbar - beta[s] mask theta[s]  (beta [S] / theta nullable)
 *   psvi_mf_reparam_hvp:  h [2P] from A_theta, A_thetadot (SURVEY A.6 last line)
 *   psvi_mf_nkl_kl:       out[0..S-1] = sampled nkl per sample, out[S] = KL (double); scratch psvi_mf_nkl_scratch_bytes(S) */
int psvi_mf_sample(int32_t S, int64_t P, const float* mu, const float* rho, const float* eps, float* theta, void* stream);
int psvi_mf_tangent(int32_t S, int64_t P, const float* rho, const float* mud, const float* rhod, const float* eps, float* thetad,
                    void* stream);
int psvi_mf_reparam_grad(int32_t S, int64_t P, const float* mu, const float* rho, const float* eps, const float* tbar,
                         const float* beta, const float* theta, const float* mask, float kl_coef, float nkl_coef, float* g,
                         void* stream);
int psvi_mf_reparam_hvp(int32_t S, int64_t P, const float* rho, const float* mud, const float* rhod, const float* eps,
                        const float* A_t, const float* A_td, const float* mask, float* h, void* stream);
size_t psvi_mf_nkl_scratch_bytes(int32_t S);
int psvi_mf_nkl_kl(int32_t S, int64_t P, const float* mu, const float* rho, const float* eps, const float* theta, const float* mask,
                   double* out, void* scratch, void* stream);

/* ---- the unrolled robust Adam of the inner loop, one fused elementwise kernel per direction.  Replace the tensor expressions
 * of DifferentiableAdam._update and what autograd replays through them (psvi/robust_higher/optim.py:303-367, the
 * _maybe_mask hook at optim.py:40-52).  k = lr / (1 - 0.9^t), sq2 = sqrt(1 - 0.999^t), t = 1-based step.
 *   psvi_adam_unroll_step:    (phi, g, m, v) -> (phi', m', v')
 *   psvi_adam_unroll_reverse: (pbar; g, m', v'; mbar, vbar in/out) -> gbar, the adjoint of this step's gradient g */
int psvi_adam_unroll_step(int64_t n, float k, float sq2, const float* phi, const float* g, const float* m, const float* v,
                          float* phi_out, float* m_out, float* v_out, void* stream);
int psvi_adam_unroll_reverse(int64_t n, float k, float sq2, const float* pbar, const float* g, const float* m_t, const float* v_t,
                             float* mbar, float* vbar, float* gbar, void* stream);

/* ---- noise: the in-kernel generator, exposed so that callers/tests can materialise the exact slabs a PHILOX-mode
 * call consumes.  out [n_slabs][S][P]. */
int psvi_philox_normal(uint64_t seed, uint32_t domain, int32_t first_slab, int32_t n_slabs, int32_t S, int32_t P,
                       float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSVI_B200_H */
