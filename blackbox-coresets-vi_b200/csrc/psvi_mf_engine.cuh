// psvi_mf_engine.cuh -- launch parameters and phase flags shared by the two cluster engines of the small regime:
// the generic one (psvi_mf_engine.cu: any MLP depth that fits shared memory) and the shape-specialised one for
// one-hidden-layer networks with tiny input / output widths (psvi_mf_fn1.cu: the BASELINE cfg2 family).
#pragma once
#include "psvi_mf_gemm.cuh"

namespace psvi_mf {

enum : int {
  F_UNROLL = 1,       // T Adam steps on the inner objective
  F_OUTER = 2,        // psvi_elbo forward + backward at the current phi
  F_REVERSE = 4,      // reverse sweep through the trajectory
  F_HVP = 8,          // one Hessian-vector pass along p.gdot
  F_NOUPDATE = 16,    // F_UNROLL: compute the gradient but do not move phi (inner_grad entry point)
  F_STORE_GOUT = 32,  // write the phase-boundary buffer (outer gradient wrt phi_T, direct u/a partials, d_s, loss)
  F_LOAD_GOUT = 64,   // start the reverse sweep from the phase-boundary buffer
  F_FINAL = 128,      // reduce ubar/abar over the cluster and write u_grad / v_grad
  F_WRITE_PHI = 256,  // write phi (and Adam moments if given) back after F_UNROLL
  F_EVAL = 512        // predictive kernels: forward only, lean shared-memory carve-up
};

struct EP {
  int L;
  int dims[MAXL + 1];
  int S, M, B, Btot;
  int G, RC, slice;
  float Nf;
  int vmode;
  float alpha;
  int flags;
  int T, step0;
  float lr;
  int adam_mode;
  float kappa;
  int noise_mode;
  unsigned long long seed;
  unsigned domain;
  float* mu;
  float* rho;
  float* adam_m;
  float* adam_v;
  const float* u;
  const int* z;
  const float* v;
  const float* roww;
  const float* xb;
  const int* yb;
  const float* eps;
  float* traj;
  float* gout;
  float* u_grad;
  float* v_grad;
  float* alpha_grad;
  float* loss_out;
  float* inner_losses;
  float* g_out;
  const float* gdot;
  float* h_phi;
  // predictive pass
  int n_rows, batch, first_slab, eval_mode, n_slabs, chunks_per_slab;
  float* eval_w;     // [n_slabs][S] LOG importance weights (softmax-ed by the consumers)
  float* eval_part;  // [n_ctas][4] per-CTA partial sums
  float* eval_out;   // [8]
  // plain forward
  float* logits_out; float* theta_out; float* nkl_out; float* kl_out;
  long long* tl;     // debug timeline (clock64 stamps of rank 0 / thread 0), normally NULL
};


// psvi_mf_fn1.cu: launches the specialised engine if the shape qualifies.  Returns PSVI_OK when launched, a negative
// PSVI_ERR_* on a CUDA error, and FN1_NOT_APPLICABLE when the caller should use the generic engine.
constexpr int FN1_NOT_APPLICABLE = 1;
int psvi_fn1_launch(EP& p, cudaStream_t stream);

}  // namespace psvi_mf
