// psvi_family.cu -- the mean-field variational family seen through its maps (sample / tangent / nkl / kl and the first- and
// second-order reparameterisation adjoints), as fused kernels over [S][P] slabs, for the streaming PSVI engine
// (psvi/inference/stream.py).  Reference: VIMixin.rsample / kl / sampled_nkl (psvi/models/neural_net.py:101-115,155-162) and
// what autograd accumulates into (weight, bias, _weight_sd, _bias_sd) through theta = mu + softplus(rho) eps -- SURVEY.md
// Appendix A.1 / A.6.  `mask` (nullable = all ones) marks the parameters that enter the KL / sampled-nkl sums (the reference
// filters those sums on VILinear, quirk Q5: conv layers carry none).  Every kernel reads each [S][P] operand once.
#include "psvi_common.cuh"

namespace {

constexpr int FT = 256;

// theta[s][i] = mu[i] + softplus(rho[i]) eps[s][i]
__global__ void __launch_bounds__(FT) mf_sample_kernel(int S, long long P, const float* __restrict__ mu, const float* __restrict__ rho,
                                                       const float* __restrict__ eps, float* __restrict__ theta) {
  const long long i = (long long)blockIdx.x * FT + threadIdx.x;
  if (i >= P) return;
  const float m = mu[i], sg = softplus_f(rho[i]);
  for (int s = 0; s < S; ++s) theta[(long long)s * P + i] = fmaf(sg, eps[(long long)s * P + i], m);
}

// thetad[s][i] = mud[i] + sigmoid(rho[i]) rhod[i] eps[s][i]
__global__ void __launch_bounds__(FT) mf_tangent_kernel(int S, long long P, const float* __restrict__ rho, const float* __restrict__ mud,
                                                        const float* __restrict__ rhod, const float* __restrict__ eps,
                                                        float* __restrict__ thetad) {
  const long long i = (long long)blockIdx.x * FT + threadIdx.x;
  if (i >= P) return;
  const float m = mud[i], sc = sigmoid_f(rho[i]) * rhod[i];
  for (int s = 0; s < S; ++s) thetad[(long long)s * P + i] = fmaf(sc, eps[(long long)s * P + i], m);
}

// g = [gmu | grho]:  tb = tbar - beta[s] mask theta  (beta / theta nullable: the d nkl_s / d theta path of the outer objective)
//   gmu[i]  = sum_s tb + mask kl_coef mu;   grho[i] = sigmoid(rho) (sum_s tb eps + mask (kl_coef (sg - 1/sg) + nkl_coef / sg))
__global__ void __launch_bounds__(FT) mf_grad_kernel(int S, long long P, const float* __restrict__ mu, const float* __restrict__ rho,
                                                     const float* __restrict__ eps, const float* __restrict__ tbar,
                                                     const float* __restrict__ beta, const float* __restrict__ theta,
                                                     const float* __restrict__ mask, float kl_coef, float nkl_coef,
                                                     float* __restrict__ g) {
  const long long i = (long long)blockIdx.x * FT + threadIdx.x;
  if (i >= P) return;
  const float mk = mask ? mask[i] : 1.f;
  float a = 0.f, b = 0.f;
  for (int s = 0; s < S; ++s) {
    const long long o = (long long)s * P + i;
    float t = tbar[o];
    if (beta) t -= beta[s] * mk * theta[o];
    a += t;
    b = fmaf(t, eps[o], b);
  }
  const float r = rho[i], sg = softplus_f(r), sig = sigmoid_f(r);
  g[i] = a + mk * kl_coef * mu[i];
  g[P + i] = sig * (b + mk * (kl_coef * (sg - 1.f / sg) + nkl_coef / sg));
}

// h = [hmu | hrho]:  hmu = sum_s A_t + mask mud;
//   hrho = sig sum_s A_t eps + sig (1 - sig) rhod sum_s A_td eps + mask [(1 + 1/sg^2) sig^2 + (sg - 1/sg) sig (1 - sig)] rhod
__global__ void __launch_bounds__(FT) mf_hvp_kernel(int S, long long P, const float* __restrict__ rho, const float* __restrict__ mud,
                                                    const float* __restrict__ rhod, const float* __restrict__ eps,
                                                    const float* __restrict__ At, const float* __restrict__ Atd,
                                                    const float* __restrict__ mask, float* __restrict__ h) {
  const long long i = (long long)blockIdx.x * FT + threadIdx.x;
  if (i >= P) return;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int s = 0; s < S; ++s) {
    const long long o = (long long)s * P + i;
    const float e = eps[o], t = At[o];
    a += t;
    b = fmaf(t, e, b);
    c = fmaf(Atd[o], e, c);
  }
  const float mk = mask ? mask[i] : 1.f, r = rho[i], sg = softplus_f(r), sig = sigmoid_f(r), rd = rhod[i];
  h[i] = a + mk * mud[i];
  h[P + i] = sig * b + sig * (1.f - sig) * rd * c + mk * ((1.f + 1.f / (sg * sg)) * sig * sig + (sg - 1.f / sg) * sig * (1.f - sig)) * rd;
}

// part[s][blk] = sum over the block's i of mask (-theta^2/2 + eps^2/2 + log sigma);  klpart[blk] = sum mask (0.5 (sg^2 + mu^2 - 1) - log sg)
constexpr int NB = 128;
__global__ void __launch_bounds__(FT) mf_nkl_kernel(int S, long long P, const float* __restrict__ mu, const float* __restrict__ rho,
                                                    const float* __restrict__ eps, const float* __restrict__ theta,
                                                    const float* __restrict__ mask, double* __restrict__ part,
                                                    double* __restrict__ klpart) {
  const int s = blockIdx.y;   // s == S: the KL row
  double acc = 0;
  for (long long i = (long long)blockIdx.x * FT + threadIdx.x; i < P; i += (long long)NB * FT) {
    const float mk = mask ? mask[i] : 1.f;
    if (mk == 0.f) continue;
    const float sg = softplus_f(rho[i]);
    if (s < S) {
      const float t = theta[(long long)s * P + i], e = eps[(long long)s * P + i];
      acc += (double)(mk * (0.5f * (e * e - t * t) + logf(sg)));
    } else {
      const float m = mu[i];
      acc += (double)(mk * (0.5f * (sg * sg + m * m - 1.f) - logf(sg)));
    }
  }
  __shared__ double red[FT];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = FT / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) (s < S ? part[(size_t)s * NB + blockIdx.x] : klpart[blockIdx.x]) = red[0];
}
__global__ void mf_nkl_finish_kernel(int S, const double* __restrict__ part, const double* __restrict__ klpart, double* __restrict__ out) {
  const int s = threadIdx.x;
  if (s > S) return;
  double a = 0;
  for (int b = 0; b < NB; ++b) a += s < S ? part[(size_t)s * NB + b] : klpart[b];
  out[s] = a;   // out[0..S-1] = nkl_s, out[S] = kl
}

// One step of the unrolled robust Adam (reference psvi/robust_higher/optim.py:303-367), elementwise over the flat
// parameter vector, in the rounding order of the reference's tensor expressions (explicit _rn intrinsics: no FMA contraction):
//   m' = m b1 + (1 - b1) g;  v' = v b2 + (1 - b2) g g;  phi' = phi - k m' / (sqrt(v' + 1e-8) / sq2 + 1e-8)
// with k = lr / (1 - b1^t), sq2 = sqrt(1 - b2^t) computed by the caller in double.
__global__ void __launch_bounds__(FT) adam_step_kernel(long long n, float k, float sq2, const float* __restrict__ phi,
                                                       const float* __restrict__ g, const float* __restrict__ m,
                                                       const float* __restrict__ v, float* __restrict__ phi_out,
                                                       float* __restrict__ m_out, float* __restrict__ v_out) {
  const long long i = (long long)blockIdx.x * FT + threadIdx.x;
  if (i >= n) return;
  const float b1 = 0.9f, b2 = 0.999f, ob1 = (float)(1.0 - 0.9), ob2 = (float)(1.0 - 0.999);
  const float gi = g[i];
  const float mn = __fadd_rn(__fmul_rn(m[i], b1), __fmul_rn(ob1, gi));
  const float vn = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(ob2, gi), gi));
  const float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(__fadd_rn(vn, 1e-8f)), sq2), 1e-8f);
  m_out[i] = mn;
  v_out[i] = vn;
  phi_out[i] = __fsub_rn(phi[i], __fmul_rn(k, __fdiv_rn(mn, den)));
}

// Reverse of that step: from the adjoint pbar of phi' and the running adjoints (mbar, vbar) of (m', v') it forms
//   mb = mbar - k pbar / den;  vb = vbar + (k pbar m' / den^2) / (2 q sq2), zeroed where v' == 0 (the _maybe_mask hook,
//   optim.py:40-52,346-347);  gbar = (1 - b1) mb + 2 (1 - b2) g vb;  mbar <- b1 mb;  vbar <- b2 vb      (q = sqrt(v' + 1e-8))
__global__ void __launch_bounds__(FT) adam_reverse_kernel(long long n, float k, float sq2, const float* __restrict__ pbar,
                                                          const float* __restrict__ g, const float* __restrict__ m_t,
                                                          const float* __restrict__ v_t, float* __restrict__ mbar,
                                                          float* __restrict__ vbar, float* __restrict__ gbar) {
  const long long i = (long long)blockIdx.x * FT + threadIdx.x;
  if (i >= n) return;
  const float b1 = 0.9f, b2 = 0.999f, ob1 = (float)(1.0 - 0.9), tob2 = (float)(2.0 * (1.0 - 0.999));
  const float vt = v_t[i], q = __fsqrt_rn(__fadd_rn(vt, 1e-8f));
  const float den = __fadd_rn(__fdiv_rn(q, sq2), 1e-8f);
  const float kp = __fmul_rn(k, pbar[i]);
  const float mb = __fsub_rn(mbar[i], __fdiv_rn(kp, den));
  float vb = __fadd_rn(vbar[i], __fdiv_rn(__fdiv_rn(__fmul_rn(kp, m_t[i]), __fmul_rn(den, den)), __fmul_rn(__fmul_rn(2.f, q), sq2)));
  if (vt == 0.f) vb = 0.f;
  gbar[i] = __fadd_rn(__fmul_rn(ob1, mb), __fmul_rn(__fmul_rn(tob2, g[i]), vb));
  mbar[i] = __fmul_rn(b1, mb);
  vbar[i] = __fmul_rn(b2, vb);
}

}  // namespace

extern "C" {

int psvi_mf_sample(int32_t S, int64_t P, const float* mu, const float* rho, const float* eps, float* theta, void* stream) {
  PSVI_REQUIRE(S >= 1 && P >= 1 && mu && rho && eps && theta, PSVI_ERR_INVALID, "bad argument");
  mf_sample_kernel<<<(unsigned)((P + FT - 1) / FT), FT, 0, (cudaStream_t)stream>>>(S, P, mu, rho, eps, theta);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_mf_tangent(int32_t S, int64_t P, const float* rho, const float* mud, const float* rhod, const float* eps, float* thetad,
                    void* stream) {
  PSVI_REQUIRE(S >= 1 && P >= 1 && rho && mud && rhod && eps && thetad, PSVI_ERR_INVALID, "bad argument");
  mf_tangent_kernel<<<(unsigned)((P + FT - 1) / FT), FT, 0, (cudaStream_t)stream>>>(S, P, rho, mud, rhod, eps, thetad);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_mf_reparam_grad(int32_t S, int64_t P, const float* mu, const float* rho, const float* eps, const float* tbar,
                         const float* beta, const float* theta, const float* mask, float kl_coef, float nkl_coef, float* g,
                         void* stream) {
  PSVI_REQUIRE(S >= 1 && P >= 1 && mu && rho && eps && tbar && g, PSVI_ERR_INVALID, "bad argument");
  PSVI_REQUIRE(!beta || theta, PSVI_ERR_INVALID, "beta needs theta");
  mf_grad_kernel<<<(unsigned)((P + FT - 1) / FT), FT, 0, (cudaStream_t)stream>>>(S, P, mu, rho, eps, tbar, beta, theta, mask, kl_coef,
                                                                              nkl_coef, g);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_mf_reparam_hvp(int32_t S, int64_t P, const float* rho, const float* mud, const float* rhod, const float* eps,
                        const float* A_t, const float* A_td, const float* mask, float* h, void* stream) {
  PSVI_REQUIRE(S >= 1 && P >= 1 && rho && mud && rhod && eps && A_t && A_td && h, PSVI_ERR_INVALID, "bad argument");
  mf_hvp_kernel<<<(unsigned)((P + FT - 1) / FT), FT, 0, (cudaStream_t)stream>>>(S, P, rho, mud, rhod, eps, A_t, A_td, mask, h);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

size_t psvi_mf_nkl_scratch_bytes(int32_t S) { return ((size_t)(S + 1) * NB + NB) * sizeof(double); }

int psvi_mf_nkl_kl(int32_t S, int64_t P, const float* mu, const float* rho, const float* eps, const float* theta, const float* mask,
                   double* out, void* scratch, void* stream) {
  PSVI_REQUIRE(S >= 1 && S <= 1023 && P >= 1 && mu && rho && eps && theta && out && scratch, PSVI_ERR_INVALID, "bad argument");
  double* part = static_cast<double*>(scratch);
  double* klpart = part + (size_t)S * NB;
  mf_nkl_kernel<<<dim3(NB, S + 1), FT, 0, (cudaStream_t)stream>>>(S, P, mu, rho, eps, theta, mask, part, klpart);
  mf_nkl_finish_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(S, part, klpart, out);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_adam_unroll_step(int64_t n, float k, float sq2, const float* phi, const float* g, const float* m, const float* v,
                          float* phi_out, float* m_out, float* v_out, void* stream) {
  PSVI_REQUIRE(n >= 1 && phi && g && m && v && phi_out && m_out && v_out, PSVI_ERR_INVALID, "bad argument");
  adam_step_kernel<<<(unsigned)((n + FT - 1) / FT), FT, 0, (cudaStream_t)stream>>>(n, k, sq2, phi, g, m, v, phi_out, m_out, v_out);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_adam_unroll_reverse(int64_t n, float k, float sq2, const float* pbar, const float* g, const float* m_t, const float* v_t,
                             float* mbar, float* vbar, float* gbar, void* stream) {
  PSVI_REQUIRE(n >= 1 && pbar && g && m_t && v_t && mbar && vbar && gbar, PSVI_ERR_INVALID, "bad argument");
  adam_reverse_kernel<<<(unsigned)((n + FT - 1) / FT), FT, 0, (cudaStream_t)stream>>>(n, k, sq2, pbar, g, m_t, v_t, mbar, vbar, gbar);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
