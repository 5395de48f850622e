// psvi_mf_stream.cu -- "medium regime" mean-field MLP path: models whose parameter vector does not fit the
// everything-replicated-in-shared-memory cluster engine (P_pad > ~3k, e.g. the baselines' fn with two 100-unit hidden
// layers, P = 10 602) but whose per-sample weights do.  Parameters stay in global memory (L2-resident), one CTA per MC
// sample keeps only ITS sampled weights + one chunk of activations on chip:
//
//   K1 sample_kernel   theta[S][P] = mu + softplus(rho) * eps                     (elementwise, Philox or external noise)
//   K2 fwdbwd_kernel   per sample: forward -> softmax/NLL -> backward over row chunks; theta_bar[S][P], loss partials
//   K3 update_kernel   g = reparameterisation + analytic-KL gradient; Adam arithmetic (torch / higher / hypergrad)
//   K4 loss_kernel     inner objective value = sum_s partial + KL
//   E  predict_kernel  per chunk of test rows: loop over samples, mixture / mean-logit predictive, NLL, argmax
//
// Covers psvi_mf_unroll (run_mfvi_subset / run_mfvi training, baselines.py:1019-1032,876-890) and psvi_mf_evaluate
// (baselines.py:1035-1043; PSVI.evaluate psvi_classes.py:1031-1108) for such models.  The bilevel step for this regime
// is future work.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "psvi_mf_gemm.cuh"

namespace cg = cooperative_groups;

using namespace psvi_mf;

namespace {

struct SP {
  int L;
  int dims[MAXL + 1];
  int S, R, RC;          // samples, rows, rows per chunk
  int noise_mode;
  unsigned long long seed;
  unsigned domain;
  const float* eps;      // external noise base [slabs][S][P]
  int slab;
  const float* mu;
  const float* rho;
  float* theta;          // [S][P]
  float* tbar;           // [S][P]
  float* loss_part;      // [S]
  const float* x;        // [R][D]
  const int* y;          // [R]
  const float* roww;     // [R] row weights (nullable -> wscale)
  float wscale;
  // generic per-sample pass (psvi_net_pass)
  const float* thetad;   // [S][P] tangent weights (dual mode)
  const float* cwm;      // [S][R] per-sample row weights
  float* nll_out;        // [S][R]
  float* tdbar;          // [S][P]
  float* xbar;           // [S][R][D]
  float* acbar;          // [S][R]
  float* logits_out;     // [S][R][C] (forward mode)
  // likelihood of the per-sample pass: 0 = categorical (softmax over C logits, int labels), 2 = Bernoulli on ONE logit (labels
  // 0. / 1. as float bits; sparse-BBVI, reference psvi/inference/utils.py:85-141), 1 = Gaussian with precision tau on
  // ONE output (the regressors, reference psvi_classes.py:1986,2034-2057): labels are float bits in y, nll = tau/2 (o - y)^2 +
  // 1/2 log(2 pi / tau); ybar [S][R] (nullable) receives d(sum_r cw nll)/dy -- or its directional derivative in the dual pass
  int like;
  float tau;
  float* ybar;
  // predictive
  int n_rows, row0, eval_mode;
  const float* lw;       // [S] log importance weights (mode 0)
  float* part;           // [ctas][4]
};

struct SLay {
  int theta, thetad, lab, cw, nll, red, w;
  int act[MAXL + 1], adj[MAXL + 1], actd[MAXL + 1], adjd[MAXL + 1];
  int total;
};

__host__ __device__ inline void make_slay(const SP& p, const Meta& m, SLay& y, bool need_adj, bool dual = false) {
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  y.theta = take(m.Pp);
  y.thetad = dual ? take(m.Pp) : 0;
  y.lab = take(p.RC); y.cw = take(p.RC); y.nll = take(p.RC); y.red = take(64); y.w = take(32);
  for (int l = 0; l <= p.L; ++l) {
    y.act[l] = take(p.RC * m.lda[l]);
    y.adj[l] = (need_adj || l == p.L) ? take(p.RC * m.lda[l]) : 0;
    y.actd[l] = dual ? take(p.RC * m.lda[l]) : 0;
    y.adjd[l] = dual ? take(p.RC * m.lda[l]) : 0;
  }
  y.total = o;
}

__device__ __forceinline__ float noise_at(const SP& p, int slab, int s, int q, int Pt) {
  if (p.noise_mode == PSVI_NOISE_PHILOX) {
    float e4[4];
    philox_normal4(p.seed, p.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)(q >> 2), e4);
    return e4[q & 3];
  }
  return __ldg(p.eps + ((size_t)slab * p.S + s) * Pt + q);
}

// ---- K1 ------------------------------------------------------------------------------------------------------------
__global__ void sample_kernel(const SP p, int Pt) {
  const long long total = (long long)p.S * Pt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i / Pt), q = (int)(i - (long long)s * Pt);
    p.theta[i] = p.mu[q] + softplus_f(p.rho[q]) * noise_at(p, p.slab, s, q, Pt);
  }
}

// ---- per-sample on-chip machinery -----------------------------------------------------------------------------------
struct Worker {
  const SP& p;
  const Meta& mt;
  const SLay& ly;
  float* sm;
  int tid;
  __device__ Worker(const SP& p_, const Meta& m_, const SLay& l_, float* s_) : p(p_), mt(m_), ly(l_), sm(s_), tid(threadIdx.x) {}
  __device__ __forceinline__ float* F(int off) const { return sm + off; }
  __device__ __forceinline__ int* I(int off) const { return reinterpret_cast<int*>(sm + off); }

  // sampled weights of sample s: TL (global) -> padded rows [W[o][0..din-1], b[o], pad]
  __device__ void load_theta(int s) { load_weights(p.theta + (size_t)s * mt.Pt, ly.theta); }
  __device__ void load_weights(const float* src, int dst_off) {
    for (int l = 1; l <= p.L; ++l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l], n = dout * din;
      float* W = F(dst_off) + mt.woff[l];
      const float* g = src + mt.tlw[l];
      // eight loads in flight per thread before the first store (a load -> store loop pays one memory latency per element)
      for (int base = 0; base < n; base += 8 * NT) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * NT + tid;
          v[u] = i < n ? __ldg(g + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = base + u * NT + tid;
          if (i < n) {
            const int oo = i / din;
            W[oo * ldw + (i - oo * din)] = v[u];
          }
        }
      }
      for (int o = tid; o < dout; o += NT) W[o * ldw + din] = __ldg(src + mt.tlb[l] + o);
    }
    __syncthreads();
  }
  __device__ void init_ones() {
    for (int i = tid; i < ly.total; i += NT) sm[i] = 0.f;
    __syncthreads();
    for (int l = 0; l < p.L; ++l)
      for (int rr = tid; rr < p.RC; rr += NT) F(ly.act[l])[rr * mt.lda[l] + p.dims[l]] = 1.f;
    __syncthreads();
  }
  __device__ void stage(const float* x, const int* y, int r0, int nr) {
    const int D = p.dims[0], ld0 = mt.lda[0];
    float* a0 = F(ly.act[0]);
    const float* src = x + (size_t)r0 * D;
    int rr = 0, c = tid;
    while (c >= D) { c -= D; ++rr; }
    const int sr = NT / D, sc = NT - sr * D;
    for (int i = tid; i < nr * D; i += NT) {
      a0[rr * ld0 + c] = __ldg(src + i);
      rr += sr; c += sc;
      if (c >= D) { c -= D; ++rr; }
    }
    for (int r2 = tid; r2 < nr; r2 += NT) I(ly.lab)[r2] = __ldg(y + r0 + r2);
    __syncthreads();
  }
  __device__ void forward(int nr) {
    for (int l = 1; l <= p.L; ++l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l], ldi = mt.lda[l - 1], ldo = mt.lda[l];
      const float* in = F(ly.act[l - 1]);
      float* out = F(ly.act[l]);
      const bool relu = l < p.L;
      GemmOp op{in, ldi, 1, F(ly.theta) + mt.woff[l], ldw, 1, nullptr, nullptr};
      small_gemm(nr, dout, din + 1, op, [&](int rr, int oo, float acc) { out[rr * ldo + oo] = relu ? fmaxf(acc, 0.f) : acc; });
      __syncthreads();
    }
  }
  // NLL per row; grad != 0: also the output adjoint cw * (softmax - onehot)
  __device__ void loss(int nr, bool grad, float* yb_out) {
    const int C = p.dims[p.L], ld = mt.lda[p.L];
    const float* o = F(ly.act[p.L]);
    float* ao = F(ly.adj[p.L]);
    if (p.like == 2) {       // -Bernoulli(logits = o).log_prob(y) = softplus(o) - y o
      for (int rr = tid; rr < nr; rr += NT) {
        const float ov = o[rr * ld], yv = __int_as_float(I(ly.lab)[rr]);
        F(ly.nll)[rr] = softplus_f(ov) - yv * ov;
        if (grad) ao[rr * ld] = F(ly.cw)[rr] * (sigmoid_f(ov) - yv);
      }
      __syncthreads();
      return;
    }
    if (p.like == 1) {
      const float hl = 0.5f * logf(6.283185307179586f / p.tau);
      for (int rr = tid; rr < nr; rr += NT) {
        const float r = o[rr * ld] - __int_as_float(I(ly.lab)[rr]);
        F(ly.nll)[rr] = 0.5f * p.tau * r * r + hl;
        if (grad) {
          const float g = F(ly.cw)[rr] * p.tau * r;
          ao[rr * ld] = g;
          if (yb_out) yb_out[rr] = -g;
        }
      }
      __syncthreads();
      return;
    }
    for (int rr = tid; rr < nr; rr += NT) {
      const float* row = o + rr * ld;
      float mx = row[0];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
      const float lse = mx + logf(se);
      const int y = I(ly.lab)[rr];
      F(ly.nll)[rr] = lse - row[y];
      if (grad) {
        const float w = F(ly.cw)[rr];
        for (int c = 0; c < C; ++c) ao[rr * ld + c] = w * (expf(row[c] - lse) - (c == y ? 1.f : 0.f));
      }
    }
    __syncthreads();
  }
  // backward; weight adjoints go straight to global theta_bar (same thread owns an element across chunks)
  __device__ void backward(int nr, float* tbar, bool first_chunk) {
    for (int l = p.L; l >= 1; --l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l], ldi = mt.lda[l - 1], ldo = mt.lda[l];
      const float* in = F(ly.act[l - 1]);
      const float* A = F(ly.adj[l]);
      float* gw = tbar + mt.tlw[l];
      float* gb = tbar + mt.tlb[l];
      GemmOp ow{A, 1, ldo, in, 1, ldi, nullptr, nullptr};
      small_gemm(dout, din + 1, nr, ow, [&](int oo, int ii, float acc) {
        float* dst = ii < din ? gw + oo * din + ii : gb + oo;
        *dst = first_chunk ? acc : *dst + acc;
      });
      if (l > 1) {
        float* Ai = F(ly.adj[l - 1]);
        GemmOp ox{A, ldo, 1, F(ly.theta) + mt.woff[l], 1, ldw, nullptr, nullptr};
        small_gemm(nr, din, dout, ox, [&](int rr, int ii, float acc) { Ai[rr * ldi + ii] = (in[rr * ldi + ii] > 0.f) ? acc : 0.f; });
      }
      __syncthreads();
    }
  }
  // ---- dual (primal + tangent) machinery: SURVEY Appendix A.6, same structure as the cluster engine ----------------
  __device__ void forward_dual(int nr) {
    for (int l = 1; l <= p.L; ++l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l], ldi = mt.lda[l - 1], ldo = mt.lda[l];
      const float* in = F(ly.act[l - 1]);
      const float* ind = F(ly.actd[l - 1]);
      const float* W = F(ly.theta) + mt.woff[l];
      const float* Wd = F(ly.thetad) + mt.woff[l];
      float* out = F(ly.act[l]);
      float* outd = F(ly.actd[l]);
      const bool relu = l < p.L;
      GemmOp op{in, ldi, 1, W, ldw, 1, nullptr, nullptr};
      small_gemm(nr, dout, din + 1, op, [&](int rr, int oo, float acc) { out[rr * ldo + oo] = relu ? fmaxf(acc, 0.f) : acc; });
      GemmOp opd{in, ldi, 1, Wd, ldw, 1, l > 1 ? ind : nullptr, l > 1 ? W : nullptr};
      // (masked by the primal ReLU: the same thread wrote out[rr][oo] -- small_gemm maps outputs to threads by shape only)
      small_gemm(nr, dout, din + 1, opd, [&](int rr, int oo, float acc) {
        outd[rr * ldo + oo] = (relu && !(out[rr * ldo + oo] > 0.f)) ? 0.f : acc;
      });
      __syncthreads();
    }
  }
  __device__ void loss_dual(int nr, float* ac_out, float* yb_out) {
    const int C = p.dims[p.L], ld = mt.lda[p.L];
    const float* o = F(ly.act[p.L]);
    const float* od = F(ly.actd[p.L]);
    float* ao = F(ly.adj[p.L]);
    float* aod = F(ly.adjd[p.L]);
    if (p.like == 2) {
      for (int rr = tid; rr < nr; rr += NT) {
        const float ov = o[rr * ld], yv = __int_as_float(I(ly.lab)[rr]), d = od[rr * ld], w = F(ly.cw)[rr];
        const float sg = sigmoid_f(ov);
        F(ly.nll)[rr] = softplus_f(ov) - yv * ov;
        aod[rr * ld] = w * (sg - yv);
        ao[rr * ld] = w * sg * (1.f - sg) * d;
        if (ac_out) ac_out[rr] = (sg - yv) * d;
      }
      __syncthreads();
      return;
    }
    if (p.like == 1) {
      const float hl = 0.5f * logf(6.283185307179586f / p.tau);
      for (int rr = tid; rr < nr; rr += NT) {
        const float r = o[rr * ld] - __int_as_float(I(ly.lab)[rr]), d = od[rr * ld], w = F(ly.cw)[rr];
        F(ly.nll)[rr] = 0.5f * p.tau * r * r + hl;
        aod[rr * ld] = w * p.tau * r;          // adjoint of the tangent output
        ao[rr * ld] = w * p.tau * d;           // second-order term: d/d eps of (tau (o - y))
        if (ac_out) ac_out[rr] = p.tau * r * d;
        if (yb_out) yb_out[rr] = -w * p.tau * d;
      }
      __syncthreads();
      return;
    }
    for (int rr = tid; rr < nr; rr += NT) {
      const float* row = o + rr * ld;
      const float* rowd = od + rr * ld;
      float mx = row[0];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
      const float lse = mx + logf(se);
      const int y = I(ly.lab)[rr];
      F(ly.nll)[rr] = lse - row[y];
      const float w = F(ly.cw)[rr];
      float pd = 0.f;
      for (int c = 0; c < C; ++c) pd += expf(row[c] - lse) * rowd[c];
      float ac = 0.f;
      for (int c = 0; c < C; ++c) {
        const float pc = expf(row[c] - lse), qc = pc - (c == y ? 1.f : 0.f);
        aod[rr * ld + c] = w * qc;
        ao[rr * ld + c] = w * pc * (rowd[c] - pd);
        ac += qc * rowd[c];
      }
      if (ac_out) ac_out[rr] = ac;
    }
    __syncthreads();
  }
  // generic backward: dual=false -> tbar; dual=true -> tbar = A_theta, tdbar = A_thetadot; xbar rows (nullable)
  __device__ void backward_any(int nr, bool dual, float* tbar, float* tdbar, float* xbar, bool first_chunk) {
    for (int l = p.L; l >= 1; --l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l], ldi = mt.lda[l - 1], ldo = mt.lda[l];
      const float* in = F(ly.act[l - 1]);
      const float* ind = F(ly.actd[l - 1]);
      const float* A = F(ly.adj[l]);
      const float* Ad = F(ly.adjd[l]);
      const float* W = F(ly.theta) + mt.woff[l];
      const float* Wd = F(ly.thetad) + mt.woff[l];
      const bool has_ind = dual && l > 1;
      const int tlw = mt.tlw[l], tlb = mt.tlb[l];
      if (dual) {   // A_theta = A^T in + Ad^T ind  and  A_thetadot = Ad^T in  in one pass over the shared operands
        GemmOp ow{A, 1, ldo, in, 1, ldi, Ad, has_ind ? ind : nullptr};
        small_gemm_pair(dout, din + 1, nr, ow, [&](int oo, int ii, float acc, float accd) {
          const int off = ii < din ? tlw + oo * din + ii : tlb + oo;
          tbar[off] = first_chunk ? acc : tbar[off] + acc;
          tdbar[off] = first_chunk ? accd : tdbar[off] + accd;
        });
      } else {
        GemmOp ow{A, 1, ldo, in, 1, ldi, nullptr, nullptr};
        small_gemm(dout, din + 1, nr, ow, [&](int oo, int ii, float acc) {
          float* dst = tbar + (ii < din ? tlw + oo * din + ii : tlb + oo);
          *dst = first_chunk ? acc : *dst + acc;
        });
      }
      if (l > 1 || xbar) {
        GemmOp ox{A, ldo, 1, W, 1, ldw, dual ? Ad : nullptr, dual ? Wd : nullptr};
        if (l > 1) {
          float* Ai = F(ly.adj[l - 1]);
          float* Aid = F(ly.adjd[l - 1]);
          small_gemm(nr, din, dout, ox, [&](int rr, int ii, float acc) { Ai[rr * ldi + ii] = (in[rr * ldi + ii] > 0.f) ? acc : 0.f; });
          if (dual) {
            GemmOp oxd{Ad, ldo, 1, W, 1, ldw, nullptr, nullptr};
            small_gemm(nr, din, dout, oxd, [&](int rr, int ii, float acc) { Aid[rr * ldi + ii] = (in[rr * ldi + ii] > 0.f) ? acc : 0.f; });
          }
        } else {
          small_gemm(nr, din, dout, ox, [&](int rr, int ii, float acc) { xbar[rr * din + ii] = acc; });
        }
      }
      __syncthreads();
    }
  }
};

// ---- K2 ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) fwdbwd_kernel(const __grid_constant__ SP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ SLay ly;
  if (threadIdx.x == 0) { make_meta(p.dims, p.L, mt); make_slay(p, mt, ly, true); }
  __syncthreads();
  Worker w(p, mt, ly, smem_dyn);
  const int s = blockIdx.x;
  w.init_ones();
  w.load_theta(s);
  float part = 0.f;
  for (int r0 = 0; r0 < p.R; r0 += p.RC) {
    const int nr = min(p.RC, p.R - r0);
    w.stage(p.x, p.y, r0, nr);
    for (int rr = threadIdx.x; rr < nr; rr += NT) w.F(ly.cw)[rr] = p.roww ? __ldg(p.roww + r0 + rr) : p.wscale;
    __syncthreads();
    w.forward(nr);
    w.loss(nr, true, nullptr);
    for (int rr = threadIdx.x; rr < nr; rr += NT) part += w.F(ly.cw)[rr] * w.F(ly.nll)[rr];
    w.backward(nr, p.tbar + (size_t)s * mt.Pt, r0 == 0);
  }
  part = block_sum(part, w.F(ly.red));
  if (threadIdx.x == 0) p.loss_part[s] = part;
}


// ---- generic per-sample pass on externally supplied weights (any variational family) ---------------------------------
// mode 0: forward only (nll); 1: gradient; 2: dual (Hessian-vector) pass
// Grid (Z, S): the rows of a sample are split over a CLUSTER of Z CTAs (Z = 1: one CTA per sample, plain launch).  With Z > 1 a
// CTA accumulates its weight adjoints in shared memory behind the activations; after a cluster barrier the Z CTAs sum the
// partials through distributed shared memory in FIXED rank order (deterministic) -- each CTA reduces every Z-th block of indices
// -- and write global memory once.  (One CTA per sample left 32 of 148 SMs busy at cfg3: S = 32.)
__global__ void __launch_bounds__(NT, 1) net_pass_kernel(const __grid_constant__ SP p, int mode) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ SLay ly;
  if (threadIdx.x == 0) { make_meta(p.dims, p.L, mt); make_slay(p, mt, ly, mode > 0, mode == 2); }
  __syncthreads();
  Worker w(p, mt, ly, smem_dyn);
  const int Z = (int)gridDim.x, z = (int)blockIdx.x, s = (int)blockIdx.y, tid = threadIdx.x;
  const int per = (p.R + Z - 1) / Z, rz0 = z * per, rz1 = min(p.R, rz0 + per);
  const int Ptp = (mt.Pt + 3) & ~3;
  const bool split = Z > 1 && mode > 0;
  float* tb = mode > 0 ? (split ? smem_dyn + ly.total : p.tbar + (size_t)s * mt.Pt) : nullptr;
  float* tdb = mode == 2 ? (split ? smem_dyn + ly.total + Ptp : p.tdbar + (size_t)s * mt.Pt) : nullptr;
  if (split && rz0 >= rz1) {   // a CTA without rows contributes zeros
    for (int i = tid; i < (mode == 2 ? 2 : 1) * Ptp; i += NT) tb[i] = 0.f;
  }
  w.init_ones();
  w.load_theta(s);
  if (mode == 2) { w.load_weights(p.thetad + (size_t)s * mt.Pt, ly.thetad); }
  const int D = p.dims[0];
  for (int r0 = rz0; r0 < rz1; r0 += p.RC) {
    const int nr = min(p.RC, rz1 - r0);
    w.stage(p.x, p.y, r0, nr);
    if (mode > 0) {
      for (int rr = tid; rr < nr; rr += NT) w.F(ly.cw)[rr] = __ldg(p.cwm + (size_t)s * p.R + r0 + rr);
      __syncthreads();
    }
    float* xb = p.xbar ? p.xbar + ((size_t)s * p.R + r0) * D : nullptr;
    if (mode == 2) {
      w.forward_dual(nr);
      w.loss_dual(nr, p.acbar ? p.acbar + (size_t)s * p.R + r0 : nullptr, p.ybar ? p.ybar + (size_t)s * p.R + r0 : nullptr);
      w.backward_any(nr, true, tb, tdb, xb, r0 == rz0);
    } else {
      w.forward(nr);
      w.loss(nr, mode == 1, p.ybar ? p.ybar + (size_t)s * p.R + r0 : nullptr);
      if (mode == 1) w.backward_any(nr, false, tb, nullptr, xb, r0 == rz0);
    }
    if (p.nll_out)
      for (int rr = tid; rr < nr; rr += NT) p.nll_out[(size_t)s * p.R + r0 + rr] = w.F(ly.nll)[rr];
    if (p.logits_out) {
      const int C = p.dims[p.L], ldc = mt.lda[p.L];
      for (int rr = 0; rr < nr; ++rr)
        for (int c = tid; c < C; c += NT) p.logits_out[((size_t)s * p.R + r0 + rr) * C + c] = w.F(ly.act[p.L])[rr * ldc + c];
    }
    __syncthreads();
  }
  if (split) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    const int nvec = mode == 2 ? 2 : 1;
    for (int v = 0; v < nvec; ++v) {
      float* local = smem_dyn + ly.total + v * Ptp;
      float* out = (v ? p.tdbar : p.tbar) + (size_t)s * mt.Pt;
      for (int i = tid + NT * z; i < mt.Pt; i += NT * Z) {
        float a = 0.f;
        for (int k = 0; k < Z; ++k) a += cluster.map_shared_rank(local, k)[i];
        out[i] = a;
      }
    }
    cluster.sync();   // nobody leaves while its partials may still be read
  }
}

// ---- K3 ------------------------------------------------------------------------------------------------------------
struct UP {
  int S, Pt, adam_mode, step;  // step is 1-based
  float lr;
  float* mu;
  float* rho;
  float* am;
  float* av;
  const float* tbar;
  float* kl_part;  // [gridDim.x]
};
__global__ void update_kernel(const SP p, const UP u) {
  const double B1 = 0.9, B2 = 0.999;
  const float b1 = (float)B1, b2 = (float)B2, omb1 = (float)(1.0 - B1), omb2 = (float)(1.0 - B2), aeps = 1e-8f;
  const double b1t = pow(B1, (double)u.step), b2t = pow(B2, (double)u.step);
  const float bc1 = (float)(1.0 - b1t), bc2 = (float)(1.0 - b2t), sq2 = (float)sqrt(1.0 - b2t);
  const float step_size = u.lr / bc1;
  float kl = 0.f;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < u.Pt; q += gridDim.x * blockDim.x) {
    const float mu = u.mu[q], rho = u.rho[q];
    const float sg = softplus_f(rho), sgm = sigmoid_f(rho);
    kl += 0.5f * (sg * sg + mu * mu - 1.f) - logf(sg);
    float s1 = 0.f, s2 = 0.f;
    for (int s = 0; s < u.S; ++s) {
      const float tb = u.tbar[(size_t)s * u.Pt + q];
      s1 += tb;
      s2 += tb * noise_at(p, p.slab, s, q, u.Pt);
    }
    const float gg[2] = {s1 + mu, sgm * (s2 + (sg - 1.f / sg))};
    const float pv[2] = {mu, rho};
    float np[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float g = gg[c];
      const float m = u.am[c * u.Pt + q] * b1 + omb1 * g;
      float v = u.av[c * u.Pt + q] * b2 + omb2 * g * g;
      if (u.adam_mode == PSVI_ADAM_ROBUST_HIGHER) np[c] = pv[c] - step_size * (m / (sqrtf(v + 1e-8f) / sq2 + aeps));
      else if (u.adam_mode == PSVI_ADAM_TORCH) np[c] = pv[c] - step_size * (m / (sqrtf(v) / sq2 + aeps));
      else { v += 1e-12f; np[c] = pv[c] - u.lr * (m / bc1 / (sqrtf(v / bc2) + aeps)); }
      u.am[c * u.Pt + q] = m;
      u.av[c * u.Pt + q] = v;
    }
    u.mu[q] = np[0];
    u.rho[q] = np[1];
  }
  __shared__ float red[32];
  kl = warp_sum(kl);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    u.kl_part[blockIdx.x] = t;
  }
}
__global__ void loss_kernel(const float* loss_part, int S, const float* kl_part, int nkl, float* out) {
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < S; ++i) t += loss_part[i];
    for (int i = 0; i < nkl; ++i) t += kl_part[i];
    *out = t;
  }
}

// ---- log importance weights (mode 0): forward over the pseudo rows, one CTA per sample ------------------------------
__global__ void __launch_bounds__(NT, 1) logweight_kernel(const __grid_constant__ SP p, float* lw_out) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ SLay ly;
  if (threadIdx.x == 0) { make_meta(p.dims, p.L, mt); make_slay(p, mt, ly, false); }
  __syncthreads();
  Worker w(p, mt, ly, smem_dyn);
  const int s = blockIdx.x;
  w.init_ones();
  w.load_theta(s);
  float ps = 0.f, nkl = 0.f;
  for (int q = threadIdx.x; q < mt.Pt; q += NT) {
    const float th = p.theta[(size_t)s * mt.Pt + q], e = noise_at(p, p.slab, s, q, mt.Pt);
    nkl += -0.5f * th * th + 0.5f * e * e + logf(softplus_f(p.rho[q]));
  }
  for (int r0 = 0; r0 < p.R; r0 += p.RC) {
    const int nr = min(p.RC, p.R - r0);
    w.stage(p.x, p.y, r0, nr);
    w.forward(nr);
    w.loss(nr, false, nullptr);
    for (int rr = threadIdx.x; rr < nr; rr += NT) ps += __ldg(p.roww + r0 + rr) * w.F(ly.nll)[rr];
    __syncthreads();
  }
  ps = block_sum(ps, w.F(ly.red));
  nkl = block_sum(nkl, w.F(ly.red));
  if (threadIdx.x == 0) lw_out[s] = ps + nkl;  // reference sign quirk Q3
}

// ---- E --------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1) predict_kernel(const __grid_constant__ SP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ SLay ly;
  if (threadIdx.x == 0) { make_meta(p.dims, p.L, mt); make_slay(p, mt, ly, false); }
  __syncthreads();
  Worker w(p, mt, ly, smem_dyn);
  const int tid = threadIdx.x;
  w.init_ones();
  if (tid == 0) {
    float mx = -INFINITY, se = 0.f;
    if (p.eval_mode == 0) {
      for (int s = 0; s < p.S; ++s) mx = fmaxf(mx, p.lw[s]);
      for (int s = 0; s < p.S; ++s) se += expf(p.lw[s] - mx);
    }
    for (int s = 0; s < p.S && s < 32; ++s) w.F(ly.w)[s] = p.eval_mode == 0 ? expf(p.lw[s] - mx) / se : 1.f / (float)p.S;
  }
  const int r0 = p.row0 + blockIdx.x * p.RC;
  const int nr = min(p.RC, p.row0 + p.n_rows - r0);
  const int C = p.dims[p.L], ld = mt.lda[p.L];
  float nll_sum = 0.f, correct = 0.f;
  if (nr > 0) {
    w.stage(p.x, p.y, r0, nr);
    float* probs = w.F(ly.adj[p.L]);
    for (int s = 0; s < p.S; ++s) {
      w.load_theta(s);
      w.forward(nr);
      const float wgt = w.F(ly.w)[s];
      const float* o = w.F(ly.act[p.L]);
      for (int rr = tid; rr < nr; rr += NT) {
        const float* row = o + rr * ld;
        if (p.eval_mode == 2) {
          for (int c = 0; c < C; ++c) probs[rr * ld + c] += wgt * row[c];
        } else {
          float mx = row[0];
          for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
          float se = 0.f;
          for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
          const float inv = wgt / se;
          for (int c = 0; c < C; ++c) probs[rr * ld + c] += inv * expf(row[c] - mx);
        }
      }
      __syncthreads();
    }
    for (int rr = tid; rr < nr; rr += NT) {
      const float* pr = probs + rr * ld;
      const int y = w.I(ly.lab)[rr];
      int am = 0;
      float best = pr[0];
      for (int c = 1; c < C; ++c)
        if (pr[c] > best) { best = pr[c]; am = c; }
      correct += (am == y) ? 1.f : 0.f;
      if (p.eval_mode == 2) {
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(pr[c] - best);
        nll_sum += best + logf(se) - pr[y];
      } else {
        float tot = 0.f;
        for (int c = 0; c < C; ++c) tot += pr[c];
        nll_sum -= logf(fminf(fmaxf(pr[y] / tot, 1.1920929e-07f), 1.f - 1.1920929e-07f));
      }
    }
  }
  nll_sum = block_sum(nll_sum, w.F(ly.red));
  correct = block_sum(correct, w.F(ly.red));
  if (tid == 0) {
    float* o = p.part + (size_t)blockIdx.x * 4;
    o[0] = nll_sum; o[1] = correct; o[2] = (float)(nr > 0 ? nr : 0); o[3] = 0.f;
  }
}

// accumulate one slab's partials into out[0..2]; diagnostics of the slab into out[3..4] (the last slab wins, Q12)
__global__ void accumulate_kernel(const float* part, int n, float* out, const float* lw, int S) {
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int i = 0; i < n; ++i) { a += part[4 * i]; b += part[4 * i + 1]; c += part[4 * i + 2]; }
    out[0] += (float)a; out[1] += (float)b; out[2] += (float)c;
    if (lw) {
      float mx = -INFINITY, se = 0.f, ent = 0.f, sw = 0.f, sw2 = 0.f;
      for (int s = 0; s < S; ++s) mx = fmaxf(mx, lw[s]);
      for (int s = 0; s < S; ++s) se += expf(lw[s] - mx);
      for (int s = 0; s < S; ++s) {
        const float w = expf(lw[s] - mx) / se;
        if (w > 0.f) ent -= logf(w) * w;
        sw += w; sw2 += w * w;
      }
      out[3] = ent;
      out[4] = sw * sw / sw2 / (float)S;
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------
int fit_rows(SP& p, const Meta& mt, int want, bool need_adj, size_t* smem_out) {
  int dev = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t budget = (size_t)smem_max - 2048;
  SLay ly;
  int lo = 0, hi = want < 1 ? 1 : want;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    p.RC = mid;
    make_slay(p, mt, ly, need_adj);
    if ((size_t)ly.total * 4 <= budget) lo = mid; else hi = mid - 1;
  }
  PSVI_REQUIRE(lo >= 1, PSVI_ERR_UNSUPPORTED,
               "model too large even for the streaming path: one sample's weights (P_pad=%d floats) plus one row of "
               "activations must fit in %zu B of shared memory", mt.Pp, budget);
  p.RC = lo;
  make_slay(p, mt, ly, need_adj);
  *smem_out = (size_t)ly.total * 4;
  return PSVI_OK;
}

void fill_sp(SP& p, const psvi_mf_model* model, const psvi_noise* noise) {
  memset(&p, 0, sizeof(p));
  p.L = model->n_layers;
  for (int l = 0; l <= p.L; ++l) p.dims[l] = model->dims[l];
  p.S = model->mc_samples;
  p.noise_mode = noise->mode; p.seed = noise->seed; p.domain = noise->domain; p.eps = noise->eps;
}

constexpr int KL_BLOCKS = 64;

}  // namespace

extern "C" {

size_t psvi_mf_stream_workspace_bytes(const psvi_mf_model* model, int32_t n_rows) {
  if (!model || model->n_layers < 1 || model->n_layers > MAXL) return 0;
  Meta mt;
  make_meta(model->dims, model->n_layers, mt);
  const size_t S = model->mc_samples;
  // theta [S][P], tbar [S][P], loss_part [S], kl_part, lw [S], partials [(n_rows + 3) / 4 + 1][4] (>= 4 rows per chunk worst case)
  return (2 * S * mt.Pt + 2 * S + KL_BLOCKS + 64 + 4 * ((size_t)(n_rows > 0 ? n_rows : 0) + 4)) * sizeof(float);
}

int psvi_mf_unroll_stream(const psvi_mf_model* model, const psvi_noise* noise, float* mu, float* rho, float* adam_m,
                          float* adam_v, int32_t step0, const float* x, const int32_t* y, const float* row_weights,
                          float row_weight_scalar, int32_t M, int32_t T, float lr, int32_t adam_mode, float* losses,
                          void* workspace, void* stream_) {
  PSVI_REQUIRE(model && noise && mu && rho && adam_m && adam_v && x && y && workspace, PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(model->n_layers >= 1 && model->n_layers <= MAXL && M > 0 && T >= 0 && step0 >= 0, PSVI_ERR_INVALID, "bad argument");
  PSVI_REQUIRE(adam_mode >= 0 && adam_mode <= 2, PSVI_ERR_INVALID, "unknown adam_mode %d", adam_mode);
  cudaStream_t stream = (cudaStream_t)stream_;
  Meta mt;
  make_meta(model->dims, model->n_layers, mt);
  SP p;
  fill_sp(p, model, noise);
  float* ws = static_cast<float*>(workspace);
  p.theta = ws; p.tbar = ws + (size_t)p.S * mt.Pt; p.loss_part = p.tbar + (size_t)p.S * mt.Pt;
  float* kl_part = p.loss_part + p.S;
  p.mu = mu; p.rho = rho; p.x = x; p.y = y; p.roww = row_weights; p.wscale = row_weight_scalar; p.R = M;
  size_t smem = 0;
  int rc = fit_rows(p, mt, M < 256 ? M : 256, true, &smem);
  if (rc) return rc;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(fwdbwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  UP u;
  u.S = p.S; u.Pt = mt.Pt; u.adam_mode = adam_mode; u.lr = lr; u.mu = mu; u.rho = rho; u.am = adam_m; u.av = adam_v;
  u.tbar = p.tbar; u.kl_part = kl_part;
  const int eb = (int)(((long long)p.S * mt.Pt + 255) / 256);
  for (int t = 0; t < T; ++t) {
    p.slab = t;
    sample_kernel<<<eb < 1184 ? eb : 1184, 256, 0, stream>>>(p, mt.Pt);
    fwdbwd_kernel<<<p.S, NT, smem, stream>>>(p);
    u.step = step0 + t + 1;
    update_kernel<<<KL_BLOCKS, 256, 0, stream>>>(p, u);
    if (losses) loss_kernel<<<1, 32, 0, stream>>>(p.loss_part, p.S, kl_part, KL_BLOCKS, losses + t);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_mf_evaluate_stream(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                            const float* u, const int32_t* z, const float* a_weights, int32_t M, const float* xt,
                            const int32_t* yt, int32_t n_rows, int32_t batch, int32_t first_slab, int32_t mode,
                            float* out, void* workspace, void* stream_) {
  PSVI_REQUIRE(model && noise && mu && rho && xt && yt && out && workspace, PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(n_rows > 0 && batch > 0 && first_slab >= 0 && mode >= 0 && mode <= 2, PSVI_ERR_INVALID, "bad argument");
  PSVI_REQUIRE(mode != 0 || (u && z && a_weights && M > 0), PSVI_ERR_INVALID, "importance-weighted mode needs pseudo-data and a = N f(v)");
  PSVI_REQUIRE(model->mc_samples <= 32, PSVI_ERR_UNSUPPORTED, "streaming predictive kernel supports S <= 32");
  cudaStream_t stream = (cudaStream_t)stream_;
  Meta mt;
  make_meta(model->dims, model->n_layers, mt);
  SP p;
  fill_sp(p, model, noise);
  float* ws = static_cast<float*>(workspace);
  p.theta = ws; p.tbar = ws + (size_t)p.S * mt.Pt; p.loss_part = p.tbar + (size_t)p.S * mt.Pt;
  float* lw = p.loss_part + p.S + KL_BLOCKS;
  p.part = lw + p.S + 32;
  p.mu = mu; p.rho = rho;
  PSVI_CUDA_CHECK(cudaMemsetAsync(out, 0, 8 * sizeof(float), stream));
  const int n_slabs = (n_rows + batch - 1) / batch;
  const int eb = (int)(((long long)p.S * mt.Pt + 255) / 256);
  for (int k = 0; k < n_slabs; ++k) {
    p.slab = first_slab + k;
    sample_kernel<<<eb < 1184 ? eb : 1184, 256, 0, stream>>>(p, mt.Pt);
    size_t smem = 0;
    if (mode == 0) {
      p.x = u; p.y = z; p.roww = a_weights; p.R = M;
      int rc = fit_rows(p, mt, M < 256 ? M : 256, false, &smem);
      if (rc) return rc;
      PSVI_CUDA_CHECK(cudaFuncSetAttribute(logweight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      logweight_kernel<<<p.S, NT, smem, stream>>>(p, lw);
    }
    p.x = xt; p.y = yt; p.row0 = k * batch; p.n_rows = (k + 1) * batch <= n_rows ? batch : n_rows - k * batch;
    p.eval_mode = mode; p.lw = lw;
    int rc = fit_rows(p, mt, p.n_rows < 128 ? p.n_rows : 128, false, &smem);
    if (rc) return rc;
    PSVI_REQUIRE(p.RC >= 4 || p.RC >= p.n_rows, PSVI_ERR_UNSUPPORTED, "fewer than 4 rows per chunk fit (P_pad=%d)", mt.Pp);
    const int ctas = (p.n_rows + p.RC - 1) / p.RC;
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    predict_kernel<<<ctas, NT, smem, stream>>>(p);
    accumulate_kernel<<<1, 32, 0, stream>>>(p.part, ctas, out, mode == 0 ? lw : nullptr, p.S);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

static int net_pass_impl(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const int32_t* y,
                         const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                         float* logits, int like, float tau, float* ybar, void* stream_);

int psvi_net_pass(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const int32_t* y,
                  const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                  float* logits, void* stream_) {
  return net_pass_impl(model, theta, thetad, x, y, cw, R, nll, tbar, tdbar, xbar, acbar, logits, 0, 0.f, nullptr, stream_);
}

int psvi_net_pass_gaussian(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const float* y,
                           const float* cw, int32_t R, float tau, float* nll, float* tbar, float* tdbar, float* xbar,
                           float* acbar, float* ybar, float* outputs, void* stream_) {
  PSVI_REQUIRE(model && model->dims[model->n_layers] == 1, PSVI_ERR_UNSUPPORTED, "the Gaussian likelihood is built for ONE output");
  PSVI_REQUIRE(tau > 0.f, PSVI_ERR_INVALID, "tau must be positive");
  return net_pass_impl(model, theta, thetad, x, reinterpret_cast<const int32_t*>(y), cw, R, nll, tbar, tdbar, xbar, acbar,
                       outputs, 1, tau, ybar, stream_);
}

int psvi_net_pass_bernoulli(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const float* y,
                            const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                            float* outputs, void* stream_) {
  PSVI_REQUIRE(model && model->dims[model->n_layers] == 1, PSVI_ERR_UNSUPPORTED, "the Bernoulli likelihood takes ONE logit");
  return net_pass_impl(model, theta, thetad, x, reinterpret_cast<const int32_t*>(y), cw, R, nll, tbar, tdbar, xbar, acbar,
                       outputs, 2, 0.f, nullptr, stream_);
}

static int net_pass_impl(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const int32_t* y,
                         const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                         float* logits, int like, float tau, float* ybar, void* stream_) {
  PSVI_REQUIRE(model && theta && x && y && R > 0, PSVI_ERR_INVALID, "null pointer or R<=0");
  PSVI_REQUIRE(model->n_layers >= 1 && model->n_layers <= MAXL, PSVI_ERR_INVALID, "bad n_layers");
  const int mode = thetad ? 2 : (tbar ? 1 : 0);
  PSVI_REQUIRE(mode == 0 || (cw && tbar), PSVI_ERR_INVALID, "gradient / dual pass needs row weights and tbar");
  PSVI_REQUIRE(mode != 2 || tdbar, PSVI_ERR_INVALID, "dual pass needs tdbar");
  PSVI_REQUIRE(mode != 0 || nll || logits, PSVI_ERR_INVALID, "forward pass needs nll or logits");
  Meta mt;
  make_meta(model->dims, model->n_layers, mt);
  SP p;
  psvi_noise dummy;
  memset(&dummy, 0, sizeof(dummy));
  fill_sp(p, model, &dummy);
  p.theta = const_cast<float*>(theta); p.thetad = thetad; p.x = x; p.y = y; p.cwm = cw; p.R = R;
  p.nll_out = nll; p.tbar = tbar; p.tdbar = tdbar; p.xbar = xbar; p.acbar = acbar; p.logits_out = logits;
  p.like = like; p.tau = tau; p.ybar = ybar;
  int dev = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t budget = (size_t)smem_max - 2048;
  // rows of a sample over a cluster of Z CTAs: >= 16 rows per CTA (cfg3, R = 100, S = 32: Z = 3 / 4 / 6 / 8 -> 4.7 / 4.6 / 4.3 /
  // 4.3 ms per outer step), at most 8 (portable cluster size), at most ~2 CTAs per SM in total; PSVI_NET_PASS_Z overrides
  // (1 = one CTA per sample)
  int sms = 1;
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int Z = R / 16;
  if (Z > 8) Z = 8;
  while (Z > 1 && p.S * Z > 2 * sms) --Z;
  if (const char* e = getenv("PSVI_NET_PASS_Z")) Z = atoi(e);
  if (Z < 1) Z = 1;
  if (Z > 8) Z = 8;
  SLay ly;
  size_t extra = 0;
  int lo = 0;
  for (;; --Z) {   // (falls back towards Z = 1 if the partial-adjoint buffers do not fit beside one row of activations)
    const int per = (R + Z - 1) / Z;
    extra = (Z > 1 && mode > 0) ? (size_t)(mode == 2 ? 2 : 1) * ((mt.Pt + 3) & ~3) : 0;
    int hi = per < 256 ? per : 256;
    lo = 0;
    while (lo < hi) {
      const int mid = (lo + hi + 1) / 2;
      p.RC = mid;
      make_slay(p, mt, ly, mode > 0, mode == 2);
      if (((size_t)ly.total + extra) * 4 <= budget) lo = mid; else hi = mid - 1;
    }
    if (lo >= 1 || Z == 1) break;
  }
  PSVI_REQUIRE(lo >= 1, PSVI_ERR_UNSUPPORTED, "one sample's weights (P_pad=%d floats%s) plus one row of activations must "
               "fit in %zu B of shared memory", mt.Pp, mode == 2 ? ", twice" : "", budget);
  p.RC = lo;
  make_slay(p, mt, ly, mode > 0, mode == 2);
  const size_t smem = ((size_t)ly.total + extra) * 4;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(net_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (Z > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)Z, (unsigned)p.S);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream_;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)Z;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PSVI_CUDA_CHECK(cudaLaunchKernelEx(&cfg, net_pass_kernel, p, mode));
  } else {
    net_pass_kernel<<<dim3(1, p.S), NT, smem, (cudaStream_t)stream_>>>(p, mode);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_net_predict(const psvi_mf_model* model, const float* theta, const float* log_weights, int32_t mode,
                     const float* xt, const int32_t* yt, int32_t n_rows, float* out, void* workspace, void* stream_) {
  PSVI_REQUIRE(model && theta && xt && yt && out && workspace && n_rows > 0, PSVI_ERR_INVALID, "null pointer or n_rows<=0");
  PSVI_REQUIRE(mode >= 0 && mode <= 2 && (mode != 0 || log_weights), PSVI_ERR_INVALID, "bad mode / missing log weights");
  PSVI_REQUIRE(model->mc_samples <= 32, PSVI_ERR_UNSUPPORTED, "streaming predictive kernel supports S <= 32");
  cudaStream_t stream = (cudaStream_t)stream_;
  Meta mt;
  make_meta(model->dims, model->n_layers, mt);
  SP p;
  psvi_noise dummy;
  memset(&dummy, 0, sizeof(dummy));
  fill_sp(p, model, &dummy);
  p.theta = const_cast<float*>(theta); p.x = xt; p.y = yt; p.row0 = 0; p.n_rows = n_rows; p.eval_mode = mode;
  p.lw = log_weights; p.part = static_cast<float*>(workspace);
  size_t smem = 0;
  int rc = fit_rows(p, mt, n_rows < 128 ? n_rows : 128, false, &smem);
  if (rc) return rc;
  PSVI_REQUIRE(p.RC >= 4 || p.RC >= n_rows, PSVI_ERR_UNSUPPORTED, "fewer than 4 rows per chunk fit (P_pad=%d)", mt.Pp);
  const int ctas = (n_rows + p.RC - 1) / p.RC;
  PSVI_CUDA_CHECK(cudaMemsetAsync(out, 0, 8 * sizeof(float), stream));
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  predict_kernel<<<ctas, NT, smem, stream>>>(p);
  accumulate_kernel<<<1, 32, 0, stream>>>(p.part, ctas, out, mode == 0 ? log_weights : nullptr, p.S);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
