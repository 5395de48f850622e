// psvi_mf_engine.cu -- the "small regime" PSVI engine for mean-field MLP BNNs (logistic_regression, fn).
//
// ONE thread-block cluster (G <= 16 CTAs, one per group of MC samples) runs a whole bilevel PSVI step on-chip:
//   T differentiable-Adam inner steps on inner_elbo  ->  outer psvi_elbo forward/backward  ->  reverse sweep through
//   the T steps with hand-written Hessian-/mixed-vector products  ->  hypergradients on (u, v).
// What the reference does with ~25k ATen launches and a double-backward autograd graph per outer step
// (psvi/inference/psvi_classes.py:541-600, psvi/robust_higher/optim.py:152-257,299-367) is a single launch here.
//
// Parallel decomposition (B200: 148 SMs, clusters with distributed shared memory):
//   * CTA `rank` owns MC samples s = rank, rank+G, ... : sampled weights, activations and adjoints of a sample never
//     leave that CTA's shared memory.
//   * the variational parameter vector (padded layout, Pp floats) is cut in G slices; CTA k is the *owner* of slice k:
//     it holds the Adam moments / reverse-sweep carries of its slice, receives every CTA's per-sample weight-gradient
//     partials through DSMEM pushes, reduces them in a fixed order (deterministic), updates, and pushes the new
//     parameters (or the next HVP direction) back into every CTA's shared memory.  Two cluster barriers per step.
//   * the trajectory (phi_t, g_t, m_t, v_t) goes to global memory, written and later re-read by the same owner thread.
// Rows (pseudo-points u, then the minibatch) are processed in chunks of RC rows so that any M / B fits.
//
// Math: SURVEY.md Appendix A (A.1 layer, A.2 objectives, A.4 Adam VJP, A.6 HVP) == oracle/psvi_oracle.py.
#include <cooperative_groups.h>
#include <math.h>

#include "psvi_mf_engine.cuh"

namespace cg = cooperative_groups;
using namespace psvi_mf;

namespace {


// shared-memory carve-up (offsets in floats)
struct Lay {
  int mu, rho, sig, sgm, gmu, grho, theta, thetad, eps, acc1, acc2, acc3;  // Pp each (padded parameter layout)
  int tl_of_pad, pad_of_tl;                                                // ints
  int recv;                                                                // [G][3][slice]
  int ost;                                                                 // owner state [10][slice]
  int a, f, ubar, abar, su, sz;                                            // coreset: a[M] f[M] ubar[M*D] abar[M] u[M*ld0] z[M]
  int lw, e, dsv, w, beta, gp;                                             // S each
  int red, lossrecv;                                                       // 64, G
  int lab, cw, nll;                                                        // RC each
  int act[MAXL + 1], actd[MAXL + 1], adj[MAXL + 1], adjd[MAXL + 1];
  int total;
};

__host__ __device__ inline void make_layout(const EP& p, const Meta& m, Lay& y) {
  const bool dual = (p.flags & (F_REVERSE | F_HVP)) != 0;
  int o = 0;
  auto take = [&](int n) {
    int r = o;
    o += (n + 3) & ~3;
    return r;
  };
  const int Pp = m.Pp;
  const bool ev = (p.flags & F_EVAL) != 0;  // forward-only kernels do not need adjoint / owner storage
  y.mu = take(Pp); y.rho = take(Pp); y.sig = take(Pp); y.sgm = take(Pp);
  y.gmu = take(ev ? 0 : Pp); y.grho = take(ev ? 0 : Pp);
  y.theta = take(Pp); y.thetad = take(ev ? 0 : Pp); y.eps = take(Pp);
  y.acc1 = take(ev ? 0 : Pp); y.acc2 = take(ev ? 0 : Pp); y.acc3 = take(ev ? 0 : Pp);
  y.tl_of_pad = take(Pp); y.pad_of_tl = take(m.Pt);
  y.recv = take(ev ? 0 : p.G * 3 * p.slice);
  y.ost = take(ev ? 0 : 10 * p.slice);
  y.a = take(p.M); y.f = take(p.M); y.ubar = take(ev ? 0 : p.M * p.dims[0]); y.abar = take(ev ? 0 : p.M);
  y.su = take(p.M * m.lda[0]); y.sz = take(p.M);
  // lw, e hold doubles
  y.lw = take(2 * p.S); y.e = take(2 * p.S); y.dsv = take(p.S); y.w = take(p.S); y.beta = take(p.S); y.gp = take(p.S);
  y.red = take(64); y.lossrecv = take(p.G);
  y.lab = take(p.RC); y.cw = take(p.RC); y.nll = take(p.RC);
  for (int l = 0; l <= p.L; ++l) {
    y.act[l] = take(p.RC * m.lda[l]);
    y.adj[l] = take(p.RC * m.lda[l]);
    y.actd[l] = dual ? take(p.RC * m.lda[l]) : 0;
    y.adjd[l] = dual ? take(p.RC * m.lda[l]) : 0;
  }
  y.total = o;
}

// ----------------------------------------------------------------------------------------------------------------
struct Engine {
  const EP& p;
  const Meta& mt;
  const Lay& ly;
  float* sm;
  cg::cluster_group cluster;
  int rank, tid;

  __device__ Engine(const EP& p_, const Meta& m_, const Lay& l_, float* s_)
      : p(p_), mt(m_), ly(l_), sm(s_), cluster(cg::this_cluster()) {
    rank = (int)cluster.block_rank();
    tid = threadIdx.x;
  }
  __device__ __forceinline__ float* F(int off) const { return sm + off; }
  __device__ __forceinline__ int* I(int off) const { return reinterpret_cast<int*>(sm + off); }
  __device__ __forceinline__ float* remote(int off, int r) { return cluster.map_shared_rank(sm + off, r); }

  // ---- refresh sigma = softplus(rho), sgm = sigmoid(rho) after phi changed -----------------------------------
  __device__ void refresh_sigma() {
    const int* top = I(ly.tl_of_pad);
    for (int pp = tid; pp < mt.Pp; pp += NT) {
      if (top[pp] >= 0) {
        const float r = F(ly.rho)[pp];
        F(ly.sig)[pp] = softplus_f(r);
        F(ly.sgm)[pp] = sigmoid_f(r);
      }
    }
    __syncthreads();
  }

  // ---- draw eps for (slab, s) and build theta (and the tangent thetad); returns this thread's partial of
  //      sampled_nkl_s = sum_i [-theta^2/2 + eps^2/2 + log sigma]  (neural_net.py:110-115)
  __device__ float sample_theta(int s, int slab, bool tangent, float fold_beta, bool want_nkl = false) {
    const int* pot = I(ly.pad_of_tl);
    float nkl = 0.f;
    const int Pt = mt.Pt;
    const int n4 = (Pt + 3) >> 2;
    for (int q4 = tid; q4 < n4; q4 += NT) {
      float e4[4];
      if (p.noise_mode == PSVI_NOISE_PHILOX) {
        philox_normal4(p.seed, p.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)q4, e4);
      } else {
        const float* src = p.eps + ((size_t)slab * p.S + s) * Pt;
#pragma unroll
        for (int j = 0; j < 4; ++j) e4[j] = (4 * q4 + j < Pt) ? __ldg(src + 4 * q4 + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = 4 * q4 + j;
        if (q < Pt) {
          const int pp = pot[q];
          const float e = e4[j], sg = F(ly.sig)[pp];
          const float th = F(ly.mu)[pp] + sg * e;
          F(ly.eps)[pp] = e;
          F(ly.theta)[pp] = th;
          if (tangent) F(ly.thetad)[pp] = F(ly.gmu)[pp] + F(ly.sgm)[pp] * F(ly.grho)[pp] * e;
          if (want_nkl) nkl += -0.5f * th * th + 0.5f * e * e + logf(sg);
          if (fold_beta != 0.f) {  // outer objective: d nkl_s / d theta = -theta, weighted by beta_s (A.2)
            const float tb = -fold_beta * th;
            F(ly.acc1)[pp] += tb;
            F(ly.acc2)[pp] += tb * e;
          }
        }
      }
    }
    __syncthreads();
    return nkl;
  }

  // ---- stage rows [r0, r0+nr) of the virtual row list (pseudo rows 0..M-1, then minibatch rows) -------------------
  // returns the pointer to the layer-0 activations of the chunk
  __device__ float* stage_rows(int r0, int nr) {
    int* lab = I(ly.lab);
    if (r0 < p.M) {  // pseudo chunk (never mixed with data rows): alias the resident copy of u
      for (int rr = tid; rr < nr; rr += NT) lab[rr] = I(ly.sz)[r0 + rr];
      return F(ly.su) + r0 * mt.lda[0];
    }
    const int D = p.dims[0], ld0 = mt.lda[0];
    float* a0 = F(ly.act[0]);
    const float* src = p.xb + (size_t)(r0 - p.M) * D;
    if (D <= 4) {  // tiny rows: one thread per row
      for (int rr = tid; rr < nr; rr += NT) {
        for (int c = 0; c < D; ++c) a0[rr * ld0 + c] = __ldg(src + rr * D + c);
        a0[rr * ld0 + D] = 1.f;
      }
    } else {       // coalesced: consecutive threads read consecutive floats of the [nr][D] block
      int rr = 0, c = tid;
      while (c >= D) { c -= D; ++rr; }
      const int step_r = NT / D, step_c = NT - step_r * D;
      for (int i = tid; i < nr * D; i += NT) {
        a0[rr * ld0 + c] = __ldg(src + i);
        rr += step_r; c += step_c;
        if (c >= D) { c -= D; ++rr; }
      }
      for (int r2 = tid; r2 < nr; r2 += NT) a0[r2 * ld0 + D] = 1.f;
    }
    if (p.yb != nullptr)   // (the module-level forward has no labels)
      for (int rr = tid; rr < nr; rr += NT) lab[rr] = __ldg(p.yb + (r0 - p.M) + rr);
    return a0;
  }

  // ---- forward through the MLP for one chunk (primal, optionally tangent) -----------------------------------------
  // bias = the weight column that multiplies the constant-one slot of the layer input (K = din + 1)
  __device__ void forward(const float* a0, int nr, bool dual) {
    for (int l = 1; l <= p.L; ++l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l];
      const int ldi = mt.lda[l - 1], ldo = mt.lda[l];
      const float* in = (l == 1) ? a0 : F(ly.act[l - 1]);
      const float* ind = F(ly.actd[l - 1]);
      const float* W = F(ly.theta) + mt.woff[l];
      const float* Wd = F(ly.thetad) + mt.woff[l];
      float* out = F(ly.act[l]);
      float* outd = F(ly.actd[l]);
      const bool relu = l < p.L;
      GemmOp op{in, ldi, 1, W, ldw, 1, nullptr, nullptr};
      small_gemm(nr, dout, din + 1, op,
                 [&](int rr, int oo, float acc) { out[rr * ldo + oo] = relu ? fmaxf(acc, 0.f) : acc; });
      if (dual) {
        // tangent pre-activation  x Wd^T (+ xd W^T for l > 1; xd of the input layer is zero), masked by the primal ReLU:
        // small_gemm maps (row, column) to threads by (nrows, ncols, K) only, so out[rr][oo] was written by this thread
        GemmOp opd{in, ldi, 1, Wd, ldw, 1, l > 1 ? ind : nullptr, l > 1 ? W : nullptr};
        small_gemm(nr, dout, din + 1, opd, [&](int rr, int oo, float acc) {
          outd[rr * ldo + oo] = (relu && !(out[rr * ldo + oo] > 0.f)) ? 0.f : acc;
        });
      }
      __syncthreads();
    }
  }

  // ---- softmax / NLL per row; fills the output adjoints.  mode 0: value only, 1: gradient, 2: HVP ------------------
  // cw[rr] must hold the row weights.  nll[rr] receives -log softmax(o)[label].  For mode 2, abar_rows (nullable)
  // accumulates A_c[r] = sum_c q_c * odot_c  (A.6) into abar[r0 + rr].
  __device__ void loss_stage(int nr, int mode, int r0) {
    const int C = p.dims[p.L], ld = mt.lda[p.L];
    const float* o = F(ly.act[p.L]);
    const float* od = F(ly.actd[p.L]);
    float* ao = F(ly.adj[p.L]);
    float* aod = F(ly.adjd[p.L]);
    const int* lab = I(ly.lab);
    const float* cw = F(ly.cw);
    for (int rr = tid; rr < nr; rr += NT) {
      const float* row = o + rr * ld;
      float mx = row[0];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, row[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += expf(row[c] - mx);
      const float lse = mx + logf(se);
      const int y = lab[rr];
      F(ly.nll)[rr] = lse - row[y];
      if (mode == 0) continue;
      const float w = cw[rr];
      if (mode == 1) {
        for (int c = 0; c < C; ++c) {
          const float pc = expf(row[c] - lse);
          ao[rr * ld + c] = w * (pc - (c == y ? 1.f : 0.f));
        }
      } else {
        const float* rowd = od + rr * ld;
        float pd = 0.f;
        for (int c = 0; c < C; ++c) pd += expf(row[c] - lse) * rowd[c];
        float ac = 0.f;
        for (int c = 0; c < C; ++c) {
          const float pc = expf(row[c] - lse);
          const float qc = pc - (c == y ? 1.f : 0.f);
          aod[rr * ld + c] = w * qc;                  // adjoint of odot
          ao[rr * ld + c] = w * pc * (rowd[c] - pd);  // adjoint of o
          ac += qc * rowd[c];
        }
        F(ly.abar)[r0 + rr] += ac;  // rows of an HVP pass are always pseudo rows
      }
    }
    __syncthreads();
  }

  // ---- backward for one chunk.  dual=false: plain gradient (acc1 += tbar, acc2 += tbar*eps);
  //      dual=true: adjoint of Ldot (acc1 += A_theta, acc2 += A_theta*eps, acc3 += A_thetadot*eps).
  //      need_x: also produce the adjoint of the layer-0 input and add it to ubar[r0 + rr]. ---------------------------
  __device__ void backward(const float* a0, int nr, bool dual, bool need_x, int r0) {
    for (int l = p.L; l >= 1; --l) {
      const int din = mt.din[l], dout = mt.dout[l], ldw = mt.ldw[l];
      const int ldi = mt.lda[l - 1], ldo = mt.lda[l];
      const float* in = (l == 1) ? a0 : F(ly.act[l - 1]);
      const float* ind = F(ly.actd[l - 1]);
      const float* A = F(ly.adj[l]);
      const float* Ad = F(ly.adjd[l]);
      const float* W = F(ly.theta) + mt.woff[l];
      const float* Wd = F(ly.thetad) + mt.woff[l];
      const float* eps = F(ly.eps) + mt.woff[l];
      float* acc1 = F(ly.acc1) + mt.woff[l];
      float* acc2 = F(ly.acc2) + mt.woff[l];
      float* acc3 = F(ly.acc3) + mt.woff[l];
      const bool has_ind = dual && l > 1;
      // -- weight + bias adjoints: C(o, i) = sum_r A[r][o] * in[r][i], i <= din (in[r][din] == 1)
      if (dual) {   // A_theta = A^T in + Ad^T ind  and  A_thetadot = Ad^T in  in one pass over the shared operands
        GemmOp ow{A, 1, ldo, in, 1, ldi, Ad, has_ind ? ind : nullptr};
        small_gemm_pair(dout, din + 1, nr, ow, [&](int oo, int ii, float acc, float accd) {
          const int pp = oo * ldw + ii;
          const float e = eps[pp];
          acc1[pp] += acc;
          acc2[pp] += acc * e;
          acc3[pp] += accd * e;
        });
      } else {
        GemmOp ow{A, 1, ldo, in, 1, ldi, nullptr, nullptr};
        small_gemm(dout, din + 1, nr, ow, [&](int oo, int ii, float acc) {
          const int pp = oo * ldw + ii;
          acc1[pp] += acc;
          acc2[pp] += acc * eps[pp];
        });
      }
      // -- input adjoints: C(r, i) = sum_o A[r][o] * W[o][i]  (+ Ad[r][o] * Wd[o][i])
      if (l > 1 || need_x) {
        float* Ai = F(ly.adj[l - 1]);
        float* Aid = F(ly.adjd[l - 1]);
        GemmOp ox{A, ldo, 1, W, 1, ldw, dual ? Ad : nullptr, dual ? Wd : nullptr};
        if (l > 1) {
          small_gemm(nr, din, dout, ox,
                     [&](int rr, int ii, float acc) { Ai[rr * ldi + ii] = (in[rr * ldi + ii] > 0.f) ? acc : 0.f; });
          if (dual) {
            GemmOp oxd{Ad, ldo, 1, W, 1, ldw, nullptr, nullptr};
            small_gemm(nr, din, dout, oxd,
                       [&](int rr, int ii, float acc) { Aid[rr * ldi + ii] = (in[rr * ldi + ii] > 0.f) ? acc : 0.f; });
          }
        } else {
          float* ub = F(ly.ubar) + r0 * din;
          small_gemm(nr, din, dout, ox, [&](int rr, int ii, float acc) { ub[rr * din + ii] += acc; });
        }
      }
      __syncthreads();
    }
  }

  // ---- push this CTA's accumulators to the slice owners, then clear them ------------------------------------------
  __device__ void push_acc(int ncomp) {
    const int slice = p.slice;
    for (int owner = 0; owner < p.G; ++owner) {
      float* r = remote(ly.recv, owner) + (size_t)rank * 3 * slice;
      const int base = owner * slice;
      const int n = min(slice, mt.Pp - base);
      for (int j = tid; j < n; j += NT) {
        const int pp = base + j;
        r[j] = F(ly.acc1)[pp];
        r[slice + j] = F(ly.acc2)[pp];
        if (ncomp > 2) r[2 * slice + j] = F(ly.acc3)[pp];
        F(ly.acc1)[pp] = 0.f;
        F(ly.acc2)[pp] = 0.f;
        if (ncomp > 2) F(ly.acc3)[pp] = 0.f;
      }
    }
  }
  __device__ __forceinline__ float recv_sum(int comp, int j) const {
    float s = 0.f;
    const float* r = sm + ly.recv + comp * p.slice + j;
    for (int c = 0; c < p.G; ++c) s += r[(size_t)c * 3 * p.slice];
    return s;
  }
  // owner writes a value of its slice element into the same padded slot of every CTA
  __device__ __forceinline__ void bcast(int off, int pp, float val) {
    for (int c = 0; c < p.G; ++c) remote(off, c)[pp] = val;
  }

  // ---- coreset weights a = N f(v)  (psvi_classes.py:476,505; f per class :111,:1358,:1486) ------------------------
  __device__ void setup_coreset() {
    float* a = F(ly.a);
    float* f = F(ly.f);
    const int M = p.M;
    if (p.roww) {
      for (int m = tid; m < M; m += NT) { a[m] = __ldg(p.roww + m); f[m] = 0.f; }
      __syncthreads();
      return;
    }
    if (p.vmode == PSVI_VMODE_IDENTITY) {
      for (int m = tid; m < M; m += NT) { f[m] = __ldg(p.v + m); a[m] = p.Nf * f[m]; }
      __syncthreads();
      return;
    }
    float mx = -INFINITY;
    for (int m = tid; m < M; m += NT) mx = fmaxf(mx, __ldg(p.v + m));
    mx = block_max(mx, F(ly.red));
    float se = 0.f;
    for (int m = tid; m < M; m += NT) se += expf(__ldg(p.v + m) - mx);
    se = block_sum(se, F(ly.red));
    const float sc = p.Nf * (p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(p.alpha) : 1.f);
    for (int m = tid; m < M; m += NT) {
      f[m] = expf(__ldg(p.v + m) - mx) / se;
      a[m] = sc * f[m];
    }
    __syncthreads();
  }

  // ---- one gradient pass of the inner objective over the pseudo rows for sample s; returns sum_m a_m nll[s,m] -------
  __device__ float inner_pass(int s, int slab, bool dual) {
    sample_theta(s, slab, dual, 0.f);
    float part = 0.f;
    for (int r0 = 0; r0 < p.M; r0 += p.RC) {
      const int nr = min(p.RC, p.M - r0);
      const float* a0 = stage_rows(r0, nr);
      for (int rr = tid; rr < nr; rr += NT) F(ly.cw)[rr] = F(ly.a)[r0 + rr];
      __syncthreads();
      forward(a0, nr, dual);
      loss_stage(nr, dual ? 2 : 1, r0);
      for (int rr = tid; rr < nr; rr += NT) part += F(ly.cw)[rr] * F(ly.nll)[rr];
      backward(a0, nr, dual, dual, r0);
    }
    return part;
  }

  __device__ void init();
  __device__ void run();
  __device__ void eval_weights();
  __device__ void eval_rows();
  __device__ void forward_rows();
};

// ----------------------------------------------------------------------------------------------------------------
// ---- init: zero shared memory, index maps, parameters, pseudo-data, coreset weights ----------------------------------
__device__ void Engine::init() {
  const int Pp = mt.Pp, Pt = mt.Pt;
  int* top = I(ly.tl_of_pad);
  int* pot = I(ly.pad_of_tl);
  for (int i = tid; i < ly.total; i += NT) sm[i] = 0.f;
  __syncthreads();
  for (int pp = tid; pp < Pp; pp += NT) top[pp] = -1;
  __syncthreads();
  for (int q = tid; q < Pt; q += NT) {
    int l = 1;
    while (l < p.L && q >= mt.tlw[l + 1]) ++l;
    int pp;
    if (q < mt.tlb[l]) {
      const int r = q - mt.tlw[l], oo = r / mt.din[l], ii = r - oo * mt.din[l];
      pp = mt.woff[l] + oo * mt.ldw[l] + ii;
    } else {
      pp = mt.boff[l] + (q - mt.tlb[l]) * mt.ldw[l];
    }
    pot[q] = pp;
    top[pp] = q;
    F(ly.mu)[pp] = p.mu[q];
    F(ly.rho)[pp] = p.rho[q];
  }
  {
    const int D = p.dims[0], ld0 = mt.lda[0];
    for (int i = tid; i < p.M * D; i += NT) {
      const int m = i / D, c = i - m * D;
      F(ly.su)[m * ld0 + c] = __ldg(p.u + i);
    }
    for (int m = tid; m < p.M; m += NT) {
      I(ly.sz)[m] = __ldg(p.z + m);
      F(ly.su)[m * ld0 + D] = 1.f;  // constant-one slot (multiplies the bias column)
    }
    for (int l = 1; l < p.L; ++l)
      for (int rr = tid; rr < p.RC; rr += NT) F(ly.act[l])[rr * mt.lda[l] + p.dims[l]] = 1.f;
  }
  __syncthreads();
  refresh_sigma();
  if (p.M > 0) setup_coreset();
}

__device__ void Engine::run() {
  const int slice = p.slice, G = p.G, Pp = mt.Pp, Pt = mt.Pt;
  const int j0 = rank * slice;  // first padded index of my slice
  int* top = I(ly.tl_of_pad);
  int* pot = I(ly.pad_of_tl);
  float* ost = F(ly.ost);
  // owner state rows: 0 pbar_mu 1 pbar_rho 2 mbar_mu 3 mbar_rho 4 vbar_mu 5 vbar_rho 6 am_mu 7 am_rho 8 av_mu 9 av_rho
  auto OST = [&](int row, int j) -> float& { return ost[row * slice + j]; };
  init();
  if (p.adam_m && (p.flags & F_UNROLL)) {
    for (int j = tid; j < slice; j += NT) {
      const int pp = j0 + j;
      const int q = pp < Pp ? top[pp] : -1;
      if (q >= 0) {
        OST(6, j) = p.adam_m[q];
        OST(7, j) = p.adam_m[Pt + q];
        OST(8, j) = p.adam_v[q];
        OST(9, j) = p.adam_v[Pt + q];
      }
    }
  }
  cluster.sync();  // every CTA's shared memory is initialised before anybody pushes into it

  // torch.optim.Adam defaults (psvi_classes.py:861).  The reference mixes Python doubles and fp32 tensors: betas and
  // (1 - beta) reach the tensors as fp32 roundings of the double values; bias corrections are double arithmetic.
  const double B1 = 0.9, B2 = 0.999;
  const float b1 = (float)B1, b2 = (float)B2, omb1 = (float)(1.0 - B1), omb2 = (float)(1.0 - B2), aeps = 1e-8f;
  const bool want_loss = p.inner_losses != nullptr;

  // =================================================================================================================
  // Phase U: T inner Adam steps  (psvi_classes.py:549-555 ; optim.py:224-229,303-367)
  // =================================================================================================================
  if (p.flags & F_UNROLL) {
    double b1t = pow(B1, (double)p.step0), b2t = pow(B2, (double)p.step0);
    for (int t = 0; t < p.T; ++t) {
      float lpart = 0.f;
      for (int s = rank; s < p.S; s += G) lpart += inner_pass(s, t, false);
      push_acc(2);
      if (want_loss) {
        // KL(q||p) of my slice at phi_t (neural_net.py:101-108), added once (Q1)
        for (int j = tid; j < slice; j += NT) {
          const int pp = j0 + j;
          if (pp < Pp && top[pp] >= 0) {
            const float sg = F(ly.sig)[pp], m = F(ly.mu)[pp];
            lpart += 0.5f * (sg * sg + m * m - 1.f) - logf(sg);
          }
        }
        lpart = block_sum(lpart, F(ly.red));
        if (tid == 0) remote(ly.lossrecv, 0)[rank] = lpart;
      }
      cluster.sync();
      if (want_loss && rank == 0 && tid == 0) {
        float s = 0.f;
        for (int c = 0; c < G; ++c) s += F(ly.lossrecv)[c];
        p.inner_losses[t] = s;
      }
      // ---- owner: gradient of my slice, Adam, trajectory, broadcast ----
      b1t *= B1;
      b2t *= B2;
      const float bc1 = (float)(1.0 - b1t);
      const float sq2 = (float)sqrt(1.0 - b2t);
      const float bc2 = (float)(1.0 - b2t);
      const float step_size = p.lr / bc1;
      for (int j = tid; j < slice; j += NT) {
        const int pp = j0 + j;
        const int q = pp < Pp ? top[pp] : -1;
        if (q < 0) continue;
        const float mu = F(ly.mu)[pp], rho = F(ly.rho)[pp], sg = F(ly.sig)[pp], sgm = F(ly.sgm)[pp];
        const float g_mu = recv_sum(0, j) + mu;
        const float g_rho = sgm * (recv_sum(1, j) + (sg - 1.f / sg));
        float nm[2], nv[2], np[2];
        const float gg[2] = {g_mu, g_rho}, pv[2] = {mu, rho};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float g = gg[c];
          float m = OST(6 + c, j) * b1 + omb1 * g;
          float v = OST(8 + c, j) * b2 + omb2 * g * g;
          float den;
          if (p.adam_mode == PSVI_ADAM_ROBUST_HIGHER) {
            den = sqrtf(v + 1e-8f) / sq2 + aeps;                     // optim.py:346-363
            np[c] = pv[c] - step_size * (m / den);
          } else if (p.adam_mode == PSVI_ADAM_TORCH) {
            den = sqrtf(v) / sq2 + aeps;
            np[c] = pv[c] - step_size * (m / den);
          } else {                                                    // hypergrad/diff_optimizers.py:184-213
            v += 1e-12f;
            den = sqrtf(v / bc2) + aeps;
            np[c] = pv[c] - p.lr * (m / bc1 / den);
          }
          nm[c] = m;
          nv[c] = v;
        }
        if (p.g_out) {
          p.g_out[q] = g_mu;
          p.g_out[Pt + q] = g_rho;
        }
        if (p.traj) {
          float* tr = p.traj + (size_t)t * 8 * Pt;
          tr[q] = mu; tr[Pt + q] = rho;
          tr[2 * Pt + q] = g_mu; tr[3 * Pt + q] = g_rho;
          tr[4 * Pt + q] = nm[0]; tr[5 * Pt + q] = nm[1];
          tr[6 * Pt + q] = nv[0]; tr[7 * Pt + q] = nv[1];
        }
        if (!(p.flags & F_NOUPDATE)) {
          OST(6, j) = nm[0]; OST(7, j) = nm[1]; OST(8, j) = nv[0]; OST(9, j) = nv[1];
          bcast(ly.mu, pp, np[0]);
          bcast(ly.rho, pp, np[1]);
        }
      }
      cluster.sync();
      refresh_sigma();
    }
    if (p.flags & F_WRITE_PHI) {
      for (int j = tid; j < slice; j += NT) {
        const int pp = j0 + j;
        const int q = pp < Pp ? top[pp] : -1;
        if (q < 0) continue;
        p.mu[q] = F(ly.mu)[pp];
        p.rho[q] = F(ly.rho)[pp];
        if (p.adam_m) {
          p.adam_m[q] = OST(6, j); p.adam_m[Pt + q] = OST(7, j);
          p.adam_v[q] = OST(8, j); p.adam_v[Pt + q] = OST(9, j);
        }
      }
    }
  }

  // =================================================================================================================
  // Phase O: outer objective psvi_elbo and its gradient at phi_T   (psvi_classes.py:445-486, A.2)
  // =================================================================================================================
  if (p.flags & F_OUTER) {
    const int slab = p.T;
    const int R = p.M + p.B;
    const float dscale = p.Nf / (float)p.Btot;
    // O1: per-sample p_s, d_s, nkl_s
    for (int s = rank; s < p.S; s += G) {
      float nkl = sample_theta(s, slab, false, 0.f, true);
      float ps = 0.f, ds = 0.f;
      for (int r0 = 0; r0 < R;) {
        const int lim = r0 < p.M ? p.M : R;
        const int nr = min(p.RC, lim - r0);
        const float* a0 = stage_rows(r0, nr);
        __syncthreads();
        forward(a0, nr, false);
        loss_stage(nr, 0, r0);
        if (r0 < p.M) {
          for (int rr = tid; rr < nr; rr += NT) ps += F(ly.a)[r0 + rr] * F(ly.nll)[rr];
        } else {
          for (int rr = tid; rr < nr; rr += NT) ds += F(ly.nll)[rr];
        }
        __syncthreads();
        r0 += nr;
      }
      // the S per-sample sums are O(N) while the importance-weight adjoints depend on their *differences*:
      // reduce and keep them in double (the fp32 reference loses ~2 digits here at init_sd=1e-6)
      const double nkl_d = block_sum_d((double)nkl, F(ly.red));
      const double ps_d = block_sum_d((double)ps, F(ly.red));
      const double ds_d = block_sum_d((double)ds, F(ly.red)) * (double)dscale;
      if (tid < G) {
        reinterpret_cast<double*>(remote(ly.lw, tid))[s] = -ps_d + nkl_d;
        reinterpret_cast<double*>(remote(ly.e, tid))[s] = ds_d - (double)p.kappa * ps_d;
        remote(ly.dsv, tid)[s] = (float)ds_d;
      }
    }
    cluster.sync();
    // O2: importance weights and adjoint seeds (every CTA, redundantly; S is tiny)
    if (tid == 0) {
      const int S = p.S;
      const double* lw = reinterpret_cast<const double*>(F(ly.lw));
      const double* ev = reinterpret_cast<const double*>(F(ly.e));
      double mx = -INFINITY;
      for (int s = 0; s < S; ++s) mx = fmax(mx, lw[s]);
      double se = 0.0;
      for (int s = 0; s < S; ++s) se += exp(lw[s] - mx);
      double ebar = 0.0, lwm = 0.0;
      for (int s = 0; s < S; ++s) {
        const double w = exp(lw[s] - mx) / se;
        F(ly.w)[s] = (float)w;
        ebar += w * ev[s];
        lwm += lw[s];
      }
      lwm /= (double)S;
      double bsum = 0.0;
      for (int s = 0; s < S; ++s) {
        const double w = exp(lw[s] - mx) / se;
        const double beta = w * (ev[s] - ebar) - (double)p.kappa / (double)S;  // dLoss/dlw_s
        F(ly.beta)[s] = (float)beta;
        F(ly.gp)[s] = (float)(-(double)p.kappa * w - beta);                    // dLoss/dp_s
        bsum += beta;
      }
      F(ly.red)[60] = (float)bsum;
      const float lossv = (float)(ebar - (double)p.kappa * lwm);
      if (rank == 0) {
        if (p.loss_out) p.loss_out[0] = lossv;
        if (p.flags & F_STORE_GOUT) {
          float* go = p.gout + 2 * Pt + p.M * p.dims[0] + p.M;
          for (int s = 0; s < S; ++s) go[s] = F(ly.dsv)[s];
          go[S] = lossv;
          go[S + 1] = (float)ebar;
          go[S + 2] = (float)lwm;
          go[S + 3] = (float)bsum;
          for (int s = 0; s < S; ++s) {  // diagnostics (rank-local, not meant to be all-reduced)
            go[S + 4 + s] = F(ly.w)[s];
            go[2 * S + 4 + s] = F(ly.beta)[s];
            go[3 * S + 4 + s] = F(ly.gp)[s];
          }
        }
      }
    }
    __syncthreads();
    const float beta_sum = F(ly.red)[60];
    // O3: backward with per-sample row weights
    for (int s = rank; s < p.S; s += G) {
      const float beta = F(ly.beta)[s], gp = F(ly.gp)[s], wd = F(ly.w)[s] * dscale;
      sample_theta(s, slab, false, beta);
      for (int r0 = 0; r0 < R;) {
        const bool pseudo = r0 < p.M;
        const int lim = pseudo ? p.M : R;
        const int nr = min(p.RC, lim - r0);
        const float* a0 = stage_rows(r0, nr);
        for (int rr = tid; rr < nr; rr += NT) F(ly.cw)[rr] = pseudo ? gp * F(ly.a)[r0 + rr] : wd;
        __syncthreads();
        forward(a0, nr, false);
        loss_stage(nr, 1, r0);
        if (pseudo) {
          for (int rr = tid; rr < nr; rr += NT) F(ly.abar)[r0 + rr] += gp * F(ly.nll)[rr];
        }
        backward(a0, nr, false, pseudo, r0);
        r0 += nr;
      }
    }
    push_acc(2);
    cluster.sync();
    // O4: owner: dLoss/dphi_T of my slice (no analytic-KL term in the outer objective)
    for (int j = tid; j < slice; j += NT) {
      const int pp = j0 + j;
      const int q = pp < Pp ? top[pp] : -1;
      if (q < 0) continue;
      const float g_mu = recv_sum(0, j);
      const float g_rho = F(ly.sgm)[pp] * (recv_sum(1, j) + beta_sum / F(ly.sig)[pp]);
      OST(0, j) = g_mu;
      OST(1, j) = g_rho;
      if (p.flags & F_STORE_GOUT) {
        p.gout[q] = g_mu;
        p.gout[Pt + q] = g_rho;
      }
    }
    cluster.sync();  // recv may be overwritten by the next phase's pushes only after every owner has read it
  }

  // =================================================================================================================
  // Phase H: a single Hessian-vector pass along gdot (building block / hyper trainer)
  // =================================================================================================================
  if (p.flags & F_HVP) {
    for (int q = tid; q < Pt; q += NT) {
      F(ly.gmu)[pot[q]] = __ldg(p.gdot + q);
      F(ly.grho)[pot[q]] = __ldg(p.gdot + Pt + q);
    }
    __syncthreads();
    for (int s = rank; s < p.S; s += G) inner_pass(s, 0, true);
    push_acc(3);
    cluster.sync();
    for (int j = tid; j < slice; j += NT) {
      const int pp = j0 + j;
      const int q = pp < Pp ? top[pp] : -1;
      if (q < 0) continue;
      const float sg = F(ly.sig)[pp], sgm = F(ly.sgm)[pp], md = F(ly.gmu)[pp], rd = F(ly.grho)[pp];
      const float isg = 1.f / sg;
      p.h_phi[q] = recv_sum(0, j) + md;
      p.h_phi[Pt + q] = sgm * recv_sum(1, j) + sgm * (1.f - sgm) * rd * recv_sum(2, j) +
                        ((1.f + isg * isg) * sgm * sgm + (sg - isg) * sgm * (1.f - sgm)) * rd;
    }
    cluster.sync();
  }

  // =================================================================================================================
  // Phase R: reverse sweep through the T Adam steps   (A.4 + A.6; replaces autograd's double backward)
  // =================================================================================================================
  if (p.flags & F_REVERSE) {
    if (p.flags & F_LOAD_GOUT) {
      for (int j = tid; j < slice; j += NT) {
        const int pp = j0 + j;
        const int q = pp < Pp ? top[pp] : -1;
        if (q < 0) continue;
        OST(0, j) = p.gout[q];
        OST(1, j) = p.gout[Pt + q];
      }
      if (rank == 0) {
        const int MD = p.M * p.dims[0];
        for (int i = tid; i < MD; i += NT) F(ly.ubar)[i] = p.gout[2 * Pt + i];
        for (int i = tid; i < p.M; i += NT) F(ly.abar)[i] = p.gout[2 * Pt + MD + i];
      }
      __syncthreads();
    }
    for (int t = p.T - 1; t >= 0; --t) {
      const double b1t = pow(B1, (double)(t + 1)), b2t = pow(B2, (double)(t + 1));
      const float k = p.lr / (float)(1.0 - b1t);
      const float sq2 = (float)sqrt(1.0 - b2t);
      const float* tr = p.traj + (size_t)t * 8 * Pt;
      // ---- owner: Adam VJP of my slice -> direction gbar; broadcast gbar and phi_t ----
      for (int j = tid; j < slice; j += NT) {
        const int pp = j0 + j;
        const int q = pp < Pp ? top[pp] : -1;
        if (q < 0) continue;
        float gb[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float g = tr[(2 + c) * Pt + q], m = tr[(4 + c) * Pt + q], v = tr[(6 + c) * Pt + q];
          const float pb = OST(0 + c, j);
          const float qd = sqrtf(v + 1e-8f);
          const float den = qd / sq2 + aeps;
          const float mbar = OST(2 + c, j) - k * pb / den;
          const float denbar = k * pb * m / (den * den);
          float vbar = OST(4 + c, j) + denbar / (2.f * qd * sq2);
          if (v == 0.f) vbar = 0.f;  // _maybe_mask hook (optim.py:40-52,346-347)
          gb[c] = omb1 * mbar + 2.f * omb2 * g * vbar;
          OST(2 + c, j) = b1 * mbar;
          OST(4 + c, j) = b2 * vbar;
        }
        bcast(ly.gmu, pp, gb[0]);
        bcast(ly.grho, pp, gb[1]);
        bcast(ly.mu, pp, tr[q]);
        bcast(ly.rho, pp, tr[Pt + q]);
      }
      cluster.sync();
      refresh_sigma();
      // ---- every CTA: HVP pass over its samples ----
      for (int s = rank; s < p.S; s += G) inner_pass(s, t, true);
      push_acc(3);
      cluster.sync();
      // ---- owner: phibar_t = phibar_{t+1} + H gbar ----
      for (int j = tid; j < slice; j += NT) {
        const int pp = j0 + j;
        const int q = pp < Pp ? top[pp] : -1;
        if (q < 0) continue;
        const float sg = F(ly.sig)[pp], sgm = F(ly.sgm)[pp], md = F(ly.gmu)[pp], rd = F(ly.grho)[pp];
        const float isg = 1.f / sg;
        OST(0, j) += recv_sum(0, j) + md;
        OST(1, j) += sgm * recv_sum(1, j) + sgm * (1.f - sgm) * rd * recv_sum(2, j) +
                     ((1.f + isg * isg) * sgm * sgm + (sg - isg) * sgm * (1.f - sgm)) * rd;
      }
      // (the next iteration's broadcasts happen before anyone pushes into recv again: the pushes come after the
      //  cluster.sync() that follows the broadcasts)
    }
    if (p.g_out) {  // dLoss/dphi_0, useful for diagnostics
      for (int j = tid; j < slice; j += NT) {
        const int pp = j0 + j;
        const int q = pp < Pp ? top[pp] : -1;
        if (q < 0) continue;
        p.g_out[q] = OST(0, j);
        p.g_out[Pt + q] = OST(1, j);
      }
    }
  }

  // =================================================================================================================
  // Final: reduce ubar / abar over the cluster (fixed order) and map abar -> v_grad through f
  // =================================================================================================================
  if (p.flags & (F_FINAL | F_STORE_GOUT)) {
    cluster.sync();
    if (rank == 0) {
      const int MD = p.M * p.dims[0];
      float* red = F(ly.red);
      for (int i = tid; i < MD + p.M; i += NT) {
        float s = 0.f;
        const int off = i < MD ? ly.ubar + i : ly.abar + (i - MD);
        for (int c = 0; c < G; ++c) s += *remote(off, c);
        if (p.flags & F_STORE_GOUT) p.gout[2 * Pt + i] = s;
        sm[off] = s;  // rank 0 now holds the totals (remote reads of rank 0 itself happened in this same iteration)
      }
      __syncthreads();
      if (p.flags & F_FINAL) {
        for (int i = tid; i < MD; i += NT) p.u_grad[i] = F(ly.ubar)[i];
        if (p.v_grad) {
          if (p.vmode == PSVI_VMODE_IDENTITY || p.roww) {
            for (int m = tid; m < p.M; m += NT) p.v_grad[m] = p.Nf * F(ly.abar)[m];
          } else {
            float dot = 0.f;
            for (int m = tid; m < p.M; m += NT) dot += F(ly.f)[m] * F(ly.abar)[m];
            dot = block_sum(dot, red);
            const float sc = p.Nf * (p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(p.alpha) : 1.f);
            for (int m = tid; m < p.M; m += NT) p.v_grad[m] = sc * F(ly.f)[m] * (F(ly.abar)[m] - dot);
            if (p.alpha_grad && tid == 0)
              p.alpha_grad[0] = p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? sc * dot : 0.f;
          }
        }
      }
    }
    cluster.sync();  // keep every CTA's shared memory alive until rank 0 has read it
  }
}

// =================================================================================================================
// Predictive pass (PSVI.evaluate, psvi_classes.py:1031-1108; run_mfvi_subset test loop, baselines.py:1035-1043)
// =================================================================================================================
// E1: one CTA per (noise slab, MC sample): the log importance weight  lw_s = sum_m a_m nll[s,m] + nkl_s.
// NB the reference's sign quirk Q3: log_weights = -(+sum_m a_m log p) + nkl = (sum_m a_m nll) + nkl.
// The consumers (E2, the tensor-core kernel) turn the S log-weights of a slab into softmax weights in their prologue.
__device__ void Engine::eval_weights() {
  init();
  const int slab_local = blockIdx.x / p.S, s = blockIdx.x - slab_local * p.S;
  const int slab = p.first_slab + slab_local;
  float nkl = sample_theta(s, slab, false, 0.f, true);
  float ps = 0.f;
  for (int r0 = 0; r0 < p.M; r0 += p.RC) {
    const int nr = min(p.RC, p.M - r0);
    const float* a0 = stage_rows(r0, nr);
    __syncthreads();
    forward(a0, nr, false);
    loss_stage(nr, 0, r0);
    for (int rr = tid; rr < nr; rr += NT) ps += F(ly.a)[r0 + rr] * F(ly.nll)[rr];
    __syncthreads();
  }
  nkl = block_sum(nkl, F(ly.red));
  ps = block_sum(ps, F(ly.red));
  if (tid == 0) p.eval_w[(size_t)slab_local * p.S + s] = ps + nkl;
}

// E2: one CTA per chunk of test rows; loops over the S samples (re-drawing the slab's noise) and accumulates the
// predictive distribution of its rows in shared memory.
__device__ void Engine::eval_rows() {
  init();
  const int slab_local = blockIdx.x / p.chunks_per_slab, chunk = blockIdx.x - slab_local * p.chunks_per_slab;
  const int slab = p.first_slab + slab_local;
  const int row_begin = slab_local * p.batch + chunk * p.RC;
  const int slab_end = min((slab_local + 1) * p.batch, p.n_rows);
  const int nr = min(p.RC, slab_end - row_begin);
  const int C = p.dims[p.L], ld = mt.lda[p.L];
  float nll_sum = 0.f, correct = 0.f;
  if (p.eval_mode == 0) {  // importance weights of this slab: softmax over the S log-weights of E1
    if (tid == 0) {
      const float* lw = p.eval_w + (size_t)slab_local * p.S;
      float mx = -INFINITY;
      for (int s = 0; s < p.S; ++s) mx = fmaxf(mx, lw[s]);
      float se = 0.f;
      for (int s = 0; s < p.S; ++s) se += expf(lw[s] - mx);
      for (int s = 0; s < p.S; ++s) F(ly.w)[s] = expf(lw[s] - mx) / se;
    }
    __syncthreads();
  }
  if (nr > 0) {
    float* probs = F(ly.adj[p.L]);  // zeroed by init()
    const float* a0 = stage_rows(p.M + row_begin, nr);
    __syncthreads();
    for (int s = 0; s < p.S; ++s) {
      sample_theta(s, slab, false, 0.f);
      forward(a0, nr, false);
      const float wgt = p.eval_mode == 0 ? F(ly.w)[s] : 1.f / (float)p.S;
      const float* o = F(ly.act[p.L]);
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
    const int* lab = I(ly.lab);
    for (int rr = tid; rr < nr; rr += NT) {
      const float* pr = probs + rr * ld;
      const int y = lab[rr];
      int am = 0;
      float best = pr[0];
      for (int c = 1; c < C; ++c)
        if (pr[c] > best) { best = pr[c]; am = c; }
      correct += (am == y) ? 1.f : 0.f;
      if (p.eval_mode == 2) {  // Categorical(logits=mean logits)
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(pr[c] - best);
        nll_sum += best + logf(se) - pr[y];
      } else {                 // Categorical(probs=...): normalise, clamp to [eps, 1-eps], log
        float tot = 0.f;
        for (int c = 0; c < C; ++c) tot += pr[c];
        float pn = pr[y] / tot;
        pn = fminf(fmaxf(pn, 1.1920929e-07f), 1.f - 1.1920929e-07f);
        nll_sum -= logf(pn);
      }
    }
  }
  nll_sum = block_sum(nll_sum, F(ly.red));
  correct = block_sum(correct, F(ly.red));
  if (tid == 0) {
    float* o = p.eval_part + (size_t)blockIdx.x * 4;
    o[0] = nll_sum; o[1] = correct; o[2] = (float)(nr > 0 ? nr : 0); o[3] = 0.f;
  }
}

// Plain sampled forward (module-level API: VILinear / nn.Sequential forward, neural_net.py:155-179): one CTA per
// (sample, row chunk); also exports the sampled weights, sampled_nkl and the analytic KL.
__device__ void Engine::forward_rows() {
  init();
  const int s = blockIdx.x / p.chunks_per_slab, chunk = blockIdx.x - s * p.chunks_per_slab;
  const int row_begin = chunk * p.RC;
  const int nr = min(p.RC, p.n_rows - row_begin);
  const int C = p.dims[p.L], ld = mt.lda[p.L];
  float nkl = sample_theta(s, 0, false, 0.f, p.nkl_out != nullptr);
  if (chunk == 0) {
    if (p.nkl_out) {
      nkl = block_sum(nkl, F(ly.red));
      if (tid == 0) p.nkl_out[s] = nkl;
    }
    if (p.theta_out) {
      const int* pot = I(ly.pad_of_tl);
      for (int q = tid; q < mt.Pt; q += NT) p.theta_out[(size_t)s * mt.Pt + q] = F(ly.theta)[pot[q]];
    }
    if (p.kl_out && s == 0) {
      const int* top = I(ly.tl_of_pad);
      float kl = 0.f;
      for (int pp = tid; pp < mt.Pp; pp += NT)
        if (top[pp] >= 0) {
          const float sg = F(ly.sig)[pp], m = F(ly.mu)[pp];
          kl += 0.5f * (sg * sg + m * m - 1.f) - logf(sg);
        }
      kl = block_sum(kl, F(ly.red));
      if (tid == 0) p.kl_out[0] = kl;
    }
  }
  if (nr <= 0) return;
  const float* a0 = stage_rows(p.M + row_begin, nr);
  __syncthreads();
  forward(a0, nr, false);
  const float* o = F(ly.act[p.L]);
  for (int i = tid; i < nr * C; i += NT) {
    const int rr = i / C, c = i - rr * C;
    p.logits_out[((size_t)s * p.n_rows + row_begin + rr) * C + c] = o[rr * ld + c];
  }
}

__global__ void __launch_bounds__(NT, 1) psvi_mf_engine_kernel(const __grid_constant__ EP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ Lay ly;
  if (threadIdx.x == 0) {
    make_meta(p.dims, p.L, mt);
    make_layout(p, mt, ly);
  }
  __syncthreads();
  Engine e(p, mt, ly, smem_dyn);
  e.run();
}


__global__ void __launch_bounds__(NT, 1) psvi_mf_eval_weights_kernel(const __grid_constant__ EP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ Lay ly;
  if (threadIdx.x == 0) {
    make_meta(p.dims, p.L, mt);
    make_layout(p, mt, ly);
  }
  __syncthreads();
  Engine e(p, mt, ly, smem_dyn);
  e.rank = 0;
  e.eval_weights();
}

__global__ void __launch_bounds__(NT, 1) psvi_mf_eval_rows_kernel(const __grid_constant__ EP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ Lay ly;
  if (threadIdx.x == 0) {
    make_meta(p.dims, p.L, mt);
    make_layout(p, mt, ly);
  }
  __syncthreads();
  Engine e(p, mt, ly, smem_dyn);
  e.rank = 0;
  e.eval_rows();
}

// Predictive rows pass for ONE hidden layer and tiny input / output widths (the cfg2 family: D, C <= 4), register form: a thread
// owns RPT test rows (inputs and the C mixture accumulators in registers) and walks the H hidden units; a unit's sampled weights
// (D + 1 + C <= 8 floats) come from shared memory as two warp-broadcast 128-bit loads shared by its RPT rows.  Per MC sample the
// CTA rebuilds theta_s = mu + sigma eps_s (same noise addressing as Engine::sample_theta: slab, sample, TL index).  The generic
// kernel above walks the same rows through shared-memory GEMM phases with a block barrier per layer and sample (3.3 G row-samples/s
// at cfg2 shapes, ~4 % of the instruction issue rate); same per-CTA partials, same reduction kernel.
constexpr int EV_RPT = 4, EV_T = 256;
template <int D, int C>
__global__ void __launch_bounds__(EV_T) psvi_mf_eval_rows_fn1_kernel(const __grid_constant__ EP p) {
  extern __shared__ __align__(16) float ev_sm[];
  const int H = p.dims[1], HD = H * D, Pt = HD + H + C * H + C;
  float* rec = ev_sm;                 // [H][8]: w1[D], b1, w2[C]
  float* b2 = rec + H * 8;            // [4]
  float* smu = b2 + 4;                // [Pt]
  float* ssg = smu + ((Pt + 3) & ~3); // [Pt]
  float* wts = ssg + ((Pt + 3) & ~3); // [S]
  __shared__ float red[64];
  const int tid = threadIdx.x;
  const int slab_local = blockIdx.x / p.chunks_per_slab, chunk = blockIdx.x - slab_local * p.chunks_per_slab;
  const int slab = p.first_slab + slab_local;
  const int row_begin = slab_local * p.batch + chunk * (EV_T * EV_RPT);
  const int slab_end = min((slab_local + 1) * p.batch, p.n_rows);
  const int nr = max(0, min(EV_T * EV_RPT, slab_end - row_begin));
  for (int q = tid; q < Pt; q += EV_T) { smu[q] = p.mu[q]; ssg[q] = softplus_f(p.rho[q]); }
  for (int i = tid; i < H * 8 + 4; i += EV_T) rec[i] = 0.f;
  if (tid == 0) {
    if (p.eval_mode == 0) {   // importance weights of this slab: softmax over the S log-weights of E1
      const float* lw = p.eval_w + (size_t)slab_local * p.S;
      float mx = -INFINITY;
      for (int s = 0; s < p.S; ++s) mx = fmaxf(mx, lw[s]);
      float se = 0.f;
      for (int s = 0; s < p.S; ++s) se += expf(lw[s] - mx);
      for (int s = 0; s < p.S; ++s) wts[s] = expf(lw[s] - mx) / se;
    } else {
      for (int s = 0; s < p.S; ++s) wts[s] = 1.f / (float)p.S;
    }
  }
  float x[EV_RPT][D], pr[EV_RPT][C];
  int yl[EV_RPT];
  bool ok[EV_RPT];
#pragma unroll
  for (int k = 0; k < EV_RPT; ++k) {
    const int r = tid + k * EV_T;
    ok[k] = r < nr;
#pragma unroll
    for (int d = 0; d < D; ++d) x[k][d] = ok[k] ? __ldg(p.xb + (size_t)(row_begin + r) * D + d) : 0.f;
    yl[k] = ok[k] ? __ldg(p.yb + row_begin + r) : 0;
#pragma unroll
    for (int c = 0; c < C; ++c) pr[k][c] = 0.f;
  }
  const int n4 = (Pt + 3) >> 2;
  for (int s = 0; s < p.S; ++s) {
    __syncthreads();   // the previous sample's records are no longer read (first pass: the staging above is visible)
    for (int q4 = tid; q4 < n4; q4 += EV_T) {
      float e4[4];
      if (p.noise_mode == PSVI_NOISE_PHILOX) {
        philox_normal4(p.seed, p.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)q4, e4);
      } else {
        const float* src = p.eps + ((size_t)slab * p.S + s) * Pt;
#pragma unroll
        for (int j = 0; j < 4; ++j) e4[j] = (4 * q4 + j < Pt) ? __ldg(src + 4 * q4 + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = 4 * q4 + j;
        if (q < Pt) {
          const float th = smu[q] + ssg[q] * e4[j];
          int r;
          if (q < HD) { const int u = q / D; r = u * 8 + (q - u * D); }
          else if (q < HD + H) r = (q - HD) * 8 + D;
          else if (q < HD + H + C * H) { const int c = (q - HD - H) / H, u = (q - HD - H) - c * H; r = u * 8 + D + 1 + c; }
          else r = H * 8 + (q - HD - H - C * H);
          rec[r] = th;
        }
      }
    }
    __syncthreads();
    float o[EV_RPT][C];
#pragma unroll
    for (int k = 0; k < EV_RPT; ++k)
#pragma unroll
      for (int c = 0; c < C; ++c) o[k][c] = b2[c];
#pragma unroll 2
    for (int u = 0; u < H; ++u) {
      const float4 wa = *reinterpret_cast<const float4*>(rec + u * 8);
      const float4 wb = *reinterpret_cast<const float4*>(rec + u * 8 + 4);
      const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int k = 0; k < EV_RPT; ++k) {
        float a = w[D];
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(w[d], x[k][d], a);
        a = fmaxf(a, 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) o[k][c] = fmaf(a, w[D + 1 + c], o[k][c]);
      }
    }
    const float wgt = wts[s];
#pragma unroll
    for (int k = 0; k < EV_RPT; ++k) {
      if (p.eval_mode == 2) {
#pragma unroll
        for (int c = 0; c < C; ++c) pr[k][c] += wgt * o[k][c];
      } else {
        float mx = o[k][0];
#pragma unroll
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, o[k][c]);
        float ex[C], se = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { ex[c] = expf(o[k][c] - mx); se += ex[c]; }
        const float inv = wgt / se;
#pragma unroll
        for (int c = 0; c < C; ++c) pr[k][c] += inv * ex[c];
      }
    }
  }
  float nll_sum = 0.f, correct = 0.f;
#pragma unroll
  for (int k = 0; k < EV_RPT; ++k) {
    if (!ok[k]) continue;
    int am = 0;
    float best = pr[k][0], py = pr[k][0];
#pragma unroll
    for (int c = 1; c < C; ++c) {
      if (pr[k][c] > best) { best = pr[k][c]; am = c; }
      py = (yl[k] == c) ? pr[k][c] : py;
    }
    correct += (am == yl[k]) ? 1.f : 0.f;
    if (p.eval_mode == 2) {  // Categorical(logits=mean logits)
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) se += expf(pr[k][c] - best);
      nll_sum += best + logf(se) - py;
    } else {                 // Categorical(probs=...): normalise, clamp to [eps, 1-eps], log
      float tot = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) tot += pr[k][c];
      float pn = py / tot;
      pn = fminf(fmaxf(pn, 1.1920929e-07f), 1.f - 1.1920929e-07f);
      nll_sum -= logf(pn);
    }
  }
  __syncthreads();
  nll_sum = block_sum(nll_sum, red);
  correct = block_sum(correct, red);
  if (tid == 0) {
    float* out = p.eval_part + (size_t)blockIdx.x * 4;
    out[0] = nll_sum; out[1] = correct; out[2] = (float)nr; out[3] = 0.f;
  }
}

__global__ void __launch_bounds__(NT, 1) psvi_mf_forward_kernel(const __grid_constant__ EP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ Meta mt;
  __shared__ Lay ly;
  if (threadIdx.x == 0) {
    make_meta(p.dims, p.L, mt);
    make_layout(p, mt, ly);
  }
  __syncthreads();
  Engine e(p, mt, ly, smem_dyn);
  e.rank = 0;
  e.forward_rows();
}

// fixed-order final reduction of the per-CTA partials (deterministic, no atomics); also the importance-weight
// diagnostics of the LAST slab (Q12): entropy -sum w log w and normalised ESS (psvi_classes.py:1085-1092)
__global__ void psvi_mf_eval_reduce_kernel(const float* part, int n, float* out, const float* lw_last, int S) {
  __shared__ double acc[3][NT];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < n; i += NT) {
    a += part[4 * i];
    b += part[4 * i + 1];
    c += part[4 * i + 2];
  }
  acc[0][threadIdx.x] = a; acc[1][threadIdx.x] = b; acc[2][threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) {
      double t = 0;
      for (int i = 0; i < NT; ++i) t += acc[k][i];
      out[k] = (float)t;
    }
    if (lw_last) {
      float mx = -INFINITY;
      for (int s = 0; s < S; ++s) mx = fmaxf(mx, lw_last[s]);
      float se = 0.f;
      for (int s = 0; s < S; ++s) se += expf(lw_last[s] - mx);
      float ent = 0.f, sw = 0.f, sw2 = 0.f;
      for (int s = 0; s < S; ++s) {
        const float w = expf(lw_last[s] - mx) / se;
        if (w > 0.f) ent -= logf(w) * w;
        sw += w;
        sw2 += w * w;
      }
      out[3] = ent;
      out[4] = sw * sw / sw2 / (float)S;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// host side
int validate_model(const psvi_mf_model* model) {
  PSVI_REQUIRE(model != nullptr, PSVI_ERR_INVALID, "model is null");
  PSVI_REQUIRE(model->n_layers >= 1 && model->n_layers <= MAXL, PSVI_ERR_UNSUPPORTED,
               "n_layers=%d outside [1,%d]", model->n_layers, MAXL);
  for (int l = 0; l <= model->n_layers; ++l)
    PSVI_REQUIRE(model->dims[l] >= 1, PSVI_ERR_INVALID, "dims[%d]=%d must be >= 1", l, model->dims[l]);
  PSVI_REQUIRE(model->mc_samples >= 1, PSVI_ERR_INVALID, "mc_samples=%d must be >= 1", model->mc_samples);
  return PSVI_OK;
}

int launch_engine(EP& p, int rows_max, cudaStream_t stream) {
  int dev = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  int smem_max = 0;
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  {  // one-hidden-layer networks with tiny input / output widths have a shape-specialised engine (psvi_mf_fn1.cu)
    const int rc = psvi_fn1_launch(p, stream);
    if (rc != FN1_NOT_APPLICABLE) return rc;
  }
  Meta mt;
  make_meta(p.dims, p.L, mt);
  // cluster size: fewest CTAs that give every CTA the minimal number of samples, capped at 16 (non-portable max);
  // function attributes are per device, so they are set on every call (cheap) rather than cached per process
  const int max_cluster = 16;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_mf_engine_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int G = p.S < max_cluster ? p.S : max_cluster;
  if (const char* e = getenv("PSVI_ENGINE_G")) {   // diagnostic: cap the cluster size (1 = the whole step in one CTA)
    const int cap = atoi(e);
    if (cap >= 1 && cap < G) G = cap;
  }
  const int per = (p.S + G - 1) / G;
  G = (p.S + per - 1) / per;
  const bool dual = (p.flags & (F_REVERSE | F_HVP)) != 0;
  for (;; --G) {
    PSVI_REQUIRE(G >= 1, PSVI_ERR_UNSUPPORTED, "no feasible cluster size");
    p.G = G;
    p.slice = (mt.Pp + G - 1) / G;
    // rows per chunk: the largest RC <= rows_max whose carve-up fits (static shared memory needs ~1 KB as well)
    const size_t budget = (size_t)smem_max - 2048;
    Lay ly;
    p.RC = 1;
    make_layout(p, mt, ly);
    if ((size_t)ly.total * 4 > budget) {
      if (G > 1) continue;  // a smaller cluster has larger slices but fewer receive buffers; try it
      psvi_set_error("model too large for the shared-memory-resident engine (P_pad=%d, M=%d: %zu B needed, %zu B available)",
                     mt.Pp, p.M, (size_t)ly.total * 4, budget);
      return PSVI_ERR_UNSUPPORTED;
    }
    int lo = 1, hi = rows_max < 1 ? 1 : rows_max;
    while (lo < hi) {
      const int mid = (lo + hi + 1) / 2;
      p.RC = mid;
      make_layout(p, mt, ly);
      if ((size_t)ly.total * 4 <= budget) lo = mid; else hi = mid - 1;
    }
    p.RC = lo;
    make_layout(p, mt, ly);
    const size_t smem = (size_t)ly.total * 4;
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_mf_engine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(G, 1, 1);
    cfg.blockDim = dim3(NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = G;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, psvi_mf_engine_kernel, &cfg);
    if (qe != cudaSuccess || ncl < 1) {
      (void)cudaGetLastError();
      if (G == 1) {
        psvi_set_error("cluster launch not possible even with one CTA: %s", cudaGetErrorString(qe));
        return PSVI_ERR_CUDA;
      }
      // try the next smaller cluster that changes the samples-per-CTA count
      continue;
    }
    cudaError_t le = cudaLaunchKernelEx(&cfg, psvi_mf_engine_kernel, p);
    if (le != cudaSuccess) {
      (void)cudaGetLastError();
      if (G == 1) {
        psvi_set_error("engine launch failed: %s", cudaGetErrorString(le));
        return PSVI_ERR_CUDA;
      }
      continue;
    }
    return PSVI_OK;
  }
}

void fill_common(EP& p, const psvi_mf_model* model, const psvi_noise* noise, int M, float N, int vmode, float alpha) {
  memset(&p, 0, sizeof(p));
  p.L = model->n_layers;
  for (int l = 0; l <= p.L; ++l) p.dims[l] = model->dims[l];
  p.S = model->mc_samples;
  p.M = M;
  p.Nf = N;
  p.vmode = vmode;
  p.alpha = alpha;
  p.kappa = 1.f;
  p.noise_mode = noise->mode;
  p.eps = noise->eps;
  p.seed = noise->seed;
  p.domain = noise->domain;
  p.adam_mode = PSVI_ADAM_ROBUST_HIGHER;
}

int check_noise(const psvi_noise* noise) {
  PSVI_REQUIRE(noise != nullptr, PSVI_ERR_INVALID, "noise is null");
  PSVI_REQUIRE(noise->mode == PSVI_NOISE_EXTERNAL || noise->mode == PSVI_NOISE_PHILOX, PSVI_ERR_INVALID,
               "unknown noise mode %d", noise->mode);
  PSVI_REQUIRE(noise->mode != PSVI_NOISE_EXTERNAL || noise->eps != nullptr, PSVI_ERR_INVALID,
               "external noise mode needs eps");
  return PSVI_OK;
}

}  // namespace

// Internal (not part of the C ABI): log importance weights lw[S] of one noise slab, used by psvi_lr_tc.cu.
int psvi_internal_eval_logweights(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                                  const float* u, const int32_t* z, const float* v, int32_t M, int32_t slab, float N,
                                  int32_t vmode, float alpha, float* lw, cudaStream_t stream) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.mu = const_cast<float*>(mu); p.rho = const_cast<float*>(rho);
  p.u = u; p.z = z; p.v = v;
  p.n_rows = 1; p.batch = 1; p.first_slab = slab; p.eval_mode = 0; p.n_slabs = 1;
  p.eval_w = lw;
  p.G = 1; p.flags = F_EVAL;
  int dev = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  Meta mt;
  make_meta(p.dims, p.L, mt);
  p.slice = mt.Pp;
  const size_t budget = (size_t)smem_max - 2048;
  Lay ly;
  int lo = 0, hi = M < 256 ? M : 256;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    p.RC = mid;
    make_layout(p, mt, ly);
    if ((size_t)ly.total * 4 <= budget) lo = mid; else hi = mid - 1;
  }
  PSVI_REQUIRE(lo >= 1, PSVI_ERR_UNSUPPORTED, "model too large for the shared-memory-resident importance-weight kernel");
  p.RC = lo;
  make_layout(p, mt, ly);
  const size_t smem = (size_t)ly.total * 4;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_mf_eval_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psvi_mf_eval_weights_kernel<<<p.S, NT, smem, stream>>>(p);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

// ================================================================================================================
extern "C" {

int64_t psvi_mf_num_theta(const psvi_mf_model* model) {
  if (validate_model(model) != PSVI_OK) return PSVI_ERR_INVALID;
  Meta mt;
  make_meta(model->dims, model->n_layers, mt);
  return mt.Pt;
}

size_t psvi_mf_traj_bytes(const psvi_mf_model* model, int32_t T) {
  const int64_t P = psvi_mf_num_theta(model);
  if (P < 0 || T < 0) return 0;
  return (size_t)T * 8 * (size_t)P * sizeof(float);
}

int64_t psvi_mf_gout_floats(const psvi_mf_model* model, int32_t M) {
  const int64_t P = psvi_mf_num_theta(model);
  if (P < 0 || M < 0) return PSVI_ERR_INVALID;
  return 2 * P + (int64_t)M * model->dims[0] + M + 4 * (int64_t)model->mc_samples + 4;
}

int psvi_mf_nested_step(const psvi_mf_model* model, const psvi_noise* noise, float* mu, float* rho, const float* u,
                        const int32_t* z, const float* v, int32_t M, const float* xb, const int32_t* yb, int32_t B,
                        int32_t n_total_rows, float N, int32_t vmode, float alpha, int32_t T, float lr,
                        float pseudo_scale, int32_t phase_mask, float* traj, float* gout, float* u_grad, float* v_grad,
                        float* alpha_grad, float* loss_out, float* inner_losses, void* stream) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  PSVI_REQUIRE(model->mc_samples > 1, PSVI_ERR_INVALID, "psvi_elbo needs mc_samples > 1 (psvi_classes.py:449)");
  PSVI_REQUIRE(mu && rho && u && z && v && M > 0, PSVI_ERR_INVALID, "null parameter / pseudo-data pointer or M<=0");
  PSVI_REQUIRE(T >= 0 && (T == 0 || traj), PSVI_ERR_INVALID, "T<0 or missing trajectory scratch");
  PSVI_REQUIRE(phase_mask & (PSVI_PHASE_UNROLL | PSVI_PHASE_REVERSE), PSVI_ERR_INVALID, "empty phase_mask");
  PSVI_REQUIRE(B >= 0 && n_total_rows >= B && n_total_rows > 0, PSVI_ERR_INVALID, "bad minibatch sizes B=%d total=%d",
               B, n_total_rows);
  PSVI_REQUIRE(B == 0 || (xb && yb), PSVI_ERR_INVALID, "minibatch pointers are null");
  const bool both = (phase_mask & PSVI_PHASE_UNROLL) && (phase_mask & PSVI_PHASE_REVERSE);
  PSVI_REQUIRE(both || gout, PSVI_ERR_INVALID, "split phases need the gout buffer");
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.B = B;
  p.Btot = n_total_rows;
  p.T = T;
  p.lr = lr;
  p.kappa = pseudo_scale;
  p.mu = mu; p.rho = rho; p.u = u; p.z = z; p.v = v; p.xb = xb; p.yb = yb;
  p.traj = traj; p.gout = gout; p.u_grad = u_grad; p.v_grad = v_grad; p.alpha_grad = alpha_grad;
  p.loss_out = loss_out; p.inner_losses = inner_losses;
  if (phase_mask & PSVI_PHASE_UNROLL) p.flags |= F_UNROLL | F_OUTER | F_WRITE_PHI;
  if (phase_mask & PSVI_PHASE_REVERSE) {
    PSVI_REQUIRE(u_grad != nullptr, PSVI_ERR_INVALID, "u_grad is null");
    p.flags |= F_REVERSE | F_FINAL;
  }
  if (!both) p.flags |= (phase_mask & PSVI_PHASE_UNROLL) ? F_STORE_GOUT : F_LOAD_GOUT;
  if (!(phase_mask & PSVI_PHASE_UNROLL)) p.inner_losses = nullptr;
  return launch_engine(p, M > B ? M : B, (cudaStream_t)stream);
}

int psvi_mf_unroll(const psvi_mf_model* model, const psvi_noise* noise, float* mu, float* rho, float* adam_m,
                   float* adam_v, int32_t step0, const float* x, const int32_t* y, const float* row_weights,
                   const float* v, int32_t M, float N, int32_t vmode, float alpha, int32_t T, float lr,
                   int32_t adam_mode, float* losses, void* stream) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  PSVI_REQUIRE(mu && rho && x && y && M > 0 && T >= 0 && step0 >= 0, PSVI_ERR_INVALID, "bad argument");
  PSVI_REQUIRE(row_weights || v, PSVI_ERR_INVALID, "need row_weights or v");
  PSVI_REQUIRE((adam_m == nullptr) == (adam_v == nullptr), PSVI_ERR_INVALID, "adam_m / adam_v must come together");
  PSVI_REQUIRE(adam_mode >= 0 && adam_mode <= 2, PSVI_ERR_INVALID, "unknown adam_mode %d", adam_mode);
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.T = T; p.step0 = step0; p.lr = lr; p.adam_mode = adam_mode;
  p.mu = mu; p.rho = rho; p.adam_m = adam_m; p.adam_v = adam_v;
  p.u = x; p.z = y; p.v = v; p.roww = row_weights;
  p.inner_losses = losses;
  p.flags = F_UNROLL | F_WRITE_PHI;
  return launch_engine(p, M, (cudaStream_t)stream);
}

int psvi_mf_outer_grad(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                       const float* u, const int32_t* z, const float* v, int32_t M, const float* xb, const int32_t* yb,
                       int32_t B, int32_t n_total_rows, float N, int32_t vmode, float alpha, float pseudo_scale,
                       float* gout, float* u_grad, float* v_grad, float* alpha_grad, float* loss_out, void* stream) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  PSVI_REQUIRE(model->mc_samples > 1, PSVI_ERR_INVALID, "psvi_elbo needs mc_samples > 1 (psvi_classes.py:449)");
  PSVI_REQUIRE(mu && rho && u && z && v && gout && u_grad && M > 0, PSVI_ERR_INVALID, "null pointer or M<=0");
  PSVI_REQUIRE(B >= 0 && n_total_rows >= B && n_total_rows > 0 && (B == 0 || (xb && yb)), PSVI_ERR_INVALID,
               "bad minibatch");
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.B = B; p.Btot = n_total_rows; p.T = 0; p.kappa = pseudo_scale;
  p.mu = const_cast<float*>(mu); p.rho = const_cast<float*>(rho);
  p.u = u; p.z = z; p.v = v; p.xb = xb; p.yb = yb;
  p.gout = gout; p.u_grad = u_grad; p.v_grad = v_grad; p.alpha_grad = alpha_grad; p.loss_out = loss_out;
  p.flags = F_OUTER | F_STORE_GOUT | F_FINAL;
  return launch_engine(p, M > B ? M : B, (cudaStream_t)stream);
}

int psvi_mf_inner_grad(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                       const float* u, const int32_t* z, const float* v, int32_t M, float N, int32_t vmode,
                       float alpha, float* grad, float* value, void* stream) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  PSVI_REQUIRE(mu && rho && u && z && v && grad && M > 0, PSVI_ERR_INVALID, "null pointer or M<=0");
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.T = 1; p.lr = 0.f;
  p.mu = const_cast<float*>(mu); p.rho = const_cast<float*>(rho);
  p.u = u; p.z = z; p.v = v;
  p.g_out = grad; p.inner_losses = value;
  p.flags = F_UNROLL | F_NOUPDATE;
  return launch_engine(p, M, (cudaStream_t)stream);
}

int psvi_mf_inner_hvp(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                      const float* u, const int32_t* z, const float* v, int32_t M, float N, int32_t vmode,
                      float alpha, const float* gdot, float* h_phi, float* h_u, float* h_v, float* h_alpha,
                      void* stream) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  PSVI_REQUIRE(mu && rho && u && z && v && gdot && h_phi && h_u && M > 0, PSVI_ERR_INVALID, "null pointer or M<=0");
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.mu = const_cast<float*>(mu); p.rho = const_cast<float*>(rho);
  p.u = u; p.z = z; p.v = v;
  p.gdot = gdot; p.h_phi = h_phi; p.u_grad = h_u; p.v_grad = h_v; p.alpha_grad = h_alpha;
  p.flags = F_HVP | F_FINAL;
  return launch_engine(p, M, (cudaStream_t)stream);
}

size_t psvi_mf_eval_scratch_bytes(const psvi_mf_model* model, int32_t n_rows, int32_t batch) {
  if (validate_model(model) != PSVI_OK || n_rows <= 0 || batch <= 0) return 0;
  const size_t n_slabs = ((size_t)n_rows + batch - 1) / batch;
  const size_t chunks = ((size_t)(batch < n_rows ? batch : n_rows) + 31) / 32;
  return (n_slabs * model->mc_samples + n_slabs * chunks * 4 + 64) * sizeof(float);
}

int psvi_mf_evaluate(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                     const float* u, const int32_t* z, const float* v, int32_t M, const float* xt, const int32_t* yt,
                     int32_t n_rows, int32_t batch, int32_t first_slab, float N, int32_t vmode, float alpha,
                     int32_t mode, float* out, void* scratch, void* stream_) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  PSVI_REQUIRE(mu && rho && xt && yt && out && scratch, PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(n_rows > 0 && batch > 0 && first_slab >= 0, PSVI_ERR_INVALID, "bad n_rows/batch/first_slab");
  PSVI_REQUIRE(mode >= 0 && mode <= 2, PSVI_ERR_INVALID, "unknown evaluate mode %d", mode);
  PSVI_REQUIRE(mode != 0 || (u && z && v && M > 0), PSVI_ERR_INVALID, "importance-weighted mode needs pseudo-data");
  PSVI_REQUIRE(mode != 0 || model->mc_samples > 1, PSVI_ERR_INVALID, "evaluate needs mc_samples > 1 (psvi_classes.py:1036)");
  if (mode != 0) M = 0;
  EP p;
  fill_common(p, model, noise, M, N, vmode, alpha);
  p.mu = const_cast<float*>(mu); p.rho = const_cast<float*>(rho);
  p.u = u; p.z = z; p.v = v; p.xb = xt; p.yb = yt;
  p.n_rows = n_rows; p.batch = batch; p.first_slab = first_slab; p.eval_mode = mode;
  p.n_slabs = (n_rows + batch - 1) / batch;
  p.eval_w = (float*)scratch;
  p.eval_part = p.eval_w + (size_t)p.n_slabs * p.S;
  p.eval_out = out;
  p.G = 1; p.flags = F_EVAL;
  int dev = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  Meta mt;
  make_meta(p.dims, p.L, mt);
  p.slice = mt.Pp;
  const size_t budget = (size_t)smem_max - 2048;
  Lay ly;
  const int rows_cap = batch < n_rows ? batch : n_rows;
  const int want = (M > rows_cap ? M : rows_cap) < 256 ? (M > rows_cap ? M : rows_cap) : 256;
  int lo = 0, hi = want;
  while (lo < hi) {  // largest RC <= want that fits
    const int mid = (lo + hi + 1) / 2;
    p.RC = mid;
    make_layout(p, mt, ly);
    if ((size_t)ly.total * 4 <= budget) lo = mid; else hi = mid - 1;
  }
  PSVI_REQUIRE(lo >= 32 || lo >= want, PSVI_ERR_UNSUPPORTED,
               "model too large for the shared-memory-resident predictive kernel (rows per chunk %d)", lo);
  p.RC = lo;
  make_layout(p, mt, ly);
  const size_t smem = (size_t)ly.total * 4;
  p.chunks_per_slab = (rows_cap + p.RC - 1) / p.RC;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_mf_eval_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_mf_eval_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (mode == 0) {
    psvi_mf_eval_weights_kernel<<<p.n_slabs * p.S, NT, smem, stream>>>(p);
    PSVI_CUDA_CHECK(cudaGetLastError());
  }
  // one hidden layer, D and C in {2, 4} (the cfg2 family), hidden layer within 48 KB of records: the register-form rows pass
  const bool fn1_rows = p.L == 2 && (p.dims[0] == 2 || p.dims[0] == 4) && (p.dims[2] == 2 || p.dims[2] == 4) &&
                        p.dims[0] + 1 + p.dims[2] <= 8 && p.dims[1] <= 1024 && !getenv("PSVI_EVAL_GENERIC");
  int nctas = p.n_slabs * p.chunks_per_slab;
  if (fn1_rows) {
    const int H = p.dims[1], Pt = mt.Pt;
    p.chunks_per_slab = (rows_cap + EV_T * EV_RPT - 1) / (EV_T * EV_RPT);
    nctas = p.n_slabs * p.chunks_per_slab;
    const size_t sm1 = (size_t)(H * 8 + 4 + 2 * ((Pt + 3) & ~3) + p.S + 4) * 4;
    void (*k)(const EP) = p.dims[0] == 2 ? (p.dims[2] == 2 ? psvi_mf_eval_rows_fn1_kernel<2, 2> : psvi_mf_eval_rows_fn1_kernel<2, 4>)
                                         : (p.dims[2] == 2 ? psvi_mf_eval_rows_fn1_kernel<4, 2> : psvi_mf_eval_rows_fn1_kernel<4, 4>);
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    k<<<nctas, EV_T, sm1, stream>>>(p);
  } else {
    psvi_mf_eval_rows_kernel<<<nctas, NT, smem, stream>>>(p);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  psvi_mf_eval_reduce_kernel<<<1, NT, 0, stream>>>(p.eval_part, nctas, out,
                                                    mode == 0 ? p.eval_w + (size_t)(p.n_slabs - 1) * p.S : nullptr, p.S);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_mf_forward(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                    const float* x, int32_t n_rows, float* logits, float* theta_out, float* nkl_out, float* kl_out,
                    void* stream_) {
  int rc = validate_model(model);
  if (rc) return rc;
  rc = check_noise(noise);
  if (rc) return rc;
  PSVI_REQUIRE(mu && rho && x && logits && n_rows > 0, PSVI_ERR_INVALID, "null pointer or n_rows<=0");
  cudaStream_t stream = (cudaStream_t)stream_;
  EP p;
  fill_common(p, model, noise, 0, 0.f, 0, 0.f);
  p.mu = const_cast<float*>(mu); p.rho = const_cast<float*>(rho);
  p.xb = x; p.n_rows = n_rows;
  p.logits_out = logits; p.theta_out = theta_out; p.nkl_out = nkl_out; p.kl_out = kl_out;
  p.G = 1; p.flags = F_EVAL;
  int dev = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  Meta mt;
  make_meta(p.dims, p.L, mt);
  p.slice = mt.Pp;
  const size_t budget = (size_t)smem_max - 2048;
  Lay ly;
  const int want = n_rows < 128 ? n_rows : 128;
  int lo = 0, hi = want;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    p.RC = mid;
    make_layout(p, mt, ly);
    if ((size_t)ly.total * 4 <= budget) lo = mid; else hi = mid - 1;
  }
  PSVI_REQUIRE(lo >= 1, PSVI_ERR_UNSUPPORTED, "model too large for the shared-memory-resident forward kernel");
  p.RC = lo;
  make_layout(p, mt, ly);
  const size_t smem = (size_t)ly.total * 4;
  p.chunks_per_slab = (n_rows + p.RC - 1) / p.RC;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_mf_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psvi_mf_forward_kernel<<<p.S * p.chunks_per_slab, NT, smem, stream>>>(p);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
