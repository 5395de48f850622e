// psvi_mf_fn1.cu -- shape-specialised cluster engine for the PSVI bilevel step of `fn` with ONE hidden layer and tiny
// input / output widths (BASELINE cfg2: halfmoon / four_blobs, D = 2, H = 100, C = 2 / 4, M = 10..50, S = 10, T = 100).
//
// Same contract, phases and arithmetic as the generic engine (psvi_mf_engine.cu; reference psvi/inference/
// psvi_classes.py:445-600, psvi/robust_higher/optim.py:299-367; math: SURVEY.md Appendix A.1-A.6) -- what changes is
// how the per-sample network pass maps onto an SM.  The generic engine runs every matrix product as a separate
// shared-memory GEMM stage with run-time strides (15 instructions per FMA, ~12 block barriers per phase).  Here
//   * (D, C, hidden units per lane) are template parameters: every product loop is fully unrolled;
//   * a warp owns two rows at a time; its 32 lanes are 2 row lanes x 16 hidden lanes, a hidden lane owns UPL hidden
//     units (H <= 16 UPL) whose weights it reads as 128-bit words of a conflict-free record layout;
//   * forward, softmax / NLL, backward (or the whole forward-over-reverse dual pass of A.6) for a row happen in
//     registers: the logits (and their tangents / the input adjoints) are the only cross-lane quantities -- four
//     shuffle steps over the 16 hidden lanes; the per-sample weight adjoints  W1bar = abar^T X,  W2bar = obar^T h
//     (and their A.6 counterparts) accumulate in REGISTERS across all rows of the pass and are combined across the
//     two row lanes (one shuffle) and the eight warps (one shared-memory sweep) once per pass;
//   * one pass costs two block barriers; a phase (all samples of the CTA + exchange) two cluster barriers, the first
//     split into arrive / wait with the next phase's Philox normals generated in its shadow.
// Exchange between the CTAs of the cluster (one per MC sample) is the generic engine's: slice owners receive the
// partial sums through DSMEM, reduce in a fixed order (deterministic), update, and push the new parameters back.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "psvi_mf_engine.cuh"

namespace cg = cooperative_groups;
using namespace psvi_mf;

namespace {

constexpr int REC = 20;        // floats per hidden-unit record: [0,8) sampled weights, [8,16) tangent weights, 4 pad
constexpr int NW = NT / 32;    // warps per CTA

// shared-memory carve-up (offsets in floats)
struct FL {
  int rec, part;                                             // (HP+1) records ; NW x (HP+1) records of partial adjoints
  int mu, rho, sig, sgm, eps, gdm, gdr, accA, accB, accC;    // Pt each (TL order)
  int q2r;                                                   // Pt ints: TL index -> float offset inside `rec`
  int recv, ost;                                             // [G][3][slice], [10][slice]
  int X, Y, cw;                                              // rows: [R][D], [R] ints, [R]
  int a, f, ubar, abar;                                      // coreset: [M], [M], [M][D], [M]
  int lw, e, dsv, w, beta, gp;                               // per-sample scalars (lw, e: doubles)
  int red, lossrecv, psc;                                    // 64, G, NW*2
  int total;
};

template <int D, int C, int UPL>
__host__ __device__ inline void make_fl(const EP& p, FL& y) {
  constexpr int HP = 16 * UPL;
  const int H = p.dims[1];
  const int Pt = H * (D + 1) + C * (H + 1);
  const int R = p.M + p.B;
  int o = 0;
  auto take = [&](int n) {
    int r = o;
    o += (n + 3) & ~3;
    return r;
  };
  y.rec = take((HP + 1) * REC);
  y.part = take(NW * (HP + 1) * REC);
  y.mu = take(Pt); y.rho = take(Pt); y.sig = take(Pt); y.sgm = take(Pt); y.eps = take(Pt);
  y.gdm = take(Pt); y.gdr = take(Pt); y.accA = take(Pt); y.accB = take(Pt); y.accC = take(Pt);
  y.q2r = take(Pt);
  y.recv = take(p.G * 3 * p.slice);
  y.ost = take(10 * p.slice);
  y.X = take(R * D); y.Y = take(R); y.cw = take(R);
  y.a = take(p.M); y.f = take(p.M); y.ubar = take(p.M * D); y.abar = take(p.M);
  y.lw = take(2 * p.S); y.e = take(2 * p.S); y.dsv = take(p.S); y.w = take(p.S); y.beta = take(p.S); y.gp = take(p.S);
  y.red = take(64); y.lossrecv = take(p.G); y.psc = take(NW * 2);
  y.total = o;
}

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() {
  cl_arrive();
  cl_wait();
}
// sum over the 16 hidden lanes of a row lane group (lane bits 0..3); every lane ends with the total
__device__ __forceinline__ float hl_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

template <int D, int C, int UPL>
struct Fn1 {
  static constexpr int HP = 16 * UPL;
  static constexpr int PS = (HP + 1) * REC;  // floats of one warp's partial-adjoint block
  static_assert(D + 1 + C <= 8, "a hidden unit's weights must fit one 8-float record half");

  const EP& p;
  const FL& y;
  float* sm;
  cg::cluster_group cluster;
  int rank, tid, warp, lane, hl, rl;
  int H, Pt, n4;

  __device__ Fn1(const EP& p_, const FL& y_, float* s_) : p(p_), y(y_), sm(s_), cluster(cg::this_cluster()) {
    rank = (int)cluster.block_rank();
    tid = threadIdx.x;
    warp = tid >> 5;
    lane = tid & 31;
    hl = lane & 15;
    rl = lane >> 4;
    H = p.dims[1];
    Pt = H * (D + 1) + C * (H + 1);
    n4 = (Pt + 3) >> 2;
  }
  __device__ __forceinline__ float* F(int off) const { return sm + off; }
  __device__ __forceinline__ int* I(int off) const { return reinterpret_cast<int*>(sm + off); }
  __device__ __forceinline__ float* remote(int off, int r) { return cluster.map_shared_rank(sm + off, r); }

  // ---- sigma = softplus(rho), sgm = sigmoid(rho) after phi changed ---------------------------------------------------
  __device__ void refresh_sigma() {
    for (int q = tid; q < Pt; q += NT) {
      const float r = F(y.rho)[q];
      F(y.sig)[q] = softplus_f(r);
      F(y.sgm)[q] = sigmoid_f(r);
    }
    __syncthreads();
  }

  // ---- the four standard normals of TL block q4 of (slab, s) ----------------------------------------------------------
  __device__ __forceinline__ void draw4(int s, int slab, int q4, float e4[4]) const {
    if (p.noise_mode == PSVI_NOISE_PHILOX) {
      philox_normal4(p.seed, p.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)q4, e4);
    } else {
      const float* src = p.eps + ((size_t)slab * p.S + s) * Pt;
#pragma unroll
      for (int j = 0; j < 4; ++j) e4[j] = (4 * q4 + j < Pt) ? __ldg(src + 4 * q4 + j) : 0.f;
    }
  }

  // ---- theta_s = mu + sigma eps_s (and the tangent) into the unit records; returns this thread's partial of
  //      sampled_nkl_s (neural_net.py:110-115).  `pre` (nullable): normals of block q4 == tid drawn ahead of time. -----
  __device__ float sample_theta(int s, int slab, bool tangent, float fold_beta, bool want_nkl, const float* pre) {
    const int* q2r = I(y.q2r);
    float* rec = F(y.rec);
    float nkl = 0.f;
    for (int q4 = tid; q4 < n4; q4 += NT) {
      float e4[4];
      if (pre != nullptr && q4 == tid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) e4[j] = pre[j];
      } else {
        draw4(s, slab, q4, e4);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = 4 * q4 + j;
        if (q < Pt) {
          const float e = e4[j], sg = F(y.sig)[q];
          const float th = F(y.mu)[q] + sg * e;
          const int r = q2r[q];
          F(y.eps)[q] = e;
          rec[r] = th;
          if (tangent) rec[r + 8] = F(y.gdm)[q] + F(y.sgm)[q] * F(y.gdr)[q] * e;
          if (want_nkl) nkl += -0.5f * th * th + 0.5f * e * e + logf(sg);
          if (fold_beta != 0.f) {  // outer objective: d nkl_s / d theta = -theta, weighted by beta_s (A.2)
            const float tb = -fold_beta * th;
            F(y.accA)[q] += tb;
            F(y.accB)[q] += tb * e;
          }
        }
      }
    }
    __syncthreads();
    return nkl;
  }

  // ---- one hidden unit's record half as 8 floats ----------------------------------------------------------------------
  __device__ __forceinline__ void load8(const float* ptr, float t[8]) const {
    const float4 a = *reinterpret_cast<const float4*>(ptr);
    const float4 b = *reinterpret_cast<const float4*>(ptr + 4);
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w;
    t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
  }
  __device__ __forceinline__ void store8(float* ptr, const float t[8]) const {
    *reinterpret_cast<float4*>(ptr) = make_float4(t[0], t[1], t[2], t[3]);
    *reinterpret_cast<float4*>(ptr + 4) = make_float4(t[4], t[5], t[6], t[7]);
  }

  // ---- primal pass over rows [0, R): MODE 0 = values only, 1 = gradient.  Row weights in cw[]; sumA += sum over pseudo
  //      rows of a_m nll, sumD += sum over data rows of nll (held by the hl == 0 lanes).  need_x: input adjoints of the
  //      pseudo rows -> ubar, and abar += gp * nll.  Leaves the per-warp weight adjoints in `part`. ---------------------
  template <int MODE>
  __device__ void rows_primal(int R, bool need_x, float gp, float& sumA, float& sumD) {
    const float* rec = F(y.rec);
    const float* X = F(y.X);
    const int* Y = I(y.Y);
    const float* cw = F(y.cw);
    const float* av = F(y.a);
    float w1[UPL][D], b1[UPL], w2[UPL][C], b2[C];
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
      float t[8];
      load8(rec + (hl + 16 * i) * REC, t);
#pragma unroll
      for (int d = 0; d < D; ++d) w1[i][d] = t[d];
      b1[i] = t[D];
#pragma unroll
      for (int c = 0; c < C; ++c) w2[i][c] = t[D + 1 + c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c) b2[c] = rec[HP * REC + c];
    float gw1[UPL][D], gb1[UPL], gw2[UPL][C], gb2[C];
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
      gb1[i] = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) gw1[i][d] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) gw2[i][c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) gb2[c] = 0.f;

    for (int rb = 2 * warp; rb < R; rb += 2 * NW) {  // warp-uniform trip count (the shuffles need the full warp)
      const int r = rb + rl;
      const bool ok = r < R;
      float x[D];
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = ok ? X[r * D + d] : 0.f;
      const int yl = ok ? Y[r] : 0;
      float h[UPL], o[C];
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = 0.f;
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        float a = b1[i];
#pragma unroll
        for (int d = 0; d < D; ++d) a = fmaf(w1[i][d], x[d], a);
        h[i] = fmaxf(a, 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = fmaf(h[i], w2[i][c], o[c]);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = hl_sum(o[c]) + b2[c];
      float mx = o[0];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, o[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) se += expf(o[c] - mx);
      const float lse = mx + logf(se);
      float oy = o[0];
#pragma unroll
      for (int c = 1; c < C; ++c) oy = (yl == c) ? o[c] : oy;
      const float nll = lse - oy;
      if (hl == 0 && ok) {
        if (r < p.M) sumA += av[r] * nll; else sumD += nll;
      }
      if (MODE == 1) {
        const float cwr = ok ? cw[r] : 0.f;
        float ob[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ob[c] = cwr * (expf(o[c] - lse) - (c == yl ? 1.f : 0.f));
        float xb[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xb[d] = 0.f;
#pragma unroll
        for (int i = 0; i < UPL; ++i) {
          float t = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            gw2[i][c] = fmaf(ob[c], h[i], gw2[i][c]);
            t = fmaf(ob[c], w2[i][c], t);
          }
          const float ab = h[i] > 0.f ? t : 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) gw1[i][d] = fmaf(ab, x[d], gw1[i][d]);
          gb1[i] += ab;
          if (need_x) {
#pragma unroll
            for (int d = 0; d < D; ++d) xb[d] = fmaf(ab, w1[i][d], xb[d]);
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) gb2[c] += ob[c];
        if (need_x && rb < p.M) {  // (rb is warp-uniform)
#pragma unroll
          for (int d = 0; d < D; ++d) xb[d] = hl_sum(xb[d]);
          if (hl == 0 && ok && r < p.M) {
#pragma unroll
            for (int d = 0; d < D; ++d) F(y.ubar)[r * D + d] += xb[d];
            F(y.abar)[r] += gp * nll;
          }
        }
      }
    }
    if (MODE == 1) {
      float* part = F(y.part) + warp * PS;
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) t[d] = gw1[i][d];
        t[D] = gb1[i];
#pragma unroll
        for (int c = 0; c < C; ++c) t[D + 1 + c] = gw2[i][c];
#pragma unroll
        for (int k = 0; k < D + 1 + C; ++k) t[k] += __shfl_xor_sync(0xffffffffu, t[k], 16);
        if (rl == 0) store8(part + (hl + 16 * i) * REC, t);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float g = gb2[c] + __shfl_xor_sync(0xffffffffu, gb2[c], 16);
        if (lane == 0) part[HP * REC + c] = g;
      }
    }
  }

  // ---- dual (forward-over-reverse) pass over the M pseudo rows, A.6: leaves A_theta in the first and A_thetadot in the
  //      second half of the `part` records; ubar += A_X, abar += A_c. ---------------------------------------------------
  __device__ void rows_dual() {
    const float* rec = F(y.rec);
    const float* X = F(y.X);
    const int* Y = I(y.Y);
    const float* cw = F(y.cw);
    const int R = p.M;
    constexpr int K = D + 1 + C;
    float g[UPL][K], gd[UPL][K], gb2[C], gb2d[C];
#pragma unroll
    for (int i = 0; i < UPL; ++i)
#pragma unroll
      for (int k = 0; k < K; ++k) { g[i][k] = 0.f; gd[i][k] = 0.f; }
#pragma unroll
    for (int c = 0; c < C; ++c) { gb2[c] = 0.f; gb2d[c] = 0.f; }
    float b2[C], b2d[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { b2[c] = rec[HP * REC + c]; b2d[c] = rec[HP * REC + 8 + c]; }

    for (int rb = 2 * warp; rb < R; rb += 2 * NW) {
      const int r = rb + rl;
      const bool ok = r < R;
      float x[D];
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = ok ? X[r * D + d] : 0.f;
      const int yl = ok ? Y[r] : 0;
      const float cwr = ok ? cw[r] : 0.f;
      float h[UPL], hd[UPL], o[C], od[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { o[c] = 0.f; od[c] = 0.f; }
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        float w[8], wd[8];
        load8(rec + (hl + 16 * i) * REC, w);
        load8(rec + (hl + 16 * i) * REC + 8, wd);
        float a = w[D], ad = wd[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { a = fmaf(w[d], x[d], a); ad = fmaf(wd[d], x[d], ad); }
        const bool k = a > 0.f;
        h[i] = k ? a : 0.f;
        hd[i] = k ? ad : 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          o[c] = fmaf(h[i], w[D + 1 + c], o[c]);
          od[c] = fmaf(hd[i], w[D + 1 + c], fmaf(h[i], wd[D + 1 + c], od[c]));
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) { o[c] = hl_sum(o[c]) + b2[c]; od[c] = hl_sum(od[c]) + b2d[c]; }
      float mx = o[0];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, o[c]);
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) se += expf(o[c] - mx);
      const float lse = mx + logf(se);
      float pc[C], pd = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) { pc[c] = expf(o[c] - lse); pd = fmaf(pc[c], od[c], pd); }
      float Aod[C], Ao[C], ac = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float qc = pc[c] - (c == yl ? 1.f : 0.f);
        Aod[c] = cwr * qc;                    // adjoint of odot
        Ao[c] = cwr * pc[c] * (od[c] - pd);   // adjoint of o
        ac = fmaf(qc, od[c], ac);
        gb2[c] += Ao[c];
        gb2d[c] += Aod[c];
      }
      float ax[D];
#pragma unroll
      for (int d = 0; d < D; ++d) ax[d] = 0.f;
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        float w[8], wd[8];
        load8(rec + (hl + 16 * i) * REC, w);
        load8(rec + (hl + 16 * i) * REC + 8, wd);
        float Ahd = 0.f, Ah = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          Ahd = fmaf(Aod[c], w[D + 1 + c], Ahd);
          Ah = fmaf(Aod[c], wd[D + 1 + c], fmaf(Ao[c], w[D + 1 + c], Ah));
          g[i][D + 1 + c] = fmaf(Aod[c], hd[i], fmaf(Ao[c], h[i], g[i][D + 1 + c]));
          gd[i][D + 1 + c] = fmaf(Aod[c], h[i], gd[i][D + 1 + c]);
        }
        const bool k = h[i] > 0.f;
        const float Aa = k ? Ah : 0.f, Aad = k ? Ahd : 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          g[i][d] = fmaf(Aa, x[d], g[i][d]);
          gd[i][d] = fmaf(Aad, x[d], gd[i][d]);
          ax[d] = fmaf(Aa, w[d], fmaf(Aad, wd[d], ax[d]));
        }
        g[i][D] += Aa;
        gd[i][D] += Aad;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) ax[d] = hl_sum(ax[d]);
      if (hl == 0 && ok) {
#pragma unroll
        for (int d = 0; d < D; ++d) F(y.ubar)[r * D + d] += ax[d];
        F(y.abar)[r] += ac;
      }
    }
    float* part = F(y.part) + warp * PS;
#pragma unroll
    for (int i = 0; i < UPL; ++i) {
      float t[8], td[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { t[k] = 0.f; td[k] = 0.f; }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        t[k] = g[i][k] + __shfl_xor_sync(0xffffffffu, g[i][k], 16);
        td[k] = gd[i][k] + __shfl_xor_sync(0xffffffffu, gd[i][k], 16);
      }
      if (rl == 0) {
        store8(part + (hl + 16 * i) * REC, t);
        store8(part + (hl + 16 * i) * REC + 8, td);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float ga = gb2[c] + __shfl_xor_sync(0xffffffffu, gb2[c], 16);
      const float gb = gb2d[c] + __shfl_xor_sync(0xffffffffu, gb2d[c], 16);
      if (lane == 0) { part[HP * REC + c] = ga; part[HP * REC + 8 + c] = gb; }
    }
  }

  // ---- fold the eight warps' partial adjoints of the sample just processed into the CTA accumulators ------------------
  __device__ void fold_part(bool dual) {
    const int* q2r = I(y.q2r);
    const float* part = F(y.part);
    for (int q = tid; q < Pt; q += NT) {
      const int r = q2r[q];
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += part[w * PS + r];
      const float e = F(y.eps)[q];
      F(y.accA)[q] += s;
      F(y.accB)[q] += s * e;
      if (dual) {
        float sd = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) sd += part[w * PS + r + 8];
        F(y.accC)[q] += sd * e;
      }
    }
    // (same q <-> thread mapping as push_acc: no barrier needed before it; `part` / `eps` are rewritten only after the
    //  block barrier at the end of the next sample_theta)
  }

  // ---- push this CTA's accumulators to the slice owners, then clear them ----------------------------------------------
  __device__ void push_acc(int ncomp) {
    const int slice = p.slice;
    for (int q = tid; q < Pt; q += NT) {
      const int owner = q / slice, j = q - owner * slice;
      float* r = remote(y.recv, owner) + (size_t)rank * 3 * slice + j;
      r[0] = F(y.accA)[q];
      r[slice] = F(y.accB)[q];
      F(y.accA)[q] = 0.f;
      F(y.accB)[q] = 0.f;
      if (ncomp > 2) {
        r[2 * slice] = F(y.accC)[q];
        F(y.accC)[q] = 0.f;
      }
    }
  }
  __device__ __forceinline__ float recv_sum(int comp, int j) const {
    float s = 0.f;
    const float* r = sm + y.recv + comp * p.slice + j;
    for (int c = 0; c < p.G; ++c) s += r[(size_t)c * 3 * p.slice];
    return s;
  }
  __device__ __forceinline__ void bcast(int off, int q, float val) {
    for (int c = 0; c < p.G; ++c) remote(off, c)[q] = val;
  }

  // ---- coreset weights a = N f(v)  (psvi_classes.py:476,505; f per class :111,:1358,:1486) ----------------------------
  __device__ void setup_coreset() {
    float* a = F(y.a);
    float* f = F(y.f);
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
    mx = block_max(mx, F(y.red));
    float se = 0.f;
    for (int m = tid; m < M; m += NT) se += expf(__ldg(p.v + m) - mx);
    se = block_sum(se, F(y.red));
    const float sc = p.Nf * (p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(p.alpha) : 1.f);
    for (int m = tid; m < M; m += NT) {
      f[m] = expf(__ldg(p.v + m) - mx) / se;
      a[m] = sc * f[m];
    }
    __syncthreads();
  }

  __device__ void init() {
    for (int i = tid; i < y.total; i += NT) sm[i] = 0.f;
    __syncthreads();
    int* q2r = I(y.q2r);
    const int HD = H * D;
    for (int q = tid; q < Pt; q += NT) {
      int r;
      if (q < HD) {
        const int j = q / D;
        r = j * REC + (q - j * D);
      } else if (q < HD + H) {
        r = (q - HD) * REC + D;
      } else if (q < HD + H + C * H) {
        const int c = (q - HD - H) / H, j = (q - HD - H) - c * H;
        r = j * REC + D + 1 + c;
      } else {
        r = HP * REC + (q - HD - H - C * H);
      }
      q2r[q] = r;
      F(y.mu)[q] = p.mu[q];
      F(y.rho)[q] = p.rho[q];
    }
    for (int i = tid; i < p.M * D; i += NT) F(y.X)[i] = __ldg(p.u + i);
    for (int m = tid; m < p.M; m += NT) I(y.Y)[m] = __ldg(p.z + m);
    for (int i = tid; i < p.B * D; i += NT) F(y.X)[p.M * D + i] = __ldg(p.xb + i);
    for (int b = tid; b < p.B; b += NT) I(y.Y)[p.M + b] = __ldg(p.yb + b);
    __syncthreads();
    refresh_sigma();
    if (p.M > 0) setup_coreset();
  }

  // ---- one gradient / dual pass of the inner objective for sample s; returns (on every thread) nothing; the weighted
  //      NLL sum of the sample is left in this thread's `sumA` partial (hl == 0 lanes) --------------------------------
  __device__ void inner_pass(int s, int slab, bool dual, const float* pre, float& sumA) {
    sample_theta(s, slab, dual, 0.f, false, pre);
    float sumD = 0.f;
    if (dual) rows_dual(); else rows_primal<1>(p.M, false, 0.f, sumA, sumD);
    __syncthreads();
    fold_part(dual);
    if (s + p.G < p.S) __syncthreads();   // more samples on this CTA: eps / rec are rewritten by other threads
  }

  __device__ void run();
};

template <int D, int C, int UPL>
__device__ void Fn1<D, C, UPL>::run() {
  const int slice = p.slice, G = p.G;
  const int j0 = rank * slice;  // first TL index of my slice
  float* ost = F(y.ost);
  // owner state rows: 0 pbar_mu 1 pbar_rho 2 mbar_mu 3 mbar_rho 4 vbar_mu 5 vbar_rho 6 am_mu 7 am_rho 8 av_mu 9 av_rho
  auto OST = [&](int row, int j) -> float& { return ost[row * slice + j]; };
  init();
  // inner passes weight the rows by a_m
  for (int m = tid; m < p.M; m += NT) F(y.cw)[m] = F(y.a)[m];
  if (p.adam_m && (p.flags & F_UNROLL)) {
    for (int j = tid; j < slice; j += NT) {
      const int q = j0 + j;
      if (q < Pt) {
        OST(6, j) = p.adam_m[q];
        OST(7, j) = p.adam_m[Pt + q];
        OST(8, j) = p.adam_v[q];
        OST(9, j) = p.adam_v[Pt + q];
      }
    }
  }
  __syncthreads();
  cl_sync();  // every CTA's shared memory is initialised before anybody pushes into it

  const double B1 = 0.9, B2 = 0.999;
  const float b1 = (float)B1, b2 = (float)B2, omb1 = (float)(1.0 - B1), omb2 = (float)(1.0 - B2), aeps = 1e-8f;
  const bool want_loss = p.inner_losses != nullptr;
  const bool can_pre = n4 <= NT;   // one Philox block per thread can be drawn ahead of time
  float pre[4] = {0.f, 0.f, 0.f, 0.f};

  // =================================================================================================================
  // Phase U: T inner Adam steps  (psvi_classes.py:549-555 ; optim.py:224-229,303-367)
  // =================================================================================================================
  if (p.flags & F_UNROLL) {
    double b1t = pow(B1, (double)p.step0), b2t = pow(B2, (double)p.step0);
    if (can_pre && tid < n4) draw4(rank, 0, tid, pre);
    for (int t = 0; t < p.T; ++t) {
      float lpart = 0.f;
      for (int s = rank; s < p.S; s += G) inner_pass(s, t, false, (can_pre && s == rank) ? pre : nullptr, lpart);
      push_acc(2);
      if (want_loss) {
        // KL(q||p) of my slice at phi_t (neural_net.py:101-108), added once (Q1)
        for (int j = tid; j < slice; j += NT) {
          const int q = j0 + j;
          if (q < Pt) {
            const float sg = F(y.sig)[q], m = F(y.mu)[q];
            lpart += 0.5f * (sg * sg + m * m - 1.f) - logf(sg);
          }
        }
        lpart = block_sum(lpart, F(y.red));
        if (tid == 0) remote(y.lossrecv, 0)[rank] = lpart;
      }
      cl_arrive();
      if (can_pre && tid < n4 && t + 1 < p.T) draw4(rank, t + 1, tid, pre);   // in the shadow of the barrier
      cl_wait();
      if (want_loss && rank == 0 && tid == 0) {
        float s = 0.f;
        for (int c = 0; c < G; ++c) s += F(y.lossrecv)[c];
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
        const int q = j0 + j;
        if (q >= Pt) continue;
        const float mu = F(y.mu)[q], rho = F(y.rho)[q], sg = F(y.sig)[q], sgm = F(y.sgm)[q];
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
          bcast(y.mu, q, np[0]);
          bcast(y.rho, q, np[1]);
        }
      }
      cl_sync();
      refresh_sigma();
    }
    if (p.flags & F_WRITE_PHI) {
      for (int j = tid; j < slice; j += NT) {
        const int q = j0 + j;
        if (q >= Pt) continue;
        p.mu[q] = F(y.mu)[q];
        p.rho[q] = F(y.rho)[q];
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
      float nkl = sample_theta(s, slab, false, 0.f, true, nullptr);
      float ps = 0.f, ds = 0.f;
      rows_primal<0>(R, false, 0.f, ps, ds);
      // the S per-sample sums are O(N) while the importance-weight adjoints depend on their *differences*:
      // reduce and keep them in double (the fp32 reference loses ~2 digits here at init_sd=1e-6)
      const double nkl_d = block_sum_d((double)nkl, F(y.red));
      const double ps_d = block_sum_d((double)ps, F(y.red));
      const double ds_d = block_sum_d((double)ds, F(y.red)) * (double)dscale;
      if (tid < G) {
        reinterpret_cast<double*>(remote(y.lw, tid))[s] = -ps_d + nkl_d;
        reinterpret_cast<double*>(remote(y.e, tid))[s] = ds_d - (double)p.kappa * ps_d;
        remote(y.dsv, tid)[s] = (float)ds_d;
      }
    }
    cl_sync();
    // O2: importance weights and adjoint seeds (every CTA, redundantly; S is tiny)
    if (tid == 0) {
      const int S = p.S;
      const double* lw = reinterpret_cast<const double*>(F(y.lw));
      const double* ev = reinterpret_cast<const double*>(F(y.e));
      double mx = -INFINITY;
      for (int s = 0; s < S; ++s) mx = fmax(mx, lw[s]);
      double se = 0.0;
      for (int s = 0; s < S; ++s) se += exp(lw[s] - mx);
      double ebar = 0.0, lwm = 0.0;
      for (int s = 0; s < S; ++s) {
        const double w = exp(lw[s] - mx) / se;
        F(y.w)[s] = (float)w;
        ebar += w * ev[s];
        lwm += lw[s];
      }
      lwm /= (double)S;
      double bsum = 0.0;
      for (int s = 0; s < S; ++s) {
        const double w = exp(lw[s] - mx) / se;
        const double beta = w * (ev[s] - ebar) - (double)p.kappa / (double)S;  // dLoss/dlw_s
        F(y.beta)[s] = (float)beta;
        F(y.gp)[s] = (float)(-(double)p.kappa * w - beta);                    // dLoss/dp_s
        bsum += beta;
      }
      F(y.red)[60] = (float)bsum;
      const float lossv = (float)(ebar - (double)p.kappa * lwm);
      if (rank == 0) {
        if (p.loss_out) p.loss_out[0] = lossv;
        if (p.flags & F_STORE_GOUT) {
          float* go = p.gout + 2 * Pt + p.M * D + p.M;
          for (int s = 0; s < S; ++s) go[s] = F(y.dsv)[s];
          go[S] = lossv;
          go[S + 1] = (float)ebar;
          go[S + 2] = (float)lwm;
          go[S + 3] = (float)bsum;
          for (int s = 0; s < S; ++s) {  // diagnostics (rank-local, not meant to be all-reduced)
            go[S + 4 + s] = F(y.w)[s];
            go[2 * S + 4 + s] = F(y.beta)[s];
            go[3 * S + 4 + s] = F(y.gp)[s];
          }
        }
      }
    }
    __syncthreads();
    const float beta_sum = F(y.red)[60];
    // O3: backward with per-sample row weights
    for (int s = rank; s < p.S; s += G) {
      const float beta = F(y.beta)[s], gp = F(y.gp)[s], wd = F(y.w)[s] * dscale;
      for (int r = tid; r < R; r += NT) F(y.cw)[r] = r < p.M ? gp * F(y.a)[r] : wd;
      sample_theta(s, slab, false, beta, false, nullptr);   // (ends with a block barrier: cw is visible)
      float ps = 0.f, ds = 0.f;
      rows_primal<1>(R, true, gp, ps, ds);
      __syncthreads();
      fold_part(false);
      if (s + G < p.S) __syncthreads();
    }
    push_acc(2);
    __syncthreads();
    for (int m = tid; m < p.M; m += NT) F(y.cw)[m] = F(y.a)[m];   // back to the inner row weights
    cl_sync();
    // O4: owner: dLoss/dphi_T of my slice (no analytic-KL term in the outer objective)
    for (int j = tid; j < slice; j += NT) {
      const int q = j0 + j;
      if (q >= Pt) continue;
      const float g_mu = recv_sum(0, j);
      const float g_rho = F(y.sgm)[q] * (recv_sum(1, j) + beta_sum / F(y.sig)[q]);
      OST(0, j) = g_mu;
      OST(1, j) = g_rho;
      if (p.flags & F_STORE_GOUT) {
        p.gout[q] = g_mu;
        p.gout[Pt + q] = g_rho;
      }
    }
    cl_sync();  // recv may be overwritten by the next phase's pushes only after every owner has read it
  }

  // =================================================================================================================
  // Phase H: a single Hessian-vector pass along gdot (building block / hyper trainer)
  // =================================================================================================================
  if (p.flags & F_HVP) {
    for (int q = tid; q < Pt; q += NT) {
      F(y.gdm)[q] = __ldg(p.gdot + q);
      F(y.gdr)[q] = __ldg(p.gdot + Pt + q);
    }
    __syncthreads();
    float dummy = 0.f;
    for (int s = rank; s < p.S; s += G) inner_pass(s, 0, true, nullptr, dummy);
    push_acc(3);
    cl_sync();
    for (int j = tid; j < slice; j += NT) {
      const int q = j0 + j;
      if (q >= Pt) continue;
      const float sg = F(y.sig)[q], sgm = F(y.sgm)[q], md = F(y.gdm)[q], rd = F(y.gdr)[q];
      const float isg = 1.f / sg;
      p.h_phi[q] = recv_sum(0, j) + md;
      p.h_phi[Pt + q] = sgm * recv_sum(1, j) + sgm * (1.f - sgm) * rd * recv_sum(2, j) +
                        ((1.f + isg * isg) * sgm * sgm + (sg - isg) * sgm * (1.f - sgm)) * rd;
    }
    cl_sync();
  }

  // =================================================================================================================
  // Phase R: reverse sweep through the T Adam steps   (A.4 + A.6; replaces autograd's double backward)
  // =================================================================================================================
  if (p.flags & F_REVERSE) {
    if (p.flags & F_LOAD_GOUT) {
      for (int j = tid; j < slice; j += NT) {
        const int q = j0 + j;
        if (q >= Pt) continue;
        OST(0, j) = p.gout[q];
        OST(1, j) = p.gout[Pt + q];
      }
      if (rank == 0) {
        const int MD = p.M * D;
        for (int i = tid; i < MD; i += NT) F(y.ubar)[i] = p.gout[2 * Pt + i];
        for (int i = tid; i < p.M; i += NT) F(y.abar)[i] = p.gout[2 * Pt + MD + i];
      }
      __syncthreads();
    }
    if (can_pre && tid < n4 && p.T > 0) draw4(rank, p.T - 1, tid, pre);
    for (int t = p.T - 1; t >= 0; --t) {
      const double b1t = pow(B1, (double)(t + 1)), b2t = pow(B2, (double)(t + 1));
      const float k = p.lr / (float)(1.0 - b1t);
      const float sq2 = (float)sqrt(1.0 - b2t);
      const float* tr = p.traj + (size_t)t * 8 * Pt;
      // ---- owner: Adam VJP of my slice -> direction gbar; broadcast gbar and phi_t ----
      for (int j = tid; j < slice; j += NT) {
        const int q = j0 + j;
        if (q >= Pt) continue;
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
        bcast(y.gdm, q, gb[0]);
        bcast(y.gdr, q, gb[1]);
        bcast(y.mu, q, tr[q]);
        bcast(y.rho, q, tr[Pt + q]);
      }
      cl_sync();
      refresh_sigma();
      // ---- every CTA: dual pass over its samples ----
      float dummy = 0.f;
      for (int s = rank; s < p.S; s += G) inner_pass(s, t, true, (can_pre && s == rank) ? pre : nullptr, dummy);
      push_acc(3);
      cl_arrive();
      if (can_pre && tid < n4 && t > 0) draw4(rank, t - 1, tid, pre);   // in the shadow of the barrier
      cl_wait();
      // ---- owner: phibar_t = phibar_{t+1} + H gbar ----
      for (int j = tid; j < slice; j += NT) {
        const int q = j0 + j;
        if (q >= Pt) continue;
        const float sg = F(y.sig)[q], sgm = F(y.sgm)[q], md = F(y.gdm)[q], rd = F(y.gdr)[q];
        const float isg = 1.f / sg;
        OST(0, j) += recv_sum(0, j) + md;
        OST(1, j) += sgm * recv_sum(1, j) + sgm * (1.f - sgm) * rd * recv_sum(2, j) +
                     ((1.f + isg * isg) * sgm * sgm + (sg - isg) * sgm * (1.f - sgm)) * rd;
      }
      // (the next iteration's broadcasts happen before anyone pushes into recv again: the pushes come after the
      //  cluster barrier that follows the broadcasts)
    }
    if (p.g_out) {  // dLoss/dphi_0, useful for diagnostics
      for (int j = tid; j < slice; j += NT) {
        const int q = j0 + j;
        if (q >= Pt) continue;
        p.g_out[q] = OST(0, j);
        p.g_out[Pt + q] = OST(1, j);
      }
    }
  }

  // =================================================================================================================
  // Final: reduce ubar / abar over the cluster (fixed order) and map abar -> v_grad through f
  // =================================================================================================================
  if (p.flags & (F_FINAL | F_STORE_GOUT)) {
    __syncthreads();
    cl_sync();
    if (rank == 0) {
      const int MD = p.M * D;
      float* red = F(y.red);
      for (int i = tid; i < MD + p.M; i += NT) {
        float s = 0.f;
        const int off = i < MD ? y.ubar + i : y.abar + (i - MD);
        for (int c = 0; c < G; ++c) s += *remote(off, c);
        if (p.flags & F_STORE_GOUT) p.gout[2 * Pt + i] = s;
        sm[off] = s;  // rank 0 now holds the totals (remote reads of rank 0 itself happened in this same iteration)
      }
      __syncthreads();
      if (p.flags & F_FINAL) {
        for (int i = tid; i < MD; i += NT) p.u_grad[i] = F(y.ubar)[i];
        if (p.v_grad) {
          if (p.vmode == PSVI_VMODE_IDENTITY || p.roww) {
            for (int m = tid; m < p.M; m += NT) p.v_grad[m] = p.Nf * F(y.abar)[m];
          } else {
            float dot = 0.f;
            for (int m = tid; m < p.M; m += NT) dot += F(y.f)[m] * F(y.abar)[m];
            dot = block_sum(dot, red);
            const float sc = p.Nf * (p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(p.alpha) : 1.f);
            for (int m = tid; m < p.M; m += NT) p.v_grad[m] = sc * F(y.f)[m] * (F(y.abar)[m] - dot);
            if (p.alpha_grad && tid == 0)
              p.alpha_grad[0] = p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? sc * dot : 0.f;
          }
        }
      }
    }
    cl_sync();  // keep every CTA's shared memory alive until rank 0 has read it
  }
}

template <int D, int C, int UPL>
__global__ void __launch_bounds__(NT, 1) psvi_mf_fn1_kernel(const __grid_constant__ EP p) {
  extern __shared__ __align__(16) float smem_dyn[];
  __shared__ FL fl;
  if (threadIdx.x == 0) make_fl<D, C, UPL>(p, fl);
  __syncthreads();
  Fn1<D, C, UPL> e(p, fl, smem_dyn);
  e.run();
}

template <int D, int C, int UPL>
int launch_inst(EP& p, cudaStream_t stream) {
  int dev = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  auto kern = psvi_mf_fn1_kernel<D, C, UPL>;
  const int H = p.dims[1];
  const int Pt = H * (D + 1) + C * (H + 1);
  // cluster size: fewest CTAs that give every CTA the minimal number of samples, capped at 16 (non-portable max)
  int G = p.S < 16 ? p.S : 16;
  const int per = (p.S + G - 1) / G;
  G = (p.S + per - 1) / per;
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (;; --G) {
    if (G < 1) return FN1_NOT_APPLICABLE;
    p.G = G;
    p.slice = (Pt + G - 1) / G;
    p.RC = 0;
    FL fl;
    make_fl<D, C, UPL>(p, fl);
    const size_t smem = (size_t)fl.total * 4;
    if (smem + 1024 > (size_t)smem_max) return FN1_NOT_APPLICABLE;   // too many rows: the generic engine chunks them
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    if (qe != cudaSuccess || ncl < 1) {
      (void)cudaGetLastError();
      if (G == 1) return FN1_NOT_APPLICABLE;
      continue;
    }
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
    if (le != cudaSuccess) {
      (void)cudaGetLastError();
      if (G == 1) {
        psvi_set_error("fn1 engine launch failed: %s", cudaGetErrorString(le));
        return PSVI_ERR_CUDA;
      }
      continue;
    }
    return PSVI_OK;
  }
}

template <int D, int C>
int launch_dc(EP& p, cudaStream_t stream) {
  const int H = p.dims[1];
  if (H <= 64) return launch_inst<D, C, 4>(p, stream);
  if (H <= 112) return launch_inst<D, C, 7>(p, stream);
  if (H <= 128) return launch_inst<D, C, 8>(p, stream);
  return FN1_NOT_APPLICABLE;
}

}  // namespace

namespace psvi_mf {

int psvi_fn1_launch(EP& p, cudaStream_t stream) {
  if (p.L != 2 || (p.flags & F_EVAL)) return FN1_NOT_APPLICABLE;
  if (p.M < 1) return FN1_NOT_APPLICABLE;
  const char* off = getenv("PSVI_DISABLE_FN1");
  if (off != nullptr && off[0] == '1') return FN1_NOT_APPLICABLE;
  const int D = p.dims[0], C = p.dims[2];
  if (D == 2 && C == 2) return launch_dc<2, 2>(p, stream);
  if (D == 2 && C == 4) return launch_dc<2, 4>(p, stream);
  return FN1_NOT_APPLICABLE;
}

}  // namespace psvi_mf
