// psvi_mf_fn1.cu -- shape-specialised cluster engine for the PSVI bilevel step of `fn` with ONE hidden layer and tiny
// input / output widths (BASELINE cfg2: halfmoon / four_blobs, D = 2, H = 100, C = 2 / 4, M = 10..50, S = 10, T = 100).
//
// Same contract, phases and arithmetic as the generic engine (psvi_mf_engine.cu; reference psvi/inference/
// psvi_classes.py:445-600, psvi/robust_higher/optim.py:299-367; math: SURVEY.md Appendix A.1-A.6) -- what changes is
// how the work maps onto the cluster (one CTA per MC sample) and onto an SM:
//   * (D, C, hidden units per lane) are template parameters: every product loop is fully unrolled;
//   * a warp owns two rows at a time; its 32 lanes are 2 row lanes x 16 hidden lanes, a hidden lane owns UPL hidden
//     units (H <= 16 UPL) whose weights it reads as 128-bit words of a conflict-free record layout;
//   * forward, softmax / NLL, backward (or the whole forward-over-reverse dual pass of A.6) for a row happen in
//     registers: the logits (and their tangents / the input adjoints) are the only cross-lane quantities -- four
//     shuffle steps over the 16 hidden lanes; the per-sample weight adjoints  W1bar = abar^T X,  W2bar = obar^T h
//     (and their A.6 counterparts) accumulate in REGISTERS across all rows of the pass and are combined across the
//     two row lanes (one shuffle) and the eight warps (one shared-memory sweep) once per pass;
//   * the variational parameter vector is cut into G slices; CTA k OWNS slice k: phi, the Adam moments, the
//     reverse-sweep carries AND the noise eps_s of its slice for EVERY sample s live only there.  The owner therefore
//     also does the sampling: after its update it pushes theta_s = mu + sigma eps_s (and the tangent) of its slice
//     straight into the weight records of the CTA that runs sample s (st.shared::cluster).  A sample CTA is left
//     with: row pass -> sum the 8 warps' partial adjoints -> push ONE vector (theta_bar_s; two for the dual pass)
//     to the owners, who form  sum_s theta_bar_s  and  sum_s theta_bar_s eps_s  in a fixed order (deterministic).
//     Two cluster barriers per inner step / reverse step; Philox normals and trajectory rows for the next step are
//     produced by otherwise idle warps while the row pass runs.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "psvi_mf_engine.cuh"

namespace cg = cooperative_groups;
using namespace psvi_mf;

namespace {

constexpr int REC = 20;        // floats per hidden-unit record: [0,8) sampled weights, [8,16) tangent weights, 4 pad
#ifndef PSVI_FN1_THREADS
// 256.  Nine warps (288) would finish the 25 row pairs of M = 50 in three rounds of the dual pass instead of four, but the
// register file is split per scheduler: the sub-partition that hosts three warps leaves 168 registers per thread, the dual
// pass (230) spills and the step takes 1.43 ms instead of 1.25 ms (measured, profiles/r2_fn1_engine_ncu_summary.md).
#define PSVI_FN1_THREADS 256
#endif
constexpr int NT1 = PSVI_FN1_THREADS;   // threads per CTA of THIS kernel (the generic engine's NT stays 256)
constexpr int NW1 = NT1 / 32;           // warps per CTA

// block-wide reductions for NW1 warps (the shared helpers of psvi_mf_gemm.cuh are written for 256 threads)
__device__ __forceinline__ float block_sum1(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < NW1) ? red[l] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ double block_sum_d1(double v, float* red_) {
  double* red = reinterpret_cast<double*>(red_ + 16);  // red[16 .. 16 + 2 NW1) as doubles (red is 16-byte aligned)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < NW1; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max1(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < NW1) ? red[l] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

// shared-memory carve-up (offsets in floats)
struct FL {
  int rec, part;                                 // [nper][(HP+1) records] ; NW1 x (HP+1) records of partial adjoints
  int q2r, pdst, pmb;                            // Pt ints: TL index -> float offset inside a record block; Pt uint32:
                                                 // shared::cluster address of recv[0][0][j] in the owner of q; of its mbarrier
  int mb;                                        // two mbarriers: [0] partials received (owner role), [1] weights received
  int mu, rho, sig, sgm, gdm, gdr;               // owner state of my slice: slice floats each
  int ost;                                       // [10][slice]
  int epsS, recv, trs;                           // [2][S][slice] noise ; [S][2][slice] received partials ; [8][slice]
  int tab;                                       // [5][T] step-size tables
  int X, Y, cw;                                  // rows: [R][D], [R] ints, [R]
  int a, f, ubar, abar;                          // coreset: [M], [M], [M][D], [M]
  int lw, e, dsv, w, beta, gp;                   // per-sample scalars (lw, e: doubles)
  int nklp;                                      // [nper][G * NCH] partial sampled-nkl sums
  int red, lossrecv;
  int total;
};

template <int D, int C, int UPL>
__host__ __device__ inline void make_fl(const EP& p, FL& y) {
  constexpr int HP = 16 * UPL;
  const int H = p.dims[1];
  const int Pt = H * (D + 1) + C * (H + 1);
  const int R = p.M + p.B;
  int spad = 32;
  while (spad < p.slice) spad <<= 1;
  const int nper = (p.S + p.G - 1) / p.G, nch = spad / 32;
  int o = 0;
  auto take = [&](int n) {
    int r = o;
    o += (n + 3) & ~3;
    return r;
  };
  y.rec = take(nper * (HP + 1) * REC);
  y.part = take(NW1 * (HP + 1) * REC);
  y.q2r = take(Pt); y.pdst = take(Pt); y.pmb = take(Pt);
  y.mb = take(4);
  y.mu = take(p.slice); y.rho = take(p.slice); y.sig = take(p.slice); y.sgm = take(p.slice);
  y.gdm = take(p.slice); y.gdr = take(p.slice);
  y.ost = take(10 * p.slice);
  y.epsS = take(2 * p.S * p.slice); y.recv = take(p.S * 2 * p.slice); y.trs = take(8 * p.slice);
  y.tab = take(6 * (p.T > 0 ? p.T : 1));
  y.X = take(R * D); y.Y = take(R); y.cw = take(R);
  y.a = take(p.M); y.f = take(p.M); y.ubar = take(p.M * D); y.abar = take(p.M);
  y.lw = take(2 * p.S); y.e = take(2 * p.S); y.dsv = take(p.S); y.w = take(p.S); y.beta = take(p.S); y.gp = take(p.S);
  y.nklp = take(nper * p.G * nch);
  y.red = take(64); y.lossrecv = take(p.G);
  y.total = o;
}

__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_sync() {
  cl_arrive();
  cl_wait();
}
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
// shared::cluster address of the same shared-memory location in CTA `r` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, int r) {
  uint32_t o;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(addr), "r"(r));
  return o;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// Asynchronous remote store that signals the DESTINATION CTA's mbarrier with the bytes written: the receiver waits on its own
// barrier for the byte count it expects, so an exchange costs one one-way DSMEM latency -- no release fence on the sender, no
// cluster-wide barrier.
__device__ __forceinline__ void st_async(uint32_t addr, float v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr),
               "r"(__float_as_uint(v)), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mb_arrive_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
  // try_wait suspends for a bounded time per attempt; a broken exchange traps instead of hanging the GPU
  for (uint32_t it = 0;; ++it) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (it > (1u << 22)) __trap();
  }
}
// sum over the 16 hidden lanes of a row lane group (lane bits 0..3); every lane ends with the total
__device__ __forceinline__ float hl_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}
// sigma = softplus(rho) (F.softplus, threshold 20; neural_net.py:131) and sigmoid(rho) from ONE exponential:
// y = e^rho, sigmoid = y / (1 + y), softplus = log1p(y) = 2 atanh(z) with z = y / (2 + y) (series to z^11: relative
// truncation error < 3e-13 for y < 1/4, i.e. sigma < 0.22; log1pf beyond).  Shortens the owner's dependent chain.
__device__ __forceinline__ void softplus_sigmoid(float r, float& sp, float& sg) {
  if (r > 20.f) {
    sp = r;
    sg = 1.f / (1.f + expf(-r));
    return;
  }
  const float yv = expf(r);
  sg = yv * __frcp_rn(1.f + yv);
  if (yv < 0.25f) {
    const float z = yv * __frcp_rn(2.f + yv), z2 = z * z;
    float pl = fmaf(z2, 1.f / 11.f, 1.f / 9.f);
    pl = fmaf(z2, pl, 1.f / 7.f);
    pl = fmaf(z2, pl, 1.f / 5.f);
    pl = fmaf(z2, pl, 1.f / 3.f);
    pl = fmaf(z2, pl, 1.f);
    sp = 2.f * z * pl;
  } else {
    sp = log1pf(yv);
  }
}

template <int D, int C, int UPL>
struct Fn1 {
  static constexpr int HP = 16 * UPL;
  static constexpr int PS = (HP + 1) * REC;  // floats of one record block (one sample's weights / one warp's partials)
  static_assert(D + 1 + C <= 8, "a hidden unit's weights must fit one 8-float record half");

  const EP& p;
  const FL& y;
  float* sm;
  cg::cluster_group cluster;
  int rank, tid, warp, lane, hl, rl;
  int H, Pt, slice, j0, nch, spad_sh;
  bool own;   // this thread owns TL index q = j0 + tid (load / store phases)
  int oj, oc; // update phases: thread (oc, oj) works on TL index j0 + oj, component oc (0 = mu, 1 = rho)
  bool own2;
  uint32_t par_recv = 0, par_rec = 0;   // phase parities of my two mbarriers
  int n_own, n_loc;                     // valid TL indices of my slice; MC samples this CTA runs

  __device__ Fn1(const EP& p_, const FL& y_, float* s_) : p(p_), y(y_), sm(s_), cluster(cg::this_cluster()) {
    rank = (int)cluster.block_rank();
    tid = threadIdx.x;
    warp = tid >> 5;
    lane = tid & 31;
    hl = lane & 15;
    rl = lane >> 4;
    H = p.dims[1];
    Pt = H * (D + 1) + C * (H + 1);
    slice = p.slice;
    j0 = rank * slice;
    spad_sh = 5;
    while ((1 << spad_sh) < slice) ++spad_sh;   // sampling phase: threads are (sample group, index in padded slice)
    nch = (1 << spad_sh) >> 5;
    own = tid < slice && j0 + tid < Pt;
    oj = tid & ((1 << spad_sh) - 1);     // the two components sit in different warps: no divergence inside a warp
    oc = tid >> spad_sh;
    own2 = oc < 2 && oj < slice && j0 + oj < Pt;
    n_own = max(0, min(slice, Pt - j0));
    n_loc = rank < p.S ? (p.S - rank + p.G - 1) / p.G : 0;
  }
  __device__ __forceinline__ float* F(int off) const { return sm + off; }
  __device__ __forceinline__ int* I(int off) const { return reinterpret_cast<int*>(sm + off); }
  __device__ __forceinline__ float* remote(int off, int r) { return cluster.map_shared_rank(sm + off, r); }
  __device__ __forceinline__ float& OST(int row) { return sm[y.ost + row * slice + tid]; }
  __device__ __forceinline__ float& OS2(int row) { return sm[y.ost + row * slice + oj]; }

  int tl_n = 0;
  __device__ __forceinline__ void stamp(int code) {
    if (p.tl != nullptr && tid == 0 && rank == 0 && tl_n < 4096) {
      p.tl[2 * tl_n] = code;
      p.tl[2 * tl_n + 1] = clock64();
      ++tl_n;
    }
  }

  // ---- owner role: wait until every sample CTA's partials of this exchange have landed in recv (ncomp vectors per sample;
  //      rank 0 also receives the G loss partials when the inner losses are logged) -----------------------------------------
  __device__ __forceinline__ void wait_recv(int ncomp, bool with_loss) {
    const uint32_t bar = smem_u32(sm + y.mb);
    if (tid == 0) mb_arrive_expect(bar, (uint32_t)((p.S * ncomp * n_own + ((with_loss && rank == 0) ? p.G : 0)) * 4));
    mb_wait(bar, par_recv);
    par_recv ^= 1u;
  }
  // ---- sample role: wait until every owner's slice of my samples' weights (and tangents / nkl partials) has landed ---------
  __device__ __forceinline__ void wait_rec(bool tangent, bool nkl) {
    const uint32_t bar = smem_u32(sm + y.mb + 2);
    if (tid == 0) mb_arrive_expect(bar, (uint32_t)(n_loc * (Pt * (tangent ? 2 : 1) + (nkl ? p.G * nch : 0)) * 4));
    mb_wait(bar, par_rec);
    par_rec ^= 1u;
  }

  // ---- standard normals of my slice for EVERY sample of noise slab `slab` -> epsS[slab & 1]; done by the LAST threads
  //      of the CTA (the first warps are the owner threads and carry the longest row loops) ---------------------------
  __device__ void gen_eps(int slab) {
    const int nb = slice >> 2;  // Philox blocks per sample (slice is a multiple of 4)
    const int items = p.S * nb;
    float* dst = F(y.epsS) + (slab & 1) * p.S * slice;
    for (int i = NT1 - 1 - tid; i < items; i += NT1) {
      const int s = i / nb, b = i - s * nb;
      const int q4 = (j0 >> 2) + b;
      float e4[4] = {0.f, 0.f, 0.f, 0.f};
      if (4 * q4 < Pt) {
        if (p.noise_mode == PSVI_NOISE_PHILOX) {
          philox_normal4(p.seed, p.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)q4, e4);
        } else {
          const float* src = p.eps + ((size_t)slab * p.S + s) * Pt;
#pragma unroll
          for (int k = 0; k < 4; ++k) e4[k] = (4 * q4 + k < Pt) ? __ldg(src + 4 * q4 + k) : 0.f;
        }
      }
      *reinterpret_cast<float4*>(dst + s * slice + 4 * b) = make_float4(e4[0], e4[1], e4[2], e4[3]);
    }
  }

  // ---- owner: theta_s = mu + sigma eps_s (and the tangent  mudot + sigmoid(rho) rhodot eps_s) of my slice for every
  //      sample, written into the records of the CTA that runs the sample.  All threads: thread = (sample group, index
  //      in the padded slice).  want_nkl: also the slice's share of  sampled_nkl_s = sum_i [-theta^2/2 + eps^2/2 +
  //      log sigma]  (neural_net.py:110-115), one partial per warp. ---------------------------------------------------
  __device__ void owner_sample(int slab, bool tangent, bool want_nkl) {
    const int j = tid & ((1 << spad_sh) - 1), sg0 = tid >> spad_sh, nsg = NT1 >> spad_sh;
    const bool ok = j < slice && j0 + j < Pt;
    const float* ep = F(y.epsS) + (slab & 1) * p.S * slice + j;
    const float mu = ok ? F(y.mu)[j] : 0.f, sg = ok ? F(y.sig)[j] : 1.f;
    const float md = (ok && tangent) ? F(y.gdm)[j] : 0.f;
    const float rs = (ok && tangent) ? F(y.sgm)[j] * F(y.gdr)[j] : 0.f;
    const float lsg = want_nkl ? logf(sg) : 0.f;
    const uint32_t base = smem_u32(F(y.rec) + (ok ? I(y.q2r)[j0 + j] : 0));
    const uint32_t nbase = smem_u32(F(y.nklp) + rank * nch + (j >> 5));
    const uint32_t mbar_rec = smem_u32(sm + y.mb + 2);
    int cta = sg0 % p.G, li = sg0 / p.G;
    const int dcta = nsg % p.G, dli = nsg / p.G;
    // (warp-uniform trip count: a warp lies inside one sample group; threads past the last full group sit out)
    for (int s = sg0 < nsg ? sg0 : p.S; s < p.S; s += nsg) {
      const float e = ok ? ep[s * slice] : 0.f;
      const float th = fmaf(sg, e, mu);
      const uint32_t rbar = mapa(mbar_rec, cta);
      if (ok) {
        const uint32_t dst = mapa(base + (uint32_t)(li * PS * 4), cta);
        st_async(dst, th, rbar);
        if (tangent) st_async(dst + 32, fmaf(rs, e, md), rbar);
      }
      if (want_nkl) {
        float v = ok ? (-0.5f * th * th + 0.5f * e * e + lsg) : 0.f;
        v = warp_sum(v);
        if (lane == 0) st_async(mapa(nbase + (uint32_t)(li * p.G * nch * 4), cta), v, rbar);
      }
      cta += dcta; li += dli;
      if (cta >= p.G) { cta -= p.G; ++li; }
    }
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

  // rows in flight per lane: independent rows are interleaved to hide the shuffle / MUFU / FMA latencies of a row's
  // serial chain (forward -> logits reduction -> softmax -> backward); bounded by the register file
  static constexpr int KW = D + 1 + C;
  static constexpr int RU_P = (UPL * KW <= 35) ? 4 : 2;
  static constexpr int RU_D = 1;
  static constexpr bool HOIST_D = (UPL * KW <= 35);   // dual pass: both weight sets stay in registers

  // ---- primal pass over rows [0, R): MODE 0 = values only, 1 = gradient.  Row weights in cw[]; sumA += sum over pseudo
  //      rows of a_m nll, sumD += sum over data rows of nll (held by the hl == 0 lanes; only if NLL).  need_x: input
  //      adjoints of the pseudo rows -> ubar, and abar += gp * nll.  Leaves the per-warp weight adjoints in `part`. ----
  template <int MODE, bool NLL>
  __device__ void rows_primal(const float* rec, int R, bool need_x, float gp, float& sumA, float& sumD) {
    constexpr int RU = RU_P;
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

    for (int rb = 2 * RU * warp; rb < R; rb += 2 * RU * NW1) {  // warp-uniform trip count (the shuffles need the full warp)
      int r[RU], yl[RU];
      bool ok[RU];
      float x[RU][D], h[RU][UPL], o[RU][C];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        r[u] = rb + 2 * u + rl;
        ok[u] = r[u] < R;
#pragma unroll
        for (int d = 0; d < D; ++d) x[u][d] = ok[u] ? X[r[u] * D + d] : 0.f;
        yl[u] = ok[u] ? Y[r[u]] : 0;
#pragma unroll
        for (int c = 0; c < C; ++c) o[u][c] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          float a = b1[i];
#pragma unroll
          for (int d = 0; d < D; ++d) a = fmaf(w1[i][d], x[u][d], a);
          h[u][i] = fmaxf(a, 0.f);
#pragma unroll
          for (int c = 0; c < C; ++c) o[u][c] = fmaf(h[u][i], w2[i][c], o[u][c]);
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c) o[u][c] = hl_sum(o[u][c]) + b2[c];
      float nll[RU], ob[RU][C];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        float mx = o[u][0];
#pragma unroll
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, o[u][c]);
        float se = 0.f, ex[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { ex[c] = expf(o[u][c] - mx); se += ex[c]; }
        nll[u] = 0.f;
        if (NLL) {   // the value is needed only by the outer objective and by the logged inner losses
          float oy = o[u][0];
#pragma unroll
          for (int c = 1; c < C; ++c) oy = (yl[u] == c) ? o[u][c] : oy;
          nll[u] = mx + logf(se) - oy;
          if (hl == 0 && ok[u]) {
            if (r[u] < p.M) sumA += av[r[u]] * nll[u]; else sumD += nll[u];
          }
        }
        if (MODE == 1) {
          const float cwr = ok[u] ? cw[r[u]] : 0.f;
          const float inv = 1.f / se;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            ob[u][c] = cwr * (ex[c] * inv - (c == yl[u] ? 1.f : 0.f));
            gb2[c] += ob[u][c];
          }
        }
      }
      if (MODE == 1) {
        float xb[RU][D];
#pragma unroll
        for (int u = 0; u < RU; ++u)
#pragma unroll
          for (int d = 0; d < D; ++d) xb[u][d] = 0.f;
#pragma unroll
        for (int i = 0; i < UPL; ++i) {
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            float t = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              gw2[i][c] = fmaf(ob[u][c], h[u][i], gw2[i][c]);
              t = fmaf(ob[u][c], w2[i][c], t);
            }
            const float ab = h[u][i] > 0.f ? t : 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) gw1[i][d] = fmaf(ab, x[u][d], gw1[i][d]);
            gb1[i] += ab;
            if (need_x) {
#pragma unroll
              for (int d = 0; d < D; ++d) xb[u][d] = fmaf(ab, w1[i][d], xb[u][d]);
            }
          }
        }
        if (need_x && rb < p.M) {  // (rb is warp-uniform)
#pragma unroll
          for (int u = 0; u < RU; ++u)
#pragma unroll
            for (int d = 0; d < D; ++d) xb[u][d] = hl_sum(xb[u][d]);
          if (hl == 0) {
#pragma unroll
            for (int u = 0; u < RU; ++u)
              if (ok[u] && r[u] < p.M) {
#pragma unroll
                for (int d = 0; d < D; ++d) F(y.ubar)[r[u] * D + d] += xb[u][d];
                F(y.abar)[r[u]] += gp * nll[u];
              }
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
  __device__ void rows_dual(const float* rec) {
    constexpr int RU = RU_D;
    constexpr bool HOIST = HOIST_D;
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

    float wh[HOIST ? UPL : 1][8], wdh[HOIST ? UPL : 1][8];
    if (HOIST) {
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        load8(rec + (hl + 16 * i) * REC, wh[i]);
        load8(rec + (hl + 16 * i) * REC + 8, wdh[i]);
      }
    }

    for (int rb = 2 * RU * warp; rb < R; rb += 2 * RU * NW1) {
      int r[RU], yl[RU];
      bool ok[RU];
      float x[RU][D], cwr[RU], h[RU][UPL], hd[RU][UPL], o[RU][C], od[RU][C];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        r[u] = rb + 2 * u + rl;
        ok[u] = r[u] < R;
#pragma unroll
        for (int d = 0; d < D; ++d) x[u][d] = ok[u] ? X[r[u] * D + d] : 0.f;
        yl[u] = ok[u] ? Y[r[u]] : 0;
        cwr[u] = ok[u] ? cw[r[u]] : 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { o[u][c] = 0.f; od[u][c] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        float w[8], wd[8];
        if (HOIST) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { w[k] = wh[i][k]; wd[k] = wdh[i][k]; }
        } else {
          load8(rec + (hl + 16 * i) * REC, w);
          load8(rec + (hl + 16 * i) * REC + 8, wd);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          float a = w[D], ad = wd[D];
#pragma unroll
          for (int d = 0; d < D; ++d) { a = fmaf(w[d], x[u][d], a); ad = fmaf(wd[d], x[u][d], ad); }
          const bool k = a > 0.f;
          h[u][i] = k ? a : 0.f;
          hd[u][i] = k ? ad : 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            o[u][c] = fmaf(h[u][i], w[D + 1 + c], o[u][c]);
            od[u][c] = fmaf(hd[u][i], w[D + 1 + c], fmaf(h[u][i], wd[D + 1 + c], od[u][c]));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c) { o[u][c] = hl_sum(o[u][c]) + b2[c]; od[u][c] = hl_sum(od[u][c]) + b2d[c]; }
      float Aod[RU][C], Ao[RU][C], ac[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        float mx = o[u][0];
#pragma unroll
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, o[u][c]);
        float se = 0.f, pc[C], pd = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { pc[c] = expf(o[u][c] - mx); se += pc[c]; }
        const float inv = 1.f / se;
#pragma unroll
        for (int c = 0; c < C; ++c) { pc[c] *= inv; pd = fmaf(pc[c], od[u][c], pd); }
        ac[u] = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float qc = pc[c] - (c == yl[u] ? 1.f : 0.f);
          Aod[u][c] = cwr[u] * qc;                       // adjoint of odot
          Ao[u][c] = cwr[u] * pc[c] * (od[u][c] - pd);   // adjoint of o
          ac[u] = fmaf(qc, od[u][c], ac[u]);
          gb2[c] += Ao[u][c];
          gb2d[c] += Aod[u][c];
        }
      }
      float ax[RU][D];
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int d = 0; d < D; ++d) ax[u][d] = 0.f;
#pragma unroll
      for (int i = 0; i < UPL; ++i) {
        float w[8], wd[8];
        if (HOIST) {
#pragma unroll
          for (int k = 0; k < 8; ++k) { w[k] = wh[i][k]; wd[k] = wdh[i][k]; }
        } else {
          load8(rec + (hl + 16 * i) * REC, w);
          load8(rec + (hl + 16 * i) * REC + 8, wd);
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          float Ahd = 0.f, Ah = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            Ahd = fmaf(Aod[u][c], w[D + 1 + c], Ahd);
            Ah = fmaf(Aod[u][c], wd[D + 1 + c], fmaf(Ao[u][c], w[D + 1 + c], Ah));
            g[i][D + 1 + c] = fmaf(Aod[u][c], hd[u][i], fmaf(Ao[u][c], h[u][i], g[i][D + 1 + c]));
            gd[i][D + 1 + c] = fmaf(Aod[u][c], h[u][i], gd[i][D + 1 + c]);
          }
          const bool k = h[u][i] > 0.f;
          const float Aa = k ? Ah : 0.f, Aad = k ? Ahd : 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            g[i][d] = fmaf(Aa, x[u][d], g[i][d]);
            gd[i][d] = fmaf(Aad, x[u][d], gd[i][d]);
            ax[u][d] = fmaf(Aa, w[d], fmaf(Aad, wd[d], ax[u][d]));
          }
          g[i][D] += Aa;
          gd[i][D] += Aad;
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u)
#pragma unroll
        for (int d = 0; d < D; ++d) ax[u][d] = hl_sum(ax[u][d]);
      if (hl == 0) {
#pragma unroll
        for (int u = 0; u < RU; ++u)
          if (ok[u]) {
#pragma unroll
            for (int d = 0; d < D; ++d) F(y.ubar)[r[u] * D + d] += ax[u][d];
            F(y.abar)[r[u]] += ac[u];
          }
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

  // ---- sum the eight warps' partial adjoints of sample s and push them to the slice owners ---------------------------
  __device__ void fold_push(int s, int ncomp) {
    const int* q2r = I(y.q2r);
    const uint32_t* pdst = reinterpret_cast<const uint32_t*>(sm + y.pdst);
    const uint32_t* pmb = reinterpret_cast<const uint32_t*>(sm + y.pmb);
    const float* part = F(y.part);
    const uint32_t soff = (uint32_t)(s * 2 * slice * 4);
    for (int q = tid; q < Pt; q += NT1) {
      const int r = q2r[q];
      float v[NW1];
#pragma unroll
      for (int w = 0; w < NW1; ++w) v[w] = part[w * PS + r];
      float s0 = v[0];
#pragma unroll
      for (int w = 1; w < NW1; ++w) s0 += v[w];
      const uint32_t dst = pdst[q] + soff, obar = pmb[q];
      st_async(dst, s0, obar);
      if (ncomp > 1) {
#pragma unroll
        for (int w = 0; w < NW1; ++w) v[w] = part[w * PS + r + 8];
        float s1 = v[0];
#pragma unroll
        for (int w = 1; w < NW1; ++w) s1 += v[w];
        st_async(dst + (uint32_t)(slice * 4), s1, obar);
      }
    }
  }

  // ---- owner: sums over the samples of the received partials of TL index j, in a fixed order (deterministic):
  //      plain sum of component `comp`, or the sum weighted by my eps_s ----------------------------------------------
  __device__ __forceinline__ float recv_sum(int j, int comp) const {
    const float* r = sm + y.recv + comp * slice + j;
    float tot = 0.f;
    for (int s0 = 0; s0 < p.S; s0 += 16) {   // 16 independent loads, then a fixed-shape tree
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {   // branch-free: clamped (always in-bounds) load, then select
        const int sk = min(s0 + k, p.S - 1);
        const float a = r[(2 * sk) * slice];
        v[k] = (s0 + k < p.S) ? a : 0.f;
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) v[k] += v[k + w];
      tot += v[0];
    }
    return tot;
  }
  __device__ __forceinline__ float recv_sum_eps(int j, int comp, int slab) const {
    const float* r = sm + y.recv + comp * slice + j;
    const float* ep = sm + y.epsS + (slab & 1) * p.S * slice + j;
    float tot = 0.f;
    for (int s0 = 0; s0 < p.S; s0 += 16) {
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int sk = min(s0 + k, p.S - 1);
        const float a = r[(2 * sk) * slice] * ep[sk * slice];
        v[k] = (s0 + k < p.S) ? a : 0.f;
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) v[k] += v[k + w];
      tot += v[0];
    }
    return tot;
  }

  // ---- coreset weights a = N f(v)  (psvi_classes.py:476,505; f per class :111,:1358,:1486) ----------------------------
  __device__ void setup_coreset() {
    float* a = F(y.a);
    float* f = F(y.f);
    const int M = p.M;
    if (p.roww) {
      for (int m = tid; m < M; m += NT1) { a[m] = __ldg(p.roww + m); f[m] = 0.f; }
      __syncthreads();
      return;
    }
    if (p.vmode == PSVI_VMODE_IDENTITY) {
      for (int m = tid; m < M; m += NT1) { f[m] = __ldg(p.v + m); a[m] = p.Nf * f[m]; }
      __syncthreads();
      return;
    }
    float mx = -INFINITY;
    for (int m = tid; m < M; m += NT1) mx = fmaxf(mx, __ldg(p.v + m));
    mx = block_max1(mx, F(y.red));
    float se = 0.f;
    for (int m = tid; m < M; m += NT1) se += expf(__ldg(p.v + m) - mx);
    se = block_sum1(se, F(y.red));
    const float sc = p.Nf * (p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(p.alpha) : 1.f);
    for (int m = tid; m < M; m += NT1) {
      f[m] = expf(__ldg(p.v + m) - mx) / se;
      a[m] = sc * f[m];
    }
    __syncthreads();
  }


  __device__ void init() {
    for (int i = tid; i < y.total; i += NT1) sm[i] = 0.f;
    __syncthreads();
    int* q2r = I(y.q2r);
    uint32_t* pdst = reinterpret_cast<uint32_t*>(sm + y.pdst);
    const int HD = H * D;
    const uint32_t recv0 = smem_u32(F(y.recv)), mb0 = smem_u32(sm + y.mb);
    uint32_t* pmb = reinterpret_cast<uint32_t*>(sm + y.pmb);
    for (int q = tid; q < Pt; q += NT1) {
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
      const int owner = q / slice;
      pdst[q] = mapa(recv0 + (uint32_t)((q - owner * slice) * 4), owner);
      pmb[q] = mapa(mb0, owner);
    }
    if (tid == 0) {
      mb_init(mb0, 1);
      mb_init(mb0 + 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (own) {
      const float r = p.rho[j0 + tid];
      F(y.mu)[tid] = p.mu[j0 + tid];
      F(y.rho)[tid] = r;
      softplus_sigmoid(r, F(y.sig)[tid], F(y.sgm)[tid]);
    }
    for (int i = tid; i < p.M * D; i += NT1) F(y.X)[i] = __ldg(p.u + i);
    for (int m = tid; m < p.M; m += NT1) I(y.Y)[m] = __ldg(p.z + m);
    for (int i = tid; i < p.B * D; i += NT1) F(y.X)[p.M * D + i] = __ldg(p.xb + i);
    for (int b = tid; b < p.B; b += NT1) I(y.Y)[p.M + b] = __ldg(p.yb + b);
    // step-size tables.  forward (rows 0..2): the reference's running products beta^t in double, as the generic engine
    // forms them; reverse (rows 3, 4): lr / (1 - beta1^(t+1)) and sqrt(1 - beta2^(t+1)) from pow().
    {
      const double B1 = 0.9, B2 = 0.999;
      float* tab = F(y.tab);
      for (int t = tid; t < p.T; t += NT1) {
        double a = pow(B1, (double)p.step0), b = pow(B2, (double)p.step0);
        for (int i = 0; i <= t; ++i) { a *= B1; b *= B2; }
        tab[t] = (float)(1.0 - a);
        tab[5 * p.T + t] = p.lr / (float)(1.0 - a);
        tab[p.T + t] = (float)sqrt(1.0 - b);
        tab[2 * p.T + t] = (float)(1.0 - b);
        const double b1t = pow(B1, (double)(t + 1)), b2t = pow(B2, (double)(t + 1));
        tab[3 * p.T + t] = p.lr / (float)(1.0 - b1t);
        tab[4 * p.T + t] = (float)sqrt(1.0 - b2t);
      }
    }
    __syncthreads();
    if (p.M > 0) setup_coreset();
    for (int m = tid; m < p.M; m += NT1) F(y.cw)[m] = F(y.a)[m];   // inner passes weight the rows by a_m
    if (p.adam_m && (p.flags & F_UNROLL) && own) {
      const int q = j0 + tid;
      OST(6) = p.adam_m[q];
      OST(7) = p.adam_m[Pt + q];
      OST(8) = p.adam_v[q];
      OST(9) = p.adam_v[Pt + q];
    }
    __syncthreads();
    cl_sync();  // every CTA's shared memory is initialised before anybody pushes into it
  }

  // ---- owner: Adam VJP of step t (A.4) for component oc of TL index oj, from a trajectory row (phi, g, m, v of both
  //      components at stride `stride`) -> direction gbar; makes phi_t the current parameters of my slice ---------------
  __device__ void owner_vjp(int t, const float* tr, int stride) {
    if (!own2) return;
    const float b1 = 0.9f, b2 = 0.999f, omb1 = (float)(1.0 - 0.9), omb2 = (float)(1.0 - 0.999), aeps = 1e-8f;
    const float k = F(y.tab)[3 * p.T + t], sq2 = F(y.tab)[4 * p.T + t];
    const int c = oc;
    const float g = tr[(2 + c) * stride], m = tr[(4 + c) * stride], v = tr[(6 + c) * stride];
    const float pb = OS2(0 + c);
    const float qd = sqrtf(v + 1e-8f);
    const float den = qd / sq2 + aeps;
    const float mbar = OS2(2 + c) - k * pb / den;
    const float denbar = k * pb * m / (den * den);
    float vbar = OS2(4 + c) + denbar / (2.f * qd * sq2);
    if (v == 0.f) vbar = 0.f;  // _maybe_mask hook (optim.py:40-52,346-347)
    const float gb = omb1 * mbar + 2.f * omb2 * g * vbar;
    OS2(2 + c) = b1 * mbar;
    OS2(4 + c) = b2 * vbar;
    const float ph = tr[c * stride];
    if (c == 0) {
      F(y.gdm)[oj] = gb;
      F(y.mu)[oj] = ph;
    } else {
      F(y.gdr)[oj] = gb;
      F(y.rho)[oj] = ph;
      softplus_sigmoid(ph, F(y.sig)[oj], F(y.sgm)[oj]);
    }
  }

  // ---- owner: component oc of (H gbar) at TL index oj from the received dual-pass partials (A.6, last line) -----------
  __device__ __forceinline__ float owner_hvp(int slab) {
    if (oc == 0) return recv_sum(oj, 0) + F(y.gdm)[oj];
    const float Ae = recv_sum_eps(oj, 0, slab), Ade = recv_sum_eps(oj, 1, slab);
    const float sg = F(y.sig)[oj], sgm = F(y.sgm)[oj], rd = F(y.gdr)[oj];
    const float isg = 1.f / sg;
    return sgm * Ae + sgm * (1.f - sgm) * rd * Ade + ((1.f + isg * isg) * sgm * sgm + (sg - isg) * sgm * (1.f - sgm)) * rd;
  }

  // ---- owner: gradient of component oc of TL index oj after inner step t, Adam arithmetic (three flavours), new
  //      parameters; trv = (phi, g, m, v) of the component for the trajectory -------------------------------------------
  __device__ void owner_forward(int t, float trv[4]) {
    const float b1 = 0.9f, b2 = 0.999f, omb1 = (float)(1.0 - 0.9), omb2 = (float)(1.0 - 0.999), aeps = 1e-8f;
    if (own2) {
      const int c = oc;
      const float sg = F(y.sig)[oj], sgm = F(y.sgm)[oj];
      const float pv = c == 0 ? F(y.mu)[oj] : F(y.rho)[oj];
      const float g = c == 0 ? recv_sum(oj, 0) + pv : sgm * (recv_sum_eps(oj, 0, t) + (sg - 1.f / sg));
      const float bc1 = F(y.tab)[t], sq2 = F(y.tab)[p.T + t], bc2 = F(y.tab)[2 * p.T + t];
      const float step_size = p.lr / bc1;
      float m = OS2(6 + c) * b1 + omb1 * g;
      float v = OS2(8 + c) * b2 + omb2 * g * g;
      float den, np;
      if (p.adam_mode == PSVI_ADAM_ROBUST_HIGHER) {
        den = sqrtf(v + 1e-8f) / sq2 + aeps;                     // optim.py:346-363
        np = pv - step_size * (m / den);
      } else if (p.adam_mode == PSVI_ADAM_TORCH) {
        den = sqrtf(v) / sq2 + aeps;
        np = pv - step_size * (m / den);
      } else {                                                    // hypergrad/diff_optimizers.py:184-213
        v += 1e-12f;
        den = sqrtf(v / bc2) + aeps;
        np = pv - p.lr * (m / bc1 / den);
      }
      trv[0] = pv; trv[1] = g; trv[2] = m; trv[3] = v;
      if (!(p.flags & F_NOUPDATE)) {
        OS2(6 + c) = m;
        OS2(8 + c) = v;
        if (c == 0) {
          F(y.mu)[oj] = np;
        } else {
          F(y.rho)[oj] = np;
          softplus_sigmoid(np, F(y.sig)[oj], F(y.sgm)[oj]);
        }
      }
    }
  }

  __device__ void run();
};

template <int D, int C, int UPL>
__device__ void Fn1<D, C, UPL>::run() {
  const int G = p.G, S = p.S;
  const int q = j0 + tid;   // my TL index as an owner thread (valid iff own)
  init();

  const float b1 = 0.9f, b2 = 0.999f, omb1 = (float)(1.0 - 0.9), omb2 = (float)(1.0 - 0.999), aeps = 1e-8f;
  const bool want_loss = p.inner_losses != nullptr;
  const float* rec0 = F(y.rec);

  // =================================================================================================================
  // Phase U: T inner Adam steps  (psvi_classes.py:549-555 ; optim.py:224-229,303-367)
  // =================================================================================================================
  if ((p.flags & F_UNROLL) && p.T > 0) {
    gen_eps(0);
    __syncthreads();
    owner_sample(0, false, false);
    wait_rec(false, false);
    for (int t = 0; t < p.T; ++t) {
      const bool follow = !(p.flags & F_NOUPDATE) && (t + 1 < p.T || (p.flags & F_OUTER));
      stamp(1);
      if (follow) gen_eps(t + 1);
      float lpart = 0.f, dsum = 0.f;
      int li = 0;
      for (int s = rank; s < S; s += G, ++li) {
        if (want_loss) rows_primal<1, true>(rec0 + li * PS, p.M, false, 0.f, lpart, dsum);
        else rows_primal<1, false>(rec0 + li * PS, p.M, false, 0.f, lpart, dsum);
        __syncthreads();
        fold_push(s, 1);
        if (s + G < S) __syncthreads();
      }
      stamp(2);
      if (want_loss) {
        // KL(q||p) of my slice at phi_t (neural_net.py:101-108), added once (Q1)
        if (own) {
          const float sg = F(y.sig)[tid], m = F(y.mu)[tid];
          lpart += 0.5f * (sg * sg + m * m - 1.f) - logf(sg);
        }
        lpart = block_sum1(lpart, F(y.red));
        if (tid == 0) st_async(mapa(smem_u32(F(y.lossrecv) + rank), 0), lpart, mapa(smem_u32(sm + y.mb), 0));
      }
      wait_recv(1, want_loss);
      stamp(3);
      if (want_loss && rank == 0 && tid == 0) {
        float sl = 0.f;
        for (int c = 0; c < G; ++c) sl += F(y.lossrecv)[c];
        p.inner_losses[t] = sl;
      }
      // ---- owner: gradient of component oc of TL index oj, Adam, new parameters; then the next sample ----
      float trv[4] = {0.f, 0.f, 0.f, 0.f};
      owner_forward(t, trv);
      stamp(31);
      if (follow) {
        __syncthreads();
        owner_sample(t + 1, false, t + 1 == p.T);
      }
      stamp(4);
      if (own2) {   // trajectory (global memory): nobody in the cluster waits for these stores
        const int q2 = j0 + oj;
        if (p.g_out) p.g_out[oc * Pt + q2] = trv[1];
        if (p.traj) {
          float* tr = p.traj + (size_t)t * 8 * Pt + q2;
#pragma unroll
          for (int k = 0; k < 4; ++k) tr[(2 * k + oc) * Pt] = trv[k];
        }
      }
      stamp(33);
      if (follow) wait_rec(false, t + 1 == p.T);
      stamp(5);
    }
    __syncthreads();   // (owner state was written with the update mapping, read below with the load / store mapping)
    if ((p.flags & F_WRITE_PHI) && own) {
      p.mu[q] = F(y.mu)[tid];
      p.rho[q] = F(y.rho)[tid];
      if (p.adam_m) {
        p.adam_m[q] = OST(6); p.adam_m[Pt + q] = OST(7);
        p.adam_v[q] = OST(8); p.adam_v[Pt + q] = OST(9);
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
    if (!((p.flags & F_UNROLL) && p.T > 0)) {   // (otherwise the last inner step has already delivered the sample)
      gen_eps(slab);
      __syncthreads();
      owner_sample(slab, false, true);
      wait_rec(false, true);
    }
    // O1: per-sample p_s, d_s, nkl_s
    {
      int li = 0;
      for (int s = rank; s < S; s += G, ++li) {
        float ps = 0.f, ds = 0.f;
        rows_primal<0, true>(rec0 + li * PS, R, false, 0.f, ps, ds);
        // the S per-sample sums are O(N) while the importance-weight adjoints depend on their *differences*:
        // reduce and keep them in double (the fp32 reference loses ~2 digits here at init_sd=1e-6)
        const double ps_d = block_sum_d1((double)ps, F(y.red));
        const double ds_d = block_sum_d1((double)ds, F(y.red)) * (double)dscale;
        double nkl_d = 0.0;
        for (int k = 0; k < G * nch; ++k) nkl_d += (double)F(y.nklp)[li * G * nch + k];
        if (tid < G) {
          reinterpret_cast<double*>(remote(y.lw, tid))[s] = -ps_d + nkl_d;
          reinterpret_cast<double*>(remote(y.e, tid))[s] = ds_d - (double)p.kappa * ps_d;
          remote(y.dsv, tid)[s] = (float)ds_d;
        }
      }
    }
    cl_sync();
    // O2: importance weights and adjoint seeds (every CTA, redundantly; S is tiny)
    if (tid == 0) {
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
    {
      int li = 0;
      for (int s = rank; s < S; s += G, ++li) {
        const float gp = F(y.gp)[s], wd = F(y.w)[s] * dscale;
        for (int r = tid; r < R; r += NT1) F(y.cw)[r] = r < p.M ? gp * F(y.a)[r] : wd;
        __syncthreads();
        float ps = 0.f, ds = 0.f;
        rows_primal<1, true>(rec0 + li * PS, R, true, gp, ps, ds);
        __syncthreads();
        fold_push(s, 1);
        if (s + G < S) __syncthreads();
      }
    }
    __syncthreads();
    for (int m = tid; m < p.M; m += NT1) F(y.cw)[m] = F(y.a)[m];   // back to the inner row weights
    wait_recv(1, false);
    // O4: owner: dLoss/dphi_T of my TL index (no analytic-KL term in the outer objective); the sampled-nkl part
    // d nkl_s / d theta = -theta_s, weighted by beta_s (A.2), is formed here from the owner's own eps_s
    if (own) {
      float A = recv_sum(tid, 0), Ae = recv_sum_eps(tid, 0, slab);
      const float mu = F(y.mu)[tid], sg = F(y.sig)[tid];
      const float* ep = F(y.epsS) + (slab & 1) * S * slice + tid;
      for (int s = 0; s < S; ++s) {
        const float e = ep[s * slice];
        const float tb = -F(y.beta)[s] * fmaf(sg, e, mu);
        A += tb;
        Ae = fmaf(tb, e, Ae);
      }
      const float g_mu = A;
      const float g_rho = F(y.sgm)[tid] * (Ae + beta_sum / sg);
      OST(0) = g_mu;
      OST(1) = g_rho;
      if (p.flags & F_STORE_GOUT) {
        p.gout[q] = g_mu;
        p.gout[Pt + q] = g_rho;
      }
    }
  }

  // =================================================================================================================
  // Phase H: a single Hessian-vector pass along gdot (building block / hyper trainer)
  // =================================================================================================================
  if (p.flags & F_HVP) {
    if (own) {
      F(y.gdm)[tid] = __ldg(p.gdot + q);
      F(y.gdr)[tid] = __ldg(p.gdot + Pt + q);
    }
    gen_eps(0);
    __syncthreads();
    owner_sample(0, true, false);
    wait_rec(true, false);
    int li = 0;
    for (int s = rank; s < S; s += G, ++li) {
      rows_dual(rec0 + li * PS);
      __syncthreads();
      fold_push(s, 2);
      if (s + G < S) __syncthreads();
    }
    wait_recv(2, false);
    if (own2) p.h_phi[oc * Pt + j0 + oj] = owner_hvp(0);
  }

  // =================================================================================================================
  // Phase R: reverse sweep through the T Adam steps   (A.4 + A.6; replaces autograd's double backward)
  // =================================================================================================================
  if ((p.flags & F_REVERSE) && p.T > 0) {
    if (p.flags & F_LOAD_GOUT) {
      if (own) {
        OST(0) = p.gout[q];
        OST(1) = p.gout[Pt + q];
      }
      if (rank == 0) {
        const int MD = p.M * D;
        for (int i = tid; i < MD; i += NT1) F(y.ubar)[i] = p.gout[2 * Pt + i];
        for (int i = tid; i < p.M; i += NT1) F(y.abar)[i] = p.gout[2 * Pt + MD + i];
      }
    }
    gen_eps(p.T - 1);
    __syncthreads();   // (OST written with the load mapping above)
    owner_vjp(p.T - 1, p.traj + (size_t)(p.T - 1) * 8 * Pt + j0 + oj, Pt);
    __syncthreads();
    owner_sample(p.T - 1, true, false);
    wait_rec(true, false);
    for (int t = p.T - 1; t >= 0; --t) {
      stamp(21);
      if (t > 0) {   // next step's noise and trajectory row, staged while the row pass runs
        gen_eps(t - 1);
        // (warps 1..3: warp 0 carries the longest row loop, the last warps draw the noise; loads first, stores after)
        const float* tr = p.traj + (size_t)(t - 1) * 8 * Pt + j0;
        if (warp >= 1 && warp <= 3) {
          for (int i0 = tid - 32; i0 < 8 * slice; i0 += 4 * 96) {
            float v[4];
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const int i = i0 + k4 * 96;
              const int k = i / slice, j = i - k * slice;
              v[k4] = (i < 8 * slice && j0 + j < Pt) ? __ldcg(tr + (size_t)k * Pt + j) : 0.f;
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              if (i0 + k4 * 96 < 8 * slice) F(y.trs)[i0 + k4 * 96] = v[k4];
          }
        }
      }
      int li = 0;
      for (int s = rank; s < S; s += G, ++li) {
        rows_dual(rec0 + li * PS);
        __syncthreads();
        fold_push(s, 2);
        if (s + G < S) __syncthreads();
      }
      stamp(22);
      wait_recv(2, false);
      stamp(23);
      // ---- owner: phibar_t = phibar_{t+1} + H gbar ; then the Adam VJP of step t-1 and its sample ----
      if (own2) OS2(oc) += owner_hvp(t);
      if (t > 0) {
        owner_vjp(t - 1, F(y.trs) + oj, slice);
        stamp(25);
        __syncthreads();
        owner_sample(t - 1, true, false);
        stamp(24);
        wait_rec(true, false);
      }
    }
    __syncthreads();
    if (p.g_out && own) {  // dLoss/dphi_0, useful for diagnostics
      p.g_out[q] = OST(0);
      p.g_out[Pt + q] = OST(1);
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
      for (int i = tid; i < MD + p.M; i += NT1) {
        float s = 0.f;
        const int off = i < MD ? y.ubar + i : y.abar + (i - MD);
        for (int c = 0; c < G; ++c) s += *remote(off, c);
        if (p.flags & F_STORE_GOUT) p.gout[2 * Pt + i] = s;
        sm[off] = s;  // rank 0 now holds the totals (remote reads of rank 0 itself happened in this same iteration)
      }
      __syncthreads();
      if (p.flags & F_FINAL) {
        for (int i = tid; i < MD; i += NT1) p.u_grad[i] = F(y.ubar)[i];
        if (p.v_grad) {
          if (p.vmode == PSVI_VMODE_IDENTITY || p.roww) {
            for (int m = tid; m < p.M; m += NT1) p.v_grad[m] = p.Nf * F(y.abar)[m];
          } else {
            float dot = 0.f;
            for (int m = tid; m < p.M; m += NT1) dot += F(y.f)[m] * F(y.abar)[m];
            dot = block_sum1(dot, red);
            const float sc = p.Nf * (p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(p.alpha) : 1.f);
            for (int m = tid; m < p.M; m += NT1) p.v_grad[m] = sc * F(y.f)[m] * (F(y.abar)[m] - dot);
            if (p.alpha_grad && tid == 0)
              p.alpha_grad[0] = p.vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? sc * dot : 0.f;
          }
        }
      }
    }
    cl_sync();  // keep every CTA's shared memory alive until rank 0 has read it
  }
}

// The shared-memory layout is computed on the host and passed as a second __grid_constant__ parameter: its fields are
// constant-bank operands of the instructions that use them (a layout struct in shared memory costs a dependent LDS per
// access).
template <int D, int C, int UPL>
__global__ void __launch_bounds__(NT1, 1) psvi_mf_fn1_kernel(const __grid_constant__ EP p, const __grid_constant__ FL fl) {
  extern __shared__ __align__(16) float smem_dyn[];
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
    p.slice = (((Pt + G - 1) / G) + 3) & ~3;   // multiple of 4: a Philox block never straddles two owners
    if (p.slice > NT1 / 2) return FN1_NOT_APPLICABLE;   // two owner threads (mu, rho) per TL index
    p.RC = 0;
    FL fl;
    make_fl<D, C, UPL>(p, fl);
    const size_t smem = (size_t)fl.total * 4;
    if (smem + 1024 > (size_t)smem_max) return FN1_NOT_APPLICABLE;   // too many rows: the generic engine chunks them
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(G, 1, 1);
    cfg.blockDim = dim3(NT1, 1, 1);
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
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p, fl);
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

static long long* g_timeline = nullptr;
extern "C" void psvi_internal_set_timeline(void* ptr) { g_timeline = (long long*)ptr; }

int psvi_fn1_launch(EP& p, cudaStream_t stream) {
  p.tl = g_timeline;
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
