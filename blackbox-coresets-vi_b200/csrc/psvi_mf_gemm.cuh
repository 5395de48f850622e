// psvi_mf_gemm.cuh -- pieces shared by the mean-field MLP kernels (cluster engine and streaming path): padded layouts,
// block reductions and the division-free shared-memory small-GEMM primitive.
#pragma once
#include "psvi_common.cuh"

namespace psvi_mf {

constexpr int NT = 256;  // threads per CTA
constexpr int MAXL = PSVI_MAX_LAYERS;

struct Meta {
  int din[MAXL + 1], dout[MAXL + 1], ldw[MAXL + 1], woff[MAXL + 1], boff[MAXL + 1], tlw[MAXL + 1], tlb[MAXL + 1];
  int lda[MAXL + 1];
  int Pt, Pp;
};

__host__ __device__ inline void make_meta(const int* dims, int L, Meta& m) {
  // Padded layouts.  Activations of layer l are rows of lda[l] = (d_l + 1) | 1 floats: d_l values, then a constant 1
  // (the tangent buffers keep 0 there), so that the bias is just one more weight column: sampled weights of layer l are
  // rows of ldw[l] = (d_{l-1} + 1) | 1 floats [W[o][0..din-1], b[o], pad].  Odd leading dimensions keep shared-memory
  // accesses conflict-free along either index.
  int pp = 0, pt = 0;
  for (int l = 0; l <= L; ++l) m.lda[l] = (dims[l] + 1) | 1;
  for (int l = 1; l <= L; ++l) {
    const int din = dims[l - 1], dout = dims[l];
    m.din[l] = din;
    m.dout[l] = dout;
    m.ldw[l] = (din + 1) | 1;
    m.woff[l] = pp;
    m.boff[l] = pp + din;  // bias of output o lives at woff + o*ldw + din
    pp += dout * m.ldw[l];
    m.tlw[l] = pt;
    pt += dout * din;
    m.tlb[l] = pt;
    pt += dout;
  }
  m.Pt = pt;
  m.Pp = pp;
}

// ----------------------------------------------------------------------------------------------------------------
// block-wide helpers
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // protect red against a previous use
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < NT / 32) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the total
}
__device__ __forceinline__ double block_sum_d(double v, float* red_) {
  double* red = reinterpret_cast<double*>(red_ + 16);  // red[16..31] as 8 doubles (red is 16-byte aligned)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < NT / 32; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < NT / 32) ? red[l] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

// Small shared-memory GEMM:  C(r, c) = sum_k A(r, k) * B(k, c)  [+ sum_k A2(r, k) * B2(k, c)]  for r < nrows, c < ncols,
// with A(r,k) = A[r*a_rs + k*a_ks], B(k,c) = B[c*b_cs + k*b_ks]; epi(r, c, value) consumes every output exactly once.
// Division-free thread mapping: the block is a (row-groups x column-lanes x k-split) grid with power-of-two extents;
// when there are fewer outputs than threads the spare threads split K (g lanes per output, combined by shuffles);
// otherwise each thread register-blocks 4 rows so that a B element is loaded once per 4 FMAs.
struct GemmOp {
  const float* A;
  int a_rs, a_ks;
  const float* B;
  int b_cs, b_ks;
  const float* A2;
  const float* B2;
};

template <class Epi>
__device__ __forceinline__ void small_gemm(int nrows, int ncols, int K, const GemmOp& op, Epi epi) {
  // smallest power of two >= min(ncols, NT)   (NT = 256 = 2^8)
  const int cl_sh = ncols <= 1 ? 0 : min(8, 32 - __clz(ncols - 1));
  const int CL = 1 << cl_sh;
  int RG = NT >> cl_sh;  // row groups before the k-split
  int g_sh = 0;
  while ((2 << g_sh) <= 32 && nrows * (2 << g_sh) <= RG && K >= (8 << g_sh)) ++g_sh;
  const int g = 1 << g_sh;
  RG >>= g_sh;
  const int tid = threadIdx.x;
  const int ks = tid & (g - 1), col = (tid >> g_sh) & (CL - 1), rg = tid >> (g_sh + cl_sh);
  const bool two = op.A2 != nullptr;
  for (int cb = 0; cb < ncols; cb += CL) {
    const int cc = cb + col;
    const bool cok = cc < ncols;
    const float* Bp = op.B + (cok ? cc : 0) * op.b_cs;
    const float* B2p = two ? op.B2 + (cok ? cc : 0) * op.b_cs : nullptr;
    if (g == 1 && K <= 4) {
      // short dot products (first layer with D <= 3 inputs + bias, data adjoints with C <= 4 classes): the column of B stays
      // in registers for all rows and the k loop is fully unrolled (predicated, no loads beyond K)
      float b[4], b2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        b[k] = (cok && k < K) ? Bp[k * op.b_ks] : 0.f;
        b2[k] = (two && cok && k < K) ? B2p[k * op.b_ks] : 0.f;
      }
      if (cok) {
        for (int rr = rg; rr < nrows; rr += RG) {
          const float* A0 = op.A + rr * op.a_rs;
          float c0 = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < K) c0 = fmaf(A0[k * op.a_ks], b[k], c0);
          if (two) {
            const float* A20 = op.A2 + rr * op.a_rs;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < K) c0 = fmaf(A20[k * op.a_ks], b2[k], c0);
          }
          epi(rr, cc, c0);
        }
      }
    } else if (g == 1) {
      int rr = rg;
      // 4-row register blocking
      for (; rr + 3 * RG < nrows; rr += 4 * RG) {
        const float* A0 = op.A + rr * op.a_rs;
        const int st = RG * op.a_rs;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
        if (cok) {
#pragma unroll 4
          for (int k = 0; k < K; ++k) {
            const float b = Bp[k * op.b_ks];
            const float* a = A0 + k * op.a_ks;
            c0 = fmaf(a[0], b, c0);
            c1 = fmaf(a[st], b, c1);
            c2 = fmaf(a[2 * st], b, c2);
            c3 = fmaf(a[3 * st], b, c3);
          }
          if (two) {
            const float* A20 = op.A2 + rr * op.a_rs;
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
              const float b = B2p[k * op.b_ks];
              const float* a = A20 + k * op.a_ks;
              c0 = fmaf(a[0], b, c0);
              c1 = fmaf(a[st], b, c1);
              c2 = fmaf(a[2 * st], b, c2);
              c3 = fmaf(a[3 * st], b, c3);
            }
          }
          epi(rr, cc, c0);
          epi(rr + RG, cc, c1);
          epi(rr + 2 * RG, cc, c2);
          epi(rr + 3 * RG, cc, c3);
        }
      }
      for (; rr < nrows; rr += RG) {
        if (cok) {
          const float* A0 = op.A + rr * op.a_rs;
          float c0 = 0.f;
#pragma unroll 4
          for (int k = 0; k < K; ++k) c0 = fmaf(A0[k * op.a_ks], Bp[k * op.b_ks], c0);
          if (two) {
            const float* A20 = op.A2 + rr * op.a_rs;
#pragma unroll 4
            for (int k = 0; k < K; ++k) c0 = fmaf(A20[k * op.a_ks], B2p[k * op.b_ks], c0);
          }
          epi(rr, cc, c0);
        }
      }
    } else {
      // k-split: every lane of a warp runs the same trip count (shuffles need the full warp)
      for (int rb = 0; rb < nrows; rb += RG) {
        const int rr = rb + rg;
        const bool ok = cok && rr < nrows;
        float c0 = 0.f;
        if (ok) {
          const float* A0 = op.A + rr * op.a_rs;
#pragma unroll 4
          for (int k = ks; k < K; k += g) c0 = fmaf(A0[k * op.a_ks], Bp[k * op.b_ks], c0);
          if (two) {
            const float* A20 = op.A2 + rr * op.a_rs;
#pragma unroll 4
            for (int k = ks; k < K; k += g) c0 = fmaf(A20[k * op.a_ks], B2p[k * op.b_ks], c0);
          }
        }
        for (int off = g >> 1; off > 0; off >>= 1) c0 += __shfl_xor_sync(0xffffffffu, c0, off);
        if (ok && ks == 0) epi(rr, cc, c0);
      }
    }
  }
}


// Two products that share their operands, in one pass (the weight adjoints of the dual backward):
//   C (r, c) = sum_k A(r, k) B(k, c) + A2(r, k) B2(k, c)      (A2 / B2 as in GemmOp; B2 may be NULL: second term dropped)
//   Cd(r, c) = sum_k A2(r, k) B(k, c)
// epi(r, c, C, Cd) consumes every output pair exactly once.  Same thread mapping as small_gemm.
template <class Epi>
__device__ __forceinline__ void small_gemm_pair(int nrows, int ncols, int K, const GemmOp& op, Epi epi) {
  const int cl_sh = ncols <= 1 ? 0 : min(8, 32 - __clz(ncols - 1));
  const int CL = 1 << cl_sh;
  int RG = NT >> cl_sh;
  int g_sh = 0;
  while ((2 << g_sh) <= 32 && nrows * (2 << g_sh) <= RG && K >= (8 << g_sh)) ++g_sh;
  const int g = 1 << g_sh;
  RG >>= g_sh;
  const int tid = threadIdx.x;
  const int ks = tid & (g - 1), col = (tid >> g_sh) & (CL - 1), rg = tid >> (g_sh + cl_sh);
  const bool two = op.B2 != nullptr;
  for (int cb = 0; cb < ncols; cb += CL) {
    const int cc = cb + col;
    const bool cok = cc < ncols;
    const float* Bp = op.B + (cok ? cc : 0) * op.b_cs;
    const float* B2p = two ? op.B2 + (cok ? cc : 0) * op.b_cs : nullptr;
    for (int rb = 0; rb < nrows; rb += RG) {
      const int rr = rb + rg;
      const bool ok = cok && rr < nrows;
      float c0 = 0.f, d0 = 0.f;
      if (ok) {
        const float* A0 = op.A + rr * op.a_rs;
        const float* A20 = op.A2 + rr * op.a_rs;
        if (two) {
#pragma unroll 4
          for (int k = ks; k < K; k += g) {
            const float b = Bp[k * op.b_ks], a2 = A20[k * op.a_ks];
            c0 = fmaf(A0[k * op.a_ks], b, fmaf(a2, B2p[k * op.b_ks], c0));
            d0 = fmaf(a2, b, d0);
          }
        } else {
#pragma unroll 4
          for (int k = ks; k < K; k += g) {
            const float b = Bp[k * op.b_ks];
            c0 = fmaf(A0[k * op.a_ks], b, c0);
            d0 = fmaf(A20[k * op.a_ks], b, d0);
          }
        }
      }
      for (int off = g >> 1; off > 0; off >>= 1) {   // (g == 1: no trips; otherwise every lane of the warp takes part)
        c0 += __shfl_xor_sync(0xffffffffu, c0, off);
        d0 += __shfl_xor_sync(0xffffffffu, d0, off);
      }
      if (ok && ks == 0) epi(rr, cc, c0, d0);
    }
  }
}


}  // namespace psvi_mf
