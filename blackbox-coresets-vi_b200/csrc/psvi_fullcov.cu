// psvi_fullcov.cu -- the two dense operations of the full-covariance layer (fn2): reference
// psvi/models/neural_net.py:408-491 (MultivariateNormalVIMixin / VILinearMultivariateNormal).
//
//   theta_s = mean + L eps_s,   L = scale_tril (:452-461): diag = softplus(_sd), strictly-lower entries of the top-left
//   (n-1) x (n-1) block = _corr in torch.tril_indices(n-1, n-1, -1) order (row-major: k = r(r-1)/2 + c, 1 <= r <= n-2,
//   c < r); the last row has no off-diagonals (quirk Q6).  The reference rebuilds a dense n x n matrix with two scatters
//   on every access; here L stays packed and is read once per product.
//
//   psvi_fc_matvec : out[s][i] = base[i] + dg[i] * eps[s][i] + sum_{c<i} off[k(i,c)] * eps[s][c]        (sample / tangent)
//   psvi_fc_outer  : g_base[i] = sum_s A[s][i],  g_dg[i] = sum_s A[s][i] eps[s][i],
//                    g_off[k(r,c)] = sum_s A[s][r] eps[s][c]                       (gradient / HVP reductions wrt L)
// Both are memory-bound streams over the packed triangle (n = 1 640 -> 5.4 MB), L2-resident between calls.
#include "psvi_common.cuh"

namespace {

// grid.x = rows, block = 256 threads = 8 warps; warp w handles samples w, w+8, ...; lanes stride over the columns c < r
__global__ void fc_matvec_kernel(int n, int S, const float* __restrict__ base, const float* __restrict__ dg,
                                 const float* __restrict__ off, const float* __restrict__ eps, int ld_eps,
                                 float* __restrict__ out, int ld_out) {
  const int r = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_off = r >= 1 && r <= n - 2;
  const float* row = off + (size_t)r * (r - 1) / 2;
  for (int s = warp; s < S; s += 8) {
    const float* e = eps + (size_t)s * ld_eps;
    float acc = 0.f;
    if (has_off)
      for (int c = lane; c < r; c += 32) acc = fmaf(__ldg(row + c), __ldg(e + c), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[(size_t)s * ld_out + r] = (base ? base[r] : 0.f) + dg[r] * e[r] + acc;
  }
}

// grid.x = rows r (+1 extra block for the vectors), threads over columns c < r
__global__ void fc_outer_kernel(int n, int S, const float* __restrict__ A, int ld_a, const float* __restrict__ eps,
                                int ld_eps, float* __restrict__ g_base, float* __restrict__ g_dg,
                                float* __restrict__ g_off) {
  const int r = blockIdx.x;
  if (r == n) {  // vector parts
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float a = 0.f, b = 0.f;
      for (int s = 0; s < S; ++s) {
        const float v = A[(size_t)s * ld_a + i];
        a += v;
        b = fmaf(v, eps[(size_t)s * ld_eps + i], b);
      }
      g_base[i] = a;
      g_dg[i] = b;
    }
    return;
  }
  if (r < 1 || r > n - 2) return;
  extern __shared__ float ar[];  // A[:, r]
  for (int s = threadIdx.x; s < S; s += blockDim.x) ar[s] = A[(size_t)s * ld_a + r];
  __syncthreads();
  float* row = g_off + (size_t)r * (r - 1) / 2;
  for (int c = threadIdx.x; c < r; c += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = fmaf(ar[s], __ldg(eps + (size_t)s * ld_eps + c), acc);
    row[c] = acc;
  }
}

// ---- the same two operations working directly on the layer's parameter block phi = [mean | _sd | _corr] (and a direction
// phidot of the same layout), with the softplus / sigmoid transforms and the KL / nkl terms folded in: the streaming
// engine's family maps for fn2 are then one launch per layer instead of ~15 elementwise launches.
//   sample : theta[s]    = mean + L eps[s],         L: diag softplus(_sd), off _corr
//   tangent: thetadot[s] = meandot + Ldot eps[s],   Ldot: diag sigmoid(_sd) _sddot, off _corrdot
__global__ void fc_sample_kernel(int n, int S, const float* __restrict__ phi, const float* __restrict__ phid,
                                 const float* __restrict__ eps, int ld_eps, float* __restrict__ out, int ld_out) {
  // longest rows first (row r costs r); a warp carries four samples at once so that a triangle element is loaded once per
  // four FMAs and five independent loads are in flight per lane
  const int r = n - 1 - blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_off = r >= 1 && r <= n - 2;
  const float* src = phid ? phid : phi;
  const float* row = src + 2 * (size_t)n + (size_t)r * (r - 1) / 2;
  const float base = src[r];
  const float dg = phid ? sigmoid_f(phi[n + r]) * phid[n + r] : softplus_f(phi[n + r]);
  for (int s0 = warp * 4; s0 < S; s0 += 32) {
    const float* e[4];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = eps + (size_t)min(s0 + j, S - 1) * ld_eps;
    if (has_off) {
      for (int c = lane; c < r; c += 32) {
        const float l = __ldg(row + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(l, __ldg(e[j] + c), acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t = warp_sum(acc[j]);
      if (lane == 0 && s0 + j < S) out[(size_t)(s0 + j) * ld_out + r] = base + dg * e[j][r] + t;
    }
  }
}

// MODE 0 (gradient): g = [sum_s A + kl m | sig (sum_s A eps + kl (d - 1/d) + nkl / d) | sum_s A[:, r] eps[:, c] + kl corr]
// MODE 1 (HVP, SURVEY A.6 for the Cholesky family): h = [sum_s A + mdot |
//          sig sum_s A eps + sig (1 - sig) sddot sum_s Ad eps + ((1 + 1/d^2) sig^2 + (d - 1/d) sig (1 - sig)) sddot |
//          sum_s A[:, r] eps[:, c] + corrdot]          (d = softplus(_sd), sig = sigmoid(_sd))
// The strictly-lower part  sum_s A[s][r] eps[s][c]  is a rank-S update of an n x n triangle: 64 x 64 output tiles (grid
// (nt, nt): tiles above the diagonal exit; the DIAGONAL tiles also produce the vector parts of their 64 indices from the same
// stage), the A / eps columns of a tile staged in shared memory, a thread owns a 4 x 4 block (8 shared loads per 16 FMAs).  The
// row-per-CTA form issued one global load per FMA, and its ONE block for the vector parts (n x S dependent strided loads) was
// the critical path of the launch: 50 us (gradient) / 114 us (HVP) for cfg3's big layer (n = 1 640).  The sum over s runs in
// ascending order for every entry, as before: results are bit-identical to the row-per-CTA kernel.
constexpr int FT = 64, FS = 32;
template <int MODE>
__global__ void __launch_bounds__(256)
fc_reparam_kernel(int n, int S, const float* __restrict__ phi, const float* __restrict__ phid,
                  const float* __restrict__ A, const float* __restrict__ Ad, int ld_a,
                  const float* __restrict__ eps, int ld_eps, float kl_coef, float nkl_coef,
                  float* __restrict__ g) {
  const int r0 = blockIdx.y * FT, c0 = blockIdx.x * FT;
  if (c0 > r0) return;   // above the diagonal
  __shared__ __align__(16) float As[FS][FT], Es[FS][FT];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4], va = 0.f, vb = 0.f, vbd = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int s0 = 0; s0 < S; s0 += FS) {
    const int ns = min(FS, S - s0);
    __syncthreads();
    {   // 16 global loads per thread in flight before the first store (one memory latency per stage)
      float av[FS * FT / 256], ev[FS * FT / 256];
#pragma unroll
      for (int u = 0; u < FS * FT / 256; ++u) {
        const int i = threadIdx.x + u * 256, ss = i / FT, k = i % FT;
        const bool ok = ss < ns;
        av[u] = (ok && r0 + k < n) ? A[(size_t)(s0 + ss) * ld_a + r0 + k] : 0.f;
        ev[u] = (ok && c0 + k < n) ? eps[(size_t)(s0 + ss) * ld_eps + c0 + k] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < FS * FT / 256; ++u) {
        const int i = threadIdx.x + u * 256;
        As[i / FT][i % FT] = av[u];
        Es[i / FT][i % FT] = ev[u];
      }
    }
    float adv[FS];   // HVP, diagonal tiles: the second adjoint's column of this thread's index, loaded as one batch
    if (MODE == 1 && c0 == r0 && threadIdx.x < FT) {
#pragma unroll
      for (int ss = 0; ss < FS; ++ss)
        adv[ss] = (ss < ns && r0 + (int)threadIdx.x < n) ? Ad[(size_t)(s0 + ss) * ld_a + r0 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    if (c0 == r0 && threadIdx.x < FT) {   // the vector parts of this tile's 64 indices (diagonal tiles only), from the same stage
      const int k = threadIdx.x;
#pragma unroll
      for (int ss = 0; ss < FS; ++ss) {
        if (ss < ns) {
          const float v = As[ss][k], e = Es[ss][k];
          va += v;
          vb = fmaf(v, e, vb);
          if (MODE == 1) vbd = fmaf(adv[ss], e, vbd);
        }
      }
    }
    for (int ss = 0; ss < ns; ++ss) {   // (a padded sample would add +0.f: skipped so that -0.f sums stay bit-identical)
      const float4 a4 = *reinterpret_cast<const float4*>(&As[ss][ty * 4]);
      const float4 e4 = *reinterpret_cast<const float4*>(&Es[ss][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], ev[j], acc[i][j]);
    }
  }
  if (c0 == r0 && threadIdx.x < FT && r0 + (int)threadIdx.x < n) {
    const int i = r0 + threadIdx.x;
    const float sd = phi[n + i], d = softplus_f(sd), sig = sigmoid_f(sd);
    if (MODE == 0) {
      g[i] = va + kl_coef * phi[i];
      g[n + i] = sig * (vb + kl_coef * (d - 1.f / d) + nkl_coef / d);
    } else {
      const float sdd = phid[n + i];
      g[i] = va + phid[i];
      g[n + i] = sig * vb + sig * (1.f - sig) * sdd * vbd +
                 ((1.f + 1.f / (d * d)) * sig * sig + (d - 1.f / d) * sig * (1.f - sig)) * sdd;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r < 1 || r > n - 2) continue;
    const size_t k0 = 2 * (size_t)n + (size_t)r * (r - 1) / 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < r) g[k0 + c] = acc[i][j] + (MODE == 0 ? kl_coef * phi[k0 + c] : phid[k0 + c]);
    }
  }
}

}  // namespace

extern "C" {

int psvi_fc_matvec(int32_t n, int32_t S, const float* base, const float* dg, const float* off, const float* eps,
                   int32_t ld_eps, float* out, int32_t ld_out, void* stream) {
  PSVI_REQUIRE(n >= 1 && S >= 1 && dg && eps && out && (n < 3 || off), PSVI_ERR_INVALID, "bad argument");
  fc_matvec_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(n, S, base, dg, off, eps, ld_eps, out, ld_out);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_fc_outer(int32_t n, int32_t S, const float* A, int32_t ld_a, const float* eps, int32_t ld_eps, float* g_base,
                  float* g_dg, float* g_off, void* stream) {
  PSVI_REQUIRE(n >= 1 && S >= 1 && A && eps && g_base && g_dg && (n < 3 || g_off), PSVI_ERR_INVALID, "bad argument");
  fc_outer_kernel<<<n + 1, 256, S * sizeof(float), (cudaStream_t)stream>>>(n, S, A, ld_a, eps, ld_eps, g_base, g_dg, g_off);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_fc_sample(int32_t n, int32_t S, const float* phi, const float* phidot, const float* eps, int32_t ld_eps, float* out,
                   int32_t ld_out, void* stream) {
  PSVI_REQUIRE(n >= 1 && S >= 1 && phi && eps && out, PSVI_ERR_INVALID, "bad argument");
  fc_sample_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(n, S, phi, phidot, eps, ld_eps, out, ld_out);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_fc_reparam_grad(int32_t n, int32_t S, const float* phi, const float* A, int32_t ld_a, const float* eps, int32_t ld_eps,
                         float kl_coef, float nkl_coef, float* g, void* stream) {
  PSVI_REQUIRE(n >= 1 && S >= 1 && phi && A && eps && g, PSVI_ERR_INVALID, "bad argument");
  const int nt = (n + FT - 1) / FT;
  fc_reparam_kernel<0><<<dim3(nt, nt), 256, 0, (cudaStream_t)stream>>>(n, S, phi, nullptr, A, nullptr, ld_a, eps, ld_eps, kl_coef,
                                                                         nkl_coef, g);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int psvi_fc_reparam_hvp(int32_t n, int32_t S, const float* phi, const float* phidot, const float* A_t, const float* A_td,
                        int32_t ld_a, const float* eps, int32_t ld_eps, float* h, void* stream) {
  PSVI_REQUIRE(n >= 1 && S >= 1 && phi && phidot && A_t && A_td && eps && h, PSVI_ERR_INVALID, "bad argument");
  const int nt = (n + FT - 1) / FT;
  fc_reparam_kernel<1><<<dim3(nt, nt), 256, 0, (cudaStream_t)stream>>>(n, S, phi, phidot, A_t, A_td, ld_a, eps, ld_eps, 0.f, 0.f, h);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
