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

}  // extern "C"
