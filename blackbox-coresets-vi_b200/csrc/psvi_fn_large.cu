// psvi_fn_large.cu -- the per-sample network pass of `fn` with ONE hidden layer in the LARGE regime (SURVEY.md section 7 /
// 8d, BASELINE config 5: D = 256, H = 1024, S = 64, M = 1000) on the Blackwell tensor path.  Same contract as psvi_net_pass
// (forward / gradient / dual Hessian-vector pass on externally supplied sampled weights theta [S][P]), so the streaming PSVI
// engine (psvi/inference/stream.py) drives inner_elbo / psvi_elbo / nested_step for models whose per-sample weights fit no
// CTA.  Reference: VILinear.forward + nn.ReLU + Categorical.log_prob + autograd / double backward
// (psvi/models/neural_net.py:155-179,267-297; psvi_classes.py:445-511,541-600; robust_higher/optim.py:224-229).
//
// Every matrix product of the pass is ONE launch of a batched "TN" GEMM kernel, C[b] = A[b] B[b]^T with
// bf16 K-major operands and fp32 accumulation: TMA (128-byte swizzle) -> 6-stage shared-memory ring -> tcgen05.mma
// (M = N = 128, K = 16) -> double-buffered TMEM accumulator -> 8 epilogue warps (tcgen05.ld, bias / ReLU / ReLU-mask,
// fp32, bf16 and transposed-bf16 stores).  Sums of two products are one GEMM over a concatenated K dimension
// ([hdot | h] [W2 | W2dot]^T etc.); products with K = C <= 16 use zero-padded K = 64 operand blocks.  Activations that feed
// a later GEMM as the K dimension are stored transposed by the producing epilogue.  Reductions over rows with a C- or 1-wide
// output (second-layer and bias gradients) are CUDA-core column reductions.
#include "psvi_tc.cuh"

using namespace psvi_tc;

namespace {

constexpr int GM = 128, GN = 128, GK = 64;
constexpr int GST = 6;                          // ring depth: 6 x (16 KB A + 16 KB B)
constexpr int G_THREADS = 128 + 256;            // 4 role warps + 8 epilogue warps
constexpr int G_TILE_BYTES = GM * GK * 2;       // 16 KB
constexpr int CW = 16, CP = 64;                 // classes padded to 16 (fp32 side) / 64 (bf16 K blocks)

struct GemmP {
  int batch, m_tiles, n_tiles, kc;              // kc = K / 64
  int M_valid, N_valid;
  int a_brows, b_brows;                         // rows per batch in the A / B tensor maps (0: operand shared by all batches)
  const float* bias; long long bias_bs;         // + bias[b * bias_bs + n]
  int relu;
  const __nv_bfloat16* mask; long long mask_bs; int mask_ld;    // * (mask[b][m][n] > 0)
  float* of; long long of_bs; int of_ld;                       // fp32 out [b][m][n], m < M_valid, n < N_valid
  __nv_bfloat16* ob; long long ob_bs; int ob_ld; int ob_rows;   // bf16 out [b][m][n], m < ob_rows (zeros for m >= M_valid)
  __nv_bfloat16* obt; long long obt_bs; int obt_ld;             // bf16 transposed out [b][n][m], m < ob_rows
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(G_THREADS, 1)
tn_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                                    // [GST][A tile | B tile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + GST * 2 * G_TILE_BYTES);
  uint64_t* full = bars;             // [GST]
  uint64_t* empty = bars + GST;      // [GST]
  uint64_t* tfull = empty + GST;     // [2]
  uint64_t* tempty = tfull + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < GST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_b = p.m_tiles * p.n_tiles, n_total = p.batch * per_b;

  if (warp == 0) {
    int st = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_total; t += gridDim.x) {
      const int b = t / per_b, rem = t - b * per_b, mt = rem / p.n_tiles, nt = rem - mt * p.n_tiles;
      const int arow = b * p.a_brows + mt * GM, brow = b * p.b_brows + nt * GN;
      for (int k = 0; k < p.kc; ++k) {
        mbar_wait(&empty[st], ph ^ 1);
        if (elect_one()) {
          uint8_t* dst = ring + st * 2 * G_TILE_BYTES;
          mbar_expect_tx(&full[st], 2 * G_TILE_BYTES);
          tma_load_2d(&map_a, &full[st], dst, k * GK, arow);
          tma_load_2d(&map_b, &full[st], dst + G_TILE_BYTES, k * GK, brow);
        }
        __syncwarp();
        if (++st == GST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_total; t += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * GN;
      for (int k = 0; k < p.kc; ++k) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(ring + st * 2 * G_TILE_BYTES), b0 = a0 + G_TILE_BYTES;
        const uint32_t acc0 = k != 0;
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < GK / 16; ++j)
            umma_bf16(tmem_d, make_desc_sw128(a0 + j * 32), make_desc_sw128(b0 + j * 32), idesc, j ? 1u : acc0);
          umma_commit(&empty[st]);
          if (k == p.kc - 1) umma_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++st == GST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, part = (warp - 4) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    for (int t = blockIdx.x; t < n_total; t += gridDim.x, ++it) {
      const int b = t / per_b, rem = t - b * per_b, mt = rem / p.n_tiles, nt = rem - mt * p.n_tiles;
      const int buf = it & 1;
      const int m = mt * GM + q * 32 + lane;
      const bool vm = m < p.M_valid;
      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int n0 = nt * GN + part * 64 + g * 32;
        float v[32];
        tmem_ld32(lane_addr + buf * GN + part * 64 + g * 32, v);
        if (n0 < p.N_valid) {
          if (p.bias) {
            const float* bp = p.bias + (long long)b * p.bias_bs + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (n0 + j < p.N_valid) ? __ldg(bp + j) : 0.f;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.mask && vm) {   // N_valid is a multiple of 32 whenever a mask / bf16 output is used
            const uint4* mp = reinterpret_cast<const uint4*>(p.mask + (long long)b * p.mask_bs + (long long)m * p.mask_ld + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 mk = __ldg(mp + i);
              const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // bf16 > 0  <=>  sign bit clear and not zero
                const uint32_t lo = w[e] & 0xFFFFu, hi = w[e] >> 16;
                if (!(lo != 0 && !(lo & 0x8000u))) v[i * 8 + e * 2] = 0.f;
                if (!(hi != 0 && !(hi & 0x8000u))) v[i * 8 + e * 2 + 1] = 0.f;
              }
            }
          }
          if (!vm) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (p.of && vm) {
            float* op = p.of + (long long)b * p.of_bs + (long long)m * p.of_ld + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N_valid) op[j] = v[j];
          }
          if (p.ob && m < p.ob_rows) {
            uint4* op = reinterpret_cast<uint4*>(p.ob + (long long)b * p.ob_bs + (long long)m * p.ob_ld + n0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[i * 8 + e * 2], v[i * 8 + e * 2 + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&t2);
              }
              op[i] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          if (p.obt && m < p.ob_rows) {
            __nv_bfloat16* op = p.obt + (long long)b * p.obt_bs + (long long)n0 * p.obt_ld + m;
#pragma unroll
            for (int j = 0; j < 32; ++j) op[(long long)j * p.obt_ld] = __float2bfloat16(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

// ------------------------------------------------------------------------------------------------ operand packing
// x fp32 [R][D] -> Xb bf16 [Rp][D] (padding rows zero) and XT bf16 [D][Rp]
__global__ void fnl_prep_x_kernel(const float* __restrict__ x, int R, int Rp, int D, __nv_bfloat16* __restrict__ Xb,
                                  __nv_bfloat16* __restrict__ XT) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, d0 = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, d = d0 + tx;
    const float v = (r < R && d < D) ? x[(size_t)r * D + d] : 0.f;
    tile[i][tx] = v;
    if (r < Rp && d < D) Xb[(size_t)r * D + d] = __float2bfloat16(v);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, r = r0 + tx;
    if (d < D && r < Rp) XT[(size_t)d * Rp + r] = __float2bfloat16(tile[tx][i]);
  }
}

// first-layer weights of theta (or thetadot) [S][P] -> W1 bf16 [S][H][D] and its transpose into W1T2 [S][D][2H] at column
// offset `toff` (0: primal, H: tangent)
__global__ void fnl_pack_w1_kernel(const float* __restrict__ theta, long long P, int D, int H, __nv_bfloat16* __restrict__ W1b,
                                   __nv_bfloat16* __restrict__ W1T2, int toff) {
  __shared__ float tile[32][33];
  const int d0 = blockIdx.x * 32, h0 = blockIdx.y * 32, s = blockIdx.z, tx = threadIdx.x, ty = threadIdx.y;
  const float* th = theta + (long long)s * P;
  for (int i = ty; i < 32; i += 8) {
    const int h = h0 + i, d = d0 + tx;
    const float v = th[(size_t)h * D + d];
    tile[i][tx] = v;
    W1b[((size_t)s * H + h) * D + d] = __float2bfloat16(v);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, h = h0 + tx;
    W1T2[((size_t)s * D + d) * (2 * H) + toff + h] = __float2bfloat16(tile[tx][i]);
  }
}

// second-layer weights: W2p bf16 [S][128][2H] rows c < C at column offset poff; W2T bf16 [S][H][128] at column offset coff
__global__ void fnl_pack_w2_kernel(const float* __restrict__ theta, long long P, int D, int H, int C, __nv_bfloat16* __restrict__ W2p,
                                   int poff, __nv_bfloat16* __restrict__ W2T, int coff) {
  const int s = blockIdx.y;
  const float* w2 = theta + (long long)s * P + (long long)H * D + H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * H; i += gridDim.x * blockDim.x) {
    const int c = i / H, h = i - c * H;
    const __nv_bfloat16 v = __float2bfloat16(w2[i]);
    W2p[((size_t)s * 128 + c) * (2 * H) + poff + h] = v;
    W2T[((size_t)s * H + h) * 128 + coff + c] = v;
  }
}

// ------------------------------------------------------------------------------------------------ softmax / NLL head
// logits o [S][Rp][16] (bias b2 still to be added) and, in dual mode, od (+ b2dot).
//   mode 0: nll;  mode 1: nll, go = cw (p - onehot) -> fp32 [S][R][16] and bf16 AA[s][r][0..63];
//   mode 2: go = cw p (od - <p, od>) -> fp32 + AA[..][64..127], god = cw (p - onehot) -> fp32 + AA[..][0..63], acbar = q . od
__global__ void fnl_head_kernel(const float* __restrict__ o, const float* __restrict__ od, const float* __restrict__ b2,
                                const float* __restrict__ b2d, long long P, const int* __restrict__ y, const float* __restrict__ cw,
                                int S, int R, int Rp, int C, int mode, float* __restrict__ nll, float* __restrict__ logits_out,
                                float* __restrict__ go, float* __restrict__ god, __nv_bfloat16* __restrict__ AA,
                                float* __restrict__ acbar) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * Rp) return;
  const int s = idx / Rp, r = idx - s * Rp;
  if (r >= R) {
    if (AA && mode >= 1) {
      uint4* a = reinterpret_cast<uint4*>(AA + (size_t)idx * 128);
      for (int i = 0; i < 16; ++i) a[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    return;
  }
  const float* lo = o + (size_t)idx * CW;
  float lg[CW], p[CW], mx = -INFINITY, se = 0.f;
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    lg[c] = c < C ? lo[c] + b2[(long long)s * P + c] : -INFINITY;
    mx = fmaxf(mx, lg[c]);
  }
#pragma unroll
  for (int c = 0; c < CW; ++c) { p[c] = c < C ? expf(lg[c] - mx) : 0.f; se += p[c]; }
  const int lab = y[r];
  float ly = 0.f;
#pragma unroll
  for (int c = 0; c < CW; ++c) if (c == lab) ly = lg[c];
  const size_t oidx = (size_t)s * R + r;
  if (nll) nll[oidx] = -(ly - mx - logf(se));
  if (logits_out)
    for (int c = 0; c < C; ++c) logits_out[oidx * C + c] = lg[c];
  if (mode == 0) return;
  const float w = cw ? cw[oidx] : 1.f, inv = 1.f / se;
  float q[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) { p[c] *= inv; q[c] = c < C ? p[c] - (c == lab ? 1.f : 0.f) : 0.f; }
  __nv_bfloat16* a = AA + (size_t)idx * 128;
  if (mode == 1) {
#pragma unroll
    for (int c = 0; c < CW; ++c) { go[oidx * CW + c] = w * q[c]; a[c] = __float2bfloat16(w * q[c]); }
    for (int c = CW; c < 128; ++c) a[c] = __float2bfloat16(0.f);
    return;
  }
  const float* ld = od + (size_t)idx * CW;
  float dd[CW], dot = 0.f, qd = 0.f;
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    dd[c] = c < C ? ld[c] + b2d[(long long)s * P + c] : 0.f;
    dot += p[c] * dd[c];
    qd += q[c] * dd[c];
  }
  for (int c = 0; c < 128; ++c) a[c] = __float2bfloat16(0.f);
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    const float g1 = c < C ? w * p[c] * (dd[c] - dot) : 0.f, g2 = w * q[c];
    go[oidx * CW + c] = g1;
    god[oidx * CW + c] = g2;
    a[CP + c] = __float2bfloat16(g1);
    a[c] = __float2bfloat16(g2);
  }
  if (acbar) acbar[oidx] = qd;
}

// out[s * P + c * H + h] (+)= sum_r W[s][r][c] * Y[s][r][h]   (W == NULL: C = 1, weight 1 -> column sums)
__global__ void __launch_bounds__(128)
fnl_colreduce_kernel(const __nv_bfloat16* __restrict__ Y, long long y_bs, int y_ld, const float* __restrict__ W, int R, int H, int C,
                     float* __restrict__ out, long long P, int accumulate) {
  const int h = blockIdx.x * 128 + threadIdx.x, s = blockIdx.y;
  if (h >= H) return;
  float acc[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) acc[c] = 0.f;
  const __nv_bfloat16* yp = Y + (long long)s * y_bs + h;
  if (W) {
    const float4* wp = reinterpret_cast<const float4*>(W + (size_t)s * R * CW);
#pragma unroll 2
    for (int r = 0; r < R; ++r) {
      const float yv = __bfloat162float(yp[(long long)r * y_ld]);
      const float4 w0 = __ldg(wp + r * 4), w1 = __ldg(wp + r * 4 + 1), w2 = __ldg(wp + r * 4 + 2), w3 = __ldg(wp + r * 4 + 3);
      acc[0] = fmaf(w0.x, yv, acc[0]); acc[1] = fmaf(w0.y, yv, acc[1]); acc[2] = fmaf(w0.z, yv, acc[2]); acc[3] = fmaf(w0.w, yv, acc[3]);
      acc[4] = fmaf(w1.x, yv, acc[4]); acc[5] = fmaf(w1.y, yv, acc[5]); acc[6] = fmaf(w1.z, yv, acc[6]); acc[7] = fmaf(w1.w, yv, acc[7]);
      acc[8] = fmaf(w2.x, yv, acc[8]); acc[9] = fmaf(w2.y, yv, acc[9]); acc[10] = fmaf(w2.z, yv, acc[10]); acc[11] = fmaf(w2.w, yv, acc[11]);
      acc[12] = fmaf(w3.x, yv, acc[12]); acc[13] = fmaf(w3.y, yv, acc[13]); acc[14] = fmaf(w3.z, yv, acc[14]); acc[15] = fmaf(w3.w, yv, acc[15]);
    }
  } else {
#pragma unroll 4
    for (int r = 0; r < R; ++r) acc[0] += __bfloat162float(yp[(long long)r * y_ld]);
  }
  for (int c = 0; c < C; ++c) {
    float* d = out + (long long)s * P + (long long)c * H + h;
    *d = accumulate ? *d + acc[c] : acc[c];
  }
}

// out[s * P + c] = sum_r W[s][r][c]
__global__ void fnl_colsum16_kernel(const float* __restrict__ W, int R, int C, float* __restrict__ out, long long P) {
  const int s = blockIdx.x, c = threadIdx.x & 15, g = threadIdx.x >> 4;   // 16 x 16 threads
  __shared__ float red[16][17];
  float t = 0.f;
  for (int r = g; r < R; r += 16) t += W[((size_t)s * R + r) * CW + c];
  red[g][c] = t;
  __syncthreads();
  if (g == 0 && c < C) {
    float a = 0.f;
    for (int k = 0; k < 16; ++k) a += red[k][c];
    out[(long long)s * P + c] = a;
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct Lws {
  __nv_bfloat16 *Xb, *XT, *W1b, *W1db, *W1T2, *W2p, *W2T, *hh, *aa, *aT, *adT, *AA;
  float *o, *od, *go, *god;
  size_t total;
};

void carve_l(int S, int R, int D, int H, uint8_t* base, Lws& w) {
  const size_t Rp = (size_t)((R + 127) / 128) * 128;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += (bytes + 1023) & ~(size_t)1023; return p; };
  w.Xb = (__nv_bfloat16*)take(Rp * D * 2);
  w.XT = (__nv_bfloat16*)take((size_t)D * Rp * 2);
  w.W1b = (__nv_bfloat16*)take((size_t)S * H * D * 2);
  w.W1db = (__nv_bfloat16*)take((size_t)S * H * D * 2);
  w.W1T2 = (__nv_bfloat16*)take((size_t)S * D * 2 * H * 2);
  w.W2p = (__nv_bfloat16*)take((size_t)S * 128 * 2 * H * 2);
  w.W2T = (__nv_bfloat16*)take((size_t)S * H * 128 * 2);
  w.hh = (__nv_bfloat16*)take((size_t)S * Rp * 2 * H * 2);
  w.aa = (__nv_bfloat16*)take((size_t)S * Rp * 2 * H * 2);
  w.aT = (__nv_bfloat16*)take((size_t)S * H * Rp * 2);
  w.adT = (__nv_bfloat16*)take((size_t)S * H * Rp * 2);
  w.AA = (__nv_bfloat16*)take((size_t)S * Rp * 128 * 2);
  w.o = (float*)take((size_t)S * Rp * CW * 4);
  w.od = (float*)take((size_t)S * Rp * CW * 4);
  w.go = (float*)take((size_t)S * R * CW * 4);
  w.god = (float*)take((size_t)S * R * CW * 4);
  w.total = off;
}

struct Operand {
  const __nv_bfloat16* base;
  uint64_t inner, outer, ld;   // K extent, rows, leading dimension (elements)
  int brows;                   // rows per batch (0: shared)
};

int launch_gemm(const Operand& A, const Operand& B, GemmP p, int sms, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_map_2d_bf16_ld(&ma, A.base, A.inner, A.outer, A.ld, GK, GM);
  if (rc) return rc;
  rc = make_map_2d_bf16_ld(&mb, B.base, B.inner, B.outer, B.ld, GK, GN);
  if (rc) return rc;
  p.kc = (int)(A.inner / GK);
  p.a_brows = A.brows;
  p.b_brows = B.brows;
  p.m_tiles = (p.M_valid + GM - 1) / GM;
  p.n_tiles = (p.N_valid + GN - 1) / GN;
  if (p.ob && p.ob_rows > p.M_valid) p.m_tiles = (p.ob_rows + GM - 1) / GM;
  const int total = p.batch * p.m_tiles * p.n_tiles;
  const int grid = total < sms ? total : sms;
  const size_t smem = (size_t)GST * 2 * G_TILE_BYTES + (2 * GST + 4) * 8 + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(tn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  tn_gemm_kernel<<<grid, G_THREADS, smem, st>>>(ma, mb, p);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

int check_large(const psvi_mf_model* model) {
  PSVI_REQUIRE(model, PSVI_ERR_INVALID, "null model");
  PSVI_REQUIRE(model->n_layers == 2, PSVI_ERR_UNSUPPORTED, "the large-regime tensor-core pass covers fn with one hidden layer");
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  PSVI_REQUIRE(D % 64 == 0 && D >= 64, PSVI_ERR_UNSUPPORTED, "D=%d must be a multiple of 64", D);
  PSVI_REQUIRE(H % 128 == 0 && H >= 128, PSVI_ERR_UNSUPPORTED, "H=%d must be a multiple of 128", H);
  PSVI_REQUIRE(C >= 1 && C <= CW && S >= 1 && S <= 64, PSVI_ERR_UNSUPPORTED, "need C <= 16 and S <= 64 (got C=%d S=%d)", C, S);
  return PSVI_OK;
}

}  // namespace

extern "C" {

size_t psvi_fnl_workspace_bytes(const psvi_mf_model* model, int32_t R) {
  if (!model || model->n_layers != 2 || R <= 0) return 0;
  Lws w;
  carve_l(model->mc_samples, R, model->dims[0], model->dims[1], nullptr, w);
  return w.total + 1024;
}

int psvi_fnl_pass(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const int32_t* y,
                  const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar, float* logits,
                  void* workspace, void* stream_) {
  PSVI_REQUIRE(theta && x && y && workspace, PSVI_ERR_INVALID, "null pointer");
  int rc = check_large(model);
  if (rc) return rc;
  PSVI_REQUIRE(R >= 1, PSVI_ERR_INVALID, "bad R");
  PSVI_REQUIRE(!thetad || (tbar && tdbar), PSVI_ERR_INVALID, "the dual pass needs tbar and tdbar");
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  const long long P = (long long)H * D + H + (long long)C * H + C;
  const long long o_b1 = (long long)H * D, o_w2 = o_b1 + H, o_b2 = o_w2 + (long long)C * H;
  const int Rp = ((R + 127) / 128) * 128;
  cudaStream_t st = (cudaStream_t)stream_;
  int dev = 0, sms = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  Lws w;
  carve_l(S, R, D, H, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023), w);
  const bool dual = thetad != nullptr;
  // ---- operands
  fnl_prep_x_kernel<<<dim3((Rp + 31) / 32, (D + 31) / 32), dim3(32, 8), 0, st>>>(x, R, Rp, D, w.Xb, w.XT);
  PSVI_CUDA_CHECK(cudaMemsetAsync(w.W2p, 0, (size_t)S * 128 * 2 * H * 2, st));
  PSVI_CUDA_CHECK(cudaMemsetAsync(w.W2T, 0, (size_t)S * H * 128 * 2, st));
  fnl_pack_w1_kernel<<<dim3(D / 32, H / 32, S), dim3(32, 8), 0, st>>>(theta, P, D, H, w.W1b, w.W1T2, 0);
  fnl_pack_w2_kernel<<<dim3(8, S), 256, 0, st>>>(theta, P, D, H, C, w.W2p, 0, w.W2T, CP);
  if (dual) {
    fnl_pack_w1_kernel<<<dim3(D / 32, H / 32, S), dim3(32, 8), 0, st>>>(thetad, P, D, H, w.W1db, w.W1T2, H);
    fnl_pack_w2_kernel<<<dim3(8, S), 256, 0, st>>>(thetad, P, D, H, C, w.W2p, H, w.W2T, 0);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  const Operand opX{w.Xb, (uint64_t)D, (uint64_t)Rp, (uint64_t)D, 0};
  const Operand opXT{w.XT, (uint64_t)Rp, (uint64_t)D, (uint64_t)Rp, 0};
  const Operand opW1{w.W1b, (uint64_t)D, (uint64_t)S * H, (uint64_t)D, H};
  const Operand opW1d{w.W1db, (uint64_t)D, (uint64_t)S * H, (uint64_t)D, H};
  const uint64_t H2 = 2 * (uint64_t)H;
  const Operand opH{w.hh + H, (uint64_t)H, (uint64_t)S * Rp, H2, Rp};            // h
  const Operand opHH{w.hh, H2, (uint64_t)S * Rp, H2, Rp};                        // [hdot | h]
  const Operand opW2{w.W2p, (uint64_t)H, (uint64_t)S * 128, H2, 128};            // W2 (rows >= C zero)
  const Operand opW22{w.W2p, H2, (uint64_t)S * 128, H2, 128};                    // [W2 | W2dot]
  const Operand opA0{w.AA, (uint64_t)CP, (uint64_t)S * Rp, 128, Rp};             // first K block of AA
  const Operand opAA{w.AA, 128, (uint64_t)S * Rp, 128, Rp};                      // [A_od | A_o]
  const Operand opW2T1{w.W2T + CP, (uint64_t)CP, (uint64_t)S * H, 128, H};       // W2^T (padded to 64 classes)
  const Operand opW2TT{w.W2T, 128, (uint64_t)S * H, 128, H};                     // [W2dot^T | W2^T]
  const Operand opA{w.aa, (uint64_t)H, (uint64_t)S * Rp, H2, Rp};                // abar / A_a
  const Operand opAAh{w.aa, H2, (uint64_t)S * Rp, H2, Rp};                       // [A_a | A_adot]
  const Operand opW1T{w.W1T2, (uint64_t)H, (uint64_t)S * D, H2, D};              // W1^T
  const Operand opW1TT{w.W1T2, H2, (uint64_t)S * D, H2, D};                      // [W1^T | W1dot^T]
  const Operand opAT{w.aT, (uint64_t)Rp, (uint64_t)S * H, (uint64_t)Rp, H};
  const Operand opADT{w.adT, (uint64_t)Rp, (uint64_t)S * H, (uint64_t)Rp, H};
  GemmP z;
  memset(&z, 0, sizeof(z));
  z.batch = S;
  const long long hh_bs = (long long)Rp * 2 * H;
  // ---- primal forward: h = relu(X W1^T + b1);  o = h W2^T
  {
    GemmP p = z;
    p.M_valid = R; p.N_valid = H; p.bias = theta + o_b1; p.bias_bs = P; p.relu = 1;
    p.ob = w.hh + H; p.ob_bs = hh_bs; p.ob_ld = 2 * H; p.ob_rows = Rp;
    rc = launch_gemm(opX, opW1, p, sms, st);
    if (rc) return rc;
    p = z;
    p.M_valid = R; p.N_valid = CW; p.of = w.o; p.of_bs = (long long)Rp * CW; p.of_ld = CW;
    rc = launch_gemm(opH, opW2, p, sms, st);
    if (rc) return rc;
  }
  const int hb = (S * Rp + 127) / 128;
  if (!tbar) {
    fnl_head_kernel<<<hb, 128, 0, st>>>(w.o, nullptr, theta + o_b2, nullptr, P, y, nullptr, S, R, Rp, C, 0, nll, logits, nullptr,
                                       nullptr, nullptr, nullptr);
    PSVI_CUDA_CHECK(cudaGetLastError());
    return PSVI_OK;
  }
  const dim3 gcol(H / 128, S);
  if (!dual) {
    // ---- gradient pass
    fnl_head_kernel<<<hb, 128, 0, st>>>(w.o, nullptr, theta + o_b2, nullptr, P, y, cw, S, R, Rp, C, 1, nll, logits, w.go, nullptr,
                                       w.AA, nullptr);
    GemmP p = z;      // abar = (obar W2) * (h > 0), also transposed
    p.M_valid = R; p.N_valid = H; p.mask = w.hh + H; p.mask_bs = hh_bs; p.mask_ld = 2 * H;
    p.ob = w.aa; p.ob_bs = hh_bs; p.ob_ld = 2 * H; p.ob_rows = Rp; p.obt = w.aT; p.obt_bs = (long long)H * Rp; p.obt_ld = Rp;
    rc = launch_gemm(opA0, opW2T1, p, sms, st);
    if (rc) return rc;
    p = z;            // W1bar = abar^T X
    p.M_valid = H; p.N_valid = D; p.of = tbar; p.of_bs = P; p.of_ld = D;
    rc = launch_gemm(opAT, opXT, p, sms, st);
    if (rc) return rc;
    fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.aa, hh_bs, 2 * H, nullptr, R, H, 1, tbar + o_b1, P, 0);
    fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.hh + H, hh_bs, 2 * H, w.go, R, H, C, tbar + o_w2, P, 0);
    fnl_colsum16_kernel<<<S, 256, 0, st>>>(w.go, R, C, tbar + o_b2, P);
    if (xbar) {
      p = z;          // xbar = abar W1
      p.M_valid = R; p.N_valid = D; p.of = xbar; p.of_bs = (long long)R * D; p.of_ld = D;
      rc = launch_gemm(opA, opW1T, p, sms, st);
      if (rc) return rc;
    }
    PSVI_CUDA_CHECK(cudaGetLastError());
    return PSVI_OK;
  }
  // ---- dual pass: tangent forward
  {
    GemmP p = z;      // hdot = (X W1dot^T + b1dot) * (h > 0)
    p.M_valid = R; p.N_valid = H; p.bias = thetad + o_b1; p.bias_bs = P; p.mask = w.hh + H; p.mask_bs = hh_bs; p.mask_ld = 2 * H;
    p.ob = w.hh; p.ob_bs = hh_bs; p.ob_ld = 2 * H; p.ob_rows = Rp;
    rc = launch_gemm(opX, opW1d, p, sms, st);
    if (rc) return rc;
    p = z;            // odot = hdot W2^T + h W2dot^T
    p.M_valid = R; p.N_valid = CW; p.of = w.od; p.of_bs = (long long)Rp * CW; p.of_ld = CW;
    rc = launch_gemm(opHH, opW22, p, sms, st);
    if (rc) return rc;
  }
  fnl_head_kernel<<<hb, 128, 0, st>>>(w.o, w.od, theta + o_b2, thetad + o_b2, P, y, cw, S, R, Rp, C, 2, nll, logits, w.go, w.god,
                                     w.AA, acbar);
  {
    GemmP p = z;      // A_a = (A_od W2dot + A_o W2) * (h > 0)
    p.M_valid = R; p.N_valid = H; p.mask = w.hh + H; p.mask_bs = hh_bs; p.mask_ld = 2 * H;
    p.ob = w.aa; p.ob_bs = hh_bs; p.ob_ld = 2 * H; p.ob_rows = Rp; p.obt = w.aT; p.obt_bs = (long long)H * Rp; p.obt_ld = Rp;
    rc = launch_gemm(opAA, opW2TT, p, sms, st);
    if (rc) return rc;
    p.ob = w.aa + H; p.obt = w.adT;      // A_adot = (A_od W2) * (h > 0)
    rc = launch_gemm(opA0, opW2T1, p, sms, st);
    if (rc) return rc;
    p = z;            // A_W1 = A_a^T X;  A_W1dot = A_adot^T X
    p.M_valid = H; p.N_valid = D; p.of = tbar; p.of_bs = P; p.of_ld = D;
    rc = launch_gemm(opAT, opXT, p, sms, st);
    if (rc) return rc;
    p.of = tdbar;
    rc = launch_gemm(opADT, opXT, p, sms, st);
    if (rc) return rc;
    if (xbar) {
      p = z;          // A_x = A_a W1 + A_adot W1dot
      p.M_valid = R; p.N_valid = D; p.of = xbar; p.of_bs = (long long)R * D; p.of_ld = D;
      rc = launch_gemm(opAAh, opW1TT, p, sms, st);
      if (rc) return rc;
    }
  }
  fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.aa, hh_bs, 2 * H, nullptr, R, H, 1, tbar + o_b1, P, 0);          // A_b1
  fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.aa + H, hh_bs, 2 * H, nullptr, R, H, 1, tdbar + o_b1, P, 0);     // A_b1dot
  fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.hh + H, hh_bs, 2 * H, w.go, R, H, C, tbar + o_w2, P, 0);         // A_o^T h
  fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.hh, hh_bs, 2 * H, w.god, R, H, C, tbar + o_w2, P, 1);            // + A_od^T hdot
  fnl_colreduce_kernel<<<gcol, 128, 0, st>>>(w.hh + H, hh_bs, 2 * H, w.god, R, H, C, tdbar + o_w2, P, 0);       // A_od^T h
  fnl_colsum16_kernel<<<S, 256, 0, st>>>(w.go, R, C, tbar + o_b2, P);
  fnl_colsum16_kernel<<<S, 256, 0, st>>>(w.god, R, C, tdbar + o_b2, P);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
