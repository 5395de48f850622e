// psvi_lr_tc.cu -- full-data predictive log-likelihood for the single-layer model (logistic_regression) on the
// Blackwell tensor path: TMA-streamed row tiles, tcgen05.mma with TMEM accumulators, fused softmax / importance-weighted
// mixture / NLL / argmax epilogue.  This is the HBM-bound member of the "Kernel D" family of SURVEY.md section 7/8d:
// per test row the kernel reads D bf16 values + one label and nothing else (all sampled weights stay in shared memory).
//
//   logits[r, s*16 + c] = sum_d X[r, d] * W_s[c, d] + b_s[c]          (one 128 x (16 S) x D UMMA per row tile)
//   probs[r, c]         = sum_s w_s softmax_c(logits[r, s, :])        (reference psvi_classes.py:1072-1080)
//   nll_r = -log clamp(probs[r, y_r] / sum_c probs[r, c]),  correct_r = [argmax_c probs == y_r]   (:1081-1083)
//
// Warp roles (128 + 128*NSPLIT threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4.. = epilogue: warp w reads TMEM lanes 32(w%4)..+31 (= rows of the tile) and the two warps that
// share a lane quarter split the MC samples (s mod NSPLIT), exchanging their partial mixtures through shared memory.
// Pipelines: up to 8 x 16 KB A stages (full/empty mbarriers), double-buffered TMEM accumulator (tmem_full/tmem_empty).
#include "psvi_tc.cuh"

using namespace psvi_tc;

// defined in psvi_mf_engine.cu (internal, not part of the C ABI)
int psvi_internal_eval_logweights(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                                  const float* u, const int32_t* z, const float* v, int32_t M, int32_t slab, float N,
                                  int32_t vmode, float alpha, float* lw, cudaStream_t stream);

namespace {

constexpr int BM = 128;      // rows per tile (UMMA M)
constexpr int BK = 64;       // bf16 elements per K chunk = one 128-byte swizzle atom
constexpr int STAGES = 8;    // maximum A-operand pipeline depth (8 x 16 KB); fewer when the weights need the room
constexpr int CP = 16;       // logit columns reserved per MC sample (C <= 16)
constexpr int NSPLIT_MAX = 4;     // epilogue warps per TMEM lane quarter (they split the MC samples): 3 or 4
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct TcParams {
  int n_rows, n_tiles, D, kc, S, C, NP;  // NP accumulator columns: sample s owns columns [s * cs, s * cs + C)
  int cs;                                // column stride per sample: C (packed) or CP
  int mode;                              // 0 importance weighted, 1 uniform weights
  int stages;                            // A-operand pipeline depth actually used (2..STAGES)
  const float* bias;                     // [NP]
  const float* wts;                      // [S] LOG importance weights (mode 0)
  const int* labels;                     // [n_rows]
  float* part;                           // [grid][4]
};

// ------------------------------------------------------------------------------------------------ the kernel
// CC = number of logit columns the epilogue touches per sample (C rounded up to 2/4/8/12/16; the padding columns carry
// bias = -inf, so they drop out of max / exp / sum without any predicate)
template <int CC, int NSPLIT>
__global__ void __launch_bounds__(128 + 128 * NSPLIT, 1)
psvi_lr_predictive_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                             const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128-byte-swizzled operand tiles need 1024-byte aligned bases (in the shared address space)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // carve-up: [B: kc chunks of NP x 128 B] [A: STAGES x 16 KB] [bias NP f32] [wts 16 f32] [barriers] [tmem slot]
  const int b_chunk_bytes = p.NP * 128;
  uint8_t* sB = smem;
  uint8_t* sA = smem + p.kc * b_chunk_bytes;  // multiple of 1024
  float* s_bias = reinterpret_cast<float*>(sA + p.stages * A_STAGE_BYTES);
  float* s_w = s_bias + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + 16);
  uint64_t* full = bars;                // [STAGES]
  uint64_t* empty = bars + STAGES;      // [STAGES]
  uint64_t* bfull = bars + 2 * STAGES;  // [1]
  uint64_t* tfull = bfull + 1;          // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);  // [4 quarters][2]
  float* s_xch = s_red + 8;  // [2 buffers][NSPLIT-1][CP][128 rows] partial mixtures of the helper warps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TC_THREADS = 128 + 128 * NSPLIT;  // 4 role warps + 4*NSPLIT epilogue warps
  for (int i = threadIdx.x; i < p.S * CP; i += TC_THREADS) s_bias[i] = p.bias[i] * 1.4426950408889634f;  // log2(e) folded in
  if (threadIdx.x == 0) {  // importance weights: softmax over the S log-weights (mode 0) or uniform
    float mx = -INFINITY, se = 0.f;
    if (p.mode == 0) {
      for (int s = 0; s < p.S; ++s) mx = fmaxf(mx, p.wts[s]);
      for (int s = 0; s < p.S; ++s) se += expf(p.wts[s] - mx);
    }
    for (int s = 0; s < 16; ++s) s_w[s] = s < p.S ? (p.mode == 0 ? expf(p.wts[s] - mx) / se : 1.f / (float)p.S) : 0.f;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(bfull, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * NSPLIT); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: 512 columns (two NP-column accumulators), allocated by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first = blockIdx.x, step = gridDim.x;
  if (warp == 0) {
    if (lane == 0) {
      // B operand: every K chunk of the stacked sampled weights, once
      mbar_expect_tx(bfull, (uint32_t)(p.kc * b_chunk_bytes));
      for (int k = 0; k < p.kc; ++k) tma_load_2d(&map_w, bfull, sB + k * b_chunk_bytes, k * BK, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first; t < p.n_tiles; t += step) {
        for (int k = 0; k < p.kc; ++k) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], A_STAGE_BYTES);
          tma_load_2d(&map_x, &full[stage], sA + stage * A_STAGE_BYTES, k * BK, t * BM);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N=NP, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      mbar_wait(bfull, 0);
      tc_fence_after();
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = first; t < p.n_tiles; t += step, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * 256;
        for (int k = 0; k < p.kc; ++k) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + stage * A_STAGE_BYTES), b0 = smem_u32(sB + k * b_chunk_bytes);
#pragma unroll
          for (int j = 0; j < BK / 16; ++j)  // 16 bf16 = 32 bytes further along K inside the swizzle atom
            umma_bf16(tmem_d, make_desc_sw128(a0 + j * 32), make_desc_sw128(b0 + j * 32), idesc, (k | j) != 0);
          umma_commit(&empty[stage]);  // frees the A stage once the MMAs above have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[buf]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;          // TMEM lane quarter == warp index % 4
    const int half = (warp - 4) >> 2;  // 0: samples 0, NSPLIT, .. + finalisation; k > 0: samples k, k + NSPLIT, ..
    const float LOG2E = 1.4426950408889634f;
    float nll_sum = 0.f, correct = 0.f;
    int it = 0;
    for (int t = first; t < p.n_tiles; t += step, ++it) {
      const int buf = it & 1;
      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      const int rl = q * 32 + lane;  // row within the tile
      const int row = t * BM + rl;
      // the label is needed only after the sample loop: issue its (long-latency) global load now
      const int y = (half == 0 && row < p.n_rows) ? __ldg(p.labels + row) : -1;
      float probs[CC];
#pragma unroll
      for (int c = 0; c < CC; ++c) probs[c] = 0.f;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256;
#pragma unroll 2
      for (int s = half; s < p.S; s += NSPLIT) {
        float v[16];
        tmem_ld16(taddr + s * p.cs, v);
        const float* bs = s_bias + s * CP;
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          v[c] = fmaf(v[c], LOG2E, bs[c]);
          mx = fmaxf(mx, v[c]);
        }
#pragma unroll
        for (int c = 0; c < CC; ++c) v[c] = ex2_approx(v[c] - mx);
        float se;
        {   // pairwise tree: a dependent chain of log2(CC) adds instead of CC
          float t[CC];
#pragma unroll
          for (int c = 0; c < CC; ++c) t[c] = v[c];
#pragma unroll
          for (int w = CC; w > 1; w = (w + 1) / 2) {
#pragma unroll
            for (int c = 0; c < w / 2; ++c) t[c] = t[c] + t[w - 1 - c];
          }
          se = t[0];
        }
        const float sc = __fdividef(s_w[s], se);
#pragma unroll
        for (int c = 0; c < CC; ++c) probs[c] = fmaf(sc, v[c], probs[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);  // 4*NSPLIT arrivals (one per epilogue warp) release the accumulator
      float* xch = s_xch + buf * ((NSPLIT - 1) * CP * BM);
      if (half > 0) {
#pragma unroll
        for (int c = 0; c < CC; ++c) xch[((half - 1) * CP + c) * BM + rl] = probs[c];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NSPLIT) : "memory");  // the epilogue warps only
      if (half == 0 && row < p.n_rows) {
        float tot = 0.f, best = -1.f, py = 0.f;
        int am = 0;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          float pc = probs[c];
#pragma unroll
          for (int h = 0; h < NSPLIT - 1; ++h) pc += xch[(h * CP + c) * BM + rl];
          tot += pc;
          if (c == y) py = pc;
          if (pc > best) { best = pc; am = c; }
        }
        const float pn = fminf(fmaxf(__fdividef(py, tot), 1.1920929e-07f), 1.f - 1.1920929e-07f);
        nll_sum -= logf(pn);
        correct += (am == y) ? 1.f : 0.f;
      }
    }
    nll_sum = warp_sum(nll_sum);
    correct = warp_sum(correct);
    if (lane == 0 && half == 0) { s_red[q * 2] = nll_sum; s_red[q * 2 + 1] = correct; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int q = 0; q < 4; ++q) { a += s_red[q * 2]; b += s_red[q * 2 + 1]; }
    int rows = 0;
    for (int t = first; t < p.n_tiles; t += step) rows += min(BM, p.n_rows - t * BM);
    float* o = p.part + (size_t)blockIdx.x * 4;
    o[0] = a; o[1] = b; o[2] = (float)rows; o[3] = 0.f;
  }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// stacked bf16 weights [16 S][D] (rows c >= C of a sample are zero) and fp32 biases from (mu, rho, noise slab)
__global__ void lr_prep_kernel(const float* mu, const float* rho, psvi_noise noise, int slab, int S, int C, int D,
                               __nv_bfloat16* Wb, float* bias, int cs, int NP) {
  const int P = C * D + C;
  const int total = NP * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int d = i % D, n = i / D, s = n / cs, c = n % cs;
    float th = 0.f;
    if (c < C && s < S) {
      const int q = c * D + d;
      float e;
      if (noise.mode == PSVI_NOISE_PHILOX) {
        float e4[4];
        philox_normal4(noise.seed, noise.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)(q >> 2), e4);
        e = e4[q & 3];
      } else {
        e = noise.eps[((size_t)slab * S + s) * P + q];
      }
      th = mu[q] + softplus_f(rho[q]) * e;
    }
    Wb[i] = __float2bfloat16(th);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S * CP; i += gridDim.x * blockDim.x) {
    const int s = i / CP, c = i % CP;
    float th = -INFINITY;  // padding logit columns vanish from max / exp / sum
    if (c < C) {
      const int q = C * D + c;
      float e;
      if (noise.mode == PSVI_NOISE_PHILOX) {
        float e4[4];
        philox_normal4(noise.seed, noise.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)(q >> 2), e4);
        e = e4[q & 3];
      } else {
        e = noise.eps[((size_t)slab * S + s) * P + q];
      }
      th = mu[q] + softplus_f(rho[q]) * e;
    }
    bias[i] = th;
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void tc_reduce_kernel(const float* part, int n, float* out, const float* lw, int S) {
  __shared__ double acc[3][256];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < n; i += 256) { a += part[4 * i]; b += part[4 * i + 1]; c += part[4 * i + 2]; }
  acc[0][threadIdx.x] = a; acc[1][threadIdx.x] = b; acc[2][threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) {
      double t = 0;
      for (int i = 0; i < 256; ++i) t += acc[k][i];
      out[k] = (float)t;
    }
    if (lw) {  // importance-weight entropy and normalised ESS (psvi_classes.py:1085-1092)
      float mx = -INFINITY, se = 0.f, ent = 0.f, sw = 0.f, sw2 = 0.f;
      for (int s = 0; s < S; ++s) mx = fmaxf(mx, lw[s]);
      for (int s = 0; s < S; ++s) se += expf(lw[s] - mx);
      for (int s = 0; s < S; ++s) {
        const float w = expf(lw[s] - mx) / se;
        if (w > 0.f) ent -= logf(w) * w;
        sw += w;
        sw2 += w * w;
      }
      out[3] = ent;
      out[4] = sw * sw / sw2 / (float)S;
    }
  }
}

}  // namespace

extern "C" {

int psvi_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  PSVI_REQUIRE(src && dst && n > 0, PSVI_ERR_INVALID, "bad argument");
  PSVI_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, PSVI_ERR_INVALID,
               "src must be 16-byte and dst 8-byte aligned");
  f32_to_bf16_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, (long long)n);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

size_t psvi_lr_predictive_tc_scratch_bytes(const psvi_mf_model* model) {
  if (!model || model->n_layers != 1) return 0;
  const size_t S = model->mc_samples, D = model->dims[0];
  // [Wb bf16 16 S D] [bias 16 S] [wts S (+pad)] [partials 148*4] [eval scratch for the importance weights]
  return S * CP * D * 2 + (S * CP + 64 + 148 * 4 + 1024) * sizeof(float) + psvi_mf_eval_scratch_bytes(model, 1, 1);
}

int psvi_lr_predictive_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                          const float* u, const int32_t* z, const float* v, int32_t M, const void* xt_bf16,
                          const int32_t* yt, int64_t n_rows, int32_t slab, float N, int32_t vmode, float alpha,
                          int32_t mode, float* out, void* scratch, void* stream_) {
  PSVI_REQUIRE(model && noise && mu && rho && xt_bf16 && yt && out && scratch, PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(model->n_layers == 1, PSVI_ERR_UNSUPPORTED, "the tensor-core predictive kernel covers the single-layer model");
  const int D = model->dims[0], C = model->dims[1], S = model->mc_samples;
  PSVI_REQUIRE(D % BK == 0 && D >= BK && D <= 256, PSVI_ERR_UNSUPPORTED, "D=%d must be a multiple of 64 in [64, 256]", D);
  PSVI_REQUIRE(C >= 1 && C <= CP && S >= 1 && S <= 16, PSVI_ERR_UNSUPPORTED, "need C <= 16 and S <= 16 (got C=%d S=%d)", C, S);
  PSVI_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) && slab >= 0, PSVI_ERR_INVALID, "bad n_rows / slab");
  PSVI_REQUIRE(mode == 0 || mode == 1, PSVI_ERR_INVALID, "mode must be 0 (importance weighted) or 1 (uniform)");
  PSVI_REQUIRE((reinterpret_cast<uintptr_t>(xt_bf16) & 15) == 0, PSVI_ERR_INVALID, "xt_bf16 must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  // epilogue width CC >= C; accumulator columns: sample s at [s * cs, s * cs + C).  Packing the samples (cs = C) instead of
  // one 16-column group each cuts the MMA N (S = 16, C = 10: 160 instead of 256 -- at N = 256 the kernel is tensor-bound).
  // The zero weight rows up to NP keep every column the epilogue touches finite (its padding columns carry bias = -inf).
  const int CCsel = C <= 2 ? 2 : C <= 4 ? 4 : C <= 6 ? 6 : C <= 8 ? 8 : C <= 10 ? 10 : C <= 12 ? 12 : 16;
  const int np_packed = (((S - 1) * C + CCsel) + 15) & ~15;
  const int cs = np_packed < CP * S ? C : CP;
  const int NP = cs == C ? np_packed : CP * S;
  uint8_t* sc = static_cast<uint8_t*>(scratch);
  __nv_bfloat16* Wb = reinterpret_cast<__nv_bfloat16*>(sc);
  float* bias = reinterpret_cast<float*>(sc + (size_t)CP * S * D * 2);
  float* wts = bias + CP * S;
  float* part = wts + 64;
  // 1. log importance weights of this slab from the pseudo-data forward (fp32, one CTA per sample)
  if (mode == 0) {
    PSVI_REQUIRE(u && z && v && M > 0, PSVI_ERR_INVALID, "importance-weighted mode needs pseudo-data");
    int rc = psvi_internal_eval_logweights(model, noise, mu, rho, u, z, v, M, slab, N, vmode, alpha, wts, stream);
    if (rc) return rc;
  }
  // 2. stacked sampled weights in bf16
  lr_prep_kernel<<<(NP * D + 255) / 256, 256, 0, stream>>>(mu, rho, *noise, slab, S, C, D, Wb, bias, cs, NP);
  PSVI_CUDA_CHECK(cudaGetLastError());
  // 3. tensor maps + the streaming kernel
  CUtensorMap map_x, map_w;
  int rc = make_map_2d_bf16(&map_x, xt_bf16, (uint64_t)D, (uint64_t)n_rows, BK, BM);
  if (rc) return rc;
  rc = make_map_2d_bf16(&map_w, Wb, (uint64_t)D, (uint64_t)NP, BK, (uint32_t)NP);
  if (rc) return rc;
  TcParams p;
  p.n_rows = (int)n_rows; p.n_tiles = (int)((n_rows + BM - 1) / BM); p.D = D; p.kc = D / BK; p.S = S; p.C = C; p.NP = NP; p.cs = cs;
  p.mode = mode; p.bias = bias; p.wts = wts; p.labels = yt; p.part = part;
  int dev = 0, sms = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.n_tiles < sms ? p.n_tiles : sms;
  PSVI_REQUIRE(grid <= 148 + 16, PSVI_ERR_UNSUPPORTED, "more SMs than the partial buffer holds");
  int smem_max = 0;
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = (size_t)p.kc * NP * 128 + (256 + 16) * 4 + (2 * STAGES + 5) * 8 + 8 + 32 + 2 * (NSPLIT_MAX - 1) * CP * BM * 4 + 1024;
  int stages = (int)(((size_t)smem_max - fixed) / A_STAGE_BYTES);
  if (stages > STAGES) stages = STAGES;
  PSVI_REQUIRE(stages >= 2, PSVI_ERR_UNSUPPORTED, "not enough shared memory for a 2-stage pipeline (S=%d, D=%d)", S, D);
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * A_STAGE_BYTES;
  // epilogue warps per lane quarter: 3 measured faster than 4 at S = 4, 10 and 16 (profiles/r2_lr_tc_summary.md); the
  // 4-way instance stays selectable for experiments
  int ns = 3;
  if (const char* e = getenv("PSVI_LR_NSPLIT")) ns = atoi(e) == 4 ? 4 : 3;
#define PSVI_TC_LAUNCH2(CCV, NSV)                                                                                        \
  do {                                                                                                                  \
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_lr_predictive_tc_kernel<CCV, NSV>,                                         \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                      \
    psvi_lr_predictive_tc_kernel<CCV, NSV><<<grid, 128 + 128 * NSV, smem, stream>>>(map_x, map_w, p);                    \
  } while (0)
#define PSVI_TC_LAUNCH(CCV)                                                                                              \
  do {                                                                                                                  \
    if (ns == 4) PSVI_TC_LAUNCH2(CCV, 4);                                                                               \
    else PSVI_TC_LAUNCH2(CCV, 3);                                                                                       \
  } while (0)
  if (C <= 2) PSVI_TC_LAUNCH(2);
  else if (C <= 4) PSVI_TC_LAUNCH(4);
  else if (C <= 6) PSVI_TC_LAUNCH(6);
  else if (C <= 8) PSVI_TC_LAUNCH(8);
  else if (C <= 10) PSVI_TC_LAUNCH(10);
  else if (C <= 12) PSVI_TC_LAUNCH(12);
  else PSVI_TC_LAUNCH(16);
#undef PSVI_TC_LAUNCH2
#undef PSVI_TC_LAUNCH
  PSVI_CUDA_CHECK(cudaGetLastError());
  // 4. fixed-order reduction + weight diagnostics
  tc_reduce_kernel<<<1, 256, 0, stream>>>(part, grid, out, mode == 0 ? wts : nullptr, S);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
