// psvi_fn_grad_tc.cuh -- full-data DATA-TERM GRADIENT of the one-hidden-layer BNN (`fn`) on the Blackwell tensor path
// (included by psvi_fn_tc.cu, inside its anonymous namespace: shares the forward kernel, the TMA / tcgen05 wrappers and
// the tile constants).
//
// What it computes (reference: the data term of psvi_elbo, psvi/inference/psvi_classes.py:477,484-486, differentiated by
// autograd): for externally sampled weights theta_s (s < S) and a shard
// of R data rows
//     dsum_s      = sum_r nll[s, r]
//     tbar_s      = coef_s * sum_r d nll[s, r] / d theta_s          (per-sample weight adjoints, [S][P] in theta layout)
// with coef_s = w_s N / B the caller's per-sample coefficient (the importance weights depend on the pseudo-data only,
// so the data rows enter linearly: every rank does its shard, ONE all-reduce of the reparameterised [2P + S] closes it).
//
// Two passes over the rows, 3 F_fwd of tensor work in total (no redundant forward):
//   pass 1  psvi_fn_forward_tc_kernel, mode 3 (psvi_fn_tc.cu): logits -> softmax -> nll sums and the output-layer adjoint
//           seeds  obar[s][r][c] = coef_s (p_c - [c == y_r])  as bf16 UMMA-operand blocks (4 KB per sample and row tile);
//   pass 2  psvi_fn_grad_tc_kernel (here): work unit = (sample s, slice of 128 hidden units j, range of row tiles).  The
//           CTA keeps the W1_s slice [128 x D] in shared memory and, per 128-row tile, with the HIDDEN UNITS ON THE TMEM
//           LANES (so that every operand is in its natural layout and the activations never leave tensor memory):
//             Ga    P  [128 j x 128 r] = W1slice X^T                      A, B K-major SW128 (smem)
//             epi-1 h^T = relu(P + b1) as bf16 A operand, in place; ReLU mask kept in registers
//             Gab   Q  [128 j x 128 r] = W2slice^T obar^T  (K = 16)       both K-major, no swizzle
//             Gw2   P' [128 j x 16 c]  = h^T obar                         A from TMEM, B = the obar block read MN-major
//             epi-2 abar^T = mask ? Q : 0 as bf16 A operand, in place; b1bar += row sums (registers); W2bar += P' (registers)
//             Gw1   ACC[128 j x D]    += abar^T X                         A from TMEM, B = the SAME X tile read MN-major
//           ACC (the W1bar slice, fp32) accumulates in TMEM across all row tiles of the unit -- TMEM: P 128 + Q 128 + ACC 256
//           = 512 columns -- and is added to tbar once per unit (a unit range never splits a slice over more than two
//           CTAs, so the two-term floating-point sums are order independent: deterministic).
// Work units are laid out in waves of one unit per SM that sweep the row tiles in the same order, so an X tile is fetched
// from HBM once per wave and served to the other CTAs from L2.
//
// Algorithmic FLOPs per row and sample: forward 2 (D H + H C), backward 2 x that = 3 F_fwd per row overall; bytes per row
// D * 2 + 4 (+ 64 S bytes of seeds written and read H / 128 times).

constexpr int G_THREADS = 384;          // warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4..11 epilogue
constexpr int G_COL_P = 0, G_COL_Q = 128, G_COL_ACC = 256;
constexpr int OB_BYTES = 4096;          // one seeds block: [128 rows x 16 classes] bf16

struct GradUnit {
  int s, j, tile0, tile1;
};

struct GradParams {
  int n_rows, n_tiles, D, kc, H, S, C, P, n_units;
  const GradUnit* units;
  const __nv_bfloat16* obar;  // [S][n_tiles][2048]
  const __nv_bfloat16* W2b;   // [S][CW][H]
  const float* b1;            // [S][H]
  float* tbar;                // [S][P], accumulated
};

// shared-memory operand descriptor, general form (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type << 61 (0 = no swizzle, 2 = 128-byte swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float a, float b) {   // max(., 0) and both conversions in one instruction
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first PTX source -> upper half: a -> low 16 bits
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(G_THREADS, 1)
psvi_fn_grad_tc_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_x, const GradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int XS = p.kc * KCH_BYTES;                 // bytes of one [128 x D] bf16 tile (kc chunks of [128 x 64])
  uint8_t* sW1 = smem;                             // [kc][128 j][64 d] SW128
  uint8_t* sX = sW1 + XS;                          // 2 stages of [kc][128 r][64 d] SW128
  uint8_t* sO = sX + 2 * XS;                       // 2 stages of the seeds block
  uint8_t* sW2T = sO + 2 * OB_BYTES;               // [128 j][16 c] bf16, canonical no-swizzle K-major
  float* sB1x = reinterpret_cast<float*>(sW2T + OB_BYTES);   // [2][128] b1bar halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB1x + 256);
  uint64_t* w1full = bars;          // W1 slice landed
  uint64_t* w1empty = bars + 1;     // every Ga of the unit done
  uint64_t* w2ready = bars + 2;     // epilogue warps have written the W2^T block
  uint64_t* xfull = bars + 3;       // [2] X tile + seeds block landed
  uint64_t* xempty = bars + 5;      // [2] Gw1 of the tile done
  uint64_t* afull = bars + 7;       // Ga done
  uint64_t* hfull = bars + 8;       // epi-1 done (8 warps)
  uint64_t* qfull = bars + 9;       // Gab, Gw2 done
  uint64_t* abfull = bars + 10;     // epi-2 done (8 warps)
  uint64_t* accfull = bars + 11;    // last Gw1 of the unit done
  uint64_t* accempty = bars + 12;   // epilogue has flushed ACC (8 warps)
  uint64_t* pread = bars + 13;      // epilogue has read the Gw2 output out of P (4 warps): the next Ga may overwrite P
  uint64_t* xck = bars + 14;        // [2][4] one per 64-column chunk of an X stage (chunk 0 also covers the seeds block): Ga
                                    // starts on the first chunk while the others are still in flight
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(w1full, 1); mbar_init(w1empty, 1); mbar_init(w2ready, 8);
    for (int i = 0; i < 2; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    mbar_init(afull, 1); mbar_init(hfull, 8); mbar_init(qfull, 1); mbar_init(abfull, 8);
    mbar_init(accfull, 1); mbar_init(accempty, 8); mbar_init(pread, 4);
    for (int i = 0; i < 8; ++i) mbar_init(&xck[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- TMA producer
    int xi = 0, ui = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ui) {
      const GradUnit gu = p.units[u];
      mbar_wait(w1empty, (ui & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(w1full, (uint32_t)XS);
        for (int k = 0; k < p.kc; ++k) tma_load_2d(&map_w1, w1full, sW1 + k * KCH_BYTES, k * BK, gu.s * p.H + gu.j * BN);
      }
      __syncwarp();
      for (int t = gu.tile0; t < gu.tile1; ++t, ++xi) {
        const int st = xi & 1;
        mbar_wait(&xempty[st], ((xi >> 1) & 1) ^ 1);
        if (elect_one()) {
          for (int k = 0; k < p.kc; ++k) {
            mbar_expect_tx(&xck[st * 4 + k], (uint32_t)(KCH_BYTES + (k == 0 ? OB_BYTES : 0)));
            tma_load_2d(&map_x, &xck[st * 4 + k], sX + st * XS + k * KCH_BYTES, k * BK, t * BM);
            if (k == 0) bulk_load_1d(sO + st * OB_BYTES, p.obar + ((size_t)gu.s * p.n_tiles + t) * 2048, OB_BYTES, &xck[st * 4]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    const uint32_t base = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);   // fp32 accum, bf16 A / B, M = 128
    const uint32_t id_a = base | ((uint32_t)(128 >> 3) << 17);                                  // N = 128, K-major A and B
    const uint32_t id_w2 = base | ((uint32_t)(CW >> 3) << 17) | (1u << 16);                     // N = 16, B MN-major
    const uint32_t id_w1 = base | ((uint32_t)(p.D >> 3) << 17) | (1u << 16);                    // N = D, B MN-major
    const uint32_t tP = tmem_base + G_COL_P, tQ = tmem_base + G_COL_Q, tACC = tmem_base + G_COL_ACC;
    // Issue order per unit (tcgen05 operations execute in issue order):
    //   Ga(t0) | for every tile t:  Gab(t), Gw2(t) | Ga(t+1) | Gw1(t)
    // so that the tensor pipe runs Ga(t+1) while the epilogue warps build abar(t), and Gw1(t) while they build h(t+1).
    int xi = 0, ui = 0;
    auto issue_ga = [&](int xj) {      // Ga of the tile with running index xj: P[j, r] = sum_d W1[j, d] X[r, d]
      const int st = xj & 1;
      const uint32_t w0 = smem_u32(sW1), x0 = smem_u32(sX + st * XS);
      for (int k = 0; k < p.kc; ++k) {
        mbar_wait(&xck[st * 4 + k], (xj >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int jj = 0; jj < BK / 16; ++jj)
            umma_bf16(tP, make_desc_sw128(w0 + k * KCH_BYTES + jj * 32), make_desc_sw128(x0 + k * KCH_BYTES + jj * 32), id_a,
                      (k | jj) ? 1u : 0u);
          if (k == p.kc - 1) umma_commit(afull);
        }
        __syncwarp();
      }
    };
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ui) {
      const GradUnit gu = p.units[u];
      mbar_wait(w1full, ui & 1);
      mbar_wait(w2ready, ui & 1);
      issue_ga(xi);
      for (int t = gu.tile0; t < gu.tile1; ++t, ++xi) {
        const int st = xi & 1;
        const uint32_t x0 = smem_u32(sX + st * XS), o0 = smem_u32(sO + st * OB_BYTES);
        mbar_wait(hfull, xi & 1);
        tc_fence_after();
        if (elect_one()) {
          // Gab: Q[j, r] = sum_c W2[c, j] obar[r, c]   (one K = 16 step; no-swizzle K-major operands: LBO between the two
          // 8-class halves, SBO between groups of 8 rows)
          umma_bf16(tQ, make_desc(smem_u32(sW2T), 2048, 128, 0), make_desc(o0, 2048, 128, 0), id_a, 0u);
          // Gw2: P'[j, c] = sum_r h[j, r] obar[r, c]   (A = bf16 pairs in P[0, 64); B = the seeds block read MN-major:
          // LBO between groups of 8 rows (K), SBO between the two 8-class halves (N))
#pragma unroll
          for (int k8 = 0; k8 < BM / 16; ++k8)
            umma_bf16_ts(tP + 64, tP + k8 * 8, make_desc(o0 + k8 * 256, 128, 2048, 0), id_w2, k8 ? 1u : 0u);
          umma_commit(qfull);
        }
        __syncwarp();
        mbar_wait(pread, xi & 1);        // P' has been read: the next Ga (of this or of the next unit) may overwrite P
        if (t + 1 < gu.tile1) {
          issue_ga(xi + 1);
        } else if (elect_one()) {
          umma_commit(w1empty);          // every Ga of the unit has been issued: the next unit's W1 slice may load
        }
        __syncwarp();
        mbar_wait(abfull, xi & 1);
        if (t == gu.tile0) mbar_wait(accempty, (ui & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          // Gw1: ACC[j, d] += sum_r abar[j, r] X[r, d]   (A = bf16 pairs in Q[0, 64); B = the X tile read MN-major: LBO
          // between the 64-column chunks (N), SBO between groups of 8 rows (K))
#pragma unroll
          for (int k8 = 0; k8 < BM / 16; ++k8)
            umma_bf16_ts(tACC, tQ + k8 * 8, make_desc(x0 + k8 * 2048, KCH_BYTES, 1024, 2), id_w1,
                         (t != gu.tile0 || k8) ? 1u : 0u);
          umma_commit(&xempty[st]);
          if (t == gu.tile1 - 1) umma_commit(accfull);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------------- epilogue warps
    const int q = warp & 3, half = (warp - 4) >> 2;   // TMEM lane quarter (hidden units); which half of the tile's rows
    const int jl = q * 32 + lane;                     // hidden unit inside the slice
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int HD = p.H * p.D;
    int xi = 0, ui = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ui) {
      const GradUnit gu = p.units[u];
      const int jg = gu.j * BN + jl;                  // hidden unit
      const float b1v = __ldg(p.b1 + (size_t)gu.s * p.H + jg);
      if (half == 0) {
        // W2^T slice as the A operand of Gab: element (j, c) at (c / 8) * 2048 + j * 16 + (c % 8) * 2 bytes
        uint32_t pk[CW / 2];
#pragma unroll
        for (int c = 0; c < CW; c += 2) {
          const __nv_bfloat16 a = p.W2b[((size_t)gu.s * CW + c) * p.H + jg], b = p.W2b[((size_t)gu.s * CW + c + 1) * p.H + jg];
          pk[c / 2] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(sW2T);
        dst[jl] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[128 + jl] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(w2ready);
      float bsum = 0.f, w2acc[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) w2acc[c] = 0.f;
      for (int t = gu.tile0; t < gu.tile1; ++t, ++xi) {
        // ---- epi-1: h^T = relu(a + b1) -> bf16 pairs over the rows, in place; mask bits of my 64 rows
        mbar_wait(afull, xi & 1);
        tc_fence_after();
        uint32_t m0 = 0u, m1 = 0u, pk[32];
        {
          float v[32];
          tmem_ld32(lane_addr + G_COL_P + half * 64, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = v[i] + b1v, a1 = v[i + 1] + b1v;
            m0 |= (a0 > 0.f ? 1u : 0u) << i;
            m0 |= (a1 > 0.f ? 1u : 0u) << (i + 1);
            pk[i / 2] = pack_relu_bf16x2(a0, a1);
          }
          tmem_ld32(lane_addr + G_COL_P + half * 64 + 32, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = v[i] + b1v, a1 = v[i + 1] + b1v;
            m1 |= (a0 > 0.f ? 1u : 0u) << i;
            m1 |= (a1 > 0.f ? 1u : 0u) << (i + 1);
            pk[16 + i / 2] = pack_relu_bf16x2(a0, a1);
          }
        }
        // the other half's warp of this lane quarter still reads columns this warp is about to overwrite
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        tmem_st32(lane_addr + G_COL_P + half * 32, pk);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(hfull);
        // ---- epi-2: abar^T = mask ? Q : 0 -> bf16 pairs, in place; bias adjoint; W2bar contribution of the tile
        mbar_wait(qfull, xi & 1);
        tc_fence_after();
        if (half == 0) {   // W2bar contribution of the tile first: it releases P for the next tile's Ga
          float w[CW];
          tmem_ld16(lane_addr + G_COL_P + 64, w);
#pragma unroll
          for (int c = 0; c < CW; ++c) w2acc[c] += w[c];
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(pread);
        }
        {
          float v[32];
          tmem_ld32(lane_addr + G_COL_Q + half * 64, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = ((m0 >> i) & 1u) ? v[i] : 0.f, a1 = ((m0 >> (i + 1)) & 1u) ? v[i + 1] : 0.f;
            bsum += a0 + a1;
            pk[i / 2] = pack_bf16(a0, a1);
          }
          tmem_ld32(lane_addr + G_COL_Q + half * 64 + 32, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a0 = ((m1 >> i) & 1u) ? v[i] : 0.f, a1 = ((m1 >> (i + 1)) & 1u) ? v[i + 1] : 0.f;
            bsum += a0 + a1;
            pk[16 + i / 2] = pack_bf16(a0, a1);
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        tmem_st32(lane_addr + G_COL_Q + half * 32, pk);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(abfull);
      }
      // ---- flush the unit: ACC -> tbar (W1bar slice), W2bar slice, b1bar slice
      mbar_wait(accfull, ui & 1);
      tc_fence_after();
      float* trow = p.tbar + (size_t)gu.s * p.P;
      {
        const int dh = p.D >> 1;   // my half of the D columns
        float* dst = trow + (size_t)jg * p.D + half * dh;
        for (int g = 0; g < dh; g += 32) {
          float v[32];
          tmem_ld32(lane_addr + G_COL_ACC + half * dh + g, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(dst + g + i, v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty);
      sB1x[half * 128 + jl] = bsum;
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
      if (half == 0) {
        atomicAdd(trow + HD + jg, sB1x[jl] + sB1x[128 + jl]);
        for (int c = 0; c < p.C; ++c) atomicAdd(trow + HD + p.H + (size_t)c * p.H + jg, w2acc[c]);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");   // sB1x is reused by the next unit
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// theta [S][P] (theta layout, fp32) -> the operand buffers of the forward / gradient kernels: W1 as bf16 [S][H][D], W2 as
// bf16 [S][CW][H] (zero rows for the padding classes), b1 [S][H], b2 [S][CW] (-inf for the padding classes)
__global__ void fn_theta_prep_kernel(const float* theta, int S, int D, int H, int C, __nv_bfloat16* W1b, __nv_bfloat16* W2b,
                                     float* b1, float* b2) {
  const int s = blockIdx.y;
  const size_t P = (size_t)H * D + H + (size_t)C * H + C;
  const float* th = theta + (size_t)s * P;
  const size_t HD = (size_t)H * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HD; i += (size_t)gridDim.x * blockDim.x)
    W1b[(size_t)s * HD + i] = __float2bfloat16_rn(th[i]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CW * H; i += gridDim.x * blockDim.x) {
    const int c = i / H, h = i - c * H;
    W2b[(size_t)s * CW * H + i] = c < C ? __float2bfloat16_rn(th[HD + H + (size_t)c * H + h]) : __float2bfloat16_rn(0.f);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += gridDim.x * blockDim.x) b1[(size_t)s * H + i] = th[HD + i];
  if (blockIdx.x == 0 && threadIdx.x < CW)
    b2[s * CW + threadIdx.x] = threadIdx.x < C ? th[HD + H + (size_t)C * H + threadIdx.x] : -INFINITY;
}

// b2bar[s][c] = sum over the rows of the seeds: block (chunk, s) sums its range of tiles in a fixed order into
// partial[s][chunk][CW]; fn_b2bar_final_kernel adds the chunks in order (deterministic, no atomics)
constexpr int B2_CHUNKS = 32;
__global__ void fn_b2bar_kernel(const __nv_bfloat16* obar, int n_tiles, float* partial) {
  __shared__ float red[CW][256];
  const int s = blockIdx.y, ch = blockIdx.x;
  const int t0 = (int)((long long)n_tiles * ch / B2_CHUNKS), t1 = (int)((long long)n_tiles * (ch + 1) / B2_CHUNKS);
  float acc[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) acc[c] = 0.f;
  // element (r, c) of tile t at t * 2048 + (c / 8) * 1024 + r * 8 + (c % 8)   (bf16 elements)
  const __nv_bfloat16* base = obar + ((size_t)s * n_tiles + t0) * 2048;
  for (size_t i = threadIdx.x; i < (size_t)(t1 - t0) * 256; i += blockDim.x) {   // one 16-byte piece (8 classes of a row) each
    const size_t t = i >> 8, rem = i & 255, hsel = rem >> 7, r = rem & 127;
    const uint4 v = *reinterpret_cast<const uint4*>(base + t * 2048 + hsel * 1024 + r * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = __uint_as_float(w[k] << 16), hi = __uint_as_float(w[k] & 0xffff0000u);
      if (hsel == 0) { acc[2 * k] += lo; acc[2 * k + 1] += hi; } else { acc[8 + 2 * k] += lo; acc[8 + 2 * k + 1] += hi; }
    }
  }
#pragma unroll
  for (int c = 0; c < CW; ++c) red[c][threadIdx.x] = acc[c];
  __syncthreads();
  if (threadIdx.x < CW) {
    double t = 0.0;
    for (int i = 0; i < 256; ++i) t += (double)red[threadIdx.x][i];
    partial[((size_t)s * B2_CHUNKS + ch) * CW + threadIdx.x] = (float)t;
  }
}
__global__ void fn_b2bar_final_kernel(const float* partial, int C, int P, int off_b2, float* tbar) {
  const int s = blockIdx.x, c = threadIdx.x;
  if (c < C) {
    double t = 0.0;
    for (int ch = 0; ch < B2_CHUNKS; ++ch) t += (double)partial[((size_t)s * B2_CHUNKS + ch) * CW + c];
    tbar[(size_t)s * P + off_b2 + c] = (float)t;
  }
}

struct GradScratch {
  FnScratch fs;            // operand buffers + per-tile partials of the forward kernel
  __nv_bfloat16* obar;     // [S][tiles][2048]
  GradUnit* units;
  size_t total;
};
constexpr int MAX_UNITS = 4096;

void carve_grad(const psvi_mf_model* model, int64_t rows, uint8_t* base, GradScratch& g) {
  const size_t S = model->mc_samples, D = model->dims[0], H = model->dims[1];
  const size_t tiles = (size_t)((rows + BM - 1) / BM);
  const int nsplit = split_for((int)tiles, (int)S, 148);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* q = base ? base + off : nullptr; off += align256(bytes); return q; };
  memset(&g.fs, 0, sizeof(g.fs));
  g.fs.W1b = reinterpret_cast<__nv_bfloat16*>(take(S * H * D * 2));
  g.fs.W2b = reinterpret_cast<__nv_bfloat16*>(take(S * CW * H * 2));
  g.fs.b1 = reinterpret_cast<float*>(take(S * H * 4));
  g.fs.b2 = reinterpret_cast<float*>(take(S * CW * 4));
  size_t part_floats = tiles * 4 * S > 4096 ? tiles * 4 * S : 4096;
  if (part_floats < S * B2_CHUNKS * CW) part_floats = S * B2_CHUNKS * CW;   // (reused for the b2bar partials)
  g.fs.part = reinterpret_cast<float*>(take(part_floats * 4));
  g.fs.probs = reinterpret_cast<float*>(take(nsplit > 1 ? (size_t)nsplit * rows * CW * 4 : 256));
  g.obar = reinterpret_cast<__nv_bfloat16*>(take(S * tiles * OB_BYTES));
  g.units = reinterpret_cast<GradUnit*>(take(MAX_UNITS * sizeof(GradUnit)));
  g.total = off;
}

// waves of one unit per SM: whole slices first; the slices left over for the last wave are cut into equal row ranges so that
// the wave still fills the machine
int build_units(int S, int NH, int tiles, int sms, GradUnit* out) {
  const int items = S * NH;
  const int full = (items / sms) * sms, rem = items - full;
  int n = 0;
  for (int i = 0; i < full; ++i) out[n++] = GradUnit{i / NH, i % NH, 0, tiles};
  if (rem > 0) {
    int parts = sms / rem;
    if (parts > 2) parts = 2;          // at most two CTAs per slice: two-term sums are order independent
    if (parts > tiles) parts = tiles;
    if (parts < 1) parts = 1;
    for (int k = 0; k < parts; ++k)
      for (int i = full; i < items; ++i)
        out[n++] = GradUnit{i / NH, i % NH, (int)((long long)tiles * k / parts), (int)((long long)tiles * (k + 1) / parts)};
  }
  return n;
}

int data_grad(const psvi_mf_model* model, const float* theta, const void* x_bf16, const int32_t* labels, const float* coef,
              int64_t n_rows, float* dsum, float* tbar, void* scratch, cudaStream_t stream) {
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  const int P = H * D + H + C * H + C;
  int dev = 0, sms = 0, smem_max = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
  GradScratch g;
  carve_grad(model, n_rows, base, g);
  const int tiles = (int)((n_rows + BM - 1) / BM);
  fn_theta_prep_kernel<<<dim3(64, S), 256, 0, stream>>>(theta, S, D, H, C, g.fs.W1b, g.fs.W2b, g.fs.b1, g.fs.b2);
  PSVI_CUDA_CHECK(cudaGetLastError());
  // pass 1: nll sums + adjoint seeds
  int rc = forward(model, g.fs, x_bf16, labels, nullptr, n_rows, 3, coef, nullptr, dsum, nullptr, stream, g.obar);
  if (rc) return rc;
  // pass 2
  PSVI_CUDA_CHECK(cudaMemsetAsync(tbar, 0, (size_t)S * P * sizeof(float), stream));
  static thread_local GradUnit host_units[MAX_UNITS];
  const int NH = H / BN;
  PSVI_REQUIRE(S * NH + sms <= MAX_UNITS, PSVI_ERR_UNSUPPORTED, "too many (sample, hidden slice) work units");
  const int n_units = build_units(S, NH, tiles, sms, host_units);
  PSVI_CUDA_CHECK(cudaMemcpyAsync(g.units, host_units, (size_t)n_units * sizeof(GradUnit), cudaMemcpyHostToDevice, stream));
  CUtensorMap map_w1, map_x;
  rc = make_map_2d_bf16(&map_w1, g.fs.W1b, (uint64_t)D, (uint64_t)S * H, BK, BN);
  if (rc) return rc;
  rc = make_map_2d_bf16(&map_x, x_bf16, (uint64_t)D, (uint64_t)n_rows, BK, BM);
  if (rc) return rc;
  GradParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.n_rows = (int)n_rows; gp.n_tiles = tiles; gp.D = D; gp.kc = D / BK; gp.H = H; gp.S = S; gp.C = C; gp.P = P;
  gp.n_units = n_units; gp.units = g.units; gp.obar = g.obar; gp.W2b = g.fs.W2b; gp.b1 = g.fs.b1; gp.tbar = tbar;
  const size_t smem = (size_t)3 * gp.kc * KCH_BYTES + 3 * OB_BYTES + 256 * 4 + 24 * 8 + 16 + 1024;
  PSVI_REQUIRE(smem <= (size_t)smem_max, PSVI_ERR_UNSUPPORTED, "gradient kernel needs %zu B of shared memory", smem);
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_fn_grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = n_units < sms ? n_units : sms;
  psvi_fn_grad_tc_kernel<<<grid, G_THREADS, smem, stream>>>(map_w1, map_x, gp);
  PSVI_CUDA_CHECK(cudaGetLastError());
  // (the per-tile partials of pass 1 have been reduced by forward(): g.fs.part is free again)
  fn_b2bar_kernel<<<dim3(B2_CHUNKS, S), 256, 0, stream>>>(g.obar, tiles, g.fs.part);
  PSVI_CUDA_CHECK(cudaGetLastError());
  fn_b2bar_final_kernel<<<S, 32, 0, stream>>>(g.fs.part, C, P, H * D + H + C * H, tbar);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}
