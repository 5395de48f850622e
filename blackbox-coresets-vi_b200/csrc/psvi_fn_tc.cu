// psvi_fn_tc.cu -- sampled-GEMM forward of the one-hidden-layer BNN (`fn`) on the Blackwell tensor path, large regime
// ("Kernel A/D large": D <= 256, H = 1024, S = 64, thousands to millions of rows).
//
//   per MC sample s and row r:  logits[s, r, :] = relu(X[r, :] W1_s^T + b1_s) W2_s^T + b2_s
//     mode 0 (objective / log-weights): part[tile, quarter, s] = sum_r cw[r] * nll[s, r]   (inner_elbo data term, psvi_classes.py:488-511;
//                                       evaluate's log-weights :1047-1057) and optionally nll[s, r] itself
//     mode 1/2 (predictive):            probs[r, c] = sum_s w_s softmax_c(logits[s, r, :]) -> NLL / argmax  (:1072-1083)
//
// Work item = (row tile of 128 rows, sample split).  The bf16 X tile stays in shared memory for all samples of the item.
// Per sample and per chunk of 128 hidden units:
//   GEMM1  tcgen05.mma kind::f16, M=128 rows, N=128 hidden units, K=D: A = X tile (smem, SW128), B = W1_s chunk streamed by
//          TMA through a 6-stage ring (all S*H*D*2 bytes of sampled first-layer weights are L2-resident: 33.5 MB at cfg5),
//          fp32 accumulator in TMEM (two buffers);
//   epilogue (8 warps): tcgen05.ld -> + b1 -> ReLU -> bf16 pairs -> tcgen05.st back into TMEM as the A operand of
//   GEMM2  tcgen05.mma (A from TMEM), M=128, N=16 classes, K=128 hidden units, B = W2_s chunk (smem, SW128), accumulated
//          over the H/128 chunks of the sample in a 16-column TMEM accumulator (two buffers across samples);
//   final  (4 warps per sample, alternating halves): tcgen05.ld 16 logits -> + b2 -> softmax / NLL / mixture.
// The hidden activations never leave the SM.  Every work item starts at a different (sample, hidden chunk) position
// (rotation by the item index) so that concurrently running CTAs stream different parts of W1 instead of hammering the same
// L2 lines in lockstep.  The MMA warp software-pipelines GEMM2(i) behind GEMM1(i+1) so the tensor pipe
// stays busy while the epilogue warps convert chunk i.
#include "psvi_tc.cuh"
#include <stdlib.h>

using namespace psvi_tc;

namespace {

constexpr int BM = 128, BK = 64, BN = 128;   // row tile, K chunk (one 128-byte swizzle atom of bf16), hidden chunk
constexpr int BST = 5;                       // W1 ring depth (5 stages of up to two [128 x 64] K-chunks = 32 KB)
constexpr int WST = 4;                       // second-layer chunk ring depth (4 x (4 KB + 512 B))
constexpr int CW = 16;                       // classes padded to 16
constexpr int EPW = 2;                       // epilogue warps per TMEM lane quarter (each owns 128 / EPW accumulator columns; 4 measured: no gain)
constexpr int CPW = BN / EPW;                // fp32 accumulator columns per epilogue warp
constexpr int FN_THREADS = 128 + 128 * EPW;  // 4 role warps + 4 * EPW epilogue warps
constexpr int KCH_BYTES = BN * BK * 2;       // 16 KB: one [128 x 64] bf16 operand tile (one K-chunk)
constexpr int STAGE_BYTES = 2 * KCH_BYTES;   // a ring stage holds two K-chunks (one when D / 64 is odd)
constexpr int W2_STAGE_BYTES = 2 * CW * 128; // two SW128 atoms of [16 rows x 64 k]
constexpr int COL_X = 0, COL_ACC = 128, COL_D2 = 384;  // TMEM columns: X tile (D/2 <= 128, bf16 pairs) | 2 x 128 fp32 (the
                                                        // bf16 hidden activations overwrite them in place) | 2 x (4 x 16) fp32 logits
constexpr int NKL_BLOCKS = 64;

struct FnParams {
  int n_rows, n_tiles, D, kc, ks, H, hc, S, nsplit, mode;   // kc = D / 64 K-chunks, ks = K-chunks per ring stage
  const __nv_bfloat16* x;  // [n_rows][D] bf16 rows
  unsigned long long* prof;  // PSVI_FN_PROF builds: [grid][4][8] cycle counters
  const float* b1;       // [S][H]
  const float* b2;       // [S][CW]  (padding classes = -inf)
  const float* cw;       // [n_rows] row weights (mode 0; nullable -> 1)
  const float* lw;       // [S] log importance weights (mode 1)
  const int* labels;     // [n_rows]
  float* nll_out;        // mode 0: [S][n_rows] (nullable)
  float* part;           // mode 0: [n_tiles][4][S] weighted nll sums; mode >= 1, nsplit == 1: [grid][4] (nll, correct, rows, 0)
  float* probs_out;      // mode >= 1, nsplit > 1: [nsplit][n_rows][CW] partial mixtures
  __nv_bfloat16* obar;   // mode 3: [S][n_tiles] blocks of 2048 bf16: output-layer adjoint seeds as a UMMA operand (see below)
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t pack_relu_bf16(float a, float b) {
  // one instruction for max(., 0) and the conversion of both values (first PTX source -> upper half): a -> low 16 bits
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// cycle-counter instrumentation of the role warps (build with -DPSVI_FN_PROF; scratch/prof_fn_tc.py prints it)
#ifdef PSVI_FN_PROF
#define PROF_DECL unsigned long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long tstart = clock64();
#define PROF_T(slot, stmt) do { const long long _t0 = clock64(); stmt; pc[slot] += (unsigned long long)(clock64() - _t0); } while (0)
#define PROF_OUT(row) do { if (p.prof && lane == 0) { unsigned long long* o = p.prof + (size_t)blockIdx.x * 32 + (row) * 8; \
    for (int i = 0; i < 7; ++i) o[i] = pc[i]; o[7] = (unsigned long long)(clock64() - tstart); } } while (0)
#else
#define PROF_DECL
#define PROF_T(slot, stmt) stmt
#define PROF_OUT(row)
#endif

// ---- thread-block-cluster helpers (CL = 2: the two CTAs of a cluster share every W1 stage through TMA multicast)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"(mask)
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <int CL>
__global__ void __launch_bounds__(FN_THREADS, 1)
psvi_fn_forward_tc_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                          const FnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // carve-up: [W1 ring: BST x 16 KB] [W2 ring: WST x 4 KB] [b1 ring: WST x 512 B] [w: 64 f32] [xch: CW x 128 f32]
  //           [red: 8 f32] [barriers] [tmem slot]
  uint8_t* sB = smem;
  uint8_t* sW2 = sB + BST * STAGE_BYTES;
  float* sB1 = reinterpret_cast<float*>(sW2 + WST * W2_STAGE_BYTES);
  float* s_w = sB1 + WST * BN;
  float* s_xch = s_w + 64;
  float* s_red = s_xch + (EPW - 1) * CW * BM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 8);
  uint64_t* xfull = bars;            // [1]  the 8 epilogue warps have stored the X tile into TMEM
  uint64_t* xempty = bars + 1;       // [1]  every GEMM1 of the item has read the X tile
  uint64_t* full = bars + 2;         // [BST] W1 stage landed
  uint64_t* empty = full + BST;      // [BST] W1 stage consumed
  uint64_t* wfull = empty + BST;     // [WST] W2 chunk + b1 chunk landed
  uint64_t* wempty = wfull + WST;    // [WST] GEMM2 done with the W2 chunk (1 commit) and the 8 epilogue warps done with b1
  uint64_t* tfull = wempty + WST;    // [2]  GEMM1 accumulator complete
  uint64_t* hfull = tfull + 2;       // [2]  the 8 epilogue warps have stored the bf16 hidden activations (in place)
  uint64_t* lfull = hfull + 2;       // [2]  logits of a sample complete
  uint64_t* lempty = lfull + 2;      // [2]  the 4 final warps have read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    float mx = -INFINITY, se = 0.f;
    if (p.mode == 1) {
      for (int s = 0; s < p.S; ++s) mx = fmaxf(mx, p.lw[s]);
      for (int s = 0; s < p.S; ++s) se += expf(p.lw[s] - mx);
    }
    for (int s = 0; s < 64; ++s)
      s_w[s] = s < p.S ? (p.mode == 1 ? expf(p.lw[s] - mx) / se : (p.mode == 3 ? p.lw[s] : 1.f / (float)p.S)) : 0.f;
    mbar_init(xfull, 8); mbar_init(xempty, 1);
    for (int i = 0; i < BST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }   // CL MMA warps release a shared stage
    for (int i = 0; i < WST; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1 + 4 * EPW); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1); mbar_init(&hfull[i], 4 * EPW);
      mbar_init(&lfull[i], 1); mbar_init(&lempty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // the peer's barriers must be initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // work items are enumerated per cluster: item w = (group of CL consecutive row tiles, sample split); CTA `crank` of the
  // cluster takes tile CL * group + crank, so both CTAs walk the same (sample, hidden chunk) sequence in lockstep
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
  const int n_items = ((p.n_tiles + CL - 1) / CL) * p.nsplit;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------- TMA producer
    // (the whole warp runs the loop so that control flow stays warp-uniform; one elected lane issues the copies)
    int st = 0, it = 0;
    uint32_t ph = 0;
    PROF_DECL
    for (int w = cid; w < n_items; w += ncl) {
      const int split = w % p.nsplit;
      const int ns = (p.S - split + p.nsplit - 1) / p.nsplit, srot = w % ns, hrot = (w / ns) % p.hc;
      for (int js = 0; js < ns; ++js) {
        const int s = split + ((js + srot) % ns) * p.nsplit;
        for (int hh = 0; hh < p.hc; ++hh, ++it) {
          const int h = (hh + hrot) % p.hc;
          const int wi = it % WST;
          PROF_T(0, mbar_wait(&wempty[wi], ((it / WST) & 1) ^ 1));
          if (elect_one()) {
            mbar_expect_tx(&wfull[wi], W2_STAGE_BYTES + BN * 4);
            tma_load_2d(&map_w2, &wfull[wi], sW2 + wi * W2_STAGE_BYTES, h * BN, s * CW);
            tma_load_2d(&map_w2, &wfull[wi], sW2 + wi * W2_STAGE_BYTES + CW * 128, h * BN + BK, s * CW);
            bulk_load_1d(sB1 + wi * BN, p.b1 + (size_t)s * p.H + h * BN, BN * 4, &wfull[wi]);
          }
          __syncwarp();
          for (int k = 0; k < p.kc; k += p.ks) {
            PROF_T(1, mbar_wait(&empty[st], ph ^ 1));
            if (elect_one()) {
              mbar_expect_tx(&full[st], (uint32_t)(p.ks * KCH_BYTES));
              for (int kk = 0; kk < p.ks; ++kk) {
                if (CL > 1)   // this CTA fetches rows [64 crank, +64) of the chunk and multicasts them into both CTAs
                  tma_load_2d_mc(&map_w1, &full[st], sB + st * STAGE_BYTES + kk * KCH_BYTES + crank * (KCH_BYTES / 2),
                                 (k + kk) * BK, s * p.H + h * BN + crank * (BN / 2), (uint16_t)3);
                else
                  tma_load_2d(&map_w1, &full[st], sB + st * STAGE_BYTES + kk * KCH_BYTES, (k + kk) * BK, s * p.H + h * BN);
              }
            }
            __syncwarp();
            if (++st == BST) { st = 0; ph ^= 1; }
          }
        }
      }
    }
    PROF_OUT(0);
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------- MMA issuer
    // warp-uniform loop; one elected lane issues tcgen05.mma / tcgen05.commit (keeps ptxas from wrapping every
    // uniform-datapath instruction in a per-thread serialisation loop).  tcgen05.mma operations execute in issue order,
    // so GEMM1(i+2) overwriting the accumulator buffer that GEMM2(i) reads as its A operand needs no barrier.
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CW >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int st = 0, it = 0, item = 0, jsamp = 0;
    uint32_t ph = 0;
    int pend_it = -1, pend_h = 0, pend_js = 0;   // GEMM2 of chunk `pend_it` is issued after GEMM1 of the next chunk
    PROF_DECL
    auto gemm2 = [&](int c_it, int c_h, int c_js) {
      const int buf = c_it & 1, wi = c_it % WST, sb = c_js & 1;
      PROF_T(0, mbar_wait(&wfull[wi], (c_it / WST) & 1));
      if (c_h == 0) PROF_T(1, mbar_wait(&lempty[sb], ((c_js >> 1) & 1) ^ 1));
      PROF_T(2, mbar_wait(&hfull[buf], (c_it >> 1) & 1));
      tc_fence_after();
      const uint32_t d2 = tmem_base + COL_D2 + sb * (4 * CW), a2 = tmem_base + COL_ACC + buf * BN;
      const uint32_t b0 = smem_u32(sW2 + wi * W2_STAGE_BYTES);
      const uint32_t acc0 = c_h != 0;
      PROF_T(3, if (elect_one()) {
#pragma unroll
        for (int j = 0; j < BN / 16; ++j)   // the bf16 activations of hidden units CPW part + 2 c, + 1 sit at column CPW part + c
          umma_bf16_ts(d2 + (j & 3) * CW, a2 + ((j * 16) / CPW) * CPW + ((j * 16) % CPW) / 2,
                       make_desc_sw128(b0 + (j >> 2) * (CW * 128) + (j & 3) * 32), idesc2, (j >> 2) ? 1u : acc0);
        umma_commit(&wempty[wi]);
        if (c_h == p.hc - 1) umma_commit(&lfull[sb]);
      });
      __syncwarp();
    };
    for (int w = cid; w < n_items; w += ncl, ++item) {
      const int split = w % p.nsplit;
      PROF_T(4, mbar_wait(xfull, item & 1));
      tc_fence_after();
      const int ns = (p.S - split + p.nsplit - 1) / p.nsplit;
      for (int js = 0; js < ns; ++js, ++jsamp) {
        for (int h = 0; h < p.hc; ++h, ++it) {   // h counts chunks in issue order (the producer rotates the actual chunk)
          const int buf = it & 1;
          const uint32_t tmem_d = tmem_base + COL_ACC + buf * BN;
          for (int k = 0; k < p.kc; k += p.ks) {
            PROF_T(5, mbar_wait(&full[st], ph));
            tc_fence_after();
            const uint32_t a0 = tmem_base + COL_X + k * (BK / 2), b0 = smem_u32(sB + st * STAGE_BYTES);
            const uint32_t acc0 = k != 0;
            PROF_T(6, if (elect_one()) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j)
                umma_bf16_ts(tmem_d, a0 + j * 8, make_desc_sw128(b0 + j * 32), idesc1, j ? 1u : acc0);
              if (p.ks == 2) {
#pragma unroll
                for (int j = 0; j < BK / 16; ++j)
                  umma_bf16_ts(tmem_d, a0 + BK / 2 + j * 8, make_desc_sw128(b0 + KCH_BYTES + j * 32), idesc1, 1u);
              }
              if (CL > 1) umma_commit_mc(&empty[st], (uint16_t)3); else umma_commit(&empty[st]);
              if (k + p.ks >= p.kc) umma_commit(&tfull[buf]);
            });
            __syncwarp();
            if (++st == BST) { st = 0; ph ^= 1; }
          }
          if (pend_it >= 0) gemm2(pend_it, pend_h, pend_js);
          pend_it = it; pend_h = h; pend_js = jsamp;
        }
      }
      // the epilogue warps store the next X tile only after they have finished this item's last sample, which needs the
      // pending GEMM2: flush it here (one short bubble per work item)
      if (pend_it >= 0) { gemm2(pend_it, pend_h, pend_js); pend_it = -1; }
      if (elect_one()) umma_commit(xempty);  // every GEMM1 that reads this X tile has completed when this arrives
      __syncwarp();
    }
    PROF_OUT(1);
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------------- epilogue warps
    const int q = warp & 3, part = (warp - 4) >> 2;  // TMEM lane quarter; which half of the 128 hidden columns
    const int rl = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    float nll_sum = 0.f, correct = 0.f;
    int it = 0, jsamp = 0, item = 0;
    PROF_DECL
    for (int w = cid; w < n_items; w += ncl, ++item) {
      const int tile = (w / p.nsplit) * CL + crank, split = w % p.nsplit;   // tile may be >= n_tiles (odd tile count): no valid rows
      const int row = tile * BM + rl;
      const bool rok = row < p.n_rows;
      const int y = rok ? __ldg(p.labels + row) : 0;
      const float cwr = (rok && (p.mode == 0 || p.mode == 3)) ? (p.cw ? __ldg(p.cw + row) : 1.f) : 0.f;
      {
        // X tile -> TMEM as the A operand of GEMM1: thread (row) stores its own bf16 row, the two warps of a lane quarter
        // take one half of the K range each (column c holds K elements 2c, 2c+1)
        mbar_wait(xempty, (item & 1) ^ 1);
        tc_fence_after();
        const int qn = part < 2 ? p.D >> 6 : 0;   // 16-column groups per half row (D/2 columns per row; warps 0 and 1 of the quarter)
        const uint4* src = reinterpret_cast<const uint4*>(p.x + (size_t)row * p.D) + part * (p.D >> 4);
        for (int g = 0; g < qn; ++g) {
          uint32_t xr[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 t = rok ? __ldg(src + g * 4 + i) : make_uint4(0u, 0u, 0u, 0u);
            xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w;
          }
          tmem_st16(lane_addr + COL_X + part * (p.D >> 2) + g * 16, xr);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && part < 2) mbar_arrive(xfull);
      }
      float probs[CW];
#pragma unroll
      for (int c = 0; c < CW; ++c) probs[c] = 0.f;
      const int ns = (p.S - split + p.nsplit - 1) / p.nsplit, srot = w % ns;
      for (int js = 0; js < ns; ++js, ++jsamp) {
        const int s = split + ((js + srot) % ns) * p.nsplit;
        for (int h = 0; h < p.hc; ++h, ++it) {
          const int buf = it & 1, wi = it % WST;
          PROF_T(0, mbar_wait(&wfull[wi], (it / WST) & 1));      // b1 chunk
          PROF_T(1, mbar_wait(&tfull[buf], (it >> 1) & 1));
          tc_fence_after();
#ifdef PSVI_FN_PROF
          const long long _tw = clock64();
#endif
          const float4* b1 = reinterpret_cast<const float4*>(sB1 + wi * BN + part * CPW);
          const uint32_t taddr = lane_addr + COL_ACC + buf * BN + part * CPW;
          uint32_t pk[CPW / 2];
#pragma unroll
          for (int g = 0; g < CPW / 32; ++g) {
            float v[32];
            tmem_ld32(taddr + g * 32, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b1[g * 8 + j];
              pk[g * 16 + j * 2] = pack_relu_bf16(v[4 * j] + bb.x, v[4 * j + 1] + bb.y);
              pk[g * 16 + j * 2 + 1] = pack_relu_bf16(v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
            }
          }
          // in place: this warp has read all CPW fp32 columns it owns; the bf16 pairs go into the first CPW / 2 of them
          if (CPW == 64) tmem_st32(taddr, pk); else { tmem_st16(taddr, pk); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&hfull[buf]); mbar_arrive(&wempty[wi]); }
#ifdef PSVI_FN_PROF
          pc[2] += (unsigned long long)(clock64() - _tw);
#endif
        }
        if (part == (jsamp & (EPW - 1))) {
          // logits of this sample: softmax / NLL / mixture (the two warps of a lane quarter alternate samples)
          const int sb = jsamp & 1;
          PROF_T(3, mbar_wait(&lfull[sb], (jsamp >> 1) & 1));
          tc_fence_after();
          float lg[CW];
          tmem_ld16(lane_addr + COL_D2 + sb * (4 * CW), lg);
#pragma unroll
          for (int a = 1; a < 4; ++a) {   // GEMM2 rotates over four accumulators (no dependent chain of tiny MMAs)
            float t[CW];
            tmem_ld16(lane_addr + COL_D2 + sb * (4 * CW) + a * CW, t);
#pragma unroll
            for (int c = 0; c < CW; ++c) lg[c] += t[c];
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&lempty[sb]);
          const float* b2 = p.b2 + (size_t)s * CW;
          float mx = -INFINITY, ly = 0.f;
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            lg[c] = (lg[c] + __ldg(b2 + c)) * LOG2E;    // padding classes carry b2 = -inf
            mx = fmaxf(mx, lg[c]);
            if (c == y) ly = lg[c];
          }
          float se = 0.f;
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            lg[c] = ex2_approx(lg[c] - mx);
            se += lg[c];
          }
          if (p.mode == 0 || p.mode == 3) {
            const float nll = rok ? (mx - ly) * LN2 + logf(se) : 0.f;
            if (p.nll_out && rok) p.nll_out[(size_t)s * p.n_rows + row] = nll;
            const float t = warp_sum(cwr * nll);
            if (lane == 0 && tile < p.n_tiles) p.part[((size_t)tile * 4 + q) * p.S + s] = t;
            if (p.mode == 3 && tile < p.n_tiles) {
              // adjoint seeds of the output layer, obar[r][c] = coef_s (softmax_c - [c == y_r]) (zero for padding classes
              // and rows past the end), as bf16 in the canonical no-swizzle UMMA layout of a [128 rows x 16 classes] tile:
              // element (r, c) at (c / 8) * 2048 + r * 16 + (c % 8) * 2 bytes -- K-major operand (N = rows, K = classes)
              // with LBO = 2048 / SBO = 128, and at the same time MN-major operand (N = classes, K = rows) with LBO = 128 /
              // SBO = 2048 (psvi_fn_grad_tc.cuh reads it both ways)
              const float cf = rok ? s_w[s] * __fdividef(1.f, se) : 0.f, cy = rok ? s_w[s] : 0.f;
              uint32_t pk[CW / 2];
#pragma unroll
              for (int c = 0; c < CW; c += 2) {
                const float o0 = fmaf(cf, lg[c], (c == y) ? -cy : 0.f), o1 = fmaf(cf, lg[c + 1], (c + 1 == y) ? -cy : 0.f);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(o0, o1);
                pk[c / 2] = *reinterpret_cast<uint32_t*>(&t2);
              }
              uint4* dst = reinterpret_cast<uint4*>(p.obar + ((size_t)s * p.n_tiles + tile) * 2048);
              dst[rl] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[128 + rl] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          } else {
            const float sc = __fdividef(s_w[s], se);
#pragma unroll
            for (int c = 0; c < CW; ++c) probs[c] = fmaf(sc, lg[c], probs[c]);
          }
        }
      }
      if (p.mode != 0) {
        // combine the mixtures of the two warps of each lane quarter
        if (part > 0) {
#pragma unroll
          for (int c = 0; c < CW; ++c) s_xch[((part - 1) * CW + c) * BM + rl] = probs[c];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EPW) : "memory");
        if (part == 0) {
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int e = 0; e < EPW - 1; ++e) probs[c] += s_xch[(e * CW + c) * BM + rl];
          if (p.nsplit > 1) {
            if (rok) {
              float* o = p.probs_out + ((size_t)split * p.n_rows + row) * CW;
#pragma unroll
              for (int c = 0; c < CW; c += 4)
                *reinterpret_cast<float4*>(o + c) = make_float4(probs[c], probs[c + 1], probs[c + 2], probs[c + 3]);
            }
          } else if (rok) {
            float tot = 0.f, best = -1.f, py = 0.f;
            int am = 0;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
              tot += probs[c];
              if (c == y) py = probs[c];
              if (probs[c] > best) { best = probs[c]; am = c; }
            }
            nll_sum -= logf(fminf(fmaxf(__fdividef(py, tot), 1.1920929e-07f), 1.f - 1.1920929e-07f));
            correct += (am == y) ? 1.f : 0.f;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EPW) : "memory");   // s_xch is reused by the next item
      }
    }
    if (warp == 4) PROF_OUT(2);
    if (warp == 4 + 4 * (EPW - 1)) PROF_OUT(3);
    if (p.mode != 0 && p.nsplit == 1 && part == 0) {
      nll_sum = warp_sum(nll_sum);
      correct = warp_sum(correct);
      if (lane == 0) { s_red[q * 2] = nll_sum; s_red[q * 2 + 1] = correct; }
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // no CTA may leave while its peer can still multicast into it
  if (threadIdx.x == 0 && p.mode != 0 && p.nsplit == 1) {
    float a = 0.f, b = 0.f;
    for (int q = 0; q < 4; ++q) { a += s_red[q * 2]; b += s_red[q * 2 + 1]; }
    float* o = p.part + (size_t)blockIdx.x * 4;
    o[0] = a; o[1] = b; o[2] = 0.f; o[3] = 0.f;
  }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ weight preparation
// sigma = softplus(rho) once per call (the S samples share it), and sum_q log sigma_q (part of every sample's nkl)
__global__ void fn_sigma_kernel(const float* rho, int P, float* sigma, double* logsig_part) {
  double acc = 0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < P; q += gridDim.x * blockDim.x) {
    const float sg = softplus_f(rho[q]);
    sigma[q] = sg;
    acc += (double)logf(sg);
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) logsig_part[blockIdx.x] = red[0];
}

// theta_s = mu + sigma * eps_s for every sample (grid.y = s), four consecutive TL indices per thread (one Philox call):
//   W1b bf16 [S][H][D];  W2b bf16 [S][16][H] (class-major rows, zero padding);  b1 f32 [S][H];  b2 f32 [S][16] (-inf padding)
//   nkl_part [S][NKL_BLOCKS] = partial sums of (-theta^2 + eps^2) / 2   (neural_net.py:110-115; log sigma added later)
__global__ void fn_prep_kernel(const float* mu, const float* sigma, psvi_noise noise, int slab, int S, int D, int H, int C,
                               __nv_bfloat16* W1b, __nv_bfloat16* W2b, float* b1, float* b2, double* nkl_part) {
  const int s = blockIdx.y;
  const int P = H * D + H + C * H + C;
  const int G4 = (P + 3) >> 2;
  const int o_b1 = H * D, o_w2 = o_b1 + H, o_b2 = o_w2 + C * H;
  double acc = 0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < G4; g += gridDim.x * blockDim.x) {
    const int q0 = g << 2;
    float e[4];
    if (noise.mode == PSVI_NOISE_PHILOX) {
      philox_normal4(noise.seed, noise.domain, (uint32_t)slab, (uint32_t)s, (uint32_t)g, e);
    } else {
      const float* ep = noise.eps + ((size_t)slab * S + s) * P + q0;
#pragma unroll
      for (int i = 0; i < 4; ++i) e[i] = q0 + i < P ? ep[i] : 0.f;
    }
    float th[4];
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + i;
      th[i] = q < P ? fmaf(sigma[q], e[i], mu[q]) : 0.f;
      if (q < P) a += 0.5f * (e[i] * e[i] - th[i] * th[i]);
    }
    acc += (double)a;
    if (q0 < o_b1) {            // W1[h][d], 4 consecutive d
      __nv_bfloat162 lo = __floats2bfloat162_rn(th[0], th[1]), hi = __floats2bfloat162_rn(th[2], th[3]);
      *reinterpret_cast<uint2*>(W1b + (size_t)s * H * D + q0) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else if (q0 < o_w2) {     // b1[h]
      *reinterpret_cast<float4*>(b1 + (size_t)s * H + (q0 - o_b1)) = make_float4(th[0], th[1], th[2], th[3]);
    } else if (q0 < o_b2) {     // W2[c][h], 4 consecutive h
      __nv_bfloat162 lo = __floats2bfloat162_rn(th[0], th[1]), hi = __floats2bfloat162_rn(th[2], th[3]);
      *reinterpret_cast<uint2*>(W2b + (size_t)s * CW * H + (q0 - o_w2)) =
          make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else {                    // b2[c]
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (q0 + i < P) b2[(size_t)s * CW + (q0 + i - o_b2)] = th[i];
    }
  }
  // zero / -inf padding of the class dimension
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (CW - C) * H; i += gridDim.x * blockDim.x)
    W2b[(size_t)s * CW * H + (size_t)C * H + i] = __float2bfloat16(0.f);
  if (blockIdx.x == 0 && (int)threadIdx.x < CW - C) b2[(size_t)s * CW + C + threadIdx.x] = -INFINITY;
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) nkl_part[(size_t)s * NKL_BLOCKS + blockIdx.x] = red[0];
}

// nkl[s] = sum of partials + sum log sigma
__global__ void fn_nkl_kernel(const double* nkl_part, const double* logsig_part, int n_logsig, int S, float* nkl) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double a = 0;
  for (int i = 0; i < n_logsig; ++i) a += logsig_part[i];
  for (int i = 0; i < NKL_BLOCKS; ++i) a += nkl_part[(size_t)s * NKL_BLOCKS + i];
  nkl[s] = (float)a;
}

// a = N f(v): coreset weights (psvi_classes.py:111, :1358-1360, :1486-1488), one block
__global__ void fn_coreset_weights_kernel(const float* v, int M, float N, int vmode, float alpha, float* a) {
  __shared__ float red[256];
  if (vmode == PSVI_VMODE_IDENTITY) {
    for (int m = threadIdx.x; m < M; m += blockDim.x) a[m] = N * v[m];
    return;
  }
  float mx = -INFINITY;
  for (int m = threadIdx.x; m < M; m += blockDim.x) mx = fmaxf(mx, v[m]);
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  float se = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) se += expf(v[m] - mx);
  red[threadIdx.x] = se;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  se = red[0];
  const float sc = N * (vmode == PSVI_VMODE_EXPALPHA_SOFTMAX ? expf(alpha) : 1.f) / se;
  for (int m = threadIdx.x; m < M; m += blockDim.x) a[m] = sc * expf(v[m] - mx);
}

// mode 0: out[s] = sum over (tile, quarter) partials (+ add[s]), fixed order
// one block per sample: thread i sums the partials i, i + 256, ... in order, then a fixed-order sum over the threads
// (deterministic, no atomics)
__global__ void fn_sum_tiles_kernel(const float* part, int n_tiles, int S, const float* add, float* out) {
  __shared__ double red[256];
  const int s = blockIdx.x;
  double a = 0.0;
  for (int t = threadIdx.x; t < n_tiles * 4; t += 256) a += (double)part[(size_t)t * S + s];
  red[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = add ? (double)add[s] : 0.0;
    for (int i = 0; i < 256; ++i) tot += red[i];
    out[s] = (float)tot;
  }
}

// mode >= 1 with sample splits: combine the per-split partial mixtures, NLL / argmax per row, fixed-order block partials
__global__ void fn_finalize_kernel(const float* probs, int nsplit, int n_rows, int C, const int* labels, float* part) {
  float nll = 0.f, corr = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
    float pr[CW];
    for (int c = 0; c < CW; ++c) pr[c] = 0.f;
    for (int k = 0; k < nsplit; ++k)
      for (int c = 0; c < CW; ++c) pr[c] += probs[((size_t)k * n_rows + r) * CW + c];
    const int y = labels[r];
    float tot = 0.f, best = -1.f, py = 0.f;
    int am = 0;
    for (int c = 0; c < CW; ++c) {
      tot += pr[c];
      if (c == y) py = pr[c];
      if (c < C && pr[c] > best) { best = pr[c]; am = c; }
    }
    nll -= logf(fminf(fmaxf(py / tot, 1.1920929e-07f), 1.f - 1.1920929e-07f));
    corr += (am == y) ? 1.f : 0.f;
  }
  __shared__ float red[2][8];
  nll = warp_sum(nll); corr = warp_sum(corr);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = nll; red[1][threadIdx.x >> 5] = corr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += red[0][i]; b += red[1][i]; }
    part[blockIdx.x * 4] = a; part[blockIdx.x * 4 + 1] = b; part[blockIdx.x * 4 + 2] = 0.f; part[blockIdx.x * 4 + 3] = 0.f;
  }
}

// out[0..2] = (sum nll, #correct, #rows); out[3..4] = importance-weight entropy / normalised ESS (psvi_classes.py:1085-1092)
__global__ void fn_reduce_kernel(const float* part, int n, int n_rows, const float* lw, int S, float* out) {
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < n; ++i) { a += part[4 * i]; b += part[4 * i + 1]; }
    out[0] = (float)a; out[1] = (float)b; out[2] = (float)n_rows;
    if (lw) {
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

// ------------------------------------------------------------------------------------------------ host side
struct FnScratch {
  __nv_bfloat16 *W1b, *W2b, *ub;
  float *b1, *b2, *sigma, *nkl, *lw, *a, *part, *probs;
  double *nkl_part, *logsig_part;
  size_t total;
};
constexpr int LOGSIG_BLOCKS = 128;

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int split_for(int tiles, int S, int sms) {
  int nsplit = tiles >= sms ? 1 : sms / tiles;
  if (nsplit > S) nsplit = S;
  return nsplit < 1 ? 1 : nsplit;
}

// carve the caller's scratch: max_rows bounds the rows of any single forward (pseudo-data or data rows)
void carve(const psvi_mf_model* model, int64_t max_rows, int M, uint8_t* base, FnScratch& sc) {
  const size_t S = model->mc_samples, D = model->dims[0], H = model->dims[1], C = model->dims[2];
  const size_t P = H * D + H + C * H + C;
  const size_t tiles = (size_t)((max_rows + BM - 1) / BM);
  const int nsplit = split_for((int)tiles, (int)S, 148);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align256(bytes); return p; };
  sc.W1b = reinterpret_cast<__nv_bfloat16*>(take(S * H * D * 2));
  sc.W2b = reinterpret_cast<__nv_bfloat16*>(take(S * CW * H * 2));
  sc.ub = reinterpret_cast<__nv_bfloat16*>(take((size_t)(M > 0 ? M : 1) * D * 2));
  sc.b1 = reinterpret_cast<float*>(take(S * H * 4));
  sc.b2 = reinterpret_cast<float*>(take(S * CW * 4));
  sc.sigma = reinterpret_cast<float*>(take(P * 4));
  sc.nkl = reinterpret_cast<float*>(take(64 * 4));
  sc.lw = reinterpret_cast<float*>(take(64 * 4));
  sc.a = reinterpret_cast<float*>(take((size_t)(M > 0 ? M : 1) * 4));
  const size_t part_floats = tiles * 4 * S > 4096 ? tiles * 4 * S : 4096;
  sc.part = reinterpret_cast<float*>(take(part_floats * 4));
  sc.probs = reinterpret_cast<float*>(take(nsplit > 1 ? (size_t)nsplit * max_rows * CW * 4 : 256));
  sc.nkl_part = reinterpret_cast<double*>(take(S * NKL_BLOCKS * 8));
  sc.logsig_part = reinterpret_cast<double*>(take(LOGSIG_BLOCKS * 8));
  sc.total = off;
}

int check_model(const psvi_mf_model* model) {
  PSVI_REQUIRE(model, PSVI_ERR_INVALID, "null model");
  PSVI_REQUIRE(model->n_layers == 2, PSVI_ERR_UNSUPPORTED, "the tensor-core sampled-GEMM forward covers fn with one hidden layer");
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  PSVI_REQUIRE(D % BK == 0 && D >= BK && D <= 256, PSVI_ERR_UNSUPPORTED, "D=%d must be a multiple of 64 in [64, 256]", D);
  PSVI_REQUIRE(H % BN == 0 && H >= BN, PSVI_ERR_UNSUPPORTED, "H=%d must be a multiple of 128", H);
  PSVI_REQUIRE(C >= 1 && C <= CW && S >= 1 && S <= 64, PSVI_ERR_UNSUPPORTED, "need C <= 16 and S <= 64 (got C=%d S=%d)", C, S);
  return PSVI_OK;
}

int prepare(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho, int slab,
            const FnScratch& sc, cudaStream_t stream) {
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  const int P = H * D + H + C * H + C;
  fn_sigma_kernel<<<LOGSIG_BLOCKS, 256, 0, stream>>>(rho, P, sc.sigma, sc.logsig_part);
  PSVI_CUDA_CHECK(cudaGetLastError());
  fn_prep_kernel<<<dim3(NKL_BLOCKS, S), 256, 0, stream>>>(mu, sc.sigma, *noise, slab, S, D, H, C, sc.W1b, sc.W2b, sc.b1, sc.b2,
                                                          sc.nkl_part);
  PSVI_CUDA_CHECK(cudaGetLastError());
  fn_nkl_kernel<<<1, 64, 0, stream>>>(sc.nkl_part, sc.logsig_part, LOGSIG_BLOCKS, S, sc.nkl);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

// one forward over `n_rows` rows with prepared weights
//   mode 0: out_s[S] = sum_r cw[r] nll[s, r] + add[s]   (nll_out optional);   mode 1 / 2: out[0..4] predictive metrics
int forward(const psvi_mf_model* model, const FnScratch& sc, const void* x_bf16, const int32_t* labels, const float* cw,
            int64_t n_rows, int mode, const float* lw, const float* add, float* out, float* nll_out, cudaStream_t stream,
            __nv_bfloat16* obar = nullptr) {
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  int dev = 0, sms = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles = (int)((n_rows + BM - 1) / BM);
  // optional (PSVI_FN_CLUSTER=2): clusters of two CTAs sharing every W1 stage by TMA multicast.  Measured on B200: correct, but not
  // faster (1.09 vs 1.17 PFLOP/s at 131 k rows) -- the L2 already de-duplicates the concurrent reads and the two CTAs run in
  // lockstep -- so one CTA per SM without clusters stays the default.
  static const int cl_env = getenv("PSVI_FN_CLUSTER") ? atoi(getenv("PSVI_FN_CLUSTER")) : 1;
  const int CLv = (cl_env == 2 && tiles >= 2) ? 2 : 1;
  const int groups = (tiles + CLv - 1) / CLv;
  const int nsplit = split_for(groups, S, (sms < 148 ? sms : 148) / CLv);
  CUtensorMap map_w1, map_w2;
  int rc = make_map_2d_bf16(&map_w1, sc.W1b, (uint64_t)D, (uint64_t)S * H, BK, BN / CLv);
  if (rc) return rc;
  rc = make_map_2d_bf16(&map_w2, sc.W2b, (uint64_t)H, (uint64_t)S * CW, BK, CW);
  if (rc) return rc;
  FnParams p;
  memset(&p, 0, sizeof(p));
  p.n_rows = (int)n_rows; p.n_tiles = tiles; p.D = D; p.kc = D / BK; p.ks = (p.kc % 2 == 0) ? 2 : 1; p.H = H; p.hc = H / BN; p.S = S; p.nsplit = nsplit;
  p.mode = mode; p.x = static_cast<const __nv_bfloat16*>(x_bf16);
  p.b1 = sc.b1; p.b2 = sc.b2; p.cw = cw; p.lw = lw; p.labels = labels; p.nll_out = nll_out; p.part = sc.part;
  p.probs_out = sc.probs;
  p.obar = obar;
  const int items = groups * nsplit;
  const int grid = CLv * (items < sms / CLv ? items : sms / CLv);
#ifdef PSVI_FN_PROF
  static unsigned long long* prof_buf = nullptr;
  const bool prof = getenv("PSVI_FN_PROF") != nullptr;
  if (prof && !prof_buf) cudaMalloc(&prof_buf, 256 * 32 * 8);
  if (prof) cudaMemsetAsync(prof_buf, 0, 256 * 32 * 8, stream);
  p.prof = prof ? prof_buf : nullptr;
#endif
  const size_t smem = (size_t)BST * STAGE_BYTES + WST * W2_STAGE_BYTES + WST * BN * 4 + 64 * 4 + (EPW - 1) * CW * BM * 4 + 8 * 4 +
                      (2 + 2 * BST + 2 * WST + 8) * 8 + 16 + 1024;
  if (CLv == 2) {
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_fn_forward_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(FN_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PSVI_CUDA_CHECK(cudaLaunchKernelEx(&cfg, psvi_fn_forward_tc_kernel<2>, map_w1, map_w2, p));
  } else {
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(psvi_fn_forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psvi_fn_forward_tc_kernel<1><<<grid, FN_THREADS, smem, stream>>>(map_w1, map_w2, p);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
#ifdef PSVI_FN_PROF
  if (prof) {
    static unsigned long long hb[256 * 32];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hb, prof_buf, sizeof(hb), cudaMemcpyDeviceToHost);
    const char* names[4] = {"producer: wempty empty - - - - - total", "mma: wfull lempty hfull g2issue xfull full g1issue total",
                            "epi w4: wfull tfull work lfull - - - total", "epi w8: same"};
    for (int b = 0; b < grid; b += (grid > 4 ? grid / 2 : 1))
      for (int r = 0; r < 4; ++r) {
        fprintf(stderr, "[prof cta %d] %s:", b, names[r]);
        for (int i = 0; i < 8; ++i) fprintf(stderr, " %llu", hb[b * 32 + r * 8 + i]);
        fprintf(stderr, "\n");
      }
  }
#endif
  if (mode == 0 || mode == 3) {
    fn_sum_tiles_kernel<<<S, 256, 0, stream>>>(sc.part, tiles, S, add, out);
  } else if (nsplit > 1) {
    const int fb = (int)((n_rows + 255) / 256) < 1024 ? (int)((n_rows + 255) / 256) : 1024;
    fn_finalize_kernel<<<fb, 256, 0, stream>>>(sc.probs, nsplit, (int)n_rows, C, labels, sc.part);
    fn_reduce_kernel<<<1, 32, 0, stream>>>(sc.part, fb, (int)n_rows, mode == 1 ? lw : nullptr, S, out);
  } else {
    fn_reduce_kernel<<<1, 32, 0, stream>>>(sc.part, grid, (int)n_rows, mode == 1 ? lw : nullptr, S, out);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

#include "psvi_fn_grad_tc.cuh"

// out[0..2] += slab[0..2] (nll sum, correct, rows); out[3..4] = the slab's importance-weight diagnostics (last slab wins, Q12)
__global__ void slab_accumulate_kernel(const float* slab, float* out) {
  const int i = threadIdx.x;
  if (i < 3) out[i] += slab[i];
  else if (i < 5) out[i] = slab[i];
}

}  // namespace

// defined in psvi_lr_tc.cu
extern "C" int psvi_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

extern "C" {

size_t psvi_fn_data_grad_tc_scratch_bytes(const psvi_mf_model* model, int64_t n_rows) {
  if (!model || model->n_layers != 2 || n_rows <= 0) return 0;
  GradScratch g;
  carve_grad(model, n_rows, nullptr, g);
  return g.total + 256;
}

int psvi_fn_data_grad_tc(const psvi_mf_model* model, const float* theta, const void* x_bf16, const int32_t* y, int64_t n_rows,
                         const float* coef, float* dsum, float* tbar, void* scratch, void* stream) {
  PSVI_REQUIRE(model && theta && x_bf16 && y && coef && dsum && tbar && scratch, PSVI_ERR_INVALID, "null pointer");
  int rc = check_model(model);
  if (rc) return rc;
  PSVI_REQUIRE(n_rows > 0 && n_rows < (1ll << 31), PSVI_ERR_INVALID, "bad n_rows");
  PSVI_REQUIRE((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, PSVI_ERR_INVALID, "x_bf16 must be 16-byte aligned");
  return data_grad(model, theta, x_bf16, y, coef, n_rows, dsum, tbar, scratch, (cudaStream_t)stream);
}

size_t psvi_fn_tc_scratch_bytes(const psvi_mf_model* model, int64_t max_rows, int32_t M) {
  if (!model || model->n_layers != 2 || max_rows <= 0) return 0;
  FnScratch sc;
  carve(model, max_rows > M ? max_rows : M, M, nullptr, sc);
  return sc.total + 256;
}

int psvi_fn_predictive_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                          const float* u, const int32_t* z, const float* v, int32_t M, const void* xt_bf16,
                          const int32_t* yt, int64_t n_rows, int32_t slab, float N, int32_t vmode, float alpha,
                          int32_t mode, float* out, void* scratch, void* stream_) {
  PSVI_REQUIRE(model && noise && mu && rho && xt_bf16 && yt && out && scratch, PSVI_ERR_INVALID, "null pointer");
  int rc = check_model(model);
  if (rc) return rc;
  PSVI_REQUIRE(mode == 0 || mode == 1, PSVI_ERR_INVALID, "mode must be 0 (importance weighted) or 1 (uniform)");
  PSVI_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) && slab >= 0, PSVI_ERR_INVALID, "bad n_rows / slab");
  PSVI_REQUIRE((reinterpret_cast<uintptr_t>(xt_bf16) & 15) == 0, PSVI_ERR_INVALID, "xt_bf16 must be 16-byte aligned");
  PSVI_REQUIRE(noise->mode == PSVI_NOISE_PHILOX || noise->eps, PSVI_ERR_INVALID, "external noise without eps");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
  FnScratch sc;
  carve(model, n_rows > M ? n_rows : M, M, base, sc);
  rc = prepare(model, noise, mu, rho, slab, sc, stream);
  if (rc) return rc;
  if (mode == 0) {
    // log importance weights from the pseudo-data forward:  lw_s = sum_m a_m nll[s, m] + nkl_s   (sign quirk Q3)
    PSVI_REQUIRE(u && z && v && M > 0, PSVI_ERR_INVALID, "importance-weighted mode needs pseudo-data");
    rc = psvi_f32_to_bf16(u, sc.ub, (int64_t)M * model->dims[0], stream_);
    if (rc) return rc;
    fn_coreset_weights_kernel<<<1, 256, 0, stream>>>(v, M, N, vmode, alpha, sc.a);
    PSVI_CUDA_CHECK(cudaGetLastError());
    rc = forward(model, sc, sc.ub, z, sc.a, M, 0, nullptr, sc.nkl, sc.lw, nullptr, stream);
    if (rc) return rc;
  }
  return forward(model, sc, xt_bf16, yt, nullptr, n_rows, mode == 0 ? 1 : 2, sc.lw, nullptr, out, nullptr, stream);
}

// defined in psvi_lr_tc.cu
int psvi_lr_predictive_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho, const float* u,
                          const int32_t* z, const float* v, int32_t M, const void* xt_bf16, const int32_t* yt, int64_t n_rows,
                          int32_t slab, float N, int32_t vmode, float alpha, int32_t mode, float* out, void* scratch, void* stream_);

int psvi_predictive_tc_slabs(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                             const float* u, const int32_t* z, const float* v, int32_t M, const void* xt_bf16,
                             const int32_t* yt, int64_t n_rows, int32_t batch, int32_t first_slab, float N, int32_t vmode,
                             float alpha, int32_t mode, float* out, void* scratch, void* stream_) {
  PSVI_REQUIRE(model && xt_bf16 && yt && out && scratch, PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(n_rows > 0 && batch > 0 && first_slab >= 0, PSVI_ERR_INVALID, "bad n_rows / batch / first_slab");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
  float* tmp = reinterpret_cast<float*>(base);
  const int D = model->dims[0];
  PSVI_CUDA_CHECK(cudaMemsetAsync(out, 0, 8 * sizeof(float), stream));
  int k = 0;
  for (int64_t a0 = 0; a0 < n_rows; a0 += batch, ++k) {
    const int64_t rows = n_rows - a0 < batch ? n_rows - a0 : batch;
    const void* xs = static_cast<const uint8_t*>(xt_bf16) + (size_t)a0 * D * 2;
    const int rc = model->n_layers == 1
                       ? psvi_lr_predictive_tc(model, noise, mu, rho, u, z, v, M, xs, yt + a0, rows, first_slab + k, N, vmode, alpha,
                                               mode, tmp, base + 256, stream_)
                       : psvi_fn_predictive_tc(model, noise, mu, rho, u, z, v, M, xs, yt + a0, rows, first_slab + k, N, vmode, alpha,
                                               mode, tmp, base + 256, stream_);
    if (rc) return rc;
    slab_accumulate_kernel<<<1, 32, 0, stream>>>(tmp, out);
    PSVI_CUDA_CHECK(cudaGetLastError());
  }
  return PSVI_OK;
}

int psvi_fn_nll_tc(const psvi_mf_model* model, const psvi_noise* noise, const float* mu, const float* rho,
                   const void* x_bf16, const int32_t* labels, const float* row_weights, int64_t n_rows, int32_t slab,
                   float* wsum_out, float* nkl_out, float* nll_out, void* scratch, void* stream_) {
  PSVI_REQUIRE(model && noise && mu && rho && x_bf16 && labels && wsum_out && scratch, PSVI_ERR_INVALID, "null pointer");
  int rc = check_model(model);
  if (rc) return rc;
  PSVI_REQUIRE(n_rows > 0 && n_rows < (1ll << 31) && slab >= 0, PSVI_ERR_INVALID, "bad n_rows / slab");
  PSVI_REQUIRE((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, PSVI_ERR_INVALID, "x_bf16 must be 16-byte aligned");
  PSVI_REQUIRE(noise->mode == PSVI_NOISE_PHILOX || noise->eps, PSVI_ERR_INVALID, "external noise without eps");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~(uintptr_t)255);
  FnScratch sc;
  carve(model, n_rows, 0, base, sc);
  rc = prepare(model, noise, mu, rho, slab, sc, stream);
  if (rc) return rc;
  rc = forward(model, sc, x_bf16, labels, row_weights, n_rows, 0, nullptr, nullptr, wsum_out, nll_out, stream);
  if (rc) return rc;
  if (nkl_out) PSVI_CUDA_CHECK(cudaMemcpyAsync(nkl_out, sc.nkl, model->mc_samples * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  return PSVI_OK;
}

}  // extern "C"
