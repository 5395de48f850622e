// psvi_fn_large.cu -- the per-sample network pass of `fn` with ONE hidden layer in the LARGE regime (SURVEY.md section 7 /
// 8d, BASELINE config 5: D = 256, H = 1024, S = 64, M = 1000) on the Blackwell tensor path.  Same contract as psvi_net_pass
// (forward / gradient / dual Hessian-vector pass on externally supplied sampled weights theta [S][P]), so the streaming PSVI
// engine (psvi/inference/stream.py) drives inner_elbo / psvi_elbo / nested_step for models whose per-sample weights fit no
// CTA.  Reference: VILinear.forward + nn.ReLU + Categorical.log_prob + autograd / double backward
// (psvi/models/neural_net.py:155-179,267-297; psvi_classes.py:445-511,541-600; robust_higher/optim.py:224-229).
//
// Every matrix product of the pass is ONE launch of a batched "TN" GEMM kernel, C[b] = A[b] B[b]^T with
// K-major operands and fp32 accumulation: TMA (128-byte swizzle) -> shared-memory ring -> tcgen05.mma (M = N = 128) ->
// double-buffered TMEM accumulator -> 8 epilogue warps (tcgen05.ld, bias / ReLU / ReLU-mask, plain fp32 stores, operand
// stores and transposed operand stores).  Three arithmetic modes:
//   precision 0 (bf16)    operands rounded to bf16, one kind::f16 MMA per K step: full tensor rate, ~1e-2 relative error per
//                         pass -- fine for values, first-order training (mfvi) and prediction;
//   precision 1 (tf32x3)  every operand is kept as an fp32 pair (hi, lo) of TF32-representable numbers (hi = rna_tf32(x),
//                         lo = rna_tf32(x - hi)) and every K step issues three kind::tf32 MMAs (hi hi + hi lo + lo hi):
//                         fp32-class accuracy (~2^-21) at 1/6 of the bf16 rate.  The unrolled hypergradient needs it:
//                         the reverse sweep through Adam divides by |g_i|, so per-coordinate gradient errors of bf16 size
//                         destroy it (measured in DESIGN.md 4.8).
//   precision 2 (bf16x3)  the same split with BF16 numbers (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits, ~2^-17) and
//                         three kind::f16 MMAs per K step: twice the MMA rate of tf32x3 and half its operand bytes;
//                         hypergradients within 1e-2 .. 1e-1 (cosine >= 0.999) of the fp64 oracle -- opt-in.
// Sums of two products are one GEMM over a concatenated K dimension ([hdot | h] [W2 | W2dot]^T etc.); products with
// K = C <= 16 use zero-padded K blocks (one 128-byte swizzle row).  Activations that feed a later GEMM as the K dimension
// are stored transposed by the producing epilogue.  Reductions over rows with a C- or 1-wide output (second-layer and
// bias gradients) are CUDA-core column reductions.
#include "psvi_tc.cuh"

using namespace psvi_tc;

namespace {

constexpr int GM = 128, GN = 128;
constexpr int G_THREADS = 128 + 256;            // 4 role warps + 8 epilogue warps
constexpr int G_TILE_BYTES = GM * 128;          // 16 KB: 128 rows of one 128-byte swizzle row
constexpr int CW = 16;                          // classes padded to 16 on the fp32 side

template <int X3> struct Prec;
template <> struct Prec<0> { static constexpr int ES = 2, GK = 64, TILES = 2, GST = 6; };   // bf16
template <> struct Prec<1> { static constexpr int ES = 4, GK = 32, TILES = 4, GST = 3; };   // tf32 (hi, lo) pairs
template <> struct Prec<2> { static constexpr int ES = 2, GK = 64, TILES = 4, GST = 3; };   // bf16 (hi, lo) pairs ("bf16x3")

struct GemmP {
  int batch, m_tiles, n_tiles, kc;              // kc = K / GK
  int M_valid, N_valid;
  int a_brows, b_brows;                         // rows per batch in the A / B tensor maps (0: operand shared by all batches)
  const float* bias; long long bias_bs;         // + bias[b * bias_bs + n]   (bias_row: + bias[b * bias_bs + m])
  int bias_row, relu;
  const void* mask; long long mask_bs; int mask_ld;             // * (mask[b][m][n] > 0)   (bf16, or the fp32 hi part)
  int mask_t;                                   // the mask array is [b][n][m]: read it transposed
  // ReLU mask as BITS: word [b][row][hidden / 32], bit = hidden % 32.  mbits_out: written by the role-swapped forward product
  // (accumulator lanes = hidden units, columns = rows: one warp ballot per row); mbits: read by the products whose epilogue
  // multiplies by (h > 0) (accumulator lanes = rows, columns = hidden units: ONE 32-bit load per 32-column chunk instead of
  // 64 / 128 bytes of h per thread).  mbits_ld = words per row, mbits_bs = words per batch
  const uint32_t* mbits; uint32_t* mbits_out; long long mbits_bs; int mbits_ld;
  int n_pad;                                    // > N_valid: columns N_valid .. n_pad - 1 are processed too and written as zeros
  float* of; long long of_bs; int of_ld;                       // fp32 out [b][m][n], m < M_valid, n < N_valid
  void *ob, *ob_lo; long long ob_bs; int ob_ld; int ob_rows;    // operand out [b][m][n], m < ob_rows (zeros for m >= M_valid)
  void *obt, *obt_lo; long long obt_bs; int obt_ld;             // transposed operand out [b][n][m], m < ob_rows
  int a_mn, b_mn;                               // split-bf16 only: the operand is read "MN-major" -- its global array is
                                                // [K rows][M (or N) contiguous], i.e. the product uses the TRANSPOSE of a
                                                // row-major array without a transposed copy (UMMA a_major / b_major = 1)
  float* colpart; long long colpart_bs; int colpart_ld;   // column sums of the (masked) output per 32-row group:
                                                // colpart[b][mt * 4 + q][n] = sum over the 32 rows of epilogue warp q of row tile
                                                // mt (fixed order); a finish kernel adds the groups -- the bias adjoint of the
                                                // gradient pass without a second sweep over the adjoint array
  int obt_stage;                                // split-bf16 only: the transposed operand store is staged in shared memory
                                                // (16-byte pieces); needs ob_rows % 32 == 0 and 16-byte aligned rows
  int ob_tma;                                   // split-bf16 only: the row-major operand store goes through shared memory and
                                                // cp.async.bulk.tensor (TMA) stores described by map_o / map_ol
};

constexpr int STG_WARP_BYTES = 2 * 32 * 64;     // per epilogue warp: (hi | lo) x 32 rows x 32 bf16 columns (64-byte swizzle)

// MN-major operand tile of one K block: two [64 K-rows x 64 MN-columns] boxes (128-byte swizzle, 8 KB each): LBO = 8 KB between
// the 64-column halves, SBO = 1 KB between 8-row groups along K; one K = 16 MMA step advances two row groups (2 KB)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((8192 >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a 2-cluster work on ONE 256 x 256 tile; each loads 128 rows of A and 128
//      rows of B (half the operand bytes per output of the 128 x 128 single-CTA tile), the even CTA issues the MMAs for both
//      (M = 256: 128 accumulator lanes in each CTA's tensor memory, N = 256 columns), commits are multicast to both CTAs.
template <int CG>
__device__ __forceinline__ void umma_tf32_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    umma_tf32(tmem_d, adesc, bdesc, idesc, acc);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_bf16_cg(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    umma_bf16(tmem_d, adesc, bdesc, idesc, acc);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if (CG == 1) {
    umma_commit(bar);
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t o;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(addr), "r"(rank));
  return o;
}
// TMA load into MY shared memory whose bytes are counted on the barrier at shared::cluster address `bar_cl` (the pair leader's)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cl, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cl), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cl) {
  // relaxed: the arrive only says "my tcgen05.ld of this accumulator have completed" (tcgen05.wait::ld ran before it); a
  // release at cluster scope would also wait for the warp's outstanding GLOBAL stores (ncu: 12 % of the samples in membar)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cl) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// x = hi + lo with hi, lo representable in TF32 (10 explicit mantissa bits), |x - hi - lo| <= 2^-22 |x|
__host__ __device__ __forceinline__ float tf32_rna(float x) {
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r & 0xFFFFE000u);
#else
  return x;
#endif
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_rna(x);
  lo = tf32_rna(x - hi);
}

// x = hi + lo with hi, lo BF16 numbers, |x - hi - lo| <= 2^-17 |x|   (precision 2)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}
// the same split for two values with the PACKED conversion (cvt.rn.bf16x2.f32: one full-rate instruction per pair; the scalar
// cvt.rn.bf16.f32 showed up as "mio" stalls in the epilogues).  Words hold (a | b << 16); results equal split_bf16's bit for bit.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

template <int X3, int CG>
__global__ void __launch_bounds__(G_THREADS, 1)
tn_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_al,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_bl,
               const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_ol, const GemmP p) {
  using PR = Prec<X3>;
  constexpr int STAGE = PR::TILES * G_TILE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ring = smem;                                    // [GST][A (| A lo) | B (| B lo)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + PR::GST * STAGE);
  uint64_t* full = bars;                 // [GST]
  uint64_t* empty = bars + PR::GST;      // [GST]
  uint64_t* tfull = empty + PR::GST;     // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // staging tiles of the TMA-stored epilogue (split-bf16): 8 warps x (hi | lo) x [32 rows][64 B], 64-byte swizzle
  uint8_t* stg = reinterpret_cast<uint8_t*>(bars) + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < PR::GST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {   // the same warp of both CTAs of the pair: 2 x 256 accumulator columns in each CTA's tensor memory
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  tc_fence_before();
  if (CG == 1) __syncthreads(); else cluster_sync_all();   // pair: the peer's barriers and tensor memory exist before use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // pair: a work item is a 256 x 256 tile; CTA `crank` of the pair owns its rows [128 crank, +128) of A (its accumulator lanes)
  // and loads rows [128 crank, +128) of the tile's B operand (accumulator columns [128 crank, +128) of BOTH CTAs)
  const int crank = CG == 2 ? (int)cluster_rank() : 0;
  const int nt_w = p.n_tiles / CG;
  const int per_b = (p.m_tiles / CG) * nt_w, n_total = p.batch * per_b;
  const int w0 = blockIdx.x / CG, wstride = gridDim.x / CG;
  constexpr int TN = GN * CG;   // accumulator columns of one work item

  if (warp == 0) {
    int st = 0;
    uint32_t ph = 0;
    for (int t = w0; t < n_total; t += wstride) {
      const int b = t / per_b, rem = t - b * per_b, mt = (rem / nt_w) * CG + crank, nt = (rem - (rem / nt_w) * nt_w) * CG + crank;
      const int arow = b * p.a_brows + mt * GM, brow = b * p.b_brows + nt * GN;
      for (int k = 0; k < p.kc; ++k) {
        mbar_wait(&empty[st], ph ^ 1);
        if (elect_one()) {
          uint8_t* dst = ring + st * STAGE;
          // pair: both CTAs' bytes are counted on the LEADER's full barrier (the MMA issuer waits there)
          if (crank == 0) mbar_expect_tx(&full[st], CG * STAGE);
          const uint32_t fbar = CG == 2 ? mapa_u32(smem_u32(&full[st]), 0) : 0u;
          auto tma_load_2d = [&](const CUtensorMap* map, uint64_t* bar, void* d, int c0, int c1) {
            if (CG == 1) psvi_tc::tma_load_2d(map, bar, d, c0, c1);
            else tma_load_2d_pair(map, fbar, d, c0, c1);
          };
          if (X3 == 2 && (p.a_mn || p.b_mn)) {
            const int ak = b * p.a_brows + k * PR::GK, bk = b * p.b_brows + k * PR::GK;
            if (p.a_mn) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                tma_load_2d(&map_a, &full[st], dst + h * 8192, mt * GM + h * 64, ak);
                tma_load_2d(&map_al, &full[st], dst + G_TILE_BYTES + h * 8192, mt * GM + h * 64, ak);
              }
            } else {
              tma_load_2d(&map_a, &full[st], dst, k * PR::GK, arow);
              tma_load_2d(&map_al, &full[st], dst + G_TILE_BYTES, k * PR::GK, arow);
            }
            if (p.b_mn) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                tma_load_2d(&map_b, &full[st], dst + 2 * G_TILE_BYTES + h * 8192, nt * GN + h * 64, bk);
                tma_load_2d(&map_bl, &full[st], dst + 3 * G_TILE_BYTES + h * 8192, nt * GN + h * 64, bk);
              }
            } else {
              tma_load_2d(&map_b, &full[st], dst + 2 * G_TILE_BYTES, k * PR::GK, brow);
              tma_load_2d(&map_bl, &full[st], dst + 3 * G_TILE_BYTES, k * PR::GK, brow);
            }
          } else if (PR::TILES == 4) {
            tma_load_2d(&map_a, &full[st], dst, k * PR::GK, arow);
            tma_load_2d(&map_al, &full[st], dst + G_TILE_BYTES, k * PR::GK, arow);
            tma_load_2d(&map_b, &full[st], dst + 2 * G_TILE_BYTES, k * PR::GK, brow);
            tma_load_2d(&map_bl, &full[st], dst + 3 * G_TILE_BYTES, k * PR::GK, brow);
          } else {
            tma_load_2d(&map_a, &full[st], dst, k * PR::GK, arow);
            tma_load_2d(&map_b, &full[st], dst + G_TILE_BYTES, k * PR::GK, brow);
          }
        }
        __syncwarp();
        if (++st == PR::GST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && crank == 0) {
    // instruction descriptor: D = f32, A = B = bf16 (format 1) or tf32 (format 2), both K-major, N = 128, M = 128 (pair: 256, 256)
    const uint32_t fmt = X3 == 1 ? 2u : 1u;
    const uint32_t idesc =
        (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((GN * CG) >> 3) << 17) | ((uint32_t)((GM * CG) >> 4) << 24);
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (int t = w0; t < n_total; t += wstride, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * TN;
      for (int k = 0; k < p.kc; ++k) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t s0 = smem_u32(ring + st * STAGE);
        const uint32_t acc0 = k != 0;
        if (elect_one()) {
          if (X3 == 1) {
            const uint32_t ah = s0, al = s0 + G_TILE_BYTES, bh = s0 + 2 * G_TILE_BYTES, bl = s0 + 3 * G_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // K = 8 fp32 = 32 bytes per step; small terms first
              umma_tf32_cg<CG>(tmem_d, make_desc_sw128(al + j * 32), make_desc_sw128(bh + j * 32), idesc, j ? 1u : acc0);
              umma_tf32_cg<CG>(tmem_d, make_desc_sw128(ah + j * 32), make_desc_sw128(bl + j * 32), idesc, 1u);
              umma_tf32_cg<CG>(tmem_d, make_desc_sw128(ah + j * 32), make_desc_sw128(bh + j * 32), idesc, 1u);
            }
          } else if (X3 == 2) {
            const uint32_t ah = s0, al = s0 + G_TILE_BYTES, bh = s0 + 2 * G_TILE_BYTES, bl = s0 + 3 * G_TILE_BYTES;
            const uint32_t id2 = idesc | ((uint32_t)(p.a_mn != 0) << 15) | ((uint32_t)(p.b_mn != 0) << 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {   // K = 16 bf16 per step: 32 bytes along K (K-major) or two 8-row groups (MN-major)
              const uint64_t dah = p.a_mn ? make_desc_mn(ah + j * 2048) : make_desc_sw128(ah + j * 32);
              const uint64_t dal = p.a_mn ? make_desc_mn(al + j * 2048) : make_desc_sw128(al + j * 32);
              const uint64_t dbh = p.b_mn ? make_desc_mn(bh + j * 2048) : make_desc_sw128(bh + j * 32);
              const uint64_t dbl = p.b_mn ? make_desc_mn(bl + j * 2048) : make_desc_sw128(bl + j * 32);
              umma_bf16_cg<CG>(tmem_d, dal, dbh, id2, j ? 1u : acc0);     // lo.hi + hi.lo + hi.hi at the bf16 rate
              umma_bf16_cg<CG>(tmem_d, dah, dbl, id2, 1u);
              umma_bf16_cg<CG>(tmem_d, dah, dbh, id2, 1u);
            }
          } else {
            const uint32_t a0 = s0, b0 = s0 + G_TILE_BYTES;
#pragma unroll
            for (int j = 0; j < 4; ++j)     // K = 16 bf16 = 32 bytes per step
              umma_bf16_cg<CG>(tmem_d, make_desc_sw128(a0 + j * 32), make_desc_sw128(b0 + j * 32), idesc, j ? 1u : acc0);
          }
          umma_commit_cg<CG>(&empty[st]);
          if (k == p.kc - 1) umma_commit_cg<CG>(&tfull[buf]);
        }
        __syncwarp();
        if (++st == PR::GST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, part = (warp - 4) >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int it = 0;
    const uint32_t tempty_cl = CG == 2 ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;   // the leader's accumulator-free barriers
    for (int t = w0; t < n_total; t += wstride, ++it) {
      const int b = t / per_b, rem = t - b * per_b, mt = (rem / nt_w) * CG + crank, ntw = rem - (rem / nt_w) * nt_w;
      const int buf = it & 1;
      const int m = mt * GM + q * 32 + lane;
      const bool vm = m < p.M_valid;
      mbar_wait(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < 2 * CG; ++g) {
        const int n0 = ntw * TN + part * (TN / 2) + g * 32;
        float v[32];
        const bool chunk_on = n0 < (p.n_pad > p.N_valid ? p.n_pad : p.N_valid);
        // loads that do not depend on the accumulator go out before the TMEM load: the bias (one value per lane, broadcast by
        // shuffles below) and the ReLU-mask word of this (row, 32 hidden units) chunk
        float bias_v = 0.f;
        uint32_t mword = 0xFFFFFFFFu;
        if (chunk_on) {
          if (p.bias && p.bias_row) bias_v = vm ? __ldg(p.bias + (long long)b * p.bias_bs + m) : 0.f;
          else if (p.bias) bias_v = (n0 + lane < p.N_valid) ? __ldg(p.bias + (long long)b * p.bias_bs + n0 + lane) : 0.f;
          if (p.mbits && vm) mword = __ldg(p.mbits + (long long)b * p.mbits_bs + (long long)m * p.mbits_ld + (n0 >> 5));
        }
        tmem_ld32(lane_addr + buf * TN + part * (TN / 2) + g * 32, v);
        if (chunk_on) {
          if (p.bias && p.bias_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bias_v;
          } else if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_v, j);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.mbits) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (!((mword >> j) & 1u)) v[j] = 0.f;
          } else if (p.mask && vm && p.mask_t) {   // mask[b][n][m]: lanes read consecutive m (coalesced)
            const long long mo = (long long)b * p.mask_bs + (long long)n0 * p.mask_ld + m;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (n0 + j >= p.N_valid) continue;
              const float mk = X3 == 1 ? __ldg(static_cast<const float*>(p.mask) + mo + (long long)j * p.mask_ld)
                                  : __bfloat162float(static_cast<const __nv_bfloat16*>(p.mask)[mo + (long long)j * p.mask_ld]);
              if (!(mk > 0.f)) v[j] = 0.f;
            }
          } else if (p.mask && vm) {   // N_valid is a multiple of 32 whenever a mask / operand output is used
            if (X3 == 1) {
              const float4* mp = reinterpret_cast<const float4*>(static_cast<const float*>(p.mask) + (long long)b * p.mask_bs +
                                                                 (long long)m * p.mask_ld + n0);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 mk = __ldg(mp + i);
                if (!(mk.x > 0.f)) v[4 * i] = 0.f;
                if (!(mk.y > 0.f)) v[4 * i + 1] = 0.f;
                if (!(mk.z > 0.f)) v[4 * i + 2] = 0.f;
                if (!(mk.w > 0.f)) v[4 * i + 3] = 0.f;
              }
            } else {
              const uint4* mp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.mask) + (long long)b * p.mask_bs +
                                                               (long long)m * p.mask_ld + n0);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 mk = __ldg(mp + i);
                const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {   // bf16 > 0  <=>  sign bit clear and not zero
                  const uint32_t lo = w[e] & 0xFFFFu, hi = w[e] >> 16;
                  if (!(lo != 0 && !(lo & 0x8000u))) v[i * 8 + e * 2] = 0.f;
                  if (!(hi != 0 && !(hi & 0x8000u))) v[i * 8 + e * 2 + 1] = 0.f;
                }
              }
            }
          }
          if (!vm) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (p.n_pad > p.N_valid && n0 + 32 > p.N_valid) {   // only the chunk that straddles N_valid (warp-uniform)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j >= p.N_valid) v[j] = 0.f;
          }
          if (p.colpart) {   // (warp-uniform; rows >= M_valid and columns >= N_valid hold zeros here)
            float* sh = reinterpret_cast<float*>(stg + (warp - 4) * STG_WARP_BYTES);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) sh[lane * 32 + (j ^ lane)] = v[j];
            __syncwarp();
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) t += sh[r * 32 + (lane ^ r)];
            if (n0 + lane < p.N_valid)
              p.colpart[(long long)b * p.colpart_bs + (long long)(mt * 4 + q) * p.colpart_ld + n0 + lane] = t;
            __syncwarp();
          }
          if (p.of) {
            // fp32 output: this warp's [32 rows x 32 columns] piece is transposed through shared memory (element (r, c) at
            // word r * 32 + (c ^ r): conflict-free both ways) so that one store instruction writes 128 contiguous bytes of
            // ONE output row -- the row-per-thread form wrote 4 bytes of 32 different rows per instruction and left the
            // epilogue warps waiting on the store queue for half of the kernel (ncu: 47 % of the samples at this line)
            float* sh = reinterpret_cast<float*>(stg + (warp - 4) * STG_WARP_BYTES);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) sh[lane * 32 + (j ^ lane)] = v[j];
            __syncwarp();
            const int mrow0 = mt * GM + q * 32;
            float* ob_ = p.of + (long long)b * p.of_bs + n0;
            if (p.N_valid <= 16) {   // 64-byte rows (of_ld = 16 keeps them contiguous): two rows per instruction
              const int c = lane & 15;
#pragma unroll 4
              for (int r = 0; r < 32; r += 2) {
                const int rr = r + (lane >> 4), mm = mrow0 + rr;
                if (mm < p.M_valid && n0 + c < p.N_valid) ob_[(long long)mm * p.of_ld + c] = sh[rr * 32 + (c ^ rr)];
              }
            } else {
              const bool vc = n0 + lane < p.N_valid;
#pragma unroll 4
              for (int r = 0; r < 32; ++r) {
                const int mm = mrow0 + r;
                if (mm < p.M_valid && vc) ob_[(long long)mm * p.of_ld + lane] = sh[r * 32 + (lane ^ r)];
              }
            }
          }
          if (p.ob && m < p.ob_rows) {
            const long long off = (long long)b * p.ob_bs + (long long)m * p.ob_ld + n0;
            if (X3 == 1) {
              float4* oh = reinterpret_cast<float4*>(static_cast<float*>(p.ob) + off);
              float4* ol = reinterpret_cast<float4*>(static_cast<float*>(p.ob_lo) + off);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float h4[4], l4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(v[4 * i + e], h4[e], l4[e]);
                oh[i] = make_float4(h4[0], h4[1], h4[2], h4[3]);
                ol[i] = make_float4(l4[0], l4[1], l4[2], l4[3]);
              }
            } else if (X3 == 2 && p.ob_tma) {
              // this warp's [32 rows x 32 columns] piece of (hi, lo): rows into swizzled shared memory (16-byte chunk c of
              // row r sits at chunk c ^ ((r >> 1) & 3): the 64-byte TMA swizzle, conflict-free for row-per-thread writes),
              // then two TMA stores; the copies of chunk g overlap the arithmetic of chunk g + 1
              uint8_t* sh = stg + (warp - 4) * STG_WARP_BYTES;
              uint8_t* sl = sh + STG_WARP_BYTES / 2;
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous copies have read it
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  split_bf16x2(v[i * 8 + e * 2], v[i * 8 + e * 2 + 1], wh[e], wl[e]);
                }
                const int phys = (i ^ ((lane >> 1) & 3)) * 16 + lane * 64;
                *reinterpret_cast<uint4*>(sh + phys) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                *reinterpret_cast<uint4*>(sl + phys) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) {
                const int row0 = b * p.ob_rows + mt * GM + q * 32;
                tma_store_2d(&map_o, sh, n0, row0);
                tma_store_2d(&map_ol, sl, n0, row0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            } else if (X3 == 2) {
              uint4* oh = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.ob) + off);
              uint4* ol = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.ob_lo) + off);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  split_bf16x2(v[i * 8 + e * 2], v[i * 8 + e * 2 + 1], wh[e], wl[e]);
                }
                oh[i] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
                ol[i] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
              }
            } else {
              uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.ob) + off);
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
          }
          if (p.mbits_out) {
            // (h > 0) for the 32 hidden units of this warp, one word per row: column j's ballot is the word of row n0 + j.
            // Lane 0 parks the 32 ballots in shared memory (128-bit stores), lane j picks up word j: two instructions per
            // column plus 8 stores (a per-column "if (lane == j)" select costs four).  Warp-uniform branch; padded hidden
            // units cannot occur (H is a multiple of 128); columns past N_valid were zeroed above.
            uint32_t* shw = reinterpret_cast<uint32_t*>(stg + (warp - 4) * STG_WARP_BYTES);
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) w4[e] = __ballot_sync(0xffffffffu, v[4 * j4 + e] > 0.f);
              if (lane == 0) *reinterpret_cast<uint4*>(shw + 4 * j4) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            __syncwarp();
            p.mbits_out[(long long)b * p.mbits_bs + (long long)(n0 + lane) * p.mbits_ld + ((mt * GM + q * 32) >> 5)] = shw[lane];
          }
          if (p.obt && m < p.ob_rows) {
            const long long off = (long long)b * p.obt_bs + (long long)n0 * p.obt_ld + m;
            if (X3 == 1) {
              float* oh = static_cast<float*>(p.obt) + off;
              float* ol = static_cast<float*>(p.obt_lo) + off;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float hi, lo;
                split_tf32(v[j], hi, lo);
                oh[(long long)j * p.obt_ld] = hi;
                ol[(long long)j * p.obt_ld] = lo;
              }
            } else if (X3 == 2 && p.obt_stage) {
              // this warp's [32 columns n][32 lanes m] piece of (hi, lo) is laid out [n][m] in shared memory (2 KB each) and
              // leaves as 16-byte pieces (8 consecutive m of one n per thread): 8 store instructions of 512 bytes instead of
              // 64 of 64 bytes (ncu: the two-byte stores were 21 % of the samples of the forward product)
              uint16_t* sh = reinterpret_cast<uint16_t*>(stg + (warp - 4) * STG_WARP_BYTES);
              uint16_t* sl = sh + 1024;
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                uint32_t wh, wl;
                split_bf16x2(v[j], v[j + 1], wh, wl);
                sh[j * 32 + lane] = (uint16_t)(wh & 0xFFFFu);
                sh[(j + 1) * 32 + lane] = (uint16_t)(wh >> 16);
                sl[j * 32 + lane] = (uint16_t)(wl & 0xFFFFu);
                sl[(j + 1) * 32 + lane] = (uint16_t)(wl >> 16);
              }
              __syncwarp();
              __nv_bfloat16* oh = static_cast<__nv_bfloat16*>(p.obt) + (long long)b * p.obt_bs + (long long)n0 * p.obt_ld + mt * GM + q * 32;
              __nv_bfloat16* ol = static_cast<__nv_bfloat16*>(p.obt_lo) + (long long)b * p.obt_bs + (long long)n0 * p.obt_ld + mt * GM + q * 32;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int piece = k * 32 + lane, n = piece >> 2, mc = piece & 3;
                *reinterpret_cast<uint4*>(oh + (long long)n * p.obt_ld + mc * 8) = *reinterpret_cast<const uint4*>(sh + n * 32 + mc * 8);
                *reinterpret_cast<uint4*>(ol + (long long)n * p.obt_ld + mc * 8) = *reinterpret_cast<const uint4*>(sl + n * 32 + mc * 8);
              }
            } else if (X3 == 2) {
              __nv_bfloat16* oh = static_cast<__nv_bfloat16*>(p.obt) + off;
              __nv_bfloat16* ol = static_cast<__nv_bfloat16*>(p.obt_lo) + off;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                __nv_bfloat16 hi, lo;
                split_bf16(v[j], hi, lo);
                oh[(long long)j * p.obt_ld] = hi;
                ol[(long long)j * p.obt_ld] = lo;
              }
            } else {
              __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.obt) + off;
#pragma unroll
              for (int j = 0; j < 32; ++j) op[(long long)j * p.obt_ld] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(&tempty[buf]);
        else mbar_arrive_cluster(tempty_cl + buf * 8);
      }
    }
    if (X3 == 2 && p.ob_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete
  }
  tc_fence_before();
  if (CG == 1) __syncthreads(); else cluster_sync_all();   // pair: no CTA leaves while its peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ operand stores
// one operand element: bf16 (X3 = 0), the (hi, lo) fp32 pair of TF32 numbers (X3 = 1) or the (hi, lo) pair of BF16 numbers (X3 = 2)
template <int X3>
__device__ __forceinline__ void put(void* hi_base, void* lo_base, size_t i, float v) {
  if (X3 == 1) {
    float h, l;
    split_tf32(v, h, l);
    static_cast<float*>(hi_base)[i] = h;
    static_cast<float*>(lo_base)[i] = l;
  } else if (X3 == 2) {
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    static_cast<__nv_bfloat16*>(hi_base)[i] = h;
    static_cast<__nv_bfloat16*>(lo_base)[i] = l;
  } else {
    static_cast<__nv_bfloat16*>(hi_base)[i] = __float2bfloat16(v);
  }
}
template <int X3>
__device__ __forceinline__ float get(const void* hi_base, const void* lo_base, size_t i) {
  if (X3 == 1) return static_cast<const float*>(hi_base)[i] + static_cast<const float*>(lo_base)[i];
  if (X3 == 2)
    return __bfloat162float(static_cast<const __nv_bfloat16*>(hi_base)[i]) + __bfloat162float(static_cast<const __nv_bfloat16*>(lo_base)[i]);
  return __bfloat162float(static_cast<const __nv_bfloat16*>(hi_base)[i]);
}

// 16 consecutive operand elements (the head of one K block; the rest of the block stays zero) at element offset i, a
// multiple of 16: 128-bit stores
template <int X3>
__device__ __forceinline__ void put16(void* hi_base, void* lo_base, size_t i, const float (&v)[16]) {
  if (X3 == 1) {
    float4* oh = reinterpret_cast<float4*>(static_cast<float*>(hi_base) + i);
    float4* ol = reinterpret_cast<float4*>(static_cast<float*>(lo_base) + i);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(v[4 * q + e], h[e], l[e]);
      oh[q] = make_float4(h[0], h[1], h[2], h[3]);
      ol[q] = make_float4(l[0], l[1], l[2], l[3]);
    }
  } else {
    uint32_t wh[8], wl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      split_bf16x2(v[2 * e], v[2 * e + 1], wh[e], wl[e]);
    }
    uint4* oh = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(hi_base) + i);
    oh[0] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
    oh[1] = make_uint4(wh[4], wh[5], wh[6], wh[7]);
    if (X3 == 2) {
      uint4* ol = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(lo_base) + i);
      ol[0] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      ol[1] = make_uint4(wl[4], wl[5], wl[6], wl[7]);
    }
  }
}

// x fp32 [R][D] -> X [Rp][D] (padding rows zero) and XT [D][Rp]
template <int X3>
__global__ void fnl_prep_x_kernel(const float* __restrict__ x, int R, int Rp, int D, void* Xh, void* Xl, void* XTh, void* XTl) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, d0 = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, d = d0 + tx;
    const float v = (r < R && d < D) ? x[(size_t)r * D + d] : 0.f;
    tile[i][tx] = v;
    if (r < Rp && d < D) put<X3>(Xh, Xl, (size_t)r * D + d, v);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, r = r0 + tx;
    if (d < D && r < Rp) put<X3>(XTh, XTl, (size_t)d * Rp + r, tile[tx][i]);
  }
}

// two consecutive operand elements at EVEN element offset i: one packed conversion and 32-bit (bf16) / 64-bit (tf32) stores
template <int X3>
__device__ __forceinline__ void put2(void* hi_base, void* lo_base, size_t i, float a, float b) {
  if (X3 == 1) {
    float ha, la, hb, lb;
    split_tf32(a, ha, la);
    split_tf32(b, hb, lb);
    *reinterpret_cast<float2*>(static_cast<float*>(hi_base) + i) = make_float2(ha, hb);
    *reinterpret_cast<float2*>(static_cast<float*>(lo_base) + i) = make_float2(la, lb);
  } else if (X3 == 2) {
    uint32_t wh, wl;
    split_bf16x2(a, b, wh, wl);
    *reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(hi_base) + i) = wh;
    *reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(lo_base) + i) = wl;
  } else {
    __nv_bfloat162 t2 = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(hi_base) + i) = *reinterpret_cast<uint32_t*>(&t2);
  }
}

// first-layer weights of theta (or thetadot) [S][P] -> W1 [S][H][D] and its transpose into W1T2 [S][D][2H] at column offset
// `toff` (0: primal, H: tangent).  64 x 64 tiles, a thread handles two consecutive elements of the contiguous dimension of
// either output (packed conversions, 128 / 256 bytes per warp store); D and H are multiples of 64.
template <int X3>
__global__ void fnl_pack_w1_kernel(const float* __restrict__ theta, long long P, int D, int H, void* W1h, void* W1l, void* Th,
                                   void* Tl, int toff) {
  __shared__ float tile[64][65];
  const int d0 = blockIdx.x * 64, h0 = blockIdx.y * 64, s = blockIdx.z, tx = threadIdx.x, ty = threadIdx.y;
  const float* th = theta + (long long)s * P;
  for (int i = ty; i < 64; i += 8) {
    const int h = h0 + i, d = d0 + 2 * tx;
    const float2 v = make_float2(th[(size_t)h * D + d], th[(size_t)h * D + d + 1]);   // (P may be odd: no 64-bit loads)
    tile[i][2 * tx] = v.x;
    tile[i][2 * tx + 1] = v.y;
    put2<X3>(W1h, W1l, ((size_t)s * H + h) * D + d, v.x, v.y);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int d = d0 + i, h = h0 + 2 * tx;
    put2<X3>(Th, Tl, ((size_t)s * D + d) * (2 * H) + toff + h, tile[2 * tx][i], tile[2 * tx + 1][i]);
  }
}

template <int X3>
__global__ void fnl_pack_w2_kernel(const float* __restrict__ theta, long long P, int D, int H, int C, void* W2ph, void* W2pl,
                                   int poff, void* W2Th, void* W2Tl, int coff, int cp2) {
  const int s = blockIdx.y;
  const float* w2 = theta + (long long)s * P + (long long)H * D + H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * H; i += gridDim.x * blockDim.x) {
    const int c = i / H, h = i - c * H;
    const float v = w2[i];
    put<X3>(W2ph, W2pl, ((size_t)s * 128 + c) * (2 * H) + poff + h, v);
    put<X3>(W2Th, W2Tl, ((size_t)s * H + h) * cp2 + coff + c, v);
  }
}

// ------------------------------------------------------------------------------------------------ softmax / NLL head
// logits o [S][Rp][16] (bias b2 still to be added) and, in dual mode, od (+ b2dot).  AA [S][Rp][2 CP]: K blocks of CP entries.
//   mode 0: nll;  mode 1: nll, go = cw (p - onehot) -> fp32 [S][R][16] and AA block 0;
//   mode 2: go = cw p (od - <p, od>) -> fp32 + AA block 1, god = cw (p - onehot) -> fp32 + AA block 0, acbar = q . od
template <int X3>
__global__ void fnl_head_kernel(const float* __restrict__ o, const float* __restrict__ od, const float* __restrict__ b2,
                                const float* __restrict__ b2d, long long P, const int* __restrict__ y, const float* __restrict__ cw,
                                int S, int R, int Rp, int C, int mode, float* __restrict__ nll, float* __restrict__ logits_out,
                                float* __restrict__ go, float* __restrict__ god, void* AAh, void* AAl, float* __restrict__ acbar) {
  constexpr int CPK = Prec<X3>::GK;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * Rp) return;
  const int s = idx / Rp, r = idx - s * Rp;
  const size_t ab = (size_t)idx * 2 * CPK;   // (AA is zeroed by the caller: only the 16-entry heads of the K blocks are written)
  if (r >= R) return;
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
  if (mode == 1) {
    float g[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) { g[c] = w * q[c]; go[oidx * CW + c] = g[c]; }
    put16<X3>(AAh, AAl, ab, g);
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
  float g1[CW], g2[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    g1[c] = c < C ? w * p[c] * (dd[c] - dot) : 0.f;
    g2[c] = w * q[c];
    go[oidx * CW + c] = g1[c];
    god[oidx * CW + c] = g2[c];
  }
  put16<X3>(AAh, AAl, ab + CPK, g1);
  put16<X3>(AAh, AAl, ab, g2);
  if (acbar) acbar[oidx] = qd;
}

// four consecutive operand elements at element offset i (a multiple of 4) as floats (hi + lo)
template <int X3>
__device__ __forceinline__ void get4(const void* hi_base, const void* lo_base, size_t i, float (&v)[4]) {
  if (X3 == 1) {
    const float4 h = *reinterpret_cast<const float4*>(static_cast<const float*>(hi_base) + i);
    const float4 l = *reinterpret_cast<const float4*>(static_cast<const float*>(lo_base) + i);
    v[0] = h.x + l.x; v[1] = h.y + l.y; v[2] = h.z + l.z; v[3] = h.w + l.w;
  } else {
    const uint2 h = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(hi_base) + i);
    v[0] = __uint_as_float(h.x << 16); v[1] = __uint_as_float(h.x & 0xFFFF0000u);
    v[2] = __uint_as_float(h.y << 16); v[3] = __uint_as_float(h.y & 0xFFFF0000u);
    if (X3 == 2) {
      const uint2 l = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(lo_base) + i);
      v[0] += __uint_as_float(l.x << 16); v[1] += __uint_as_float(l.x & 0xFFFF0000u);
      v[2] += __uint_as_float(l.y << 16); v[3] += __uint_as_float(l.y & 0xFFFF0000u);
    }
  }
}

// part[z][s][c][h] = sum over the z-th row chunk of W[s][r][c] * Y[s][r][h]   (W == NULL: C = 1, weight 1 -> column sums);
// fnl_colreduce_finish_kernel adds the chunks in fixed order: out[s * P + c * H + h] (+)= sum_z part[z][s][c][h].
// A thread owns FOUR consecutive columns (one 8- / 16-byte load per row and operand part, the row's 16 weights -- four
// broadcast 128-bit loads -- feed 64 FMAs): enough bytes in flight per SM to stream Y at HBM speed.
constexpr int RSPLIT = 8, CR_T = 64;   // row chunks; threads per CTA (CR_T * 4 columns)
template <int X3>
__global__ void __launch_bounds__(CR_T)
fnl_colreduce_kernel(const void* Yh, const void* Yl, long long y_off, long long y_bs, int y_ld, const float* __restrict__ W, int R,
                     int H, int C, float* __restrict__ part) {
  const int h_raw = (blockIdx.x * CR_T + threadIdx.x) * 4, s = blockIdx.y, zc = blockIdx.z;
  const bool live = h_raw < H;
  const int h = live ? h_raw : 0;   // (idle threads of a partial column block shadow column 0 and do not write)
  const int rpc = (R + RSPLIT - 1) / RSPLIT, r0 = zc * rpc, r1 = min(R, r0 + rpc);
  float acc[4][CW];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[k][c] = 0.f;
  const size_t y0 = (size_t)(y_off + (long long)s * y_bs + h);
  if (W) {
    // the row weights of a sub-chunk are staged in shared memory once (coalesced) and read back as broadcast 128-bit loads
    constexpr int SUB = 64;
    __shared__ float4 s_w[SUB * (CW / 4)];
    const float4* wp = reinterpret_cast<const float4*>(W + (size_t)s * R * CW);
#pragma unroll 1
    for (int rs = r0; rs < r1; rs += SUB) {
      const int nr = min(SUB, r1 - rs);
      __syncthreads();
      for (int i = threadIdx.x; i < nr * (CW / 4); i += CR_T) s_w[i] = __ldg(wp + (size_t)rs * (CW / 4) + i);
      __syncthreads();
#pragma unroll 4
      for (int rr = 0; rr < nr; ++rr) {
        float yv[4];
        get4<X3>(Yh, Yl, y0 + (size_t)(rs + rr) * y_ld, yv);
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          const float4 wv = s_w[rr * (CW / 4) + q];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[k][4 * q + 0] = fmaf(wv.x, yv[k], acc[k][4 * q + 0]);
            acc[k][4 * q + 1] = fmaf(wv.y, yv[k], acc[k][4 * q + 1]);
            acc[k][4 * q + 2] = fmaf(wv.z, yv[k], acc[k][4 * q + 2]);
            acc[k][4 * q + 3] = fmaf(wv.w, yv[k], acc[k][4 * q + 3]);
          }
        }
      }
    }
  } else {
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      float yv[4];
      get4<X3>(Yh, Yl, y0 + (size_t)r * y_ld, yv);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k][0] += yv[k];
    }
  }
#pragma unroll
  for (int c = 0; c < CW; ++c)
    if (c < C && live)
      *reinterpret_cast<float4*>(part + (((size_t)zc * gridDim.y + s) * C + c) * H + h) =
          make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
}
__global__ void fnl_colreduce_finish_kernel(const float* __restrict__ part, int S, int H, int C, float* __restrict__ out, long long P,
                                            int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;   // i = c * H + h
  if (i >= C * H) return;
  float t = 0.f;
#pragma unroll
  for (int zc = 0; zc < RSPLIT; ++zc) t += part[((size_t)zc * S + s) * C * H + i];
  float* d = out + (long long)s * P + i;
  *d = accumulate ? *d + t : t;
}

// out[s * P + h] = sum_g part[s][g][h]  (the per-row-group column sums written by the GEMM epilogue, fixed order)
__global__ void fnl_colpart_finish_kernel(const float* __restrict__ part, int G, int H, float* __restrict__ out, long long P) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (h >= H) return;
  const float* pp = part + (size_t)s * G * H + h;
  float t = 0.f;
  for (int g = 0; g < G; ++g) t += pp[(size_t)g * H];
  out[(long long)s * P + h] = t;
}

// The three weighted column reductions of the dual pass in ONE sweep over hh = [hdot | h] (each element read once):
//   part[z][s][0][c][h] = sum_r (Wo[s][r][c] h[s][r][h] + Wod[s][r][c] hdot[s][r][h])     (A_o^T h + A_od^T hdot)
//   part[z][s][1][c][h] = sum_r Wod[s][r][c] h[s][r][h]                                   (A_od^T h)
// A thread owns two consecutive columns (2 x 2 x 16 accumulators).
template <int X3>
__device__ __forceinline__ void get2(const void* hi_base, const void* lo_base, size_t i, float (&v)[2]) {
  if (X3 == 1) {
    const float2 h = *reinterpret_cast<const float2*>(static_cast<const float*>(hi_base) + i);
    const float2 l = *reinterpret_cast<const float2*>(static_cast<const float*>(lo_base) + i);
    v[0] = h.x + l.x; v[1] = h.y + l.y;
  } else {
    const uint32_t h = *reinterpret_cast<const uint32_t*>(static_cast<const __nv_bfloat16*>(hi_base) + i);
    v[0] = __uint_as_float(h << 16); v[1] = __uint_as_float(h & 0xFFFF0000u);
    if (X3 == 2) {
      const uint32_t l = *reinterpret_cast<const uint32_t*>(static_cast<const __nv_bfloat16*>(lo_base) + i);
      v[0] += __uint_as_float(l << 16); v[1] += __uint_as_float(l & 0xFFFF0000u);
    }
  }
}
constexpr int CRD_T = 128;   // threads per CTA of the dual reduction (CRD_T * 2 columns)
template <int X3>
__global__ void __launch_bounds__(CRD_T)
fnl_colreduce_dual_kernel(const void* Yh, const void* Yl, long long y_bs, const float* __restrict__ Wo,
                          const float* __restrict__ Wod, int R, int H, int C, float* __restrict__ part) {
  const int h_raw = (blockIdx.x * CRD_T + threadIdx.x) * 2, s = blockIdx.y, zc = blockIdx.z;
  const bool live = h_raw < H;
  const int h = live ? h_raw : 0;   // (idle threads of a partial column block shadow column 0 and do not write)
  const int rpc = (R + RSPLIT - 1) / RSPLIT, r0 = zc * rpc, r1 = min(R, r0 + rpc);
  const size_t y0 = (size_t)((long long)s * y_bs + h);
  const float4* wo = reinterpret_cast<const float4*>(Wo + (size_t)s * R * CW);
  const float4* wd = reinterpret_cast<const float4*>(Wod + (size_t)s * R * CW);
  float a1[2][CW], a2[2][CW];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int c = 0; c < CW; ++c) a1[k][c] = a2[k][c] = 0.f;
  // the row weights of a sub-chunk are staged in shared memory once (coalesced) and read back as broadcast 128-bit loads
  constexpr int SUB = 64;
  __shared__ float4 s_wo[SUB * (CW / 4)], s_wd[SUB * (CW / 4)];
#pragma unroll 1
  for (int rs = r0; rs < r1; rs += SUB) {
    const int nr = min(SUB, r1 - rs);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * (CW / 4); i += CRD_T) {
      s_wo[i] = __ldg(wo + (size_t)rs * (CW / 4) + i);
      s_wd[i] = __ldg(wd + (size_t)rs * (CW / 4) + i);
    }
    __syncthreads();
#pragma unroll 2
    for (int rr = 0; rr < nr; ++rr) {
      const int r = rs + rr;
      float hd[2], hv[2];
      get2<X3>(Yh, Yl, y0 + (size_t)r * 2 * H, hd);
      get2<X3>(Yh, Yl, y0 + (size_t)r * 2 * H + H, hv);
#pragma unroll
      for (int q = 0; q < CW / 4; ++q) {
        const float4 u = s_wo[rr * (CW / 4) + q], v = s_wd[rr * (CW / 4) + q];
        const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            a1[k][4 * q + e] = fmaf(uu[e], hv[k], fmaf(vv[e], hd[k], a1[k][4 * q + e]));
            a2[k][4 * q + e] = fmaf(vv[e], hv[k], a2[k][4 * q + e]);
          }
      }
    }
  }
  float* p1 = part + (((size_t)zc * gridDim.y + s) * 2) * C * H;
#pragma unroll
  for (int c = 0; c < CW; ++c)
    if (c < C && live) {
      *reinterpret_cast<float2*>(p1 + (size_t)c * H + h) = make_float2(a1[0][c], a1[1][c]);
      *reinterpret_cast<float2*>(p1 + (size_t)(C + c) * H + h) = make_float2(a2[0][c], a2[1][c]);
    }
}
// out1[s * P + i] = sum_z part[z][s][0][i], out2[s * P + i] = sum_z part[z][s][1][i]   (i = c * H + h)
__global__ void fnl_colreduce_dual_finish_kernel(const float* __restrict__ part, int S, int H, int C, float* __restrict__ out1,
                                                 float* __restrict__ out2, long long P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (i >= C * H) return;
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int zc = 0; zc < RSPLIT; ++zc) {
    const float* q = part + (((size_t)zc * S + s) * 2) * C * H;
    t1 += q[i];
    t2 += q[(size_t)C * H + i];
  }
  out1[(long long)s * P + i] = t1;
  out2[(long long)s * P + i] = t2;
}

// out[s * P + h] = sum_r YT[s][h][r]   (row sums of a transposed operand array; one warp per row, fixed order)
template <int X3>
__global__ void __launch_bounds__(256)
fnl_rowsum_kernel(const void* Yh, const void* Yl, int H, int Rp, int R, float* __restrict__ out, long long P) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), s = blockIdx.y, lane = threadIdx.x & 31;
  if (row >= H) return;
  const size_t base = ((size_t)s * H + row) * Rp;
  float t = 0.f;
  for (int r4 = lane * 4; r4 < R; r4 += 128) {   // four consecutive r per lane and load (Rp is a multiple of 128)
    float v[4];
    get4<X3>(Yh, Yl, base + r4, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) t += r4 + k < R ? v[k] : 0.f;
  }
  t = warp_sum(t);
  if (lane == 0) out[(long long)s * P + row] = t;
}

// column sums of the row-major split-bf16 adjoint array Y[s][r][0 .. 2H): columns < H -> out1[s * P + c], columns >= H ->
// out2[s * P + c - H]  (the bias adjoints A_b1 / A_b1dot without a transposed copy; a thread owns two adjacent columns)
__global__ void __launch_bounds__(128)
fnl_colsum_pair_kernel(const __nv_bfloat16* __restrict__ Yh, const __nv_bfloat16* __restrict__ Yl, long long bs, int ld, int R, int H,
                       float* __restrict__ out1, float* __restrict__ out2, long long P) {
  const int c = 2 * (blockIdx.x * 128 + threadIdx.x), s = blockIdx.y;
  if (c >= 2 * H) return;
  const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(Yh + (size_t)s * bs + c);
  const __nv_bfloat162* pl = reinterpret_cast<const __nv_bfloat162*>(Yl + (size_t)s * bs + c);
  const size_t step = (size_t)ld / 2;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  int r = 0;
  for (; r + 8 <= R; r += 8) {   // 16 independent 4-byte loads in flight per thread (HBM-bound sweep over 0.5 GB at cfg5)
    __nv_bfloat162 h[8], l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      h[k] = ph[(size_t)(r + k) * step];
      l[k] = pl[(size_t)(r + k) * step];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 hf = __bfloat1622float2(h[k]), lf = __bfloat1622float2(l[k]);
      a0 += hf.x; a1 += hf.y; b0 += lf.x; b1 += lf.y;
    }
  }
  for (; r < R; ++r) {
    const float2 h = __bfloat1622float2(ph[(size_t)r * step]), l = __bfloat1622float2(pl[(size_t)r * step]);
    a0 += h.x; a1 += h.y; b0 += l.x; b1 += l.y;
  }
  float* o = c < H ? out1 + (long long)s * P + c : out2 + (long long)s * P + (c - H);
  o[0] = a0 + b0;
  o[1] = a1 + b1;
}

// G2[s][r][0..64) = [go[s][r][0..16) | god[s][r][0..16) | 0 ...] as split-bf16 pairs, rows r >= R zero: the B operand (N = 64
// columns, K = rows, MN-major) of the W2 adjoint product  hh^T [go | god]
__global__ void fnl_pack_g2_kernel(const float* __restrict__ go, const float* __restrict__ god, int R, int Rp, __nv_bfloat16* __restrict__ Gh,
                                   __nv_bfloat16* __restrict__ Gl) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over S * Rp * 64
  const int c = (int)(i & 63);
  const long long sr = i >> 6;
  const int r = (int)(sr % Rp);
  const long long sidx = sr / Rp;
  float v = 0.f;
  if (r < R && c < 2 * CW) v = (c < CW ? go : god)[((size_t)sidx * R + r) * CW + (c & (CW - 1))];
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  Gh[i] = h;
  Gl[i] = l;
}
// W2 adjoints from Cm[s][2H][32] = hh^T [go | god]  (rows j < H: hdot, rows H + j: h):
// out1[s * P + c * H + j] = (h . go)[j][c] + (hdot . god)[j][c],  out2[...] = (h . god)[j][c]
__global__ void fnl_w2adj_finish_kernel(const float* __restrict__ Cm, int H, int C, float* __restrict__ out1, float* __restrict__ out2,
                                        long long P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (i >= C * H) return;
  const int c = i / H, j = i - c * H;
  const float* q = Cm + (size_t)s * 2 * H * 32;
  out1[(long long)s * P + i] = q[(size_t)(H + j) * 32 + c] + q[(size_t)j * 32 + CW + c];
  out2[(long long)s * P + i] = q[(size_t)(H + j) * 32 + CW + c];
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
// an operand array: the hi part (or the bf16 array) and, in tf32x3 mode, the lo part right behind it
struct Buf {
  uint8_t *hi, *lo;
};
struct Lws {
  Buf X, XT, W1, W1d, W1T2, W2p, W2T, hh, aa, aT, adT, AA, G2;
  float *o, *od, *go, *god, *cpart;
  uint32_t* mbits;     // ReLU mask bits [S][Rp][H / 32]
  float* bpart;        // column sums of abar per 32-row group [S][Rp / 32][H]
  size_t total;
};

void carve_l(int S, int R, int D, int H, int x3, uint8_t* base, Lws& w) {
  const size_t Rp = (size_t)((R + 127) / 128) * 128, es = x3 == 1 ? 4 : 2, cp2 = x3 == 1 ? 64 : 128;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += (bytes + 1023) & ~(size_t)1023; return p; };
  auto takeb = [&](size_t elems) { Buf b; b.hi = take(elems * es); b.lo = x3 ? take(elems * es) : nullptr; return b; };
  w.X = takeb(Rp * D);
  w.XT = takeb((size_t)D * Rp);
  w.W1 = takeb((size_t)S * H * D);
  w.W1d = takeb((size_t)S * H * D);
  w.W1T2 = takeb((size_t)S * D * 2 * H);
  w.W2p = takeb((size_t)S * 128 * 2 * H);
  w.W2T = takeb((size_t)S * H * cp2);
  w.hh = takeb((size_t)S * Rp * 2 * H);
  w.aa = takeb((size_t)S * Rp * 2 * H);
  w.aT = takeb((size_t)S * H * Rp);
  w.adT = takeb((size_t)S * H * Rp);
  w.AA = takeb((size_t)S * Rp * cp2);
  w.G2 = takeb((size_t)S * Rp * 64);      // [A_o | A_od | 0] as a split-bf16 operand (the W2 adjoint product of the dual pass)
  w.o = (float*)take((size_t)S * Rp * CW * 4);
  w.od = (float*)take((size_t)S * Rp * CW * 4);
  w.go = (float*)take((size_t)S * R * CW * 4);
  w.god = (float*)take((size_t)S * R * CW * 4);
  w.cpart = (float*)take((size_t)RSPLIT * S * 2 * CW * H * 4);
  w.mbits = (uint32_t*)take((size_t)S * Rp * (H / 32) * 4);
  w.bpart = (float*)take((size_t)S * (Rp / 32) * H * 4);
  w.total = off;
}

struct Operand {
  Buf buf;
  uint64_t eoff;               // element offset of the view inside the buffer
  uint64_t inner, outer, ld;   // K extent, rows, leading dimension (elements)
  int brows;                   // rows per batch (0: shared)
  int mn;                      // MN-major view (split-bf16): inner = M (or N) extent, outer / brows count K rows
};

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_map_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                 uint32_t box_outer) {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    PSVI_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr));
    PSVI_REQUIRE(ptr != nullptr && qr == cudaDriverEntryPointSuccess, PSVI_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
    fn = reinterpret_cast<EncodeFn>(ptr);
  }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld * 4};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PSVI_REQUIRE(r == CUDA_SUCCESS, PSVI_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return PSVI_OK;
}

// [outer rows][inner bf16 columns] output view, boxes of 32 x 32 with the 64-byte swizzle (the TMA-stored epilogue)
int make_map_out_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qr;
  PSVI_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr));
  PSVI_REQUIRE(ptr != nullptr && qr == cudaDriverEntryPointSuccess, PSVI_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = reinterpret_cast<EncodeFn>(ptr)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                                                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PSVI_REQUIRE(r == CUDA_SUCCESS, PSVI_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)r);
  return PSVI_OK;
}

template <int X3>
int launch_gemm(const Operand& A, const Operand& B, GemmP p, int sms, cudaStream_t st) {
  using PR = Prec<X3>;
  CUtensorMap ma, mal, mb, mbl, mo, mol;
  int rc;
  PSVI_REQUIRE(X3 == 2 || (!A.mn && !B.mn), PSVI_ERR_INVALID, "MN-major operands are built for the split-bf16 kernels");
  if (X3 == 2) {
    // K-major: boxes of [GM rows x GK K-elements]; MN-major: boxes of [GK K-rows x 64 MN-elements] (two per tile)
    if ((rc = make_map_2d_bf16_ld(&ma, A.buf.hi + A.eoff * 2, A.inner, A.outer, A.ld, A.mn ? 64 : PR::GK, A.mn ? PR::GK : GM))) return rc;
    if ((rc = make_map_2d_bf16_ld(&mal, A.buf.lo + A.eoff * 2, A.inner, A.outer, A.ld, A.mn ? 64 : PR::GK, A.mn ? PR::GK : GM))) return rc;
    if ((rc = make_map_2d_bf16_ld(&mb, B.buf.hi + B.eoff * 2, B.inner, B.outer, B.ld, B.mn ? 64 : PR::GK, B.mn ? PR::GK : GN))) return rc;
    if ((rc = make_map_2d_bf16_ld(&mbl, B.buf.lo + B.eoff * 2, B.inner, B.outer, B.ld, B.mn ? 64 : PR::GK, B.mn ? PR::GK : GN))) return rc;
  } else if (X3 == 1) {
    if ((rc = make_map_f32(&ma, A.buf.hi + A.eoff * 4, A.inner, A.outer, A.ld, PR::GK, GM))) return rc;
    if ((rc = make_map_f32(&mal, A.buf.lo + A.eoff * 4, A.inner, A.outer, A.ld, PR::GK, GM))) return rc;
    if ((rc = make_map_f32(&mb, B.buf.hi + B.eoff * 4, B.inner, B.outer, B.ld, PR::GK, GN))) return rc;
    if ((rc = make_map_f32(&mbl, B.buf.lo + B.eoff * 4, B.inner, B.outer, B.ld, PR::GK, GN))) return rc;
  } else {
    if ((rc = make_map_2d_bf16_ld(&ma, A.buf.hi + A.eoff * 2, A.inner, A.outer, A.ld, PR::GK, GM))) return rc;
    if ((rc = make_map_2d_bf16_ld(&mb, B.buf.hi + B.eoff * 2, B.inner, B.outer, B.ld, PR::GK, GN))) return rc;
    mal = ma;
    mbl = mb;
  }
  p.kc = (int)((A.mn ? (A.brows ? (uint64_t)A.brows : A.outer) : A.inner) / PR::GK);
  p.a_brows = A.brows;
  p.b_brows = B.brows;
  p.a_mn = A.mn;
  p.b_mn = B.mn;
  p.m_tiles = (p.M_valid + GM - 1) / GM;
  p.n_tiles = ((p.n_pad > p.N_valid ? p.n_pad : p.N_valid) + GN - 1) / GN;
  if ((p.ob || p.obt) && p.ob_rows > p.M_valid) p.m_tiles = (p.ob_rows + GM - 1) / GM;
  const int total = p.batch * p.m_tiles * p.n_tiles;
  const int grid = total < sms ? total : sms;
  size_t smem = (size_t)PR::GST * PR::TILES * G_TILE_BYTES + (2 * PR::GST + 4) * 8 + 16 + 1024;
  // fp32 outputs are transposed through 8 x 4 KB of staging behind the barriers (the same area the TMA-stored epilogue uses)
  // (one user of the per-warp staging area per launch: not together with a TMA-stored row-major operand output)
  p.obt_stage = X3 == 2 && p.obt && p.obt_lo && !p.ob && p.ob_rows % 32 == 0 && p.obt_ld % 8 == 0 && p.obt_bs % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(p.obt) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.obt_lo) & 15) == 0;
  if (p.colpart) p.obt_stage = 0;
  if (p.of || p.mbits_out || p.obt_stage || p.colpart) smem = (size_t)PR::GST * PR::TILES * G_TILE_BYTES + 1024 + 8 * STG_WARP_BYTES + 1024;
  mo = ma;
  mol = ma;
  p.ob_tma = 0;
  if (X3 == 2 && !p.of && p.ob && p.ob_lo && p.ob_rows % GM == 0 && p.ob_bs == (long long)p.ob_rows * p.ob_ld && !getenv("PSVI_FNL_NO_TMASTORE")) {
    const uint64_t ncols = (uint64_t)(p.n_pad > p.N_valid ? p.n_pad : p.N_valid);
    if ((rc = make_map_out_bf16(&mo, p.ob, ncols, (uint64_t)p.batch * p.ob_rows, (uint64_t)p.ob_ld))) return rc;
    if ((rc = make_map_out_bf16(&mol, p.ob_lo, ncols, (uint64_t)p.batch * p.ob_rows, (uint64_t)p.ob_ld))) return rc;
    p.ob_tma = 1;
    smem = (size_t)PR::GST * PR::TILES * G_TILE_BYTES + 1024 + 8 * STG_WARP_BYTES + 1024;
  }
  // CTA pairs (cta_group::2, 256 x 256 tiles: half the operand traffic per output) whenever the tile grid is even both ways
  const char* cg2_env = getenv("PSVI_FNL_CG2");   // (read per call: the parity tests switch it inside one process)
  const int cg2_mode = cg2_env ? atoi(cg2_env) : 1;
  const bool pair = X3 != 0 && cg2_mode && p.m_tiles % 2 == 0 && p.n_tiles % 2 == 0 && sms >= 2 &&
                    (cg2_mode == 2 || p.kc * PR::GK >= 256);
  if (pair) {
    const int items = total / 4;
    const int grid2 = 2 * (items < sms / 2 ? items : sms / 2);
    // function attributes are per device: set on every call (cheap), not cached per process
    PSVI_CUDA_CHECK(cudaFuncSetAttribute(tn_gemm_kernel<X3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid2);
    cfg.blockDim = dim3(G_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    PSVI_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tn_gemm_kernel<X3, 2>, ma, mal, mb, mbl, mo, mol, p));
    return PSVI_OK;
  }
  PSVI_CUDA_CHECK(cudaFuncSetAttribute(tn_gemm_kernel<X3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tn_gemm_kernel<X3, 1><<<grid, G_THREADS, smem, st>>>(ma, mal, mb, mbl, mo, mol, p);
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

template <int X3>
int fnl_pass_impl(const psvi_mf_model* model, const float* theta, const float* thetad, const float* x, const int32_t* y,
                  const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar, float* logits,
                  void* workspace, cudaStream_t st) {
  constexpr int CP = Prec<X3>::GK;      // classes padded to one K block (64 bf16 / 32 fp32)
  constexpr size_t ES = Prec<X3>::ES;
  const int D = model->dims[0], H = model->dims[1], C = model->dims[2], S = model->mc_samples;
  const long long P = (long long)H * D + H + (long long)C * H + C;
  const long long o_b1 = (long long)H * D, o_w2 = o_b1 + H, o_b2 = o_w2 + (long long)C * H;
  const int Rp = ((R + 127) / 128) * 128;
  int dev = 0, sms = 0, rc;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  Lws w;
  carve_l(S, R, D, H, X3, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023), w);
  const bool dual = thetad != nullptr;
  // ---- operands
  fnl_prep_x_kernel<X3><<<dim3((Rp + 31) / 32, (D + 31) / 32), dim3(32, 8), 0, st>>>(x, R, Rp, D, w.X.hi, w.X.lo, w.XT.hi, w.XT.lo);
  for (int part = 0; part < (X3 ? 2 : 1); ++part) {
    PSVI_CUDA_CHECK(cudaMemsetAsync(part ? w.W2p.lo : w.W2p.hi, 0, (size_t)S * 128 * 2 * H * ES, st));
    PSVI_CUDA_CHECK(cudaMemsetAsync(part ? w.W2T.lo : w.W2T.hi, 0, (size_t)S * H * 2 * CP * ES, st));
  }
  fnl_pack_w1_kernel<X3><<<dim3(D / 64, H / 64, S), dim3(32, 8), 0, st>>>(theta, P, D, H, w.W1.hi, w.W1.lo, w.W1T2.hi, w.W1T2.lo, 0);
  fnl_pack_w2_kernel<X3><<<dim3(8, S), 256, 0, st>>>(theta, P, D, H, C, w.W2p.hi, w.W2p.lo, 0, w.W2T.hi, w.W2T.lo, CP, 2 * CP);
  if (dual) {
    fnl_pack_w1_kernel<X3><<<dim3(D / 64, H / 64, S), dim3(32, 8), 0, st>>>(thetad, P, D, H, w.W1d.hi, w.W1d.lo, w.W1T2.hi,
                                                                         w.W1T2.lo, H);
    fnl_pack_w2_kernel<X3><<<dim3(8, S), 256, 0, st>>>(thetad, P, D, H, C, w.W2p.hi, w.W2p.lo, H, w.W2T.hi, w.W2T.lo, 0, 2 * CP);
  }
  PSVI_CUDA_CHECK(cudaGetLastError());
  const uint64_t H2 = 2 * (uint64_t)H, uH = (uint64_t)H, uD = (uint64_t)D, uRp = (uint64_t)Rp, uS = (uint64_t)S;
  const Operand opX{w.X, 0, uD, uRp, uD, 0};
  const Operand opXT{w.XT, 0, uRp, uD, uRp, 0};
  const Operand opW1{w.W1, 0, uD, uS * H, uD, H};
  const Operand opW1d{w.W1d, 0, uD, uS * H, uD, H};
  const Operand opH{w.hh, uH, uH, uS * Rp, H2, Rp};                   // h
  const Operand opHH{w.hh, 0, H2, uS * Rp, H2, Rp};                   // [hdot | h]
  const Operand opW2{w.W2p, 0, uH, uS * 128, H2, 128};                // W2 (rows >= C zero)
  const Operand opW22{w.W2p, 0, H2, uS * 128, H2, 128};               // [W2 | W2dot]
  const Operand opA0{w.AA, 0, (uint64_t)CP, uS * Rp, 2 * (uint64_t)CP, Rp};          // first K block of AA
  const Operand opAA{w.AA, 0, 2 * (uint64_t)CP, uS * Rp, 2 * (uint64_t)CP, Rp};      // [A_od | A_o]
  const Operand opW2T1{w.W2T, (uint64_t)CP, (uint64_t)CP, uS * H, 2 * (uint64_t)CP, H};   // W2^T (classes padded to one K block)
  const Operand opW2TT{w.W2T, 0, 2 * (uint64_t)CP, uS * H, 2 * (uint64_t)CP, H};          // [W2dot^T | W2^T]
  const Operand opA{w.aa, 0, uH, uS * Rp, H2, Rp};                    // abar / A_a
  const Operand opAAh{w.aa, 0, H2, uS * Rp, H2, Rp};                  // [A_a | A_adot]
  const Operand opW1T{w.W1T2, 0, uH, uS * D, H2, D};                  // W1^T
  const Operand opW1TT{w.W1T2, 0, H2, uS * D, H2, D};                 // [W1^T | W1dot^T]
  const Operand opAT{w.aT, 0, uRp, uS * H, uRp, H};
  const Operand opADT{w.adT, 0, uRp, uS * H, uRp, H};
  GemmP z;
  memset(&z, 0, sizeof(z));
  z.batch = S;
  const long long hh_bs = (long long)Rp * 2 * H;
  auto at = [&](const Buf& b, size_t eoff, bool lo) -> void* { return (lo ? b.lo : b.hi) ? (lo ? b.lo : b.hi) + eoff * ES : nullptr; };
  // the ReLU mask travels as bits (written by the forward product below); PSVI_FNL_NO_MBITS=1 reads the stored h instead
  const bool use_mbits = !getenv("PSVI_FNL_NO_MBITS");
  auto set_mask_h = [&](GemmP& p) {
    if (use_mbits) { p.mbits = w.mbits; p.mbits_bs = (long long)Rp * (H / 32); p.mbits_ld = H / 32; }
    else { p.mask = at(w.hh, H, false); p.mask_bs = hh_bs; p.mask_ld = 2 * H; }
  };
  auto set_ob = [&](GemmP& p, const Buf& b, size_t eoff) {
    p.ob = at(b, eoff, false); p.ob_lo = at(b, eoff, true); p.ob_bs = hh_bs; p.ob_ld = 2 * H; p.ob_rows = Rp;
  };
  auto set_obt = [&](GemmP& p, const Buf& b) {
    p.obt = b.hi; p.obt_lo = b.lo; p.obt_bs = (long long)H * Rp; p.obt_ld = Rp;
  };
  // ---- primal forward: h = relu(X W1^T + b1);  o = h W2^T
  {
    // roles swapped (C^T = W1 X^T, output through the transposed-store path): the row-major store of h is then coalesced
    // (lanes = consecutive hidden units) and the bias is one scalar per thread
    GemmP p = z;
    p.M_valid = H; p.N_valid = R; p.n_pad = Rp; p.bias = theta + o_b1; p.bias_bs = P; p.bias_row = 1; p.relu = 1;
    p.obt = at(w.hh, H, false); p.obt_lo = at(w.hh, H, true); p.obt_bs = hh_bs; p.obt_ld = 2 * H; p.ob_rows = H;
    if (use_mbits && tbar) { p.mbits_out = w.mbits; p.mbits_bs = (long long)Rp * (H / 32); p.mbits_ld = H / 32; }
    if ((rc = launch_gemm<X3>(opW1, opX, p, sms, st))) return rc;
    p = z;
    p.M_valid = R; p.N_valid = CW; p.of = w.o; p.of_bs = (long long)Rp * CW; p.of_ld = CW;
    if ((rc = launch_gemm<X3>(opH, opW2, p, sms, st))) return rc;
  }
  const int hb = (S * Rp + 127) / 128;
  if (!tbar) {
    fnl_head_kernel<X3><<<hb, 128, 0, st>>>(w.o, nullptr, theta + o_b2, nullptr, P, y, nullptr, S, R, Rp, C, 0, nll, logits, nullptr,
                                           nullptr, nullptr, nullptr, nullptr);
    PSVI_CUDA_CHECK(cudaGetLastError());
    return PSVI_OK;
  }
  for (int part = 0; part < (X3 ? 2 : 1); ++part)   // K blocks of the output adjoints: zero except the 16-entry heads
    PSVI_CUDA_CHECK(cudaMemsetAsync(part ? w.AA.lo : w.AA.hi, 0, (size_t)S * Rp * 2 * CP * ES, st));
  auto colreduce = [&](const Buf& Y, long long y_off, const float* W, int Cc, float* out, int accumulate) {
    fnl_colreduce_kernel<X3><<<dim3((H + 4 * CR_T - 1) / (4 * CR_T), S, RSPLIT), CR_T, 0, st>>>(Y.hi, Y.lo, y_off, hh_bs, 2 * H, W, R, H, Cc, w.cpart);
    fnl_colreduce_finish_kernel<<<dim3((Cc * H + 255) / 256, S), 256, 0, st>>>(w.cpart, S, H, Cc, out, P, accumulate);
  };
  if (!dual) {
    // ---- gradient pass
    fnl_head_kernel<X3><<<hb, 128, 0, st>>>(w.o, nullptr, theta + o_b2, nullptr, P, y, cw, S, R, Rp, C, 1, nll, logits, w.go, nullptr,
                                           w.AA.hi, w.AA.lo, nullptr);
    GemmP p = z;      // abar = (obar W2) * (h > 0): transposed (K-major for the next GEMM); row-major only if xbar needs it
    p.M_valid = R; p.N_valid = H;
    set_mask_h(p); set_obt(p, w.aT);
    p.ob_rows = Rp;
    if (xbar) set_ob(p, w.aa, 0);
    // tf32x3: the bias adjoint b1bar = column sums of abar comes out of this product's epilogue (per 32-row group, then a
    // finish kernel) instead of a row-sum sweep over the 0.5 GB transposed array (PSVI_FNL_NO_COLPART=1: the sweep)
    const bool colpart = X3 == 1 && !getenv("PSVI_FNL_NO_COLPART");
    if (colpart) { p.colpart = w.bpart; p.colpart_bs = (long long)(Rp / 32) * H; p.colpart_ld = H; }
    if ((rc = launch_gemm<X3>(opA0, opW2T1, p, sms, st))) return rc;
    p = z;            // W1bar = abar^T X
    p.M_valid = H; p.N_valid = D; p.of = tbar; p.of_bs = P; p.of_ld = D;
    if ((rc = launch_gemm<X3>(opAT, opXT, p, sms, st))) return rc;
    if (colpart) fnl_colpart_finish_kernel<<<dim3((H + 255) / 256, S), 256, 0, st>>>(w.bpart, Rp / 32, H, tbar + o_b1, P);
    else fnl_rowsum_kernel<X3><<<dim3((H + 7) / 8, S), 256, 0, st>>>(w.aT.hi, w.aT.lo, H, Rp, R, tbar + o_b1, P);   // b1bar
    colreduce(w.hh, H, w.go, C, tbar + o_w2, 0);
    fnl_colsum16_kernel<<<S, 256, 0, st>>>(w.go, R, C, tbar + o_b2, P);
    if (xbar) {
      p = z;          // xbar = abar W1
      p.M_valid = R; p.N_valid = D; p.of = xbar; p.of_bs = (long long)R * D; p.of_ld = D;
      if ((rc = launch_gemm<X3>(opA, opW1T, p, sms, st))) return rc;
    }
    PSVI_CUDA_CHECK(cudaGetLastError());
    return PSVI_OK;
  }
  // ---- dual pass: tangent forward
  {
    GemmP p = z;      // hdot = (X W1dot^T + b1dot) * (h > 0)   (not role-swapped: the transposed mask read costs more than
    p.M_valid = R; p.N_valid = H; p.bias = thetad + o_b1; p.bias_bs = P;     //  the coalesced store saves -- measured in
                                                                            //  tf32x3 (round 1) and in split-bf16: the dual
                                                                            //  passes of a cfg5 step take 29.4 ms swapped
                                                                            //  against 21.8 ms as written)
    set_mask_h(p); set_ob(p, w.hh, 0);
    if ((rc = launch_gemm<X3>(opX, opW1d, p, sms, st))) return rc;
    p = z;            // odot = hdot W2^T + h W2dot^T
    p.M_valid = R; p.N_valid = CW; p.of = w.od; p.of_bs = (long long)Rp * CW; p.of_ld = CW;
    if ((rc = launch_gemm<X3>(opHH, opW22, p, sms, st))) return rc;
  }
  fnl_head_kernel<X3><<<hb, 128, 0, st>>>(w.o, w.od, theta + o_b2, thetad + o_b2, P, y, cw, S, R, Rp, C, 2, nll, logits, w.go, w.god,
                                         w.AA.hi, w.AA.lo, acbar);
  {
    // split-bf16: the products with A_a^T / A_adot^T read the ROW-MAJOR adjoints as MN-major UMMA operands (and X likewise),
    // so no transposed copies are written (PSVI_FNL_NO_MN=1 restores the transposed-copy form for comparison)
    const bool mn = X3 == 2 && !getenv("PSVI_FNL_NO_MN");
    GemmP p = z;      // A_a = (A_od W2dot + A_o W2) * (h > 0)
    p.M_valid = R; p.N_valid = H;
    set_mask_h(p); set_ob(p, w.aa, 0);
    if (!mn) set_obt(p, w.aT);
    if ((rc = launch_gemm<X3>(opAA, opW2TT, p, sms, st))) return rc;
    set_ob(p, w.aa, H);                         // A_adot = (A_od W2) * (h > 0)
    if (!mn) set_obt(p, w.adT);
    if ((rc = launch_gemm<X3>(opA0, opW2T1, p, sms, st))) return rc;
    p = z;            // A_W1 = A_a^T X;  A_W1dot = A_adot^T X
    p.M_valid = H; p.N_valid = D; p.of = tbar; p.of_bs = P; p.of_ld = D;
    if (mn) {
      const Operand opAmn{w.aa, 0, uH, uS * Rp, H2, Rp, 1}, opAdmn{w.aa, uH, uH, uS * Rp, H2, Rp, 1};   // M = hidden, K = rows
      const Operand opXmn{w.X, 0, uD, uRp, uD, 0, 1};                                                     // N = D, K = rows
      if ((rc = launch_gemm<X3>(opAmn, opXmn, p, sms, st))) return rc;
      p.of = tdbar;
      if ((rc = launch_gemm<X3>(opAdmn, opXmn, p, sms, st))) return rc;
    } else {
      if ((rc = launch_gemm<X3>(opAT, opXT, p, sms, st))) return rc;
      p.of = tdbar;
      if ((rc = launch_gemm<X3>(opADT, opXT, p, sms, st))) return rc;
    }
    if (xbar) {
      p = z;          // A_x = A_a W1 + A_adot W1dot
      p.M_valid = R; p.N_valid = D; p.of = xbar; p.of_bs = (long long)R * D; p.of_ld = D;
      if ((rc = launch_gemm<X3>(opAAh, opW1TT, p, sms, st))) return rc;
    }
  }
  // A_b1 / A_b1dot: row sums of the transposed adjoints (coalesced along r);  A_W2 = A_o^T h + A_od^T hdot and
  // A_W2dot = A_od^T h in one sweep over hh
  if (X3 == 2 && !getenv("PSVI_FNL_NO_MN")) {
    fnl_colsum_pair_kernel<<<dim3((H + 127) / 128, S), 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(w.aa.hi),
                                                                      reinterpret_cast<const __nv_bfloat16*>(w.aa.lo), hh_bs, 2 * H, R, H,
                                                                      tbar + o_b1, tdbar + o_b1, P);
  } else {
    fnl_rowsum_kernel<X3><<<dim3((H + 7) / 8, S), 256, 0, st>>>(w.aT.hi, w.aT.lo, H, Rp, R, tbar + o_b1, P);
    fnl_rowsum_kernel<X3><<<dim3((H + 7) / 8, S), 256, 0, st>>>(w.adT.hi, w.adT.lo, H, Rp, R, tdbar + o_b1, P);
  }
  if (X3 == 2 && !getenv("PSVI_FNL_NO_MN") && !getenv("PSVI_FNL_NO_W2GEMM")) {
    // A_W2 = A_o^T h + A_od^T hdot,  A_W2dot = A_od^T h  as ONE tensor-core product  hh^T [A_o | A_od]  (M = 2H hidden columns of
    // the row-major hh read as an MN-major operand, N = 64, K = rows) instead of a CUDA-core sweep over hh
    const long long n2 = (long long)S * Rp * 64;
    fnl_pack_g2_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(w.go, w.god, R, Rp, reinterpret_cast<__nv_bfloat16*>(w.G2.hi),
                                                                      reinterpret_cast<__nv_bfloat16*>(w.G2.lo));
    const Operand opHmn{w.hh, 0, H2, uS * Rp, H2, Rp, 1}, opGmn{w.G2, 0, 64, uS * Rp, 64, Rp, 1};
    GemmP p = z;
    p.M_valid = 2 * H; p.N_valid = 32; p.of = w.cpart; p.of_bs = (long long)2 * H * 32; p.of_ld = 32;
    if ((rc = launch_gemm<X3>(opHmn, opGmn, p, sms, st))) return rc;
    fnl_w2adj_finish_kernel<<<dim3((C * H + 255) / 256, S), 256, 0, st>>>(w.cpart, H, C, tbar + o_w2, tdbar + o_w2, P);
  } else {
    fnl_colreduce_dual_kernel<X3><<<dim3((H + 2 * CRD_T - 1) / (2 * CRD_T), S, RSPLIT), CRD_T, 0, st>>>(w.hh.hi, w.hh.lo, hh_bs, w.go, w.god, R, H, C, w.cpart);
    fnl_colreduce_dual_finish_kernel<<<dim3((C * H + 255) / 256, S), 256, 0, st>>>(w.cpart, S, H, C, tbar + o_w2, tdbar + o_w2, P);
  }
  fnl_colsum16_kernel<<<S, 256, 0, st>>>(w.go, R, C, tbar + o_b2, P);
  fnl_colsum16_kernel<<<S, 256, 0, st>>>(w.god, R, C, tdbar + o_b2, P);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // namespace

extern "C" {

size_t psvi_fnl_workspace_bytes(const psvi_mf_model* model, int32_t R, int32_t precision) {
  if (!model || model->n_layers != 2 || R <= 0) return 0;
  Lws w;
  carve_l(model->mc_samples, R, model->dims[0], model->dims[1], precision, nullptr, w);
  return w.total + 1024;
}

int psvi_fnl_pass(const psvi_mf_model* model, int32_t precision, const float* theta, const float* thetad, const float* x,
                  const int32_t* y, const float* cw, int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar,
                  float* logits, void* workspace, void* stream_) {
  PSVI_REQUIRE(theta && x && y && workspace, PSVI_ERR_INVALID, "null pointer");
  int rc = check_large(model);
  if (rc) return rc;
  PSVI_REQUIRE(R >= 1, PSVI_ERR_INVALID, "bad R");
  PSVI_REQUIRE(precision >= 0 && precision <= 2, PSVI_ERR_INVALID, "precision must be 0 (bf16), 1 (tf32x3) or 2 (bf16x3)");
  PSVI_REQUIRE(!thetad || (tbar && tdbar), PSVI_ERR_INVALID, "the dual pass needs tbar and tdbar");
  cudaStream_t st = (cudaStream_t)stream_;
  if (precision == 2)
    return fnl_pass_impl<2>(model, theta, thetad, x, y, cw, R, nll, tbar, tdbar, xbar, acbar, logits, workspace, st);
  if (precision == 1)
    return fnl_pass_impl<1>(model, theta, thetad, x, y, cw, R, nll, tbar, tdbar, xbar, acbar, logits, workspace, st);
  return fnl_pass_impl<0>(model, theta, thetad, x, y, cw, R, nll, tbar, tdbar, xbar, acbar, logits, workspace, st);
}

}  // extern "C"
