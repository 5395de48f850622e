// psvi_lenet.cu -- the convolutional family of the PSVI hot path: the per-sample network pass
// of `lenet` (reference psvi/models/neural_net.py:194-246 VIConv2d, :249-255 BatchMaxPool2d, :334-359 make_lenet) on
// externally supplied sampled weights theta [S][P], in the three flavours the streaming PSVI engine needs
// (psvi/inference/stream.py; same contract as psvi_net_pass):
//   forward          -> nll [S][R] (and logits)
//   gradient pass    -> tbar = d/dtheta_s, xbar = d/dx per sample, with per-sample row weights cw [S][R]
//   dual (HVP) pass  -> tbar = A_theta, tdbar = A_thetadot, xbar = A_x, acbar = adjoint of the row weights  (SURVEY A.6,
//                       generalised: a conv layer is a linear map, ReLU + max-pool a fixed selection)
// Network: conv(1->6, 5x5, pad 2) ReLU pool2 | conv(6->16, 5x5) ReLU pool2 | flatten 400 | fc 120 ReLU | fc 84 ReLU | fc 10.
// theta layout (TL): per layer weight then bias: W1[6][1][5][5] b1[6] W2[16][6][5][5] b2[16] W3[120][400] b3[120]
// W4[84][120] b4[84] W5[10][84] b5[10]  (P = 61 706).
// fp32 on the CUDA cores: the products have K = 25 / 150 and N = 6 / 16 -- far below a 128-wide UMMA tile -- and the
// reference computes this family in fp32 (and bf16 operands would destroy the unrolled hypergradient).  A CTA
// keeps the feature maps of a few rows of one sample in shared memory: conv + ReLU + pool are fused (the full-resolution maps
// never reach global memory; only the pooled map and a 3-bit selection code per pooled element do), the backward kernels
// scatter through that code on the fly, and every conv kernel is register-tiled so that the FMA pipe, not shared memory, is
// the limit (patch reuse across channels, 64/128-bit shared loads); the fully connected layers go through one batched
// small-GEMM kernel.
#include <algorithm>
#include "psvi_common.cuh"

namespace {

constexpr int LN_P = 61706;
constexpr int O_W1 = 0, O_B1 = 150, O_W2 = 156, O_B2 = 2556, O_W3 = 2572, O_B3 = 50572, O_W4 = 50692, O_B4 = 60772,
              O_W5 = 60856, O_B5 = 61696;
constexpr int N_P1 = 6 * 14 * 14, N_P2 = 16 * 5 * 5, N_H3 = 120, N_H4 = 84, N_O = 10, N_X = 28 * 28;
constexpr int N_CHUNKS = 128;   // row chunks of the conv weight-gradient kernels (up to 128 x S CTAs)

// ------------------------------------------------------------------------------------------------ conv + ReLU + pool
// out[s][r][co][py][px]: primal: max over the 2x2 window of relu(conv(in1, w1) [+ conv(in2, w2)] + bias), selection code
// sel = argmax (first maximum in row-major window order, as torch's max_pool2d) | (max > 0) << 2;
// tangent: the same selection applied to conv(in1, w1) + conv(in2, w2) + bias (all of them tangent quantities).
// One CTA per (IMGS rows, sample); a thread owns one pooling window of COG output channels: the 6x6 input patch of an input
// channel is read once (18 64-bit loads) and reused for COG x 100 FMAs, the 25 (padded to 28) weights of a channel pair come in
// as 7 warp-broadcast 128-bit loads, and the pooling is thread-local, so the full-resolution map is never materialised.
constexpr int WPAD = 28;   // taps of one (co, ci) filter in shared memory, padded for 128-bit loads
template <int CI, int CO, int HIN, int PAD, int IMGS, int COG>
struct ConvFwdCfg {
  static constexpr int HP = HIN + 2 * PAD, HO = HP - 4, HQ = HO / 2, NIN = CI * HP * HP, NWP = CO * CI * WPAD;
  static constexpr int TPI = (CO / COG) * HQ * HQ, NT = (IMGS * TPI + 31) / 32 * 32;
  static constexpr size_t SMEM = (size_t)(2 * IMGS * NIN + 2 * NWP) * sizeof(float);
};
template <int CI, int CO, int HIN, int PAD, int IMGS, int COG>
__global__ void __launch_bounds__((ConvFwdCfg<CI, CO, HIN, PAD, IMGS, COG>::NT))
conv_pool_fwd_kernel(const float* __restrict__ in1, size_t ss1, const float* __restrict__ w1, const float* __restrict__ in2,
                     size_t ss2, const float* __restrict__ w2, const float* __restrict__ bias, int R, uint8_t* sel,
                     int tangent, float* __restrict__ out) {
  using C = ConvFwdCfg<CI, CO, HIN, PAD, IMGS, COG>;
  constexpr int HP = C::HP, HQ = C::HQ, NIN = C::NIN, NWP = C::NWP, TPI = C::TPI, NT = C::NT;
  static_assert(HP % 2 == 0 && NIN % 4 == 0 && CO % COG == 0, "alignment of the vector loads");
  extern __shared__ __align__(16) float smem_f[];
  float* s_in1 = smem_f;               // [IMGS][CI][HP][HP]
  float* s_in2 = s_in1 + IMGS * NIN;
  float* s_w1 = s_in2 + IMGS * NIN;    // [CO][CI][WPAD]
  float* s_w2 = s_w1 + NWP;
  const int s = blockIdx.y, tid = threadIdx.x, r_base = blockIdx.x * IMGS;
  // staging: 64-bit copies of the interior (HIN, PAD, HP even), zero border only where there is one (PAD > 0), and nothing at
  // all for the second term when it is absent (primal passes: in2 == w2 == nullptr -- s_in2 / s_w2 are never read then)
  const bool two = in2 != nullptr;
  if (PAD > 0) {
    for (int i = tid; i < IMGS * NIN / 4; i += NT) {
      reinterpret_cast<float4*>(s_in1)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (two) reinterpret_cast<float4*>(s_in2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
  }
  {
    constexpr int HH = HIN / 2, PER_IMG = CI * HIN * HH;
    for (int i = tid; i < IMGS * PER_IMG; i += NT) {
      const int img = i / PER_IMG, j = i % PER_IMG, r = r_base + img;
      const int ci = j / (HIN * HH), yy = (j / HH) % HIN, x2 = j % HH;
      const size_t off = (size_t)r * CI * HIN * HIN + (ci * HIN + yy) * HIN + 2 * x2;
      const int so = img * NIN + (ci * HP + yy + PAD) * HP + 2 * x2 + PAD;
      const bool ok = r < R;
      *reinterpret_cast<float2*>(s_in1 + so) = ok ? *reinterpret_cast<const float2*>(in1 + (size_t)s * ss1 + off) : make_float2(0.f, 0.f);
      if (two) *reinterpret_cast<float2*>(s_in2 + so) = ok ? *reinterpret_cast<const float2*>(in2 + (size_t)s * ss2 + off) : make_float2(0.f, 0.f);
    }
  }
  for (int i = tid; i < NWP; i += NT) {
    const int f = i / WPAD, t = i % WPAD;
    s_w1[i] = t < 25 ? w1[(size_t)s * LN_P + f * 25 + t] : 0.f;
    if (w2) s_w2[i] = t < 25 ? w2[(size_t)s * LN_P + f * 25 + t] : 0.f;
  }
  __syncthreads();
  const int img = tid / TPI, rem = tid % TPI, r = r_base + img;
  if (img >= IMGS || r >= R) return;
  const int g = rem / (HQ * HQ), py = (rem / HQ) % HQ, px = rem % HQ;
  float acc[COG][4];
#pragma unroll
  for (int c = 0; c < COG; ++c) {
    const float b0 = bias ? bias[(size_t)s * LN_P + g * COG + c] : 0.f;
    acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = b0;
  }
#pragma unroll 1
  for (int term = 0; term < 2; ++term) {
    if (term == 1 && !in2) break;
    const float* sin = (term ? s_in2 : s_in1) + img * NIN;
    const float* sw = term ? s_w2 : s_w1;
#pragma unroll 1
    for (int ci = 0; ci < CI; ++ci) {
      const float* a = sin + (ci * HP + 2 * py) * HP + 2 * px;
      float pt[6][6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float2 v = *reinterpret_cast<const float2*>(a + i * HP + 2 * j);
          pt[i][2 * j] = v.x;
          pt[i][2 * j + 1] = v.y;
        }
#pragma unroll
      for (int c = 0; c < COG; ++c) {
        const float4* w4 = reinterpret_cast<const float4*>(sw + ((g * COG + c) * CI + ci) * WPAD);
        float wv[WPAD];
#pragma unroll
        for (int i = 0; i < WPAD / 4; ++i) {
          const float4 v = w4[i];
          wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const float wk = wv[ky * 5 + kx];
            acc[c][0] = fmaf(pt[ky][kx], wk, acc[c][0]);
            acc[c][1] = fmaf(pt[ky][kx + 1], wk, acc[c][1]);
            acc[c][2] = fmaf(pt[ky + 1][kx], wk, acc[c][2]);
            acc[c][3] = fmaf(pt[ky + 1][kx + 1], wk, acc[c][3]);
          }
      }
    }
  }
  const size_t ob = ((size_t)s * R + r) * (CO * HQ * HQ);
#pragma unroll
  for (int c = 0; c < COG; ++c) {
    const size_t o = ob + ((g * COG + c) * HQ + py) * HQ + px;
    if (!tangent) {
      float best = fmaxf(acc[c][0], 0.f);
      int k = 0;
      if (fmaxf(acc[c][1], 0.f) > best) { best = fmaxf(acc[c][1], 0.f); k = 1; }
      if (fmaxf(acc[c][2], 0.f) > best) { best = fmaxf(acc[c][2], 0.f); k = 2; }
      if (fmaxf(acc[c][3], 0.f) > best) { best = fmaxf(acc[c][3], 0.f); k = 3; }
      sel[o] = (uint8_t)(k | ((best > 0.f) ? 4 : 0));
      out[o] = best;
    } else {
      const int cd = sel[o], k = cd & 3;
      const float v = k == 0 ? acc[c][0] : (k == 1 ? acc[c][1] : (k == 2 ? acc[c][2] : acc[c][3]));
      out[o] = (cd & 4) ? v : 0.f;
    }
  }
}

// d/d(input map) of the fused conv + ReLU + pool: the pooled adjoints pb1 (and pb2) are scattered to full resolution through
// the selection code into a zero-padded map, then out[ci][y][x] = sum_co sum_k a1[co][y+PAD-ky][x+PAD-kx] w1[co][ci][ky][kx]
// (+ the same with a2, w2).  A thread owns a 2x2 block of ALL CI input channels of one image: the 6x6 adjoint patch of an
// output channel is read once (18 64-bit loads) for CI x 100 FMAs; the output channels are staged COCH at a time.
template <int CI, int CO, int HIN, int PAD, int IMGS, int COCH>
struct ConvBwdCfg {
  static constexpr int HP = HIN + 2 * PAD, HO = HP - 4, HQ = HO / 2, LP = 4 - PAD, HA = HO + 2 * LP, HB = HIN / 2;
  static constexpr int NA = COCH * HA * HA, NWP = CO * CI * WPAD, TPI = HB * HB, NT = (IMGS * TPI + 31) / 32 * 32;
  static constexpr size_t SMEM = (size_t)(IMGS * NA + NWP) * sizeof(float);
};
template <int CI, int CO, int HIN, int PAD, int IMGS, int COCH>
__global__ void __launch_bounds__((ConvBwdCfg<CI, CO, HIN, PAD, IMGS, COCH>::NT))
conv_bwd_data_kernel(const float* __restrict__ pb1, const float* __restrict__ w1, const float* __restrict__ pb2,
                     const float* __restrict__ w2, const uint8_t* __restrict__ sel, int R, float* __restrict__ out) {
  using C = ConvBwdCfg<CI, CO, HIN, PAD, IMGS, COCH>;
  constexpr int HQ = C::HQ, LP = C::LP, HA = C::HA, HB = C::HB, NA = C::NA, NWP = C::NWP, TPI = C::TPI, NT = C::NT;
  static_assert(HA % 2 == 0 && NA % 4 == 0 && CO % COCH == 0 && HIN % 2 == 0, "alignment of the vector loads");
  extern __shared__ __align__(16) float smem_f[];
  float* s_a = smem_f;              // [IMGS][COCH][HA][HA]
  float* s_w = s_a + IMGS * NA;     // [CO][CI][WPAD]
  const int s = blockIdx.y, tid = threadIdx.x, r_base = blockIdx.x * IMGS;
  const int img = tid / TPI, rem = tid % TPI, r = r_base + img;
  const bool active = img < IMGS && r < R;
  const int by = rem / HB, bx = rem % HB;
  float acc[CI][4];
#pragma unroll
  for (int ci = 0; ci < CI; ++ci) acc[ci][0] = acc[ci][1] = acc[ci][2] = acc[ci][3] = 0.f;
#pragma unroll 1
  for (int term = 0; term < 2; ++term) {
    const float* pb = term ? pb2 : pb1;
    const float* wg = term ? w2 : w1;
    if (!pb) break;
    __syncthreads();
    for (int i = tid; i < NWP; i += NT) {
      const int f = i / WPAD, t = i % WPAD;
      s_w[i] = t < 25 ? wg[(size_t)s * LN_P + f * 25 + t] : 0.f;
    }
#pragma unroll 1
    for (int c0 = 0; c0 < CO; c0 += COCH) {
      if (c0) __syncthreads();
      for (int i = tid; i < IMGS * NA; i += NT) s_a[i] = 0.f;
      __syncthreads();
      for (int i = tid; i < IMGS * COCH * HQ * HQ; i += NT) {
        const int im = i / (COCH * HQ * HQ), j = i % (COCH * HQ * HQ), rr = r_base + im;
        if (rr < R) {
          const size_t gi = ((size_t)s * R + rr) * (CO * HQ * HQ) + c0 * HQ * HQ + j;
          const int cd = sel[gi];
          if (cd & 4) {
            const int c = j / (HQ * HQ), py = (j / HQ) % HQ, px = j % HQ, k = cd & 3;
            s_a[im * NA + (c * HA + LP + 2 * py + (k >> 1)) * HA + LP + 2 * px + (k & 1)] = pb[gi];
          }
        }
      }
      __syncthreads();
      if (active) {
#pragma unroll 1
        for (int c = 0; c < COCH; ++c) {
          const float* a = s_a + img * NA + (c * HA + 2 * by) * HA + 2 * bx;
          float pt[6][6];
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const float2 v = *reinterpret_cast<const float2*>(a + i * HA + 2 * j);
              pt[i][2 * j] = v.x;
              pt[i][2 * j + 1] = v.y;
            }
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) {
            const float4* w4 = reinterpret_cast<const float4*>(s_w + ((c0 + c) * CI + ci) * WPAD);
            float wv[WPAD];
#pragma unroll
            for (int i = 0; i < WPAD / 4; ++i) {
              const float4 v = w4[i];
              wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
              for (int kx = 0; kx < 5; ++kx) {
                const float wk = wv[ky * 5 + kx];
                acc[ci][0] = fmaf(pt[4 - ky][4 - kx], wk, acc[ci][0]);
                acc[ci][1] = fmaf(pt[4 - ky][5 - kx], wk, acc[ci][1]);
                acc[ci][2] = fmaf(pt[5 - ky][4 - kx], wk, acc[ci][2]);
                acc[ci][3] = fmaf(pt[5 - ky][5 - kx], wk, acc[ci][3]);
              }
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) {
      float* o = out + ((size_t)s * R + r) * (CI * HIN * HIN) + (ci * HIN + 2 * by) * HIN + 2 * bx;
      *reinterpret_cast<float2*>(o) = make_float2(acc[ci][0], acc[ci][1]);
      *reinterpret_cast<float2*>(o + HIN) = make_float2(acc[ci][2], acc[ci][3]);
    }
  }
}

template <int CI, int CO, int HIN, int PAD, int IMGS, int COG>
static void launch_conv_fwd(int S, int R, cudaStream_t st, const float* in1, size_t ss1, const float* w1, const float* in2,
                            size_t ss2, const float* w2, const float* bias, uint8_t* sel, int tangent, float* out) {
  using C = ConvFwdCfg<CI, CO, HIN, PAD, IMGS, COG>;
  auto k = conv_pool_fwd_kernel<CI, CO, HIN, PAD, IMGS, COG>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  k<<<dim3((R + IMGS - 1) / IMGS, S), C::NT, C::SMEM, st>>>(in1, ss1, w1, in2, ss2, w2, bias, R, sel, tangent, out);
}
template <int CI, int CO, int HIN, int PAD, int IMGS, int COCH>
static void launch_conv_bwd_data(int S, int R, cudaStream_t st, const float* pb1, const float* w1, const float* pb2,
                                 const float* w2, const uint8_t* sel, float* out) {
  using C = ConvBwdCfg<CI, CO, HIN, PAD, IMGS, COCH>;
  auto k = conv_bwd_data_kernel<CI, CO, HIN, PAD, IMGS, COCH>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  k<<<dim3((R + IMGS - 1) / IMGS, S), C::NT, C::SMEM, st>>>(pb1, w1, pb2, w2, sel, R, out);
}
#define CONV1_FWD launch_conv_fwd<1, 6, 28, 2, 2, 6>
#define CONV2_FWD launch_conv_fwd<6, 16, 14, 0, 5, 8>
#define CONV1_BWD launch_conv_bwd_data<1, 6, 28, 2, 2, 6>
#define CONV2_BWD launch_conv_bwd_data<6, 16, 14, 0, 5, 8>

// d/d(weights, bias) of the fused conv + ReLU + pool.  One CTA per (row chunk, sample) computes the contribution of its
// rows to ALL CO*CI*25 weight gradients and to the CO bias gradients, and writes a partial [chunk][s][CO*CI*25 + CO];
// conv_wgrad_reduce_kernel sums the chunks in fixed order:
//   wbar[s][co][ci][ky][kx] = sum_r sum_{y,x} a1[s][r][co][y][x] in1[(s)][r][ci][y+ky-PAD][x+kx-PAD] (+ the same with a2, in2)
//   bbar[s][co] = sum a1
// A thread owns one filter row (ci, ky) -- its 5 taps for ALL CO channels, 5*CO accumulators -- over the output rows of its
// row group g: per four output columns it reads 8 inputs (two 128-bit loads) and CO adjoint quads (warp-broadcast 128-bit
// loads) for 20*CO FMAs, so the loop is FMA-bound instead of shared-memory-bound.
template <int CI, int CO, int HIN, int PAD, int NG>
struct ConvWgCfg {
  static constexpr int HP = HIN + 2 * PAD, HO = HP - 4, HQ = HO / 2, HOP = (HO + 3) / 4 * 4;
  static constexpr int HPMIN = HP > HOP + 4 ? HP : HOP + 4;
  static constexpr int HPP = (HPMIN + 3) / 8 * 8 + 4;   // smallest pitch >= HPMIN with pitch % 8 == 4 (conflict-free 128-bit rows)
  static constexpr int NTR = CI * 5, NT = (NTR * NG + 31) / 32 * 32, NOUT = CO * CI * 25;
};
template <int CI, int CO, int HIN, int PAD, int NG>
__global__ void __launch_bounds__((ConvWgCfg<CI, CO, HIN, PAD, NG>::NT), 3)
conv_bwd_weight_kernel(const float* __restrict__ pb1, const float* __restrict__ in1, size_t ss1, const float* __restrict__ pb2,
                       const float* __restrict__ in2, size_t ss2, const uint8_t* __restrict__ sel, int R, int rows_per_chunk,
                       float* __restrict__ part) {
  using C = ConvWgCfg<CI, CO, HIN, PAD, NG>;
  constexpr int HP = C::HP, HO = C::HO, HQ = C::HQ, HOP = C::HOP, HPP = C::HPP, NTR = C::NTR, NT = C::NT, NOUT = C::NOUT;
  constexpr int NW = NT / 32, NB = (CO + NW - 1) / NW;
  static_assert(HPP >= C::HPMIN && HPP % 8 == 4, "input pitch");
  static_assert(HIN % 2 == 0 && PAD % 2 == 0 && HOP % 2 == 0, "64-bit staging");
  __shared__ __align__(16) float s_a[CO * HO * HOP], s_in[CI * HP * HPP], s_red[NOUT];
  const int chunk = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const int r0 = chunk * rows_per_chunk, r1 = min(R, r0 + rows_per_chunk);
  const int tr = tid % NTR, g = tid / NTR;
  const bool active = g < NG;
  const int ci = tr / 5, ky = tr % 5;
  const int warp = tid >> 5, lane = tid & 31;
  float acc[5][CO], bacc[NB];
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[k][co] = 0.f;
#pragma unroll
  for (int h = 0; h < NB; ++h) bacc[h] = 0.f;
  // the zero padding (border of the input, pad columns of the adjoint) is written once; the per-row staging only touches
  // the interior
  for (int i = tid; i < CI * HP * HPP; i += NT) s_in[i] = 0.f;
  for (int i = tid; i < CO * HO * HOP; i += NT) s_a[i] = 0.f;
  __syncthreads();
  // Work items = (row, term).  The global data of item i + 1 (its input map and pooled adjoints with their selection codes)
  // is fetched into REGISTERS right after the barrier that publishes item i, i.e. the loads fly while item i is multiplied
  // (ncu of the load -> barrier -> multiply form: 33 % of the warp samples waited on these loads).
  constexpr int NI = CI * HIN * (HIN / 2), NIT = (NI + NT - 1) / NT;   // 64-bit pieces of the input interior, per thread
  constexpr int NA = CO * HQ * HQ, NAT = (NA + NT - 1) / NT;            // pooled adjoints, per thread
  float2 pin[NIT];
  float pv[NAT];
  int pc[NAT];
  const int nterm = pb2 ? 2 : 1, n_items = (r1 - r0) * nterm;
  auto fetch = [&](int item) {
    const int r = r0 + item / nterm, term = item % nterm;
    const float* pb = term ? pb2 : pb1;
    const float* ip = (term ? in2 + (size_t)s * ss2 : in1 + (size_t)s * ss1) + (size_t)r * CI * HIN * HIN;
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int i = tid + k * NT;
      if (i < NI) {
        const int c = i / (HIN * (HIN / 2)), yy = (i / (HIN / 2)) % HIN, x2 = i % (HIN / 2);
        pin[k] = *reinterpret_cast<const float2*>(ip + (c * HIN + yy) * HIN + 2 * x2);
      }
    }
    const size_t pbase = ((size_t)s * R + r) * (CO * HQ * HQ);
#pragma unroll
    for (int k = 0; k < NAT; ++k) {
      const int i = tid + k * NT;
      if (i < NA) {
        pc[k] = sel[pbase + i];
        pv[k] = pb[pbase + i];
      }
    }
  };
  if (n_items > 0) fetch(0);
#pragma unroll 1
  for (int item = 0; item < n_items; ++item) {
    const int term = item % nterm;
    {
#pragma unroll
      for (int k = 0; k < NIT; ++k) {   // interior of the padded input, 64 bits at a time
        const int i = tid + k * NT;
        if (i < NI) {
          const int c = i / (HIN * (HIN / 2)), yy = (i / (HIN / 2)) % HIN, x2 = i % (HIN / 2);
          *reinterpret_cast<float2*>(s_in + (c * HP + yy + PAD) * HPP + 2 * x2 + PAD) = pin[k];
        }
      }
#pragma unroll
      for (int k = 0; k < NAT; ++k) {   // each pooled adjoint owns its 2x2 window of the full-resolution map
        const int i = tid + k * NT;
        if (i < NA) {
          const int co = i / (HQ * HQ), py = (i / HQ) % HQ, px = i % HQ, c = pc[k];
          const float v = (c & 4) ? pv[k] : 0.f;
          const int kk = c & 3;
          float* d = s_a + (co * HO + 2 * py) * HOP + 2 * px;
          *reinterpret_cast<float2*>(d) = make_float2(kk == 0 ? v : 0.f, kk == 1 ? v : 0.f);
          *reinterpret_cast<float2*>(d + HOP) = make_float2(kk == 2 ? v : 0.f, kk == 3 ? v : 0.f);
        }
      }
      __syncthreads();
      if (item + 1 < n_items) fetch(item + 1);
      if (active) {
#pragma unroll 1
        for (int y = g; y < HO; y += NG) {
          const float* irow = s_in + (ci * HP + y + ky) * HPP;
          const float* arow = s_a + y * HOP;
#pragma unroll 1
          for (int x4 = 0; x4 < HOP; x4 += 4) {
            const float4 i0 = *reinterpret_cast<const float4*>(irow + x4);
            const float4 i1 = *reinterpret_cast<const float4*>(irow + x4 + 4);
            const float iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
            for (int co = 0; co < CO; ++co) {
              const float4 a4 = *reinterpret_cast<const float4*>(arow + co * HO * HOP + x4);
              const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
              for (int k = 0; k < 5; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[k][co] = fmaf(av[j], iv[j + k], acc[k][co]);
            }
          }
        }
      }
      if (term == 0) {
#pragma unroll
        for (int h = 0; h < NB; ++h) {
          const int co = warp + NW * h;
          if (co < CO) {
            float t = 0.f;
            for (int pos = lane; pos < HO * HOP; pos += 32) t += s_a[co * HO * HOP + pos];
            bacc[h] += warp_sum(t);
          }
        }
      }
      __syncthreads();
    }
  }
  // reduce over the row groups in fixed order through shared memory, then write this chunk's partial
#pragma unroll 1
  for (int gg = 0; gg < NG; ++gg) {
    if (active && g == gg) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          float* d = s_red + (co * CI + ci) * 25 + ky * 5 + k;
          *d = gg ? *d + acc[k][co] : acc[k][co];
        }
    }
    __syncthreads();
  }
  float* dst = part + ((size_t)chunk * gridDim.y + s) * (NOUT + CO);
  for (int o = tid; o < NOUT; o += NT) dst[o] = s_red[o];
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < NB; ++h)
      if (warp + NW * h < CO) dst[NOUT + warp + NW * h] = bacc[h];
  }
}

// CTAs of `kernel` that can be resident on the device at once (one wave)
template <typename K>
static int resident_ctas(K kernel, int threads) {
  int per_sm = 1, dev = 0, sms = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return std::max(1, per_sm) * std::max(1, sms);
}

// wbar[s][..] (+)= sum_chunks part;  bbar[s][co] = sum_chunks  (bbar nullable)
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ part, int n_chunks, int S, int NOUT, int CO,
                                         float* __restrict__ wbar, float* __restrict__ bbar, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (i >= NOUT + CO) return;
  float t = 0.f;
  for (int c = 0; c < n_chunks; ++c) t += part[((size_t)c * S + s) * (NOUT + CO) + i];
  if (i < NOUT) {
    float* d = wbar + (size_t)s * LN_P + i;
    *d = accumulate ? *d + t : t;
  } else if (bbar) {
    bbar[(size_t)s * LN_P + (i - NOUT)] = t;
  }
}

// ------------------------------------------------------------------------------------------------ fully connected layers
// The three matrix products of a fully connected layer (forward, data adjoint, weight gradient) are all instances of one
// per-sample batched product   out[s][m][n] (+)= sum_k A1[s](m, k) B1[s](n, k) (+ sum_k A2[s](m, k) B2[s](n, k))
// with either operand stored k-contiguous (KC) or m/n-contiguous.  The matrices are small (at most 200 x 400 x 120 per
// sample) and L2-resident, so the kernel is built for latency: 32 x 32 output tiles (many CTAs), 32-deep k tiles in a
// 4-stage cp.async ring (three tiles in flight while one is multiplied out of shared memory), 2 x 2 register micro-tiles
// on 256 threads (4 x 4 on 64 threads left the SMs with two warps each: measured 3x slower).  Epilogue: + bias[n], relu, mask by (maskfrom > 0), accumulate into out; optional
// column sums  colsum[s][m] = sum_k A1(m, k)  (the bias gradient of the weight-gradient product), in fixed order.
struct BGemm {
  const float *A1, *B1, *A2, *B2;
  long long sA1, sB1, sA2, sB2;   // per-sample strides (elements)
  int lda1, ldb1, lda2, ldb2;     // pitch of the non-contiguous dimension
  int M, N, K;
  const float* bias; long long sBias;
  int relu;
  const float* maskfrom;
  float* out; long long sO; int ldo; int accumulate;
  float* colsum; long long sCol;
};
constexpr int G_T = 32, G_P = 36;   // tile edge (m, n and k) and shared-memory pitch

constexpr int G_ST = 4;   // k tiles in flight (cp.async ring)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;   // 0 source bytes: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <bool AKC, bool BKC, int TM, int TN>
__global__ void __launch_bounds__((G_T / TM) * (G_T / TN))
bgemm_kernel(const BGemm p) {
  constexpr int NT = (G_T / TM) * (G_T / TN), PER = G_T * G_T / NT;   // threads, tile elements copied per thread and operand
  static_assert((TM == 2 || TM == 4) && (TN == 2 || TN == 4), "micro-tile");
  __shared__ __align__(16) float As[G_ST][G_T][G_P], Bs[G_ST][G_T][G_P];   // [stage][k][m], [stage][k][n]
  const int tid = threadIdx.x, tx = tid % (G_T / TN), ty = tid / (G_T / TN);
  const int m0 = blockIdx.x * G_T, n0 = blockIdx.y * G_T, s = blockIdx.z;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float csum = 0.f;
  const bool want_csum = p.colsum && blockIdx.y == 0;
  const int nkt = (p.K + G_T - 1) / G_T, ntiles = (p.A2 ? 2 : 1) * nkt;   // the two terms form one sequence of k tiles
  // 4-byte asynchronous copies place every element at its transposed position directly (zero-filled out of range)
  auto issue = [&](int it) {
    if (it < ntiles) {
      const int term = it >= nkt, k0 = (it - term * nkt) * G_T, st = it % G_ST;
      const float* A = (term ? p.A2 : p.A1) + (size_t)s * (term ? p.sA2 : p.sA1);
      const float* B = (term ? p.B2 : p.B1) + (size_t)s * (term ? p.sB2 : p.sB1);
      const int lda = term ? p.lda2 : p.lda1, ldb = term ? p.ldb2 : p.ldb1;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int e = tid + j * NT;
        {
          const int mm = AKC ? (e >> 5) : (e & 31), kk = AKC ? (e & 31) : (e >> 5);
          const bool ok = m0 + mm < p.M && k0 + kk < p.K;
          const size_t off = AKC ? (size_t)(m0 + mm) * lda + k0 + kk : (size_t)(k0 + kk) * lda + m0 + mm;
          cp_async4(&As[st][kk][mm], ok ? A + off : A, ok);
        }
        {
          const int nn = BKC ? (e >> 5) : (e & 31), kk = BKC ? (e & 31) : (e >> 5);
          const bool ok = n0 + nn < p.N && k0 + kk < p.K;
          const size_t off = BKC ? (size_t)(n0 + nn) * ldb + k0 + kk : (size_t)(k0 + kk) * ldb + n0 + nn;
          cp_async4(&Bs[st][kk][nn], ok ? B + off : B, ok);
        }
      }
    }
    cp_async_commit();   // (possibly empty: keeps the group count uniform)
  };
#pragma unroll
  for (int it = 0; it < G_ST - 1; ++it) issue(it);
#pragma unroll 1
  for (int it = 0; it < ntiles; ++it) {
    cp_async_wait<G_ST - 2>();   // tile `it` has landed (this thread's copies) ...
    __syncthreads();             // ... and everybody's; everybody is also done with the stage refilled next
    issue(it + G_ST - 1);
    const int st = it % G_ST;
    if (want_csum && it < nkt && tid < G_T) {
#pragma unroll 8
      for (int k = 0; k < G_T; ++k) csum += As[st][k][tid];
    }
#pragma unroll
    for (int k = 0; k < G_T; ++k) {
      float av[TM], bv[TN];
      if (TM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[TM - 2] = a.z; av[TM - 1] = a.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(&As[st][k][ty * 2]);
        av[0] = a.x; av[1] = a.y;
      }
      if (TN == 4) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
        bv[0] = b.x; bv[1] = b.y; bv[TN - 2] = b.z; bv[TN - 1] = b.w;
      } else {
        const float2 b = *reinterpret_cast<const float2*>(&Bs[st][k][tx * 2]);
        bv[0] = b.x; bv[1] = b.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= p.M) break;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.N) break;
      const size_t dst = (size_t)s * p.sO + (size_t)m * p.ldo + n;
      float v = acc[i][j] + (p.bias ? p.bias[(size_t)s * p.sBias + n] : 0.f);
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.maskfrom && !(p.maskfrom[dst] > 0.f)) v = 0.f;
      p.out[dst] = p.accumulate ? p.out[dst] + v : v;
    }
  }
  if (want_csum && tid < G_T && m0 + tid < p.M) p.colsum[(size_t)s * p.sCol + m0 + tid] = csum;
}
constexpr int G_TM = 2, G_TN = 2, G_NT = (G_T / G_TM) * (G_T / G_TN);   // micro-tile of the launches below

// ------------------------------------------------------------------------------------------------ softmax / NLL head
// mode 0: nll;  mode 1: nll, g_o = cw (p - onehot);  mode 2 (dual, needs od): G_o = cw p (od - <p, od>), G_od = cw (p - onehot),
// acbar = (p - onehot) . od     
__global__ void lenet_head_kernel(const float* __restrict__ o, const float* __restrict__ od, const int* __restrict__ y,
                                  const float* __restrict__ cw, int S, int R, int mode, float* __restrict__ nll,
                                  float* __restrict__ g_o, float* __restrict__ g_od, float* __restrict__ acbar) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * R) return;
  const int r = idx % R;
  const float* lo = o + (size_t)idx * N_O;
  float p[N_O], mx = -INFINITY, se = 0.f;
#pragma unroll
  for (int c = 0; c < N_O; ++c) mx = fmaxf(mx, lo[c]);
#pragma unroll
  for (int c = 0; c < N_O; ++c) { p[c] = expf(lo[c] - mx); se += p[c]; }
  const int lab = y[r];
  if (nll) nll[idx] = -(lo[lab] - mx - logf(se));
  if (mode == 0) return;
  const float w = cw ? cw[idx] : 1.f, inv = 1.f / se;
#pragma unroll
  for (int c = 0; c < N_O; ++c) p[c] *= inv;
  if (mode == 1) {
#pragma unroll
    for (int c = 0; c < N_O; ++c) g_o[(size_t)idx * N_O + c] = w * (p[c] - (c == lab ? 1.f : 0.f));
    return;
  }
  const float* ld = od + (size_t)idx * N_O;
  float dot = 0.f, qd = 0.f;
#pragma unroll
  for (int c = 0; c < N_O; ++c) { dot += p[c] * ld[c]; qd += (p[c] - (c == lab ? 1.f : 0.f)) * ld[c]; }
#pragma unroll
  for (int c = 0; c < N_O; ++c) {
    g_o[(size_t)idx * N_O + c] = w * p[c] * (ld[c] - dot);
    g_od[(size_t)idx * N_O + c] = w * (p[c] - (c == lab ? 1.f : 0.f));
  }
  if (acbar) acbar[idx] = qd;
}

// predictive metrics from per-sample logits [S][R][C] (psvi_classes.py:1072-1092): one CTA, fixed-order reduction
__global__ void __launch_bounds__(256)
logits_predict_kernel(const float* __restrict__ logits, const float* __restrict__ lw, int mode, const int* __restrict__ y, int S,
                      int R, int C, float* __restrict__ out, float* __restrict__ probs_out) {
  __shared__ float s_w[64], s_red[2][256];
  const int tid = threadIdx.x;
  if (tid == 0) {
    float mx = -INFINITY, se = 0.f;
    if (mode == 0) {
      for (int s = 0; s < S; ++s) mx = fmaxf(mx, lw[s]);
      for (int s = 0; s < S; ++s) se += expf(lw[s] - mx);
    }
    for (int s = 0; s < S; ++s) s_w[s] = mode == 0 ? expf(lw[s] - mx) / se : 1.f / (float)S;
  }
  __syncthreads();
  float nll = 0.f, corr = 0.f;
  for (int r = tid; r < R; r += 256) {
    float pr[16];
    for (int c = 0; c < C; ++c) pr[c] = 0.f;
    if (mode == 2) {   // softmax of the mean logits (mfvi baselines)
      for (int s = 0; s < S; ++s)
        for (int c = 0; c < C; ++c) pr[c] += logits[((size_t)s * R + r) * C + c] / (float)S;
      float mx = -INFINITY, se = 0.f;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, pr[c]);
      for (int c = 0; c < C; ++c) { pr[c] = expf(pr[c] - mx); se += pr[c]; }
      for (int c = 0; c < C; ++c) pr[c] /= se;
    } else {
      for (int s = 0; s < S; ++s) {
        const float* lo = logits + ((size_t)s * R + r) * C;
        float mx = -INFINITY, se = 0.f;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, lo[c]);
        for (int c = 0; c < C; ++c) se += expf(lo[c] - mx);
        const float sc = s_w[s] / se;
        for (int c = 0; c < C; ++c) pr[c] += sc * expf(lo[c] - mx);
      }
    }
    float tot = 0.f, best = -1.f;
    int am = 0;
    for (int c = 0; c < C; ++c) {
      tot += pr[c];
      if (pr[c] > best) { best = pr[c]; am = c; }
      if (probs_out) probs_out[(size_t)r * C + c] = pr[c];
    }
    if (!y) continue;
    const int lab = y[r];
    nll -= logf(fminf(fmaxf(pr[lab] / tot, 1.1920929e-07f), 1.f - 1.1920929e-07f));
    corr += (am == lab) ? 1.f : 0.f;
  }
  s_red[0][tid] = nll; s_red[1][tid] = corr;
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 256; ++i) { a += s_red[0][i]; b += s_red[1][i]; }
    out[0] = (float)a; out[1] = (float)b; out[2] = (float)R; out[3] = 0.f; out[4] = 0.f;
    if (mode == 0) {
      float ent = 0.f, sw = 0.f, sw2 = 0.f;
      for (int s = 0; s < S; ++s) {
        const float w = s_w[s];
        if (w > 0.f) ent -= logf(w) * w;
        sw += w; sw2 += w * w;
      }
      out[3] = ent;
      out[4] = sw * sw / sw2 / (float)S;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host sequencing
struct Ws {
  float *p1, *p2, *h3, *h4, *o, *pd1, *pd2, *hd3, *hd4, *od;
  float *g1, *g1d, *g2, *g2d, *g3, *g3d, *g4, *g4d, *go, *god;
  uint8_t *sel1, *sel2;
  float* wpart;   // conv weight-gradient partials [N_CHUNKS][S][2416]
  size_t total;
};
void carve_ws(int S, int R, uint8_t* base, Ws& w) {
  size_t off = 0;
  const size_t n = (size_t)S * R;
  auto takef = [&](size_t floats) { float* p = base ? reinterpret_cast<float*>(base + off) : nullptr; off += ((floats * 4 + 255) & ~(size_t)255); return p; };
  w.p1 = takef(n * N_P1); w.p2 = takef(n * N_P2); w.h3 = takef(n * N_H3); w.h4 = takef(n * N_H4); w.o = takef(n * N_O);
  w.pd1 = takef(n * N_P1); w.pd2 = takef(n * N_P2); w.hd3 = takef(n * N_H3); w.hd4 = takef(n * N_H4); w.od = takef(n * N_O);
  w.g1 = takef(n * N_P1); w.g1d = takef(n * N_P1); w.g2 = takef(n * N_P2); w.g2d = takef(n * N_P2);
  w.g3 = takef(n * N_H3); w.g3d = takef(n * N_H3); w.g4 = takef(n * N_H4); w.g4d = takef(n * N_H4);
  w.go = takef(n * N_O); w.god = takef(n * N_O);
  w.wpart = takef((size_t)N_CHUNKS * S * 2416);
  w.sel1 = base ? base + off : nullptr; off += ((n * N_P1 + 255) & ~(size_t)255);
  w.sel2 = base ? base + off : nullptr; off += ((n * N_P2 + 255) & ~(size_t)255);
  w.total = off;
}

#define LN_CHECK() PSVI_CUDA_CHECK(cudaGetLastError())

}  // namespace

extern "C" {

int64_t psvi_lenet_num_theta(void) { return LN_P; }

size_t psvi_lenet_workspace_bytes(int32_t S, int32_t R) {
  if (S <= 0 || R <= 0) return 0;
  Ws w;
  carve_ws(S, R, nullptr, w);
  return w.total + 256;
}

// The weight-gradient kernels of a backward pass only produce outputs: nothing later in the pass reads them.  They run on a side
// stream, forked from the caller's stream after the kernel that produced their adjoint and joined at the end of the pass, so that
// they overlap the (small-grid) data-adjoint chain.  The fork / join pattern is plain event record / wait: it is captured as
// parallel branches when the caller's stream is being captured into a CUDA graph.  One side stream and event set per device,
// created on first use (the first call of a process is never inside a capture: the graph path warms up eagerly).
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[16];
  cudaEvent_t done = nullptr;
};
static SideStream* side_stream() {
  static SideStream tab[32];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) return nullptr;
  SideStream& t = tab[dev];
  if (!t.s) {
    if (cudaStreamCreateWithFlags(&t.s, cudaStreamNonBlocking) != cudaSuccess) { t.s = nullptr; return nullptr; }
    for (auto& e : t.ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming);
  }
  return &t;
}

int psvi_lenet_pass(int32_t S, const float* theta, const float* thetad, const float* x, const int32_t* y, const float* cw,
                    int32_t R, float* nll, float* tbar, float* tdbar, float* xbar, float* acbar, float* logits,
                    void* workspace, void* stream_) {
  PSVI_REQUIRE(theta && x && y && workspace, PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(S >= 1 && S <= 64 && R >= 1, PSVI_ERR_INVALID, "bad S / R");
  PSVI_REQUIRE(!thetad || (tbar && tdbar), PSVI_ERR_INVALID, "the dual pass needs tbar and tdbar");
  PSVI_REQUIRE(R * 4 * 4 <= 200 * 1024, PSVI_ERR_UNSUPPORTED, "at most 12800 rows per call");
  cudaStream_t st = (cudaStream_t)stream_;
  SideStream* side = (tbar && !getenv("PSVI_LENET_NO_FORK")) ? side_stream() : nullptr;
  cudaStream_t wst = side ? side->s : st;     // where the weight-gradient kernels go
  int n_ev = 0;
  auto fork = [&]() {                         // everything launched on st so far is visible to what follows on wst
    if (side) { cudaEventRecord(side->ev[n_ev], st); cudaStreamWaitEvent(wst, side->ev[n_ev], 0); ++n_ev; }
  };
  auto join = [&]() {
    if (side) { cudaEventRecord(side->done, wst); cudaStreamWaitEvent(st, side->done, 0); }
  };
  Ws w;
  carve_ws(S, R, reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255), w);
  const size_t sR = (size_t)R;
  auto tiles = [](int n) { return (n + G_T - 1) / G_T; };
  // out[s][r][o] = bias[s][o] + sum_i x1[s][r][i] w1[s][o][i] (+ x2 . w2); relu; * (mask > 0)
  auto lin_fwd = [&](const float* x1, const float* w1, const float* x2, const float* w2, const float* b, int IN, int OUT,
                     int relu, const float* mask, float* out) {
    BGemm g{};
    g.A1 = x1; g.B1 = w1; g.A2 = x2; g.B2 = w2;
    g.sA1 = g.sA2 = (long long)sR * IN; g.sB1 = g.sB2 = LN_P;
    g.lda1 = g.lda2 = g.ldb1 = g.ldb2 = IN;
    g.M = R; g.N = OUT; g.K = IN;
    g.bias = b; g.sBias = LN_P; g.relu = relu; g.maskfrom = mask;
    g.out = out; g.sO = (long long)sR * OUT; g.ldo = OUT;
    bgemm_kernel<true, true, G_TM, G_TN><<<dim3(tiles(R), tiles(OUT), S), G_NT, 0, st>>>(g);
  };
  // out[s][r][i] = sum_o y1[s][r][o] w1[s][o][i] (+ y2 . w2); * (mask > 0)
  auto lin_bwd_data = [&](const float* y1, const float* w1, const float* y2, const float* w2, int IN, int OUT, const float* mask,
                          float* out) {
    BGemm g{};
    g.A1 = y1; g.B1 = w1; g.A2 = y2; g.B2 = w2;
    g.sA1 = g.sA2 = (long long)sR * OUT; g.sB1 = g.sB2 = LN_P;
    g.lda1 = g.lda2 = OUT; g.ldb1 = g.ldb2 = IN;
    g.M = R; g.N = IN; g.K = OUT;
    g.maskfrom = mask;
    g.out = out; g.sO = (long long)sR * IN; g.ldo = IN;
    bgemm_kernel<true, false, G_TM, G_TN><<<dim3(tiles(R), tiles(IN), S), G_NT, 0, st>>>(g);
  };
  // wbar[s][o][i] = sum_r y1[s][r][o] x1[s][r][i] (+ y2 . x2);  bbar[s][o] = sum_r y1[s][r][o]
  auto lin_bwd_weight = [&](const float* y1, const float* x1, const float* y2, const float* x2, int IN, int OUT, float* wb,
                            float* bb) {
    BGemm g{};
    g.A1 = y1; g.B1 = x1; g.A2 = y2; g.B2 = x2;
    g.sA1 = g.sA2 = (long long)sR * OUT; g.sB1 = g.sB2 = (long long)sR * IN;
    g.lda1 = g.lda2 = OUT; g.ldb1 = g.ldb2 = IN;
    g.M = OUT; g.N = IN; g.K = R;
    g.out = wb; g.sO = LN_P; g.ldo = IN;
    g.colsum = bb; g.sCol = LN_P;
    bgemm_kernel<false, false, G_TM, G_TN><<<dim3(tiles(OUT), tiles(IN), S), G_NT, 0, wst>>>(g);
  };
  // row chunks of the weight-gradient kernels: as many as fit in ONE wave of resident CTAs (at most N_CHUNKS)
  auto chunking = [&](int slots, int& rpc, int& nch) {
    const int target = std::max(1, std::min(N_CHUNKS, slots / S));
    rpc = (R + target - 1) / target;
    nch = (R + rpc - 1) / rpc;
  };
  auto conv2_wgrad = [&](const float* pb1, const float* i1, const float* pb2, const float* i2, float* wb, float* bb) {
    using C = ConvWgCfg<6, 16, 14, 0, 5>;
    auto k = conv_bwd_weight_kernel<6, 16, 14, 0, 5>;
    static const int slots = resident_ctas(k, C::NT);
    int rpc, nch;
    chunking(slots, rpc, nch);
    k<<<dim3(nch, S), C::NT, 0, wst>>>(pb1, i1, sR * N_P1, pb2, i2, sR * N_P1, w.sel2, R, rpc, w.wpart);
    conv_wgrad_reduce_kernel<<<dim3((2416 + 127) / 128, S), 128, 0, wst>>>(w.wpart, nch, S, 2400, 16, wb, bb, 0);
  };
  auto conv1_wgrad = [&](const float* pbv, float* wb, float* bb) {
    using C = ConvWgCfg<1, 6, 28, 2, 28>;
    auto k = conv_bwd_weight_kernel<1, 6, 28, 2, 28>;
    static const int slots = resident_ctas(k, C::NT);
    int rpc, nch;
    chunking(slots, rpc, nch);
    k<<<dim3(nch, S), C::NT, 0, wst>>>(pbv, x, 0, nullptr, nullptr, 0, w.sel1, R, rpc, w.wpart);
    conv_wgrad_reduce_kernel<<<dim3((156 + 127) / 128, S), 128, 0, wst>>>(w.wpart, nch, S, 150, 6, wb, bb, 0);
  };
  // ---- primal forward
  CONV1_FWD(S, R, st, x, 0, theta + O_W1, nullptr, 0, nullptr, theta + O_B1, w.sel1, 0, w.p1);
  CONV2_FWD(S, R, st, w.p1, sR * N_P1, theta + O_W2, nullptr, 0, nullptr, theta + O_B2,
            w.sel2, 0, w.p2);
  lin_fwd(w.p2, theta + O_W3, nullptr, nullptr, theta + O_B3, N_P2, N_H3, 1, nullptr, w.h3);
  lin_fwd(w.h3, theta + O_W4, nullptr, nullptr, theta + O_B4, N_H3, N_H4, 1, nullptr, w.h4);
  lin_fwd(w.h4, theta + O_W5, nullptr, nullptr, theta + O_B5, N_H4, N_O, 0, nullptr, w.o);
  LN_CHECK();
  if (logits) PSVI_CUDA_CHECK(cudaMemcpyAsync(logits, w.o, sR * S * N_O * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int hb = (S * R + 127) / 128;
  if (!tbar) {
    lenet_head_kernel<<<hb, 128, 0, st>>>(w.o, nullptr, y, nullptr, S, R, 0, nll, nullptr, nullptr, nullptr);
    LN_CHECK();
    return PSVI_OK;
  }
  if (!thetad) {
    // ---- gradient pass
    lenet_head_kernel<<<hb, 128, 0, st>>>(w.o, nullptr, y, cw, S, R, 1, nll, w.go, nullptr, nullptr);
    fork();
    lin_bwd_weight(w.go, w.h4, nullptr, nullptr, N_H4, N_O, tbar + O_W5, tbar + O_B5);
    lin_bwd_data(w.go, theta + O_W5, nullptr, nullptr, N_H4, N_O, w.h4, w.g4);
    fork();
    lin_bwd_weight(w.g4, w.h3, nullptr, nullptr, N_H3, N_H4, tbar + O_W4, tbar + O_B4);
    lin_bwd_data(w.g4, theta + O_W4, nullptr, nullptr, N_H3, N_H4, w.h3, w.g3);
    fork();
    lin_bwd_weight(w.g3, w.p2, nullptr, nullptr, N_P2, N_H3, tbar + O_W3, tbar + O_B3);
    lin_bwd_data(w.g3, theta + O_W3, nullptr, nullptr, N_P2, N_H3, nullptr, w.g2);
    fork();
    conv2_wgrad(w.g2, w.p1, nullptr, nullptr, tbar + O_W2, tbar + O_B2);
    CONV2_BWD(S, R, st, w.g2, theta + O_W2, nullptr, nullptr, w.sel2, w.g1);
    fork();
    conv1_wgrad(w.g1, tbar + O_W1, tbar + O_B1);
    if (xbar) CONV1_BWD(S, R, st, w.g1, theta + O_W1, nullptr, nullptr, w.sel1, xbar);
    join();
    LN_CHECK();
    return PSVI_OK;
  }
  // ---- dual pass: tangent forward (x itself carries no tangent)
  CONV1_FWD(S, R, st, x, 0, thetad + O_W1, nullptr, 0, nullptr, thetad + O_B1, w.sel1, 1, w.pd1);
  CONV2_FWD(S, R, st, w.pd1, sR * N_P1, theta + O_W2, w.p1, sR * N_P1, thetad + O_W2,
            thetad + O_B2, w.sel2, 1, w.pd2);
  lin_fwd(w.pd2, theta + O_W3, w.p2, thetad + O_W3, thetad + O_B3, N_P2, N_H3, 0, w.h3, w.hd3);
  lin_fwd(w.hd3, theta + O_W4, w.h3, thetad + O_W4, thetad + O_B4, N_H3, N_H4, 0, w.h4, w.hd4);
  lin_fwd(w.hd4, theta + O_W5, w.h4, thetad + O_W5, thetad + O_B5, N_H4, N_O, 0, nullptr, w.od);
  lenet_head_kernel<<<hb, 128, 0, st>>>(w.o, w.od, y, cw, S, R, 2, nll, w.go, w.god, acbar);
  LN_CHECK();
  // layer 5
  fork();
  lin_bwd_weight(w.go, w.h4, w.god, w.hd4, N_H4, N_O, tbar + O_W5, tbar + O_B5);
  lin_bwd_weight(w.god, w.h4, nullptr, nullptr, N_H4, N_O, tdbar + O_W5, tdbar + O_B5);
  lin_bwd_data(w.go, theta + O_W5, w.god, thetad + O_W5, N_H4, N_O, w.h4, w.g4);
  lin_bwd_data(w.god, theta + O_W5, nullptr, nullptr, N_H4, N_O, w.h4, w.g4d);
  // layer 4
  fork();
  lin_bwd_weight(w.g4, w.h3, w.g4d, w.hd3, N_H3, N_H4, tbar + O_W4, tbar + O_B4);
  lin_bwd_weight(w.g4d, w.h3, nullptr, nullptr, N_H3, N_H4, tdbar + O_W4, tdbar + O_B4);
  lin_bwd_data(w.g4, theta + O_W4, w.g4d, thetad + O_W4, N_H3, N_H4, w.h3, w.g3);
  lin_bwd_data(w.g4d, theta + O_W4, nullptr, nullptr, N_H3, N_H4, w.h3, w.g3d);
  // layer 3
  fork();
  lin_bwd_weight(w.g3, w.p2, w.g3d, w.pd2, N_P2, N_H3, tbar + O_W3, tbar + O_B3);
  lin_bwd_weight(w.g3d, w.p2, nullptr, nullptr, N_P2, N_H3, tdbar + O_W3, tdbar + O_B3);
  lin_bwd_data(w.g3, theta + O_W3, w.g3d, thetad + O_W3, N_P2, N_H3, nullptr, w.g2);
  lin_bwd_data(w.g3d, theta + O_W3, nullptr, nullptr, N_P2, N_H3, nullptr, w.g2d);
  LN_CHECK();
  // conv 2
  fork();
  conv2_wgrad(w.g2, w.p1, w.g2d, w.pd1, tbar + O_W2, tbar + O_B2);
  conv2_wgrad(w.g2d, w.p1, nullptr, nullptr, tdbar + O_W2, tdbar + O_B2);
  CONV2_BWD(S, R, st, w.g2, theta + O_W2, w.g2d, thetad + O_W2, w.sel2, w.g1);
  CONV2_BWD(S, R, st, w.g2d, theta + O_W2, nullptr, nullptr, w.sel2, w.g1d);
  // conv 1
  fork();
  conv1_wgrad(w.g1, tbar + O_W1, tbar + O_B1);
  conv1_wgrad(w.g1d, tdbar + O_W1, tdbar + O_B1);
  if (xbar) CONV1_BWD(S, R, st, w.g1, theta + O_W1, w.g1d, thetad + O_W1, w.sel1, xbar);
  join();
  LN_CHECK();
  return PSVI_OK;
}

int psvi_logits_predict(const float* logits, const float* log_weights, int32_t mode, const int32_t* yt, int32_t S, int32_t R,
                        int32_t C, float* out, float* probs_out, void* stream) {
  PSVI_REQUIRE(logits && out && (yt || probs_out), PSVI_ERR_INVALID, "null pointer");
  PSVI_REQUIRE(S >= 1 && S <= 64 && C >= 1 && C <= 16 && R >= 1, PSVI_ERR_INVALID, "bad S / C / R");
  PSVI_REQUIRE(mode >= 0 && mode <= 2 && (mode != 0 || log_weights), PSVI_ERR_INVALID, "bad mode / missing log weights");
  logits_predict_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, log_weights, mode, yt, S, R, C, out, probs_out);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
