// Latency of an exchange between two CTAs that are NOT in the same cluster (through L2: payload + release flag, acquire spin on
// the other side), against the same exchange between two CTAs of one cluster through distributed shared memory
// (st.async + mbarrier complete_tx, as psvi_mf_fn1.cu does).  Question behind it: would splitting the rows
// of a sample over two clusters pay, given that every phase of the cfg2 step would then need one L2 exchange?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scratch/l2_exchange_latency.cu -o scratch/l2_exchange_latency
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// two CTAs (any SMs): ping-pong.  Each side writes `words` floats of payload, then a release flag; the other side spins on the
// flag with acquire loads, reads the payload, answers.
__global__ void l2_pingpong(float* buf, unsigned* flags, int words, int iters, long long* cycles, float* sink) {
  const int me = blockIdx.x, other = me ^ 1;
  float* mine = buf + me * 4096;
  const float* theirs = buf + other * 4096;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    if (me == 0) {
      for (int i = threadIdx.x; i < words; i += blockDim.x) mine[i] = (float)(it + i);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        st_release(flags + 0, (unsigned)it);
        while (ld_acquire(flags + 1) < (unsigned)it) {}
      }
      __syncthreads();
      for (int i = threadIdx.x; i < words; i += blockDim.x) acc += __ldcg(theirs + i);
    } else {
      if (threadIdx.x == 0) while (ld_acquire(flags + 0) < (unsigned)it) {}
      __syncthreads();
      for (int i = threadIdx.x; i < words; i += blockDim.x) acc += __ldcg(theirs + i);
      for (int i = threadIdx.x; i < words; i += blockDim.x) mine[i] = acc + (float)i;
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        st_release(flags + 1, (unsigned)it);
      }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && me == 0) cycles[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// the same ping-pong inside one 2-CTA cluster: st.async of the payload with complete_tx on the receiver's mbarrier
__global__ void __cluster_dims__(2, 1, 1) dsmem_pingpong(int words, int iters, long long* cycles, float* sink) {
  __shared__ __align__(16) float recv[4096];
  __shared__ __align__(8) unsigned long long bar;
  cg::cluster_group cluster = cg::this_cluster();
  const int me = (int)cluster.block_rank(), other = me ^ 1;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();
  unsigned rrecv, rbar;
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rrecv) : "r"(smem_u32(recv)), "r"(other));
  asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar)), "r"(other));
  float acc = 0.f;
  unsigned parity = 0;
  auto send = [&](int it) {
    for (int i = threadIdx.x; i < words; i += blockDim.x)
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(rrecv + 4 * i),
                   "r"(__float_as_uint((float)(it + i))), "r"(rbar)
                   : "memory");
  };
  auto wait = [&]() {
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(words * 4) : "memory");
    unsigned done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(parity)
                   : "memory");
    parity ^= 1u;
    for (int i = threadIdx.x; i < words; i += blockDim.x) acc += recv[i];
    __syncthreads();
  };
  const long long t0 = clock64();
  for (int it = 1; it <= iters; ++it) {
    if (me == 0) { send(it); wait(); } else { wait(); send(it); }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && me == 0) cycles[0] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  cluster.sync();
}

int main() {
  float *buf, *sink;
  unsigned* flags;
  long long* cyc;
  cudaMalloc(&buf, 2 * 4096 * 4);
  cudaMalloc(&sink, 4096 * 4);
  cudaMalloc(&flags, 64);
  cudaMallocManaged(&cyc, 8);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 2000;
  for (int words : {1, 64, 512}) {
    cudaMemset(flags, 0, 64);
    cudaMemset(buf, 0, 2 * 4096 * 4);
    l2_pingpong<<<2, 256>>>(buf, flags, words, iters, cyc, sink);
    cudaDeviceSynchronize();
    const double l2 = (double)cyc[0] / iters;
    dsmem_pingpong<<<2, 256>>>(words, iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    const double ds = (double)cyc[0] / iters;
    printf("payload %4d words: round trip through L2 (release flag / acquire spin) %.0f cycles = %.2f us; inside a cluster "
           "(st.async + mbarrier) %.0f cycles = %.2f us   [%s]\n",
           words, l2, l2 / (clk * 1e-3), ds, ds / (clk * 1e-3), cudaGetErrorString(e));
  }
  return 0;
}
