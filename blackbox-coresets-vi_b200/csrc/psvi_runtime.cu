// psvi_runtime.cu -- error plumbing, device probe and the stand-alone Philox normal generator of libpsvi_b200.
#include <stdarg.h>

#include "psvi_common.cuh"

namespace {
thread_local char g_err[1024] = "";
}

void psvi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
__global__ void philox_normal_kernel(uint64_t seed, uint32_t domain, int first_slab, int n_slabs, int S, int P,
                                     float* out) {
  const int n4 = (P + 3) >> 2;
  const long long total = (long long)n_slabs * S * n4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q4 = (int)(i % n4);
    const long long r = i / n4;
    const int s = (int)(r % S), sl = (int)(r / S);
    float e[4];
    philox_normal4(seed, domain, (uint32_t)(first_slab + sl), (uint32_t)s, (uint32_t)q4, e);
    float* o = out + ((size_t)sl * S + s) * P;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * q4 + j < P) o[4 * q4 + j] = e[j];
  }
}
}  // namespace

extern "C" {

const char* psvi_last_error(void) { return g_err; }

int psvi_device_sm_count(void) {
  int dev = 0, n = 0;
  PSVI_CUDA_CHECK(cudaGetDevice(&dev));
  PSVI_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

int psvi_philox_normal(uint64_t seed, uint32_t domain, int32_t first_slab, int32_t n_slabs, int32_t S, int32_t P,
                       float* out, void* stream) {
  PSVI_REQUIRE(out && n_slabs > 0 && S > 0 && P > 0 && first_slab >= 0, PSVI_ERR_INVALID, "bad argument");
  const long long total = (long long)n_slabs * S * ((P + 3) / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  philox_normal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(seed, domain, first_slab, n_slabs, S, P, out);
  PSVI_CUDA_CHECK(cudaGetLastError());
  return PSVI_OK;
}

}  // extern "C"
