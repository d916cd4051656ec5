// Device peaks measured in place (tno_device_peak): the FP64 denominators of the secondary roofline bench.py reports.
// MEASURED_PEAKS.json carries the HBM copy bandwidth and a bf16 GEMM figure only; north_star asks for the FP64 pipe and
// the FP64 tensor pipe, so both are measured here with register-resident kernels (no memory traffic):
//   what = 0  DFMA:  8 independent fused multiply-add chains per thread
//   what = 1  DMMA:  mma.sync.aligned.m8n8k4.row.col.f64 (the FP64 tensor-core instruction), 4 accumulators per warp
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "tno_plan.hpp"

namespace tno {

__global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters) {
  double a[8];
  const double b = 1.0 + 1e-9 * threadIdx.x, c = 1e-12;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_peak_dmma(double* out, int iters) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

}  // namespace tno

using namespace tno;

extern "C" int tno_device_peak(int32_t device, int32_t what, double* value) {
  if (!value || what < 0 || what > 1) {
    set_error("tno_device_peak: bad argument");
    return TNO_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    set_error("no usable CUDA device");
    return TNO_ENODEVICE;
  }
  TNO_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  TNO_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  TNO_CUDA_OK(cudaMalloc((void**)&out, 64));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = prop.multiProcessorCount * 8, threads = 256;
  double best = 0.0;
  int iters = 2000;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    if (what == 0)
      k_peak_dfma<<<grid, threads>>>(out, iters);
    else
      k_peak_dmma<<<grid, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) {
      set_error(std::string("tno_device_peak: ") + cudaGetErrorString(err));
      cudaFree(out);
      return TNO_ECUDA;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    // flops per launch: DFMA 2 flops x 32 per thread-iteration; DMMA 2 * 8 * 8 * 4 = 512 flops x 16 per warp-iteration
    const double flops = what == 0 ? 2.0 * 32.0 * iters * (double)grid * threads : 512.0 * 16.0 * iters * (double)grid * (threads / 32);
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    if (ms < 5.f) iters *= 4;   // make a launch last a few milliseconds
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *value = best;
  return TNO_OK;
}
