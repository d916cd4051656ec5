// Micro-benchmarks of the latencies that bound k_onchip (run: nvcc -arch=sm_100a -O3 -o ubench ubench.cu && ./ubench)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, int n) {
  __shared__ double sm[4096];
  __shared__ unsigned short idx[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) { sm[i] = 1.0 + i * 1e-9; idx[i] = (unsigned short)((i * 7 + 1) & 4095); }
  __syncthreads();
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) a = fma(a, b, c);
  t1 = clock64();
  if (threadIdx.x == 0) t[0] = (t1 - t0);
  // 2. dependent LDS chain (pointer chasing, u16 index)
  int p = threadIdx.x & 4095;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) p = idx[p];
  t1 = clock64();
  if (threadIdx.x == 0) t[1] = (t1 - t0);
  a += p;
  // 3. LDS.64 -> DFMA dependent (address independent)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) a = fma(sm[(threadIdx.x + i) & 4095], b, a);
  t1 = clock64();
  if (threadIdx.x == 0) t[2] = (t1 - t0);
  // 4. DP division chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < 64; ++i) a = 1.0 / (a + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0) t[3] = (t1 - t0);
  // 5. __syncthreads
  t0 = clock64();
  for (int i = 0; i < 64; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) t[4] = (t1 - t0);
  // 6. shuffle chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1);
  t1 = clock64();
  if (threadIdx.x == 0) t[5] = (t1 - t0);
  // 7. smem RMW chain: load, fma, store to same address
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < 256; ++i) { double v = sm[threadIdx.x]; v = fma(v, b, c); sm[threadIdx.x] = v; }
  t1 = clock64();
  if (threadIdx.x == 0) t[6] = (t1 - t0);
  // 8. global (L2) dependent load chain
  const int* g = reinterpret_cast<const int*>(out + 1024);
  int q = threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < 64; ++i) q = g[q & 0xffff] + (q & 1023);
  t1 = clock64();
  if (threadIdx.x == 0) t[7] = (t1 - t0);
  // 9. __syncwarp
  t0 = clock64();
  for (int i = 0; i < 256; ++i) __syncwarp();
  t1 = clock64();
  if (threadIdx.x == 0) t[8] = (t1 - t0);
  out[threadIdx.x + blockIdx.x * blockDim.x] = a + q + sm[5];
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 1 << 22); cudaMemset(out, 0, 1 << 22); cudaMalloc(&t, 128);
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, t, 0); k<<<1, threads>>>(out, t, 0);
    long long h[9]; cudaMemcpy(h, t, sizeof(h), cudaMemcpyDeviceToHost);
    printf("threads=%d: DFMA %.1f | LDS chase %.1f | LDS->DFMA %.1f | DDIV %.1f | syncthreads %.1f | SHFL+DADD %.1f | smem RMW %.1f | L2 chase %.1f | syncwarp %.1f  (cycles each)\n",
           threads, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 64.0, h[4] / 64.0, h[5] / 256.0, h[6] / 256.0, h[7] / 64.0, h[8] / 256.0);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
