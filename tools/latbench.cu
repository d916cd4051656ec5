// Latency micro-benchmarks on one SM (sm_100a): dependent DFMA, shared-memory pointer chase, named barriers, rcp.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/latbench tools/latbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, int iters, int nwarp_bar) {
  __shared__ int chase[1024];
  __shared__ double sd[1024];
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += blockDim.x) { chase[i] = (i + 33) & 1023; sd[i] = 1.0 + i * 1e-9; }
  __syncthreads();
  double a = 1.0 + tid * 1e-12, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  if (tid < 32) {
    // dependent DFMA chain
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
      a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c);
    }
    t1 = clock64();
    if (tid == 0) t[0] = (t1 - t0) / (8LL * iters);
    // 8 independent DFMA streams (throughput per warp)
    double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      x0 = fma(x0, b, c); x1 = fma(x1, b, c); x2 = fma(x2, b, c); x3 = fma(x3, b, c);
      x4 = fma(x4, b, c); x5 = fma(x5, b, c); x6 = fma(x6, b, c); x7 = fma(x7, b, c);
    }
    t1 = clock64();
    if (tid == 0) t[1] = (t1 - t0) * 100 / (8LL * iters);
    a = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    // shared-memory pointer chase
    int p = tid;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) { p = chase[p]; p = chase[p]; p = chase[p]; p = chase[p]; }
    t1 = clock64();
    if (tid == 0) t[2] = (t1 - t0) / (4LL * iters);
    a += p;
    // LDS.64 -> DFMA -> STS -> LDS dependent loop
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      double v = sd[(tid + i) & 1023];
      v = fma(v, b, c);
      sd[(tid + i + 1) & 1023] = v;
      __syncwarp();
    }
    t1 = clock64();
    if (tid == 0) t[3] = (t1 - t0) / iters;
    // rcp: MUFU.RCP64H + 2 Newton steps, dependent
    double r = a;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      double x;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(r));
      double e = fma(-r, x, 1.0); x = fma(x, e, x); e = fma(-r, x, 1.0); x = fma(x, e, x);
      r = x + 1.5;
    }
    t1 = clock64();
    if (tid == 0) t[4] = (t1 - t0) / iters;
    a += r;
    // IEEE division and sqrt
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) r = 1.0 / r + 1.5;
    t1 = clock64();
    if (tid == 0) t[7] = (t1 - t0) / iters;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) r = sqrt(r) + 1.5;
    t1 = clock64();
    if (tid == 0) t[8] = (t1 - t0) / iters;
    a += r;
  }
  __syncthreads();
  // named barrier among nwarp_bar warps, back to back
  if (tid < 32 * nwarp_bar) {
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) asm volatile("bar.sync 8, %0;" ::"r"(32 * nwarp_bar) : "memory");
    t1 = clock64();
    if (tid == 0) t[5] = (t1 - t0) / iters;
  }
  __syncthreads();
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64();
  if (tid == 0) t[6] = (t1 - t0) / iters;
  // exchange step: k stores, barrier, dependent load (what one column step of the dense chain op is made of)
  for (int ks = 0; ks <= 4; ++ks) {
    for (int nw = 2; nw <= 8; nw += 3) {
      __syncthreads();
      if (tid < 32 * nw) {
        double v = a;
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
          for (int q = 0; q < ks; ++q) sd[(tid * 2 + 64 * q + i) & 1023] = v;
          asm volatile("bar.sync 8, %0;" ::"r"(32 * nw) : "memory");
          v = fma(sd[(tid * 2 + 2 + i) & 1023], b, c);
        }
        t1 = clock64();
        if (tid == 0) t[16 + ks * 3 + (nw - 2) / 3] = (t1 - t0) / iters;
        a += v;
      }
    }
  }
  // the same exchange step with other synchronisation primitives (1 store per thread)
  __shared__ unsigned long long mbar;
  __shared__ unsigned int cnt;
  for (int var = 0; var < 5; ++var)
    for (int nw = 2; nw <= 8; nw += 3) {
      __syncthreads();
      if (tid == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&mbar)), "r"(32 * nw));
        cnt = 0;
      }
      __syncthreads();
      if (tid < 32 * nw) {
        double v = a;
        unsigned parity = 0, target = 0;
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
          sd[(tid * 2 + i) & 1023] = v;
          if (var == 0) {
            asm volatile("bar.sync 8, %0;" ::"r"(32 * nw) : "memory");
          } else if (var == 1) {   // mbarrier, every thread arrives
            unsigned long long st;
            asm volatile("mbarrier.arrive.shared.b64 %0, [%1];" : "=l"(st) : "r"(mb) : "memory");
            unsigned done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(mb), "r"(parity) : "memory");
            parity ^= 1;
          } else if (var == 2) {   // fence + shared counter, every thread
            __threadfence_block();
            atomicAdd(&cnt, 1u);
            target += 32 * nw;
            while (*(volatile unsigned*)&cnt < target) {}
          } else if (var == 3) {   // syncwarp, one lane per warp arrives on the counter
            __syncwarp();
            target += nw;
            if ((tid & 31) == 0) {
              __threadfence_block();
              atomicAdd(&cnt, 1u);
              while (*(volatile unsigned*)&cnt < target) {}
            }
            __syncwarp();
          } else {                 // bar.arrive early is impossible here; plain __syncthreads-like with .aligned
            asm volatile("barrier.sync.aligned 9, %0;" ::"r"(32 * nw) : "memory");
          }
          v = fma(sd[(tid * 2 + 64 + i) & 1023], b, c);
        }
        t1 = clock64();
        if (tid == 0) t[32 + var * 3 + (nw - 2) / 3] = (t1 - t0) / iters;
        a += v;
      }
    }
  __syncthreads();
  out[blockIdx.x * blockDim.x + tid] = a;
}
int main() {
  double* out; long long* t; cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&t, 64 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 256>>>(out, t, 2000, 3);
  long long h[64]; cudaMemcpy(h, t, sizeof(h), cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency %lld cyc | 8 independent DFMA: %.2f cyc each per warp | LDS chase %lld cyc | LDS->DFMA->STS->syncwarp loop %lld cyc | rcp (mufu + 4 dfma + add) %lld cyc | div+add %lld | sqrt+add %lld | named barrier 3 warps %lld | __syncthreads 8 warps %lld\n",
         h[0], h[1] / 100.0, h[2], h[3], h[4], h[7], h[8], h[5], h[6]);
  for (int ks = 0; ks <= 4; ++ks) printf("exchange step with %d STS.64 + bar + LDS + DFMA: %lld cyc (2 warps) %lld (5 warps) %lld (8 warps)\n", ks, h[16 + ks * 3], h[17 + ks * 3], h[18 + ks * 3]);
  const char* vn[5] = {"bar.sync", "mbarrier arrive + try_wait", "fence + counter (all threads)", "syncwarp + fence + counter (lane 0)", "barrier.sync.aligned"};
  for (int v = 0; v < 5; ++v) printf("1 STS + %s + LDS + DFMA: %lld cyc (2 warps) %lld (5 warps) %lld (8 warps)\n", vn[v], h[32 + v * 3], h[33 + v * 3], h[34 + v * 3]);
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
