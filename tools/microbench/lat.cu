// Dependent-chain latencies on sm_100a (one warp): DFMA, DMUL, F2F, LDS, DMMA; DMMA throughput per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n) {
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = 1.0 + threadIdx.x * 1e-9;
  __syncthreads();
  double x = out[0], y = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, y, 1e-9);
  long long t1 = clock64();
  for (int i = 0; i < n; ++i) x = x * y;
  long long t2 = clock64();
  for (int i = 0; i < n; ++i) x = (double)((float)x) + 1e-9;
  long long t3 = clock64();
  int idx = (int)x & 63;
  for (int i = 0; i < n; ++i) idx = (int)sm[idx] & 63;
  long long t4 = clock64();
  double d0 = x, d1 = x;
  for (int i = 0; i < n; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(y), "d"(y));
  long long t5 = clock64();
  float f = (float)x;
  for (int i = 0; i < n; ++i) f = rsqrtf(f) + 1.0f;
  long long t6 = clock64();
  for (int i = 0; i < n; ++i) x = rsqrt(x) + 1.0;
  long long t7 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6;
  }
  out[threadIdx.x + blockIdx.x * blockDim.x] = x + idx + d0 + d1 + f;
}
__global__ void thr(double* out, int n) {  // DMMA / DFMA throughput: many warps, independent chains
  double d[8];
  for (int i = 0; i < 8; ++i) d[i] = threadIdx.x * 1e-3 + i;
  double y = 1.0000001;
  if (blockIdx.y == 0) {
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[2 * j]), "+d"(d[2 * j + 1]) : "d"(y), "d"(y));
  } else {
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = fma(d[j], y, 1e-9);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += d[i];
  out[threadIdx.x + blockDim.x * (blockIdx.x + gridDim.x * blockIdx.y)] = s;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 64);
  cudaMemset(out, 0, 1 << 24);
  int n = 4096;
  k<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize();
  k<<<1, 32>>>(out, cyc, n); cudaDeviceSynchronize();
  long long h[7]; cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
  const char* nm[7] = {"DFMA", "DMUL", "F2F.f32<->f64 + DADD", "LDS dependent", "DMMA dependent", "rsqrtf+FADD", "rsqrt(double)+DADD"};
  for (int i = 0; i < 7; ++i) printf("%-24s %.1f cycles/iter\n", nm[i], (double)h[i] / n);
  for (int which = 0; which < 2; ++which) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int nn = 20000; dim3 grid(148 * 2, 1);
    thr<<<dim3(148 * 2, 2), 512>>>(out, 10);
    cudaEventRecord(e0);
    if (which == 0) thr<<<dim3(148 * 2, 1), 512>>>(out, nn);
    else { // DFMA only: launch y=1 slice by offsetting: reuse kernel with gridDim.y=2 would run both; so run separate kernel config
      thr<<<dim3(148 * 2, 2), 512>>>(out, nn);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = 148.0 * 2 * 16;
    if (which == 0) printf("DMMA throughput: %.2f TFLOP/s\n", warps * nn * 4 * 512.0 / (ms * 1e-3) / 1e12);
    else printf("DMMA+DFMA concurrently: %.2f ms (DMMA alone above); DFMA-equivalent total %.2f TFLOP/s\n", ms, (warps * nn * 4 * 512.0 + warps * nn * 8 * 64.0) / (ms * 1e-3) / 1e12);
  }
  return 0;
}
