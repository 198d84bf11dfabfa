// Cost of one inner round of the block-Jacobi sub-problem with ONE WARP per sub-problem (warp-synchronous variant),
// for 1..8 concurrent warps and with parts disabled.   nvcc -arch=sm_100a -O3 phase1_w1.cu -o phase1_w1
#include <cstdio>
#include <cuda_runtime.h>
#define MLD 24
#define QLD 20
// VAR bit0: skip rsqrt chain; bit1: skip Q update; bit2: skip M update; bit3: no syncwarp before stores;
// bit4: branch-free rotation (rsqrt.approx + 2 Newton steps) and Q loads batched before the stores
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
template <int VAR>
__global__ void k(double* out, long long* cyc, int rounds, int nwarps) {
  __shared__ double Ms[8][16 * MLD], Qs[8][16 * QLD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  double* M = Ms[warp]; double* Q = Qs[warp];
  for (int e = lane; e < 256; e += 32) {
    int r = e >> 4, c = e & 15;
    M[r * MLD + c] = (r == c) ? 1.0 + r + 0.1 * warp : 0.01 / (1 + abs(r - c));
    Q[r * QLD + c] = r == c;
  }
  __syncwarp();
  const int kk = lane & 7;
  const int qrow = 2 * ((lane >> 3) & 1) + (lane >> 4);
  double pc = 1.0, ps = 0.0; int pp_ = 0, pq_ = 8;
  long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
    const int t = it & 7;
    const int p = kk, q = 8 + ((kk + t) & 7);
    const double app = M[p * MLD + p], aqq = M[q * MLD + q], apq = M[p * MLD + q];
    int p1[2], q1[2];
    p1[0] = lane >> 3; q1[0] = 8 + ((p1[0] + t) & 7);
    p1[1] = (lane >> 3) + 4; q1[1] = 8 + ((p1[1] + t) & 7);
    double a00[2], a01[2], a10[2], a11[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      a00[i] = M[p1[i] * MLD + p]; a01[i] = M[p1[i] * MLD + q]; a10[i] = M[q1[i] * MLD + p]; a11[i] = M[q1[i] * MLD + q];
    }
    double cs_ = 1.0, sn = 0.0;
    if (VAR & 16) {
      const double d = aqq - app, x = fma(d, d, 4.0 * apq * apq);
      const double xs = fmax(x, 1e-280);
      const double ir = rsqrt_nr(xs), c2 = fma(0.5 * fabs(d), ir, 0.5), ic = rsqrt_nr(c2);
      const bool ok = x > 1e-280;
      cs_ = ok ? c2 * ic : 1.0; sn = ok ? copysign(apq * ir, d * apq) * ic : 0.0;
    } else if (!(VAR & 1)) {
      const double d = aqq - app, x = d * d + 4.0 * apq * apq;
      if (x > 1e-280) {
        const double ir = rsqrt(x), c2 = 0.5 + 0.5 * fabs(d) * ir, ic = rsqrt(c2);
        cs_ = c2 * ic; sn = copysign(apq * ir, d * apq) * ic;
      }
    } else { cs_ = 0.8 + 1e-3 * apq; sn = 0.6 - app * 1e-9 + aqq * 1e-9; }
    if (it > 0 && (VAR & 16)) {
      double vp[4], vq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { vp[j] = Q[(qrow + 4 * j) * QLD + pp_]; vq[j] = Q[(qrow + 4 * j) * QLD + pq_]; }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Q[(qrow + 4 * j) * QLD + pp_] = pc * vp[j] - ps * vq[j]; Q[(qrow + 4 * j) * QLD + pq_] = ps * vp[j] + pc * vq[j];
      }
    } else if (it > 0 && !(VAR & 2)) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qi = qrow + 4 * j;
        const double vp = Q[qi * QLD + pp_], vq = Q[qi * QLD + pq_];
        Q[qi * QLD + pp_] = pc * vp - ps * vq; Q[qi * QLD + pq_] = ps * vp + pc * vq;
      }
    }
    if (!(VAR & 8)) __syncwarp();
    if (!(VAR & 4)) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int kp = (lane >> 3) + 4 * i;
        const double c1 = __shfl_sync(0xffffffffu, cs_, kp), s1 = __shfl_sync(0xffffffffu, sn, kp);
        const double x00 = c1 * a00[i] - s1 * a10[i], x01 = c1 * a01[i] - s1 * a11[i];
        const double x10 = s1 * a00[i] + c1 * a10[i], x11 = s1 * a01[i] + c1 * a11[i];
        double y00 = cs_ * x00 - sn * x01, y01 = sn * x00 + cs_ * x01, y10 = cs_ * x10 - sn * x11, y11 = sn * x10 + cs_ * x11;
        if (kp == kk) { y01 = 0.0; y10 = 0.0; }
        M[p1[i] * MLD + p] = y00; M[p1[i] * MLD + q] = y01; M[q1[i] * MLD + p] = y10; M[q1[i] * MLD + q] = y11;
      }
    }
    pc = cs_; ps = sn; pp_ = p; pq_ = q;
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) cyc[warp] = t1 - t0;
  out[threadIdx.x] = M[lane] + Q[lane] + pc;
}
template <int VAR>
void run(const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 4096 * 8); cudaMalloc(&cyc, 64 * 8);
  const int rounds = 8000;
  for (int nw : {1, 2, 4, 5, 8}) {
    k<VAR><<<1, 512>>>(out, cyc, rounds, nw);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps %d: %.0f cycles / inner round\n", name, nw, (double)h[0] / rounds);
  }
}
int main() {
  run<0>("full");
  run<16>("batched Q + branch-free rot");
  run<24>("same, one syncwarp");
  run<18>("branch-free rot, no Q");
  return 0;
}
