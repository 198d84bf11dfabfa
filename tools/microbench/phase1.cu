// Cost of one inner round of the block-Jacobi sub-problem (128 threads, 16x16 M and Q in shared memory).
#include <cstdio>
#include <cuda_runtime.h>
#define MLD 17
#define QLD 20
__device__ __forceinline__ void gbar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id + 1) : "memory"); }
template <int VAR>
__global__ void k(double* out, long long* cyc, int rounds) {
  __shared__ double M[16 * MLD], Q[16 * QLD], rcs[16];
  __shared__ int rpq[16];
  const int gt = threadIdx.x & 127, grp = threadIdx.x >> 7;
  for (int e = gt; e < 256; e += 128) {
    int r = e >> 4, c = e & 15;
    M[r * MLD + c] = (r == c) ? 1.0 + r : 0.01 / (1 + abs(r - c));
    Q[r * QLD + c] = r == c;
  }
  gbar(grp);
  long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
    const int t = it & 7;
    if (gt < 8) {
      const int l = gt;
      const int p = l, q = 8 + ((l + t) & 7);
      const double app = M[p * MLD + p], aqq = M[q * MLD + q], apq = M[p * MLD + q];
      double cs_ = 1.0, sn = 0.0;
      if (VAR != 3) {
        const double d = aqq - app;
        const double x = d * d + 4.0 * apq * apq;
        if (x > 1e-280) {
          const double y = d + copysign(x * rsqrt(x), d);
          const double tt = 2.0 * apq * copysign(rsqrt(y * y), y);
          cs_ = rsqrt(1.0 + tt * tt);
          sn = tt * cs_;
        }
      } else { cs_ = 0.8 + 1e-3 * apq; sn = 0.6 - app * 1e-9 + aqq * 1e-9; }
      rcs[2 * l] = cs_; rcs[2 * l + 1] = sn; rpq[2 * l] = p; rpq[2 * l + 1] = q;
    }
    gbar(grp);
    if (VAR != 2) {
      const int kp = (gt >> 3) & 7, kq = gt & 7, qi = gt >> 3;
      const int p1 = rpq[2 * kp], q1 = rpq[2 * kp + 1], p2 = rpq[2 * kq], q2 = rpq[2 * kq + 1];
      const double c1 = rcs[2 * kp], s1 = rcs[2 * kp + 1], c2 = rcs[2 * kq], s2 = rcs[2 * kq + 1];
      const double vp = Q[qi * QLD + p2], vq = Q[qi * QLD + q2];
      double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
      if (gt < 64) { a00 = M[p1 * MLD + p2]; a01 = M[p1 * MLD + q2]; a10 = M[q1 * MLD + p2]; a11 = M[q1 * MLD + q2]; }
      if (VAR != 1) { Q[qi * QLD + p2] = c2 * vp - s2 * vq; Q[qi * QLD + q2] = s2 * vp + c2 * vq; }
      if (gt < 64) {
        const double x00 = c1 * a00 - s1 * a10, x01 = c1 * a01 - s1 * a11;
        const double x10 = s1 * a00 + c1 * a10, x11 = s1 * a01 + c1 * a11;
        double y00 = c2 * x00 - s2 * x01, y01 = s2 * x00 + c2 * x01;
        double y10 = c2 * x10 - s2 * x11, y11 = s2 * x10 + c2 * x11;
        if (kp == kq) { y01 = 0.0; y10 = 0.0; }
        M[p1 * MLD + p2] = y00; M[p1 * MLD + q2] = y01; M[q1 * MLD + p2] = y10; M[q1 * MLD + q2] = y11;
      }
    }
    gbar(grp);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = M[gt % 16 * MLD + (gt / 16) % 16] + Q[gt % 16];
}
template <int VAR> void run(const char* name, int nthreads, double* out, long long* cyc) {
  int rounds = 4000;
  k<VAR><<<148, nthreads>>>(out, cyc, rounds);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s threads/CTA %4d: %.0f cycles per inner round\n", name, nthreads, (double)h / rounds);
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
  for (int nt : {128, 640}) {
    run<0>("full inner round", nt, out, cyc);
    run<1>("without Q update", nt, out, cyc);
    run<2>("rotation + barriers only", nt, out, cyc);
    run<3>("trivial rotation (no rsqrt chain)", nt, out, cyc);
  }
  return 0;
}
