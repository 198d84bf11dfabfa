// Cost of one inner round of the block-Jacobi sub-problem as in xtb_scf.cu (3-warp groups, 16x16 M in smem),
// for 1..10 concurrent groups and with parts disabled, to find the contended resource.
#include <cstdio>
#include <cuda_runtime.h>
#define MLD 24
#define QLD 20
// (static shared memory is capped at 48 KB: the Q rows overlap a little, irrelevant for timing)
__device__ __forceinline__ void gbar(int id) { asm volatile("bar.sync %0, 96;" ::"r"(id + 1) : "memory"); }
// VAR bit0: skip rsqrt chain; bit1: skip Q update; bit2: skip M update; bit3: M params via one LDS.128 pair only
template <int VAR>
__global__ void k(double* out, long long* cyc, int rounds, int ngroups) {
  __shared__ double Ms[8][16 * MLD], Qs[8][16 * QLD];
  __shared__ double2 rcss[8][16];
  __shared__ int2 rpqs[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / 3, vt = threadIdx.x - 96 * grp;
  if (grp >= ngroups) return;
  double* M = Ms[grp]; double* Q = Qs[grp]; double2* rcs = rcss[grp]; int2* rpq = rpqs[grp];
  for (int e = vt; e < 256; e += 96) {
    int r = e >> 4, c = e & 15;
    M[r * MLD + c] = (r == c) ? 1.0 + r + 0.1 * grp : 0.01 / (1 + abs(r - c));
    Q[r * QLD + c] = r == c;
  }
  gbar(grp);
  long long t0 = clock64();
  for (int it = 0; it < rounds; ++it) {
    const int t = it & 7, buf = it & 1;
    if (vt < 8) {
      const int l = vt, p = l, q = 8 + ((l + t) & 7);
      const double app = M[p * MLD + p], aqq = M[q * MLD + q], apq = M[p * MLD + q];
      double cs_ = 1.0, sn = 0.0;
      if (!(VAR & 1)) {
        const double d = aqq - app, x = d * d + 4.0 * apq * apq;
        if (x > 1e-280) {
          const double ir = rsqrt(x), c2 = 0.5 + 0.5 * fabs(d) * ir, ic = rsqrt(c2);
          cs_ = c2 * ic; sn = copysign(apq * ir, d * apq) * ic;
        }
      } else { cs_ = 0.8 + 1e-3 * apq; sn = 0.6 - app * 1e-9 + aqq * 1e-9; }
      rcs[8 * buf + l] = make_double2(cs_, sn); rpq[8 * buf + l] = make_int2(p, q);
    } else if (vt >= 32 && it > 0 && !(VAR & 2)) {
      for (int i2 = vt - 32; i2 < 128; i2 += 64) {
        const int qk = (i2 & 3) + 4 * ((i2 >> 4) & 1), qi = ((i2 >> 2) & 3) + 4 * (i2 >> 5);
        const int2 pqq = rpq[8 * (buf ^ 1) + qk]; const double2 csq = rcs[8 * (buf ^ 1) + qk];
        const double vp = Q[qi * QLD + pqq.x], vq = Q[qi * QLD + pqq.y];
        Q[qi * QLD + pqq.x] = csq.x * vp - csq.y * vq; Q[qi * QLD + pqq.y] = csq.y * vp + csq.x * vq;
      }
    }
    gbar(grp);
    if (vt < 64 && !(VAR & 4)) {
      const int kp = vt >> 3, kq = vt & 7;
      const int2 pq1 = rpq[8 * buf + kp], pq2 = rpq[8 * buf + kq];
      const double2 cs1 = rcs[8 * buf + kp], cs2 = rcs[8 * buf + kq];
      const int p1 = pq1.x, q1 = pq1.y, p2 = pq2.x, q2 = pq2.y;
      const double c1 = cs1.x, s1 = cs1.y, c2 = cs2.x, s2 = cs2.y;
      const double a00 = M[p1 * MLD + p2], a01 = M[p1 * MLD + q2], a10 = M[q1 * MLD + p2], a11 = M[q1 * MLD + q2];
      const double x00 = c1 * a00 - s1 * a10, x01 = c1 * a01 - s1 * a11, x10 = s1 * a00 + c1 * a10, x11 = s1 * a01 + c1 * a11;
      double y00 = c2 * x00 - s2 * x01, y01 = s2 * x00 + c2 * x01, y10 = c2 * x10 - s2 * x11, y11 = s2 * x10 + c2 * x11;
      if (kp == kq) { y01 = 0.0; y10 = 0.0; }
      M[p1 * MLD + p2] = y00; M[p1 * MLD + q2] = y01; M[q1 * MLD + p2] = y10; M[q1 * MLD + q2] = y11;
    }
    gbar(grp);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = M[vt % 16 * MLD + (vt / 16) % 16] + Q[vt % 16];
}
template <int VAR> void run(const char* name, int ng, double* out, long long* cyc) {
  int rounds = 4000;
  k<VAR><<<148, 1024>>>(out, cyc, rounds, ng);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-40s groups %2d: %5.0f cycles / inner round\n", name, ng, (double)h / rounds);
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
  for (int ng : {1, 2, 5, 8}) {
    run<0>("full", ng, out, cyc);
    run<1>("no rsqrt chain", ng, out, cyc);
    run<2>("no Q update", ng, out, cyc);
    run<3>("no rsqrt, no Q", ng, out, cyc);
    run<4>("no M update (rot+Q+barriers)", ng, out, cyc);
  }
  return 0;
}
