// Shared device helpers for the GFN1-xTB kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "xtb_b200.h"

#define XTB_DEV __device__ __forceinline__

namespace xtb {

constexpr double kEps = 2.220446049250313e-16;    // torch.finfo(float64).eps
constexpr double kTiny = 2.2250738585072014e-308;  // torch.finfo(float64).tiny
constexpr int kMaxGridY = 65535;  // CUDA limit of gridDim.y: launches with the molecule index in y are chunked
constexpr double kPi = 3.14159265358979323846;
constexpr double kSqrtPi3 = 5.568327996831707845;  // sqrt(pi)^3

// tad-mctc storch.cdist (p=2): sqrt(clamp(sum (xi-xj)^2, min=eps))
XTB_DEV double safe_dist(double dx, double dy, double dz) { return sqrt(fmax(dx * dx + dy * dy + dz * dz, kEps)); }

XTB_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
XTB_DEV double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions; `red` is >= 32 doubles of shared scratch. All threads get the result.
XTB_DEV double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : 0.0;
  r = warp_sum(r);
  return r;
}
XTB_DEV double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = (lane < nw) ? red[lane] : -1.0e300;
  r = warp_max(r);
  return r;
}

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace xtb
