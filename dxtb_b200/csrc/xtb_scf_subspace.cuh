// Occupied-subspace solve for the INTERMEDIATE SCF map evaluations of the one-CTA kernel (xtb_scf.cu).
//
// The reference diagonalises the Fock matrix in every iteration (scf/unrolling/base.py:141-175) although the iteration only
// needs the charges of the density it defines (scf/base.py:651-675).  For a closed-shell molecule whose HOMO-LUMO gap is
// many kT wide the Fermi occupations (wavefunction/filling.py:201-366) are 2 / 0 to round-off, so the density is the projector
// onto the occupied subspace and no individual eigenvector is needed.  In the basis C of an earlier (partial) diagonalisation
// the projected Fock matrix A = C^T F C = [[Aoo, Aov], [Avo, Avv]] is nearly block diagonal; the occupied subspace is the
// graph span([1; X]) of the solution of the algebraic Riccati equation
//     R(X) = Avo + Avv X - X (Aoo + Aov X) = 0,
// which the diagonally preconditioned fixed point X <- X - R(X) / (d_a - d_i) reaches in 3-9 iterations of two small
// tensor-core GEMMs (a Jacobi sweep costs ~10x as much), warm-started from the previous map evaluation.  Then
//     P = 2 Y (1 + X^T X)^-1 Y^T,   Y = C_o + C_v X,
// with the inverse by Newton's iteration Z <- Z (2 - G Z), also warm-started.  Everything is GEMM-shaped (DMMA).
//
// Safeguards (each falls back to a Jacobi sweep, i.e. to the path that was there before):
//  * integer occupations are CERTIFIED per map evaluation: with Gershgorin discs inside the two diagonal blocks and Cauchy
//    interlacing, gap(A) >= min_a (d_a - sum_{b != a} |A_ab|) - max_i (d_i + sum_{j != i} |A_ij|); the path is taken only if
//    that bound is >= subspace_gap * kT (50 kT: occupation error exp(-25) ~ 1.4e-11);
//  * the o / v classification follows the ranks of the current diagonal; a change permutes C and restarts X;
//  * a stalled, diverging or large (|X| >= 1) fixed point, or a failed Newton iteration, triggers a sweep.
// The final solve that defines energies, charges, P, W and the orbital energies is always the full Jacobi solve.
#pragma once
#include "xtb_scf_core.cuh"

namespace {

// Block-wide maxima of two values with one pair of barriers (`red` holds 32 doubles; at most 16 warps).
__device__ __forceinline__ void block_max2(double& a, double& b, double* red) {
  static_assert(NT <= 512, "block_max2 assumes at most 16 warps");
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_max(a);
  b = warp_max(b);
  __syncthreads();
  if (lane == 0) { red[w] = a; red[16 + w] = b; }
  __syncthreads();
  double r = ((lane & 15) < NW) ? red[lane] : -1.0e300;  // lanes 0-15: a of the warps, lanes 16-31: b
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  a = __shfl_sync(0xffffffffu, r, 0);
  b = __shfl_sync(0xffffffffu, r, 16);
}

// Certified gap between the occupied and the virtual class of the diagonal of A (header comment; diagonal -> c.eps).
// RANKED = false: classes by position (first `no` columns occupied) -- valid for ANY partition, cheap, used once the basis is
// in occupied-first order.  RANKED = true: classes by the ranks of the diagonal (-> c.occl); needs_perm tells the caller
// that the occupied class is not the first `no` columns.
template <bool RANKED>
XTB_CTX_FN double subspace_certify(Ctx& c, const double* __restrict__ A, bool& needs_perm) {
  const int n = c.n, ld = c.ld, no = c.sub.no;
  int* rank = c.occl;
  for (int k = threadIdx.x; k < n; k += NT) c.eps[k] = A[(size_t)k * ld + k];
  __syncthreads();
  needs_perm = false;
  if (RANKED) {
    int bad = 0;
    for (int k = threadIdx.x; k < n; k += NT) {
      const double e = c.eps[k];
      int rk = 0;
      for (int j = 0; j < n; ++j) {
        const double ej = c.eps[j];
        rk += (ej < e) || (ej == e && j < k);
      }
      rank[k] = rk;
      bad |= (rk < no) != (k < no);
    }
    needs_perm = __syncthreads_or(bad) != 0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double hi = -1.0e300, nlo = -1.0e300;  // max over occupied rows of d + r;  max over virtual rows of -(d - r)
  for (int i = warp; i < n; i += NT / 32) {
    const double* row = A + (size_t)i * ld;
    const bool oi = (RANKED ? rank[i] : i) < no;
    double s = 0.0;
    for (int j = lane; j < n; j += 32)
      if (j != i && ((RANKED ? rank[j] : j) < no) == oi) s += fabs(row[j]);
    s = warp_sum(s);
    if (oi) hi = fmax(hi, c.eps[i] + s);
    else nlo = fmax(nlo, s - c.eps[i]);
  }
  block_max2(hi, nlo, c.red);
  return -nlo - hi;
}

// Sort the basis by the ranks in c.occl: column k of C -> column rank[k], A permuted on both sides, c.eps alike.
// Uses the X buffer as the copy target.
XTB_CTX_FN void subspace_permute(Ctx& c) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  const int* rank = c.occl;
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int mu = t / ne, k = t - mu * ne;
    c.X[(size_t)mu * ld + (k < n ? rank[k] : k)] = c.C[(size_t)mu * ld + k];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int mu = t / ne, k = t - mu * ne;
    c.C[(size_t)mu * ld + k] = c.X[(size_t)mu * ld + k];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int i = t / ne, j = t - i * ne;
    c.X[(size_t)(i < n ? rank[i] : i) * ld + (j < n ? rank[j] : j)] = c.A[(size_t)i * ld + j];
  }
  for (int k = threadIdx.x; k < n; k += NT) c.srt[rank[k]] = c.eps[k];
  __syncthreads();
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int i = t / ne, j = t - i * ne;
    c.A[(size_t)i * ld + j] = c.X[(size_t)i * ld + j];
  }
  for (int k = threadIdx.x; k < n; k += NT) c.eps[k] = c.srt[k];
  __syncthreads();
}

// Fixed-point iteration of the Riccati equation in the occupied-first basis (A in the A buffer, its diagonal in c.eps,
// T = A(:, v) X with Lambda = Aoo + Aov X in its first `no` rows in the X buffer).  Returns true when max |R| <= tol.
template <int MODE>
XTB_CTX_FN bool subspace_riccati(Ctx& c, const xtb_scf_opts& o) {
  const int n = c.n, ld = c.ld, no = c.sub.no, nv = c.sub.nv, lds = c.sub.lds;
  const double* __restrict__ A = in_shared<MODE != 0>(c.A);
  double* __restrict__ T = in_shared<MODE == 1>(MODE == 1 ? c.X : c.sub.T);
  double* __restrict__ X = in_shared<MODE == 1>(c.sub.X);
  const double* __restrict__ dg = in_shared<true>(c.eps);
  if (!c.sub.xvalid) {
    for (int t = threadIdx.x; t < nv * lds; t += NT) X[t] = 0.0;
    c.sub.xvalid = true;
    c.sub.zvalid = false;
    __syncthreads();
  }
  double rprev = 1.0e300;
  for (int it = 0; it < o.subspace_maxiter; ++it) {
#ifdef XTB_PROFILE_PHASES
    const long long tq0 = clock64();
#endif
    // T[r][i] = sum_b A[no + b][r] X[b][i]  (A symmetric);  rows r < no: + Aoo -> Lambda
    gemm_small<2, 1>(n, no, nv, Operand{A + (size_t)no * ld, ld, 1}, Operand{X, lds, 1},
                     [&](int r, int i, double v) { T[r * lds + i] = (r < no) ? v + A[(size_t)r * ld + i] : v; });
    __syncthreads();
#ifdef XTB_PROFILE_PHASES
    const long long tq1 = clock64();
#endif
    // R = Avo + T1 - X Lambda;  X_new = X - R / (d_a - d_i) -> written over T1 (every element has one owner)
    double rmax = 0.0, xmax = 0.0;
    gemm_small<1, 1>(
        nv, no, no, Operand{X, 1, lds}, Operand{T, lds, 1},
        [&](int a, int i, double t3) {
          const double r = A[(size_t)(no + a) * ld + i] + T[(no + a) * lds + i] - t3;
          const double xn = X[a * lds + i] - r / (dg[no + a] - dg[i]);
          T[(no + a) * lds + i] = xn;
          rmax = fmax(rmax, fabs(r));
          xmax = fmax(xmax, fabs(xn));
        });
#ifdef XTB_PROFILE_PHASES
    const long long tq2 = clock64();
#endif
    block_max2(rmax, xmax, c.red);  // its barriers also order the X_new stores before the copy
    for (int t = threadIdx.x; t < nv * no; t += NT) {
      const int a = t / no, i = t - a * no;
      X[a * lds + i] = T[(no + a) * lds + i];
    }
    __syncthreads();
    ++c.sub.nric;
#ifdef XTB_PROFILE_PHASES
    c.tr1 += tq1 - tq0; c.tr2 += tq2 - tq1; c.tr3 += clock64() - tq2;
#endif
#ifdef XTB_DEBUG_SUBSPACE
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("    riccati it %d: max|R| %.3e max|X| %.3e\n", it, rmax, xmax);
#endif
    if (!(xmax < 1.0) || !(rmax < 1.0e300)) return false;  // large rotation / NaN: not the regime of this path
    // X already carries the update computed from this residual: its own residual is ~ (contraction rate) x rmax.  Stop when
    // that prediction (observed rate of the last step, at most 1/2, safety factor 2) meets the tolerance.
    const double rate = it > 0 ? fmin(0.5, 2.0 * rmax / rprev) : 1.0;
    if (rmax * rate <= o.subspace_tol) return true;
    if (it >= 2 && rmax > 2.0 * rprev) return false;  // diverging
    rprev = rmax;
  }
  return false;
}

// P = 2 Y Z Y^T with Y = C_o + C_v X and Z = (1 + X^T X)^-1; overwrites the A and X buffers and returns the buffer that
// holds P (row stride ld), or nullptr if the Newton iteration for Z failed (A is lost then: the caller rebuilds it).
//   A buffer: G, E (no x lds each) during the Newton iteration, then Y^T and W^T = (Y Z)^T ([k][mu], no x ld each);
//   X buffer: Z, Z' (ping-pong), then P.
// Shared-memory pointers are marked with in_shared<>, not XTB_ASSUME_SHARED: with assumptions on G / E / Z nvcc 12.9 treated
// the whole function as unreachable for the shared-memory variants (and dropped the `return true` of subspace_riccati with
// it) -- found in the PTX.
template <int MODE>
XTB_CTX_FN double* subspace_density(Ctx& c) {
  const int n = c.n, ld = c.ld, no = c.sub.no, nv = c.sub.nv, lds = c.sub.lds;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  const double* __restrict__ C = in_shared<CS>(c.C);
  const double* __restrict__ X = in_shared<CS>(c.sub.X);
  double* G = in_shared<AS>(c.A);
  double* E = G + (size_t)no * lds;
  double* Zc = in_shared<CS>(MODE == 1 ? c.X : c.sub.T);  // (T is free once the fixed point has converged)
  double* Zn = Zc + (size_t)no * lds;
  // G = 1 + X^T X
  gemm_small<1, 1>(no, no, nv, Operand{X, lds, 1}, Operand{X, lds, 1},
                   [&](int i, int j, double v) { G[i * lds + j] = (i == j) ? 1.0 + v : v; });
  if (c.sub.zvalid)
    for (int t = threadIdx.x; t < no * lds; t += NT) Zc[t] = c.sub.Zg[t];
  __syncthreads();
  bool ok = false;
  for (int it = 0; it < 12; ++it) {
    if (!c.sub.zvalid) {  // Z0 = 2 - G = 1 - X^T X: residual (X^T X)^2
      for (int t = threadIdx.x; t < no * no; t += NT) {
        const int i = t / no, j = t - i * no;
        Zc[i * lds + j] = (i == j ? 2.0 : 0.0) - G[i * lds + j];
      }
      c.sub.zvalid = true;
      __syncthreads();
    }
    double emax = 0.0;
    gemm_small<1, 1>(
        no, no, no, Operand{G, 1, lds}, Operand{Zc, lds, 1},
        [&](int i, int j, double v) {
          const double e = (i == j ? 1.0 : 0.0) - v;
          E[i * lds + j] = e;
          emax = fmax(emax, fabs(e));
        });
    emax = block_max(emax, c.red);  // (its barriers order the E stores before the next GEMM)
    if (emax <= 1e-12) { ok = true; break; }
    if (!(emax < 0.5)) {
      if (it > 0) break;       // the cold start does not contract either: give up
      c.sub.zvalid = false;    // warm start too far off: restart from Z0
      continue;
    }
    gemm_small<1, 1>(no, no, no, Operand{Zc, 1, lds}, Operand{E, lds, 1},
                     [&](int i, int j, double v) { Zn[i * lds + j] = Zc[i * lds + j] + v; });
    __syncthreads();
    double* t = Zc; Zc = Zn; Zn = t;
    ++c.sub.nnewt;
    if ((double)no * emax * emax <= 1e-13) { ok = true; break; }  // the new residual is E^2: no need to form it
  }
  if (!ok) {
    c.sub.zvalid = false;
    return nullptr;
  }
  for (int t = threadIdx.x; t < no * no; t += NT) {
    const int i = t / no, j = t - i * no;
    c.sub.Zg[i * lds + j] = Zc[i * lds + j];
  }
  // Yt[k][mu] = C[mu][k] + sum_b X[b][k] C[mu][no + b]
  double* Yt = G;
  double* Wt = G + (size_t)no * ld;
  gemm_small<1, 2>(no, n, nv, Operand{X, lds, 1}, Operand{C + no, 1, ld},
                   [&](int k, int mu, double v) { Yt[(size_t)k * ld + mu] = v + C[(size_t)mu * ld + k]; });
  __syncthreads();
  // Wt[k][mu] = sum_j Z[j][k] Yt[j][mu]
  gemm_small<1, 2>(no, n, no, Operand{Zc, lds, 1}, Operand{Yt, ld, 1}, [&](int k, int mu, double v) { Wt[(size_t)k * ld + mu] = v; });
  __syncthreads();
  // P[mu][nu] = 2 sum_k Wt[k][mu] Yt[k][nu]  -> X buffer (Z is dead)
  double* Pb = in_shared<CS>(c.X);
  gemm_small<2, 2, true>(n, n, no, Operand{Wt, ld, 1}, Operand{Yt, ld, 1}, [&](int mu, int nu, double v) {
    if (mu >= nu) {  // lower tiles computed, mirrored: P exactly symmetric
      Pb[(size_t)mu * ld + nu] = 2.0 * v;
      Pb[(size_t)nu * ld + mu] = 2.0 * v;
    }
  });
  __syncthreads();
  return Pb;
}

}  // namespace
