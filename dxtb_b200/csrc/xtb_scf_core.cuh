// Device code shared by the one-CTA-per-molecule SCF kernel (xtb_scf.cu) and the multi-CTA path for large systems
// (xtb_scf_large.cu): per-molecule context, fp64 tensor-core helpers, the in-CTA blocked Jacobi eigensolver, Fermi
// filling, potential, Anderson mixing and the result writer.  Everything has internal linkage (anonymous namespace).
#pragma once
#include "xtb_common.cuh"
#ifdef XTB_PROFILE_PHASES
#include <cstdio>
#endif

using namespace xtb;

namespace {

// xtb_scf.cu is compiled twice (dxtb_b200/build.py), both with 512 threads per CTA:
//   primary   (XTB_MINB = 1): 1 CTA/SM, up to 128 registers per thread (no spills; measured 8 % faster than 1024
//                             threads x 64 registers) -- the shared-memory variant, the global-memory variant for
//                             small buckets, and the C entry points;
//   secondary (-DXTB_SECONDARY -DXTB_MINB=2): the global-memory and hybrid variants at 2 CTAs/SM (64 registers), so that
//                             one molecule's latency-bound sub-problem phase overlaps the other's L2-bound tensor-core
//                             passes.  Exports only xtb_scf_launch_2cta().
#ifndef XTB_NT
#define XTB_NT 512
#endif
#ifndef XTB_MINB
#define XTB_MINB 1
#endif
constexpr int NT = XTB_NT;  // threads per CTA

// State of the occupied-subspace solve of the intermediate SCF map evaluations (see subspace_* below).
struct Subspace {
  bool eligible;   // closed shell with an integer number of doubly occupied orbitals and room for the scratch layout
  bool xvalid;     // X belongs to the current basis C (reset by every Jacobi sweep / permutation)
  bool layout;     // the basis has been put in occupied-first order at least once (cheap certificate first)
  bool zvalid;     // Zg holds (1 + X^T X)^-1 of an earlier X in the same basis (warm start of the Newton iteration)
  int no, nv, lds; // occupied / virtual orbitals, leading dimension of the no-column matrices (== 4 mod 16)
  double gapmin;   // certified HOMO-LUMO gap required for integer occupations
  double* X;       // [nv][lds] graph of the occupied subspace in the basis C (shared Jacobi scratch or workspace)
  double* T;       // scratch of the fixed point / Newton iterates: the X buffer, or shared memory behind the carve-up (MODE 0 / 2)
  double* Zg;      // [no][lds] workspace copy of Z
  int nfast, nric, nnewt;  // diagnostics: map evaluations on this path, fixed-point / Newton iterations
};

struct Ctx {
  int n, ne, ld, ns, na, np;
  Subspace sub;
  int o0, s0, a0;
  double *C, *A, *X;             // ne x ld matrices (shared or global)
  const double *S, *H0, *gam;    // global, n x n / ns x ns
  double *eps, *srt, *focc, *v, *vnew, *q, *n0, *eorb, *qsh, *vsh, *qat, *red, *cs;
  int *pp, *qq, *occl;
  double *jq, *jm, *jr;          // block-Jacobi scratch: accumulated rotations / sub-problem copies / (unused)
  bool smem;                     // matrices live in shared memory
  int ng;                        // Jacobi: number of 16x16 sub-problem copies = warps that solve sub-problems concurrently
  bool defer;                    // Jacobi: Q double buffered, V pass deferred into the next round's sub-problem phase
#ifdef XTB_PROFILE_PHASES
  long long tcert, tric, tden, tfock, tmull, tr1, tr2, tr3;
#endif
  long long tp1, tp2, tjac, tsub;  // XTB_PROFILE_PHASES: cycles in the sub-problem phase / rotation pass / whole eigensolver / subspace solve
  const int *ao_sh, *sh_atom, *at_sh0, *at_nsh, *sh_ao, *sh_l;
  const double* gam3;            // at_par base (stride XTB_ATPAR)
  double *xh, *fh;               // Anderson history [gen+1][n] (global)
  int status;
  int sweeps;                    // diagnostic: total Jacobi sweeps of this molecule
  double ef[2];                  // Fermi level per spin channel of the last fermi_fill (scf_response)
  bool spin_on[2];               // channel holds electrons
};

// Every function that takes the per-molecule context is force-inlined: one call that is not inlined puts the whole Ctx (60
// pointers) into local memory and every c.xyz access of the kernel becomes a stack load (seen in the SASS when fcn<> stopped
// being inlined at its three call sites).  The big ones (fcn, jacobi) have a single call site.
#define XTB_CTX_FN __device__ __forceinline__

// Address-space hint: lets the compiler emit LDS/STS (32-bit addressing) instead of generic LD/ST.
#define XTB_ASSUME_SHARED(ptr) __builtin_assume(__isShared(ptr))

// fp64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4],
// D[l/4][2*(l%4) + {0,1}].  SASS: DMMA.8x8x4.
XTB_DEV void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Out[i][j] = sum_{k<K} L[k*ld+i] * R[k*ld+j], i,j < ne (ne % 16 == 0), on the fp64 tensor cores:
// one warp per 16x16 output tile (2x2 DMMA tiles), K consumed 4 at a time.  With ld == 4 (mod 16) both
// fragment loads (4 consecutive k rows x 8 consecutive columns per half-warp) are bank-conflict free.
// WK: row k of L is scaled by wk[k] (weighted inner product over k).
template <bool LS, bool RS, bool WK = false>
__device__ void gemm_tn(int ne, int K, const double* __restrict__ L, const double* __restrict__ R, int ld, double* __restrict__ Out,
                        int ldo, int nout, const double* __restrict__ wk = nullptr) {
  if (LS) XTB_ASSUME_SHARED(L);
  if (RS) XTB_ASSUME_SHARED(R);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int nt = ne >> 4;
  for (int t = warp; t < nt * nt; t += NT / 32) {
    const int ti = t / nt, tj = t - ti * nt;
    const int i0 = ti << 4, j0 = tj << 4;
    double d[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) d[a][b][0] = d[a][b][1] = 0.0;
    for (int k0 = 0; k0 < K; k0 += 4) {
      const bool kv = k0 + tg < K;
      const double* lr = L + (size_t)(k0 + tg) * ld + i0 + g;
      const double* rr = R + (size_t)(k0 + tg) * ld + j0 + g;
      double a0 = kv ? lr[0] : 0.0, a1 = kv ? lr[8] : 0.0;
      const double b0 = kv ? rr[0] : 0.0, b1 = kv ? rr[8] : 0.0;
      if (WK) {
        const double wv = kv ? wk[k0 + tg] : 0.0;
        a0 *= wv;
        a1 *= wv;
      }
      dmma884(d[0][0][0], d[0][0][1], a0, b0);
      dmma884(d[0][1][0], d[0][1][1], a0, b1);
      dmma884(d[1][0][0], d[1][0][1], a1, b0);
      dmma884(d[1][1][0], d[1][1][1], a1, b1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int i = i0 + 8 * a + g, j = j0 + 8 * b + 2 * tg;
        if (i < nout) {
          if (j < nout) Out[(size_t)i * ldo + j] = d[a][b][0];
          if (j + 1 < nout) Out[(size_t)i * ldo + j + 1] = d[a][b][1];
        }
      }
  }
  __syncthreads();
}

// Symmetric product Out = L^T R for operands with L^T R == (L^T R)^T in exact arithmetic (A = C^T (F C)): only the tiles on and
// below the diagonal are computed (15 instead of 25 for ne = 80: one round of the 16 warps instead of two) and mirrored, so Out
// is EXACTLY symmetric; rows / columns >= n (padding) are written as exact zeros.
template <bool LS, bool RS, bool OS>
__device__ void gemm_tn_sym(int ne, int n, const double* __restrict__ L, const double* __restrict__ R, int ld, double* __restrict__ Out) {
  if (LS) XTB_ASSUME_SHARED(L);
  if (RS) XTB_ASSUME_SHARED(R);
  if (OS) XTB_ASSUME_SHARED(Out);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int nt = ne >> 4;
  for (int t = warp; t < nt * (nt + 1) / 2; t += NT / 32) {
    int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;  // tj <= ti
    const int i0 = ti << 4, j0 = tj << 4;
    double d[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) d[a][b][0] = d[a][b][1] = 0.0;
    for (int k0 = 0; k0 < ne; k0 += 4) {
      const double* lr = L + (size_t)(k0 + tg) * ld + i0 + g;
      const double* rr = R + (size_t)(k0 + tg) * ld + j0 + g;
      const double a0 = lr[0], a1 = lr[8];
      const double b0 = rr[0], b1 = rr[8];
      dmma884(d[0][0][0], d[0][0][1], a0, b0);
      dmma884(d[0][1][0], d[0][1][1], a0, b1);
      dmma884(d[1][0][0], d[1][0][1], a1, b0);
      dmma884(d[1][1][0], d[1][1][1], a1, b1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = i0 + 8 * a + g, j = j0 + 8 * b + 2 * tg + e;
          if (ti != tj || i >= j) {
            const double v = (i < n && j < n) ? d[a][b][e] : 0.0;
            Out[(size_t)i * ld + j] = v;
            Out[(size_t)j * ld + i] = v;
          }
        }
  }
  __syncthreads();
}

// Operand of gemm_small: element (k, m) at p[k * sk + m * sm] (either orientation of a row-major matrix, any offset).
struct Operand {
  const double* p;
  int sk, sm;
};

// Out(i, j) = sum_{k < K} L(k, i) R(k, j) for i < M, j < N on the fp64 tensor cores: one warp per (8 TM) x (8 TN) tile,
// out-of-range operand elements are zero, results are handed to st(i, j, value).  With leading dimensions == 4 (mod 8)
// row-major operands are bank-conflict free in both orientations.  Operands are (pointer, strides) so that a k step is a
// pointer increment and the row / column predicates are loop invariant: the first version took element lambdas and spent
// ~29 instructions per k step (index arithmetic rematerialised at the 128-register cap) for 2 DMMAs.  No trailing barrier.
// SYM (M == N, TM == TN, symmetric result): only the tiles on and below the diagonal are computed; st decides what to mirror.
template <int TM, int TN, bool SYM = false, class ST>
__device__ __forceinline__ void gemm_small(int M, int N, int K, Operand L, Operand R, ST st) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int mt = (M + 8 * TM - 1) / (8 * TM), nt = (N + 8 * TN - 1) / (8 * TN);
  const int stepl = 4 * L.sk, stepr = 4 * R.sk;
  const int kmain = K & ~3;
  const int ntile = SYM ? mt * (mt + 1) / 2 : mt * nt;
  for (int t = warp; t < ntile; t += NT / 32) {
    int ti, tj;
    if (SYM) {
      ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      while (ti * (ti + 1) / 2 > t) --ti;
      tj = t - ti * (ti + 1) / 2;
    } else {
      ti = t / nt;
      tj = t - ti * nt;
    }
    const int i0 = ti * 8 * TM, j0 = tj * 8 * TN;
    const double* pa[TM];
    const double* pb[TN];
    bool va[TM], vb[TN];
#pragma unroll
    for (int x = 0; x < TM; ++x) {
      const int i = i0 + 8 * x + g;
      va[x] = i < M;
      pa[x] = L.p + (va[x] ? i : 0) * L.sm + tg * L.sk;
    }
#pragma unroll
    for (int y = 0; y < TN; ++y) {
      const int j = j0 + 8 * y + g;
      vb[y] = j < N;
      pb[y] = R.p + (vb[y] ? j : 0) * R.sm + tg * R.sk;
    }
    double d[TM][TN][2];
#pragma unroll
    for (int x = 0; x < TM; ++x)
#pragma unroll
      for (int y = 0; y < TN; ++y) d[x][y][0] = d[x][y][1] = 0.0;
#pragma unroll 4
    for (int k0 = 0; k0 < kmain; k0 += 4) {  // unrolled: the operand loads of four k steps are in flight together
      double a[TM], b[TN];
#pragma unroll
      for (int x = 0; x < TM; ++x) {
        const double v = *pa[x];
        a[x] = va[x] ? v : 0.0;
        pa[x] += stepl;
      }
#pragma unroll
      for (int y = 0; y < TN; ++y) {
        const double v = *pb[y];
        b[y] = vb[y] ? v : 0.0;
        pb[y] += stepr;
      }
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TN; ++y) dmma884(d[x][y][0], d[x][y][1], a[x], b[y]);
    }
    if (kmain < K) {  // last, partial k step
      const bool kv = kmain + tg < K;
      double a[TM], b[TN];
#pragma unroll
      for (int x = 0; x < TM; ++x) a[x] = (kv && va[x]) ? *pa[x] : 0.0;
#pragma unroll
      for (int y = 0; y < TN; ++y) b[y] = (kv && vb[y]) ? *pb[y] : 0.0;
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TN; ++y) dmma884(d[x][y][0], d[x][y][1], a[x], b[y]);
    }
#pragma unroll
    for (int x = 0; x < TM; ++x)
#pragma unroll
      for (int y = 0; y < TN; ++y) {
        const int i = i0 + 8 * x + g, j = j0 + 8 * y + 2 * tg;
        if (i < M) {
          if (j < N) st(i, j, d[x][y][0]);
          if (j + 1 < N) st(i, j + 1, d[x][y][1]);
        }
      }
  }
}

// p, marked as a shared-memory pointer when SH (address-space conversion round trip: a definition, not an assumption --
// XTB_ASSUME_SHARED on the pointers of subspace_density made nvcc 12.9 drop the function as unreachable).
template <bool SH, class T>
__device__ __forceinline__ T* in_shared(T* p) {
  if (SH) return (T*)__cvta_shared_to_generic(__cvta_generic_to_shared(p));
  return p;
}

#ifndef XTB_DEFER_NBP_MAX
#define XTB_DEFER_NBP_MAX 16
#endif
constexpr int XTB_DEFER_NBP = XTB_DEFER_NBP_MAX;  // up to this many block pairs (nao <= 256) the accumulated rotations are double buffered
#ifndef XTB_NGRP
#define XTB_NGRP 16
#endif
constexpr int NGRP = XTB_NGRP;  // upper bound of sub-problems solved concurrently (one warp each with its own 16x16 copy; Ctx::ng)

#ifndef XTB_JACOBI_SMALL_SIN
#define XTB_JACOBI_SMALL_SIN 0.1  // sub-problems whose rotations all have |sin| below this take the small-angle shortcut
#endif
constexpr int JB = 8;        // Jacobi block size
constexpr int JB2 = 2 * JB;  // indices of a block pair
constexpr int MLD = 24;      // leading dimension of the 16x16 sub-problem copy (== 8 mod 16: conflict-free 2x2-block updates)
constexpr int QLD = 20;      // leading dimension of the accumulated 16x16 rotation (== 4 mod 16: conflict-free DMMA fragments)

// 1/sqrt(x) for normal x > 0: MUFU seed (rsqrt.approx.ftz.f64, ~2^-22) + two Newton steps, straight-line code (the
// library rsqrt() has a slow path behind a branch, which stops the scheduler from interleaving independent work).
XTB_DEV double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}

// Global index of local index l (0..15) of block pair (I, J).
XTB_DEV int bp_index(int I, int J, int l) { return (l < JB ? I * JB : J * JB - JB) + l; }

// BLOCKED two-sided Jacobi (block size 8) on the symmetric ne x ne matrix A (ne % 16 == 0), accumulating
// the transformation into the columns of V (ne rows).
//   Per block round (round-robin over the ne/8 blocks, ne/16 disjoint block pairs):
//   1. one warp per block pair copies its 16x16 sub-matrix, runs the scalar rotations of the pair on the copy
//      (warp-synchronous: no CTA or named barrier) and accumulates them into a 16x16 orthogonal Q;
//   2. all warps apply the Q's with fp64 tensor-core MMAs: every 16x16 block of A two-sided (Q_P^T B Q_R), V <- V Q.
//   A sweep = one "self" round (block pairs (0,1),(2,3),..: the 2 x 28 in-block index pairs) followed by the
//   nblk-1 round-robin rounds in which the 64 cross pairs of a block pair are rotated: every index pair once.
// Compared with rotating the full matrix after every scalar rotation round this moves A and V through
// shared memory ~10x instead of ~80x per sweep.  Returns the number of sweeps, or -sweeps if not converged.
template <bool AS, bool VS, bool DEFER = false>
XTB_CTX_FN int jacobi(Ctx& c, double* __restrict__ A, double* __restrict__ V, int nrow, double tol, int maxsweeps) {
  const int ne = c.ne, ld = c.ld;
  const int nblk = ne / JB, nbp = nblk / 2, ntile = ne / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = NT / 32;
  const int NG = c.ng;  // warps that solve sub-problems concurrently (each has its own 16x16 copy in shared memory)
  double* Qs0 = c.jq;  // [2][nbp][16][QLD]  accumulated rotations, double buffered by round parity
  double* Ms = c.jm;   // [NG][16][MLD]  sub-problem copy of a warp
  int* bij0 = c.pp;    // [2][nbp][2] blocks of the pairs of a round
  XTB_ASSUME_SHARED(bij0);
  XTB_ASSUME_SHARED(Qs0); XTB_ASSUME_SHARED(Ms);
  if (AS) XTB_ASSUME_SHARED(A);
  if (VS) XTB_ASSUME_SHARED(V);
  (void)nrow;
  const int nga = nbp < NG ? nbp : NG;  // warps busy with sub-problems; the others apply the previous round's Q to V
  constexpr bool defer = DEFER;  // compile time: the non-deferred instantiation keeps one Q buffer and no pool logic
  // V[:, idx] <- V[:, idx] Q (m8 n16 k16 tensor-core units) for the round whose rotations are in buffer `buf`
  auto v_unit = [&](int buf, int u) {
    const int g = lane >> 2, tg = lane & 3;
    const double* Qb = Qs0 + (size_t)buf * nbp * (JB2 * QLD);
    const int* bb = bij0 + buf * 2 * nbp;
    const int k = (int)__fdividef((float)u + 0.5f, (float)ntile), rt = u - k * ntile;
    const int I = bb[2 * k], J = bb[2 * k + 1];
    const double* Q = Qb + k * (JB2 * QLD);
    double* row = V + (size_t)(rt * 8 + g) * ld;
    double af[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) af[kk] = row[bp_index(I, J, 4 * kk + tg)];
    double d[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) dmma884(d[nt][0], d[nt][1], af[kk], Q[(4 * kk + tg) * QLD + 8 * nt + g]);
    }
    __syncwarp();  // in-place update of the 8 x 16 strip
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = bp_index(I, J, 8 * nt + 2 * tg);  // 2*tg and 2*tg+1 are in the same 8-block: contiguous
      row[col] = d[nt][0];
      row[col + 1] = d[nt][1];
    }
  };
  // deferred mode: the V units of the previous round are a pool (counter set `cs`: [0] next unit, [1] sub-problem warps
  // that finished this round); warps without a sub-problem work on it until the sub-problems are done, everybody drains
  // the rest after the A update
  int* ctr = bij0 + 4 * nbp;  // [2][2]
  auto v_pool = [&](int buf, int cs, bool until_done) {
    volatile int* vc = ctr + 2 * cs;
    for (;;) {
      if (until_done && vc[1] >= nga) break;
      int u = 0;
      if (lane == 0) u = atomicAdd(ctr + 2 * cs, 1);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= ntile * nbp) break;
      v_unit(buf, u);
    }
  };
  if (threadIdx.x < 4) ctr[threadIdx.x] = 0;  // made visible by the barriers of the first off-diagonal test
  int sweep = 0, rc = 0;
  bool pending = false;  // the V pass of the last finished round (buffer (rc - 1) & 1) is still to be done
  auto flush = [&]() {
    if (pending) {
      v_pool((rc - 1) & 1, rc & 1, false);
      pending = false;
      __syncthreads();
      if (threadIdx.x < 2) ctr[2 * (rc & 1) + threadIdx.x] = 0;  // the set is used again by the next round
      __syncthreads();
    }
  };
  for (;;) {
    double off = 0.0;
    for (int i = warp; i < ne; i += NW) {
      const double* row = A + (size_t)i * ld;
      for (int j = lane; j < ne; j += 32)
        if (i != j) off = fmax(off, fabs(row[j]));
    }
    off = block_max(off, c.red);
    if (off <= tol) { flush(); return sweep; }
    if (sweep >= maxsweeps) { flush(); return -sweep; }
    ++sweep;
    for (int r = -1; r < nblk - 1; ++r, ++rc) {
      const int buf = defer ? (rc & 1) : 0;
      double* Qs = Qs0 + (size_t)buf * nbp * (JB2 * QLD);
      int* bij = bij0 + buf * 2 * nbp;
#ifdef XTB_PROFILE_PHASES
      const long long tq0 = clock64();
#endif
      // ---- 1. sub-problems: ONE WARP per block pair, warp-synchronous (no CTA or named barriers) ------------------
      //   Every lane computes the rotation k = lane & 7 of the inner round (4 redundant copies, so there is no divergent
      //   branch and the parameters of any rotation are one shuffle away); the warp then updates the 16x16 copy M on the
      //   8x8 grid of 2x2 blocks (2 blocks per lane) and, one inner round behind and therefore off the dependent chain
      //   rot(t) -> M(t) -> rot(t+1), accumulates the rotations into Q (4 row items per lane, its own rotation).
      for (int w = warp; w < nbp && warp < NG; w += NG) {
        int I, J;
        if (r < 0) { I = 2 * w; J = 2 * w + 1; }
        else if (w == 0) { I = r; J = nblk - 1; }
        else {
          // (r + w) mod (nblk - 1), (r - w) mod (nblk - 1) with 0 <= r, w < nblk - 1: one conditional correction each
          I = r + w; if (I >= nblk - 1) I -= nblk - 1;
          J = r - w; if (J < 0) J += nblk - 1;
        }
        if (I > J) { const int t = I; I = J; J = t; }
        double* M = Ms + warp * (JB2 * MLD);
        double* Q = Qs + w * (JB2 * QLD);
        if (lane == 0) { bij[2 * w] = I; bij[2 * w + 1] = J; }
        const int nin = (r < 0) ? JB - 1 : JB;
        const bool self = r < 0;
        // index pair (p, q) of rotation k in inner round t
        auto pair_of = [&](int t, int k, int& p, int& q) {
          if (self) {
            // self round: the 28 in-block pairs of block I (k < 4) and of block J (k >= 4), round-robin on 8 indices
            const int kk = k & 3, h = (k >> 2) * JB;
            if (kk == 0) { p = t; q = JB - 1; }
            else { p = t + kk; if (p >= JB - 1) p -= JB - 1; q = t - kk; if (q < 0) q += JB - 1; }
            if (p > q) { const int x = p; p = q; q = x; }
            p += h; q += h;
          } else {
            p = k; q = JB + ((k + t) & (JB - 1));
          }
        };
        const int k = lane & 7;
        // Q items of this lane: rotation k, rows qrow + 4 j (lanes 0-7: row 0, 8-15: row 2, 16-23: row 1, 24-31: row 3,
        // so that the two rows of a half-warp are 2 apart: conflict free with QLD == 4 mod 16)
        const int qrow = 2 * ((lane >> 3) & 1) + (lane >> 4);
        double pc = 1.0, ps = 0.0;  // rotation of the previous inner round (for the Q update)
        int pp_ = 0, pq_ = 0;
        // Small-angle shortcut: all rotations of the sub-problem from the UN-UPDATED copy, two per lane (inner rounds
        // lane >> 3 and (lane >> 3) + 4), i.e. without the dependent chain rot(t) -> M(t) -> rot(t+1).  Rotations by small
        // angles commute to second order, so the sub-matrix is left with O(angle^2) off-diagonals exactly as after the
        // sequential rounds (quadratic regime); Q is still a product of exact Givens rotations.  Taken only if every angle
        // of the sub-problem is small (warp-uniform decision); otherwise the sequential rounds below.
        bool small_angles = false;
        if (XTB_JACOBI_SMALL_SIN > 0.0) {
          double cr[2], sr[2];
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            const int t = (lane >> 3) + 4 * sl;
            int p, q;
            pair_of(t < nin ? t : 0, k, p, q);
            // rotation inputs straight from A: the 16x16 copy is only made if the sequential rounds are needed
            const size_t gp = (size_t)bp_index(I, J, p), gq = (size_t)bp_index(I, J, q);
            const double app = A[gp * ld + gp], aqq = A[gq * ld + gq], apq = A[gp * ld + gq];
            const double d = aqq - app;
            const double x = fma(d, d, 4.0 * apq * apq);
            const double ir = rsqrt_nr(fmax(x, 1e-280));
            const double c2 = fma(0.5 * fabs(d), ir, 0.5);
            const double ic = rsqrt_nr(c2);
            const bool rot = x > 1e-280 && t < nin;
            cr[sl] = rot ? c2 * ic : 1.0;
            sr[sl] = rot ? copysign(apq * ir, d * apq) * ic : 0.0;
          }
          small_angles = __all_sync(0xffffffffu, fabs(sr[0]) <= XTB_JACOBI_SMALL_SIN && fabs(sr[1]) <= XTB_JACOBI_SMALL_SIN);
          if (small_angles) {
            for (int t = 0; t < nin; ++t) {
              const int src = k + 8 * (t & 3);
              const double cv = __shfl_sync(0xffffffffu, (t >> 2) ? cr[1] : cr[0], src);
              const double sv = __shfl_sync(0xffffffffu, (t >> 2) ? sr[1] : sr[0], src);
              int p, q;
              pair_of(t, k, p, q);
              if (t == 0) {
                // first layer on the identity: columns p, q of J_0 written directly (the 8 rotations cover all 16 columns)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int qi = qrow + 4 * j;
                  Q[qi * QLD + p] = (qi == p) ? cv : (qi == q) ? -sv : 0.0;
                  Q[qi * QLD + q] = (qi == p) ? sv : (qi == q) ? cv : 0.0;
                }
              } else {
                double vp[4], vq[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { vp[j] = Q[(qrow + 4 * j) * QLD + p]; vq[j] = Q[(qrow + 4 * j) * QLD + q]; }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  Q[(qrow + 4 * j) * QLD + p] = cv * vp[j] - sv * vq[j];
                  Q[(qrow + 4 * j) * QLD + q] = sv * vp[j] + cv * vq[j];
                }
              }
              __syncwarp();
            }
          }
        }
        if (!small_angles) {
          for (int e = lane; e < JB2 * JB2; e += 32) {
            const int rr = e >> 4, cc = e & 15;
            M[rr * MLD + cc] = A[(size_t)bp_index(I, J, rr) * ld + bp_index(I, J, cc)];
            Q[rr * QLD + cc] = (rr == cc) ? 1.0 : 0.0;
          }
          __syncwarp();
        }
        for (int t = 0; t < nin && !small_angles; ++t) {
          int p, q;
          pair_of(t, k, p, q);
          const double app = M[p * MLD + p], aqq = M[q * MLD + q], apq = M[p * MLD + q];
          // M items of this lane: (kp, kq) = (lane >> 3, lane & 7) and (kp + 4, kq); loads issued before the rotation
          // parameters are known
          int p1[2], q1[2], p2, q2;
          pair_of(t, lane >> 3, p1[0], q1[0]);
          pair_of(t, (lane >> 3) + 4, p1[1], q1[1]);
          p2 = p; q2 = q;
          double a00[2], a01[2], a10[2], a11[2];
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            a00[it] = M[p1[it] * MLD + p2]; a01[it] = M[p1[it] * MLD + q2];
            a10[it] = M[q1[it] * MLD + p2]; a11[it] = M[q1[it] * MLD + q2];
          }
          // Small-angle Jacobi rotation from the double-angle identities (two dependent rsqrt instead of three
          // divisions/square roots):
          //   r = sqrt(d^2 + 4 apq^2), cos 2t = |d| / r, sin 2t = sign(d) 2 apq / r,
          //   c^2 = (1 + cos 2t) / 2,  c = c^2 rsqrt(c^2),  s = sin 2t / (2 c).
          // s has no cancellation for small angles; c^2 + s^2 = 1 holds to a few ulp.
          // branch-free (straight-line code lets the scheduler interleave the Q update below with this dependent chain)
          const double d = aqq - app;
          const double x = fma(d, d, 4.0 * apq * apq);
          const double ir = rsqrt_nr(fmax(x, 1e-280));
          const double c2 = fma(0.5 * fabs(d), ir, 0.5);
          const double ic = rsqrt_nr(c2);
          const bool rot = x > 1e-280;
          const double cs_ = rot ? c2 * ic : 1.0;
          const double sn = rot ? copysign(apq * ir, d * apq) * ic : 0.0;
          // Q <- Q J(t-1): rotation k of the previous inner round on 4 rows (independent of the chain above; all loads
          // before the stores, the compiler cannot prove that the rows do not alias)
          if (t > 0) {
            double vp[4], vq[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { vp[j] = Q[(qrow + 4 * j) * QLD + pp_]; vq[j] = Q[(qrow + 4 * j) * QLD + pq_]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              Q[(qrow + 4 * j) * QLD + pp_] = pc * vp[j] - ps * vq[j];
              Q[(qrow + 4 * j) * QLD + pq_] = ps * vp[j] + pc * vq[j];
            }
          }
          // M <- J^T M J on the 8x8 grid of 2x2 blocks; column rotation kq = own rotation, row rotation kp by shuffle
          __syncwarp();  // every lane has read its rotation inputs and block values before any lane stores
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int kp = (lane >> 3) + 4 * it;
            const double c1 = __shfl_sync(0xffffffffu, cs_, kp), s1 = __shfl_sync(0xffffffffu, sn, kp);
            const double x00 = c1 * a00[it] - s1 * a10[it], x01 = c1 * a01[it] - s1 * a11[it];
            const double x10 = s1 * a00[it] + c1 * a10[it], x11 = s1 * a01[it] + c1 * a11[it];
            double y00 = cs_ * x00 - sn * x01, y01 = sn * x00 + cs_ * x01;
            double y10 = cs_ * x10 - sn * x11, y11 = sn * x10 + cs_ * x11;
            if (kp == k) { y01 = 0.0; y10 = 0.0; }
            M[p1[it] * MLD + p2] = y00; M[p1[it] * MLD + q2] = y01;
            M[q1[it] * MLD + p2] = y10; M[q1[it] * MLD + q2] = y11;
          }
          pc = cs_; ps = sn; pp_ = p; pq_ = q;
          __syncwarp();
        }
        // rotations of the last inner round of the sequential path
#pragma unroll
        for (int j = 0; j < 4 && !small_angles; ++j) {
          const int qi = qrow + 4 * j;
          const double vp = Q[qi * QLD + pp_], vq = Q[qi * QLD + pq_];
          Q[qi * QLD + pp_] = pc * vp - ps * vq;
          Q[qi * QLD + pq_] = ps * vp + pc * vq;
        }
      }
      // meanwhile the warps without a sub-problem apply the PREVIOUS round's rotations to V (off the critical path:
      // only A feeds the next sub-problems)
      if (defer) {
        if (warp < nga) {
          __syncwarp();
          if (lane == 0) atomicAdd(ctr + 2 * (rc & 1) + 1, 1);
        } else if (pending) {
          v_pool((rc - 1) & 1, rc & 1, true);
        }
      }
      __syncthreads();
#ifdef XTB_PROFILE_PHASES
      const long long tq1 = clock64();
#endif
      // ---- 2. apply the rotations with fp64 tensor-core MMAs, one phase ---------------------------------------
      //   A: every 16x16 block (pair P, pair R), P >= R, is transformed on BOTH sides in registers, B' = Q_P^T (B Q_R),
      //      and mirrored into (R, P): A is read (half) and written once per block round (the intermediate T = B Q_R is re-laid out from the accumulator
      //      to the B-operand fragment layout with warp shuffles);
      //   V: V[:, idx] <- V[:, idx] Q  (m8 n16 k16 units).
      {
        const int g = lane >> 2, tg = lane & 3;
        // A is symmetric: only the blocks P >= R are computed, P > R blocks are mirrored into (R, P)
        const int nfused = nbp * (nbp + 1) / 2, nunit = nfused + (defer ? 0 : ntile * nbp);
        for (int u = warp; u < nunit; u += NW) {
          if (u >= nfused) {
            v_unit(0, u - nfused);  // single Q buffer: V in the same phase
          } else {
            int P = (int)((sqrtf(8.0f * (float)u + 1.0f) - 1.0f) * 0.5f);
            while ((P + 1) * (P + 2) / 2 <= u) ++P;
            while (P * (P + 1) / 2 > u) --P;
            const int R = u - P * (P + 1) / 2;
            const int IP = bij[2 * P], JP = bij[2 * P + 1], IR = bij[2 * R], JR = bij[2 * R + 1];
            const double* QP = Qs + P * (JB2 * QLD);
            const double* QR = Qs + R * (JB2 * QLD);
            // T = B Q_R
            double t[2][2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) t[mt][nt][0] = t[mt][nt][1] = 0.0;
            const double* r0 = A + (size_t)bp_index(IP, JP, g) * ld;
            const double* r1 = A + (size_t)bp_index(IP, JP, 8 + g) * ld;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int col = bp_index(IR, JR, 4 * kk + tg);
              const double a0 = r0[col], a1 = r1[col];
              const double b0 = QR[(4 * kk + tg) * QLD + g], b1 = QR[(4 * kk + tg) * QLD + 8 + g];
              dmma884(t[0][0][0], t[0][0][1], a0, b0);
              dmma884(t[0][1][0], t[0][1][1], a0, b1);
              dmma884(t[1][0][0], t[1][0][1], a1, b0);
              dmma884(t[1][1][0], t[1][1][1], a1, b1);
            }
            // B' = Q_P^T T:  A-operand [m][k] = Q_P[k][m];  B-operand [k][n] = T[k][n] fetched from the accumulator
            // layout (row g', cols 2 tg', 2 tg' + 1) of lane 4 g' + tg' with g' = 4 (kk & 1) + tg, tg' = g >> 1.
            double d[2][2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) d[mt][nt][0] = d[mt][nt][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int src = ((4 * (kk & 1) + tg) << 2) | (g >> 1);
              double bfr[2];
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                const double v0 = __shfl_sync(0xffffffffu, t[kk >> 1][nt][0], src);
                const double v1 = __shfl_sync(0xffffffffu, t[kk >> 1][nt][1], src);
                bfr[nt] = (g & 1) ? v1 : v0;
              }
              const double a0 = QP[(4 * kk + tg) * QLD + g], a1 = QP[(4 * kk + tg) * QLD + 8 + g];
              dmma884(d[0][0][0], d[0][0][1], a0, bfr[0]);
              dmma884(d[0][1][0], d[0][1][1], a0, bfr[1]);
              dmma884(d[1][0][0], d[1][0][1], a1, bfr[0]);
              dmma884(d[1][1][0], d[1][1][1], a1, bfr[1]);
            }
            __syncwarp();  // every lane's reads of this block precede the MMAs; made explicit for the in-place update
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              double* row = A + (size_t)bp_index(IP, JP, 8 * mt + g) * ld;
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                const int col = bp_index(IR, JR, 8 * nt + 2 * tg);
                row[col] = d[mt][nt][0];
                row[col + 1] = d[mt][nt][1];
                if (P != R) {  // mirror: A[col][row]
                  A[(size_t)col * ld + bp_index(IP, JP, 8 * mt + g)] = d[mt][nt][0];
                  A[(size_t)(col + 1) * ld + bp_index(IP, JP, 8 * mt + g)] = d[mt][nt][1];
                }
              }
            }
          }
        }
      }
      if (defer) {
        if (pending) v_pool((rc - 1) & 1, rc & 1, false);               // drain what the idle warps left
        if (threadIdx.x < 2) ctr[2 * ((rc + 1) & 1) + threadIdx.x] = 0;  // counters of the next round (unused in this one)
        pending = true;
      }
      __syncthreads();
#ifdef XTB_PROFILE_PHASES
      c.tp1 += tq1 - tq0;
      c.tp2 += clock64() - tq1;
#endif
    }
  }
}

// Fermi smearing, both spin channels in lockstep (wavefunction/filling.py:201-366).
// focc[k] = f_alpha + f_beta; returns G = kT sum ln(f^f (1-f)^(1-f)) (scf/base.py:586-594).
XTB_CTX_FN double fermi_fill(Ctx& c, double nel_a, double nel_b, const xtb_scf_opts& o) {
  const int n = c.n;
  // rank sort of the eigenvalues (ascending; ties broken by index)
  for (int k = threadIdx.x; k < n; k += NT) {
    const double e = c.eps[k];
    int rk = 0;
    for (int j = 0; j < n; ++j) {
      const double ej = c.eps[j];
      rk += (ej < e) || (ej == e && j < k);
    }
    c.srt[rk] = e;
  }
  __syncthreads();
  double nel[2] = {nel_a, nel_b}, ef[2], ef_used[2], hom[2];
  bool ne_[2];
  c.spin_on[0] = c.spin_on[1] = false;  // aufbau / empty: no Fermi-function derivative
  c.ef[0] = c.ef[1] = 0.0;
  if (fabs(nel_a + nel_b) < kEps) {
    for (int k = threadIdx.x; k < n; k += NT) c.focc[k] = 0.0;
    __syncthreads();
    return 0.0;
  }
  if (o.kt < 3e-7) {  // aufbau (scf/base.py:889)
    for (int k = threadIdx.x; k < n; k += NT) {
      const double e = c.eps[k];
      int rk = 0;
      for (int j = 0; j < n; ++j) rk += (c.eps[j] < e) || (c.eps[j] == e && j < k);
      c.focc[k] = (rk < nel_a ? 1.0 : 0.0) + (rk < nel_b ? 1.0 : 0.0);
    }
    __syncthreads();
    return 0.0;
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    int h = (int)ceil(nel[s] - 5.0e-15) - 1;
    if (h < 0) h = 0;
    if (h > n - 1) h = 0;
    const int l = (n - 1 <= h) ? h : h + 1;
    hom[s] = (double)h;
    ne_[s] = nel[s] != 0.0;
    ef[s] = ne_[s] ? 0.5 * (c.srt[h] + c.srt[l]) : 0.0;
    ef_used[s] = ef[s];
  }
  bool conv = false;
  for (int it = 0; it < o.fermi_maxiter; ++it) {
    double sf[2] = {0.0, 0.0}, sd[2] = {0.0, 0.0};
    for (int k = threadIdx.x; k < n; k += NT) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const double e = ne_[s] ? c.eps[k] : 0.0;
        const double ex = (e - ef[s]) / o.kt;
        if (ex < 50.0) {
          const double et = exp(ex);
          sf[s] += 1.0 / (et + 1.0);
          sd[s] += et / (o.kt * (et + 1.0) * (et + 1.0));
        } else {
          sd[s] += kEps;
        }
      }
    }
    const double f0 = block_sum(sf[0], c.red), f1 = block_sum(sf[1], c.red);
    const double d0 = block_sum(sd[0], c.red), d1 = block_sum(sd[1], c.red);
    const double r0 = hom[0] - f0 + 1.0, r1 = hom[1] - f1 + 1.0;
    ef_used[0] = ef[0];
    ef_used[1] = ef[1];
    ef[0] += r0 / d0;
    ef[1] += r1 / d1;
    if (fabs(r0) <= o.fermi_thresh && fabs(r1) <= o.fermi_thresh) { conv = true; break; }
  }
  if (!conv) c.status |= XTB_STATUS_FERMI_FAILED;
  c.ef[0] = ef_used[0]; c.ef[1] = ef_used[1];
  c.spin_on[0] = ne_[0]; c.spin_on[1] = ne_[1];
  double g = 0.0;
  for (int k = threadIdx.x; k < n; k += NT) {
    double ft = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      double f = 0.0;
      if (ne_[s]) {
        const double ex = (c.eps[k] - ef_used[s]) / o.kt;
        if (ex < 50.0) f = 1.0 / (exp(ex) + 1.0);
      }
      ft += f;
      const double o1 = fmax(f, kEps), o2 = fmax(1.0 - f, kEps);
      g += o1 * log(o1) + o2 * log(o2);
    }
    c.focc[k] = ft;
  }
  g = block_sum(g, c.red);
  __syncthreads();
  return g * o.kt;
}

// q (orbital charges) -> shell/atom charges -> potential vout (scf/base.py:702-727, interactions/base.py:134-181)
XTB_CTX_FN void potential(Ctx& c, const double* __restrict__ q, double* __restrict__ vout) {
  for (int a = threadIdx.x; a < c.na; a += NT) {
    const int s0 = c.at_sh0[a], nsa = c.at_nsh[a];
    double qa = 0.0;
    for (int k = 0; k < nsa; ++k) {
      double qs = 0.0;
      const int sh = s0 + k;
      const int mu0 = c.sh_ao[sh], nmu = 2 * c.sh_l[sh] + 1;  // AOs of a shell are contiguous
      for (int mu = mu0; mu < mu0 + nmu; ++mu) qs += q[mu];
      c.qsh[sh] = qs;
      qa += qs;
    }
    c.qat[a] = qa;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = w; k < c.ns; k += NT / 32) {
    const double* gr = c.gam + (size_t)k * c.ns;
    double acc = 0.0;
    for (int l = lane; l < c.ns; l += 32) acc += gr[l] * c.qsh[l];
    acc = warp_sum(acc);
    if (lane == 0) {
      const int a = c.sh_atom[k];
      const double qa = c.qat[a];
      c.vsh[k] = acc + c.gam3[(size_t)a * XTB_ATPAR + XTB_AT_GAM3] * qa * qa;
    }
  }
  __syncthreads();
  for (int mu = threadIdx.x; mu < c.n; mu += NT) vout[mu] = c.vsh[c.ao_sh[mu]];
  __syncthreads();
}

// In-CTA right-looking Cholesky S = L L^T (lower triangle, in the A buffer), X = L^{-1} by forward substitution
// (thread per column, X buffer), C = X^T.  Returns false if S is not positive definite.
template <int MODE>
XTB_CTX_FN bool cholesky_start_basis(Ctx& c) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  double* A = c.A; double* X = c.X; double* C = c.C; double* lk = c.srt;
  if (MODE != 0) XTB_ASSUME_SHARED(A);
  if (MODE == 1) { XTB_ASSUME_SHARED(X); XTB_ASSUME_SHARED(C); }
  XTB_ASSUME_SHARED(lk);
  for (int t = threadIdx.x; t < ne * ld; t += NT) {
    const int i = t / ld, j = t - i * ld;
    A[t] = (i < n && j < n) ? c.S[(size_t)i * n + j] : 0.0;
    X[t] = 0.0;
    C[t] = 0.0;
  }
  __syncthreads();
  bool ok = true;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = 0; k < n; ++k) {
    const double akk = A[(size_t)k * ld + k];
    if (!(akk > 0.0)) ok = false;
    const double inv = (akk > 0.0) ? rsqrt(akk) : 0.0;
    // column k of L, also staged contiguously in lk[]
    for (int i = k + threadIdx.x; i < n; i += NT) lk[i] = (i == k) ? akk * inv : A[(size_t)i * ld + k] * inv;
    __syncthreads();
    for (int i = k + threadIdx.x; i < n; i += NT) A[(size_t)i * ld + k] = lk[i];
    // trailing update of the lower triangle: A[i][j] -= L[i][k] L[j][k], k < j <= i
    for (int i = k + 1 + warp; i < n; i += NT / 32) {
      const double li = lk[i];
      double* row = A + (size_t)i * ld;
      for (int j = k + 1 + lane; j <= i; j += 32) row[j] = fma(-li, lk[j], row[j]);
    }
    __syncthreads();
  }
  // X = L^{-1}: thread j owns column j (consecutive threads -> consecutive words: no bank conflicts)
  for (int j = threadIdx.x; j < n; j += NT) {
    for (int i = j; i < n; ++i) {
      const double* lrow = A + (size_t)i * ld;
      double s0 = (i == j) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = j;
      for (; k + 3 < i; k += 4) {
        s0 = fma(-lrow[k], X[(size_t)k * ld + j], s0);
        s1 = fma(-lrow[k + 1], X[(size_t)(k + 1) * ld + j], s1);
        s2 = fma(-lrow[k + 2], X[(size_t)(k + 2) * ld + j], s2);
        s3 = fma(-lrow[k + 3], X[(size_t)(k + 3) * ld + j], s3);
      }
      for (; k < i; ++k) s0 = fma(-lrow[k], X[(size_t)k * ld + j], s0);
      X[(size_t)i * ld + j] = ((s0 + s1) + (s2 + s3)) / lrow[i];
    }
  }
  __syncthreads();
  // C0 = X^T  (upper triangular): C0^T S C0 = L^{-1} L L^T L^{-T} = I
  for (int t = threadIdx.x; t < n * n; t += NT) {
    const int a = t / n, b = t - a * n;
    C[(size_t)a * ld + b] = (b >= a) ? X[(size_t)b * ld + a] : 0.0;
  }
  __syncthreads();
  return ok;
}

// 5x5 (or smaller) solve with partial pivoting, thread 0 only
__device__ void small_solve(int n, double (*a)[5], double* b, double* x) {
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = fabs(a[k][k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(a[i][k]) > best) { best = fabs(a[i][k]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < n; ++j) { const double t = a[k][j]; a[k][j] = a[piv][j]; a[piv][j] = t; }
      const double t = b[k]; b[k] = b[piv]; b[piv] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      const double f = a[i][k] / a[k][k];
      for (int j = k + 1; j < n; ++j) a[i][j] -= f * a[k][j];
      b[i] -= f * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= a[i][j] * x[j];
    x[i] = s / a[i][i];
  }
}

struct Mixer {
  int step, head;  // head: physical slot of logical history index 0
};

// Anderson / simple mixing (mixer/anderson.py:163-317, mixer/simple.py:88-151).  x_old is in c.v, x_new in
// c.vnew; the mixed vector is written to c.v.  Returns true if converged (mixer/base.py:229-256).
XTB_CTX_FN bool mix(Ctx& c, Mixer& mx, const xtb_scf_opts& o, double* sm_theta) {
  const int n = c.n, G1 = o.generations + 1;
  auto slot = [&](int i) { return (mx.head + i) % G1; };
  double* f0 = c.fh + (size_t)slot(0) * n;
  double* x0 = c.xh + (size_t)slot(0) * n;
  if (mx.step == 0)
    for (int k = threadIdx.x; k < n; k += NT) x0[k] = c.v[k];
  mx.step += 1;
  double l2 = 0.0, li = 0.0;
  for (int k = threadIdx.x; k < n; k += NT) {
    const double d = c.vnew[k] - c.v[k];
    f0[k] = d;
    l2 += d * d;
    li = fmax(li, fabs(d));
  }
  l2 = block_sum(l2, c.red);
  li = block_max(li, c.red);
  __syncthreads();
  const bool conv = (sqrt(l2) < o.x_atol) && (li < o.x_atol_max);

  const bool anderson = o.mixer == 0 && (mx.step > o.generations || (mx.step > 1 && !o.soft_start));
  if (o.mixer == 1) {
    for (int k = threadIdx.x; k < n; k += NT) c.v[k] = c.v[k] + o.damp * f0[k];
  } else if (anderson) {
    int nh = mx.step - 1;
    if (nh > o.generations) nh = o.generations;
    if (nh > 5) nh = 5;
    // a_ij = <dF_i, dF_j>, b_i = <dF_i, F0>, dF_i = F0 - F_i   (20 dot products, one warp each)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ndot = nh * (nh + 1) / 2 + nh;
    for (int d = w; d < ndot; d += NT / 32) {
      int i, j;
      if (d < nh) { i = d; j = -1; }
      else {
        int r = d - nh; i = 0;
        while (r > i) { r -= i + 1; ++i; }
        j = r;
      }
      const double* fi = c.fh + (size_t)slot(i + 1) * n;
      const double* fj = (j >= 0) ? c.fh + (size_t)slot(j + 1) * n : nullptr;
      double acc = 0.0;
      for (int k = lane; k < n; k += 32) {
        const double di = f0[k] - fi[k];
        const double dj = (j >= 0) ? f0[k] - fj[k] : f0[k];
        acc = fma(di, dj, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        if (j < 0) sm_theta[25 + i] = acc;
        else { sm_theta[i * 5 + j] = acc; sm_theta[j * 5 + i] = acc; }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a[5][5], bb[5], th[5];
      for (int i = 0; i < nh; ++i) {
        for (int j = 0; j < nh; ++j) a[i][j] = sm_theta[i * 5 + j];
        a[i][i] *= 1.0 + o.diag_offset * o.diag_offset;
        bb[i] = sm_theta[25 + i];
      }
      small_solve(nh, a, bb, th);
      for (int i = 0; i < nh; ++i) sm_theta[30 + i] = th[i];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += NT) {
      double xb = x0[k], fb = f0[k];
      for (int i = 0; i < nh; ++i) {
        const double th = sm_theta[30 + i];
        xb += th * (c.xh[(size_t)slot(i + 1) * n + k] - x0[k]);
        fb -= th * (f0[k] - c.fh[(size_t)slot(i + 1) * n + k]);
      }
      c.v[k] = xb + o.damp * fb;
    }
  } else {
    for (int k = threadIdx.x; k < n; k += NT) c.v[k] = x0[k] + o.damp_init * f0[k];
  }
  __syncthreads();
  // roll histories and store the mixed vector as x_hist[0]
  mx.head = (mx.head + G1 - 1) % G1;
  double* xn = c.xh + (size_t)slot(0) * n;
  for (int k = threadIdx.x; k < n; k += NT) xn[k] = c.v[k];
  __syncthreads();
  return conv;
}


// ---------------------------------------------------------------------------------------------------------------------
// First-order response of a NOT fully converged SCF state in the nuclear gradient.
//
// The reference's forces are autograd through the unrolled SCF (calculators/types/autograd.py:80-201): the exact derivative
// of E = E[v_in(R), R], v_in being the un-mixed potential that enters the final solve (scf/base.py:497-501) and
// v_out = V(q_out) the potential of its charges.  With Omega = sum f eps + G stationary in orbitals and occupations,
//     dE/dR = [Hellmann-Feynman + Pulay terms] + (v_out - v_in) . dq_out/dR ;
// the last term is first order in the SCF residual (1e-6..1e-5 Eh/bohr at dxtb's default thresholds) and absent from the
// converged-SCF formula (analytical.py:63-222).  dq_out/dR is expanded with the response of the converged fixed point
// (coupled-perturbed equations in adjoint form):
//     y = (1 - chi K)^-1 chi dv,  u = dv + K y,  chi w = -diag(Z_w S),  Z_w = C [(C^T A_w C) o G] C^T,
//     A_w = -1/2 S o (w (+) w),  G_ij = (f_j - f_i)/(e_j - e_i)  (diagonal: Fermi-function derivative per spin channel with
//     the Fermi-level shift projected out),  K = dV/dq = gamma + 2 Gamma q_A,
// and the gradient kernels are fed P + Z_u, W + ZW_u, v_out + K y and y_sh (ES2 cross term).  4 in-CTA GEMMs per
// application of chi, Anderson-accelerated (same mixer code), <= 12 applications.  Same arithmetic as
// oracle/gfn1_oracle.py:_scf_response.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kResponseMaxIter = 12;
constexpr double kResponseTol = 1e-8;  // max-norm of the potential-space residual (the accepted iterate is ~4x better); force error of that size

// w = K y: shell/atom sums of y, then gamma y_sh + 2 Gamma q_A y_A (linearised `potential`); leaves y_sh in c.qsh
XTB_CTX_FN void potential_lin(Ctx& c, const double* __restrict__ y, const double* __restrict__ qat_final, double* __restrict__ wout) {
  for (int a = threadIdx.x; a < c.na; a += NT) {
    const int s0 = c.at_sh0[a], nsa = c.at_nsh[a];
    double ya = 0.0;
    for (int k = 0; k < nsa; ++k) {
      double ys = 0.0;
      const int sh = s0 + k;
      const int mu0 = c.sh_ao[sh], nmu = 2 * c.sh_l[sh] + 1;
      for (int mu = mu0; mu < mu0 + nmu; ++mu) ys += y[mu];
      c.qsh[sh] = ys;
      ya += ys;
    }
    c.qat[a] = ya;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = w; k < c.ns; k += NT / 32) {
    const double* gr = c.gam + (size_t)k * c.ns;
    double acc = 0.0;
    for (int l = lane; l < c.ns; l += 32) acc += gr[l] * c.qsh[l];
    acc = warp_sum(acc);
    if (lane == 0) {
      const int a = c.sh_atom[k];
      c.vsh[k] = acc + 2.0 * c.gam3[(size_t)a * XTB_ATPAR + XTB_AT_GAM3] * qat_final[a] * c.qat[a];
    }
  }
  __syncthreads();
  for (int mu = threadIdx.x; mu < c.n; mu += NT) wout[mu] = c.vsh[c.ao_sh[mu]];
  __syncthreads();
}

// Buffers of the response solver: C (eigenvectors, untouched), X = S C (set up once; the operand of both contractions of an
// application, so it lives next to C -- in shared memory in the shared-memory variant), A (work), and one matrix in the global
// workspace (G2: temporary of the two back-transformations at the end).  With A_w = -1/2 (diag(w) S + S diag(w)):
//     At = C^T A_w C = -1/2 (M + M^T),   M = C^T diag(w) (S C)                       -- ONE weighted GEMM
//     chi w = -diag(Z S) = -rowdot(C Zt, S C)                                         -- ONE GEMM with the row dots in its epilogue
// i.e. 2 GEMMs per application of chi instead of 4 GEMMs + 2 transposes; the full Z / ZW matrices (2 more GEMMs each) are
// only formed once at the end.  (Until round 2b S C and the product C Zt lived in the global workspace and X held C^T: an
// application cost ~100 k cycles, most of it L2 latency of the GEMM operand S C and of the C Zt round trip.)
struct RespBuf {
  double* SC;  // unused since round 2b (S C is kept in the X buffer); the workspace region stays reserved
  double* G2;  // GEMM temporary / row-dot partial sums (global workspace)
};

// A <- Zt (WMAT = false) or ZWt (WMAT = true) of the perturbation w, in the eigenvector basis
template <int MODE, bool WMAT>
XTB_CTX_FN void response_zt(Ctx& c, const RespBuf& rb, const double* __restrict__ w, const double* __restrict__ fp0,
                            const double* __restrict__ fp1, bool have_m = false) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  // M[i][j] = sum_{mu < n} w_mu C[mu][i] SC[mu][j]   (X = S C); have_m: the caller has put M into the A buffer already
  if (!have_m) gemm_tn<CS, CS, true>(ne, n, c.C, c.X, ld, c.A, ld, ne, w);
  // Fermi-level shift per spin channel: abar_s = sum_k f'_s(k) At_kk / sum_k f'_s(k),  At_kk = -M_kk
  double s0 = 0.0, s1 = 0.0, n0 = 0.0, n1 = 0.0;
  for (int k = threadIdx.x; k < n; k += NT) {
    const double d = -c.A[(size_t)k * ld + k];
    s0 += fp0[k] * d; n0 += fp0[k];
    s1 += fp1[k] * d; n1 += fp1[k];
  }
  s0 = block_sum(s0, c.red); n0 = block_sum(n0, c.red);
  s1 = block_sum(s1, c.red); n1 = block_sum(n1, c.red);
  const double ab0 = fabs(n0) > kTiny ? s0 / n0 : 0.0, ab1 = fabs(n1) > kTiny ? s1 / n1 : 0.0;
  __syncthreads();
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int i = t / ne, j = t - i * ne;
    if (i > j) continue;
    double val = 0.0;
    if (j < n) {
      const double ei = c.eps[i], ej = c.eps[j], fi = c.focc[i], fj = c.focc[j];
      const double fpi = fp0[i] + fp1[i], fpj = fp0[j] + fp1[j];
      if (i == j) {
        const double d = -c.A[(size_t)i * ld + i];
        const double zd = (d - ab0) * fp0[i] + (d - ab1) * fp1[i];
        val = WMAT ? d * fi + zd * ei : zd;
      } else {
        const double a = -0.5 * (c.A[(size_t)i * ld + j] + c.A[(size_t)j * ld + i]);
        const double de = ej - ei;
        const bool close = fabs(de) <= 1e-9;
        double gf;
        if (WMAT) gf = close ? 0.5 * ((fi + ei * fpi) + (fj + ej * fpj)) : (fj * ej - fi * ei) / de;
        else gf = close ? 0.5 * (fpi + fpj) : (fj - fi) / de;
        val = a * gf;
      }
    }
    c.A[(size_t)i * ld + j] = val;
    c.A[(size_t)j * ld + i] = val;
  }
  __syncthreads();
  (void)AS;
}

// out[mu] = add[mu] + chi w = add[mu] - sum_q (C Zt)[mu][q] SC[mu][q]: the product C Zt is never stored -- every 16 x 16 tile
// is multiplied with its tile of S C in registers and reduced along q (two shuffles over the four lanes of a row); the
// per-tile-column partial sums are added in a fixed order (deterministic: get_forces and autograd.grad stay bit-equal).
template <int MODE>
XTB_CTX_FN void response_charges(Ctx& c, const RespBuf& rb, const double* __restrict__ w, const double* __restrict__ fp0,
                                 const double* __restrict__ fp1, const double* __restrict__ add, double* __restrict__ out) {
  const int n = c.n, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  response_zt<MODE, false>(c, rb, w, fp0, fp1);
  const double* __restrict__ C = in_shared<CS>(c.C);
  const double* __restrict__ Zt = in_shared<AS>(c.A);
  const double* __restrict__ SC = in_shared<CS>(c.X);
  double* part = MODE == 1 ? in_shared<true>(c.jq) : rb.G2;  // [tile column][mu]; MODE 1: ne^2 / 16 <= 20 ne doubles of Jacobi scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int nt = (n + 15) >> 4, kmain = n & ~3;
  for (int t = warp; t < nt * nt; t += NT / 32) {
    const int ti = t / nt, tj = t - ti * nt;
    const int i0 = ti << 4, j0 = tj << 4;
    // (C Zt)[mu][q] = sum_p C[mu][p] Zt[p][q]: A operand C[mu][p] (row major in mu), B operand Zt[p][q]
    const bool va0 = i0 + g < n, va1 = i0 + 8 + g < n, vb0 = j0 + g < n, vb1 = j0 + 8 + g < n;
    const double* pa0 = C + (size_t)(va0 ? i0 + g : 0) * ld + tg;
    const double* pa1 = C + (size_t)(va1 ? i0 + 8 + g : 0) * ld + tg;
    const double* pb0 = Zt + (size_t)tg * ld + (vb0 ? j0 + g : 0);
    const double* pb1 = Zt + (size_t)tg * ld + (vb1 ? j0 + 8 + g : 0);
    double d[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll 4
    for (int k0 = 0; k0 < kmain; k0 += 4) {
      const double x0 = *pa0, x1 = *pa1, y0 = *pb0, y1 = *pb1;
      const double a0 = va0 ? x0 : 0.0, a1 = va1 ? x1 : 0.0, b0 = vb0 ? y0 : 0.0, b1 = vb1 ? y1 : 0.0;
      pa0 += 4; pa1 += 4; pb0 += 4 * ld; pb1 += 4 * ld;
      dmma884(d[0][0][0], d[0][0][1], a0, b0);
      dmma884(d[0][1][0], d[0][1][1], a0, b1);
      dmma884(d[1][0][0], d[1][0][1], a1, b0);
      dmma884(d[1][1][0], d[1][1][1], a1, b1);
    }
    if (kmain < n) {
      const bool kv = kmain + tg < n;
      const double a0 = (kv && va0) ? *pa0 : 0.0, a1 = (kv && va1) ? *pa1 : 0.0;
      const double b0 = (kv && vb0) ? *pb0 : 0.0, b1 = (kv && vb1) ? *pb1 : 0.0;
      dmma884(d[0][0][0], d[0][0][1], a0, b0);
      dmma884(d[0][1][0], d[0][1][1], a0, b1);
      dmma884(d[1][0][0], d[1][0][1], a1, b0);
      dmma884(d[1][1][0], d[1][1][1], a1, b1);
    }
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int mu = i0 + 8 * x + g;
      double sacc = 0.0;
      if (mu < n) {
#pragma unroll
        for (int y = 0; y < 2; ++y) {
          const int q = j0 + 8 * y + 2 * tg;
          const double* sr = SC + (size_t)mu * ld + q;
          if (q < n) sacc = fma(d[x][y][0], sr[0], sacc);
          if (q + 1 < n) sacc = fma(d[x][y][1], sr[1], sacc);
        }
      }
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
      if (tg == 0 && mu < n) part[(size_t)tj * c.ne + mu] = sacc;
    }
  }
  __syncthreads();
  for (int mu = threadIdx.x; mu < n; mu += NT) {
    double acc = 0.0;
    for (int tj = 0; tj < nt; ++tj) acc += part[(size_t)tj * c.ne + mu];
    out[mu] = (add ? add[mu] : 0.0) - acc;
  }
  __syncthreads();
}

// A <- Z_w (or ZW_w) in the AO basis, Z = C Zt C^T.  Called twice at the end with the SAME perturbation (Z, then ZW): the
// two only differ in the occupation factors applied to M, so M is computed once (first call: saved to the global G2; second
// call: reloaded) and the X buffer, whose S C is no longer needed once M exists, is the temporary of the back-transformation.
template <int MODE, bool WMAT>
XTB_CTX_FN void response_density(Ctx& c, const RespBuf& rb, const double* __restrict__ w, const double* __restrict__ fp0,
                                 const double* __restrict__ fp1) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  if (!WMAT) {
    gemm_tn<CS, CS, true>(ne, n, c.C, c.X, ld, c.A, ld, ne, w);
    for (int t = threadIdx.x; t < ne * ld; t += NT) rb.G2[t] = c.A[t];
  } else {
    __syncthreads();
    for (int t = threadIdx.x; t < ne * ld; t += NT) c.A[t] = rb.G2[t];
    __syncthreads();
  }
  response_zt<MODE, WMAT>(c, rb, w, fp0, fp1, true);
  const double* C = in_shared<CS>(c.C);
  double* A = in_shared<AS>(c.A);
  double* T = in_shared<CS>(c.X);
  // T[q][mu] = sum_p Zt[p][q] C[mu][p]
  gemm_small<2, 2>(n, n, n, Operand{A, ld, 1}, Operand{C, 1, ld}, [&](int q, int mu, double v) { T[(size_t)q * ld + mu] = v; });
  __syncthreads();
  // Z[mu][nu] = sum_q T[q][mu] C[nu][q]
  gemm_small<2, 2>(n, n, n, Operand{T, ld, 1}, Operand{C, 1, ld}, [&](int mu, int nu, double v) { A[(size_t)mu * ld + nu] = v; });
  __syncthreads();
}

// Runs after the final solve and after P, W were written: adds Z_u, ZW_u to Pm, Wm and writes v_out + K y / y_sh.
// Vector scratch (all free after emit_results): dv -> eorb, z0 -> q, w -> v, g(w) -> vnew, y -> n0, f'_0 -> srt, f'_1 -> cs.
// The coupled-perturbed equations are iterated in potential space, w = K y:  w = g(w) = K (z0 + chi w), Anderson-accelerated
// from an empty history (re-using the SCF's own Anderson history as search directions was measured SLOWER: 12 instead of 8
// applications of chi to 1e-9, its differences carry the non-linearity of the early SCF iterations).
template <int MODE>
XTB_CTX_FN void scf_response(Ctx& c, const xtb_scf_opts& o, const RespBuf& rb, const double* __restrict__ v_out_g,
                             const double* __restrict__ qat_final, double* __restrict__ Pm, double* __restrict__ Wm,
                             double* __restrict__ v_grad, double* __restrict__ y_sh, double* sm_theta) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  double* dv = c.eorb; double* z0 = c.q; double* y = c.n0; double* fp0 = c.srt; double* fp1 = c.cs;
  double dvmax = 0.0;
  for (int k = threadIdx.x; k < n; k += NT) {
    dv[k] = c.vnew[k] - c.v[k];
    dvmax = fmax(dvmax, fabs(dv[k]));
    double f[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      f[s] = 0.0;
      if (c.spin_on[s] && o.kt >= 3e-7) {
        const double ex = (c.eps[k] - c.ef[s]) / o.kt;
        if (ex < 50.0) f[s] = 1.0 / (exp(ex) + 1.0);
      }
    }
    fp0[k] = -(f[0] * (1.0 - f[0])) / o.kt;
    fp1[k] = -(f[1] * (1.0 - f[1])) / o.kt;
  }
  dvmax = block_max(dvmax, c.red);
  __syncthreads();
  if (dvmax < kResponseTol) {  // converged to the last digit: nothing to add
    for (int k = threadIdx.x; k < n; k += NT) v_grad[k] = v_out_g[k];
    for (int k = threadIdx.x; k < c.ns; k += NT) y_sh[k] = 0.0;
    return;
  }
#ifdef XTB_PROFILE_PHASES
  const long long tr0 = clock64();
  int nit = 0;
  c.tr1 = c.tr2 = c.tr3 = 0;  // (re-used: the subspace split was printed before)
#endif
  // set-up: X = S C
  for (int t = threadIdx.x; t < ne * ld; t += NT) {
    const int i = t / ld, j = t - i * ld;
    c.A[t] = (i < n && j < n) ? c.S[(size_t)i * n + j] : 0.0;
  }
  __syncthreads();
  gemm_tn<AS, CS>(ne, ne, c.A, c.C, ld, c.X, ld, ne);  // SC[mu][q] = sum_nu S[nu][mu] C[nu][q]

  response_charges<MODE>(c, rb, dv, fp0, fp1, nullptr, z0);
  potential_lin(c, z0, qat_final, c.v);  // w0 = K z0
  xtb_scf_opts o2 = o;
  o2.mixer = 0; o2.soft_start = 0; o2.damp = 0.5; o2.damp_init = 0.5; o2.diag_offset = 0.01;
  o2.x_atol = 0.0; o2.x_atol_max = 0.0;  // the stop test is the max-norm below
  Mixer mx;
  mx.step = 0; mx.head = 0;
  for (int it = 0; it < kResponseMaxIter; ++it) {
#ifdef XTB_PROFILE_PHASES
    const long long ta0 = clock64();
#endif
    response_charges<MODE>(c, rb, c.v, fp0, fp1, z0, y);  // y = z0 + chi w
#ifdef XTB_PROFILE_PHASES
    const long long ta1 = clock64();
#endif
    potential_lin(c, y, qat_final, c.vnew);               // g(w) = K y; leaves y_sh in c.qsh
#ifdef XTB_PROFILE_PHASES
    const long long ta2 = clock64();
    c.tr1 += ta1 - ta0; c.tr2 += ta2 - ta1;
#endif
    double res = 0.0;
    for (int k = threadIdx.x; k < n; k += NT) res = fmax(res, fabs(c.vnew[k] - c.v[k]));
    res = block_max(res, c.red);
    __syncthreads();
#ifdef XTB_PROFILE_PHASES
    ++nit;
#ifdef XTB_DEBUG_SUBSPACE
    if (threadIdx.x == 0 && blockIdx.x == 0) printf("  response it %d: residual %.3e\n", it, res);
#endif
#endif
    if (res < kResponseTol) {
      for (int k = threadIdx.x; k < n; k += NT) c.v[k] = c.vnew[k];
      __syncthreads();
      break;
    }
#ifdef XTB_PROFILE_PHASES
    const long long ta3 = clock64();
#endif
    mix(c, mx, o2, sm_theta);
#ifdef XTB_PROFILE_PHASES
    c.tr3 += clock64() - ta3;
#endif
  }
#ifdef XTB_PROFILE_PHASES
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("  response: %d applications in the loop, %lld cycles so far (in the loop: chi w %lld, K y %lld, mixer %lld)\n", nit, clock64() - tr0, c.tr1,
           c.tr2, c.tr3);
#endif
  for (int k = threadIdx.x; k < c.ns; k += NT) y_sh[k] = c.qsh[k];
  for (int k = threadIdx.x; k < n; k += NT) {
    v_grad[k] = v_out_g[k] + c.v[k];
    c.vnew[k] = dv[k] + c.v[k];  // u = dv + K y
  }
  __syncthreads();
  response_density<MODE, false>(c, rb, c.vnew, fp0, fp1);
  for (int t = threadIdx.x; t < n * n; t += NT) {
    const int i = t / n, j = t - i * n;
    Pm[t] += c.A[(size_t)i * ld + j];
  }
  __syncthreads();
  response_density<MODE, true>(c, rb, c.vnew, fp0, fp1);
  for (int t = threadIdx.x; t < n * n; t += NT) {
    const int i = t / n, j = t - i * n;
    Wm[t] += c.A[(size_t)i * ld + j];
  }
  __syncthreads();
#ifdef XTB_PROFILE_PHASES
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("  response total %lld cycles\n", clock64() - tr0);
#endif
}

// Per-molecule results of the converged SCF: charges, potential, orbital energies / occupations, atom-resolved energies.
XTB_CTX_FN void emit_results(Ctx& c, const xtb_batch& b, int m, double g, int iters, double* __restrict__ q_orb, double* __restrict__ q_sh,
                             double* __restrict__ q_at, double* __restrict__ v_orb, double* __restrict__ e_atom,
                             double* __restrict__ fenergy, double* __restrict__ emo, double* __restrict__ occ,
                             int32_t* __restrict__ iterations, int32_t* __restrict__ status) {
  const int n = c.n;
  for (int mu = threadIdx.x; mu < n; mu += NT) {
    q_orb[c.o0 + mu] = c.q[mu];
    v_orb[c.o0 + mu] = c.vnew[mu];  // potential of the final charges (scf/base.py:468)
    emo[c.o0 + mu] = c.eps[mu];
    occ[c.o0 + mu] = c.focc[mu];
  }
  for (int k = threadIdx.x; k < c.ns; k += NT) q_sh[c.s0 + k] = c.qsh[k];
  // atom-resolved energies (scf/base.py:514-534; interactions/base.py:305-360; secondorder.py:374-382; thirdorder.py:246-276)
  for (int a = threadIdx.x; a < c.na; a += NT) {
    double e = 0.0;
    const int sh0 = c.at_sh0[a], nsa = c.at_nsh[a];
    for (int k = 0; k < nsa; ++k) {
      const int mu0 = c.sh_ao[sh0 + k], nmu = 2 * c.sh_l[sh0 + k] + 1;
      for (int mu = mu0; mu < mu0 + nmu; ++mu) e += c.eorb[mu];
    }
    const double qa = c.qat[a];
    const double g3 = c.gam3[(size_t)a * XTB_ATPAR + XTB_AT_GAM3];
    for (int k = 0; k < nsa; ++k) {
      const double ves2 = c.vsh[sh0 + k] - g3 * qa * qa;  // gamma.q_sh part of the shell potential
      e += 0.5 * c.qsh[sh0 + k] * ves2;
    }
    e += g3 * qa * qa * qa / 3.0;
    e += g / (double)c.na;
    e_atom[c.a0 + a] = e;
    q_at[c.a0 + a] = qa;
  }
  if (threadIdx.x == 0) {
    fenergy[m] = g;
    iterations[m] = iters;
    status[m] = c.status | (c.sweeps << 8);  // bits 8..: total Jacobi sweeps (diagnostic)
  }
}

}  // namespace
