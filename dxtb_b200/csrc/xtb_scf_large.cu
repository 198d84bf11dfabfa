// Multi-CTA self-consistent field path for ONE large molecule (BASELINE config 4: sh3, nao 3104).
//
// The one-CTA-per-molecule kernel (xtb_scf.cu) stops scaling when a single system should use the whole device, and its
// per-orbital vectors no longer fit in shared memory beyond ~700 AOs.  Here every stage of the SCF map is a grid-wide
// kernel over the padded ne x ne matrices (ne = n rounded up to 128, row-major, ld = ne) in a torch-owned workspace:
//   Fock build -> A = C^T F C (two fp64 tensor-core GEMMs, 128x128 tiles) -> two-level block Jacobi (outer blocks of 32:
//   one CTA per block pair solves the 64x64 sub-problem with the in-CTA blocked Jacobi of xtb_scf_core.cuh, then all
//   CTAs apply the accumulated 64x64 rotations Q to A (two-sided, symmetric half + mirror) and C with 64^3 DMMA GEMMs)
//   -> Fermi filling / potential / Anderson mixing (single-CTA kernels reusing the device functions of the batch path)
//   -> density P = Y^T Y (GEMM) -> Mulliken populations.
// The host drives the loop (one stream synchronisation per Jacobi sweep and per SCF iteration to read the convergence
// scalars); same arithmetic conventions as the batch kernel (eigenvector basis of the previous iteration, final solve
// with the un-mixed potential, scf/base.py:497-501).  The S-orthonormal start basis is U s^{-1/2} from the Jacobi
// eigendecomposition S = U s U^T (any S-orthonormal basis gives the same SCF trajectory up to round-off).
#include "xtb_scf_core.cuh"

#include <cstdio>
#include <mutex>
#include <vector>

namespace {

constexpr int OB = 32;       // outer Jacobi block
constexpr int OP = 2 * OB;   // indices of an outer block pair = sub-problem dimension
constexpr int SLD = OP + 4;  // shared-memory leading dimension (== 4 mod 16: conflict-free DMMA fragments)
constexpr int LPAD = 128;    // ne is a multiple of the GEMM tile
constexpr int GT = 128;      // GEMM CTA tile
constexpr int GK = 16;       // GEMM K chunk
constexpr int GLD = GT + 4;  // GEMM shared-memory leading dimension

struct LargeState {  // device-resident scalars of one large-molecule SCF
  double g;          // electronic free energy of the last solve
  double off;        // max |off-diagonal| (Jacobi sweep test), compared as raw bits (non-negative doubles)
  int mixer_step, mixer_head;
  int status, sweeps, nocc, converged;
  double ef[2];      // Fermi levels of the last solve (SCF response)
  double abar[2];    // Fermi-level shifts of the current response density
  int spin_on[2];
  int sub_bad, pad;  // occupied-subspace solve: the occupied class is not the first `no` columns
  // maxima of the subspace kernels, stored biased by kSubBias and compared as raw bits: max over occupied rows of d + r,
  // max over virtual rows of r - d (Gershgorin), max |Riccati residual|, max |X|, max |1 - G Z|
  double sub_hi, sub_nlo, sub_rmax, sub_xmax, sub_emax;
};

struct Layout {  // workspace offsets in doubles
  size_t C, A, X, Q, vec, hist, bij, state, total;
  // occupied-subspace solve (xtb_scf_subspace.cuh, grid-wide): nh = ne/2 rounded up to the GEMM tile bounds the occupied count
  size_t sX, sXt, sT, sT3, sG, sE, sZ, sCt, sYt, sWt;
  int ne, nbp, nh;
};

__host__ __device__ inline Layout layout(int n, int ns, int na, int gen) {
  Layout l;
  l.ne = (n + LPAD - 1) / LPAD * LPAD;
  l.nbp = l.ne / OP;
  const size_t m = (size_t)l.ne * l.ne;
  size_t p = 0;
  l.C = p; p += m;
  l.A = p; p += m;
  l.X = p; p += m;
  l.Q = p; p += (size_t)2 * l.nbp * OP * OP;  // accumulated rotations, double buffered by round parity
  l.vec = p; p += (size_t)8 * (n + 2) + 2 * ns + na + 32 + 36 + (n + 4) / 2 + 2 + (n + 2);  // + f'_beta of the SCF response
  p += p & 1;
  l.hist = p; p += (size_t)2 * (gen + 1) * n;
  p += p & 1;
  l.bij = p; p += l.nbp + 1;
  p += p & 1;
  l.state = p; p += sizeof(LargeState) / 8 + 1;
  p += p & 1;
  l.nh = (l.ne / 2 + LPAD - 1) / LPAD * LPAD;
  const size_t mh = (size_t)l.ne * l.nh, hh = (size_t)l.nh * l.nh;
  l.sX = p; p += mh;    // X  [nv][no]  (ld nh)
  l.sXt = p; p += mh;   // X^T [no][nv] (ld ne)
  l.sT = p; p += mh;    // A(:, v) X, Lambda in its first no rows (ld nh); Newton: Z E
  l.sT3 = p; p += mh;   // X Lambda (ld nh)
  l.sG = p; p += hh;
  l.sE = p; p += hh;
  l.sZ = p; p += hh;
  l.sCt = p; p += m;    // C^T
  l.sYt = p; p += mh;   // Y^T [no][n] (ld ne)
  l.sWt = p; p += mh;   // (Y Z)^T
  l.total = p;
  return l;
}

// Context of molecule m with all per-orbital vectors in the workspace (generic address space).
__device__ void large_ctx(Ctx& c, const xtb_batch& b, int m, double* work, int gen, const double* S, const double* H0,
                          const double* gamma) {
  c.o0 = b.ao_off[m]; c.s0 = b.sh_off[m]; c.a0 = b.at_off[m];
  c.n = b.ao_off[m + 1] - c.o0;
  c.ns = b.sh_off[m + 1] - c.s0;
  c.na = b.at_off[m + 1] - c.a0;
  const Layout l = layout(c.n, c.ns, c.na, gen);
  c.ne = l.ne; c.ld = l.ne; c.np = l.ne / 2;
  c.status = 0; c.sweeps = 0; c.smem = false;
  c.C = work + l.C; c.A = work + l.A; c.X = work + l.X;
  const int nmx = c.n + 2;
  double* p = work + l.vec;
  c.eps = p; p += nmx; c.srt = p; p += nmx; c.focc = p; p += nmx; c.v = p; p += nmx; c.vnew = p; p += nmx;
  c.q = p; p += nmx; c.n0 = p; p += nmx; c.eorb = p; p += nmx;
  c.qsh = p; p += c.ns; c.vsh = p; p += c.ns; c.qat = p; p += c.na;
  c.red = p; p += 32;
  c.cs = p; p += 36;  // Anderson small system (sm_theta of the batch kernel)
  c.occl = (int*)p; p += (c.n + 4) / 2 + 2;
  c.pp = c.qq = nullptr; c.jq = c.jm = nullptr;
  c.jr = p;  // f'_beta of the SCF response (the batch kernel keeps it in its Jacobi scratch c.cs)
  c.xh = work + l.hist;
  c.fh = c.xh + (size_t)(gen + 1) * c.n;
  c.S = S + b.mat_off[m];
  c.H0 = H0 + b.mat_off[m];
  c.gam = gamma + b.gam_off[m];
  c.ao_sh = b.ao_sh + c.o0;
  c.sh_atom = b.sh_atom + c.s0;
  c.sh_ao = b.sh_ao + c.s0;
  c.sh_l = b.sh_l + c.s0;
  c.at_sh0 = b.at_sh0 + c.a0;
  c.at_nsh = b.at_nsh + c.a0;
  c.gam3 = b.at_par + (size_t)c.a0 * XTB_ATPAR;
}

// ---- elementwise / reduction kernels over the padded matrices ----------------------------------------------------

// dst (ne x ne) = src (n x n) zero padded; pad diagonal = dpad
__global__ void kl_load(double* __restrict__ dst, const double* __restrict__ src, int n, int ne, double dpad) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    dst[t] = (i < n && j < n) ? src[(size_t)i * n + j] : (i == j ? dpad : 0.0);
  }
}

__global__ void kl_pack(double* __restrict__ dst, const double* __restrict__ src, int n, int ne) {
  const size_t tot = (size_t)n * n;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
    dst[t] = src[(size_t)i * ne + j];
  }
}

// F = H0 - 1/2 S (v_i + v_j), zero padded (scf/base.py:651-675)
__global__ void kl_fock(double* __restrict__ A, const double* __restrict__ H0, const double* __restrict__ S, const double* __restrict__ v,
                        int n, int ne) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    double f = 0.0;
    if (i < n && j < n) {
      const size_t ij = (size_t)i * n + j;
      f = H0[ij] - 0.5 * S[ij] * (v[i] + v[j]);
    }
    A[t] = f;
  }
}

// symmetrise (round-off of the two GEMMs) and keep the pad rows / columns exactly zero
__global__ void kl_symmetrize(double* __restrict__ A, int n, int ne) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    if (i < j) {
      const double a = (i < n && j < n) ? 0.5 * (A[(size_t)i * ne + j] + A[(size_t)j * ne + i]) : 0.0;
      A[(size_t)i * ne + j] = a;
      A[(size_t)j * ne + i] = a;
    } else if (i == j && i >= n) {
      A[t] = 0.0;
    }
  }
}

__global__ void kl_offmax(const double* __restrict__ A, int ne, LargeState* st) {
  __shared__ double red[32];
  const size_t tot = (size_t)ne * ne;
  double off = 0.0;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    if (i != j) off = fmax(off, fabs(A[t]));
  }
  off = block_max(off, red);
  if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(&st->off), (unsigned long long)__double_as_longlong(off));
}

// start basis: C[:, k] = U[:, k] / sqrt(s_k) (columns of the eigenvectors of S scaled by the inverse root of the eigenvalue)
__global__ void kl_scale_cols(double* __restrict__ C, const double* __restrict__ A, int n, int ne, LargeState* st) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), k = (int)(t - (size_t)i * ne);
    if (k < n && i < n) {
      const double s = A[(size_t)k * ne + k];
      if (!(s > 0.0)) {
        if (i == 0) atomicOr(&st->status, XTB_STATUS_S_NOT_POSDEF);
        C[t] = 0.0;
      } else {
        C[t] *= rsqrt(s);
      }
    } else {
      C[t] = 0.0;
    }
  }
}

// M = 3/2 I - 1/2 sym(G) (Newton-Schulz factor of the re-orthonormalisation)
__global__ void kl_ns_factor(double* __restrict__ A, int ne) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    if (i <= j) {
      const double g = 0.5 * (A[(size_t)i * ne + j] + A[(size_t)j * ne + i]);
      const double m = (i == j ? 1.5 : 0.0) - 0.5 * g;
      A[(size_t)i * ne + j] = m;
      A[(size_t)j * ne + i] = m;
    }
  }
}

// dst = src^T (ne x ne, 32 x 32 tiles)
__global__ void kl_transpose(double* __restrict__ dst, const double* __restrict__ src, int ne) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = src[(size_t)(i0 + r) * ne + j0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) dst[(size_t)(j0 + r) * ne + i0 + tx] = tile[tx][r];
}

// Y[kk][i] = w_k C[i][k] for the occupied orbitals k = occl[kk] (rows nocc..kocc-1 zero); which = 0: w = sqrt(f) -> X;
// which = 1: w = f eps -> X and w = 1 -> A (energy-weighted density, W = Y1^T Y2)
__global__ void kl_build_y(double* __restrict__ X, double* __restrict__ A2, const double* __restrict__ C, const double* __restrict__ focc,
                           const double* __restrict__ eps, const int* __restrict__ occl, int n, int ne, int nocc, int kocc, int which) {
  __shared__ double tile[32][33];
  // 32 x 32 transposing tiles: block (bx over i, by over kk)
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, kk = k0 + tx;
    double v = 0.0;
    if (kk < nocc && i < n) v = C[(size_t)i * ne + occl[kk]];
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int kk = k0 + r, i = i0 + tx;
    if (kk < kocc) {
      const double cv = tile[tx][r];
      double w = 0.0;
      if (kk < nocc) {
        const int k = occl[kk];
        w = which == 0 ? sqrt(focc[k]) : focc[k] * eps[k];
      }
      X[(size_t)kk * ne + i] = w * cv;
      if (which == 1) A2[(size_t)kk * ne + i] = cv;
    }
  }
}

// Mulliken populations and orbital-resolved H0 energies: one warp per row of P (wavefunction/mulliken.py)
__global__ void kl_mulliken(const double* __restrict__ P, const double* __restrict__ S, const double* __restrict__ H0,
                            const double* __restrict__ n0, double* __restrict__ q, double* __restrict__ eorb, int n, int ne) {
  const int lane = threadIdx.x & 31;
  const int mu = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (mu >= n) return;
  const double* pr = P + (size_t)mu * ne;
  const double* sr = S + (size_t)mu * n;
  const double* hr = H0 + (size_t)mu * n;
  double pop = 0.0, e = 0.0;
  for (int nu = lane; nu < n; nu += 32) {
    const double p = pr[nu];
    pop = fma(p, sr[nu], pop);
    e = fma(p, hr[nu], e);
  }
  pop = warp_sum(pop);
  e = warp_sum(e);
  if (lane == 0) {
    q[mu] = n0[mu] - pop;
    eorb[mu] = e;
  }
}

// ---- fp64 tensor-core GEMM  Out[i][j] = sum_k L[k][i] R[k][j]  (all ld = ne, ne % 128 == 0, K % 16 == 0) ------------
// 128 x 128 tile per CTA, 16 warps with 32 x 32 warp tiles (4 x 4 DMMA tiles), K chunks of 16 double buffered through
// shared memory with register staging of the next chunk.
__global__ void __launch_bounds__(NT, 1)
kl_gemm_tn(const double* __restrict__ L, const double* __restrict__ R, double* __restrict__ Out, int ld, int K) {
  extern __shared__ double gsm[];
  double* Ls = gsm;                  // [2][GK][GLD]
  double* Rs = gsm + 2 * GK * GLD;   // [2][GK][GLD]
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int wi = (warp >> 2) * 32, wj = (warp & 3) * 32;
  // staging: thread -> (row = t / 32, 4 consecutive columns)
  const int srow = threadIdx.x >> 5, scol = (threadIdx.x & 31) * 4;
  double d[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) d[a][b][0] = d[a][b][1] = 0.0;
  double2 l0, l1, r0, r1;
  auto gload = [&](int k0) {
    const double* lp = L + (size_t)(k0 + srow) * ld + i0 + scol;
    const double* rp = R + (size_t)(k0 + srow) * ld + j0 + scol;
    l0 = *reinterpret_cast<const double2*>(lp); l1 = *reinterpret_cast<const double2*>(lp + 2);
    r0 = *reinterpret_cast<const double2*>(rp); r1 = *reinterpret_cast<const double2*>(rp + 2);
  };
  auto sstore = [&](int buf) {
    double* lp = Ls + (buf * GK + srow) * GLD + scol;
    double* rp = Rs + (buf * GK + srow) * GLD + scol;
    *reinterpret_cast<double2*>(lp) = l0; *reinterpret_cast<double2*>(lp + 2) = l1;
    *reinterpret_cast<double2*>(rp) = r0; *reinterpret_cast<double2*>(rp + 2) = r1;
  };
  gload(0);
  sstore(0);
  __syncthreads();
  const int nchunk = K / GK;
  for (int ch = 0; ch < nchunk; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunk) gload((ch + 1) * GK);
    const double* lb = Ls + buf * GK * GLD;
    const double* rb = Rs + buf * GK * GLD;
#pragma unroll
    for (int k4 = 0; k4 < GK / 4; ++k4) {
      double a[4], b[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        a[t] = lb[(4 * k4 + tg) * GLD + wi + 8 * t + g];
        b[t] = rb[(4 * k4 + tg) * GLD + wj + 8 * t + g];
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(d[mt][nt][0], d[mt][nt][1], a[mt], b[nt]);
    }
    if (ch + 1 < nchunk) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      double* o = Out + (size_t)(i0 + wi + 8 * mt + g) * ld + j0 + wj + 8 * nt + 2 * tg;
      *reinterpret_cast<double2*>(o) = make_double2(d[mt][nt][0], d[mt][nt][1]);
    }
}

// ---- two-level block Jacobi ------------------------------------------------------------------------------------------

// global index of local index l (0..63) of the outer block pair (I, J)
XTB_DEV int op_index(int I, int J, int l) { return (l < OB ? I * OB : J * OB - OB) + l; }

// round-robin pairing of nblk blocks (nblk even): pair w of round r
XTB_DEV void rr_pair(int nblk, int r, int w, int& I, int& J) {
  if (w == 0) { I = r; J = nblk - 1; }
  else {
    I = (r + w) % (nblk - 1);
    J = (r - w + 2 * (nblk - 1)) % (nblk - 1);
  }
  if (I > J) { const int t = I; I = J; J = t; }
}

// block paired with block x in round r of the same schedule
XTB_DEV int rr_partner(int nblk, int r, int x) {
  const int nb1 = nblk - 1;
  if (x == nb1) return r;
  int y = (2 * r - x) % nb1;
  if (y < 0) y += nb1;
  return y == x ? nb1 : y;
}

constexpr int SUB_MAT = OP * SLD;  // doubles of a 64 x 68 shared-memory matrix
constexpr int SUB_NBP = OP / JB2;  // block pairs of the in-CTA solver on a 64 x 64 sub-problem
constexpr int SUB_SMEM = (2 * SUB_MAT + SUB_NBP * JB2 * QLD + SUB_NBP * JB2 * MLD + 32 + 2 * SUB_NBP + 2) * 8;

// One CTA per outer block pair: 64 x 64 sub-problem on the diagonal tile, one in-CTA Jacobi sweep, Q -> global.
__global__ void __launch_bounds__(NT, 1)
kl_jacobi_sub(const double* __restrict__ A, double* __restrict__ Qs, int ne, int r, double tol) {
  extern __shared__ double ssm[];
  double* As = ssm;
  double* Vs = ssm + SUB_MAT;
  Ctx c;
  c.ne = OP; c.ld = SLD; c.n = OP; c.np = OP / 2;
  double* p = ssm + 2 * SUB_MAT;
  c.jq = p; p += SUB_NBP * JB2 * QLD;
  c.ng = SUB_NBP;
  c.jm = p; p += SUB_NBP * JB2 * MLD;
  c.jr = nullptr;
  c.red = p; p += 32;
  c.pp = (int*)p;
  c.status = 0; c.sweeps = 0; c.defer = false;
  const int w = blockIdx.x, nblk = ne / OB;
  int I, J;
  rr_pair(nblk, r, w, I, J);
  for (int t = threadIdx.x; t < OP * OP; t += NT) {
    const int a = t >> 6, b = t & 63;
    As[a * SLD + b] = A[(size_t)op_index(I, J, a) * ne + op_index(I, J, b)];
    Vs[a * SLD + b] = (a == b) ? 1.0 : 0.0;
  }
  __syncthreads();
  jacobi<true, true>(c, As, Vs, OP, tol, 1);
  __syncthreads();
  double* Q = Qs + (size_t)w * OP * OP;
  for (int t = threadIdx.x; t < OP * OP; t += NT) Q[t] = Vs[(t >> 6) * SLD + (t & 63)];
}

constexpr int PASS_SMEM = 2 * SUB_MAT * 8;

// 16-byte asynchronous copy global -> shared (LDGSTS), no register staging
XTB_DEV void cp_async16(double* smem_dst, const double* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
XTB_DEV void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

// dst[r][c] (ld SLD) = src[rowbase(r)][colbase(c)] for a 64 x 64 tile whose rows / columns are two 32-blocks (I, J)
XTB_DEV void load_tile_async(double* dst, const double* __restrict__ src, int ne, int rI, int rJ, int cI, int cJ) {
  for (int t = threadIdx.x; t < OP * OP / 2; t += NT) {
    const int r = t >> 5, c = (t & 31) << 1;
    cp_async16(dst + r * SLD + c, src + (size_t)op_index(rI, rJ, r) * ne + op_index(cI, cJ, c));
  }
}
XTB_DEV void load_q_async(double* dst, const double* __restrict__ Q) {
  for (int t = threadIdx.x; t < OP * OP / 2; t += NT) {
    const int r = t >> 5, c = (t & 31) << 1;
    cp_async16(dst + r * SLD + c, Q + r * OP + c);
  }
}

// 16 x 16 warp tile of a 64 x 64 x 64 product with both operands in shared memory (ld SLD, conflict-free either way):
//   LT == false: d[i][j] = sum_k L[k][i] R[k][j];   LT == true: d[i][j] = sum_k L[i][k] R[k][j]
template <bool LT>
XTB_DEV void tile_gemm64(const double* __restrict__ L, const double* __restrict__ R, int i0, int j0, int g, int tg, double (&d)[2][2][2]) {
  XTB_ASSUME_SHARED(L); XTB_ASSUME_SHARED(R);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) d[a][b][0] = d[a][b][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < OP; k0 += 4) {
    double a0, a1;
    if (LT) { a0 = L[(i0 + g) * SLD + k0 + tg]; a1 = L[(i0 + 8 + g) * SLD + k0 + tg]; }
    else { a0 = L[(k0 + tg) * SLD + i0 + g]; a1 = L[(k0 + tg) * SLD + i0 + 8 + g]; }
    const double b0 = R[(k0 + tg) * SLD + j0 + g], b1 = R[(k0 + tg) * SLD + j0 + 8 + g];
    dmma884(d[0][0][0], d[0][0][1], a0, b0);
    dmma884(d[0][1][0], d[0][1][1], a0, b1);
    dmma884(d[1][0][0], d[1][0][1], a1, b0);
    dmma884(d[1][1][0], d[1][1][1], a1, b1);
  }
}

// Apply the accumulated rotations of one outer round: A <- Q^T A Q (tiles P >= R, mirrored) and V <- V Q.  One 64 x 64
// tile per CTA, 16 warps x (16 x 16) DMMA warp tiles, operands staged with cp.async into two shared-memory buffers
// (3 CTAs per SM overlap the loads of one tile with the tensor-core work of the others); results go from the
// accumulators straight to global memory.
__global__ void __launch_bounds__(NT, 3)
kl_jacobi_pass(double* __restrict__ A, double* __restrict__ V, const double* __restrict__ Qs, int ne, int nbp, int r, int phase) {
  extern __shared__ double psm[];
  double* B0 = psm;
  double* B1 = psm + SUB_MAT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int i0 = (warp >> 2) << 4, j0 = (warp & 3) << 4;
  const int nfused = nbp * (nbp + 1) / 2;
  const int u = blockIdx.x;
  double d[2][2][2];
  if (u < nfused) {
    int P = (int)((sqrtf(8.0f * (float)u + 1.0f) - 1.0f) * 0.5f);
    while ((P + 1) * (P + 2) / 2 <= u) ++P;
    while (P * (P + 1) / 2 > u) --P;
    const int R = u - P * (P + 1) / 2;
    const int nblk = ne / OB;
    int IP, JP, IR, JR;
    rr_pair(nblk, r, P, IP, JP);
    rr_pair(nblk, r, R, IR, JR);
    if (phase != 0) {
      // tiles holding a block pair of the NEXT round (incl. all diagonal tiles) are updated first (phase 1), so that the
      // next sub-problems can start while the remaining tiles (phase 2) are still being rotated
      const int pi = rr_partner(nblk, r + 1, IP), pj = rr_partner(nblk, r + 1, JP);
      const bool needed = P == R || pi == IR || pi == JR || pj == IR || pj == JR;
      if (needed != (phase == 1)) return;
    }
    // B0[i][k] = A[P_i][R_k], B1 = Q_R
    load_tile_async(B0, A, ne, IP, JP, IR, JR);
    load_q_async(B1, Qs + (size_t)R * OP * OP);
    cp_async_wait_all();
    __syncthreads();
    tile_gemm64<true>(B0, B1, i0, j0, g, tg, d);  // T = A_PR Q_R
    __syncthreads();
    load_q_async(B1, Qs + (size_t)P * OP * OP);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
        *reinterpret_cast<double2*>(B0 + (i0 + 8 * a + g) * SLD + j0 + 8 * b + 2 * tg) = make_double2(d[a][b][0], d[a][b][1]);
    cp_async_wait_all();
    __syncthreads();
    tile_gemm64<false>(B1, B0, i0, j0, g, tg, d);  // B' = Q_P^T T
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int gi = op_index(IP, JP, i0 + 8 * a + g);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int gj = op_index(IR, JR, j0 + 8 * b + 2 * tg);  // 2 tg and 2 tg + 1 lie in the same 32-block
        *reinterpret_cast<double2*>(A + (size_t)gi * ne + gj) = make_double2(d[a][b][0], d[a][b][1]);
        if (P != R) {  // mirror
          A[(size_t)gj * ne + gi] = d[a][b][0];
          A[(size_t)(gj + 1) * ne + gi] = d[a][b][1];
        }
      }
    }
  } else {
    if (phase == 1) return;
    const int rem = u - nfused;
    const int k = rem % nbp, rt = rem / nbp;
    int I, J;
    rr_pair(ne / OB, r, k, I, J);
    // B0[i][c] = V[64 rt + i][idx_c], B1 = Q
    for (int t = threadIdx.x; t < OP * OP / 2; t += NT) {
      const int r = t >> 5, c = (t & 31) << 1;
      cp_async16(B0 + r * SLD + c, V + (size_t)(rt * OP + r) * ne + op_index(I, J, c));
    }
    load_q_async(B1, Qs + (size_t)k * OP * OP);
    cp_async_wait_all();
    __syncthreads();
    tile_gemm64<true>(B0, B1, i0, j0, g, tg, d);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
        *reinterpret_cast<double2*>(V + (size_t)(rt * OP + i0 + 8 * a + g) * ne + op_index(I, J, j0 + 8 * b + 2 * tg)) =
            make_double2(d[a][b][0], d[a][b][1]);
  }
}

// ---- SCF response of the nuclear gradient (see scf_response in xtb_scf_core.cuh), grid-wide stages ----------------------
// A = -1/2 S (w_i + w_j), zero padded
__global__ void kl_resp_fock(double* __restrict__ A, const double* __restrict__ S, const double* __restrict__ w, int n, int ne) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    A[t] = (i < n && j < n) ? -0.5 * S[(size_t)i * n + j] * (w[i] + w[j]) : 0.0;
  }
}

// Fermi-level shift per spin channel of the response density: abar_s = sum f'_s(k) At_kk / sum f'_s(k)   (one CTA)
__global__ void kl_resp_abar(const double* __restrict__ A, const double* __restrict__ fp0, const double* __restrict__ fp1, int n, int ne,
                             LargeState* st) {
  __shared__ double red[32];
  double s0 = 0.0, s1 = 0.0, n0 = 0.0, n1 = 0.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double d = A[(size_t)k * ne + k];
    s0 += fp0[k] * d; n0 += fp0[k];
    s1 += fp1[k] * d; n1 += fp1[k];
  }
  s0 = block_sum(s0, red); n0 = block_sum(n0, red);
  s1 = block_sum(s1, red); n1 = block_sum(n1, red);
  if (threadIdx.x == 0) {
    st->abar[0] = fabs(n0) > kTiny ? s0 / n0 : 0.0;
    st->abar[1] = fabs(n1) > kTiny ? s1 / n1 : 0.0;
  }
}

// At -> Zt = At o G (wmat = 0) or ZWt = At o Ge (wmat = 1), symmetrised, pads zero; same formulas as response_density()
__global__ void kl_resp_scale(double* __restrict__ A, const double* __restrict__ eps, const double* __restrict__ focc,
                              const double* __restrict__ fp0, const double* __restrict__ fp1, int n, int ne, int wmat,
                              const LargeState* st) {
  const size_t tot = (size_t)ne * ne;
  const double ab0 = st->abar[0], ab1 = st->abar[1];
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    if (i > j) continue;
    double val = 0.0;
    if (j < n) {
      const double ei = eps[i], ej = eps[j], fi = focc[i], fj = focc[j];
      const double fpi = fp0[i] + fp1[i], fpj = fp0[j] + fp1[j];
      if (i == j) {
        const double d = A[t];
        const double zd = (d - ab0) * fp0[i] + (d - ab1) * fp1[i];
        val = wmat ? d * fi + zd * ei : zd;
      } else {
        const double a = 0.5 * (A[(size_t)i * ne + j] + A[(size_t)j * ne + i]);
        const double de = ej - ei;
        const bool close = fabs(de) <= 1e-9;
        double gf;
        if (wmat) gf = close ? 0.5 * ((fi + ei * fpi) + (fj + ej * fpj)) : (fj * ej - fi * ei) / de;
        else gf = close ? 0.5 * (fpi + fpj) : (fj - fi) / de;
        val = a * gf;
      }
    }
    A[(size_t)i * ne + j] = val;
    A[(size_t)j * ne + i] = val;
  }
}

// out[mu] = add[mu] - sum_nu Z[mu][nu] S[mu][nu]: one warp per row
__global__ void kl_resp_charges(const double* __restrict__ Z, const double* __restrict__ S, const double* __restrict__ add,
                                double* __restrict__ out, int n, int ne) {
  const int lane = threadIdx.x & 31;
  const int mu = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (mu >= n) return;
  const double* zr = Z + (size_t)mu * ne;
  const double* sr = S + (size_t)mu * n;
  double acc = 0.0;
  for (int nu = lane; nu < n; nu += 32) acc = fma(zr[nu], sr[nu], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[mu] = (add ? add[mu] : 0.0) - acc;
}

// dst (n x n) += / = src (ne x ne)
__global__ void kl_pack_add(double* __restrict__ dst, const double* __restrict__ src, int n, int ne, int add) {
  const size_t tot = (size_t)n * n;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t - (size_t)i * n);
    const double v = src[(size_t)i * ne + j];
    dst[t] = add ? dst[t] + v : v;
  }
}

// ---- single-CTA vector stages (reuse the device functions of the batch kernel) ----------------------------------------
enum Phase { PH_INIT = 0, PH_FERMI = 1, PH_POT = 2, PH_MIX = 3, PH_COPYV = 4, PH_EMIT = 5,
             PH_RESP_INIT = 6, PH_RESP_W0 = 7, PH_RESP_STEP = 8, PH_RESP_FINAL = 9 };

__global__ void __launch_bounds__(NT, 1)
kl_vec(int phase, const xtb_batch b, const xtb_scf_opts o, int m, const double* __restrict__ S, const double* __restrict__ H0,
       const double* __restrict__ gamma, const double* __restrict__ nel_ab, const double* __restrict__ q0_at, double* __restrict__ work,
       double* __restrict__ q_orb, double* __restrict__ q_sh, double* __restrict__ q_at, double* __restrict__ v_orb,
       double* __restrict__ e_atom, double* __restrict__ fenergy, double* __restrict__ emo, double* __restrict__ occ,
       int32_t* __restrict__ iterations, int32_t* __restrict__ status, int iters, double* __restrict__ resp) {
  Ctx c;
  large_ctx(c, b, m, work, o.generations, S, H0, gamma);
  const Layout l = layout(c.n, c.ns, c.na, o.generations);
  LargeState* st = reinterpret_cast<LargeState*>(work + l.state);
  const int n = c.n;
  if (phase == PH_INIT) {
    for (int mu = threadIdx.x; mu < n; mu += NT) {
      const int sh = c.ao_sh[mu];
      const int a = c.sh_atom[sh];
      const double deg = (double)(2 * b.sh_l[c.s0 + sh] + 1);
      c.n0[mu] = b.sh_par[(size_t)(c.s0 + sh) * XTB_SHPAR + XTB_SH_REFOCC] / deg;   // scf/iterator.py:147-170
      c.q[mu] = q0_at[c.a0 + a] / (double)c.at_nsh[a] / deg;                          // scf/guess.py:122-182
    }
    if (threadIdx.x == 0) {
      st->g = 0.0; st->off = 0.0; st->mixer_step = 0; st->mixer_head = 0; st->status = 0; st->sweeps = 0; st->nocc = 0;
      st->converged = 0;
    }
    __syncthreads();
    potential(c, c.q, c.v);
  } else if (phase == PH_FERMI) {
    for (int k = threadIdx.x; k < n; k += NT) c.eps[k] = c.A[(size_t)k * c.ld + k];
    __syncthreads();
    const double g = fermi_fill(c, nel_ab[2 * m], nel_ab[2 * m + 1], o);
    if (threadIdx.x == 0) {
      int no = 0;
      for (int k = 0; k < n; ++k)
        if (c.focc[k] > 0.0) c.occl[no++] = k;
      c.occl[n] = no;
      st->nocc = no;
      st->g = g;
      st->status |= c.status;
      st->ef[0] = c.ef[0]; st->ef[1] = c.ef[1];
      st->spin_on[0] = c.spin_on[0]; st->spin_on[1] = c.spin_on[1];
    }
  } else if (phase == PH_POT) {
    potential(c, c.q, c.vnew);
  } else if (phase == PH_MIX) {
    Mixer mx;
    mx.step = st->mixer_step;
    mx.head = st->mixer_head;
    __syncthreads();
    const bool conv = mix(c, mx, o, c.cs);
    if (threadIdx.x == 0) {
      st->mixer_step = mx.step;
      st->mixer_head = mx.head;
      st->converged = conv ? 1 : 0;
    }
  } else if (phase == PH_COPYV) {
    for (int k = threadIdx.x; k < n; k += NT) c.v[k] = c.vnew[k];
  } else if (phase == PH_EMIT) {
    c.status = st->status;
    c.sweeps = st->sweeps;
    emit_results(c, b, m, st->g, iters, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo, occ, iterations, status);
  } else if (phase == PH_RESP_INIT) {
    // SCF response (xtb_scf_core.cuh:scf_response): dv -> eorb, f'_alpha -> srt, f'_beta -> jr; runs after PH_EMIT
    double dvmax = 0.0;
    for (int k = threadIdx.x; k < n; k += NT) {
      c.eorb[k] = c.vnew[k] - c.v[k];
      dvmax = fmax(dvmax, fabs(c.eorb[k]));
      double f[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        f[s] = 0.0;
        if (st->spin_on[s] && o.kt >= 3e-7) {
          const double ex = (c.eps[k] - st->ef[s]) / o.kt;
          if (ex < 50.0) f[s] = 1.0 / (exp(ex) + 1.0);
        }
      }
      c.srt[k] = -(f[0] * (1.0 - f[0])) / o.kt;
      c.jr[k] = -(f[1] * (1.0 - f[1])) / o.kt;
    }
    dvmax = block_max(dvmax, c.red);
    if (dvmax < kResponseTol) {  // nothing to add: hand the plain potential to the gradient
      for (int k = threadIdx.x; k < n; k += NT) resp[c.o0 + k] = v_orb[c.o0 + k];
      for (int k = threadIdx.x; k < c.ns; k += NT) resp[b.nao_tot + c.s0 + k] = 0.0;
    }
    if (threadIdx.x == 0) st->converged = dvmax < kResponseTol ? 1 : 0;
  } else if (phase == PH_RESP_W0) {
    potential_lin(c, c.q, q_at + c.a0, c.v);  // w0 = K z0
  } else if (phase == PH_RESP_STEP) {
    potential_lin(c, c.n0, q_at + c.a0, c.vnew);  // g(w) = K y, y = z0 + chi w; leaves y_sh in c.qsh
    Mixer mx;
    mx.step = iters == 0 ? 0 : st->mixer_step;  // first response step: empty history
    mx.head = iters == 0 ? 0 : st->mixer_head;
    double res = 0.0;
    for (int k = threadIdx.x; k < n; k += NT) res = fmax(res, fabs(c.vnew[k] - c.v[k]));
    res = block_max(res, c.red);
    __syncthreads();
    if (res < kResponseTol) {
      for (int k = threadIdx.x; k < n; k += NT) c.v[k] = c.vnew[k];
      if (threadIdx.x == 0) st->converged = 1;
    } else {
      xtb_scf_opts o2 = o;
      o2.mixer = 0; o2.soft_start = 0; o2.damp = 0.5; o2.damp_init = 0.5; o2.diag_offset = 0.01;
      o2.x_atol = 0.0; o2.x_atol_max = 0.0;
      mix(c, mx, o2, c.cs);
      if (threadIdx.x == 0) st->converged = 0;
    }
    if (threadIdx.x == 0) { st->mixer_step = mx.step; st->mixer_head = mx.head; }
  } else if (phase == PH_RESP_FINAL) {
    for (int k = threadIdx.x; k < c.ns; k += NT) resp[b.nao_tot + c.s0 + k] = c.qsh[k];
    for (int k = threadIdx.x; k < n; k += NT) {
      resp[c.o0 + k] = v_orb[c.o0 + k] + c.v[k];
      c.vnew[k] = c.eorb[k] + c.v[k];  // u = dv + K y
    }
  }
}


// ---- occupied-subspace solve of the intermediate map evaluations, grid-wide (algorithm and safeguards: xtb_scf_subspace.cuh) ----
// All operands are zero padded to the GEMM tile (128 columns, K multiple of 16), so that the GEMMs need no predicates: rows /
// columns of X, X^T beyond (nv, no) stay exactly zero (the update kernel only writes the valid block).
constexpr double kSubBias = 4096.0;  // maxima of possibly negative values are kept as raw bits of (value + bias) >= 0

__device__ __forceinline__ void atomic_max_biased(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v + kSubBias));
}

// Out[i][j] = sum_{k < K} L[k][i] R[k][j] with separate leading dimensions: grid (N / 128, M / 128), K % 16 == 0.
__global__ void __launch_bounds__(NT, 1)
kl_gemm_tn2(const double* __restrict__ L, int ldl, const double* __restrict__ R, int ldr, double* __restrict__ Out, int ldo, int K) {
  extern __shared__ double gsm[];
  double* Ls = gsm;                  // [2][GK][GLD]
  double* Rs = gsm + 2 * GK * GLD;   // [2][GK][GLD]
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int wi = (warp >> 2) * 32, wj = (warp & 3) * 32;
  const int srow = threadIdx.x >> 5, scol = (threadIdx.x & 31) * 4;
  double d[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) d[a][b][0] = d[a][b][1] = 0.0;
  double2 l0, l1, r0, r1;
  auto gload = [&](int k0) {
    const double* lp = L + (size_t)(k0 + srow) * ldl + i0 + scol;
    const double* rp = R + (size_t)(k0 + srow) * ldr + j0 + scol;
    l0 = *reinterpret_cast<const double2*>(lp); l1 = *reinterpret_cast<const double2*>(lp + 2);
    r0 = *reinterpret_cast<const double2*>(rp); r1 = *reinterpret_cast<const double2*>(rp + 2);
  };
  auto sstore = [&](int buf) {
    double* lp = Ls + (buf * GK + srow) * GLD + scol;
    double* rp = Rs + (buf * GK + srow) * GLD + scol;
    *reinterpret_cast<double2*>(lp) = l0; *reinterpret_cast<double2*>(lp + 2) = l1;
    *reinterpret_cast<double2*>(rp) = r0; *reinterpret_cast<double2*>(rp + 2) = r1;
  };
  gload(0);
  sstore(0);
  __syncthreads();
  const int nchunk = K / GK;
  for (int ch = 0; ch < nchunk; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunk) gload((ch + 1) * GK);
    const double* lb = Ls + buf * GK * GLD;
    const double* rb = Rs + buf * GK * GLD;
#pragma unroll
    for (int k4 = 0; k4 < GK / 4; ++k4) {
      double a[4], b[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        a[t] = lb[(4 * k4 + tg) * GLD + wi + 8 * t + g];
        b[t] = rb[(4 * k4 + tg) * GLD + wj + 8 * t + g];
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(d[mt][nt][0], d[mt][nt][1], a[mt], b[nt]);
    }
    if (ch + 1 < nchunk) sstore(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      double* o = Out + (size_t)(i0 + wi + 8 * mt + g) * ldo + j0 + wj + 8 * nt + 2 * tg;
      *reinterpret_cast<double2*>(o) = make_double2(d[mt][nt][0], d[mt][nt][1]);
    }
}

__global__ void kl_sub_diag(const double* __restrict__ A, double* __restrict__ eps, int n, int ne) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) eps[k] = A[(size_t)k * ne + k];
}

// ranks of the diagonal (ascending, ties by index); sub_bad: the `no` lowest are not the first `no` positions
__global__ void kl_sub_rank(const double* __restrict__ eps, int* __restrict__ rank, int n, int no, LargeState* st) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double e = eps[k];
  int rk = 0;
  for (int j = 0; j < n; ++j) {
    const double ej = eps[j];
    rk += (ej < e) || (ej == e && j < k);
  }
  rank[k] = rk;
  if ((rk < no) != (k < no)) atomicOr(&st->sub_bad, 1);
}

// Gershgorin bounds inside the two classes (rank == nullptr: classes by position), one warp per row
__global__ void kl_sub_gersh(const double* __restrict__ A, const double* __restrict__ eps, const int* __restrict__ rank, int n, int ne, int no,
                             LargeState* st) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const double* row = A + (size_t)i * ne;
  const bool oi = (rank ? rank[i] : i) < no;
  double s = 0.0;
  for (int j = lane; j < n; j += 32)
    if (j != i && ((rank ? rank[j] : j) < no) == oi) s += fabs(row[j]);
  s = warp_sum(s);
  if (lane == 0) {
    if (oi) atomic_max_biased(&st->sub_hi, eps[i] + s);
    else atomic_max_biased(&st->sub_nlo, s - eps[i]);
  }
}

// dst[mu][rank[k]] = src[mu][k] (columns k >= n stay in place)
__global__ void kl_sub_perm_cols(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ rank, int n, int ne) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int mu = (int)(t / ne), k = (int)(t - (size_t)mu * ne);
    dst[(size_t)mu * ne + (k < n ? rank[k] : k)] = src[t];
  }
}

// dst[rank[i]][rank[j]] = src[i][j]
__global__ void kl_sub_perm_sym(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ rank, int n, int ne) {
  const size_t tot = (size_t)ne * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ne), j = (int)(t - (size_t)i * ne);
    dst[(size_t)(i < n ? rank[i] : i) * ne + (j < n ? rank[j] : j)] = src[t];
  }
}

__global__ void kl_sub_perm_vec(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ rank, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) dst[rank[k]] = src[k];
}

// Lambda = Aoo + Aov X: T[r][i] += A[r][i] for r, i < no
__global__ void kl_sub_lambda(double* __restrict__ T, const double* __restrict__ A, int no, int ne, int ldt) {
  const size_t tot = (size_t)no * no;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(t / no), i = (int)(t - (size_t)r * no);
    T[(size_t)r * ldt + i] += A[(size_t)r * ne + i];
  }
}

// R = Avo + T1 - X Lambda;  X <- X - R / (d_a - d_i)  (X and X^T);  maxima of |R| and |X_new|
__global__ void kl_sub_update(double* __restrict__ X, double* __restrict__ Xt, const double* __restrict__ A, const double* __restrict__ T,
                              const double* __restrict__ T3, const double* __restrict__ eps, int no, int nv, int ne, int ldx, LargeState* st) {
  __shared__ double red[32];
  const size_t tot = (size_t)nv * no;
  double rmax = 0.0, xmax = 0.0;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int a = (int)(t / no), i = (int)(t - (size_t)a * no);
    const double r = A[(size_t)(no + a) * ne + i] + T[(size_t)(no + a) * ldx + i] - T3[(size_t)a * ldx + i];
    const double xn = X[(size_t)a * ldx + i] - r / (eps[no + a] - eps[i]);
    X[(size_t)a * ldx + i] = xn;
    Xt[(size_t)i * ne + a] = xn;
    rmax = fmax(rmax, fabs(r));
    xmax = fmax(xmax, fabs(xn));
  }
  rmax = block_max(rmax, red);
  xmax = block_max(xmax, red);
  if (threadIdx.x == 0) {
    // NaN compares false everywhere: map it to a huge value so that the host sees the failure
    atomic_max_biased(&st->sub_rmax, rmax == rmax ? rmax : 1.0e300);
    atomic_max_biased(&st->sub_xmax, xmax == xmax ? xmax : 1.0e300);
  }
}

// G <- G + 1 on the whole padded diagonal (the pad block of G, and with it of Z, is the identity)
__global__ void kl_sub_addone(double* __restrict__ G, int np, int ld) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) G[(size_t)i * ld + i] += 1.0;
}

// Z0 = 2 - G
__global__ void kl_sub_zinit(double* __restrict__ Z, const double* __restrict__ G, int np, int ld) {
  const size_t tot = (size_t)np * np;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / np), j = (int)(t - (size_t)i * np);
    Z[(size_t)i * ld + j] = (i == j ? 2.0 : 0.0) - G[(size_t)i * ld + j];
  }
}

// E <- 1 - E (E held G Z); max |E|
__global__ void kl_sub_eres(double* __restrict__ E, int np, int ld, LargeState* st) {
  __shared__ double red[32];
  const size_t tot = (size_t)np * np;
  double emax = 0.0;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / np), j = (int)(t - (size_t)i * np);
    const double e = (i == j ? 1.0 : 0.0) - E[(size_t)i * ld + j];
    E[(size_t)i * ld + j] = e;
    emax = fmax(emax, fabs(e));
  }
  emax = block_max(emax, red);
  if (threadIdx.x == 0) atomic_max_biased(&st->sub_emax, emax == emax ? emax : 1.0e300);
}

// Z <- Z + D on the np x np block
__global__ void kl_sub_add(double* __restrict__ Z, const double* __restrict__ D, int np, int ld) {
  const size_t tot = (size_t)np * np;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / np), j = (int)(t - (size_t)i * np);
    Z[(size_t)i * ld + j] += D[(size_t)i * ld + j];
  }
}

// Y^T[k][mu] += C^T[k][mu] for the occupied rows k < no
__global__ void kl_sub_addrows(double* __restrict__ Yt, const double* __restrict__ Ct, int no, int ne) {
  const size_t tot = (size_t)no * ne;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) Yt[t] += Ct[t];
}

__global__ void kl_sub_scale(double* __restrict__ p, size_t count, double f) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < count; t += (size_t)gridDim.x * blockDim.x) p[t] *= f;
}

struct SideStream {
  int dev;
  cudaStream_t main, side;
  cudaEvent_t ev_a, ev_s;
};

// thread-safe lookup / creation of the helper stream of (device, caller stream)
SideStream* side_stream_for(int dev, cudaStream_t main) {
  static std::mutex mtx;
  static std::vector<SideStream*> all;
  std::lock_guard<std::mutex> lock(mtx);
  for (SideStream* s : all)
    if (s->dev == dev && s->main == main) return s;
  SideStream* s = new SideStream{dev, main, nullptr, nullptr, nullptr};
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (cudaStreamCreateWithPriority(&s->side, cudaStreamNonBlocking, hi) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_a, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_s, cudaEventDisableTiming) != cudaSuccess) {
    delete s;
    return nullptr;
  }
  all.push_back(s);
  return s;
}

template <typename F>
int set_smem(F f, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

extern "C" int64_t xtb_scf_large_workspace_bytes(int32_t nao, int32_t nsh, int32_t nat, int32_t generations) {
  if (nao <= 0 || nsh <= 0 || nat <= 0 || generations < 1) return -1;
  return (int64_t)layout(nao, nsh, nat, generations).total * 8 + 256;
}

// SCF of molecule `mol` of the batch on the whole device (host-driven; synchronises `stream`).
extern "C" int xtb_scf_run_large(const xtb_batch* b, const xtb_scf_opts* o, int32_t mol, int32_t nao, int32_t nsh, int32_t nat,
                                 const double* S, const double* H0, const double* gamma, const double* nel_ab, const double* q0_at,
                                 void* work_, double* q_orb, double* q_sh, double* q_at, double* v_orb, double* e_atom, double* fenergy,
                                 double* emo, double* occ, int32_t* iterations, int32_t* status, double* P, double* W,
                                 double* resp, int64_t mat_off, void* stream) {
  if (!b || !o || !S || !H0 || !gamma || !nel_ab || !q0_at || !work_ || !q_orb || !q_sh || !q_at || !v_orb || !e_atom || !fenergy ||
      !emo || !occ || !iterations || !status)
    return -1;
  if (o->want_density && (!P || !W)) return -1;
  if (o->generations > 5 || o->generations < 1) return -3;
  if (mol < 0 || mol >= b->nb) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  double* work = (double*)work_;
  const Layout l = layout(nao, nsh, nat, o->generations);
  const int n = nao, ne = l.ne, nbp = l.nbp, nblk = ne / OB;
  double *C = work + l.C, *A = work + l.A, *X = work + l.X, *Qs = work + l.Q;
  LargeState* dst = reinterpret_cast<LargeState*>(work + l.state);
  const double* Sm = S + mat_off;
  const double* Hm = H0 + mat_off;
  double* vecp = work + l.vec;
  const int nmx = n + 2;
  double *eps = vecp, *srt = vecp + (size_t)nmx, *focc = vecp + 2 * (size_t)nmx, *v = vecp + 3 * (size_t)nmx, *vnew = vecp + 4 * (size_t)nmx,
         *q = vecp + 5 * (size_t)nmx, *n0 = vecp + 6 * (size_t)nmx, *eorb = vecp + 7 * (size_t)nmx;
  const int* occl = (const int*)(vecp + 8 * (size_t)nmx + 2 * (size_t)nsh + nat + 32 + 36);
  double* fp1 = vecp + 8 * (size_t)nmx + 2 * (size_t)nsh + nat + 32 + 36 + (size_t)(n + 4) / 2 + 2;  // large_ctx: c.jr

  static bool configured[64] = {};  // function attributes are per device
  {
    static std::mutex cfg_mtx;
    std::lock_guard<std::mutex> lock(cfg_mtx);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -5;
    if (!configured[dev]) {
      int e = set_smem(kl_gemm_tn, 4 * GK * GLD * 8);
      if (!e) e = set_smem(kl_gemm_tn2, 4 * GK * GLD * 8);
      if (!e) e = set_smem(kl_jacobi_sub, SUB_SMEM);
      if (!e) e = set_smem(kl_jacobi_pass, PASS_SMEM);
      if (e) return e;
      configured[dev] = true;
    }
  }
  int n_sm = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  // second stream (highest priority) and events for the sub-problem / pass overlap of the Jacobi rounds: one set per
  // (device, caller stream), so that several large molecules can be driven concurrently from different host threads on
  // different streams (a 550-AO molecule fills only 25..155 CTAs per launch)
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  SideStream* ss = side_stream_for(cur_dev, st);
  if (!ss) return -5;
  cudaStream_t side = ss->side;
  cudaEvent_t ev_a = ss->ev_a, ev_s = ss->ev_s;
  const int ew_grid = 4 * n_sm;  // elementwise kernels: grid-stride
  LargeState hs;
  auto read_state = [&]() -> int {
    cudaError_t e = cudaMemcpyAsync(&hs, dst, sizeof(LargeState), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);
    return e == cudaSuccess ? 0 : (int)e;
  };
  auto vec = [&](int phase, int iters) {
    kl_vec<<<1, NT, 0, st>>>(phase, *b, *o, mol, S, H0, gamma, nel_ab, q0_at, work, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo, occ,
                             iterations, status, iters, resp);
  };
  auto gemm = [&](const double* L, const double* R, double* Out, int K) {
    kl_gemm_tn<<<dim3(ne / GT, ne / GT), NT, 4 * GK * GLD * 8, st>>>(L, R, Out, ne, K);
  };
  int total_sweeps = 0, hstatus = 0;
  // Jacobi on (Am, Vm) until max |off-diagonal| <= tol
  // (diagonal: set when the off-diagonal test passed; with a null pointer running out of sweeps is an error status)
  auto jacobi_large = [&](double* Am, double* Vm, double tol, int maxsweeps, bool* diagonal = nullptr) -> int {
    const int npass = nbp * (nbp + 1) / 2 + (ne / OP) * nbp;
    if (diagonal) *diagonal = false;
    for (int sweep = 0;; ++sweep) {
      cudaMemsetAsync(&dst->off, 0, sizeof(double), st);
      kl_offmax<<<ew_grid, 256, 0, st>>>(Am, ne, dst);
      if (int e = read_state()) return e;
      if (hs.off <= tol) {
        if (diagonal) *diagonal = true;
        return 0;
      }
      if (sweep >= maxsweeps) {
        if (!diagonal) hstatus |= XTB_STATUS_JACOBI_NOT_CONVERGED;
        return 0;
      }
      ++total_sweeps;
      // Round r: sub-problems -> Q (buffer r & 1), then the rotation pass.  The pass first updates the few tiles the
      // sub-problems of round r + 1 read (phase 1); those sub-problems (nbp CTAs, a third of the SMs) then run on a second,
      // higher-priority stream concurrently with the bulk of the pass (phase 2).
      const size_t qsz = (size_t)nbp * OP * OP;
      kl_jacobi_sub<<<nbp, NT, SUB_SMEM, st>>>(Am, Qs, ne, 0, tol);
      for (int r = 0; r < nblk - 1; ++r) {
        const double* Qr = Qs + (size_t)(r & 1) * qsz;
        if (r + 1 < nblk - 1) {
          kl_jacobi_pass<<<nbp * (nbp + 1) / 2, NT, PASS_SMEM, st>>>(Am, Vm, Qr, ne, nbp, r, 1);
          cudaEventRecord(ev_a, st);
          cudaStreamWaitEvent(side, ev_a, 0);
          kl_jacobi_sub<<<nbp, NT, SUB_SMEM, side>>>(Am, Qs + (size_t)((r + 1) & 1) * qsz, ne, r + 1, tol);
          cudaEventRecord(ev_s, side);
          kl_jacobi_pass<<<npass, NT, PASS_SMEM, st>>>(Am, Vm, Qr, ne, nbp, r, 2);
          cudaStreamWaitEvent(st, ev_s, 0);
        } else {
          kl_jacobi_pass<<<npass, NT, PASS_SMEM, st>>>(Am, Vm, Qr, ne, nbp, r, 0);
        }
      }
    }
  };
  // ---- occupied-subspace solve of the intermediate map evaluations (host-driven; see xtb_scf_subspace.cuh) ----------------
  static const bool debug = getenv("DXTB_B200_DEBUG_LARGE") != nullptr;  // developer printout of the certificate / fixed point
  struct {
    bool eligible, layout, xvalid, zvalid, ctvalid;
    int no, nv, nop, nvp, ko, kv;
    double gapmin;
  } sb{};
  double *sX = work + l.sX, *sXt = work + l.sXt, *sT = work + l.sT, *sT3 = work + l.sT3, *sG = work + l.sG, *sE = work + l.sE, *sZ = work + l.sZ,
         *sCt = work + l.sCt, *sYt = work + l.sYt, *sWt = work + l.sWt;
  const int nh = l.nh;
  int* rank = const_cast<int*>(occl);  // the occupation list of the Fermi stage is only needed after the final solve
  {
    double nelh[2] = {0.0, 1.0};
    if (cudaMemcpyAsync(nelh, nel_ab + 2 * (size_t)mol, 2 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -6;
    if (cudaStreamSynchronize(st) != cudaSuccess) return -6;
    sb.no = (int)rint(nelh[0]);
    sb.nv = n - sb.no;
    sb.eligible = o->subspace != 0 && o->maxiter > 0 && nelh[0] == nelh[1] && fabs(nelh[0] - (double)sb.no) < 1e-9 && sb.no >= 1 && sb.nv >= 1 &&
                  2 * sb.no <= ne;
    sb.nop = (sb.no + GT - 1) / GT * GT;
    sb.nvp = (sb.nv + GT - 1) / GT * GT;
    sb.ko = (sb.no + GK - 1) / GK * GK;
    sb.kv = (sb.nv + GK - 1) / GK * GK;
    sb.gapmin = fmax(o->subspace_gap * o->kt, 0.02);
    if (sb.nop > nh || sb.nvp > ne || sb.no + sb.kv > ne) sb.eligible = false;
  }
  auto gemm2 = [&](const double* L, int ldl, const double* R, int ldr, double* Out, int ldo, int M, int N, int K) {
    kl_gemm_tn2<<<dim3(N / GT, M / GT), NT, 4 * GK * GLD * 8, st>>>(L, ldl, R, ldr, Out, ldo, K);
  };
  // certified gap between the classes (rank: by the ranks of the diagonal, else by position)
  auto certify = [&](bool ranked, double* gapc, bool* needs_perm) -> int {
    cudaMemsetAsync(&dst->sub_bad, 0, 2 * sizeof(int) + 2 * sizeof(double), st);  // sub_bad, pad, sub_hi, sub_nlo
    kl_sub_diag<<<(n + 255) / 256, 256, 0, st>>>(A, eps, n, ne);
    if (ranked) kl_sub_rank<<<(n + 255) / 256, 256, 0, st>>>(eps, rank, n, sb.no, dst);
    kl_sub_gersh<<<(n + 7) / 8, 256, 0, st>>>(A, eps, ranked ? rank : nullptr, n, ne, sb.no, dst);
    if (int e = read_state()) return e;
    *gapc = -(hs.sub_nlo - kSubBias) - (hs.sub_hi - kSubBias);
    *needs_perm = ranked && hs.sub_bad != 0;
    return 0;
  };
  auto permute = [&]() {
    kl_sub_perm_cols<<<ew_grid, 256, 0, st>>>(X, C, rank, n, ne);
    cudaMemcpyAsync(C, X, (size_t)ne * ne * 8, cudaMemcpyDeviceToDevice, st);
    kl_sub_perm_sym<<<ew_grid, 256, 0, st>>>(X, A, rank, n, ne);
    cudaMemcpyAsync(A, X, (size_t)ne * ne * 8, cudaMemcpyDeviceToDevice, st);
    kl_sub_perm_vec<<<(n + 255) / 256, 256, 0, st>>>(srt, eps, rank, n);
    cudaMemcpyAsync(eps, srt, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
  };
  // Riccati fixed point; *ok = converged
  auto riccati = [&](bool* ok) -> int {
    *ok = false;
    if (!sb.xvalid) {
      cudaMemsetAsync(sX, 0, (size_t)ne * nh * 8, st);
      cudaMemsetAsync(sXt, 0, (size_t)ne * nh * 8, st);
      sb.xvalid = true;
      sb.zvalid = false;
    }
    double rprev = 1.0e300;
    for (int it = 0; it < o->subspace_maxiter; ++it) {
      gemm2(A + (size_t)sb.no * ne, ne, sX, nh, sT, nh, ne, sb.nop, sb.kv);  // T[r][i] = sum_b A[no + b][r] X[b][i]
      kl_sub_lambda<<<ew_grid, 256, 0, st>>>(sT, A, sb.no, ne, nh);             // rows r < no: + Aoo -> Lambda
      gemm2(sXt, ne, sT, nh, sT3, nh, sb.nvp, sb.nop, sb.ko);                  // T3[a][i] = sum_j X[a][j] Lambda[j][i]
      cudaMemsetAsync(&dst->sub_rmax, 0, 2 * sizeof(double), st);              // sub_rmax, sub_xmax
      kl_sub_update<<<ew_grid, 256, 0, st>>>(sX, sXt, A, sT, sT3, eps, sb.no, sb.nv, ne, nh, dst);
      if (int e = read_state()) return e;
      const double rmax = hs.sub_rmax - kSubBias, xmax = hs.sub_xmax - kSubBias;
      if (!(xmax < 1.0) || !(rmax < 1.0e200)) return 0;
      const double rate = it > 0 ? fmin(0.5, 2.0 * rmax / rprev) : 1.0;
      if (debug) fprintf(stderr, "    riccati it %d: max|R| %.3e max|X| %.3e\n", it, rmax, xmax);
      if (rmax * rate <= o->subspace_tol) { *ok = true; return 0; }
      if (it >= 2 && rmax > 2.0 * rprev) return 0;
      rprev = rmax;
    }
    return 0;
  };
  // P = 2 Y Z Y^T -> A; *ok = false if the Newton iteration for Z failed (A untouched then)
  auto density = [&](bool* ok) -> int {
    *ok = false;
    const int np = sb.nop;
    gemm2(sX, nh, sX, nh, sG, nh, np, np, sb.kv);  // X^T X
    kl_sub_addone<<<(np + 255) / 256, 256, 0, st>>>(sG, np, nh);
    for (int it = 0; it < 12; ++it) {
      if (!sb.zvalid) {
        kl_sub_zinit<<<ew_grid, 256, 0, st>>>(sZ, sG, np, nh);
        sb.zvalid = true;
      }
      gemm2(sG, nh, sZ, nh, sE, nh, np, np, np);  // G Z (G symmetric)
      cudaMemsetAsync(&dst->sub_emax, 0, sizeof(double), st);
      kl_sub_eres<<<ew_grid, 256, 0, st>>>(sE, np, nh, dst);
      if (int e = read_state()) return e;
      const double emax = hs.sub_emax - kSubBias;
      if (emax <= 1e-12) { *ok = true; break; }
      if (!(emax < 0.5)) {
        if (it > 0) break;
        sb.zvalid = false;
        continue;
      }
      gemm2(sZ, nh, sE, nh, sT, nh, np, np, np);  // Z E (Z symmetric up to the Newton residual)
      kl_sub_add<<<ew_grid, 256, 0, st>>>(sZ, sT, np, nh);
      if ((double)sb.no * emax * emax <= 1e-13) { *ok = true; break; }
    }
    if (!*ok) {
      sb.zvalid = false;
      return 0;
    }
    if (!sb.ctvalid) {
      kl_transpose<<<dim3(ne / 32, ne / 32), 256, 0, st>>>(sCt, C, ne);
      sb.ctvalid = true;
    }
    gemm2(sX, nh, sCt + (size_t)sb.no * ne, ne, sYt, ne, np, ne, sb.kv);  // Yt[k][mu] = sum_b X[b][k] C[mu][no + b]
    kl_sub_addrows<<<ew_grid, 256, 0, st>>>(sYt, sCt, sb.no, ne);           // + C[mu][k]
    gemm2(sZ, nh, sYt, ne, sWt, ne, np, ne, np);                           // Wt[k][mu] = sum_j Z[j][k] Yt[j][mu]
    kl_sub_scale<<<ew_grid, 256, 0, st>>>(sWt, (size_t)np * ne, 2.0);
    gemm2(sWt, ne, sYt, ne, A, ne, ne, ne, np);                            // P = 2 Wt^T Yt
    return 0;
  };

  // one SCF map evaluation v -> q -> vnew; final_solve: always the full eigendecomposition
  auto fcn = [&](double jtol, bool final_solve) -> int {
    kl_fock<<<ew_grid, 256, 0, st>>>(A, Hm, Sm, v, n, ne);
    gemm(A, C, X, ne);  // X = F C (F symmetric)
    gemm(C, X, A, ne);  // A = C^T X
    kl_symmetrize<<<ew_grid, 256, 0, st>>>(A, n, ne);
    bool fast = false, diagonal = false;
    if (!final_solve && sb.eligible) {
      // Jacobi sweeps only until the gap between the diagonal blocks is certified, then the Riccati fixed point
      for (int sweeps_here = 0;;) {
        double gapc = -1.0;
        bool needs_perm = false;
        if (sb.layout)
          if (int e = certify(false, &gapc, &needs_perm)) return e;
        if (gapc < sb.gapmin)
          if (int e = certify(true, &gapc, &needs_perm)) return e;
        if (debug) fprintf(stderr, "  large: certified gap %.5f (need %.5f) after %d sweeps of this solve (%d in total)\n", gapc, sb.gapmin, sweeps_here, total_sweeps);
        if (gapc >= sb.gapmin) {
          if (needs_perm) {
            permute();
            sb.xvalid = sb.ctvalid = false;
          }
          sb.layout = true;
          bool ok = false;
          if (int e = riccati(&ok)) return e;
          if (ok) {
            if (int e = density(&ok)) return e;
            if (ok) { fast = true; break; }
          }
        }
        if (sweeps_here >= o->jacobi_max_sweeps) break;
        sb.xvalid = sb.ctvalid = false;
        if (int e = jacobi_large(A, C, jtol, 1, &diagonal)) return e;
        ++sweeps_here;
        if (diagonal) break;
      }
    }
    if (!fast) {
      if (!diagonal) {
        sb.xvalid = sb.ctvalid = false;
        if (int e = jacobi_large(A, C, jtol, o->jacobi_max_sweeps)) return e;
      }
      vec(PH_FERMI, 0);
      if (int e = read_state()) return e;
      const int nocc = hs.nocc, kocc = (nocc + GK - 1) / GK * GK;
      if (kocc > 0) {
        kl_build_y<<<dim3(ne / 32, (kocc + 31) / 32), 256, 0, st>>>(X, nullptr, C, focc, eps, occl, n, ne, nocc, kocc, 0);
        gemm(X, X, A, kocc);  // P = Y^T Y
      } else {
        cudaMemsetAsync(A, 0, (size_t)ne * ne * 8, st);
      }
    }
    kl_mulliken<<<(n + 7) / 8, 256, 0, st>>>(A, Sm, Hm, n0, q, eorb, n, ne);
    vec(PH_POT, 0);
    return launch_status();
  };

  // one Newton-Schulz step C <- C (3/2 I - 1/2 C^T S C), see reorthonormalize() in xtb_scf.cu
  auto reorthonormalize = [&]() {
    kl_load<<<ew_grid, 256, 0, st>>>(A, Sm, n, ne, 0.0);
    gemm(A, C, X, ne);  // X = S C
    gemm(C, X, A, ne);  // A = C^T S C
    kl_ns_factor<<<ew_grid, 256, 0, st>>>(A, ne);
    kl_transpose<<<dim3(ne / 32, ne / 32), 256, 0, st>>>(X, C, ne);
    gemm(X, A, C, ne);  // C = C M
  };

  vec(PH_INIT, 0);
  // start basis from the eigendecomposition of S (pad diagonal 1 keeps the padded matrix positive definite)
  kl_load<<<ew_grid, 256, 0, st>>>(A, Sm, n, ne, 1.0);
  kl_load<<<ew_grid, 256, 0, st>>>(C, Sm, 0, ne, 1.0);  // identity
  if (int e = jacobi_large(A, C, o->jacobi_tol, 2 * o->jacobi_max_sweeps)) return e;
  kl_scale_cols<<<ew_grid, 256, 0, st>>>(C, A, n, ne, dst);
  reorthonormalize();

  int iters = 1;
  bool converged = true;
  if (int e = fcn(o->maxiter > 0 ? o->jacobi_tol_iter : o->jacobi_tol, o->maxiter <= 0)) return e;
  if (o->maxiter > 0) {
    converged = false;
    vec(PH_MIX, 0);  // mix_guess (unrolling/default.py:93-94); convergence is not tested here
    for (int it = 0; it < o->maxiter; ++it) {
      if (int e = fcn(o->jacobi_tol_iter, false)) return e;
      ++iters;
      vec(PH_MIX, 0);
      if (int e = read_state()) return e;
      if (hs.converged) { converged = true; break; }
    }
    vec(PH_COPYV, 0);  // converged_to_charges: one more solve with the un-mixed potential
    reorthonormalize();
    if (int e = fcn(o->jacobi_tol, true)) return e;
  }
  if (!converged) hstatus |= XTB_STATUS_SCF_NOT_CONVERGED;
  // fold the host-side status / sweep count into the device state, then emit
  if (int e = read_state()) return e;
  hs.status |= hstatus;
  hs.sweeps = total_sweeps;
  cudaMemcpyAsync(dst, &hs, sizeof(LargeState), cudaMemcpyHostToDevice, st);
  vec(PH_EMIT, iters);
  if (o->want_density) {
    kl_pack<<<ew_grid, 256, 0, st>>>(P + mat_off, A, n, ne);
    const bool response = resp != nullptr && o->maxiter > 0;
    if (response) {
      // First-order response of the SCF residual (xtb_scf_core.cuh:scf_response), every stage grid-wide; C is scratch of
      // the back-transformation and restored by the second transpose.  Z (or ZW) of the perturbation w ends up in A.
      auto respond = [&](const double* w, int wmat) {
        kl_resp_fock<<<ew_grid, 256, 0, st>>>(A, Sm, w, n, ne);
        gemm(A, C, X, ne);  // X = A_w C
        gemm(C, X, A, ne);  // A = C^T A_w C
        kl_resp_abar<<<1, 512, 0, st>>>(A, srt, fp1, n, ne, dst);
        kl_resp_scale<<<ew_grid, 256, 0, st>>>(A, eps, focc, srt, fp1, n, ne, wmat, dst);
        kl_transpose<<<dim3(ne / 32, ne / 32), 256, 0, st>>>(X, C, ne);  // X = C^T
        gemm(A, X, C, ne);                                               // C <- T = Zt C^T
        gemm(X, C, A, ne);                                               // A = C T = Z
        kl_transpose<<<dim3(ne / 32, ne / 32), 256, 0, st>>>(C, X, ne);  // C restored
      };
      vec(PH_RESP_INIT, 0);
      if (int e = read_state()) return e;
      if (!hs.converged) {
        respond(eorb, 0);
        kl_resp_charges<<<(n + 7) / 8, 256, 0, st>>>(A, Sm, nullptr, q, n, ne);  // z0 = chi dv
        vec(PH_RESP_W0, 0);
        for (int it = 0; it < kResponseMaxIter; ++it) {
          respond(v, 0);
          kl_resp_charges<<<(n + 7) / 8, 256, 0, st>>>(A, Sm, q, n0, n, ne);  // y = z0 + chi w
          vec(PH_RESP_STEP, it);
          if (int e = read_state()) return e;
          if (hs.converged) break;
        }
        vec(PH_RESP_FINAL, 0);
        respond(vnew, 0);
        kl_pack_add<<<ew_grid, 256, 0, st>>>(P + mat_off, A, n, ne, 1);  // P += Z_u
        respond(vnew, 1);
        kl_pack_add<<<ew_grid, 256, 0, st>>>(W + mat_off, A, n, ne, 0);  // W = ZW_u (+ the density-weighted part below)
      } else {
        cudaMemsetAsync(W + mat_off, 0, (size_t)n * n * 8, st);
      }
    }
    const int nocc = hs.nocc, kocc = (nocc + GK - 1) / GK * GK;
    if (kocc > 0) {
      // W = C diag(f eps) C^T = Y1^T Y2 (X, A buffers) -> C buffer (no longer needed)
      kl_build_y<<<dim3(ne / 32, (kocc + 31) / 32), 256, 0, st>>>(X, A, C, focc, eps, occl, n, ne, nocc, kocc, 1);
      // C is read by kl_build_y and overwritten by the GEMM: same stream, ordered
      gemm(X, A, C, kocc);
      kl_pack_add<<<ew_grid, 256, 0, st>>>(W + mat_off, C, n, ne, response ? 1 : 0);
    } else if (!response) {
      cudaMemsetAsync(W + mat_off, 0, (size_t)n * n * 8, st);
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);  // hs lives on this stack frame
  if (e != cudaSuccess) return (int)e;
  return launch_status();
}
