// On-device self-consistent field loop of the GFN1-xTB single point: ONE CTA PER MOLECULE runs the whole
// SCF (Fock build, eigensolve, Fermi filling, density, Mulliken populations, ES2/ES3 potential, Anderson
// mixing, per-molecule convergence, final energies) with no host round trip (scf/unrolling/default.py:71-136).
//
// Eigensolver: the generalised problem F C = S C eps is solved in the S-orthonormal basis of the previous
// iteration's eigenvectors (C^T S C = I), i.e. A = C^T F C is nearly diagonal after the first iterations and the
// blocked two-sided Jacobi of xtb_scf_core.cuh (warp-synchronous 16x16 sub-problems + fp64 tensor-core rotation
// passes) converges in 2-3 sweeps instead of 7-9.  The rotations are applied to C directly, so there is no
// back-transformation.  The first basis is C0 = L^-T from an in-CTA Cholesky factorisation S = L L^T, followed by
// one Newton-Schulz re-orthonormalisation step.
//
// Kernel variants (template parameter MODE = xtb_scf_opts::use_smem): 1 = C, A/F/P and X in shared memory (nao <= 80),
// 2 = only A in shared memory, C and X in the L2-resident workspace (nao <= ~128), 0 = all three in the workspace.
// Leading dimension ld = ne + 4 (== 4 mod 16): every DMMA fragment load is bank-conflict free.
#include "xtb_scf_core.cuh"
#include "xtb_scf_subspace.cuh"

#include <cstdlib>
#ifdef XTB_PROFILE_PHASES
#include <cstdio>
#endif

namespace {

// One Newton-Schulz step C <- C (3/2 I - 1/2 C^T S C): restores C^T S C = I to second order.  The Cholesky start basis
// is S-orthonormal only to cond(S) eps and the Jacobi rotations accumulated over the SCF let it drift further; the defect
// d enters the energy as sum_k f_k eps_k d_kk (3e-10 Eh for the 550-AO vancoh2, 1e-8 Eh for 3104 AOs).  Uses A and X.
template <int MODE>
XTB_CTX_FN void reorthonormalize(Ctx& c) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  for (int t = threadIdx.x; t < ne * ld; t += NT) {
    const int i = t / ld, j = t - i * ld;
    c.A[t] = (i < n && j < n) ? c.S[(size_t)i * n + j] : 0.0;
  }
  __syncthreads();
  gemm_tn<AS, CS>(ne, ne, c.A, c.C, ld, c.X, ld, ne);  // X = S C
  gemm_tn<CS, CS>(ne, ne, c.C, c.X, ld, c.A, ld, ne);  // A = C^T S C
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int i = t / ne, j = t - i * ne;
    if (i <= j) {
      const double g = 0.5 * (c.A[(size_t)i * ld + j] + c.A[(size_t)j * ld + i]);
      const double m = (i == j ? 1.5 : 0.0) - 0.5 * g;
      c.A[(size_t)i * ld + j] = m;
      c.A[(size_t)j * ld + i] = m;
    }
  }
  for (int t = threadIdx.x; t < ne * ne; t += NT) {
    const int k = t / ne, i = t - k * ne;
    c.X[(size_t)k * ld + i] = c.C[(size_t)i * ld + k];  // X = C^T
  }
  __syncthreads();
  gemm_tn<CS, AS>(ne, ne, c.X, c.A, ld, c.C, ld, ne);  // C = C M
}

// Start basis of a molecule from the final eigenvectors of the previous molecule of this CTA (persistent launch over a batch
// of equally sized molecules, i.e. conformers): C_prev is S_prev-orthonormal; Newton-Schulz steps C <- C (3/2 - 1/2 C^T S C)
// make it S-orthonormal for the new overlap (defect 0.03 for 0.05 bohr perturbations: 0.03 -> 7e-4 -> 4e-7 -> 1e-13).  In that
// basis the first projected Fock matrix is nearly diagonal, the gap certificate holds after 1-2 Jacobi sweeps instead of 4-5.
// Returns false (C is garbage then: the caller falls back to the Cholesky start basis) if the first defect is >= 1/2 -- a
// different molecule with the same dimensions, or an overlap that is not positive definite -- or after 6 steps.
template <int MODE>
XTB_CTX_FN bool warm_start_basis(Ctx& c, const double* __restrict__ prevC) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;
  if (prevC != c.C) {  // hybrid / global variants: every molecule has its own matrices in the workspace
    for (int t = threadIdx.x; t < ne * ld; t += NT) c.C[t] = prevC[t];
    __syncthreads();
  }
  for (int it = 0; it < 6; ++it) {
    for (int t = threadIdx.x; t < ne * ld; t += NT) {
      const int i = t / ld, j = t - i * ld;
      c.A[t] = (i < n && j < n) ? c.S[(size_t)i * n + j] : 0.0;
    }
    __syncthreads();
    gemm_tn<AS, CS>(ne, ne, c.A, c.C, ld, c.X, ld, ne);  // X = S C
    gemm_tn<CS, CS>(ne, ne, c.C, c.X, ld, c.A, ld, ne);  // A = C^T S C
    double defect = 0.0;
    for (int t = threadIdx.x; t < n * n; t += NT) {
      const int i = t / n, j = t - i * n;
      defect = fmax(defect, fabs(c.A[(size_t)i * ld + j] - (i == j ? 1.0 : 0.0)));
    }
    defect = block_max(defect, c.red);
    __syncthreads();
    if (!(defect < 0.5)) return false;
    if (defect < 1e-12) return true;
    for (int t = threadIdx.x; t < ne * ne; t += NT) {
      const int i = t / ne, j = t - i * ne;
      if (i <= j) {
        const double g = 0.5 * (c.A[(size_t)i * ld + j] + c.A[(size_t)j * ld + i]);
        const double m = (i == j ? 1.5 : 0.0) - 0.5 * g;
        c.A[(size_t)i * ld + j] = m;
        c.A[(size_t)j * ld + i] = m;
      }
    }
    for (int t = threadIdx.x; t < ne * ne; t += NT) {
      const int k = t / ne, i = t - k * ne;
      c.X[(size_t)k * ld + i] = c.C[(size_t)i * ld + k];  // X = C^T
    }
    __syncthreads();
    gemm_tn<CS, AS>(ne, ne, c.X, c.A, ld, c.C, ld, ne);  // C = C M
    if (defect < 1e-6) return true;  // the step squares the defect: no need to measure it again
  }
  return false;
}

// A = C^T F C with F = H0 - 1/2 S o (v_i + v_j), exactly symmetric, pad rows / columns exactly zero (X buffer: F C).
template <int MODE>
XTB_CTX_FN void build_projected_fock(Ctx& c, const double* __restrict__ v) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool AS = MODE != 0, CS = MODE == 1;  // address space of the A buffer / of the C and X buffers
  // F = H0 - 1/2 S (v_i + v_j)   -> A buffer (symmetric, zero padded)
  for (int t = threadIdx.x; t < ne * ld; t += NT) {
    const int i = t / ld, j = t - i * ld;
    double f = 0.0;
    if (i < n && j < n) {
      const size_t ij = (size_t)i * n + j;
      f = c.H0[ij] - 0.5 * c.S[ij] * (v[i] + v[j]);
    }
    c.A[t] = f;
  }
  __syncthreads();
  gemm_tn<AS, CS>(ne, ne, c.A, c.C, ld, c.X, ld, ne);  // X = F C   (F symmetric)
  gemm_tn_sym<CS, CS, AS>(ne, n, c.C, c.X, ld, c.A);   // A = C^T X: lower tiles computed and mirrored, padding exactly zero
}

template <int MODE>
XTB_CTX_FN int jacobi_mode(Ctx& c, double tol, int maxsweeps) {
  constexpr bool AS = MODE != 0, CS = MODE == 1;
#ifdef XTB_PROFILE_PHASES
  const long long tj0 = clock64();
#endif
  c.sub.xvalid = false;  // the basis changes (and X may live in the Jacobi scratch)
  const int sw = (MODE != 1 && c.defer) ? jacobi<AS, CS, MODE != 1>(c, c.A, c.C, c.ne, tol, maxsweeps)
                                        : jacobi<AS, CS, false>(c, c.A, c.C, c.ne, tol, maxsweeps);
#ifdef XTB_PROFILE_PHASES
  c.tjac += clock64() - tj0;
#endif
  c.sweeps += sw < 0 ? -sw : sw;
  return sw;
}

// One SCF map evaluation v -> q -> vnew (scf/base.py:651-675, 818-907, 765-792).
// Returns the electronic free energy of this solve (0 on the occupied-subspace path, whose occupations are integer).
// final_solve: the solve that defines the results -- always the full eigendecomposition.
template <int MODE>
XTB_CTX_FN double fcn(Ctx& c, const double* __restrict__ v, const xtb_scf_opts& o, double nel_a, double nel_b, double jtol,
                      bool final_solve) {
  const int n = c.n, ne = c.ne, ld = c.ld;
  constexpr bool CS = MODE == 1;
  const double* Pb = c.A;  // buffer that holds the density after the solve
  bool fast = false;
  double g = 0.0;
  // Occupied-subspace path (xtb_scf_subspace.cuh): Jacobi sweeps only until the gap between the two diagonal blocks is
  // certified, then the Riccati fixed point for the occupied subspace.  Otherwise (final solve, open shell, small gap):
  // sweeps until A is diagonal.  One loop, so that the projected Fock build and the Jacobi solver have ONE call site each.
  bool try_sub = !final_solve && c.sub.eligible, rebuild = true;
  int sweeps_here = 0;
  for (;;) {
    if (rebuild) {
#ifdef XTB_PROFILE_PHASES
      const long long tf0 = clock64();
#endif
      build_projected_fock<MODE>(c, v);
#ifdef XTB_PROFILE_PHASES
      c.tfock += clock64() - tf0;
#endif
      rebuild = false;
    }
    if (try_sub) {
      bool needs_perm;
#ifdef XTB_PROFILE_PHASES
      const long long tc0 = clock64();
#endif
      double gapc = c.sub.layout ? subspace_certify<false>(c, c.A, needs_perm) : -1.0;
      if (gapc < c.sub.gapmin) gapc = subspace_certify<true>(c, c.A, needs_perm);  // (re)classify by the ranks of the diagonal
#ifdef XTB_PROFILE_PHASES
      c.tcert += clock64() - tc0;
#endif
#ifdef XTB_DEBUG_SUBSPACE
      if (threadIdx.x == 0 && blockIdx.x == 0) printf("  certified gap %.4f (need %.4f) permute %d sweeps so far %d\n", gapc, c.sub.gapmin, (int)needs_perm, c.sweeps);
#endif
      if (gapc >= c.sub.gapmin) {
        if (needs_perm) {
          subspace_permute(c);
          c.sub.xvalid = false;
        }
        c.sub.layout = true;
#ifdef XTB_PROFILE_PHASES
        const long long ts0 = clock64();
#endif
        double* pb = nullptr;
        const bool ok = subspace_riccati<MODE>(c, o);
#ifdef XTB_PROFILE_PHASES
        const long long ts1 = clock64();
        c.tric += ts1 - ts0;
#endif
        if (ok) pb = subspace_density<MODE>(c);
#ifdef XTB_PROFILE_PHASES
        c.tsub += clock64() - ts0;
        c.tden += clock64() - ts1;
#endif
        if (pb != nullptr) {
          Pb = pb;
          fast = true;
          ++c.sub.nfast;
          break;
        }
        if (ok) {  // the density step overwrote A before it failed (not seen in practice): this molecule diagonalises from now on
          c.sub.eligible = try_sub = false;
          rebuild = true;
          continue;
        }
      }
    }
    const int sw = jacobi_mode<MODE>(c, jtol, try_sub ? 1 : o.jacobi_max_sweeps);
    sweeps_here += sw < 0 ? -sw : sw;
    if (sw >= 0) break;  // A is diagonal: finish on the standard path
    if (!try_sub || sweeps_here >= o.jacobi_max_sweeps) {
      c.status |= XTB_STATUS_JACOBI_NOT_CONVERGED;
      break;
    }
  }
  if (!fast) {
    for (int k = threadIdx.x; k < n; k += NT) c.eps[k] = c.A[(size_t)k * ld + k];
    __syncthreads();
    g = fermi_fill(c, nel_a, nel_b, o);
    // compact list of occupied orbitals
    if (threadIdx.x == 0) {
      int no = 0;
      for (int k = 0; k < n; ++k)
        if (c.focc[k] > 0.0) c.occl[no++] = k;
      c.occl[n] = no;
    }
    __syncthreads();
    const int nocc = c.occl[n];
    // Y[kk][i] = sqrt(f_k) C[i][k]  -> X buffer;  P = Y^T Y -> A buffer
    for (int t = threadIdx.x; t < nocc * ld; t += NT) {
      const int kk = t / ld, i = t - kk * ld;
      const int k = c.occl[kk];
      c.X[t] = (i < n) ? sqrt(c.focc[k]) * c.C[(size_t)i * ld + k] : 0.0;
    }
    __syncthreads();
    gemm_tn<CS, CS>(ne, nocc, c.X, c.X, ld, c.A, ld, ne);
  }
  // Mulliken populations and orbital-resolved H0 energies: one warp per row
#ifdef XTB_PROFILE_PHASES
  const long long tm0 = clock64();
#endif
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int mu = w; mu < n; mu += NT / 32) {
    const double* pr = Pb + (size_t)mu * ld;
    const double* sr = c.S + (size_t)mu * n;
    const double* hr = c.H0 + (size_t)mu * n;
    double pop = 0.0, e = 0.0;
    if (final_solve) {
      for (int nu = lane; nu < n; nu += 32) {
        const double p = pr[nu];
        pop = fma(p, sr[nu], pop);
        e = fma(p, hr[nu], e);
      }
      e = warp_sum(e);
    } else {  // the energies of intermediate iterations are never used
      for (int nu = lane; nu < n; nu += 32) pop = fma(pr[nu], sr[nu], pop);
    }
    pop = warp_sum(pop);
    if (lane == 0) {
      c.q[mu] = c.n0[mu] - pop;
      c.eorb[mu] = e;
    }
  }
  __syncthreads();
  potential(c, c.q, c.vnew);
#ifdef XTB_PROFILE_PHASES
  c.tmull += clock64() - tm0;
#endif
  return g;
}

__host__ __device__ inline int64_t vec_smem_bytes(int64_t nao_max, int64_t nsx, int64_t nax, bool v_global) {
  const int64_t nmx = nao_max + 2;
  int64_t d = 2 * (nmx + (nmx & 1)) + 8 * nmx + 2 * nsx + nax + 32 + 36 + (nmx + 2) / 2 + 1;
  d += d & 1;
  const int64_t nbpx = (nao_max + 15) / 16;
  d += (v_global && nbpx <= XTB_DEFER_NBP ? 2 : 1) * nbpx * JB2 * QLD + (nbpx < NGRP ? nbpx : NGRP) * JB2 * MLD;  // block-Jacobi scratch
  d += d & 1;
  return d * 8;
}

// vectors + Jacobi scratch + the matrices the mode keeps in shared memory
__host__ __device__ inline int64_t mode_smem_base(int mode, int64_t nao_max, int64_t nsx, int64_t nax) {
  const int64_t nex = (nao_max + 15) & ~15;
  const int64_t nmat = mode == 1 ? 3 : mode == 2 ? 1 : 0;
  return vec_smem_bytes(nao_max, nsx, nax, mode != 1) + nmat * nex * (nex + 4) * 8;
}

// Hybrid / global-memory variants: scratch of the occupied-subspace solve (T = A(:, v) X, n x lds with no <= ne / 2; afterwards
// the Newton iterates) in shared memory behind everything else, when it fits -- the operands of the small GEMMs then come from
// shared memory instead of L2 (the fixed point was 27 % of the capsaicin single point with T and X in the workspace).
__host__ __device__ inline int64_t sub_scratch_bytes(int mode, int64_t nao_max, int64_t nsx, int64_t nax) {
  if (mode == 1) return 0;
  const int64_t nex = (nao_max + 15) & ~15;
  const int64_t ldsx = ((nex / 2 + 15) & ~15) + 4;
  const int64_t ext = nex * ldsx * 8;
  return mode_smem_base(mode, nao_max, nsx, nax) + ext <= XTB_SMEM_LIMIT ? ext : 0;
}

// dynamic shared memory of a launch
__host__ __device__ inline int64_t mode_smem_bytes(int mode, int64_t nao_max, int64_t nsx, int64_t nax) {
  return mode_smem_base(mode, nao_max, nsx, nax) + sub_scratch_bytes(mode, nao_max, nsx, nax);
}

template <int MODE>
__global__ void __launch_bounds__(NT, XTB_MINB)
k_scf(const xtb_batch b, const xtb_scf_opts o, const double* __restrict__ S, const double* __restrict__ H0,
      const double* __restrict__ gamma, const double* __restrict__ nel_ab, const double* __restrict__ q0_at, double* __restrict__ work,
      double* __restrict__ q_orb, double* __restrict__ q_sh, double* __restrict__ q_at, double* __restrict__ v_orb,
      double* __restrict__ e_atom, double* __restrict__ fenergy, double* __restrict__ emo, double* __restrict__ occ,
      int32_t* __restrict__ iterations, int32_t* __restrict__ status, double* __restrict__ Pout, double* __restrict__ Wout,
      double* __restrict__ resp) {
  extern __shared__ double sm[];
  // one molecule per CTA, or (o.persistent: equally sized molecules, gridDim.x = number of SMs) a fixed-stride walk over the
  // list with the eigenvector warm start between consecutive molecules of the CTA (warm_start_basis)
  const int nslot = o.mol_list ? o.list_len : b.nb;
  int prev_n = -1;
  const double* prev_c = nullptr;
  for (int slot = blockIdx.x; slot < nslot; slot += gridDim.x) {
  const int m = o.mol_list ? o.mol_list[slot] : slot;
  const int lnao = o.mol_list ? o.list_nao_max : b.nao_max, lnsh = o.mol_list ? o.list_nsh_max : b.nsh_max,
            lnat = o.mol_list ? o.list_nat_max : b.nat_max;
  Ctx c;
  c.o0 = b.ao_off[m]; c.s0 = b.sh_off[m]; c.a0 = b.at_off[m];
  c.n = b.ao_off[m + 1] - c.o0;
  c.ns = b.sh_off[m + 1] - c.s0;
  c.na = b.at_off[m + 1] - c.a0;
  c.ne = (c.n + 15) & ~15;  // padded to a multiple of 16 (block-pair size of the Jacobi solver)
  c.ld = c.ne + 4;          // == 4 (mod 16): conflict-free tensor-core fragment loads
  c.np = c.ne / 2;
  c.status = 0;
  c.sweeps = 0;
  c.tp1 = c.tp2 = c.tjac = c.tsub = 0;
#ifdef XTB_PROFILE_PHASES
  c.tcert = c.tric = c.tden = c.tfock = c.tmull = c.tr1 = c.tr2 = c.tr3 = 0;
#endif
#ifdef XTB_PROFILE_PHASES
  const long long tk0 = clock64();
#endif
  const int n = c.n, ne = c.ne, ld = c.ld;
  // shared-memory carve-up (sizes by batch maxima so the layout is launch-uniform)
  const int nmx = lnao + 2, nsx = lnsh, nax = lnat;
  double* p = sm;
  c.cs = p; p += nmx + (nmx & 1);  // double2 per pair: keep 16-byte alignment
  c.pp = (int*)p; p += nmx + (nmx & 1);  // int2 per pair (over-allocated, keeps the matrices 16-byte aligned)
  c.qq = c.pp;
  c.eps = p; p += nmx; c.srt = p; p += nmx; c.focc = p; p += nmx; c.v = p; p += nmx; c.vnew = p; p += nmx;
  c.q = p; p += nmx; c.n0 = p; p += nmx; c.eorb = p; p += nmx;
  c.qsh = p; p += nsx; c.vsh = p; p += nsx; c.qat = p; p += nax;
  c.red = p; p += 32;
  double* sm_theta = p; p += 36;
  c.occl = (int*)p; p += (nmx + 2) / 2 + 1;
  p += ((p - sm) & 1);  // 16-byte alignment for the double2 / int2 scratch and the matrices
  int jcap = 0;  // doubles of block-Jacobi scratch (free between sweeps: home of the occupied-subspace X, MODE 1)
  {
    // block-Jacobi scratch, always in shared memory: accumulated rotations Q per block pair, sub-problem copy and
    // rotation parameters per concurrently working thread group
    const int nbpx = (lnao + 15) / 16;
    c.defer = MODE != 1 && nbpx <= XTB_DEFER_NBP;  // pays off when V lives in the L2-resident workspace
    c.jq = p; p += (c.defer ? 2 : 1) * nbpx * JB2 * QLD;  // double buffered by round parity if the V pass is deferred
    c.ng = nbpx < NGRP ? nbpx : NGRP;
    c.jm = p; p += c.ng * JB2 * MLD;
    c.jr = nullptr;
    jcap = (int)(p - c.jq);
  }
  p += ((p - sm) & 1);
  c.smem = MODE == 1;
  const size_t msz = (size_t)ne * ld;
  // workspace: [Anderson history][2 response matrices per molecule if want_density == 2][occupied-subspace X, Z: 1 matrix
  // bound per molecule][3 matrices per molecule, MODE != 1]
  const size_t mat_bound = (size_t)b.mat_total + 34 * (size_t)b.nao_tot + 285 * (size_t)b.nb;
  const size_t resp_region = o.want_density == 2 ? 2 * mat_bound : 0;
  const size_t mol_bound = (size_t)b.mat_off[m] + 34 * (size_t)c.o0 + 285 * (size_t)m;  // sum of (n+15)(n+19) over the molecules before m
  double* persist = work + (size_t)(o.generations + 1) * 2 * b.nao_tot + resp_region + mol_bound;
  if (MODE == 1) {
    const size_t nex = (size_t)((lnao + 15) & ~15);
    const size_t mszx = nex * (nex + 4);
    c.C = p; c.A = p + mszx; c.X = p + 2 * mszx;
  } else {
    double* wm = work + (size_t)(o.generations + 1) * 2 * b.nao_tot + resp_region + mat_bound + 3 * mol_bound;  // (n+15)(n+19) bounds ne*ld
    c.C = wm; c.A = wm + msz; c.X = wm + 2 * msz;
    if (MODE == 2) c.A = p;  // hybrid: the Fock / A / density buffer (sub-problem gathers, two-sided updates) in shared memory
  }
  c.xh = work + (size_t)(o.generations + 1) * 2 * c.o0;
  c.fh = c.xh + (size_t)(o.generations + 1) * n;
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
  const double nel_a = nel_ab[2 * m], nel_b = nel_ab[2 * m + 1];
  {
    // occupied-subspace path of the intermediate map evaluations (xtb_scf_subspace.cuh): closed shell, integer occupation,
    // Y^T and (Y Z)^T side by side in the A buffer (2 no <= ne); not below 33 AOs, where a Jacobi sweep is one or two block
    // rounds and costs less than the barriers of the fixed point (measured: H2O 0.77x, SiH4 0.99x, MB16_43_01 1.45x)
    Subspace& sb = c.sub;
    sb.no = (int)rint(nel_a);
    sb.nv = c.n - sb.no;
    sb.lds = ((sb.no + 15) & ~15) + 4;
    sb.xvalid = sb.zvalid = sb.layout = false;
    sb.nfast = sb.nric = sb.nnewt = 0;
    sb.gapmin = fmax(o.subspace_gap * o.kt, 0.02);
    sb.eligible = o.subspace != 0 && o.maxiter > 0 && nel_a == nel_b && fabs(nel_a - (double)sb.no) < 1e-9 && sb.no >= 1 && sb.nv >= 1 &&
                  2 * sb.no <= c.ne && c.ne >= 48;
    sb.Zg = persist;  // [no][lds]
    // X [nv][lds]: in the shared-memory variant the Jacobi scratch (free between sweeps), else the workspace
    // ((no + nv) lds <= n (n + 19)).  MODE 1 must never see the workspace pointer here: an address-space hint on X
    // propagates to EVERY source of the pointer (with a select between the two nvcc treated the whole workspace region, Zg
    // included, as shared memory: memcheck "invalid __shared__ write").
    sb.T = c.X;
    if (MODE == 1) {
      sb.X = c.jq;
      if (sb.nv * sb.lds > jcap) sb.eligible = false;
    } else {
      // hybrid / global-memory variants: X in the Jacobi scratch when it fits (no address-space hints on it in these
      // variants), T behind the carve-up when the launch reserved it (sub_scratch_bytes)
      sb.X = sb.nv * sb.lds <= jcap ? c.jq : persist + (size_t)sb.no * sb.lds;
      if (sub_scratch_bytes(MODE, lnao, lnsh, lnat) > 0) sb.T = sm + mode_smem_base(MODE, lnao, lnsh, lnat) / 8;
    }
  }

  // reference occupation per AO (scf/iterator.py:147-170)
  for (int mu = threadIdx.x; mu < n; mu += NT) {
    const int sh = c.ao_sh[mu];
    c.n0[mu] = b.sh_par[(size_t)(c.s0 + sh) * XTB_SHPAR + XTB_SH_REFOCC] / (double)(2 * b.sh_l[c.s0 + sh] + 1);
  }
  // S-orthonormal start basis C0 = L^{-T} from the Cholesky factor S = L L^T (done once; the reference
  // re-factorises S in every iteration inside storch.eighb, scf/unrolling/base.py:141-175)
#ifdef XTB_PROFILE_PHASES
  const long long tc0 = clock64();
#endif
  const bool warm = o.persistent != 0 && prev_n == n && warm_start_basis<MODE>(c, prev_c);
  if (!warm && !cholesky_start_basis<MODE>(c)) c.status |= XTB_STATUS_S_NOT_POSDEF;
#ifdef XTB_PROFILE_PHASES
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("   start basis (%s) %lld\n", warm ? "warm start: Newton-Schulz on the previous eigenvectors" : "Cholesky", clock64() - tc0);
#endif

  // guess: atomic charges spread equally over shells, then over the AOs of a shell (scf/guess.py:122-182)
  for (int mu = threadIdx.x; mu < n; mu += NT) {
    const int sh = c.ao_sh[mu];
    const int a = c.sh_atom[sh];
    c.q[mu] = q0_at[c.a0 + a] / (double)c.at_nsh[a] / (double)(2 * b.sh_l[c.s0 + sh] + 1);
  }
  __syncthreads();
  potential(c, c.q, c.v);  // guess potential (scf/base.py:363-420)

  Mixer mx;
  mx.step = 0;
  mx.head = 0;
  int iters = 0;
  bool converged = o.maxiter <= 0;
  double g = 0.0;
  // scf/unrolling/default.py:71-136 as ONE loop (the map evaluation, the mixer and the re-orthonormalisation have a single call
  // site each, see XTB_CTX_FN): stage 0 = the evaluation outside the reference's loop (default.py:81) followed by mix_guess
  // (default.py:93-94, convergence not tested), stage 1 = the iterations, stage 2 = converged_to_charges: one more solve with
  // the UN-MIXED potential (scf/base.py:497-501, default.py:111-114), which defines the results.
  // Intermediate map evaluations only steer the SCF trajectory: a looser eigensolver tolerance saves the last (verification)
  // sweep; the final solve that defines charges / energies / P / W uses the tight one.
  for (int stage = 0, it = 0;;) {
    const bool final_solve = stage == 2 || o.maxiter <= 0;
    if (stage == 2) {
      for (int k = threadIdx.x; k < n; k += NT) c.v[k] = c.vnew[k];
      __syncthreads();
    }
    // one Newton-Schulz step before the solve that defines the results: the Cholesky start basis is S-orthonormal only to
    // cond(S) eps and the accumulated rotations drift (defect 1e-13 .. 1e-10: irrelevant for the intermediate charges, squared
    // by the step)
    if (stage == 2 || o.maxiter <= 0) reorthonormalize<MODE>(c);
    g = fcn<MODE>(c, c.v, o, nel_a, nel_b, final_solve ? o.jacobi_tol : o.jacobi_tol_iter, final_solve);
    if (stage != 2) ++iters;
    if (final_solve) break;
    const bool conv = mix(c, mx, o, sm_theta);
    if (stage == 0) {
      stage = 1;
    } else if (conv) {
      converged = true;
      stage = 2;
    } else if (++it >= o.maxiter) {
      stage = 2;
    }
  }
  if (!converged) c.status |= XTB_STATUS_SCF_NOT_CONVERGED;

  emit_results(c, b, m, g, iters, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo, occ, iterations, status);
#ifdef XTB_PROFILE_PHASES
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("k_scf phases (cycles): total %lld  jacobi %lld  sub-problems %lld  pass %lld  sweeps %d  subspace %lld (%d map evaluations of %d, %d fixed-point, %d Newton iterations)\n",
           clock64() - tk0, c.tjac, c.tp1, c.tp2, c.sweeps, c.tsub, c.sub.nfast, iters, c.sub.nric, c.sub.nnewt);
  if (threadIdx.x == 0 && blockIdx.x == 0)
    printf("   certify %lld  fixed point %lld (T GEMM %lld, residual GEMM %lld, reduction + copy %lld)  density %lld  | projected Fock (all map evaluations) %lld  Mulliken + potential %lld\n",
           c.tcert, c.tric, c.tr1, c.tr2, c.tr3, c.tden, c.tfock, c.tmull);
#endif
  if (o.want_density) {
    double* Pm = Pout + b.mat_off[m];
    double* Wm = Wout + b.mat_off[m];
    for (int t = threadIdx.x; t < n * n; t += NT) {
      const int i = t / n, j = t - i * n;
      Pm[t] = c.A[(size_t)i * ld + j];
    }
    __syncthreads();
    // W = C diag(f eps) C^T = Y1^T Y2 with Y1[kk][i] = f_k eps_k C[i][k] (X buffer), Y2[kk][i] = C[i][k] (A buffer)
    const int nocc = c.occl[n];
    for (int t = threadIdx.x; t < nocc * ld; t += NT) {
      const int kk = t / ld, i = t - kk * ld;
      const int k = c.occl[kk];
      const double cv = (i < n) ? c.C[(size_t)i * ld + k] : 0.0;
      c.X[t] = c.focc[k] * c.eps[k] * cv;
      c.A[t] = cv;
    }
    __syncthreads();
    gemm_tn<MODE == 1, MODE != 0>(ne, nocc, c.X, c.A, ld, Wm, n, n);
    // first-order response of the SCF residual (forces of the reference = autograd through the unrolled SCF)
    if (resp != nullptr && o.maxiter > 0 && o.want_density == 2) {
      RespBuf rb;
      rb.SC = work + (size_t)(o.generations + 1) * 2 * b.nao_tot + 2 * ((size_t)b.mat_off[m] + 34 * (size_t)c.o0 + 285 * (size_t)m);
      rb.G2 = rb.SC + msz;
      scf_response<MODE>(c, o, rb, v_orb + c.o0, q_at + c.a0, Pm, Wm, resp + c.o0, resp + b.nao_tot + c.s0, sm_theta);
    }
  }
  prev_n = n;
  prev_c = c.C;  // the final eigenvectors (untouched by the P / W output and by the response solver)
  __syncthreads();
  }  // molecules of this CTA
}

#define XTB_SCF_ARGS                                                                                                              \
  const xtb_batch *b, const xtb_scf_opts *o, int nblocks, int lnao, int lnsh, int lnat, const double *S, const double *H0,       \
      const double *gamma, const double *nel_ab, const double *q0_at, double *work, double *q_orb, double *q_sh, double *q_at,    \
      double *v_orb, double *e_atom, double *fenergy, double *emo, double *occ, int32_t *iterations, int32_t *status, double *P, \
      double *W, double *resp, cudaStream_t st

template <int MODE>
int launch_mode(XTB_SCF_ARGS) {
  const int64_t smem = mode_smem_bytes(MODE, lnao, lnsh, lnat);
  static int64_t configured[64] = {};  // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return -5;
  if (smem > configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_scf<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured[dev] = smem;
  }
  int grid = nblocks;
  if (o->persistent) {  // one CTA per SM (the kernel occupies a whole SM: 512 threads x 128 registers)
    int n_sm = 0;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm > 0 && grid > n_sm * XTB_MINB) grid = n_sm * XTB_MINB;
  }
  k_scf<MODE><<<grid, NT, (size_t)smem, st>>>(*b, *o, S, H0, gamma, nel_ab, q0_at, work, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo,
                                               occ, iterations, status, P, W, resp);
  return launch_status();
}

}  // namespace

#ifdef XTB_SECONDARY
// secondary build: launcher of the 2-CTA/SM variants (mode 0: global memory, mode 2: hybrid)
int xtb_scf_launch_2cta(int mode, XTB_SCF_ARGS) {
  if (mode == 2)
    return launch_mode<2>(b, o, nblocks, lnao, lnsh, lnat, S, H0, gamma, nel_ab, q0_at, work, q_orb, q_sh, q_at, v_orb, e_atom, fenergy,
                          emo, occ, iterations, status, P, W, resp, st);
  return launch_mode<0>(b, o, nblocks, lnao, lnsh, lnat, S, H0, gamma, nel_ab, q0_at, work, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo,
                        occ, iterations, status, P, W, resp, st);
}
#else
int xtb_scf_launch_2cta(int mode, XTB_SCF_ARGS);

extern "C" int64_t xtb_scf_smem_bytes_mode(int32_t mode, int32_t nao_max, int32_t nsh_max, int32_t nat_max) {
  if (mode < 0 || mode > 2) return -1;
  return mode_smem_bytes(mode, nao_max, nsh_max, nat_max);
}

extern "C" int64_t xtb_scf_smem_bytes_for(int32_t nao_max, int32_t nsh_max, int32_t nat_max) {
  return mode_smem_bytes(1, nao_max, nsh_max, nat_max);
}

extern "C" int64_t xtb_scf_smem_bytes(const xtb_batch* b) {
  if (!b) return -1;
  return xtb_scf_smem_bytes_for(b->nao_max, b->nsh_max, b->nat_max);
}

extern "C" int64_t xtb_scf_workspace_bytes(const xtb_batch* b, const xtb_scf_opts* o) {
  if (!b || !o) return -1;
  int64_t d = (int64_t)(o->generations + 1) * 2 * b->nao_tot;
  d += b->mat_total + 34 * (int64_t)b->nao_tot + 285 * (int64_t)b->nb;  // occupied-subspace X and Z per molecule
  if (o->want_density == 2) d += 2 * (b->mat_total + 34 * (int64_t)b->nao_tot + 285 * (int64_t)b->nb);  // SCF response: S C and a GEMM output
  if (o->use_smem != 1) {
    // 3 matrices of at most (n+15)(n+19) per molecule: 3 (sum n^2 + 34 sum n + 285 nb)
    d += 3 * (b->mat_total + 34 * (int64_t)b->nao_tot + 285 * (int64_t)b->nb);
  }
  return d * 8 + 256;
}

extern "C" int xtb_scf_run(const xtb_batch* b, const xtb_scf_opts* o, const double* S, const double* H0, const double* gamma,
                           const double* nel_ab, const double* q0_at, void* work, double* q_orb, double* q_sh, double* q_at,
                           double* v_orb, double* e_atom, double* fenergy, double* emo, double* occ, int32_t* iterations,
                           int32_t* status, double* P, double* W, double* resp, void* stream) {
  if (!b || !o || !S || !H0 || !gamma || !nel_ab || !q0_at || !work || !q_orb || !q_sh || !q_at || !v_orb || !e_atom || !fenergy ||
      !emo || !occ || !iterations || !status)
    return -1;
  if (o->want_density && (!P || !W)) return -1;
  if (resp && o->want_density != 2) return -1;  // the workspace must have been sized for the response (want_density = 2)
  if (o->generations > 5 || o->generations < 1) return -3;
  if (o->use_smem < 0 || o->use_smem > 2) return -4;
  if (b->nb == 0) return 0;
  const int nblocks = o->mol_list ? o->list_len : b->nb;
  if (nblocks <= 0) return 0;
  const int lnao = o->mol_list ? o->list_nao_max : b->nao_max, lnsh = o->mol_list ? o->list_nsh_max : b->nsh_max,
            lnat = o->mol_list ? o->list_nat_max : b->nat_max;
  cudaStream_t st = (cudaStream_t)stream;
  double* wk = (double*)work;
  const int mode = o->use_smem;
  if (mode == 1)
    return launch_mode<1>(b, o, nblocks, lnao, lnsh, lnat, S, H0, gamma, nel_ab, q0_at, wk, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo,
                          occ, iterations, status, P, W, resp, st);
  // global-memory and hybrid variants at two CTAs per SM (secondary build, 64 registers): one molecule's latency-bound
  // sub-problems overlap the other's tensor-core passes.  Round 2 measurement (tools/ab_variants.py, 1024 capsaicin
  // conformers, nao 142): 4.50 k SP/s at 2 CTAs/SM against 5.52 k at 1 CTA/SM (128 registers, no spills, the matrices of 148
  // instead of 296 molecules in L2) -- so it is off unless DXTB_B200_2CTA is set.
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  static const bool want_2cta = getenv("DXTB_B200_2CTA") != nullptr;  // developer A/B knob
  const bool two = want_2cta && 2 * nblocks >= 3 * n_sm && mode_smem_bytes(mode, lnao, lnsh, lnat) <= XTB_SMEM_2CTA;
  if (two)
    return xtb_scf_launch_2cta(mode, b, o, nblocks, lnao, lnsh, lnat, S, H0, gamma, nel_ab, q0_at, wk, q_orb, q_sh, q_at, v_orb, e_atom,
                               fenergy, emo, occ, iterations, status, P, W, resp, st);
  if (mode == 2)
    return launch_mode<2>(b, o, nblocks, lnao, lnsh, lnat, S, H0, gamma, nel_ab, q0_at, wk, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo,
                          occ, iterations, status, P, W, resp, st);
  return launch_mode<0>(b, o, nblocks, lnao, lnsh, lnat, S, H0, gamma, nel_ab, q0_at, wk, q_orb, q_sh, q_at, v_orb, e_atom, fenergy, emo,
                        occ, iterations, status, P, W, resp, st);
}
#endif  // XTB_SECONDARY
