// McMurchie-Davidson contracted-Gaussian overlap per shell pair, fused with the GFN1 H0 build, and the
// matching analytic-derivative kernel.  One thread per shell pair of a fixed (l_i, l_j) class: the class
// is a template parameter, so the Hermite recursion, the cartesian accumulators and the cartesian->
// spherical transform are fully unrolled into registers; shells are visited in l-sorted order
// (sh_by_l) so every warp runs a single class with no divergence.  Compute-bound on the fp64 pipe
// (exp + FMA); the only HBM traffic is the write of S and H0 (16 B per matrix element).
#include "xtb_common.cuh"

using namespace xtb;

namespace {

template <int L> struct Cart;
template <> struct Cart<0> { static constexpr int n = 1; };
template <> struct Cart<1> { static constexpr int n = 3; };
template <> struct Cart<2> { static constexpr int n = 6; };

// cartesian exponents in the order of integral/driver/pytorch/impls/md/trafo.py:144-161
// (p: y,z,x; d: xx,yy,zz,xy,xz,yz), as a constexpr function so that unrolled loops index registers
template <int L> XTB_DEV constexpr int nlm(int m, int d) {
  if (L == 0) return 0;
  if (L == 1) return (m == 0 && d == 1) || (m == 1 && d == 2) || (m == 2 && d == 0) ? 1 : 0;
  // L == 2
  if (m < 3) return m == d ? 2 : 0;
  if (m == 3) return d < 2 ? 1 : 0;
  if (m == 4) return d != 1 ? 1 : 0;
  return d > 0 ? 1 : 0;
}

// E^{ij}_0 for one cartesian direction by the MD recursion (explicit.py:360-422, 533-622, 830-937):
//   E^{i+1,j}_t = x E^{ij}_{t-1} + A E^{ij}_t + (t+1) E^{ij}_{t+1},  same with B for j+1.
template <int IMAX, int JMAX> XTB_DEV void ecoef0(double x, double A, double B, double (&E0)[IMAX + 1][JMAX + 1]) {
  constexpr int TM = IMAX + JMAX + 2;
  double E[IMAX + 1][JMAX + 1][TM];
#pragma unroll
  for (int i = 0; i <= IMAX; ++i)
#pragma unroll
    for (int j = 0; j <= JMAX; ++j)
#pragma unroll
      for (int t = 0; t < TM; ++t) E[i][j][t] = 0.0;
  E[0][0][0] = 1.0;
#pragma unroll
  for (int i = 0; i <= IMAX; ++i) {
    if (i > 0) {
#pragma unroll
      for (int t = 0; t <= i; ++t)
        E[i][0][t] = (t > 0 ? x * E[i - 1][0][t - 1] : 0.0) + A * E[i - 1][0][t] + (t + 1) * E[i - 1][0][t + 1];
    }
#pragma unroll
    for (int j = 1; j <= JMAX; ++j) {
#pragma unroll
      for (int t = 0; t <= i + j; ++t)
        E[i][j][t] = (t > 0 ? x * E[i][j - 1][t - 1] : 0.0) + B * E[i][j - 1][t] + (t + 1) * E[i][j - 1][t + 1];
    }
  }
#pragma unroll
  for (int i = 0; i <= IMAX; ++i)
#pragma unroll
    for (int j = 0; j <= JMAX; ++j) E0[i][j] = E[i][j][0];
}

// cartesian -> spherical (trafo.py:63-79); d order [0, +1, -1, +2, -2]
template <int L> XTB_DEV void to_sph_left(const double* in, int ncol, double* out);
template <> XTB_DEV void to_sph_left<0>(const double* in, int ncol, double* out) {
  for (int c = 0; c < ncol; ++c) out[c] = in[c];
}
template <> XTB_DEV void to_sph_left<1>(const double* in, int ncol, double* out) {
  for (int c = 0; c < 3 * ncol; ++c) out[c] = in[c];
}
template <> XTB_DEV void to_sph_left<2>(const double* in, int ncol, double* out) {
  // md/trafo.py:37-38, 65-75: the reference builds TRAFO with torch.tensor(...) at the default dtype, so sqrt(3) and
  // sqrt(3)/2 are float32-rounded values cast to fp64 (replicated for parity; see also param.py slater_to_gauss)
  const double s3 = 1.7320507764816284, s34 = 0.8660253882408142;
  for (int c = 0; c < ncol; ++c) {
    const double xx = in[0 * ncol + c], yy = in[1 * ncol + c], zz = in[2 * ncol + c];
    const double xy = in[3 * ncol + c], xz = in[4 * ncol + c], yz = in[5 * ncol + c];
    out[0 * ncol + c] = -0.5 * xx - 0.5 * yy + zz;
    out[1 * ncol + c] = s3 * xz;
    out[2 * ncol + c] = s3 * yz;
    out[3 * ncol + c] = s34 * xx - s34 * yy;
    out[4 * ncol + c] = s3 * xy;
  }
}

// Cartesian block -> spherical block: out (2LI+1)x(2LJ+1) = T_i * in * T_j^T
template <int LI, int LJ> XTB_DEV void cart2sph(const double (&in)[Cart<LI>::n][Cart<LJ>::n], double (&out)[2 * LI + 1][2 * LJ + 1]) {
  constexpr int NCI = Cart<LI>::n, NCJ = Cart<LJ>::n, NI = 2 * LI + 1, NJ = 2 * LJ + 1;
  double t1[NI][NCJ];
  to_sph_left<LI>(&in[0][0], NCJ, &t1[0][0]);
  // transpose, transform the other index, transpose back
  double t2[NCJ][NI];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NCJ; ++j) t2[j][i] = t1[i][j];
  double t3[NJ][NI];
  to_sph_left<LJ>(&t2[0][0], NI, &t3[0][0]);
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) out[i][j] = t3[j][i];
  (void)NCI;
}

// Contracted overlap block (and optionally its derivative w.r.t. the centre of shell i).
// v = R_j - R_i is what the reference passes to md_explicit (impls/overlap.py:225-226).
template <int LI, int LJ, bool GRAD>
XTB_DEV void shell_pair_overlap(const double* __restrict__ gi, const double* __restrict__ gj, double vx, double vy, double vz,
                                double (&s)[2 * LI + 1][2 * LJ + 1], double (&ds)[GRAD ? 3 : 1][2 * LI + 1][2 * LJ + 1]) {
  constexpr int NCI = Cart<LI>::n, NCJ = Cart<LJ>::n;
  constexpr int IM = LI + (GRAD ? 1 : 0);
  double acc[NCI][NCJ];
  double dacc[GRAD ? 3 : 1][NCI][NCJ];
#pragma unroll
  for (int i = 0; i < NCI; ++i)
#pragma unroll
    for (int j = 0; j < NCJ; ++j) {
      acc[i][j] = 0.0;
#pragma unroll
      for (int d = 0; d < (GRAD ? 3 : 1); ++d) dacc[d][i][j] = 0.0;
    }
  const int npi = (int)gi[0], npj = (int)gj[0];
  const double r2 = vx * vx + vy * vy + vz * vz;
  for (int pa = 0; pa < npi; ++pa) {
    const double a = gi[1 + pa], ca = gi[1 + XTB_MAXPRIM + pa];
    for (int pb = 0; pb < npj; ++pb) {
      const double bb = gj[1 + pb], cb = gj[1 + XTB_MAXPRIM + pb];
      const double o = 1.0 / (a + bb);
      const double x = 0.5 * o;
      const double est = a * bb * o * r2;
      const double sij = exp(-est) * kSqrtPi3 * o * sqrt(o) * ca * cb;
      double E[3][IM + 1][LJ + 1];
      ecoef0<IM, LJ>(x, vx * bb * o, -vx * a * o, E[0]);
      ecoef0<IM, LJ>(x, vy * bb * o, -vy * a * o, E[1]);
      ecoef0<IM, LJ>(x, vz * bb * o, -vz * a * o, E[2]);
#pragma unroll
      for (int mi = 0; mi < NCI; ++mi) {
        const int ix = nlm<LI>(mi, 0), iy = nlm<LI>(mi, 1), iz = nlm<LI>(mi, 2);
#pragma unroll
        for (int mj = 0; mj < NCJ; ++mj) {
          const int jx = nlm<LJ>(mj, 0), jy = nlm<LJ>(mj, 1), jz = nlm<LJ>(mj, 2);
          const double ex = E[0][ix][jx], ey = E[1][iy][jy], ez = E[2][iz][jz];
          acc[mi][mj] += sij * ex * ey * ez;
          if (GRAD) {
            // F^{ij} = 2a E^{i+1,j} - i E^{i-1,j}   (explicit.py:196-203)
            const double two_a = 2.0 * a;
            double fx = two_a * E[0][ix + 1][jx];
            if (ix > 0) fx -= ix * E[0][ix - 1][jx];
            double fy = two_a * E[1][iy + 1][jy];
            if (iy > 0) fy -= iy * E[1][iy - 1][jy];
            double fz = two_a * E[2][iz + 1][jz];
            if (iz > 0) fz -= iz * E[2][iz - 1][jz];
            dacc[0][mi][mj] += sij * fx * ey * ez;
            dacc[1][mi][mj] += sij * ex * fy * ez;
            dacc[2][mi][mj] += sij * ex * ey * fz;
          }
        }
      }
    }
  }
  cart2sph<LI, LJ>(acc, s);
  if (GRAD) {
#pragma unroll
    for (int d = 0; d < 3; ++d) cart2sph<LI, LJ>(dacc[d], ds[d]);
  }
}

struct PairInfo {
  int I, J, A, B;       // molecule-local shell / atom ids
  int aoI, aoJ;         // first AOs
  double vx, vy, vz, dist;
  bool valid;
};

template <int LI, int LJ> XTB_DEV PairInfo pair_setup(const xtb_batch& b, int m, const double* __restrict__ pos) {
  PairInfo pi;
  pi.valid = false;
  const int* cnt = b.nsh_l + 3 * m;
  const int nI = cnt[LI], nJ = cnt[LJ];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nI * nJ) return pi;
  const int ia = t / nJ, ja = t - ia * nJ;
  if (LI == LJ && ja >= ia) return pi;
  const int s0 = b.sh_off[m];
  const int offI = (LI > 0 ? cnt[0] : 0) + (LI > 1 ? cnt[1] : 0);
  const int offJ = (LJ > 0 ? cnt[0] : 0) + (LJ > 1 ? cnt[1] : 0);
  pi.I = b.sh_by_l[s0 + offI + ia];
  pi.J = b.sh_by_l[s0 + offJ + ja];
  pi.A = b.sh_atom[s0 + pi.I];
  pi.B = b.sh_atom[s0 + pi.J];
  if (pi.A == pi.B) return pi;
  const int a0 = b.at_off[m];
  const double* pa = pos + 3 * (size_t)(a0 + pi.A);
  const double* pb = pos + 3 * (size_t)(a0 + pi.B);
  // v = -(pos_i - pos_j)
  pi.vx = pb[0] - pa[0];
  pi.vy = pb[1] - pa[1];
  pi.vz = pb[2] - pa[2];
  pi.dist = safe_dist(pi.vx, pi.vy, pi.vz);
  if (!(pi.dist < b.int_cutoff && pi.dist > 0.1)) return pi;
  pi.aoI = b.sh_ao[s0 + pi.I];
  pi.aoJ = b.sh_ao[s0 + pi.J];
  pi.valid = true;
  return pi;
}

// Off-atom shell-pair factor Pi * K of xtb/base.py:294-334 and the mean self energy :339-343.
struct H0Factors {
  double var_pi, var_k, hmean, rr, tmp_a, tmp_b, shp_a, shp_b, kcn_a, kcn_b;
};

XTB_DEV H0Factors h0_factors(const xtb_batch& b, int m, const PairInfo& pi, const double* __restrict__ cn) {
  H0Factors f;
  const int s0 = b.sh_off[m], a0 = b.at_off[m];
  const double* spa = b.sh_par + (size_t)(s0 + pi.I) * XTB_SHPAR;
  const double* spb = b.sh_par + (size_t)(s0 + pi.J) * XTB_SHPAR;
  const double* apa = b.at_par + (size_t)(a0 + pi.A) * XTB_ATPAR;
  const double* apb = b.at_par + (size_t)(a0 + pi.B) * XTB_ATPAR;
  f.kcn_a = spa[XTB_SH_KCN];
  f.kcn_b = spb[XTB_SH_KCN];
  const double ha = spa[XTB_SH_LEVEL] - f.kcn_a * cn[a0 + pi.A];
  const double hb = spb[XTB_SH_LEVEL] - f.kcn_b * cn[a0 + pi.B];
  f.hmean = 0.5 * (ha + hb);
  f.rr = sqrt(pi.dist / (apa[XTB_AT_RAD] + apb[XTB_AT_RAD]));
  f.shp_a = spa[XTB_SH_SHPOLY];
  f.shp_b = spb[XTB_SH_SHPOLY];
  f.tmp_a = 1.0 + f.shp_a * f.rr;
  f.tmp_b = 1.0 + f.shp_b * f.rr;
  f.var_pi = f.tmp_a * f.tmp_b;
  const int ta = b.sh_type[s0 + pi.I], tb = b.sh_type[s0 + pi.J];
  const double hs = b.hscale[ta * 6 + tb];
  if (ta < 3 && tb < 3) {
    const double den = apa[XTB_AT_EN] - apb[XTB_AT_EN];
    const double kp = b.kpair[b.at_species[a0 + pi.A] * b.nspecies + b.at_species[a0 + pi.B]];
    f.var_k = hs * kp * (1.0 + b.enscale * den * den);
  } else {
    f.var_k = hs;
  }
  return f;
}

// Basis data staged in shared memory: the contracted-Gaussian table of the batch (one 16-double row per unique (element,
// shell): nprim, exponents, coefficients) is copied once per CTA; every thread then reads its two rows from there in the
// primitive double loop.  Falls back to the global table if the batch has more than kCgtoSmemRows unique shells.
constexpr int kCgtoSmemRows = 96;  // 12 kB

XTB_DEV const double* stage_cgto(const xtb_batch& b, double* s_cgto) {
  if (b.ncgto > kCgtoSmemRows) return b.cgto;
  for (int t = threadIdx.x; t < b.ncgto * XTB_CGTO; t += blockDim.x) s_cgto[t] = b.cgto[t];
  __syncthreads();
  return s_cgto;
}

template <int LI, int LJ>
__global__ void __launch_bounds__(128) k_overlap_h0(const xtb_batch b, const double* __restrict__ pos, const double* __restrict__ cn,
                                                    double* __restrict__ S, double* __restrict__ H0, int mol0) {
  __shared__ double s_cgto[kCgtoSmemRows * XTB_CGTO];
  const double* cgto = stage_cgto(b, s_cgto);
  const int m = mol0 + blockIdx.y;
  const PairInfo pi = pair_setup<LI, LJ>(b, m, pos);
  if (!pi.valid) return;
  const int s0 = b.sh_off[m];
  const double* gi = cgto + (size_t)b.sh_cgto[s0 + pi.I] * XTB_CGTO;
  const double* gj = cgto + (size_t)b.sh_cgto[s0 + pi.J] * XTB_CGTO;
  double s[2 * LI + 1][2 * LJ + 1];
  double dummy[1][2 * LI + 1][2 * LJ + 1];
  shell_pair_overlap<LI, LJ, false>(gi, gj, pi.vx, pi.vy, pi.vz, s, dummy);
  const H0Factors f = h0_factors(b, m, pi, cn);
  const double hsh = f.var_pi * f.var_k * f.hmean;
  const int n = b.ao_off[m + 1] - b.ao_off[m];
  double* Sm = S + b.mat_off[m];
  double* Hm = H0 + b.mat_off[m];
#pragma unroll
  for (int r = 0; r < 2 * LI + 1; ++r)
#pragma unroll
    for (int c = 0; c < 2 * LJ + 1; ++c) {
      const double v = s[r][c];
      const size_t ij = (size_t)(pi.aoI + r) * n + pi.aoJ + c, ji = (size_t)(pi.aoJ + c) * n + pi.aoI + r;
      Sm[ij] = v;
      Sm[ji] = v;
      Hm[ij] = v * hsh;
      Hm[ji] = v * hsh;
    }
}

// diagonal: S = 1 (impls/overlap.py:241-242), H0 = self energy (xtb/base.py:287-292)
__global__ void k_diag(const xtb_batch b, const double* __restrict__ cn, double* __restrict__ S, double* __restrict__ H0, int mol0) {
  const int m = mol0 + blockIdx.y;
  const int o0 = b.ao_off[m], n = b.ao_off[m + 1] - o0;
  const int s0 = b.sh_off[m], a0 = b.at_off[m];
  for (int mu = blockIdx.x * blockDim.x + threadIdx.x; mu < n; mu += gridDim.x * blockDim.x) {
    const int sh = b.ao_sh[o0 + mu];
    const double* sp = b.sh_par + (size_t)(s0 + sh) * XTB_SHPAR;
    S[b.mat_off[m] + (size_t)mu * n + mu] = 1.0;
    H0[b.mat_off[m] + (size_t)mu * n + mu] = sp[XTB_SH_LEVEL] - sp[XTB_SH_KCN] * cn[a0 + b.sh_atom[s0 + sh]];
  }
}

// Shell-pair part of the analytic gradient (xtb/gfn1.py:311-408): overlap-derivative term with
// sval = 2(P*Hsh - W) - P (v_mu + v_nu), the dPi/dR term, and dE/dCN.
template <int LI, int LJ>
__global__ void __launch_bounds__(128) k_grad_pair(const xtb_batch b, const double* __restrict__ pos, const double* __restrict__ cn,
                                                   const double* __restrict__ P, const double* __restrict__ W,
                                                   const double* __restrict__ v_orb, double* __restrict__ pairbuf, int mol0) {
  __shared__ double s_cgto[kCgtoSmemRows * XTB_CGTO];
  const double* cgto = stage_cgto(b, s_cgto);
  const int m = mol0 + blockIdx.y;
  const PairInfo pi = pair_setup<LI, LJ>(b, m, pos);
  if (!pi.valid) return;
  const int s0 = b.sh_off[m], a0 = b.at_off[m], o0 = b.ao_off[m];
  const double* gi = cgto + (size_t)b.sh_cgto[s0 + pi.I] * XTB_CGTO;
  const double* gj = cgto + (size_t)b.sh_cgto[s0 + pi.J] * XTB_CGTO;
  double s[2 * LI + 1][2 * LJ + 1];
  double ds[3][2 * LI + 1][2 * LJ + 1];
  shell_pair_overlap<LI, LJ, true>(gi, gj, pi.vx, pi.vy, pi.vz, s, ds);
  const H0Factors f = h0_factors(b, m, pi, cn);
  const double hsh = f.var_pi * f.var_k * f.hmean;
  const int n = b.ao_off[m + 1] - o0;
  const double* Pm = P + b.mat_off[m];
  const double* Wm = W + b.mat_off[m];
  double gx = 0.0, gy = 0.0, gz = 0.0, ps = 0.0;
#pragma unroll
  for (int r = 0; r < 2 * LI + 1; ++r)
#pragma unroll
    for (int c = 0; c < 2 * LJ + 1; ++c) {
      const size_t ij = (size_t)(pi.aoI + r) * n + pi.aoJ + c;
      const double p = Pm[ij];
      const double sval = 2.0 * (p * hsh - Wm[ij]) - p * (v_orb[o0 + pi.aoI + r] + v_orb[o0 + pi.aoJ + c]);
      gx += sval * ds[0][r][c];
      gy += sval * ds[1][r][c];
      gz += sval * ds[2][r][c];
      ps += p * s[r][c];
    }
  // dPi/dR term: 2 * (P.H.S)_sh * dPi/Pi * (R_A - R_B);  R_A - R_B = -v
  const double dvar_pi = (f.tmp_a * f.shp_b + f.tmp_b * f.shp_a) * f.rr * 0.5 / (pi.dist * pi.dist);
  const double dpi = 2.0 * ps * hsh * dvar_pi / f.var_pi;
  gx += dpi * (-pi.vx);
  gy += dpi * (-pi.vy);
  gz += dpi * (-pi.vz);
  // Deterministic accumulation: every shell pair owns the slot (hi, lo) = (max(I,J), min(I,J)) of an nsh x nsh
  // table; k_grad_atoms sums the slots of an atom in a fixed order (no atomics, bit-reproducible forces).
  // Stored: derivative w.r.t. the atom of shell `hi`, and Pi*K*(P.S)_sh for dE/dCN.
  const int ns = b.sh_off[m + 1] - s0;
  const double sg = pi.I > pi.J ? 1.0 : -1.0;
  const int hi = pi.I > pi.J ? pi.I : pi.J, lo = pi.I > pi.J ? pi.J : pi.I;
  double* slot = pairbuf + 4 * ((size_t)b.gam_off[m] + (size_t)hi * ns + lo);
  slot[0] = sg * gx;
  slot[1] = sg * gy;
  slot[2] = sg * gz;
  slot[3] = ps * f.var_pi * f.var_k;
  (void)a0;
}

template <int LI, int LJ> int launch_overlap(const xtb_batch* b, const double* pos, const double* cn, double* S, double* H0, cudaStream_t st) {
  const int nt = 128;
  const int npair = b->nsh_l_max[LI] * b->nsh_l_max[LJ];  // upper bound on n_I * n_J of this class over the shard
  if (npair == 0) return 0;                               // e.g. no d shells anywhere in the shard
  for (int m0 = 0; m0 < b->nb; m0 += kMaxGridY) {  // gridDim.y is capped at 65535
    dim3 grid((npair + nt - 1) / nt, b->nb - m0 < kMaxGridY ? b->nb - m0 : kMaxGridY);
    k_overlap_h0<LI, LJ><<<grid, nt, 0, st>>>(*b, pos, cn, S, H0, m0);
  }
  return launch_status();
}
template <int LI, int LJ>
int launch_grad_pair(const xtb_batch* b, const double* pos, const double* cn, const double* P, const double* W, const double* v,
                     double* pairbuf, cudaStream_t st) {
  const int nt = 128;
  const int npair = b->nsh_l_max[LI] * b->nsh_l_max[LJ];
  if (npair == 0) return 0;
  for (int m0 = 0; m0 < b->nb; m0 += kMaxGridY) {
    dim3 grid((npair + nt - 1) / nt, b->nb - m0 < kMaxGridY ? b->nb - m0 : kMaxGridY);
    k_grad_pair<LI, LJ><<<grid, nt, 0, st>>>(*b, pos, cn, P, W, v, pairbuf, m0);
  }
  return launch_status();
}

}  // namespace

int xtb_launch_grad_atoms(const xtb_batch* b, const double* pos, const double* P, const double* q_sh, const double* y_sh,
                          const double* gamma, const double* pairbuf, double* dedcn, const double* ge, double* grad, cudaStream_t st);
int xtb_launch_d3_grad(const xtb_batch* b, const double* pos, const double* d3w, const double* ge, double* dedcn, double* grad,
                       cudaStream_t st);

extern "C" int xtb_overlap_h0_fwd(const xtb_batch* b, const double* pos, const double* cn, double* S, double* H0, void* stream) {
  if (!b || !pos || !cn || !S || !H0) return -1;
  if (b->nb == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // same-atom blocks and pairs beyond the cutoff stay zero (impls/overlap.py:186-194)
  cudaMemsetAsync(S, 0, sizeof(double) * (size_t)b->mat_total, st);
  cudaMemsetAsync(H0, 0, sizeof(double) * (size_t)b->mat_total, st);
  int rc;
  if ((rc = launch_overlap<0, 0>(b, pos, cn, S, H0, st))) return rc;
  if ((rc = launch_overlap<1, 0>(b, pos, cn, S, H0, st))) return rc;
  if ((rc = launch_overlap<1, 1>(b, pos, cn, S, H0, st))) return rc;
  if ((rc = launch_overlap<2, 0>(b, pos, cn, S, H0, st))) return rc;
  if ((rc = launch_overlap<2, 1>(b, pos, cn, S, H0, st))) return rc;
  if ((rc = launch_overlap<2, 2>(b, pos, cn, S, H0, st))) return rc;
  int gx = (b->nao_max + 127) / 128;
  for (int m0 = 0; m0 < b->nb; m0 += kMaxGridY)
    k_diag<<<dim3(gx, b->nb - m0 < kMaxGridY ? b->nb - m0 : kMaxGridY), 128, 0, st>>>(*b, cn, S, H0, m0);
  return launch_status();
}

extern "C" int xtb_grad_bwd(const xtb_batch* b, const double* pos, const double* cn, const double* S, const double* P,
                            const double* W, const double* v_orb, const double* q_sh, const double* gamma, const double* ge,
                            const double* d3w, const double* y_sh, double* pairbuf, double* dedcn, double* grad, void* stream) {
  (void)S;
  if (!b || !pos || !cn || !P || !W || !v_orb || !q_sh || !gamma || !ge || !pairbuf || !dedcn || !grad) return -1;
  if (b->nb == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dedcn, 0, sizeof(double) * (size_t)b->nat_tot, st);
  cudaMemsetAsync(grad, 0, sizeof(double) * 3 * (size_t)b->nat_tot, st);
  cudaMemsetAsync(pairbuf, 0, sizeof(double) * 4 * (size_t)b->gam_total, st);
  int rc;
  if ((rc = launch_grad_pair<0, 0>(b, pos, cn, P, W, v_orb, pairbuf, st))) return rc;
  if ((rc = launch_grad_pair<1, 0>(b, pos, cn, P, W, v_orb, pairbuf, st))) return rc;
  if ((rc = launch_grad_pair<1, 1>(b, pos, cn, P, W, v_orb, pairbuf, st))) return rc;
  if ((rc = launch_grad_pair<2, 0>(b, pos, cn, P, W, v_orb, pairbuf, st))) return rc;
  if ((rc = launch_grad_pair<2, 1>(b, pos, cn, P, W, v_orb, pairbuf, st))) return rc;
  if ((rc = launch_grad_pair<2, 2>(b, pos, cn, P, W, v_orb, pairbuf, st))) return rc;
  // dispersion: direct part into grad, dE/dCN into dedcn (the exp-count CN is shared with H0); one writer per atom
  if (d3w && (rc = xtb_launch_d3_grad(b, pos, d3w, ge, dedcn, grad, st))) return rc;
  return xtb_launch_grad_atoms(b, pos, P, q_sh, y_sh, gamma, pairbuf, dedcn, ge, grad, st);
}
