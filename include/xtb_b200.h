/*
 * xtb_b200.h -- C ABI of the B200-native GFN1-xTB single-point hot path.
 *
 * Every entry point replaces one stage of dxtb's Python hot path (paths relative to
 * /root/reference/src/dxtb/_src, dxtb v0.4.0).  The reference has no FFI of its own (it is pure
 * Python/PyTorch); the binding a maintainer would add is the ctypes stub in INTEGRATION.md, which
 * is exactly what dxtb_b200/_abi.py does.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; the library never allocates
 *     persistent memory: every buffer (incl. workspaces) is owned by the caller (torch tensors)
 *   - all floating point data is fp64, indices int32, matrix offsets int64
 *   - molecules are stored ragged (CSR): atom a of molecule m is at_off[m] + a, etc.;
 *     matrices S/H0/P/W of molecule m start at mat_off[m] (row-major nao x nao), gamma at gam_off[m]
 *   - functions return 0 on success, <0 on argument errors, >0 = cudaError_t of the launch
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and never synchronise
 */
#ifndef XTB_B200_H
#define XTB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XTB_ATPAR 12 /* doubles per atom in at_par   */
#define XTB_SHPAR 6  /* doubles per shell in sh_par  */
#define XTB_CGTO 16  /* doubles per CGTO table entry */
#define XTB_MAXPRIM 7

/* at_par[a][..] : replaces the per-species gathers of xtb/base.py:138-155, repulsion/base.py:229-243,
 * thirdorder.py:232-244, halogen/hal.py:164-165 and the EEQ parameter lookup of tad-multicharge */
enum { XTB_AT_RAD = 0, XTB_AT_RCOV, XTB_AT_EN, XTB_AT_AREP, XTB_AT_ZEFF, XTB_AT_GAM3, XTB_AT_XBOND,
       XTB_AT_EEQ_CHI, XTB_AT_EEQ_ETA, XTB_AT_EEQ_KCN, XTB_AT_EEQ_RAD, XTB_AT_R4R2 /* D3: sqrt(Z) r4/r2 */ };
/* sh_par[s][..] : xtb/base.py:141-155 (levels, kcn in Hartree; shpoly; refocc), secondorder.py:838-839 (eta) */
enum { XTB_SH_LEVEL = 0, XTB_SH_KCN, XTB_SH_SHPOLY, XTB_SH_ETA, XTB_SH_REFOCC, XTB_SH_PAD };

/* Immutable batch descriptor (the IndexHelper + parameter gathers of basis/indexhelper.py:336-491,
 * built once per calculator by the host and uploaded). */
typedef struct xtb_batch {
  int32_t nb;                         /* molecules in this shard */
  int32_t nat_tot, nsh_tot, nao_tot;  /* totals over the shard */
  int32_t nat_max, nsh_max, nao_max;  /* maxima over the shard */
  int32_t nspecies;                   /* unique elements in the shard */
  int32_t ncgto;                      /* rows of the cgto table */
  int32_t has_xb;                     /* 1: some atom has a non-zero halogen-bond strength (Br, I, At with "hal" not excluded) */
  int32_t nsh_l_max[4];               /* maxima over the shard of the number of s / p / d shells of a molecule ([3] unused): size the
                                         (l_i, l_j) class launches of the integral / gradient pair kernels, empty classes are skipped */
  int64_t mat_total, gam_total, eeq_total; /* host copies of mat_off[nb], gam_off[nb], eeq_off[nb] */
  const int32_t* at_off;  /* [nb+1] */
  const int32_t* sh_off;  /* [nb+1] */
  const int32_t* ao_off;  /* [nb+1] */
  const int64_t* mat_off; /* [nb+1] prefix of nao^2 */
  const int64_t* gam_off; /* [nb+1] prefix of nsh^2 */
  const int64_t* eeq_off; /* [nb+1] prefix of (nat+1)^2 */
  const int32_t* at_z;       /* [nat_tot] atomic number */
  const int32_t* at_species; /* [nat_tot] index into kpair table */
  const int32_t* at_sh0;     /* [nat_tot] first shell (molecule-local) */
  const int32_t* at_nsh;     /* [nat_tot] number of shells */
  const double* at_par;      /* [nat_tot][XTB_ATPAR] */
  const int32_t* sh_atom;    /* [nsh_tot] atom (molecule-local) */
  const int32_t* sh_l;       /* [nsh_tot] angular momentum */
  const int32_t* sh_ao;      /* [nsh_tot] first AO (molecule-local) */
  const int32_t* sh_cgto;    /* [nsh_tot] row of cgto table */
  const int32_t* sh_type;    /* [nsh_tot] l + 3*(non-valence) : row of hscale */
  const int32_t* sh_by_l;    /* [nsh_tot] molecule-local shell ids sorted by l */
  const int32_t* nsh_l;      /* [nb][3] number of s/p/d shells */
  const double* sh_par;      /* [nsh_tot][XTB_SHPAR] */
  const int32_t* ao_sh;      /* [nao_tot] shell (molecule-local) */
  const double* cgto;        /* [ncgto][XTB_CGTO]: nprim, alpha[7], coeff[7], 0 */
  const double* kpair;       /* [nspecies][nspecies] */
  double hscale[36];         /* [6][6] by sh_type (xtb/gfn1.py:66-165) */
  double enscale;            /* hamiltonian.xtb.enscale */
  double rep_kexp;           /* repulsion.effective.kexp */
  double xb_damp, xb_rscale; /* halogen.classical */
  double gexp;               /* charge.effective.gexp (only 2.0 is implemented) */
  double int_cutoff, rep_cutoff, xb_cutoff, cn_cutoff;
  double kcn_d3;             /* 16.0 */
  /* D3(BJ) dispersion (tad-dftd3; dispersion/d3.py:93-211).  NULL tables = dispersion excluded. */
  const double* d3_refcn;    /* [nspecies][7] reference CNs, -1 = missing */
  const double* d3_c6;       /* [nspecies][nspecies][7][7] reference C6 */
  double d3_s6, d3_s8, d3_a1, d3_a2, d3_cutoff, d3_wf;
} xtb_batch;

/* SCF options: resolved defaults of constants/defaults.py */
typedef struct xtb_scf_opts {
  int32_t maxiter;          /* 100 */
  int32_t mixer;            /* 0 = Anderson (default path), 1 = simple */
  int32_t generations;      /* 5 */
  int32_t soft_start;       /* 1 */
  int32_t fermi_maxiter;    /* 200 */
  int32_t want_density;     /* 1: write P, W (needed by xtb_grad_bwd); 2: the same + workspace for the SCF response (resp != NULL) */
  int32_t use_smem;         /* kernel variant: 1 = C, A, X matrices in shared memory; 2 = hybrid (A in shared memory, C and X in the
                               workspace); 0 = all in the workspace.  The host checks capacity (xtb_scf_smem_bytes_mode) */
  int32_t jacobi_max_sweeps;/* 30 */
  int32_t subspace;         /* 1: intermediate SCF map evaluations of closed-shell molecules with a certified gap >= subspace_gap kT
                               take the occupied-subspace (Riccati) solve instead of the full eigendecomposition; 0: always full */
  int32_t subspace_maxiter; /* 16: fixed-point iterations before falling back to a Jacobi sweep */
  double damp;              /* 0.5 */
  double damp_init;         /* 0.1 */
  double diag_offset;       /* 0.01 */
  double x_atol;            /* 1e-4 (L2) */
  double x_atol_max;        /* 1e-5 (Linf) */
  double kt;                /* fermi_etemp * KELVIN2AU */
  double fermi_thresh;      /* sqrt(eps) */
  double jacobi_tol;        /* 1e-13: max |off-diagonal| of the final solve */
  double jacobi_tol_iter;   /* 2e-9: the same for intermediate SCF map evaluations */
  double subspace_tol;      /* 1e-10: max |Riccati residual| (Eh) of the occupied-subspace solve */
  double subspace_gap;      /* 50: certified HOMO-LUMO gap in units of kT below which occupations are not taken as integer */
  /* optional size bucket: this launch handles molecules mol_list[0..list_len) only (device pointer; NULL = all).
   * list_*_max are the maxima over the bucket (they size the shared-memory layout). */
  const int32_t* mol_list;
  int32_t list_len, list_nao_max, list_nsh_max, list_nat_max;
  /* 1: the launch is a batch of equally sized molecules (conformers): one persistent CTA per SM walks the list with a fixed
   * stride and starts every molecule after its first from the eigenvectors of the previous one, re-orthonormalised against the
   * new overlap (Newton-Schulz), instead of the Cholesky start basis -- the first solve then needs 1-2 Jacobi sweeps, not 4-5.
   * Results are those of the Cholesky start to solver tolerance (same iteration counts); falls back per molecule. */
  int32_t persistent, reserved0;
} xtb_scf_opts;

/* status bits written per molecule by xtb_scf_run */
#define XTB_STATUS_SCF_NOT_CONVERGED 1
#define XTB_STATUS_FERMI_FAILED 2
#define XTB_STATUS_JACOBI_NOT_CONVERGED 4
#define XTB_STATUS_S_NOT_POSDEF 8
#define XTB_STATUS_SWEEPS_SHIFT 8 /* bits 8..: total Jacobi sweeps of the molecule (diagnostic) */

int xtb_version(void);

/* sizeof checks for the ctypes mirror */
int xtb_sizeof_batch(void);
int xtb_sizeof_scf_opts(void);

/* Classical, geometry-only stage: D3-type coordination number (tad-mctc cn_d3/exp_count, called at
 * xtb/gfn1.py:57-64), repulsion (classicals/repulsion/base.py:269-334) and halogen-bond correction
 * (classicals/halogen/hal.py:209-364).  pos [nat_tot][3]; cn, e_rep, e_xb [nat_tot]. */
int xtb_geometry_fwd(const xtb_batch* b, const double* pos, double* cn, double* e_rep, double* e_xb, void* stream);

/* EEQ guess charges (tad-multicharge get_eeq_charges, called at scf/guess.py:118-120).
 * chrg [nb]; work [eeq_off[nb] + 2*(nat_tot+nb)]; q_at [nat_tot]. */
int xtb_eeq_guess(const xtb_batch* b, const double* pos, const double* chrg, double* work, double* q_at, void* stream);

/* Shell-resolved Coulomb matrix (coulomb/secondorder.py:799-870). gamma [gam_off[nb]]. */
int xtb_gamma_fwd(const xtb_batch* b, const double* pos, double* gamma, void* stream);

/* Overlap (integral/driver/pytorch/impls/overlap.py:147-243 + md/explicit.py:61-187) fused with the
 * GFN1 H0 build (xtb/base.py:251-360).  S, H0 [mat_off[nb]]. */
int xtb_overlap_h0_fwd(const xtb_batch* b, const double* pos, const double* cn, double* S, double* H0, void* stream);

/* Bytes of workspace xtb_scf_run needs for this batch / option set. */
int64_t xtb_scf_workspace_bytes(const xtb_batch* b, const xtb_scf_opts* o);
/* Diagnostic (bench.py): SM clock in MHz measured on the device over ~20 us (clock64 / globaltimer), written to the
   device double *mhz_out; no reference counterpart. */
int xtb_clock_probe(double* mhz_out, void* stream);

/* EEQ guess of one molecule with nat >= XTB_EEQ_LARGE_NAT on the whole device (xtb_eeq_guess skips those): same system
   and elimination order, every stage a grid-wide kernel.  Replaces the same reference call as xtb_eeq_guess
   (scf/guess.py:118-120 -> tad-multicharge get_eeq_charges).  at_off / eeq_off = at_off[mol] / eeq_off[mol]. */
#define XTB_EEQ_LARGE_NAT 256
int xtb_eeq_guess_large(const xtb_batch* b, int32_t mol, int32_t nat, int64_t at_off, int64_t eeq_off, const double* pos,
                        const double* chrg, double* work, double* q_at, void* stream);

/* Large-system SCF (BASELINE config 4, e.g. sh3 with 3104 AOs): molecule `mol` of the batch runs on the WHOLE device
   as a sequence of grid-wide kernels (tensor-core GEMMs + two-level block Jacobi) driven from the host; synchronises
   `stream`.  Same inputs / outputs / status bits as xtb_scf_run; `mat_off` = offset of the molecule in S/H0/P/W (doubles);
   `work` holds xtb_scf_large_workspace_bytes(nao, nsh, nat, generations) bytes.  Replaces the same reference loop as
   xtb_scf_run (scf/unrolling/default.py:64-137 with scf/base.py:818-907) for systems the one-CTA kernel cannot hold.
   resp: as in xtb_scf_run (whole-batch array [nao_tot + nsh_tot], this molecule's slices are written), or NULL. */
int64_t xtb_scf_large_workspace_bytes(int32_t nao, int32_t nsh, int32_t nat, int32_t generations);
int xtb_scf_run_large(const xtb_batch* b, const xtb_scf_opts* o, int32_t mol, int32_t nao, int32_t nsh, int32_t nat, const double* S,
                      const double* H0, const double* gamma, const double* nel_ab, const double* q0_at, void* work, double* q_orb,
                      double* q_sh, double* q_at, double* v_orb, double* e_atom, double* fenergy, double* emo, double* occ,
                      int32_t* iterations, int32_t* status, double* P, double* W, double* resp, int64_t mat_off, void* stream);

/* Dynamic shared memory the SCF kernel needs with use_smem=1 (host compares with the device limit). */
int64_t xtb_scf_smem_bytes(const xtb_batch* b);
int64_t xtb_scf_smem_bytes_for(int32_t nao_max, int32_t nsh_max, int32_t nat_max);
/* Same for kernel variant `mode` (the use_smem values).  With DXTB_B200_2CTA set, variants 0 and 2 run two CTAs per SM when the
   launch has at least 1.5 molecules per SM and this is at most XTB_SMEM_2CTA (measured slower in round 2; off by default). */
int64_t xtb_scf_smem_bytes_mode(int32_t mode, int32_t nao_max, int32_t nsh_max, int32_t nat_max);
#define XTB_SMEM_2CTA (113 * 1024)
#define XTB_SMEM_LIMIT (227 * 1024) /* per-CTA opt-in shared memory on sm_100 */

/* The whole SCF (scf/iterator.py:51-144, scf/unrolling/default.py:71-136, scf/base.py:651-907,
 * mixer/anderson.py:163-317, wavefunction/filling.py:201-366): one CTA per molecule, no host round-trips.
 *   nel_ab [nb][2]  alpha/beta electron numbers; q0_at [nat_tot] guess atomic charges
 *   outputs: q_orb [nao_tot], q_sh [nsh_tot], q_at [nat_tot], v_orb [nao_tot] (potential of the final charges),
 *            e_atom [nat_tot] (electronic + ES2 + ES3 + G/nat), fenergy [nb], emo [nao_tot], occ [nao_tot],
 *            iterations [nb], status [nb]; P, W [mat_off[nb]] if want_density
 *   resp [nao_tot + nsh_tot] or NULL (needs want_density): first-order response of the SCF residual for the nuclear
 *            gradient.  The reference's forces are autograd through the unrolled SCF (calculators/types/autograd.py:80-201),
 *            i.e. they contain (v_out - v_in) . dq/dR of the not fully converged state; with resp != NULL the kernel solves
 *            the coupled-perturbed equations for it, ADDS the response densities to P and W, and writes
 *            resp[0..nao_tot) = v_out + K y (potential to hand to xtb_grad_bwd instead of v_orb) and
 *            resp[nao_tot..) = y_sh (ES2 cross term of xtb_grad_bwd).  See xtb_scf_core.cuh:scf_response. */
int xtb_scf_run(const xtb_batch* b, const xtb_scf_opts* o, const double* S, const double* H0, const double* gamma,
                const double* nel_ab, const double* q0_at, void* work, double* q_orb, double* q_sh, double* q_at,
                double* v_orb, double* e_atom, double* fenergy, double* emo, double* occ, int32_t* iterations,
                int32_t* status, double* P, double* W, double* resp, void* stream);

/* Analytic nuclear gradient of the converged single point (calculators/types/analytical.py:63-222,
 * xtb/gfn1.py:185-408, secondorder.py:873-926, repulsion/base.py:337-406, ncoord/utils.py:30-52):
 * grad[a] = ge[m] * dE_m/dR_a, ge [nb] being the upstream gradient of the molecular energies.
 * dedcn [nat_tot] and pairbuf [4 * gam_off[nb]] are scratch; d3w = weights of xtb_d3_fwd, or NULL when
 * dispersion is excluded.  y_sh [nsh_tot] = response shell charges of xtb_scf_run (resp + nao_tot), or NULL: adds
 * y . dV/dR|_q = sum y_s dgamma_st/dR q_t to the second-order term; v_orb is then resp[0..nao_tot) and P, W carry the
 * response densities.  No atomics: the result is bit-reproducible. */
int xtb_grad_bwd(const xtb_batch* b, const double* pos, const double* cn, const double* S, const double* P,
                 const double* W, const double* v_orb, const double* q_sh, const double* gamma, const double* ge,
                 const double* d3w, const double* y_sh, double* pairbuf, double* dedcn, double* grad, void* stream);

/* D3(BJ) two-body dispersion energy (tad-dftd3 dftd3: weight_references, atomic_c6, dispersion with
 * rational_damping; s9 = 0 for GFN1).  cn is the exp-count CN of xtb_geometry_fwd; d3w [nat_tot][14] receives
 * the Gaussian reference weights and their CN derivatives (input of xtb_grad_bwd); e_disp [nat_tot]. */
int xtb_d3_fwd(const xtb_batch* b, const double* pos, const double* cn, double* d3w, double* e_disp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XTB_B200_H */
