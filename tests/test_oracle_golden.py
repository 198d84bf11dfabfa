"""Pin the CPU oracle against the reference's own golden vectors.

Sources: test/test_overlap/overlap.npz, test/test_hamiltonian/h0.npz (float32), test/test_scf/samples.py and
test/test_singlepoint/samples.py (tblite fp64 literals), test/test_scf/grad.npz (float32), test/test_scf/test_guess.py.
tblite converts eV with 1 Eh = 27.21138505 eV; with that constant the oracle reproduces tblite's SCF energies
to <= 5e-9 Eh, which pins every electronic-structure stage of the path.
"""
import numpy as np
import pytest

from oracle import gfn1_oracle as O

TBLITE_EV2AU = 1.0 / 27.21138505
TIGHT = dict(x_atol=1e-10, x_atol_max=1e-10)


@pytest.fixture()
def tblite_units():
    par = O.params()
    old = par.ev2au
    par.ev2au = TBLITE_EV2AU
    yield
    par.ev2au = old


def _geom(mols, name):
    m = mols[name]
    return np.array(m["numbers"]), np.array(m["positions"]), m["charge"]


@pytest.mark.parametrize("name", ["H2", "LiH", "H2O", "CH4", "SiH4", "LYS_xao"])
def test_overlap_and_h0_goldens(mols, goldens, name, tblite_units):
    nums, pos, _ = _geom(mols, name)
    m = O.make_mol(nums)
    S, _ = O.overlap(m, pos)
    H = O.h0(m, pos, S)
    ref_s, ref_h = goldens[f"overlap/{name}"], goldens[f"h0/{name}"]
    # goldens are float32: tolerance = float32 rounding of O(1) numbers
    assert np.abs(S - ref_s).max() < 1e-7
    assert np.abs(H - ref_h).max() < 1e-7
    assert np.allclose(S, S.T, atol=0) and np.allclose(np.diag(S), 1.0)


@pytest.mark.parametrize("name", ["H", "H2", "LiH", "H2O", "CH4", "SiH4", "MB16_43_01", "LYS_xao"])
def test_scf_energy_tblite(mols, energies, name, tblite_units):
    nums, pos, chrg = _geom(mols, name)
    r = O.singlepoint(nums, pos, chrg, opts=dict(exclude=("disp",), **TIGHT))
    assert r.converged
    assert abs(r.e_scf - energies["scf_gfn1_tblite"][name]) < 1e-9


def test_h_atom_total_energy_pins_ev2au(mols, energies, tblite_units):
    nums, pos, chrg = _geom(mols, "H")
    r = O.singlepoint(nums, pos, chrg, opts=dict(exclude=("disp",), guess="sad"))
    assert abs(r.energy - energies["total_gfn1_tblite"]["H"]) < 1e-14


@pytest.mark.parametrize("name", ["H2", "LiH", "CH4", "SiH4"])
def test_scf_gradient_tblite(mols, goldens, name, tblite_units):
    nums, pos, chrg = _geom(mols, name)
    r = O.singlepoint(nums, pos, chrg, opts=dict(exclude=("disp", "rep", "hal"), **TIGHT), grad=True)
    assert np.abs(r.gradient - goldens[f"scf_grad/{name}"]).max() < 2e-7  # float32 golden


def test_gradient_is_derivative_of_energy(mols):
    nums, pos, chrg = _geom(mols, "H2O")
    o = dict(exclude=("disp",), x_atol=1e-12, x_atol_max=1e-12)
    r = O.singlepoint(nums, pos, chrg, opts=o, grad=True)
    h = 1e-4
    for a, x in [(0, 2), (1, 0), (2, 2)]:
        p = pos.copy(); p[a, x] += h
        ep = O.singlepoint(nums, p, chrg, opts=o).energy
        p = pos.copy(); p[a, x] -= h
        em = O.singlepoint(nums, p, chrg, opts=o).energy
        assert abs((ep - em) / (2 * h) - r.gradient[a, x]) < 5e-8


def test_eeq_guess_known_answer(mols, energies):
    nums, pos, chrg = _geom(mols, "CH_guess")
    m = O.make_mol(nums)
    q = O.guess_orbital_charges(m, O.eeq_charges(m, pos, chrg))
    # reference tolerance (test/test_scf/test_guess.py:57-71)
    assert np.abs(q - np.array(energies["eeq_guess_CH"])).max() < 1e-5
    assert abs(q.sum()) < 1e-12


@pytest.mark.parametrize("name,tol", [("H2", 2e-7), ("LiH", 3e-7), ("H2O", 2e-7), ("CH4", 2e-7), ("SiH4", 2e-7),
                                      ("MB16_43_01", 2e-6), ("LYS_xao", 6e-6)])
def test_cn_derivative_golden(mols, goldens, name, tol):
    """dE/dCN of the converged density (`dedcn` of GFN1Hamiltonian.get_gradient, xtb/gfn1.py:185-408) and its CN chain
    rule term (`get_dcn(cn_d3_gradient, dedcn)`, ncoord/utils.py:30-52) against test/test_hamiltonian/grad_no_overlap.npz
    (`*_dedcn`, `*_dcn`; float32, generated at x_atol = 1e-6): pins the exp-count derivative and its sign conventions."""
    nums, pos, chrg = _geom(mols, name)
    m = O.make_mol(nums)
    cn, dcf = O.cn_d3(m, pos, grad=True)
    assert np.abs(dcf + dcf.transpose(1, 0, 2)).max() < 1e-14  # antisymmetry of the pair derivative
    r = O.singlepoint(nums, pos, float(chrg), opts={"exclude": ("disp",), "x_atol": 1e-6, "x_atol_max": 1e-6}, grad=True)
    assert np.abs(r.gradient_parts["h0_dedcn"] - goldens[f"dedcn/{name}"]).max() < tol
    assert np.abs(r.gradient_parts["h0_dcn"] - goldens[f"dcn/{name}"]).max() < tol


def test_anderson_matches_simple_for_soft_start():
    rng = np.random.default_rng(0)
    a = O.Anderson(7)
    x_old = rng.normal(size=7)
    for _ in range(5):
        x_new = rng.normal(size=7)
        mixed = a.iter(x_new, x_old)
        assert np.allclose(mixed, x_old + 0.1 * (x_new - x_old))
        x_old = mixed


def test_fermi_occupation_counts():
    emo = np.linspace(-1.0, 1.0, 12)
    occ = O.fermi_occupation(np.array([4.0, 3.0]), emo, 300 * O.params().kelvin2au)
    assert abs(occ[0].sum() - 4) < 1e-7 and abs(occ[1].sum() - 3) < 1e-7
    assert (np.diff(occ[0]) <= 1e-15).all()


def test_default_path_iterations_and_charges(mols):
    nums, pos, chrg = _geom(mols, "caffeine")
    r = O.singlepoint(nums, pos, chrg, opts=dict(exclude=("disp",)))
    assert r.converged and r.iterations == 13
    assert abs(r.q_at.sum() - chrg) < 1e-10


def test_d3_gradient_is_derivative_of_energy(mols):
    """D3(BJ) restatement: analytic gradient (direct + CN chain) vs finite differences, synthetic table."""
    tab = O.synthetic_d3_table()
    nums, pos, _ = _geom(mols, "CH4")
    m = O.make_mol(nums)
    cn, dcf = O.cn_d3(m, pos, grad=True)
    e, gd, dedcn = O.d3_dispersion(nums, pos, tab, cn=cn, grad=True)
    g = gd + (dcf * (dedcn[:, None] + dedcn[None, :])[:, :, None]).sum(1)
    h = 1e-5
    for a, x in [(0, 0), (1, 2), (4, 1)]:
        p = pos.copy(); p[a, x] += h
        ep = O.d3_dispersion(nums, p, tab)[0].sum()
        p = pos.copy(); p[a, x] -= h
        em = O.d3_dispersion(nums, p, tab)[0].sum()
        assert abs((ep - em) / (2 * h) - g[a, x]) < 1e-9
    assert e.sum() < 0


def test_atomic_scf_energies_all_elements(energies, tblite_units):
    """Neutral atoms Z = 1..86 (test/test_scf/test_elements_gfn1.py): pins the element parametrisation blob
    (levels, hardnesses, third-order terms, reference occupations) and the Fermi-smeared open-shell SCF.
    Mn converges to a different SCF solution of the d shell (the reference's own tolerance for atoms is 1e-2)."""
    import warnings

    refs = energies["scf_gfn1_tblite_atoms"]
    worst = 0.0
    for z in range(1, 87):
        if z == 25:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = O.singlepoint([z], [[0.0, 0.0, 0.0]], opts=dict(exclude=("disp",), x_atol=1e-9, x_atol_max=1e-9, maxiter=300, guess="sad"))
        worst = max(worst, abs(r.e_scf - refs[z - 1]))
    assert worst < 1e-7


def test_readme_batch_example_bounds_the_missing_dispersion_term():
    """examples/batch-2.py:83-93 prints dxtb's own default-path energies (10 digits, D3(BJ) included) for the formamide
    dimer / monomer.  The D3 reference table is third-party data, so the oracle can only be compared
    without dispersion: the implied D3 energies must be small, negative and attractive for the dimer -- a bound, not a pin
    (it catches any error of the SCF / repulsion part beyond ~1 mEh and a wrong sign or unit of the total)."""
    sym = {"C": 6, "N": 7, "H": 1, "O": 8}
    n1 = [sym[s] for s in "C C N N H H H H H H O O".split()]
    p1 = np.array([[-3.81469488143921, 0.09993441402912, 0], [3.81469488143921, -0.09993441402912, 0],
                   [-2.66030049324036, -2.15898251533508, 0], [2.66030049324036, 2.15898251533508, 0],
                   [-0.73178529739380, -2.28237795829773, 0], [-5.89039325714111, -0.02589114569128, 0],
                   [-3.71254944801331, -3.73605775833130, 0], [3.71254944801331, 3.73605775833130, 0],
                   [0.73178529739380, 2.28237795829773, 0], [5.89039325714111, 0.02589114569128, 0],
                   [-2.74426102638245, 2.16115570068359, 0], [2.74426102638245, -2.16115570068359, 0]])
    n2 = [sym[s] for s in "C O N H H H".split()]
    p2 = np.array([[-0.55569743203406, 1.09030425468557, 0], [0.51473634678469, 3.15152550263611, 0],
                   [0.59869690244446, -1.16861263789477, 0], [-0.45355203669134, -2.74568780438064, 0],
                   [2.52721209544999, -1.29200800956867, 0], [-2.63139587595376, 0.96447869452240, 0]])
    ed = []
    for n, p, ref in ((n1, p1, -23.2835232516), (n2, p2, -11.6302093800)):
        r = O.singlepoint(np.array(n), p, 0.0, opts={"exclude": ("disp",)})
        assert r.converged
        ed.append(ref - r.energy)  # implied D3(BJ) energy
    assert -6.0e-3 < ed[0] < -2.0e-3 and -2.0e-3 < ed[1] < -0.5e-3
    assert -3.0e-3 < ed[0] - 2 * ed[1] < -0.5e-3  # dispersion part of the dimer interaction energy is attractive


@pytest.mark.parametrize("name", ["H2", "H2O", "SiH4", "LYS_xao", "MB16_43_01"])
def test_repulsion_energy_vs_reference_literals(mols, energies, name):
    """Classical repulsion (components/classicals/repulsion/base.py:250-335) against the fp64 literals of
    test/test_classical/test_repulsion/samples.py (same geometries as test/test_singlepoint/mols)."""
    m = mols[name]
    r = O.singlepoint(np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"]), opts={"exclude": ("disp",), "maxiter": 1})
    assert abs(r.e_rep - energies["repulsion_gfn1"][name]) < 1e-13


@pytest.mark.parametrize("name", ["SiH4", "LiH"])
def test_es2_shell_energy_and_gradient_at_fixed_charges(mols, energies, name):
    """Shell-resolved second-order electrostatics (secondorder.py:374-382, 799-926; harmonic Hubbard average, gexp = 2):
    energy 1/2 q^T gamma q and its nuclear gradient at the fixed shell charges of test/test_coulomb/samples.py:576-660."""
    ref = energies["es2_shell_gfn1"][name]
    m = mols[name]
    mol = O.make_mol(np.array(m["numbers"]))
    pos = np.array(m["positions"])
    q = np.array(ref["q_shell"])
    assert q.shape == (mol.nsh,)

    def e(p):
        return 0.5 * q @ O.gamma_shell(mol, p) @ q

    if name != "LiH":  # the LiH energy literal is a 0.0 placeholder in the reference file
        assert abs(e(pos) - ref["es2"]) < 1e-14
    g = np.array(ref["grad"]).reshape(-1, 3)
    fd, h = np.zeros_like(pos), 1e-5
    for a in range(mol.nat):
        for x in range(3):
            pp, pm = pos.copy(), pos.copy()
            pp[a, x] += h
            pm[a, x] -= h
            fd[a, x] = (e(pp) - e(pm)) / (2 * h)
    assert np.abs(fd - g).max() < 1e-10


def test_es3_atom_energy_at_fixed_charges(mols, energies):
    """Third-order on-site electrostatics (thirdorder.py:246-276), 1/3 sum Gamma_A q_A^3, with the (float32-precision)
    atomic charges of test/test_coulomb/samples.py:60-110."""
    ref = energies["es3_atom_gfn1"]["MB16_43_01"]
    mol = O.make_mol(np.array(mols["MB16_43_01"]["numbers"]))
    q = np.array(ref["q_atom"])
    assert abs(float((O.gam3(mol) * q**3).sum() / 3.0) - ref["es3"]) < 1e-8


def test_fermi_filling_known_answer_5000K(energies):
    """Fermi smearing at 5000 K on the SiH4 orbital energies (wavefunction/filling.py:201-366) and the electronic free
    energy kT sum ln(f^f (1-f)^(1-f)) (scf/base.py:586-594) against test/test_wavefunction/test_filling.py:185-268
    (reference tolerance there: 1.5e-7)."""
    ref = energies["fermi_sih4_5000K"]
    par = O.params()
    kt = ref["kelvin"] * par.kelvin2au
    f = O.fermi_occupation(np.array([ref["nel"] / 2, ref["nel"] / 2]), np.array(ref["emo"]), kt)
    assert np.abs(f.sum(0) - np.array(ref["focc"])).max() < 1e-8
    o1, o2 = np.maximum(f, O.EPS), np.maximum(1.0 - f, O.EPS)
    g = kt * float(np.sum(np.log(o1**o1 * o2**o2)))
    assert abs(g - ref["fenergy"]) < 1e-8


def _wiberg(mol, P, S):
    ps = P @ S
    t = ps * ps.T
    w = np.zeros((mol.nat, mol.nat))
    np.add.at(w, (mol.ao_atom[:, None], mol.ao_atom[None, :]), t)
    np.fill_diagonal(w, 0.0)
    return w


@pytest.mark.parametrize("name", ["H2", "LiH", "SiH4"])
def test_converged_density_vs_reference_wiberg_and_mulliken(mols, energies, name):
    """Wiberg bond orders (wavefunction/wiberg.py:33-60) and Mulliken atomic charges of the converged SCF density against
    the literals of test/test_wavefunction/samples.py: pins P (and through it the whole SCF) at the 1e-8 level."""
    ref = energies["wiberg_gfn1"][name]
    m = mols[name]
    r = O.singlepoint(np.array(m["numbers"]), np.array(m["positions"]), 0.0,
                      opts={"exclude": ("disp",), "x_atol": 1e-10, "x_atol_max": 1e-10, "maxiter": 100})
    mol = O.make_mol(np.array(m["numbers"]))
    w = _wiberg(mol, r.P, r.S)
    assert np.abs(w - np.array(ref["wiberg"]).reshape(mol.nat, mol.nat)).max() < 1e-8
    assert np.abs(r.q_at - np.array(ref["mulliken_charges"])).max() < 6e-6  # the reference lists 5 digits


@pytest.mark.parametrize("name,tol", [("H2", 2e-7), ("LiH", 3e-7), ("H2O", 2e-7), ("CH4", 2e-7), ("SiH4", 2e-7),
                                      ("MB16_43_01", 2e-6), ("LYS_xao", 6e-6)])
def test_h0_gradient_part_vs_reference_goldens(mols, goldens, name, tol):
    """Overlap-derivative + scaling-function part of the electronic gradient (= `dedr` of GFN1Hamiltonian.get_gradient,
    xtb/gfn1.py:185-408) against test/test_hamiltonian/grad.npz (float32, generated at x_atol = 1e-6)."""
    m = mols[name]
    r = O.singlepoint(np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"]),
                      opts={"exclude": ("disp",), "x_atol": 1e-6, "x_atol_max": 1e-6}, grad=True)
    assert np.abs(r.gradient_parts["h0_dedr"] - goldens[f"h0_grad/{name}"]).max() < tol


@pytest.mark.parametrize("name", ["CH3Br_NH3", "CH3I_OCH2", "Br2_NH3", "CH2BrI_cluster"])
def test_halogen_gradient_is_derivative_of_energy(name):
    """Halogen-bond term (hal.py:209-364) on molecules where it is non-zero: analytic dE/dR == central finite
    difference of the oracle energy (the reference obtains this derivative by autograd, classicals/base.py:118-156)."""
    from halogen_mols import halogen_mol

    nums, pos = halogen_mol(name)
    m = O.make_mol(nums)
    e, g = O.halogen(m, pos, grad=True)
    assert abs(e.sum()) > 1e-4
    h = 1e-5
    fd = np.zeros_like(pos)
    for a in range(m.nat):
        for c in range(3):
            pp, pm = pos.copy(), pos.copy()
            pp[a, c] += h
            pm[a, c] -= h
            fd[a, c] = (O.halogen(m, pp).sum() - O.halogen(m, pm).sum()) / (2 * h)
    assert np.abs(fd - g).max() < 1e-9
    assert np.abs(g.sum(0)).max() < 1e-14  # translational invariance


def test_total_gradient_with_halogen_bond_vs_finite_difference():
    """Total oracle gradient (incl. the halogen term) == finite difference of the total energy at tight SCF."""
    from halogen_mols import halogen_mol

    nums, pos = halogen_mol("CH3Br_NH3")
    o = {"exclude": ("disp",), "x_atol": 1e-11, "x_atol_max": 1e-11}
    r = O.singlepoint(nums, pos, opts=o, grad=True)
    h = 1e-4
    for a, c in [(4, 2), (5, 0), (0, 2), (4, 1)]:
        pp, pm = pos.copy(), pos.copy()
        pp[a, c] += h
        pm[a, c] -= h
        fd = (O.singlepoint(nums, pp, opts=o).energy - O.singlepoint(nums, pm, opts=o).energy) / (2 * h)
        assert abs(fd - r.gradient[a, c]) < 2e-8
