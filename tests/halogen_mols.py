"""In-tree test molecules with a non-zero halogen-bond energy (Br / I next to N, O, P, S).

The reference's halogen goldens (test/test_classical/test_halogen/samples.py) carry no geometry that lives in the tree
together with a Br/I-containing literal usable offline, and every golden molecule extracted so far has E_xb = 0 (only
Br, I, At have a non-zero ``xbond`` in gfn1-xtb.toml).  These hand-built complexes close that hole; coordinates in Angstrom.
"""
import numpy as np

AA2AU = 1.0 / 0.529177210903

_CH3 = [[0.0, 0.0, 0.0], [1.03, 0.0, -0.36], [-0.515, 0.892, -0.36], [-0.515, -0.892, -0.36]]

HALOGEN_MOLS = {
    # methyl bromide ... ammonia (sigma-hole contact, slightly bent)
    "CH3Br_NH3": ([6, 1, 1, 1, 35, 7, 1, 1, 1],
                  _CH3 + [[0.0, 0.0, 1.94], [0.30, 0.20, 4.90], [1.20, 0.30, 5.30], [-0.20, 1.00, 5.30], [-0.10, -0.70, 5.30]]),
    # methyl iodide ... formaldehyde
    "CH3I_OCH2": ([6, 1, 1, 1, 53, 8, 6, 1, 1],
                  _CH3 + [[0.0, 0.0, 2.14], [0.20, 0.10, 5.00], [0.30, 0.20, 6.21], [1.25, 0.30, 6.78], [-0.58, 0.25, 6.82]]),
    # dibromine ... ammonia: two halogens, the far one sees the base past its neighbour
    "Br2_NH3": ([35, 35, 7, 1, 1, 1],
                [[0.0, 0.0, 0.0], [0.0, 0.0, 2.28], [0.10, -0.20, 4.95], [1.02, -0.30, 5.35], [-0.40, 0.62, 5.38], [-0.35, -1.02, 5.30]]),
    # CH2BrI with H2S, PH3 and H2O around it: several bases per halogen, S and P among them
    "CH2BrI_cluster": ([6, 1, 1, 35, 53, 16, 1, 1, 15, 1, 1, 1, 8, 1, 1],
                       [[0.0, 0.0, 0.0], [0.62, 0.89, -0.12], [0.62, -0.89, -0.12], [-1.12, 0.0, -1.58], [-1.22, 0.0, 1.76],
                        [-3.10, 0.45, -3.90], [-3.95, 1.35, -4.35], [-3.85, -0.62, -4.25],
                        [-2.95, -0.35, 4.95], [-2.35, -1.55, 5.35], [-4.35, -0.55, 5.05], [-2.65, 0.55, 6.05],
                        [3.35, 0.15, 0.35], [3.85, 0.75, -0.22], [3.80, -0.70, 0.30]]),
}


def halogen_mol(name):
    z, xyz = HALOGEN_MOLS[name]
    return np.array(z, dtype=np.int64), np.array(xyz, dtype=np.float64) * AA2AU
