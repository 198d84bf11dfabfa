"""File readers and command line: xyz / Turbomole coord / .CHRG / .UHF -> tensors -> single point."""
import json

import numpy as np
import pytest
import torch

from dxtb_b200 import io


def _write_xyz(path, numbers, pos_bohr, comment=""):
    lines = [str(len(numbers)), comment]
    for z, p in zip(numbers, pos_bohr):
        lines.append(f"{io.SYMBOLS[z]} {p[0] / io.AA2AU:.14f} {p[1] / io.AA2AU:.14f} {p[2] / io.AA2AU:.14f}")
    path.write_text("\n".join(lines) + "\n")


def _write_coord(path, numbers, pos_bohr, angs=False):
    s = 1.0 / io.AA2AU if angs else 1.0
    lines = ["$coord angs" if angs else "$coord"]
    for z, p in zip(numbers, pos_bohr):
        lines.append(f"  {p[0] * s:.14f} {p[1] * s:.14f} {p[2] * s:.14f} {io.SYMBOLS[z].lower()}")
    lines += ["$user-defined bonds", "$end"]
    path.write_text("\n".join(lines) + "\n")


def test_xyz_and_coord_round_trip(tmp_path, mols):
    m = mols["H2O"]
    ref = np.array(m["positions"])
    _write_xyz(tmp_path / "h2o.xyz", m["numbers"], ref)
    _write_coord(tmp_path / "coord", m["numbers"], ref)
    _write_coord(tmp_path / "h2o.coord", m["numbers"], ref, angs=True)
    for f in ("h2o.xyz", "coord", "h2o.coord"):
        numbers, pos = io.read_structure(tmp_path / f)
        assert numbers.dtype == torch.int64 and pos.dtype == torch.float64
        assert numbers.tolist() == m["numbers"]
        assert np.abs(pos.numpy() - ref).max() < 1e-11
    assert io.read_chrg(tmp_path / "coord") == 0 and io.read_spin(tmp_path / "coord") == 0
    (tmp_path / ".CHRG").write_text("1\n")
    (tmp_path / ".UHF").write_text("2\n")
    assert io.read_chrg(tmp_path / "coord") == 1 and io.read_spin(tmp_path / "h2o.xyz") == 2


def test_multi_frame_xyz_and_errors(tmp_path, mols):
    m = mols["LiH"]
    a = np.array(m["positions"])
    with open(tmp_path / "ens.xyz", "w") as f:
        for k in range(3):
            _write_xyz(tmp_path / "one.xyz", m["numbers"], a + 0.01 * k)
            f.write((tmp_path / "one.xyz").read_text())
    numbers, pos = io.read_xyz(tmp_path / "ens.xyz", frame=None)
    assert numbers.shape == (3, 2) and pos.shape == (3, 2, 3)
    assert np.abs(pos[2].numpy() - (a + 0.02)).max() < 1e-11
    n1, p1 = io.read_xyz(tmp_path / "ens.xyz", frame=1)
    assert torch.equal(p1, pos[1])
    (tmp_path / "bad.xyz").write_text("2\n\nXx 0 0 0\nH 0 0 1\n")
    with pytest.raises(ValueError):
        io.read_xyz(tmp_path / "bad.xyz")
    (tmp_path / "trunc.xyz").write_text("3\n\nH 0 0 0\n")
    with pytest.raises(ValueError):
        io.read_xyz(tmp_path / "trunc.xyz")
    with pytest.raises(ValueError):
        io.read_structure(tmp_path / "mol.pdb")
    (tmp_path / "p.coord").write_text("$coord frac\n 0 0 0 h\n$end\n")
    with pytest.raises(NotImplementedError):
        io.read_coord(tmp_path / "p.coord")


def test_pack_zero_pads(mols):
    s = [(torch.tensor(mols[n]["numbers"]), torch.tensor(mols[n]["positions"], dtype=torch.float64)) for n in ("H2O", "LiH")]
    numbers, pos = io.pack(s)
    assert numbers.shape == (2, 3) and pos.shape == (2, 3, 3)
    assert numbers[1].tolist() == [3, 1, 0] and float(pos[1, 2].abs().sum()) == 0.0


def test_cli_options_map_to_calculator_opts():
    from dxtb_b200 import cli

    a = cli.parser().parse_args(["x.xyz", "--etemp", "0", "--maxiter", "50", "--xtol", "1e-6", "--exclude", "disp", "hal", "--grad"])
    o = cli.options(a)
    assert o == {"fermi_etemp": 0.0, "maxiter": 50, "x_atol": 1e-6, "x_atol_max": 1e-7, "exclude": ["disp", "hal"]}
    assert a.grad is True and a.chrg is None


@pytest.mark.gpu
def test_cli_single_point_batch(tmp_path, mols):
    """Two files (one with a .CHRG next to it) as one batch; energies and gradients against the oracle."""
    from dxtb_b200 import cli
    from oracle import gfn1_oracle as O

    d1, d2 = tmp_path / "a", tmp_path / "b"
    d1.mkdir(), d2.mkdir()
    _write_xyz(d1 / "h2o.xyz", mols["H2O"]["numbers"], np.array(mols["H2O"]["positions"]))
    _write_coord(d2 / "coord", mols["AD7en+"]["numbers"], np.array(mols["AD7en+"]["positions"]))
    (d2 / ".CHRG").write_text("1\n")
    out = tmp_path / "res.json"
    res = cli.run([str(d1 / "h2o.xyz"), str(d2 / "coord"), "--exclude", "disp", "--grad", "--json", str(out)])
    saved = json.loads(out.read_text())
    assert [s["charge"] for s in saved["systems"]] == [0, 1]
    for s, name in zip(res["systems"], ("H2O", "AD7en+")):
        m = mols[name]
        r = O.singlepoint(np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"]), opts={"exclude": ["disp"]}, grad=True)
        assert abs(s["energy"] - r.energy) < 1e-9
        assert s["scf_iterations"] == r.iterations
        assert np.abs(np.array(s["gradient"]) - r.gradient).max() < 1e-7
        assert np.abs(np.array(s["charges"]) - r.q_at).max() < 1e-7
