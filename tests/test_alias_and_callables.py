"""Drop-in import surface (`pantea.*` alias of `pantea_b200.*`) and the public symmetry-function callables
(reference `descriptors/acsf/cutoff.py:64-110`, `radial.py:39-61`, `angular.py:51-107`), checked on the CPU against the
dense oracle's own formulas; plus the extended-xyz trajectory writer."""
import math

import numpy as np
import torch


def test_pantea_alias_resolves_to_the_same_modules():
    import pantea
    import pantea.atoms
    import pantea.potentials.nnp.potential as p1
    import pantea_b200.atoms
    import pantea_b200.potentials.nnp.potential as p2
    from pantea.atoms import Box, ElementMap, Neighbor, Structure  # noqa: F401
    from pantea.descriptors import ACSF, DescriptorScaler, ScalerParams  # noqa: F401
    from pantea.descriptors.acsf import G1, G2, G3, G9, CutoffFunction, NeighborElements  # noqa: F401
    from pantea.potentials import NNP, NeuralNetworkPotential
    from pantea.simulation import (BrendsenThermostat, LJPotential, MCSimulator, MDSimulator, System,  # noqa: F401
                                   simulate)
    from pantea.types import default_dtype  # noqa: F401
    from pantea.units import units  # noqa: F401

    assert pantea.atoms is pantea_b200.atoms and p1 is p2
    assert NNP is NeuralNetworkPotential is p2.NeuralNetworkPotential


def test_cutoff_function_call_matches_the_formulas():
    from pantea.descriptors.acsf import CutoffFunction

    r = torch.tensor([0.3, 0.9, 5.0, 11.999, 12.0, 12.5], dtype=torch.float64)
    rc = 12.0
    pre = ((math.e + 1 / math.e) / (math.e - 1 / math.e)) ** 3
    expect = {
        "hard": torch.ones_like(r),
        "tanhu": torch.tanh(1 - r / rc) ** 3,
        "tanh": pre * torch.tanh(1 - r / rc) ** 3,
        "cos": 0.5 * (torch.cos(math.pi * r / rc) + 1),
        "exp": torch.exp(1 - 1 / (1 - (r / rc) ** 2)),
        "poly1": (2 * r - 3) * r**2 + 1,            # raw r, not r / rc (cutoff.py:105-110)
        "poly2": ((15 - 6 * r) * r - 10) * r**3 + 1,
    }
    for kind, val in expect.items():
        got = CutoffFunction.from_type(kind, rc)(r)
        want = torch.where(r < rc, val, torch.zeros_like(r))
        assert torch.allclose(got, want, rtol=1e-14, atol=0), kind
        assert float(got[-1]) == 0.0 and float(got[-2]) == 0.0  # strict r < rc


def test_symmetry_function_calls_match_the_dense_oracle_terms():
    from pantea.descriptors.acsf import G1, G2, G3, G9, CutoffFunction

    cfn = CutoffFunction.from_type("tanhu", 12.0)
    rij, rik, rjk = (torch.tensor(v, dtype=torch.float64) for v in ([1.8, 4.0], [2.9, 6.5], [1.8, 7.2]))
    cost = torch.tensor([0.3, -0.7], dtype=torch.float64)
    assert torch.equal(G1(cfn)(rij), cfn(rij))
    assert torch.allclose(G2(cfn, 0.5, 0.01)(rij), torch.exp(-0.01 * (rij - 0.5) ** 2) * cfn(rij), rtol=1e-15)
    g3 = G3(cfn, 0.07, 4.0, -1.0, 12.0)(rij, rik, rjk, cost)
    want = 2.0 ** (1 - 4.0) * (1 - cost) ** 4 * torch.exp(-0.07 * (rij**2 + rik**2 + rjk**2)) * cfn(rij) * cfn(rik) * cfn(rjk)
    assert torch.allclose(g3, want, rtol=1e-14)
    g9 = G9(cfn, 0.07, 2.0, 1.0, 12.0)(rij, rik, rjk, cost)   # r_shift stored, ignored; no r_jk terms
    want = 2.0 ** (1 - 2.0) * (1 + cost) ** 2 * torch.exp(-0.07 * (rij**2 + rik**2)) * cfn(rij) * cfn(rik)
    assert torch.allclose(g9, want, rtol=1e-14)


def test_xyz_trajectory_writer_round_trip(tmp_path):
    """`simulate(..., filename=...)` appends extended-xyz frames without ase (the reference goes through ase.io,
    simulate.py:80-82): atom count, Lattice header in Angstrom, one `El x y z` line per atom, frames appended."""
    from types import SimpleNamespace

    from pantea_b200.simulation.simulate import _append_xyz
    from pantea_b200.units import units

    pos = np.array([[0.0, 0.1, 0.2], [1.5, 1.6, 1.7], [3.0, 3.1, 3.2]])
    lattice = np.diag([10.0, 11.0, 12.0])
    struct = SimpleNamespace(positions=torch.tensor(pos), get_elements=lambda: ["O", "H", "H"],
                             box=SimpleNamespace(lattice=torch.tensor(lattice)))
    path = tmp_path / "traj.xyz"
    _append_xyz(path, struct)
    _append_xyz(path, struct)
    lines = path.read_text().splitlines()
    assert len(lines) == 10 and lines[0].strip() == "3" and lines[5].strip() == "3"
    assert lines[1].startswith('Lattice="') and 'pbc="T T T"' in lines[1]
    lat = [float(v) for v in lines[1].split('"')[1].split()]
    assert np.allclose(np.array(lat).reshape(3, 3), lattice * units.TO_ANGSTROM, atol=1e-7)
    parsed = np.array([[float(v) for v in ln.split()[1:]] for ln in lines[2:5]])
    assert [ln.split()[0] for ln in lines[2:5]] == ["O", "H", "H"]
    assert np.allclose(parsed, pos * units.TO_ANGSTROM, atol=1e-7)
    struct.box = None
    _append_xyz(path, struct)
    assert path.read_text().splitlines()[11] == ""
