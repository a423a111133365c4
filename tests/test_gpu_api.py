"""The reference's own API-level tests, re-stated against pantea_b200's mirror of that API (GPU).

Mirrors /root/reference/tests/test_acsf.py, test_nnp.py, test_md.py (API behaviour) and the notebook cells whose
printed outputs serve as golden vectors; the numbers come from tests/golden/reference_vectors.json."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vec(golden_dir):
    return json.loads((golden_dir / "reference_vectors.json").read_text())


@pytest.fixture(scope="module")
def h2o_structure(golden_dir):
    from pantea_b200.datasets import Dataset
    return Dataset.from_runner(golden_dir / "h2o.data")[0]


def _np(t):
    return t.detach().cpu().numpy()


def test_acsf_without_pbc(vec):
    """reference tests/test_acsf.py:123-151"""
    from pantea_b200.atoms import Structure
    from pantea_b200.descriptors.acsf import ACSF, G2, CutoffFunction, NeighborElements
    v = vec["ne2_g2"]
    cfn = CutoffFunction.from_type("tanhu", r_cutoff=3.0)
    acsf = ACSF("Ne", tuple((G2(cfn, eta=1.0, r_shift=rs), NeighborElements("Ne")) for rs in v["r_shifts"]), ())
    s = Structure.from_dict({"positions": v["positions"], "elements": v["elements"]})
    out = acsf(s)
    assert tuple(out.shape) == (2, 5)
    np.testing.assert_allclose(_np(out), np.tile(v["expected_row"], (2, 1)), rtol=1e-8)


def test_acsf_with_pbc_and_index_validation(vec, h2o_structure):
    """reference tests/test_acsf.py:153-173 and acsf.py:66-80, 103-113"""
    from pantea_b200.descriptors.acsf import ACSF, G2, G3, CutoffFunction, NeighborElements
    v = vec["h2o_pbc_g2_g3"]
    cfn = CutoffFunction.from_type("tanhu", r_cutoff=v["cutoff"][1])
    acsf = ACSF("O", ((G2(cfn, r_shift=0.0, eta=0.001), NeighborElements("H")),),
                ((G3(cfn, eta=0.07, zeta=1.0, lambda0=1.0, r_shift=0.0), NeighborElements("H", "H")),))
    assert tuple(acsf(h2o_structure).shape) == tuple(v["shape"])
    np.testing.assert_allclose(_np(acsf(h2o_structure, atom_index=0)), [v["expected_atom0"]], rtol=0, atol=6e-11)
    with pytest.raises(ValueError):
        acsf(h2o_structure, atom_index=1)             # atom 1 is H, the descriptor is O-centred
    with pytest.raises(ValueError):
        acsf.grad(h2o_structure, atom_index=12)       # out of range
    assert tuple(acsf.grad(h2o_structure).shape) == (12, 2, 3)          # all atoms when atom_index is None
    assert tuple(acsf.grad(h2o_structure, atom_index=[0, 3]).shape) == (2, 2, 3)


def test_notebook_descriptor_values_gradient_distances_neighbors(vec, h2o_structure):
    """examples/getting_started.ipynb cells 17, 20, 25-29"""
    from pantea_b200.atoms import Neighbor, calculate_distances
    from pantea_b200.descriptors.acsf import ACSF, G2, G3, CutoffFunction, NeighborElements
    cfn = CutoffFunction.from_type("tanhu", r_cutoff=12.0)
    acsf = ACSF("O", ((G2(cfn, 0.0, 0.001), NeighborElements("H")), (G2(cfn, 0.0, 0.01), NeighborElements("H"))),
                ((G3(cfn, 0.2, 1.0, 1.0, 0.0), NeighborElements("H", "H")), (G3(cfn, 0.2, 1.0, 1.0, 0.0), NeighborElements("H", "O"))))
    v = vec["notebook_acsf"]
    np.testing.assert_allclose(_np(acsf(h2o_structure)), v["expected_values"], rtol=2e-8)
    np.testing.assert_allclose(_np(acsf.grad(h2o_structure)[:1])[0], v["expected_grad_atom0"], atol=6e-9)
    d = calculate_distances(h2o_structure)
    np.testing.assert_allclose(_np(d[0, :5]), vec["notebook_distances"]["expected"], atol=5e-9)
    d2, dx = calculate_distances(h2o_structure, atom_index=[0, 1], neighbor_atom_index=[2, 3, 4], with_aux=True)
    assert tuple(d2.shape) == (2, 3) and tuple(dx.shape) == (2, 3, 3)
    np.testing.assert_allclose(_np(d2), _np(d[:2, 2:5]), rtol=0, atol=0)
    nb = Neighbor.from_structure(h2o_structure, r_cutoff=vec["notebook_neighbors"]["r_cutoff"])
    assert int(nb.masks[0].sum()) == vec["notebook_neighbors"]["expected_count_atom0"]
    assert nb.masks.shape == (12, 12) and not bool(nb.masks.diagonal().any())
    assert repr(nb) == "Neighbor(r_cutoff=10.0)"


@pytest.mark.parametrize("dtype,e_atol,f_rtol", [(torch.float64, 3e-7, 1e-5), (torch.float32, 1e-6, 3e-5)])
def test_nnp_outputs(vec, golden_dir, dtype, e_atol, f_rtol):
    """reference tests/test_nnp.py:44-81 (its FLOATX is float32; both modes are checked here)"""
    from pantea_b200.datasets import Dataset
    from pantea_b200.potentials import NeuralNetworkPotential
    from pantea_b200.types import default_dtype
    old = default_dtype.FLOATX
    default_dtype.FLOATX = dtype
    try:
        nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
        assert nnp.num_elements == 2 and nnp.elements == ("H", "O")
        structure = Dataset.from_runner(golden_dir / "h2o.data")[0]
        with pytest.raises(ValueError):
            nnp(structure)                                   # scaler parameters not loaded yet
        nnp.load_scaler()
        nnp.load_model()
        v = vec["nnp_fp32"]
        energy, forces = nnp(structure), nnp.compute_forces(structure)
        assert energy.ndim == 0 and tuple(forces.shape) == (12, 3) and forces.dtype == dtype
        np.testing.assert_allclose(float(energy), v["energy"], rtol=0, atol=e_atol)
        np.testing.assert_allclose(_np(forces).astype(np.float64), v["forces"], rtol=f_rtol, atol=2e-7)
        e2, f2 = nnp.compute_energy_and_forces(structure)
        assert float(e2) == float(energy) and torch.equal(f2, forces)
    finally:
        default_dtype.FLOATX = old


def test_structure_missing_element_raises(golden_dir):
    from pantea_b200.atoms import Structure
    from pantea_b200.potentials import NeuralNetworkPotential
    nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
    nnp.load()
    only_o = Structure.from_dict({"positions": [[0.0, 0, 0], [3.0, 0, 0]], "elements": ["O", "O"],
                                  "lattice": np.diag([30.0, 30.0, 30.0])})
    with pytest.raises(KeyError):
        nnp(only_o)                                           # reference: positions['H'] KeyError (energy.py:54-60)


def _water_system(n_atoms=192, temperature=300.0):
    from pantea_b200.atoms import Structure
    from pantea_b200.potentials import NeuralNetworkPotential
    from pantea_b200.simulation import System
    from pantea_b200.utils.synthetic import water_box
    from tests.conftest import GOLDEN
    nnp = NeuralNetworkPotential.from_runner(GOLDEN / "h2o.json")
    nnp.load()
    p, t, box = water_box(n_atoms)
    s = Structure.from_dict({"positions": p, "elements": ["H" if x == 1 else "O" for x in t], "lattice": np.diag(box)})
    return System.from_structure(s, nnp, temperature=temperature, seed=2024), nnp


def test_md_simulator_api_and_device_loop_agree():
    """reference tests/test_md.py:99-126 (step counter, elapsed time, COM velocity) + one-step vs device-loop equality"""
    from pantea_b200.simulation import MDSimulator, simulate
    from pantea_b200.units import units
    sys_a, _ = _water_system()
    sys_b, _ = _water_system()
    # rescaled to 300 K, then the COM velocity is removed (system.py:94-95): slightly below the target
    assert 295.0 < float(sys_a.get_temperature()) <= 300.0 + 1e-9
    np.testing.assert_allclose(_np(sys_a.get_center_of_mass_velocity()), 0.0, atol=1e-12)
    md_a, md_b = MDSimulator(time_step=0.25), MDSimulator(time_step=0.25)
    for _ in range(3):
        md_a.simulate_one_step(sys_a)
    md_b.simulate_steps(sys_b, 3)
    assert md_a.step == md_b.step == 3 and md_a.elapsed_time == pytest.approx(0.75)
    assert torch.equal(sys_a.positions, sys_b.positions) and torch.equal(sys_a.velocities, sys_b.velocities)
    assert torch.equal(sys_a.forces, sys_b.forces)
    line = md_a.repr_physical_params(sys_a)
    assert line.startswith("3 ") and "Temp[K]:" in line and "Etot[Ha]:" in line and "Pres[kb]:" in line
    assert float(sys_a.get_total_energy()) == pytest.approx(float(sys_a.get_potential_energy()) + float(sys_a.get_kinetic_energy()))
    simulate(sys_b, md_b, num_steps=4, output_freq=2)
    assert md_b.step == 7
    assert units.TO_PICO_SECOND * md_b.elapsed_time == pytest.approx(units.TO_PICO_SECOND * 1.75)


def test_thermostat_pulls_temperature_towards_target():
    from pantea_b200.simulation import BrendsenThermostat, MDSimulator
    system, _ = _water_system(temperature=600.0)
    md = MDSimulator(time_step=0.25, thermostat=BrendsenThermostat(target_temperature=300.0, time_constant=2.5))
    t0 = float(system.get_temperature())
    v_before = system.velocities.clone()
    scaled = md.thermostat.get_rescaled_velocities(md, system)
    factor = 1.0 / np.sqrt(1.0 + (0.25 / 2.5) * (t0 / 300.0 - 1.0))           # thermostat.py:16-21
    np.testing.assert_allclose(_np(scaled), _np(v_before) * factor, rtol=1e-13)
    sys_loop, _ = _water_system(temperature=600.0)
    md_loop = MDSimulator(time_step=0.25, thermostat=BrendsenThermostat(300.0, 2.5))
    md.simulate_one_step(system)
    md_loop.simulate_steps(sys_loop, 1)
    np.testing.assert_allclose(_np(system.velocities), _np(sys_loop.velocities), rtol=1e-12, atol=1e-15)


def test_mc_simulator_follows_numpy_stream():
    """reference monte_carlo.py:65-91: displacements, indices, then the acceptance draw from numpy's global stream"""
    from pantea_b200.simulation import MCSimulator
    system, nnp = _water_system(n_atoms=24)
    e0 = float(system.structure.total_energy)
    pos0 = system.positions.clone()
    mc = MCSimulator(translate_step=0.05, target_temperature=300.0, movements_per_step=3, seed=12345)
    rng = np.random.RandomState(12345)
    disp = rng.uniform(-0.05, 0.05, size=(3, 3))
    idx = rng.randint(0, system.natoms, size=(3,))
    mc.simulate_one_step(system)
    assert mc.step == 1
    trial = pos0.clone()
    trial.index_add_(0, torch.as_tensor(idx, device=trial.device), torch.as_tensor(disp, device=trial.device))
    e_trial = float(nnp(system.structure.replace(positions=trial)))
    if e_trial <= e0:
        accepted = True
    else:
        accepted = np.exp(-(e_trial - e0) / (3.166811563e-6 * 300.0)) >= rng.uniform(0.0, 1.0)
    expected = system.structure.replace(positions=trial).positions if accepted else pos0
    assert torch.equal(system.positions, expected)
    assert float(system.structure.total_energy) == pytest.approx(e_trial if accepted else e0)
    assert mc.repr_physical_params(system).startswith("1 ")


# ------------------------------------------------------------------------------------------ Lennard-Jones drivers
def _helium(vec):
    from pantea_b200.atoms import Structure
    from pantea_b200.simulation import LJPotential
    from pantea_b200.units import units
    v = vec["lj_helium"]
    d = v["d_angstrom"]
    pos = [[d / 2 + d * i, d / 2 + d * j, d / 2 + d * k] for i in range(2) for j in range(2) for k in range(2)]
    s = Structure.from_dict({"positions": np.asarray(pos) * units.FROM_ANGSTROM, "elements": ["He"] * 8,
                             "lattice": np.diag([2 * d] * 3) * units.FROM_ANGSTROM})
    lj = LJPotential(sigma=v["sigma_angstrom"] * units.FROM_ANGSTROM, epsilon=v["epsilon_ev"] * units.FROM_ELECTRON_VOLT,
                     r_cutoff=v["r_cutoff_angstrom"] * units.FROM_ANGSTROM)
    return s, lj


def test_lj_energy_and_gradient_match_dense_restatement(vec):
    from oracle import dense_oracle as D
    from pantea_b200.atoms import Structure
    from pantea_b200.simulation import LJPotential
    s, lj = _helium(vec)
    # printed (not asserted) by the reference's test with 7 digits: -4.575687e-06
    np.testing.assert_allclose(float(lj(s)), vec["lj_helium"]["initial_energy"], rtol=1e-6)
    rng = np.random.default_rng(4)
    pos = rng.uniform(0, 30.0, size=(200, 3))
    gas = Structure.from_dict({"positions": pos, "elements": ["He"] * 200, "lattice": np.diag([30.0, 28.0, 33.0])})
    pot = LJPotential(sigma=4.77, epsilon=1.73e-5, r_cutoff=11.9)
    e_o, g_o = D.lj_energy_and_gradient(torch.as_tensor(_np(gas.positions)), torch.tensor([30.0, 28.0, 33.0], dtype=torch.float64),
                                        4.77, 1.73e-5, 11.9)
    assert abs(float(pot(gas)) - float(e_o)) < 1e-12 * abs(float(e_o))
    f = _np(pot.compute_forces(gas))
    assert np.abs(f - g_o.numpy()).max() < 1e-11 * np.abs(g_o.numpy()).max()
    with pytest.raises(ValueError):
        LJPotential(1.0, 1.0, 1.0, gradient_method="nope")


def test_reference_mc_golden_energy(vec):
    """reference tests/test_mc.py:66-89: potential energy after one MC step (seed 12345, 10 moves of <= 0.3 A)"""
    from pantea_b200.simulation import MCSimulator, System
    from pantea_b200.units import units
    v = vec["lj_helium"]["mc"]
    s, lj = _helium(vec)
    mc = MCSimulator(translate_step=v["translate_step_angstrom"] * units.FROM_ANGSTROM,
                     target_temperature=v["target_temperature"], movements_per_step=v["movements_per_step"])
    system = System.from_structure(structure=s, potential=lj, temperature=300.0)
    assert mc.step == 0
    mc.simulate_one_step(system)
    assert mc.step == 1
    np.testing.assert_allclose(float(system.get_potential_energy()), v["energy_after_one_step"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(float(system.get_potential_energy()), v["energy_after_one_step"], rtol=2e-6)


def test_reference_md_lj_attributes(vec):
    """reference tests/test_md.py:55-126: step counter, elapsed time, COM velocity and position around one MD step"""
    from pantea_b200.atoms import ElementMap
    from pantea_b200.simulation import BrendsenThermostat, MDSimulator, System
    from pantea_b200.units import units
    v = vec["lj_helium"]["md"]
    s, lj = _helium(vec)
    dt = v["time_step_fs"] * units.FROM_FEMTO_SECOND
    md = MDSimulator(time_step=dt, thermostat=BrendsenThermostat(target_temperature=300.0, time_constant=100 * dt))
    system = System.from_structure(structure=s, potential=lj, temperature=300.0, seed=2023)
    assert md.step == 0 and md.elapsed_time == 0.0 and md.time_step == pytest.approx(dt)
    np.testing.assert_allclose(_np(system.get_center_of_mass_velocity()), 0.0, atol=1e-12)
    np.testing.assert_allclose(_np(system.get_center_of_mass_position()), v["com_position"], rtol=1e-8)
    np.testing.assert_allclose(_np(system.positions), _np(s.positions))
    np.testing.assert_allclose(_np(system.masses).ravel(), _np(ElementMap.get_masses_from_structure(s)))
    md.simulate_one_step(system)
    assert md.step == 1 and md.elapsed_time == pytest.approx(dt)
    np.testing.assert_allclose(_np(system.get_center_of_mass_velocity()), 0.0, atol=1e-10)
    np.testing.assert_allclose(_np(system.get_center_of_mass_position()), v["com_position"], rtol=1e-6)


# ------------------------------------------------------------------------------------------ scaler fitting (8(f)-2)
@pytest.mark.parametrize("dtype,n_rows,n_cols", [(torch.float64, 1, 3), (torch.float64, 777, 27), (torch.float64, 20011, 70),
                                                 (torch.float32, 5000, 33)])
def test_scaler_statistics_kernel_matches_definitions(dtype, n_rows, n_cols):
    """pantea_scaler_stats (two-pass mean / population sigma / min / max per feature) against numpy float64, through
    DescriptorScaler.fit on a device-resident batch; also a strided (non-contiguous rows) view."""
    from pantea_b200.descriptors import DescriptorScaler
    rng = np.random.default_rng(9)
    host = (rng.normal(1.5, 2.0, size=(n_rows, n_cols + 5)) * rng.uniform(1e-3, 1e3, size=n_cols + 5)).astype(
        np.float64 if dtype == torch.float64 else np.float32)
    dev = torch.as_tensor(host, device="cuda")
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    for view, ref in ((dev[:, :n_cols], host[:, :n_cols].astype(np.float64)), (dev.contiguous(), host.astype(np.float64))):
        p = DescriptorScaler.fit(view)
        assert int(p.nsamples) == n_rows and int(p.dimension) == view.shape[1] and p.mean.dtype == dtype
        scale = np.abs(ref).max(0)  # per feature: the columns span six orders of magnitude
        assert (np.abs(_np(p.mean) - ref.mean(0)) <= tol * scale).all()
        assert (np.abs(_np(p.sigma) - ref.std(0)) <= max(tol, 1e-11) * scale).all()
        np.testing.assert_array_equal(_np(p.minval).astype(np.float64), ref.min(0))
        np.testing.assert_array_equal(_np(p.maxval).astype(np.float64), ref.max(0))
    again = DescriptorScaler.fit(dev[:, :n_cols])
    assert torch.equal(again.sigma, DescriptorScaler.fit(dev[:, :n_cols]).sigma)  # fixed-order reduction


def test_fit_scaler_over_a_dataset_matches_oracle_descriptors(golden_dir):
    """trainer.fit_scaler (reference trainer.py:68-88): statistics of the ACSF descriptors of every element over a
    dataset of water boxes, against numpy statistics of the oracle's descriptors of the same structures; a 2-way split
    of the dataset merged with the reference's partial_fit rule gives the same numbers."""
    from oracle import c_oracle
    from oracle.spec import load_potential, water_box
    from pantea_b200.atoms import Structure
    from pantea_b200.descriptors import DescriptorScaler
    from pantea_b200.potentials import NeuralNetworkPotential
    from pantea_b200.potentials.nnp import NeuralNetworkPotentialTrainer
    specs = {s.atom_type: s for s in load_potential(golden_dir / "h2o.json")}
    names = {1: "H", 2: "O"}
    structures, expected = [], {"H": [], "O": []}
    for seed, n_atoms in ((1, 192), (2, 192), (3, 81), (4, 648)):
        pos, types, box = water_box(n_atoms, seed=seed)
        structures.append(Structure.from_dict({"elements": [names[int(t)] for t in types], "positions": pos,
                                               "lattice": np.diag(box)}, dtype=torch.float64))
        for t, name in names.items():
            G, _ = c_oracle.acsf(specs[t], pos, types, box, centres=np.nonzero(types == t)[0], grad=False)
            expected[name].append(G)
    nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
    params = NeuralNetworkPotentialTrainer(nnp).fit_scaler(structures)
    for name in ("H", "O"):
        ref = np.concatenate(expected[name])
        p = params[name]
        assert int(p.nsamples) == len(ref) and int(p.dimension) == ref.shape[1]
        np.testing.assert_allclose(_np(p.mean), ref.mean(0), rtol=1e-10)
        np.testing.assert_allclose(_np(p.sigma), ref.std(0), rtol=1e-9)
        np.testing.assert_allclose(_np(p.minval), ref.min(0), rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(_np(p.maxval), ref.max(0), rtol=1e-10)
    # the same dataset in two shards (what two ranks would hold), merged
    shards = []
    for r in range(2):
        part = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
        shards.append(NeuralNetworkPotentialTrainer(part).fit_scaler(structures[r::2]))
    for name in ("H", "O"):
        merged = DescriptorScaler.merge(shards[0][name], shards[1][name])
        np.testing.assert_allclose(_np(merged.mean), _np(params[name].mean), rtol=1e-12)
        np.testing.assert_allclose(_np(merged.sigma), _np(params[name].sigma), rtol=1e-10)
        assert int(merged.nsamples) == int(params[name].nsamples)
    # fitted parameters are usable: the potential evaluates with them once model weights are loaded
    nnp.load_model()
    assert torch.isfinite(nnp(structures[0]))
