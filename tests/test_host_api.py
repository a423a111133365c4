"""Host-side logic mirrored from the reference's own tests (tests/test_structure.py, test_runner.py, test_nn.py,
test_nnp.py::test_settings, test_acsf.py::test_acsf_attributes).  CPU only: no compute kernels are called."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from pantea_b200.atoms import Box, ElementMap, Structure
from pantea_b200.datasets import Dataset, RunnerDataSource
from pantea_b200.descriptors import ACSF, DescriptorScaler
from pantea_b200.descriptors.acsf import G2, G3, G9, CutoffFunction, NeighborElements
from pantea_b200.models import NeuralNetworkModel
from pantea_b200.potentials import NeuralNetworkPotential
from pantea_b200.potentials.nnp.settings import NeuralNetworkPotentialSettings
from pantea_b200.types import default_dtype
from pantea_b200.units import units
from pantea_b200.utils.tokenize import tokenize

H2O = {
    "lattice": [[11.8086403654, 0.0, 0.0], [0.0, 11.8086403654, 0.0], [0.0, 0.0, 11.8086403654]],
    "positions": [[0.2958498542, -0.8444146738, 1.9618569793], [1.7226932399, -1.8170359274, 0.5237867306],
                  [1.2660050151, -4.3958431356, 0.822408813]],
    "elements": ["O", "H", "H"],
    "charges": [0.1, 0.2, 0.3], "energies": [0.0, 0.0, 0.0], "forces": [[0, 0, 0]] * 3,
    "total_energy": [-32.0], "total_charge": [0.6],
}


def test_structure_from_dict_wraps_and_types():
    s = Structure.from_dict(H2O)
    assert s.natoms == 3 and s.dtype == torch.float64
    assert s.atom_types.tolist() == [2, 1, 1]                       # types by ascending Z (element.py:99-108)
    assert s.get_unique_elements() == ("H", "O") and s.get_elements() == ("O", "H", "H")
    L = 11.8086403654
    np.testing.assert_allclose(s.positions[0].cpu().numpy(), [0.2958498542, L - 0.8444146738, 1.9618569793], rtol=1e-14)
    assert float(s.positions.min()) >= 0.0 and float(s.positions.max()) < L   # wrapped (structure.py:80-81)
    assert s.select("H").tolist() == [1, 2] and s.select("O").tolist() == [0]
    assert s.total_energy.ndim == 0 and float(s.total_energy) == -32.0
    np.testing.assert_allclose(s.lattice.cpu().numpy(), H2O["lattice"])
    d = s.to_dict()
    assert d["elements"] == ["O", "H", "H"]
    s2 = Structure.from_dict(d)
    assert torch.equal(s2.positions, s.positions) and torch.equal(s2.atom_types, s.atom_types)
    assert repr(s) == "Structure(natoms=3, elements=('H', 'O'), dtype=torch.float64)"


def test_structure_without_box_and_float32():
    s = Structure.from_dict({"positions": [[0.0, 0.0, 0.0], [0.588897275] * 3], "elements": ["Ne", "Ne"]},
                            dtype=torch.float32)
    assert s.box is None and s.lattice is None and s.dtype == torch.float32
    assert s.forces.numel() == 0                                       # missing keys become empty arrays
    assert s.atom_types.tolist() == [1, 1]


def test_energy_offsets():
    s = Structure.from_dict(H2O)
    s.add_energy_offset({"O": 2.4, "H": 1.2})
    assert float(s.total_energy) == pytest.approx(-32.0 + 4.8)
    s.remove_energy_offset({"O": 2.4, "H": 1.2})
    assert float(s.total_energy) == pytest.approx(-32.0)


def test_element_map_and_masses(golden_dir):
    vec = json.loads((golden_dir / "reference_vectors.json").read_text())["masses"]
    for el in ("H", "O", "Ne"):
        assert ElementMap.get_atomic_mass_from_element(el) == pytest.approx(vec[el], rel=1e-14)
    em = ElementMap.from_list(["O", "H", "Ne", "H"])
    assert em.element_to_atom_type == {"H": 1, "O": 2, "Ne": 3}
    assert em.get_element_from_atom_type(3) == "Ne" and ElementMap.get_atomic_number_from_element("O") == 8
    assert ElementMap.get_element_from_atomic_number(10) == "Ne"
    s = Structure.from_dict(H2O)
    m = ElementMap.get_masses_from_structure(s)
    np.testing.assert_allclose(m.cpu().numpy(), [vec["O"], vec["H"], vec["H"]], rtol=1e-14)


def test_box():
    box = Box.from_list([10.0, 0, 0, 0, 8.0, 0, 0, 0, 6.0])
    assert float(box.lx) == 10.0 and float(box.ly) == 8.0 and float(box.lz) == 6.0 and float(box.volume) == 480.0
    dx = torch.tensor([[5.1, -4.1, 2.9], [-5.1, 4.1, -3.1]], dtype=torch.float64)
    np.testing.assert_allclose(box.apply_pbc(dx).numpy(), [[-4.9, 3.9, 2.9], [4.9, -3.9, 2.9]], rtol=1e-12)
    np.testing.assert_allclose(box.wrap_into_box(torch.tensor([[-1.0, 9.0, 6.5]], dtype=torch.float64)).numpy(),
                               [[9.0, 1.0, 0.5]], rtol=1e-12)


def test_runner_dataset(golden_dir):
    ds = Dataset.from_runner(golden_dir / "h2o.data")
    assert len(ds) == 2 and not ds.cache
    s = ds[1]
    assert s.natoms == 12 and s.get_unique_elements() == ("H", "O")
    frames = (golden_dir / "h2o.data").read_text().split("begin")[2]
    first_atom = [float(x) for x in frames.split("atom")[1].split()[:3]]
    np.testing.assert_allclose(s.positions[0].cpu().numpy(), np.remainder(first_atom, 11.8086403654), rtol=1e-12)
    assert float(s.total_energy) != 0.0 and s.forces.shape == (12, 3) and s.charges.shape == (12,)
    with pytest.raises(IndexError):
        ds[2]
    ds2 = Dataset.from_runner(golden_dir / "h2o.data", persist=True)
    assert ds2[0] is ds2[0] and 0 in ds2.cache
    ds2.preload()
    assert sorted(ds2.cache) == [0, 1]
    assert len(list(RunnerDataSource(golden_dir / "h2o.data").read_structures())) == 2
    s32 = Dataset.from_runner(golden_dir / "h2o.data", dtype=torch.float32)[0]
    assert s32.dtype == torch.float32


def test_tokenize():
    assert tokenize("atom 1 2 3\n") == ("atom", ["1", "2", "3"])
    assert tokenize("# comment", comment="#") == (None, [])
    assert tokenize("Cutoff_Type 2 # tanhu\n", comment="#") == ("cutoff_type", ["2"])
    assert tokenize("\n") == (None, [])


def test_settings_json_and_nn(golden_dir, tmp_path):
    s = NeuralNetworkPotentialSettings.from_file(golden_dir / "h2o.json")
    assert s.elements == ["H", "O"] and s.number_of_elements == 2 and s.cutoff_type == "tanhu"
    assert s.scaler_save_format == "scaling.{:03d}.json" and s.model_save_format == "weights.{:03d}.pkl"  # extras ignored
    assert s["scale_type"] == "center" and len(s.symfunction_short) == 7
    ang = s.symfunction_short[-1]
    assert (ang.central_element, ang.acsf_type, ang.neighbor_element_j, ang.neighbor_element_k) == ("O", 3, "O", "O")
    assert (ang.eta, ang.r_cutoff, ang.lambda0, ang.zeta) == (0.001, 12.0, -1.0, 4.0)
    nn = tmp_path / "input.nn"
    nn.write_text(
        "# RuNNer style\nnumber_of_elements 2\nelements O H\natom_energy H -0.45\nglobal_hidden_layers_short 2\n"
        "global_nodes_short 5 5\nglobal_activation_short t t l\ncutoff_type 2\nscale_symmetry_functions\n"
        "scale_min_short 0.0\nscale_max_short 1.0\nrandom_seed 7\nunknown_keyword 1\n"
        "symfunction_short H 2 O 0.001 0.5 12.0\nsymfunction_short O 3 H H 0.2 1.0 4.0 11.0\n")
    t = NeuralNetworkPotentialSettings.from_file(nn)
    assert t.elements == ["H", "O"] and t.global_activation_short == ["tanh", "tanh", "identity"]
    assert t.global_nodes_short == [5, 5] and t.cutoff_type == "tanhu" and t.random_seed == 7
    assert t.scale_type == "center"                  # the scaler switch keywords are dropped by the reader (App. B 12)
    rad, ang = t.symfunction_short
    assert (rad.eta, rad.r_shift, rad.r_cutoff) == (0.001, 0.5, 12.0)          # .nn order: eta r_shift r_cutoff
    assert (ang.eta, ang.lambda0, ang.zeta, ang.r_cutoff, ang.r_shift) == (0.2, 1.0, 4.0, 11.0, 0.0)
    with pytest.raises(ValueError):
        NeuralNetworkPotentialSettings.from_file(tmp_path / "pot.txt")
    out = tmp_path / "dump.json"
    s.to_json(out)
    assert NeuralNetworkPotentialSettings.from_json(out).symfunction_short == s.symfunction_short


def test_acsf_attributes_and_records():
    cfn = CutoffFunction.from_type("tanhu", r_cutoff=3.0)
    acsf = ACSF("Ne", tuple((G2(cfn, eta=1.0, r_shift=rs), NeighborElements("Ne")) for rs in (0.0, 0.25, 0.5, 0.75, 1.0)), ())
    assert (acsf.central_element, acsf.num_radial_symmetry_functions, acsf.num_angular_symmetry_functions,
            acsf.num_symmetry_functions, acsf.r_cutoff) == ("Ne", 5, 0, 5, 3.0)
    g2, g3, g9 = G2(cfn, 0.0, 0.001), G3(cfn, 0.2, 1.0, 1.0, 0.0), G9(cfn, 0.2, 4.0, -1.0, 0.0)  # positional orders
    assert (g2.r_shift, g2.eta) == (0.0, 0.001) and (g3.eta, g3.zeta, g3.lambda0) == (0.2, 1.0, 1.0) and g9.kind == 9
    a = ACSF("O", ((g2, NeighborElements("H")),), ((g3, NeighborElements("H", "H")), (g9, NeighborElements("H", "O"))))
    recs = a.symfunc_records()
    assert [r.kind for r in recs] == [2, 3, 9] and recs[2].neighbor_k == "O" and recs[0].neighbor_k is None
    with pytest.raises(KeyError):
        CutoffFunction.from_type("nope", 1.0)


def test_nn_model_param_layout_and_weights(golden_dir):
    model = NeuralNetworkModel(hidden_layers=((5, "tanh"), (5, "tanh")))
    assert model.output_layer == (1, "identity")
    shapes = model.param_shapes(3)
    assert shapes == {"layers_0": {"kernel": (3, 5), "bias": (5,)}, "layers_2": {"kernel": (5, 5), "bias": (5,)},
                      "layers_4": {"kernel": (5, 1), "bias": (1,)}}      # reference tests/test_nn.py:97-138
    params = model.load(golden_dir / "weights.001.pkl")                    # jax pickle read without jax
    ref = np.load(golden_dir / "weights.001.npz")
    for layer in shapes:
        assert params[layer]["kernel"].dtype == np.float32
        np.testing.assert_array_equal(params[layer]["kernel"], ref[f"{layer}.kernel"])
        assert not params[layer]["bias"].any()
    sizes, acts, w = model.flatten(params, 3)
    assert sizes == [3, 5, 5, 1] and acts == [1, 1, 0] and w.dtype == np.float64 and w.size == 15 + 5 + 25 + 5 + 5 + 1
    init = model.init_params(3, seed=1, weights_range=(-0.5, 0.5))
    assert abs(init["layers_0"]["kernel"]).max() <= 0.5 and not init["layers_2"]["bias"].any()
    with pytest.raises(ValueError):
        model.flatten(params, 4)
    with pytest.raises(KeyError):
        NeuralNetworkModel(hidden_layers=((5, "swish"),))


def test_potential_from_runner_host_side(golden_dir):
    nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
    assert nnp.num_elements == 2 and nnp.elements == ("H", "O") and nnp.r_cutoff == 12.0  # tests/test_nnp.py:36-42
    assert nnp.descriptors["H"].num_symmetry_functions == 3 and nnp.descriptors["O"].num_symmetry_functions == 4
    assert [type(sf).__name__ for sf, _ in nnp.descriptors["O"].radial_symmetry_functions] == ["G2", "G2"]
    assert nnp.models["O"].hidden_layers == ((5, "tanh"), (5, "tanh")) and nnp.scalers["H"].scale_type == "center"
    with pytest.raises(ValueError):                                   # scaler params not loaded (potential.py:319-327)
        nnp(Structure.from_dict(H2O))
    nnp.load()
    assert float(nnp.scalers_params["O"].dimension) == 4
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            nnp(Structure.from_dict(H2O))


def test_scaler_fit_partial_fit_and_transforms():
    """Reference tests/test_scaler.py: fit / partial_fit reproduce mean, sigma, min, max over batches."""
    rng = np.random.default_rng(0)
    data = torch.as_tensor(rng.normal(size=(18, 4)), dtype=torch.float64)
    p = DescriptorScaler.fit(data[:7])
    p = DescriptorScaler.partial_fit(p, data[7:17])
    p = DescriptorScaler.partial_fit(p, data[17])
    assert int(p.nsamples) == 18 and int(p.dimension) == 4
    np.testing.assert_allclose(p.mean.numpy(), data.mean(0).numpy(), rtol=1e-12)
    np.testing.assert_allclose(p.sigma.numpy(), data.std(0, unbiased=False).numpy(), rtol=1e-10)
    np.testing.assert_allclose(p.minval.numpy(), data.min(0).values.numpy())
    np.testing.assert_allclose(p.maxval.numpy(), data.max(0).values.numpy())
    x = data[:3]
    for kind in ("center", "scale", "scale_center", "scale_center_sigma"):
        sc = DescriptorScaler.from_type(kind, 0.0, 1.0)
        shift, slope, offset = sc.affine(p)
        np.testing.assert_allclose(sc(p, x).numpy(), offset + slope * (x.numpy() - shift), rtol=1e-12, atol=1e-14)
    sig = DescriptorScaler.from_type("scale_center_sigma")(p, x).numpy()
    np.testing.assert_allclose(sig, -(x.numpy() - p.mean.numpy()) / p.sigma.numpy(), rtol=1e-12)   # sign as written
    w = DescriptorScaler.check_warnings(p, data.max(0).values + 1.0, DescriptorScaler.initialize_warnings(0, 5))
    assert w.number_of_warnings == 1
    with pytest.raises(ValueError):
        DescriptorScaler.from_type("center", 1.0, 0.0)


def test_units_and_dtype_switch():
    assert units.BOLTZMANN_CONSTANT == 3.166811563e-6 and units.TO_FEMTO_SECOND == pytest.approx(2.418884326e-2)
    assert units.FROM_ATOMIC_MASS == pytest.approx(1.0 / 5.48579957163e-4)
    old = default_dtype.FLOATX
    try:
        default_dtype.FLOATX = torch.float32
        assert Structure.from_dict(H2O).dtype == torch.float32
    finally:
        default_dtype.FLOATX = old


def test_extension_switches_are_validated_on_the_host():
    """The extensions beyond the reference are opt-in and validated before any device work: unknown force definitions
    are rejected, HaloMD refuses to run without the CUDA library / a device (no CPU fallback), brick grids reject
    inconsistent layouts."""
    import pytest
    from pantea_b200.halo import BrickGrid, HaloMD, brick_dims
    from pantea_b200.simulation import MDSimulator
    sim = MDSimulator(time_step=0.5)
    assert sim.mass_scaled is False and sim.forces == "reference"          # defaults = the reference's integrator
    assert MDSimulator(0.5, mass_scaled=True, forces="full").forces == "full"
    with pytest.raises(ValueError):
        MDSimulator(0.5, forces="newton")
    assert brick_dims(8, [10.0, 10.0, 10.0]) == (2, 2, 2) and brick_dims(4, [10.0, 10.0, 40.0]) == (1, 1, 4)
    with pytest.raises(ValueError):
        BrickGrid([10.0, 10.0, 10.0], 8, dims=(2, 2, 1))
    with pytest.raises(ValueError):
        BrickGrid([10.0, 10.0, 10.0], 2, dims=(2, 1, 1), cuts=[[7.0, 3.0], [], []])
    grid = BrickGrid([10.0, 10.0, 10.0], 2, dims=(2, 1, 1), cuts=[[4.0], [], []])
    assert grid.bounds(0) == ([0.0, 0.0, 0.0], [4.0, 10.0, 10.0]) and grid.bounds(1)[0][0] == 4.0
    pos = torch.tensor([[3.999, 1.0, 1.0], [4.0, 1.0, 1.0], [9.999, 9.0, 9.0], [10.0, 0.0, 0.0]], dtype=torch.float64)
    assert grid.owner(pos).tolist() == [0, 1, 1, 0]                         # x == L wraps onto the first brick
    if not torch.cuda.is_available():
        z = torch.zeros((3, 3), dtype=torch.float64)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            HaloMD(None, z, z, torch.ones(3, dtype=torch.float64), torch.ones(3, dtype=torch.int32), [10.0] * 3, 0.25)
