"""GPU parity tests of the two SURVEY 8(e)/(f) extensions of the hot path:

* brick decomposition with ghost atoms (`pantea_b200.halo`): every brick's local evaluation -- owned + ghost atoms in
  the global periodic box, owned range only -- must reproduce the single-GPU neighbour sets (as sets of global ids,
  bit-exact) and forces / energies (<= 1e-10 relative), for 2, 4 and 8 bricks emulated on one GPU; the `HaloMD` driver
  (world = 1 here; NCCL transport in tests/mgpu_check.py, host logic under gloo in test_distributed_gloo.py) must follow
  the oracle's MD loop, including the skin / rollback machinery;
* `PANTEA_FORCE_FULL`: -dE/dr of the total energy against reverse-mode autograd through the dense oracle
  (`oracle/dense_oracle.py::energy_and_full_forces`), plus size-independent properties at 3 000 atoms.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle, dense_oracle
from oracle.spec import ElementSpec, SymFuncSpec, load_potential, md_velocities, rune_width_potential, water_box, water_masses
from tests.helpers import csr_rows, cuda, device_potential_from_specs, rel_err

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-10
FP32_TOL = 1e-5


@pytest.fixture(scope="module")
def pot(golden_dir):
    return load_potential(golden_dir / "h2o.json")


def _workspace(dev_pot, n, dtype=torch.float64, cap=None):
    from pantea_b200 import engine
    return engine.Workspace(dev_pot, max(n, 64), cap or min(max(n - 1, 32), 400), dtype)


# ------------------------------------------------------------------------------------------ brick decomposition
@pytest.mark.parametrize("n_atoms,world,balanced", [(3000, 2, False), (12000, 4, True), (12000, 8, False), (24000, 8, True)])
def test_brick_local_evaluation_matches_global(n_atoms, world, balanced, pot):
    from pantea_b200.halo import BrickGrid
    pos, types, box = water_box(n_atoms)
    dev = device_potential_from_specs(pot)
    rc = dev.r_cutoff
    p, t = cuda(pos), cuda(types, torch.int32)
    ws = _workspace(dev, n_atoms)
    ws.bind(p, t, box, rc)
    row_ptr, col = ws.neighbor_lists()
    rows_global = csr_rows(row_ptr.cpu().numpy(), col.cpu().numpy())
    _, ea_full, f_full = ws.energy_forces(True, True, True)
    ea_full, f_full = ea_full.cpu().numpy(), f_full.cpu().numpy()

    grid = BrickGrid.balanced(list(box), world, p) if balanced else BrickGrid(list(box), world)
    owner = grid.owner(p)
    if balanced:   # boundaries at the atom-count quantiles: the bricks own (almost) the same number of atoms
        counts = torch.bincount(owner, minlength=world)
        assert int(counts.max() - counts.min()) <= 0.05 * n_atoms / world + 8
    seen = np.zeros(n_atoms, dtype=int)
    e_sum = 0.0
    for r in range(world):
        own = torch.nonzero(owner == r, as_tuple=True)[0]
        ghost = torch.nonzero(grid.ghost_mask(p, owner, r, rc), as_tuple=True)[0]
        gid = torch.cat([own, ghost])
        n_own, n_loc = int(own.numel()), int(gid.numel())
        assert n_loc < n_atoms or world == 2          # a brick really sees only part of the box
        wl = _workspace(dev, n_loc)
        wl.bind(p[gid].contiguous(), t[gid].contiguous(), box, rc, owned=(0, n_own))
        rp_l, col_l = wl.neighbor_lists()
        rows_local = csr_rows(rp_l.cpu().numpy(), col_l.cpu().numpy())
        gid_h = gid.cpu().numpy()
        for a in range(0, n_own, max(1, n_own // 400)):      # neighbour sets as global ids: bit-exact
            assert np.array_equal(np.sort(gid_h[rows_local[a]]), rows_global[gid_h[a]])
        assert all(len(rows_local[a]) == 0 for a in range(n_own, n_loc, max(1, (n_loc - n_own) // 50 + 1)))
        e, ea, f = wl.energy_forces(True, True, True)
        own_h = own.cpu().numpy()
        assert rel_err(ea[:n_own].cpu().numpy(), ea_full[own_h]) < FP64_TOL
        assert np.abs(f[:n_own].cpu().numpy() - f_full[own_h]).max() < FP64_TOL * np.abs(f_full).max()
        e_sum += float(e)
        seen[own_h] += 1
    assert (seen == 1).all()
    assert abs(e_sum - ea_full.sum()) < FP64_TOL * np.abs(ea_full).sum()


def _md_arrays(n_atoms):
    pos, types, box = water_box(n_atoms)
    return pos, types, box, md_velocities(types), water_masses(types)


def _settled(make):
    """HaloMD built (and rebuilt) until no device-side capacity flag is raised by its first evaluation."""
    from pantea_b200 import _lib
    md = make()
    for _ in range(5):
        try:
            md.check_capacity()
            return md
        except _lib.CapacityError:
            md = make()
    raise AssertionError("capacities did not settle")


@pytest.mark.parametrize("n_atoms,n_steps,skin,every", [(192, 8, 0.0, 1), (3000, 6, 0.0, 1), (3000, 6, 1.0, 3),
                                                       (3000, 8, 0.02, 4)])
def test_halo_md_world1_follows_oracle(n_atoms, n_steps, skin, every, pot):
    """The HaloMD step sequence on one rank (no ghosts) against the oracle's MD loop; with skin 0.02 / rebuild every 4
    steps the ghost-shell criterion is violated and the driver must roll back and repeat the segment."""
    from pantea_b200.halo import HaloMD
    pos, types, box, vel, mass = _md_arrays(n_atoms)
    dev = device_potential_from_specs(pot)
    dt = 0.25
    args = (dev, cuda(pos), cuda(vel), cuda(mass), cuda(types, torch.int32), list(box), dt)
    md = _settled(lambda: HaloMD(*args, skin=skin, rebuild_every=every))
    for _ in range(n_steps):
        md.step()
    md.validate()
    md.check_capacity()
    assert md.steps == n_steps
    if skin == 0.02:
        assert md.rollbacks >= 1 and md.rebuild_every < every
    po, vo, fo, _ = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps)
    x = md.gather_owned(md.pos).cpu().numpy()
    d = x - po
    d -= np.asarray(box) * np.rint(d / np.asarray(box))
    assert np.abs(d).max() < 1e-9
    assert rel_err(md.gather_owned(md.vel).cpu().numpy(), vo) < 1e-8
    assert rel_err(md.gather_owned(md.frc).cpu().numpy(), fo) < 1e-7


# ------------------------------------------------------------------------------------------ full forces
def _full_forces_gpu(specs, pos, types, box, dtype=torch.float64, owned=None):
    dev = device_potential_from_specs(specs)
    ws = _workspace(dev, len(pos), dtype)
    ws.bind(cuda(pos, dtype), cuda(types, torch.int32), box, dev.r_cutoff, owned=owned)
    e, ea, f = ws.energy_forces(True, True, True, force_mode=1)
    return float(e), ea.cpu().numpy(), f.cpu().numpy()


def _full_forces_oracle(specs, pos, types, box):
    models = dense_oracle.models_from_specs(specs)
    b = torch.from_numpy(np.asarray(box, dtype=np.float64)) if box is not None else None
    e, f = dense_oracle.energy_and_full_forces(models, torch.from_numpy(np.asarray(pos, dtype=np.float64)),
                                               torch.from_numpy(np.asarray(types)), b)
    return float(e), f.numpy()


@pytest.mark.parametrize("n_atoms,use_box", [(24, True), (96, True), (192, True), (81, False)])
def test_full_forces_match_autograd_oracle(n_atoms, use_box, pot):
    pos, types, box = water_box(n_atoms, seed=4)
    box = box if use_box else None
    e, _, f = _full_forces_gpu(pot, pos, types, box)
    eo, fo = _full_forces_oracle(pot, pos, types, box)
    assert abs(e - eo) < FP64_TOL * max(abs(eo), 1.0)
    assert np.abs(f - fo).max() < FP64_TOL * np.abs(fo).max()
    assert np.abs(f.sum(0)).max() < 1e-12 * np.abs(f).sum()              # Newton's third law


def test_full_forces_wide_potential_and_mixed_kinds():
    """30 symmetry functions per element (G2, G3 with lambda = -1 / zeta = 2, G9 with zeta = 4, scale_center scaler,
    30-25-25-1 networks), then a potential mixing cutoff types, G1, a shorter second cutoff and a non-integer zeta."""
    pos, types, box = water_box(96, seed=3)
    specs = rune_width_potential()
    e, _, f = _full_forces_gpu(specs, pos, types, box)
    eo, fo = _full_forces_oracle(specs, pos, types, box)
    assert abs(e - eo) < FP64_TOL * max(abs(eo), 1.0) and np.abs(f - fo).max() < FP64_TOL * np.abs(fo).max()

    rng = np.random.default_rng(3)
    mixed = []
    for t in (1, 2):
        sfs = [SymFuncSpec(1, "cos", 9.0, 1), SymFuncSpec(2, "tanhu", 12.0, 2, 0, 0.02, 1.5),
               SymFuncSpec(2, "exp", 9.0, 1, 0, 0.05, 0.0), SymFuncSpec(3, "cos", 9.0, 1, 1, 0.01, 0.0, 1.0, 1.5),
               SymFuncSpec(3, "tanhu", 12.0, 1, 2, 0.02, 0.0, 1.0, 3.0), SymFuncSpec(9, "tanh", 12.0, 2, 2, 0.005, 0.0, 1.0, 1.0),
               SymFuncSpec(9, "cos", 9.0, 1, 2, 0.01, 0.0, -1.0, 2.0)]
        n = len(sfs)
        layers = [(rng.uniform(-1, 1, (n, 6)) / 3.0, rng.uniform(-0.1, 0.1, 6), "softplus"),
                  (rng.uniform(-1, 1, (6, 1)), np.zeros(1), "identity")]
        scaler = {"mean": rng.uniform(0, 1, n), "sigma": rng.uniform(0.5, 1, n), "minval": rng.uniform(-1, 0, n),
                  "maxval": rng.uniform(1, 2, n)}
        mixed.append(ElementSpec(t, sfs, "scale_center_sigma", scaler, 0.0, 1.0, layers))
    e, _, f = _full_forces_gpu(mixed, pos, types, box)
    eo, fo = _full_forces_oracle(mixed, pos, types, box)
    assert abs(e - eo) < FP64_TOL * max(abs(eo), 1.0) and np.abs(f - fo).max() < FP64_TOL * np.abs(fo).max()


def test_full_forces_properties_3000_atoms(pot):
    """Size-independent checks where the dense oracle is too slow: same energy as the reference mode, zero net force,
    central difference of the total energy along a random direction, FP32 mode, ownership blocks adding up."""
    pos, types, box = water_box(3000)
    e_ref, _, f_ref = c_oracle.energy_forces(pot, pos, types, box)
    e, ea, f = _full_forces_gpu(pot, pos, types, box)
    assert abs(e - e_ref) < FP64_TOL * np.abs(ea).sum()
    assert np.abs(f.sum(0)).max() < 1e-12 * np.abs(f).sum()
    assert np.abs(f - f_ref).max() > 1e-3 * np.abs(f_ref).max()          # it is a different quantity than the reference force
    rng = np.random.default_rng(1)
    u = rng.standard_normal(pos.shape)
    u /= np.linalg.norm(u)
    h = 1e-4
    ep = c_oracle.energy_forces(pot, np.remainder(pos + h * u, box), types, box, want_forces=False)[0]
    em = c_oracle.energy_forces(pot, np.remainder(pos - h * u, box), types, box, want_forces=False)[0]
    assert abs(-(ep - em) / (2 * h) - float((f * u).sum())) < 1e-6 * np.abs(f).max() * np.sqrt(f.size)
    _, _, f32 = _full_forces_gpu(pot, pos, types, box, torch.float32)
    assert np.abs(f32 - f).max() < FP32_TOL * np.abs(f).max()
    parts = sum(_full_forces_gpu(pot, pos, types, box, owned=blk)[2] for blk in ((0, 1100), (1100, 3000)))
    assert np.abs(parts - f).max() < FP64_TOL * np.abs(f).max()


def test_compute_forces_full_through_the_public_api(golden_dir):
    from pantea_b200.datasets import Dataset
    from pantea_b200.potentials import NeuralNetworkPotential
    nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
    nnp.load()
    s = Dataset.from_runner(golden_dir / "h2o.data")[0]
    f_ref = nnp.compute_forces(s).cpu().numpy()
    f_full = nnp.compute_forces(s, forces="full").cpu().numpy()
    specs = load_potential(golden_dir / "h2o.json")
    frame_pos = s.positions.cpu().numpy()
    box = torch.diagonal(s.box.lattice).cpu().numpy()
    types = np.asarray([1 if el == "H" else 2 for el in s.get_elements()], dtype=np.int32)
    _, fo = _full_forces_oracle(specs, frame_pos, types, box)
    assert np.abs(f_full - fo).max() < FP64_TOL * np.abs(fo).max()
    import json
    fx = json.loads((golden_dir / "extension_vectors.json").read_text())          # committed known answer
    assert np.abs(f_full - np.asarray(fx["full_forces"])).max() < FP64_TOL * np.abs(fo).max()
    assert np.abs(f_full.sum(0)).max() < 1e-13 and np.abs(f_ref.sum(0)).max() > 1e-3   # SURVEY App. C remark
    with pytest.raises(ValueError):
        nnp.compute_forces(s, forces="newton")


# ------------------------------------------------------------------------------------------ mass-scaled / full-force MD
def _md_run_gpu(specs, pos, vel, mass, types, box, dt, n_steps, mass_scaled, force_mode, use_graph=1):
    import ctypes as C
    from pantea_b200 import _lib
    dev = device_potential_from_specs(specs)
    n = len(pos)
    ws = _workspace(dev, n)
    p, v, m, t = cuda(pos), cuda(vel), cuda(mass), cuda(types, torch.int32)
    ws.bind(p, t, box, dev.r_cutoff)
    _, _, f = ws.energy_forces(False, True, force_mode=force_mode)
    scal = torch.zeros((n_steps, 2), dtype=torch.float64, device="cuda")
    params = _lib.MDParams(dt, 0.0, 0.0, 3.166811563e-6, 1, use_graph, 1 if mass_scaled else 0, force_mode)
    _lib.check(_lib.load().pantea_md_run(ws.handle, _lib.ptr(p), _lib.ptr(v), _lib.ptr(f), _lib.ptr(m), _lib.ptr(t), n,
                                         _lib.box_arg(box), n_steps, C.byref(params), _lib.ptr(scal), _lib.stream_ptr()))
    _lib.check(_lib.load().pantea_neighbor_status(ws.handle, None, _lib.stream_ptr()))
    return p.cpu().numpy(), v.cpu().numpy(), f.cpu().numpy(), scal.cpu().numpy()


@pytest.mark.parametrize("use_graph", [0, 1])
def test_md_run_mass_scaled_matches_oracle(use_graph, pot):
    """Integrator extension alone: reference (central-role) forces, accelerations F/m, against the oracle's MD loop."""
    n_atoms, n_steps, dt = 192, 12, 5.0
    pos, types, box, vel, mass = _md_arrays(n_atoms)
    po, vo, fo, sc = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps, mass_scaled=True)
    p, v, f, s = _md_run_gpu(pot, pos, vel, mass, types, box, dt, n_steps, True, 0, use_graph)
    d = p - po
    d -= np.asarray(box) * np.rint(d / np.asarray(box))
    assert np.abs(d).max() < 1e-10 and rel_err(v, vo) < 1e-10 and rel_err(f, fo) < 1e-9
    assert rel_err(s[:, 0], sc[1:, 0]) < 1e-10 and rel_err(s[:, 1], sc[1:, 1]) < 1e-10
    # and it is a different trajectory than the reference integrator's
    po_ref, _, _, _ = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps)
    assert np.abs(po_ref - po).max() > 1e-3


def test_md_full_forces_mass_scaled_follows_dense_oracle_and_conserves_energy(pot):
    """Full force + F/m = the usual velocity Verlet: parity with the dense autograd oracle, and E_pot + E_kin stays
    put while both change by ~4 Ha (the reference force is not a gradient of E: with it the same run does not conserve)."""
    n_atoms, n_steps, dt = 48, 60, 5.0
    pos, types, box, vel, mass = _md_arrays(n_atoms)
    models = dense_oracle.models_from_specs(pot)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    xo, vo, fo, sc = dense_oracle.md_run_full(models, T(pos), T(vel), T(mass), torch.from_numpy(types), T(box), dt, n_steps)
    p, v, f, s = _md_run_gpu(pot, pos, vel, mass, types, box, dt, n_steps, True, 1)
    d = p - xo.numpy()
    d -= np.asarray(box) * np.rint(d / np.asarray(box))
    assert np.abs(d).max() < 1e-8 and rel_err(v, vo.numpy()) < 1e-8 and rel_err(f, fo.numpy()) < 1e-7
    sc = sc.numpy()
    swing = sc[:, 1].max() - sc[:, 1].min()
    assert rel_err(s[:, 0], sc[1:, 0]) < 1e-8 and rel_err(s[:, 1], sc[1:, 1]) < 1e-8
    e_tot = s.sum(1)
    assert swing > 1.0 and np.abs(e_tot - sc[0].sum()).max() < 5e-3 * swing
    _, _, _, s_ref = _md_run_gpu(pot, pos, vel, mass, types, box, dt, n_steps, True, 0)
    assert np.abs(s_ref.sum(1) - sc[0].sum()).max() > 5 * np.abs(e_tot - sc[0].sum()).max()


def test_md_simulator_extensions_through_the_public_api(golden_dir):
    """MDSimulator(mass_scaled=True, forces="full"): the call-by-call path and the device-resident loop agree."""
    from pantea_b200.atoms import Structure
    from pantea_b200.potentials import NeuralNetworkPotential
    from pantea_b200.simulation import MDSimulator, System
    nnp = NeuralNetworkPotential.from_runner(golden_dir / "h2o.json")
    nnp.load()
    pos, types, box = water_box(96)
    make = lambda: Structure.from_dict({"positions": pos, "elements": ["H" if x == 1 else "O" for x in types],  # noqa: E731
                                        "lattice": np.diag(box)})
    a = System.from_structure(make(), nnp, temperature=300.0, seed=3)
    b = System.from_structure(make(), nnp, temperature=300.0, seed=3)
    sim_a = MDSimulator(time_step=5.0, mass_scaled=True, forces="full")
    sim_b = MDSimulator(time_step=5.0, mass_scaled=True, forces="full")
    for _ in range(6):
        sim_a.simulate_one_step(a)
    sim_b.simulate_steps(b, 6)
    assert (a.positions - b.positions).abs().max() < 1e-10
    assert (a.velocities - b.velocities).abs().max() < 1e-10 * a.velocities.abs().max() + 1e-15
    with pytest.raises(ValueError):
        MDSimulator(time_step=1.0, forces="newton")

