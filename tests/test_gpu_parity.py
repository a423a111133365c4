"""Parity tests proper: CUDA path (through the C ABI) against the oracle on the same inputs.

Tolerances: neighbour sets bit-exact; descriptors, gradients, energies, forces within 1e-10
relative in FP64 mode and 1e-5 in FP32 mode (BASELINE.json north_star).  "Relative" is measured
against the largest magnitude of the compared quantity (per symmetry-function column for
descriptors), which is how a force/descriptor array is meaningfully compared.
"""
import json

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle.spec import (ElementSpec, SymFuncSpec, load_potential, md_velocities, read_runner, rune_width_potential,
                         type_map, water_box, water_masses, KB)
from tests.helpers import csr_rows, cuda, device_potential_from_specs, rel_err

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-10
FP32_TOL = 1e-5


@pytest.fixture(scope="module")
def vec(golden_dir):
    return json.loads((golden_dir / "reference_vectors.json").read_text())


@pytest.fixture(scope="module")
def pot(golden_dir):
    return load_potential(golden_dir / "h2o.json")


@pytest.fixture(scope="module")
def h2o(golden_dir):
    f = read_runner(golden_dir / "h2o.data")[0]
    f["positions"] = np.remainder(f["positions"], f["box"])
    return f


def _workspace(dev_pot, n, dtype=torch.float64, cap=None):
    from pantea_b200 import engine
    return engine.Workspace(dev_pot, max(n, 64), cap or min(max(n - 1, 32), 400), dtype)


# ------------------------------------------------------------------------------------------ neighbours
@pytest.mark.parametrize("n_atoms,rc,use_box", [(12, 12.0, True), (192, 12.0, True), (3000, 12.0, True),
                                               (3000, 6.5, True), (648, 9.0, False), (6000, 12.0, True)])
def test_neighbor_sets_bit_exact(n_atoms, rc, use_box, h2o):
    if n_atoms == 12:
        pos, types, box = h2o["positions"], h2o["types"], h2o["box"]
    else:
        pos, types, box = water_box(n_atoms)
    box_arg = box if use_box else None
    row_ptr_o, col_o = c_oracle.neighbors(pos, types, box_arg, rc)
    ws = _workspace(None, n_atoms)
    ws.bind(cuda(pos), cuda(types, torch.int32), box_arg, rc)
    row_ptr, col = ws.neighbor_lists()
    assert np.array_equal(row_ptr.cpu().numpy(), row_ptr_o)
    assert np.array_equal(col.cpu().numpy(), col_o)


def _assert_sets_equal(pos, types, box, rc, cap=None):
    row_ptr_o, col_o = c_oracle.neighbors(pos, types, box, rc)
    n = len(pos)
    ws = _workspace(None, n, cap=cap)
    ws.bind(cuda(pos), cuda(types, torch.int32), box, rc)
    row_ptr, col = ws.neighbor_lists()
    assert np.array_equal(row_ptr.cpu().numpy(), row_ptr_o)
    assert np.array_equal(col.cpu().numpy(), col_o)
    return row_ptr_o


def test_neighbor_sets_distances_exactly_at_cutoff():
    """Simple-cubic lattice, spacing 3 Bohr, rc = 12: (4,0,0)a lies exactly on the cutoff sphere (included: r <= rc),
    and a copy of the lattice nudged by single ulps puts pairs one rounding step either side of it."""
    m, a, rc = 15, 3.0, 12.0
    g = np.arange(m, dtype=np.float64) * a
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    types = np.where(np.arange(len(pos)) % 3 == 0, 2, 1).astype(np.int32)
    box = np.array([m * a] * 3)
    rp = _assert_sets_equal(pos, types, box, rc, cap=320)
    assert (np.diff(rp) == 256).all()  # 257 lattice points within the closed ball, minus the atom itself
    rng = np.random.default_rng(5)
    nudged = pos.copy()
    sel = rng.random(len(pos)) < 0.5
    nudged[sel] = np.nextafter(nudged[sel], np.where(rng.random((sel.sum(), 3)) < 0.5, -np.inf, np.inf))
    _assert_sets_equal(nudged, types, box, rc, cap=320)


@pytest.mark.parametrize("shift,frac", [(0.2, 1.0), (-0.2, 1.0), (0.6, 0.3), (3.3, 0.05), (-1.7, 1.0)])
def test_neighbor_sets_atoms_outside_the_box(shift, frac):
    """The reference applies one box shift to raw coordinate differences (box.py:112-117); atoms far outside the cell
    therefore lose neighbours.  The FP32 screening pass measures true minimum-image distances and must hand over to
    the exact scan exactly when the two notions can differ."""
    pos, types, box = water_box(3000)
    rng = np.random.default_rng(11)
    moved = rng.random(len(pos)) < frac
    pos = pos.copy()
    pos[moved] += shift * box
    _assert_sets_equal(pos, types, box, 12.0)


def test_neighbor_sets_pairs_within_rounding_of_cutoff_in_a_large_box():
    """Pairs at r = rc (1 + delta), |delta| from 1e-15 to 1e-4, at coordinates up to 400 Bohr: the FP32 screen (absolute
    coordinate error ~2e-5 Bohr here) must route every one of them through the exact FP64 predicate."""
    rng = np.random.default_rng(17)
    L, rc = 400.0, 12.0
    deltas = np.array([0.0, 1e-15, -1e-15, 1e-13, -1e-13, 1e-10, -1e-10, 1e-8, -1e-8, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5,
                       1e-4, -1e-4, 1e-3, -1e-3])
    sites = rng.random((40 * len(deltas), 3)) * L
    u = rng.standard_normal(sites.shape)
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    partner = np.remainder(sites + rc * (1.0 + np.tile(deltas, 40))[:, None] * u, L)
    filler = rng.random((1500, 3)) * L
    pos = np.concatenate([sites, partner, filler])
    types = rng.integers(1, 3, len(pos)).astype(np.int32)
    _assert_sets_equal(pos, types, np.array([L, L, L]), rc, cap=64)


def test_neighbor_capacity_overflow_is_reported_and_grown():
    from pantea_b200 import _lib, engine
    pos, types, box = water_box(192)
    ws = engine.Workspace(None, 192, 32, torch.float64)
    with pytest.raises(_lib.CapacityError):
        ws.bind(cuda(pos), cuda(types, torch.int32), box, 12.0, check=False)
        import ctypes as C
        _lib.check(_lib.load().pantea_neighbor_status(ws.handle, C.byref(C.c_int32(0)), _lib.stream_ptr()))
    ws.bind(cuda(pos), cuda(types, torch.int32), box, 12.0)  # grows and succeeds
    row_ptr_o, col_o = c_oracle.neighbors(pos, types, box, 12.0)
    row_ptr, col = ws.neighbor_lists()
    assert np.array_equal(col.cpu().numpy(), col_o)


def test_distances_match_oracle(h2o, vec):
    ws = _workspace(None, 12)
    ws.bind(cuda(h2o["positions"]), cuda(h2o["types"], torch.int32), h2o["box"], 1e-6, check=False)
    r, d = ws.distances(None, None, True)
    assert np.array_equal(r.cpu().numpy(), c_oracle.distances(h2o["positions"], h2o["box"]))
    np.testing.assert_allclose(r[0, :5].cpu().numpy(), vec["notebook_distances"]["expected"], atol=5e-9)
    assert d.shape == (12, 12, 3)


# ------------------------------------------------------------------------------------------ descriptors
def _acsf_gpu(spec, pos, types, box, centres=None, dtype=torch.float64):
    dev = device_potential_from_specs([spec])
    ws = _workspace(dev, len(pos), dtype)
    rc = max(s.r_cutoff for s in spec.symfuncs)
    ws.bind(cuda(pos, dtype), cuda(types, torch.int32), box, rc)
    c = None if centres is None else cuda(centres, torch.int32)
    G, dG = ws.acsf(spec.atom_type - 1, len(spec.symfuncs), c, True, True)
    return G.cpu().numpy(), dG.cpu().numpy()


def _assert_descriptor_close(G, dG, G_o, dG_o, tol):
    col_scale = np.abs(G_o).max(axis=0) + 1e-300
    assert np.abs((G - G_o) / col_scale).max() < tol
    g_scale = np.abs(dG_o).max(axis=(0, 2))[None, :, None] + 1e-300
    assert np.abs((dG - dG_o) / g_scale).max() < tol


def test_acsf_golden_vectors(vec, h2o):
    tm = type_map(h2o["elements"])
    v = vec["h2o_pbc_g2_g3"]
    ct, rc = v["cutoff"]
    spec = ElementSpec(tm["O"], [
        SymFuncSpec(2, ct, rc, tm["H"], 0, v["radial"]["eta"], v["radial"]["r_shift"]),
        SymFuncSpec(3, ct, rc, tm["H"], tm["H"], v["angular"]["eta"], 0.0, v["angular"]["lambda0"], v["angular"]["zeta"])])
    centres = np.nonzero(h2o["types"] == tm["O"])[0]
    G, _ = _acsf_gpu(spec, h2o["positions"], h2o["types"], h2o["box"], centres)
    assert G.shape == tuple(v["shape"])
    np.testing.assert_allclose(G[0], v["expected_atom0"], rtol=0, atol=6e-11)
    # notebook: values for O atoms and gradient row of atom 0 (all atoms are centres for grad)
    v = vec["notebook_acsf"]
    ct, rc = v["cutoff"]
    sfs = [SymFuncSpec(2, ct, rc, tm[r["neighbor"]], 0, r["eta"], r["r_shift"]) for r in v["radial"]]
    sfs += [SymFuncSpec(a["kind"], ct, rc, tm[a["neighbors"][0]], tm[a["neighbors"][1]], a["eta"], 0.0, a["lambda0"],
                        a["zeta"]) for a in v["angular"]]
    spec = ElementSpec(tm["O"], sfs)
    G, dG = _acsf_gpu(spec, h2o["positions"], h2o["types"], h2o["box"], None)
    np.testing.assert_allclose(G[centres], v["expected_values"], rtol=2e-8)
    np.testing.assert_allclose(dG[0], v["expected_grad_atom0"], atol=6e-9)


def test_acsf_ne2_without_box(vec):
    v = vec["ne2_g2"]
    spec = ElementSpec(1, [SymFuncSpec(2, v["cutoff"][0], v["cutoff"][1], 1, 0, v["eta"], rs) for rs in v["r_shifts"]])
    G, _ = _acsf_gpu(spec, np.asarray(v["positions"]), np.ones(2, dtype=np.int32), None)
    np.testing.assert_allclose(G, np.tile(v["expected_row"], (2, 1)), rtol=1e-8)


@pytest.mark.parametrize("cutoff", ["hard", "cos", "tanhu", "tanh", "exp", "poly1", "poly2"])
def test_acsf_all_cutoffs_vs_oracle(cutoff):
    pos, types, box = water_box(192, seed=11)
    rc = 0.9 if cutoff.startswith("poly") else 9.0
    if cutoff.startswith("poly"):
        pos, box = pos / 12.0, box / 12.0
    spec = ElementSpec(2, [
        SymFuncSpec(1, cutoff, rc, 1), SymFuncSpec(2, cutoff, rc * 0.8, 2, 0, 0.05, 0.5),
        SymFuncSpec(3, cutoff, rc, 1, 1, 0.01, 0.0, -1.0, 4.0), SymFuncSpec(3, cutoff, rc * 0.9, 1, 2, 0.02, 0.0, 1.0, 2.0),
        SymFuncSpec(9, cutoff, rc, 2, 2, 0.005, 0.0, 1.0, 1.0), SymFuncSpec(9, cutoff, rc, 2, 1, 0.03, 0.0, -1.0, 3.0),
        SymFuncSpec(3, cutoff, rc, 2, 2, 0.03, 0.0, 1.0, 2.5)])
    G_o, dG_o = c_oracle.acsf(spec, pos, types, box)
    G, dG = _acsf_gpu(spec, pos, types, box)
    _assert_descriptor_close(G, dG, G_o, dG_o, FP64_TOL)


# ------------------------------------------------------------------------------------------ energy / forces
def _energy_forces_gpu(specs, pos, types, box, dtype=torch.float64):
    dev = device_potential_from_specs(specs)
    ws = _workspace(dev, len(pos), dtype)
    ws.bind(cuda(pos, dtype), cuda(types, torch.int32), box, dev.r_cutoff)
    e, ea, f = ws.energy_forces(True, True, True)
    return float(e), ea.cpu().numpy(), f.cpu().numpy()


def test_nnp_golden_energy_forces(vec, h2o, pot):
    e, ea, f = _energy_forces_gpu(pot, h2o["positions"], h2o["types"], h2o["box"])
    v = vec["nnp_fp32"]
    np.testing.assert_allclose(e, v["energy"], rtol=0, atol=3e-7)   # float32 golden, see test_oracle_golden.py
    np.testing.assert_allclose(f, v["forces"], rtol=1e-5, atol=1e-7)
    eo, eao, fo = c_oracle.energy_forces(pot, h2o["positions"], h2o["types"], h2o["box"])
    assert rel_err(ea, eao) < FP64_TOL and rel_err(f, fo) < FP64_TOL
    assert abs(e - eo) < FP64_TOL * np.abs(eao).sum()


@pytest.mark.parametrize("n_atoms", [24, 192, 3000])
def test_energy_forces_water_vs_oracle(n_atoms, pot):
    pos, types, box = water_box(n_atoms)
    e, ea, f = _energy_forces_gpu(pot, pos, types, box)
    eo, eao, fo = c_oracle.energy_forces(pot, pos, types, box)
    assert rel_err(ea, eao) < FP64_TOL
    assert rel_err(f, fo) < FP64_TOL
    assert abs(e - eo) < FP64_TOL * np.abs(eao).sum()


@pytest.mark.parametrize("scale_type", ["scale", "scale_center", "scale_center_sigma"])
def test_energy_forces_scalers(scale_type, pot):
    pos, types, box = water_box(96, seed=5)
    specs = [ElementSpec(s.atom_type, s.symfuncs, scale_type, s.scaler, 0.0, 1.0, s.layers) for s in pot]
    e, ea, f = _energy_forces_gpu(specs, pos, types, box)
    eo, eao, fo = c_oracle.energy_forces(specs, pos, types, box)
    assert rel_err(ea, eao) < FP64_TOL and rel_err(f, fo) < FP64_TOL


def test_energy_forces_wide_potential_all_activations():
    pos, types, box = water_box(192, seed=3)
    specs = rune_width_potential()
    acts = ["logistic", "softplus", "gaussian", "cos", "exp", "harmonic", "relu", "tanh"]
    for i, s in enumerate(specs):
        s.layers = [(k, b, acts[(2 * i + l) % len(acts)] if l < 2 else "identity") for l, (k, b, _) in enumerate(s.layers)]
    e, ea, f = _energy_forces_gpu(specs, pos, types, box)
    eo, eao, fo = c_oracle.energy_forces(specs, pos, types, box)
    assert rel_err(ea, eao) < FP64_TOL and rel_err(f, fo) < FP64_TOL


def test_energy_forces_no_box(pot):
    pos, types, _ = water_box(81, seed=9)
    e, ea, f = _energy_forces_gpu(pot, pos, types, None)
    eo, eao, fo = c_oracle.energy_forces(pot, pos, types, None)
    assert rel_err(ea, eao) < FP64_TOL and rel_err(f, fo) < FP64_TOL


def test_energy_forces_fp32_mode(pot):
    pos, types, box = water_box(192)
    e, ea, f = _energy_forces_gpu(pot, pos, types, box, torch.float32)
    eo, eao, fo = c_oracle.energy_forces(pot, pos, types, box)
    assert rel_err(ea, eao) < FP32_TOL and rel_err(f, fo) < FP32_TOL


def test_results_are_bitwise_reproducible(pot):
    pos, types, box = water_box(3000)
    a = _energy_forces_gpu(pot, pos, types, box)
    b = _energy_forces_gpu(pot, pos, types, box)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("n_atoms,blocks", [(3000, ((0, 1700), (1700, 3000))),
                                            (12000, ((0, 1500), (1500, 7000), (7000, 12000)))])
def test_owned_range_partition_matches_full(n_atoms, blocks, pot):
    """Multi-GPU ownership: evaluating the atoms in index blocks gives bitwise the same per-atom results as one pass
    over all atoms (12 000 atoms: the blocks fall below the size where a lone system would switch to several warps per
    atom -- the choice must follow the system, not the share)."""
    pos, types, box = water_box(n_atoms)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms)
    p, t = cuda(pos), cuda(types, torch.int32)
    ws.bind(p, t, box, dev.r_cutoff)
    _, ea_full, f_full = ws.energy_forces(True, True, True)
    ea, f = torch.zeros_like(ea_full), torch.zeros_like(f_full)
    for lo, hi in blocks:
        ws.bind(p, t, box, dev.r_cutoff, owned=(lo, hi))
        e_part, ea_part, f_part = ws.energy_forces(True, True, True)
        ea[lo:hi], f[lo:hi] = ea_part[lo:hi], f_part[lo:hi]
        assert abs(float(e_part) - float(ea_full[lo:hi].sum())) < 1e-12
    assert torch.equal(ea, ea_full) and torch.equal(f, f_full)


# ------------------------------------------------------------------------------------------ batch (dataset preprocessing)
def test_batched_structures_match_single(pot):
    structs = [water_box(192, seed=2024 + s) for s in range(5)]
    pos = np.concatenate([s[0] for s in structs])
    types = np.concatenate([s[1] for s in structs])
    boxes = np.stack([s[2] for s in structs])
    ptr = np.arange(6, dtype=np.int32) * 192
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, len(pos))
    ws.bind_batch(cuda(pos), cuda(types, torch.int32), cuda(ptr, torch.int32), cuda(boxes), dev.r_cutoff)
    for spec in pot:
        G, dG = ws.acsf(spec.atom_type - 1, len(spec.symfuncs), None, True, True)
        G, dG = G.cpu().numpy(), dG.cpu().numpy()
        for s, (p_s, t_s, b_s) in enumerate(structs):
            G_o, dG_o = c_oracle.acsf(spec, p_s, t_s, b_s)
            _assert_descriptor_close(G[s * 192:(s + 1) * 192], dG[s * 192:(s + 1) * 192], G_o, dG_o, FP64_TOL)


# ------------------------------------------------------------------------------------------ MD
@pytest.mark.parametrize("n_atoms,n_steps,tau", [(24, 20, 0.0), (192, 10, 25.0), (3000, 4, 0.0)])
def test_md_run_matches_oracle(n_atoms, n_steps, tau, pot):
    import ctypes as C
    from pantea_b200 import _lib
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    dt, t0 = 0.25, 300.0
    p_o, v_o, f_o, sc_o = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps, t0, tau, KB)
    dev = device_potential_from_specs(pot)
    for use_graph in (0, 1):
        ws = _workspace(dev, n_atoms)
        p, v, t, m = cuda(pos), cuda(vel), cuda(types, torch.int32), cuda(mass)
        ws.bind(p, t, box, dev.r_cutoff)
        _, _, f = ws.energy_forces(False, True)
        scal = torch.zeros((n_steps, 2), dtype=torch.float64, device="cuda")
        params = _lib.MDParams(dt, t0, tau, KB, 1, use_graph)
        _lib.check(_lib.load().pantea_md_run(ws.handle, _lib.ptr(p), _lib.ptr(v), _lib.ptr(f), _lib.ptr(m), _lib.ptr(t),
                                             n_atoms, _lib.box_arg(box), n_steps, C.byref(params), _lib.ptr(scal),
                                             _lib.stream_ptr()))
        torch.cuda.synchronize()
        np.testing.assert_allclose(p.cpu().numpy(), p_o, rtol=0, atol=1e-10)
        assert rel_err(v.cpu().numpy(), v_o) < 1e-9
        assert rel_err(f.cpu().numpy(), f_o) < 1e-8
        np.testing.assert_allclose(scal.cpu().numpy()[:, 0], sc_o[1:, 0], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(scal.cpu().numpy()[:, 1], sc_o[1:, 1], rtol=1e-9)


def test_md_energy_curve_3000_atoms_matches_oracle(pot):
    """BASELINE.json configs[2] (bounded): 3000-atom water box, NVE, dt = 0.25 a.u.  The potential- and
    kinetic-energy curves of the device-resident MD loop must follow the oracle's step by step."""
    import ctypes as C
    from pantea_b200 import _lib
    n_atoms, n_steps, dt = 3000, 60, 0.25
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    _, _, _, sc_o = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps, 0.0, 0.0, KB)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms)
    p, v, t, m = cuda(pos), cuda(vel), cuda(types, torch.int32), cuda(mass)
    ws.bind(p, t, box, dev.r_cutoff)
    _, _, f = ws.energy_forces(False, True)
    scal = torch.zeros((n_steps, 2), dtype=torch.float64, device="cuda")
    params = _lib.MDParams(dt, 0.0, 0.0, KB, 1, 1)
    _lib.check(_lib.load().pantea_md_run(ws.handle, _lib.ptr(p), _lib.ptr(v), _lib.ptr(f), _lib.ptr(m), _lib.ptr(t), n_atoms,
                                         _lib.box_arg(box), n_steps, C.byref(params), _lib.ptr(scal), _lib.stream_ptr()))
    _lib.check(_lib.load().pantea_neighbor_status(ws.handle, None, _lib.stream_ptr()))
    s = scal.cpu().numpy()
    e_tot, e_tot_o = s[:, 0] + s[:, 1], sc_o[1:, 0] + sc_o[1:, 1]
    scale = np.abs(sc_o[:, 0]).max()
    assert np.abs(s[:, 0] - sc_o[1:, 0]).max() < 1e-9 * scale
    assert np.abs(e_tot - e_tot_o).max() < 1e-9 * scale          # same drift curve (the reference MD is not conservative)
    assert np.abs(e_tot_o - e_tot_o[0]).max() > 1e-6 * scale     # ... and it does drift: the test is not vacuous


# ------------------------------------------------------------------------------------------ Verlet skin
@pytest.mark.parametrize("use_graph,skin", [(0, 0.6), (1, 0.6), (1, 0.05)])
def test_md_run_with_verlet_skin_matches_oracle(use_graph, skin, pot):
    """rows gathered with rc + skin and reused until an atom moved more than skin / 2 (decided on the
    device).  Neighbours beyond a cutoff contribute exactly zero, so the trajectory equals the oracle's (which rebuilds
    every step) up to summation order.  skin = 0.05 forces several device-side rebuilds within the run."""
    import ctypes as C
    from pantea_b200 import _lib
    n_atoms, n_steps, dt = 3000, 30, 0.25
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    p_o, v_o, f_o, sc_o = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps, 0.0, 0.0, KB)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms)
    ws.set_skin(skin)
    p, v, t, m = cuda(pos), cuda(vel), cuda(types, torch.int32), cuda(mass)
    ws.bind(p, t, box, dev.r_cutoff)
    _, _, f = ws.energy_forces(False, True)
    scal = torch.zeros((n_steps, 2), dtype=torch.float64, device="cuda")
    params = _lib.MDParams(dt, 0.0, 0.0, KB, 1, use_graph)
    for _ in range(2):  # a capacity report after the first attempt raises the capacity; the run is then repeated
        p.copy_(cuda(pos)); v.copy_(cuda(vel))
        ws.bind(p, t, box, dev.r_cutoff)
        _, _, f = ws.energy_forces(False, True)
        b0, r0 = ws.rebuild_counts()
        _lib.check(_lib.load().pantea_md_run(ws.handle, _lib.ptr(p), _lib.ptr(v), _lib.ptr(f), _lib.ptr(m), _lib.ptr(t),
                                             n_atoms, _lib.box_arg(box), n_steps, C.byref(params), _lib.ptr(scal),
                                             _lib.stream_ptr()))
        code = _lib.load().pantea_neighbor_status(ws.handle, None, _lib.stream_ptr())
        if code != _lib.PANTEA_ECAPACITY:
            _lib.check(code)
            break
    builds, rebuilds = ws.rebuild_counts()
    assert builds - b0 == n_steps
    # the reference integrator carries no mass: atoms cover 0.3 Bohr within ~6 steps, so even the wide skin rebuilds
    if skin > 0.5:
        assert 1 <= rebuilds - r0 <= n_steps // 3  # rows reused for several steps at a time
    else:
        assert n_steps // 3 < rebuilds - r0 <= n_steps  # rebuilt on demand
    np.testing.assert_allclose(p.cpu().numpy(), p_o, rtol=0, atol=1e-9)
    assert rel_err(v.cpu().numpy(), v_o) < 1e-9
    assert rel_err(f.cpu().numpy(), f_o) < 1e-8
    np.testing.assert_allclose(scal.cpu().numpy()[:, 0], sc_o[1:, 0], rtol=1e-9, atol=1e-10)


def test_verlet_skin_single_evaluations_and_rebuild_trigger(pot):
    """Energy / forces with a skin equal the oracle's for the bound structure, after small moves (rows reused) and after
    a move beyond skin / 2 (rows rebuilt); exact-set queries are refused while a skin is active."""
    n_atoms, skin = 3000, 0.5
    pos, types, box = water_box(n_atoms)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms)
    ws.set_skin(skin)
    t = cuda(types, torch.int32)
    rng = np.random.default_rng(3)
    expected_rebuilds = 0
    for step, amp in enumerate([0.0, 0.05, 0.05, 0.4, 0.02]):
        pos = np.remainder(pos + amp * rng.uniform(-1, 1, pos.shape) / np.sqrt(3.0), box)
        p = cuda(pos)
        ws.bind(p, t, box, dev.r_cutoff)
        e, _, f = ws.energy_forces(True, True)
        e_o, _, f_o = c_oracle.energy_forces(pot, pos, types, box)
        assert abs(float(e) - e_o) <= FP64_TOL * abs(e_o)
        assert rel_err(f.cpu().numpy(), f_o) < FP64_TOL
    builds, rebuilds = ws.rebuild_counts()
    assert builds >= 5 and 2 <= rebuilds < builds  # first build + the 0.4-Bohr move (cumulative drift may add one)
    with pytest.raises(ValueError):
        ws.neighbor_lists()
    ws.set_skin(0.0)
    ws.bind(cuda(pos), t, box, dev.r_cutoff)
    row_ptr_o, col_o = c_oracle.neighbors(pos, types, box, dev.r_cutoff)
    row_ptr, col = ws.neighbor_lists()
    assert np.array_equal(col.cpu().numpy(), col_o)


@pytest.mark.parametrize("tau", [0.0, 25.0])
def test_md_600_steps_energy_curves_follow_oracle(tau, pot):
    """Longer trajectory (648 atoms, 600 steps, NVE and Berendsen-NVT) through the CUDA-graph MD loop.  The reference
    integrator carries no mass, so kinetic energy grows by orders of magnitude and the dynamics is chaotic: an initial
    1e-13 perturbation reaches 1e-8 Bohr after 400 steps (measured with the oracle).  Potential- and kinetic-energy
    curves must still follow the oracle's to 1e-7 of their scale over the whole run -- summation-order differences are
    amplified, anything systematic (a wrong force, a missed neighbour) would be off by many orders more."""
    import ctypes as C
    from pantea_b200 import _lib
    n_atoms, n_steps, dt, t0 = 648, 600, 0.25, 300.0
    pos, types, box = water_box(n_atoms)
    vel, mass = md_velocities(types), water_masses(types)
    _, _, _, sc_o = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, n_steps, t0, tau, KB)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms, cap=n_atoms - 1)  # rows can hold every other atom: the box densifies along the run
    t, m = cuda(types, torch.int32), cuda(mass)
    scal = torch.zeros((n_steps, 2), dtype=torch.float64, device="cuda")
    params = _lib.MDParams(dt, t0, tau, KB, 1, 1)
    reports = []
    for _ in range(6):  # the box densifies along the run: repeat until the capacities have been raised far enough
        p, v = cuda(pos), cuda(vel)
        ws.bind(p, t, box, dev.r_cutoff)
        _, _, f = ws.energy_forces(False, True)
        _lib.check(_lib.load().pantea_md_run(ws.handle, _lib.ptr(p), _lib.ptr(v), _lib.ptr(f), _lib.ptr(m), _lib.ptr(t), n_atoms,
                                             _lib.box_arg(box), n_steps, C.byref(params), _lib.ptr(scal), _lib.stream_ptr()))
        code = _lib.load().pantea_neighbor_status(ws.handle, None, _lib.stream_ptr())
        if code != _lib.PANTEA_ECAPACITY:
            _lib.check(code)
            break
        reports.append(_lib.load().pantea_last_error().decode())
    else:
        raise AssertionError(f"capacities did not settle: {reports}")
    s = scal.cpu().numpy()
    for col in (0, 1):
        scale = np.abs(sc_o[:, col]).max()
        assert np.abs(s[:, col] - sc_o[1:, col]).max() < 1e-7 * scale
    assert sc_o[-1, 1] > 50 * sc_o[0, 1] or tau > 0  # NVE: the kinetic energy really runs away (the test is not vacuous)


# ------------------------------------------------------------------------------------------ edge cases
def test_tiny_structures_and_atoms_without_neighbours(pot):
    """One atom (with and without a box), two atoms further apart than the cutoff, one water molecule: empty neighbour
    rows must give the bare network value scale(0) -> MLP, zero forces, and no out-of-range access."""
    cases = [(np.array([[1.0, 2.0, 3.0]]), np.array([2], dtype=np.int32), np.array([30.0, 30.0, 30.0])),
             (np.array([[1.0, 2.0, 3.0]]), np.array([1], dtype=np.int32), None),
             (np.array([[0.0, 0.0, 0.0], [20.0, 0.0, 0.0]]), np.array([1, 2], dtype=np.int32), None),
             (np.array([[0.0, 0.0, 0.0], [1.43, 1.1, 0.0], [-1.43, 1.1, 0.0]]), np.array([2, 1, 1], dtype=np.int32), None),
             (np.array([[5.0, 5.0, 5.0], [6.43, 6.1, 5.0], [3.57, 6.1, 5.0]]), np.array([2, 1, 1], dtype=np.int32),
              np.array([40.0, 41.0, 42.0]))]
    for pos, types, box in cases:
        e, ea, f = _energy_forces_gpu(pot, pos, types, box)
        eo, eao, fo = c_oracle.energy_forces(pot, pos, types, box)
        assert np.abs(ea - eao).max() < 1e-12 and np.abs(f - fo).max() < 1e-12 and abs(e - eo) < 1e-12
        e2, _, f2 = _energy_forces_gpu(pot, pos, types, box, torch.float32)
        assert abs(e2 - eo) < 1e-5 and np.abs(f2 - fo).max() < 1e-5


def test_atoms_of_an_element_unknown_to_the_potential(pot):
    """Types outside 1..n_elements are legal atoms no symmetry function refers to (reference: an element missing from
    the potential's neighbour lists contributes nothing): zero energy and force for them, and they must not disturb the
    descriptors of the others."""
    pos, types, box = water_box(192, seed=6)
    types = types.copy()
    types[::7] = 3
    e, ea, f = _energy_forces_gpu(pot, pos, types, box)
    eo, eao, fo = c_oracle.energy_forces(pot, pos, types, box)
    assert rel_err(ea, eao) < FP64_TOL and rel_err(f, fo) < FP64_TOL
    assert (ea[::7] == 0).all() and (f[::7] == 0).all()
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, len(pos))
    ws.bind(cuda(pos), cuda(types, torch.int32), box, dev.r_cutoff)
    for spec in pot:
        G, dG = ws.acsf(spec.atom_type - 1, len(spec.symfuncs), None, True, True)
        G_o, dG_o = c_oracle.acsf(spec, pos, types, box)
        _assert_descriptor_close(G.cpu().numpy(), dG.cpu().numpy(), G_o, dG_o, FP64_TOL)
    G, dG = ws.acsf(0, len(pot[0].symfuncs), torch.zeros(0, dtype=torch.int32, device="cuda"), True, True)
    assert G.shape == (0, len(pot[0].symfuncs)) and dG.shape == (0, len(pot[0].symfuncs), 3)   # empty centre list


def test_ragged_batch_of_structures(pot):
    """Dataset preprocessing launch over structures of very different sizes (1 ... 648 atoms, each with its own box)."""
    sizes = [3, 648, 12, 1 * 3, 192, 81]
    structs = [water_box(n, seed=40 + i) for i, n in enumerate(sizes)]
    pos = np.concatenate([s[0] for s in structs])
    types = np.concatenate([s[1] for s in structs])
    boxes = np.stack([s[2] for s in structs])
    ptr = np.concatenate([[0], np.cumsum([len(s[0]) for s in structs])]).astype(np.int32)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, len(pos), cap=647)
    ws.bind_batch(cuda(pos), cuda(types, torch.int32), cuda(ptr, torch.int32), cuda(boxes), dev.r_cutoff)
    for spec in pot:
        G, dG = ws.acsf(spec.atom_type - 1, len(spec.symfuncs), None, True, True)
        G, dG = G.cpu().numpy(), dG.cpu().numpy()
        for s, (p_s, t_s, b_s) in enumerate(structs):
            G_o, dG_o = c_oracle.acsf(spec, p_s, t_s, b_s)
            _assert_descriptor_close(G[ptr[s]:ptr[s + 1]], dG[ptr[s]:ptr[s + 1]], G_o, dG_o, FP64_TOL)



# ------------------------------------------------------------------------------------------ coincident neighbours
@pytest.mark.parametrize("n_atoms", [24, 12000])
def test_coincident_neighbours_are_excluded_exactly(n_atoms, pot):
    """Two neighbours j != k of a centre at bit-identical positions have r_jk = 0 and the reference drops the triplet
    (acsf.py:325).  A few atoms are moved onto the position of another atom of the same and of the other element (small
    system: all-pairs kernels; large system: cell list, fast path with the binning's coincidence flag)."""
    pos, types, box = water_box(n_atoms)
    pos = pos.copy()
    rng = np.random.default_rng(3)
    idx = rng.choice(n_atoms, size=6, replace=False)
    for a, b in zip(idx[:3], idx[3:]):
        pos[a] = pos[b]
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms)
    ws.bind(cuda(pos), cuda(types, torch.int32), box, dev.r_cutoff)
    _, e_atom, f = ws.energy_forces(False, True, True)
    _, ea_o, f_o = c_oracle.energy_forces(pot, pos, types, box)
    assert rel_err(e_atom.cpu().numpy(), ea_o) < 1e-10
    assert rel_err(f.cpu().numpy(), f_o) < 1e-10


# ------------------------------------------------------------------------------------------ fast path / mixed precision
@pytest.mark.parametrize("n_atoms", [12000, 99999])
def test_fast_path_and_mixed_precision_against_oracle(n_atoms, pot):
    """Specialised kernels (csrc/acsf2.cu) on the configurations they serve: double evaluation <= 1e-10 against the
    oracle with and without Gaussian screening, identical neighbour handling as the generic kernels; mixed mode (FP32
    symmetry functions on FP64 state, pantea_workspace_set_compute_precision) <= 1e-5 -- element-wise with the rms force
    as the absolute floor."""
    from pantea_b200 import _lib
    lib = _lib.load()
    pos, types, box = water_box(n_atoms)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, n_atoms)
    ws.bind(cuda(pos), cuda(types, torch.int32), box, dev.r_cutoff)
    m = min(n_atoms, 6000)
    _, ea_o, f_o = c_oracle.energy_forces(pot, pos, types, box, begin=0, end=m)
    f_o = torch.as_tensor(f_o[:m], device="cuda")
    rms = float(f_o.pow(2).mean().sqrt())

    def err_over(f, rtol):
        return float(((f[:m] - f_o).abs() / (rtol * (f_o.abs() + rms))).max())

    try:
        lib.pantea_set_fast_path(0)
        _, _, f_gen = ws.energy_forces(False, True)
        f_gen = f_gen.clone()
        lib.pantea_set_fast_path(1)
        old = lib.pantea_set_gauss_screen(0.0)
        _, e_atom, f_fast = ws.energy_forces(False, True, True)
        f_fast, e_fast = f_fast.clone(), e_atom.clone()
        lib.pantea_set_gauss_screen(40.0)
        _, e_atom, f_scr = ws.energy_forces(False, True, True)
        f_scr, e_scr = f_scr.clone(), e_atom.clone()
        ws.set_compute_precision(32)
        _, _, f_mix = ws.energy_forces(False, True)
        f_mix = f_mix.clone()
    finally:
        ws.set_compute_precision(64)
        lib.pantea_set_fast_path(1)
        lib.pantea_set_gauss_screen(40.0)
    assert err_over(f_gen, 1e-10) <= 1.0
    assert err_over(f_fast, 1e-10) <= 1.0
    assert err_over(f_scr, 1e-10) <= 1.0
    assert rel_err(e_fast[:m].cpu().numpy(), ea_o[:m]) < 1e-10 and rel_err(e_scr[:m].cpu().numpy(), ea_o[:m]) < 1e-10
    assert float((f_scr - f_fast).abs().max()) < 1e-13 * float(f_fast.abs().max())  # what screening drops is below 1e-14 of G
    assert err_over(f_mix, 1e-5) <= 1.0


@pytest.mark.parametrize("jitter", [0.0, 1e-9, 1e-4])
def test_fast_path_pairs_at_the_cutoff_distance(jitter, pot):
    """Pair lists of the fast path are built from TF32 distance tiles with an inclusive margin (csrc/acsf2.cu): a cubic
    lattice of spacing r_c / 3 puts thousands of neighbour pairs (j, k) exactly at r_jk = r_c, and 30 neighbours of every
    atom exactly at r_ij = r_c; a jitter moves them to either side within rounding (1e-9) or within the margin (1e-4).
    Forces must still match the oracle to 1e-10 (a dropped live pair or a kept dead one shows at 1e-4 and above)."""
    from pantea_b200 import _lib
    lib = _lib.load()
    n, a = 15, 4.0  # 3 375 atoms, box 60 Bohr >= 4 r_c, 122 lattice sites inside r_c = 12
    g = np.arange(n)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([ix, iy, iz], -1).reshape(-1, 3) * a
    rng = np.random.default_rng(5)
    pos = np.remainder(pos + jitter * rng.standard_normal(pos.shape) + 0.5 * a, n * a)
    t_h, t_o = pot[0].atom_type, pot[1].atom_type
    types = np.where((ix + iy + iz).reshape(-1) % 3 == 0, t_o, t_h).astype(np.int64)
    box = np.array([n * a] * 3)
    dev = device_potential_from_specs(pot)
    ws = _workspace(dev, len(pos), cap=160)
    ws.bind(cuda(pos), cuda(types, torch.int32), box, dev.r_cutoff)
    _, ea_o, f_o = c_oracle.energy_forces(pot, pos, types, box)
    f_o = torch.as_tensor(f_o, device="cuda")
    scale = float(f_o.abs().max()) + 1e-30
    try:
        lib.pantea_set_fast_path(1)
        _, e_atom, f = ws.energy_forces(False, True, True)
        f, e_atom = f.clone(), e_atom.clone()
        lib.pantea_set_fast_path(0)
        _, _, f_gen = ws.energy_forces(False, True)
    finally:
        lib.pantea_set_fast_path(1)
    # the jittered lattice has (near-)cancelling forces: the element-wise test uses the per-atom energy as well
    assert rel_err(e_atom.cpu().numpy(), ea_o) < 1e-10
    assert float((f - f_o).abs().max()) < 1e-10 * max(scale, 1e-3)
    assert float((f_gen - f_o).abs().max()) < 1e-10 * max(scale, 1e-3)
