"""Pins the two oracle tiers to every known answer the reference holds for the hot path
(tests/golden/reference_vectors.json cites the reference file:line of each vector), then
checks the analytic C oracle against the faithful dense restatement on other inputs.
CPU only -- this is the checker being checked, not the product."""
import json

import numpy as np
import pytest
import torch

from oracle import c_oracle, dense_oracle as D
from oracle.spec import (ElementSpec, SymFuncSpec, load_potential, md_velocities, read_runner, rune_width_potential,
                         type_map, water_box, water_masses, FROM_ATOMIC_MASS, KB, MASS_U)


@pytest.fixture(scope="module")
def vec(golden_dir):
    return json.loads((golden_dir / "reference_vectors.json").read_text())


@pytest.fixture(scope="module")
def h2o(golden_dir):
    f = read_runner(golden_dir / "h2o.data")[0]
    f["positions"] = np.remainder(f["positions"], f["box"])  # Structure.__post_init__ wrap (structure.py:80-81)
    return f


@pytest.fixture(scope="module")
def pot(golden_dir):
    return load_potential(golden_dir / "h2o.json")


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float64)


def test_ne2_g2_no_pbc(vec):
    v = vec["ne2_g2"]
    pos, types = np.asarray(v["positions"]), np.ones(2, dtype=np.int32)
    spec = ElementSpec(1, [SymFuncSpec(2, v["cutoff"][0], v["cutoff"][1], 1, 0, v["eta"], rs) for rs in v["r_shifts"]])
    dense = D.acsf_values(D.symfuncs_from_spec(spec), _t(pos), torch.as_tensor(types), None, torch.arange(2))
    np.testing.assert_allclose(dense.numpy(), np.tile(v["expected_row"], (2, 1)), rtol=1e-8)
    G, _ = c_oracle.acsf(spec, pos, types, None)
    np.testing.assert_allclose(G, np.tile(v["expected_row"], (2, 1)), rtol=1e-8)


def test_h2o_pbc_g2_g3(vec, h2o):
    v = vec["h2o_pbc_g2_g3"]
    tm = type_map(h2o["elements"])
    ct, rc = v["cutoff"]
    spec = ElementSpec(tm["O"], [
        SymFuncSpec(2, ct, rc, tm[v["radial"]["neighbor"]], 0, v["radial"]["eta"], v["radial"]["r_shift"]),
        SymFuncSpec(3, ct, rc, tm["H"], tm["H"], v["angular"]["eta"], 0.0, v["angular"]["lambda0"], v["angular"]["zeta"]),
    ])
    centres = np.nonzero(h2o["types"] == tm["O"])[0]
    dense = D.acsf_values(D.symfuncs_from_spec(spec), _t(h2o["positions"]), torch.as_tensor(h2o["types"]),
                          _t(h2o["box"]), torch.as_tensor(centres))
    assert tuple(dense.shape) == tuple(v["shape"])
    np.testing.assert_allclose(dense[0].numpy(), v["expected_atom0"], rtol=0, atol=6e-11)  # 10 printed decimals
    G, _ = c_oracle.acsf(spec, h2o["positions"], h2o["types"], h2o["box"], centres)
    np.testing.assert_allclose(G[0], v["expected_atom0"], rtol=0, atol=6e-11)
    np.testing.assert_allclose(G, dense.numpy(), rtol=1e-12, atol=1e-18)


def test_notebook_distances_and_neighbors(vec, h2o):
    d = c_oracle.distances(h2o["positions"], h2o["box"])
    np.testing.assert_allclose(d[0, :5], vec["notebook_distances"]["expected"], atol=5e-9)
    r, _ = D.distances_with_aux(_t(h2o["positions"]), _t(h2o["positions"]), _t(h2o["box"]))
    assert np.array_equal(r.numpy(), d)  # the two tiers agree bit for bit on distances
    row_ptr, col = c_oracle.neighbors(h2o["positions"], h2o["types"], h2o["box"], vec["notebook_neighbors"]["r_cutoff"])
    assert row_ptr[1] - row_ptr[0] == vec["notebook_neighbors"]["expected_count_atom0"]
    mask = D.cutoff_mask(r, vec["notebook_neighbors"]["r_cutoff"]).numpy()
    for i in range(len(d)):
        assert np.array_equal(np.nonzero(mask[i])[0], col[row_ptr[i]:row_ptr[i + 1]])


def _notebook_spec(v, tm):
    ct, rc = v["cutoff"]
    sfs = [SymFuncSpec(2, ct, rc, tm[r["neighbor"]], 0, r["eta"], r["r_shift"]) for r in v["radial"]]
    sfs += [SymFuncSpec(a["kind"], ct, rc, tm[a["neighbors"][0]], tm[a["neighbors"][1]], a["eta"], 0.0, a["lambda0"],
                        a["zeta"]) for a in v["angular"]]
    return ElementSpec(tm["O"], sfs)


def test_notebook_acsf_values_and_grad(vec, h2o):
    v = vec["notebook_acsf"]
    tm = type_map(h2o["elements"])
    spec = _notebook_spec(v, tm)
    centres = np.nonzero(h2o["types"] == tm["O"])[0]
    args = (D.symfuncs_from_spec(spec), _t(h2o["positions"]), torch.as_tensor(h2o["types"]), _t(h2o["box"]))
    dense = D.acsf_values(*args, torch.as_tensor(centres)).numpy()
    np.testing.assert_allclose(dense, v["expected_values"], rtol=2e-8)
    # ACSF.grad(structure)[:1]: all atoms are centres, first row is atom 0 (acsf.py:103-104)
    gd = D.acsf_grad(*args, torch.arange(len(h2o["types"]))).numpy()
    np.testing.assert_allclose(gd[0], v["expected_grad_atom0"], atol=6e-9)
    G, dG = c_oracle.acsf(spec, h2o["positions"], h2o["types"], h2o["box"], None)
    np.testing.assert_allclose(G[centres], dense, rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(dG, gd, rtol=1e-10, atol=1e-16)


def test_nnp_energy_forces_fp32_goldens(vec, h2o, pot):
    v = vec["nnp_fp32"]
    e, e_atom, f = D.energy_and_forces(D.models_from_specs(pot), _t(h2o["positions"]), torch.as_tensor(h2o["types"]),
                                       _t(h2o["box"]))
    # goldens are float32 results compared with jnp.allclose defaults (rtol 1e-5, atol 1e-8)
    # E is a sum of 12 atomic energies of magnitude ~0.2 that cancel to -0.0072: the float32 golden carries an
    # accumulation error of ~12 * 0.2 * 2^-24 ~ 1.5e-7 (a float32 run of this oracle gives -0.00721341)
    np.testing.assert_allclose(float(e), v["energy"], rtol=0, atol=3e-7)
    np.testing.assert_allclose(f.numpy(), v["forces"], rtol=1e-5, atol=1e-7)
    ec, eac, fc = c_oracle.energy_forces(pot, h2o["positions"], h2o["types"], h2o["box"])
    np.testing.assert_allclose(ec, float(e), rtol=1e-12)
    np.testing.assert_allclose(eac, e_atom.numpy(), rtol=1e-11, atol=1e-16)
    np.testing.assert_allclose(fc, f.numpy(), rtol=1e-10, atol=1e-15)
    # the reference force is the central-role partial derivative: it does not sum to zero
    assert np.abs(f.numpy().sum(axis=0)).max() > 1e-2


def test_masses(vec):
    for el in ("H", "O", "Ne"):
        assert MASS_U[el] * FROM_ATOMIC_MASS == pytest.approx(vec["masses"][el], rel=1e-14)


@pytest.mark.parametrize("cutoff", ["hard", "cos", "tanhu", "tanh", "exp", "poly1", "poly2"])
def test_c_vs_dense_all_cutoffs_g9_zeta(cutoff):
    """Unpinned-by-reference features: every cutoff type, G1, G9, zeta > 1, lambda = -1,
    mixed cutoff radii -- analytic C gradients against dense autograd."""
    pos, types, box = water_box(48, seed=11)
    rc = 0.9 if cutoff in ("poly1", "poly2") else 9.0  # poly* act on raw r: only sensible for r < 1
    if cutoff in ("poly1", "poly2"):
        pos, box = pos / 12.0, box / 12.0
    spec = ElementSpec(2, [
        SymFuncSpec(1, cutoff, rc, 1), SymFuncSpec(2, cutoff, rc * 0.8, 2, 0, 0.05, 0.5),
        SymFuncSpec(3, cutoff, rc, 1, 1, 0.01, 0.0, -1.0, 4.0), SymFuncSpec(3, cutoff, rc * 0.9, 1, 2, 0.02, 0.0, 1.0, 2.0),
        SymFuncSpec(9, cutoff, rc, 2, 2, 0.005, 0.0, 1.0, 1.0), SymFuncSpec(9, cutoff, rc, 2, 1, 0.03, 0.0, -1.0, 3.0),
        SymFuncSpec(3, cutoff, rc, 2, 2, 0.03, 0.0, 1.0, 2.5),
    ])
    args = (D.symfuncs_from_spec(spec), _t(pos), torch.as_tensor(types), _t(box))
    centres = torch.arange(len(types))
    gd, dgd = D.acsf_values(*args, centres).numpy(), D.acsf_grad(*args, centres).numpy()
    G, dG = c_oracle.acsf(spec, pos, types, box)
    scale = np.abs(gd).max(axis=0) + 1e-300
    assert np.abs((G - gd) / scale).max() < 1e-12
    gscale = np.abs(dgd).max(axis=(0, 2))[None, :, None] + 1e-300
    assert np.abs((dG - dgd) / gscale).max() < 1e-11


@pytest.mark.parametrize("scale_type", ["center", "scale", "scale_center", "scale_center_sigma"])
def test_c_vs_dense_energy_forces_water96(scale_type, pot):
    pos, types, box = water_box(96, seed=5)
    specs = [ElementSpec(s.atom_type, s.symfuncs, scale_type, s.scaler, 0.0, 1.0, s.layers) for s in pot]
    e, e_atom, f = D.energy_and_forces(D.models_from_specs(specs), _t(pos), torch.as_tensor(types), _t(box))
    ec, eac, fc = c_oracle.energy_forces(specs, pos, types, box)
    np.testing.assert_allclose(eac, e_atom.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(fc, f.numpy(), rtol=1e-9, atol=1e-13)
    assert abs(ec - float(e)) < 1e-12 * max(1.0, abs(float(e)))


def test_c_vs_dense_wide_potential_and_activations():
    pos, types, box = water_box(48, seed=3)
    specs = rune_width_potential()
    acts = ["logistic", "softplus", "gaussian", "cos", "exp", "harmonic", "relu", "tanh"]
    for i, s in enumerate(specs):
        s.layers = [(k, b, acts[(2 * i + l) % len(acts)] if l < 2 else "identity") for l, (k, b, _) in enumerate(s.layers)]
    e, e_atom, f = D.energy_and_forces(D.models_from_specs(specs), _t(pos), torch.as_tensor(types), _t(box))
    ec, eac, fc = c_oracle.energy_forces(specs, pos, types, box)
    np.testing.assert_allclose(eac, e_atom.numpy(), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(fc, f.numpy(), rtol=1e-9, atol=1e-12)


def test_no_box_structure(pot):
    pos, types, _ = water_box(24, seed=9)
    e, e_atom, f = D.energy_and_forces(D.models_from_specs(pot), _t(pos), torch.as_tensor(types), None)
    ec, eac, fc = c_oracle.energy_forces(pot, pos, types, None)
    np.testing.assert_allclose(eac, e_atom.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(fc, f.numpy(), rtol=1e-9, atol=1e-13)


def test_md_c_oracle_matches_dense_steps(pot):
    """Velocity Verlet without mass + Berendsen (molecular_dynamics.py:16-77, thermostat.py:12-22)."""
    pos, types, box = water_box(24, seed=2)
    vel = md_velocities(types)
    mass = water_masses(types)
    dt, tau, t0, kb = 0.25, 25.0, 300.0, 3.166811563e-6
    p1, v1, f1, sc = c_oracle.md_run(pot, pos, vel, mass, types, box, dt, 3, t0, tau, kb)
    models = D.models_from_specs(pot)
    x, v, m = _t(pos), _t(vel), _t(mass)[:, None]
    tb, ty = _t(box), torch.as_tensor(types)
    _, _, f = D.energy_and_forces(models, x, ty, tb)
    for _ in range(3):
        x = D.wrap_into_box(D.verlet_positions(x, v, f, dt), tb)
        _, _, fn = D.energy_and_forces(models, x, ty, tb)
        v = D.verlet_velocities(v, f, fn, dt)
        f = fn
        v = D.berendsen_scale(v, dt, tau, D.temperature(v, m, kb), t0)
    np.testing.assert_allclose(p1, x.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(v1, v.numpy(), rtol=1e-10, atol=1e-14)
    assert sc.shape == (4, 3) and np.isfinite(sc).all()


def test_c_oracle_cell_list_is_bitwise_identical_to_all_pairs(pot):
    pos, types, box = water_box(3000)
    c_oracle.set_use_cells(False)
    try:
        e0, ea0, f0 = c_oracle.energy_forces(pot, pos, types, box)
    finally:
        c_oracle.set_use_cells(True)
    e1, ea1, f1 = c_oracle.energy_forces(pot, pos, types, box)
    assert np.array_equal(ea0, ea1) and np.array_equal(f0, f1)


def test_full_force_oracle_is_the_total_derivative(h2o, pot):
    """The oracle of the PANTEA_FORCE_FULL extension (autograd through both roles of the dense restatement): zero net
    force, central differences of the total energy, and the contrast SURVEY.md App. C records for the reference fixture
    (the full force differs from the reference's central-role force by up to 0.145)."""
    from oracle import dense_oracle
    models = dense_oracle.models_from_specs(pot)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    pos, types, box = T(h2o["positions"]), torch.from_numpy(h2o["types"]), T(h2o["box"])
    e, f_full = dense_oracle.energy_and_full_forces(models, pos, types, box)
    e_c, _, f_central = dense_oracle.energy_and_forces(models, pos, types, box)
    assert abs(float(e) - float(e_c)) < 1e-15
    assert f_full.sum(0).abs().max() < 1e-15
    assert abs(float((f_full - f_central).abs().max()) - 0.145) < 1e-3
    h = 1e-5
    for i, c in ((0, 0), (5, 1), (11, 2)):
        pp, pm = pos.clone(), pos.clone()
        pp[i, c] += h
        pm[i, c] -= h
        ep = dense_oracle.energy_and_forces(models, pp, types, box, want_forces=False)[0]
        em = dense_oracle.energy_and_forces(models, pm, types, box, want_forces=False)[0]
        assert abs(float(-(ep - em) / (2 * h)) - float(f_full[i, c])) < 1e-8


def test_mass_scaled_full_force_oracle_md_conserves_energy(pot):
    """Oracle of the `mass_scaled` + full-force MD extension: ordinary velocity Verlet, so E_pot + E_kin is conserved to
    O(dt^2) while both change by several Ha; the C oracle's mass-scaled loop (reference force) shares the integrator."""
    from oracle import dense_oracle
    pos, types, box = water_box(24)
    vel, mass = md_velocities(types), water_masses(types)
    models = dense_oracle.models_from_specs(pot)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    drift = {}
    for dt in (2.5, 5.0):
        _, _, _, sc = dense_oracle.md_run_full(models, T(pos), T(vel), T(mass), torch.from_numpy(types), T(box), dt,
                                               int(100 / dt))
        e_tot = sc.sum(1)
        drift[dt] = float((e_tot - e_tot[0]).abs().max())
        swing = float(sc[:, 1].max() - sc[:, 1].min())
        assert swing > 0.1
    assert drift[5.0] < 5e-3 * swing and drift[2.5] < 0.4 * drift[5.0]   # second-order integrator
    # integrator arithmetic of the C oracle's mass-scaled loop: one step by hand
    po, vo, fo, _ = c_oracle.md_run(pot, pos, vel, mass, types, box, 5.0, 1, mass_scaled=True)
    _, _, f0 = c_oracle.energy_forces(pot, pos, types, box)
    x1 = np.remainder(pos + vel * 5.0 + 0.5 * (f0 / mass[:, None]) * 25.0, box)
    np.testing.assert_allclose(po, x1, rtol=0, atol=1e-12)
    _, _, f1 = c_oracle.energy_forces(pot, x1, types, box)
    np.testing.assert_allclose(vo, vel + 0.5 * (f0 / mass[:, None] + f1 / mass[:, None]) * 5.0, rtol=1e-12, atol=1e-18)


def test_extension_vectors_fixture_is_reproduced_by_the_oracle(golden_dir, h2o, pot):
    """tests/golden/extension_vectors.json (made by tests/golden/make_extension_vectors.py) holds the known answers the GPU
    tests of the full-force / mass-scaled extensions compare with; the oracle must still reproduce it."""
    from oracle import dense_oracle
    fx = json.loads((golden_dir / "extension_vectors.json").read_text())
    models = dense_oracle.models_from_specs(pot)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    pos, types, box = T(h2o["positions"]), torch.from_numpy(h2o["types"]), T(h2o["box"])
    e, f = dense_oracle.energy_and_full_forces(models, pos, types, box)
    assert abs(float(e) - fx["energy"]) < 1e-15
    np.testing.assert_allclose(f.numpy(), np.asarray(fx["full_forces"]), rtol=0, atol=1e-14)
    assert abs(fx["max_abs_difference_to_reference_force"] - 0.145) < 1e-3
    md = fx["md_full_mass_scaled"]
    x, v, _, sc = dense_oracle.md_run_full(models, pos, T(md["velocities0"]), T(md["masses"]), types, box, md["dt"],
                                           md["n_steps"])
    np.testing.assert_allclose(x.numpy(), np.asarray(md["positions"]), rtol=0, atol=1e-12)
    np.testing.assert_allclose(v.numpy(), np.asarray(md["velocities"]), rtol=1e-10, atol=1e-16)
    e_tot = np.asarray(md["e_pot_e_kin"]).sum(1)
    assert np.abs(e_tot - e_tot[0]).max() < 1e-4 and np.ptp(np.asarray(md["e_pot_e_kin"])[:, 1]) > 0.03


@pytest.mark.parametrize("case", ["h2o_box", "h2o_open", "wide"])
def test_c_oracle_full_forces_match_dense_autograd(case, pot):
    """Tier 2 of the full-force extension oracle (analytic, any size) against tier 1 (autograd through both roles)."""
    from oracle import dense_oracle
    specs, n, use_box = {"h2o_box": (pot, 96, True), "h2o_open": (pot, 81, False), "wide": (rune_width_potential(), 48, True)}[case]
    pos, types, box = water_box(n, seed=4)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    e, _, f = c_oracle.energy_full_forces(specs, pos, types, box if use_box else None)
    eo, fo = dense_oracle.energy_and_full_forces(dense_oracle.models_from_specs(specs), T(pos), torch.from_numpy(types),
                                                 T(box) if use_box else None)
    assert abs(e - float(eo)) < 1e-12 * max(1.0, abs(float(eo)))
    assert np.abs(f - fo.numpy()).max() < 1e-12 * np.abs(fo.numpy()).max()
    assert np.abs(f.sum(0)).max() < 1e-13


def test_md10k_fixture_head_is_reproduced_by_the_oracle(golden_dir, pot):
    """tests/golden/md10k_nve_192.json (10 000-step NVE curve of the C oracle, ~1 min to regenerate): its first 500 steps
    are recomputed here; the rest is trusted to the committed generating script."""
    fx = json.loads((golden_dir / "md10k_nve_192.json").read_text())
    pos, types, box = water_box(fx["n_atoms"])
    vel, mass = md_velocities(types), water_masses(types)
    _, _, _, sc = c_oracle.md_run(pot, pos, vel, mass, types, box, fx["dt"], 500, 300.0, 0.0, KB)
    ref = np.asarray(fx["e_pot_e_kin"])
    assert fx["steps"][:3] == [0, 250, 500]
    np.testing.assert_allclose(sc[[0, 250, 500], :2], ref[:3], rtol=1e-9)
    assert fx["spread"]["max_abs_log_ratio_e_kin"] < 0.2 and fx["spread"]["max_rel_dev_first_500_steps"] < 1e-8
    assert ref[-1, 1] > 1e6 * ref[0, 1]
