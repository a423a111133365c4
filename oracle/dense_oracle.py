"""TEST INFRASTRUCTURE ONLY -- faithful dense CPU restatement of the reference HDNNP hot path.

This module is the parity *oracle* (tier 1).  It restates, in float64 torch-on-CPU, exactly the
algorithm of the reference JAX code -- dense N x N minimum-image distances, masked O(N^3)
angular sums, automatic differentiation with respect to the *central* positions only -- so
that the analytic-gradient oracle (`oracle/hdnnp_oracle.c`) and the CUDA kernels can be
checked against the reference's own semantics.  Nothing under `pantea_b200/` imports it; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`/`--impl reference` legs may.

Pinned (tests/test_oracle_golden.py) against every golden vector the reference holds for this
path: `tests/test_acsf.py:123-173`, `tests/test_nnp.py:44-81`, notebook cells 17/27/29 of
`examples/getting_started.ipynb`.  The real reference cannot be imported here (jax/flax/ase are
not installed and there is no network), so those vectors are the anchor.

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

TANH_PRE = ((math.e + 1 / math.e) / (math.e - 1 / math.e)) ** 3  # pantea/descriptors/acsf/cutoff.py:82

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- geometry
def apply_pbc(dx: Tensor, box: Tensor) -> Tensor:
    """Single-shift minimum image -- pantea/atoms/box.py:112-117."""
    dx = torch.where(dx > 0.5 * box, dx - box, dx)
    dx = torch.where(dx < -0.5 * box, dx + box, dx)
    return dx


def wrap_into_box(positions: Tensor, box: Tensor) -> Tensor:
    """Floored remainder -- pantea/atoms/box.py:123-126."""
    return torch.remainder(positions, box)


def distances_with_aux(pos_c: Tensor, pos_all: Tensor, box: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """r_ij and d_ij = r_i - r_j (+PBC) with the zero-vector guard -- pantea/atoms/distance.py:63-78.

    pos_c [n,3], pos_all [N,3] -> (r [n,N], d [n,N,3]).  The accumulation order of the norm is
    (dx^2 + dy^2) + dz^2, which is what the analytic oracle and the CUDA kernels use too.
    """
    d = pos_c[:, None, :] - pos_all[None, :, :]
    if box is not None:
        d = apply_pbc(d, box)
    is_zero = (d == 0.0).all(dim=-1, keepdim=True)
    dm = torch.where(is_zero, torch.ones_like(d), d)
    r = torch.sqrt((dm[..., 0] * dm[..., 0] + dm[..., 1] * dm[..., 1]) + dm[..., 2] * dm[..., 2])
    r = torch.where(is_zero[..., 0], torch.zeros_like(r), r)
    return r, d


def cutoff_mask(r: Tensor, rc: float) -> Tensor:
    """Neighbour predicate (r <= rc) & (r > 0) -- pantea/atoms/neighbor.py:102-107."""
    return (r <= rc) & (r > 0.0)


# ----------------------------------------------------------------------------- cutoff functions
def cutoff_function(kind: str, r: Tensor, rc: float) -> Tensor:
    """fc(r) = where(r < rc, f(r), 0) -- pantea/descriptors/acsf/cutoff.py:67-110."""
    if kind == "hard":
        f = torch.ones_like(r)
    elif kind == "tanhu":
        f = torch.tanh(1.0 - r / rc) ** 3
    elif kind == "tanh":
        f = TANH_PRE * torch.tanh(1.0 - r / rc) ** 3
    elif kind == "cos":
        f = 0.5 * (torch.cos(math.pi * r / rc) + 1.0)
    elif kind == "exp":
        # guard the unused branch (r >= rc) against 1/0 -> the where() below discards it
        u = torch.where(r < rc, (r / rc) ** 2, torch.zeros_like(r))
        f = torch.exp(1.0 - 1.0 / (1.0 - u))
    elif kind == "poly1":
        f = (2.0 * r - 3.0) * r**2 + 1.0  # raw r, not r/rc (cutoff.py:105-106)
    elif kind == "poly2":
        f = ((15.0 - 6.0 * r) * r - 10) * r**3 + 1.0  # raw r (cutoff.py:109-110)
    else:
        raise KeyError(kind)
    return torch.where(r < rc, f, torch.zeros_like(r))


# ----------------------------------------------------------------------------- symmetry functions
class SymFunc:
    """One symmetry function with its neighbour elements.

    kind: 1 (G1), 2 (G2), 3 (G3), 9 (G9); type_j/type_k are atom types of the structure
    (1-based); r_shift is carried but ignored by G3/G9 (pantea/descriptors/acsf/angular.py:58-65,
    100-107).
    """

    def __init__(self, kind: int, cutoff_type: str, r_cutoff: float, type_j: int, type_k: int = 0,
                 eta: float = 0.0, r_shift: float = 0.0, lambda0: float = 0.0, zeta: float = 0.0) -> None:
        self.kind, self.cutoff_type, self.r_cutoff = kind, cutoff_type, float(r_cutoff)
        self.type_j, self.type_k = int(type_j), int(type_k)
        self.eta, self.r_shift, self.lambda0, self.zeta = float(eta), float(r_shift), float(lambda0), float(zeta)

    @property
    def is_radial(self) -> bool:
        return self.kind in (1, 2)

    def fc(self, r: Tensor) -> Tensor:
        return cutoff_function(self.cutoff_type, r, self.r_cutoff)

    def radial(self, r: Tensor) -> Tensor:
        """G1/G2 -- pantea/descriptors/acsf/radial.py:39-61."""
        if self.kind == 1:
            return self.fc(r)
        return torch.exp(-self.eta * (r - self.r_shift) ** 2) * self.fc(r)

    def angular(self, rij: Tensor, rik: Tensor, rjk: Tensor, cost: Tensor) -> Tensor:
        """G3/G9 -- pantea/descriptors/acsf/angular.py:51-107."""
        pre = 2.0 ** (1.0 - self.zeta) * torch.pow(1 + self.lambda0 * cost, self.zeta)
        if self.kind == 3:
            return pre * torch.exp(-self.eta * (rij**2 + rik**2 + rjk**2)) * self.fc(rij) * self.fc(rik) * self.fc(rjk)
        return pre * torch.exp(-self.eta * (rij**2 + rik**2)) * self.fc(rij) * self.fc(rik)


def acsf_descriptor(symfuncs: Sequence[SymFunc], pos_c: Tensor, pos_all: Tensor, types: Tensor,
                    box: Optional[Tensor]) -> Tensor:
    """Descriptor rows for the centres `pos_c` -- pantea/descriptors/acsf/acsf.py:163-202, 231-330.

    Radial functions first, then angular, in the order given.  Differentiable w.r.t. `pos_c`
    only when `pos_all` is a constant (the reference's jacfwd/grad argnums=1 semantics).
    """
    r_i, d_i = distances_with_aux(pos_c, pos_all, box)  # [n,N], [n,N,3]
    cols: List[Tensor] = []
    for sf in symfuncs:
        mask_rc = cutoff_mask(r_i, sf.r_cutoff)
        if sf.is_radial:
            m = mask_rc & (types == sf.type_j)[None, :]  # acsf.py:238-241
            cols.append(torch.where(m, sf.radial(r_i), torch.zeros_like(r_i)).sum(dim=1))  # acsf.py:242-246
            continue
        mask_ij = mask_rc & (types == sf.type_j)[None, :]  # acsf.py:263-265
        mask_ik = mask_rc & (types == sf.type_k)[None, :]  # acsf.py:267-269
        # acsf.py:307-315 -- cos(theta) with zero guards; masked k -> cost = 1
        operand = r_i[:, :, None] * r_i[:, None, :]  # [n,j,k]
        is_zero = operand == 0.0
        true_op = torch.where(is_zero, torch.ones_like(operand), operand)
        inner = (d_i[:, :, None, :] * d_i[:, None, :, :]).sum(dim=-1)
        cost = torch.where(mask_ik[:, None, :], inner / true_op, torch.ones_like(operand))
        cost = torch.where(is_zero, torch.zeros_like(cost), cost)
        # acsf.py:316-320 -- r_jk = || pbc(d_ij - d_ik) || with the same zero-vector guard
        d_jk = d_i[:, :, None, :] - d_i[:, None, :, :]
        if box is not None:
            d_jk = apply_pbc(d_jk, box)
        zero_jk = (d_jk == 0.0).all(dim=-1)
        djm = torch.where(zero_jk[..., None], torch.ones_like(d_jk), d_jk)
        r_jk = torch.sqrt((djm[..., 0] * djm[..., 0] + djm[..., 1] * djm[..., 1]) + djm[..., 2] * djm[..., 2])
        r_jk = torch.where(zero_jk, torch.zeros_like(r_jk), r_jk)
        r_jk = torch.where(mask_ik[:, None, :], r_jk, torch.zeros_like(r_jk))
        # acsf.py:321-329 -- sum over k (mask_ik & r_jk > 0), then over j (mask_ij)
        val = sf.angular(r_i[:, :, None], r_i[:, None, :], r_jk, cost)
        keep = mask_ij[:, :, None] & mask_ik[:, None, :] & (r_jk > 0.0)
        total = torch.where(keep, val, torch.zeros_like(val)).sum(dim=(1, 2))
        if sf.type_j == sf.type_k:  # acsf.py:285-290
            total = total * 0.5
        cols.append(total)
    if not cols:
        return pos_c.new_zeros((pos_c.shape[0], 0))
    return torch.stack(cols, dim=1)


def acsf_values(symfuncs: Sequence[SymFunc], positions: Tensor, types: Tensor, box: Optional[Tensor],
                centres: Tensor, chunk: int = 16) -> Tensor:
    """`ACSF.__call__` -- pantea/descriptors/acsf/acsf.py:47-86 (centres given explicitly)."""
    out = []
    with torch.no_grad():
        for s in range(0, len(centres), chunk):
            out.append(acsf_descriptor(symfuncs, positions[centres[s:s + chunk]], positions, types, box))
    return torch.cat(out) if out else positions.new_zeros((0, len(symfuncs)))


def acsf_grad(symfuncs: Sequence[SymFunc], positions: Tensor, types: Tensor, box: Optional[Tensor],
              centres: Tensor, chunk: int = 8) -> Tensor:
    """`ACSF.grad` = dG_i/dr_i, central role only -- pantea/descriptors/acsf/acsf.py:88-120, 215-228."""
    n_sf = len(symfuncs)
    grads = positions.new_zeros((len(centres), n_sf, 3))
    const = positions.detach()
    for s in range(0, len(centres), chunk):
        idx = centres[s:s + chunk]
        pc = const[idx].clone().requires_grad_(True)
        g = acsf_descriptor(symfuncs, pc, const, types, box)  # [c, n_sf]
        for k in range(n_sf):
            (gk,) = torch.autograd.grad(g[:, k].sum(), pc, retain_graph=k + 1 < n_sf)
            grads[s:s + len(idx), k, :] = gk
    return grads


# ----------------------------------------------------------------------------- scaler + MLP
def scale(kind: str, params: Dict[str, Tensor], x: Tensor, smin: float = 0.0, smax: float = 1.0) -> Tensor:
    """Scaler transforms -- pantea/descriptors/scaler.py:206-246 (sign of `scale_center_sigma` as written)."""
    if kind == "center":
        return x - params["mean"]
    if kind == "scale":
        return smin + (smax - smin) * (x - params["minval"]) / (params["maxval"] - params["minval"])
    if kind == "scale_center":
        return smin + (smax - smin) * (x - params["mean"]) / (params["maxval"] - params["minval"])
    if kind == "scale_center_sigma":
        return smin + (smin - smax) * (x - params["mean"]) / params["sigma"]
    raise KeyError(kind)


def activation(name: str, x: Tensor) -> Tensor:
    """Activation table -- pantea/models/nn/activation.py:7-60 (`exp` is exp(-x), sic)."""
    if name == "identity":
        return x
    if name == "tanh":
        return torch.tanh(x)
    if name == "logistic":
        return 1.0 / (1.0 + torch.exp(-x))
    if name == "softplus":
        return torch.nn.functional.softplus(x)
    if name == "relu":
        return torch.relu(x)
    if name == "gaussian":
        return torch.exp(-0.5 * x**2)
    if name == "cos":
        return torch.cos(x)
    if name == "exp":
        return torch.exp(-x)
    if name == "harmonic":
        return x * x
    raise KeyError(name)


def mlp(layers: Sequence[Tuple[Tensor, Tensor, str]], x: Tensor) -> Tensor:
    """Dense stack x -> act(x W + b) -- pantea/models/nn/model.py:40-58; kernels are [in,out]."""
    for w, b, act in layers:
        x = activation(act, x @ w + b)
    return x


class ElementModel:
    """Descriptor + scaler + MLP of one element (pantea/potentials/nnp/atomic_potential.py:12-22)."""

    def __init__(self, atom_type: int, symfuncs: Sequence[SymFunc], scale_type: str, scaler_params: Dict[str, Tensor],
                 layers: Sequence[Tuple[Tensor, Tensor, str]], smin: float = 0.0, smax: float = 1.0) -> None:
        self.atom_type, self.symfuncs = atom_type, list(symfuncs)
        self.scale_type, self.scaler_params, self.layers = scale_type, scaler_params, list(layers)
        self.smin, self.smax = smin, smax

    def energies(self, pos_c: Tensor, pos_all: Tensor, types: Tensor, box: Optional[Tensor]) -> Tensor:
        """Per-atom energies -- pantea/potentials/nnp/energy.py:23-39."""
        x = acsf_descriptor(self.symfuncs, pos_c, pos_all, types, box)
        x = scale(self.scale_type, self.scaler_params, x, self.smin, self.smax)
        return mlp(self.layers, x)[:, 0]


def energy_and_forces(models: Sequence[ElementModel], positions: Tensor, types: Tensor, box: Optional[Tensor],
                      chunk: int = 8, want_forces: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """Total energy, per-atom energies and the reference's *central-role* forces.

    E = sum_el sum_{i in el} E_i (pantea/potentials/nnp/energy.py:45-63); forces are
    -dE/d(central positions) with the neighbour copy of the positions held constant
    (pantea/potentials/nnp/force.py:16-43, potential.py:83-102).
    """
    const = positions.detach()
    e_atom = const.new_zeros(len(const))
    forces = const.new_zeros((len(const), 3))
    for model in models:
        idx_all = torch.nonzero(types == model.atom_type, as_tuple=True)[0]
        for s in range(0, len(idx_all), chunk):
            idx = idx_all[s:s + chunk]
            pc = const[idx].clone().requires_grad_(want_forces)
            e = model.energies(pc, const, types, box)
            e_atom[idx] = e.detach()
            if want_forces:
                (g,) = torch.autograd.grad(e.sum(), pc)
                forces[idx] = -g
    return e_atom.sum(), e_atom, forces


def energy_and_full_forces(models: Sequence[ElementModel], positions: Tensor, types: Tensor, box: Optional[Tensor],
                           chunk: int = 16) -> Tuple[Tensor, Tensor]:
    """Total energy and F = -dE/dr with respect to ALL occurrences of the positions (centre and neighbour roles).

    NOT a reference mode: the reference differentiates the central copy only (force.py:16-43, see
    `energy_and_forces`).  This is the oracle of the library's PANTEA_FORCE_FULL extension:
    the same energy expression as above, plain reverse-mode autograd through both roles.
    """
    pos = positions.detach().clone().requires_grad_(True)
    total = pos.new_zeros(())
    for model in models:
        idx_all = torch.nonzero(types == model.atom_type, as_tuple=True)[0]
        for s in range(0, len(idx_all), chunk):
            idx = idx_all[s:s + chunk]
            total = total + model.energies(pos[idx], pos, types, box).sum()
    (g,) = torch.autograd.grad(total, pos)
    return total.detach(), -g


def md_run_full(models: Sequence[ElementModel], positions: Tensor, velocities: Tensor, masses: Tensor, types: Tensor,
                box: Optional[Tensor], dt: float, n_steps: int, mass_scaled: bool = True):
    """Velocity Verlet with the FULL force (and, by default, accelerations F/m): the oracle of the library's
    `pantea_md_params.{force_mode = FULL, mass_scaled}` extensions -- NOT a reference mode (the reference integrates
    the central-role force without mass, molecular_dynamics.py:16-30).  Same update order as `verlet_positions` /
    `verlet_velocities`.  Returns (positions, velocities, forces, scalars [n_steps + 1, 2] = (E_pot, E_kin))."""
    x, v = positions.clone(), velocities.clone()
    m = masses.reshape(-1, 1)
    inv = 1.0 / m if mass_scaled else torch.ones_like(m)
    e, f = energy_and_full_forces(models, x, types, box)
    scal = [(float(e), float(kinetic_energy(v, m)))]
    for _ in range(n_steps):
        x = x + v * dt + 0.5 * (f * inv) * dt * dt
        if box is not None:
            x = wrap_into_box(x, box)
        e, f_new = energy_and_full_forces(models, x, types, box)
        v = v + 0.5 * (f * inv + f_new * inv) * dt
        f = f_new
        scal.append((float(e), float(kinetic_energy(v, m))))
    return x, v, f, torch.tensor(scal, dtype=torch.float64)


# ----------------------------------------------------------------------------- MD pieces
def verlet_positions(x: Tensor, v: Tensor, f: Tensor, dt: float) -> Tensor:
    """pantea/simulation/molecular_dynamics.py:16-21 (no mass division)."""
    return x + v * dt + 0.5 * f * dt * dt


def verlet_velocities(v: Tensor, f: Tensor, f_new: Tensor, dt: float) -> Tensor:
    """pantea/simulation/molecular_dynamics.py:23-30."""
    return v + 0.5 * (f + f_new) * dt


def kinetic_energy(v: Tensor, m: Tensor) -> Tensor:
    """pantea/simulation/system.py:20-22; m is [N,1]."""
    return 0.5 * torch.sum(m * v * v)


def temperature(v: Tensor, m: Tensor, kb: float) -> Tensor:
    """pantea/simulation/system.py:25-29."""
    return 2 * kinetic_energy(v, m) / (3 * v.shape[0] * kb)


def berendsen_scale(v: Tensor, dt: float, tau: float, t_now: Tensor, t_target: float) -> Tensor:
    """pantea/simulation/thermostat.py:12-22."""
    return v * (1.0 / torch.sqrt(1.0 + (dt / tau) * (t_now / t_target - 1.0)))


# ----------------------------------------------------------------------------- adaptors
def symfuncs_from_spec(spec) -> List[SymFunc]:
    """oracle.spec.ElementSpec -> list of SymFunc."""
    return [SymFunc(s.kind, s.cutoff_type, s.r_cutoff, s.type_j, s.type_k, s.eta, s.r_shift, s.lambda0, s.zeta)
            for s in spec.symfuncs]


def models_from_specs(specs) -> List[ElementModel]:
    out = []
    for spec in specs:
        params = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in spec.scaler.items()}
        layers = [(torch.as_tensor(k, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64), a)
                  for k, b, a in spec.layers]
        out.append(ElementModel(spec.atom_type, symfuncs_from_spec(spec), spec.scale_type, params, layers,
                                spec.scale_min, spec.scale_max))
    return out


# ----------------------------------------------------------------------------- Lennard-Jones (driver tests)
def lj_energy_and_gradient(positions: Tensor, box: Optional[Tensor], sigma: float, epsilon: float,
                           r_cutoff: float) -> Tuple[Tensor, Tensor]:
    """Dense LJ energy and the reference's "forces" (= +dE/dr) -- pantea/simulation/lennard_jones.py:73-123."""
    r, d = distances_with_aux(positions, positions, box)
    mask = cutoff_mask(r, r_cutoff)
    rs = torch.where(mask, r, torch.ones_like(r))
    t6 = (sigma / rs) ** 6
    e_pair = torch.where(mask, 4.0 * epsilon * t6 * (t6 - 1.0), torch.zeros_like(r))
    coef = torch.where(mask, -24.0 * epsilon / (rs * rs) * t6 * (2.0 * t6 - 1.0), torch.zeros_like(r))
    return 0.5 * e_pair.sum(), (coef[..., None] * d).sum(dim=1)
