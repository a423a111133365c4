"""Generates tests/golden/extension_vectors.json: known answers for the two extensions the reference does not have
(PANTEA_FORCE_FULL and mass-scaled integration), computed by the dense float64 oracle (autograd through both roles of
the restated reference energy) on the reference's own fixture `h2o.data[0]` + `h2o.json`.

    python tests/golden/make_extension_vectors.py

The reference cannot produce these numbers (it differentiates the central copy of the positions only and integrates
without mass); what ties them to it is the energy expression, pinned by tests/test_oracle_golden.py, and the contrast
SURVEY.md Appendix C records for this structure (full force differs from the reference force by up to 0.145).
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import dense_oracle  # noqa: E402
from oracle.spec import load_potential, md_velocities, read_runner, water_masses  # noqa: E402

GOLDEN = Path(__file__).resolve().parent


def main() -> None:
    pot = load_potential(GOLDEN / "h2o.json")
    models = dense_oracle.models_from_specs(pot)
    frame = read_runner(GOLDEN / "h2o.data")[0]
    box = frame["box"]
    pos = np.remainder(frame["positions"], box)
    T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))  # noqa: E731
    types = torch.from_numpy(frame["types"])
    e, f_full = dense_oracle.energy_and_full_forces(models, T(pos), types, T(box))
    _, _, f_ref = dense_oracle.energy_and_forces(models, T(pos), types, T(box))
    vel, mass = md_velocities(frame["types"]), water_masses(frame["types"])
    dt, n_steps = 5.0, 8
    x, v, f, sc = dense_oracle.md_run_full(models, T(pos), T(vel), T(mass), types, T(box), dt, n_steps)
    out = {
        "source": "oracle/dense_oracle.py (energy_and_full_forces, md_run_full) on tests/golden/h2o.data[0] + h2o.json",
        "energy": float(e),
        "full_forces": f_full.numpy().tolist(),
        "max_abs_difference_to_reference_force": float((f_full - f_ref).abs().max()),
        "md_full_mass_scaled": {"dt": dt, "n_steps": n_steps, "velocities0": vel.tolist(), "masses": mass.tolist(),
                                "positions": x.numpy().tolist(), "velocities": v.numpy().tolist(),
                                "e_pot_e_kin": sc.numpy().tolist()},
    }
    (GOLDEN / "extension_vectors.json").write_text(json.dumps(out, indent=1))
    print("wrote", GOLDEN / "extension_vectors.json")


if __name__ == "__main__":
    main()
