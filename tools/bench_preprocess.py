"""BASELINE.json configs[1]: ACSF descriptor + gradient batch over synthetic 192-atom water structures
(dataset preprocessing).  Many structures go through one batched launch sequence (all-pairs neighbour mode,
per-structure box).  Prints one JSON line: structures/s and descriptor rows/s; checks a sample against the oracle.

    python tools/bench_preprocess.py [--structures 2000] [--reps 5]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pantea_b200 import engine  # noqa: E402
from pantea_b200.potentials import NeuralNetworkPotential  # noqa: E402
from pantea_b200.utils.synthetic import water_box  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--structures", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
    nnp.load()
    pot = nnp.device_potential()
    dev = torch.device("cuda")
    S, n = args.structures, 192
    structs = [water_box(n, seed=2024 + s) for s in range(S)]
    pos = torch.as_tensor(np.concatenate([s[0] for s in structs]), device=dev)
    types = torch.as_tensor(np.concatenate([s[1] for s in structs]), dtype=torch.int32, device=dev)
    boxes = torch.as_tensor(np.stack([s[2] for s in structs]), device=dev)
    ptr = torch.arange(S + 1, dtype=torch.int32, device=dev) * n
    ws = engine.Workspace(pot, S * n, 191, torch.float64)
    idx = {el: torch.nonzero(types == pot.type_of[el]).flatten().to(torch.int32) for el in ("H", "O")}

    def run():
        ws.bind_batch(pos, types, ptr, boxes, pot.r_cutoff, check=False)
        out = {}
        for el in ("H", "O"):
            out[el] = ws.acsf(pot.slot(el), pot.n_symfunc[el], idx[el], True, True)
        return out

    out = run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        out = run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.reps
    # oracle spot check on the first structure
    from oracle import c_oracle
    from oracle.spec import load_potential
    specs = load_potential(ROOT / "tests" / "golden" / "h2o.json")
    p0, t0_, b0 = structs[0]
    ok = True
    for spec, el in zip(specs, ("H", "O")):
        centres = np.nonzero(t0_ == spec.atom_type)[0]
        G_o, dG_o = c_oracle.acsf(spec, p0, t0_, b0, centres)
        G, dG = out[el]
        k = len(centres)
        ok &= np.abs(G[:k].cpu().numpy() - G_o).max() < 1e-10 * np.abs(G_o).max()
        ok &= np.abs(dG[:k].cpu().numpy() - dG_o).max() < 1e-10 * np.abs(dG_o).max()
    print(json.dumps({"workload": f"{S} x 192-atom water structures, ACSF values + gradients (h2o.json)",
                      "structures_per_s": S / dt, "atoms_per_s": S * n / dt, "ms_per_batch": dt * 1e3,
                      "parity_first_structure": bool(ok), "dtype": "f64"}))


if __name__ == "__main__":
    main()
