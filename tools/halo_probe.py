"""Diagnostic (one GPU): how far does the fastest atom of the bench water box move within s steps?  Sizes the ghost-shell
skin / rebuild interval of HaloMD (an atom may move skin / 2 between two rebuilds).  usage: python tools/halo_probe.py [atoms]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pantea_b200.distributed import ReplicatedMD  # noqa: E402
from pantea_b200.potentials import NeuralNetworkPotential  # noqa: E402
from pantea_b200.utils.synthetic import md_velocities, water_box, water_masses  # noqa: E402

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 99999
dev = torch.device("cuda", 0)
nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
nnp.load()
pot = nnp.device_potential()
pos, types, box = water_box(n_atoms)
t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)  # noqa: E731
md = ReplicatedMD(pot, t(pos), t(md_velocities(types)), t(water_masses(types)), t(types, torch.int32), list(box), 0.25)
bx = torch.tensor(list(box), dtype=torch.float64, device=dev)
hist = [md.pos.clone()]
for s in range(25):
    md.step()
    hist.append(md.pos.clone())


def disp(a, b):
    d = a - b
    d -= bx * torch.round(d / bx)
    return float(d.norm(dim=1).max())


for span in (1, 2, 3, 4, 6, 8, 12, 25):
    worst = max(disp(hist[s + span], hist[s]) for s in range(0, 26 - span))
    print(f"max displacement over any {span:2d}-step window of the 25-step segment: {worst:.4f} Bohr  (needs skin >= {2 * worst:.3f})")
