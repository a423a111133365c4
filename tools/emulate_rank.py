"""Per-kernel cost of ONE rank's share of a W-way block-owned MD step, measured on one GPU (no collectives):
python tools/emulate_rank.py [--atoms 100000] [--world 8].  Diagnostic for the strong-scaling overheads."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from pantea_b200 import _lib, engine  # noqa: E402
from pantea_b200.potentials import NeuralNetworkPotential  # noqa: E402
from pantea_b200.utils.synthetic import water_box  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=100000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--all-ranks", action="store_true", help="time every rank's share (load balance), events only")
    ap.add_argument("--shuffle", action="store_true", help="random atom order: every block is a random sample of the box")
    args = ap.parse_args()
    nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
    nnp.load()
    pot = nnp.device_potential()
    pos_h, types_h, box_h = water_box(args.atoms - args.atoms % 3)
    n = len(pos_h)
    if args.shuffle:
        perm = np.random.default_rng(0).permutation(n)
        pos_h, types_h = pos_h[perm], types_h[perm]
    dev = torch.device("cuda")
    pos = torch.as_tensor(pos_h, dtype=torch.float64, device=dev)
    types = torch.as_tensor(types_h, dtype=torch.int32, device=dev)
    box = [float(b) for b in box_h]
    density = n / (box[0] * box[1] * box[2])
    ws = engine.Workspace(pot, n, engine.estimate_max_neighbors(pot.r_cutoff, density, n), torch.float64)
    per = (n + args.world - 1) // args.world
    owned = (0, per)
    frc = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    for _ in range(3):
        ws.bind(pos, types, box, pot.r_cutoff, owned=owned)
        ws.energy_forces(False, True, out_forces=frc)
    if args.all_ranks:
        times = []
        for r in range(args.world):
            own = (r * per, min(n, (r + 1) * per))
            for _ in range(2):
                ws.bind(pos, types, box, pot.r_cutoff, check=False, owned=own)
                ws.energy_forces(False, True, out_forces=frc)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(10):
                ws.bind(pos, types, box, pot.r_cutoff, check=False, owned=own)
                _lib.check(_lib.load().pantea_energy_forces(ws.handle, None, _lib.ptr(frc), None, 0, _lib.stream_ptr()))
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) / 10)
        print(f"world={args.world} shuffle={args.shuffle} per-rank build+force ms: " + " ".join(f"{t:.4f}" for t in times)
              + f" | max {max(times):.4f} mean {sum(times) / len(times):.4f} max/mean {max(times) * len(times) / sum(times):.3f}")
        return
    from torch.profiler import ProfilerActivity, profile
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        ws.bind(pos, types, box, pot.r_cutoff, check=False, owned=owned)
        _lib.check(_lib.load().pantea_energy_forces(ws.handle, None, _lib.ptr(frc), None, 0, _lib.stream_ptr()))
    ev1.record()
    torch.cuda.synchronize()
    print(f"owned {per} of {n} atoms: {ev0.elapsed_time(ev1) / reps:.4f} ms per build + force evaluation (events)")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ws.bind(pos, types, box, pot.r_cutoff, check=False, owned=owned)
            _lib.check(_lib.load().pantea_energy_forces(ws.handle, None, _lib.ptr(frc), None, 0, _lib.stream_ptr()))
        torch.cuda.synchronize()
    tot = 0.0
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
        tot += e.device_time_total / 5
        print(f"  {e.key[:70]:70s} n={e.count:3d} avg={e.device_time_total / e.count / 1e3:8.4f} ms")
    print(f"  sum of kernel times per step: {tot / 1e3:.4f} ms")


if __name__ == "__main__":
    main()
