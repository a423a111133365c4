"""Fast path (csrc/acsf2.cu) against the generic kernels and the C oracle on the benchmark box; per-kernel times.
usage (GPU box): python tools/v2_check.py [n_atoms] [oracle_sample]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from pantea_b200 import _lib, engine  # noqa: E402
from pantea_b200.potentials import NeuralNetworkPotential  # noqa: E402
from pantea_b200.utils.synthetic import water_box  # noqa: E402

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 99999
n_oracle = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
lib = _lib.load()
dev = torch.device("cuda", 0)
nnp = NeuralNetworkPotential.from_runner(ROOT / "tests" / "golden" / "h2o.json")
nnp.load()
pot = nnp.device_potential()
pos_h, types_h, box_h = water_box(n_atoms)
n = len(pos_h)
pos = torch.as_tensor(pos_h, dtype=torch.float64, device=dev)
types = torch.as_tensor(types_h, dtype=torch.int32, device=dev)
box = [float(b) for b in box_h]
ws = engine.Workspace(pot, n, engine.estimate_max_neighbors(pot.r_cutoff, n / np.prod(box), n), torch.float64)
ws.bind(pos, types, box, pot.r_cutoff)


def forces(fast):
    lib.pantea_set_fast_path(1 if fast else 0)
    e, ea, f = ws.energy_forces(want_energy=True, want_forces=True, want_atomic=True)
    torch.cuda.synchronize()
    return float(e), ea.clone(), f.clone()


def timed(fast, reps=5):
    lib.pantea_set_fast_path(1 if fast else 0)
    f = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    for _ in range(2):
        _lib.check(lib.pantea_energy_forces(ws.handle, None, _lib.ptr(f), None, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        _lib.check(lib.pantea_energy_forces(ws.handle, None, _lib.ptr(f), None, 0, _lib.stream_ptr()))
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


lib.pantea_set_gauss_screen(0.0)
e0, ea0, f0 = forces(False)
e2, ea2, f2 = forces(True)
lib.pantea_set_gauss_screen(40.0)
e1, ea1, f1 = forces(True)
r2 = ((f2 - f1).abs() / f2.abs().clamp_min(1e-300)).max().item()
print(f"screening on vs off (fast path): max elementwise rel {r2:.3e}, max abs {(f2 - f1).abs().max().item():.3e}, dE {abs(e2 - e1):.3e}")
cnt = torch.zeros(4, dtype=torch.int64, device=dev)
_lib.check(lib.pantea_workspace_set_counters(ws.handle, _lib.ptr(cnt)))
for thr in (0.0, 40.0):
    cnt.zero_()
    lib.pantea_set_gauss_screen(thr)
    forces(True)
    print(f"screen {thr}: list entries evaluated {int(cnt[3])} ({int(cnt[3]) / n:.1f} per atom)")
_lib.check(lib.pantea_workspace_set_counters(ws.handle, None))
den = f0.abs().clamp_min(1e-300)
rel = ((f1 - f0).abs() / den)
print(f"n={n} generic E={e0:.12f} fast E={e1:.12f} dE/E={abs(e1 - e0) / abs(e0):.2e}")
print(f"fast vs generic forces: max elementwise rel {rel.max().item():.3e}, max abs {(f1 - f0).abs().max().item():.3e}, "
      f"max|F| {f0.abs().max().item():.3e}; e_atom max abs {(ea1 - ea0).abs().max().item():.3e}")
print(f"force evaluation: generic {timed(False):.4f} ms, fast {timed(True):.4f} ms")
ws.set_compute_precision(32)
e3, ea3, f3 = forces(True)
rms = float(f0.pow(2).mean().sqrt())
print(f"mixed (FP32 symmetry functions) vs generic: max |dF| / (|F| + rms) {((f3 - f0).abs() / (f0.abs() + rms)).max().item():.3e}, "
      f"dE/E {abs(e3 - e0) / abs(e0):.2e}; force evaluation {timed(True):.4f} ms")
ws.set_compute_precision(64)
if n_oracle > 0:
    from oracle import c_oracle
    from oracle.spec import load_potential
    specs = load_potential(ROOT / "tests" / "golden" / "h2o.json")
    t0 = time.perf_counter()
    _, ea_o, f_o = c_oracle.energy_forces(specs, pos_h, types_h, box_h, begin=0, end=min(n, n_oracle))
    print(f"oracle {min(n, n_oracle)} atoms in {time.perf_counter() - t0:.1f} s")
    m = min(n, n_oracle)
    f_o = torch.as_tensor(f_o[:m], device=dev)
    for name, f in (("generic", f0), ("fast", f1)):
        r = ((f[:m] - f_o).abs() / f_o.abs().clamp_min(1e-300)).max().item()
        print(f"{name} vs oracle: max elementwise rel {r:.3e}, max abs {(f[:m] - f_o).abs().max().item():.3e}")
from torch.profiler import ProfilerActivity, profile
for fast in (False, True, 32):
    lib.pantea_set_fast_path(1 if fast else 0)
    ws.set_compute_precision(32 if fast == 32 else 64)
    f = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ws.bind(pos, types, box, pot.r_cutoff, check=False)
            _lib.check(lib.pantea_energy_forces(ws.handle, None, _lib.ptr(f), None, 0, _lib.stream_ptr()))
        torch.cuda.synchronize()
    for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
        print(f"[{'mixed32' if fast == 32 else 'fast' if fast else 'generic'}] {ev.key[:60]:60s} n={ev.count:3d} avg={ev.device_time_total / ev.count / 1e3:8.4f} ms")
lib.pantea_set_fast_path(1)
ws.set_compute_precision(64)
