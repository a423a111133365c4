"""Brick-decomposed molecular dynamics over NVLink peer memory: host side of `csrc/mgpu.cu`.

One process per GPU.  `torch.distributed` is used for set-up only (exchange of the CUDA-IPC handles, barriers around a
state reset, gathers for diagnostics); the steps themselves are replays of one CUDA graph per rank in which the
integration kernel stores ghost positions straight into the peers' mailboxes and a wait kernel synchronises on
step-number flags -- no collective library call, no host synchronisation (see the header of `csrc/mgpu.cu`).

The brick boundaries sit at the atom-count quantiles of the initial configuration (`halo.BrickGrid.balanced`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from pantea_b200.halo import BrickGrid


class BrickMD:
    """Velocity-Verlet MD (reference integrator: no mass, NVE) of one periodic box on `world` GPUs."""

    def __init__(self, device_potential, positions: torch.Tensor, velocities: torch.Tensor, types: torch.Tensor,
                 box: Sequence[float], time_step: float, rank: int = 0, world: int = 1,
                 grid: Optional[BrickGrid] = None, own_slack: float = 1.25) -> None:
        from pantea_b200 import _lib, engine

        self._lib, self.lib = _lib, _lib.load()
        self.pot = device_potential
        self.rank, self.world = int(rank), int(world)
        n = int(positions.shape[0])
        self.n, self.dt, self.box = n, float(time_step), [float(b) for b in box]
        self.dtype, self.code = positions.dtype, _lib.dtype_code(positions.dtype)
        self.device = positions.device
        self.types = types.to(torch.int32).contiguous()
        self.grid = grid if grid is not None else BrickGrid.balanced(self.box, self.world, positions)
        density = n / (self.box[0] * self.box[1] * self.box[2])
        self.ws = engine.Workspace(device_potential, n, engine.estimate_max_neighbors(device_potential.r_cutoff, density, n),
                                   self.dtype)
        self.own_cap = n if self.world == 1 else min(n, int(own_slack * n / self.world) + 1024)
        self.handle = C.c_void_p()
        self._create()
        self.steps = 0
        self.reset(positions, velocities)

    # ------------------------------------------------------------------------------------------ set-up
    def _create(self) -> None:
        _lib, lib = self._lib, self.lib
        dims = (C.c_int32 * 3)(*self.grid.dims)
        cuts = [(C.c_double * max(len(c), 1))(*(c if c else [0.0])) for c in self.grid.cuts]
        _lib.check(lib.pantea_mgpu_create(self.ws.handle, self.rank, self.world, self.n, _lib.box_arg(self.box), dims,
                                          cuts[0] if self.grid.cuts[0] else None, cuts[1] if self.grid.cuts[1] else None,
                                          cuts[2] if self.grid.cuts[2] else None, float(self.pot.r_cutoff), self.dt,
                                          int(self.own_cap), C.byref(self.handle)))
        if self.world > 1:
            nb = int(lib.pantea_mgpu_handle_bytes())
            mine = (C.c_ubyte * nb)()
            _lib.check(lib.pantea_mgpu_export_handle(self.handle, mine))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine))
            blob = b"".join(gathered)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            _lib.check(lib.pantea_mgpu_connect(self.handle, buf))

    def _barrier(self) -> None:
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()

    def reset(self, positions: torch.Tensor, velocities: torch.Tensor) -> None:
        """(Re)start from full-length replicated arrays; all ranks call it together (barriers on both sides)."""
        _lib = self._lib
        self._barrier()
        pos = positions.to(self.dtype).contiguous()
        vel = velocities.to(self.dtype).contiguous()
        for attempt in range(4):  # the first evaluation may have to grow a capacity
            _lib.check(self.lib.pantea_mgpu_set_state(self.handle, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(self.types),
                                                      _lib.stream_ptr()))
            mx = C.c_int32(0)
            code = self.lib.pantea_neighbor_status(self.ws.handle, C.byref(mx), _lib.stream_ptr())
            if code != _lib.PANTEA_ECAPACITY:
                _lib.check(code)
                break
            if mx.value > (self.ws.max_neighbors + 31) // 32 * 32:
                raise _lib.CapacityError(_lib.last_error())
        self.steps = 0
        self._barrier()

    # ------------------------------------------------------------------------------------------ stepping
    def run(self, n_steps: int, use_graph: bool = True) -> None:
        _lib = self._lib
        _lib.check(self.lib.pantea_mgpu_run(self.handle, int(n_steps), 1 if use_graph else 0, _lib.stream_ptr()))
        self.steps += int(n_steps)

    def step(self) -> None:
        self.run(1)

    def check_capacity(self) -> int:
        """Sticky device-side capacity flags (see ReplicatedMD.check_capacity) and the engine's own status."""
        _lib = self._lib
        status = C.c_int32(0)
        _lib.check(self.lib.pantea_mgpu_read(self.handle, None, None, None, None, C.byref(status), _lib.stream_ptr()))
        if status.value != 0:
            raise RuntimeError(f"brick engine status {status.value} (1 + r: peer r never published a step; 1000: own_cap exceeded)")
        mx = C.c_int32(0)
        _lib.check(self.lib.pantea_neighbor_status(self.ws.handle, C.byref(mx), _lib.stream_ptr()))
        return int(mx.value)

    # ------------------------------------------------------------------------------------------ diagnostics
    def read(self):
        """(positions, velocities, forces, roles) of this rank: full-length arrays, see pantea_mgpu_read."""
        _lib = self._lib
        pos = torch.zeros((self.n, 3), dtype=self.dtype, device=self.device)
        vel, frc = torch.zeros_like(pos), torch.zeros_like(pos)
        role = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.pantea_mgpu_read(self.handle, _lib.ptr(pos), _lib.ptr(vel), _lib.ptr(frc), _lib.ptr(role), None,
                                             _lib.stream_ptr()))
        return pos, vel, frc, role

    def gather(self):
        """Global (positions, velocities, forces) assembled from every rank's owned rows, plus the owner count per atom
        (must be exactly one everywhere)."""
        pos, vel, frc, role = self.read()
        own = (role == 2)
        out = [torch.where(own[:, None], t, torch.zeros_like(t)) for t in (pos, vel, frc)]
        count = own.to(torch.int32)
        if self.world > 1:
            for t in out:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dist.all_reduce(count, op=dist.ReduceOp.SUM)
        return out[0], out[1], out[2], count

    def owned_count(self) -> int:
        return int((self.read()[3] == 2).sum().item())

    def close(self) -> None:
        if self.handle:
            self.lib.pantea_mgpu_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
