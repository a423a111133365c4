"""world_size-2 gloo test of the multi-GPU host logic: block ownership, the in-place position all-gather and the
scalar reductions used by ReplicatedMD (the CUDA kernels themselves are covered by the gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pantea_b200.distributed import BlockLayout, all_reduce_max, all_reduce_sum


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_atoms, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layout = BlockLayout(n_atoms, rank, world)
        ref = torch.arange(n_atoms * 3, dtype=torch.float64).reshape(n_atoms, 3)
        buf = layout.allocate(torch.zeros_like(ref))
        lo, hi = layout.owned()
        buf[lo:hi] = ref[lo:hi] + 1.0          # every rank "integrates" only its own block ...
        layout.exchange(buf)                   # ... and one in-place all-gather makes everybody consistent
        ok_gather = torch.equal(buf[:n_atoms], ref + 1.0)
        s = all_reduce_sum(torch.tensor([float(hi - lo)], dtype=torch.float64))
        m = all_reduce_max(torch.tensor([float(rank)], dtype=torch.float64))
        out[rank] = (ok_gather, float(s), float(m), lo, hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_atoms", [10, 11, 96])
def test_block_layout_exchange_world2(n_atoms):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_atoms, out), nprocs=world, join=True)
        res = dict(out)
    assert all(res[r][0] for r in range(world))
    assert res[0][1] == n_atoms and res[0][2] == world - 1
    assert res[0][3] == 0 and res[0][4] == res[1][3] and res[1][4] == n_atoms     # contiguous, complete cover


def test_block_layout_properties():
    for n, w in ((10, 4), (99999, 8), (7, 8), (1, 1)):
        covered = []
        for r in range(w):
            lo, hi = BlockLayout(n, r, w).owned()
            assert 0 <= lo <= hi <= n and hi - lo <= BlockLayout(n, r, w).block
            covered += list(range(lo, hi))
        assert covered == list(range(n))
        assert BlockLayout(n, 0, w).padded % w == 0 and BlockLayout(n, 0, w).padded >= n


def _scaler_worker(rank, world, port, out):
    from pantea_b200.descriptors.scaler import DescriptorScaler
    from pantea_b200.distributed import merge_scaler_params
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = torch.from_numpy(np.random.default_rng(42).normal(2.0, 3.0, size=(37, 5)))
        mine = data[rank::world]                       # structures are split index mod world
        params = DescriptorScaler.fit(mine[:3])
        params = DescriptorScaler.partial_fit(params, mine[3:])
        merged = merge_scaler_params(params)
        empty = merge_scaler_params(params if rank == 0 else None)   # a rank without samples of the element
        out[rank] = (merged.nsamples.item(), merged.mean.numpy(), merged.sigma.numpy(), merged.minval.numpy(),
                     merged.maxval.numpy(), empty.nsamples.item(), empty.mean.numpy())
    finally:
        dist.destroy_process_group()


def test_scaler_statistics_merge_across_ranks_world2():
    """SURVEY 8(e) dataset preprocessing: per-rank scaler statistics merged into the whole-dataset ones."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_scaler_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    data = np.random.default_rng(42).normal(2.0, 3.0, size=(37, 5))
    for r in range(world):
        n, mean, sigma, mn, mx, n_e, mean_e = res[r]
        assert n == 37
        np.testing.assert_allclose(mean, data.mean(0), rtol=1e-13)
        np.testing.assert_allclose(sigma, data.std(0), rtol=1e-12)
        np.testing.assert_array_equal(mn, data.min(0))
        np.testing.assert_array_equal(mx, data.max(0))
        assert n_e == len(data[0::2])
        np.testing.assert_allclose(mean_e, data[0::2].mean(0), rtol=1e-13)
    for k in range(1, 5):
        np.testing.assert_array_equal(res[0][k], res[1][k])        # every rank holds bitwise the same statistics
