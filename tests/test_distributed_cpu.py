"""Host-side multi-rank logic (frame sharding, S2 planning, one-frame halo exchange) on CPU with
the gloo backend, world_size = 2.  The kernels themselves need no collective: frames are
independent units; only sweep 1 of schedule S2 needs the left neighbour's last sweep-0 frame."""

import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keypoints2body_b200.distributed import (exchange_halo, pack_halo, plan_two_sweep, shard_range,
                                             unpack_halo)


def test_shard_range_covers_all_frames():
    for n in (1, 7, 8, 4096, 8 * 1048576 + 3):
        for world in (1, 2, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_plan_two_sweep():
    seq = torch.tensor([5, 6, 0, 1, 2, 0, 1], dtype=torch.int32)
    it0, it1, keep, starts = plan_two_sweep(seq, 30, 10)
    assert it0.tolist() == [30] * 7
    assert it1.tolist() == [10, 10, 0, 10, 10, 0, 10]
    assert keep.tolist() == [1, 1, 0, 1, 1, 0, 1]
    assert starts.tolist() == [2, 5]


def test_halo_pack_roundtrip():
    g = torch.Generator().manual_seed(0)
    last = {"global_orient": torch.randn(1, 3, generator=g), "body_pose": torch.randn(1, 69, generator=g),
            "betas": torch.randn(1, 10, generator=g), "transl": torch.randn(1, 3, generator=g)}
    row = pack_halo(last)
    assert row.shape == (95,)
    back = unpack_halo(row, with_expression=False)
    for k, v in last.items():
        assert torch.equal(back[k], v)
    assert "expression" not in back


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank "fits" its shard: sweep-0 parameters of frame g are a function of g only
        lo, hi = shard_range(10, rank, world)
        g = torch.arange(lo, hi, dtype=torch.float32)
        last = {"global_orient": g[-1:].reshape(1, 1).expand(1, 3).clone(),
                "body_pose": (100 + g[-1:]).reshape(1, 1).expand(1, 69).clone(),
                "betas": torch.zeros(1, 10), "transl": (-g[-1:]).reshape(1, 1).expand(1, 3).clone()}
        halo = exchange_halo(last)
        if rank == 0:
            assert halo is None
        else:
            prev = float(lo - 1)              # the left neighbour's last global frame
            assert torch.all(halo["global_orient"] == prev)
            assert torch.all(halo["body_pose"] == 100 + prev)
            assert torch.all(halo["transl"] == -prev)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_halo_exchange_world_size_2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert dict(out) == {0: 1, 1: 1}
