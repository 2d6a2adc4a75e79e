"""Host-side logic of the multi-GPU paths on CPU: slab partitioning, the all-reduce(MAX) composite over a
world_size-2/3 gloo group (with the oracle standing in for the per-rank renderer), frame sharding."""
import os
import socket

import numpy as np
import pytest

import scenes
from spimagine_b200.multigpu import (composite_max, frames_for_rank, partition_slabs, partition_slabs_multi,
                                     slab_with_halo)


def test_partition_slabs_covers_everything_once():
    for nz, world in ((61, 2), (61, 3), (2048, 8), (8, 8), (9, 4)):
        parts = partition_slabs(nz, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == nz
        for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
            assert a1 == b0 and a1 > a0
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition_slabs(3, 4)
    assert slab_with_halo(0, 10, 40) == (0, 11)
    assert slab_with_halo(10, 20, 40) == (9, 21)
    assert slab_with_halo(30, 40, 40) == (29, 40)


def test_serpentine_slab_assignment():
    for nz, world, k in ((2048, 8, 2), (61, 3, 3), (64, 1, 4), (37, 4, 1)):
        per_rank = partition_slabs_multi(nz, world, k)
        assert len(per_rank) == world and all(len(p) == k for p in per_rank)
        flat = sorted(sum(per_rank, []))
        assert flat == partition_slabs(nz, world * k)  # every slice exactly once
    # front and back slabs pair up: rank r gets slab r and slab 2*world-1-r
    parts = partition_slabs(2048, 16)
    per_rank = partition_slabs_multi(2048, 8, 2)
    for r in range(8):
        assert per_rank[r] == [parts[r], parts[15 - r]]


def test_frame_sharding():
    world = 8
    owned = [frames_for_rank(100, r, world) for r in range(world)]
    assert sorted(sum(owned, [])) == list(range(100))
    assert all(f % world == r for r in range(world) for f in owned[r])
    assert max(map(len, owned)) - min(map(len, owned)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        data = scenes.vol_g(0, np.uint16, seed=4, shape=(37, 40, 44))
        M, P = scenes.gui_camera(0.7, 3.1)
        o = oracle.OracleRenderer((64, 48), kind="port", pos_mode=2, weight_bits=8)
        o.set_data(data)
        o.set_modelView(M)
        o.set_projection(P)
        z0, z1 = partition_slabs(37, world)[rank]
        part = torch.from_numpy(o.render_raw(z0, z1))
        composite_max(part)                      # all_reduce(MAX) over the gloo group
        full = o.render_raw()
        ok = bool(np.array_equal(part.numpy(), full))
        # every rank ends with the same image
        gathered = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)
        same = all(bool(torch.equal(g, part)) for g in gathered)
        q.put((rank, ok, same))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sort_last_composite_over_gloo(oracle_mod, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == list(range(world))
    assert all(ok and same for _, ok, same in results)
