"""N>1 path on CPU: world_size-2 gloo run of the shard / gather plumbing (no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dxtb_b200.parallel import gather_by_index, gather_results, shard_bounds, shard_by_cost


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1024, 8192):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_cost_balances():
    cost = torch.tensor([float(n) ** 3 for n in range(30, 650, 7)])
    parts = shard_by_cost(cost, 8)
    assert sorted(torch.cat(parts).tolist()) == list(range(cost.numel()))
    loads = [float(cost[p].sum()) for p in parts]
    assert max(loads) / min(loads) < 1.25


def _worker(rank, world, port, n_total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_bounds(n_total, world, rank)
    # stand-in for the per-shard single point: a deterministic function of the global molecule index
    idx = torch.arange(a, b, dtype=torch.float64)
    e_local = -idx * 1.5 - 7.0
    f_local = torch.stack([idx, -idx, 2 * idx], dim=-1).unsqueeze(1).expand(-1, 4, -1).contiguous()
    e = gather_results(e_local, n_total)
    f = gather_results(f_local, n_total)
    full = torch.arange(n_total, dtype=torch.float64)
    assert torch.equal(e, -full * 1.5 - 7.0)
    assert f.shape == (n_total, 4, 3) and torch.equal(f[:, 0, 2], 2 * full)
    # cost-balanced (non-contiguous) shards: the same gather by global molecule id
    cost = torch.tensor([float((7 * i) % 5 + 1) ** 3 for i in range(n_total)])
    parts = shard_by_cost(cost, world)
    mine = parts[rank].to(torch.float64)
    e2 = gather_by_index(-mine * 1.5 - 7.0, parts, n_total)
    f2 = gather_by_index(torch.stack([mine, -mine, 2 * mine], dim=-1).unsqueeze(1).expand(-1, 4, -1).contiguous(), parts, n_total)
    assert torch.equal(e2, e) and torch.equal(f2, f)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 11), nprocs=2, join=True)
