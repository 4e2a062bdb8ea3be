"""Host-side multi-GPU logic on CPU: target partitioning, action localisation and the per-step
all-gather of packed vis words + summaries, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest


def test_partition_covers_every_target_once():
    from clockpunch_b200.sharding import TargetPartition
    for n, w in ((10, 1), (10, 3), (7, 8), (1_000_000, 8), (0, 2), (125, 4)):
        p = TargetPartition(n, w)
        seen = np.zeros(n, int)
        for r in range(w):
            lo, hi = p.bounds(r)
            assert hi - lo == p.count(r) <= p.cap
            seen[lo:hi] += 1
            for t in (lo, hi - 1):
                if lo < hi:
                    assert p.owner(t) == r
        assert (seen == 1).all()
        assert sum(p.count(r) for r in range(w)) == n
    with pytest.raises(ValueError):
        TargetPartition(5, 0)


def test_localize_actions():
    from clockpunch_b200.sharding import TargetPartition
    p = TargetPartition(10, 3)          # shards [0,4) [4,7) [7,10)
    a = np.array([0, 3, 4, 6, 7, 9, 10])   # 10 = inaction
    np.testing.assert_array_equal(p.localize_actions(a, 0), [0, 3, 4, 4, 4, 4, 4])
    np.testing.assert_array_equal(p.localize_actions(a, 1), [3, 3, 0, 2, 3, 3, 3])
    np.testing.assert_array_equal(p.localize_actions(a, 2), [3, 3, 3, 3, 0, 2, 3])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, m, out_dir):
    import torch
    import torch.distributed as dist
    from clockpunch_b200.sharding import StepGather, TargetPartition
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        part = TargetPartition(n_total, world)
        lo, hi = part.bounds(rank)
        g = np.random.default_rng(123)          # same global arrays on every rank
        wpr = (m + 31) // 32
        glob = {"vis_est": g.integers(-2**31, 2**31 - 1, size=(n_total, wpr)).astype(np.int32),
                "tr_pos": g.random(n_total).astype(np.float32), "tr_vel": g.random(n_total).astype(np.float32),
                "last_time_tasked": g.random(n_total).astype(np.float32),
                "num_tasked": g.integers(0, 50, n_total).astype(np.int64)}
        local = {k: torch.from_numpy(v[lo:hi].copy()) for k, v in glob.items()}
        sg = StepGather(part, rank, local)
        for it in range(2):                     # buffers are reused step after step
            if it == 1:
                local["num_tasked"] += 1
                glob["num_tasked"] = glob["num_tasked"] + 1
            out = sg.gather()
            for k, v in glob.items():
                np.testing.assert_array_equal(out[k].numpy(), v, err_msg=f"{k} rank {rank} it {it}")
        assert sg.bytes_per_step == world * part.cap * (4 * wpr + 4 + 4 + 4 + 8)
        # the packed single-collective form used by bench.py (one summary buffer per shard)
        from clockpunch_b200.engine import CatalogEngine
        from clockpunch_b200.sharding import PackedGather
        if n_total % world == 0:
            n = n_total // world
            buf = torch.zeros(n * (8 + 4 * wpr + 12), dtype=torch.uint8)
            v = CatalogEngine.summary_views(buf, n, wpr)
            for k in v:
                v[k].copy_(torch.from_numpy(glob[k][lo:hi].copy()))
            pg = PackedGather(world, rank, buf, n, wpr)
            pg.gather()
            for r in range(world):
                a, b = part.bounds(r)
                for k, t in pg.fields(r).items():
                    np.testing.assert_array_equal(t.numpy(), glob[k][a:b], err_msg=f"packed {k} from rank {r}")
            assert pg.bytes_per_step == world * n * (8 + 4 * wpr + 12)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 37])       # equal shards (zero-copy send) and ragged ones
def test_step_gather_gloo_world2(tmp_path, n_total):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, 100, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_peer_push_layout_matches_summary_views():
    """The gathered buffer written by the peer-push path is [parity][rank][summary bytes]; its
    slots are read with the same field views as a local summary buffer (host-side check of the
    offsets pc_step_fused derives from the five summary pointers)."""
    import torch
    from clockpunch_b200.engine import CatalogEngine
    n, wpr, world = 40, 4, 3
    nbytes = n * (8 + 4 * wpr + 12)
    gathered = torch.zeros((2, world, nbytes), dtype=torch.uint8)
    base = gathered[1, 2]
    v = CatalogEngine.summary_views(base, n, wpr)
    offs = {k: t.data_ptr() - base.data_ptr() for k, t in v.items()}
    assert offs == {"num_tasked": 0, "vis_est": 8 * n, "tr_pos": 8 * n + 4 * wpr * n,
                    "tr_vel": 8 * n + 4 * wpr * n + 4 * n, "last_time_tasked": 8 * n + 4 * wpr * n + 8 * n}
    assert sum(t.numel() * t.element_size() for t in v.values()) == nbytes       # no padding: slots compare bytewise
    v["vis_est"][7, 3] = 0x55
    assert gathered[1, 2, 8 * n + 4 * (7 * wpr + 3)].item() == 0x55
