"""Multi-GPU host logic on the CPU: world_size-2 gloo processes exercise the sharding arithmetic, the
statistics all-reduce and the one-message actor broadcast of ddpg_trucktrailer_b200.dist."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from ddpg_trucktrailer_b200 import dist as ttd
    from ddpg_trucktrailer_b200.agent import init_actor_state_dict
    r, w, _ = ttd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    off, cnt = ttd.shard(1001, r, w)
    stats = torch.zeros(16, dtype=torch.float64)
    stats[0] = cnt; stats[1] = 1 + r; stats[3] = 10.0 * (r + 1); stats[4] = 100.0 * (r + 1)
    ttd.all_reduce_stats(stats)
    sd = init_actor_state_dict(seed=100 + r)          # different weights per rank before the broadcast
    sd = ttd.broadcast_actor(sd, src=0)
    q.put((rank, off, cnt, stats.tolist(), float(ttd.flatten_actor(sd).double().sum()), ttd.flatten_actor(sd).numel()))
    dist.destroy_process_group()


def test_world2_gloo_shard_stats_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(30) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    (r0, off0, cnt0, st0, sum0, n0), (r1, off1, cnt1, st1, sum1, n1) = res
    assert (off0, cnt0, off1, cnt1) == (0, 501, 501, 500)            # contiguous global env ranges
    assert st0 == st1 and st0[0] == 1001 and st0[1] == 3 and st0[3] == 30.0 and st0[4] == 300.0
    assert sum0 == sum1 and n0 == n1 == 131601                        # rank 0's weights everywhere, one message
    from ddpg_trucktrailer_b200.agent import init_actor_state_dict
    from ddpg_trucktrailer_b200 import dist as ttd
    assert sum0 == float(ttd.flatten_actor(init_actor_state_dict(seed=100)).double().sum())


def _worker_overlap(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from ddpg_trucktrailer_b200 import dist as ttd
    ttd.init_from_env("gloo")
    sync = ttd.OverlappedSync("cpu", src=0)
    got = []
    for it in range(5):                                    # the consumer sees the all-reduced statistics one iteration later
        st = torch.zeros(16, dtype=torch.float64); st[0] = 100 * (rank + 1) + it; st[1] = it
        prev = sync.push_stats(st)
        got.append(None if prev is None else prev.tolist())
    last = sync.flush_stats().tolist()
    flat = torch.full((131601,), float(rank + 1))          # rank 0 holds the learner's vector
    for it in range(3):
        if rank == 0:
            flat += 1.0                                    # "learner step"
        sync.push_policy(flat)
        sync.wait_policy()
    q.put((rank, got, last, float(flat[0]), float(flat[-1])))
    dist.destroy_process_group()


def test_world2_gloo_overlapped_sync():
    """OverlappedSync: the per-iteration statistics all-reduce is consumed one iteration behind; the policy broadcast moves
    the learner rank's flat parameter vector in place."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(30) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for rank, got, last, f0, f1 in res:
        assert got[0] is None
        for it in range(1, 5):
            assert got[it][0] == 300 + 2 * (it - 1) and got[it][1] == 2 * (it - 1)
        assert last[0] == 300 + 8 and last[1] == 8
        assert f0 == f1 == 4.0                              # rank 0's vector after three "learner steps", on both ranks


def test_shard_covers_everything():
    from ddpg_trucktrailer_b200 import dist as ttd
    for total in (1, 7, 8, 1 << 22, 12345):
        for world in (1, 2, 4, 8):
            parts = [ttd.shard(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def test_summarize():
    from ddpg_trucktrailer_b200 import dist as ttd
    s = torch.zeros(16, dtype=torch.float64)
    s[1], s[2], s[3], s[4] = 4, 1, 40.0, 600.0
    d = ttd.summarize(s)
    assert d["mean_return"] == 10.0 and abs(d["std_return"] - np.sqrt(50.0)) < 1e-12 and d["success_rate"] == 0.25
