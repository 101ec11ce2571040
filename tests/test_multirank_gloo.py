"""world_size-2 gloo test of the N>1 host logic (env sharding + the single throughput all-gather)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from mdt_policy_b200 import dist as D
    r, lr, w = D.init_from_env("gloo")
    lo, hi = D.shard_range(2 * 256 + 1, r, w)
    D.barrier()
    agg = D.aggregate_throughput(local_units=(hi - lo) * 10, local_seconds=1.0 + r, device="cpu")
    q.put((r, lo, hi, agg))
    torch.distributed.destroy_process_group()


def test_two_rank_sharding_and_counter_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    (r0, lo0, hi0, a0), (r1, lo1, hi1, a1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 257, 257, 513)
    assert a0["units"] == a1["units"] == 5130.0
    assert a0["seconds"] == 2.0 and abs(a0["throughput"] - 2565.0) < 1e-9
    assert a0["per_rank"] == [(2570.0, 1.0), (2560.0, 2.0)]


def test_single_process_is_a_noop():
    from mdt_policy_b200 import dist as D
    assert D.shard_range(10, 0, 1) == (0, 10)
    assert [D.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert D.aggregate_throughput(100, 2.0)["throughput"] == 50.0
