"""CPU, world_size 2, gloo: the N>1 host logic of the unit-sharded path (no data-path collective)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ait_b200 import synth
    from ait_b200.sharding import gather_results, shard_units
    mine = shard_units(n_units, rank, world)
    # stand-in for the per-unit result of the head: a deterministic function of the unit's inputs
    local = [(u, float(synth.query_feat(u, channels=8).sum())) for u in mine]
    # the max-over-ranks timing reduction of bench.py
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = gather_results(local, world)
    if rank == 0:
        out_q.put((gathered, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    from ait_b200 import synth
    world, n_units = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [u for u, _ in gathered] == list(range(n_units))                # unit order preserved
    expect = [float(synth.query_feat(u, channels=8).sum()) for u in range(n_units)]
    assert [v for _, v in gathered] == expect
    assert tmax == 11.0
