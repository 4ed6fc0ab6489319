"""The N>1 host logic on CPU: 2 gloo ranks shard the batches round robin, run them (the oracle stands in for the GPU
here), gather by batch sequence and must reproduce the single-process output byte for byte."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_c
from excord_lr_b200 import shard, synth
from excord_lr_b200.batch import ExlrParams


def _runner(p):
    def run(hb):
        r = oracle_c.run(hb, p)
        assert r.status == 0
        return oracle_c.format_lines(hb, r.events)
    return run


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hb = synth.config(0, 0.3)
    p = ExlrParams.make(**synth.CONFIGS[0]["params"])
    plan = shard.plan_batches(hb.n_reads, 257)
    mine = shard.run_rank(hb, plan, rank, world, _runner(p))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    n = torch.tensor([len(mine)])
    dist.all_reduce(n)
    if rank == 0:
        q.put((shard.merge_ordered(gathered), int(n.item()), len(plan)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_round_robin_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    merged, n_batches, n_plan = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    hb = synth.config(0, 0.3)
    p = ExlrParams.make(**synth.CONFIGS[0]["params"])
    assert n_batches == n_plan and n_plan > 5
    assert merged == _runner(p)(hb)


def test_plan_and_merge_edge_cases():
    assert shard.plan_batches(0, 10) == []
    assert shard.plan_batches(10, 10) == [(0, 0, 10)]
    assert shard.plan_batches(11, 10) == [(0, 0, 10), (1, 10, 11)]
    plan = shard.plan_batches(95, 10)
    assert sorted(shard.rank_batches(plan, 0, 4) + shard.rank_batches(plan, 1, 4) + shard.rank_batches(plan, 2, 4) +
                  shard.rank_batches(plan, 3, 4)) == plan
    with pytest.raises(ValueError):
        shard.merge_ordered([[(0, b"a")], [(2, b"c")]])
