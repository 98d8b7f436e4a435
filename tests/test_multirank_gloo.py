"""CPU tier, world_size 2 over gloo: the sharding rule bench.py and a multi-GPU host
use (disjoint Philox trial ranges per rank + one all-reduce of the per-pair counts,
SURVEY.md 8(e)) gives exactly the single-rank answer.  The per-rank compute here is
the oracle (this is a test of the plumbing, not of the kernels)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tempfile

    import drone_collision_rp_b200 as pk
    import oracle_api as oa
    from drone_collision_rp_b200 import capi, scenario

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oa.Oracle()
    with tempfile.TemporaryDirectory() as tmp:
        tracks, _ = scenario.synthetic_tracks(tmp, 1)
    ring = scenario.load_ring(pk.RING_1KM)
    t = tracks[0]
    box = capi.target_box(t["y"][0], t["z"][0])
    drones = [41, 42, 43, 44, 53]
    total_trials, begin0 = 150001, 17
    begin, count = scenario.partition_trials(begin0, total_trials, rank, world)
    counts = torch.zeros(len(drones), dtype=torch.int64)
    ids = []
    for k, j in enumerate(drones):
        h, hit_ids = o.run_philox(t, box, ring[j], 5, j, 0, begin, count)
        counts[k] = h
        ids += [(j, int(i)) for i in hit_ids]
    dist.all_reduce(counts)                       # the one exchange step of the path
    gathered = [None] * world
    dist.all_gather_object(gathered, ids)
    if rank == 0:
        q.put((counts.tolist(), sorted(sum(gathered, []))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(oracle, tracks, ring1):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, ids = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from drone_collision_rp_b200 import capi

    t = tracks[0]
    box = capi.target_box(t["y"][0], t["z"][0])
    want_counts, want_ids = [], []
    for j in [41, 42, 43, 44, 53]:
        h, hit_ids = oracle.run_philox(t, box, ring1[j], 5, j, 0, 17, 150001)
        want_counts.append(h)
        want_ids += [(j, int(i)) for i in hit_ids]
    assert counts == want_counts and ids == sorted(want_ids)
    assert sum(counts) > 0
