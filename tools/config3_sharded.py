"""BASELINE config 3 at full size: 1.0e10 trials (1 km and 5 km rings, one track), sharded over the ranks
of a torchrun launch as contiguous Philox trial ranges, counts all-reduced (NCCL).  Rank 0 then runs the
whole range alone and checks that the sharded counts are identical (results never depend on the sharding).

  python -m torch.distributed.run --nproc-per-node 8 ... tools/config3_sharded.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

N_PER_DRONE = 101_010_102            # x 99 drones = 1.0000000098e10 trials


def main():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = capi.Engine(local)
    tracks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 1)
    out = {"workload": "BASELINE config 3: %d trials x 99 drones = %.4e trials per ring, one track" % (
        N_PER_DRONE, 99.0 * N_PER_DRONE), "n_gpus": world, "rings": {}}
    for name, path in (("1km", pk.RING_1KM), ("5km", pk.RING_5KM)):
        ring = scenario.load_ring(path)
        begin, count = scenario.partition_trials(0, N_PER_DRONE, rank, world)
        eng.simulate(tracks, ring, seed=20170630, trial_begin=begin, trial_count=1000, precision=capi.MIXED)   # warm-up
        res = eng.simulate(tracks, ring, seed=20170630, trial_begin=begin, trial_count=count, precision=capi.MIXED,
                           record_capacity=1 << 18)
        counts = torch.from_numpy(res.counts.astype(np.int64)).cuda()
        agg = torch.tensor([res.stats["ms_device_total"], float(res.stats["tests_evaluated"]),
                            float(res.stats["trials"])], dtype=torch.float64, device="cuda")
        mx, sm = agg.clone(), agg.clone()
        if world > 1:
            dist.all_reduce(counts)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        if rank == 0:
            full = eng.simulate(tracks, ring, seed=20170630, trial_count=N_PER_DRONE, precision=capi.MIXED,
                                record_capacity=1 << 18)
            same = bool(np.array_equal(full.counts.astype(np.int64), counts.cpu().numpy()))
            ms = mx[0].item()
            out["rings"][name] = {
                "trials": sm[2].item(), "hits": int(counts.sum().item()), "p": counts.sum().item() / sm[2].item(),
                "ms_sharded_max_over_ranks": ms, "tests_per_s": sm[1].item() / (ms * 1e-3),
                "trials_per_s": sm[2].item() / (ms * 1e-3), "ms_single_gpu": full.stats["ms_device_total"],
                "speedup_vs_single_gpu": full.stats["ms_device_total"] / ms,
                "counts_identical_to_single_gpu": same}
            assert same
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
