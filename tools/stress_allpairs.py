"""BASELINE config 5 (synthetic stress) at a chosen scale: N synthetic tracks of S samples x
M rays per ring drone, every ray against every track (dcrp_allpairs), collision-record compaction on.
Under torchrun the rays are sharded over the ranks (disjoint ray ranges), tracks replicated, and the
per-(track, drone) counts all-reduced.

  python tools/stress_allpairs.py --tracks 1000 --steps 10000 --rays-per-drone 1000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario


def synthetic_long_tracks(n_tracks, n_steps, seed):
    """Smooth random wandering within ~8 km of the airport, 27L band at index 0, altitude 0 on the
    ground roll then up to 3 km: geometry only matters for the (tiny) level-2 rate."""
    rng = np.random.default_rng(seed)
    out = []
    t = np.arange(n_steps, dtype=np.float64)
    for k in range(n_tracks):
        ph = rng.uniform(0, 2 * np.pi, 4)
        w = rng.uniform(2e-4, 2e-3, 4)
        x = 675000.0 + 5000.0 * np.sin(w[0] * t + ph[0]) + 2500.0 * np.sin(3 * w[1] * t + ph[1])
        y = 5705300.0 + 5000.0 * np.sin(w[2] * t + ph[2]) + 2500.0 * np.cos(3 * w[3] * t + ph[3])
        y[0] = 5704545.0 + rng.uniform(-5, 5)
        z = np.clip(9.0 * (t - 30.0), 0.0, 1500.0 + 1500.0 * rng.random())
        out.append({"x": x, "y": y, "z": z, "takeoff_t": 30, "n_steps": n_steps})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=200)
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--rays-per-drone", type=int, default=1000)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = capi.Engine(local)
    tracks = synthetic_long_tracks(args.tracks, args.steps, 1234)
    ring = scenario.load_ring(pk.RING_1KM)
    box = capi.target_box(tracks[0]["y"][0], 0.0)
    t0 = time.time()
    eng.upload(tracks, ring)
    begin, count = scenario.partition_trials(0, args.rays_per_drone, rank, world)
    best = None
    for rep in range(args.repeat):
        res = eng.allpairs(20170630, begin, count, 30, box, record_capacity=1 << 16)
        if best is None or res.stats["ms_trials"] < best.stats["ms_trials"]:
            best = res
    st = best.stats
    counts = best.counts.astype(np.int64)
    ms = st["ms_trials"]
    tests = st["tests_evaluated"]
    if world > 1:
        import torch
        import torch.distributed as dist

        c = torch.from_numpy(counts).cuda()
        dist.all_reduce(c)
        agg = torch.tensor([ms, float(tests), float(st["hits"])], dtype=torch.float64, device="cuda")
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, tests, hits = mx[0].item(), sm[1].item(), sm[2].item()
        counts = c.cpu().numpy()
    else:
        hits = st["hits"]
    if rank == 0:
        line = {"workload": "all-pairs stress: %d tracks x %d steps x %d rays (99 drones x %d)" % (
                    args.tracks, args.steps, 99 * args.rays_per_drone, args.rays_per_drone),
                "n_gpus": world, "tests_evaluated": tests, "scan_ms": ms,
                "tests_per_s": tests / (ms * 1e-3), "tflops_11": tests * 11 / (ms * 1e-3) * 1e-12,
                "issued_over_evaluated": st["tests_issued"] / max(1, st["tests_evaluated"]),
                "level2": st["level2"], "candidates": st["candidates"], "hits": hits,
                "records": int(len(best.records)), "resolve_ms": st["ms_confirm"],
                "track_bytes_fp32": args.tracks * args.steps * 16, "upload_s": None,
                "pairs_with_hits": int((counts > 0).sum())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
