"""BASELINE config 4 at full size: every departure of a (synthetic) day in ONE scenario x 1e9 trials,
fp64 vs mixed vs fp32 mode on one GPU.

  python tools/day_batch.py [--flights 600]
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flights", type=int, default=600)
    ap.add_argument("--trials", type=float, default=1e9)
    args = ap.parse_args()
    eng = capi.Engine(0)
    day, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), args.flights, seed=4)
    ring = scenario.load_ring(pk.RING_1KM)
    n_pairs = len(day) * len(ring)
    n = int(np.ceil(args.trials / n_pairs))
    out = {"workload": "BASELINE config 4: %d departures x 99 drones (%d pairs, one launch) x %d trials/pair = %.3e trials" % (
        len(day), n_pairs, n, float(n_pairs) * n), "modes": {}}
    scn = capi.Scenario(day, ring)
    res = {}
    for name, prec in (("fp64", capi.FP64), ("mixed", capi.MIXED), ("mixed, 2-D screen (round-1 kernel)", capi.MIXED),
                       ("fp32", capi.FP32), ("mixed+reach_cull", capi.MIXED)):
        flags = capi.FLAG_REACH_CULL if name.endswith("cull") else (capi.FLAG_SCREEN_2D if "2-D" in name else 0)
        eng.simulate(scn, seed=4, trial_count=64, precision=prec, flags=flags)                  # warm-up
        r = eng.simulate(scn, seed=4, trial_count=n, precision=prec, flags=flags, record_capacity=1 << 20)
        res[name] = r
        st = r.stats
        out["modes"][name] = {"hits": int(r.hits), "ms_device_total": st["ms_device_total"],
                              "tests_evaluated": st["tests_evaluated"], "tests_culled": st["tests_culled"],
                              "tests_per_s": (st["tests_evaluated"] + st["tests_culled"]) / (st["ms_device_total"] * 1e-3),
                              "trials_per_s": st["trials"] / (st["ms_device_total"] * 1e-3),
                              "candidates": st["candidates"], "level2": st["level2"], "slab_survivors": st["slab_survivors"],
                              "ms_trials": st["ms_trials"]}
    a, b, c, d = res["fp64"], res["mixed"], res["fp32"], res["mixed+reach_cull"]
    b2 = res["mixed, 2-D screen (round-1 kernel)"]
    assert np.array_equal(a.counts, b2.counts)
    out["mixed_equals_fp64"] = bool(np.array_equal(a.counts, b.counts) and np.array_equal(
        a.records[["track", "drone", "trial", "first_hit"]], b.records[["track", "drone", "trial", "first_hit"]]))
    out["cull_equals_fp64"] = bool(np.array_equal(a.counts, d.counts) and np.array_equal(
        a.records[["track", "drone", "trial", "first_hit"]], d.records[["track", "drone", "trial", "first_hit"]]))
    out["fp32_hit_difference"] = int(c.hits) - int(a.hits)
    out["fp32_pairs_with_different_count"] = int((a.counts != c.counts).sum())
    assert out["mixed_equals_fp64"] and out["cull_equals_fp64"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
