"""Writes tests/golden/ref_config2_5km.json: the answers of the UNMODIFIED reference (oracle/_ref, built from
/root/reference) for BASELINE configs[1] at full size -- 5 km ring x the bench's synthetic departure x 101 011
trials per drone (1.0e7 trials), draws = the Philox table of seed 20170630 (oracle/dcrp_oracle.c, the same
table dcrp_dump_draws produces on the GPU: tests/test_gpu_parity.py::test_draw_dump_is_oracle_philox).
Per drone: hit count, a SHA-256 over all per-trial flags and first-hit indices, min / sum of the per-trial
minimum distance, the four closest trials; plus every colliding (drone, trial, first_hit, min_dist).  Runs in the BUILD container only.

usage: python tests/golden/make_config2_fixture.py"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import config2_common as c2  # noqa: E402
import oracle_api as oa  # noqa: E402
import drone_collision_rp_b200 as pk  # noqa: E402
from drone_collision_rp_b200 import capi, scenario  # noqa: E402


def main():
    assert oa.RefReplay.available(), "build oracle/_ref first (make -C oracle)"
    o = oa.Oracle()
    with tempfile.TemporaryDirectory() as tmp:
        tracks, _ = scenario.synthetic_tracks(tmp, 1)
    t = tracks[0]
    ring = scenario.load_ring(pk.RING_5KM)
    box = capi.target_box(t["y"][0], t["z"][0])
    n = c2.TRIALS_PER_DRONE
    draws = [o.draws(c2.SEED, 0, n, j, 0, t["takeoff_t"], box) for j in range(c2.N_DRONES)]
    hit, first, mind = c2.replay_all(t, ring, draws, use_ref=True)
    hit_o, first_o, mind_o = c2.replay_all(t, ring, draws, use_ref=False)       # the C restatement agrees bit for bit
    assert np.array_equal(hit, hit_o) and np.array_equal(first, first_o) and np.array_equal(mind, mind_o)
    per_drone = []
    for j in range(c2.N_DRONES):
        per_drone.append({"hits": int(hit[j].sum()), "sha256_hit_first": c2.per_trial_digest(hit[j], first[j]),
                          "min_dist_min": float(mind[j].min()), "min_dist_sum": float(mind[j].sum()),
                          "closest": [[int(k), float(mind[j, k])] for k in np.argsort(mind[j], kind="stable")[:4]],
                          "draws_sha256": __import__("hashlib").sha256(draws[j].tobytes()).hexdigest()})
    jj, kk = np.nonzero(hit)
    out = {"what": "unmodified reference Drone.cpp (oracle/_ref, replay shim) on BASELINE configs[1] at full size",
           "seed": c2.SEED, "trials_per_drone": n, "n_drones": c2.N_DRONES, "ring": "drone_inital_positions_5km.csv",
           "track_sha256": c2.track_digest(t), "track_n_steps": int(t["n_steps"]), "track_takeoff_t": int(t["takeoff_t"]),
           "total_hits": int(hit.sum()),
           "collisions": [{"drone": int(j), "trial": int(k), "first_hit": int(first[j, k]), "min_dist": float(mind[j, k])}
                          for j, k in zip(jj, kk)],
           "per_drone": per_drone}
    with open(c2.FIXTURE, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", c2.FIXTURE, "hits", out["total_hits"], "closest", min(p["min_dist_min"] for p in per_drone))


if __name__ == "__main__":
    main()
