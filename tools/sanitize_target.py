"""Small fixed workload for compute-sanitizer (memcheck / racecheck): every kernel of the engine once
on tiny inputs -- Monte-Carlo path in all three precisions, replay, long (multi-tile) tracks, all-pairs."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

eng = capi.Engine(0)
tracks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 3)
ring = scenario.load_ring(pk.RING_1KM)[40:52]
for prec in (capi.FP64, capi.MIXED, capi.FP32):
    r = eng.simulate(tracks, ring, seed=1, trial_count=3000, precision=prec)
    print("prec", prec, "hits", r.hits, flush=True)
# the slab-first screen with everything it can do in one launch: ragged items, a full survivor queue (re-scan rounds),
# level 2, the in-kernel fp64 decision and the warp-aggregated record append (a 60 m "aircraft": many hits), the
# direction choice (second run of >= 20 000 trials per pair on the scenario), and the 2-D screen / culled variants
m = capi.default_model()
m.aircraft_radius = 60.0
for flags in (0, 0, capi.FLAG_SCREEN_2D, capi.FLAG_REACH_CULL):
    r = eng.simulate(tracks[:2], ring[:3], seed=5, trial_count=20011, precision=capi.MIXED, flags=flags, model=m,
                     record_capacity=1 << 16)
    print("dense flags", flags, "hits", r.hits, "survivors", r.stats["slab_survivors"], "records", len(r.records), flush=True)
uid = capi.Engine.comm_unique_id()
eng.attach_comm(uid, 0, 1)
eng.allreduce_counts()
eng.detach_comm()
eng.upload(tracks[:1], ring[:2])
draws = eng.dump_draws(0, 1, 1, 0, 2500)
rep = np.concatenate([eng.dump_draws(0, 0, 1, 0, 2500), draws])
r = eng.simulate(tracks[:1], ring[:2], seed=0, trial_count=2500, precision=capi.MIXED, replay=rep)
print("replay hits", r.hits, flush=True)
t = tracks[0]
n = 1100
k = np.arange(1, n - len(t["x"]) + 1)
long_t = dict(t, x=np.concatenate([t["x"], t["x"][-1] - 80.0 * k]), y=np.concatenate([t["y"], t["y"][-1] + 0 * k]),
              z=np.concatenate([t["z"], np.minimum(t["z"][-1] + 9.0 * k, 2500.0)]))
r = eng.simulate([long_t, tracks[1]], ring, seed=2, trial_count=2100, precision=capi.MIXED)
print("long hits", r.hits, flush=True)
eng.upload([long_t, tracks[1], tracks[2]], ring)
r = eng.allpairs(3, 0, 700, 30, capi.target_box(t["y"][0], 0.0))
print("allpairs hits", r.hits, "level2", r.stats["level2"], flush=True)
if len(r.records):
    print(eng.trajectories(r.records[:2]).shape)
eng.close()
print("done")
