"""Short, fixed command for ncu: BASELINE configs[1] (5 km ring x 1 track x 1e7
trials, mixed precision), 5 runs, then one 1e9-trial run.  `--fp64` profiles k_exact,
`--2d` the sorted 2-D screen (k_screen2), `--cull` the reach-culled variant, `--1km` the 1 km ring."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

prec = capi.FP64 if "--fp64" in sys.argv else capi.MIXED
flags = capi.FLAG_REACH_CULL if "--cull" in sys.argv else 0
flags |= capi.FLAG_SCREEN_2D if "--2d" in sys.argv else 0
eng = capi.Engine(0)
tracks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 1)
ring = scenario.load_ring(pk.RING_1KM if "--1km" in sys.argv else pk.RING_5KM)
eng.upload(tracks, ring)
sizes = [101011] * 5 + ([] if "--short" in sys.argv else [10_000_000])
for i, n in enumerate(sizes):
    run = eng.run_async(seed=20170630, trial_begin=i * 101011, trial_count=n, precision=prec, flags=flags)
    r = eng.fetch(run)
    print(i, "trials %.3e ms_trials %.4f  tests/s %.4e  TFLOP11 %.2f  issued/eval %.4f hits %d cand %d level2 %d slab_surv %d" % (
        r.stats["trials"], r.stats["ms_trials"], r.stats["tests_evaluated"] / (r.stats["ms_trials"] * 1e-3),
        r.stats["tests_evaluated"] * 11 / (r.stats["ms_trials"] * 1e-3) * 1e-12,
        r.stats["tests_issued"] / max(1, r.stats["tests_evaluated"]), r.stats["hits"],
        r.stats["candidates"], r.stats["level2"], r.stats["slab_survivors"]), flush=True)
eng.close()
