"""ms_trials (the trial kernel alone, CUDA events) of the mixed path against the trial count: fixed cost vs slope."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

flags = capi.FLAG_SCREEN_2D if "--2d" in sys.argv else 0
eng = capi.Engine(0)
tracks = [scenario.load_flight(os.path.join(pk.DATA_DIR, "synthetic_day_1_20170630.csv"), 0)]
ring = scenario.load_ring(pk.RING_1KM if "--ring1" in sys.argv else pk.RING_5KM)
eng.upload(tracks, ring)
eng.fetch(eng.run_async(seed=1, trial_count=101011, precision=capi.MIXED, flags=flags))
for n in ((101011, 1010110) if "--short" in sys.argv else (1, 256, 2560, 10000, 25600, 50000, 101011, 202022, 404044, 1010110)):
    best, tot = 1e9, 1e9
    for k in range(7):
        r = eng.fetch(eng.run_async(seed=1, trial_begin=k * n, trial_count=n, precision=capi.MIXED, flags=flags))
        best = min(best, r.stats["ms_trials"]); tot = min(tot, r.stats["ms_device_total"])
    print("n/drone %8d  trials %.3e  kernel %.4f ms  (device total %.4f)  per 1e7 trials %.4f ms" % (
        n, 99.0 * n, best, tot, best * 1e7 / (99.0 * n)), flush=True)
