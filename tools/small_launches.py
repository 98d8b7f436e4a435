"""k_screen3 launches of chosen sizes (trials per drone on the 5 km ring), 3 per size, ms_trials printed: the
command ncu wraps to compare a short launch with a long one ('what bounds what is left')."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

sizes = [int(a) for a in sys.argv[1:]] or [1, 256, 2560, 10000, 25600, 101011]
eng = capi.Engine(0)
tracks = [scenario.load_flight(os.path.join(pk.DATA_DIR, "synthetic_day_1_20170630.csv"), 0)]
ring = scenario.load_ring(pk.RING_5KM)
eng.upload(tracks, ring)
eng.fetch(eng.run_async(seed=1, trial_count=101011, precision=capi.MIXED))
for n in sizes:
    for k in range(3):
        r = eng.fetch(eng.run_async(seed=1, trial_begin=k * n, trial_count=n, precision=capi.MIXED))
    print(n, r.stats["ms_trials"], r.stats["ms_device_total"])
