"""DCRP_FLAG_SUBSTEP (interpolated separation) on BASELINE configs[1]'s shape, both rings: the slab-first screen
(k_screen3<., SUB>) against the 2-D screen (k_screen2, DCRP_FLAG_SCREEN_2D) and the sample mode -- device time, how many
trials reach each level.  python tools/substep_speed.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

eng = capi.Engine(0)
tracks = [scenario.load_flight(os.path.join(pk.DATA_DIR, "synthetic_day_1_20170630.csv"), 0)]
n = 101011
for ring_name, ring_csv in (("5 km", pk.RING_5KM), ("1 km", pk.RING_1KM)):
    ring = scenario.load_ring(ring_csv)
    eng.upload(tracks, ring)
    for name, flags in (("sample mode, slab-first", 0), ("sub-step, slab-first", capi.FLAG_SUBSTEP),
                        ("sub-step, 2-D screen", capi.FLAG_SUBSTEP | capi.FLAG_SCREEN_2D)):
        best, r = 1e9, None
        for k in range(8):
            r = eng.fetch(eng.run_async(seed=1, trial_begin=k * n, trial_count=n, precision=capi.MIXED, flags=flags))
            best = min(best, r.stats["ms_device_total"])
        st = r.stats
        print("%s ring  %-26s %.4f ms  %.3e tests/s  survivors %.4f  level2 %.2e  candidates %.2e  hits %d" % (
            ring_name, name, best, st["tests_evaluated"] / (best * 1e-3), st["slab_survivors"] / st["trials"],
            st["level2"] / st["trials"], st["candidates"] / st["trials"], st["hits"]), flush=True)
