"""Where the fixed cost of a short k_screen3 launch sits: per-warp %globaltimer stamps of one launch.

Needs the instrumented library (not part of the product build):
  nvcc -DDCRP_TIMELINE ... -o drone-collision-rp_b200/libdcrp_timeline.so drone-collision-rp_b200/csrc/engine.cu
  python tools/warp_timeline.py [trials_per_drone]
Prints the ramp (entry and first-staging times of the warps), the tail (when warps leave the item loop and end,
relative to the last warp) and the warp-time lost to both.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

capi.LIBDCRP = os.path.join(pk.PKG_DIR, "libdcrp_timeline.so")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 101011
eng = capi.Engine(0)
stream = torch.cuda.Stream(torch.device("cuda", 0))      # flush and run on ONE stream (ordered), as bench.py does
sh = stream.cuda_stream
tracks = [scenario.load_flight(os.path.join(pk.DATA_DIR, "synthetic_day_1_20170630.csv"), 0)]
ring = scenario.load_ring(pk.RING_5KM)
eng.upload(tracks, ring)
L = capi.lib()
L.dcrp_debug_timeline.argtypes = [C.c_void_p, C.c_int]
SLOTS, WARPS = 8, 8192
buf = np.zeros(WARPS * SLOTS, dtype=np.uint64)


def pct(a, qs=(0, 1, 10, 50, 90, 99, 100)):
    return "  ".join("p%d %.1f" % (q, np.percentile(a, q)) for q in qs)


for flush in (False, True):
    for k in range(6):
        if flush:
            capi.flush_l2(0, sh)
        buf[:] = 0
        r = eng.fetch(eng.run_async(seed=1, trial_begin=k * n, trial_count=n, precision=capi.MIXED, stream=sh))
    assert L.dcrp_debug_timeline(buf.ctypes.data_as(C.c_void_p), buf.size) == 0
    t = buf.reshape(WARPS, SLOTS).astype(np.int64)
    t = t[t[:, 0] > 0]
    w = len(t)
    t0 = t[:, 0].min()
    end = t[:, 4].max()
    us = lambda x: (x - t0) * 1e-3
    dur = (end - t0) * 1e-3
    print("---- %d trials/drone, L2 %s: %d warps, first entry -> last end %.1f us, events say %.1f us" % (
        n, "flushed" if flush else "warm", w, dur, 1e3 * r.stats["ms_device_total"]))
    print("entry                :", pct(us(t[:, 0])))
    print("first staging done   :", pct(us(t[:, 1])))
    print("first item done      :", pct(us(t[:, 2])))
    print("left the item loop   :", pct(us(t[:, 3])), " (before the end: " + pct((end - t[:, 3]) * 1e-3, (50, 90, 99)) + ")")
    print("warp end             :", pct(us(t[:, 4])))
    print("items per warp       :", pct(t[:, 5], (0, 10, 50, 90, 100)))
    lost_tail = float(np.sum(end - t[:, 4])) * 1e-3 / w
    lost_ramp = float(np.sum(t[:, 1] - t0)) * 1e-3 / w
    print("mean warp-time idle at the tail %.2f us, before the first Philox %.2f us (of %.1f)" % (lost_tail, lost_ramp, dur))
    # per SM sub-partition (warp slot mod 4 of a CTA's 8 warps -> its 4 schedulers): how evenly the tail ends
    gw = np.nonzero(buf.reshape(WARPS, SLOTS)[:, 0] > 0)[0]
    smsp = t[:, 6] * 4 + (gw % 8) % 4
    ends, items, lastlen = [], [], []
    for q in np.unique(smsp):
        m = smsp == q
        ends.append(us(t[m, 4].max()))
        items.append(t[m, 5].sum())
    print("per SM sub-partition : last warp ends at", pct(np.array(ends), (0, 10, 50, 90, 100)), "| items", pct(np.array(items), (0, 10, 50, 90, 100)))
    print("last item of a warp  : duration", pct((t[:, 3] - t[:, 7]) * 1e-3, (0, 10, 50, 90, 100)), "| started", pct(us(t[:, 7]), (0, 10, 50, 90, 100)))
    print("final re-scan        : duration", pct((t[:, 4] - t[:, 3]) * 1e-3, (0, 10, 50, 90, 100)))
    d = np.diff(np.unique(t[:, :5].ravel()))
    print("globaltimer: smallest step between distinct stamps %d ns" % (d[d > 0].min() if len(d) else -1))
eng.close()
