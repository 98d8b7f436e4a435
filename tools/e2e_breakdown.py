"""Fixed cost of one dcrp_simulate call (host buffers in/out): wall time vs the device-side event times."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

# host-side stamps inside one dcrp_simulate call (instrumented build only: libdcrp_timeline.so, make EXTRA=-DDCRP_TIMELINE)
import ctypes as C
import numpy as np
tl = os.path.join(pk.PKG_DIR, "libdcrp_timeline.so")
if "--host" in sys.argv and os.path.exists(tl):
    capi.LIBDCRP = tl          # (the ONLY engine library of this process: its extern "C" calls bind to the first one loaded)
    tracks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 1)
    ring = scenario.load_ring(pk.RING_5KM)
    scn = capi.Scenario(tracks, ring)
    eng = capi.Engine(0)
    L = capi.lib()
    L.dcrp_debug_host_times.argtypes = [C.c_void_p, C.c_int]
    call = eng.prepared_simulate(scn, precision=capi.MIXED, record_capacity=4096, seed=1)
    names = ["entry", "resident check done", "run_async: before ev_begin", "ev_begin recorded", "k_run_init launched",
             "k_screen3 launched", "ev_end recorded", "fetch: before copy", "copy queued", "stream done", "fetch done"]
    for n in (1, 101011):
        rows, walls, dev = [], [], []
        for k in range(60):
            t0 = time.perf_counter()
            _, _, st = call(k * n, n)
            walls.append(time.perf_counter() - t0)
            buf = np.zeros(16)
            L.dcrp_debug_host_times(buf.ctypes.data_as(C.c_void_p), 16)
            rows.append(buf[:11].copy()); dev.append(st.ms_device_total)
        med = np.median(np.array(rows[10:]), axis=0) * 1e-3
        print("host stamps n=%d (us, medians; wall %.1f, device_total %.1f):" % (n, 1e6 * np.median(walls[10:]), 1e3 * np.median(dev[10:])))
        for nm, v in zip(names, med):
            print("   %-28s %8.2f" % (nm, v))
    eng.close()
    sys.exit(0)

eng = capi.Engine(0)
tracks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 1)
ring = scenario.load_ring(pk.RING_5KM)
scn = capi.Scenario(tracks, ring)
for flags, name in ((0, "plain"), (capi.FLAG_REACH_CULL, "cull")):
    for n in (1, 101011):
        walls, st = [], None
        for k in range(30):
            t0 = time.perf_counter()
            r = eng.simulate(scn, seed=1, trial_begin=k * n, trial_count=n, precision=capi.MIXED, record_capacity=4096,
                             flags=flags)
            walls.append(time.perf_counter() - t0)
            st = r.stats
        walls = sorted(walls[5:])
        print("%-5s n=%-7d wall median %.1f us  min %.1f us | prepare %.1f trials %.1f confirm %.1f device_total %.1f us" % (
            name, n, 1e6 * walls[len(walls) // 2], 1e6 * walls[0], 1e3 * st["ms_prepare"], 1e3 * st["ms_trials"],
            1e3 * st["ms_confirm"], 1e3 * st["ms_device_total"]))
eng.close()

# the same split into its two halves (Python-side clocks around the ctypes calls; scenario resident)
import statistics
eng = capi.Engine(0)
eng.upload(tracks, ring)
for n in (1, 101011):
    ta, tf = [], []
    for k in range(40):
        t0 = time.perf_counter()
        run = eng.run_async(seed=1, trial_begin=k * n, trial_count=n, precision=capi.MIXED)
        t1 = time.perf_counter()
        r = eng.fetch(run)
        t2 = time.perf_counter()
        ta.append(t1 - t0); tf.append(t2 - t1)
    print("split n=%-7d run_async median %.1f us   fetch (wait + copy + unpack) median %.1f us   device_total %.1f us" % (
        n, 1e6 * statistics.median(ta[5:]), 1e6 * statistics.median(tf[5:]), 1e3 * r.stats["ms_device_total"]))
eng.close()
