"""First-contact probe for a GPU box: exercises every kernel once, smallest
first, printing as it goes so that a fault is attributable; writes
gpurun_out/probe.json.  Not a test and not the bench."""
import json
import os
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario
import oracle_api as oa

out = {}


def stage(name):
    print("== " + name, flush=True)


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    stage("engine")
    eng = capi.Engine(0)
    out["device"] = eng.device_info()
    print(out["device"], flush=True)
    o = oa.Oracle()
    tmp = tempfile.mkdtemp()
    tracks, csv = scenario.synthetic_tracks(tmp, 6)
    ring1 = scenario.load_ring(pk.RING_1KM)
    ring5 = scenario.load_ring(pk.RING_5KM)

    stage("peaks")
    out["peaks_tflops"] = {}
    for which, name in ((0, "ffma"), (1, "ffma2"), (2, "dfma"), (3, "scan_mix_11flop")):
        tf, ms = capi.measure_peak(0, which)
        out["peaks_tflops"][name] = tf
        print(name, tf, "TFLOP/s", ms, "ms", flush=True)

    stage("draws")
    eng.upload(tracks[:2], ring1)
    d = eng.dump_draws(1, 57, 7, 12345, 1000)
    t = tracks[1]
    w = o.draws(7, 12345, 1000, 57, 1, t["takeoff_t"], capi.target_box(t["y"][0], t["z"][0]))
    out["draws_equal"] = bool(np.array_equal(d, w))
    print("draws equal oracle:", out["draws_equal"], flush=True)

    t = tracks[0]
    box = capi.target_box(t["y"][0], t["z"][0])
    n = 20000
    want = np.array([o.run_philox(t, box, ring1[j], 7, j, 0, 0, n)[0] for j in range(99)], dtype=np.uint64)
    for prec, name in ((capi.FP64, "fp64"), (capi.MIXED, "mixed"), (capi.FP32, "fp32")):
        stage("simulate " + name)
        r = eng.simulate(tracks[:1], ring1, seed=7, trial_count=n, precision=prec)
        ok = bool(np.array_equal(r.counts, want))
        out["small_" + name] = {"hits": r.hits, "oracle_hits": int(want.sum()), "counts_equal": ok, "stats": r.stats}
        print(name, "hits", r.hits, "oracle", int(want.sum()), "equal", ok, r.stats, flush=True)

    stage("throughput sweep (mixed, 1 track x 99 drones)")
    out["sweep"] = []
    eng.upload(tracks[:1], ring5)
    for per in (10000, 101011, 1000000, 10000000):
        for rep in range(2):
            run = eng.run_async(seed=1, trial_count=per, precision=capi.MIXED)
            r = eng.fetch(run)
        ms = r.stats["ms_trials"]
        row = {"trials": r.stats["trials"], "ms_trials": ms, "ms_confirm": r.stats["ms_confirm"],
               "ms_total": r.stats["ms_device_total"],
               "tests_per_s": r.stats["tests_evaluated"] / (ms * 1e-3),
               "trials_per_s": r.stats["trials"] / (ms * 1e-3),
               "tflops_11": r.stats["tests_evaluated"] * 11 / (ms * 1e-3) * 1e-12,
               "issued_over_eval": r.stats["tests_issued"] / max(1, r.stats["tests_evaluated"]),
               "cand": r.stats["candidates"], "hits": r.stats["hits"]}
        out["sweep"].append(row)
        print(row, flush=True)
    stage("throughput fp64 / fp32 at 1e6 per drone")
    for prec, name in ((capi.FP64, "fp64"), (capi.FP32, "fp32")):
        run = eng.run_async(seed=1, trial_count=1000000, precision=prec)
        r = eng.fetch(run)
        ms = r.stats["ms_trials"]
        row = {"mode": name, "trials": r.stats["trials"], "ms_trials": ms,
               "tests_per_s": r.stats["tests_evaluated"] / (ms * 1e-3), "hits": r.stats["hits"]}
        out["sweep_" + name] = row
        print(row, flush=True)

    stage("multi-track batch (6 tracks x 99 drones x 100k, mixed vs fp64)")
    a = eng.simulate(tracks, ring1, seed=3, trial_count=100000, precision=capi.MIXED)
    b = eng.simulate(tracks, ring1, seed=3, trial_count=100000, precision=capi.FP64)
    out["batch"] = {"mixed_hits": a.hits, "fp64_hits": b.hits, "equal": bool(np.array_equal(a.counts, b.counts)),
                    "ms_mixed": a.stats["ms_trials"], "ms_fp64": b.stats["ms_trials"]}
    print(out["batch"], flush=True)

    stage("trajectories")
    xyz = eng.trajectories(a.records[:4])
    print(xyz[0, :, :3], flush=True)
    out["ok"] = True


try:
    main()
except Exception:
    traceback.print_exc()
    out["error"] = traceback.format_exc()
finally:
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1, default=str)
