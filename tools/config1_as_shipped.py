"""BASELINE configs[0] -- "reference run as shipped": the first departure of the day x the 99 drones of the 1 km ring x
10 000 trials each (Monte_Carlo.cpp:11-18,28,33), the one workload the reference runs on its own.  Timed three ways
on one box, on the same aircraft CSV (the synthetic stand-in for the undistributed ADS-B file):

  reference   oracle/_ref/ref_asshipped, ONE process: the unmodified Aircraft.cpp + Drone.cpp + the loop of
              Monte_Carlo.cpp as shipped (std::random_device per draw), seconds measured inside, after the CSV load
  drop-in     drone-collision-rp_b200/Monte_Carlo: the same main() on the new classes -- 99 Drone::Simulation calls,
              each a dcrp_simulate of one (aircraft, drone) pair; process wall time (CUDA context creation included)
              and, with a warm engine in this process, the loop of 99 dcrp_simulate calls alone
  batch       Monte_Carlo --batch: the 99 pairs in ONE launch

  python tools/config1_as_shipped.py        (one JSON line; oracle/ is executed here as the timed reference only)
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

RUNS = 10000


def main():
    import oracle_api as oa

    work = tempfile.mkdtemp()
    shutil.copy(pk.RING_1KM, os.path.join(work, "drone_inital_positions_1km.csv"))
    out = {"workload": "BASELINE configs[0]: aircraft 0 x 99 drones (1 km ring) x %d trials = %d trials" % (RUNS, 99 * RUNS)}

    def run_mc(extra):
        t0 = time.perf_counter()
        p = subprocess.run([pk.MONTE_CARLO, "--synthetic", "16"] + extra, cwd=work, capture_output=True, text=True, timeout=600)
        dt = time.perf_counter() - t0
        assert p.returncode == 0, p.stdout + p.stderr
        line = [l for l in p.stdout.splitlines() if l.startswith("Number of collisions")][-1]
        return dt, float(line.split(":")[1])

    run_mc([])                                         # writes the synthetic day, pages the binaries in
    wall, hits = min(run_mc([]) for _ in range(3))
    out["drop_in_process"] = {"seconds": wall, "collisions": hits, "trials_per_s": 99 * RUNS / wall,
                              "what": "./Monte_Carlo (99 Drone::Simulation calls), process wall, best of 3"}
    wall_b, hits_b = min(run_mc(["--batch"]) for _ in range(3))
    assert hits_b == hits
    out["batch_process"] = {"seconds": wall_b, "collisions": hits_b, "trials_per_s": 99 * RUNS / wall_b,
                            "what": "./Monte_Carlo --batch (one launch), process wall, best of 3"}

    # the loop alone, engine warm: what Drone::Simulation costs per (aircraft, drone) pair
    csv = os.path.join(work, [f for f in os.listdir(work) if f.startswith("TAKEOFF")][0])
    track = scenario.load_flight(csv, 0)
    ring = scenario.load_ring(pk.RING_1KM)
    eng = capi.Engine(0)
    calls = [eng.prepared_simulate(capi.Scenario([track], ring[j:j + 1]), precision=capi.MIXED, record_capacity=4096,
                                   flags=capi.FLAG_NO_TIMING) for j in range(99)]
    best, total = None, 0
    for rep in range(5):
        total = 0
        t0 = time.perf_counter()
        for j, call in enumerate(calls):
            counts, _, _ = call(0, RUNS)
            total += int(counts[0])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out["drop_in_loop_warm"] = {"seconds": best, "us_per_pair_call": 1e6 * best / 99, "trials_per_s": 99 * RUNS / best,
                                "what": "99 dcrp_simulate calls of one pair x %d trials (new scenario each call), engine "
                                        "warm, best of 5" % RUNS}

    if os.path.exists(oa.REF_ASSHIPPED):
        p = subprocess.run([oa.REF_ASSHIPPED, csv, "22", "0", pk.RING_1KM, "99", str(RUNS), "1", os.path.join(work, "ref")],
                           check=True, capture_output=True, text=True)
        j = json.loads(p.stdout.strip().split("\n")[-1])
        out["reference_as_shipped"] = {"seconds": j["seconds"], "collisions": j["hits"], "trials_per_s": j["trials"] / j["seconds"],
                                       "cores": 1, "what": "unmodified reference, one process (it is single-threaded), "
                                                           "seconds inside the binary after the CSV load"}
        out["speedup_process_wall"] = j["seconds"] / wall
        out["speedup_loop_warm"] = j["seconds"] / best
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
