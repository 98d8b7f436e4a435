"""BASELINE configs[1] at FULL size as a parity case: the 5 km ring x the bench's synthetic departure x
101 011 trials per drone = 1.0e7 trials.  Shared by the fixture generator (tests/golden/make_config2_fixture.py,
runs the UNMODIFIED reference on CPU), the GPU test (tests/test_gpu_config2_full.py) and bench.py's warm-up
guard.  Test infrastructure."""
import hashlib
import json
import multiprocessing as mp
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "ref_config2_5km.json")
SEED = 20170630
TRIALS_PER_DRONE = 101011
N_DRONES = 99


def track_digest(t):
    h = hashlib.sha256()
    for k in "xyz":
        h.update(np.ascontiguousarray(t[k], dtype=np.float64).tobytes())
    h.update(np.int32([t["takeoff_t"], t["n_steps"]]).tobytes())
    return h.hexdigest()


def per_trial_digest(hit, first_hit):
    """One hash over the per-trial flags and first-hit indices of a drone (bit-exact quantities)."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(hit, dtype=np.uint8).tobytes())
    h.update(np.ascontiguousarray(first_hit, dtype=np.int32).tobytes())
    return h.hexdigest()


def load_fixture():
    with open(FIXTURE) as f:
        return json.load(f)


# ---- replay of one drone's table through a CPU checker, in forked workers (the workers never touch CUDA)
_G = {}


def _worker(j):
    import oracle_api as oa
    import drone_collision_rp_b200 as pk

    t, ring, draws = _G["track"], _G["ring"], _G["draws"]
    if _G["use_ref"]:
        r = oa.RefReplay().run(pk.RING_5KM, j, 0, t, draws[j])
        assert r["range_violations"] == 0
        return j, r["hit"], r["first_hit"], r["min_dist"]
    hit, first, mind = oa.Oracle().run_trials(t, ring[j], draws[j])
    return j, hit, first, mind


def replay_all(track, ring, draws, use_ref, nproc=None):
    """draws: list of per-drone DRAW_DTYPE arrays.  Returns (hit, first_hit, min_dist) arrays [n_drones, n]."""
    _G.update(track=track, ring=ring, draws=draws, use_ref=use_ref)
    n_dr, n = len(draws), len(draws[0])
    hit = np.zeros((n_dr, n), np.uint8)
    first = np.zeros((n_dr, n), np.int32)
    mind = np.zeros((n_dr, n), np.float64)
    nproc = nproc or min(n_dr, os.cpu_count() or 1)
    with mp.get_context("fork").Pool(nproc) as pool:
        for j, h, f, m in pool.imap_unordered(_worker, range(n_dr)):
            hit[j], first[j], mind[j] = h, f, m
    return hit, first, mind
