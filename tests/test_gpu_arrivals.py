"""GPU tier, SURVEY 8(f4): ARRIVAL flows -- the depart_or_arrive = 1 branch of the reference
(/root/reference/Drone.cpp:47-52 picks it from altitude[0] != 0; :138-141 and :152-155 put the target box at
[676345.60, 676345.60 + 5000] east of the threshold).  Synthetic arrivals on both runways
(host/synth_day.cpp: 3-degree approach from the east, touchdown, roll-out), drones on a small ring around
the point where the approach passes ~110 m (the shipped rings sit 4 km west of it and could never collide:
a drone starts at 100 m and only climbs).  Bars as in test_gpu_parity.py: per-trial flags and first-hit
indices bit-exact against the UNMODIFIED reference on replayed draws, min distance <= 1e-9 m; mixed ==
fp64 record for record."""
import os

import numpy as np
import pytest

from drone_collision_rp_b200 import capi, scenario

pytestmark = pytest.mark.gpu

SEED = 20170630
POS_TOL = 1e-9
N_RING = 12
R_AC = 12.0        # a wider aircraft so that a few 1e4 trials per pair hold enough conflicts to compare


@pytest.fixture(scope="module")
def arrivals(tmp_path_factory):
    d = tmp_path_factory.mktemp("arrivals")
    tracks, _ = scenario.synthetic_tracks(str(d), 6, arrival_every=3)     # flights 2 (27L) and 5 (27R) arrive
    out = []
    for a in (2, 5):
        t = tracks[a]
        assert t["z"][0] != 0.0
        istar = t["takeoff_t"] - 30
        cx, cy = t["x"][istar], t["y"][istar]
        path = os.path.join(str(d), "ring_arrival_%d.csv" % a)
        with open(path, "w", newline="") as f:           # the format of drone_inital_positions_*.csv (CRLF, header)
            f.write(",longitude,latitude,heading\r\n")
            for k in range(N_RING + 1):
                b = 2 * np.pi * (k % N_RING) / N_RING
                f.write("%d,%.10f,%.10f,%.15f\r\n" % (k, cx + 350 * np.sin(b), cy + 350 * np.cos(b),
                                                      (b + np.pi) % (2 * np.pi)))
        out.append((a, t, path, scenario.load_ring(path, n=N_RING)))
    return out


def _model():
    m = capi.default_model()
    m.aircraft_radius = R_AC
    return m


def test_arrival_box_is_the_references(arrivals, ref):
    for a, t, _, _ in arrivals:
        box = capi.target_box(t["y"][0], t["z"][0])
        assert box[0] == 676345.60 and box[1] == 676345.60 + 5000          # Drone.cpp:138-141 / :152-155
        assert np.array_equal(box, ref.cubed_volume(t["y"][0], t["z"][0]))
    assert arrivals[0][1]["y"][0] < 5705065.74 < arrivals[1][1]["y"][0]      # one per runway (27L, 27R)


def test_arrivals_fp64_per_trial_matches_unmodified_reference(engine, ref, arrivals):
    n = 30000
    total = 0
    for a, t, ring_csv, ring in arrivals:
        for j in range(N_RING):
            res = engine.simulate([t], ring[j:j + 1], seed=SEED, trial_count=n, precision=capi.FP64,
                                  flags=capi.FLAG_PER_TRIAL, track_id_base=a, drone_id_base=j, model=_model())
            draws = engine.dump_draws(0, 0, SEED, 0, n, track_id_base=a, drone_id_base=j)
            assert draws["tx"].min() >= 676345.60 and draws["tx"].max() < 676345.60 + 5000
            r = ref.run(ring_csv, j, a, t, draws, r_ac=R_AC)
            assert r["range_violations"] == 0
            assert np.array_equal(res.per_trial["hit"][0], r["hit"]), (a, j)
            assert np.array_equal(res.per_trial["first_hit"][0], r["first_hit"]), (a, j)
            assert np.abs(res.per_trial["min_dist"][0] - r["min_dist"]).max() < POS_TOL
            assert int(res.counts[0]) == r["hits"]
            total += r["hits"]
    assert total > 1000          # not a vacuous comparison


def test_arrivals_mixed_equals_fp64_and_oracle(engine, oracle, arrivals):
    """Both arrivals x the whole ring in one scenario: mixed (slab-first and 2-D screens), culled and fp64
    runs agree record for record, and the counts are the oracle's."""
    import oracle_api as oa

    n = 40000
    keys = ["track", "drone", "trial", "t1st", "first_hit"]
    om = oa.Model()
    oracle.L.orc_default_model(om)
    om.aircraft_radius = R_AC
    for a, t, _, ring in arrivals:
        kw = dict(seed=SEED, trial_begin=5 * n, trial_count=n, model=_model(), track_id_base=a, record_capacity=1 << 18)
        f64 = engine.simulate([t], ring, precision=capi.FP64, **kw)
        assert not f64.overflow
        for flags in (0, capi.FLAG_SCREEN_2D, capi.FLAG_REACH_CULL):
            mx = engine.simulate([t], ring, precision=capi.MIXED, flags=flags, **kw)
            assert np.array_equal(f64.counts, mx.counts), flags
            assert np.array_equal(f64.records[keys], mx.records[keys]), flags
            assert np.array_equal(f64.records["min_dist"], mx.records["min_dist"]), flags
            assert np.array_equal(f64.first_conflict, mx.first_conflict), flags
        box = capi.target_box(t["y"][0], t["z"][0])
        for j in (0, 3, 7, 9):
            h, ids = oracle.run_philox(t, box, ring[j], SEED, j, a, 5 * n, n, cap=1 << 18, model=om)
            assert int(f64.counts[j]) == h
            got = f64.records[f64.records["drone"] == j]["trial"]
            assert np.array_equal(np.sort(got), np.sort(ids))
        assert f64.hits > 100
