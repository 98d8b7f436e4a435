"""GPU tier: DCRP_FLAG_REACH_CULL, the opt-in exact culling by reachability (SURVEY 7.6 "exact culling of
impossible indices").  The bound decides tests instead of computing them, so the bar is identity with
the un-culled run: counts, first-conflict indices, colliding (track, drone, trial) ids, records and
min distances record for record -- and with the oracle, which knows nothing of the bound."""
import numpy as np
import pytest

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu

SEED = 20170630
KEYS = ["track", "drone", "trial", "t1st", "first_hit"]


def _same(a, b):
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.first_conflict, b.first_conflict)
    assert np.array_equal(a.records[KEYS], b.records[KEYS])
    assert np.array_equal(a.records["min_dist"], b.records["min_dist"])
    assert a.hits == b.hits


def test_cull_equals_plain_and_oracle(engine, oracle, tracks, ring1):
    """Two tracks x 99 drones x 20 000 trials (the whole-ring oracle case of test_gpu_parity)."""
    n = 20000
    tr = tracks[:2]
    plain = engine.simulate(tr, ring1, seed=SEED, trial_begin=1000, trial_count=n, precision=capi.MIXED)
    cull = engine.simulate(tr, ring1, seed=SEED, trial_begin=1000, trial_count=n, precision=capi.MIXED,
                           flags=capi.FLAG_REACH_CULL)
    _same(plain, cull)
    got = cull.counts.reshape(len(tr), len(ring1))
    ids = []
    for a, t in enumerate(tr):
        box = capi.target_box(t["y"][0], t["z"][0])
        for j in range(len(ring1)):
            h, hit_ids = oracle.run_philox(t, box, ring1[j], SEED, j, a, 1000, n)
            assert int(got[a, j]) == h, (a, j)
            ids += [(a, j, int(i)) for i in hit_ids]
    assert [(int(r["track"]), int(r["drone"]), int(r["trial"])) for r in cull.records] == ids
    st = cull.stats
    assert st["trials"] == plain.stats["trials"]
    assert st["tests_issued"] < plain.stats["tests_issued"] // 4          # the point of the exercise
    assert st["tests_culled"] > 0
    # evaluated and culled are reported separately; cleared pairs report a lower bound
    assert st["tests_evaluated"] + st["tests_culled"] <= plain.stats["tests_evaluated"]
    assert st["tests_evaluated"] <= st["tests_issued"]
    assert plain.stats["trials_culled"] == 0 and plain.stats["tests_culled"] == 0


def test_cull_at_scale_both_rings(engine, tracks, ring1, ring5):
    """BASELINE config 2 size (1e7 trials) on the 1 km and 5 km rings."""
    for ring in (ring1, ring5):
        plain = engine.simulate(tracks[:1], ring, seed=SEED, trial_count=101011, precision=capi.MIXED)
        cull = engine.simulate(tracks[:1], ring, seed=SEED, trial_count=101011, precision=capi.MIXED,
                               flags=capi.FLAG_REACH_CULL)
        _same(plain, cull)
        st = cull.stats
        assert st["trials_culled"] % 101011 == 0                               # whole pairs only
        assert st["trials_culled"] < st["trials"]
        print("ring %s: pairs cleared %d / 99, issued %.3e of %.3e tests, screen %.4f ms vs %.4f ms" % (
            "1km" if ring is ring1 else "5km", st["trials_culled"] // 101011, st["tests_issued"],
            plain.stats["tests_issued"], st["ms_trials"], plain.stats["ms_trials"]))
    assert st["trials_culled"] >= 50 * 101011       # the 5 km ring: most drones can never reach the track


def test_cull_whole_day_batch(engine, ring1, tmp_path):
    """48 departures x 99 drones in one launch: identical to the un-culled mixed run and to fp64."""
    from drone_collision_rp_b200 import scenario

    day, _ = scenario.synthetic_tracks(str(tmp_path), 48, seed=99)
    n = 6000
    ref = engine.simulate(day, ring1, seed=5, trial_count=n, precision=capi.FP64)
    cull = engine.simulate(day, ring1, seed=5, trial_count=n, precision=capi.MIXED, flags=capi.FLAG_REACH_CULL)
    assert ref.hits > 100
    assert np.array_equal(ref.counts, cull.counts)
    assert np.array_equal(ref.records[["track", "drone", "trial", "first_hit"]],
                          cull.records[["track", "drone", "trial", "first_hit"]])
    assert np.array_equal(ref.first_conflict, cull.first_conflict)


def test_cull_keeps_stage1_conflicts(engine, oracle, tracks, ring1):
    """An aircraft that shadows the drone's straight first stage collides BEFORE the kink: such pairs
    keep their full range (the trials must be drawn to know which of them kink after the conflict)."""
    t0 = tracks[0]
    n_steps, T = int(t0["n_steps"]), int(t0["takeoff_t"])
    x0, y0, h0 = ring1[7]
    i = np.arange(n_steps)
    # on the drone's stage-1 line at index 5 only (5 <= T - 1), far away otherwise
    x = np.array(t0["x"], dtype=np.float64).copy()
    y = np.array(t0["y"], dtype=np.float64).copy()
    z = np.array(t0["z"], dtype=np.float64).copy()
    x[5] = x0 + np.sin(h0) * 19.0 * 5
    y[5] = y0 + np.cos(h0) * 19.0 * 5
    z[5] = 100.0
    trk = dict(t0, x=x, y=y, z=z)
    n = 5000
    # the box depends on sample 0, which is untouched
    plain = engine.simulate([trk], ring1, seed=3, trial_count=n, precision=capi.MIXED, record_capacity=1 << 16)
    cull = engine.simulate([trk], ring1, seed=3, trial_count=n, precision=capi.MIXED, record_capacity=1 << 16,
                           flags=capi.FLAG_REACH_CULL)
    _same(plain, cull)
    assert int(cull.counts[7]) > n // 2        # every trial of drone 7 with t1st > 5
    h, _ = oracle.run_philox(trk, capi.target_box(y[0], z[0]), ring1[7], 3, 7, 0, 0, n)
    assert h == int(cull.counts[7])


def test_cull_flag_is_ignored_where_it_does_not_apply(engine, tracks, ring1):
    """fp64 mode, per-trial outputs and the sub-step extension run un-culled (same results, no culled stats)."""
    for kw in (dict(precision=capi.FP64), dict(precision=capi.MIXED, flags=capi.FLAG_SUBSTEP)):
        flags = kw.pop("flags", 0)
        a = engine.simulate(tracks[:1], ring1, seed=9, trial_count=4000, flags=flags, **kw)
        b = engine.simulate(tracks[:1], ring1, seed=9, trial_count=4000, flags=flags | capi.FLAG_REACH_CULL, **kw)
        assert np.array_equal(a.counts, b.counts)
        assert b.stats["trials_culled"] == 0 and b.stats["tests_culled"] == 0
        assert b.stats["tests_evaluated"] == a.stats["tests_evaluated"]


def test_cull_long_tracks_and_large_takeoff(engine, oracle, tracks, ring1):
    """Multi-tile tracks (700 and 1300 samples: the culled range is clamped tile by tile) and a takeoff
    time of 200 (> the 128-entry shared-memory tables: the reach ranges then come from global memory),
    with a shadowing aircraft whose conflicts sit deep in the second tile."""
    from test_gpu_allpairs import _extend

    n, t_ref = 1300, 200
    drones = ring1[40:60]
    box0 = capi.target_box(tracks[0]["y"][0], 0.0)
    dummy = {"x": np.zeros(n), "y": np.zeros(n), "z": np.zeros(n), "takeoff_t": t_ref}
    star = oracle.draws(21, 5, 1, 10, 2, t_ref, box0)[0]
    sx, sy, sz = oracle.trajectory(dummy, drones[10], star)
    i = np.arange(n)
    loiter = {"x": sx + 0.5 * np.maximum(0, 900 - i), "y": sy.copy(), "z": sz.copy(), "takeoff_t": t_ref,
              "box": box0, "n_steps": n}
    cases = [_extend(tracks[0], 129), _extend(tracks[1], 700, wobble=30.0), loiter]
    runs = 3000
    ref = engine.simulate(cases, drones, seed=21, trial_count=runs, precision=capi.FP64)
    plain = engine.simulate(cases, drones, seed=21, trial_count=runs, precision=capi.MIXED)
    cull = engine.simulate(cases, drones, seed=21, trial_count=runs, precision=capi.MIXED, flags=capi.FLAG_REACH_CULL)
    _same(plain, cull)
    assert np.array_equal(ref.counts, cull.counts)
    assert cull.hits > 0 and cull.records["first_hit"].max() > 512
    assert cull.stats["tests_issued"] < plain.stats["tests_issued"]
    got = cull.counts.reshape(len(cases), len(drones))
    h, _ = oracle.run_philox(loiter, box0, drones[10], 21, 10, 2, 0, runs)
    assert h == int(got[2, 10]) > 0


def test_cull_sliced_and_repeated_runs(engine, tracks, ring5):
    """The reach table is built once per scenario and reused by later runs; trial ranges compose."""
    eng = engine
    eng.upload(tracks[:2], ring5)
    tot = None
    for k in range(3):
        run = eng.run_async(seed=SEED, trial_begin=k * 50000, trial_count=50000, precision=capi.MIXED,
                            flags=capi.FLAG_REACH_CULL)
        r = eng.fetch(run)
        tot = r.counts.copy() if tot is None else tot + r.counts
    whole = eng.simulate(tracks[:2], ring5, seed=SEED, trial_count=150000, precision=capi.MIXED)
    assert np.array_equal(tot, whole.counts)
