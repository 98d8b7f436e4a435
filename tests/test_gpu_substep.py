"""GPU tier: DCRP_FLAG_SUBSTEP, the interpolated-separation EXTENSION.  The reference
has no such mode (Drone.cpp:254-259 compares sample i with sample i), so the checker is the oracle's
own restatement of the same definition (oracle/dcrp_oracle.c: parity unpinned for this function);
what IS anchored in the reference: sub-step conflicts are a superset of the sample conflicts."""
import numpy as np
import pytest

from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu
SEED = 20170630


def test_substep_fp64_per_trial_matches_oracle_definition(engine, oracle, tracks, ring1):
    n = 60000
    more = 0
    for a, j in ((0, 41), (1, 13), (3, 4)):
        t = tracks[a]
        res = engine.simulate([t], ring1[j:j + 1], seed=SEED, trial_count=n, precision=capi.FP64,
                              flags=capi.FLAG_PER_TRIAL | capi.FLAG_SUBSTEP, track_id_base=a, drone_id_base=j)
        draws = engine.dump_draws(0, 0, SEED, 0, n, track_id_base=a, drone_id_base=j)
        hit, first, ftime, mind = oracle.run_trials_substep(t, ring1[j], draws)
        assert np.array_equal(res.per_trial["hit"][0], hit)
        assert np.array_equal(res.per_trial["first_hit"][0], first)
        assert np.abs(res.per_trial["min_dist"][0] - mind).max() < 1e-9
        rec = {int(r["trial"]): r for r in res.records}
        for k in np.nonzero(hit)[0]:
            assert abs(float(rec[int(k)]["first_time"]) - ftime[k]) < 1e-6      # north star: times to 1e-6 s
            assert int(rec[int(k)]["first_hit"]) == first[k]
        # superset of the reference's sample-point conflicts
        shit, sfirst, _ = oracle.run_trials(t, ring1[j], draws)
        assert ((shit == 1) <= (hit == 1)).all()
        more += int(hit.sum()) - int(shit.sum())
        s_res = engine.simulate([t], ring1[j:j + 1], seed=SEED, trial_count=n, precision=capi.FP64,
                                track_id_base=a, drone_id_base=j)
        assert np.array_equal(s_res.records["first_time"], s_res.records["first_hit"].astype(np.float64))
    assert more > 0        # interpolation finds conflicts between the samples


def test_substep_mixed_equals_fp64(engine, tracks, ring1, ring5):
    for trks, ring, n in ((tracks[:3], ring1, 40000), (tracks[:1], ring5, 101011)):
        a = engine.simulate(trks, ring, seed=9, trial_count=n, precision=capi.FP64, flags=capi.FLAG_SUBSTEP)
        b = engine.simulate(trks, ring, seed=9, trial_count=n, precision=capi.MIXED, flags=capi.FLAG_SUBSTEP)
        assert np.array_equal(a.counts, b.counts)
        assert np.array_equal(a.records[["track", "drone", "trial", "first_hit"]],
                              b.records[["track", "drone", "trial", "first_hit"]])
        assert np.array_equal(a.records["first_time"], b.records["first_time"])
        assert np.array_equal(a.records["min_dist"], b.records["min_dist"])
        # the same through the 2-D screen (k_screen2): both screens serve the sub-step mode
        b2 = engine.simulate(trks, ring, seed=9, trial_count=n, precision=capi.MIXED,
                             flags=capi.FLAG_SUBSTEP | capi.FLAG_SCREEN_2D)
        assert np.array_equal(a.counts, b2.counts) and np.array_equal(a.records["first_time"], b2.records["first_time"])
        assert np.array_equal(a.records[["track", "drone", "trial", "first_hit"]],
                              b2.records[["track", "drone", "trial", "first_hit"]])
        assert b.stats["slab_survivors"] > 0 and b2.stats["slab_survivors"] == 0
        assert b.stats["slab_survivors"] >= b.stats["level2"] >= b.stats["candidates"] >= b.hits
        s = engine.simulate(trks, ring, seed=9, trial_count=n, precision=capi.MIXED)
        assert b.hits >= s.hits
        assert b.stats["level2"] > s.stats["level2"]          # the wider level-1 radius passes more trials on
    assert a.hits > 0


def test_substep_crossing_between_samples(engine, oracle, tracks, ring1):
    """An aircraft that crosses the drone's path BETWEEN two samples: 60 m away at the sample
    before, 60 m away on the other side at the sample after.  No sample conflict, one sub-step
    conflict for exactly the shadowed trial."""
    t = tracks[0]
    j = 50
    box = capi.target_box(t["y"][0], 0.0)
    n = len(t["x"])
    star = oracle.draws(5, 7, 1, 0, 0, t["takeoff_t"], box)[0]
    dummy = {"x": np.zeros(n), "y": np.zeros(n), "z": np.zeros(n), "takeoff_t": t["takeoff_t"]}
    sx, sy, sz = oracle.trajectory(dummy, ring1[j], star)
    i = np.arange(n)
    k = 70
    # the aircraft sits far away except around index k, where it sweeps through the drone's segment
    ax = sx.copy()
    ay = sy + 5000.0              # 5 km abeam everywhere else: approaches and leaves sideways
    az = sz.copy()
    ax[k] = 0.5 * (sx[k] + sx[k + 1]) - 60.0
    ax[k + 1] = 0.5 * (sx[k] + sx[k + 1]) + 60.0
    ay[k] = ay[k + 1] = 0.5 * (sy[k] + sy[k + 1])
    az[k] = az[k + 1] = 0.5 * (sz[k] + sz[k + 1])
    trk = {"x": ax, "y": ay, "z": az, "takeoff_t": t["takeoff_t"], "box": box}
    runs = 2000
    plain = engine.simulate([trk], ring1[j:j + 1], seed=5, trial_count=runs, precision=capi.MIXED)
    sub = engine.simulate([trk], ring1[j:j + 1], seed=5, trial_count=runs, precision=capi.MIXED,
                          flags=capi.FLAG_SUBSTEP)
    assert plain.hits == 0
    assert sub.hits >= 1 and 7 in sub.records["trial"].tolist()
    r = sub.records[sub.records["trial"] == 7][0]
    assert int(r["first_hit"]) == k and k < float(r["first_time"]) < k + 1 and float(r["min_dist"]) < 3.69


def test_substep_flag_errors(engine, tracks, ring1):
    with pytest.raises(capi.DcrpError) as e:
        engine.simulate(tracks[:1], ring1, trial_count=10, precision=capi.FP32, flags=capi.FLAG_SUBSTEP)
    assert e.value.status == capi.ERR_INVALID


def test_substep_minimum_on_the_last_stage1_segment(engine, oracle, tracks, ring1):
    """The closest approach falls INSIDE a first-stage segment (aircraft parked 10 m beside the middle of
    segment 5 -> 6 of the drone's straight run): every kink index, in particular the kink at sample 6
    whose last stage-1 segment that is, must report the same interpolated minimum as the oracle."""
    t = tracks[0]
    j = 23
    x0, y0, h0 = ring1[j]
    n = len(t["x"])
    box = capi.target_box(t["y"][0], 0.0)
    ux, uy = np.sin(h0), np.cos(h0)
    px = x0 + ux * 19.0 * 5.5 + uy * 10.0          # abeam of the middle of segment 5 -> 6
    py = y0 + uy * 19.0 * 5.5 - ux * 10.0
    trk = {"x": np.full(n, px), "y": np.full(n, py), "z": np.full(n, 100.0), "takeoff_t": t["takeoff_t"], "box": box}
    runs = 4000
    res = engine.simulate([trk], ring1[j:j + 1], seed=11, trial_count=runs, precision=capi.FP64,
                          flags=capi.FLAG_PER_TRIAL | capi.FLAG_SUBSTEP, drone_id_base=j)
    draws = engine.dump_draws(0, 0, 11, 0, runs, drone_id_base=j)
    hit, first, ftime, mind = oracle.run_trials_substep(trk, ring1[j], draws)
    assert np.array_equal(res.per_trial["hit"][0], hit)
    assert np.abs(res.per_trial["min_dist"][0] - mind).max() < 1e-9
    late = draws["t1st"] >= 7
    assert late.any() and np.abs(mind[late] - 10.0).max() < 1e-6
