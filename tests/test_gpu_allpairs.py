"""GPU tier: long tracks (multi-tile TMA streaming) in the Monte-Carlo path, and the all-pairs
stress path of BASELINE config 5 (every ray against every track), both against the oracle."""
import numpy as np
import pytest

from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu
AP_TRACK_ID = 0xFFFFFFFF


def _extend(track, n_steps, wobble=0.0):
    """Continue a synthetic departure at its final velocity up to n_steps samples."""
    x, y, z = (track[k].copy() for k in "xyz")
    n0 = len(x)
    k = np.arange(1, n_steps - n0 + 1)
    vx, vy, vz = x[-1] - x[-2], y[-1] - y[-2], z[-1] - z[-2]
    x = np.concatenate([x, x[-1] + vx * k])
    y = np.concatenate([y, y[-1] + vy * k + wobble * np.sin(k / 40.0)])
    z = np.concatenate([z, np.minimum(z[-1] + vz * k, 3000.0)])
    return dict(track, x=x, y=y, z=z, n_steps=n_steps)


def test_long_tracks_stream_through_tiles(engine, oracle, tracks, ring1):
    """n_steps of 129 (just past the single-tile limit), 700 (2 tiles of 512) and 1300 (3 tiles):
    mixed == fp64 == oracle, with hits deep inside the second tile."""
    n, t_ref = 1300, 30
    drones = ring1[40:60]
    box0 = capi.target_box(tracks[0]["y"][0], 0.0)
    # a "shadowing" aircraft: it flies the path of one particular trial (drone 10, trial 5 of track
    # id 2), displaced sideways by an offset that only closes after index ~890 -> first hits deep in
    # the second tile, for that trial and its near neighbours
    dummy = {"x": np.zeros(n), "y": np.zeros(n), "z": np.zeros(n), "takeoff_t": t_ref}
    star = oracle.draws(21, 5, 1, 10, 2, t_ref, box0)[0]
    sx, sy, sz = oracle.trajectory(dummy, drones[10], star)
    i = np.arange(n)
    loiter = {"x": sx + 0.5 * np.maximum(0, 900 - i), "y": sy.copy(), "z": sz.copy(), "takeoff_t": t_ref,
              "box": box0, "n_steps": n}
    cases = [_extend(tracks[0], 129), _extend(tracks[1], 700, wobble=30.0), loiter]
    runs = 3000
    a = engine.simulate(cases, drones, seed=21, trial_count=runs, precision=capi.FP64)
    b = engine.simulate(cases, drones, seed=21, trial_count=runs, precision=capi.MIXED)
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.records[["track", "drone", "trial", "first_hit"]],
                          b.records[["track", "drone", "trial", "first_hit"]])
    got = a.counts.reshape(len(cases), len(drones))
    for ta, t in enumerate(cases):
        box = t["box"] if "box" in t else capi.target_box(t["y"][0], t["z"][0])
        for jj in (0, 10, 19):
            h, _ = oracle.run_philox(t, box, drones[jj], 21, jj, ta, 0, runs)
            assert h == int(got[ta, jj]), (ta, jj)
    assert a.hits > 0 and int(got[2, 10]) > 0            # the shadowed trial (and neighbours) hit
    assert a.records["first_hit"].max() > 512            # ... beyond the first tile
    assert a.stats["tests_evaluated"] == b.stats["tests_evaluated"]


def _oracle_allpairs(oracle, trks, drones, seed, begin, n, t_ref, box):
    counts = np.zeros((len(trks), len(drones)), dtype=np.int64)
    ids = []
    for j in range(len(drones)):
        draws = oracle.draws(seed, begin, n, j, AP_TRACK_ID, t_ref, box)
        for a, t in enumerate(trks):
            hit, first, _ = oracle.run_trials(t, drones[j], draws)
            counts[a, j] = int(hit.sum())
            ids += [(a, j, begin + int(r), int(first[r])) for r in np.nonzero(hit)[0]]
    return counts, sorted(ids)


def test_allpairs_matches_oracle(engine, oracle, tracks, ring1):
    """Every ray of 30 drones against 7 tracks (5 departures, one 700-step and one 1300-step
    track): counts, colliding (track, drone, ray) ids and first-hit indices equal the oracle's."""
    trks = [tracks[0], tracks[1], tracks[2], tracks[3], tracks[4], _extend(tracks[1], 700, wobble=25.0),
            _extend(tracks[0], 1300)]
    drones = np.concatenate([ring1[0:10], ring1[40:60]])
    t_ref = max(t["takeoff_t"] for t in trks)
    box = capi.target_box(tracks[0]["y"][0], 0.0)
    n, begin, seed = 4000, 1000, 77
    engine.upload(trks, drones)
    res = engine.allpairs(seed, begin, n, t_ref, box)
    want_counts, want_ids = _oracle_allpairs(oracle, trks, drones, seed, begin, n, t_ref, box)
    got = res.counts.reshape(len(trks), len(drones)).astype(np.int64)
    assert np.array_equal(got, want_counts)
    got_ids = sorted((int(r["track"]), int(r["drone"]), int(r["trial"]), int(r["first_hit"])) for r in res.records)
    assert got_ids == want_ids and len(want_ids) > 5
    st = res.stats
    assert st["trials"] == len(drones) * n and st["hits"] == len(want_ids)
    assert st["level2"] >= st["candidates"] >= st["hits"]
    # evaluated tests = sum over (ray, track) of max(0, n_steps - t1st)
    t1 = np.concatenate([oracle.draws(seed, begin, n, j, AP_TRACK_ID, t_ref, box)["t1st"] for j in range(len(drones))])
    want_tests = sum(int(np.maximum(0, len(t["x"]) - t1).sum()) for t in trks)
    assert st["tests_evaluated"] == want_tests
    assert st["tests_issued"] >= want_tests


def test_allpairs_ragged_and_errors(engine, tracks, ring1):
    engine.upload(tracks[:3], ring1[40:45])
    box = capi.target_box(tracks[0]["y"][0], 0.0)
    t_ref = max(t["takeoff_t"] for t in tracks[:3])
    for n in (1, 511, 513):                        # not multiples of the 512-ray block
        r = engine.allpairs(3, 0, n, t_ref, box)
        assert r.stats["trials"] == 5 * n
    r0 = engine.allpairs(3, 0, 0, t_ref, box)
    assert r0.hits == 0 and r0.stats["trials"] == 0
    # splitting the ray range changes nothing
    full = engine.allpairs(3, 0, 6000, t_ref, box)
    lo = engine.allpairs(3, 0, 2500, t_ref, box)
    hi = engine.allpairs(3, 2500, 3500, t_ref, box)
    assert np.array_equal(full.counts, lo.counts + hi.counts)
    with pytest.raises(capi.DcrpError) as e:
        engine.allpairs(3, 0, 10, t_ref + 1, box)
    assert e.value.status == capi.ERR_INVALID
