"""GPU parity tests proper: every call goes through the C ABI of libdcrp.so and is
compared with the oracle (and, where oracle/_ref travelled, with the unmodified
reference) on the same draws.  Bars: bit-exact for draws, flags, counts, first-hit
indices and colliding-trial ids; 1e-9 m for fp64 positions / distances (device
libm vs glibc in atan/sin/cos; the north-star bound is 1e-4 m)."""
import numpy as np
import pytest

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu

SEED = 20170630
POS_TOL = 1e-9     # metres, fp64 mode vs reference (north star: 1e-4 m)


def _oracle_counts(oracle, tracks, ring, seed, begin, n, drones=None, track_ids=None):
    drones = range(len(ring)) if drones is None else drones
    counts, ids = {}, []
    for a, t in enumerate(tracks):
        box = capi.target_box(t["y"][0], t["z"][0])
        for j in drones:
            h, hit_ids = oracle.run_philox(t, box, ring[j], seed, j, a, begin, n)
            counts[(a, j)] = h
            ids += [(a, j, int(i)) for i in hit_ids]
    return counts, ids


def test_draw_dump_is_oracle_philox(engine, oracle, tracks, ring1):
    engine.upload(tracks[:2], ring1)
    for a, j, begin in ((0, 0, 0), (1, 57, 2**32 - 5), (1, 98, 2**40 + 123)):
        got = engine.dump_draws(a, j, SEED, begin, 4096)
        t = tracks[a]
        want = oracle.draws(SEED, begin, 4096, j, a, t["takeoff_t"], capi.target_box(t["y"][0], t["z"][0]))
        assert np.array_equal(got, want)


def test_fp64_per_trial_matches_oracle_on_replayed_draws(engine, oracle, tracks, ring1):
    """Dump the Philox table, replay it through the oracle: flags and first-hit
    indices bit-exact, min distance within POS_TOL."""
    n = 50000
    for a, j in ((0, 41), (1, 13), (3, 4)):
        t = tracks[a]
        res = engine.simulate([t], ring1[j:j + 1], seed=SEED, trial_count=n, precision=capi.FP64,
                              flags=capi.FLAG_PER_TRIAL, track_id_base=a, drone_id_base=j)
        draws = engine.dump_draws(0, 0, SEED, 0, n, track_id_base=a, drone_id_base=j)
        hit, first, mind = oracle.run_trials(t, ring1[j], draws)
        assert np.array_equal(res.per_trial["hit"][0], hit)
        assert np.array_equal(res.per_trial["first_hit"][0], first)
        assert np.abs(res.per_trial["min_dist"][0] - mind).max() < POS_TOL
        assert int(res.counts[0]) == int(hit.sum())


def test_fp64_matches_unmodified_reference(engine, ref, tracks, ring1):
    n = 50000
    a, j = 3, 4
    t = tracks[a]
    engine.upload([t], ring1[j:j + 1])
    draws = engine.dump_draws(0, 0, SEED, 0, n, track_id_base=a, drone_id_base=j)
    r = ref.run(pk.RING_1KM, j, a, t, draws)
    assert r["range_violations"] == 0
    res = engine.simulate([t], ring1[j:j + 1], seed=SEED, trial_count=n, precision=capi.FP64,
                          flags=capi.FLAG_PER_TRIAL, track_id_base=a, drone_id_base=j)
    assert np.array_equal(res.per_trial["hit"][0], r["hit"])
    assert np.array_equal(res.per_trial["first_hit"][0], r["first_hit"])
    assert np.abs(res.per_trial["min_dist"][0] - r["min_dist"]).max() < POS_TOL
    assert r["hits"] > 0


def test_golden_reference_vectors_replayed_on_gpu(engine, golden, tracks, ring1):
    """The committed answers of the unmodified reference (tests/golden) through the
    GPU replay path, fp64 and mixed."""
    for p, (a, j) in enumerate(golden["pairs"]):
        draws = golden["pair%d_draws" % p]
        t = tracks[a]
        res = engine.simulate([t], ring1[j:j + 1], seed=0, trial_count=len(draws), precision=capi.FP64,
                              flags=capi.FLAG_PER_TRIAL, replay=draws)
        assert np.array_equal(res.per_trial["hit"][0], golden["pair%d_hit" % p])
        assert np.array_equal(res.per_trial["first_hit"][0], golden["pair%d_first_hit" % p])
        assert np.abs(res.per_trial["min_dist"][0] - golden["pair%d_min_dist" % p]).max() < POS_TOL
        mixed = engine.simulate([t], ring1[j:j + 1], seed=0, trial_count=len(draws), precision=capi.MIXED,
                                replay=draws)
        assert int(mixed.counts[0]) == int(golden["pair%d_hit" % p].sum())
        want_ids = np.nonzero(golden["pair%d_hit" % p])[0]
        assert np.array_equal(np.sort(mixed.records["trial"]).astype(np.int64), want_ids)


def test_mixed_equals_fp64_and_oracle_whole_ring(engine, oracle, tracks, ring1):
    """All 99 drones x 2 tracks x 20 000 trials: counts and colliding (track, drone,
    trial) ids identical across fp64, mixed and the oracle."""
    n = 20000
    tr = tracks[:2]
    want, ids = _oracle_counts(oracle, tr, ring1, SEED, 1000, n)
    for prec in (capi.FP64, capi.MIXED):
        res = engine.simulate(tr, ring1, seed=SEED, trial_begin=1000, trial_count=n, precision=prec)
        got = res.counts.reshape(len(tr), len(ring1))
        for (a, j), h in want.items():
            assert int(got[a, j]) == h, (prec, a, j)
        assert [(int(r["track"]), int(r["drone"]), int(r["trial"])) for r in res.records] == ids
        assert res.stats["hits"] == len(ids)
        assert res.stats["trials"] == len(tr) * len(ring1) * n


def test_mixed_equals_fp64_at_scale(engine, tracks, ring1, ring5):
    """1e7-trial configs (BASELINE config 2 size): mixed == fp64 record for record."""
    for ring, n in ((ring1, 101011), (ring5, 101011)):
        a = engine.simulate(tracks[:1], ring, seed=SEED, trial_count=n, precision=capi.FP64)
        b = engine.simulate(tracks[:1], ring, seed=SEED, trial_count=n, precision=capi.MIXED)
        assert np.array_equal(a.counts, b.counts)
        assert np.array_equal(a.records[["track", "drone", "trial", "t1st", "first_hit"]],
                              b.records[["track", "drone", "trial", "t1st", "first_hit"]])
        assert np.array_equal(a.records["min_dist"], b.records["min_dist"])
        assert np.array_equal(a.first_conflict, b.first_conflict)
        assert b.stats["tests_evaluated"] == a.stats["tests_evaluated"]
        assert b.stats["candidates"] >= b.stats["hits"]
    assert a.stats["trials"] == 99 * 101011


def test_whole_day_batch_fp32_vs_fp64(engine, oracle, ring1, ring5, tmp_path):
    """BASELINE config 4 in miniature: every departure of a (synthetic, 48-flight) day in ONE
    launch -- 4752 (track, drone) pairs -- fp64 == mixed exactly, fp32 within knife-edge flips,
    and two arbitrary pairs re-derived by the oracle."""
    from drone_collision_rp_b200 import scenario

    day, _ = scenario.synthetic_tracks(str(tmp_path), 48, seed=99)
    n = 6000
    a = engine.simulate(day, ring1, seed=5, trial_count=n, precision=capi.FP64)
    b = engine.simulate(day, ring1, seed=5, trial_count=n, precision=capi.MIXED)
    c = engine.simulate(day, ring1, seed=5, trial_count=n, precision=capi.FP32)
    assert a.hits > 100
    assert np.array_equal(a.counts, b.counts)
    assert np.array_equal(a.records[["track", "drone", "trial", "first_hit"]],
                          b.records[["track", "drone", "trial", "first_hit"]])
    assert abs(int(c.hits) - int(a.hits)) <= max(2, a.hits // 50)
    got = a.counts.reshape(len(day), len(ring1))
    for (ta, j) in ((17, int(np.argmax(got[17]))), (40, int(np.argmax(got[40])))):
        t = day[ta]
        h, _ = oracle.run_philox(t, capi.target_box(t["y"][0], t["z"][0]), ring1[j], 5, j, ta, 0, n)
        assert h == int(got[ta, j])
    # the 5 km ring barely reaches the departures (SURVEY: vacuous counts): compare records anyway
    a5 = engine.simulate(day[:8], ring5, seed=5, trial_count=n, precision=capi.FP64)
    b5 = engine.simulate(day[:8], ring5, seed=5, trial_count=n, precision=capi.MIXED)
    assert np.array_equal(a5.counts, b5.counts) and np.array_equal(a5.first_conflict, b5.first_conflict)


def test_config3_scale_1e10_trials(engine, tracks, ring1, ring5):
    """BASELINE config 3 at full size on one GPU: 1.0e10 trials against one track, 1 km and 5 km
    rings.  Size-independent properties: the count over the whole range equals the sum over two
    disjoint halves (what two ranks would compute), every record is reproduced, ids are unique."""
    n = 101_010_102                      # x 99 drones = 1.0000000098e10 trials
    for ring in (ring1, ring5):
        full = engine.simulate(tracks[:1], ring, seed=SEED, trial_count=n, precision=capi.MIXED,
                               record_capacity=1 << 18)
        lo = engine.simulate(tracks[:1], ring, seed=SEED, trial_begin=0, trial_count=n // 2, precision=capi.MIXED,
                             record_capacity=1 << 18)
        hi = engine.simulate(tracks[:1], ring, seed=SEED, trial_begin=n // 2, trial_count=n - n // 2,
                             precision=capi.MIXED, record_capacity=1 << 18)
        assert full.stats["trials"] == 99 * n >= 10**10
        assert np.array_equal(full.counts, lo.counts + hi.counts)
        assert not full.overflow and len(full.records) == full.hits
        ids = lambda r: set(zip(r.records["drone"].tolist(), r.records["trial"].tolist()))
        assert ids(full) == ids(lo) | ids(hi) and len(ids(full)) == full.hits
        assert full.stats["tests_evaluated"] == lo.stats["tests_evaluated"] + hi.stats["tests_evaluated"]
        print("ring %s: %d hits in %.3e trials, p = %.3e, screen %.1f ms" % (
            "1km" if ring is ring1 else "5km", full.hits, 99 * n, full.hits / (99 * n), full.stats["ms_trials"]))
    assert full.hits >= 0


def test_fp32_mode_is_close(engine, tracks, ring1):
    n = 200000
    a = engine.simulate(tracks[:1], ring1, seed=SEED, trial_count=n, precision=capi.FP64)
    b = engine.simulate(tracks[:1], ring1, seed=SEED, trial_count=n, precision=capi.FP32)
    ia = set(zip(a.records["drone"].tolist(), a.records["trial"].tolist()))
    ib = set(zip(b.records["drone"].tolist(), b.records["trial"].tolist()))
    assert a.hits > 50
    assert len(ia ^ ib) <= max(2, a.hits // 50)          # knife-edge flips only
    common = sorted(ia & ib)
    ra = {(int(r["drone"]), int(r["trial"])): r for r in a.records}
    rb = {(int(r["drone"]), int(r["trial"])): r for r in b.records}
    assert max(abs(ra[k]["min_dist"] - rb[k]["min_dist"]) for k in common) < 0.02


def test_trial_range_and_id_base_invariance(engine, tracks, ring1):
    """Results depend only on (seed, trial id, drone id, track id): splitting the
    trial range, or running one pair alone with id bases, changes nothing."""
    n = 60000
    full = engine.simulate(tracks[:2], ring1, seed=11, trial_begin=5, trial_count=n, precision=capi.MIXED)
    lo = engine.simulate(tracks[:2], ring1, seed=11, trial_begin=5, trial_count=n // 3, precision=capi.MIXED)
    hi = engine.simulate(tracks[:2], ring1, seed=11, trial_begin=5 + n // 3, trial_count=n - n // 3,
                         precision=capi.MIXED)
    assert np.array_equal(full.counts, lo.counts + hi.counts)
    ids = lambda r: [(int(x["track"]), int(x["drone"]), int(x["trial"])) for x in r.records]
    assert sorted(ids(full)) == sorted(ids(lo) + ids(hi))
    a, j = 1, 13
    one = engine.simulate([tracks[a]], ring1[j:j + 1], seed=11, trial_begin=5, trial_count=n,
                          precision=capi.MIXED, track_id_base=a, drone_id_base=j)
    assert int(one.counts[0]) == int(full.counts.reshape(2, -1)[a, j])
    assert [int(x["trial"]) for x in one.records] == [t for (aa, jj, t) in ids(full) if (aa, jj) == (a, j)]


def test_trajectories_match_oracle(engine, oracle, tracks, ring1):
    res = engine.simulate(tracks[:2], ring1, seed=SEED, trial_count=50000, precision=capi.MIXED)
    assert len(res.records) > 5
    xyz = engine.trajectories(res.records)
    for k, r in enumerate(res.records[:12]):
        t = tracks[int(r["track"])]
        draw = (int(r["t1st"]), 0, float(r["tx"]), float(r["ty"]), float(r["tz"]))
        ox, oy, oz = oracle.trajectory(t, ring1[int(r["drone"])], draw)
        n = len(ox)
        assert np.abs(xyz[k, 0, :n] - ox).max() < POS_TOL
        assert np.abs(xyz[k, 1, :n] - oy).max() < POS_TOL
        assert np.abs(xyz[k, 2, :n] - oz).max() < POS_TOL
        i = int(r["first_hit"])
        d = np.sqrt((ox[i] - t["x"][i]) ** 2 + (oy[i] - t["y"][i]) ** 2 + (oz[i] - t["z"][i]) ** 2)
        assert d < 3.69 and abs(float(r["min_dist"]) - d) < 3.69


def test_edge_cases(engine, tracks, ring1):
    t = tracks[0]
    # empty run
    r = engine.simulate([t], ring1[:3], seed=1, trial_count=0)
    assert r.hits == 0 and r.stats["trials"] == 0
    # ragged: one trial, and a count that is not a multiple of any chunk size
    for n in (1, 2047, 2049):
        a = engine.simulate([t], ring1[40:44], seed=1, trial_count=n, precision=capi.FP64)
        b = engine.simulate([t], ring1[40:44], seed=1, trial_count=n, precision=capi.MIXED)
        assert np.array_equal(a.counts, b.counts) and a.stats["trials"] == 4 * n
        assert a.stats["tests_evaluated"] == b.stats["tests_evaluated"]
    # takeoff_t == 1: stage 1 never runs (fixture aircraft 6 / run 4433 in the reference data)
    t1 = dict(t, takeoff_t=1)
    a = engine.simulate([t1], ring1, seed=3, trial_count=5000, precision=capi.FP64, flags=capi.FLAG_PER_TRIAL)
    b = engine.simulate([t1], ring1, seed=3, trial_count=5000, precision=capi.MIXED)
    assert np.array_equal(a.counts, b.counts)
    assert a.stats["tests_shared"] == 99 * 5000
    # takeoff_t == n_steps: t1st may equal n_steps (empty stage 2)
    tn = dict(t, takeoff_t=len(t["x"]))
    a = engine.simulate([tn], ring1[40:60], seed=3, trial_count=4000, precision=capi.FP64)
    b = engine.simulate([tn], ring1[40:60], seed=3, trial_count=4000, precision=capi.MIXED)
    assert np.array_equal(a.counts, b.counts)
    # a drone parked on the runway under the aircraft: stage-1 hits, shared per pair
    park = np.array([[t["x"][3] - 0.0, t["y"][3] + 19.0 * 3, np.pi]])     # flies south onto sample 3
    m = capi.default_model()
    m.start_alt = 0.0
    a = engine.simulate([t], park, seed=3, trial_count=3000, precision=capi.FP64, model=m, flags=capi.FLAG_PER_TRIAL)
    b = engine.simulate([t], park, seed=3, trial_count=3000, precision=capi.MIXED, model=m)
    assert a.hits > 0 and np.array_equal(a.counts, b.counts)
    assert set(a.per_trial["first_hit"][0][a.per_trial["hit"][0] == 1].tolist()) == {3}
    # a huge radius: everything collides; candidate list overflows -> fp64 rerun inside fetch; records overflow flagged
    m = capi.default_model()
    m.aircraft_radius = 50000.0
    b = engine.simulate([t], ring1, seed=3, trial_count=60000, precision=capi.MIXED, model=m, record_capacity=1000)
    assert b.hits == 99 * 60000 and b.overflow and len(b.records) == 1000
    assert (b.first_conflict == 0).all()


def test_error_behaviour(engine, tracks, ring1):
    t = tracks[0]
    with pytest.raises(capi.DcrpError) as e:
        engine.simulate([dict(t, takeoff_t=0)], ring1, trial_count=10)
    assert e.value.status == capi.ERR_INVALID
    with pytest.raises(capi.DcrpError) as e:
        engine.simulate([dict(t, takeoff_t=len(t["x"]) + 1)], ring1, trial_count=10)
    assert e.value.status == capi.ERR_INVALID
    with pytest.raises(capi.DcrpError) as e:
        capi.target_box(5.0e6, 0.0)
    assert e.value.status == capi.ERR_RUNWAY
    with pytest.raises(capi.DcrpError) as e:
        engine.simulate([t], ring1, trial_count=10, precision=7)
    assert e.value.status == capi.ERR_INVALID
    with pytest.raises(capi.DcrpError) as e:
        engine.simulate([t], ring1, trial_count=10, precision=capi.MIXED, flags=capi.FLAG_PER_TRIAL)
    assert e.value.status == capi.ERR_INVALID
    bad = dict(t, x=t["x"].copy())
    bad["x"][5] = np.nan
    with pytest.raises(capi.DcrpError):
        engine.simulate([bad], ring1, trial_count=10)
    fresh = capi.Engine(0)
    with pytest.raises(capi.DcrpError) as e:
        fresh.run_async(trial_count=10)
    assert e.value.status == capi.ERR_STATE
    fresh.close()


def test_free_running_probability_vs_reference_as_shipped(engine, tracks, ring1):
    """p_hat(GPU, Philox) vs p_hat(reference, its own random_device RNG; counts
    committed in tests/golden/ref_freerun.json) on the same inputs: |z| < 3."""
    import json
    import os

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_freerun.json")))
    n = 2_000_000                      # per drone -> 1.98e8 trials per aircraft
    for entry in ref["aircraft"]:
        a = entry["index"]
        res = engine.simulate([tracks[a]], ring1, seed=SEED, trial_count=n, precision=capi.MIXED,
                              track_id_base=a)
        n_gpu, n_ref = 99 * n, entry["trials"]
        p_gpu, p_ref = res.hits / n_gpu, entry["hits"] / n_ref
        p = (res.hits + entry["hits"]) / (n_gpu + n_ref)
        sigma = np.sqrt(p * (1 - p) * (1 / n_gpu + 1 / n_ref))
        z = (p_gpu - p_ref) / sigma
        print("aircraft %d: p_gpu %.3e (%d hits)  p_ref %.3e (%d hits)  z %.2f" % (
            a, p_gpu, res.hits, p_ref, entry["hits"], z))
        assert abs(z) < 3.0


def test_sliced_launches_equal_one_launch(tracks, ring1):
    """Runs whose item count would exceed 2^30 are cut into slices of consecutive trials, one launch each.
    DCRP_MAX_ITEMS (a test aid read once per process) forces that at a small size, in a child process:
    fp64, mixed and culled runs of 40 000 trials in slices of <= 3 items per pair must equal the un-sliced ones."""
    import json
    import subprocess
    import sys

    code = r'''
import json, sys
import numpy as np
sys.path.insert(0, %r)
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario
import tempfile
eng = capi.Engine(0)
tracks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 2)
ring = scenario.load_ring(pk.RING_1KM)[:30]
out = {}
for name, prec, flags in (("fp64", capi.FP64, 0), ("mixed", capi.MIXED, 0), ("cull", capi.MIXED, capi.FLAG_REACH_CULL)):
    r = eng.simulate(tracks, ring, seed=31, trial_count=40000, precision=prec, flags=flags)
    out[name] = {"counts": r.counts.tolist(), "launches": r.stats["kernel_launches"], "trials": r.stats["trials"],
                 "ids": [[int(x["track"]), int(x["drone"]), int(x["trial"]), int(x["first_hit"])] for x in r.records]}
print(json.dumps(out))
''' % pk.REPO_ROOT
    import os

    def run(env_extra):
        env = dict(os.environ, **env_extra)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        return json.loads(p.stdout.strip().split("\n")[-1])

    whole = run({})
    sliced = run({"DCRP_MAX_ITEMS": "180"})          # 60 pairs -> 3 items per pair per launch
    for name in ("fp64", "mixed", "cull"):
        assert sliced[name]["counts"] == whole[name]["counts"], name
        assert sliced[name]["ids"] == whole[name]["ids"], name
        assert sliced[name]["trials"] == whole[name]["trials"] == 60 * 40000
        assert sliced[name]["launches"] > whole[name]["launches"] + 2, name      # several screen launches
    assert whole["fp64"]["counts"] == whole["mixed"]["counts"] == whole["cull"]["counts"]
    assert sum(whole["fp64"]["counts"]) > 0
