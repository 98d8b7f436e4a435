"""GPU tier: hit-dense stress of the concurrent bookkeeping -- the warp-aggregated record append and the
per-pair atomics that replace the reference's sequential `Output(r); *total_collisions += 1`
(/root/reference/Drone.cpp:274-276).  With a 300 m "aircraft" a third of the trials collide, so thousands of
warps append records and bump counters at once; below capacity every hit must own exactly one record slot:
n_records == hits == sum(counts), all (track, drone, trial) unique, first_conflict == min(first_hit) per
pair, and the records are the oracle's.  Also: the status bits that make the perf cliffs visible
(dcrp_stats.status) and the replay-table validation."""
import numpy as np
import pytest

from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu
SEED = 424242


def _dense_model():
    m = capi.default_model()
    m.aircraft_radius = 300.0
    return m


def _check_bookkeeping(res, n_pairs):
    rec = res.records
    assert not res.overflow
    assert len(rec) == res.hits == res.stats["hits"] == int(res.counts.sum())
    key = (rec["track"].astype(np.uint64) << np.uint64(52)) | (rec["drone"].astype(np.uint64) << np.uint64(40)) | rec["trial"]
    assert len(np.unique(key)) == len(rec)                       # no slot written twice, no hit dropped
    assert np.all(np.diff(key.astype(np.int64)) > 0)            # sorted (track, drone, trial)
    n_drones = n_pairs // (int(rec["track"].max()) + 1 if len(rec) else 1)
    pair = rec["track"].astype(np.int64) * n_drones + rec["drone"].astype(np.int64)
    cnt = np.bincount(pair, minlength=n_pairs)
    assert np.array_equal(cnt.astype(np.uint64), res.counts)
    first = np.full(n_pairs, capi.INT32_MAX, dtype=np.int64)
    np.minimum.at(first, pair, rec["first_hit"].astype(np.int64))
    assert np.array_equal(first.astype(np.int32), res.first_conflict)
    assert (rec["first_hit"] >= 0).all() and (rec["min_dist"] < 300.19).all()


def test_hit_dense_bookkeeping_all_paths(engine, oracle, tracks, ring1):
    import oracle_api as oa

    trks, drones, n = tracks[:2], ring1[30:60], 6000
    kw = dict(seed=SEED, trial_begin=777, trial_count=n, model=_dense_model(), record_capacity=1 << 20)
    runs = {"fp64": engine.simulate(trks, drones, precision=capi.FP64, **kw),
            "mixed": engine.simulate(trks, drones, precision=capi.MIXED, **kw),
            "mixed2d": engine.simulate(trks, drones, precision=capi.MIXED, flags=capi.FLAG_SCREEN_2D, **kw),
            "cull": engine.simulate(trks, drones, precision=capi.MIXED, flags=capi.FLAG_REACH_CULL, **kw)}
    ref = runs["fp64"]
    assert 0.05 < ref.hits / (2 * 30 * n) < 0.95                 # dense, but both outcomes occur
    keys = ["track", "drone", "trial", "t1st", "first_hit"]
    for name, r in runs.items():
        _check_bookkeeping(r, 60)
        assert np.array_equal(r.records[keys], ref.records[keys]), name
        assert np.array_equal(r.records["min_dist"], ref.records["min_dist"]), name
    om = oa.Model()
    oracle.L.orc_default_model(om)
    om.aircraft_radius = 300.0
    for a, j in ((0, 0), (1, 17), (1, 29)):                      # every record of these pairs re-derived by the oracle
        t = trks[a]
        h, ids = oracle.run_philox(t, capi.target_box(t["y"][0], t["z"][0]), drones[j], SEED, j, a, 777, n,
                                   cap=1 << 20, model=om)
        mine = ref.records[(ref.records["track"] == a) & (ref.records["drone"] == j)]
        assert h == len(mine) and np.array_equal(mine["trial"], ids)


def test_hit_dense_bookkeeping_allpairs(engine, oracle, tracks, ring1):
    trks, drones, n = tracks[:3], ring1[40:48], 3000
    engine.upload(trks, drones, model=_dense_model())
    box = capi.target_box(trks[0]["y"][0], 0.0)
    t_ref = max(t["takeoff_t"] for t in trks)
    res = engine.allpairs(SEED, 99, n, t_ref, box, record_capacity=1 << 20)
    _check_bookkeeping(res, 24)
    assert res.hits > 0.05 * 24 * n


def test_status_bits_make_the_cliffs_visible(engine, tracks, ring1):
    t = tracks[0]
    plain = engine.simulate([t], ring1[:8], seed=1, trial_count=5000, precision=capi.MIXED, flags=capi.FLAG_REACH_CULL)
    assert plain.stats["status"] & capi.STAT_CULL_IGNORED == 0
    for prec, flags in ((capi.FP64, capi.FLAG_REACH_CULL), (capi.FP32, capi.FLAG_REACH_CULL),
                        (capi.MIXED, capi.FLAG_REACH_CULL | capi.FLAG_SUBSTEP)):
        r = engine.simulate([t], ring1[:8], seed=1, trial_count=5000, precision=prec, flags=flags)
        assert r.stats["status"] & capi.STAT_CULL_IGNORED, (prec, flags)
        assert r.stats["tests_culled"] == 0
    # candidate-list overflow of the 2-D screen path: the whole run is redone in fp64 and says so
    m = capi.default_model()
    m.aircraft_radius = 50000.0
    big = engine.simulate([t], ring1, seed=3, trial_count=60000, precision=capi.MIXED, flags=capi.FLAG_SCREEN_2D,
                          model=m, record_capacity=1000)
    assert big.hits == 99 * 60000 and big.overflow
    assert big.stats["status"] & capi.STAT_FP64_FALLBACK
    ok = engine.simulate([t], ring1, seed=3, trial_count=60000, precision=capi.MIXED, model=m, record_capacity=1000)
    assert ok.hits == 99 * 60000 and ok.stats["status"] & capi.STAT_FP64_FALLBACK == 0     # slab-first path: no list
    # the scenario cache of dcrp_simulate
    scn = capi.Scenario([t], ring1)
    first = engine.simulate(scn, seed=5, trial_count=1000)
    again = engine.simulate(scn, seed=5, trial_begin=1000, trial_count=1000)
    assert first.stats["status"] & capi.STAT_SCENARIO_CACHED == 0 and first.stats["h2d_bytes"] > 0
    assert again.stats["status"] & capi.STAT_SCENARIO_CACHED and again.stats["h2d_bytes"] == 0
    moved = dict(t, x=t["x"] + 1.0)
    other = engine.simulate([moved], ring1, seed=5, trial_count=1000)
    assert other.stats["status"] & capi.STAT_SCENARIO_CACHED == 0
    m2 = capi.default_model()
    m2.v_ascend = 9.0
    assert engine.simulate([moved], ring1, seed=5, trial_count=1000, model=m2).stats["status"] & capi.STAT_SCENARIO_CACHED == 0
    assert engine.simulate([moved], ring1, seed=5, trial_count=1000, model=m2).stats["status"] & capi.STAT_SCENARIO_CACHED


def test_replay_table_is_validated(engine, tracks, ring1):
    """A row that does not belong to the track (t1st outside [1, takeoff_t], non-finite target) is
    DCRP_ERR_INVALID before anything is launched, not an out-of-bounds access on the device."""
    t = tracks[0]
    engine.upload([t], ring1[:2])
    good = np.concatenate([engine.dump_draws(0, j, 1, 0, 64) for j in range(2)])
    for field, value in (("t1st", 0), ("t1st", t["takeoff_t"] + 1), ("t1st", -7), ("tx", np.nan), ("tz", np.inf)):
        bad = good.copy()
        bad[field][70] = value
        for prec in (capi.FP64, capi.MIXED):
            with pytest.raises(capi.DcrpError) as e:
                engine.simulate([t], ring1[:2], seed=0, trial_count=64, precision=prec, replay=bad)
            assert e.value.status == capi.ERR_INVALID
    ok = engine.simulate([t], ring1[:2], seed=0, trial_count=64, precision=capi.MIXED, replay=good)
    assert ok.stats["trials"] == 128


def test_two_engines_share_a_device(tracks, ring1):
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) belongs to the (kernel, device) pair, not to an engine: a
    second engine running a smaller scenario on the same device must not leave the first one unable to launch
    (found on a 2-GPU box in round 2: 'k_screen3 launch: invalid argument')."""
    big = [dict(t) for t in tracks[:4]]                                  # larger per-warp tables
    long_t = dict(tracks[0], takeoff_t=len(tracks[0]["x"]))
    a, b = capi.Engine(0), capi.Engine(0)
    for flags in (0, capi.FLAG_SCREEN_2D):
        r1 = a.simulate(big + [long_t], ring1, seed=9, trial_count=3000, flags=flags)
        small = dict(tracks[1], x=tracks[1]["x"][:40].copy(), y=tracks[1]["y"][:40].copy(), z=tracks[1]["z"][:40].copy(),
                     takeoff_t=5, n_steps=40)
        b.simulate([small], ring1[:1], seed=9, trial_count=3000, flags=flags)
        r2 = a.simulate(big + [long_t], ring1, seed=9, trial_count=3000, flags=flags)
        assert np.array_equal(r1.counts, r2.counts)
    a.close()
    b.close()
