"""GPU tier: randomised scenario shapes.  For every shape the fp64 kernel (the parity kernel, checked
against the oracle and the unmodified reference in test_gpu_parity.py) is the yardstick; the mixed
screen, the reach-culled screen and the sub-step pair must reproduce it exactly: counts, first-conflict
indices and the colliding (track, drone, trial, first_hit) records.  Shapes cover what the fixed tests
do not combine: 1-sample and 2-sample tracks, takeoff_t = 1 and = n_steps, tracks just around the
128 / 512 sample tile limits, takeoff_t beyond the 128-entry shared tables, aircraft that shadow a
drone (many hits), huge and tiny radii, trial counts around the item size."""
import os

import numpy as np
import pytest

from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu
KEYS = ["track", "drone", "trial", "first_hit"]


def _random_track(rng, base, n_steps, takeoff_t, ring, shadow, v_straight=19.0, start_alt=100.0):
    """A smooth random walk through the neighbourhood of the ring, sample 0 kept in the 27L band."""
    t = np.arange(n_steps, dtype=np.float64)
    x0 = np.mean([r[0] for r in ring]) + rng.uniform(-1500, 1500)
    y0 = float(base["y"][0])
    heading = rng.uniform(0, 2 * np.pi)
    speed = rng.uniform(0.0, 90.0)
    x = x0 + np.sin(heading) * speed * t + 40.0 * np.sin(t * rng.uniform(0.01, 0.3))
    y = y0 + np.cos(heading) * speed * t * rng.uniform(0.0, 1.0)
    z = np.clip(rng.uniform(2.0, 12.0) * (t - rng.uniform(0, takeoff_t)), 0.0, 600.0)
    if shadow is not None:
        # sit on a drone's straight first-stage line for a while, at its altitude: stage-1 and early stage-2 hits
        xs, ys, hs = shadow
        k = min(n_steps, rng.integers(1, 12))
        lo = rng.integers(0, max(1, n_steps - k))
        idx = np.arange(lo, lo + k)
        x[idx] = xs + np.sin(hs) * v_straight * idx + rng.uniform(-3, 3)
        y[idx] = ys + np.cos(hs) * v_straight * idx + rng.uniform(-3, 3)
        z[idx] = start_alt + rng.uniform(-2, 2)
    y[0] = y0
    return {"x": x, "y": y, "z": z, "takeoff_t": int(takeoff_t), "n_steps": int(n_steps),
            "box": capi.target_box(float(base["y"][0]), 0.0)}


def _same(a, b, what):
    assert np.array_equal(a.counts, b.counts), what
    assert np.array_equal(a.first_conflict, b.first_conflict), what
    assert np.array_equal(a.records[KEYS], b.records[KEYS]), what
    assert np.array_equal(a.records["min_dist"], b.records["min_dist"]), what


# DCRP_FUZZ_CASES=n widens the sweep for a one-off soak run (2000 cases were run on B200 in round 1)
@pytest.mark.parametrize("case", range(int(os.environ.get("DCRP_FUZZ_CASES", "64"))))
def test_random_shapes_mixed_and_cull_equal_fp64(engine, tracks, ring1, ring5, case):
    rng = np.random.default_rng(1000 + case)
    ring = ring1 if case % 3 else ring5
    n_drones = int(rng.integers(1, 9))
    j0 = int(rng.integers(0, 99 - n_drones))
    drones = ring[j0:j0 + n_drones]
    model = capi.default_model()
    model.aircraft_radius = float(rng.choice([3.5, 0.5, 40.0]))
    if case >= 24:       # other kinematic constants: the reach bound must follow them (incl. v_ascend > v_straight)
        model.v_straight = float(rng.choice([19.0, 10.0, 30.0]))
        model.v_ascend = float(rng.choice([8.0, 25.0, 0.5]))
        model.start_alt = float(rng.choice([100.0, 30.0]))
    lengths = [1, 2, 3, 31, 127, 128, 129, 200, 511, 512, 513, 700, 1100]
    n_tracks = int(rng.integers(1, 4))
    trks = []
    for _ in range(n_tracks):
        n_steps = int(rng.choice(lengths))
        takeoff_t = int(rng.choice([1, n_steps, max(1, n_steps // 2), min(n_steps, 30), min(n_steps, 150)]))
        shadow = drones[int(rng.integers(0, n_drones))] if rng.random() < 0.6 else None
        trks.append(_random_track(rng, tracks[0], n_steps, takeoff_t, drones, shadow, model.v_straight, model.start_alt))
    n = int(rng.choice([1, 63, 3071, 3072, 3073, 5000, 12289]))
    kw = dict(seed=77 + case, trial_begin=int(rng.integers(0, 2**40)), trial_count=n, model=model,
              record_capacity=1 << 18)
    ref = engine.simulate(trks, drones, precision=capi.FP64, **kw)
    if ref.overflow:
        pytest.skip("more hits than the record buffer: counts-only case")
    mixed = engine.simulate(trks, drones, precision=capi.MIXED, **kw)                  # slab-first screen where eligible
    mixed2d = engine.simulate(trks, drones, precision=capi.MIXED, flags=capi.FLAG_SCREEN_2D, **kw)   # sorted 2-D screen
    cull = engine.simulate(trks, drones, precision=capi.MIXED, flags=capi.FLAG_REACH_CULL, **kw)
    _same(ref, mixed, "mixed")
    _same(ref, mixed2d, "mixed, 2-D screen")
    _same(ref, cull, "cull")
    assert mixed2d.stats["slab_survivors"] == 0
    eligible = max(t["n_steps"] for t in trks) <= 512 and max(t["takeoff_t"] for t in trks) <= 128
    if eligible:      # every trial the later levels see is a slab survivor
        assert mixed.stats["slab_survivors"] >= mixed.stats["level2"] >= mixed.stats["candidates"] >= ref.hits
    mixed = mixed2d
    # (with takeoff_t > 64 a t1st bucket holds several values and its internal order comes from atomics:
    # the issued lane-steps then vary by a few warp-steps from run to run)
    assert cull.stats["tests_issued"] <= mixed.stats["tests_issued"] * 1.01 + 4096
    if max(t["n_steps"] for t in trks) < 2:
        with pytest.raises(capi.DcrpError):           # a segment needs two samples
            engine.simulate(trks, drones, precision=capi.FP64, flags=capi.FLAG_SUBSTEP, **kw)
        return
    sref = engine.simulate(trks, drones, precision=capi.FP64, flags=capi.FLAG_SUBSTEP, **kw)
    smix = engine.simulate(trks, drones, precision=capi.MIXED, flags=capi.FLAG_SUBSTEP, **kw)       # slab-first (k_screen3)
    smix2d = engine.simulate(trks, drones, precision=capi.MIXED, flags=capi.FLAG_SUBSTEP | capi.FLAG_SCREEN_2D, **kw)
    assert smix2d.stats["slab_survivors"] == 0
    if not sref.overflow:
        for name, got in (("substep", smix), ("substep, 2-D screen", smix2d)):
            _same(sref, got, name)
            assert np.array_equal(sref.records["first_time"], got.records["first_time"]), name
        if eligible:
            assert smix.stats["slab_survivors"] >= smix.stats["level2"] >= smix.stats["candidates"] >= sref.hits
    if min(t["n_steps"] for t in trks) >= 2:         # a 1-sample track has no segment to interpolate on
        assert sref.hits >= ref.hits


@pytest.mark.parametrize("case", range(int(os.environ.get("DCRP_FUZZ_AP_CASES", "12"))))
def test_random_shapes_allpairs_equals_oracle(engine, oracle, tracks, ring1, case):
    """The all-pairs path (every ray against every track) on random shapes: tracks around the 512-sample
    tile limit, 1- and 2-sample tracks, shadowing aircraft, ray counts around the 1024-ray block, reference
    takeoff times of 1 and beyond the shared-memory table: counts and (track, drone, ray, first_hit) ids
    equal the oracle's."""
    rng = np.random.default_rng(5000 + case)
    n_drones = int(rng.integers(1, 5))
    j0 = int(rng.integers(0, 99 - n_drones))
    drones = ring1[j0:j0 + n_drones]
    trks = []
    for _ in range(int(rng.integers(1, 4))):
        n_steps = int(rng.choice([1, 2, 40, 511, 512, 513, 1030]))
        takeoff_t = int(rng.choice([1, max(1, n_steps // 3), min(n_steps, 30), min(n_steps, 140)]))
        shadow = drones[int(rng.integers(0, n_drones))] if rng.random() < 0.7 else None
        trks.append(_random_track(rng, tracks[0], n_steps, takeoff_t, drones, shadow))
    t_ref = max(t["takeoff_t"] for t in trks)
    box = capi.target_box(tracks[0]["y"][0], 0.0)
    n = int(rng.choice([1, 100, 1023, 1024, 1025, 2100]))
    begin = int(rng.integers(0, 2**33))
    model = capi.default_model()
    model.aircraft_radius = float(rng.choice([3.5, 25.0]))
    engine.upload(trks, drones, model=model)
    res = engine.allpairs(9 + case, begin, n, t_ref, box, record_capacity=1 << 18)
    if res.overflow:
        pytest.skip("more hits than the record buffer")
    import oracle_api as oa

    om = oa.Model()
    om.aircraft_radius, om.drone_radius = model.aircraft_radius, model.drone_radius
    om.start_alt, om.v_straight, om.v_ascend = model.start_alt, model.v_straight, model.v_ascend
    want_counts = np.zeros((len(trks), len(drones)), dtype=np.int64)
    want_ids = []
    for j in range(len(drones)):
        draws = oracle.draws(9 + case, begin, n, j, 0xFFFFFFFF, t_ref, box)      # all-pairs rays: track id 2^32 - 1
        for a, t in enumerate(trks):
            hit, first, _ = oracle.run_trials(t, drones[j], draws, model=om)
            want_counts[a, j] = int(hit.sum())
            want_ids += [(a, j, begin + int(k), int(first[k])) for k in np.nonzero(hit)[0]]
    want_ids.sort()
    got = res.counts.reshape(len(trks), len(drones)).astype(np.int64)
    assert np.array_equal(got, want_counts)
    got_ids = sorted((int(r["track"]), int(r["drone"]), int(r["trial"]), int(r["first_hit"])) for r in res.records)
    assert got_ids == want_ids
