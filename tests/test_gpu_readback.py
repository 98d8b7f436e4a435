"""GPU tier: reading a run back.  The same MIXED run with engine-owned counts (output arena copied to the host by
dcrp_fetch), with caller-owned device counts (DCRP_FLAG_ZERO_COUNTS; what bench.py's timed step uses) and as an
fp64 run must hand the caller the same counts, first-conflict indices, records and statistics -- also back to back
on one engine, with more records than travel with the first copy, with the record buffer overflowing, and when a
run is sliced into several launches.  Replaces nothing in the reference (/root/reference/Drone.cpp:274-276
accumulates on the host); it guards the readback of the counter that line increments."""
import os
import subprocess
import sys

import numpy as np
import pytest

from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu
SEED = 20170630
KEYS = ["track", "drone", "trial", "t1st", "first_hit"]
# (tests_issued depends on how the dynamic scheduling batches the re-scans: not compared)
STATS = ["trials", "tests_evaluated", "tests_shared", "drone_steps", "hits", "candidates", "level2", "slab_survivors"]


def _same(a, b, what):
    assert np.array_equal(a.counts, b.counts), what
    assert np.array_equal(a.first_conflict, b.first_conflict), what
    assert a.overflow == b.overflow, what
    assert np.array_equal(a.records[KEYS], b.records[KEYS]), what
    assert np.array_equal(a.records["min_dist"], b.records["min_dist"]), what


def test_own_counts_equal_caller_counts_and_fp64(engine, tracks, ring1):
    import torch

    trks = tracks[:2]
    engine.upload(trks, ring1)
    n_pairs = len(trks) * len(ring1)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    ext = torch.zeros(n_pairs, dtype=torch.int64, device=dev)
    for k, n in enumerate((1, 255, 256, 257, 3000, 20011)):
        kw = dict(seed=SEED, trial_begin=1000 * k, trial_count=n, record_capacity=4096, stream=stream.cuda_stream)
        pub = engine.fetch(engine.run_async(precision=capi.MIXED, **kw))                     # engine-owned counts
        run = engine.run_async(precision=capi.MIXED, d_counts=ext.data_ptr(), flags=capi.FLAG_ZERO_COUNTS, **kw)
        cop = engine.fetch(run)                                                             # caller-owned device counts
        f64 = engine.fetch(engine.run_async(precision=capi.FP64, **kw))
        stream.synchronize()
        assert np.array_equal(ext.cpu().numpy().astype(np.uint64), pub.counts), n
        assert np.array_equal(pub.first_conflict, cop.first_conflict) and np.array_equal(pub.records[KEYS], cop.records[KEYS])
        _same(pub, f64, n)
        for key in STATS:
            assert pub.stats[key] == cop.stats[key], (n, key)
        assert pub.stats["d2h_bytes"] > 0 and pub.stats["ms_device_total"] > 0


def test_no_timing_flag_changes_nothing_but_the_times(engine, tracks, ring5):
    """DCRP_FLAG_NO_TIMING: dcrp_stats.ms_* stay 0 (a run read back by k_publish then returns without waiting for its
    end-of-run event), everything else is the run's; without the flag the times are CUDA-event times."""
    kw = dict(seed=SEED, trial_begin=123, trial_count=20000, precision=capi.MIXED, record_capacity=4096)
    timed = engine.simulate(tracks[:1], ring5, **kw)
    quiet = engine.simulate(tracks[:1], ring5, flags=capi.FLAG_NO_TIMING, **kw)
    _same(timed, quiet, "no-timing")
    for key in ("trials", "tests_evaluated", "tests_shared", "drone_steps", "hits"):     # (survivor counts depend on the slab
        assert timed.stats[key] == quiet.stats[key], key                                 #  directions, refined at the second run)
    assert 0 < timed.stats["ms_trials"] <= timed.stats["ms_device_total"] < 50
    assert quiet.stats["ms_trials"] == 0 and quiet.stats["ms_device_total"] == 0 and quiet.stats["ms_confirm"] == 0
    with pytest.raises(capi.DcrpError):
        engine.last_run_ms()                                                     # the last run recorded no events
    f64 = engine.simulate(tracks[:1], ring5, seed=SEED, trial_begin=123, trial_count=20000, precision=capi.FP64,
                          record_capacity=4096, flags=capi.FLAG_NO_TIMING)
    _same(f64, quiet, "fp64 no-timing")
    assert f64.stats["ms_device_total"] == 0


def test_runs_back_to_back(engine, tracks, ring5):
    """40 runs on one engine, no other call between: every result is the run's own (counts of disjoint trial ranges
    add up to the count of the whole range, ids partition)."""
    engine.upload(tracks[:1], ring5)
    n, parts = 30000, 40
    whole = engine.fetch(engine.run_async(seed=SEED, trial_begin=0, trial_count=n * parts, precision=capi.FP64,
                                          record_capacity=1 << 16))
    total = np.zeros_like(whole.counts)
    ids = []
    for k in range(parts):
        r = engine.fetch(engine.run_async(seed=SEED, trial_begin=k * n, trial_count=n, precision=capi.MIXED,
                                          record_capacity=1 << 16))
        assert r.stats["trials"] == n * len(ring5)
        assert all(k * n <= int(t) < (k + 1) * n for t in r.records["trial"])
        total += r.counts
        ids += [(int(x["drone"]), int(x["trial"])) for x in r.records]
    assert np.array_equal(total, whole.counts)
    assert sorted(ids) == [(int(x["drone"]), int(x["trial"])) for x in whole.records]


def test_many_records_and_overflow(engine, tracks, ring1):
    """More hits than the 32 records that travel with the first copy (the rest is copied separately), and more hits
    than the record buffer holds (overflow flag, counts still exact)."""
    m = capi.default_model()
    m.aircraft_radius = 300.0
    trks, drones, n = tracks[:1], ring1[40:56], 4000
    kw = dict(seed=7, trial_begin=5, trial_count=n, model=m)
    f64 = engine.simulate(trks, drones, precision=capi.FP64, record_capacity=1 << 20, **kw)
    assert f64.hits > 2000
    big = engine.simulate(trks, drones, precision=capi.MIXED, record_capacity=1 << 20, **kw)
    _same(big, f64, "all records")
    for cap in (1, 31, 32, 33, 100):
        r = engine.simulate(trks, drones, precision=capi.MIXED, record_capacity=cap, **kw)
        assert r.overflow and len(r.records) == cap
        assert np.array_equal(r.counts, f64.counts) and np.array_equal(r.first_conflict, f64.first_conflict)
        want = set(zip(f64.records["drone"].tolist(), f64.records["trial"].tolist()))
        got = list(zip(r.records["drone"].tolist(), r.records["trial"].tolist()))
        assert len(set(got)) == cap and set(got) <= want


def test_sliced_launch_agrees(tracks, ring1):
    """DCRP_MAX_ITEMS forces several launches per run, each with its own block schedule and a re-armed work counter.
    Same answers.  (The variable is read once per process: run in a child.)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import tempfile
import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario
trks, _ = scenario.synthetic_tracks(tempfile.mkdtemp(), 2)
ring = scenario.load_ring(pk.RING_1KM)
e = capi.Engine(0)
a = e.simulate(trks, ring, seed=3, trial_count=5000, precision=capi.MIXED, record_capacity=4096)
b = e.simulate(trks, ring, seed=3, trial_count=5000, precision=capi.FP64, record_capacity=4096)
assert a.stats["kernel_launches"] > 3, a.stats["kernel_launches"]
assert np.array_equal(a.counts, b.counts) and np.array_equal(a.records["trial"], b.records["trial"])
print("sliced ok", a.stats["kernel_launches"], int(a.counts.sum()))
""" % (root, root)
    env = dict(os.environ, DCRP_MAX_ITEMS="1000")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "sliced ok" in out.stdout, out.stdout + out.stderr


def test_published_runs_across_scenario_sizes(engine, tracks, ring1):
    """The flag k_publish raises sits behind the output arena in the host's pinned mirror, so its place moves with the
    number of pairs -- onto words an earlier, larger run filled with counts and first-conflict indices: small
    integers, like the flag's sequence numbers.  On a fresh engine (sequence numbers from 1) a hit-dense 396-pair
    scenario and a one-pair scenario alternate; every result must be the run's own (dcrp_run_async clears the flag's
    word before it queues k_publish).  Both parities of the sequence number meet the stale words."""
    m = capi.default_model()
    m.aircraft_radius = 300.0
    big, small = (tracks[:4], ring1), (tracks[:1], ring1[47:48])
    kw = dict(seed=11, precision=capi.MIXED, model=m, record_capacity=1 << 15, flags=capi.FLAG_NO_TIMING)
    want_big = engine.simulate(*big, seed=11, trial_count=96, precision=capi.FP64, model=m, record_capacity=1 << 15)
    assert want_big.hits > 3000 and 2 < np.median(want_big.counts) < 96
    fresh = capi.Engine(0)
    try:
        for k in range(70):
            a = fresh.simulate(*big, trial_count=96, **kw)
            _same(a, want_big, ("big", k))
            for rep in range(2 if k == 35 else 1):          # (one extra run flips the parity of the small runs' numbers)
                n0 = 1000 * k + 500 * rep
                b = fresh.simulate(*small, trial_begin=n0, trial_count=700, **kw)
                want = engine.simulate(*small, seed=11, trial_begin=n0, trial_count=700, precision=capi.FP64, model=m,
                                       record_capacity=1 << 15)
                _same(b, want, ("small", k, rep))
                assert b.stats["trials"] == 700 and b.stats["hits"] == want.stats["hits"]
    finally:
        fresh.close()


def test_a_failed_run_cannot_be_fetched(engine, tracks, ring1):
    """dcrp_run_async that fails (here: a replay row of another track) leaves nothing to fetch -- not the previous
    run's arena read against the failed run's parameters."""
    engine.upload(tracks[:1], ring1[:4])
    ok = engine.fetch(engine.run_async(seed=5, trial_count=300, precision=capi.MIXED))
    assert ok.stats["trials"] == 1200
    bad = np.zeros(4 * 50, dtype=capi.DRAW_DTYPE)
    bad["t1st"] = tracks[0]["takeoff_t"] + 1
    with pytest.raises(capi.DcrpError) as e:
        engine.run_async(seed=5, trial_count=50, precision=capi.FP64, flags=capi.FLAG_PER_TRIAL, replay=bad)
    assert e.value.status == capi.ERR_INVALID
    with pytest.raises(capi.DcrpError) as e:
        engine.fetch(capi.Engine._run_struct(5, 0, 50, capi.FP64))
    assert e.value.status == capi.ERR_STATE
    again = engine.fetch(engine.run_async(seed=5, trial_count=300, precision=capi.MIXED))
    assert np.array_equal(again.counts, ok.counts)
