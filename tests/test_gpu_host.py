"""GPU tier: the C++ drop-in (host/Aircraft, host/Drone, Monte_Carlo) end to end --
the reference's own call sequence (Monte_Carlo.cpp:26-38) through the new classes,
checked against the oracle and against the batched C-ABI call."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu
SEED = 20170630


def _parse_blocks(text):
    blocks = []
    for chunk in text.split("\n\n"):
        rows = [l.split(",") for l in chunk.split("\n") if l.strip()]
        if rows:
            blocks.append((int(rows[0][3]), np.array([[float(v) for v in r[:3]] for r in rows])))
            assert all(int(r[3]) == int(rows[0][3]) for r in rows)
    return blocks


def test_monte_carlo_loop_matches_batch_and_oracle(engine, oracle, day, ring1, tmp_path):
    tracks, csv = day
    n_air, runs = 2, 10000
    out_dir = tmp_path / "Drone_Collisions"
    out_dir.mkdir()
    total = C.c_double(0.0)
    exact = C.c_uint64(0)
    capi._hcheck(capi.host().dcrph_monte_carlo(csv.encode(), 22, 0, n_air, pk.RING_1KM.encode(), 4, 99, runs, 3.5,
                                               0.19, SEED, capi.MIXED, str(out_dir).encode(), C.byref(total),
                                               C.byref(exact)))
    # (1) one Drone::Simulation call per pair == one batched dcrp_simulate over the same ids
    batch = engine.simulate(tracks[:n_air], ring1, seed=SEED, trial_count=runs, precision=capi.MIXED)
    assert int(total.value) == batch.hits == exact.value and batch.hits > 0
    # (2) the files: blocks ordered (drone asc, run asc), rows == oracle trajectories at print precision
    for a in range(n_air):
        text = open(out_dir / ("Drone_coords_%d.csv" % a)).read()
        blocks = _parse_blocks(text)
        recs = [r for r in batch.records if int(r["track"]) == a]
        assert [b[0] for b in blocks] == [int(r["trial"]) for r in recs]
        assert text.count("\n\n") == len(recs)
        for (run_no, rows), r in zip(blocks, recs):
            t = tracks[a]
            ox, oy, oz = oracle.trajectory(t, ring1[int(r["drone"])],
                                           (int(r["t1st"]), 0, float(r["tx"]), float(r["ty"]), float(r["tz"])))
            assert rows.shape == (len(ox), 3)
            assert np.abs(rows[:, 0] - ox).max() < 1.1e-4      # 10 significant digits of ~6.7e5
            assert np.abs(rows[:, 1] - oy).max() < 1.1e-3      # 10 significant digits of ~5.7e6
            assert np.abs(rows[:, 2] - oz).max() < 1.1e-6
    # (3) the oracle agrees on the count (same Philox draws)
    want = 0
    for a in range(n_air):
        t = tracks[a]
        box = capi.target_box(t["y"][0], t["z"][0])
        want += sum(oracle.run_philox(t, box, ring1[j], SEED, j, a, 0, runs)[0] for j in range(99))
    assert want == batch.hits


def test_batch_driver_equals_per_pair_loop(day, tmp_path):
    """One launch for the whole day (Batch.cpp) == the reference-style per-pair loop: same count,
    byte-identical collision files."""
    tracks, csv = day
    loop_dir, batch_dir = tmp_path / "loop", tmp_path / "batch"
    loop_dir.mkdir(); batch_dir.mkdir()
    total = C.c_double(0.0)
    exact = C.c_uint64(0)
    capi._hcheck(capi.host().dcrph_monte_carlo(csv.encode(), 22, 1, 4, pk.RING_1KM.encode(), 4, 99, 8000, 3.5, 0.19,
                                               SEED, capi.MIXED, str(loop_dir).encode(), C.byref(total),
                                               C.byref(exact)))
    btotal = C.c_uint64(0)
    capi._hcheck(capi.host().dcrph_monte_carlo_batch(csv.encode(), 22, 1, 4, pk.RING_1KM.encode(), 4, 99, 8000, 3.5,
                                                     0.19, SEED, capi.MIXED, str(batch_dir).encode(),
                                                     C.byref(btotal)))
    assert btotal.value == exact.value > 0
    for a in (1, 2, 3):
        name = "Drone_coords_%d.csv" % a
        assert open(loop_dir / name).read() == open(batch_dir / name).read()
    # the same two drivers with DCRP_FLAG_REACH_CULL: same count, byte-identical files
    cull_dir = tmp_path / "cull"
    cull_dir.mkdir()
    capi.host().dcrph_set_run_flags(capi.FLAG_REACH_CULL)
    try:
        ctotal = C.c_uint64(0)
        capi._hcheck(capi.host().dcrph_monte_carlo_batch(csv.encode(), 22, 1, 4, pk.RING_1KM.encode(), 4, 99, 8000,
                                                         3.5, 0.19, SEED, capi.MIXED, str(cull_dir).encode(),
                                                         C.byref(ctotal)))
    finally:
        capi.host().dcrph_set_run_flags(0)
    assert ctotal.value == exact.value
    for a in (1, 2, 3):
        name = "Drone_coords_%d.csv" % a
        assert open(loop_dir / name).read() == open(cull_dir / name).read()


def test_monte_carlo_binary_as_the_reference_is_run(engine, tracks, ring1, tmp_path):
    """./Monte_Carlo in a directory laid out like the reference checkout: default
    constants (first flight, 1 km ring, 99 drones, 10 000 runs), stdout line, file."""
    os.symlink(pk.RING_1KM, tmp_path / "drone_inital_positions_1km.csv")
    r = subprocess.run([pk.MONTE_CARLO, "--synthetic", "6"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"Number of collisions: (\d+)", r.stdout)
    assert m
    batch = engine.simulate(tracks[:1], ring1, seed=SEED, trial_count=10000, precision=capi.MIXED)
    assert int(m.group(1)) == batch.hits
    blocks = _parse_blocks(open(tmp_path / "Drone_Collisions" / "Drone_coords_0.csv").read())
    assert len(blocks) == batch.hits
    # other constants through flags (the reference needs a recompile for these), fp64 kernel
    r2 = subprocess.run([pk.MONTE_CARLO, "--synthetic", "6", "--precision", "fp64", "--runs", "2500", "--drones", "60"],
                        cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0
    part = engine.simulate(tracks[:1], ring1[:60], seed=SEED, trial_count=2500, precision=capi.FP64)
    assert int(re.search(r"Number of collisions: (\d+)", r2.stdout).group(1)) == part.hits
    r3 = subprocess.run([pk.MONTE_CARLO, "--synthetic", "6", "--batch"], cwd=tmp_path, capture_output=True, text=True,
                        timeout=300)
    assert r3.returncode == 0 and int(re.search(r"Number of collisions: (\d+)", r3.stdout).group(1)) == batch.hits
    r4 = subprocess.run([pk.MONTE_CARLO, "--synthetic", "6", "--cull"], cwd=tmp_path, capture_output=True, text=True,
                        timeout=300)
    assert r4.returncode == 0 and int(re.search(r"Number of collisions: (\d+)", r4.stdout).group(1)) == batch.hits


def test_out_of_band_runway_is_an_error_not_stale_values(engine, tracks, ring1, tmp_path):
    """Drone::CubedVolume leaves six stale bounds when latitude[0] is in neither band
    (Drone.cpp:135-160); the drop-in reports it."""
    t = tracks[0]
    bad = dict(t, y=t["y"] + 800.0)        # between 27L and 27R
    with pytest.raises(capi.DcrpError) as e:
        engine.simulate([bad], ring1, trial_count=100)
    assert e.value.status == capi.ERR_RUNWAY


def test_batch_driver_on_all_gpus_of_the_process(day, tmp_path):
    """Batch.cpp with n_gpus > 1: one process, the trial range split over its devices -- same count and
    byte-identical collision files as one device.  Skipped on a single-GPU box."""
    n = C.c_int(0)
    capi._check(capi.lib().dcrp_device_count(C.byref(n)))
    if n.value < 2:
        pytest.skip("needs at least 2 GPUs in the process")
    tracks, csv = day
    one_dir, all_dir = tmp_path / "one", tmp_path / "all"
    one_dir.mkdir(); all_dir.mkdir()
    a, b = C.c_uint64(0), C.c_uint64(0)
    capi._hcheck(capi.host().dcrph_monte_carlo_batch(csv.encode(), 22, 0, 4, pk.RING_1KM.encode(), 4, 99, 30001, 3.5,
                                                     0.19, SEED, capi.MIXED, str(one_dir).encode(), C.byref(a)))
    capi.host().dcrph_set_gpus(0)
    try:
        capi._hcheck(capi.host().dcrph_monte_carlo_batch(csv.encode(), 22, 0, 4, pk.RING_1KM.encode(), 4, 99, 30001,
                                                         3.5, 0.19, SEED, capi.MIXED, str(all_dir).encode(),
                                                         C.byref(b)))
    finally:
        capi.host().dcrph_set_gpus(1)
    assert a.value == b.value > 0
    for k in range(4):
        name = "Drone_coords_%d.csv" % k
        assert open(one_dir / name).read() == open(all_dir / name).read()
