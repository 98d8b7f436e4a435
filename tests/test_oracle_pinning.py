"""Pins the oracle (oracle/dcrp_oracle.c) before anything is compared with it:
  1. Philox4x32-10 against the Random123 known-answer vectors;
  2. kinematics + writer format against the 52 collision blocks the reference ships
     (tests/golden/ref_collision_blocks.json, from Drone_Collisions/Drone_coords_*.csv);
  3. flags / first-hit / min-distance / trajectories against the answers of the
     UNMODIFIED reference Drone.cpp on replayed draws, both from the committed
     fixture (tests/golden/ref_replay_vectors.npz) and, where oracle/_ref is
     built, live."""
import json
import os

import numpy as np
import pytest

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

# aircraft k -> [(run, drone j, t1st)], derived there from the shipped files
APPENDIX_A = {
    0: [(2885, 47, 14), (5038, 47, 14), (9581, 47, 4), (9371, 50, 14), (7677, 52, 11), (5050, 53, 10),
        (8441, 53, 10), (3317, 64, 13)],
    1: [(7573, 3, 17), (8953, 5, 15), (1874, 8, 14), (5597, 95, 21)],
    4: [(2871, 94, 12)],
    6: [(4433, 88, 1), (1698, 93, 12), (23, 94, 18), (3584, 95, 19)],
}


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_mapping_ranges(oracle, tracks):
    t = tracks[0]
    box = capi.target_box(t["y"][0], t["z"][0])
    d = oracle.draws(1, 0, 20000, 5, 0, t["takeoff_t"], box)
    assert d["t1st"].min() == 1 and d["t1st"].max() == t["takeoff_t"]
    assert (d["tx"] >= box[0]).all() and (d["tx"] < box[1]).all()
    assert (d["ty"] >= box[2]).all() and (d["ty"] < box[3]).all()
    assert (d["tz"] >= box[4]).all() and (d["tz"] < box[5]).all()
    # uniformity of t1st (chi-square, generous)
    cnt = np.bincount(d["t1st"], minlength=t["takeoff_t"] + 1)[1:]
    exp = 20000 / t["takeoff_t"]
    assert ((cnt - exp) ** 2 / exp).sum() < 3 * t["takeoff_t"]


def _blocks():
    return json.load(open(os.path.join(GOLDEN, "ref_collision_blocks.json")))


def test_reference_collision_blocks_kinematics(oracle):
    """Replay every shipped block: drone row from the start point, t1st from the
    altitude kink, a target recovered from the stage-2 step; every row within the
    files' print precision."""
    g = _blocks()
    rows = json.load(open(os.path.join(GOLDEN, "ref_ring_rows.json")))["drone_inital_positions_1km.csv"]
    ring = np.array(rows)
    assert len(g["blocks"]) == 52
    per_aircraft = {}
    worst = 0.0
    for b in g["blocks"]:
        x, y, z = (np.array(b[k]) for k in "xyz")
        n = len(x)
        j = int(np.argmin(np.hypot(ring[:99, 0] - x[0], ring[:99, 1] - y[0])))
        assert np.hypot(ring[j, 0] - x[0], ring[j, 1] - y[0]) < 1e-3
        kink = np.nonzero(z != 100.0)[0]
        t1st = int(kink[0])
        assert (z[:t1st] == 100.0).all() and t1st >= 1
        per_aircraft.setdefault(b["aircraft"], []).append((b["run"], j, t1st))
        trk = {"x": np.zeros(n), "y": np.zeros(n), "z": np.zeros(n), "takeoff_t": max(t1st, 1)}
        # Recover a target that reproduces the ray.  The path's slope is sin(pitch), not
        # tan(pitch) (the horizontal step is not scaled by cos(pitch), Drone.cpp:225-227), so
        # the target is NOT on the flown ray: pitch = asin(dz/|dh|), target = kink + m*(dir, tan(pitch)).
        step = np.array([x[-1] - x[t1st - 1], y[-1] - y[t1st - 1], z[-1] - z[t1st - 1]]) / (n - t1st)
        vfh = np.hypot(step[0], step[1])
        pitch = np.arcsin(step[2] / vfh)
        m = 1000.0
        target = (x[t1st - 1] + m * step[0] / vfh, y[t1st - 1] + m * step[1] / vfh, 100.0 + m * np.tan(pitch))
        ox, oy, oz = oracle.trajectory(trk, ring[j], (t1st, 0) + target)
        err = max(np.abs(ox - x).max(), np.abs(oy - y).max(), np.abs(oz - z).max())
        worst = max(worst, err)
        assert err < 2e-3, (b["aircraft"], b["run"], err)
        # the speed law vf = 19 - (22/pi) * pitch, Drone.cpp:220, from the block alone
        if t1st < n - 1:
            step = np.array([x[-1] - x[t1st], y[-1] - y[t1st], z[-1] - z[t1st]]) / (n - 1 - t1st)
            vf = np.hypot(step[0], step[1])
            pitch = np.arcsin(step[2] / vf)
            assert abs(vf - (19 - 22 / np.pi * pitch)) < 1e-4
    for k, want in APPENDIX_A.items():
        assert per_aircraft[k] == want
    counts = [len(per_aircraft[k]) for k in range(10)]
    assert counts == [8, 4, 8, 8, 1, 3, 4, 3, 8, 5]
    print("worst replay error %.2e m" % worst)


def test_reference_collision_blocks_writer_format(oracle):
    """Drone::Output format (Drone.cpp:237-241): parse -> format reproduces the
    shipped text byte for byte, in the oracle and in the C++ host writer."""
    g = _blocks()
    b = g["blocks"][0]
    text = oracle.format_block(b["x"], b["y"], b["z"], b["run"])
    assert text == g["first_block_text"]
    import ctypes as C
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "Drone_coords_0.csv")
        for blk in g["blocks"][:8]:
            x, y, z = (np.array(blk[k]) for k in "xyz")
            assert capi.host().dcrph_write_block(path.encode(), capi.dptr(x), capi.dptr(y), capi.dptr(z), len(x),
                                                 blk["run"]) == 0
        got = open(path).read()
    want = "".join(oracle.format_block(blk["x"], blk["y"], blk["z"], blk["run"]) for blk in g["blocks"][:8])
    assert got == want and got.startswith(g["first_block_text"])


def test_oracle_equals_committed_reference_answers(oracle, golden, tracks, ring1):
    """Bit-exact against what the unmodified reference answered on these draws."""
    total = 0
    for p, (a, j) in enumerate(golden["pairs"]):
        t = tracks[a]
        assert np.array_equal(t["x"], golden["track%d_x" % a])      # same synthetic day
        hit, first, mind = oracle.run_trials(t, ring1[j], golden["pair%d_draws" % p])
        assert np.array_equal(hit, golden["pair%d_hit" % p])
        assert np.array_equal(first, golden["pair%d_first_hit" % p])
        assert np.array_equal(mind, golden["pair%d_min_dist" % p])
        total += int(hit.sum())
        # Philox draw table itself
        idx = golden["pair%d_trial" % p]
        box = capi.target_box(t["y"][0], t["z"][0])
        d = oracle.draws(int(golden["seed"]), int(idx[0]), 1, j, a, t["takeoff_t"], box)
        assert d[0] == golden["pair%d_draws" % p][0]
        # full trajectories
        sel = golden["pair%d_traj_trial" % p]
        pos = {int(tr): i for i, tr in enumerate(idx)}
        for s, tr in enumerate(sel):
            ox, oy, oz = oracle.trajectory(t, ring1[j], golden["pair%d_draws" % p][pos[int(tr)]])
            assert np.array_equal(np.stack([ox, oy, oz]), golden["pair%d_traj" % p][s])
    assert total == sum(int(golden["pair%d_total_hits" % p]) for p in range(len(golden["pairs"]))) and total > 100


def test_oracle_equals_live_reference(oracle, ref, tracks, ring1):
    n = 40000
    for a, j in ((0, 44), (1, 8), (5, 10)):
        t = tracks[a]
        box = capi.target_box(t["y"][0], t["z"][0])
        assert np.array_equal(box, ref.cubed_volume(t["y"][0], t["z"][0]))
        draws = oracle.draws(99, 7 * n, n, j, a, t["takeoff_t"], box)
        r = ref.run(pk.RING_1KM, j, a, t, draws, want_traj=False)
        hit, first, mind = oracle.run_trials(t, ring1[j], draws)
        assert r["range_violations"] == 0
        assert np.array_equal(hit, r["hit"]) and np.array_equal(first, r["first_hit"])
        assert np.array_equal(mind, r["min_dist"])


def test_reference_simulation_loop_and_output_file(oracle, ref, tracks, ring1, tmp_path, monkeypatch):
    """The reference's own Drone::Simulation + Output on replayed draws: count and
    file equal what the oracle predicts (this is the file the new host must match)."""
    a, j = 3, 4
    t = tracks[a]
    box = capi.target_box(t["y"][0], t["z"][0])
    n = 60000
    draws = oracle.draws(20170630, 0, n, j, a, t["takeoff_t"], box)
    monkeypatch.chdir(tmp_path)
    os.mkdir("Drone_Collisions")
    total = ref.simulation(pk.RING_1KM, j, a, t, draws, clear_output=1)
    hit, first, mind = oracle.run_trials(t, ring1[j], draws)
    assert total == hit.sum() and total > 0
    want = ""
    for r in np.nonzero(hit)[0]:
        ox, oy, oz = oracle.trajectory(t, ring1[j], draws[r])
        want += oracle.format_block(ox, oy, oz, int(r))
    assert open("Drone_Collisions/Drone_coords_%d.csv" % a).read() == want


def test_target_box_matches_reference(oracle, ref):
    for ay0, az0 in ((5705983.0, 0.0), (5704545.0, 0.0), (5705983.0, 12.0), (5704545.0, 300.0),
                     (5706340.62, 0.0), (5705792.58, 0.0), (5705065.74, 0.0), (5704388.38, 0.0)):
        rc, box = oracle.target_box(ay0, az0)
        assert rc == 0
        assert np.array_equal(box, ref.cubed_volume(ay0, az0))
        assert np.array_equal(box, capi.target_box(ay0, az0))
    rc, _ = oracle.target_box(5705500.0, 0.0)          # between the runways: reference leaves stale values
    assert rc == -1
    assert np.isnan(ref.cubed_volume(5705500.0, 0.0)[:4]).all()


def _substep_independent(traj, trk, radius):
    """Interpolated separation of one trial, written independently of oracle/dcrp_oracle.c: positions of drone and
    aircraft interpolated linearly between samples; per segment the closest approach in closed form (numpy longdouble)
    and by brute force on 4097 points.  Returns (min distance, first conflict time in index units or -1,
    brute-force min distance, longest relative step)."""
    ld = np.longdouble
    D = np.stack([traj[0].astype(ld) - trk["x"].astype(ld), traj[1].astype(ld) - trk["y"].astype(ld),
                  traj[2].astype(ld) - trk["z"].astype(ld)], axis=1)            # relative position per sample
    D0, E = D[:-1], D[1:] - D[:-1]
    a = (E * E).sum(1)
    b = (D0 * E).sum(1)
    t = np.where(a > 0, np.clip(-b / np.where(a > 0, a, 1), 0, 1), 0)
    closest = np.sqrt(((D0 + t[:, None] * E) ** 2).sum(1))
    ts = np.linspace(0.0, 1.0, 4097).astype(ld)
    dist = np.sqrt(((D0[:, None, :] + ts[None, :, None] * E[:, None, :]) ** 2).sum(2))      # [segment, sample]
    first_time = -1.0
    inside = dist < ld(radius)
    if inside.any():
        seg = int(np.argmax(inside.any(1)))
        k = int(np.argmax(inside[seg]))
        if k == 0:
            first_time = float(seg)
        else:                                   # entry between two brute-force points: bisect the exact distance
            lo, hi = ts[k - 1], ts[k]
            for _ in range(60):
                mid = (lo + hi) / 2
                if np.sqrt(((D0[seg] + mid * E[seg]) ** 2).sum()) < ld(radius):
                    hi = mid
                else:
                    lo = mid
            first_time = float(seg + hi)
    return float(closest.min()), first_time, float(dist.min()), float(np.sqrt(a.max()))


def test_substep_extension_anchored_on_reference_trajectories(oracle, golden, tracks, ring1):
    """The sub-step mode is an EXTENSION: the reference compares sample i with sample i only (Drone.cpp:254-259), so no
    reference output can pin it and oracle/dcrp_oracle.h says so.  What can be anchored on the reference is:
      (1) at the samples it reduces to the reference -- on the reference's own replayed draws every reference hit is a
          sub-step hit, found in the segment that ends or starts at the reference's first-hit index, and the continuous
          minimum distance never exceeds the reference's per-trial minimum;
      (2) on trajectories WRITTEN BY THE UNMODIFIED Drone.cpp (golden pair*_traj) and on oracle trajectories (bit-equal
          to the reference's, test above) an independent numpy evaluation of "both tracks interpolated linearly" --
          closed form in longdouble, and brute force on 4097 points per segment -- gives the oracle's minimum distance
          and first-conflict time, for the reference's radius and for one that makes most trials conflicts."""
    om = type(oracle.model)()
    oracle.L.orc_default_model(om)
    wide = type(oracle.model)()
    oracle.L.orc_default_model(wide)
    wide.aircraft_radius = 60.0
    n_conf = n_traj = 0
    for p, (a, j) in enumerate(golden["pairs"]):
        t = tracks[a]
        draws = golden["pair%d_draws" % p]
        hit_s, first_s, ftime, mind_s = oracle.run_trials_substep(t, ring1[j], draws)
        ref_hit, ref_first, ref_min = golden["pair%d_hit" % p] == 1, golden["pair%d_first_hit" % p], golden["pair%d_min_dist" % p]
        # (1) reduces to the reference at the samples
        assert hit_s[ref_hit].all()
        assert (mind_s <= ref_min * (1 + 1e-12) + 1e-9).all()
        assert ((first_s[ref_hit] == ref_first[ref_hit]) | (first_s[ref_hit] == ref_first[ref_hit] - 1)).all()
        assert (ftime[ref_hit] <= ref_first[ref_hit]).all() and (ftime[ref_hit] >= first_s[ref_hit]).all()
        assert (first_s[hit_s == 0] == -1).all() and (ftime[hit_s == 0] == -1).all()
        # (2a) the reference's own trajectories
        pos = {int(tr): i for i, tr in enumerate(golden["pair%d_trial" % p])}
        cases = [(golden["pair%d_traj" % p][s], draws[pos[int(tr)]]) for s, tr in enumerate(golden["pair%d_traj_trial" % p])]
        # (2b) more geometry: oracle trajectories of the first 25 draws
        cases += [(np.stack(oracle.trajectory(t, ring1[j], d)), d) for d in draws[:25]]
        for traj, d in cases:
            one = np.array([d], dtype=draws.dtype)
            for model in (om, wide):
                radius = model.aircraft_radius + model.drone_radius
                h, f, ft, md = oracle.run_trials_substep(t, ring1[j], one, model=model)
                want_min, want_time, brute_min, step = _substep_independent(traj, t, radius)
                assert abs(md[0] - want_min) <= 1e-9 * max(1.0, want_min)
                assert want_min - 1e-9 <= brute_min <= want_min + step / 4096
                knife = abs(want_min - radius) < 1e-6
                if not knife:
                    assert bool(h[0]) == (want_min < radius)
                    if h[0]:
                        assert abs(ft[0] - want_time) <= 1e-6 and f[0] == int(np.floor(ft[0] if ft[0] < len(t["x"]) - 1 else ft[0] - 1e-12))
                        n_conf += 1
                n_traj += 1
    assert n_traj > 500 and n_conf > 100
