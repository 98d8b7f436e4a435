"""CPU tier: the __host__ __device__ arithmetic the kernels execute
(drone-collision-rp_b200/csrc/dcrp_math.cuh, compiled here with g++ by
tests/hostsim) against the oracle -- Philox, draw mapping, the fp64 trial, and the
error of the fp32 screen relative to the margin the engine adds to the radius."""
import numpy as np

from drone_collision_rp_b200 import capi


def test_philox_and_draws_equal_oracle(hostsim, oracle, tracks):
    for ctr, key in (([0, 0, 0, 0], [0, 0]), ([1, 2, 3, 4], [5, 6]), ([0xffffffff] * 4, [0xffffffff] * 2)):
        assert hostsim.philox(ctr, key) == oracle.philox(ctr, key)
    t = tracks[2]
    box = capi.target_box(t["y"][0], t["z"][0])
    for begin in (0, 2**32 - 3, 2**45):
        assert np.array_equal(hostsim.draws(9, begin, 2000, 17, 2, t["takeoff_t"], box),
                              oracle.draws(9, begin, 2000, 17, 2, t["takeoff_t"], box))


def test_exact_trial_statements_equal_reference_answers(hostsim, golden, tracks, ring1):
    """Stage-1 tables + stage2_ray + sep2 + the sqrt-free threshold reproduce the
    unmodified reference's flags, first-hit indices and min distances bit for bit
    (host libm on both sides)."""
    for p, (a, j) in enumerate(golden["pairs"]):
        first, mind = hostsim.exact(tracks[a], ring1[j], golden["pair%d_draws" % p])
        assert np.array_equal(first, golden["pair%d_first_hit" % p])
        assert np.array_equal(first >= 0, golden["pair%d_hit" % p] == 1)
        assert np.array_equal(mind, golden["pair%d_min_dist" % p])


def test_sqrt_free_threshold_is_exact(hostsim):
    for r in (3.69, 0.19 + 3.5, 1.0, 2.0, 1e-3, 12345.678):
        c = hostsim.L.hs_thr_d2(r)
        below = np.nextafter(c, 0.0)
        assert np.sqrt(c) >= r and np.sqrt(below) < r


def test_magic_number_division_is_exact(hostsim):
    """FastDiv (pair / n_drones and item / chunks_per_pair in k_screen2) equals integer division for every divisor
    class: 1, powers of two, odd, just below / above powers of two, huge."""
    import ctypes as C

    hostsim.L.hs_fastdiv_mismatches.restype = C.c_long
    hostsim.L.hs_fastdiv_mismatches.argtypes = [C.c_uint32, C.c_uint32]
    for d in (1, 2, 3, 5, 7, 25, 33, 99, 100, 127, 128, 129, 296, 1000, 4095, 4096, 4097, 65535, 65536, 1000003,
              0x7fffffff, 0x80000000, 0x80000001, 0xfffffffe, 0xffffffff):
        assert hostsim.L.hs_fastdiv_mismatches(d, 65521) == 0, d


def test_polynomial_atan_error_bound(hostsim):
    """atan_nonneg (the pitch of the screen's ray) against fp64 atan: <= 2.5e-7 rad everywhere the set-up
    can land (0 .. 1e30, including exactly 1 and the range-reduction seam), monotone enough to be a speed law."""
    import ctypes as C

    x = np.concatenate([np.linspace(0.0, 1.0, 400001), np.linspace(1.0, 64.0, 400001), np.logspace(-30, 30, 200001),
                        np.array([0.0, 1.0, np.nextafter(np.float32(1.0), np.float32(2.0)), 1e38])]).astype(np.float32)
    out = np.empty_like(x)
    hostsim.L.hs_atan_nonneg(x.ctypes.data_as(C.c_void_p), C.c_long(len(x)), out.ctypes.data_as(C.c_void_p))
    err = np.abs(out.astype(np.float64) - np.arctan(x.astype(np.float64)))
    assert err.max() <= 2.5e-7, err.max()
    assert out[0] == 0.0 and abs(float(out[-1]) - np.pi / 2) < 2e-7


def test_fp32_screen_levels_are_conservative_and_accurate(hostsim, golden, tracks, ring1):
    """Level 1 (horizontal, incremental) never exceeds level 2 (3-D, fused) beyond rounding, level 2
    agrees with the fp64 distance on every golden near miss (min in stage 2, < 25 m), and every trial
    the reference counts as a hit in stage 2 passes both levels with room to spare.  The engine's margin
    is 0.02 m + sqrt(3)*(tile/2 + 4) ulp32(extent) ~ 0.08 m."""
    worst3 = 0.0
    passed1 = total = 0
    for p, (a, j) in enumerate(golden["pairs"]):
        t = tracks[a]
        draws = golden["pair%d_draws" % p]
        mind = golden["pair%d_min_dist" % p]
        first = golden["pair%d_first_hit" % p]
        origin = (np.floor(0.5 * (671819.02 + 678600.0)), np.floor(0.5 * (5704300.0 + 5706300.0)), 250.0)
        s2, s3 = hostsim.screen(t, ring1[j], draws, origin)
        s2 = np.sqrt(s2.astype(np.float64))
        s3 = np.sqrt(s3.astype(np.float64))
        assert (s2 <= s3 + 2e-2).all()                       # horizontal <= 3-D (up to accumulation error)
        near = mind < 25.0
        # the screen only covers stage 2; where the fp64 minimum sits in stage 1 the screen is >= it
        assert (s3[near] >= mind[near] - 5e-3).all()
        same = near & (np.abs(s3 - mind) < 1.0)
        assert same.sum() > 0.9 * near.sum()
        worst3 = max(worst3, np.abs(s3[same] - mind[same]).max())
        hit2 = (first >= draws["t1st"])                      # reference hit, found in stage 2
        assert (s2[hit2] < 3.69 + 2e-2).all() and (s3[hit2] < 3.69 + 5e-3).all()
        passed1 += int((s2 < 3.78).sum())
        total += len(draws)
    print("worst level-2 error %.2e m; level 1 passes %d of %d golden (near-miss-enriched) trials" % (
        worst3, passed1, total))
    assert worst3 < 5e-3


def test_slab_level_is_conservative_and_within_its_margin(hostsim, golden, tracks, ring1):
    """Level 1a of k_screen3 (drone-collision-rp_b200/csrc/screen3.cu), the statements of dcrp_math.cuh::slab_ray32 and
    the incremental projected scan, on the golden (near-miss-enriched) trials and 16 directions:
      * exactness: |n . D(i)| <= |D_xy(i)|, so the slab minimum never exceeds the 2-D level-1 minimum (same trials,
        a superset of the indices) -- beyond fp32 rounding;
      * every trial the reference counts as a stage-2 hit is flagged with room to spare;
      * accuracy: against the fp64 value of min_i |n . D(i)| the fp32 scan is within the margin the engine adds to the
        radius (engine.cu: 0.02 + 76 ulp32(sqrt2 * extent) + n_steps * vmax * 2^-20 ~ 0.10 m at 8 km extent)."""
    origin = (np.floor(0.5 * (671819.02 + 678600.0)), np.floor(0.5 * (5704300.0 + 5706300.0)), 250.0)
    worst = 0.0
    for p, (a, j) in enumerate(golden["pairs"]):
        t = tracks[a]
        draws = golden["pair%d_draws" % p][:4000]
        first = golden["pair%d_first_hit" % p][:4000]
        s2, _ = hostsim.screen(t, ring1[j], draws, origin)
        s2 = np.sqrt(s2.astype(np.float64))
        # fp64 restatement of the projected separation at every index >= 1 (ray extrapolated backwards before the kink)
        x0, y0, h = ring1[j]
        last = draws["t1st"].astype(np.int64) - 1
        kx = x0 + np.cumsum(np.r_[0.0, np.full(t["takeoff_t"] - 1, np.sin(h) * 19.0)])[last]
        ky = y0 + np.cumsum(np.r_[0.0, np.full(t["takeoff_t"] - 1, np.cos(h) * 19.0)])[last]
        dl, db = draws["tx"] - kx, draws["ty"] - ky
        m = np.hypot(dl, db)
        vf = 19.0 - (2.0 * (19.0 - 8.0) / np.pi) * np.arctan(np.abs(draws["tz"] - 100.0) / m)
        idx = np.arange(1, t["n_steps"])
        rel = idx[None, :] - last[:, None]
        dx = kx[:, None] + rel * (dl / m * vf)[:, None] - t["x"][None, 1:]
        dy = ky[:, None] + rel * (db / m * vf)[:, None] - t["y"][None, 1:]
        hit2 = first >= draws["t1st"]
        for k in range(16):
            n = (np.cos(np.pi * k / 16), np.sin(np.pi * k / 16))
            got = hostsim.slab(t, ring1[j], draws, origin, n).astype(np.float64)
            want = np.abs(n[0] * dx + n[1] * dy).min(axis=1)
            worst = max(worst, np.abs(got - want).max())
            assert (got <= s2 + 2e-2).all(), (p, k)                    # never above the 2-D level
            assert (got[hit2] < 3.69 + 2e-2).all(), (p, k)             # reference hits are always flagged
    print("worst |fp32 slab - fp64| = %.2e m" % worst)
    assert worst < 0.08


def test_substep_slab_threshold_is_conservative(hostsim, oracle, golden, tracks, ring1):
    """Level 1a of k_screen3<., SUB> (DCRP_FLAG_SUBSTEP): on a segment in conflict |n . D| at ONE of its two ends is
    below R + half the projected relative step, so the engine compares the slab minimum -- taken over every sample,
    index 0 included -- against thr_slab_sub = R + slab margin + (max aircraft step + vmax) / 2 (engine.cu).  Checked
    against the oracle's sub-step decisions (stage-2 conflicts; a stage-1 segment in conflict forces the trial) for
    the reference's radius and for a 40 m one, on 16 directions: every conflict is flagged."""
    origin = (np.floor(0.5 * (671819.02 + 678600.0)), np.floor(0.5 * (5704300.0 + 5706300.0)), 250.0)
    n_conf = 0
    for p, (a, j) in enumerate(golden["pairs"][:6]):
        t = tracks[a]
        draws = golden["pair%d_draws" % p][:3000]
        amax = np.hypot(np.diff(t["x"]), np.diff(t["y"])).max()
        x0, y0, h = ring1[j]
        last = draws["t1st"].astype(np.int64) - 1
        # the ray at index 0 in fp64 (the fp32 scan of hostsim.slab starts at index 1, the device's sub-step build at 0)
        kx = x0 + np.cumsum(np.r_[0.0, np.full(t["takeoff_t"] - 1, np.sin(h) * 19.0)])[last]
        ky = y0 + np.cumsum(np.r_[0.0, np.full(t["takeoff_t"] - 1, np.cos(h) * 19.0)])[last]
        dl, db = draws["tx"] - kx, draws["ty"] - ky
        m = np.hypot(dl, db)
        vf = 19.0 - (2.0 * (19.0 - 8.0) / np.pi) * np.arctan(np.abs(draws["tz"] - 100.0) / m)
        d0x = kx - last * (dl / m * vf) - t["x"][0]
        d0y = ky - last * (db / m * vf) - t["y"][0]
        for r_ac in (3.5, 40.0):
            om = type(oracle.model)()
            oracle.L.orc_default_model(om)
            om.aircraft_radius = r_ac
            radius = r_ac + 0.19
            hit, first_seg, _, _ = oracle.run_trials_substep(t, ring1[j], draws, model=om)
            stage2 = (hit == 1) & (first_seg >= last)
            thr = radius + 0.10 + 0.5 * (amax + 19.0) * (1 + 1e-6)
            for k in range(16):
                n = (np.cos(np.pi * k / 16), np.sin(np.pi * k / 16))
                got = hostsim.slab(t, ring1[j], draws, origin, n).astype(np.float64)
                got = np.minimum(got, np.abs(n[0] * d0x + n[1] * d0y))
                assert (got[stage2] < thr).all(), (p, r_ac, k, got[stage2].max(), thr)
            n_conf += int(stage2.sum())
    assert n_conf > 200


def test_substep_level2_far_segment_skip_is_safe(oracle, golden, tracks, ring1):
    """Level 2 of the sub-step mode (kernels.cu min3d_seg_fp32) skips a segment whose two ends are both farther than
    thr2_near_sub = ((R + margin) + E / 2)^2 with E = longest 3-D aircraft step + sqrt(2) vmax (engine.cu): on
    reference-exact trajectories (oracle.trajectory) the relative step of every stage-2 segment is within E, and no
    skipped segment's closest approach comes within the level-2 radius."""
    n_skipped = n_near = 0
    for p, (a, j) in enumerate(golden["pairs"][:6]):
        t = tracks[a]
        amax3 = np.sqrt(np.diff(t["x"]) ** 2 + np.diff(t["y"]) ** 2 + np.diff(t["z"]) ** 2).max()
        E = amax3 + 1.4143 * 19.0
        for r_ac in (3.5, 60.0):
            level2_radius = r_ac + 0.19 + 0.10                       # R + margin (+ the 4e-6 (R + E)^2 fp32 term, negligible)
            near = (level2_radius + 0.5 * E) * (1 + 1e-4) + 0.05
            for d in golden["pair%d_draws" % p][:150]:
                x, y, z = oracle.trajectory(t, ring1[j], d)
                D = np.stack([x - t["x"], y - t["y"], z - t["z"]], axis=1)
                last = int(d["t1st"]) - 1
                D0, D1 = D[last:-1], D[last + 1:]
                Ev = D1 - D0
                assert np.sqrt((Ev ** 2).sum(1)).max() <= E + 1e-9
                aa, bb = (Ev ** 2).sum(1), (D0 * Ev).sum(1)
                tt = np.clip(-bb / np.where(aa > 0, aa, 1.0), 0.0, 1.0)
                closest = np.sqrt(((D0 + tt[:, None] * Ev) ** 2).sum(1))
                ends = np.minimum(np.sqrt((D0 ** 2).sum(1)), np.sqrt((D1 ** 2).sum(1)))
                skipped = ends >= near
                assert (closest[skipped] >= level2_radius).all()
                n_skipped += int(skipped.sum()); n_near += int((~skipped).sum())
    assert n_skipped > 10 * n_near > 0          # even on these near-miss-enriched trials all but a few segments per trial are far
