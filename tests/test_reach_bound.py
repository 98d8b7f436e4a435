"""CPU tier: the inequality DCRP_FLAG_REACH_CULL rests on, checked on the ORACLE's trajectories (the
restatement pinned against the unmodified reference), independently of any GPU code.

After the kink at index L the drone's step is (sin b * vf, cos b * vf, sin p * vf) with
vf = v_straight - coef * pitch and pitch in [0, pi/2] (Drone.cpp:218-227), so at every index i > L
    |P_i - K|_xy <= (i - L) * vmax          and          start_alt <= z_i <= start_alt + (i - L) * vmax,
vmax = max(v_straight, v_ascend).  k_reach (csrc/kernels.cu) clears a (track, drone, t1st, i) when the aircraft
sample i is outside that set grown by R + 1 cm; this test shows the trajectories never leave the set (with
eight orders of magnitude to spare on the 1 cm) for the shipped constants and for others."""
import numpy as np
import pytest

import oracle_api as oa
from drone_collision_rp_b200 import capi


@pytest.mark.parametrize("vs,va,alt", [(19.0, 8.0, 100.0), (10.0, 25.0, 100.0), (30.0, 0.5, 30.0)])
def test_trajectories_stay_inside_the_reach_set(oracle, tracks, ring1, ring5, vs, va, alt):
    m = oa.Model()
    m.aircraft_radius, m.drone_radius, m.start_alt, m.v_straight, m.v_ascend = 3.5, 0.19, alt, vs, va
    vmax = max(vs, va)
    worst_xy = worst_z = -np.inf
    rng = np.random.default_rng(7)
    for ring in (ring1, ring5):
        for t in tracks[:3]:
            n = len(t["x"])
            box = capi.target_box(t["y"][0], t["z"][0])
            for j in rng.integers(0, 99, 4):
                draws = oracle.draws(123, int(rng.integers(0, 2**40)), 60, int(j), 0, t["takeoff_t"], box)
                for d in draws:
                    x, y, z = oracle.trajectory(t, ring[j], d, model=m)
                    L = int(d["t1st"]) - 1
                    i = np.arange(L + 1, n)
                    if len(i) == 0:
                        continue
                    r_xy = np.hypot(x[i] - x[L], y[i] - y[L])
                    worst_xy = max(worst_xy, float(np.max(r_xy - (i - L) * vmax)))
                    worst_z = max(worst_z, float(np.max(z[i] - (alt + (i - L) * vmax))), float(np.max(alt - z[i])))
                    # stage 1 is the straight run at the start altitude (what k_stage1 tabulates)
                    assert np.all(z[:L + 1] == alt)
    assert worst_xy <= 1e-9, worst_xy          # metres beyond the bound (the kernel allows 1e-2)
    assert worst_z <= 1e-9, worst_z
