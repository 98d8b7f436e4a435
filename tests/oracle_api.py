"""ctypes access to the CHECKERS: oracle/_build/liboracle.so (C restatement),
oracle/_ref/*.so (unmodified reference + replay shim) and tests/hostsim.
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIBORACLE = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
LIBREF_REPLAY = os.path.join(ORACLE_DIR, "_ref", "libref_replay.so")
LIBREF_AIRCRAFT = os.path.join(ORACLE_DIR, "_ref", "libref_aircraft.so")
REF_ASSHIPPED = os.path.join(ORACLE_DIR, "_ref", "ref_asshipped")
HOSTSIM_SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
LIBHOSTSIM = os.path.join(ROOT, "tests", "hostsim", "_build", "libhostsim.so")
CSRC = os.path.join(ROOT, "drone-collision-rp_b200", "csrc")

DRAW_DTYPE = np.dtype([("t1st", "<i4"), ("reserved", "<i4"), ("tx", "<f8"), ("ty", "<f8"), ("tz", "<f8")])
dp = C.POINTER(C.c_double)


class Model(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("aircraft_radius", "drone_radius", "start_alt", "v_straight", "v_ascend")]


def _d(a):
    return a.ctypes.data_as(dp)


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


def build_hostsim():
    os.makedirs(os.path.dirname(LIBHOSTSIM), exist_ok=True)
    srcs = [HOSTSIM_SRC, os.path.join(CSRC, "dcrp_math.cuh")]
    if os.path.exists(LIBHOSTSIM) and all(os.path.getmtime(LIBHOSTSIM) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-I", CSRC,
                    HOSTSIM_SRC, "-o", LIBHOSTSIM], check=True, capture_output=True)


class Oracle:
    def __init__(self):
        if not os.path.exists(LIBORACLE):
            build_oracle()
        L = C.CDLL(LIBORACLE)
        L.orc_default_model.argtypes = [C.POINTER(Model)]
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.orc_make_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, dp, C.c_void_p]
        L.orc_make_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32, C.c_int32, dp, C.c_void_p]
        L.orc_target_box.argtypes = [C.c_double, C.c_double, dp]
        L.orc_run_trials.restype = C.c_int64
        L.orc_run_trials.argtypes = [dp, dp, dp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                     C.c_int64, C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_run_philox.restype = C.c_int64
        L.orc_run_philox.argtypes = [dp, dp, dp, C.c_int32, C.c_int32, dp, C.c_double, C.c_double, C.c_double,
                                     C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int64, C.POINTER(Model),
                                     C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_run_trials_substep.restype = C.c_int64
        L.orc_run_trials_substep.argtypes = [dp, dp, dp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                             C.c_int64, C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        L.orc_trial.argtypes = [dp, dp, dp, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                C.POINTER(Model), dp, dp, dp, C.c_void_p]
        L.orc_format_block.restype = C.c_int64
        L.orc_format_block.argtypes = [dp, dp, dp, C.c_int32, C.c_int32, C.c_char_p, C.c_int64]
        self.L = L
        self.model = Model()
        L.orc_default_model(C.byref(self.model))

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.L.orc_philox4x32_10(c, k, o)
        return list(o)

    def target_box(self, ay0, az0):
        box = (C.c_double * 6)()
        rc = self.L.orc_target_box(float(ay0), float(az0), box)
        return rc, np.array(box[:])

    def draws(self, seed, trial_begin, n, drone, track, takeoff_t, box):
        out = np.zeros(n, dtype=DRAW_DTYPE)
        b = np.ascontiguousarray(box, dtype=np.float64)
        self.L.orc_make_draws(int(seed), int(trial_begin), int(n), int(drone), int(track), int(takeoff_t), _d(b),
                              out.ctypes.data)
        return out

    def run_trials(self, trk, drone, draws, model=None):
        n = len(draws)
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        hit, first, mind = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.float64)
        m = model or self.model
        self.L.orc_run_trials(_d(x), _d(y), _d(z), len(x), float(drone[0]), float(drone[1]), float(drone[2]),
                              draws.ctypes.data, n, C.byref(m), hit.ctypes.data, first.ctypes.data, mind.ctypes.data)
        return hit, first, mind

    def run_trials_substep(self, trk, drone, draws, model=None):
        """EXTENSION (no reference counterpart): interpolated separation, see dcrp_oracle.h."""
        n = len(draws)
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        hit, first = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        ftime, mind = np.zeros(n, np.float64), np.zeros(n, np.float64)
        m = model or self.model
        self.L.orc_run_trials_substep(_d(x), _d(y), _d(z), len(x), float(drone[0]), float(drone[1]), float(drone[2]),
                                      draws.ctypes.data, n, C.byref(m), hit.ctypes.data, first.ctypes.data,
                                      ftime.ctypes.data, mind.ctypes.data)
        return hit, first, ftime, mind

    def run_philox(self, trk, box, drone, seed, drone_id, track_id, trial_begin, n, cap=4096, model=None):
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        b = np.ascontiguousarray(box, dtype=np.float64)
        ids = np.zeros(cap, np.uint64)
        nid = C.c_int64()
        m = model or self.model
        hits = self.L.orc_run_philox(_d(x), _d(y), _d(z), len(x), int(trk["takeoff_t"]), _d(b), float(drone[0]),
                                     float(drone[1]), float(drone[2]), int(seed), int(drone_id), int(track_id),
                                     int(trial_begin), int(n), C.byref(m), ids.ctypes.data, cap, C.byref(nid))
        return hits, ids[:nid.value].copy()

    def trajectory(self, trk, drone, draw, model=None):
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        n = len(x)
        tx, ty, tz = np.zeros(n), np.zeros(n), np.zeros(n)
        d = np.ascontiguousarray(np.array([draw], dtype=DRAW_DTYPE))
        res = (C.c_int32 * 4)()
        m = model or self.model
        self.L.orc_trial(_d(x), _d(y), _d(z), n, float(drone[0]), float(drone[1]), float(drone[2]),
                         d.ctypes.data, C.byref(m), _d(tx), _d(ty), _d(tz), C.cast(res, C.c_void_p))
        return tx, ty, tz

    def format_block(self, x, y, z, run_no):
        x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
        cap = 64 * len(x) + 16
        buf = C.create_string_buffer(cap)
        n = self.L.orc_format_block(_d(x), _d(y), _d(z), len(x), int(run_no), buf, cap)
        assert n >= 0
        return buf.raw[:n].decode()


class RefReplay:
    """The unmodified reference Drone.cpp under the replay shim (oracle/_ref)."""

    @staticmethod
    def available():
        return os.path.exists(LIBREF_REPLAY)

    def __init__(self):
        L = C.CDLL(LIBREF_REPLAY)
        L.ref_replay_run.restype = C.c_long
        L.ref_replay_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp,
                                     C.c_double, C.c_double, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_long)]
        L.ref_replay_simulation.restype = C.c_double
        L.ref_replay_simulation.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp,
                                            C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int]
        L.ref_cubed_volume.argtypes = [C.c_double, C.c_double, dp]
        self.L = L

    def run(self, ring_csv, drone_index, aircraft_index, trk, draws, r_ac=3.5, r_dr=0.19, want_traj=False):
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        n = len(draws)
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        hit, first, mind = np.zeros(n, np.uint8), np.zeros(n, np.int32), np.zeros(n, np.float64)
        traj = np.zeros((n, 3, len(x))) if want_traj else None
        viol = C.c_long()
        hits = self.L.ref_replay_run(ring_csv.encode(), 4, int(drone_index), int(aircraft_index), len(x),
                                     int(trk["takeoff_t"]), _d(x), _d(y), _d(z), r_ac, r_dr, draws.ctypes.data, n,
                                     hit.ctypes.data, first.ctypes.data, mind.ctypes.data,
                                     traj.ctypes.data if want_traj else None, C.byref(viol))
        return {"hits": hits, "hit": hit, "first_hit": first, "min_dist": mind, "traj": traj,
                "range_violations": viol.value}

    def simulation(self, ring_csv, drone_index, aircraft_index, trk, draws, clear_output, r_ac=3.5, r_dr=0.19):
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        return self.L.ref_replay_simulation(ring_csv.encode(), 4, int(drone_index), int(aircraft_index), len(x),
                                            int(trk["takeoff_t"]), _d(x), _d(y), _d(z), r_ac, r_dr,
                                            draws.ctypes.data, len(draws), int(clear_output))

    def cubed_volume(self, ay0, az0):
        box = (C.c_double * 6)()
        self.L.ref_cubed_volume(float(ay0), float(az0), box)
        return np.array(box[:])


class RefAircraft:
    @staticmethod
    def available():
        return os.path.exists(LIBREF_AIRCRAFT)

    def __init__(self):
        L = C.CDLL(LIBREF_AIRCRAFT)
        L.ref_aircraft_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int), dp, dp, dp, dp, C.c_int]
        self.L = L

    def load(self, path, index, n_cols=22, cap=4096):
        vl, tk, size = C.c_int(), C.c_int(), C.c_int()
        x, y, z, gs = (np.zeros(cap) for _ in range(4))
        n = self.L.ref_aircraft_load(path.encode(), n_cols, int(index), C.byref(vl), C.byref(tk), C.byref(size),
                                     _d(x), _d(y), _d(z), _d(gs), cap)
        return {"x": x[:n].copy(), "y": y[:n].copy(), "z": z[:n].copy(), "gs": gs[:n].copy(),
                "takeoff_t": tk.value, "n_steps": vl.value, "index_size": size.value}


class HostSim:
    def __init__(self):
        build_hostsim()
        L = C.CDLL(LIBHOSTSIM)
        L.hs_philox.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.hs_make_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, dp, C.c_void_p]
        L.hs_thr_d2.restype = C.c_double
        L.hs_thr_d2.argtypes = [C.c_double]
        L.hs_exact.argtypes = [dp, dp, dp, C.c_int, C.c_int] + [C.c_double] * 7 + [C.c_void_p, C.c_long,
                                                                                   C.c_void_p, C.c_void_p]
        L.hs_screen.argtypes = [dp, dp, dp, C.c_int, C.c_int] + [C.c_double] * 9 + [C.c_void_p, C.c_long, C.c_void_p,
                                                                                    C.c_void_p]
        L.hs_slab.argtypes = [dp, dp, C.c_int, C.c_int] + [C.c_double] * 10 + [C.c_void_p, C.c_long, C.c_void_p]
        self.L = L

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.L.hs_philox(c, k, o)
        return list(o)

    def slab(self, trk, drone, draws, origin, direction, start_alt=100.0, v_straight=19.0, v_ascend=8.0):
        """min over i >= 1 of |n . D(i)| in the fp32 arithmetic of k_screen3's level 1a."""
        x, y = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xy")
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        out = np.zeros(len(draws), np.float32)
        self.L.hs_slab(_d(x), _d(y), len(x), int(trk["takeoff_t"]), float(drone[0]), float(drone[1]), float(drone[2]),
                       start_alt, v_straight, v_ascend, float(origin[0]), float(origin[1]), float(direction[0]),
                       float(direction[1]), draws.ctypes.data, len(draws), out.ctypes.data)
        return out

    def draws(self, seed, trial_begin, n, drone, track, takeoff_t, box):
        out = np.zeros(n, dtype=DRAW_DTYPE)
        b = np.ascontiguousarray(box, dtype=np.float64)
        for r in range(n):
            self.L.hs_make_draw(int(seed), int(trial_begin + r), int(drone), int(track), int(takeoff_t), _d(b),
                                out.ctypes.data + r * 32)
        return out

    def exact(self, trk, drone, draws, radius=3.69, start_alt=100.0, v_straight=19.0, v_ascend=8.0):
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        n = len(draws)
        first, mind = np.zeros(n, np.int32), np.zeros(n, np.float64)
        self.L.hs_exact(_d(x), _d(y), _d(z), len(x), int(trk["takeoff_t"]), float(drone[0]), float(drone[1]),
                        float(drone[2]), radius, start_alt, v_straight, v_ascend, draws.ctypes.data, n,
                        first.ctypes.data, mind.ctypes.data)
        return first, mind

    def screen(self, trk, drone, draws, origin, start_alt=100.0, v_straight=19.0, v_ascend=8.0):
        x, y, z = (np.ascontiguousarray(trk[k], dtype=np.float64) for k in "xyz")
        draws = np.ascontiguousarray(draws, dtype=DRAW_DTYPE)
        n = len(draws)
        out2, out3 = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self.L.hs_screen(_d(x), _d(y), _d(z), len(x), int(trk["takeoff_t"]), float(drone[0]), float(drone[1]),
                         float(drone[2]), start_alt, v_straight, v_ascend, float(origin[0]), float(origin[1]),
                         float(origin[2]), draws.ctypes.data, n, out2.ctypes.data, out3.ctypes.data)
        return out2, out3
