"""ctypes binding of include/dcrp.h and include/dcrp_host.h.

Loading fails loudly when the native libraries are missing; nothing here can
compute a trial.
"""
import ctypes as C
import os

import numpy as np

from . import LIBDCRP, LIBDCRP_BENCH, LIBDCRP_HOST

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_RUNWAY, ERR_STATE, ERR_IO, ERR_COMM = -1, -2, -3, -4, -5, -6, -7
COMM_NCCL_ONLY = 1
FP64, MIXED, FP32 = 0, 1, 2
FLAG_PER_TRIAL = 1
FLAG_SUBSTEP = 2
FLAG_REACH_CULL = 4
FLAG_SCREEN_2D = 8
FLAG_ZERO_COUNTS = 16
FLAG_NO_TIMING = 32
STAT_SCENARIO_CACHED, STAT_FP64_FALLBACK, STAT_CULL_IGNORED = 1, 2, 4
INT32_MAX = 2**31 - 1

c_double_p = C.POINTER(C.c_double)


class Track(C.Structure):
    _fields_ = [("x", c_double_p), ("y", c_double_p), ("z", c_double_p),
                ("n_steps", C.c_int32), ("takeoff_t", C.c_int32), ("box", C.c_double * 6)]


class Drone(C.Structure):
    _fields_ = [("x0", C.c_double), ("y0", C.c_double), ("heading0", C.c_double)]


class Draw(C.Structure):
    _fields_ = [("t1st", C.c_int32), ("reserved", C.c_int32),
                ("tx", C.c_double), ("ty", C.c_double), ("tz", C.c_double)]


DRAW_DTYPE = np.dtype([("t1st", "<i4"), ("reserved", "<i4"), ("tx", "<f8"), ("ty", "<f8"), ("tz", "<f8")])
RECORD_DTYPE = np.dtype([("track", "<u4"), ("drone", "<u4"), ("trial", "<u8"), ("t1st", "<i4"),
                         ("first_hit", "<i4"), ("tx", "<f8"), ("ty", "<f8"), ("tz", "<f8"),
                         ("min_dist", "<f8"), ("first_time", "<f8")])
assert DRAW_DTYPE.itemsize == 32 and RECORD_DTYPE.itemsize == 64


class Model(C.Structure):
    _fields_ = [("aircraft_radius", C.c_double), ("drone_radius", C.c_double),
                ("start_alt", C.c_double), ("v_straight", C.c_double), ("v_ascend", C.c_double)]


class Run(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("trial_begin", C.c_uint64), ("trial_count", C.c_uint64),
                ("precision", C.c_int32), ("flags", C.c_uint32), ("record_capacity", C.c_uint32),
                ("track_id_base", C.c_uint32), ("drone_id_base", C.c_uint32), ("reserved", C.c_uint32),
                ("replay", C.c_void_p)]


class AllPairsRun(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("ray_begin", C.c_uint64), ("ray_count", C.c_uint64),
                ("ref_takeoff_t", C.c_int32), ("record_capacity", C.c_uint32), ("drone_id_base", C.c_uint32),
                ("reserved", C.c_uint32), ("ref_box", C.c_double * 6)]


class Stats(C.Structure):
    _fields_ = [("trials", C.c_uint64), ("tests_evaluated", C.c_uint64), ("tests_shared", C.c_uint64),
                ("tests_issued", C.c_uint64), ("drone_steps", C.c_uint64), ("level2", C.c_uint64),
                ("candidates", C.c_uint64),
                ("hits", C.c_uint64), ("kernel_launches", C.c_uint32), ("status", C.c_uint32),
                ("ms_prepare", C.c_float), ("ms_trials", C.c_float), ("ms_confirm", C.c_float),
                ("ms_device_total", C.c_float), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("trials_culled", C.c_uint64), ("tests_culled", C.c_uint64), ("slab_survivors", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class Result(C.Structure):
    _fields_ = [("counts", C.POINTER(C.c_uint64)), ("first_conflict", C.POINTER(C.c_int32)),
                ("records", C.c_void_p), ("n_records", C.c_uint32), ("overflow", C.c_int32),
                ("trial_hit", C.POINTER(C.c_uint8)), ("trial_first_hit", C.POINTER(C.c_int32)),
                ("trial_min_dist", c_double_p), ("stats", Stats)]


class DcrpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("dcrp status %d: %s" % (status, message))
        self.status = status


_lib = None
_host = None

# every symbol include/dcrp.h declares (tests check the library exports them all)
DCRP_SYMBOLS = [
    "dcrp_abi_version", "dcrp_last_error", "dcrp_default_model", "dcrp_engine_create",
    "dcrp_engine_destroy", "dcrp_target_box", "dcrp_simulate", "dcrp_scenario_upload",
    "dcrp_run_async", "dcrp_fetch", "dcrp_last_run_ms", "dcrp_allpairs", "dcrp_dump_draws", "dcrp_trajectories",
    "dcrp_comm_unique_id", "dcrp_engine_attach_comm", "dcrp_engine_detach_comm", "dcrp_allreduce_counts", "dcrp_comm_info",
    "dcrp_device_count", "dcrp_engine_reserve_sms", "dcrp_launch_count", "dcrp_device_info",
    "dcrp_debug_block_schedule",
]
DCRP_HOST_SYMBOLS = [
    "dcrph_last_error", "dcrph_reference_aircraft_csv", "dcrph_synth_day_write", "dcrph_synth_day_write_mixed",
    "dcrph_ring_write",
    "dcrph_aircraft_load",
    "dcrph_ring_row", "dcrph_monte_carlo", "dcrph_monte_carlo_batch", "dcrph_set_run_flags", "dcrph_set_gpus", "dcrph_write_block",
]


def lib():
    """libdcrp.so (CUDA engine).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBDCRP):
            raise FileNotFoundError(
                LIBDCRP + " is missing: build it (python -c 'import __graft_entry__ as g; g.build()'); "
                "there is no CPU fallback")
        L = C.CDLL(LIBDCRP, mode=C.RTLD_GLOBAL)
        L.dcrp_last_error.restype = C.c_char_p
        L.dcrp_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.dcrp_engine_destroy.argtypes = [C.c_void_p]
        L.dcrp_default_model.argtypes = [C.POINTER(Model)]
        L.dcrp_default_model.restype = None
        L.dcrp_target_box.argtypes = [C.c_double, C.c_double, c_double_p]
        L.dcrp_simulate.argtypes = [C.c_void_p, C.POINTER(Track), C.c_int32, C.POINTER(Drone), C.c_int32,
                                    C.POINTER(Model), C.POINTER(Run), C.POINTER(Result)]
        L.dcrp_scenario_upload.argtypes = [C.c_void_p, C.POINTER(Track), C.c_int32, C.POINTER(Drone),
                                           C.c_int32, C.POINTER(Model)]
        L.dcrp_run_async.argtypes = [C.c_void_p, C.POINTER(Run), C.c_void_p, C.c_void_p]
        L.dcrp_fetch.argtypes = [C.c_void_p, C.POINTER(Result)]
        L.dcrp_allpairs.argtypes = [C.c_void_p, C.POINTER(AllPairsRun), C.POINTER(Result)]
        L.dcrp_last_run_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.dcrp_dump_draws.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Run), C.c_void_p]
        L.dcrp_trajectories.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, c_double_p, C.POINTER(C.c_int32)]
        L.dcrp_comm_unique_id.argtypes = [C.c_void_p]
        L.dcrp_engine_attach_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
        L.dcrp_engine_detach_comm.argtypes = [C.c_void_p]
        L.dcrp_allreduce_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.dcrp_comm_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.dcrp_device_count.argtypes = [C.POINTER(C.c_int)]
        L.dcrp_engine_reserve_sms.argtypes = [C.c_void_p, C.c_int]
        L.dcrp_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.dcrp_device_info.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def host():
    """libdcrp_host.so (C++ Aircraft / Drone / Monte_Carlo mirror)."""
    global _host
    if _host is None:
        lib()
        if not os.path.exists(LIBDCRP_HOST):
            raise FileNotFoundError(LIBDCRP_HOST + " is missing: build it first")
        H = C.CDLL(LIBDCRP_HOST)
        H.dcrph_last_error.restype = C.c_char_p
        H.dcrph_reference_aircraft_csv.restype = C.c_char_p
        H.dcrph_synth_day_write.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_int]
        H.dcrph_synth_day_write_mixed.argtypes = [C.c_char_p, C.c_int, C.c_uint64, C.c_int, C.c_int]
        H.dcrph_ring_write.argtypes = [C.c_char_p, C.c_double, C.c_int]
        H.dcrph_aircraft_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int), c_double_p, c_double_p, c_double_p, c_double_p,
                                          C.c_int]
        H.dcrph_ring_row.argtypes = [C.c_char_p, C.c_int, C.c_int, c_double_p]
        H.dcrph_monte_carlo.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                        C.c_int, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_char_p,
                                        c_double_p, C.POINTER(C.c_uint64)]
        H.dcrph_monte_carlo_batch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                              C.c_uint64, C.c_double, C.c_double, C.c_uint64, C.c_int, C.c_char_p,
                                              C.POINTER(C.c_uint64)]
        H.dcrph_set_run_flags.argtypes = [C.c_uint32]
        H.dcrph_set_run_flags.restype = None
        H.dcrph_set_gpus.argtypes = [C.c_int]
        H.dcrph_set_gpus.restype = None
        H.dcrph_write_block.argtypes = [C.c_char_p, c_double_p, c_double_p, c_double_p, C.c_int, C.c_long]
        _host = H
    return _host


_bench = None


def bench_lib():
    """libdcrp_bench.so: roofline micro-kernels and the L2 flush (measurement only)."""
    global _bench
    if _bench is None:
        if not os.path.exists(LIBDCRP_BENCH):
            raise FileNotFoundError(LIBDCRP_BENCH + " is missing: build it first")
        B = C.CDLL(LIBDCRP_BENCH)
        B.dcrpb_last_error.restype = C.c_char_p
        B.dcrpb_measure_peak.argtypes = [C.c_int, C.c_int, c_double_p, C.POINTER(C.c_float)]
        B.dcrpb_flush_l2.argtypes = [C.c_int, C.c_void_p]
        _bench = B
    return _bench


def measure_peak(device, which):
    tf, ms = C.c_double(), C.c_float()
    if bench_lib().dcrpb_measure_peak(int(device), int(which), C.byref(tf), C.byref(ms)) != 0:
        raise RuntimeError(bench_lib().dcrpb_last_error().decode())
    return tf.value, ms.value


def flush_l2(device, stream=None):
    if bench_lib().dcrpb_flush_l2(int(device), C.c_void_p(stream or 0)) != 0:
        raise RuntimeError(bench_lib().dcrpb_last_error().decode())


def _check(status):
    if status < 0:
        raise DcrpError(status, lib().dcrp_last_error().decode())
    return status


def _hcheck(status):
    if status < 0:
        raise DcrpError(status, host().dcrph_last_error().decode())
    return status


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def default_model():
    m = Model()
    lib().dcrp_default_model(C.byref(m))
    return m


def target_box(ay0, az0):
    box = (C.c_double * 6)()
    _check(lib().dcrp_target_box(float(ay0), float(az0), box))
    return np.array(box[:], dtype=np.float64)


class RunResult:
    """Host-side view of a dcrp_result."""

    def __init__(self, counts, first_conflict, records, overflow, stats, per_trial=None):
        self.counts = counts
        self.first_conflict = first_conflict
        self.records = records
        self.overflow = overflow
        self.stats = stats
        self.per_trial = per_trial

    @property
    def hits(self):
        return int(self.counts.sum())


class Scenario:
    """Tracks + drones marshalled once into the C structs of include/dcrp.h (keeps the
    numpy arrays the structs point into alive).  Pass it wherever (tracks, drones) go."""

    def __init__(self, tracks, drones):
        self.T, self.D, self._keep = Engine._pack(tracks, drones)
        self.n_tracks, self.n_drones = len(self.T), len(self.D)


class Engine:
    """One dcrp_engine on one CUDA device."""

    def __init__(self, device=0):
        self.device = int(device)
        self._h = C.c_void_p()
        _check(lib().dcrp_engine_create(int(device), C.byref(self._h)))
        self.n_tracks = self.n_drones = 0
        self._keep = None

    def close(self):
        if self._h:
            lib().dcrp_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- marshalling
    @staticmethod
    def _pack(tracks, drones):
        """tracks: list of dict(x,y,z,takeoff_t[,box]); drones: (n,3) array x0,y0,heading0."""
        keep = []
        T = (Track * len(tracks))()
        for i, t in enumerate(tracks):
            x = np.ascontiguousarray(t["x"], dtype=np.float64)
            y = np.ascontiguousarray(t["y"], dtype=np.float64)
            z = np.ascontiguousarray(t["z"], dtype=np.float64)
            keep += [x, y, z]
            T[i].x, T[i].y, T[i].z = dptr(x), dptr(y), dptr(z)
            T[i].n_steps = len(x)
            T[i].takeoff_t = int(t["takeoff_t"])
            box = t.get("box")
            if box is None:
                box = target_box(y[0], z[0])
            for k in range(6):
                T[i].box[k] = float(box[k])
        d = np.ascontiguousarray(drones, dtype=np.float64).reshape(-1, 3)
        D = (Drone * len(d))()
        for j in range(len(d)):
            D[j].x0, D[j].y0, D[j].heading0 = d[j]
        return T, D, keep

    @staticmethod
    def _run_struct(seed, trial_begin, trial_count, precision, flags=0, record_capacity=0, replay=None,
                    track_id_base=0, drone_id_base=0):
        r = Run()
        r.seed, r.trial_begin, r.trial_count = int(seed), int(trial_begin), int(trial_count)
        r.precision, r.flags, r.record_capacity = int(precision), int(flags), int(record_capacity)
        r.track_id_base, r.drone_id_base = int(track_id_base), int(drone_id_base)
        r.replay = replay.ctypes.data if replay is not None else None
        return r

    def _result_buffers(self, n_pairs, run):
        cap = run.record_capacity or (1 << 16)
        counts = np.zeros(n_pairs, dtype=np.uint64)
        first = np.empty(n_pairs, dtype=np.int32)
        recs = np.empty(cap, dtype=RECORD_DTYPE)
        res = Result()
        res.counts = counts.ctypes.data_as(C.POINTER(C.c_uint64))
        res.first_conflict = first.ctypes.data_as(C.POINTER(C.c_int32))
        res.records = recs.ctypes.data
        per = None
        if run.flags & FLAG_PER_TRIAL:
            n = n_pairs * run.trial_count
            per = {"hit": np.zeros(n, np.uint8), "first_hit": np.zeros(n, np.int32), "min_dist": np.zeros(n, np.float64)}
            res.trial_hit = per["hit"].ctypes.data_as(C.POINTER(C.c_uint8))
            res.trial_first_hit = per["first_hit"].ctypes.data_as(C.POINTER(C.c_int32))
            res.trial_min_dist = dptr(per["min_dist"])
        return res, counts, first, recs, per

    @staticmethod
    def _wrap(res, counts, first, recs, per, n_pairs, trial_count):
        if per is not None:
            per = {k: v.reshape(n_pairs, trial_count) for k, v in per.items()}
        return RunResult(counts, first, recs[:res.n_records].copy(), bool(res.overflow), res.stats.as_dict(), per)

    # ---- the ABI
    def simulate(self, tracks, drones=None, seed=20170630, trial_begin=0, trial_count=10000, precision=MIXED,
                 flags=0, record_capacity=0, replay=None, model=None, track_id_base=0, drone_id_base=0):
        """dcrp_simulate: host buffers in, host buffers out (blocking).  `tracks` may be a
        Scenario (then `drones` is ignored)."""
        if isinstance(tracks, Scenario):
            T, D = tracks.T, tracks.D
        else:
            T, D, keep = self._pack(tracks, drones)
        m = model or default_model()
        run = self._run_struct(seed, trial_begin, trial_count, precision, flags, record_capacity, replay,
                               track_id_base, drone_id_base)
        n_pairs = len(T) * len(D)
        res, counts, first, recs, per = self._result_buffers(n_pairs, run)
        _check(lib().dcrp_simulate(self._h, T, len(T), D, len(D), C.byref(m), C.byref(run), C.byref(res)))
        self.n_tracks, self.n_drones = len(T), len(D)
        return self._wrap(res, counts, first, recs, per, n_pairs, run.trial_count)

    def prepared_simulate(self, scenario, precision=MIXED, flags=0, record_capacity=4096, seed=20170630, model=None):
        """dcrp_simulate with everything marshalled ONCE: returns call(trial_begin, trial_count) -> (counts, n_records,
        stats struct).  The per-call work is the C call itself (what a C/C++ caller pays); buffers are reused."""
        T, D = scenario.T, scenario.D
        m = model or default_model()
        run = self._run_struct(seed, 0, 0, precision, flags, record_capacity)
        n_pairs = len(T) * len(D)
        res, counts, first, recs, per = self._result_buffers(n_pairs, run)
        fn, h, nT, nD = lib().dcrp_simulate, self._h, len(T), len(D)
        pm, prun, pres = C.byref(m), C.byref(run), C.byref(res)
        self.n_tracks, self.n_drones = nT, nD

        def call(trial_begin, trial_count):
            run.trial_begin, run.trial_count = trial_begin, trial_count
            counts.fill(0)
            _check(fn(h, T, nT, D, nD, pm, prun, pres))
            return counts, res.n_records, res.stats

        call.keep = (T, D, m, run, res, counts, first, recs, scenario)
        return call

    def upload(self, tracks, drones=None, model=None):
        if isinstance(tracks, Scenario):
            T, D = tracks.T, tracks.D
        else:
            T, D, keep = self._pack(tracks, drones)
        m = model or default_model()
        _check(lib().dcrp_scenario_upload(self._h, T, len(T), D, len(D), C.byref(m)))
        self.n_tracks, self.n_drones = len(T), len(D)

    def run_async(self, seed=20170630, trial_begin=0, trial_count=10000, precision=MIXED, flags=0,
                  record_capacity=0, replay=None, stream=None, d_counts=None, track_id_base=0,
                  drone_id_base=0):
        run = self._run_struct(seed, trial_begin, trial_count, precision, flags, record_capacity, replay,
                               track_id_base, drone_id_base)
        self._keep = (run, replay)
        _check(lib().dcrp_run_async(self._h, C.byref(run), C.c_void_p(stream or 0), C.c_void_p(d_counts or 0)))
        return run

    def fetch(self, run):
        n_pairs = self.n_tracks * self.n_drones
        res, counts, first, recs, per = self._result_buffers(n_pairs, run)
        _check(lib().dcrp_fetch(self._h, C.byref(res)))
        return self._wrap(res, counts, first, recs, per, n_pairs, run.trial_count)

    def allpairs(self, seed, ray_begin, ray_count, ref_takeoff_t, ref_box, record_capacity=0, drone_id_base=0):
        """dcrp_allpairs on the uploaded scenario: every ray against every track."""
        ar = AllPairsRun()
        ar.seed, ar.ray_begin, ar.ray_count = int(seed), int(ray_begin), int(ray_count)
        ar.ref_takeoff_t, ar.record_capacity, ar.drone_id_base = int(ref_takeoff_t), int(record_capacity), int(drone_id_base)
        for k in range(6):
            ar.ref_box[k] = float(ref_box[k])
        run = self._run_struct(seed, ray_begin, ray_count, MIXED, 0, record_capacity)
        n_pairs = self.n_tracks * self.n_drones
        res, counts, first, recs, per = self._result_buffers(n_pairs, run)
        _check(lib().dcrp_allpairs(self._h, C.byref(ar), C.byref(res)))
        return self._wrap(res, counts, first, recs, per, n_pairs, run.trial_count)

    def last_run_ms(self):
        ms = C.c_float()
        _check(lib().dcrp_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    def dump_draws(self, track, drone, seed, trial_begin, n, track_id_base=0, drone_id_base=0):
        out = np.zeros(n, dtype=DRAW_DTYPE)
        run = self._run_struct(seed, trial_begin, n, FP64, track_id_base=track_id_base,
                               drone_id_base=drone_id_base)
        _check(lib().dcrp_dump_draws(self._h, int(track), int(drone), C.byref(run), out.ctypes.data))
        return out

    def trajectories(self, records):
        recs = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
        stride = C.c_int32()
        _check(lib().dcrp_trajectories(self._h, None, 0, None, C.byref(stride)))
        xyz = np.zeros((len(recs), 3, stride.value), dtype=np.float64)
        if len(recs):
            _check(lib().dcrp_trajectories(self._h, recs.ctypes.data, len(recs), dptr(xyz), C.byref(stride)))
        return xyz

    # ---- multi-process exchange
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        _check(lib().dcrp_comm_unique_id(buf))
        return buf.raw

    def attach_comm(self, unique_id, rank, nranks, flags=0):
        assert len(unique_id) == 128
        _check(lib().dcrp_engine_attach_comm(self._h, C.c_char_p(unique_id), int(rank), int(nranks), int(flags)))

    def detach_comm(self):
        _check(lib().dcrp_engine_detach_comm(self._h))

    def allreduce_counts(self, stream=None, d_counts=None):
        _check(lib().dcrp_allreduce_counts(self._h, C.c_void_p(stream or 0), C.c_void_p(d_counts or 0)))

    def comm_info(self):
        a = (C.c_int32 * 8)()
        _check(lib().dcrp_comm_info(self._h, a))
        return {"attached": bool(a[0]), "rank": a[1], "nranks": a[2], "peer_memory": bool(a[3]), "nccl_version": a[4],
                "allreduces_peer_memory": a[5], "allreduces_nccl": a[6], "peer_timeout": bool(a[7])}

    def reserve_sms(self, n):
        _check(lib().dcrp_engine_reserve_sms(self._h, int(n)))

    def launch_count(self):
        n = C.c_uint64()
        _check(lib().dcrp_launch_count(self._h, C.byref(n)))
        return n.value

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, maj, mn, khz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _check(lib().dcrp_device_info(self._h, name, 256, C.byref(sm), C.byref(maj), C.byref(mn), C.byref(khz)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": "%d.%d" % (maj.value, mn.value),
                "clock_khz": khz.value}
