"""Input helpers for tests and bench.py: everything goes through the C++ host
loaders (libdcrp_host.so), i.e. the same code a C++ user of the drop-in runs."""
import ctypes as C
import os

import numpy as np

from . import RING_1KM, RING_5KM, capi

AIRCRAFT_COLS = 22     # Monte_Carlo.cpp:12
DRONE_COLS = 4         # Monte_Carlo.cpp:16
DRONE_TOTAL = 99       # Monte_Carlo.cpp:17
DEFAULT_SEED = 20170630
SYNTH_DAY_SEED = 20170630


def load_ring(path=RING_1KM, n=DRONE_TOTAL, total_cols=DRONE_COLS):
    """(n,3) array of x0, y0, heading0: rows 1..n of the ring CSV as Drone binds them."""
    out = np.zeros((n, 3), dtype=np.float64)
    row = (C.c_double * 3)()
    for j in range(n):
        capi._hcheck(capi.host().dcrph_ring_row(path.encode(), total_cols, j, row))
        out[j] = row[:]
    return out


def write_synthetic_day(path, n_flights, seed=SYNTH_DAY_SEED, trailing_dummy=True, arrival_every=0):
    """arrival_every=K > 0 makes every K-th flight an arrival (first altitude != 0)."""
    capi._hcheck(capi.host().dcrph_synth_day_write_mixed(path.encode(), int(n_flights), int(seed),
                                                         int(trailing_dummy), int(arrival_every)))
    return path


def flight_count(path, n_cols=AIRCRAFT_COLS):
    size = C.c_int()
    capi._hcheck(capi.host().dcrph_aircraft_load(path.encode(), n_cols, -1, None, None, C.byref(size),
                                                None, None, None, None, 0))
    return size.value


def load_flight(path, index, n_cols=AIRCRAFT_COLS, cap=65536):
    """dict(x, y, z, gs, takeoff_t) of flight `index` through the host Aircraft class."""
    vl, tk, size = C.c_int(), C.c_int(), C.c_int()
    x, y, z, gs = (np.zeros(cap) for _ in range(4))
    n = capi._hcheck(capi.host().dcrph_aircraft_load(path.encode(), n_cols, int(index), C.byref(vl), C.byref(tk),
                                                     C.byref(size), capi.dptr(x), capi.dptr(y), capi.dptr(z),
                                                     capi.dptr(gs), cap))
    return {"x": x[:n].copy(), "y": y[:n].copy(), "z": z[:n].copy(), "gs": gs[:n].copy(),
            "takeoff_t": tk.value, "n_steps": vl.value}


def synthetic_tracks(workdir, n_flights, seed=SYNTH_DAY_SEED, arrival_every=0):
    """Write a synthetic day under workdir and load its first n_flights tracks."""
    os.makedirs(workdir, exist_ok=True)
    path = os.path.join(workdir, "synthetic_day_%d_%d_%d.csv" % (n_flights, seed, arrival_every))
    if not os.path.exists(path):
        write_synthetic_day(path, n_flights, seed, True, arrival_every)
    return [load_flight(path, i) for i in range(n_flights)], path


def partition_trials(trial_begin, trial_count, rank, world):
    """Contiguous, disjoint, exhaustive split of [trial_begin, trial_begin+trial_count)
    over `world` ranks; returns (begin, count) of `rank`.  Philox
    counters carry the global trial id, so the union of the ranks' results equals
    the single-GPU result bit for bit."""
    base, extra = divmod(int(trial_count), int(world))
    count = base + (1 if rank < extra else 0)
    begin = int(trial_begin) + rank * base + min(rank, extra)
    return begin, count
