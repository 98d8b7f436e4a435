"""CPU tier: the C-ABI libraries load and export every declared symbol, the C++
host mirror (Aircraft / Drone CSV handling, synthetic day, writer) behaves like the
reference's, and nothing pretends to compute without a GPU."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

ROOT = pk.REPO_ROOT
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_libraries_export_every_declared_symbol():
    for header, libpath, listed in (("dcrp.h", pk.LIBDCRP, capi.DCRP_SYMBOLS),
                                    ("dcrp_host.h", pk.LIBDCRP_HOST, capi.DCRP_HOST_SYMBOLS)):
        text = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"\b(dcrph?_[a-z0-9_]+)\s*\(", text))
        assert declared == set(listed), (declared ^ set(listed))
        nm = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
        exported = set(re.findall(r" T (dcrph?_[a-z0-9_]+)", nm))
        assert declared <= exported, declared - exported
    assert capi.lib().dcrp_abi_version() == 5


def test_abi_struct_sizes_match_header():
    src = r'''
    #include <stdio.h>
    #include "dcrp.h"
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(dcrp_track), sizeof(dcrp_drone), sizeof(dcrp_draw),
      sizeof(dcrp_model), sizeof(dcrp_run), sizeof(dcrp_record), sizeof(dcrp_stats), sizeof(dcrp_result)); return 0; }'''
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(tmp, "s.c"), "-o",
                        os.path.join(tmp, "s")], check=True)
        sizes = [int(v) for v in subprocess.run([os.path.join(tmp, "s")], capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(capi.Track), C.sizeof(capi.Drone), C.sizeof(capi.Draw), C.sizeof(capi.Model), C.sizeof(capi.Run),
            capi.RECORD_DTYPE.itemsize, C.sizeof(capi.Stats), C.sizeof(capi.Result)]
    assert sizes == want


@pytest.mark.skipif(_has_gpu(), reason="CPU-only behaviour")
def test_no_gpu_means_loud_failure_not_fallback(tmp_path):
    with pytest.raises(capi.DcrpError) as e:
        capi.Engine(0)
    assert e.value.status == capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    # the C++ driver too: error + non-zero exit, no collision count printed
    for f in ("drone_inital_positions_1km.csv",):
        os.symlink(os.path.join(pk.DATA_DIR, f), tmp_path / f)
    r = subprocess.run([pk.MONTE_CARLO, "--synthetic", "3", "--runs", "10", "--drones", "2"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode != 0
    assert "ERROR - no CUDA device" in r.stdout and "Number of collisions" not in r.stdout


def test_ring_rows_as_the_reference_binds_them():
    want = json.load(open(os.path.join(GOLDEN, "ref_ring_rows.json")))
    for name, path in (("drone_inital_positions_1km.csv", pk.RING_1KM), ("drone_inital_positions_5km.csv", pk.RING_5KM)):
        got = scenario.load_ring(path, n=100)
        assert np.array_equal(got, np.array(want[name]))
        assert np.array_equal(got[99], got[0]) or np.allclose(got[99, :2], got[0, :2], atol=1e-6)   # row 99 == row 0
    with pytest.raises(capi.DcrpError):
        scenario.load_ring(pk.RING_1KM, n=101)          # row 101 does not exist
    with pytest.raises(capi.DcrpError):
        scenario.load_ring(os.path.join(pk.DATA_DIR, "nope.csv"), n=1)


def test_ring_generator_reproduces_the_shipped_rings(tmp_path):
    """rings of other radii.  Calibration check: 1 km and 5 km rings against the shipped files."""
    for radius, shipped in ((1000.0, pk.RING_1KM), (5000.0, pk.RING_5KM)):
        path = str(tmp_path / ("ring_%d.csv" % radius))
        capi._hcheck(capi.host().dcrph_ring_write(path.encode(), radius, 100))
        got = scenario.load_ring(path, n=100)
        want = scenario.load_ring(shipped, n=100)
        assert np.abs(got[:, 2] - want[:, 2]).max() < 1e-12                   # headings: pi + 2 pi k / 99
        assert np.hypot(got[:, 0] - want[:, 0], got[:, 1] - want[:, 1]).max() < 1e-3 * radius
        assert open(path, "rb").read().count(b"\r\n") == 101
    path = str(tmp_path / "ring_2500.csv")
    capi._hcheck(capi.host().dcrph_ring_write(path.encode(), 2500.0, 100))
    r = scenario.load_ring(path, n=99)
    d = np.hypot(r[:, 0] - 676346.3657, r[:, 1] - 5705295.6395)
    assert (d > 2500.0).all() and (d < 2510.0).all()


def test_synthetic_day_schema_and_determinism(tmp_path):
    a = scenario.write_synthetic_day(str(tmp_path / "a.csv"), 5, 77, True)
    b = scenario.write_synthetic_day(str(tmp_path / "b.csv"), 5, 77, True)
    c = scenario.write_synthetic_day(str(tmp_path / "c.csv"), 5, 78, True)
    assert open(a).read() == open(b).read() != open(c).read()
    lines = open(a).read().strip().split("\n")
    assert all(len(l.split(",")) == 22 for l in lines)                 # Monte_Carlo.cpp:12
    assert scenario.flight_count(a) == 6                               # 5 + trailing dummy
    for i in range(5):
        f = scenario.load_flight(a, i)
        assert 92 <= f["n_steps"] <= 105 and 23 <= f["takeoff_t"] <= 33
        assert f["z"][0] == 0.0 and (f["z"][: f["takeoff_t"] + 1] == 0).all() and f["z"][f["takeoff_t"] + 1] > 0
        box = capi.target_box(f["y"][0], f["z"][0])                    # inside a runway band
        assert box[1] == 676819.02                                     # departure


def test_aircraft_loader_matches_golden_and_reference(day, golden):
    import oracle_api as oa

    tracks, csv = day
    assert scenario.flight_count(csv) == int(golden["ref_index_size"]) == 7
    for a, t in enumerate(tracks):
        for k in "xyz":
            assert np.array_equal(t[k], golden["track%d_%s" % (a, k)])
        assert t["takeoff_t"] == int(golden["track%d_takeoff_t" % a])
    if oa.RefAircraft.available():
        ra = oa.RefAircraft()
        for a in range(6):
            r = ra.load(csv, a)
            t = scenario.load_flight(csv, a)
            assert r["n_steps"] == t["n_steps"] and r["takeoff_t"] == t["takeoff_t"]
            for k in ("x", "y", "z", "gs"):
                assert np.array_equal(r[k], t[k])
        # the reference reaches the last flight only through index == N (Aircraft.cpp:73-82);
        # the new loader serves it at N-1 as well
        last_ref = ra.load(csv, 7)
        assert np.array_equal(last_ref["x"], scenario.load_flight(csv, 7)["x"])
        assert np.array_equal(last_ref["x"], scenario.load_flight(csv, 6)["x"])


def test_aircraft_loader_errors(tmp_path):
    with pytest.raises(capi.DcrpError):
        scenario.flight_count(str(tmp_path / "missing.csv"))
    p = scenario.write_synthetic_day(str(tmp_path / "d.csv"), 2, 1, False)
    assert scenario.flight_count(p) == 2
    with pytest.raises(capi.DcrpError):
        scenario.load_flight(p, 5)
    # takeoff_t never reads past the end when the on-ground flag never flips
    rows = open(p).read().split("\n")
    flat = [rows[0]] + [r.replace(",False,", ",True,", 1) if False else r for r in rows[1:]]
    open(tmp_path / "e.csv", "w").write("\n".join(flat))
    assert scenario.load_flight(str(tmp_path / "e.csv"), 1)["n_steps"] > 0


def test_partition_trials_is_disjoint_and_exhaustive():
    for begin, count, world in ((0, 10, 3), (5, 101011, 8), (2**40, 7, 8), (0, 0, 4), (3, 1, 2)):
        parts = [scenario.partition_trials(begin, count, r, world) for r in range(world)]
        assert sum(c for _, c in parts) == count
        pos = begin
        for b, c in parts:
            assert b == pos
            pos += c
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_slab_screen_block_schedule_deals_every_item_exactly_once():
    """k_screen3's work counter numbers BLOCKS of consecutive items whose size tapers 32, 16, ... 1 with the block
    number (csrc/engine.cu s3_schedule, read by screen3.cu grab_take).  For launch sizes from empty to 2^30 items:
    the blocks tile [n_static, n_items) in order with no gap and no overlap, sizes never grow, full-size blocks are
    powers of two, and the last 4 * n_warps items (at least) are dealt out one by one."""
    L = capi.lib()
    L.dcrp_debug_block_schedule.restype = C.c_int64
    L.dcrp_debug_block_schedule.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int64]
    rng = np.random.default_rng(5)
    cases = [(0, 8, 0), (1, 8, 8), (7, 8, 8), (9, 8, 8), (39105, 3552, 7104), (390654, 4736, 9472), (3920400, 4736, 9472),
             (2**20 + 3, 3552, 7104), (5000, 8, 16), (100000, 8, 16)]
    for _ in range(60):
        w = int(rng.integers(1, 600)) * 8
        fb = int(rng.integers(1, 3))
        cases.append((int(rng.integers(0, 3_000_000)), w, w * fb))
    cap = 3_200_000
    b = np.zeros(cap, dtype=np.uint32)
    e = np.zeros(cap, dtype=np.uint32)
    for n_items, n_warps, n_static in cases:
        nb = L.dcrp_debug_block_schedule(n_items, n_warps, n_static, b.ctypes.data, e.ctypes.data, cap)
        assert 0 <= nb <= cap, (n_items, n_warps, n_static, nb)
        if n_items <= n_static:
            assert nb == 0
            continue
        bb, ee = b[:nb].astype(np.int64), e[:nb].astype(np.int64)
        assert bb[0] == n_static and ee[-1] == n_items
        assert np.array_equal(bb[1:], ee[:-1])                       # contiguous, in order: every item exactly once
        size = ee - bb
        assert size.min() >= 1 and size.max() <= 32
        assert np.all(np.diff(size[:-1]) <= 0) or nb < 3             # sizes taper (the very last block may be cut short)
        full = size[:-1] if nb > 1 else size
        assert np.all((full & (full - 1)) == 0)                      # powers of two
        tail = min(n_items - n_static, 4 * n_warps - 1)
        assert np.all(size[bb >= n_items - tail] == 1)               # the end of the launch goes item by item
    # the whole 2^30-item launch limit: block count and coverage by the phase arithmetic alone (no arrays)
    assert L.dcrp_debug_block_schedule(2**30, 4736, 9472, None, None, 0) > 0


def test_committed_workload_csv_is_what_the_generator_writes(tmp_path):
    """bench.py (both arms) reads data/synthetic_day_1_20170630.csv; it must stay byte for byte what
    host/synth_day.cpp writes for one flight with the default seed."""
    import drone_collision_rp_b200 as pk
    from drone_collision_rp_b200 import scenario

    fresh = scenario.write_synthetic_day(str(tmp_path / "day.csv"), 1)
    committed = os.path.join(pk.DATA_DIR, "synthetic_day_1_20170630.csv")
    assert open(fresh, "rb").read() == open(committed, "rb").read()
    import json
    fx = json.load(open(os.path.join(pk.REPO_ROOT, "tests", "golden", "ref_config2_5km.json")))
    t = scenario.load_flight(committed, 0)
    assert (t["n_steps"], t["takeoff_t"]) == (fx["track_n_steps"], fx["track_takeoff_t"])
