"""GPU tier, needs >= 2 GPUs (skipped on the one-GPU test box; run with `gpurun --gpus 2`, log under profiles/):
one PROCESS per GPU, the engine's own communicator -- no torch anywhere.  The per-pair counts of two ranks that
ran disjoint Philox trial ranges, all-reduced over peer memory and by ncclAllReduce, equal one GPU's counts of the
whole range; the C++ driver's multi-process mode prints the same count and writes byte-identical files."""
import multiprocessing as mp
import os
import subprocess

import numpy as np
import pytest

import drone_collision_rp_b200 as pk
from drone_collision_rp_b200 import capi, scenario

pytestmark = pytest.mark.gpu


def _n_gpus():
    import ctypes as C

    n = C.c_int()
    return n.value if capi.lib().dcrp_device_count(C.byref(n)) == 0 else 0


def _rank_main(rank, world, uid, flags, day_csv, n, out_q):
    try:
        eng = capi.Engine(rank)
        tracks = [scenario.load_flight(day_csv, a) for a in range(2)]
        ring = scenario.load_ring(pk.RING_1KM)
        eng.attach_comm(uid, rank, world, flags)
        eng.upload(tracks, ring)
        out = {}
        for rep in range(3):                                   # several exchanges: both mailbox parities, sequence numbers
            begin, count = scenario.partition_trials(rep * n, n, rank, world)
            run = eng.run_async(seed=5, trial_begin=begin, trial_count=count, precision=capi.MIXED)
            eng.allreduce_counts()
            res = eng.fetch(run)
            out[rep] = res.counts.copy()
        out["info"] = eng.comm_info()
        eng.detach_comm()
        eng.close()
        out_q.put((rank, out))
    except Exception as exc:                                   # surface the failure instead of a hung join
        out_q.put((rank, {"error": repr(exc)}))


@pytest.mark.parametrize("route", ["peer", "nccl"])
def test_two_processes_allreduce_equals_one_gpu(route, tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _, day_csv = scenario.synthetic_tracks(str(tmp_path), 2)
    n = 400000
    ctx = mp.get_context("spawn")
    uid = capi.Engine.comm_unique_id()
    q = ctx.Queue()
    flags = capi.COMM_NCCL_ONLY if route == "nccl" else 0
    procs = [ctx.Process(target=_rank_main, args=(r, 2, uid, flags, day_csv, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert "error" not in got[0] and "error" not in got[1], got
    eng = capi.Engine(0)
    tracks = [scenario.load_flight(day_csv, a) for a in range(2)]
    ring = scenario.load_ring(pk.RING_1KM)
    for rep in range(3):
        whole = eng.simulate(tracks, ring, seed=5, trial_begin=rep * n, trial_count=n, precision=capi.FP64)
        assert whole.hits > 0
        assert np.array_equal(got[0][rep], whole.counts) and np.array_equal(got[1][rep], whole.counts), rep
    info = got[0]["info"]
    assert info["nranks"] == 2 and info["peer_memory"] == (route == "peer")
    assert (info["allreduces_peer_memory"], info["allreduces_nccl"]) == ((3, 0) if route == "peer" else (0, 3))
    eng.close()


def test_cpp_driver_one_process_per_gpu(tmp_path):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    ring = pk.RING_1KM

    def run(extra, cwd):
        os.makedirs(cwd, exist_ok=True)
        return subprocess.Popen([pk.MONTE_CARLO, "--synthetic", "3", "--aircraft", "3", "--drone-csv", ring, "--runs", "300000",
                                 "--batch"] + extra, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)

    one = run([], str(tmp_path / "one"))
    out_one = one.communicate(timeout=300)[0]
    assert one.returncode == 0, out_one
    rv = str(tmp_path / "rendezvous")
    ps = [run(["--ranks", "2", "--rank", str(r), "--rendezvous", rv], str(tmp_path / ("r%d" % r))) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in ps]
    assert all(p.returncode == 0 for p in ps), outs
    count = [l for l in out_one.splitlines() if "Number of collisions" in l][0].split(":")[1].strip()
    assert float(count) > 0
    for r in range(2):                                         # every rank reports the GLOBAL count
        line = [l for l in outs[r].splitlines() if "Number of collisions" in l][0]
        assert line.startswith("rank %d of 2" % r) and line.split(":")[-1].strip() == count
    for a in range(3):                                         # rank 0 wrote the merged files: same bytes as one GPU
        f = "Drone_Collisions/Drone_coords_%d.csv" % a
        assert open(tmp_path / "r0" / f).read() == open(tmp_path / "one" / f).read()
