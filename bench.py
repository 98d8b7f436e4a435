#!/usr/bin/env python
"""bench.py -- separation tests/s of the Monte-Carlo trial path on N B200s.

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1],
the 5 km ring (99 drones) against one LHR-style departure track, 101 011 trials
per drone = 1.0e7 trials, DCRP_PRECISION_MIXED (fp32 packed screen + fp64
confirm: same answers as the fp64 kernel).  With N GPUs every rank runs the same
number of trials on a disjoint Philox trial range (weak scaling) and the per-pair
hit counts are combined once per step by the engine's own communicator
(dcrp_engine_attach_comm / dcrp_allreduce_counts: peer-memory kernels over NVLink,
NCCL for the bootstrap and as the fallback).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's engine
  python bench.py --impl reference [...]                       # the reference's CPU path

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "drone-aircraft separation tests/sec"
UNIT = "tests/s"
SEED = 20170630
TRIALS_PER_DRONE = 101011          # x 99 drones = 1.0000089e7 trials  (BASELINE config 2: "1e7 trials")
FLOP_PER_TEST = 11                 # 3 add + 3 sub + 3 mul + 2 add, FMA = 2
SUSTAIN_TRIALS_PER_DRONE = 50_000_000   # 4.95e9 trials per launch for the steady-state figure
SUSTAIN_LAUNCHES = 10
STRONG_TRIALS_PER_DRONE = 101_010_102   # x 99 drones = 1.0e10 trials: the fixed job of the strong-scaling figure
COMM_ROUTE = os.environ.get("DCRP_BENCH_COMM", "peer")    # "peer" (default) or "nccl": dcrp_engine_attach_comm flags
DATA = os.path.join(ROOT, "data")
DAY_CSV = os.path.join(DATA, "synthetic_day_1_20170630.csv")   # the workload's track as written by host/synth_day.cpp
RING_CSV = os.path.join(DATA, "drone_inital_positions_5km.csv")
FIXTURE = os.path.join(ROOT, "tests", "golden", "ref_config2_5km.json")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on
# stdout.  Keep the real stdout aside, point fd 1 at stderr for the whole run, emit the line last.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def workload():
    """The engine arm's inputs, through the C++ host loaders (the committed CSV is byte for byte what
    host/synth_day.cpp writes for 1 flight, seed 20170630: tests/test_host_logic.py)."""
    from drone_collision_rp_b200 import scenario

    tracks = [scenario.load_flight(DAY_CSV, 0)]
    ring = scenario.load_ring(RING_CSV)
    return tracks, ring, DAY_CSV


def track_shape():
    """n_steps / takeoff_t of the workload's track without loading any native library (reference arm)."""
    with open(FIXTURE) as f:
        fx = json.load(f)
    return {"n_steps": fx["track_n_steps"], "takeoff_t": fx["track_takeoff_t"]}


def config_dict(n_gpus, track):
    return {
        "workload": "BASELINE configs[1]: drone_inital_positions_5km.csv (99 drones) x 1 synthetic LHR A320 "
                    "departure track (%d samples, takeoff_t %d; the real ADS-B CSV is not distributed) x "
                    "%d trials/drone = %d trials per step per GPU" % (
                        track["n_steps"], track["takeoff_t"], TRIALS_PER_DRONE, 99 * TRIALS_PER_DRONE),
        "precision": "mixed (fp32 packed screen of every trial + fp64 confirm of candidates; results == fp64 mode)",
        "seed": SEED,
        "sharding": "rank g runs trials [(s*N+g)*n, +n) of every (track, drone); counts all-reduced once per step by "
                    "the engine's communicator (dcrp_allreduce_counts) on a high-priority side stream, overlapped "
                    "with the next step's kernels (one exposed all-reduce per job is added to the timed total)",
        "l2": "inputs (3.3 KB track + 2.4 KB ring) are shared-memory/L2 resident by design; a 256 MB buffer "
              "is written between timed steps to flush L2 anyway",
        "parallelism": "dp%d over trial ranges" % n_gpus,
    }


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                sm.append(float(c[1])); smmax.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in the load window"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smmax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the reference's CPU path (cpu_baseline leg and --impl reference): the ONLY code here that may
# execute anything under oracle/.  Nothing in this section imports the package or loads libdcrp*.so.
# ----------------------------------------------------------------------------------------------
def oracle_paths():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa          # pure ctypes glue around oracle/_build and oracle/_ref

    return oa


def read_ring_rows(path, n=99):
    """Rows 1..n of a drone ring CSV as (x0, y0, heading0) -- plain text parsing, no native code."""
    rows = []
    with open(path, newline="") as f:
        lines = f.read().replace("\r", "").split("\n")
    for line in lines[1:1 + n]:
        c = line.split(",")
        rows.append((float(c[1]), float(c[2]), float(c[3])))
    return rows


def cpu_reference_run(csv, ring_csv, runs_per_drone, nproc, scratch):
    """oracle/_ref/ref_asshipped: unmodified Aircraft.cpp + Drone.cpp, as shipped, nproc processes."""
    oa = oracle_paths()
    out = subprocess.run([oa.REF_ASSHIPPED, csv, "22", "0", ring_csv, "99", str(runs_per_drone), str(nproc), scratch],
                         check=True, capture_output=True, text=True).stdout
    j = json.loads(out.strip().split("\n")[-1])
    tests = j["trials"] * j["vector_length"]          # Drone::Collision walks every index (Drone.cpp:254)
    return {"trials": j["trials"], "tests": tests, "seconds": j["seconds"], "hits": j["hits"], "kind": "reference"}


_REPLAY = {}


def _replay_worker(j):
    """One drone's table through the UNMODIFIED reference Drone.cpp under the replay shim (no random_device)."""
    oa = oracle_paths()
    t0 = time.perf_counter()
    r = oa.RefReplay().run(RING_CSV, j, 0, _REPLAY["track"], _REPLAY["draws"][j])
    return r["hits"], time.perf_counter() - t0


def _port_worker(j):
    oa = oracle_paths()
    t0 = time.perf_counter()
    hit, _, _ = oa.Oracle().run_trials(_REPLAY["track"], _REPLAY["ring"][j], _REPLAY["draws"][j])
    return int(hit.sum()), time.perf_counter() - t0


def cpu_replay_run(track, ring_rows, runs_per_drone, nproc, use_ref):
    """The reference's arithmetic on a replayable RNG: Philox tables made by the C
    oracle (untimed), replayed by nproc forked processes through oracle/_ref/libref_replay.so (or, when _ref did
    not travel, the C restatement).  Returns trials, tests, wall seconds of the replay."""
    import multiprocessing as mp

    oa = oracle_paths()
    o = oa.Oracle()
    rc, box = o.target_box(track["y"][0], track["z"][0])
    draws = [o.draws(SEED, 0, runs_per_drone, j, 0, track["takeoff_t"], box) for j in range(99)]
    _REPLAY.update(track=track, ring=ring_rows, draws=draws)
    fn = _replay_worker if use_ref else _port_worker
    t0 = time.perf_counter()
    if nproc == 1:
        res = [fn(j) for j in range(99)]
    else:
        with mp.get_context("fork").Pool(nproc) as pool:
            res = pool.map(fn, range(99), chunksize=1)
    secs = time.perf_counter() - t0
    trials = 99 * runs_per_drone
    return {"trials": trials, "tests": trials * track["n_steps"], "seconds": secs, "hits": sum(r[0] for r in res),
            "kind": "reference" if use_ref else "port", "busy_seconds": sum(r[1] for r in res)}


def load_track_plain():
    """The workload's track parsed from the committed CSV in pure Python (reference arm: no native loader)."""
    xs, ys, zs, og = [], [], [], []
    with open(DAY_CSV) as f:
        rows = [l.rstrip("\n").split(",") for l in f][1:]
    cs = rows[0][3]
    for r in rows:
        if r[3] != cs:
            break
        xs.append(float(r[14])); ys.append(float(r[13])); zs.append(float(r[7])); og.append(r[15])
    takeoff = 0
    while takeoff + 1 < len(og) and og[takeoff] == og[takeoff + 1]:      # Aircraft.cpp:102-110
        takeoff += 1
    import numpy as np

    return {"x": np.array(xs), "y": np.array(ys), "z": np.array(zs), "takeoff_t": takeoff, "n_steps": len(xs)}


def cpu_baseline(runs_per_drone, tmp, replay_runs=20000, one_core=True):
    """All the CPU figures of SURVEY 8(d) / BASELINE.md 3 on this box's host cores, bounded samples."""
    oa = oracle_paths()
    nproc = os.cpu_count() or 1
    have_ref = os.path.exists(oa.REF_ASSHIPPED) and os.path.exists(oa.LIBREF_REPLAY)
    track = load_track_plain()
    ring_rows = read_ring_rows(RING_CSV)
    if have_ref:
        r = cpu_reference_run(DAY_CSV, RING_CSV, runs_per_drone, nproc, os.path.join(tmp, "asshipped"))
        what = ("unmodified reference Aircraft.cpp+Drone.cpp as shipped (std::random_device per draw), "
                "%d processes" % nproc)
    else:
        r = cpu_replay_run(track, ring_rows, runs_per_drone, nproc, use_ref=False)
        what = "oracle/dcrp_oracle.c (C restatement, Philox draws), %d processes" % nproc
    cpu = {"value": r["tests"] / r["seconds"], "unit": UNIT, "cores": nproc, "kind": r["kind"],
           "sample": "99 drones x %d trials (%.1f s); %s" % (runs_per_drone, r["seconds"], what),
           "trials_per_s": r["trials"] / r["seconds"], "seconds": r["seconds"], "tests": r["tests"], "trials": r["trials"]}
    try:
        rp = cpu_replay_run(track, ring_rows, replay_runs, nproc, use_ref=have_ref)
        cpu["all_cores_replay"] = {
            "tests_per_s": rp["tests"] / rp["seconds"], "trials_per_s": rp["trials"] / rp["seconds"], "cores": nproc,
            "kind": rp["kind"],
            "sample": "%s on replayed Philox draws (no std::random_device), 99 drones x %d trials, %d processes "
                      "(%.1f s wall, %.1f s busy)" % ("unmodified reference Drone.cpp (oracle/_ref/libref_replay.so)"
                                                      if have_ref else "oracle/dcrp_oracle.c", replay_runs, nproc,
                                                      rp["seconds"], rp["busy_seconds"])}
        if one_core:
            if have_ref:
                one = cpu_reference_run(DAY_CSV, RING_CSV, 1500, 1, os.path.join(tmp, "asshipped1"))
                cpu["one_core_as_shipped"] = {"tests_per_s": one["tests"] / one["seconds"],
                                              "trials_per_s": one["trials"] / one["seconds"],
                                              "sample": "99 drones x 1500 trials, 1 process (%.1f s)" % one["seconds"]}
            p1 = cpu_replay_run(track, ring_rows, 4000, 1, use_ref=have_ref)
            cpu["one_core_replay"] = {"tests_per_s": p1["tests"] / p1["seconds"], "trials_per_s": p1["trials"] / p1["seconds"],
                                      "kind": p1["kind"], "sample": "99 drones x 4000 trials, 1 process (%.1f s)" % p1["seconds"]}
    except Exception as exc:          # reported baselines only: never fail the bench line over them
        cpu["replay_note"] = "unavailable: %s" % exc
    return cpu


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        runs = 4000                                    # 99 x 4000 = 396 000 trials per step (bounded sample)
        for _ in range(max(0, args.warmup)):
            cpu_baseline(max(200, runs // 10), tmp, replay_runs=500, one_core=False)
        secs, tests, trials, last = 0.0, 0, 0, None
        for k in range(args.steps):
            last = cpu_baseline(runs, tmp, replay_runs=20000 if k == args.steps - 1 else 500, one_core=False)
            secs += last["seconds"]; tests += last["tests"]; trials += last["trials"]
        value = tests / secs
        cb = {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
              "sample": "%d steps x 99 drones x %d trials (a bounded sample of the %d-trial workload); %s" % (
                  args.steps, runs, 99 * TRIALS_PER_DRONE, last["sample"].split("; ", 1)[1])}
        if "all_cores_replay" in last:
            cb["all_cores_replay"] = last["all_cores_replay"]
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.gpus, track_shape()),
            "trials_per_s": trials / secs,
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# this repo's engine
# ----------------------------------------------------------------------------------------------
def engine_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from drone_collision_rp_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def allmax(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def allsum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    tmp_ctx = tempfile.TemporaryDirectory()
    tmp = tmp_ctx.name
    tracks, ring, csv = workload()
    eng = capi.Engine(local)
    info = eng.device_info()
    n_pairs = len(tracks) * len(ring)
    n = TRIALS_PER_DRONE
    # a real (non-default) torch stream: the engine treats a NULL handle as "use my own
    # stream", and torch's events must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    sh = stream.cuda_stream
    assert sh != 0

    # ---- roofline denominators measured on this device
    peaks = {}
    if rank == 0:
        for which, name in ((0, "ffma"), (1, "ffma2"), (2, "dfma"), (3, "scan_mix_11flop")):
            tf, ms = capi.measure_peak(local, which)
            peaks[name] = tf
        log("peaks TFLOP/s:", peaks)

    # ---- the engine's own communicator: torch.distributed only carries the 128-byte id to the ranks
    comm_info = None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Engine.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        eng.attach_comm(bytes(uid.cpu().numpy().tobytes()), rank, world, capi.COMM_NCCL_ONLY if COMM_ROUTE == "nccl" else 0)
        comm_info = eng.comm_info()
    eng.upload(tracks, ring)

    # ---- guard: the timed configuration's first 1.0e7 trials give the counts and colliding ids the UNMODIFIED
    # reference gives on the same draws (tests/golden/ref_config2_5km.json; the full per-trial comparison is
    # tests/test_gpu_config2_full.py).  Every rank checks its own GPU.
    with open(FIXTURE) as f:
        fx = json.load(f)
    g = eng.fetch(eng.run_async(seed=SEED, trial_begin=0, trial_count=n, precision=capi.MIXED, stream=sh))
    want = sorted((c["drone"], c["trial"], c["first_hit"]) for c in fx["collisions"])
    got = sorted((int(r["drone"]), int(r["trial"]), int(r["first_hit"])) for r in g.records)
    if [int(c) for c in g.counts] != [p["hits"] for p in fx["per_drone"]] or got != want:
        raise SystemExit("bench.py: the timed configuration does not reproduce the reference fixture: %s vs %s" % (got, want))
    guard = {"fixture": "tests/golden/ref_config2_5km.json", "hits": int(g.hits), "ids_match": True}

    counts = [torch.zeros(n_pairs, dtype=torch.int64, device=dev) for _ in range(2)]

    # Multi-GPU: the counts of step s are all-reduced by the engine (dcrp_allreduce_counts) on a high-priority side
    # stream while step s+1 computes; a step's timed region ends when its own kernels are done AND the previous
    # step's reduction has landed.  The last reduction's tail is timed once after the loop and added to the total.
    comm = torch.cuda.Stream(dev, priority=-1) if world > 1 else None
    kdone = [torch.cuda.Event() for _ in range(2)]
    reduced = [torch.cuda.Event() for _ in range(2)]
    reduced_valid = [False, False]

    def step(s):
        b = s & 1
        c = counts[b]
        if reduced_valid[b]:
            stream.wait_event(reduced[b])          # the reduction that last used this buffer is done
        begin = (s * world + rank) * n
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run = eng.run_async(seed=SEED, trial_begin=begin, trial_count=n, precision=capi.MIXED, stream=sh,
                            d_counts=c.data_ptr(), flags=capi.FLAG_ZERO_COUNTS)
        if world > 1:
            kdone[b].record(stream)
            if reduced_valid[b ^ 1]:
                stream.wait_event(reduced[b ^ 1])  # previous step's counts are globally reduced
            e1.record(stream)
            comm.wait_event(kdone[b])
            eng.allreduce_counts(comm.cuda_stream, c.data_ptr())
            reduced[b].record(comm)
            reduced_valid[b] = True
        else:
            e1.record(stream)
        return run, e0, e1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    for s in range(args.warmup):
        run, e0, e1 = step(s)
        eng.fetch(run)
        capi.flush_l2(local, sh)
    barrier()

    launches0 = eng.launch_count()
    t_load0 = time.time()
    wall0 = time.perf_counter()
    keys = ("tests_evaluated", "tests_shared", "tests_issued", "drone_steps", "trials", "candidates", "hits", "level2",
            "slab_survivors")
    step_ms, trial_ms, stats_sum = [], [], {k: 0 for k in keys}
    untimed_launches = 0        # engine launches inside the loop that are not part of the timed steps
    for s in range(args.warmup, args.warmup + args.steps):
        run, e0, e1 = step(s)
        res = eng.fetch(run)                      # syncs the stream; outside the event pair
        step_ms.append(e0.elapsed_time(e1))
        trial_ms.append(res.stats["ms_trials"])
        for k in stats_sum:
            stats_sum[k] += res.stats[k]
        capi.flush_l2(local, sh)                  # between timed steps, outside the event pairs (libdcrp_bench.so)
    if world == 1:      # the caller-owned device counters (cleared by DCRP_FLAG_ZERO_COUNTS) hold the last step's hits
        assert int(counts[(args.warmup + args.steps - 1) & 1].sum().item()) == res.stats["hits"]
    tail_ms = 0.0
    if world > 1:
        # the last step's reduction has no next step to hide behind: time one isolated all-reduce of
        # the same buffer and charge it to the job
        barrier()                      # the ranks' skew is already in the max-over-ranks of the step times:
        t0e = torch.cuda.Event(enable_timing=True)     # time the collective itself
        t1e = torch.cuda.Event(enable_timing=True)
        t0e.record(stream)
        eng.allreduce_counts(sh, counts[(args.warmup + args.steps - 1) & 1].data_ptr())
        t1e.record(stream)
        torch.cuda.synchronize()
        tail_ms = t0e.elapsed_time(t1e)
        untimed_launches += 2 if (comm_info and comm_info["peer_memory"]) else 0
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launch_count() - launches0 - untimed_launches

    total_ms = sum(step_ms) + tail_ms
    per_rank = None
    if world > 1:
        mine = torch.tensor([sum(step_ms) / len(step_ms), sum(trial_ms) / len(trial_ms), tail_ms],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        # per-rank step and kernel times: tells a slow GPU from an exposed collective when the scaling is off
        per_rank = {"step_ms": [round(t[0].item(), 4) for t in allr], "screen_kernel_ms": [round(t[1].item(), 4) for t in allr],
                    "last_allreduce_ms": [round(t[2].item(), 4) for t in allr]}
    total_ms = allmax([total_ms])[0]
    tests, trials, dsteps, shared = allsum([float(stats_sum[k]) for k in ("tests_evaluated", "trials", "drone_steps", "tests_shared")])
    secs = total_ms * 1e-3
    value = tests / secs

    # ---- cross-check of the per-step timing: K steps back to back between ONE event pair, no fetch and no L2 flush in
    # between (the all-reduces overlap exactly as in a production loop; nothing can hide in an untimed gap)
    barrier()
    b0 = torch.cuda.Event(enable_timing=True)
    b1 = torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for s in range(args.warmup + args.steps, args.warmup + 2 * args.steps):
        step(s)
    if world > 1:
        stream.wait_event(reduced[(args.warmup + 2 * args.steps - 1) & 1])
    b1.record(stream)
    torch.cuda.synchronize()
    b2b_ms = allmax([b0.elapsed_time(b1)])[0]
    back_to_back = {"steps": args.steps, "ms_per_step": b2b_ms / args.steps,
                    "tests_per_s": tests / (b2b_ms * 1e-3),
                    "note": "same steps, one CUDA-event pair around all of them, every all-reduce inside, no fetch / flush between"}
    reduced_valid[0] = reduced_valid[1] = False

    # ---- steady state: one long launch per rank (also gives the clock sampler something to see)
    sus = {k: 0 for k in ("tests_evaluated", "trials", "tests_issued", "hits", "candidates", "slab_survivors")}
    sus_ms = 0.0
    for k in range(SUSTAIN_LAUNCHES):
        run = eng.run_async(seed=SEED, trial_begin=10**12 + (k * world + rank) * SUSTAIN_TRIALS_PER_DRONE,
                            trial_count=SUSTAIN_TRIALS_PER_DRONE, precision=capi.MIXED, stream=sh)
        res_s = eng.fetch(run)
        sus_ms += eng.last_run_ms()        # CUDA events around the launch on its stream (dcrp_last_run_ms)
        for q in sus:
            sus[q] += res_s.stats[q]
    t_load1 = time.time()
    sustain = {"launches": SUSTAIN_LAUNCHES, "trials": sus["trials"], "ms_trials_kernel": sus_ms,
               "tests_per_s": sus["tests_evaluated"] / (sus_ms * 1e-3), "trials_per_s": sus["trials"] / (sus_ms * 1e-3),
               "issued_over_evaluated": sus["tests_issued"] / max(1, sus["tests_evaluated"]), "hits": sus["hits"],
               "candidates": sus["candidates"], "slab_survivors_fraction": sus["slab_survivors"] / max(1, sus["trials"])}

    # ---- strong scaling (BASELINE config 3): ONE fixed job of 1.0e10 trials split over the ranks as contiguous trial
    # ranges; device time from the first kernel to the landed all-reduce, max over ranks
    strong = None
    if not args.no_strong:
        tot = STRONG_TRIALS_PER_DRONE
        lo, hi = tot * rank // world, tot * (rank + 1) // world
        best = None
        for rep in range(2):
            barrier()
            s0 = torch.cuda.Event(enable_timing=True)
            s1 = torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            run = eng.run_async(seed=SEED, trial_begin=2 * 10**12 + lo, trial_count=hi - lo, precision=capi.MIXED, stream=sh)
            if world > 1:
                eng.allreduce_counts(sh, None)
            s1.record(stream)
            r = eng.fetch(run)
            ms = allmax([s0.elapsed_time(s1)])[0]
            best = ms if best is None else min(best, ms)
        st_tests = allsum([float(r.stats["tests_evaluated"])])[0]
        strong = {"job": "BASELINE configs[2] on the 5 km ring: 99 drones x %d trials = %.4e trials, one track, split into "
                         "%d contiguous trial ranges" % (tot, 99.0 * tot, world),
                  "ms": best, "tests_per_s": st_tests / (best * 1e-3), "trials_per_s": 99.0 * tot / (best * 1e-3),
                  "global_hits": int(r.counts.sum()),
                  "note": "device time from the first kernel to the all-reduced counts, max over ranks, best of 2; the "
                          "speed-up over N=1 is this figure's ratio between two runs of bench.py"}

    # ---- end to end through the public C-ABI call with HOST buffers (dcrp_simulate: scenario check / upload,
    # kernels, download per step), on every rank at once; aggregate = all ranks' tests / slowest rank's wall time
    scn = capi.Scenario(tracks, ring)            # ctypes marshalling of the caller's arrays, done once
    # (DCRP_FLAG_NO_TIMING: the caller reads no dcrp_stats.ms_*, as the reference-facing Drone::Simulation does not)
    call = eng.prepared_simulate(scn, precision=capi.MIXED, record_capacity=4096, seed=SEED, flags=capi.FLAG_NO_TIMING)
    e2e_secs, e2e_tests, h2d, d2h, cached_steps = 0.0, 0, 0, 0, 0
    first_call = None
    for s in range(args.warmup + args.steps):
        if s == args.warmup and world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        _, _, st = call((s * world + rank) * n, n)
        dt = time.perf_counter() - t0
        if s == 0:
            first_call = {"ms": dt * 1e3, "h2d_bytes": int(st.h2d_bytes), "scenario_cached": bool(st.status & capi.STAT_SCENARIO_CACHED)}
        if s >= args.warmup:
            e2e_secs += dt
            e2e_tests += st.tests_evaluated
            h2d, d2h = int(st.h2d_bytes), int(st.d2h_bytes)
            cached_steps += 1 if st.status & capi.STAT_SCENARIO_CACHED else 0
    e2e_secs = allmax([e2e_secs])[0]
    e2e_tests = allsum([float(e2e_tests)])[0]
    e2e_value = e2e_tests / e2e_secs
    # the same with the scenario cache defeated (the track nudged by 1e-9 m every step): every call validates, packs,
    # uploads and rebuilds the stage-1 tables -- what a caller with a NEW scenario per call pays
    import copy
    cold_secs, cold_h2d = 0.0, 0
    for s in range(min(args.steps, 5) + 1):
        t2 = copy.deepcopy(tracks)
        t2[0]["x"] = t2[0]["x"] + 1e-9 * (s + 1)
        scn2 = capi.Scenario(t2, ring)
        call2 = eng.prepared_simulate(scn2, precision=capi.MIXED, record_capacity=4096, seed=SEED)
        t0 = time.perf_counter()
        _, _, st2 = call2((s * world + rank) * n, n)
        dt = time.perf_counter() - t0
        if s >= 1:
            cold_secs += dt
            cold_h2d = int(st2.h2d_bytes)
    cold_ms = allmax([1e3 * cold_secs / min(args.steps, 5)])[0]

    # ---- the same workload with the OPT-IN exact reach culling (DCRP_FLAG_REACH_CULL): identical counts,
    # ids and records (tests/test_gpu_reach_cull.py); reported beside the headline, never as it
    eng.upload(tracks, ring)
    rc_ms, rc_tests, rc_issued, rc_culled_trials, rc_trials, rc_hits, rc_eval = 0.0, 0, 0, 0, 0, 0, 0
    for s in range(args.warmup + args.steps):
        run = eng.run_async(seed=SEED, trial_begin=(s * world + rank) * n, trial_count=n, precision=capi.MIXED,
                            flags=capi.FLAG_REACH_CULL, stream=sh)
        r = eng.fetch(run)
        if s >= args.warmup:
            rc_ms += r.stats["ms_device_total"]; rc_tests += r.stats["tests_evaluated"] + r.stats["tests_culled"]
            rc_eval += r.stats["tests_evaluated"]
            rc_issued += r.stats["tests_issued"]; rc_culled_trials += r.stats["trials_culled"]
            rc_trials += r.stats["trials"]; rc_hits += r.stats["hits"]
        capi.flush_l2(local, sh)
    callc = eng.prepared_simulate(scn, precision=capi.MIXED, flags=capi.FLAG_REACH_CULL, record_capacity=4096, seed=SEED)
    rc_e2e_secs, rc_e2e_tests = 0.0, 0
    for s in range(args.warmup + args.steps):
        if s == args.warmup and world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        _, _, st = callc((s * world + rank) * n, n)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            rc_e2e_secs += dt
            rc_e2e_tests += st.tests_evaluated + st.tests_culled
    rc_ms, rc_e2e_secs = allmax([rc_ms, rc_e2e_secs])
    rc_tests, rc_issued, rc_culled_trials, rc_trials, rc_e2e_tests, rc_eval = allsum(
        [float(v) for v in (rc_tests, rc_issued, rc_culled_trials, rc_trials, rc_e2e_tests, rc_eval)])
    reach_cull = {
        "flag": "DCRP_FLAG_REACH_CULL (opt-in; results identical to the headline path)",
        "value": rc_tests / (rc_ms * 1e-3), "unit": "tests decided/s (computed or proven impossible by the reach bound; "
                                                    "pairs cleared as a whole count their guaranteed minimum)",
        "ms_per_step": rc_ms / args.steps, "trials_per_s": rc_trials / (rc_ms * 1e-3),
        "tests_evaluated_per_s": rc_eval / (rc_ms * 1e-3),
        "tests_evaluated_fraction": rc_eval / max(1.0, rc_tests),
        "trials_in_cleared_pairs_fraction": rc_culled_trials / max(1.0, rc_trials),
        "e2e": {"value": rc_e2e_tests / rc_e2e_secs, "ms_per_step": 1e3 * rc_e2e_secs / args.steps},
    }
    # ---- the same workload in the sub-step mode (DCRP_FLAG_SUBSTEP, the interpolated-separation extension of SURVEY
    # 8(f3): a test is then one (trial, segment) closest approach); slab-first screen and, for comparison, the 2-D screen
    sub = {}
    for name, fl in (("slab_first", capi.FLAG_SUBSTEP), ("screen_2d", capi.FLAG_SUBSTEP | capi.FLAG_SCREEN_2D)):
        ms, tests_s, hits_s, surv = 0.0, 0, 0, 0
        for s in range(args.warmup + args.steps):
            r = eng.fetch(eng.run_async(seed=SEED, trial_begin=(s * world + rank) * n, trial_count=n, precision=capi.MIXED,
                                        flags=fl, stream=sh))
            if s >= args.warmup:
                ms += r.stats["ms_device_total"]; tests_s += r.stats["tests_evaluated"]; hits_s += r.stats["hits"]
                surv += r.stats["slab_survivors"]
            capi.flush_l2(local, sh)
        ms = allmax([ms])[0]
        tests_s, hits_s, surv = allsum([float(tests_s), float(hits_s), float(surv)])
        sub[name] = {"ms_per_step": ms / args.steps, "tests_per_s": tests_s / (ms * 1e-3), "hits": int(hits_s),
                     "slab_survivors_fraction": surv / (world * args.steps * 99.0 * n)}
    assert sub["slab_first"]["hits"] == sub["screen_2d"]["hits"]
    substep = {"flag": "DCRP_FLAG_SUBSTEP (extension: drone and aircraft interpolated between samples; results equal the "
                       "fp64 kernel's, tests/test_gpu_substep.py)", **sub}
    if comm_info is not None:
        comm_info = eng.comm_info()

    if rank != 0:
        if world > 1:
            dist.barrier()
            eng.detach_comm()
            dist.destroy_process_group()
        return 0

    clocks = sampler.stop(t_load0, t_load1)

    # ---- roofline of the dominant kernel (k_screen3), live CUDA-event durations of the timed steps
    k_ms = sum(trial_ms) / len(trial_ms)
    tests_per_launch = stats_sum["tests_evaluated"] / args.steps
    achieved = tests_per_launch * FLOP_PER_TEST / (k_ms * 1e-3) * 1e-12
    peak = peaks["ffma"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
    # executed fp32 flops: the slab level issues n_steps lane-steps per trial at 2 flop (advance + difference; the
    # min is not a flop), the 2-D re-scan of the survivors 7 flop per lane-step
    slab_steps = stats_sum["trials"] * float(tracks[0]["n_steps"])
    rescan_steps = max(0.0, stats_sum["tests_issued"] - slab_steps)
    exec_per_test = (2.0 * slab_steps + 7.0 * rescan_steps) / max(1.0, stats_sum["tests_evaluated"])
    roofline = {
        "bound": "fp32", "kernel": "k_screen3<MINB=2> (256 threads, 128 registers, 2 CTAs per SM; slab level + queued 2-D re-scan + level 2 + fp64 decision)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic,
        "traffic_note": "DRAM bytes per launch from one ncu --set full capture (profiles/r02_traffic.json); the "
                        "kernel is dispatch/FMA-pipe bound, its inputs (~4 KB) live in shared memory",
        "frac_is": "ALGORITHMIC: 11 flop per decided stage-2 test (SURVEY 8(d)) over the measured FFMA peak; the kernel "
                   "executes far fewer (frac_executed) because level 1a tests every (trial, index) along ONE direction",
        "executed_fp32_flop_per_test": exec_per_test,
        "levels": "level 1a (every test): |n . D(i)| along one direction per (track, drone) pair, 2 executed flop (advance, "
                  "difference) + half an FMNMX3; level 1b (the %.2f %% of trials it cannot clear): horizontal fp32 "
                  "separation, 7 flop; level 2 (3-D fp32) on ~5e-5 of trials; level 3 (fp64, reference arithmetic) on "
                  "~3e-7, all inside the one kernel" % (100.0 * stats_sum["slab_survivors"] / max(1, stats_sum["trials"])),
        "peak_source": "FFMA-only micro-kernel measured on this device in this run (MEASURED_PEAKS.json has no "
                       "FP32 figure); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
        "algorithmic_flop_per_test": FLOP_PER_TEST, "tests_per_launch": tests_per_launch, "kernel_ms": k_ms,
        "kernel_ms_is": "CUDA events on the launch stream around k_run_init + k_screen3: the screen kernel is launched "
                        "programmatically dependent on the 2 us clearing launch and overlaps it, so no event can stand "
                        "between the two; k_screen3 alone is ~2 us less (ncu: profiles/r02_bench_launches.csv)",
        "peaks_measured_tflops": peaks,
        "achieved_executed": achieved * exec_per_test / 11.0, "frac_executed": achieved * exec_per_test / 11.0 / peak,
        "issued_over_evaluated": stats_sum["tests_issued"] / max(1, stats_sum["tests_evaluated"]),
        "sustained": {"achieved": sustain["tests_per_s"] * FLOP_PER_TEST * 1e-12,
                      "frac": sustain["tests_per_s"] * FLOP_PER_TEST * 1e-12 / peak, **sustain},
        "hbm": {"note": "track 3.3 KB + ring tables per launch; records 64 B per hit: not HBM-bound",
                "achieved_gbs": (traffic / (k_ms * 1e-3) * 1e-9) if traffic else None,
                "achieved_is": "DRAM bytes per launch (ncu, roofline.traffic) / the live kernel time: trajectory tiles, tables "
                               "and collision records together",
                "peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0},
    }

    # ---- the reference's CPU path on this box's host cores (bounded samples)
    cpu = cpu_baseline(20000, tmp)
    for k in ("seconds", "tests", "trials"):
        cpu.pop(k, None)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(world, tracks[0]),
        "trials_per_s": trials / secs, "drone_steps_per_s": dsteps / secs,
        "tests_per_s_incl_shared_stage1": (tests + shared) / secs,
        "tests_evaluated_per_step": tests / args.steps, "hits_in_timed_steps": stats_sum["hits"],
        "candidates_in_timed_steps": stats_sum["candidates"], "level2_in_timed_steps": stats_sum["level2"],
        "slab_survivors_in_timed_steps": stats_sum["slab_survivors"],
        "wall_ms_timed_region": wall * 1e3,
        "reference_guard": guard,
        "back_to_back": back_to_back,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "call": "dcrp_simulate (host buffers in/out; DCRP_FLAG_NO_TIMING, as host/Drone.cpp calls it)",
                "ms_per_step": 1e3 * e2e_secs / args.steps,
                "scenario_cache": "%d of %d timed calls found the scenario resident (bit-for-bit comparison with the pinned "
                                  "mirror of the last upload): h2d_bytes_per_step = 0 is the truth for them; the first call "
                                  "of the loop uploaded %d bytes in %.3f ms" % (
                                      cached_steps, args.steps, first_call["h2d_bytes"], first_call["ms"]),
                "new_scenario_every_call": {"ms_per_step": cold_ms, "h2d_bytes_per_step": cold_h2d,
                                            "tests_per_s": world * tests / args.steps / (cold_ms * 1e-3)}},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "reach_cull": reach_cull,
        "substep": substep,
        "strong_scaling": strong,
        "collective": comm_info,
        "device": info,
        "per_rank": per_rank,          # last: survives a truncated tail
    }
    emit(line)
    if world > 1:
        dist.barrier()
        eng.detach_comm()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dcrp", choices=["dcrp", "reference"])
    ap.add_argument("--no-strong", action="store_true", help="skip the 1e10-trial strong-scaling job")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return engine_arm(args)


if __name__ == "__main__":
    sys.exit(main())
