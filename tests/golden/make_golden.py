"""Regenerates the committed golden fixtures.  Runs in the BUILD container only
(it reads /root/reference and runs the reference rebuilt under oracle/_ref);
the fixtures it writes travel to the GPU box, the reference tree does not.

  ref_collision_blocks.json  the 52 collision blocks the reference ships
                             (Drone_Collisions/Drone_coords_{0..9}.csv), parsed, plus the
                             exact text of the first block (writer format check)
  ref_ring_rows.json         rows the reference binds from drone_inital_positions_*.csv
  ref_replay_vectors.npz     synthetic-day tracks + Philox draws + the UNMODIFIED reference
                             Drone.cpp's answers on those replayed draws (hit, first index,
                             min distance, a few full trajectories), and the reference
                             Aircraft loader's view of the synthetic day
  ref_freerun.json           hit counts of the reference AS SHIPPED (its own random_device
                             RNG) on the synthetic day -- the free-running statistical anchor

usage:  python tests/golden/make_golden.py [--freerun-runs N]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import oracle_api as oa  # noqa: E402
import drone_collision_rp_b200 as pk  # noqa: E402
from drone_collision_rp_b200 import capi, scenario  # noqa: E402

GOLDEN_SEED = 20170630
N_FLIGHTS = 6
PAIRS = [(0, 41), (0, 53), (1, 13), (1, 98), (2, 39), (3, 4), (3, 96), (4, 50), (5, 8)]
TRIALS_PER_PAIR = 200000


def collision_blocks():
    blocks = []
    first_text = None
    for k in range(10):
        path = os.path.join(REF, "Drone_Collisions", "Drone_coords_%d.csv" % k)
        text = open(path).read()
        for chunk in text.split("\n\n"):
            lines = [l for l in chunk.split("\n") if l.strip()]
            if not lines:
                continue
            rows = [l.split(",") for l in lines]
            blk = {"aircraft": k, "run": int(rows[0][3]), "x": [float(r[0]) for r in rows],
                   "y": [float(r[1]) for r in rows], "z": [float(r[2]) for r in rows]}
            blocks.append(blk)
            if first_text is None:
                first_text = chunk + "\n\n"
    return {"source": "reference Drone_Collisions/Drone_coords_{0..9}.csv", "blocks": blocks,
            "first_block_text": first_text}


def ring_rows():
    out = {}
    for name in ("drone_inital_positions_1km.csv", "drone_inital_positions_5km.csv"):
        rows = []
        for line in open(os.path.join(REF, name)).read().split("\n")[1:]:
            cells = line.strip("\r").split(",")
            if len(cells) == 4:
                rows.append([float(c) for c in cells[1:]])
        out[name] = rows
    return out


def replay_vectors(tmp):
    o = oa.Oracle()
    ref = oa.RefReplay()
    refair = oa.RefAircraft()
    tracks, csv = scenario.synthetic_tracks(tmp, N_FLIGHTS, GOLDEN_SEED)
    ring1 = scenario.load_ring(pk.RING_1KM)
    out = {"n_flights": N_FLIGHTS, "seed": GOLDEN_SEED, "pairs": np.array(PAIRS, dtype=np.int32)}
    for a, t in enumerate(tracks):
        r = refair.load(csv, a)
        for key in "xyz":
            assert np.array_equal(r[key], t[key]), "host loader differs from the reference loader"
        assert r["takeoff_t"] == t["takeoff_t"] and r["n_steps"] == t["n_steps"]
        out["track%d_x" % a], out["track%d_y" % a], out["track%d_z" % a] = t["x"], t["y"], t["z"]
        out["track%d_takeoff_t" % a] = np.int32(t["takeoff_t"])
    out["ref_index_size"] = np.int32(refair.load(csv, 0)["index_size"])
    for p, (a, j) in enumerate(PAIRS):
        t = tracks[a]
        box = capi.target_box(t["y"][0], t["z"][0])
        assert np.array_equal(box, ref.cubed_volume(t["y"][0], t["z"][0]))
        draws = o.draws(GOLDEN_SEED, 0, TRIALS_PER_PAIR, j, a, t["takeoff_t"], box)
        r = ref.run(pk.RING_1KM, j, a, t, draws)
        assert r["range_violations"] == 0
        # keep every hit, every near miss (< 25 m) and a thin uniform sample
        keep = np.zeros(TRIALS_PER_PAIR, bool)
        keep[r["min_dist"] < 25.0] = True
        keep[::97] = True
        idx = np.nonzero(keep)[0]
        out["pair%d_trial" % p] = idx.astype(np.int64)
        out["pair%d_draws" % p] = draws[idx]
        out["pair%d_hit" % p] = r["hit"][idx]
        out["pair%d_first_hit" % p] = r["first_hit"][idx]
        out["pair%d_min_dist" % p] = r["min_dist"][idx]
        out["pair%d_total_hits" % p] = np.int64(r["hits"])
        hits = np.nonzero(r["hit"])[0][:3]
        sel = np.concatenate([hits, idx[:2]]).astype(np.int64)
        rt = ref.run(pk.RING_1KM, j, a, t, draws[sel], want_traj=True)
        out["pair%d_traj_trial" % p] = sel
        out["pair%d_traj" % p] = rt["traj"]
        print("pair", p, (a, j), "hits", r["hits"], "kept", len(idx))
    return out


def freerun(tmp, runs):
    tracks, csv = scenario.synthetic_tracks(tmp, N_FLIGHTS, GOLDEN_SEED)
    res = {"note": "reference as shipped (std::random_device), oracle/_ref/ref_asshipped", "seed_of_day": GOLDEN_SEED,
           "n_flights_in_day": N_FLIGHTS, "ring": "drone_inital_positions_1km.csv", "drone_total": 99,
           "runs_per_drone": runs, "aircraft": []}
    for a in (0, 1):
        scratch = os.path.join(tmp, "asshipped%d" % a)
        os.makedirs(scratch, exist_ok=True)
        outp = subprocess.run([oa.REF_ASSHIPPED, csv, "22", str(a), pk.RING_1KM, "99", str(runs),
                               str(os.cpu_count()), scratch], check=True, capture_output=True, text=True).stdout
        j = json.loads(outp.strip().split("\n")[-1])
        print("as shipped, aircraft", a, j)
        res["aircraft"].append({"index": a, "trials": j["trials"], "hits": j["hits"], "seconds": j["seconds"],
                                "nproc": j["nproc"]})
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--freerun-runs", type=int, default=200000)
    ap.add_argument("--skip-freerun", action="store_true")
    args = ap.parse_args()
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)
    with tempfile.TemporaryDirectory() as tmp:
        json.dump(collision_blocks(), open(os.path.join(HERE, "ref_collision_blocks.json"), "w"))
        json.dump(ring_rows(), open(os.path.join(HERE, "ref_ring_rows.json"), "w"))
        np.savez_compressed(os.path.join(HERE, "ref_replay_vectors.npz"), **replay_vectors(tmp))
        if not args.skip_freerun:
            json.dump(freerun(tmp, args.freerun_runs), open(os.path.join(HERE, "ref_freerun.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
