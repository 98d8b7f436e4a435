"""Pipe micro-benchmarks (dcrp_measure_peak 0..9) -> gpurun_out/micro.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from drone_collision_rp_b200 import capi

names = ["ffma", "ffma2", "dfma", "scan_mix", "fadd2_only", "fmul2_only", "mix_lop3_instead_of_fmnmx3",
         "mix_2xfmnmx", "fmnmx3_only_ginstr", "mix_unpacked", "philox_only_Tcalls", "mix_2xfmnmx_plus_philox_round", "mix_incremental",
         "scan2d_lds128_dup", "scan2d_lds64_bcast", "scan2d_regs_only", "scan2d_const_uniform",
         "l1_euclid_lds_np1", "l1_euclid_lds_np2", "l1_euclid_lds_np4", "l1_cheb_regs_np1", "l1_cheb_lds_np1",
         "l1_cheb_lds_np2", "l1_cheb_lds_np4", "l1_cheb_regs_np2"]
eng = capi.Engine(0)
out = {}
for which, name in enumerate(names):
    tf, ms = capi.measure_peak(0, which)
    out[name] = {"tflops_or_ginstr": tf, "ms": ms}
    print("%-32s %8.3f  (%.3f ms)" % (name, tf, ms), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "micro.json"), "w"), indent=1)
