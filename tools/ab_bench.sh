#!/bin/bash
# A/B of two builds of libdcrp.so on ONE box (boxes differ by 1-3 %): bench.py --no-strong, alternating, twice each.
#   tools/ab_bench.sh <libA.so> <libB.so> <out prefix>      (bench.py picks the library up from DCRP_LIBDCRP)
A=$1; B=$2; OUT=${3:-gpurun_out/ab}
for i in 1 2; do
  DCRP_LIBDCRP=$A python bench.py --no-strong > ${OUT}_A_$i.json 2>/dev/null
  DCRP_LIBDCRP=$B python bench.py --no-strong > ${OUT}_B_$i.json 2>/dev/null
done
python - "$OUT" <<'PY'
import json, sys
for f in ("A_1", "B_1", "A_2", "B_2"):
    d = json.load(open("%s_%s.json" % (sys.argv[1], f)))
    r = d["roofline"]
    print(f, "step %.5f kernel %.5f sustained %.4f e2e %.5f" % (d["ms_per_step"], r["kernel_ms"], r["sustained"]["frac"], d["e2e"]["ms_per_step"]))
PY
