"""Summarise an .ncu-rep (raw metrics + SASS-level opcode / stall breakdown) as text.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else ""
WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if kre and kre not in name:
        continue
    print("=== " + name[:90])
    for w in WANT:
        if w in hdr:
            print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    extra = [h for h in hdr if ("pipe" in h and "pct_of_peak_sustained_active" in h and h not in WANT)]
    for w in extra:
        print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kre] if kre else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
first_addr = None
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[iE].isdigit():
        continue
    if first_addr is None:
        first_addr = r[0]
    elif r[0] == first_addr:
        break                      # the report repeats the kernel per captured launch
    data.append(r)
tot = sum(int(r[iE]) for r in data)
totS = sum(int(r[iN]) for r in data)
ops, ops_s, st = collections.Counter(), collections.Counter(), collections.Counter()
reg, regs = 0, collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
for r in data:
    src_ = r[iS].strip()
    tok = src_.split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    if op == "BAR":
        reg += 1
    e, n = int(r[iE]), int(r[iN])
    ops[op] += e
    ops_s[op] += n
    regs[reg][0] += e
    regs[reg][1] += n
    regs[reg][3][op] += e
    for h in stalls:
        try:
            v = int(r[hdr.index(h)])
        except ValueError:
            v = 0
        st[h] += v
        regs[reg][2][h[6:]] += v
print("instructions %d  samples %d  sass lines %d" % (tot, totS, len(data)))
print("opcode share (inst%% / samples%%):")
for o, c in ops.most_common(28):
    print("   %-8s %5.1f%% %5.1f%%" % (o, 100 * c / tot, 100 * ops_s[o] / max(1, totS)))
ss = sum(st.values())
print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100 * c / ss) for h, c in st.most_common(9)))
print("regions between BAR.SYNCs (inst% / samples% / top stalls):")
for k, v in regs.items():
    if v[1] * 200 > totS:
        t = sum(v[2].values()) or 1
        print("   region %d: %5.1f%% %5.1f%%  %s" % (k, 100 * v[0] / tot, 100 * v[1] / totS,
              ", ".join("%s %.0f%%" % (h, 100 * c / t) for h, c in v[2].most_common(5))))
        print("        ops: " + ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in v[3].most_common(12)))
