"""Development aid: per-instruction stall samples of a kernel's hot loop from an .ncu-rep (source page).
usage: python scripts/ncu_loop.py <report.ncu-rep> <kernel regex> [min fraction of max exec count, default 0.5] [--all]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 0.5
show_all = "--all" in sys.argv
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__cycles_active.min", "sm__cycles_elapsed.max", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed_op_ldgsts.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:70s} {units[i]:10s}", [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r)][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] in seen:
        continue
    seen.add(r[0]); data.append(r)
def num(r, k):
    try: return int(r[ix[k]])
    except Exception: return 0
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, {k[6:]: sum(num(r, k) for r in data) for k in reasons if sum(num(r, k) for r in data)})
mx = max(num(r, "Instructions Executed") for r in data)
loop = [r for r in data if num(r, "Instructions Executed") >= frac * mx]
print("max exec", mx, "loop instrs", len(loop), "samples in loop", sum(num(r, "# Samples") for r in loop))
print({k[6:]: sum(num(r, k) for r in loop) for k in reasons if sum(num(r, k) for r in loop)})
for r in loop:
    n = num(r, "# Samples")
    if show_all or n >= 8:
        rs = [(k[6:], num(r, k)) for k in reasons if num(r, k) > 0]
        print(r[0][-5:], n, num(r, "Instructions Executed"), r[1][:80], rs)
