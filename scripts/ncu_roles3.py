#!/usr/bin/env python
"""Development aid: split an ncu source-page CSV of the v2-style kernel (producer / probe / accumulate regions, boundaries =
USETMAXREG instructions) and print stall shares + hottest instructions.   usage: ncu_roles3.py rep.ncu-rep [top_n]"""
import csv, subprocess, sys, tempfile, os
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tmp = tempfile.mkdtemp()
subprocess.run(f"ncu -i {rep} --page source --csv > {tmp}/src.csv 2>/dev/null", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page raw --csv > {tmp}/raw.csv 2>/dev/null", shell=True, check=True)
rows = list(csv.reader(open(f"{tmp}/src.csv"))); hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
bounds = [n for n, r in enumerate(data) if 'USETMAXREG' in r[ix['Source']]]
names = ['pre', 'producer', 'probe', 'accum'][: len(bounds) + 1]
edges = [0] + bounds + [len(data)]
regions = [(names[i], edges[i], edges[i + 1]) for i in range(len(names))]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(rows[0][1][:100])
for name, a, b in regions:
    s = {k: sum(f(data[n], k) for n in range(a, b)) for k in stalls}
    tot = sum(s.values()); ex = sum(f(data[n], 'Instructions Executed') for n in range(a, b))
    print(f"{name}: samples {tot:.0f}, warp-instr {ex:.0f}")
    if tot: print("   " + ", ".join(f"{k[6:]} {100*v/tot:.0f}%" for k, v in sorted(s.items(), key=lambda x: -x[1])[:8]))
top = sorted(range(len(data)), key=lambda n: -f(data[n], '# Samples'))[:topn]
for n in sorted(top):
    r = data[n]; st = max(stalls, key=lambda k: f(r, k))
    role = [nm for nm, a, b in regions if a <= n < b][0]
    print(f"{n:5d} {role[:4]:4s} {int(f(r,'# Samples')):5d} {int(f(r,'Instructions Executed')):8d} {st[6:]:16s} {r[ix['Source']][:72]}")
rr = list(csv.reader(open(f"{tmp}/raw.csv"))); h, u, v = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, k in enumerate(h):
    if k in want: print(f"{k:70s} {v[i]:>14s} {u[i]}")
