#!/usr/bin/env python
"""Development aid: split an ncu source-page CSV of the warp-specialised kernel into probe / accumulate regions
(boundary = USETMAXREG.TRY_ALLOC) and print stall shares + hottest instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; ncu -i rep.ncu-rep --page raw --csv > raw.csv ; ncu_roles.py src.csv raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
b = next(n for n, r in enumerate(data) if 'USETMAXREG.TRY_ALLOC' in r[ix['Source']])
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(rows[0][1][:90])
for name, rng in (('probe', range(0, b)), ('accumulate', range(b, len(data)))):
    s = {k: sum(f(data[n], k) for n in rng) for k in stalls}
    tot = sum(s.values()); ex = sum(f(data[n], 'Instructions Executed') for n in rng)
    print(f"{name}: samples {tot:.0f}, warp instructions executed {ex:.0f}")
    print("   " + ", ".join(f"{k[6:]} {100*v/tot:.0f}%" for k, v in sorted(s.items(), key=lambda x: -x[1])[:9]))
top = sorted(range(len(data)), key=lambda n: -f(data[n], '# Samples'))[:22]
for n in sorted(top):
    r = data[n]; st = max(stalls, key=lambda k: f(r, k))
    print(f"{n:5d} {'P' if n < b else 'A'} {int(f(r,'# Samples')):5d} {int(f(r,'Instructions Executed')):8d} {st[6:]:18s} {r[ix['Source']][:64]}")
if len(sys.argv) > 2:
    rr = list(csv.reader(open(sys.argv[2]))); h, u, v = rr[0], rr[1], rr[2]
    want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "launch__registers_per_thread"]
    for i, k in enumerate(h):
        if k in want: print(f"{k:75s} {v[i]:>16s} {u[i]}")
