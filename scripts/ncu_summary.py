#!/usr/bin/env python
"""Turns an `ncu --set full --import-source on` capture of the warp-specialised factor kernel into the markdown summary kept
under profiles/ (key metrics, stall shares split by warp role, hottest SASS instructions) + the DRAM traffic JSON.

usage: scripts/ncu_summary.py <rep.ncu-rep> <out.md> "<title>" "<command>" [traffic.json key]"""
import csv, json, os, subprocess, sys, tempfile

rep, out_md, title, command = sys.argv[1:5]
traffic_key = sys.argv[5] if len(sys.argv) > 5 else None
tmp = tempfile.mkdtemp()
raw, src = os.path.join(tmp, "raw.csv"), os.path.join(tmp, "src.csv")
subprocess.run(f"ncu -i {rep} --page raw --csv > {raw} 2>/dev/null", shell=True, check=True)
subprocess.run(f"ncu -i {rep} --page source --csv > {src} 2>/dev/null", shell=True, check=True)

rr = list(csv.reader(open(raw)))
h, u, v = rr[0], rr[1], rr[2]
val = {k: (v[i], u[i]) for i, k in enumerate(h)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(open(src)))
kernel = rows[0][1]
hdr, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]] or 0)
    except Exception:
        return 0.0


b = next((n for n, r in enumerate(data) if "USETMAXREG.TRY_ALLOC" in r[ix["Source"]]), len(data))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
L = [f"# {title}", "", f"Command (1 x B200, under gpurun): `{command}`", "", f"Kernel: `{kernel[:150]}`", "",
     "ncu replays the launch with cold caches and serialised; numbers printed by a run under ncu are never bench values.", "",
     "| metric | value | unit |", "|---|---:|---|"]
for k in want:
    if k in val:
        L.append(f"| `{k}` | {val[k][0]} | {val[k][1]} |")
L += ["", "## Warp stall sampling by warp role (boundary = `USETMAXREG.TRY_ALLOC`: probe warps before, accumulate warps after)", ""]
for name, rng in (("probe warps", range(0, b)), ("accumulate warps", range(b, len(data)))):
    s = {k: sum(f(data[n], k) for n in rng) for k in stalls}
    tot = sum(s.values()) or 1.0
    ex = sum(f(data[n], "Instructions Executed") for n in rng)
    L.append(f"**{name}**: {tot:.0f} samples, {ex:.0f} warp instructions executed")
    L.append("")
    L.append("| reason | samples | share |")
    L.append("|---|---:|---:|")
    for k, x in sorted(s.items(), key=lambda kv: -kv[1])[:9]:
        L.append(f"| {k} | {x:.0f} | {100 * x / tot:.1f}% |")
    L.append("")
L += ["## Hottest SASS instructions (by samples)", "", "| role | samples | executed | dominant stall | instruction |", "|---|---:|---:|---|---|"]
top = sorted(range(len(data)), key=lambda n: -f(data[n], "# Samples"))[:18]
for n in sorted(top):
    r = data[n]
    st = max(stalls, key=lambda k: f(r, k))
    L.append(f"| {'probe' if n < b else 'accumulate'} | {int(f(r, '# Samples'))} | {int(f(r, 'Instructions Executed'))} | {st} | `{r[ix['Source']].strip()[:70]}` |")
open(out_md, "w").write("\n".join(L) + "\n")
if traffic_key:
    rd = float(val["dram__bytes_read.sum"][0]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[val["dram__bytes_read.sum"][1]]
    wr = float(val["dram__bytes_write.sum"][0]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[val["dram__bytes_write.sum"][1]]
    tj = os.path.join(os.path.dirname(out_md), "traffic.json")
    d = json.load(open(tj)) if os.path.exists(tj) else {}
    d[traffic_key] = rd + wr
    d["dram_bytes_read"], d["dram_bytes_write"], d["source"] = rd, wr, os.path.basename(out_md) + " (ncu --set full, 1 launch, bench.py workload)"
    json.dump(d, open(tj, "w"), indent=1)
print("wrote", out_md)
