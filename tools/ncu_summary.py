#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the committed evidence under
profiles/: a text summary (launch config, time, occupancy, issue utilisation, SIMT efficiency,
cache hit rates, DRAM bytes, stall breakdown, hottest source lines) and, with --json, the numbers
bench.py quotes as `roofline.traffic`.
usage: ncu_summary.py <report.ncu-rep> <out.txt> [--views N] [--json profiles/ncu_summary.json]"""
import csv
import io
import json
import subprocess
import sys

rep, out_txt = sys.argv[1], sys.argv[2]
views = int(sys.argv[sys.argv.index("--views") + 1]) if "--views" in sys.argv else 1
json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def g(name, default=None):
    return m.get(name, (default, ""))[0]


def gf(name):
    try:
        return float(g(name).replace(",", ""))
    except Exception:  # noqa: BLE001
        return float("nan")


def to_bytes(name):
    v, u = m.get(name, ("nan", ""))
    f = float(v.replace(",", "")) if v else float("nan")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    return f * scale


def to_ms(name):
    v, u = m.get(name, ("nan", ""))
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)


dur_ms = to_ms("gpu__time_duration.sum")
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
lines = []
lines.append(f"report            : {rep}")
lines.append(f"kernel            : {g('Kernel Name')}")
lines.append(f"grid x block      : {g('launch__grid_size')} x {g('launch__block_size')}   regs/thread {g('launch__registers_per_thread')}"
             f"   dyn smem/CTA {g('launch__shared_mem_per_block_dynamic')} {m.get('launch__shared_mem_per_block_dynamic', ('',''))[1]}")
lines.append(f"occupancy limits  : regs {g('launch__occupancy_limit_registers')} smem {g('launch__occupancy_limit_shared_mem')} "
             f"warps {g('launch__occupancy_limit_warps')} CTAs/SM;  achieved warps active {gf('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % of 64")
lines.append(f"duration          : {dur_ms:.4f} ms for {views} view(s) of 800x800 = {dur_ms / views:.4f} ms/frame "
             f"(under ncu, --clock-control none)")
lines.append(f"issue utilisation : {gf('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} % of issue slots, "
             f"IPC/SM {gf('sm__inst_executed.avg.per_cycle_active'):.2f}")
lines.append(f"SIMT efficiency   : {gf('smsp__thread_inst_executed_per_inst_executed.ratio'):.2f} active threads / instruction")
lines.append(f"warp instructions : {gf('smsp__inst_executed.sum'):.0f} ({gf('smsp__inst_executed.sum') / views / 1e6:.1f} M per frame)")
lines.append(f"L1 hit rate       : {gf('l1tex__t_sector_hit_rate.pct'):.1f} %   L2 hit rate {gf('lts__t_sector_hit_rate.pct'):.1f} %")
lines.append(f"DRAM traffic      : read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / views / 1e6:.1f} MB per frame; "
             f"{(rd + wr) / (dur_ms * 1e-3) / 1e9:.0f} GB/s = {gf('dram__bytes_read.sum.pct_of_peak_sustained_elapsed'):.1f} % of DRAM peak (read)")
lines.append("stall reasons (warps stalled per issue-active cycle):")
for k, (v, u) in sorted(m.items()):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            f = float(v)
        except ValueError:
            continue
        if f >= 0.1:
            lines.append(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {f:.2f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
try:
    hi = next(i for i, r in enumerate(srows) if "Instructions Executed" in r)
    h = srows[hi]
    iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    ls = []
    for r in srows[hi + 1:]:
        if r[0] != "":
            try:
                ls.append((int(r[0]), r[1].strip()[:96], int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)))
            except ValueError:
                pass
    tI, tS = sum(x[3] for x in ls), sum(x[2] for x in ls)
    lines.append("hottest source lines (vr_march.cuh):  inst%  stall-sample%  active-threads  source")
    for x in sorted(ls, key=lambda x: -x[3])[:18]:
        lines.append(f"   L{x[0]:<4d} {100 * x[3] / tI:5.2f}% {100 * x[2] / tS:6.2f}% {x[4] / max(x[3], 1):5.1f}  {x[1]}")
    lines.append("lines with the most stall samples:")
    for x in sorted(ls, key=lambda x: -x[2])[:8]:
        lines.append(f"   L{x[0]:<4d} {100 * x[3] / tI:5.2f}% {100 * x[2] / tS:6.2f}% {x[4] / max(x[3], 1):5.1f}  {x[1]}")
except StopIteration:
    lines.append("(no source correlation in this report)")
open(out_txt, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if json_out:
    json.dump({"report": rep, "kernel": g("Kernel Name"), "views": views, "duration_ms": dur_ms,
               "dram_bytes_per_frame": (rd + wr) / views, "dram_read_bytes": rd, "dram_write_bytes": wr,
               "l1_hit_pct": gf("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": gf("lts__t_sector_hit_rate.pct"),
               "issue_active_pct": gf("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "threads_per_inst": gf("smsp__thread_inst_executed_per_inst_executed.ratio"),
               "registers": g("launch__registers_per_thread"),
               "warps_active_pct": gf("sm__warps_active.avg.pct_of_peak_sustained_active")}, open(json_out, "w"), indent=1)
