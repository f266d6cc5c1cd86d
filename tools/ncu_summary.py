#!/usr/bin/env python
"""Summarise an .ncu-rep (raw metrics + per-opcode instruction mix + top stall lines) into text for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [frames_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 65536
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:90])
    for i, h in enumerate(hdr):
        if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            print(f"  {h:78s} {units[i]:14s} {r[i]}")
    try:
        t = float(r[hdr.index("gpu__time_duration.sum")])
        unit = units[hdr.index("gpu__time_duration.sum")]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
        rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index("dram__bytes_read.sum")]]
        mulw = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index("dram__bytes_write.sum")]]
        print(f"  -> duration {t_s*1e3:.4f} ms (under ncu, cold), dram traffic {(rd*mul+wr*mulw)/1e6:.1f} MB, "
              f"{frames/t_s/1e6:.1f} M frames/s")
        if "--traffic" in sys.argv:
            import json, os
            out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
            json.dump({"kernel": r[hdr.index("Kernel Name")], "dram_bytes_per_launch": rd * mul + wr * mulw,
                       "dram_read": rd * mul, "dram_write": wr * mulw, "source": os.path.basename(rep)}, open(out, "w"))
    except Exception as ex:
        print("  (derived failed)", ex)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        secs.append((r[1], cur))
        continue
    if cur is not None:
        cur.append(r)
for name, sec in secs[:1]:
    hdr = sec[0]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    byop, samp, tot, nl, tots = collections.Counter(), collections.Counter(), 0, 0, 0
    lines = []
    for r in sec[1:]:
        try:
            e, s = int(r[iE]), int(r[iN])
        except Exception:
            continue
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        byop[op] += e; samp[op] += s; tot += e; tots += s; nl += 1
        lines.append((s, e, r[iS].strip()))
    print(f"== SASS mix for {name[:80]}: {nl} SASS lines, {tot/frames:.0f} warp-instr/frame")
    for op, c in byop.most_common(24):
        print(f"  {op:10s} {c/frames:8.1f}/frame {100*c/tot:5.1f}%  stall-samples {100*samp[op]/max(tots,1):5.1f}%")
    print("== top stall-sample lines")
    for s, e, txt in sorted(lines, reverse=True)[:14]:
        print(f"  {100*s/max(tots,1):5.2f}%  exec/frame {e/frames:6.2f}  {txt[:100]}")
