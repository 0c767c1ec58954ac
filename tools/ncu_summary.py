#!/usr/bin/env python
"""Developer tool: turn an ncu report (.ncu-rep, captured with --set full --import-source on) into
the short markdown summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xyz.md
"""
import collections
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print("# %s\n" % title)
print("Source: `ncu --set full --clock-control none --import-source on`, read with `ncu -i ... --page raw/source --csv`.\n")
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("## launch: `%s`\n" % d.get("Kernel Name", "?")[:120])
    print("| metric | value | unit |\n|---|---|---|")
    for k in keys[1:]:
        if k in d and d[k] != "":
            print("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
    print("\nWarp stall reasons (warps per issued instruction):\n")
    print("| reason | ratio |\n|---|---|")
    st = [(h, float(d[h].replace(",", ""))) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]]
    for h, v in sorted(st, key=lambda x: -x[1])[:8]:
        print("| %s | %.3f |" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page lists the launches one after the other, each introduced by a "Kernel Name" row and a header row
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
seen = set()
for b in blocks:
    key = (b["name"], len(b["rows"]), b["rows"][0][1] if b["rows"] else "")
    if key in seen:                                        # every launch is listed twice (the same SASS under two views)
        continue
    seen.add(key)
    ix = {h: i for i, h in enumerate(b["hdr"])}
    byop, ex = collections.Counter(), collections.Counter()
    hot = []
    for r in b["rows"]:
        s = r[ix["Source"]].strip()
        if not s:
            continue
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        n_s, n_e = int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)
        byop[op] += n_s
        ex[op] += n_e
        hot.append((n_s, n_e, s))
    tot, te = sum(byop.values()) or 1, sum(ex.values()) or 1
    print("### SASS of `%s`\n" % b["name"][:110])
    print("%d SASS instructions, %d warp-level instructions executed, %d stall samples.\n" % (len(hot), te, tot))
    print("| opcode | executed | share | stall samples |\n|---|---|---|---|")
    for op, c in ex.most_common(12):
        print("| %s | %d | %.1f%% | %.1f%% |" % (op, c, 100.0 * c / te, 100.0 * byop[op] / tot))
    tma = [op for op in ex if op.startswith(("UBLKCP", "UTMA", "SYNCS"))]
    print("\nTMA / mbarrier opcodes present: %s\n" % (", ".join(sorted(tma)) or "none"))
    print("Hottest instructions (stall samples, share, times executed):\n")
    print("| samples | share | executed | instruction |\n|---|---|---|---|")
    for n_s, n_e, s in sorted(hot, key=lambda x: -x[0])[:top_n]:
        print("| %d | %.1f%% | %d | `%s` |" % (n_s, 100.0 * n_s / tot, n_e, s[:90]))
    print()
