"""Summarise ncu captures into profiles/ (tracked): launch list shares, key counters of the top kernel, per-function
stall samples.  Usage: python scripts/make_profile_summary.py <tag> <launches.csv> <report.ncu-rep> [bench.json]"""
import collections, csv, json, os, subprocess, sys

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
bench = sys.argv[4] if len(sys.argv) > 4 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = [f"# ncu summary {tag}\n"]
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn, mv = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:80]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out.append(f"## Launch list (`{os.path.basename(launches)}`: `ncu --metrics gpu__time_duration.sum --clock-control none` on "
           "`python bench.py --steps 2 --warmup 1 --no-cpu-baseline`; cold-cache, serialised -- compare SHARES)\n")
out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} % |")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, vals = r[0], r[1], r[2]
get = lambda name: next((f"{vals[i]} {units[i]}" for i, x in enumerate(hdr) if x == name), "n/a")
num = lambda name: next((float(vals[i].replace(',', '')) for i, x in enumerate(hdr) if x == name), None)
out.append(f"\n## Top kernel counters (`{os.path.basename(rep)}`, one launch of `{vals[hdr.index('Kernel Name')][:60]}`)\n")
out.append("Captured with `ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section Occupancy "
           "--section LaunchStats --section MemoryWorkloadAnalysis --section SpeedOfLight --metrics "
           "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --import-source on "
           "-k regex:k_trace_forward -s 3 -c 1` on `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (after the same "
           "command had exited 0 without ncu).  The section list replaces `--set full` because a full-set capture of this "
           "kernel takes ≈8 GPU-minutes of the round's budget (done once earlier in the round: same DRAM traffic, 0.39 GB per "
           "launch); the DRAM byte counters are requested explicitly.\n")
for m in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]:
    out.append(f"- `{m}` = {get(m)}")
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
ur = units[hdr.index("dram__bytes_read.sum")] if "dram__bytes_read.sum" in hdr else ""
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(ur, 1)
uw = units[hdr.index("dram__bytes_write.sum")] if "dram__bytes_write.sum" in hdr else ""
scale_w = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(uw, 1)
traffic = None
if rd is not None and wr is not None:
    traffic = rd * scale + wr * scale_w
    out.append(f"- DRAM traffic per launch = {traffic / 1e9:.3f} GB (algorithmic bytes per launch: see bench line)")
    json.dump({"k_trace_forward_dram_bytes_per_launch": traffic, "source": os.path.basename(rep)},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))
reg = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), rep], capture_output=True, text=True).stdout
out.append("\n## Warp-stall samples per device function (source-correlated, -lineinfo)\n\n```\n" + reg + "```")
if bench and os.path.exists(bench):
    out.append("\n## Bench line of the same build (no profiler attached)\n\n```json\n" + open(bench).read().strip() + "\n```")
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:1500])
