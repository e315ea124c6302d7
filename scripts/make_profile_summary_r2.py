"""Summarise the round-2 ncu captures into profiles/ (tracked).
Usage: python scripts/make_profile_summary_r2.py <tag> <launches.csv> <forward.ncu-rep> <adjoint.ncu-rep> [bench.json]
  launches.csv     `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of the bench command
  forward.ncu-rep  one launch of k_trace_forward (sections + DRAM byte counters + source)
  adjoint.ncu-rep  one launch each of the recording forward and of the adjoint passes"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

tag, launches, fwd_rep, adj_rep = sys.argv[1:5]
bench_json = sys.argv[5] if len(sys.argv) > 5 else None
out = [f"# ncu summary {tag}  (kernel sources stamp {bench.source_stamp()})\n"]

rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn, mv = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:70]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
out.append(f"## Launch list (`{os.path.basename(launches)}`: `ncu --metrics gpu__time_duration.sum --clock-control none` on "
           "`python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras`; cold-cache, serialised -- compare SHARES)\n")
out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} % |")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}


def report(rep, title, how):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    out.append(f"\n## {title} (`{os.path.basename(rep)}`)\n\n{how}\n")
    traffic = {}
    for vals in r[2:]:
        name = vals[hdr.index('Kernel Name')].replace('void ', '').replace('<unnamed>::', '').split('(')[0][:70]
        out.append(f"### `{name}`\n")
        for m in METRICS:
            if m in hdr and vals[hdr.index(m)] not in ("", "n/a"):
                out.append(f"- `{m}` = {vals[hdr.index(m)]} {units[hdr.index(m)]}")
        try:
            rd = float(vals[hdr.index("dram__bytes_read.sum")].replace(',', '')) * SCALE.get(units[hdr.index("dram__bytes_read.sum")], 1)
            wr = float(vals[hdr.index("dram__bytes_write.sum")].replace(',', '')) * SCALE.get(units[hdr.index("dram__bytes_write.sum")], 1)
            out.append(f"- DRAM traffic of this launch = {(rd + wr) / 1e9:.3f} GB")
            traffic[name] = rd + wr
        except Exception:
            pass
        out.append("")
    return traffic


HOW = ("Captured with `ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section Occupancy --section "
       "LaunchStats --section MemoryWorkloadAnalysis --section SpeedOfLight --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
       "gpu__time_duration.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum --clock-control none` on "
       "`python scripts/ab.py cfg2` (workload = bench.py's cfg2), after the same command had exited 0 without ncu.  (The section list "
       "replaces `--set full`: a full-set capture replays the 1.4 GB working set of these kernels ~40 times.)")
t = report(fwd_rep, "k_trace_forward (plain, then recording) and the adjoint passes, one launch each" if fwd_rep == adj_rep
           else "Top kernel: k_trace_forward, one launch", HOW)
fwd_traffic = next((v for k, v in t.items() if "k_trace_forward" in k), None)   # first match = the plain (bench) instantiation
if adj_rep != fwd_rep:
    report(adj_rep, "Recording forward + the adjoint passes, one launch each", HOW)
if fwd_traffic is not None:
    json.dump({"k_trace_forward_dram_bytes_per_launch": fwd_traffic, "source": os.path.basename(fwd_rep),
               "source_stamp": bench.source_stamp()}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"))
for kern in ["k_trace_forward", "k_adjoint_rows_dense", "k_adjoint_gather"]:
    reg = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), fwd_rep, kern], capture_output=True, text=True).stdout
    out.append(f"\n## {kern}: warp-stall samples per device function (source-correlated, -lineinfo)\n\n```\n" + reg + "```")
if bench_json and os.path.exists(bench_json):
    out.append("\n## Bench line of the same build (no profiler attached)\n\n```json\n" + open(bench_json).read().strip().splitlines()[-1] + "\n```")
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
