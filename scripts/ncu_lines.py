"""Aggregate an ncu report's warp-stall samples by CUDA source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg = None, None, []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0].isdigit() and hdr:
        d = dict(zip(hdr, r))
        try: s = int(d['Warp Stall Sampling (All Samples)'])
        except Exception: continue
        ie, te = float(d.get('Instructions Executed', 0) or 0), float(d.get('Thread Instructions Executed', 0) or 0)
        agg.append((s, cur, int(r[0]), r[1].strip()[:100], ie, te))
tot = sum(a[0] for a in agg); ti = sum(a[4] for a in agg)
print('total samples', tot, 'warp insts', ti)
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[4]/ti:5.1f}% {a[1]}:{a[2]:4d} thr/inst={a[5]/a[4] if a[4] else 0:4.1f} {a[3]}")
