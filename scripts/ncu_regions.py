"""Aggregate ncu stall samples / instructions per device function of vp_trace.cu (by source line ranges)."""
import csv, re, subprocess, sys
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None      # optional: only this kernel (regex), first launch in the report
src = open('/root/repo/volprim_balance_b200/csrc/vp_trace.cu').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    if l.startswith('__device__') or l.startswith('__global__'):
        m = re.search(r'\b(\w+)\(', re.sub(r'__launch_bounds__\([^)]*\)', '', l))
        if m: marks.append((i, m.group(1)))
def fn(line):
    name = 'header'
    for i, n in marks:
        if line >= i - 3: name = n
    return name
flt = ["-k", "regex:" + kern, "-c", "1"] if kern else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, R = None, None, {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0].isdigit() and hdr:
        d = dict(zip(hdr, r))
        try: s = int(d['Warp Stall Sampling (All Samples)'])
        except Exception: continue
        k = fn(int(r[0])) if cur == 'vp_trace.cu' else cur
        a = R.setdefault(k, [0, 0.0, 0.0]); a[0] += s; a[1] += float(d.get('Instructions Executed', 0) or 0); a[2] += float(d.get('Thread Instructions Executed', 0) or 0)
tot = sum(v[0] for v in R.values()); ti = sum(v[1] for v in R.values())
for k, v in sorted(R.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*v[0]/tot:5.1f}% samples {100*v[1]/ti:5.1f}% warp-inst lanes {v[2]/max(v[1],1):4.1f}  {k}")
