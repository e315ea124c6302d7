"""Aggregate ncu stall samples / instructions per device function of vp_trace.cu (by source line ranges)."""
import csv, re, subprocess, sys
rep = sys.argv[1]
src = open('/root/repo/volprim_balance_b200/csrc/vp_trace.cu').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:__device__ __forceinline__|__global__|template <.*>\s*$)?.*?\b(exact_isect|fast_isect|slab|list_insert_key|list_insert|drain_list|walk_ray|tile_capsule|tile_prism|prism_may_hit|capsule_children|walk_tile|sh_basis|sh_color|rf_eval|gauss_density_integral|epan_density_integral|srgb_to_linear|ray_index|flush_counters|k_trace_forward|rf_adjoint_hit|tomo_adjoint_hit|k_trace_adjoint|k_raygen)\(', l)
    if m and (l.startswith('__device__') or l.startswith('__global__')): marks.append((i, m.group(1)))
def fn(line):
    name = 'header'
    for i, n in marks:
        if line >= i - 3: name = n
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
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
