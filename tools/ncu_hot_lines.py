"""Stall samples / executed instructions of one kernel by source line, function and IPM phase.
usage: ncu_hot_lines.py file.ncu-rep [title]   (the capture needs --import-source on and a -lineinfo build)"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]; title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, data = None, None, []          # (file, line, samples, inst)
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1]; hdr = None; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or cur is None: continue
    try:
        ln = int(r[0]); smp = int(r[hdr["# Samples"]] or 0); ins = int(r[hdr["Instructions Executed"]] or 0)
    except (ValueError, IndexError, KeyError):
        continue
    if smp or ins: data.append((cur, ln, smp, ins))
tot_s = sum(d[2] for d in data) or 1; tot_i = sum(d[3] for d in data) or 1
def func_map(path):
    """line -> enclosing function / phase label (definitions at column 0, phase banners '// ===== (n)')"""
    try: src = open(path).read().split("\n")
    except OSError: return {}
    m, name, phase = {}, "?", ""
    for i, l in enumerate(src, 1):
        g = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:QS_HD|__global__|__device__|static|inline)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if g and not l.startswith(" "): name, phase = g.group(1), ""
        p = re.search(r"// =+ \(([0-9/\-]+)\)", l)
        if p and name == "qw_iterate": phase = " (" + p.group(1) + ")"
        p2 = re.search(r"// ---- \(([a-d])\)", l)
        if p2 and name == "qp_warp_solve": phase = " (" + p2.group(1) + ")"
        m[i] = name + phase
    return m
maps = {}
byf = collections.Counter(); byf_i = collections.Counter()
for f, ln, smp, ins in data:
    if f not in maps: maps[f] = func_map(f)
    key = (f.split("/")[-1], maps[f].get(ln, "?"))
    byf[key] += smp; byf_i[key] += ins
print(f"# {title}: warp-stall samples and executed warp instructions by source function / phase\n")
print(f"total samples {tot_s}, total warp instructions {tot_i}\n")
print("| file | function (phase) | samples | share | instructions | share |\n|---|---|---:|---:|---:|---:|")
for key, v in byf.most_common(40):
    print(f"| {key[0]} | {key[1]} | {v} | {100*v/tot_s:.1f}% | {byf_i[key]} | {100*byf_i[key]/tot_i:.1f}% |")
print("\nTop source lines by samples:\n\n| file:line | samples | share | instructions |\n|---|---:|---:|---:|")
for f, ln, smp, ins in sorted(data, key=lambda d: -d[2])[:30]:
    print(f"| {f.split('/')[-1]}:{ln} | {smp} | {100*smp/tot_s:.1f}% | {ins} |")
