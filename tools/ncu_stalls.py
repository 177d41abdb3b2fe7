"""Stall-reason samples per source-line range (warp role) of a kernel from an .ncu-rep captured with --import-source on.
  python tools/ncu_stalls.py rep.ncu-rep --ranges 553-631:producer,632-730:mma,... [--file conv_tc.cu]"""
import argparse, csv, io, subprocess, collections
ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("--file", default="conv_tc.cu"); ap.add_argument("--ranges", required=True)
ap.add_argument("--top", type=int, default=25)
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur = None; in_file = False
per_line = collections.defaultdict(lambda: collections.Counter())
src = {}
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"):
        in_file = r[1].endswith(a.file); continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not in_file: continue
    if r[0].isdigit():
        cur = int(r[0]); src[cur] = r[1].strip(); continue
    if cur is None: continue
    for i, h in enumerate(hdr):
        if i < len(r) and (h.startswith("stall_") and "Not Issued" not in h or h in ("# Samples", "Instructions Executed")):
            try: per_line[cur][h] += int(r[i] or 0)
            except ValueError: pass
tot = sum(v["# Samples"] for v in per_line.values()) or 1
toti = sum(v["Instructions Executed"] for v in per_line.values()) or 1
print(f"total samples {tot}, warp-instructions {toti}")
for spec in a.ranges.split(","):
    rng, name = spec.split(":"); lo, hi = (int(x) for x in rng.split("-"))
    c = collections.Counter()
    for k, v in per_line.items():
        if lo <= k <= hi: c.update(v)
    s = c["# Samples"]
    top = ", ".join(f"{k[6:]} {100*v/max(s,1):.0f}%" for k, v in sorted(((k, v) for k, v in c.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:6])
    print(f"  {name:14s} {lo:4d}-{hi:4d}: instr {100*c['Instructions Executed']/toti:5.1f}%  samples {100*s/tot:5.1f}%  | {top}")
print("hottest lines by samples:")
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1]["# Samples"])[: a.top]:
    top = ", ".join(f"{kk[6:]} {vv}" for kk, vv in sorted(((kk, vv) for kk, vv in v.items() if kk.startswith("stall_")), key=lambda kv: -kv[1])[:3])
    print(f"  {100*v['# Samples']/tot:5.1f}%  instr {100*v['Instructions Executed']/toti:4.1f}%  line {k:4d}  {src.get(k,'')[:90]}   [{top}]")
