"""Instructions executed / stall samples per SOURCE LINE of a kernel from an .ncu-rep captured with --import-source on
(read on the CPU box).  The per-role instruction budget of a warp-specialised kernel falls out of the line ranges.
  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--file conv_tc.cu] [--launch 0] [--top 40]"""
import argparse, csv, io, subprocess, collections

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--file", default="conv_tc.cu")
ap.add_argument("--launch", type=int, default=0)
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--ranges", default="", help="comma separated lo-hi:name line ranges to total, e.g. 460-560:mma,560-700:xform")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launch, cur_file, hdr = -1, None, None
per_line = collections.OrderedDict()
seen_kernel_files = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        if cur_file in seen_kernel_files:  # a new launch repeats the file list
            seen_kernel_files = set()
            launch_changed = True
        continue
    if r[0] == "Function Name":
        key = cur_file
        if key in seen_kernel_files:
            pass
        seen_kernel_files.add(key)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur_file is None or not cur_file.endswith(a.file):
        continue
    d = dict(zip(hdr, r))
    # header has two "Source" columns (cuda, sass); csv.reader keeps positional order: use indices
    try:
        line_no = r[0]
        ie = hdr.index("Instructions Executed")
        smp = hdr.index("# Samples")
    except ValueError:
        continue
    if line_no.isdigit():
        cur_line = int(line_no)
        per_line.setdefault(cur_line, [r[1].strip(), 0, 0, 0])
    else:
        # sass row of the current source line
        try:
            per_line[cur_line][1] += int(r[ie] or 0)
            per_line[cur_line][2] += int(r[smp] or 0)
            per_line[cur_line][3] += 1
        except (ValueError, KeyError, NameError):
            pass
tot_i = sum(v[1] for v in per_line.values()) or 1
tot_s = sum(v[2] for v in per_line.values()) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s} (all launches in the report are summed)")
if a.ranges:
    for spec in a.ranges.split(","):
        rng, name = spec.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        ii = sum(v[1] for k, v in per_line.items() if lo <= k <= hi)
        ss = sum(v[2] for k, v in per_line.items() if lo <= k <= hi)
        print(f"  {name:12s} lines {lo:4d}-{hi:4d}: {100 * ii / tot_i:5.1f}% of instructions, {100 * ss / tot_s:5.1f}% of samples")
print("hottest lines: instr%  samples%  sass-instrs  line  source")
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[: a.top]:
    print(f"  {100 * v[1] / tot_i:5.1f}% {100 * v[2] / tot_s:5.1f}% {v[3]:4d}  {k:4d}  {v[0][:120]}")
