"""profiles/traffic.json from an ncu launch list (csv) of tools/profile_step.py captured with
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
One eager guided Euler step launches every conv of both U-Nets once (RGFM_PROFILE_REPS=1): the per-launch DRAM traffic of
the dominant kernel family (conv_tc) is the mean over those launches, next to the algorithmic bytes bench.py counts.
  python tools/make_traffic.py gpurun_out/traffic_launches.csv profiles/traffic.json"""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = d["Metric Unit"]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)
    per.setdefault(d["ID"], {"name": d["Kernel Name"]})[d["Metric Name"]] = v * mult
convs = [v for v in per.values() if "conv_tc_kernel" in v["name"] and "dram__bytes_read.sum" in v]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 68  # conv launches of one eager guided step (the first n in launch order)
convs = convs[:n]
rd = sum(c["dram__bytes_read.sum"] for c in convs)
wr = sum(c["dram__bytes_write.sum"] for c in convs)
us = sum(c["gpu__time_duration.sum"] for c in convs)
out = dict(dram_bytes_per_launch=(rd + wr) / len(convs), kernel="conv_tc_kernel", launches=len(convs), dram_read_bytes_total=rd, dram_write_bytes_total=wr,
           duration_us_total_under_ncu=us, source="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum on tools/profile_step.py "
           "(one eager guided Euler step, B=1024, fp16); mean over the step's conv launches",
           note="cold-cache, serialised launches: every tensor is read from DRAM, so this is an upper bound of the in-graph traffic")
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
