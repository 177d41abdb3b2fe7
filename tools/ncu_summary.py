"""Text summary of an .ncu-rep (read on the CPU box): key metrics, stall-reason histogram, hottest source lines.
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/name.txt"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(run([path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d:
                print(f"  {k:82s} {d[k]:>16s} {u[k]}")
        stalls = sorted(((int(float(v)), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()
                         if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "n/a")),
                        reverse=True)
        tot = sum(s for s, _ in stalls) or 1
        print("  warp-state samples:", ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in stalls[:8]))
    src = list(csv.reader(io.StringIO(run([path, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    lines, seen = [], 0
    for r in src:
        if r and r[0] == "Function Name":
            seen += 1
            if seen == 2:
                break
        if len(r) > 7 and r[0].isdigit():
            try:
                lines.append((int(r[7] or 0), int(r[6] or 0), int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    if lines:
        ti = sum(l[0] for l in lines) or 1
        ts = sum(l[1] for l in lines) or 1
        print("  hottest source lines (first launch): instructions%, samples%, line, source")
        for ins, smp, ln, text in sorted(lines, reverse=True)[:14]:
            print(f"    {100 * ins / ti:5.1f}% {100 * smp / ts:5.1f}%  {ln:4d}  {text}")


if __name__ == "__main__":
    main(sys.argv[1])
