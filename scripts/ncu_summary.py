"""Summarise one `ncu --set full` report as JSON (development aid): python scripts/ncu_summary.py report.ncu-rep [note]"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, unit, val = rows[0], rows[1], rows[2]
    d = {"Kernel Name": val[head.index("Kernel Name")]}
    for k in WANT:
        if k in head:
            i = head.index(k)
            d[k] = f"{val[i]} {unit[i]}".strip()
    stalls = {}
    for i, k in enumerate(head):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(val[i])
    d["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = head.index(k)
        tot += float(val[i]) * SCALE[unit[i]]
    d["dram_bytes_per_launch"] = tot
    if len(sys.argv) > 2:
        d["note"] = sys.argv[2]
    print(json.dumps(d, indent=1))

main()
