"""Summarise ncu outputs for profiles/ (run in the build container; needs only the ncu CLI).

  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep   > profiles/rN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        agg[row["Kernel Name"]][0] += 1
        agg[row["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms total")
    print(f"# (cold-cache, serialised: compare SHARES, not absolutes)")
    print(f"{'us total':>12} {'launches':>8} {'share':>7} {'us/launch':>10}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {v[0]:8d} {100*v[1]/tot:6.1f}% {v[1]/v[0]:10.1f}  {k[:110]}")


def full(path, every=False):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        if r[ni] in seen and not every:   # one instance per kernel name (repeats of a timing loop are identical)
            continue
        seen.add(r[ni])
        print("kernel:", r[ni])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:75s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and r[i]:
                stalls.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per_issue")[0]))
        print("  warp stall reasons (warps per issue-active cycle):",
              ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
