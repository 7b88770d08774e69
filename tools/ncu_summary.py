"""Per-launch summary of an ncu report: duration, DRAM bytes, achieved DRAM GB/s against the measured (6551 GB/s) and
nominal (8000 GB/s) HBM peaks, tensor-pipe activity, registers.
    ncu -i file.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [--json out.json]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, key, default=0.0):
    i = col.get(key)
    if i is None or r[i] in ("", "n/a"):
        return default
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9,
             "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(u, 1.0)
    return v * scale


out = []
print(f"{'kernel':46s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'%6551':>6s} {'%8000':>6s} {'tensor%':>8s} {'regs':>5s}")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("sg::", "")[:46]
    t = num(r, "gpu__time_duration.sum")
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / t / 1e9 if t else 0.0
    tens = num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    regs = num(r, "launch__registers_per_thread")
    print(f"{short:46s} {t * 1e6:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:8.0f} {100 * gbs / 6551.4:6.1f} {100 * gbs / 8000:6.1f} {tens:8.1f} {regs:5.0f}")
    out.append({"kernel": short, "us": t * 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_gbs": gbs,
                "tensor_pipe_active_pct": tens, "registers": regs})
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
