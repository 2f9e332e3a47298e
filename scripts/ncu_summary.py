#!/usr/bin/env python
"""Summarise ncu outputs into profiles/ (tracked).
  python scripts/ncu_summary.py full   gpurun_out/X.ncu-rep   profiles/NAME.md
  python scripts/ncu_summary.py launch gpurun_out/X_launches.csv profiles/NAME.md [skip_launches]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n(per launch; `ncu -i <rep> --page raw --csv`; clocks not locked)\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:90]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in FULL_METRICS:
                if m in idx:
                    f.write(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |\n")
            if "dram__bytes_read.sum" in idx:
                def mb(v, u):
                    v = float(v)
                    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
                t = mb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + mb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
                f.write(f"| traffic (dram read+write) | {t:.3f} | Mbyte |\n")
            f.write("\n")
    print("wrote", out)


def launch(csvf, out, skip=0):
    lines = [l for l in open(csvf) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"][skip:]
    agg = OrderedDict()
    tot = 0.0
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v_us = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1e-3)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v_us
        tot += v_us
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{csvf}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                f"(cold-cache, serialised: compare SHARES). {len(rows)} launches after skipping {skip}.\n\n"
                "| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
        f.write(f"| **all** | {len(rows)} | {tot:.1f} | 100% |\n")
    print("wrote", out)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        launch(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
