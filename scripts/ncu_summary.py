#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV/JSON for profiles/.

usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/X_summary.csv [traffic.json dtype]
"""
import csv
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
]  # fmt: skip


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * m
    return v


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["Kernel Name"] + [w for w in WANT if w in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow(["unit"] + [units[idx[c]] for c in cols[1:]])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])
    if len(sys.argv) > 3:
        tj = sys.argv[3]
        try:
            d = json.load(open(tj))
        except Exception:
            d = {}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            dt = "f64" if "<double" in name else "f32"
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            d[dt] = {"kernel": name[:80], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                     "gpu_time_us": float(r[idx["gpu__time_duration.sum"]].replace(",", "")), "source": rep}
        json.dump(d, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
