"""Key metrics of `ncu -i x.ncu-rep --page raw --csv` exports (one row per captured launch): duration, DRAM bytes and
throughput, achieved HBM GB/s against MEASURED_PEAKS.json, tensor-pipe and issue activity, registers, occupancy."""
import csv
import json
import os
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print("%s: no captured launch" % path)
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("== %s" % os.path.basename(path))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        t = num(r[ix["gpu__time_duration.sum"]])
        t_us = t / 1e3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else t
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = num(r[ix["dram__bytes_read.sum"]]) * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
        wr = num(r[ix["dram__bytes_write.sum"]]) * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
        gbs = (rd + wr) / (t_us * 1e-6) / 1e9
        print("  %s" % name[:110])
        print("     duration %.1f us | dram read %.1f MB + write %.1f MB = %.0f GB/s (%.2f of the measured %.0f GB/s copy peak) | dram %s %% | sm %s %% | "
              "tensor pipe %s %% | issue active %s %% | warps active %s %% | regs %s | grid %s x %s"
              % (t_us, rd / 1e6, wr / 1e6, gbs, gbs / hbm, hbm, r[ix[KEYS[3]]], r[ix[KEYS[4]]], r[ix[KEYS[5]]], r[ix[KEYS[6]]], r[ix[KEYS[7]]],
                 r[ix[KEYS[8]]], r[ix[KEYS[9]]], r[ix[KEYS[10]]]))
