#!/usr/bin/env python
"""Turn one `ncu --set full` capture of a bench.py step into the tracked summary + profiles/traffic.json.

    ncu --set full --clock-control none -k regex:msda_fwd -s 12 -c 12 -f -o gpurun_out/x/bench_full \
        python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline
    python tools/ncu_bench_summary.py gpurun_out/x/bench_full.ncu-rep profiles/r01_bench_ncu_full_summary.csv [frames]

One row per launch with the counters DESIGN.md quotes; traffic.json gets the mean DRAM bytes (read + write) of the
encoder launches (the long ones: Lq = S), which bench.py reports as roofline.traffic.
"""
import csv
import json
import os
import subprocess
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    frames = sys.argv[3] if len(sys.argv) > 3 else "8"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow(["ns" if hdr[i] == "gpu__time_duration.sum" else ("byte" if units[i].endswith("byte") else units[i]) for i in idx])
        enc, recs = [], []
        for r in rows[2:]:
            vals = []
            for i in idx:
                v = r[i]
                if units[i] in SCALE and hdr[i] != "Kernel Name":
                    v = "%.0f" % (float(v.replace(",", "")) * SCALE[units[i]])
                vals.append(v)
            w.writerow(vals)
            rec = dict(zip([hdr[i] for i in idx], vals))
            recs.append(rec)
        # An encoder launch (Lq = S) is the TMA window kernel plus the register-gather kernel that follows it for the
        # coarse query levels (round 2 default), or one long register-gather kernel (round 1 / small launches).
        tmax = max(float(r["gpu__time_duration.sum"]) for r in recs)
        dram = lambda r: float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"])
        i = 0
        while i < len(recs):
            rec = recs[i]
            if "pipelined" in rec["Kernel Name"]:
                t = dram(rec)
                if i + 1 < len(recs) and "fast_kernel" in recs[i + 1]["Kernel Name"]:
                    t += dram(recs[i + 1])
                    i += 1
                enc.append(t)
            elif float(rec["gpu__time_duration.sum"]) > 0.5 * tmax:
                enc.append(dram(rec))
            i += 1
    if enc:
        tj = os.path.join(os.path.dirname(os.path.abspath(out)), "traffic.json")
        json.dump({"encoder_dram_bytes_per_launch_at_frames": {frames: sum(enc) / len(enc)},
                   "source": "%s: mean dram__bytes_read.sum + dram__bytes_write.sum over the %d encoder launches of one "
                             "bench.py step (ncu --set full --clock-control none, F=%s)" % (
                                 os.path.relpath(out, os.path.dirname(os.path.dirname(os.path.abspath(out)))), len(enc), frames)},
                  open(tj, "w"), indent=1)
        print("encoder launches: %d, mean DRAM bytes %.1f MB" % (len(enc), sum(enc) / len(enc) / 1e6))


if __name__ == "__main__":
    main()
