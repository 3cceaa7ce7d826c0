#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/ (tracked).
   python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
   python tools/summarize_ncu.py kernels  gpurun_out/prof_matmult_r1.ncu-rep profiles/r1_matmult.txt
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        name = re.sub(r"Layout<1, 1, 2, 4, 0, 1, 1>", "L310", name)
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {src}\n")
        f.write(f"# one bench step (solar + thermal band, 256x256x80, 3_10); total kernel time {tot:.1f} ms\n")
        f.write("%-58s %6s %10s %9s %7s\n" % ("kernel", "n", "ms", "avg_ms", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-58s %6d %10.3f %9.4f %6.1f%%\n" % (k[:58], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.pct", "smsp__average_warp_latency_issue_stalled_lg_throttle.pct",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.pct", "smsp__average_warp_latency_issue_stalled_barrier.pct",
        "smsp__average_warp_latency_issue_stalled_wait.pct", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.pct",
        "smsp__average_warp_latency_issue_stalled_drain.pct", "smsp__average_warp_latency_issue_stalled_not_selected.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def kernels(src, dst):
    # src: an .ncu-rep, or the output of `ncu -i <rep> --page raw --csv` taken on the GPU box (reports of many kernels exceed
    # what gpurun copies back)
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on : {src}\n")
        for row in r[2:]:
            name = row[hdr.index("Kernel Name")]
            name = re.sub(r"Layout<1, 1, 2, 4, 0, 1, 1>", "L310", name)
            f.write(f"\n== {name[:110]}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("   %-66s %16s %s\n" % (w, row[i], units[i]))
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2], sys.argv[3])
