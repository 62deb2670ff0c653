#!/usr/bin/env python
"""Turn what tools/collect_profiles.sh brought back in gpurun_out/ into the tracked summaries
under profiles/ (run here, no GPU needed): key ncu metrics per captured kernel, launch shares,
the bench sweep and the two bench lines. (profiles/traffic_rNN.json, the DRAM bytes per algorithmic
byte of the kernel that streams the input, is written by hand from ncu_frontend_fused_rNN.txt.)"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def raw_page(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    # key metrics of every full capture that came back (tools/collect_profiles.sh)
    for rep in sorted(glob.glob(os.path.join(GO, "prof_*_%s.ncu-rep" % R))):
        tag = os.path.basename(rep)[len("prof_"):-len("_%s.ncu-rep" % R)]
        ks, units = raw_page(rep)
        with open(os.path.join(OUT, "ncu_%s_%s.txt" % (tag, R)), "w") as f:
            for d in ks:
                f.write("== %s  grid %s block %s\n" % (d.get("Kernel Name"), d.get("Grid Size"), d.get("Block Size")))
                for k in KEYS:
                    if k in d:
                        f.write("  %-80s %s %s\n" % (k, d[k], units.get(k, "")))
    # launch shares
    lc = os.path.join(GO, "launches_%s.csv" % R)
    if os.path.exists(lc):
        lines = [l for l in open(lc) if l.startswith('"')]
        rows = list(csv.reader(lines))
        hdr = rows[0]
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = {}
        for r in rows[1:]:
            try:
                v = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            name = r[ik].split("(")[0]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(OUT, "launch_shares_%s.txt" % R), "w") as f:
            f.write("kernel, launches, total (ncu gpu__time_duration.sum, unit as reported), share "
                    "(default bench, cold-cache serialised launches)\n")
            for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write("%-70s %4d %12.1f %5.1f%%\n" % (name[:70], a[0], a[1], 100 * a[1] / tot))
        open(os.path.join(OUT, "launches_%s.csv" % R), "w").writelines(lines)
    # sweep + bench lines
    sweep = []
    for fn in sorted(glob.glob(os.path.join(GO, "sweep_*_%s.json" % R))):
        try:
            sweep.append(json.loads(open(fn).read().strip()))
        except Exception:
            pass
    if sweep:
        json.dump(sweep, open(os.path.join(OUT, "sweep_%s.json" % R), "w"), indent=1)
    for nm in ("bench_default", "bench_reference"):
        fn = os.path.join(GO, "%s_%s.json" % (nm, R))
        if os.path.exists(fn) and open(fn).read().strip().startswith("{"):
            open(os.path.join(OUT, "%s_%s.json" % (nm, R)), "w").write(open(fn).read())
    print("profiles/ updated:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
