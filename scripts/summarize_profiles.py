#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep + launches.csv into the tracked summaries under profiles/ (run here, no GPU needed)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def raw(rep):
    """Rows of the raw page: from a .ncu-rep (via `ncu -i`) or from a CSV exported on the GPU box (*.raw.csv; reports embed
    the whole module and weigh ~17 MB each, more than gpurun merges back, so only the dominant kernel's report travels)."""
    if rep.endswith(".csv"):
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    os.makedirs(OUT, exist_ok=True)
    gout = os.path.join(ROOT, "gpurun_out")
    traffic = {}
    lines = []
    for f in sorted(os.listdir(gout)):
        if not (f.endswith(".ncu-rep") or f.endswith(".raw.csv")):
            continue
        rows, units = raw(os.path.join(gout, f))
        for d in rows:
            name = re.sub(r"\(.*", "", d.get("Kernel Name", "?"))
            lines.append("== %s   [%s]" % (d.get("Kernel Name", "?"), f))
            for k in KEYS:
                if k in d:
                    lines.append("   %-80s %s %s" % (k, d[k], units.get(k, "")))
            try:
                def tobytes(key):
                    v = float(d[key].replace(",", ""))
                    u = units.get(key, "byte").lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
                short = name.split("::")[-1].split("<")[0].strip()
                if "k_jit_" in d["Kernel Name"]:
                    short = re.sub(r"\(.*", "", d["Kernel Name"]).strip()   # k_jit_grid4 / k_jit_grid1 / k_jit_centers
                elif "GenGrid" in d["Kernel Name"]:
                    short = "k_eval<GenGrid>"
                elif "GenCenters" in d["Kernel Name"]:
                    short = "k_eval<GenCenters>"
                traffic[short] = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
                if "smsp__inst_executed.sum" in d:  # warp instructions per launch (issue-slot roofline in bench.py)
                    traffic[short + ".warp_inst"] = float(d["smsp__inst_executed.sum"].replace(",", ""))
            except Exception:
                pass
    open(os.path.join(OUT, "%s_ncu_full_summary.txt" % TAG), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(OUT, "traffic.json"), "w"), indent=1, sort_keys=True)
    # launch list: per-kernel mean duration and share of one step
    lpath = os.path.join(gout, "launches.csv")
    if os.path.exists(lpath):
        rows = list(csv.DictReader(l for l in open(lpath) if not l.startswith("==")))
        agg = collections.OrderedDict()
        for r in rows:
            n = r["Kernel Name"]
            if "vectorized_elementwise" in n or "at::" in n:
                continue
            key = re.sub(r"\(.*", "", n)
            if "k_eval" in n:
                key = "k_eval<%s>" % ("GenCenters" if "GenCenters" in n else "GenGrid" if "GenGrid" in n else "other")
            agg.setdefault(key, []).append(float(r["Metric Value"].replace(",", "")))
        # steady state: the last occurrences
        out = ["kernel,launches,mean_ns_last3,share_of_step"]
        means = {k: sum(v[-3:]) / len(v[-3:]) for k, v in agg.items()}
        tot = sum(means.values())
        for k, v in agg.items():
            out.append("%s,%d,%.0f,%.3f" % (k, len(v), means[k], means[k] / tot))
        out.append("TOTAL,,%.0f,1.000" % tot)
        open(os.path.join(OUT, "%s_launches_summary.csv" % TAG), "w").write("\n".join(out) + "\n")
        open(os.path.join(OUT, "%s_launches_raw.csv" % TAG), "w").write(open(lpath).read())
        print("\n".join(out))


if __name__ == "__main__":
    main()
