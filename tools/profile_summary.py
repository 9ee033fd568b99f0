#!/usr/bin/env python3
"""Turns the raw ncu artefacts a gpurun call brought back (gpurun_out/, scratch) into the
summaries kept under profiles/ (tracked):

  python tools/profile_summary.py <tag> <launches.csv> <full.ncu-rep> [bench.json]
  (PROFILE_WORKLOAD = the bench workload the captures were taken on, default "encode";
   PROFILE_LAUNCH_CMD / PROFILE_FULL_CMD = the commands, quoted in the summaries)

  profiles/<tag>_launches.csv / _launches_summary.md   per-kernel totals of the launch list
  profiles/<tag>_kernels_full.md                        key metrics of the `--set full` captures
  profiles/<tag>_traffic.json                           DRAM bytes per launch of the dominant kernel
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

RAW = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("static smem/block", "launch__shared_mem_per_block_static"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("ALU pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("LSU pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("FMA pipe (IMAD) %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("instruction cache hit %", "sm__icc_requests_lookup_hit.avg.pct"),
    ("stall barrier / issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall short scoreboard (smem) / issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall long scoreboard (global) / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall math pipe / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall no instruction / issue", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall mio throttle / issue", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
]


def short(name):
    name = name.replace("void ", "").replace("xvcb::", "")
    return name.split("(")[0].replace("(int)", "")


def launches(tag, path, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(short(r[ik]), []).append(float(r[iv].replace(",", "")) / 1000.0)
    total = sum(sum(v) for v in per.values())
    shutil.copy(path, os.path.join(PROF, tag + "_launches.csv"))
    with open(os.path.join(PROF, tag + "_launches_summary.md"), "w") as f:
        f.write("# Launch list, %s\n\n`%s`\n(per-launch times under ncu are serialised and cold-cache -- shares, not absolutes; the T/Q kernels overlap on one side stream per shape class and the three sub-pel classes on 3 streams in the real step).\n\n" % (tag, cmd))
        f.write("| kernel | launches | total us | share | us / launch |\n|---|---|---|---|---|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f %% | %.1f |\n" % (k, len(v), sum(v), 100 * sum(v) / total, sum(v) / len(v)))


def full(tag, rep, cmd):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    kernels, seen = [], set()
    for k in rows[2:]:              # one column per kernel (first capture of each), the search kernel first
        if short(k[ik]) not in seen:
            seen.add(short(k[ik]))
            kernels.append(k)
    kernels.sort(key=lambda k: "tz_search" not in k[ik])
    with open(os.path.join(PROF, tag + "_kernels_full.md"), "w") as f:
        f.write("# `ncu --set full` captures, %s\n\n`%s`\n(one 1080p picture of bench.py, workload '%s'; the .ncu-rep stays in gpurun_out/, scratch).\n\n" % (tag, cmd, WORKLOAD))
        f.write("| metric | " + " | ".join("`%s`" % short(k[ik]) for k in kernels) + " |\n|---|" + "---|" * len(kernels) + "\n")
        for label, m in RAW:
            if m not in hdr:
                continue
            i = hdr.index(m)
            f.write("| %s | " % label + " | ".join("%s %s" % (k[i], units[i]) for k in kernels) + " |\n")
    dom = kernels[0]
    traffic = sum(float(dom[hdr.index(m)].replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[units[hdr.index(m)]]
                  for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    json.dump({"kernel": short(dom[ik]).replace("_kernel", ""), "dram_bytes_per_launch": traffic,
               "warp_instructions_per_launch": float(dom[hdr.index("smsp__inst_executed.sum")].replace(",", "")),
               "shared_wavefronts_per_launch": float(dom[hdr.index("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")].replace(",", "")),
               "workload": WORKLOAD,
               "source": "profiles/%s_kernels_full.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tag},
              open(os.path.join(PROF, tag + "_traffic.json"), "w"))


WORKLOAD = os.environ.get("PROFILE_WORKLOAD", "encode")

if __name__ == "__main__":
    tag, lcsv, rep = sys.argv[1:4]
    launches(tag, lcsv, os.environ.get("PROFILE_LAUNCH_CMD", "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --gop off --steps 2 --warmup 3"))
    full(tag, rep, os.environ.get("PROFILE_FULL_CMD", "ncu --set full --import-source on --clock-control none -k regex:'tz_search|subpel_team|bi_search|partition_kernel|motion_compensate' -c 12 python bench.py --gop off --steps 1 --warmup 3"))
    if len(sys.argv) > 4:
        shutil.copy(sys.argv[4], os.path.join(PROF, tag + "_bench.json"))
