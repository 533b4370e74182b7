#!/usr/bin/env python
"""Summarise an `ncu --set full` report into profiles/: one markdown table per capture plus traffic.json (DRAM bytes
per launch keyed by kernel name, read by bench.py for roofline.traffic).

  python scripts/ncu_summary.py gpurun_out/prof_TAG.ncu-rep profiles/ncu_rNN_TAG.md [--traffic profiles/traffic.json]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stalled warps per issue: long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stalled warps per issue: barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stalled warps per issue: short_scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stalled warps per issue: mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stalled warps per issue: lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stalled warps per issue: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stalled warps per issue: math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stalled warps per issue: not_selected"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stalled warps per issue: membar"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stalled warps per issue: no_instruction"),
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def short(name):
    """'void adfem::k_tile_fwd<(int)2, (int)1, (int)0, (bool)1>(...)' -> 'k_tile_fwd<2,1,LAPLACE,1>' (the names bench.py uses)."""
    m = re.match(r"(?:void )?(?:adfem::)?(\w+)<([^>]*)>", name)
    if not m:
        return name.split("(")[0]
    ops = {"0": "LAPLACE", "1": "MASS", "2": "STIFFNESS"}
    a = [re.sub(r"\((?:int|bool)\)", "", x).strip() for x in m.group(2).split(",")]
    kname = m.group(1)
    if kname.startswith("k_grid"):            # <OP, MINB>
        a = [ops.get(a[0], a[0])]
    elif len(a) >= 3:                         # <DIM, DEG, OP[, KPRE]>
        a[2] = ops.get(a[2], a[2])
    return "%s<%s>" % (kname, ",".join(a))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full --clock-control none: %s" % os.path.basename(rep), ""]
    if note:
        lines += [note, ""]
    names = [short(r[col["Kernel Name"]]) for r in data]
    lines.append("| metric | " + " | ".join(names) + " |")
    lines.append("|---|" + "---|" * len(names))
    traffic = {}
    for key, label in METRICS:
        if key not in col:
            continue
        u = units[col[key]]
        lines.append("| %s (`%s`, %s) | " % (label, key, u or "-") + " | ".join(r[col[key]] for r in data) + " |")
    for r, nm in zip(data, names):
        rd = float(r[col["dram__bytes_read.sum"]]) * SCALE[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * SCALE[units[col["dram__bytes_write.sum"]]]
        us = float(r[col["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[col["gpu__time_duration.sum"]]]
        traffic[nm] = rd + wr
        lines.append("")
        lines.append("* `%s`: DRAM read %.3f GB + write %.3f GB = **%.3f GB per launch**, %.1f us under ncu -> %.0f GB/s of DRAM traffic"
                     % (nm, rd / 1e9, wr / 1e9, (rd + wr) / 1e9, us, (rd + wr) / us / 1e3))
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_path:
        old = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
        old.update(traffic)
        json.dump(old, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
