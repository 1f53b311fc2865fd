#!/usr/bin/env python3
"""Key metrics per kernel from an `ncu --page raw --csv` export (one row per captured launch)."""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"), ("launch__shared_mem_per_block_static", "ssmem"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("launch__grid_size", "grid"), ("launch__occupancy_limit_registers", "lim_regs"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
seen = {}
for r in rows[2:]:
    name = r[ki].split("(")[0]
    if name in seen and "--all" not in sys.argv:
        continue
    seen[name] = 1
    print("==", name)
    out = []
    for k, label in KEYS:
        if k in hdr:
            out.append(f"{label}={r[hdr.index(k)]}{units[hdr.index(k)]}")
    print("  ", "  ".join(out))
    st = [(h[len(STALL):].replace("_per_issue_active.ratio", ""), float(r[i])) for i, h in enumerate(hdr)
          if h.startswith(STALL) and h.endswith("per_issue_active.ratio")]
    st.sort(key=lambda kv: -kv[1])
    print("   stalls/issue:", "  ".join(f"{k}={v:.2f}" for k, v in st[:8]))
