#!/usr/bin/env python3
"""Turn the two ncu outputs of tools/gpu_check.sh into the files committed under profiles/.

    python tools/summarise_ncu.py TAG OUT_PREFIX      e.g.  r1k profiles/r1_k

  gpurun_out/launches_TAG.csv   (ncu --metrics gpu__time_duration.sum ... --csv)  ->  OUT_PREFIX_launches.csv and
                                OUT_PREFIX_step_kernel_share.json (kernels of ONE resident step: the launches
                                between two consecutive L2-flush fills of the timed loop)
  gpurun_out/prof_TAG.ncu-rep   (ncu --set full -k regex:composite_)              ->  OUT_PREFIX_composite_raw.csv,
                                OUT_PREFIX_ncu_summary.json (per kernel: duration, DRAM bytes, warp instructions,
                                issue-slot utilisation, pipes, occupancy, registers, shared memory)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0]


def launches(tag, out):
    src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    shutil.copy(src, out + "_launches.csv")
    seq = [(short(r[4]), r[8], float(r[-1]) / 1e3) for r in rows]            # name, grid, microseconds
    # the timed loop of bench.py fills a 256 MiB buffer (grid 65536) before every step
    marks = [i for i, (n, grid, _) in enumerate(seq) if "FillFunctor<float>" in n and "65536" in grid]
    steps = [seq[a + 1:b] for a, b in zip(marks, marks[1:]) if 5 < b - a - 1 < 40]
    if not steps:
        return None
    step = steps[len(steps) // 2]
    share = {}
    for n, _, us in step:
        share[n] = round(share.get(n, 0.0) + us, 2)
    total = round(sum(share.values()), 2)
    res = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py "
                      "--steps 3 --warmup 3 --no-cpu-baseline",
           "note": "one resident step (cold-cache, serialised under ncu): microseconds per kernel, summed over its "
                   "launches in the step",
           "kernels_us": share, "total_us": total,
           "share": {k: round(v / total, 3) for k, v in sorted(share.items(), key=lambda kv: -kv[1])}}
    json.dump(res, open(out + "_step_kernel_share.json", "w"), indent=1)
    return res


WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__shared_mem_per_block_static": "static_smem_per_block",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_heavy_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_pipe_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}


def full(tag, out):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(out + "_composite_raw.csv", "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        d = {}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                try:
                    d[WANT[h]] = [float(v.replace(",", "")), u]
                except ValueError:
                    pass
        res.setdefault(name, d)
    json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:composite_ -s 6 -c 2 "
                          "python bench.py --steps 3 --warmup 3 --no-cpu-baseline", "kernels": res},
              open(out + "_ncu_summary.json", "w"), indent=1)
    return res


if __name__ == "__main__":
    tag, out = sys.argv[1], os.path.join(ROOT, sys.argv[2])
    sh = launches(tag, out)
    if sh:
        print("step:", sh["total_us"], "us;", list(sh["share"].items())[:4])
    for k, d in full(tag, out).items():
        print(k, {a: b[0] for a, b in d.items() if a in ("duration", "dram_read", "dram_write", "warp_instructions",
                                                         "issue_slots_busy_pct")})
