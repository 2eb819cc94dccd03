#!/usr/bin/env python3
"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full`) into a small JSON + markdown under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/ncu_r01_step_dyn_capsule

Reads the report here (no GPU needed): `ncu -i <rep> --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "no_instruction", "barrier", "not_selected", "branch_resolving",
               "math_pipe_throttle", "dispatch_stall", "lg_throttle", "mio_throttle"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in KEYS.items():
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    d[name] = float(v)
                except ValueError:
                    d[name] = v
                u = units[hdr.index(k)]
                if name in ("dram_read", "dram_write"):
                    d[name + "_unit"] = u
        d["stalls_warps_per_issue"] = {}
        for s in STALL_NAMES:
            k = STALLS % s
            if k in hdr:
                try:
                    d["stalls_warps_per_issue"][s] = float(r[hdr.index(k)])
                except ValueError:
                    pass
        kernels.append(d)
    json.dump({"report": rep, "kernels": kernels}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep}` (ncu --set full --clock-control none; per-launch, cold cache, serialised)\n\n")
        for d in kernels:
            f.write(f"## {d['kernel']}\n\n")
            for k, v in d.items():
                if k not in ("kernel", "stalls_warps_per_issue"):
                    f.write(f"- {k}: {v}\n")
            f.write("- stalls (warps per issue slot): " + ", ".join(f"{k} {v:.2f}" for k, v in d["stalls_warps_per_issue"].items()) + "\n\n")
    print("wrote", out + ".json", out + ".md")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
