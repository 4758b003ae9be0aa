"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON for profiles/: headline metrics of one kernel
launch plus the PC-sampling stall shares.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.json [idx]"""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_global_loads",
        "sass__inst_executed_global_stores", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__inst_executed_pipe_fp64.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2 + idx]
    d = {"kernel": vals[h.index("Kernel Name")]}
    stalls = {}
    for i, n in enumerate(h):
        if n in KEEP:
            d[n] = {"unit": units[i], "value": vals[i]}
        if "pcsamp_warps_issue_stalled" in n and "not_issued" not in n:
            try:
                stalls[n.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(vals[i])
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1.0
    d["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.002}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
