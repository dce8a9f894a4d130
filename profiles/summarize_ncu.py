"""Turns `ncu --set full` reports (gpurun_out/*.ncu-rep, scratch) into the small JSON summaries committed under profiles/.

    python profiles/summarize_ncu.py <report.ncu-rep> <out.json> "<what was run>" [kernel-substring]

Per captured launch: duration, DRAM bytes read / written, L1 and L2 hit rates, registers, achieved occupancy, issue
utilisation, active lanes per instruction, warp instructions, and the stall reasons per issued instruction that matter
for these kernels.  Numbers under a profiler are cold-cache and serialised: they explain a kernel, they are not bench
values (bench values come from bench.py, CUDA events)."""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "smsp__inst_executed_op_local_ld.sum": "local_loads", "smsp__inst_executed_op_local_st.sum": "local_stores",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
    pat = sys.argv[4] if len(sys.argv) > 4 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    launches = []
    for r in rows[2:]:
        if pat and pat not in r[ki]:
            continue
        d = {"kernel": r[ki].split("(")[0].replace("void ", "")}
        for k, name in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if name in ("dram_read", "dram_write"):
                    d[name + "_bytes"] = v * SCALE.get(u, 1.0)
                elif name == "duration":
                    d["duration_us"] = v * SCALE.get(u, 1.0) * 1e6
                else:
                    d[name] = v
        if "dram_read_bytes" in d and "duration_us" in d:
            d["dram_GBps"] = (d["dram_read_bytes"] + d["dram_write_bytes"]) / (d["duration_us"] * 1e-6) / 1e9
        launches.append(d)
    json.dump({"what": what, "tool": "ncu --set full --clock-control none (one launch each, under the profiler)", "launches": launches},
              open(out, "w"), indent=1)
    print("wrote", out, len(launches), "launches")


if __name__ == "__main__":
    main()
