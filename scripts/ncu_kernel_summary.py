"""Summarise one kernel of an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into the small JSON committed under
profiles/ (and read by bench.py for the issue-bound roofline).  Usage:
    python scripts/ncu_kernel_summary.py gpurun_out/r02_knn.ncu-rep k_sor_knn profiles/r02_knn_ncu.json n=10000000 kind=mixed hash=i32wrap
Runs here (no GPU needed to read a report)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_ns",
    "smsp__inst_executed.sum": "warp_instructions_per_launch",
    "smsp__thread_inst_executed.sum": "thread_instructions_per_launch",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed": "mem_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_pipe_wavefronts_shuffles",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum": "global_load_wavefronts",
    "sm__inst_executed_pipe_alu.sum": "pipe_alu",
    "sm__inst_executed_pipe_fma.sum": "pipe_fma",
    "sm__inst_executed_pipe_lsu.sum": "pipe_lsu",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio": "stall_long_scoreboard",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio": "stall_short_scoreboard",
    "smsp__average_warp_latency_issue_stalled_wait.ratio": "stall_wait",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio": "stall_not_selected",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio": "stall_math_throttle",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio": "stall_branch",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio": "stall_barrier",
}

UNIT_SCALE = {"ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0,
              "second": 1e9}


def main():
    rep, kernel, out = sys.argv[1:4]
    extra = dict(a.split("=", 1) for a in sys.argv[4:])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    kcol = hdr.index("Kernel Name")
    sel = [r for r in body if kernel in r[kcol]]
    if not sel:
        raise SystemExit(f"no launch of {kernel} in {rep}")
    r = sel[0]
    res = {"kernel": r[kcol], "source": rep.split("/")[-1], "launches_in_report": len(sel)}
    for k, v in extra.items():
        res[k] = int(v) if v.isdigit() else v
    for i, name in enumerate(hdr):
        if name in WANT:
            try:
                val = float(r[i].replace(",", ""))
            except ValueError:
                continue
            val *= UNIT_SCALE.get(units[i], 1.0) if ("bytes" in name or "duration" in name) else 1.0
            res[WANT[name]] = val
    if "dram_read_bytes" in res:
        res["dram_bytes_per_launch"] = res["dram_read_bytes"] + res.get("dram_write_bytes", 0.0)
    if "queries" not in res and "n" in res:
        res["queries"] = res["n"]
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
