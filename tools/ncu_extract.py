"""Summarise a `ncu --set full` report as the per-kernel JSON kept under profiles/ (run where ncu is installed).

usage: python tools/ncu_extract.py gpurun_out/r1i_full.ncu-rep > profiles/r1i_ncu_metrics.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "time": "gpu__time_duration.sum",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "sm__inst_executed.sum.pct_of_peak_sustained_elapsed" ,
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lsu_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smem_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_math_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
}


def main(path):
    # a .csv is the raw page already exported on the GPU box (`ncu -i rep --page raw --csv`): reports with many launches
    # are too large to bring back
    raw = open(path).read() if path.endswith(".csv") else subprocess.run(
        ["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("nlb::", "")
        d = {}
        for key, metric in WANT.items():
            hits = [i for i, h in enumerate(hdr) if h == metric]
            if hits:
                d[key] = (r[hits[0]] + " " + units[hits[0]]).strip()
        out[name] = d
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1])
