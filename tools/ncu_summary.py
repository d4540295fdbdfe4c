#!/usr/bin/env python
"""Summarise an Nsight Compute report (one `ncu --set full` capture) as JSON: one object per kernel launch with the counters
DESIGN.md / profiles/README.md quote.  Runs where `ncu` is installed (no GPU needed to READ a report).
Usage: python tools/ncu_summary.py report.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
    "launch__cluster_size": "cluster",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active": "pipe_uniform_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar",
    "nvltx__bytes.sum": "nvlink_tx_bytes",
    "nvlrx__bytes.sum": "nvlink_rx_bytes",
    "lts__t_sectors_srcunit_tex_aperture_peer.sum": "l2_sectors_to_peer",
    "lts__t_sectors_srcunit_tex_aperture_peer_op_write.sum": "l2_sectors_to_peer_write",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for full, short in WANT.items():
            if full in hdr:
                i = hdr.index(full)
                try:
                    d[short] = float(r[i].replace(",", ""))
                    if units[i]:
                        d[short + "_unit"] = units[i]
                except ValueError:
                    pass
        res.append(d)
    txt = json.dumps(res, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt + "\n")
    else:
        print(txt)


if __name__ == "__main__":
    main()
