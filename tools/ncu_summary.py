"""Summarise .ncu-rep captures (read on the CPU box with `ncu -i`): python tools/ncu_summary.py out.md rep1 rep2 ..."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
]


def main(out, reps):
    lines = ["# ncu captures (`ncu --set full --clock-control none --import-source on`, one launch each)\n"]
    traffic = {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for k_idx, vals in enumerate(r for r in rows[2:] if len(r) == len(hdr)):
            name = vals[col["Kernel Name"]]
            tag = rep.split('/')[-1] + (f" (launch {k_idx + 1})" if k_idx else "")
            lines.append(f"\n## {tag}\n\n`{name}`\n\n| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in col:
                    lines.append(f"| {label} (`{k}`) | {vals[col[k]]} {units[col[k]]} |")
            try:
                rd = float(vals[col["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(vals[col["dram__bytes_write.sum"]].replace(",", ""))
                scale_r = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[col["dram__bytes_read.sum"]]]
                scale_w = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[col["dram__bytes_write.sum"]]]
                if k_idx == 0:
                    traffic[rep.split("/")[-1]] = rd * scale_r + wr * scale_w
                lines.append(f"| **DRAM traffic per launch** | {(rd * scale_r + wr * scale_w) / 1e9:.3f} GB |")
            except Exception:
                pass
    open(out, "w").write("\n".join(lines) + "\n")
    print(traffic)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
