#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into profiles/<name>.md: per kernel the
duration, DRAM bytes, DRAM / SM throughput %, occupancy, registers, stall picture.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_x.md"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM rate"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle")]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, lines = {}, []
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if seen.get(name, 0) >= int(__import__('os').environ.get('NCU_PER_KERNEL', '1')):
            continue
        seen[name] = seen.get(name, 0) + 1
        lines.append(f"### {name}\n")
        lines.append("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in ix and r[ix[k]] != "":
                lines.append(f"| {label} | {r[ix[k]]} {units[ix[k]]} |")
        lines.append("")
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary of `{rep}` (first launch of each kernel; `--clock-control none`)\n\n")
        fh.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
