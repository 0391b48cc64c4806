"""Turn an ncu report (gpurun_out/*.ncu-rep) into the text summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r1_validate_kernel.txt
"""
import csv
import io
import subprocess
import sys

RAW_KEYS = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
)


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(rep, out):
    lines = []
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    if len(raw) >= 3:
        hdr, units = raw[0], raw[1]
        for row in raw[2:]:
            name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            lines.append(f"== kernel: {name}")
            for h, u, v in zip(hdr, units, row):
                if h in RAW_KEYS:
                    lines.append(f"  {h} [{u}] = {v}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, hdr, agg = None, None, {}
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
            iB, iL = hdr.index("stall_barrier"), hdr.index("stall_long_sb")
        elif hdr and len(r) > iL and r[0] not in ("", "Function Name") and r[2] == "-":
            try:
                agg[(cur, int(r[0]))] = (float(r[iS] or 0), float(r[iI] or 0), float(r[iB] or 0), float(r[iL] or 0), r[1][:100])
            except ValueError:
                pass
    tot = sum(v[0] for v in agg.values()) or 1.0
    lines.append(f"== warp-stall samples by source line (total {tot:.0f}; top 30)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        lines.append(f"  {k[0]}:{k[1]:<4d} {100 * v[0] / tot:5.1f}%  barrier {v[2]:7.0f} long_sb {v[3]:7.0f} inst {v[1] / 1e6:8.2f}M | {v[4]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
