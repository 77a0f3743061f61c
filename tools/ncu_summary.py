#!/usr/bin/env python3
"""Summarise an Nsight Compute report (``ncu --set full``) into a small CSV + markdown table for ``profiles/``.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_edge_full

Reads the report with ``ncu -i ... --page raw --csv`` (works without a GPU) and keeps the metrics the roofline
discussion in DESIGN.md uses: duration, DRAM bytes (read + write = `traffic`), issue / pipe utilisation, occupancy,
shared-memory wavefronts and bank conflicts, stall reasons.
"""
import csv
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conflicts"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma_cyc_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "pipe_tensor_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "not_selected", "selected", "dispatch_stall", "branch_resolving", "no_instruction", "sleeping", "membar"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}
    names = ["kernel"] + [s for _, s in KEEP] + ["stall_" + s for s in STALLS]
    table = []
    for r in body:
        rec = {"kernel": r[col["Kernel Name"]].replace("void ", "").split("(")[0]}
        for k, s in KEEP:
            rec[s] = (r[col[k]] + " " + units[col[k]]).strip() if k in col else ""
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            rec["stall_" + s] = r[col[k]] if k in col else ""
        table.append(rec)
    with open(out + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=names)
        w.writeheader()
        w.writerows(table)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n(stall_* = average warps stalled for that reason per issue-active cycle)\n\n")
        for rec in table:
            f.write(f"## {rec['kernel']}\n\n| metric | value |\n|---|---|\n")
            for n in names[1:]:
                if rec[n] != "":
                    f.write(f"| {n} | {rec[n]} |\n")
            f.write("\n")
    print("wrote", out + ".csv", out + ".md")


if __name__ == "__main__":
    main()
