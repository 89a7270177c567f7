#!/usr/bin/env python
"""Per-kernel pipe utilisation and warp-stall mix of an .ncu-rep (ncu --set full), as kept under profiles/*_kernel_pipes_and_stalls.txt.
usage: ncu_pipes.py <report.ncu-rep> <out.txt> [note]"""
import csv
import subprocess
import sys

PIPES = [("time ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("issue %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         ("LSU wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
         ("shared-load wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
         ("of which bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
         ("TEX wavefronts %", "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed"),
         ("FMA pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         ("L1 hit %", "l1tex__t_sector_hit_rate.pct"), ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
         ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"), ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("threads per instr", "smsp__thread_inst_executed_per_inst_executed.ratio"),
         ("warp instr", "smsp__inst_executed.sum")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"Per-kernel pipe utilisation and warp-stall mix from {rep} (ncu --set full --clock-control none). {note}", ""]
    for r in rows[2:]:
        lines.append("== " + r[ix["Kernel Name"]][:110])
        parts = []
        for label, m in PIPES:
            if m in ix and r[ix[m]] not in ("", "n/a"):
                u = units[ix[m]]
                parts.append(f"{label} {r[ix[m]]}{(' ' + u) if u not in ('', '%') and 'byte' in u else ''}")
        lines.append("   " + "; ".join(parts))
        st = {}
        for h, i in ix.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        tot = sum(st.values())
        if tot > 0:
            lines.append("   stalls: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(rows) - 2, "kernels")


if __name__ == "__main__":
    main()
