#!/usr/bin/env python
"""Condenses the round's `ncu --set full` capture into profiles/r2_counters.json: the per-unit counters bench.py scales by the work
it counts live (bench.py cannot run under ncu).  For each window-traversal launch: DRAM bytes per face, shared-memory load
wavefronts per node test and per LDS instruction, issue rate, share of warp samples stalled on the long scoreboard; for the
other kernels: DRAM bytes per launch and duration under ncu.

usage: ncu_counters.py <report.ncu-rep> <faces_per_launch> <bench.json with work_per_step at the same faces_per_launch> [git head of the captured build]
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main():
    rep, faces, bench = sys.argv[1], int(sys.argv[2]), json.loads(Path(sys.argv[3]).read_text().strip().splitlines()[-1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def get(r, name, as_bytes=False):
        if name not in ix:
            return None
        v = num(r[ix[name]])
        if v is None:
            return None
        return v * scale.get(units[ix[name]], 1) if as_bytes else v

    work = bench["work_per_step"]
    bench_faces = bench["config"]["faces_per_gpu"]
    out = {"git_head": sys.argv[4] if len(sys.argv) > 4 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip(),
           "report": Path(rep).name, "faces_per_launch": faces, "kernels": []}
    trav = 0
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        dram = (get(r, "dram__bytes_read.sum", True) or 0) + (get(r, "dram__bytes_write.sum", True) or 0)
        rec = {"kernel": name, "duration_ms_under_ncu": get(r, "gpu__time_duration.sum"), "dram_bytes_per_launch": dram,
               "issue_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "lts_throughput_pct": get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "l1tex_throughput_pct": get(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
               "registers": get(r, "launch__registers_per_thread"), "block": get(r, "launch__block_size"), "grid": get(r, "launch__grid_size")}
        if units[ix["gpu__time_duration.sum"]] in ("us", "usecond"):
            rec["duration_ms_under_ncu"] /= 1e3
        elif units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond"):
            rec["duration_ms_under_ncu"] /= 1e6
        out["kernels"].append(rec)
        if "k_traverse_win" in name:
            key = "k_traverse_win_hp" if trav == 0 else "k_traverse_win_ffd"    # launch order: head pose, then FFD
            tests = work["hp_node_tests" if trav == 0 else "ffd_node_tests"] * faces / bench_faces
            trav += 1
            wf = get(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum")
            # requests of the shared-memory loads = wavefronts - bank conflicts (one wavefront per conflict-free request)
            confl = get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum")
            lds = wf - confl if (wf and confl is not None) else None
            stall = None
            tot = sum(v for v in (get(r, h) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h) if v)
            lsb = get(r, "smsp__pcsamp_warps_issue_stalled_long_scoreboard")
            if tot and lsb is not None:
                stall = 100.0 * lsb / tot
            out[key] = {"kernel": name, "dram_bytes_per_face": dram / faces, "node_tests_per_launch": tests,
                        "lds_wavefronts_per_node_test": wf / tests if wf else None, "lds_wavefronts_per_load": wf / lds if (wf and lds) else None,
                        "lds_pipe_pct": get(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed"),
                        "tex_pipe_pct": get(r, "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                        "issue_active_pct": rec["issue_active_pct"], "long_scoreboard_stall_pct": stall, "duration_ms_under_ncu": rec["duration_ms_under_ncu"]}
    (ROOT / "profiles" / "r2_counters.json").write_text(json.dumps(out, indent=1))
    print(json.dumps({k: v for k, v in out.items() if k != "kernels"}, indent=1))


if __name__ == "__main__":
    main()
