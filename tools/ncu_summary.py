#!/usr/bin/env python
"""Summarise an ncu report (read here, on the CPU box) into profiles/*.md.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_x.md "command that was profiled"
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global store sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global store requests"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{rep}`", "", f"command: `{cmd}`", "", "`ncu --set full --clock-control none --import-source on` (cold cache, serialised replays: compare shares, not absolutes)", ""]
    for r in rows[2:]:
        lines += [f"## {r[idx['Kernel Name']]}", "", "| metric | value |", "|---|---|"]
        for k, label in KEYS:
            if k in idx:
                lines.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        st = [(h[33:], float(r[idx[h]].replace(",", ""))) for h in hdr
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[idx[h]] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1.0
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for n, v in sorted(st, key=lambda x: -x[1])[:6])
        lines += [f"| top warp stall reasons (pc sampling) | {top} |", ""]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
