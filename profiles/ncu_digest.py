#!/usr/bin/env python
"""Digest of an ncu report: headline metrics per kernel (raw page) and the hottest source lines (source page)."""
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'launch__grid_size', 'launch__block_size', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'sm__cycles_elapsed.avg',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, top=25):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for i, h in enumerate(hdr):
            if h in KEEP:
                print(f"  {h} = {r[i]} {units[i]}")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur_file, cur_fn, hdr2 = None, None, None
    lines = {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0] == "Line No":
            hdr2 = r
        elif hdr2 and r[0].strip().isdigit():
            ci = hdr2.index("# Samples")
            ii = hdr2.index("Instructions Executed")
            ti = hdr2.index("Thread Instructions Executed")
            key = (cur_fn[:60], cur_file, int(r[0]), r[1][:110])
            try:
                v = (float(r[ci] or 0), float(r[ii] or 0), float(r[ti] or 0))
            except (ValueError, IndexError):
                continue  # a source line with unescaped quotes breaks the CSV row
            a = lines.setdefault(key, [0.0, 0.0, 0.0])
            a[0] += v[0]
            a[1] += v[1]
            a[2] += v[2]
    fns = sorted({k[0] for k in lines})
    for fn in fns:
        sel = {k: v for k, v in lines.items() if k[0] == fn}
        tot = sum(v[0] for v in sel.values()) or 1
        toti = sum(v[1] for v in sel.values()) or 1
        print(f"-- {fn}: hottest source lines by stall samples (total {tot:.0f} samples, {toti:.3g} warp instructions)")
        for k, v in sorted(sel.items(), key=lambda x: -x[1][0])[:top]:
            print(f"  {100 * v[0] / tot:5.1f}% smp {100 * v[1] / toti:5.1f}% inst lanes {v[2] / max(v[1], 1):4.1f}  {k[1]}:{k[2]:<4d} {k[3]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
