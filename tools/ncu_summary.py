#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (read here, no GPU needed) into two small CSVs:
   <out>_ncu_full.csv      selected metrics of the first matching launch (raw page)
   <out>_sass_opcodes.csv  executed warp instructions and stall samples by SASS opcode (source page)
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_mh_kernel [kernel-substring]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio",
        "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    row = next(r for r in rows[2:] if pat in r[kcol])
    with open(out + "_ncu_full.csv", "w") as f:
        f.write("metric,unit,value\n")
        f.write("kernel,,\"%s\"\n" % row[kcol])
        for h, u, v in zip(hdr, units, row):
            if h in KEEP:
                f.write("%s,%s,%s\n" % (h, u, v))
    src = ncu(rep, "source")
    # the source page holds one table per launch; take the first that matches
    lines = src.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and pat in l)
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    tab = list(csv.reader(io.StringIO("\n".join(lines[start + 1:end]))))
    h = tab[0]
    ci, cs, cn = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    inst, stall = defaultdict(int), defaultdict(int)
    for r in tab[1:]:
        if len(r) <= max(ci, cs, cn):
            continue
        toks = r[ci].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        try:
            inst[op] += int(r[cn]); stall[op] += int(r[cs])
        except ValueError:
            pass
    tot = sum(inst.values()) or 1
    with open(out + "_sass_opcodes.csv", "w") as f:
        f.write("opcode,warp_instructions_executed,share,stall_samples\n")
        for op in sorted(inst, key=lambda o: -inst[o]):
            f.write("%s,%d,%.5f,%d\n" % (op, inst[op], inst[op] / tot, stall[op]))
    print("wrote", out + "_ncu_full.csv", out + "_sass_opcodes.csv")


if __name__ == "__main__":
    main()
