#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md cites.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.txt] [--sass sass.txt]"""
import csv, subprocess, sys, io

KEEP = ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__cycles_active.avg','sm__cycles_elapsed.max','smsp__sass_average_branch_targets_threads_uniform.pct','launch__grid_size','launch__block_size',
 'launch__shared_mem_per_block_dynamic','smsp__inst_executed_op_shared_atom.sum','sm__inst_issued.avg.per_cycle_active',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','sm__cycles_active.avg','gpc__cycles_elapsed.avg.per_second']

def raw(rep):
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]

def sass(rep):
    out = subprocess.run(['ncu','-i',rep,'--page','source','--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]; ix = {h:i for i,h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        if len(r) < 10: continue
        res.append((r[ix['Source']].strip(), int(r[ix['Instructions Executed']]), int(r[ix['Thread Instructions Executed']]), int(r[ix['# Samples']])))
    return res

if __name__ == '__main__':
    rep = sys.argv[1]
    hdr, units, vals = raw(rep)
    lines = [f"{h} [{u}] = {v}" for h,u,v in zip(hdr,units,vals) if h in KEEP]
    txt = "\n".join(lines)
    if len(sys.argv) > 2 and not sys.argv[2].startswith('--'):
        open(sys.argv[2],'w').write(f"# ncu --set full summary of {rep}\n" + txt + "\n")
    print(txt)
    if '--sass' in sys.argv:
        path = sys.argv[sys.argv.index('--sass')+1]
        res = sass(rep)
        tot = sum(r[1] for r in res)
        with open(path,'w') as f:
            f.write(f"# idx  warp-instr-executed  avg-active-threads  pc-samples  SASS   (total warp instr {tot})\n")
            for i,(s,n,t,smp) in enumerate(res):
                f.write(f"{i:4d} {n:12d} {t/max(n,1):5.1f} {smp:6d}  {s}\n")
