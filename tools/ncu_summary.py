"""Summarise .ncu-rep files (read here, no GPU): key per-launch metrics as a markdown table."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__warps_eligible.avg.per_cycle_active','sm__cycles_active.avg','sm__cycles_elapsed.avg','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
def table(rep):
    txt = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(txt.splitlines())); hdr = rows[0]
    ki = hdr.index('Kernel Name')
    names = [r[ki] for r in rows[2:]]
    out = ["kernels: " + "; ".join(f"[{i}] {n[:90]}" for i,n in enumerate(names)), "", "| metric | unit | " + " | ".join(f"[{i}]" for i in range(len(names))) + " |", "|---|---|" + "---|"*len(names)]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            vals = []
            for r in rows[2:]:
                try: vals.append(f"{float(r[i].replace(',','')):.4g}")
                except ValueError: vals.append(r[i])
            out.append(f"| {w.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio','')} | {rows[1][i]} | " + " | ".join(vals) + " |")
    return "\n".join(out)
if __name__ == "__main__":
    for rep in sys.argv[1:]:
        print(f"### {rep}\n"); print(table(rep)); print()
