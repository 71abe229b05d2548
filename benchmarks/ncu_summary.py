# usage: python benchmarks/ncu_summary.py <report.ncu-rep>   -- key metrics + top stall reasons per kernel from the raw page
import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
want=["Kernel Name","gpu__time_duration.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sectors_op_red.sum","lts__t_sectors_op_atom.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
stall=[h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=====")
    for w in want:
        if w in idx: print(f"{w:75s} {r[idx[w]]} {units[idx[w]]}")
    st=sorted(((float(r[idx[h]] or 0),h) for h in stall), reverse=True)[:8]
    for v,h in st: print(f"   stall {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:.2f}")
