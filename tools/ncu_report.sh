#!/bin/bash
# tools/ncu_report.sh <report.ncu-rep> "<title>" > profiles/<name>.txt : headline metrics, stall reasons, per-line and SASS summaries of one capture
rep=$1; title=$2
tmp=$(mktemp)
ncu -i $rep --page raw --csv 2>/dev/null > $tmp
echo "# $title"
echo "## headline metrics (tools/ncu_metrics.py)"
python tools/ncu_metrics.py < $tmp
python - $tmp <<'P'
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr,units,r=rows[0],rows[1],rows[2]
print("## stall reasons per issued instruction, pipes, memory")
keys=('sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active',
      'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
      'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
      'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
      'lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio',
      'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')
for i,h in enumerate(hdr):
    if 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h:
        try: v=float(r[i])
        except: continue
        if v>0.1: print(f"{h:<95}{v:8.2f}")
    if h in keys: print(f"{h:<95}{r[i]:>16} {units[i]}")
P
echo "## per CUDA line (tools/ncu_lines.py, >= 1.5 %)"
ncu -i $rep --page source --print-source cuda,sass --csv 2>/dev/null | python tools/ncu_lines.py 1.5
echo "## SASS opcode mix (tools/ncu_sass_summary.py)"
ncu -i $rep --page source --print-source sass --csv 2>/dev/null | python tools/ncu_sass_summary.py | head -40
rm -f $tmp
