#!/usr/bin/env python
"""Digest of one ncu report: headline counters, opcode mix and stall reasons of the first kernel in it.
Usage: python profiles/ncu_digest.py gpurun_out/prof.ncu-rep"""
import csv, sys, subprocess, re
from collections import Counter
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__cycles_elapsed.max','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_lsu.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum','sm__inst_executed_pipe_uniform.sum']
for k in keys:
    for i,h in enumerate(hdr):
        if h==k: print('%-75s %-10s %s'%(k,units[i],vals[i]))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}; data=rows[2:]
tot=sum(float(r[ix['Instructions Executed']] or 0) for r in data)
print('total warp instr',tot)
c=Counter(); 
for r in data:
    t=r[ix['Source']].split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    c[op.split('.')[0]]+=float(r[ix['Instructions Executed']] or 0)
print(' '.join('%s %.1f%%'%(o,100*v/tot) for o,v in c.most_common(18)))
st={k:sum(float(r[ix[k]] or 0) for r in data) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k}
ts=sum(st.values())
print(' '.join('%s %.1f%%'%(k[6:],100*v/ts) for k,v in sorted(st.items(),key=lambda x:-x[1])[:10]))
