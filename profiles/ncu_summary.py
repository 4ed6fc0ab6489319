import csv, sys, subprocess, collections
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); H=rows[0]; r=rows[2]
for w in ['gpu__time_duration.sum','smsp__inst_executed.sum','launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']:
    if w in H: print(f"{w:90s} {r[H.index(w)]}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
blocks=[i for i,r in enumerate(rows) if r and r[0]=='Line No']
H=rows[blocks[0]]; end = blocks[1]-2 if len(blocks)>1 else len(rows)
li=H.index('Line No'); ie=H.index('Instructions Executed'); ss=H.index('# Samples')
per={}; srcs={}
for r in rows[blocks[0]+1:end]:
    if len(r)<=ie: continue
    try: ln=int(r[li]); n=int(r[ie] or 0); s=int(r[ss] or 0)
    except: continue
    a=per.setdefault(ln,[0,0]); a[0]+=n; a[1]+=s; srcs[ln]=r[1]
tot=sum(v[0] for v in per.values()); tots=sum(v[1] for v in per.values())
print('total inst',tot,'samples',tots)
key = (lambda kv:-kv[1][1]) if len(sys.argv)>2 and sys.argv[2]=='samp' else (lambda kv:-kv[1][0])
for ln,(n,s) in sorted(per.items(), key=key)[:int(sys.argv[3]) if len(sys.argv)>3 else 30]:
    print(f"{ln:5d} {n:10d} {100*n/tot:5.1f}%  samp {100*s/max(1,tots):5.1f}%  {srcs[ln][:115]}")
