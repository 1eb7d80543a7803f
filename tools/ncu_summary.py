import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units=rows[0],rows[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__block_size','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i=hdr.index(w); print(f'{w} = {r[i][:100]} {units[i]}')
    print('---')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address']
hdr=rows[hi[0]]; blk=rows[hi[0]+1:(hi[1]-1 if len(hi)>1 else len(rows))]
sc=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot={hdr[i]:0 for i in sc}; ops={}; nexec=0
iex=hdr.index('Instructions Executed')
for r in blk:
    if len(r)<len(hdr): continue
    try: ex=int(r[iex])
    except: ex=0
    nexec+=ex
    toks=r[1].split()
    op=(toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')).split('.')[0]
    ops[op]=ops.get(op,0)+ex
    for i in sc:
        try: tot[hdr[i]]+=int(r[i])
        except: pass
s=sum(tot.values()) or 1
print('warp-instr executed',nexec)
print(' '.join(f'{k[6:]}={100*v/s:.1f}%' for k,v in sorted(tot.items(),key=lambda kv:-kv[1])[:10]))
print(' '.join(f'{k}={100*v/nexec:.1f}%' for k,v in sorted(ops.items(),key=lambda kv:-kv[1])[:22]))
