"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports: the metrics quoted in profiles/r01_ncu_manytarg.txt (duration, DRAM bytes
and throughput, DMMA / FP64 pipe activity, issue rate, shared-memory conflicts, L2 read sectors) and every warp-stall reason above
0.2 per issued instruction, for the LAST kernel of each file.  Usage: python tools/summarize_ncu.py a_raw.csv [b_raw.csv ...]"""
import csv,sys
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','gpc__cycles_elapsed.avg.per_second','lts__t_sector_hit_rate.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ldgsts.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','dram__cycles_active.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_read.sum']
for f in sys.argv[1:]:
    rows=list(csv.reader(open(f)))
    hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    hdr=rows[hi]; units=rows[hi+1]; idx={h:i for i,h in enumerate(hdr)}
    r=rows[-1]
    print('====',f, r[idx['Kernel Name']][:40])
    for w in want:
        if w in idx: print('  %-90s %s %s'%(w, r[idx[w]], units[idx[w]]))
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            v=float(r[idx[h]])
            if v>0.2: print('  stall %-60s %.2f'%(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), v))
