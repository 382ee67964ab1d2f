import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; 
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','launch__grid_size','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    d=dict(zip(hdr,vals))
    print('==',d.get('Kernel Name','')[:60])
    for k in keys:
        if k in d: print(f"  {k:75s} {d[k]}")
    for h in hdr:
        if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            try:
                v=float(d[h]); 
                if v>0.15: print(f"  stall {h.split('issue_stalled_')[1].split('_per')[0]:30s} {v:.2f}")
            except: pass
