import sys,csv
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__t_bytes.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__inst_executed_pipe_alu.sum','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('---',r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr: print(f"  {w:70s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st=[(float(r[i].replace(',','')),h) for i,h in enumerate(hdr) if h.startswith('smsp__average_warp') and 'per_issue_active' in h and r[i] not in ('','n/a')]
    st=[(v,h) for v,h in st if 'not_issued' not in h]
    for v,h in sorted(st,reverse=True)[:10]: print(f"  STALL {v:8.2f} {h}")
