import csv, sys, subprocess, io
rep = sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
want=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__cycles_active.avg','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('---')
    for w in want:
        if w in idx: print(f"  {w}: {r[idx[w]][:90]} {units[idx[w]]}")
