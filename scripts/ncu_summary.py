"""Summarise an .ncu-rep (run in the build container): key raw metrics + SASS hot-spot buckets."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum','l1tex__throughput.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg','sm__cycles_active.avg','sm__cycles_elapsed.max','smsp__inst_executed.sum']
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:80])
    for k in KEYS:
        if k in hdr:
            print(f"  {k} = {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
    st = [(h, r[i]) for i, h in enumerate(hdr) if 'smsp__pcsamp_warps_issue_stalled' in h and not h.endswith('not_issued')]
    st.sort(key=lambda x: -float(x[1] or 0))
    print("  stall samples:", ", ".join(f"{h.split('stalled_')[1]}={v}" for h, v in st[:8]))
