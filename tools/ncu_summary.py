"""Prints the metrics we care about from an .ncu-rep: python tools/ncu_summary.py file.ncu-rep [kernel substring]
(the first launch whose name contains the substring; default: the first launch)"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
if len(sys.argv) > 2:
    ki = h.index("Kernel Name")
    sel = [r for r in rows[2:] if len(r) > ki and sys.argv[2] in r[ki]]
    if not sel:
        sys.exit("no launch of " + sys.argv[2])
    rows = [rows[0], rows[1], sel[0]]
    print("kernel:", sel[0][ki][:100])
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__sectors_read.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'local_load_transactions', 'derived__l1tex__lsu_writeback_bytes_mem_lg.sum']
for w in want:
    if w in h:
        i = h.index(w)
        print(f"{w:72s} {rows[1][i]:>10s} {rows[2][i]}")
print("-- stall reasons (warps per issue active)")
for i, n in enumerate(h):
    if 'issue_stalled' in n and n.endswith('per_issue_active.ratio') and 'not_issued' not in n:
        try:
            v = float(rows[2][i])
        except Exception:
            continue
        if v > 0.15:
            print(f"{v:8.2f} {n}")
