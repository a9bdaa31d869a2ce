"""Key metrics per kernel from an ncu report: python scripts/ncu_raw_summary.py <file.ncu-rep>"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:70], 'grid', r[idx['Grid Size']] if 'Grid Size' in idx else '', 'block', r[idx['Block Size']] if 'Block Size' in idx else '')
    for w in want:
        if w in idx:
            print(f"  {w:66s} {r[idx[w]][:30]:>16s} {units[idx[w]]}")
