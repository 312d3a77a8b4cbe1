"""Print the key metrics + top stalled SASS lines of an .ncu-rep (read here, no GPU needed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print('-' * 60)
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print('  %-62s %-8s %s' % (w, units[i], r[i][:90]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", sys.argv[3]] if len(sys.argv) > 3 else []),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[hi]
    si = hdr.index('# Samples'); so = hdr.index('Source'); ie = hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    tot = sum(int(r[si]) for r in data)
    agg = {}
    for r in data:
        for i in stall:
            if r[i] not in ('', '0'):
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print('samples', tot, sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(data, key=lambda r: -int(r[si]))[:int(sys.argv[2])]:
        st = sorted(((hdr[i], int(r[i])) for i in stall if r[i] not in ('', '0')), key=lambda kv: -kv[1])[:2]
        print('%7s %9s  %-70s %s' % (r[si], r[ie], r[so].strip()[:70], st))
