"""Summarise an exported ncu report: python scripts/ncu_summary.py raw.csv src.csv  (from `ncu -i X.ncu-rep --page raw|source --csv`)."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.per_cycle_active', 'sm__cycles_active.avg', 'sm__cycles_active.max',
        'sm__cycles_active.min', 'sm__cycles_elapsed.avg', 'launch__registers_per_thread', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed')
for i, h in enumerate(hdr):
    if h in keep or ('issue_stalled' in h and 'per_issue_active' in h):
        print(h, '=', vals[i], units[i])
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(int(r[ix['# Samples']] or 0) for r in data)
    print('total samples', tot, 'SASS instructions', len(data))
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
        print(r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Instructions Executed']], 'long', r[ix['stall_long_sb']], 'short', r[ix['stall_short_sb']],
              'wait', r[ix['stall_wait']], 'math', r[ix['stall_math']], 'br', r[ix['stall_branch_resolving']], r[ix['Source']][:80])
    c = collections.Counter()
    tot = 0
    for r in data:
        n = int(r[ix['Instructions Executed']] or 0)
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
        op = m.group(2).split('.')[0] if m else r[ix['Source']][:10]
        c[op] += n
        tot += n
    print('total warp instructions', tot)
    for op, n in c.most_common(16):
        print(op, n, '%.1f%%' % (100 * n / tot))
