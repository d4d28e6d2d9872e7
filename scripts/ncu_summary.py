"""Summarise an .ncu-rep: per-launch key metrics (raw page) and the top stall instructions (source page)."""
import csv, subprocess, sys, io

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', r[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('   %-75s %s %s' % (w, r[i], units[i]))


def source(rep, skip, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(skip), '--launch-count', '1'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
    tot = sum(int(r[iS]) for r in data)
    agg = {}
    for r in data:
        for i in stall:
            if r[i] not in ('', '0'):
                agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
    print('total samples', tot, 'instructions', len(data), 'executed', sum(int(r[iEx]) for r in data))
    print('stall totals:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    for r in sorted(data, key=lambda r: -int(r[iS]))[:top]:
        st = sorted(((hdr[i][6:], int(r[i])) for i in stall if r[i] not in ('', '0')), key=lambda kv: -kv[1])[:3]
        print('%6d %5.1f%% ex=%9s %-58s %s' % (int(r[iS]), 100 * int(r[iS]) / max(tot, 1), r[iEx], r[iSrc].strip()[:58], st))


if __name__ == '__main__':
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        source(rep, int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    else:
        raw(rep)
