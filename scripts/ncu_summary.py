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


def to_json(rep, out_path, kernel, source_note):
    """Average DRAM traffic / duration / tensor-pipe activity of the launches whose name contains `kernel`, as the JSON
    bench.py reads its roofline.traffic from."""
    import json
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    sel = [r for r in rows[2:] if kernel in r[hdr.index('Kernel Name')]]

    def col(name):
        i = hdr.index(name)
        scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9}.get(units[i], 1.0)
        return [float(r[i].replace(',', '')) * scale for r in sel]
    rd, wr, dur = col('dram__bytes_read.sum'), col('dram__bytes_write.sum'), col('gpu__time_duration.sum')
    tp = col('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
    d = {'kernel': kernel, 'launches': len(sel), 'dram_bytes_per_launch': sum(a + b for a, b in zip(rd, wr)) / len(sel),
         'dram_bytes_read': rd, 'dram_bytes_write': wr, 'duration_s': dur, 'tensor_pipe_active_pct': tp,
         'kernel_names': [r[hdr.index('Kernel Name')][:80] for r in sel], 'source': source_note}
    with open(out_path, 'w') as f:
        json.dump(d, f, indent=1)
    print('wrote', out_path, 'launches', len(sel), 'avg DRAM bytes', d['dram_bytes_per_launch'])


if __name__ == '__main__':
    rep = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == '--json':
        to_json(rep, sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else rep)
    elif len(sys.argv) > 2:
        source(rep, int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    else:
        raw(rep)
