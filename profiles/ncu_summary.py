"""Summarise an ncu report (raw page metrics + top stall instructions) — used to write profiles/*.md."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'smsp__inst_executed.sum',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__waves_per_multiprocessor', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'local_load_bytes', 'smsp__inst_executed_op_local_ld.sum']


def main(rep, top=18):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr = r[0]
    print('kernel:', r[2][hdr.index('Kernel Name')][:100] if 'Kernel Name' in hdr else '')
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print('%-62s %-14s %s' % (w, r[1][i], [x[i] for x in r[2:]]))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(src.splitlines()))
    hi = next(i for i, row in enumerate(r) if row and row[0] == 'Address')
    hdr, rows = r[hi], r[hi + 1:]
    # only the first kernel instance's rows (until the next 'Kernel Name' line)
    cut = next((i for i, row in enumerate(rows) if row and row[0] == 'Kernel Name'), len(rows))
    rows = rows[:cut]
    si = hdr.index('Warp Stall Sampling (All Samples)')
    tot = sum(int(x[si]) for x in rows if len(x) > si and x[si].isdigit())
    print('instructions', len(rows), 'stall samples', tot)
    cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {c: sum(int(x[hdr.index(c)]) for x in rows if len(x) > hdr.index(c) and x[hdr.index(c)].isdigit()) for c in cols}
    s = sum(agg.values()) or 1
    print('stalls: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * v / s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    t = sorted([(int(x[si]), x[1].strip()[:80], i) for i, x in enumerate(rows) if len(x) > si and x[si].isdigit()], reverse=True)[:top]
    for smp, ins, i in t:
        print('%6d %5.1f%% @%5d %s' % (smp, 100 * smp / max(tot, 1), i, ins))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 18)
