"""Stall samples and executed instructions bucketed by SASS index range / source line: where does the kernel's time go.
usage: python profiles/ncu_buckets.py report.ncu-rep [bucket_size]"""
import csv, subprocess, sys, signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
rep = sys.argv[1]
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 250
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(src.splitlines()))
hi = next(i for i, row in enumerate(r) if row and row[0] == 'Address')
hdr, rows = r[hi], r[hi + 1:]
cut = next((i for i, row in enumerate(rows) if row and row[0] == 'Kernel Name'), len(rows))
rows = rows[:cut]
si = hdr.index('Warp Stall Sampling (All Samples)')
ei = hdr.index('Instructions Executed') if 'Instructions Executed' in hdr else None
ti = hdr.index('Thread Instructions Executed') if 'Thread Instructions Executed' in hdr else None
tot = sum(int(x[si]) for x in rows if x[si].isdigit())
tote = sum(int(x[ei]) for x in rows if ei is not None and x[ei].isdigit())
print('total samples', tot, 'warp instr executed', tote)
for b in range(0, len(rows), bs):
    chunk = rows[b:b + bs]
    s = sum(int(x[si]) for x in chunk if x[si].isdigit())
    e = sum(int(x[ei]) for x in chunk if ei is not None and x[ei].isdigit())
    t = sum(int(x[ti]) for x in chunk if ti is not None and x[ti].isdigit())
    print('%5d-%5d  samples %5.1f%%  instr %5.1f%%  lanes/instr %4.1f   first: %s' % (b, b + len(chunk), 100 * s / max(tot, 1), 100 * e / max(tote, 1), t / max(e, 1), chunk[0][1].strip()[:50]))
