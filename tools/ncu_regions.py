"""Summarise an ncu report's source page: executed warp instructions / active threads / stall samples per code region.
usage: ncu_regions.py report.ncu-rep [chunk_bytes]"""
import csv, subprocess, sys
rep = sys.argv[1]
chunk = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x200
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, isrc, iex, ith, ismp = (hdr.index(k) for k in ('Address', 'Source', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
base = int(data[0][ia], 16)
tot = sum(int(r[iex]) for r in data); tth = sum(int(r[ith]) for r in data); tsm = sum(int(r[ismp]) for r in data)
print('total warp inst', tot, 'thread inst', tth, 'avg threads %.2f' % (tth / tot), 'samples', tsm)
acc = {}
for r in data:
    k = (int(r[ia], 16) - base) // chunk
    a = acc.setdefault(k, [0, 0, 0])
    a[0] += int(r[iex]); a[1] += int(r[ith]); a[2] += int(r[ismp])
for k in sorted(acc):
    if acc[k][0] > tot * 0.002:
        print('%6s warp_inst %5.2f%%  avg thr %4.1f  samples %5.2f%%' % (hex(k * chunk), 100 * acc[k][0] / tot, acc[k][1] / max(1, acc[k][0]), 100 * acc[k][2] / tsm))
if len(sys.argv) > 3:   # dump instructions with counts
    for r in data:
        print('%6s %10s %5s %s' % (hex(int(r[ia], 16) - base), r[iex], r[ismp], r[isrc].strip()))
