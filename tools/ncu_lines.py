"""Executed warp instructions per source line of an ncu report captured with --import-source on (kernels built with
-lineinfo).  usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif len(r) > 5 and r[0] == 'Line No':
        hdr = r
        iex, ith, ismp = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    elif hdr and len(r) > 5 and r[0] != '':
        num = lambda x: int(x) if x.strip('-') else 0
        lines.append((cur, int(r[0]), r[1].strip(), num(r[iex]), num(r[ith]), num(r[ismp])))
tot = sum(l[3] for l in lines); tsm = sum(l[5] for l in lines)
print('total warp inst', tot, 'samples', tsm)
byfile = {}
for f, n, s, ex, th, sm in lines:
    a = byfile.setdefault(f, [0, 0, 0]); a[0] += ex; a[1] += th; a[2] += sm
for f, a in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print('%-24s warp_inst %5.1f%%  avg thr %4.1f  samples %5.1f%%' % (f, 100 * a[0] / tot, a[1] / max(1, a[0]), 100 * a[2] / max(1, tsm)))
for f, n, s, ex, th, sm in sorted(lines, key=lambda l: -l[3])[:top]:
    print('%5.2f%% thr %4.1f smp %5.2f%%  %s:%d  %s' % (100 * ex / tot, th / max(1, ex), 100 * sm / max(1, tsm), f, n, s[:110]))
