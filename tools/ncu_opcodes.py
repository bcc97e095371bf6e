"""Executed-instruction histogram by SASS opcode for one kernel of an ncu report (source page, --import-source not needed).
   python tools/ncu_opcodes.py <report.ncu-rep> <kernel regex> [top N]"""
import csv, io, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address' and 'Source' in r:
        if hdr: break
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
iS = hdr.index('Source'); iI = hdr.index('Instructions Executed')
agg = {}; tot = 0
for r in data:
    src = re.sub(r'^\s*@!?U?P\d+\s+', '', r[iS].strip())
    op = src.split()[0].rstrip(';') if src else '?'
    base = op.split('.')[0]
    key = op if base in ('IMAD', 'LDS', 'STS', 'LDG', 'STG', 'SYNCS', 'ATOMS', 'SHF', 'LEA') else base
    n = int(r[iI] or 0); agg[key] = agg.get(key, 0) + n; tot += n
print('kernel', kern, 'warp instructions', tot)
for k, n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print('%6.2f%%  %12d  %s' % (100.0 * n / tot, n, k))
