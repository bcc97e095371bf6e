"""SASS listing of one kernel with executed warp-instruction counts, restricted to a source-line range (outermost call site).
   python tools/ncu_sass.py <report.ncu-rep> <kernel regex> <lo> <hi> [library.so]"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
so = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'u-vip-slam_b200', 'libuvip_orb.so')
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address' and 'Source' in r:
        if hdr: break
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
tmp = tempfile.mkdtemp(); subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
best = None
for cb in sorted(os.listdir(tmp)):
    if not cb.endswith('.cubin') or '-' in cb: continue
    asm = subprocess.run(['nvdisasm', '--print-line-info-inline', os.path.join(tmp, cb)], capture_output=True, text=True).stdout
    cur = None; line = ('', 0)
    for ln in asm.split('\n'):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            cur = [] if re.search(kern, m.group(1)) else None
            if cur is not None and (best is None or abs(len(best) - len(data)) > 0): best = cur
            continue
        if cur is None: continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            mi = re.findall(r'inlined at "(.*?)", line (\d+)', ln)
            line = (os.path.basename(mi[-1][0]), int(mi[-1][1])) if mi else (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', ln): cur.append(line)
tot = sum(int(r[iI] or 0) for r in data); sel = 0
for i, r in enumerate(data):
    fn, ln = best[i] if i < len(best) else ('', 0)
    if lo <= ln <= hi:
        n = int(r[iI] or 0); sel += n
        print('%5d L%-4d %10d %5.2f%% smp %5s  %s' % (i, ln, n, 100.0 * n / tot, r[iN], r[iS].strip()[:90]))
print('range total %d = %.2f%% of %d' % (sel, 100.0 * sel / tot, tot))
