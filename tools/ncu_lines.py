"""Per-source-line instruction / stall-sample summary of one kernel from an ncu report.
ncu's CSV source page is SASS-level; line numbers come from nvdisasm --print-line-info on the cubin embedded in the
library (the i-th SASS instruction of the function in both listings).
   python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N] [library.so]"""
import csv, io, os, re, subprocess, sys, tempfile

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        'u-vip-slam_b200', 'libuvip_orb.so')
nth_args = ['--launch-skip', os.environ['NTH'], '--launch-count', '1'] if os.environ.get('NTH') else []       # NTH = which matching launch
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern] + nth_args,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; nth = 0; seen = -1
for r in rows:
    if r and r[0] == 'Address' and 'Source' in r:
        seen += 1
        if seen > nth:
            break
        hdr = r; data = []; continue
    if hdr and len(r) == len(hdr):
        data.append(r)
iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iT = hdr.index('Thread Instructions Executed'); iN = hdr.index('# Samples')
# line info
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
lines = []
for cb in sorted(os.listdir(tmp)):
    if not cb.endswith('.cubin') or '-' in cb:
        continue
    asm = subprocess.run(['nvdisasm', '--print-line-info-inline', os.path.join(tmp, cb)], capture_output=True, text=True).stdout
    cur = None; line = ('', 0); fn = None
    for ln in asm.split('\n'):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            fn = m.group(1); cur = [] if re.search(kern, fn) else None
            if cur is not None:
                lines.append((fn, cur))
            continue
        if cur is None:
            continue
        m = re.search(r'//## File "(.*?)", line (\d+)', ln)
        if m:
            # an inlined intrinsic / helper is charged to the outermost call site ("... inlined at <file>, line N")
            mi = re.findall(r'inlined at "(.*?)", line (\d+)', ln)
            line = (os.path.basename(mi[-1][0]), int(mi[-1][1])) if mi else (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+\S', ln):
            cur.append(line)
best = None
for fn, l in lines:
    if len(l) == len(data):
        best = l
if best is None and lines:
    # nvdisasm lists a few trailing padding instructions that ncu omits: align from the front
    fn, l = min(lines, key=lambda t: abs(len(t[1]) - len(data)))
    print('note: listing lengths differ (%d vs %d), aligned from the front' % (len(l), len(data)))
    best = l[:len(data)] + [l[-1]] * max(0, len(data) - len(l))
if best is None:
    print('could not align SASS listings:', [(fn, len(l)) for fn, l in lines], len(data)); sys.exit(1)
srcs = {}
def src_line(fn, ln):
    if fn not in srcs:
        try:
            srcs[fn] = open(os.path.join(os.path.dirname(so), 'csrc', fn)).read().split('\n')
        except Exception:
            srcs[fn] = None
    t = srcs[fn]
    return t[ln - 1].strip()[:105] if t and 0 < ln <= len(t) else '<' + fn + '>'
agg = {}
for r, ln in zip(data, best):
    a = agg.setdefault(ln, [0, 0, 0])
    a[0] += int(r[iI] or 0); a[1] += int(r[iT] or 0); a[2] += int(r[iN] or 0)
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print('kernel', kern, 'SASS instructions', len(data), 'warp-instr executed', tot, 'samples', tots)
key = (lambda kv: -kv[1][2]) if os.environ.get('BY_SAMPLES') else (lambda kv: -kv[1][0])
for (fn, ln), a in sorted(agg.items(), key=key)[:top]:
    print('%5.1f%% inst %5.1f%% smp thr/warp=%4.1f  L%-4d %s' % (a[0] * 100 / max(tot, 1), a[2] * 100 / max(tots, 1), a[1] / max(a[0], 1), ln, src_line(fn, ln)))

if os.environ.get('RANGES'):
    # RANGES="name:lo-hi,name:lo-hi" sums instruction share per source-line range of extractor.cu / matcher.cu
    for spec in os.environ['RANGES'].split(','):
        name, rng = spec.split(':'); lo, hi = [int(v) for v in rng.split('-')]
        ins = sum(a[0] for (fn, ln), a in agg.items() if fn.endswith('.cu') and lo <= ln <= hi)
        smp = sum(a[2] for (fn, ln), a in agg.items() if fn.endswith('.cu') and lo <= ln <= hi)
        print('%-12s %5.1f%% inst %5.1f%% samples' % (name, ins * 100 / max(tot, 1), smp * 100 / max(tots, 1)))
    ins = sum(a[0] for (fn, ln), a in agg.items() if not fn.endswith('.cu'))
    print('%-12s %5.1f%% inst (CUDA headers: intrinsics, atomics)' % ('headers', ins * 100 / max(tot, 1)))
