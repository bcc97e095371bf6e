"""Regenerate profiles/kernel_pipes.json and profiles/roofline_traffic.json from an ncu launch list (CSV, one step at batch 256):
   python tools/pipes_from_launches.py profiles/r1_launches_bench_b256_v2.csv
The launch list is the output of
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,\
smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,\
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,\
sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 45 -c 15 --csv python bench.py --steps 1 --warmup 3 --no-extras"""
import csv, hashlib, json, os, sys

src = sys.argv[1]
rows = list(csv.reader(open(src)))
hdr = [r for r in rows if 'Kernel Name' in r][0]
i0 = rows.index(hdr)
ki, mi, vi, idi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
launches = {}
for r in rows[i0 + 1:]:
    if len(r) > vi:
        launches.setdefault(int(r[idi]), {'name': r[ki]})[r[mi]] = float(r[vi].replace(',', ''))
order = sorted(launches)
# one step = import, 7 x resize, fast, quadtree, blur, select, describe, knn2: take the LAST complete step of the list
names = [launches[i]['name'] for i in order]
last_desc = max(i for i, n in enumerate(names) if 'k_describe' in n)
start = max(i for i, n in enumerate(names) if 'k_import' in n and i < last_desc)        # the last COMPLETE step of the window
nxt = [i for i, n in enumerate(names) if 'k_import' in n and i > start]
step = [launches[order[i]] for i in range(start, nxt[0] if nxt else len(order))]
# the kNN of a step follows the describe of the same step; if the list ends before it, take the one in front of the import
knn = [l for l in step if 'k_knn2' in l['name']] or [launches[order[i]] for i in range(start) if 'k_knn2' in names[i]][-1:]
def entry(l, extra=None):
    e = {'issue_slot_pct': l.get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
         'alu_pipe_pct': l.get('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
         'fma_pipe_pct': l.get('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
         'xu_pipe_pct': l.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
         'warps_active_pct': l.get('sm__warps_active.avg.pct_of_peak_sustained_active'),
         'warp_instructions': l.get('smsp__inst_executed.sum'),
         'ncu_time_us': l.get('gpu__time_duration.sum', 0.0) / 1000.0}
    if extra:
        e.update(extra)
    return e
traffic = lambda l: l.get('dram__bytes_read.sum', 0.0) + l.get('dram__bytes_write.sum', 0.0)
pipes = {'_note': 'per-kernel pipe utilisation from the committed ncu launch list %s (batch 256, one step; pct of peak sustained while active)' % src}
traf = {'_note': 'dram__bytes_read.sum + dram__bytes_write.sum per launch at batch 256 (ncu launch list %s, cold cache); pyramid = import + 7 resizes' % src}
rs = 0
pyr = 0.0
for l in step:
    n = l['name']
    if 'k_import' in n: pipes['import'] = entry(l); pyr += traffic(l)
    elif 'k_resize' in n: rs += 1; pipes['resize_l%d' % rs] = entry(l); pyr += traffic(l)
    elif 'k_fast' in n and 'k_fast_roi' not in n: pipes['fast'] = entry(l); traf['fast'] = traffic(l)
    elif 'k_quadtree' in n: pipes['quadtree'] = entry(l); traf['quadtree'] = traffic(l)
    elif 'k_blur' in n: pipes['blur'] = entry(l); traf['blur'] = traffic(l)
    elif 'k_select' in n: pipes['select'] = entry(l); traf['select'] = traffic(l)
    elif 'k_describe' in n: pipes['describe'] = entry(l); traf['describe'] = traffic(l)
traf['pyramid'] = pyr
if knn:
    pipes['knn2'] = entry(knn[0], {'xu_pipe_pct_popc': knn[0].get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')})
    traf['knn2'] = traffic(knn[0])
total = sum(v['ncu_time_us'] for k, v in pipes.items() if not k.startswith('_'))
for k, v in pipes.items():
    if not k.startswith('_'):
        v['share_of_step_pct'] = round(100.0 * v['ncu_time_us'] / total, 2)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the kernel sources this evidence was captured from (bench.py quotes it only while they are unchanged)
STEP_UNRELATED = ('klt.cu', 'synth.cu')       # same rule as bench.py kernel_source_hash()
h = hashlib.sha256()
d = os.path.join(root, 'u-vip-slam_b200', 'csrc')
for fn in sorted(os.listdir(d)):
    if fn.endswith(('.cu', '.cuh', '.inc')) and fn not in STEP_UNRELATED:
        h.update(fn.encode()); h.update(open(os.path.join(d, fn), 'rb').read())
pipes['_source_hash'] = traf['_source_hash'] = h.hexdigest()[:16]
json.dump(pipes, open(os.path.join(root, 'profiles', 'kernel_pipes.json'), 'w'), indent=1)
json.dump(traf, open(os.path.join(root, 'profiles', 'roofline_traffic.json'), 'w'), indent=1)
print(json.dumps({k: (v['ncu_time_us'], v['share_of_step_pct']) for k, v in pipes.items() if not k.startswith('_')}))
