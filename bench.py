#!/usr/bin/env python
"""bench.py — ORB front-end throughput on B200 (BASELINE.json metric), one JSON line on stdout.

A step = one pass of the hot path over one batch of synthetic frames: ORBextractor over `--batch` frames of
752x480 at 1000 keypoints (8 levels, 1.2, FAST 20/7 — the shape the metric is quoted on) followed by brute-force
Hamming kNN2 between consecutive frames' descriptors.  `value` is measured with inputs resident in HBM through the
device-pointer C-ABI; `e2e` goes through the host-buffer C-ABI calls (pinned host memory, H2D + D2H inside the timed
region).  N > 1: one process per GPU (torchrun), frames sharded, no data-path collective (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'ORB frames/s @752x480,1000kp; Hamming pairs/s; % of HBM/popc roofline'
W, H, NFEAT, NLEVELS, SCALE, FAST_TH = 752, 480, 1000, 8, 1.2, 20
B_FRAME_BYTES = 1177367          # 360 960 in + 756 407 pyramid out + 1000 x 60 B of keypoints/descriptors
FALLBACK_HBM_GBS = 6650.0        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='frames per step and per GPU')
    ap.add_argument('--ref-frames', type=int, default=0, help='frames per step of the reference arm (0 = --batch, the same step as our arm)')
    ap.add_argument('--e2e-chunk', type=int, default=0, help='frames per pipeline chunk of the host-buffer entry point (0 = the batch)')
    ap.add_argument('--no-extras', action='store_true', help='skip the Hamming-only and cpu_baseline legs')
    ap.add_argument('--seq-frames', type=int, default=4096, help='frames per sequence of --workload cfg5')
    ap.add_argument('--workload', default='frames', choices=['frames', 'knn', 'search', 'next', 'cfg5'],
                    help="'frames' = the headline metric; 'knn' = BASELINE config 4: database-sharded brute-force kNN2 with NCCL top-2 merge; 'search' = BASELINE config 3: batched grid-windowed SearchByProjection; "
                         "'cfg5' = BASELINE config 5 as written: one 4096-frame 1280x1024 sequence per GPU, streamed")
    ap.add_argument('--shape', default='euroc', choices=['euroc', 'aqualoc', 'hd'],
                    help='frame shape of --workload frames: euroc 752x480/1000 kp (the metric, config 1 shape), aqualoc 640x512/1500 kp (config 2), '
                         'hd 1280x1024/2000 kp (config 5)')
    ap.add_argument('--knn-n', type=int, default=262144, help='rows of the query and train sets for --workload knn (config 4 is 1048576)')
    return ap.parse_args()


SHAPES = {'euroc': (752, 480, 1000), 'aqualoc': (640, 512, 1500), 'hd': (1280, 1024, 2000)}


def set_shape(name):
    """select the frame shape; algorithmic bytes per frame = input + pyramid levels 1..7 written once + 60 B per keypoint"""
    global W, H, NFEAT, B_FRAME_BYTES
    import numpy as np
    W, H, NFEAT = SHAPES[name]
    inv, isf, tot = np.float32(1), np.float32(1.0 / np.float64(np.float32(SCALE))), 0
    for _ in range(1, NLEVELS):
        inv = np.float32(inv * isf)
        tot += int(np.rint(np.float32(W) * inv)) * int(np.rint(np.float32(H) * inv))
    B_FRAME_BYTES = W * H + tot + NFEAT * 60
    return 'ORBextractor %dx%d / %d kp / 8 levels / 1.2 / FAST 20-7' % (W, H, NFEAT)


def make_frames(synth, n, seed0):
    """n distinct synthetic EuRoC-shaped frames; generated in a few worker processes (numpy, ~50 ms each)"""
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        frames = list(pool.map(lambda i: synth.synth_frame(seed0 + i, W, H), range(n)))
    return np.stack(frames)


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process
    every few milliseconds (the timed region of a default run is well under a second, too short for `nvidia-smi -lms`
    to deliver a sample); nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index, period_s=0.004):
        self.index, self.period, self.proc, self.lines = index, period_s, None, []
        self.sm, self.mx, self.reasons, self.stop_flag, self.th, self.nvml = [], None, set(), False, None, None

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            try:
                return int(vis.split(',')[self.index])
            except Exception:
                pass
        return self.index

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            self.h = N.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.nvml = N
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self._physical_index()), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        N = self.nvml
        names = (('hw_slowdown', N.nvmlClocksEventReasonHwSlowdown), ('hw_thermal_slowdown', N.nvmlClocksEventReasonHwThermalSlowdown),
                 ('sw_thermal_slowdown', N.nvmlClocksEventReasonSwThermalSlowdown), ('sw_power_cap', N.nvmlClocksEventReasonSwPowerCap))
        while not self.stop_flag:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                r = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in names:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=2)
            sm = sorted(self.sm)
            return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.mx, 'reasons': sorted(self.reasons),
                    'samples': len(sm), 'source': 'nvml'}
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml and nvidia-smi unavailable'], 'samples': 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(',')]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi'}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'measured'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback'


# sources that hold no kernel of the benchmark step (KLT / RANSAC row N1, the synthetic-frame generator): changing them does not
# invalidate the ncu evidence of the step
STEP_UNRELATED = ('klt.cu', 'synth.cu')


def kernel_source_hash():
    """sha256 over the sources of the step's kernels: the committed ncu evidence names the sources it was captured from (tools/pipes_from_launches.py
    writes the same hash), so a bench run on changed kernels reports that evidence as stale instead of quoting it"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'u-vip-slam_b200', 'csrc')
    for fn in sorted(os.listdir(d)):
        if fn.endswith(('.cu', '.cuh', '.inc')) and fn not in STEP_UNRELATED:
            h.update(fn.encode()); h.update(open(os.path.join(d, fn), 'rb').read())
    return h.hexdigest()[:16]


def _static_evidence(name):
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', name)))
    except Exception:
        return {}, True
    return d, d.get('_source_hash') != kernel_source_hash()


def ncu_pipes():
    """per-kernel issue-slot / ALU-pipe utilisation and warp-instruction counts from the committed ncu launch list (static evidence:
    the instruction count of a kernel is a property of the code and the input, the utilisation figures are not measured live);
    returns (dict, stale)"""
    return _static_evidence('kernel_pipes.json')


def ncu_traffic():
    """dram bytes per launch of every stage's kernel from the committed ncu launch list; returns (dict, stale)"""
    return _static_evidence('roofline_traffic.json')


def _cv2_version():
    try:
        import cv2
        return cv2.__version__
    except Exception:
        return 'absent'


def cv2_composite(frames, threads):
    """SURVEY 8(d) cross-check (iii): the OpenCV primitives the reference calls, through the cv2 wheel of this image (whole-level
    FAST instead of per-cell calls, no quadtree / orientation / descriptors): pyramid (7 resize + 8 copyMakeBorder), FAST th 20
    with NMS and GaussianBlur 7x7 sigma 2 on all 8 levels, then BFMatcher.knnMatch(k=2) on 1000 x 1000 descriptors per frame
    pair.  Returns frames/s, or None when cv2 is missing."""
    try:
        import cv2
        import numpy as np
    except Exception:
        return None
    cv2.setNumThreads(threads)
    fast = cv2.FastFeatureDetector_create(threshold=FAST_TH, nonmaxSuppression=True)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    rng = np.random.default_rng(0)
    da = rng.integers(0, 256, (NFEAT, 32), dtype=np.uint8); db = rng.integers(0, 256, (NFEAT, 32), dtype=np.uint8)
    inv = 1.0
    sizes = []
    for _ in range(1, NLEVELS):
        inv /= SCALE
        sizes.append((int(round(W * inv)), int(round(H * inv))))

    def one(img):
        lv = img
        for l in range(NLEVELS):
            if l:
                lv = cv2.resize(lv, sizes[l - 1], interpolation=cv2.INTER_LINEAR)
            cv2.copyMakeBorder(lv, 16, 16, 16, 16, cv2.BORDER_REFLECT_101)
            fast.detect(lv, None)
            cv2.GaussianBlur(lv, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        bf.knnMatch(da, db, k=2)
    one(frames[0])
    t0 = time.perf_counter(); n = 0
    while n < 4 or time.perf_counter() - t0 < 1.5:
        one(frames[n % len(frames)]); n += 1
    return n / (time.perf_counter() - t0)


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_kind():
    """'reference' when oracle/_ref (the reference's own ORBextractor.cc, compiled in the build container) is present,
    else 'port' (the C oracle restating it)"""
    from oracle import reference as R
    return 'reference' if os.path.exists(R.SO) else 'port'


CPU_KIND_NOTE = {
    'reference': "extraction = the reference's own src/ORBextractor.cc (oracle/_ref, OpenCV primitives through stand-ins backed by the C "
                 "oracle); kNN2 = the C oracle's BFMatcher restatement (the reference's matcher needs its whole map/ROS stack)",
    'port': 'the C oracle restating the reference (oracle/_ref absent)'}


def cpu_reference(frames, threads, want_knn=True):
    """the reference's CPU path for one batch: ORBextractor over all frames (OpenMP over frames, one extractor per
    thread) + brute-force kNN2 between consecutive frames.  Returns seconds."""
    import numpy as np
    from oracle import oracle as O
    from oracle import reference as R
    t0 = time.perf_counter()
    if os.path.exists(R.SO):
        kps, n, desc = R.extract_batch(frames, NFEAT, SCALE, NLEVELS, FAST_TH, threads=threads)
    else:
        kps, n, desc = O.extract_batch(frames, NFEAT, SCALE, NLEVELS, FAST_TH, threads=threads)
    if want_knn:
        for f in range(len(frames) - 1):
            O.knn2(desc[f, :n[f]], desc[f + 1, :n[f + 1]], threads=threads)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle import oracle as O
    O.lib()
    cores = os.cpu_count() or 1
    nfr = args.ref_frames or args.batch
    frames = make_frames(pkg.synth, nfr, 1)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_reference(frames[:max(2, cores)], cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_reference(frames, cores)
    fps = nfr * args.steps / t
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': workload_config(args, nfr, args.gpus),
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': cpu_kind(),
                         'sample': '%d frames per step, OpenMP over frames; %s' % (nfr, CPU_KIND_NOTE[cpu_kind()])},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, frames_per_step, world):
    """the `config` object of both arms (identical for the same command line)"""
    return {'workload': set_shape(args.shape) + ' + brute-force Hamming kNN2 of consecutive frames (BASELINE %s shape, batched)'
                        % {'euroc': 'config 1', 'aqualoc': 'config 2', 'hd': 'config 5'}[args.shape],
            'frames_per_step_per_gpu': frames_per_step, 'keypoints': NFEAT,
            'l2': 'inputs alternate between two %d MB batches and each step streams a %.1f GB pyramid working set (> 126 MB L2)'
                  % (frames_per_step * W * H >> 20, 2.0 * frames_per_step * B_FRAME_BYTES / 1e9),
            'parallelism': 'frames sharded over %d GPU(s), no collective' % world}


# ------------------------------------------------------------------------------------------------ sharded kNN (config 4)
def sharded_knn_measure(pkg, torch, dist, dev, rank, world, local, n, passes, stream):
    """BASELINE config 4: n x n brute-force kNN2, train rows sharded contiguously over the ranks, queries replicated, per-query top-2
    all-gathered over NCCL and merged by (distance, global index).  Times the three phases separately (CUDA events on `stream`,
    which must be torch's current stream: the collectives are ordered against it), max over ranks, and ALWAYS checks the first
    2048 queries against the oracle's scan of the FULL train set on rank 0.  Returns the record on rank 0, None elsewhere."""
    import ctypes as C
    import numpy as np
    L = pkg.capi.lib(); chk = pkg.capi.check
    T, Q = pkg.synth.knn_database(n, n)
    b = pkg.sharding.shard_bounds(n, world)
    dq = torch.from_numpy(Q).to(dev); dt = torch.from_numpy(np.ascontiguousarray(T[b[rank]:b[rank + 1]])).to(dev)
    m = pkg.ORBmatcher(0.75, True, device=local)
    nq = n
    oi = torch.empty((nq, 2), dtype=torch.int32, device=dev); od = torch.empty_like(oi)
    gi = torch.empty((world, nq, 2), dtype=torch.int32, device=dev); gd = torch.empty_like(gi)
    mi = torch.empty_like(oi); md = torch.empty_like(od)
    sp = C.c_void_p(stream.cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one(rec=None):
        e = [ev() for _ in range(4)] if rec is not None else None
        if e: e[0].record(stream)
        chk(L.uvip_knn2_device(m.h, P(dq), nq, P(dt), dt.shape[0], int(b[rank]), P(oi), P(od), sp))
        if e: e[1].record(stream)
        if world > 1:
            dist.all_gather_into_tensor(gi, oi); dist.all_gather_into_tensor(gd, od)
        if e: e[2].record(stream)
        if world > 1:
            chk(L.uvip_knn2_merge_device(m.h, P(gi), P(gd), world, nq * 2, nq, P(mi), P(md), sp))
        if e: e[3].record(stream); rec.append(e)
    for _ in range(2):
        one()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    rec = []
    t0 = ev(); t1 = ev()
    t0.record(stream)
    for _ in range(passes):
        one(rec)
    t1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    tot = t0.elapsed_time(t1)
    ph = [sum(e[i].elapsed_time(e[i + 1]) for e in rec) for i in range(3)]
    if world > 1:
        tt = torch.tensor([tot] + ph, dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tot, ph = float(tt[0].item()), [float(v) for v in tt[1:].tolist()]
    ri_dev = (mi if world > 1 else oi)[:2048].cpu().numpy(); rd_dev = (md if world > 1 else od)[:2048].cpu().numpy()
    if rank != 0:
        return None
    from oracle import oracle as O
    tc = time.perf_counter()
    ri, rd = O.knn2(Q[:2048], T, threads=os.cpu_count() or 1)
    tc = time.perf_counter() - tc
    ok = bool(np.array_equal(ri_dev, ri) and np.array_equal(rd_dev, rd))
    ties = int((rd[:, 0] == rd[:, 1]).sum())
    popc = C.c_double(); chk(L.uvip_popc_peak(local, 4096, C.byref(popc)))
    pps = float(n) * n * passes / (tot * 1e-3)
    return {'workload': 'BASELINE config 4: %d x %d 256-bit descriptors (SURVEY Appendix B row 4 generator), train rows sharded over %d GPU(s), queries '
                        'replicated, NCCL all-gather of the per-query top-2 (16 B per query and rank) + (distance, global index) merge' % (n, n, world),
            'pairs_per_s': pps, 'unit': 'pairs/s', 'n_gpus': world, 'passes': passes, 'ms_per_pass': tot / passes,
            'phase_ms_per_pass': {'shard_knn2_kernel': ph[0] / passes, 'nccl_all_gather_x2': ph[1] / passes, 'merge_kernel': ph[2] / passes},
            'scaling': 'strong', 'matches_oracle_first_2048_queries': ok, 'exact_distance_ties_in_checked_queries': ties,
            'oracle_check_seconds': tc, 'popc_per_pair_issued': 4,
            'popc_frac_issued_of_n_gpus': pps * 4 / (popc.value * world), 'popc_frac_8_per_pair_accounting': pps * 8 / (popc.value * world),
            'popc_peak_measured_per_s_one_gpu': popc.value}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    L = pkg.capi.lib()
    chk = pkg.capi.check
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    cap = NFEAT + 8 * NLEVELS + 24

    frames = make_frames(pkg.synth, B, 1 + 100000 * rank)
    host_in = torch.from_numpy(frames).pin_memory()
    d_in = [host_in.to(dev, non_blocking=True), torch.roll(host_in, 1, 0).to(dev, non_blocking=True)]   # 2 x 92 MB > L2 together
    d_kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(B, dtype=torch.int32, device=dev)
    d_idx = torch.zeros((B, cap, 2), dtype=torch.int32, device=dev); d_dist = torch.zeros_like(d_idx)

    ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, pkg.ORBextractor.FAST_SCORE, FAST_TH, device=local, max_width=W, max_height=H,
                          max_batch=B)
    m = pkg.ORBmatcher(0.75, True, device=local)
    stream = torch.cuda.Stream(dev)              # an explicit stream: torch events and our kernels share it
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    knn_ev = []
    # Schedule of the timed region: two batches in flight, as in the host-buffer pipeline below.  Two extractor / matcher handle pairs
    # work on alternate steps, each pair on its own stream (extraction, then the kNN of the same step), so that every kernel's tail and
    # the latency-bound stages of one step are filled by the other step's kernels (measured: 131.2 k frames/s on one stream, 134.5 k with
    # only the kNN on a second stream, 141.7 k with this schedule).  UVIP_SERIAL=1 times the one-stream schedule instead.
    SERIAL = bool(os.environ.get('UVIP_SERIAL'))
    outs = [(d_kps, d_desc, d_n, d_idx, d_dist)]
    exs, mts, sts = [ex], [m], [stream]
    for _ in range(0 if SERIAL else max(1, int(os.environ.get('UVIP_INFLIGHT', '2'))) - 1):
        outs.append(tuple(torch.zeros_like(t) for t in outs[0]))
        exs.append(pkg.ORBextractor(NFEAT, SCALE, NLEVELS, pkg.ORBextractor.FAST_SCORE, FAST_TH, device=local, max_width=W, max_height=H, max_batch=B))
        mts.append(pkg.ORBmatcher(0.75, True, device=local))
        sts.append(torch.cuda.Stream(dev))
    sps = [C.c_void_p(t.cuda_stream) for t in sts]

    def step(i, timed=False, serial=False):
        j = 0 if serial else i % len(exs)
        o_kps, o_desc, o_n, o_idx, o_dist = outs[j]
        chk(L.uvip_extract_batch_device(exs[j].h, C.c_void_p(d_in[i & 1].data_ptr()), B, W, H, W, W * H, C.c_void_p(o_kps.data_ptr()),
                                        C.c_void_p(o_n.data_ptr()), cap, C.c_void_p(o_desc.data_ptr()), sps[j]))
        if timed:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(sts[j])
        # frame f (queries) against frame f+1 (train): B-1 consecutive pairs of the sequence
        chk(L.uvip_knn2_batch_device(mts[j].h, C.c_void_p(o_desc.data_ptr()), C.c_void_p(o_n.data_ptr()), cap * 32,
                                     C.c_void_p(o_desc.data_ptr() + cap * 32), C.c_void_p(o_n.data_ptr() + 4), cap * 32,
                                     B - 1, cap, C.c_void_p(o_idx.data_ptr()), C.c_void_p(o_dist.data_ptr()), cap, sps[j]))
        if timed:
            e1.record(sts[j]); knn_ev.append((e0, e1))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    for i in range(max(Wm, 2 * len(exs))):
        step(i)
    torch.cuda.synchronize(dev)
    for e in exs:
        e.status()                                         # loud if any capacity flag was raised
    n_host = d_n.cpu().numpy()
    assert n_host.min() >= NFEAT, 'extractor returned fewer than nfeatures keypoints: %d' % n_host.min()

    ex.profile(True)                                       # stage timers of the first handle: the dominant kernel's duration inside the timed region
    count = lambda: sum(e.launch_count() for e in exs) + sum(t.launch_count() for t in mts)
    launches0 = count()
    clocks = ClockSampler(local); clocks.start()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(sts[0])
    for t in sts[1:]:
        t.wait_event(t0)
    for i in range(K):
        step(i)
    for t in sts[1:]:
        sts[0].wait_stream(t)
    t1.record(sts[0])
    barrier()
    ms = t0.elapsed_time(t1)
    clk = clocks.stop()
    launches = count() - launches0
    for e in exs:
        e.status()
    region_ms, region_groups = ex.stage_ms()
    # stage times and the kNN time on ONE stream (a short extra pass outside the timed region): with two steps in flight a stage timer
    # also contains the other step's kernels, so only the dominant kernel's in-region duration above is kept from the timed region
    ex.profile(True)
    KS = K if SERIAL else max(3, min(K, 20))
    for i in range(KS):
        step(i, timed=True, serial=True)
    torch.cuda.synchronize(dev)
    stage_ms, ngroups = ex.stage_ms()
    ex.profile(False)
    knn_ms = sum(a.elapsed_time(b) for a, b in knn_ev)
    pairs_step = float((n_host[:-1].astype(np.int64) * n_host[1:].astype(np.int64)).sum())
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    fps = world * B * K / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI (pinned host memory in, host memory out), copies inside the timed region.
    # Each step uploads its own 256 frames and downloads keypoints, descriptors and kNN results; steps are issued through
    # uvip_extract_batch_submit / _wait so that the upload of step i+1 overlaps the kernels of step i (two host buffer sets).
    hp = lambda t: C.c_void_p(t.data_ptr())
    hb = []
    for j in range(2):
        hb.append(dict(inp=(host_in if j == 0 else torch.roll(host_in, 1, 0).pin_memory()),
                       k=torch.zeros((B, cap, 28), dtype=torch.uint8).pin_memory(), d=torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory(),
                       n=torch.zeros(B, dtype=torch.int32).pin_memory(), i=torch.zeros((B, cap, 2), dtype=torch.int32).pin_memory(),
                       dd=torch.zeros((B, cap, 2), dtype=torch.int32).pin_memory()))
    # UVIP_E2E_HANDLES (default 2) extractor / matcher handle pairs take alternate steps, UVIP_E2E_INFLIGHT (default 2 per handle) batches
    # are in flight.  A handle runs its kernels on ONE stream; with two handles the kernels of consecutive steps overlap like in the
    # HBM-resident schedule above, and with two batches queued per handle the upload of a step never waits for the host.  Measured on one
    # B200 (tools/e2e_sweep.sh): 1 handle / 2 in flight / 128-frame chunks (round 1's schedule) 120.9 k frames/s; 2 / 2 / 128: 124.6 k;
    # 2 / 4 / 128: 132.3 k; 2 / 4 / 64: 120.8 k; 3 / 6 / 128: 142.1 k; 2 / 4 / 256: 143.7 k = 96 % of the 149.4 k link ceiling
    NH = max(1, int(os.environ.get('UVIP_E2E_HANDLES', '2')))
    ex_e2es = [pkg.ORBextractor(NFEAT, SCALE, NLEVELS, pkg.ORBextractor.FAST_SCORE, FAST_TH, device=local, max_width=W, max_height=H,
                                max_batch=args.e2e_chunk or B) for _ in range(NH)]
    m_e2es = [m] + [pkg.ORBmatcher(0.75, True, device=local) for _ in range(NH - 1)]
    ex_e2e = ex_e2es[0]
    while len(hb) < max(2, int(os.environ.get('UVIP_E2E_INFLIGHT', str(2 * NH)))):
        hb.append(dict(inp=hb[len(hb) % 2]['inp'], k=torch.zeros((B, cap, 28), dtype=torch.uint8).pin_memory(), d=torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory(),
                       n=torch.zeros(B, dtype=torch.int32).pin_memory(), i=torch.zeros((B, cap, 2), dtype=torch.int32).pin_memory(),
                       dd=torch.zeros((B, cap, 2), dtype=torch.int32).pin_memory()))

    # uvip_extract_match_batch_submit chains the consecutive-frame kNN2 behind the extraction ON THE DEVICE: the descriptors a step
    # downloads are not uploaded again (round 1's e2e leg paid 8.9 MB of H2D per step for that: uvip_extract_batch_wait -> uvip_knn2_batch)
    NB = len(hb)                                           # batches in flight = host buffer sets; step i uses set i % NB on handle i % NH

    def submit(i):
        j = i % NB
        t = C.c_int(-1)
        chk(L.uvip_extract_match_batch_submit(ex_e2es[i % NH].h, m_e2es[i % NH].h, hp(hb[j]['inp']), B, W, H, W, W * H, hp(hb[j]['k']), hp(hb[j]['n']), cap,
                                              hp(hb[j]['d']), hp(hb[j]['i']), hp(hb[j]['dd']), C.byref(t)))
        return t.value

    def e2e_run(nsteps, mark_from=None, mark_to=None):
        """steps 0 .. nsteps-1 through the pipeline (NB batches in flight); returns the host time between the COMPLETION of step
        mark_from - 1 and the completion of step mark_to - 1: the pipeline is full at both ends, every timed step's upload and download
        lie inside (overlapped with its neighbours'), and exactly mark_to - mark_from steps complete in between"""
        tickets = {}
        ta = tb = None
        for i in range(min(NB, nsteps)):
            tickets[i] = submit(i)
        for i in range(nsteps):
            chk(L.uvip_extract_batch_wait(ex_e2es[i % NH].h, tickets.pop(i)))
            if mark_from is not None and i == mark_from - 1:
                ta = time.perf_counter()
            if mark_to is not None and i == mark_to - 1:
                tb = time.perf_counter()
            if i + NB < nsteps:
                tickets[i + NB] = submit(i + NB)
        return (tb - ta) if ta is not None and tb is not None else None

    Ke = max(3, K)
    e2e_run(NB)
    barrier()
    e2e_s = e2e_run(NB + Ke + NB, NB, NB + Ke)             # NB steps fill the pipeline, Ke are timed, NB keep it full behind them
    barrier()
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_fps = world * B * Ke / e2e_s
    h2d = B * W * H
    d2h = B * cap * 60 + B * 4 + 2 * (B - 1) * cap * 8
    # the chained results equal the two-call path's on the last step's buffers (cheap: the kNN of 3 pairs through the host entry point)
    jl = (NB + Ke + NB - 1) % NB
    chk_i = np.zeros((3, cap, 2), np.int32); chk_d = np.zeros((3, cap, 2), np.int32)
    chk(L.uvip_knn2_batch(m.h, hp(hb[jl]['d']), hp(hb[jl]['n']), cap * 32, C.c_void_p(hb[jl]['d'].data_ptr() + cap * 32),
                          C.c_void_p(hb[jl]['n'].data_ptr() + 4), cap * 32, 3, cap, C.c_void_p(chk_i.ctypes.data), C.c_void_p(chk_d.ctypes.data), cap))
    nn3 = hb[jl]['n'][:3].numpy()
    e2e_knn_ok = all(np.array_equal(hb[jl]['i'][f, :nn3[f]].numpy(), chk_i[f, :nn3[f]]) and np.array_equal(hb[jl]['dd'][f, :nn3[f]].numpy(), chk_d[f, :nn3[f]])
                     for f in range(3))
    # ---- link ceiling: the step's copies alone (pinned memory, both directions at once, every rank at the same time), i.e. what the
    # end-to-end path would reach if the kernels were free; at N > 1 this measures the HOST side shared by the N processes
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    def copies(n):
        for i in range(n):
            with torch.cuda.stream(s_up):
                d_in[i & 1].copy_(hb[i & 1]['inp'], non_blocking=True)
            with torch.cuda.stream(s_dn):
                hb[i & 1]['k'].copy_(d_kps, non_blocking=True); hb[i & 1]['d'].copy_(d_desc, non_blocking=True)
                hb[i & 1]['i'].copy_(d_idx, non_blocking=True); hb[i & 1]['dd'].copy_(d_dist, non_blocking=True)
    copies(2)
    barrier()
    tl0 = time.perf_counter()
    copies(Ke)
    barrier()
    link_s = time.perf_counter() - tl0
    if world > 1:
        tt = torch.tensor([link_s], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); link_s = float(tt.item())
    link_fps = world * B * Ke / link_s
    # single-frame latency of the reference-shaped call (uvip_extract = operator(), host image in, host keypoints out)
    one_k = np.zeros((cap, 7), np.float32); one_d = np.zeros((cap, 32), np.uint8)
    lat = []
    for i in range(40):
        nn = C.c_int(0)
        tl = time.perf_counter()
        chk(L.uvip_extract(ex_e2e.h, C.c_void_p(host_in[i % B].data_ptr()), W, H, W, C.c_void_p(one_k.ctypes.data), C.byref(nn), cap,
                           C.c_void_p(one_d.ctypes.data), None, 0, 0, 0, 1, 0))
        lat.append(time.perf_counter() - tl)
    lat_ms = float(np.median(lat[8:]) * 1e3)

    # ---- roofline of the dominant extraction kernel (algorithmic bytes of SURVEY 8(d) / its measured duration)
    peak, peak_kind = hbm_peak()
    # dominant KERNEL: named by the one-stream stage times ('pyramid' is eight launches, the largest a quarter of the stage, so it never
    # names it); its duration is the one measured INSIDE the timed region (stage timers of the first handle pair)
    dom = max((k for k in stage_ms if k != 'pyramid'), key=stage_ms.get)
    dom_ms = region_ms[dom] / max(region_groups, 1)
    achieved = B_FRAME_BYTES * B / (dom_ms * 1e-3) / 1e9
    traffic, traffic_stale = ncu_traffic()
    pipes, pipes_stale = ncu_pipes()
    static_ok = args.shape == 'euroc' and B == 256          # the committed launch list is one step of this workload at batch 256
    sm_hz = (clk.get('sm_mhz') or clk.get('sm_max_mhz') or 1965.0) * 1e6
    issue_peak = 148 * 4 * sm_hz                            # warp instructions per second: 4 schedulers per SM, one issue per clock each
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic.get(dom) if static_ok and not traffic_stale else None,
                'peak_source': peak_kind + ' (MEASURED_PEAKS.json hbm_gbs)' if peak_kind == 'measured' else 'fallback 6.65 TB/s',
                'algorithmic_bytes_per_launch': B_FRAME_BYTES * B, 'launch_ms': dom_ms,
                'launch_ms_one_stream': stage_ms[dom] / max(ngroups, 1),
                'stage_ms_per_step': {k: v / max(ngroups, 1) for k, v in stage_ms.items()}, 'knn_ms_per_step': knn_ms / KS,
                'stage_ms_source': 'one-stream pass of %d steps after the timed region (with two steps in flight a stage timer also contains the '
                                   'other step\'s kernels); launch_ms is the dominant kernel inside the timed region' % KS,
                'whole_step_frac': (B_FRAME_BYTES * B * K / (ms * 1e-3) / 1e9) / peak,
                'schedule': 'one stream' if SERIAL else 'two handle pairs on two streams, alternate steps (two batches in flight)'}
    if static_ok and not traffic_stale:
        tsum = sum(v for k, v in traffic.items() if not k.startswith('_'))
        roofline['traffic_all_stages'] = {'bytes_per_step': tsum, 'vs_algorithmic': tsum / (B_FRAME_BYTES * B),
                                          'per_stage': {k: v for k, v in traffic.items() if not k.startswith('_')}}
    elif static_ok:
        roofline['traffic_note'] = 'profiles/roofline_traffic.json was captured from other kernel sources (hash mismatch): not quoted'
    if static_ok and dom in pipes and not pipes_stale:
        # the extraction kernels are instruction-bound.  Second roofline, LIVE in its time term: the warp instructions the
        # kernel executes per launch (a property of code + input: ncu launch list of this command, same sources by hash) / the duration
        # measured in this run / the chip's issue rate at the SM clock sampled in this run
        wi = pipes[dom]['warp_instructions']
        step_wi = sum(v['warp_instructions'] for k, v in pipes.items() if not k.startswith('_'))
        roofline['issue'] = {'bound': 'issue', 'kernel': dom, 'warp_instructions_per_launch': wi,
                             'achieved': wi / (dom_ms * 1e-3), 'peak': issue_peak, 'unit': 'warp-instr/s', 'frac': wi / (dom_ms * 1e-3) / issue_peak,
                             'thread_instructions_per_pyramid_pixel': wi * 32.0 / (256 * 1117367.0),
                             'alu_pipe_pct_ncu': pipes[dom]['alu_pipe_pct'], 'issue_slot_pct_ncu': pipes[dom]['issue_slot_pct'],
                             'whole_step_warp_instructions': step_wi, 'whole_step_frac': step_wi * K / (ms * 1e-3) / issue_peak,
                             'source': 'instruction counts: profiles/kernel_pipes.json (sources hash %s); durations and SM clock: this run' % kernel_source_hash()}
    elif static_ok:
        roofline['issue'] = {'bound': 'issue', 'stale': True, 'note': 'profiles/kernel_pipes.json was captured from other kernel sources (hash mismatch)'}

    line = {
        'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms / K,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': workload_config(args, B, world),
        'clocks': clk,
        'e2e': {'value': e2e_fps, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': Ke, 'timing': 'steady state: from the completion of the step in front of the %d timed steps to the completion of the last of them, pipeline full at both ends' % Ke,
                'chunk_frames': args.e2e_chunk or B, 'single_frame_latency_ms': lat_ms,
                'pipeline': 'uvip_extract_match_batch_submit / _wait, %d batches in flight over %d handle pair(s); consecutive-frame kNN2 chained on the device (no descriptor re-upload)' % (NB, NH),
                'chained_knn_equals_two_call_path': bool(e2e_knn_ok),
                'link_ceiling_frames_per_s': link_fps, 'link_ceiling_gb_per_s_h2d': world * B * W * H * Ke / link_s / 1e9,
                'link_ceiling_what': 'the same pinned H2D (frames) and D2H (keypoints, descriptors, kNN results) copies of %d steps with no kernels, '
                                     'both directions at once, all %d rank(s) simultaneously (max over ranks)' % (Ke, world),
                'frac_of_min_device_link': e2e_fps / min(fps, link_fps)},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'hamming': {'pairs_per_s_in_step': pairs_step * KS / (knn_ms * 1e-3) if knn_ms > 0 else None, 'pairs_per_step': pairs_step},
    }

    if world > 1 and not args.no_extras:
        # BASELINE config 4 on the same ranks: the one path with a real exchange step (NCCL all-gather + merge), parity-checked
        nk = int(os.environ.get('UVIP_KNN_SHARDED_N', '1048576'))
        rec = sharded_knn_measure(pkg, torch, dist, dev, rank, world, local, nk, 2, stream)
        if rank == 0:
            line['knn_sharded'] = rec
        # BASELINE config 3 on the same ranks: grid-windowed search, frames sharded over the GPUs (no collective)
        rec3 = search_measure(pkg, torch, dist, dev, rank, world, local, B, 10, 3)
        if rank == 0:
            line['search_sharded'] = {k: rec3[k] for k in ('metric', 'value', 'unit', 'n_gpus', 'ms_per_step', 'scaling', 'matches_oracle_frame0', 'cpu_baseline')
                                      if k in rec3}
            line['search_sharded']['workload'] = rec3['config']['workload']
            line['search_sharded']['frames_per_step_per_gpu'] = B
    if rank == 0 and not args.no_extras:
        # Hamming-only leg: database-scale kNN2 (cfg4 shape at 64k x 64k) against the measured popc-pipe peak
        nq = nt = 65536
        T, Q = pkg.synth.knn_database(nt, nq)
        dT = torch.from_numpy(T).to(dev); dQ = torch.from_numpy(Q).to(dev)
        oi = torch.zeros((nq, 2), dtype=torch.int32, device=dev); od = torch.zeros_like(oi)
        run = lambda: chk(L.uvip_knn2_device(m.h, C.c_void_p(dQ.data_ptr()), nq, C.c_void_p(dT.data_ptr()), nt, 0,
                                             C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), sp))
        for _ in range(3):
            run()
        torch.cuda.synchronize(dev)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5):
            run()
        b.record(stream); torch.cuda.synchronize(dev)
        pps = 5.0 * nq * nt / (a.elapsed_time(b) * 1e-3)
        popc = C.c_double()
        chk(L.uvip_popc_peak(local, 4096, C.byref(popc)))
        # the kernel ISSUES 4 POPC per pair (carry-save tree over the 8 XOR words); the SURVEY's accounting of 8 per pair is secondary
        line['hamming'].update({'knn2_64k_pairs_per_s': pps, 'popc_per_pair_issued': 4, 'popc_peak_measured_per_s': popc.value,
                                'popc_frac_issued': pps * 4 / popc.value, 'popc_frac_8_per_pair_accounting': pps * 8 / popc.value,
                                'popc_peak_nominal_per_s': 148 * 16 * (clk.get('sm_max_mhz') or 1965.0) * 1e6})
        if world == 1:
            # cpu_baseline: the oracle on this box's host cores, bounded sample of the same workload
            cores = os.cpu_count() or 1
            ns = min(B, max(16, 8 * cores))
            sample = frames[:ns]
            cpu_reference(sample[:max(2, cores)], cores)
            tc = cpu_reference(sample, cores)
            reps = max(1, min(40, int(10.0 / max(tc, 1e-3))))          # ~10 s of all-core CPU work in total
            for _ in range(reps - 1):
                tc += cpu_reference(sample, cores)
            n1 = max(4, min(ns, 128))
            t1c = cpu_reference(sample[:n1], 1)
            line['cpu_baseline'] = {'value': reps * len(sample) / tc, 'unit': 'frames/s', 'cores': cores, 'kind': cpu_kind(), 'what': CPU_KIND_NOTE[cpu_kind()],
                                    'sample': '%d passes over %d frames of the same workload (extraction + consecutive-frame kNN2), OpenMP over '
                                              'frames on all host cores; single-thread figure on %d frames' % (reps, len(sample), n1),
                                    'single_thread_frames_per_s': n1 / t1c,
                                    'cv2_composite_frames_per_s': {'1_thread': cv2_composite(sample, 1), 'all_threads': cv2_composite(sample, cores),
                                                                   'what': 'SURVEY 8(d) cross-check (iii): cv2 %s resize / copyMakeBorder / FAST / GaussianBlur on 8 '
                                                                           'levels + BFMatcher.knnMatch(k=2) 1000 x 1000 per frame (no quadtree, orientation or '
                                                                           'descriptors): bounds how soft the scalar OpenCV stand-ins of the reference build are' % _cv2_version()}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_knn(args):
    """config 4: nq x nt brute-force kNN2, train rows sharded contiguously over the ranks, queries replicated, per-query
    top-2 all-gathered over NCCL and merged by (distance, global index).  value = descriptor pairs per second."""
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    K, Wm = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local); clocks.start()
    rec = sharded_knn_measure(pkg, torch, dist, dev, rank, world, local, args.knn_n, K, stream)
    clk = clocks.stop()
    if rank == 0:
        print(json.dumps({'metric': 'Hamming pairs/s (brute-force kNN2, database sharded over GPUs, NCCL top-2 merge)', 'value': rec['pairs_per_s'],
                          'unit': 'pairs/s', 'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': rec['ms_per_pass'], 'higher_is_better': True,
                          'scaling': 'strong', 'vs_baseline': None, 'dtype': 'u32 popc', 'data': 'synthetic',
                          'config': {'workload': rec['workload']}, 'clocks': clk,
                          'roofline': {'bound': 'popc', 'popc_per_pair_issued': 4, 'achieved_popc_per_s': rec['pairs_per_s'] * 4,
                                       'peak_popc_per_s_measured_one_gpu': rec['popc_peak_measured_per_s_one_gpu'],
                                       'frac': rec['popc_frac_issued_of_n_gpus'], 'frac_8_per_pair_accounting': rec['popc_frac_8_per_pair_accounting']},
                          'phase_ms_per_pass': rec['phase_ms_per_pass'],
                          'matches_oracle_first_2048_queries': rec['matches_oracle_first_2048_queries'],
                          'exact_distance_ties_in_checked_queries': rec['exact_distance_ties_in_checked_queries']}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_search(args):
    """BASELINE config 3 as a throughput workload: SearchByProjection-style grid-windowed matching, 10 000 projected map points
    against a 2000-keypoint frame (radius search on the 64 x 48 frame grid, top-2 + ratio, claims), `--batch` frames per step in
    one launch pair (grid build + search), frames sharded over the GPUs with no collective.  value = map points per second."""
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    line = search_measure(pkg, torch, dist, dev, rank, world, local, args.batch, args.steps, max(args.warmup, 3))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def search_measure(pkg, torch, dist, dev, rank, world, local, F, K, Wm):
    """the measurement of run_search on an existing process group; returns the record on rank 0, None elsewhere"""
    import ctypes as C
    import numpy as np
    Wf, Hf, NQ, NK = 752, 480, 10000, 2000
    ndistinct = min(F, 16)
    cases = [pkg.synth.projection_case(seed_f=3 + 10 * i + 1000 * rank, seed_p=4 + 10 * i + 1000 * rank) for i in range(ndistinct)]
    m = pkg.ORBmatcher(0.8, True, device=local)
    sf = pkg.ORBextractor(1000, 1.2, 8, 1, 20, device=local, max_width=Wf, max_height=Hf).tables()[0]
    rad = [m.projection_radius(c['view_cos'], c['level'], sf, 1.0) for c in cases]
    pick = [i % ndistinct for i in range(F)]
    stack = lambda fn, dt: torch.from_numpy(np.ascontiguousarray(np.stack([fn(i) for i in pick]), dt)).to(dev)
    t = dict(u=stack(lambda i: cases[i]['u'], np.float32), v=stack(lambda i: cases[i]['v'], np.float32), r=stack(lambda i: rad[i], np.float32),
             lo=stack(lambda i: cases[i]['level'] - 1, np.int32), hi=stack(lambda i: cases[i]['level'], np.int32),
             qd=stack(lambda i: cases[i]['qdesc'], np.uint8), x=stack(lambda i: cases[i]['kx'], np.float32),
             y=stack(lambda i: cases[i]['ky'], np.float32), o=stack(lambda i: cases[i]['octave'], np.int32),
             kd=stack(lambda i: cases[i]['kdesc'], np.uint8),
             nq=torch.full((F,), NQ, dtype=torch.int32, device=dev), nk=torch.full((F,), NK, dtype=torch.int32, device=dev),
             taken=torch.full((F, NK), -1, dtype=torch.int32, device=dev), match=torch.zeros((F, NQ), dtype=torch.int32, device=dev),
             counts=torch.zeros((F, 2), dtype=torch.int32, device=dev))
    P = lambda k: t[k].data_ptr()
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)

    def step():
        t['taken'].fill_(-1)                                 # every step starts from a frame without map points
        m.search_window_batch_device(0, 100, (0, Wf, 0, Hf), F, (P('u'), P('v'), P('r'), P('lo'), P('hi'), P('qd')), P('nq'), NQ,
                                     (P('x'), P('y'), P('o'), P('kd')), P('nk'), NK, P('taken'), P('match'), P('counts'), stream.cuda_stream)
    for _ in range(Wm):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    launches0 = m.launch_count()
    clocks = ClockSampler(local); clocks.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = m.launch_count() - launches0 + K              # + the fill kernel of every step
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = float(tt.item())
    mps = float(world) * F * NQ * K / (ms * 1e-3)
    counts = t['counts'].cpu().numpy()
    # ---- end to end: the single-frame host-buffer call (uvip_search_frame, H2D + D2H inside), one frame after another
    c0 = cases[0]
    te = time.perf_counter(); ne = 0
    while ne < 20 or time.perf_counter() - te < 0.5:
        c = cases[ne % ndistinct]
        n1, match1, taken1 = m.search_frame(0, 100, c['u'], c['v'], rad[ne % ndistinct], c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'],
                                            c['octave'], c['kdesc'], c['bounds'])
        ne += 1
    e2e_s = time.perf_counter() - te
    if world > 1:
        tt = torch.tensor([e2e_s / ne], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); e2e_s = float(tt.item()) * ne
    line = None
    if rank == 0:
        from oracle import oracle as O
        # candidate pairs per frame (one 8-popc distance each) from the oracle's GetFeaturesInArea on the distinct frames
        inv_w = np.float32(64.0) / np.float32(Wf); inv_h = np.float32(48.0) / np.float32(Hf)
        c = cases[0]
        start, items = O.grid_build(c['kx'], c['ky'], 0.0, 0.0, float(inv_w), float(inv_h))
        ncand = sum(len(O.features_in_area(c['kx'], c['ky'], c['octave'], start, items, 0.0, 0.0, float(inv_w), float(inv_h), float(c['u'][q]),
                                           float(c['v'][q]), float(rad[0][q]), int(c['level'][q]) - 1, int(c['level'][q]))) for q in range(NQ))
        on, om, otk = O.search_window(0, 100, 0.8, c['u'], c['v'], rad[0], c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'], c['octave'],
                                      c['kdesc'], start, items, 0.0, 0.0, float(inv_w), float(inv_h))
        ok = bool(np.array_equal(t['match'][0].cpu().numpy(), om) and counts[0, 0] == on and n1 >= 0)
        # algorithmic bytes: per query 32 + 16 read, 16 written; per frame 2000 x 44 B + grid CSR
        bytes_frame = NQ * (32 + 16 + 16) + NK * 44 + (64 * 48 + 1) * 4 + NK * 4
        peak, peak_kind = hbm_peak()
        achieved = bytes_frame * F * K / (ms * 1e-3) / 1e9
        popc = C.c_double(); pkg.capi.check(pkg.capi.lib().uvip_popc_peak(local, 4096, C.byref(popc)))
        line = {'metric': 'SearchByProjection map points/s (BASELINE config 3: 10k projected map points vs 2000-kp frame)', 'value': mps,
                'unit': 'map points/s', 'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32 geometry + u32 popc', 'data': 'synthetic',
                'config': {'workload': 'BASELINE config 3: grid-windowed SearchByProjection, 10000 map points x 2000 keypoints per frame, th=1, ratio 0.8',
                           'frames_per_step_per_gpu': F, 'distinct_frames': ndistinct, 'l2': 'the batch (%.0f MB) exceeds nothing: this stage is latency-bound, '
                           'not bandwidth-bound; inputs stay L2-resident by design' % (F * (NQ * 56 + NK * 48) / 1e6)},
                'clocks': clk, 'gpu_launches': int(launches),
                'e2e': {'value': world * NQ * ne / e2e_s, 'unit': 'map points/s', 'h2d_bytes_per_step': NQ * 52 + NK * 48 + (64 * 48 + 1) * 4 + NK * 8,
                        'd2h_bytes_per_step': NQ * 4 + NK * 4 + (64 * 48 + 1) * 4 + NK * 4 + 8, 'what': 'single-frame host-buffer call uvip_search_frame '
                        '(grid + search on the device), one frame after another (the shape of the reference call)', 'frames': ne, 'ms_per_frame': 1e3 * e2e_s / ne},
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                             'peak_source': peak_kind, 'algorithmic_bytes_per_launch': bytes_frame * F,
                             'candidate_pairs_per_s': float(world) * ncand * F * K / (ms * 1e-3), 'candidate_pairs_per_frame': ncand,
                             'popc_frac_8_per_pair': float(ncand) * F * K / (ms * 1e-3) * 8 / popc.value,
                             'claim_rounds_per_frame': float(counts[:, 1].mean()), 'note': 'latency-bound gather: one CTA per frame, queries re-run until '
                             'the claim table is a fixed point'},
                'matches_oracle_frame0': ok}
        # cpu_baseline: the reference's own compiled ORBmatcher::SearchByProjection (oracle/_ref), single thread as the reference runs it
        from oracle import reference as R
        if R.matcher_available():
            sfl = [float(x) for x in sf]
            tot = 0.0; nfr = 0
            while tot < 1.0 or nfr < 8:
                c = cases[nfr % ndistinct]
                R.search_by_projection_mps(c['kx'], c['ky'], c['octave'], c['kdesc'], [0, Wf, 0, Hf], sfl, c['u'], c['v'], c['level'], c['view_cos'],
                                           c['qdesc'], 1.0, 0.8)
                tot += R.last_call_seconds(); nfr += 1
            line['cpu_baseline'] = {'value': NQ * nfr / tot, 'unit': 'map points/s', 'cores': 1, 'kind': 'reference',
                                    'sample': "%d frames through the reference's own compiled ORBmatcher::SearchByProjection (oracle/_ref, scene "
                                              'construction excluded); the reference runs this on one Tracking thread' % nfr}
        else:
            tc = time.perf_counter(); nfr = 0
            while time.perf_counter() - tc < 1.0:
                O.search_window(0, 100, 0.8, c['u'], c['v'], rad[0], c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'], c['octave'], c['kdesc'],
                                start, items, 0.0, 0.0, float(inv_w), float(inv_h)); nfr += 1
            line['cpu_baseline'] = {'value': NQ * nfr / (time.perf_counter() - tc), 'unit': 'map points/s', 'cores': 1, 'kind': 'port',
                                    'sample': '%d frames through the C oracle' % nfr}
    return line


def run_cfg5(args):
    """BASELINE config 5 as written: 8 concurrent synthetic sequences x 4096 frames of 1280x1024, sequence s ->
    GPU s, frame f = synth_frame(100000 s + f, dx = 3f mod 29, dy = 2f mod 23), extractor(2000, 1.2, 8, FAST 20/7) + brute-force kNN2
    between consecutive frames of the sequence.  The sequence is generated on the device (uvip_synth_frames_device, byte-identical to
    the numpy generator) and parked in pinned host memory.  e2e: ONE uvip_extract_match_batch_submit over the whole sequence — the
    library streams it in chunks of `--batch` frames through its copy / compute pipeline, matching frames across chunk boundaries on
    the device.  value: the same chunks from an HBM-resident copy through the device-pointer entry points.  One chunk is checked
    against the reference's own compiled extractor (oracle/_ref) and the oracle's kNN2 on rank 0."""
    import ctypes as C
    import numpy as np
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    set_shape('hd')
    L = pkg.capi.lib(); chk = pkg.capi.check
    NF, CH = args.seq_frames, min(args.batch, 64) if args.batch == 256 else args.batch
    cap = NFEAT + 8 * NLEVELS + 24
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    # ---- the sequence: generated on the device chunk by chunk, kept in HBM (5.4 GB) and in pinned host memory
    fidx = torch.arange(NF, dtype=torch.int64)
    seeds = (100000 * rank + fidx).to(dev)
    dxy = torch.stack([(3 * fidx) % 29, (2 * fidx) % 23], 1).to(torch.int32).to(dev).contiguous()
    d_seq = torch.empty((NF, H, W), dtype=torch.uint8, device=dev)
    tg = time.perf_counter()
    for f0 in range(0, NF, 256):
        nb = min(256, NF - f0)
        chk(L.uvip_synth_frames_device(P(seeds[f0:]), P(dxy[f0:]), P(seeds[f0:]), nb, W, H, P(d_seq[f0:]), W * H, sp))
    torch.cuda.synchronize(dev)
    gen_s = time.perf_counter() - tg
    h_seq = torch.empty((NF, H, W), dtype=torch.uint8).pin_memory()
    h_seq.copy_(d_seq); torch.cuda.synchronize(dev)
    h_k = torch.zeros((NF, cap, 28), dtype=torch.uint8).pin_memory(); h_d = torch.zeros((NF, cap, 32), dtype=torch.uint8).pin_memory()
    h_n = torch.zeros(NF, dtype=torch.int32).pin_memory()
    h_i = torch.zeros((NF - 1, cap, 2), dtype=torch.int32).pin_memory(); h_dd = torch.zeros((NF - 1, cap, 2), dtype=torch.int32).pin_memory()
    ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, pkg.ORBextractor.FAST_SCORE, FAST_TH, device=local, max_width=W, max_height=H, max_batch=CH)
    m = pkg.ORBmatcher(0.75, True, device=local)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- end to end: the whole sequence through one submit / wait (host frames in, host keypoints / descriptors / matches out)
    def e2e_pass(nframes):
        t = C.c_int(-1)
        chk(L.uvip_extract_match_batch_submit(ex.h, m.h, P(h_seq), nframes, W, H, W, W * H, P(h_k), P(h_n), cap, P(h_d), P(h_i), P(h_dd), C.byref(t)))
        chk(L.uvip_extract_batch_wait(ex.h, t.value))
    e2e_pass(min(NF, 4 * CH))                              # warm-up: plan, staging buffers, first-use allocations
    barrier()
    clocks = ClockSampler(local); clocks.start()
    te = time.perf_counter()
    e2e_pass(NF)
    barrier()
    e2e_s = time.perf_counter() - te
    clk = clocks.stop()
    n_host = h_n.numpy().copy()
    # ---- HBM-resident: chunks through the device-pointer entry points, two output sets, the pair across a chunk boundary matched too
    outs = [dict(k=torch.zeros((CH, cap, 28), dtype=torch.uint8, device=dev), d=torch.zeros((CH, cap, 32), dtype=torch.uint8, device=dev),
                 n=torch.zeros(CH, dtype=torch.int32, device=dev), i=torch.zeros((CH, cap, 2), dtype=torch.int32, device=dev),
                 dd=torch.zeros((CH, cap, 2), dtype=torch.int32, device=dev)) for _ in range(2)]
    launches0 = ex.launch_count() + m.launch_count()

    def dev_pass(nframes):
        for c, f0 in enumerate(range(0, nframes, CH)):
            nb = min(CH, nframes - f0); o = outs[c & 1]; pv = outs[(c & 1) ^ 1]
            chk(L.uvip_extract_batch_device(ex.h, P(d_seq[f0:]), nb, W, H, W, W * H, P(o['k']), P(o['n']), cap, P(o['d']), sp))
            if c:                                         # last frame of the previous chunk against the first frame of this one
                chk(L.uvip_knn2_batch_device(m.h, C.c_void_p(pv['d'].data_ptr() + (CH - 1) * cap * 32), C.c_void_p(pv['n'].data_ptr() + 4 * (CH - 1)), cap * 32,
                                             P(o['d']), P(o['n']), cap * 32, 1, cap, P(pv['i'][CH - 1:]), P(pv['dd'][CH - 1:]), cap, sp))
            if nb > 1:
                chk(L.uvip_knn2_batch_device(m.h, P(o['d']), P(o['n']), cap * 32, C.c_void_p(o['d'].data_ptr() + cap * 32), C.c_void_p(o['n'].data_ptr() + 4),
                                             cap * 32, nb - 1, cap, P(o['i']), P(o['dd']), cap, sp))
    dev_pass(min(NF, 4 * CH))
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    dev_pass(NF)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ex.launch_count() + m.launch_count() - launches0
    ex.status()
    if world > 1:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms, e2e_s = [float(v) for v in tt.tolist()]
    if rank == 0:
        # ---- parity spot check of one chunk of the streamed (end-to-end) results: the reference's compiled extractor + the oracle's kNN2
        from oracle import oracle as O
        from oracle import reference as R
        c0 = min(10 * CH, max(0, NF - CH)) if NF > CH else 0
        nchk = min(8, NF - c0)
        fr = h_seq[c0:c0 + nchk].numpy()
        if os.path.exists(R.SO):
            rk, rn, rd = R.extract_batch(fr, NFEAT, SCALE, NLEVELS, FAST_TH, threads=os.cpu_count() or 1, cap=cap)
            kind = 'reference (oracle/_ref)'
        else:
            rk, rn, rd = O.extract_batch(fr, NFEAT, SCALE, NLEVELS, FAST_TH, cap=cap)
            kind = 'oracle port'
        ok_ext = bool(np.array_equal(rn, n_host[c0:c0 + nchk]))
        kview = h_k.numpy().view(pkg.capi.KP_DTYPE).reshape(NF, cap)
        for f in range(nchk):
            nn = int(rn[f])
            ok_ext = ok_ext and bool(np.array_equal(kview[c0 + f, :nn]['x'], rk[f, :nn]['x']) and np.array_equal(kview[c0 + f, :nn]['y'], rk[f, :nn]['y'])
                                     and np.array_equal(kview[c0 + f, :nn]['angle'], rk[f, :nn]['angle']) and np.array_equal(h_d[c0 + f, :nn].numpy(), rd[f, :nn]))
        ok_knn = True
        pairs = list(range(c0 - 1 if c0 > 0 else c0, c0 + nchk - 1))      # includes the pair that crosses the chunk boundary in front of c0
        for f in pairs:
            na = int(n_host[f]); nb_ = int(n_host[f + 1])
            oi, od = O.knn2(h_d[f, :na].numpy(), h_d[f + 1, :nb_].numpy(), threads=os.cpu_count() or 1)
            ok_knn = ok_knn and bool(np.array_equal(h_i[f, :na].numpy(), oi) and np.array_equal(h_dd[f, :na].numpy(), od))
        peak, peak_kind = hbm_peak()
        fps = world * NF / (ms * 1e-3)
        line = {'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': 1, 'warmup': 1, 'ms_per_step': ms, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': {'workload': 'BASELINE config 5 as written: %d concurrent synthetic sequences x %d frames of %dx%d (seed 100000 s + f, dx = 3f mod 29, '
                                       'dy = 2f mod 23), sequence -> GPU, ORBextractor %d kp / 8 levels / 1.2 / FAST 20-7 + brute-force kNN2 of consecutive frames, '
                                       'streamed in chunks of %d frames with matching across chunk boundaries' % (world, NF, W, H, NFEAT, CH),
                           'frames_per_sequence': NF, 'chunk_frames': CH, 'keypoints': NFEAT, 'input_gb_per_gpu': NF * W * H / 1e9,
                           'l2': 'every chunk streams a %.0f MB pyramid working set (> 126 MB L2)' % (2.0 * CH * B_FRAME_BYTES / 1e6),
                           'parallelism': 'one sequence per GPU, no collective'},
                'clocks': clk, 'gpu_launches': int(launches),
                'e2e': {'value': world * NF / e2e_s, 'unit': 'frames/s', 'h2d_bytes_per_step': NF * W * H, 'd2h_bytes_per_step': NF * cap * 60 + NF * 4 + 2 * (NF - 1) * cap * 8,
                        'seconds_per_sequence': e2e_s, 'what': 'one uvip_extract_match_batch_submit + wait over the whole sequence from pinned host memory'},
                'roofline': {'bound': 'hbm', 'achieved': B_FRAME_BYTES * NF / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                             'frac': B_FRAME_BYTES * NF / (ms * 1e-3) / 1e9 / peak, 'traffic': None, 'what': 'whole sequence pass (all stages + kNN2), algorithmic bytes of SURVEY 8(d)',
                             'algorithmic_bytes_per_frame': B_FRAME_BYTES},
                'keypoints_per_frame': {'min': int(n_host.min()), 'mean': float(n_host.mean())},
                'generator': {'seconds_for_sequence_on_device': gen_s},
                'parity_spot_check': {'frames': [c0, c0 + nchk - 1], 'pairs': [pairs[0], pairs[-1]] if pairs else None, 'against': kind,
                                      'extraction_identical': ok_ext, 'knn2_identical_incl_chunk_boundary_pair': ok_knn}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_next(args):
    """the SURVEY 8f rows built so far (CLAHE, DBoW2 descent, KLT), each timed next to its oracle on the host cores"""
    import ctypes as C
    import numpy as np
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from oracle import oracle as O
    dev = torch.device('cuda', 0); torch.cuda.set_device(0)
    L = pkg.capi.lib(); chk = pkg.capi.check
    stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    K = args.steps
    out = {'metric': 'next rows (SURVEY 8f)', 'n_gpus': 1, 'steps': K, 'data': 'synthetic'}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(K):
            fn()
        b.record(stream); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / K

    # CLAHE: 256 frames resident in HBM, in place
    B = 256
    frames = make_frames(pkg.synth, B, 1)
    d = torch.from_numpy(frames).to(dev)
    ex = pkg.ORBextractor(NFEAT, SCALE, NLEVELS, 1, FAST_TH, device=0, max_width=W, max_height=H, max_batch=1)
    ms = timed(lambda: chk(L.uvip_clahe_batch_device(ex.h, C.c_void_p(d.data_ptr()), B, W, H, W, W * H, 4.0, 12, 12, C.c_void_p(d.data_ptr()), W, W * H, sp)))
    t0 = time.perf_counter(); [O.clahe(f) for f in frames[:32]]; tc = (time.perf_counter() - t0) / 32
    peak, _ = hbm_peak()
    out['clahe'] = {'frames_per_s': B / (ms * 1e-3), 'ms_per_256_frames': ms, 'hbm_frac_of_measured': (2.0 * W * H * B / (ms * 1e-3) / 1e9) / peak,
                    'cpu_oracle_frames_per_s_1thread': 1.0 / tc}
    # DBoW2 descent: k=10, L=5 synthetic vocabulary (111k nodes), 256 x 1000 descriptors
    tree, _ = pkg.synth.synthetic_vocabulary(10, 5, seed=3)
    voc = pkg.ORBVocabulary(tree)
    nd = 256000
    desc = torch.from_numpy(pkg.synth.random_descriptors(5, nd)).to(dev)
    wid = torch.empty(nd, dtype=torch.int32, device=dev); nid = torch.empty_like(wid); wt = torch.empty(nd, dtype=torch.float64, device=dev)
    ms = timed(lambda: chk(L.uvip_bow_transform_device(voc.h, C.c_void_p(desc.data_ptr()), nd, 4, C.c_void_p(wid.data_ptr()), C.c_void_p(nid.data_ptr()),
                                                       C.c_void_p(wt.data_ptr()), sp)))
    hd = desc[:20000].cpu().numpy()
    t0 = time.perf_counter(); O.bow_transform(tree, hd, 4); tc = time.perf_counter() - t0
    out['bow_descent'] = {'descriptors_per_s': nd / (ms * 1e-3), 'ms_per_256k': ms, 'distances_per_descriptor': 10 * 5,
                          'cpu_oracle_descriptors_per_s_1thread': 20000 / tc}
    # KLT: 1000 points, 21x21, 5 levels, through the host-buffer calls (copies inside)
    a_img = pkg.synth.synth_frame(1, W, H); b_img = pkg.synth.synth_frame(1, W, H, dx=5, dy=3, noise_seed=2)
    kps, _ = ex(a_img)
    p0 = np.stack([kps['x'], kps['y']], 1).astype(np.float32)[:1000]
    klt = pkg.KLTTracker(W, H, 21, 5)
    t0 = time.perf_counter()
    for _ in range(K):
        klt.build_pyramid(0, a_img); klt.build_pyramid(1, b_img)
        p1, st, err = klt.track(0, 1, p0, p0)
    tg = (time.perf_counter() - t0) / K
    t0 = time.perf_counter()
    P0 = O.LKPyramid(a_img, 21, 5); P1 = O.LKPyramid(b_img, 21, 5); O.lk_track(P0, P1, p0, p0, 21, 5, 30, 0.01, 8)
    tc = time.perf_counter() - t0
    out['klt'] = {'frame_pairs_per_s_e2e': 1.0 / tg, 'points': len(p0), 'tracked': int(st.sum()), 'cpu_oracle_frame_pairs_per_s_1thread': 1.0 / tc}
    # batched, device-resident: a 64-frame sequence (pyramids of all frames, then the 63 consecutive pairs in one launch), 1000 points per pair
    NS = 64
    seq = torch.from_numpy(np.stack([pkg.synth.synth_frame(1, W, H, dx=(3 * f) % 29, dy=(2 * f) % 23, noise_seed=100 + f) for f in range(NS)])).to(dev)
    klt_b = pkg.KLTTracker(W, H, 21, 5, nslots=NS)
    prevp = torch.from_numpy(np.tile(p0[None], (NS - 1, 1, 1))).to(dev).contiguous(); nextp = prevp.clone()
    stt = torch.zeros((NS - 1, len(p0)), dtype=torch.uint8, device=dev); err_t = torch.zeros((NS - 1, len(p0)), dtype=torch.float32, device=dev)
    Pp = lambda t: C.c_void_p(t.data_ptr())
    def klt_step():
        nextp.copy_(prevp)
        chk(L.uvip_klt_track_sequence_device(klt_b.h, Pp(seq), NS, W, H, W, W * H, Pp(prevp), Pp(nextp), len(p0), 5, 30, 0.01, 8, 1e-4, Pp(stt), Pp(err_t), sp))
    ms = timed(klt_step)
    # algorithmic bytes of the tracker: per frame the level-0 image + its pyramid (4/3) is read once, the int16 derivative pair planes (4 B per
    # pixel) are written and read once; the windows themselves are L2 / shared-memory traffic
    klt_bytes = NS * W * H * (4.0 / 3.0) * (1 + 1 + 4 + 4)
    out['klt_sequence_device'] = {'frames': NS, 'pairs': NS - 1, 'points_per_pair': len(p0), 'ms_per_sequence': ms, 'frame_pairs_per_s': (NS - 1) / (ms * 1e-3),
                                  'points_per_s': (NS - 1) * len(p0) / (ms * 1e-3), 'tracked_frac': float(stt.float().mean().item()),
                                  'roofline': {'bound': 'hbm', 'achieved': klt_bytes / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                                               'frac': klt_bytes / (ms * 1e-3) / 1e9 / peak, 'traffic': None,
                                               'note': 'the tracker is arithmetic/latency-bound (30 Lucas-Kanade iterations over a 21x21 window per point and level, '
                                                       'one warp per point): see profiles/ for the ncu capture of k_klt_track'}}
    # ---- N4 (descriptor half): batched MapPoint::ComputeDistinctiveDescriptors, 100k map points with 2..30 observations
    rng = np.random.default_rng(4)
    sizes = rng.integers(2, 31, 100000)
    start = np.zeros(len(sizes) + 1, np.int32); start[1:] = np.cumsum(sizes)
    dd = pkg.synth.random_descriptors(8, int(start[-1]))
    mm = pkg.ORBmatcher(0.6, True, device=0)
    mm.distinctive_descriptors(dd, start)
    tg = time.perf_counter()
    for _ in range(K):
        bi, _ = mm.distinctive_descriptors(dd, start)
    tg = (time.perf_counter() - tg) / K
    ns = 5000
    tc = time.perf_counter(); obi, _ = O.distinctive_descriptors(dd[:start[ns]], start[:ns + 1]); tc = time.perf_counter() - tc
    assert np.array_equal(bi[:ns], obi)
    out['distinctive_descriptors'] = {'map_points_per_s_e2e': len(sizes) / tg, 'observations': int(start[-1]),
                                      'cpu_oracle_map_points_per_s_1thread': ns / tc}
    print(json.dumps(out), flush=True)
    return 0


if __name__ == '__main__':
    a = parse()
    set_shape(a.shape)
    assert a.shape != 'euroc' or B_FRAME_BYTES == 1177367
    if a.impl == 'reference':
        sys.exit(run_reference(a))
    sys.exit(run_knn(a) if a.workload == 'knn' else run_next(a) if a.workload == 'next' else run_search(a) if a.workload == 'search'
             else run_cfg5(a) if a.workload == 'cfg5' else run_ours(a))
