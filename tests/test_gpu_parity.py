"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI (ctypes -> libuvip_orb.so), against the CPU
oracle on the same seeded inputs.  Bars (BASELINE.json north_star): pyramid, blur, FAST keypoint sets, Hamming
distances and match indices bit-exact; angles within 1e-3 rad; descriptor bits >= 99.9 %."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = 1e-3 * 180.0 / np.pi       # 1e-3 rad
DESC_BITS_MIN = 0.999


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope='module')
def gpu(pkg):
    if pkg.capi.lib().uvip_device_count() < 1:
        pytest.fail('no CUDA device: the gpu-marked tests must run on the B200 box')
    return pkg


def compare_frame(gpu_ex, ora_ex, img, golden=None, tag=None, **kw):
    kps, desc = gpu_ex(img, **kw)
    okw = dict(full_detect=kw.get('FullDetect', True), num_needed=kw.get('num_featsneeded', 0),
               min_px_dist=kw.get('min_px_dist', 1))
    okps, odesc = ora_ex(img, keypoints=kw.get('keypoints'), grid=kw.get('ogrid'), **okw)
    nlev = ora_ex.nlevels
    for l in range(nlev):
        lv = gpu_ex.level(l)
        assert np.array_equal(lv, ora_ex.level(l)), ('pyramid level', l)
        if golden is not None:
            assert sha(lv) == golden['pyr_sha_' + tag][l]
        assert np.array_equal(gpu_ex.level(l, blurred=True), ora_ex.level(l, blurred=True)), ('blurred level', l)
        gx, gy, gs = gpu_ex.raw_corners(l)
        ox, oy, os_ = ora_ex.raw_corners(l)
        assert len(gx) == len(ox), ('raw corner count', l, len(gx), len(ox))
        assert np.array_equal(gx, ox) and np.array_equal(gy, oy) and np.array_equal(gs, os_), ('raw corners + order', l)
        wx, wy, ws = gpu_ex.level_keypoints(l)
        ok = ora_ex.level_keypoints(l)
        assert len(wx) == len(ok), ('quadtree winners', l, len(wx), len(ok))
        assert np.array_equal(wx, ok['x'].astype(np.int32)) and np.array_equal(wy, ok['y'].astype(np.int32)), ('winner order', l)
        assert np.array_equal(ws, ok['response'].astype(np.int32))
    assert len(kps) == len(okps)
    for fld in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
        assert np.array_equal(kps[fld], okps[fld]), fld
    if len(kps):
        assert np.abs(kps['angle'] - okps['angle']).max() <= ANGLE_TOL_DEG
        agree = 1.0 - np.unpackbits(desc ^ odesc).mean()
        assert agree >= DESC_BITS_MIN, agree
    return kps, desc, okps, odesc


def test_extract_cfg1_euroc_frame(gpu, oracle, synth, golden):
    """BASELINE config 1: 752x480, 1000 kp, 8 levels, 1.2, FAST 20/7."""
    img = synth.synth_frame(1, 752, 480)
    ex = gpu.ORBextractor(1000, 1.2, 8, gpu.ORBextractor.FAST_SCORE, 20, max_width=752, max_height=480)
    oex = oracle.Extractor(1000, 1.2, 8, 1, 20)
    kps, desc, okps, odesc = compare_frame(ex, oex, img, golden, '752x480')
    assert len(kps) >= 1000
    # measured expectation: angles and descriptors are bit-identical, not merely within tolerance
    assert np.array_equal(kps['angle'], okps['angle'])
    assert np.array_equal(desc, odesc)
    assert ex.GetLevels() == 8 and ex.GetScaleFactor() == np.float32(1.2)
    sc, inv, quota, umax = ex.tables()
    osc, oinv, oquota, oumax = oex.tables()
    assert np.array_equal(sc, osc) and np.array_equal(inv, oinv) and np.array_equal(quota, oquota) and np.array_equal(umax, oumax)


@pytest.mark.parametrize('seed,W,H,nf,th', [(1000, 640, 512, 1500, 20), (100000, 1280, 1024, 2000, 20), (7, 320, 240, 400, 7),
                                            (8, 968, 608, 800, 10), (9, 401, 307, 600, 1), (10, 640, 480, 1000, 40)])
def test_extract_other_shapes(gpu, oracle, synth, seed, W, H, nf, th):
    img = synth.synth_frame(seed, W, H)
    ex = gpu.ORBextractor(nf, 1.2, 8, 0, th, max_width=W, max_height=H)
    compare_frame(ex, oracle.Extractor(nf, 1.2, 8, 0, th), img)


def test_extract_hard_images(gpu, oracle, synth):
    ex = gpu.ORBextractor(500, 1.2, 8, 1, 20, max_width=400, max_height=300)
    oex = oracle.Extractor(500, 1.2, 8, 1, 20)
    flat = np.full((300, 400), 90, np.uint8)                      # constant image -> 0 keypoints, descriptors released
    kps, desc = ex(flat)
    assert len(kps) == 0 and len(desc) == 0
    noise = (synth.draw(5, np.arange(1, 400 * 300 + 1, dtype=np.uint64)) % np.uint64(256)).astype(np.uint8).reshape(300, 400)
    compare_frame(ex, oex, noise)                                 # white noise: densest possible corner field
    ramp = np.tile(np.arange(400, dtype=np.uint8), (300, 1))
    compare_frame(ex, oex, ramp)
    sparse = np.full((300, 400), 50, np.uint8); sparse[100:140, 120:180] = 58; sparse[200:203, 300:303] = 66
    compare_frame(ex, oex, sparse)                                # only retry-threshold (7) corners exist


def test_extract_empty_image_and_strided_roi(gpu, oracle, synth):
    ex = gpu.ORBextractor(300, 1.2, 8, 1, 20, max_width=400, max_height=300)
    marker = np.zeros(3, gpu.capi.KP_DTYPE)
    out_k, out_d = ex(np.zeros((0, 0), np.uint8), keypoints=marker)
    assert out_k is marker and out_d is None                      # empty image: silent return, outputs untouched
    with pytest.raises(AssertionError):
        ex(np.zeros((10, 10), np.float32))
    big = synth.synth_frame(3, 500, 400)
    roi = big[50:350, 60:460]                                     # image is a ROI of a larger buffer (stride > width)
    k1, d1 = ex(roi)
    k2, d2 = ex(np.ascontiguousarray(roi))
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
    ok, od = oracle.Extractor(300, 1.2, 8, 1, 20)(np.ascontiguousarray(roi))
    assert np.array_equal(k1['x'], ok['x']) and np.array_equal(d1, od)


def test_extract_occupancy_grid_path(gpu, oracle, synth):
    """FullDetect=false: greedy occupancy filter on the caller's column-major grid, incoming keypoints kept
    (src/ORBextractor.cc:872-910, Tracking.cc:901-946)."""
    W, H, mpd = 752, 480, 20
    img = synth.synth_frame(21, W, H)
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=W, max_height=H)
    oex = oracle.Extractor(1000, 1.2, 8, 1, 20)
    for n_in, need in ((0, 400), (25, 150), (5, 1000)):
        grid = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F')
        rng = np.random.default_rng(n_in)
        inc = np.zeros(n_in, gpu.capi.KP_DTYPE)
        inc['x'] = rng.integers(40, W - 40, n_in); inc['y'] = rng.integers(40, H - 40, n_in)
        inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 7
        for k in inc:
            grid[int(k['y'] / mpd), int(k['x'] / mpd)] += 1
        g1 = grid.copy(order='F'); g2 = grid.copy(order='F')
        kps, desc = ex(img, keypoints=inc, grid_2d=g1, min_px_dist=mpd, FullDetect=False, num_featsneeded=need)
        okps, odesc = oex(img, keypoints=inc, grid=g2, min_px_dist=mpd, full_detect=False, num_needed=need)
        assert np.array_equal(g1, g2)
        assert len(kps) == len(okps) and len(kps) >= n_in
        for fld in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
            assert np.array_equal(kps[fld], okps[fld]), fld
        assert np.abs(kps['angle'] - okps['angle']).max() <= ANGLE_TOL_DEG
        assert 1.0 - np.unpackbits(desc ^ odesc).mean() >= DESC_BITS_MIN
    # FullDetect=true leaves the grid untouched and drops incoming keypoints
    g3 = grid.copy(order='F')
    kps, _ = ex(img, keypoints=inc, grid_2d=g3, min_px_dist=mpd, FullDetect=True, num_featsneeded=10)
    assert np.array_equal(g3, grid) and len(kps) >= 1000


def test_extract_match_batch_chains_knn_on_the_device(gpu, oracle, synth):
    """uvip_extract_match_batch_submit: extraction + consecutive-frame kNN2 without the descriptor round trip; chunk sizes 3, 3, 1
    exercise the pair across a chunk boundary (query frame in the other staging set) and a one-frame chunk"""
    nfr, W, H, NF = 7, 640, 512, 800
    frames = synth.synth_batch(2000, nfr, W, H)
    ex = gpu.ORBextractor(NF, 1.2, 8, 1, 20, max_width=W, max_height=H, max_batch=3)
    m = gpu.ORBmatcher(0.75, True)
    cap = NF + 8 * 8 + 64
    for rep in range(2):                                       # the second batch reuses the staging sets and the chunk counter parity
        kps = np.zeros((nfr, cap), gpu.capi.KP_DTYPE); desc = np.zeros((nfr, cap, 32), np.uint8); n = np.zeros(nfr, np.int32)
        ki = np.full((nfr - 1, cap, 2), -7, np.int32); kd = np.full((nfr - 1, cap, 2), -7, np.int32)
        fr = frames if rep == 0 else np.ascontiguousarray(frames[::-1])
        t = ex.extract_match_batch_submit(m, fr, kps, n, desc, ki, kd)
        ex.extract_batch_wait(t)
        okps, on, odesc = oracle.extract_batch(fr, NF, 1.2, 8, 20, cap=cap)
        assert np.array_equal(n, on) and n.min() >= NF
        for f in range(nfr):
            assert np.array_equal(desc[f, :n[f]], odesc[f, :n[f]])
        for f in range(nfr - 1):
            oi, od = oracle.knn2(odesc[f, :n[f]], odesc[f + 1, :n[f + 1]])
            assert np.array_equal(ki[f, :n[f]], oi) and np.array_equal(kd[f, :n[f]], od), (rep, f)


def test_incoming_keypoints_near_the_border(gpu, oracle, synth):
    """incoming level-0 keypoints (ComputeKeyPointsCopy, src/ORBextractor.cc:523-534) may lie anywhere inside the image: their
    IC_Angle disc and descriptor pattern reach into the reference's 16-px reflect-101 border, which the CUDA path materialises
    in full for such calls.  Points at least 2 px inside the image must equal the oracle and the reference's compiled operator()."""
    from oracle import reference as R
    W, H, mpd = 752, 480, 20
    img = synth.synth_frame(33, W, H)
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=W, max_height=H)
    oex = oracle.Extractor(1000, 1.2, 8, 1, 20)
    xs = [2, 3, 7, 15, 16, 40, W - 41, W - 17, W - 16, W - 8, W - 4, W - 3]
    ys = [2, 5, 15, 16, 17, H - 17, H - 16, H - 6, H - 3]
    pts = [(x, y) for x in xs for y in ys]
    inc = np.zeros(len(pts), gpu.capi.KP_DTYPE)
    inc['x'] = [p[0] for p in pts]; inc['y'] = [p[1] for p in pts]
    inc['x'] += 0.25; inc['y'] -= 0.25                                       # cvRound of non-integer coordinates
    inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 3
    grid = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F')
    g1 = grid.copy(order='F'); g2 = grid.copy(order='F')
    kps, desc = ex(img, keypoints=inc, grid_2d=g1, min_px_dist=mpd, FullDetect=False, num_featsneeded=200)
    okps, odesc = oex(img, keypoints=inc, grid=g2, min_px_dist=mpd, full_detect=False, num_needed=200)
    assert np.array_equal(g1, g2) and len(kps) == len(okps) > len(inc)
    n = len(inc)
    assert np.array_equal(kps['x'][:n], inc['x']) and np.array_equal(kps['class_id'][:n], inc['class_id'])
    assert np.array_equal(kps['angle'], okps['angle'])
    assert np.array_equal(desc, odesc)
    if R.available():
        g3 = grid.copy(order='F')
        rk, rd = R.Extractor(1000, 1.2, 8, 1, 20)(img, keypoints=inc, grid=g3, min_px_dist=mpd, full_detect=False, num_needed=200)
        assert np.array_equal(rk['angle'], kps['angle']) and np.array_equal(rd, desc) and np.array_equal(g3, g1)
    # the same extractor goes back to border-free calls and full detection unchanged
    k2, d2 = ex(img)
    ok2, od2 = oex(img)
    assert np.array_equal(d2, od2)


def test_incoming_keypoints_outside_the_image_are_refused(gpu, synth):
    W, H, mpd = 320, 240, 20
    img = synth.synth_frame(5, W, H)
    ex = gpu.ORBextractor(300, 1.2, 8, 1, 20, max_width=W, max_height=H)
    for bad in ((-1.0, 50.0), (W + 0.0, 50.0), (50.0, H + 3.0), (float('nan'), 10.0), (float('inf'), 10.0), (1e12, 5.0)):
        inc = np.zeros(2, gpu.capi.KP_DTYPE)
        inc['x'] = [60.0, bad[0]]; inc['y'] = [60.0, bad[1]]
        grid = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F')
        with pytest.raises(gpu.capi.UvipError) as e:
            ex(img, keypoints=inc, grid_2d=grid, min_px_dist=mpd, FullDetect=False, num_featsneeded=50)
        assert e.value.code == gpu.capi.ERR_ARG
        assert not grid.any()                                                   # a refused call leaves the caller's grid alone
    k, d = ex(img)                                                              # the handle is still healthy (no sticky CUDA error)
    assert len(k) >= 300


def test_single_frame_graph_is_captured_once_per_call_shape(gpu, oracle, synth):
    """src/Tracking.cc:946 changes num_featsneeded (and the grid contents) at every frame: the captured CUDA graph of the
    single-frame call must survive that — one capture per call shape, results still equal to the oracle's"""
    W, H, mpd = 752, 480, 20
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=W, max_height=H)
    oex = oracle.Extractor(1000, 1.2, 8, 1, 20)
    c0 = ex.graph_captures()
    for i, need in enumerate((400, 137, 1000, 3, 250, 999)):
        img = synth.synth_frame(50 + i, W, H)
        g1 = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F'); g1[i::7, ::3] = 1
        g2 = g1.copy(order='F')
        kps, desc = ex(img, grid_2d=g1, min_px_dist=mpd, FullDetect=False, num_featsneeded=need)
        okps, odesc = oex(img, grid=g2, min_px_dist=mpd, full_detect=False, num_needed=need)
        assert np.array_equal(g1, g2) and np.array_equal(kps['x'], okps['x']) and np.array_equal(desc, odesc)
    assert ex.graph_captures() - c0 == 1
    ex(synth.synth_frame(1, W, H))                                             # FullDetect=true is another shape
    ex(synth.synth_frame(2, W, H))
    assert ex.graph_captures() - c0 == 2


def test_extract_batch_matches_single_frames(gpu, oracle, synth):
    """BASELINE config 2 shape: 640x512, 1500 kp, one HBM-resident batch; each frame equals its single-frame result."""
    nfr = 6
    frames = synth.synth_batch(1000, nfr, 640, 512)
    ex = gpu.ORBextractor(1500, 1.2, 8, 1, 20, max_width=640, max_height=512, max_batch=2)   # 6 frames -> 3 pipelined chunks (double-buffered staging)
    kps, n, desc = ex.extract_batch(frames)
    okps, on, odesc = oracle.extract_batch(frames, 1500, 1.2, 8, 20, cap=kps.shape[1])
    assert np.array_equal(n, on) and n.min() >= 1500
    for f in range(nfr):
        a, b = kps[f, :n[f]], okps[f, :n[f]]
        for fld in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
            assert np.array_equal(a[fld], b[fld]), (f, fld)
        assert np.abs(a['angle'] - b['angle']).max() <= ANGLE_TOL_DEG
        assert 1.0 - np.unpackbits(desc[f, :n[f]] ^ odesc[f, :n[f]]).mean() >= DESC_BITS_MIN
    # size-independent property: a batch is order-equivariant
    kps2, n2, desc2 = ex.extract_batch(frames[::-1].copy())
    assert np.array_equal(n2, n[::-1]) and np.array_equal(desc2[0, :n2[0]], desc[nfr - 1, :n[nfr - 1]])
    # submit/wait: two batches in flight give the same bytes as the synchronous call; a third submit and a single-frame
    # call are refused while tickets are outstanding
    cap = kps.shape[1]
    fa, fb = frames[:5].copy(), frames[1:].copy()
    outs = [(np.zeros((5, cap), gpu.KP_DTYPE), np.zeros(5, np.int32), np.zeros((5, cap, 32), np.uint8)) for _ in range(2)]
    ta = ex.extract_batch_submit(fa, *outs[0]); tb = ex.extract_batch_submit(fb, *outs[1])
    assert {ta, tb} == {0, 1}
    with pytest.raises(gpu.UvipError):
        ex.extract_batch_submit(fa, *outs[0])
    with pytest.raises(gpu.UvipError):
        ex(frames[0])
    ex.extract_batch_wait(tb); ex.extract_batch_wait(ta)
    with pytest.raises(gpu.UvipError):
        ex.extract_batch_wait(ta)
    for (k_, n_, d_), lo in ((outs[0], 0), (outs[1], 1)):
        assert np.array_equal(n_, n[lo:lo + 5])
        for f in range(5):
            assert np.array_equal(k_[f, :n_[f]], kps[lo + f, :n_[f]]) and np.array_equal(d_[f, :n_[f]], desc[lo + f, :n_[f]])
    k1, d1 = ex(frames[0])                                   # the single-frame path works again once the tickets are back
    assert np.array_equal(d1, desc[0, :n[0]])


def test_capacity_overflow_is_loud(gpu, synth):
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=752, max_height=480)
    with pytest.raises(gpu.capi.UvipError) as e:
        ex(synth.synth_frame(1, 752, 480), cap=100)              # 1000+ keypoints do not fit 100 rows
    assert e.value.code == gpu.capi.ERR_CAPACITY
    with pytest.raises(gpu.capi.UvipError) as e:
        ex(synth.synth_frame(1, 800, 480))                       # wider than max_width
    assert e.value.code in (gpu.capi.ERR_ARG, gpu.capi.ERR_UNSUPPORTED)
    with pytest.raises(gpu.capi.UvipError):
        gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=120, max_height=100)   # top level would be < 64 px


# ------------------------------------------------------------------------------------------------ matcher
def test_descriptor_distance(gpu, oracle, synth):
    m = gpu.ORBmatcher()
    a = synth.random_descriptors(1, 300); b = synth.random_descriptors(2, 300)
    b[0] = a[0]; b[1] = ~a[1]
    d = m.DescriptorDistance(a, b)
    ref = np.array([oracle.descriptor_distance(x, y) for x, y in zip(a, b)], np.int32)
    assert np.array_equal(d, ref) and d[0] == 0 and d[1] == 256
    assert m.DescriptorDistance(a[5], b[5]) == ref[5]
    assert m.TH_HIGH == 100 and m.TH_LOW == 50 and m.HISTO_LENGTH == 30


def test_knn2_cfg1_frame_pair_and_ratio_and_histogram(gpu, oracle, synth):
    """config 1: brute-force kNN2 between a frame and its shifted twin, ratios {0.6,0.75,0.8,0.9}, rotation histogram."""
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=752, max_height=480)
    ka, da = ex(synth.synth_frame(1, 752, 480))
    kb, db = ex(synth.synth_frame(1, 752, 480, dx=5, dy=3, noise_seed=2))
    m = gpu.ORBmatcher(0.75, True)
    idx, dist = m.knn2(da, db)
    oidx, odist = oracle.knn2(da, db)
    assert np.array_equal(idx, oidx) and np.array_equal(dist, odist)
    for ratio in (0.6, 0.75, 0.8, 0.9):
        got = m.ratio_filter(idx, dist, ratio)
        ref = oracle.ratio_filter(oidx, odist, ratio)
        assert np.array_equal(got, ref)
        g2 = m.rot_hist_filter(got, ka['angle'], kb['angle'])
        r2 = oracle.rot_hist_filter(ref, ka['angle'], kb['angle'])
        assert np.array_equal(g2, r2)
        assert np.array_equal(m.ratioMatching(da, db, ratio, ka['angle'], kb['angle']), r2)
    assert (m.ratio_filter(idx, dist, 0.75) >= 0).sum() > 200      # the twin really matches


def test_knn2_golden_ties_and_ragged_sizes(gpu, oracle, synth, golden):
    m = gpu.ORBmatcher()
    q = synth.random_descriptors(31, 200); t = synth.random_descriptors(32, 2000)
    t[100] = t[7]; t[300] = q[5]; t[301] = q[5]; t[1999] = q[9]
    idx, dist = m.knn2(q, t)
    assert np.array_equal(idx, golden['knn_idx']) and np.array_equal(dist, golden['knn_dist'])
    for nq, nt in ((1, 1), (1, 2), (3, 255), (129, 256), (130, 257), (257, 513), (5, 70000), (1000, 1)):
        qq = synth.random_descriptors(100 + nq, nq); tt = synth.random_descriptors(200 + nt, nt)
        if nt > 600:
            tt[nt - 1] = tt[0]; tt[65536 % nt] = tt[0]          # ties across tiles and across the 65536-row index window
        i1, d1 = m.knn2(qq, tt)
        i2, d2 = oracle.knn2(qq, tt)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2), (nq, nt)
    i, d = m.knn2(q[:4], t[:1])
    assert i[:, 1].tolist() == [-1] * 4 and d[:, 1].tolist() == [257] * 4   # a single train row has no second neighbour
    i, d = m.knn2(q[:0], t)
    assert i.shape == (0, 2)


def test_knn2_shard_invariance_via_device_api(gpu, oracle, synth):
    """cfg4 shape in miniature: database rows sharded contiguously, per-shard top-2 with global indices, merged by
    (distance, index) — identical for every shard count (the multi-GPU path runs the same two entry points)."""
    import ctypes as C
    import torch
    L = gpu.capi.lib()
    T, Q = synth.knn_database(8192, 2048)
    ref_i, ref_d = oracle.knn2(Q, T)
    m = gpu.ORBmatcher()
    dq = torch.from_numpy(Q).cuda(); dt = torch.from_numpy(T).cuda()
    for G in (1, 2, 3, 8):
        bounds = [(g * len(T)) // G for g in range(G + 1)]
        pi = torch.empty((G, len(Q), 2), dtype=torch.int32, device='cuda'); pd = torch.empty_like(pi)
        for g in range(G):
            sh = dt[bounds[g]:bounds[g + 1]]
            gpu.capi.check(L.uvip_knn2_device(m.h, C.c_void_p(dq.data_ptr()), len(Q), C.c_void_p(sh.data_ptr()), sh.shape[0],
                                              bounds[g], C.c_void_p(pi[g].data_ptr()), C.c_void_p(pd[g].data_ptr()), None))
        oi = torch.empty((len(Q), 2), dtype=torch.int32, device='cuda'); od = torch.empty_like(oi)
        gpu.capi.check(L.uvip_knn2_merge_device(m.h, C.c_void_p(pi.data_ptr()), C.c_void_p(pd.data_ptr()), G, len(Q) * 2, len(Q),
                                                C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), None))
        torch.cuda.synchronize()
        assert np.array_equal(oi.cpu().numpy(), ref_i) and np.array_equal(od.cpu().numpy(), ref_d), G


def test_grid_and_search_by_projection_cfg3(gpu, oracle, synth):
    """config 3: 10k projected map points vs a 2000-keypoint frame; claims replayed in the reference's order."""
    c = synth.projection_case()
    m = gpu.ORBmatcher(0.8, True)
    grid = m.grid_build(c['kx'], c['ky'], c['bounds'])
    ostart, oitems = oracle.grid_build(c['kx'], c['ky'], grid['minX'], grid['minY'], grid['inv_w'], grid['inv_h'])
    assert np.array_equal(grid['start'], ostart) and np.array_equal(grid['items'], oitems)
    sf = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=752, max_height=480).tables()[0]
    frame = dict(kx=c['kx'], ky=c['ky'], octave=c['octave'], kdesc=c['kdesc'], grid=grid, scale_factors=sf)
    mps = dict(u=c['u'], v=c['v'], level=c['level'], view_cos=c['view_cos'], desc=c['qdesc'])
    for th in (1.0, 3.0):
        n, match, taken = m.SearchByProjection(frame, mps, th)
        r = np.array([oracle.lib().uo_radius_by_viewing_cos(float(v)) for v in c['view_cos']], np.float32)
        if th != 1.0:
            r = (r * np.float32(th)).astype(np.float32)
        r = (r * sf[c['level']]).astype(np.float32)
        on, omatch, otaken = oracle.search_window(0, 100, 0.8, c['u'], c['v'], r, c['level'] - 1, c['level'], c['qdesc'],
                                                  c['kx'], c['ky'], c['octave'], c['kdesc'], ostart, oitems,
                                                  grid['minX'], grid['minY'], grid['inv_w'], grid['inv_h'])
        assert n == on and n > 500
        assert np.array_equal(match, omatch) and np.array_equal(taken, otaken)
        n2, match2, taken2 = m.search_frame(0, 100, c['u'], c['v'], r, c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'], c['octave'],
                                            c['kdesc'], c['bounds'])                     # grid built on the device by the same call
        assert n2 == on and np.array_equal(match2, omatch) and np.array_equal(taken2, otaken)
    # M5 semantics: th=10 window, levels [l-1, l+1], best only, ORBdist 100, pre-taken keypoints, then the histogram
    lvl = c['level']
    r = (np.float32(10) * sf[lvl]).astype(np.float32)
    pre = np.full(len(c['kx']), -1, np.int32); pre[::7] = -2
    n, match, taken = m.search_window(1, 100, c['u'], c['v'], r, lvl - 1, lvl + 1, c['qdesc'], c['kx'], c['ky'], c['octave'],
                                      c['kdesc'], grid, taken=pre)
    on, omatch, otaken = oracle.search_window(1, 100, 0.8, c['u'], c['v'], r, lvl - 1, lvl + 1, c['qdesc'], c['kx'], c['ky'],
                                              c['octave'], c['kdesc'], ostart, oitems, grid['minX'], grid['minY'],
                                              grid['inv_w'], grid['inv_h'], taken=pre)
    assert n == on and np.array_equal(match, omatch) and np.array_equal(taken, otaken)
    assert (match[match >= 0] % 7 != 0).all()                     # pre-taken keypoints are never claimed
    g = m.rot_hist_filter(match, c['qangle'], c['kangle'])
    o = oracle.rot_hist_filter(omatch, c['qangle'], c['kangle'])
    assert np.array_equal(g, o) and (g >= 0).sum() < (match >= 0).sum()
    # M8 Fuse semantics (src/ORBmatcher.cc:1075-1100): th=2.5 window, levels [l-1, l], best only, TH_LOW, NO claims: taken[] is
    # neither read nor written and several map points may land on the same keypoint
    r = (np.float32(2.5) * sf[lvl]).astype(np.float32)
    n4, match4, taken4 = m.search_window(4, 50, c['u'], c['v'], r, lvl - 1, lvl, c['qdesc'], c['kx'], c['ky'], c['octave'],
                                         c['kdesc'], grid, taken=pre)
    on4, omatch4, otaken4 = oracle.search_window(4, 50, 0.8, c['u'], c['v'], r, lvl - 1, lvl, c['qdesc'], c['kx'], c['ky'],
                                                 c['octave'], c['kdesc'], ostart, oitems, grid['minX'], grid['minY'],
                                                 grid['inv_w'], grid['inv_h'], taken=pre)
    assert n4 == on4 == int((match4 >= 0).sum()) and np.array_equal(match4, omatch4)
    assert np.array_equal(taken4, pre) and np.array_equal(otaken4, pre)
    hit = match4[match4 >= 0]
    assert len(np.unique(hit)) < len(hit) and (hit % 7 == 0).any()   # shared keypoints, pre-taken ones included


def test_search_window_claim_chain(gpu, oracle, synth):
    """adversarial claims: every query wants the same few keypoints, so results depend on the sequential order."""
    nk, nq = 40, 300
    kx = (100 + (np.arange(nk) % 8) * 0.5).astype(np.float32); ky = (100 + (np.arange(nk) // 8) * 0.5).astype(np.float32)
    octave = np.zeros(nk, np.int32)
    kdesc = synth.random_descriptors(77, nk)
    qdesc = synth.flip_bits(np.repeat(kdesc[:1], nq, 0), 900, (np.arange(nq) % 5).tolist())
    m = gpu.ORBmatcher(0.99, True)
    grid = m.grid_build(kx, ky, (0, 752, 0, 480))
    qu = np.full(nq, 101, np.float32); qv = np.full(nq, 101, np.float32); qr = np.full(nq, 6, np.float32)
    ml = np.full(nq, -1, np.int32)
    for mode, th in ((0, 256), (1, 256), (0, 100)):
        n, match, taken = m.search_window(mode, th, qu, qv, qr, ml, ml, qdesc, kx, ky, octave, kdesc, grid)
        on, omatch, otaken = oracle.search_window(mode, th, 0.99, qu, qv, qr, ml, ml, qdesc, kx, ky, octave, kdesc,
                                                  grid['start'], grid['items'], grid['minX'], grid['minY'], grid['inv_w'], grid['inv_h'])
        assert n == on and np.array_equal(match, omatch) and np.array_equal(taken, otaken), mode
    assert n <= nk


def test_search_frame_candidate_cache_equals_sequential_scan(gpu, oracle, synth):
    """uvip_search_frame replays per-query candidate lists in the claim rounds after the first (csrc/matcher.cu, SearchCtx::qcache).
    Cases the list must get right: queries with 0, a few and more than six candidates (the last are searched in full every round),
    octaves beyond what an entry holds (> 15), pre-taken keypoints, and heavy competition for the same keypoints (many rounds);
    with and without level gates, best-only and top-2 modes.  The switch UVIP_SEARCH_NOCACHE must not change a single entry."""
    rng = np.random.RandomState(5)
    nk, nq = 400, 1500
    kx = (60 + rng.rand(nk) * 120).astype(np.float32); ky = (60 + rng.rand(nk) * 90).astype(np.float32)
    kdesc = synth.random_descriptors(91, nk)
    src = rng.randint(0, nk, nq)
    qdesc = synth.flip_bits(kdesc[src], 901, rng.randint(0, 40, nq).tolist())
    qu = (kx[src] + rng.randn(nq) * 2).astype(np.float32); qv = (ky[src] + rng.randn(nq) * 2).astype(np.float32)
    qr = np.where(rng.rand(nq) < 0.5, 2.5, np.where(rng.rand(nq) < 0.5, 9.0, 30.0)).astype(np.float32)     # few / several / dozens of candidates
    qu[::50] += 4000                                                                                        # windows outside the grid
    bounds = (0, 752, 0, 480)
    m = gpu.ORBmatcher(0.9, True)
    pre = np.full(nk, -1, np.int32); pre[::11] = 12345
    for octmax in (8, 21):
        octave = rng.randint(0, octmax, nk).astype(np.int32)
        start, items = oracle.grid_build(kx, ky, 0.0, 0.0, float(np.float32(64) / np.float32(752)), float(np.float32(48) / np.float32(480)))
        for gated in (False, True):
            lo = (octave[src] - 1).astype(np.int32) if gated else np.full(nq, -1, np.int32)
            hi = (octave[src] + 1).astype(np.int32) if gated else np.full(nq, -1, np.int32)
            for mode, th in ((0, 256), (0, 60), (1, 256), (6, 256), (4, 100)):
                # (mode 6, the level-free top-2 of WindowSearch, has no oracle restatement: it is pinned through the shim against the
                #  reference's compiled functions; here the cached and the plain search must agree)
                want = None if mode == 6 else oracle.search_window(mode, th, 0.9, qu, qv, qr, lo, hi, qdesc, kx, ky, octave, kdesc, start, items, 0.0, 0.0,
                                                                   float(np.float32(64) / np.float32(752)), float(np.float32(48) / np.float32(480)), taken=pre)
                got = m.search_frame(mode, th, qu, qv, qr, lo, hi, qdesc, kx, ky, octave, kdesc, bounds, taken=pre)
                os.environ['UVIP_SEARCH_NOCACHE'] = '1'
                try:
                    plain = m.search_frame(mode, th, qu, qv, qr, lo, hi, qdesc, kx, ky, octave, kdesc, bounds, taken=pre)
                finally:
                    os.environ.pop('UVIP_SEARCH_NOCACHE', None)
                for a, b, c in zip(got, want or got, plain):
                    assert np.array_equal(a, b) and np.array_equal(a, c), (octmax, gated, mode, th)
                assert got[0] > 50


def test_search_by_bow_node_restricted_lists(gpu, oracle, synth):
    """M6: SearchByBoW inner loops (src/ORBmatcher.cc:186-245, :751-811) over explicit candidate lists: top-2 inside a
    vocabulary node, TH_LOW, ratio, sequential claims, rotation histogram."""
    rng = np.random.default_rng(11)
    n1, n2 = 1500, 1700
    d1 = synth.random_descriptors(61, n1)
    src = rng.integers(0, n1, n2)
    d2 = synth.flip_bits(d1[src], 700, rng.integers(0, 40, n2).tolist())
    fresh = rng.random(n2) < 0.3
    d2[fresh] = synth.random_descriptors(62, n2)[fresh]
    nodes1 = rng.integers(0, 60, n1)
    nodes2 = np.where(rng.random(n2) < 0.9, nodes1[src], rng.integers(0, 60, n2))
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    a2 = np.mod(a1[src] + rng.normal(0, 4, n2), 360).astype(np.float32)
    m = gpu.ORBmatcher(0.75, True)
    # flat lists in the reference's node-major order
    order = np.lexsort((np.arange(n1), nodes1))
    lists = [np.nonzero(nodes2 == nodes1[i])[0] for i in order]
    cs = np.zeros(n1 + 1, np.int32); cs[1:] = np.cumsum([len(l) for l in lists]); ci = np.concatenate(lists).astype(np.int32)
    pre = np.full(n2, -1, np.int32); pre[::9] = -2
    for mode, th in ((2, 50), (3, 50), (1, 50), (2, 256)):
        n, match, taken = m.search_lists(mode, th, d1[order], cs, ci, d2, taken=pre)
        on, omatch, otaken = oracle.search_lists(mode, th, np.float32(0.75), d1[order], cs, ci, d2, taken=pre)
        assert n == on and np.array_equal(match, omatch) and np.array_equal(taken, otaken), mode
        assert (match[match >= 0] % 9 != 0).all()
    assert n > 300
    full = m.SearchByBoW(d1, nodes1, a1, d2, nodes2, a2)
    on, omatch, _ = oracle.search_lists(2, 50, np.float32(0.75), d1[order], cs, ci, d2)
    ofull = np.full(n1, -1, np.int32); ofull[order] = omatch
    assert np.array_equal(full, oracle.rot_hist_filter(ofull, a1, a2))
    assert 0 < (full >= 0).sum() <= (ofull >= 0).sum()


@pytest.mark.parametrize('seed,trials,wmax,hmax,sfs', [(2024, 24, 900, 700, (1.1, 1.2, 1.2, 1.3, 1.5)),
                                                      (77, 40, 1500, 1100, (1.1, 1.15, 1.2, 1.25, 1.3, 1.4, 1.5, 1.6, 1.7))])
def test_extract_parameter_sweep(gpu, oracle, synth, seed, trials, wmax, hmax, sfs):
    """randomised sweep over shapes, level counts, scale factors, quotas and thresholds: keypoints (position, order,
    response, octave) and descriptors must equal the oracle's for every draw (the second set reaches widths whose last FAST /
    resize tile is a sliver and scale factors that need the three-word resize window)"""
    rng = np.random.default_rng(seed)
    for trial in range(trials):
        W = int(rng.integers(160, wmax)); H = int(rng.integers(140, hmax))
        if W < 0.55 * H:
            continue
        nlev = int(rng.integers(1, 9)); sf = float(rng.choice(sfs))
        while min(W, H) / sf ** (nlev - 1) < 70:
            nlev -= 1
        nf = int(rng.integers(50, 2500)); th = int(rng.choice([5, 7, 12, 20, 20, 30]))
        if nlev == 1 and nf > 1900:
            nf = 1900                                             # one level holding the whole quota: quadtree node capacity
        img = synth.synth_frame(int(rng.integers(1, 1 << 30)), W, H)
        if trial % 5 == 4:
            img = (img // 4 + 90).astype(np.uint8)                # low contrast: many cells fall back to the retry threshold
        ex = gpu.ORBextractor(nf, sf, nlev, 1, th, max_width=W, max_height=H)
        oex = oracle.Extractor(nf, sf, nlev, 1, th)
        kps, desc = ex(img, cap=nf + 8 * nlev + 4096)
        okps, odesc = oex(img)
        tag = (trial, W, H, nlev, sf, nf, th)
        assert len(kps) == len(okps), tag
        for fld in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
            assert np.array_equal(kps[fld], okps[fld]), (tag, fld)
        if len(kps):
            assert np.abs(kps['angle'] - okps['angle']).max() <= ANGLE_TOL_DEG, tag
            assert 1.0 - np.unpackbits(desc ^ odesc).mean() >= DESC_BITS_MIN, tag
        ex.close()


@pytest.mark.parametrize('sf,nlev,W,H', [(1.7, 4, 640, 480), (1.75, 3, 752, 480), (1.6, 4, 514, 402), (1.25, 8, 1026, 770)])
def test_extract_wide_scale_factors(gpu, oracle, synth, sf, nlev, W, H):
    """scale factors above 1.5 take the resize kernel's three-word byte selection for every column (k_resize<false>), and widths
    such as 514 / 1026 leave a last tile narrower than the mirrored border columns (ring pixels owned by the tile before it):
    pyramid, blurred planes, corners, winners, keypoints and descriptors must still equal the oracle's"""
    img = synth.synth_frame(31 + nlev, W, H)
    ex = gpu.ORBextractor(800, sf, nlev, 1, 20, max_width=W, max_height=H)
    compare_frame(ex, oracle.Extractor(800, sf, nlev, 1, 20), img)
    ex.close()


def test_extract_scale_factor_beyond_the_resize_window_is_loud(gpu, synth):
    """a scale factor whose source box or 4-column source span leaves the kernel's envelope is refused, not approximated"""
    img = synth.synth_frame(5, 640, 480)
    for sf in (2.0, 2.6):                                   # 2.0: source box of a 128-column tile wider than a TMA box; 2.6: window
        with pytest.raises(gpu.capi.UvipError) as e:
            ex = gpu.ORBextractor(500, sf, 2, 1, 20, max_width=640, max_height=480)
            ex(img, cap=8192)
        assert e.value.code == gpu.capi.ERR_UNSUPPORTED


def test_clahe_preprocessing(gpu, oracle, synth, golden):
    """next row N3: CLAHE (clip 4, 12x12 tiles) bit-exact against the oracle and the cv2 golden hashes, then the full
    Enhance -> extract chain of Tracking::GrabImage (src/Tracking.cc:425-446)"""
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=1280, max_height=1024)
    for seed, W, H in ((1, 752, 480), (1000, 640, 512), (100000, 1280, 1024)):
        img = synth.synth_frame(seed, W, H)
        out = ex.clahe(img, 4.0, (12, 12))
        assert sha(out) == str(golden['clahe_sha_%dx%d' % (W, H)])
        assert np.array_equal(out, oracle.clahe(img, 4.0, (12, 12)))
    assert np.array_equal(ex.clahe(synth.synth_frame(17, 97, 61), 2.0, (4, 3)), golden['clahe_small'])
    for clip, tiles in ((2.0, (8, 8)), (40.0, (12, 12)), (0.0, (5, 7))):
        img = synth.synth_frame(9, 401, 307)
        assert np.array_equal(ex.clahe(img, clip, tiles), oracle.clahe(img, clip, tiles)), (clip, tiles)
    img = synth.synth_frame(1, 752, 480)
    kps, desc = ex(ex.clahe(img))
    okps, odesc = oracle.Extractor(1000, 1.2, 8, 1, 20)(oracle.clahe(img))
    assert np.array_equal(kps['x'], okps['x']) and np.array_equal(kps['y'], okps['y']) and np.array_equal(desc, odesc)


def _distinctive_case(synth, seed=5):
    """ragged map-point observation lists: N = 0, 1, 2, 3, ... incl. > 32 and > 64, clustered descriptors + exact duplicates"""
    rng = np.random.default_rng(seed)
    sizes = [0, 1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 40, 64, 65, 100, 2, 6] + list(rng.integers(1, 20, 200))
    start = np.zeros(len(sizes) + 1, np.int32); start[1:] = np.cumsum(sizes)
    base = synth.random_descriptors(seed, len(sizes))
    rows = []
    for p, n in enumerate(sizes):
        if n == 0:
            continue
        d = np.repeat(base[p:p + 1], n, 0)
        d = synth.flip_bits(d, seed * 1000 + p, [int(v) for v in rng.integers(0, 60, n)])
        if n >= 4:
            d[n - 1] = d[0]                                  # exact duplicates -> zero distances and median ties
        rows.append(d)
    return np.concatenate(rows), start


def test_distinctive_descriptors_batch(gpu, oracle, synth):
    """next row N4 (descriptor half): MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:197-270; bit-exact"""
    desc, start = _distinctive_case(synth)
    m = gpu.ORBmatcher(0.6, True)
    bi, bm = m.distinctive_descriptors(desc, start)
    obi, obm = oracle.distinctive_descriptors(desc, start)
    assert np.array_equal(bi, obi) and np.array_equal(bm, obm)
    assert bi[0] == -1 and bi[1] == 0 and bm[1] == 0
    # size-independent property: the least median does not depend on the order of a point's observations
    p = 15                                                   # N = 100
    sl = slice(start[p], start[p + 1])
    perm = np.random.default_rng(1).permutation(100)
    d2 = desc.copy(); d2[sl] = desc[sl][perm]
    bi2, bm2 = m.distinctive_descriptors(d2, start)
    assert bm2[p] == bm[p] and np.array_equal(np.delete(bm2, p), np.delete(bm, p))


def test_search_by_projection_batch_device(gpu, oracle, synth):
    """config 3 as a batch: several frames (ragged sizes) searched by one launch, each identical to the sequential oracle."""
    import torch
    dev = torch.device('cuda', 0)
    W, H = 752, 480
    cases = [synth.projection_case(seed_f=3 + 10 * i, seed_p=4 + 10 * i, nk=2000 - 150 * i, nq=10000 - 900 * i) for i in range(5)]
    m = gpu.ORBmatcher(0.8, True)
    sf = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=W, max_height=H).tables()[0]
    QS, KS, F = 10000, 2000, len(cases)
    qf = {k: np.zeros((F, QS), np.float32) for k in ('u', 'v', 'r')}; qi = {k: np.zeros((F, QS), np.int32) for k in ('lo', 'hi')}
    qd = np.zeros((F, QS, 32), np.uint8)
    kf = {k: np.zeros((F, KS), np.float32) for k in ('x', 'y')}; ko = np.zeros((F, KS), np.int32); kd = np.zeros((F, KS, 32), np.uint8)
    nq = np.zeros(F, np.int32); nk = np.zeros(F, np.int32)
    taken = np.full((F, KS), -1, np.int32)
    for f, c in enumerate(cases):
        a, b = len(c['u']), len(c['kx'])
        nq[f], nk[f] = a, b
        qf['u'][f, :a] = c['u']; qf['v'][f, :a] = c['v']; qf['r'][f, :a] = m.projection_radius(c['view_cos'], c['level'], sf, 1.0)
        qi['lo'][f, :a] = c['level'] - 1; qi['hi'][f, :a] = c['level']; qd[f, :a] = c['qdesc']
        kf['x'][f, :b] = c['kx']; kf['y'][f, :b] = c['ky']; ko[f, :b] = c['octave']; kd[f, :b] = c['kdesc']
        taken[f, :b:9] = -2
    T = lambda a: torch.from_numpy(a).to(dev)
    t = dict(u=T(qf['u']), v=T(qf['v']), r=T(qf['r']), lo=T(qi['lo']), hi=T(qi['hi']), qd=T(qd), x=T(kf['x']), y=T(kf['y']), o=T(ko), kd=T(kd),
             nq=T(nq), nk=T(nk), taken=T(taken), match=torch.full((F, QS), -7, dtype=torch.int32, device=dev),
             counts=torch.zeros((F, 2), dtype=torch.int32, device=dev))
    P = lambda k: t[k].data_ptr()
    m.search_window_batch_device(0, 100, (0, W, 0, H), F, (P('u'), P('v'), P('r'), P('lo'), P('hi'), P('qd')), P('nq'), QS,
                                 (P('x'), P('y'), P('o'), P('kd')), P('nk'), KS, P('taken'), P('match'), P('counts'))
    torch.cuda.synchronize()
    g_match = t['match'].cpu().numpy(); g_taken = t['taken'].cpu().numpy(); g_counts = t['counts'].cpu().numpy()
    inv_w = np.float32(64.0) / np.float32(W); inv_h = np.float32(48.0) / np.float32(H)
    for f, c in enumerate(cases):
        a, b = nq[f], nk[f]
        start, items = oracle.grid_build(c['kx'], c['ky'], 0.0, 0.0, float(inv_w), float(inv_h))
        on, om, otk = oracle.search_window(0, 100, 0.8, c['u'], c['v'], qf['r'][f, :a], c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'],
                                           c['octave'], c['kdesc'], start, items, 0.0, 0.0, float(inv_w), float(inv_h), taken=taken[f, :b])
        assert g_counts[f, 0] == on and on > 500, f
        assert np.array_equal(g_match[f, :a], om) and np.array_equal(g_taken[f, :b], otk), f
        assert (g_match[f, a:] == -7).all()                       # nothing written beyond the frame's own queries


def test_new_entry_points_edge_cases(gpu, oracle, synth):
    """empty and degenerate inputs of the entry points added late in round 1: nothing crashes, nothing is written out of bounds,
    results equal the oracle's on the same degenerate input"""
    import ctypes as C
    L = gpu.capi.lib()
    m = gpu.ORBmatcher(0.8, True)
    c = synth.projection_case(nk=50, nq=80)
    r = m.projection_radius(c['view_cos'], c['level'], [1.0, 1.2, 1.44, 1.728, 2.0736, 2.48832, 2.985984, 3.5831808], 1.0)
    # uvip_search_frame: no queries, no keypoints, all keypoints pre-taken, queries far outside the image
    n, match, taken = m.search_frame(0, 100, [], [], [], [], [], np.zeros((0, 32), np.uint8), c['kx'], c['ky'], c['octave'], c['kdesc'], c['bounds'])
    assert n == 0 and len(match) == 0 and (taken == -1).all()
    n, match, taken = m.search_frame(0, 100, c['u'], c['v'], r, c['level'] - 1, c['level'], c['qdesc'], [], [], [], np.zeros((0, 32), np.uint8), c['bounds'])
    assert n == 0 and (match == -1).all()
    n, match, taken = m.search_frame(0, 100, c['u'], c['v'], r, c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'], c['octave'], c['kdesc'], c['bounds'],
                                     taken=np.full(50, -2, np.int32))
    assert n == 0 and (match == -1).all() and (taken == -2).all()
    n, match, taken = m.search_frame(1, 256, c['u'] + 5000, c['v'] - 5000, r, c['level'] - 1, c['level'], c['qdesc'], c['kx'], c['ky'], c['octave'], c['kdesc'],
                                     c['bounds'])
    assert n == 0 and (match == -1).all()
    # one query, one keypoint, identical position: found in every mode
    for mode in (0, 1, 4):
        n, match, taken = m.search_frame(mode, 256, c['kx'][:1], c['ky'][:1], [3.0], [-1], [-1], c['kdesc'][:1], c['kx'][:1], c['ky'][:1], c['octave'][:1], c['kdesc'][:1],
                                         c['bounds'])
        assert n == 1 and match[0] == 0 and taken[0] == (0 if mode != 4 else -1)
    # epipolar list search: empty lists, a degenerate line (den == 0), threshold 0
    qd = c['qdesc'][:4]; kd = c['kdesc'][:6]
    kthr = np.full(6, 3.84, np.float64)
    cs = np.array([0, 0, 2, 4, 6], np.int32); ci = np.array([0, 1, 2, 3, 4, 5], np.int32)
    ql = np.array([[0, 1, -c['ky'][0], 1], [0, 0, 0, 0], [0, 1, -c['ky'][2], 1], [1, 0, -c['kx'][4], 1]], np.float32)
    for th in (256, 0):
        on, om, otk = oracle.search_lists_epipolar(th, qd, ql, cs, ci, kd, c['kx'][:6], c['ky'][:6], kthr)
        match = np.full(4, -7, np.int32); taken = np.full(6, -1, np.int32); nn = C.c_int(-1)
        gpu.capi.check(L.uvip_search_lists_epipolar(m.h, th, qd.ctypes.data, ql.ctypes.data, 4, cs.ctypes.data, ci.ctypes.data, kd.ctypes.data,
                                                    np.ascontiguousarray(c['kx'][:6]).ctypes.data, np.ascontiguousarray(c['ky'][:6]).ctypes.data, kthr.ctypes.data, 6,
                                                    taken.ctypes.data, match.ctypes.data, C.byref(nn)))
        assert nn.value == on and np.array_equal(match, om) and np.array_equal(taken, otk), th
    assert om[0] == -1 and om[1] == -1                     # empty list, degenerate line
    # haloc hash: an empty set among non-empty ones; match against an empty table
    proj = (np.arange(3 * 64, dtype=np.float32).reshape(3, 64) % 7 - 3) / np.float32(8)
    start = np.array([0, 0, 5, 5, 9], np.int32)
    h = m.haloc_hash(c['kdesc'][:9], start, proj)
    assert (h[0] == 0).all() and (h[2] == 0).all()
    assert np.array_equal(h[1], oracle.haloc_hash(c['kdesc'][:5], proj)) and np.array_equal(h[3], oracle.haloc_hash(c['kdesc'][5:9], proj))
    assert len(m.haloc_match(h[1], np.zeros((0, 96), np.float32))) == 0
    with pytest.raises(gpu.capi.UvipError):               # more rows than projection entries: loud, not truncated
        m.haloc_hash(c['kdesc'][:9], np.array([0, 9], np.int32), proj[:, :4])
    # Harris: a point too close to the level's edge is refused, not read out of bounds
    ex = gpu.ORBextractor(300, 1.2, 4, 0, 20, max_width=320, max_height=240)
    ex(synth.synth_frame(2, 320, 240))
    assert len(ex.harris_responses(1, [], [])) == 0
    with pytest.raises(gpu.capi.UvipError):
        ex.harris_responses(0, [-1.0], [100.0])               # the 4-px reflect-101 ring covers x >= 0, not x = -1
    assert np.isfinite(ex.harris_responses(3, [20.0], [20.0])).all()
    # batched search with zero frames
    m.search_window_batch_device(0, 100, (0, 752, 0, 480), 0, (0, 0, 0, 0, 0, 0), 0, 1, (0, 0, 0, 0), 0, 1, 0, 0, 0)


def test_device_frame_generator_equals_numpy(gpu, synth):
    """uvip_synth_frames_device (SURVEY Appendix B synth_frame on the device, used to produce BASELINE config 5 on the GPU box) gives
    the bytes of the numpy generator, shifts and separate noise seeds included"""
    import ctypes as C
    import torch
    L = gpu.capi.lib()
    for (W, H), cases in (((752, 480), [(1, 0, 0, 1), (1, 5, 3, 2), (77, -4, 9, 77)]), ((1280, 1024), [(300007, 21, 14, 300007)]), ((321, 243), [(9, 28, 22, 5)])):
        seeds = torch.tensor([c[0] for c in cases], dtype=torch.int64, device='cuda')
        dxy = torch.tensor([[c[1], c[2]] for c in cases], dtype=torch.int32, device='cuda')
        nse = torch.tensor([c[3] for c in cases], dtype=torch.int64, device='cuda')
        out = torch.zeros((len(cases), H, W), dtype=torch.uint8, device='cuda')
        gpu.capi.check(L.uvip_synth_frames_device(C.c_void_p(seeds.data_ptr()), C.c_void_p(dxy.data_ptr()), C.c_void_p(nse.data_ptr()), len(cases), W, H,
                                                  C.c_void_p(out.data_ptr()), W * H, None))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for k, (seed, dx, dy, ns) in enumerate(cases):
            assert np.array_equal(got[k], synth.synth_frame(seed, W, H, dx=dx, dy=dy, noise_seed=ns)), (W, H, seed, dx, dy, ns)
