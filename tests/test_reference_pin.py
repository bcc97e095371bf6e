"""Parity pin against the reference's OWN code.

oracle/_ref/libref_orbextractor.so is USLAM::ORBextractor compiled from /root/reference/src/ORBextractor.cc where it
lies (`make -C oracle ref`), against stand-in OpenCV/Eigen/ROS headers (oracle/ref_shim/): cell geometry, empty-cell
retry, DistributeOctTree/DivideNode, IC_Angle, computeOrbDescriptor, operator() orchestration and the occupancy-grid
filter are the reference's compiled code; the OpenCV primitives behind the stand-ins are the C oracle's restatements,
themselves pinned against cv2 (tests/test_oracle_golden.py).

CPU tests: the C oracle must reproduce the reference bit for bit (keypoints incl. order, angles, descriptors, grid).
GPU tests (-m gpu): the CUDA path through the C-ABI against the reference directly.  The library is prebuilt in the
build container and travels to the GPU box; nothing here reads /root/reference at run time."""
import numpy as np
import pytest

from oracle import reference as R

needs_ref = pytest.mark.skipif(not R.available(), reason='oracle/_ref not built and /root/reference absent')

ANGLE_TOL_DEG = 1e-3 * 180.0 / np.pi
FIELDS = ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')


def assert_same(a, b, what=''):
    (ak, ad), (bk, bd) = a, b
    assert len(ak) == len(bk), (what, len(ak), len(bk))
    for f in FIELDS:
        assert np.array_equal(ak[f], bk[f]), (what, f)
    assert np.array_equal(ad, bd), (what, 'descriptors')


@needs_ref
def test_reference_builds_and_reports_its_tables(oracle):
    ex = R.Extractor(1000, 1.2, 8, 1, 20)
    assert ex.GetLevels() == 8 and ex.GetScaleFactor() == np.float32(1.2)


@needs_ref
@pytest.mark.parametrize('seed,W,H,nf,th', [(1, 752, 480, 1000, 20), (1000, 640, 512, 1500, 20), (100000, 1280, 1024, 2000, 20),
                                            (7, 320, 240, 400, 7), (8, 968, 608, 800, 10), (9, 401, 307, 600, 1),
                                            (10, 640, 480, 1000, 40)])
def test_oracle_equals_reference_on_benchmark_shapes(oracle, synth, seed, W, H, nf, th):
    img = synth.synth_frame(seed, W, H)
    assert_same(oracle.Extractor(nf, 1.2, 8, 1, th)(img), R.Extractor(nf, 1.2, 8, 1, th)(img), (seed, W, H))


@needs_ref
def test_oracle_equals_reference_parameter_sweep(oracle, synth):
    rng = np.random.default_rng(2026)
    for i in range(16):
        W, H = int(rng.integers(200, 800)), int(rng.integers(160, 600))
        nf = int(rng.integers(100, 2500)); sf = float(rng.choice([1.1, 1.2, 1.25, 1.5, 2.0])); nl = int(rng.integers(1, 9))
        th = int(rng.choice([5, 7, 12, 20, 30, 60]))
        while min(W, H) / sf ** (nl - 1) < 70:          # keep every level larger than the 2 x 16 px border + one cell
            nl -= 1
        img = synth.synth_frame(500 + i, W, H)
        assert_same(oracle.Extractor(nf, sf, nl, i & 1, th)(img), R.Extractor(nf, sf, nl, i & 1, th)(img), (i, W, H, nf, sf, nl, th))


@needs_ref
def test_oracle_equals_reference_on_hard_images(oracle, synth):
    oex, rex = oracle.Extractor(500, 1.2, 8, 1, 20), R.Extractor(500, 1.2, 8, 1, 20)
    flat = np.full((300, 400), 90, np.uint8)
    noise = (synth.draw(5, np.arange(1, 400 * 300 + 1, dtype=np.uint64)) % np.uint64(256)).astype(np.uint8).reshape(300, 400)
    ramp = np.tile(np.arange(400, dtype=np.uint8), (300, 1))
    sparse = np.full((300, 400), 50, np.uint8); sparse[100:140, 120:180] = 58; sparse[200:203, 300:303] = 66
    for name, img in (('flat', flat), ('noise', noise), ('ramp', ramp), ('sparse', sparse)):
        assert_same(oex(img), rex(img), name)
    assert len(rex(flat)[0]) == 0


@needs_ref
def test_oracle_equals_reference_occupancy_grid_sequence(oracle, synth):
    """FullDetect=false over a short sequence that carries the caller's grid from call to call, with incoming keypoints
    (src/ORBextractor.cc:863,872-910; the call site is src/Tracking.cc:901-946)."""
    W, H, mpd = 752, 480, 20
    oex, rex = oracle.Extractor(1000, 1.2, 8, 1, 20), R.Extractor(1000, 1.2, 8, 1, 20)
    g_o = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F'); g_r = g_o.copy(order='F')
    rng = np.random.default_rng(5)
    for step, (n_in, need) in enumerate(((0, 400), (25, 150), (5, 1000), (60, 37))):
        img = synth.synth_frame(21 + step, W, H)
        inc = np.zeros(n_in, oracle.KP_DTYPE)
        inc['x'] = rng.integers(40, W - 40, n_in); inc['y'] = rng.integers(40, H - 40, n_in)
        inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 7
        a = oex(img, keypoints=inc, grid=g_o, min_px_dist=mpd, full_detect=False, num_needed=need)
        b = rex(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=False, num_needed=need)
        assert_same(a, b, step)
        assert np.array_equal(g_o, g_r), step
        assert len(a[0]) >= n_in
    # FullDetect=true drops the incoming keypoints and leaves the grid alone (:911-913)
    before = g_r.copy(order='F')
    a = oex(img, keypoints=inc, grid=g_o, min_px_dist=mpd, full_detect=True, num_needed=10)
    b = rex(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=True, num_needed=10)
    assert_same(a, b, 'full')
    assert np.array_equal(g_r, before)


@needs_ref
def test_oracle_equals_reference_incoming_keypoints_near_border(oracle, synth):
    """incoming level-0 keypoints 2..15 px from the image border: IC_Angle and the descriptor pattern read the 16-px reflect-101
    border of the level-0 buffer (src/ORBextractor.cc:523-534, :996-997); oracle and compiled reference agree bit for bit"""
    W, H, mpd = 752, 480, 20
    img = synth.synth_frame(33, W, H)
    pts = [(x, y) for x in (2, 3, 7, 15, 16, 40, W - 41, W - 17, W - 16, W - 8, W - 4, W - 3) for y in (2, 5, 15, 16, 17, H - 17, H - 16, H - 6, H - 3)]
    inc = np.zeros(len(pts), oracle.KP_DTYPE)
    inc['x'] = [p[0] for p in pts]; inc['y'] = [p[1] for p in pts]
    inc['x'] += 0.25; inc['y'] -= 0.25
    inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 3
    g_o = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F'); g_r = g_o.copy(order='F')
    a = oracle.Extractor(1000, 1.2, 8, 1, 20)(img, keypoints=inc, grid=g_o, min_px_dist=mpd, full_detect=False, num_needed=200)
    b = R.Extractor(1000, 1.2, 8, 1, 20)(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=False, num_needed=200)
    assert_same(a, b, 'border')
    assert np.array_equal(g_o, g_r) and len(a[0]) > len(inc)


@needs_ref
def test_reference_pointer_tiebreak_is_the_only_freedom(oracle, synth):
    """With plain malloc addresses the reference's (size, node pointer) sort (src/ORBextractor.cc:1151) may order
    equal-size nodes differently; the result may then differ from the pinned one only in a few keypoints per level."""
    img = synth.synth_frame(1, 752, 480)
    pk, _ = R.Extractor(1000, 1.2, 8, 1, 20)(img)
    mk, _ = R.Extractor(1000, 1.2, 8, 1, 20)(img, arena_mb=-1)
    a = set(zip(pk['x'], pk['y'], pk['octave'])); b = set(zip(mk['x'], mk['y'], mk['octave']))
    assert len(a - b) <= 0.05 * len(a)
    assert abs(len(pk) - len(mk)) <= 8


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope='module')
def gpu(pkg):
    if pkg.capi.lib().uvip_device_count() < 1:
        pytest.fail('no CUDA device: the gpu-marked tests must run on the B200 box')
    return pkg


def assert_cuda_matches(gk, gd, rk, rd, what=''):
    assert len(gk) == len(rk), (what, len(gk), len(rk))
    for f in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
        assert np.array_equal(gk[f], rk[f]), (what, f)
    if len(gk):
        assert np.abs(gk['angle'] - rk['angle']).max() <= ANGLE_TOL_DEG, what
        assert 1.0 - np.unpackbits(gd ^ rd).mean() >= 0.999, what


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('seed,W,H,nf,th', [(1, 752, 480, 1000, 20), (1000, 640, 512, 1500, 20), (100000, 1280, 1024, 2000, 20),
                                            (9, 401, 307, 600, 1)])
def test_cuda_equals_reference(gpu, synth, seed, W, H, nf, th):
    img = synth.synth_frame(seed, W, H)
    gk, gd = gpu.ORBextractor(nf, 1.2, 8, 1, th, max_width=W, max_height=H)(img)
    rk, rd = R.Extractor(nf, 1.2, 8, 1, th)(img)
    assert_cuda_matches(gk, gd, rk, rd, (seed, W, H))
    assert np.array_equal(gk['angle'], rk['angle']) and np.array_equal(gd, rd)      # measured: bit-identical, not merely within tolerance


@needs_ref
@pytest.mark.gpu
def test_cuda_equals_reference_occupancy_grid(gpu, synth):
    W, H, mpd = 752, 480, 20
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=W, max_height=H)
    rex = R.Extractor(1000, 1.2, 8, 1, 20)
    g_g = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F'); g_r = g_g.copy(order='F')
    rng = np.random.default_rng(5)
    for step, (n_in, need) in enumerate(((0, 400), (25, 150), (5, 1000))):
        img = synth.synth_frame(21 + step, W, H)
        inc = np.zeros(n_in, gpu.capi.KP_DTYPE)
        inc['x'] = rng.integers(40, W - 40, n_in); inc['y'] = rng.integers(40, H - 40, n_in)
        inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 7
        gk, gd = ex(img, keypoints=inc, grid_2d=g_g, min_px_dist=mpd, FullDetect=False, num_featsneeded=need)
        rk, rd = rex(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=False, num_needed=need)
        assert_cuda_matches(gk, gd, rk, rd, step)
        assert np.array_equal(g_g, g_r), step


# ---------------------------------------------------------------------------------------------------- matcher
# oracle/_ref/libref_orbmatcher.so = the reference's own src/ORBmatcher.cc over stand-in FrameKTL/KeyFrame/MapPoint data
# holders (oracle/ref_shim/slam_standin.h).  The oracle's matcher restatement must give the same claims.
needs_mref = pytest.mark.skipif(not R.matcher_available(), reason='oracle/_ref matcher not built and /root/reference absent')


def scale_factors(n=8, s=1.2):
    sf = [np.float32(1)]
    for _ in range(n - 1):
        sf.append(np.float32(sf[-1] * np.float32(s)))
    return np.array(sf, np.float32)


@needs_mref
def test_descriptor_distance_equals_reference(oracle):
    rng = np.random.default_rng(1)
    for i in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        if i == 0:
            b = a.copy()
        if i == 1:
            b = ~a
        assert R.descriptor_distance(a, b) == oracle.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())


def m4_oracle(oracle, c, th, ratio, W, H, taken=None, use=None):
    sf = scale_factors()
    rad = np.array([oracle.lib().uo_radius_by_viewing_cos(float(x)) for x in c['view_cos']], np.float32)
    if th != 1.0:
        rad = rad * np.float32(th)
    r_ = (rad * sf[c['level']]).astype(np.float32)
    inv_w = np.float32(64.0) / np.float32(W); inv_h = np.float32(48.0) / np.float32(H)
    start, items = oracle.grid_build(c['kx'], c['ky'], 0.0, 0.0, float(inv_w), float(inv_h))
    q = np.arange(len(c['u'])) if use is None else np.nonzero(use)[0]
    n, match, tk = oracle.search_window(0, 100, np.float32(ratio), c['u'][q], c['v'][q], r_[q], c['level'][q] - 1, c['level'][q], c['qdesc'][q],
                                        c['kx'], c['ky'], c['octave'], c['kdesc'], start, items, 0.0, 0.0, float(inv_w), float(inv_h), taken=taken)
    owner = np.where(tk >= 0, q[np.clip(tk, 0, len(q) - 1)], tk).astype(np.int32)
    return n, owner


@needs_mref
@pytest.mark.parametrize('th,ratio', [(1.0, 0.8), (3.0, 0.8), (5.0, 0.6), (1.0, 0.95)])
def test_search_by_projection_cfg3_equals_reference(oracle, synth, th, ratio):
    """BASELINE config 3: 10 000 projected map points vs a 2000-keypoint frame (src/ORBmatcher.cc:49-125)."""
    W, H = 752, 480
    c = synth.projection_case()
    nq, nk = len(c['u']), len(c['kx'])
    in_view = (np.arange(nq) % 17 != 3); bad = (np.arange(nq) % 29 == 7)
    taken = np.where(np.arange(nk) % 13 == 2, -2, -1).astype(np.int32)
    on, oowner = m4_oracle(oracle, c, th, ratio, W, H, taken=taken, use=in_view & ~bad)
    rn, rowner = R.search_by_projection_mps(c['kx'], c['ky'], c['octave'], c['kdesc'], [0, W, 0, H], scale_factors(), c['u'], c['v'], c['level'],
                                            c['view_cos'], c['qdesc'], th, ratio, taken=taken, in_view=in_view, bad=bad)
    assert on == rn and rn > 1000
    assert np.array_equal(oowner, rowner)


def frame_for_matching(oracle, synth, seed=1, W=752, H=480):
    img = synth.synth_frame(seed, W, H)
    return oracle.Extractor(1000, 1.2, 8, 1, 20)(img)


def m5_scene(kps, W, H, sf):
    """keyframe map points that project near the frame's own keypoints under a small pose change; replay of
    src/ORBmatcher.cc:1626-1672 with cv::Mat arithmetic as OpenCV 3.4 evaluates it (R*x+t: float products and sums;
    -R.t()*t and cv::norm: double accumulation)"""
    f32 = np.float32
    n = len(kps)
    fx, fy, cx, cy = f32(458.0), f32(457.0), f32(367.0), f32(248.0)
    c_, s_ = f32(0.99995), f32(0.0099998)
    T = np.array([[c_, -s_, 0, f32(0.01)], [s_, c_, 0, f32(-0.02)], [0, 0, 1, f32(0.03)], [0, 0, 0, 1]], np.float32)
    Rm, tv = T[:3, :3], T[:3, 3]
    Ow = np.array([f32(-1.0 * sum(float(Rm[r, c]) * float(tv[r]) for r in range(3))) for c in range(3)], np.float32)
    has_mp = (np.arange(n) % 9 != 0); bad = (np.arange(n) % 23 == 0); found = (np.arange(n) % 10 == 1)
    pos = np.zeros((n, 3), np.float32); min_dist = np.ones(n, np.float32)
    q = dict(u=[], v=[], r=[], lo=[], hi=[], who=[])
    for i in range(n):
        z = f32(2.0) + f32(i % 7) * f32(0.5)
        X = np.array([(kps['x'][i] - cx) / fx * z, (kps['y'][i] - cy) / fy * z, z], np.float32)
        pos[i] = X
        min_dist[i] = z / sf[kps['octave'][i]] * f32(1.05)
        if not has_mp[i] or bad[i] or found[i]:
            continue
        xc3 = np.zeros(3, np.float32)
        for r in range(3):
            t = f32(Rm[r, 0] * X[0])
            t = f32(t + f32(Rm[r, 1] * X[1])); t = f32(t + f32(Rm[r, 2] * X[2]))
            xc3[r] = f32(float(t) + float(tv[r]))
        invz = f32(1.0 / float(xc3[2]))
        u_ = f32(f32(f32(fx * xc3[0]) * invz) + cx); v_ = f32(f32(f32(fy * xc3[1]) * invz) + cy)
        if u_ < 0 or u_ > W or v_ < 0 or v_ > H:
            continue
        po = (X - Ow).astype(np.float32)
        d3 = f32(np.sqrt(sum(float(p) * float(p) for p in po)))
        ratio = f32(d3 / min_dist[i])
        lv = min(int(np.searchsorted(sf, ratio, side='left')), len(sf) - 1)
        q['u'].append(u_); q['v'].append(v_); q['lo'].append(lv - 1); q['hi'].append(lv + 1); q['who'].append(i); q['r'].append(lv)
    return dict(T=T, intr=[fx, fy, cx, cy], has_mp=has_mp, bad=bad, found=found, pos=pos, min_dist=min_dist, q=q)


@needs_mref
@pytest.mark.parametrize('th,orb_dist,check_ori', [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_keyframe_equals_reference(oracle, synth, th, orb_dist, check_ori):
    """src/ORBmatcher.cc:1622-1746: pose projection, lower_bound level prediction, levels [l-1, l+1], best-only, claims,
    rotation histogram (ComputeThreeMaxima :1748-1789)."""
    W, H = 752, 480
    kps, desc = frame_for_matching(oracle, synth)
    n = len(kps); sf = scale_factors()
    S = m5_scene(kps, W, H, sf); q = S['q']; who = np.array(q['who'])
    taken0 = np.where(np.arange(n) % 31 == 5, -2, -1).astype(np.int32)
    inv_w = np.float32(64.0) / np.float32(W); inv_h = np.float32(48.0) / np.float32(H)
    start, items = oracle.grid_build(kps['x'], kps['y'], 0.0, 0.0, float(inv_w), float(inv_h))
    qr = (np.float32(th) * sf[np.array(q['r'])]).astype(np.float32)
    on, om, _ = oracle.search_window(1, orb_dist, np.float32(0.9), q['u'], q['v'], qr, q['lo'], q['hi'], desc[who], kps['x'], kps['y'],
                                     kps['octave'].astype(np.int32), desc, start, items, 0.0, 0.0, float(inv_w), float(inv_h), taken=taken0)
    kept = oracle.rot_hist_filter(om, kps['angle'][who], kps['angle']) if check_ori else om
    expect = np.where(np.arange(n) % 31 == 5, -2, -1).astype(np.int32)
    for qi, k in enumerate(kept):
        if k >= 0:
            expect[k] = who[qi]
    rn, rowner = R.search_by_projection_kf(kps['x'], kps['y'], kps['octave'], kps['angle'], desc, [0, W, 0, H], sf, S['T'], S['intr'], S['has_mp'],
                                           S['bad'], S['found'], S['pos'], S['min_dist'], desc, kps['angle'], th, orb_dist, 0.9, check_ori, taken=taken0)
    assert rn == int((kept >= 0).sum()) and rn > 100
    assert np.array_equal(rowner, expect)


def bow_scene(kps, n):
    node = (kps['x'] / np.float32(64)).astype(np.int64) + 16 * (kps['y'] / np.float32(64)).astype(np.int64)
    return node


@needs_mref
@pytest.mark.parametrize('ratio,check_ori', [(0.9, True), (0.7, True), (0.9, False)])
def test_search_by_bow_keyframe_frame_equals_reference(oracle, synth, ratio, check_ori):
    """src/ORBmatcher.cc:155-284: merge-join of the two FeatureVectors, top-2 inside a node, TH_LOW, ratio, claims, histogram."""
    kps, desc = frame_for_matching(oracle, synth)
    kps2, desc2 = frame_for_matching(oracle, synth, seed=1)       # same scene: the frame's descriptors with a few bits flipped
    n = len(kps)
    rng = np.random.default_rng(3)
    fdesc = desc.copy()
    flip = rng.integers(0, 256, (n, 12))
    for j in range(12):
        sel = rng.random(n) < 0.7
        fdesc[sel, flip[sel, j] >> 3] ^= (1 << (flip[sel, j] & 7)).astype(np.uint8)
    fangle = np.mod(kps['angle'] + rng.normal(0, 8, n).astype(np.float32) + np.where(rng.random(n) < 0.2, 90, 0), 360).astype(np.float32)
    node = bow_scene(kps, n)
    kf_fv, f_fv = {}, {}
    for k in range(n):
        kf_fv.setdefault(int(node[k]), []).append(k)
        if k % 17 != 3:
            f_fv.setdefault(int(node[k]) + (1000 if k % 29 == 0 else 0), []).append(k)
    has_mp = (np.arange(n) % 7 != 0); bad = (np.arange(n) % 19 == 0)
    queries, cs, ci = [], [0], []
    for nd in sorted(kf_fv):
        if nd not in f_fv:
            continue
        for k in kf_fv[nd]:
            if not has_mp[k] or bad[k]:
                continue
            queries.append(k); ci.extend(f_fv[nd]); cs.append(len(ci))
    queries = np.array(queries)
    on, om, _ = oracle.search_lists(2, 50, np.float32(ratio), desc[queries], np.array(cs, np.int32), np.array(ci, np.int32), fdesc)
    if check_ori:
        om = oracle.rot_hist_filter(om, kps['angle'][queries], fangle)
    expect = np.full(n, -1, np.int32)
    for qi, k in enumerate(om):
        if k >= 0:
            expect[k] = queries[qi]
    rn, rmatch = R.search_by_bow_kf_frame(desc, kps['angle'], has_mp, bad, kf_fv, fdesc, fangle, f_fv, ratio, check_ori)
    assert rn == int((om >= 0).sum()) and rn > 200
    assert np.array_equal(rmatch, expect)


@needs_mref
@pytest.mark.parametrize('ratio,check_ori', [(0.9, True), (0.75, False)])
def test_search_by_bow_keyframe_keyframe_equals_reference(oracle, synth, ratio, check_ori):
    """src/ORBmatcher.cc:715-850: strict best < TH_LOW, ratio, claims on the second keyframe, candidates without a map point
    dropped, histogram rollback."""
    kps, desc = frame_for_matching(oracle, synth)
    n = len(kps)
    rng = np.random.default_rng(4)
    d2 = desc.copy()
    flip = rng.integers(0, 256, (n, 10))
    for j in range(10):
        sel = rng.random(n) < 0.6
        d2[sel, flip[sel, j] >> 3] ^= (1 << (flip[sel, j] & 7)).astype(np.uint8)
    a2 = np.mod(kps['angle'] + rng.normal(0, 6, n).astype(np.float32) + np.where(rng.random(n) < 0.15, 120, 0), 360).astype(np.float32)
    node = bow_scene(kps, n)
    fv1, fv2 = {}, {}
    for k in range(n):
        fv1.setdefault(int(node[k]), []).append(k)
        if k % 15 != 4:
            fv2.setdefault(int(node[k]) + (2000 if k % 37 == 0 else 0), []).append(k)
    has1 = (np.arange(n) % 7 != 0); bad1 = (np.arange(n) % 19 == 0)
    has2 = (np.arange(n) % 6 != 1); bad2 = (np.arange(n) % 21 == 2)
    q, cs, ci = [], [0], []
    for nd in sorted(fv1):
        if nd not in fv2:
            continue
        for k in fv1[nd]:
            if not has1[k] or bad1[k]:
                continue
            q.append(k); ci.extend(j for j in fv2[nd] if has2[j] and not bad2[j]); cs.append(len(ci))
    q = np.array(q)
    on, om, _ = oracle.search_lists(3, 50, np.float32(ratio), desc[q], np.array(cs, np.int32), np.array(ci or [0], np.int32), d2)
    if check_ori:
        om = oracle.rot_hist_filter(om, kps['angle'][q], a2)
    expect = np.full(n, -1, np.int32); expect[q[om >= 0]] = om[om >= 0]
    rn, r12 = R.search_by_bow_kf_kf(desc, kps['angle'], has1, bad1, fv1, d2, a2, has2, bad2, fv2, ratio, check_ori)
    assert rn == int((om >= 0).sum()) and rn > 150
    assert np.array_equal(r12, expect)


@needs_mref
@pytest.mark.gpu
@pytest.mark.parametrize('th', [1.0, 3.0])
def test_cuda_search_by_projection_cfg3_equals_reference(gpu, synth, th):
    """BASELINE config 3 on the GPU (k_grid_build + k_search_window through the C-ABI) against the reference's compiled
    ORBmatcher::SearchByProjection(FrameKTL&, vector<MapPoint*>&, th)."""
    W, H = 752, 480
    c = synth.projection_case()
    m = gpu.ORBmatcher(0.8, True)
    grid = m.grid_build(c['kx'], c['ky'], c['bounds'])
    sf = scale_factors()
    frame = dict(kx=c['kx'], ky=c['ky'], octave=c['octave'], kdesc=c['kdesc'], grid=grid, scale_factors=sf)
    mps = dict(u=c['u'], v=c['v'], level=c['level'], view_cos=c['view_cos'], desc=c['qdesc'])
    n, match, taken = m.SearchByProjection(frame, mps, th)
    rn, rowner = R.search_by_projection_mps(c['kx'], c['ky'], c['octave'], c['kdesc'], [0, W, 0, H], sf, c['u'], c['v'], c['level'],
                                            c['view_cos'], c['qdesc'], th, 0.8)
    assert n == rn and n > 1000
    assert np.array_equal(taken, rowner)


@needs_mref
@pytest.mark.gpu
def test_cuda_search_by_bow_lists_equals_reference(gpu, oracle, synth):
    """uvip_search_lists mode 2 + uvip_rot_hist_filter on the GPU against the reference's compiled SearchByBoW(KeyFrame*, FrameKTL&)."""
    kps, desc = frame_for_matching(oracle, synth)
    n = len(kps)
    rng = np.random.default_rng(3)
    fdesc = desc.copy()
    flip = rng.integers(0, 256, (n, 12))
    for j in range(12):
        sel = rng.random(n) < 0.7
        fdesc[sel, flip[sel, j] >> 3] ^= (1 << (flip[sel, j] & 7)).astype(np.uint8)
    fangle = np.mod(kps['angle'] + rng.normal(0, 8, n).astype(np.float32), 360).astype(np.float32)
    node = bow_scene(kps, n)
    kf_fv, f_fv = {}, {}
    for k in range(n):
        kf_fv.setdefault(int(node[k]), []).append(k)
        if k % 17 != 3:
            f_fv.setdefault(int(node[k]), []).append(k)
    has_mp = (np.arange(n) % 7 != 0); bad = (np.arange(n) % 19 == 0)
    queries, cs, ci = [], [0], []
    for nd in sorted(kf_fv):
        if nd not in f_fv:
            continue
        for k in kf_fv[nd]:
            if has_mp[k] and not bad[k]:
                queries.append(k); ci.extend(f_fv[nd]); cs.append(len(ci))
    queries = np.array(queries)
    m = gpu.ORBmatcher(0.9, True)
    gn, gm, _ = m.search_lists(2, 50, desc[queries], np.array(cs, np.int32), np.array(ci, np.int32), fdesc)
    gm = m.rot_hist_filter(gm, kps['angle'][queries], fangle)
    expect = np.full(n, -1, np.int32)
    for qi, k in enumerate(gm):
        if k >= 0:
            expect[k] = queries[qi]
    rn, rmatch = R.search_by_bow_kf_frame(desc, kps['angle'], has_mp, bad, kf_fv, fdesc, fangle, f_fv, 0.9, True)
    assert rn == int((gm >= 0).sum()) and rn > 200
    assert np.array_equal(rmatch, expect)


def fuse_scene(oracle, kps, desc, W, H, sf, th):
    """map points that project near the keyframe's keypoints; numpy replay of the geometry of src/ORBmatcher.cc:1036-1086 and
    the oracle's best-only search without claims (mode 4) for :1088-1113"""
    f32 = np.float32
    n = len(kps)
    fx, fy, cx, cy = f32(458.0), f32(457.0), f32(367.0), f32(248.0)
    c_, s_ = f32(0.99995), f32(0.0099998)
    Rm = np.array([[c_, -s_, 0], [s_, c_, 0], [0, 0, 1]], np.float32); tv = np.array([0.01, -0.02, 0.03], np.float32)
    Ow = np.array([f32(-1.0 * sum(float(Rm[r, c]) * float(tv[r]) for r in range(3))) for c in range(3)], np.float32)
    npnt = 2 * n
    src = np.arange(npnt) % n
    rng = np.random.default_rng(11)
    is_null = (np.arange(npnt) % 14 == 3); bad = (np.arange(npnt) % 25 == 6); in_kf = (np.arange(npnt) % 9 == 2)
    pos = np.zeros((npnt, 3), np.float32); normal = np.zeros((npnt, 3), np.float32)
    min_dist = np.ones(npnt, np.float32); max_dist = np.ones(npnt, np.float32)
    pdesc = desc[src].copy()
    flip = rng.integers(0, 256, (npnt, 14))
    for j in range(14):
        sel = rng.random(npnt) < 0.6
        pdesc[sel, flip[sel, j] >> 3] ^= (1 << (flip[sel, j] & 7)).astype(np.uint8)
    q = dict(u=[], v=[], r=[], lo=[], hi=[], who=[])
    for i in range(npnt):
        k = src[i]
        z = f32(2.0) + f32(i % 7) * f32(0.5)
        if i % 41 == 0:
            z = f32(-1.0)                                  # behind the camera
        jx = f32(((i * 7) % 5) - 2) * f32(0.8); jy = f32(((i * 3) % 5) - 2) * f32(0.6)
        X = np.array([(kps['x'][k] + jx - cx) / fx * z, (kps['y'][k] + jy - cy) / fy * z, z], np.float32)
        pos[i] = X
        min_dist[i] = abs(z) / sf[kps['octave'][k]] * f32(1.05) * (f32(3.0) if i % 37 == 1 else f32(1.0))     # some too close
        max_dist[i] = abs(z) * f32(1.4) * (f32(0.5) if i % 43 == 2 else f32(1.0))                                # some too far
        nrm = X - Ow
        if i % 31 == 4:
            nrm = np.array([1, 0, 0], np.float32) * np.linalg.norm(nrm)                                        # viewed from the side
        normal[i] = (nrm / np.linalg.norm(nrm)).astype(np.float32)
        if is_null[i] or bad[i] or in_kf[i]:
            continue
        xc3 = np.zeros(3, np.float32)
        for r in range(3):
            t = f32(Rm[r, 0] * X[0]); t = f32(t + f32(Rm[r, 1] * X[1])); t = f32(t + f32(Rm[r, 2] * X[2]))
            xc3[r] = f32(float(t) + float(tv[r]))
        if xc3[2] < 0:
            continue
        invz = f32(f32(1) / xc3[2])
        x_ = f32(xc3[0] * invz); y_ = f32(xc3[1] * invz)
        u_ = f32(f32(fx * x_) + cx); v_ = f32(f32(fy * y_) + cy)
        if not (u_ >= 0 and u_ < W and v_ >= 0 and v_ < H):
            continue
        po = (X - Ow).astype(np.float32)
        d3 = f32(np.sqrt(sum(float(p) * float(p) for p in po)))
        if d3 < min_dist[i] or d3 > max_dist[i]:
            continue
        if sum(float(a) * float(b) for a, b in zip(po, normal[i])) < 0.5 * float(d3):
            continue
        ratio = f32(d3 / min_dist[i])
        lv = min(int(np.searchsorted(sf, ratio, side='left')), len(sf) - 1)
        q['u'].append(u_); q['v'].append(v_); q['r'].append(f32(f32(th) * sf[lv])); q['lo'].append(lv - 1); q['hi'].append(lv); q['who'].append(i)
    return dict(Rm=Rm, tv=tv, Ow=Ow, intr=[fx, fy, cx, cy], is_null=is_null, bad=bad, in_kf=in_kf, pos=pos, normal=normal,
                min_dist=min_dist, max_dist=max_dist, pdesc=pdesc, q=q)


@needs_mref
@pytest.mark.parametrize('th', [2.5, 4.0])
def test_fuse_equals_reference(oracle, synth, th):
    """M8: ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>&, th) (src/ORBmatcher.cc:1016-1134) — frustum / distance / viewing-angle
    gates, level prediction, KeyFrame::GetFeaturesInArea window, levels [l-1, l], best-only at TH_LOW WITHOUT claims (the
    oracle's / kernel's mode 4), then the host-side Replace / AddObservation bookkeeping in map-point order."""
    W, H = 752, 480
    kps, desc = frame_for_matching(oracle, synth)
    n = len(kps); sf = scale_factors()
    S = fuse_scene(oracle, kps, desc, W, H, sf, th); q = S['q']; who = np.array(q['who'])
    inv_w = np.float32(64.0) / np.float32(W); inv_h = np.float32(48.0) / np.float32(H)
    start, items = oracle.grid_build(kps['x'], kps['y'], 0.0, 0.0, float(inv_w), float(inv_h))
    on, om, _ = oracle.search_window(4, 50, np.float32(0.6), q['u'], q['v'], q['r'], q['lo'], q['hi'], S['pdesc'][who], kps['x'], kps['y'],
                                     kps['octave'].astype(np.int32), desc, start, items, 0.0, 0.0, float(inv_w), float(inv_h))
    kf_has = (np.arange(n) % 3 != 0); kf_bad = (np.arange(n) % 16 == 5)
    slot = [(-1 if not kf_has[k] else k) for k in range(n)]          # k = the keyframe's own map point, -2-j = map point j added by Fuse
    slot_bad = [bool(kf_bad[k]) for k in range(n)]
    e_action = np.zeros(len(S['is_null']), np.int32); e_target = np.full(len(S['is_null']), -1, np.int32)
    fused = 0
    for qi, k in enumerate(om):
        if k < 0:
            continue
        i = who[qi]
        fused += 1
        if slot[k] != -1:
            if not slot_bad[k]:
                e_action[i] = 1; e_target[i] = slot[k]
        else:
            e_action[i] = 2; e_target[i] = k
            slot[k] = -2 - i; slot_bad[k] = False
    rn, action, target = R.fuse(kps['x'], kps['y'], kps['octave'], desc, [0, W, 0, H], sf, S['Rm'], S['tv'], S['Ow'], S['intr'], kf_has, kf_bad,
                                S['is_null'], S['bad'], S['in_kf'], S['pos'], S['normal'], S['min_dist'], S['max_dist'], S['pdesc'], th)
    assert rn == fused == on and rn > 300
    assert (e_action == 1).sum() > 50 and (e_action == 2).sum() > 50 and (e_target < -1).sum() > 0     # every branch exercised
    assert np.array_equal(action, e_action) and np.array_equal(target, e_target)


# ---------------------------------------------------------------------------------------------------- row M8 through the shim
def m8_bundle(oracle, synth, which, th=None):
    """scene for the keyframe-side searches (array order: oracle/ref_shim/m8_scene.h); th < 0 makes the scene code call
    Fuse with its DEFAULT th argument (include/ORBmatcher.h:85,88)"""
    f32 = np.float32
    W, H = 752, 480
    kps, desc = frame_for_matching(oracle, synth)
    n = len(kps); sf = scale_factors()
    rng = np.random.default_rng(100 + which)
    fx, fy, cx, cy = f32(458.0), f32(457.0), f32(367.0), f32(248.0)

    def rot(ax, ay, az):
        cxr, sxr, cyr, syr, czr, szr = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
        Rx = np.array([[1, 0, 0], [0, cxr, -sxr], [0, sxr, cxr]]); Ry = np.array([[cyr, 0, syr], [0, 1, 0], [-syr, 0, cyr]])
        Rz = np.array([[czr, -szr, 0], [szr, czr, 0], [0, 0, 1]])
        return (Rz @ Ry @ Rx).astype(np.float32)
    R1, t1 = rot(0.004, -0.003, 0.01), np.array([0.01, -0.02, 0.03], np.float32)
    R2, t2 = rot(-0.002, 0.006, 0.004), np.array([-0.04, 0.01, 0.02], np.float32)
    ow = lambda R, t: (-(R.astype(np.float64).T @ t.astype(np.float64))).astype(np.float32)
    O1, O2 = ow(R1, t1), ow(R2, t2)
    z = (2.0 + (np.arange(n) % 7) * 0.5).astype(np.float32)
    Xc = np.stack([(kps['x'] - cx) / fx * z, (kps['y'] - cy) / fy * z, z], 1).astype(np.float64)
    Xw = ((Xc - t1.astype(np.float64)) @ R1.astype(np.float64)).astype(np.float32)           # R1^T (Xc - t1)

    def geo(X, Ocam, octave, i):
        d = np.linalg.norm((X - Ocam).astype(np.float64), axis=1)
        mind = (d / (sf[octave] * 0.98)).astype(np.float32)
        maxd = (mind * sf[-1] * 1.3).astype(np.float32)
        nrm = ((X - Ocam) / d[:, None]).astype(np.float32)
        mind = np.where(i % 37 == 1, mind * 3, mind).astype(np.float32)                      # some too close
        maxd = np.where(i % 43 == 2, mind * 0.5, maxd).astype(np.float32)                    # some too far
        nrm[i % 31 == 4] = np.array([1, 0, 0], np.float32)                                   # some seen from the side
        return np.concatenate([X, nrm, mind[:, None], maxd[:, None]], 1).astype(np.float32)

    def flips(d, k, p):
        out = d.copy()
        pos = rng.integers(0, 256, (len(d), k))
        for j in range(k):
            sel = rng.random(len(d)) < p
            out[sel, pos[sel, j] >> 3] ^= (1 << (pos[sel, j] & 7)).astype(np.uint8)
        return out

    idx = np.arange(n)
    k1_xyoa = np.stack([kps['x'], kps['y'], kps['octave'].astype(np.float32), kps['angle']], 1).astype(np.float32)
    k1_mp = np.stack([(idx % 3 != 0), (idx % 16 == 5), np.where(idx % 13 == 6, idx, -1)], 1).astype(np.int32)
    k1_geo = geo(Xw, O1, kps['octave'], idx)
    # keyframe 2 sees the same 3-D points from a slightly different pose
    Xc2 = (Xw.astype(np.float64) @ R2.astype(np.float64).T + t2.astype(np.float64))
    u2 = (fx * Xc2[:, 0] / Xc2[:, 2] + cx + rng.normal(0, 0.6, n)).astype(np.float32)
    v2 = (fy * Xc2[:, 1] / Xc2[:, 2] + cy + rng.normal(0, 0.6, n)).astype(np.float32)
    k2_xyoa = np.stack([u2, v2, kps['octave'].astype(np.float32), np.mod(kps['angle'] + rng.normal(0, 5, n), 360).astype(np.float32)], 1).astype(np.float32)
    k2_desc = flips(desc, 10, 0.6)
    k2_mp = np.stack([(idx % 4 != 1), (idx % 18 == 7), np.full(n, -1)], 1).astype(np.int32)
    k2_geo = geo(Xw, O2, kps['octave'], idx + 5)
    # candidate map points: two per keypoint, jittered around the keypoint's 3-D point
    npnt = 2 * n
    pi = np.arange(npnt); src = pi % n
    jit = np.stack([((pi * 7) % 5 - 2) * 0.004, ((pi * 3) % 5 - 2) * 0.003, np.zeros(npnt)], 1)
    Pw = (Xw[src].astype(np.float64) + jit * z[src, None]).astype(np.float32)
    Pw[pi % 41 == 0] = (O1 - (Xw[src][pi % 41 == 0] - O1)).astype(np.float32)                # behind the camera
    p_flags = np.stack([(pi % 14 == 3), (pi % 25 == 6), (pi % 9 == 2), np.zeros(npnt)], 1).astype(np.int32)
    p_geo = geo(Pw, O1, kps['octave'][src], pi + 11)
    p_desc = flips(desc[src], 14, 0.6)
    k1_matched = np.where(idx % 11 == 4, (idx * 2 + 1) % npnt, -1).astype(np.int32)
    # Sim3 between the cameras (s12 = 1): p1 = R12 p2 + t12;  Scw = 1.7 * [R1 | t1]
    R12 = (R1.astype(np.float64) @ R2.astype(np.float64).T).astype(np.float32)
    t12 = (t1.astype(np.float64) - R12.astype(np.float64) @ t2.astype(np.float64)).astype(np.float32)
    Scw = np.eye(4, dtype=np.float32); Scw[:3, :3] = f32(1.7) * R1; Scw[:3, 3] = f32(1.7) * t1
    th = [3.0, 4.0, 10.0, 7.5, 0.0, 12.0, 12.0, 14.0, 7.0, 15.0][which] if th is None else th
    pose = lambda R, t, O: np.concatenate([R.ravel(), t, O]).astype(np.float32)
    # fundamental matrix of the pair, F12 = K^-T [t12]x R12 K^-1 (src/LocalMapping.cc ComputeF12), and vocabulary nodes
    Kinv = np.linalg.inv(np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64))
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]], np.float64)
    F12 = (Kinv.T @ tx @ R12.astype(np.float64) @ Kinv).astype(np.float32)
    node1 = ((kps['x'] / 64).astype(np.int32) + 16 * (kps['y'] / 64).astype(np.int32)).astype(np.int32)
    node2 = np.where(idx % 15 == 4, -1, node1 + np.where(idx % 37 == 0, 2000, 0)).astype(np.int32)
    return [np.array([th, 1.0, 0.6], np.float32), sf, np.array([fx, fy, cx, cy], np.float32), np.array([0, W, 0, H], np.int32), Scw, R12, t12,
            k1_xyoa, desc, pose(R1, t1, O1), k1_mp, k1_geo, k2_xyoa, k2_desc, pose(R2, t2, O2), k2_mp, k2_geo, p_flags, p_geo, p_desc, k1_matched,
            F12, node1, node2]


M8_NAMES = ['Fuse(KF, MPs, th)', 'Fuse(KF, Scw, MPs, th)', 'SearchByProjection(KF, Scw, MPs, vpMatched, th)', 'SearchBySim3', 'SearchForTriangulation',
            'WindowSearch(F1, F2, windowSize, matches2)', 'SearchByProjection(F1, F2, windowSize, matches2)', 'SearchForInitialization',
            'SearchByProjection(CurrentFrame, LastFrame, th)', 'WindowSearch(F1, F2, windowSize, matches2, 1, 3)']
M8_MIN = [150, 150, 150, 150, 25, 100, 20, 100, 20, 50]


@needs_mref
@pytest.mark.parametrize('which', [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_m8_reference_runs_every_branch(oracle, synth, which, tmp_path):
    """the scenes must exercise the reference's code paths (otherwise the GPU comparison below proves little)"""
    R.write_bundle(str(tmp_path / 's.bin'), m8_bundle(oracle, synth, which))
    r, (ret, slot_owner, replaced, obs_slot) = R.m8_run(str(tmp_path / 's.bin'), str(tmp_path / 'o.bin'), which)
    assert ret[0] == r and r > M8_MIN[which], (M8_NAMES[which], r)
    if which == 4:
        assert (slot_owner >= 0).sum() == r and len(replaced) == 2 * r and obs_slot.all()
        assert (np.diff(replaced[0::2]) > 0).all()                                # pairs come out in ascending order of the first index
    elif which in (0, 1):
        assert (replaced >= 0).sum() > 30 and (obs_slot >= 0).sum() > 30          # both bookkeeping branches
        assert (slot_owner >= 200000).sum() == (obs_slot >= 0).sum()              # added points sit in the keyframe's slots
        if which == 1:
            assert (replaced[:len(slot_owner)] >= 200000).sum() > 30              # Fuse(Scw) replaces the KEYFRAME's points
    elif which == 2:
        assert (slot_owner >= 200000).sum() >= r
    elif which == 3:
        assert (slot_owner >= 100000).sum() >= r
    elif which == 7:
        assert (slot_owner >= 0).sum() == r and len(replaced) == 2 * len(slot_owner)      # vnMatches12 and the updated vbPrevMatched
    else:
        assert (slot_owner >= 0).sum() >= r                                             # frame 2's slots hold frame 1's map points


@needs_mref
@pytest.mark.parametrize('which', [0, 1])
def test_m8_reference_fuse_default_is_2_5(oracle, synth, which, tmp_path):
    """both Fuse overloads default to th = 2.5 (include/ORBmatcher.h:85,88; src/LocalMapping.cc:1236,1261 rely on it): the scene
    called WITHOUT th equals th = 2.5 and differs from th = 3.0, so a drop-in header with another default cannot pass
    test_shim_m8_fuse_defaults_equal_reference"""
    out = {}
    for name, th in (('default', -1.0), ('2.5', 2.5), ('3.0', 3.0)):
        R.write_bundle(str(tmp_path / 's.bin'), m8_bundle(oracle, synth, which, th=th))
        out[name] = R.m8_run(str(tmp_path / 's.bin'), str(tmp_path / 'o.bin'), which)
    assert out['default'][0] == out['2.5'][0] and all(np.array_equal(a, b) for a, b in zip(out['default'][1], out['2.5'][1]))
    assert not all(np.array_equal(a, b) for a, b in zip(out['default'][1], out['3.0'][1]))


@needs_mref
@pytest.mark.gpu
@pytest.mark.parametrize('which', [0, 1])
def test_shim_m8_fuse_defaults_equal_reference(gpu, oracle, synth, which, tmp_path):
    """Fuse(pKF, vpMapPoints) / Fuse(pKF, Scw, vpPoints) with the default th through the shim == the reference's compiled call"""
    import os, subprocess
    assert os.path.exists(R.M8_EXE)
    R.write_bundle(str(tmp_path / 's.bin'), m8_bundle(oracle, synth, which, th=-1.0))
    r, ref = R.m8_run(str(tmp_path / 's.bin'), str(tmp_path / 'ref.bin'), which)
    p = subprocess.run([R.M8_EXE, str(tmp_path / 's.bin'), str(tmp_path / 'shim.bin'), str(which)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    shim = R.read_bundle(str(tmp_path / 'shim.bin'))
    assert int(p.stdout.strip()) == r and r > 100
    for a, b, name in zip(shim, ref, ('return', 'slot owner', 'replaced', 'observation')):
        assert np.array_equal(a, b), (M8_NAMES[which], name, int((a != b).sum()))


@needs_mref
def test_shim_m8_driver_refuses_without_device(pkg, oracle, synth, tmp_path):
    import os, subprocess
    if pkg.capi.lib().uvip_device_count() > 0 or not os.path.exists(R.M8_EXE):
        pytest.skip('a CUDA device is present, or the prebuilt driver is missing')
    R.write_bundle(str(tmp_path / 's.bin'), m8_bundle(oracle, synth, 0))
    p = subprocess.run([R.M8_EXE, str(tmp_path / 's.bin'), str(tmp_path / 'o.bin'), '0'], capture_output=True, text=True)
    assert p.returncode == 3 and 'no CPU fallback' in p.stderr


@needs_mref
@pytest.mark.gpu
@pytest.mark.parametrize('which', [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_shim_m8_equals_reference(gpu, oracle, synth, which, tmp_path):
    """row M8 end to end: the drop-in shim (host geometry + uvip_search_window on the GPU) against the reference's compiled
    ORBmatcher, same scene code, same stand-in SLAM types, same call signatures; result bundles must be identical.  which >= 5:
    the four public overloads of include/ORBmatcher.h:49-72 that this fork never calls (WindowSearch, SearchByProjection(F1, F2, ...),
    SearchForInitialization, SearchByProjection(CurrentFrame, LastFrame, th))."""
    import os, subprocess
    assert os.path.exists(R.M8_EXE), 'oracle/_ref/test_shim_m8 is prebuilt by `make -C oracle ref` in the build container'
    R.write_bundle(str(tmp_path / 's.bin'), m8_bundle(oracle, synth, which))
    r, ref = R.m8_run(str(tmp_path / 's.bin'), str(tmp_path / 'ref.bin'), which)
    p = subprocess.run([R.M8_EXE, str(tmp_path / 's.bin'), str(tmp_path / 'shim.bin'), str(which)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    shim = R.read_bundle(str(tmp_path / 'shim.bin'))
    assert int(p.stdout.strip()) == r and r > M8_MIN[which], (M8_NAMES[which], p.stdout, r)
    for a, b, name in zip(shim, ref, ('return', 'slot owner', 'replaced', 'observation')):
        assert np.array_equal(a, b), (M8_NAMES[which], name, int((a != b).sum()))


@needs_mref
def test_search_for_triangulation_core_equals_reference(oracle, synth, tmp_path):
    """the oracle's restatement of SearchForTriangulation's matching core (src/ORBmatcher.cc:893-952, :136-153) against the
    reference's compiled function on the M8 scene: same pairs, histogram applied afterwards"""
    B = m8_bundle(oracle, synth, 4)
    R.write_bundle(str(tmp_path / 's.bin'), B)
    r, (ret, slot_owner, replaced, obs_slot) = R.m8_run(str(tmp_path / 's.bin'), str(tmp_path / 'o.bin'), 4)
    k1_xyoa, d1, k1_mp, k2_xyoa, d2, k2_mp, F12, node1, node2, sf = B[7], B[8], B[10], B[12], B[13], B[15], B[21], B[22], B[23], B[1]
    f32 = np.float32
    q, ql, cs, ci = [], [], [0], []
    by2 = {}
    for k in range(len(node2)):
        if node2[k] >= 0:
            by2.setdefault(int(node2[k]), []).append(k)
    for nd in sorted(set(node1.tolist())):
        if nd not in by2:
            continue
        for k in np.nonzero(node1 == nd)[0]:
            if k1_mp[k, 0]:
                continue
            x, y = k1_xyoa[k, 0], k1_xyoa[k, 1]
            a = f32(f32(f32(x * F12[0, 0]) + f32(y * F12[1, 0])) + F12[2, 0]); b = f32(f32(f32(x * F12[0, 1]) + f32(y * F12[1, 1])) + F12[2, 1])
            c = f32(f32(f32(x * F12[0, 2]) + f32(y * F12[1, 2])) + F12[2, 2])
            q.append(k); ql.append([a, b, c, f32(f32(a * a) + f32(b * b))])
            ci.extend(j for j in by2[nd] if not k2_mp[j, 0]); cs.append(len(ci))
    q = np.array(q)
    sig2 = (sf * sf).astype(np.float32)
    kthr = 3.84 * sig2[k2_xyoa[:, 2].astype(np.int32)].astype(np.float64)
    on, om, _ = oracle.search_lists_epipolar(50, d1[q], np.array(ql, np.float32), np.array(cs, np.int32), np.array(ci or [0], np.int32), d2,
                                             k2_xyoa[:, 0], k2_xyoa[:, 1], kthr)
    om = oracle.rot_hist_filter(om, k1_xyoa[q, 3], k2_xyoa[:, 3])
    expect = np.full(len(node1), -1, np.int32); expect[q[om >= 0]] = om[om >= 0]
    assert r == int((om >= 0).sum()) and np.array_equal(slot_owner, expect)


# ---------------------------------------------------------------------------------------------------- row E8, scoring half
@needs_ref
def test_harris_responses_equal_reference_dead_path(oracle, synth):
    """HarrisResponses (src/ORBextractor.cc:80-121) is only reachable from the dead ComputeKeyPoints path (:536-746); the
    reference's compiled code is driven into it through a subclass and the responses it leaves must be reproduced bit for bit"""
    for seed, W, H in ((1, 752, 480), (1000, 640, 512)):
        img = synth.synth_frame(seed, W, H)
        rex = R.Extractor(1000, 1.2, 8, 0, 20)              # HARRIS_SCORE
        oex = oracle.Extractor(1000, 1.2, 8, 0, 20); oex(img)
        total = 0
        for l in range(8):
            k = R.dead_path_keypoints(rex, img, l)
            assert len(k) > 20 and (k['octave'] == l).all()
            assert np.array_equal(oracle.harris_responses(oex.level(l), k['x'], k['y']), k['response']), (seed, l)
            total += len(k)
        assert total > 800


@needs_ref
@pytest.mark.gpu
def test_cuda_harris_responses_equal_reference_dead_path(gpu, synth):
    img = synth.synth_frame(1, 752, 480)
    ex = gpu.ORBextractor(1000, 1.2, 8, 0, 20, max_width=752, max_height=480)
    ex(img)
    rex = R.Extractor(1000, 1.2, 8, 0, 20)
    for l in range(8):
        k = R.dead_path_keypoints(rex, img, l)
        assert np.array_equal(ex.harris_responses(l, k['x'], k['y']), k['response']), l


# ---------------------------------------------------------------------------------------------------- next row N4: haloc hash
needs_href = pytest.mark.skipif(not R.hash_available(), reason='oracle/_ref hash not built and /root/reference absent')


def haloc_sets(oracle, synth):
    sets = [frame_for_matching(oracle, synth, seed=s)[1] for s in (1, 2, 3)]
    sets.append(sets[0][:1]); sets.append(sets[1][:137])
    return sets


@needs_href
def test_haloc_hash_and_match_equal_reference(oracle, synth):
    """haloc::Hash::getHash / match (src/hash.cpp:57-85, :190-206) of the reference's compiled code on its own projection vectors"""
    rh = R.HalocHash(3)
    sets = haloc_sets(oracle, synth)
    ref = [rh.get_hash(d) for d in sets]
    proj = rh.projections()
    assert proj.shape == (3, 6000) and abs(float((proj[0].astype(np.float64) ** 2).sum()) - 1.0) < 1e-3      # unit vectors
    assert abs(float((proj[0].astype(np.float64) * proj[1].astype(np.float64)).sum())) < 1e-3                 # orthogonalised (:112-146)
    ours = [oracle.haloc_hash(d, proj) for d in sets]
    for a, b in zip(ours, ref):
        assert np.array_equal(a, b)
    for i in range(len(sets)):
        for j in range(len(sets)):
            assert oracle.haloc_match(ours[i], ours[j]) == rh.match(ref[i], ref[j])


@needs_href
@pytest.mark.gpu
def test_cuda_haloc_hash_and_match_equal_reference(gpu, oracle, synth):
    rh = R.HalocHash(3)
    sets = haloc_sets(oracle, synth)
    ref = np.stack([rh.get_hash(d) for d in sets])
    proj = rh.projections()
    start = np.zeros(len(sets) + 1, np.int32); start[1:] = np.cumsum([len(d) for d in sets])
    m = gpu.ORBmatcher(0.6, True)
    got = m.haloc_hash(np.concatenate(sets), start, proj)
    assert np.array_equal(got, ref)
    scores = m.haloc_match(ref[0], ref)
    assert np.array_equal(scores, np.array([rh.match(ref[0], r) for r in ref], np.float32))


# ---------------------------------------------------------------------------------------------------- next row N2: DBoW2
needs_vref = pytest.mark.skipif(not R.dbow_available(), reason='oracle/_ref DBoW2 not built and /root/reference absent')


def voc_case(synth, k, L, ragged, seed, n):
    tree, lines = synth.synthetic_vocabulary(k, L, seed=seed, ragged=ragged)
    desc = synth.random_descriptors(19 + seed, n)
    src = np.random.default_rng(2).integers(1, len(tree['word']), n // 2)
    desc[:n // 2] = synth.flip_bits(tree['desc'][src], 44, np.random.default_rng(3).integers(0, 60, n // 2).tolist())
    return tree, lines, desc


@needs_vref
@pytest.mark.parametrize('k,L,ragged', [(10, 3, False), (6, 4, True), (10, 5, False)])
def test_vocabulary_transform_equals_reference_dbow2(pkg, oracle, synth, tmp_path, k, L, ragged):
    """the reference's own DBoW2 (text loader :1338-1420, transform :1175-1259, L1-normalised TF-IDF) against the oracle descent
    + the host-side BowVector / FeatureVector bookkeeping of frontend.ORBVocabulary"""
    tree, lines, desc = voc_case(synth, k, L, ragged, 5 + k, 600)
    path = tmp_path / 'voc.txt'
    # no trailing newline: the reference's loader (`while(!f.eof())`, :1376) turns a final empty line into one more child of the
    # root whose leaf flag is an uninitialised variable — undefined behaviour, not a target
    path.write_text('\n'.join(lines))
    rv = R.Vocabulary(path)
    assert rv.size() == int((tree['word'] >= 0).sum())
    parsed = pkg.ORBVocabulary.parse_text(str(path))
    # ragged trees: a descent that ends above level L - levelsup never writes `nid` in the reference (:1230-1253), whose caller
    # then files the feature under an uninitialised variable (:1152) — undefined behaviour, so only levels every descent reaches
    for levelsup in ((0, 2, 4, L, L + 3) if not ragged else (L - 1, L, L + 3)):
        rbow, rfv = rv.transform(desc, levelsup)
        wid, nid, w = oracle.bow_transform(parsed, desc, levelsup)
        bow, fv = pkg.ORBVocabulary.accumulate(wid, nid, w, parsed['weighting'], parsed['scoring'])
        assert sorted(bow) == sorted(rbow) and all(bow[x] == rbow[x] for x in bow), (k, L, levelsup)      # values bit-identical (double)
        assert {a: list(b) for a, b in fv.items()} == rfv, (k, L, levelsup)


@needs_vref
@pytest.mark.gpu
def test_cuda_vocabulary_transform_equals_reference_dbow2(gpu, synth, tmp_path):
    tree, lines, desc = voc_case(synth, 10, 5, False, 15, 3000)
    path = tmp_path / 'voc.txt'
    path.write_text('\n'.join(lines))
    rv = R.Vocabulary(path)
    voc = gpu.ORBVocabulary(); voc.loadFromTextFile(str(path))
    for levelsup in (4, 2):
        bow, fv = voc.transform(desc, levelsup)
        rbow, rfv = rv.transform(desc, levelsup)
        assert sorted(bow) == sorted(rbow) and all(bow[x] == rbow[x] for x in bow)
        assert {a: list(b) for a, b in fv.items()} == rfv
    voc.close()


# ---------------------------------------------------------------------------------------------------- next row N4, first half
needs_pref = pytest.mark.skipif(not R.mappoint_available(), reason='oracle/_ref MapPoint not built and /root/reference absent')


def distinctive_case(synth):
    rng = np.random.default_rng(8)
    sizes = [0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 33, 64, 100] + rng.integers(1, 40, 120).tolist()
    start = np.zeros(len(sizes) + 1, np.int32); start[1:] = np.cumsum(sizes)
    base = synth.random_descriptors(123, len(sizes))
    rows = []
    for p, n in enumerate(sizes):                       # observations of one map point: its descriptor with a few bits flipped, some exact duplicates
        d = np.repeat(base[p:p + 1], n, 0)
        d = synth.flip_bits(d, 1000 + p, rng.integers(0, 25, n).tolist()) if n else d
        if n > 3:
            d[n - 1] = d[0]
        rows.append(d)
    return np.concatenate(rows), start, sizes


@needs_pref
def test_distinctive_descriptors_equal_reference_mappoint(oracle, synth):
    """the reference's REAL MapPoint class (src/MapPoint.cc compiled over stand-in KeyFrame / Map): the descriptor
    ComputeDistinctiveDescriptors (:197-270) keeps must be the one the oracle's best index names"""
    desc, start, sizes = distinctive_case(synth)
    rdesc, chosen = R.distinctive_descriptors(desc, start)
    best, med = oracle.distinctive_descriptors(desc, start)
    for p, n in enumerate(sizes):
        assert chosen[p] == (1 if n else 0)
        if n:
            assert np.array_equal(rdesc[p], desc[start[p] + best[p]]), (p, n)


@needs_pref
@pytest.mark.gpu
def test_cuda_distinctive_descriptors_equal_reference_mappoint(gpu, synth):
    desc, start, sizes = distinctive_case(synth)
    rdesc, chosen = R.distinctive_descriptors(desc, start)
    best, med = gpu.ORBmatcher(0.6, True).distinctive_descriptors(desc, start)
    for p, n in enumerate(sizes):
        if n:
            assert np.array_equal(rdesc[p], desc[start[p] + best[p]]), (p, n)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('score_type,seed,W,H,nf', [(0, 1, 752, 480, 1000), (1, 1, 752, 480, 1000), (0, 1000, 640, 512, 1500), (1, 7, 320, 240, 500)])
def test_cuda_quota_detector_equals_reference_dead_path(gpu, synth, score_type, seed, W, H, nf):
    """row E8 as an optional mode: uvip_compute_keypoints_quota against the reference's compiled ComputeKeyPoints
    (src/ORBextractor.cc:536-746, reached through a subclass): quota cells, FAST + retry at 5, Harris / FAST responses, quota
    redistribution, retainBest per cell and per level (std::nth_element on both sides), orientation — same keypoints, same order"""
    img = synth.synth_frame(seed, W, H)
    ex = gpu.ORBextractor(nf, 1.2, 8, score_type, 20, max_width=W, max_height=H)
    ex(img)
    got = ex.compute_keypoints_quota()
    rex = R.Extractor(nf, 1.2, 8, score_type, 20)
    total = 0
    for l in range(8):
        ref = R.dead_path_keypoints(rex, img, l)
        assert len(got[l]) == len(ref), (l, len(got[l]), len(ref))
        for f in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
            assert np.array_equal(got[l][f], ref[f]), (l, f)
        assert np.abs(got[l]['angle'] - ref['angle']).max() <= ANGLE_TOL_DEG, l
        assert np.array_equal(got[l]['angle'], ref['angle']), l
        total += len(ref)
    assert total > 0.8 * nf


# ---------------------------------------------------------------------------------------------------- the frame grid, for real
needs_fref = pytest.mark.skipif(not R.frontend_available(), reason='oracle/_ref mini front-end not built and /root/reference absent')


def local_points_scene(oracle, synth, th):
    f32 = np.float32
    W, H = 752, 480
    kps, desc = frame_for_matching(oracle, synth)
    n = len(kps)
    # keypoints away from integer grid lines too: undistorted coordinates are not integers in the reference
    rng = np.random.default_rng(17)
    kx = (kps['x'] + rng.uniform(-0.49, 0.49, n)).astype(np.float32); ky = (kps['y'] + rng.uniform(-0.49, 0.49, n)).astype(np.float32)
    kxyoa = np.stack([kx, ky, kps['octave'].astype(np.float32), kps['angle']], 1).astype(np.float32)
    fx, fy, cx, cy = f32(458.0), f32(457.0), f32(367.0), f32(248.0)
    c_, s_ = f32(0.99995), f32(0.0099998)
    T = np.array([[c_, -s_, 0, 0.01], [s_, c_, 0, -0.02], [0, 0, 1, 0.03], [0, 0, 0, 1]], np.float32)
    R1, t1 = T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64)
    O1 = (-(R1.T @ t1)).astype(np.float32)
    npnt = 3 * n
    pi = np.arange(npnt); src = pi % n
    z = (2.0 + (pi % 7) * 0.5)
    jit = np.stack([((pi * 7) % 5 - 2) * 0.9, ((pi * 3) % 5 - 2) * 0.7], 1)
    Xc = np.stack([(kx[src] + jit[:, 0] - cx) / fx * z, (ky[src] + jit[:, 1] - cy) / fy * z, z], 1).astype(np.float64)
    Xw = ((Xc - t1) @ R1).astype(np.float32)
    Xw[pi % 41 == 0] = (O1 - (Xw[pi % 41 == 0] - O1)).astype(np.float32)                 # behind the camera
    Xw[pi % 53 == 1] += np.array([40, 0, 0], np.float32)                                 # outside the image
    ref_Ow = (O1 + np.array([0.3, -0.1, 0.05], np.float32)).astype(np.float32)           # the observing keyframe sits elsewhere
    obs_level = np.clip(kps['octave'][src] + (pi % 3) - 1, 0, 7).astype(np.int32)
    pdesc = desc[src].copy()
    flip = rng.integers(0, 256, (npnt, 16))
    for j in range(16):
        sel = rng.random(npnt) < 0.6
        pdesc[sel, flip[sel, j] >> 3] ^= (1 << (flip[sel, j] & 7)).astype(np.uint8)
    return dict(W=W, H=H, kxyoa=kxyoa, kdesc=desc, intr=[fx, fy, cx, cy], T=T, pos=Xw, obs_level=obs_level, ref_Ow=ref_Ow, pdesc=pdesc, th=th)


def replay_local_points(oracle, S, ref, search):
    """grid + search of the product / oracle on what the reference's isInFrustum produced"""
    W, H = S['W'], S['H']
    sf = scale_factors()
    inv = np.nonzero(ref['inview'])[0]
    rad = np.array([oracle.lib().uo_radius_by_viewing_cos(float(c)) for c in ref['viewcos'][inv]], np.float32)
    if S['th'] != 1.0:
        rad = rad * np.float32(S['th'])
    r = (rad * sf[ref['level'][inv]]).astype(np.float32)
    n, match, taken = search(ref['u'][inv], ref['v'][inv], r, ref['level'][inv] - 1, ref['level'][inv], S['pdesc'][inv])
    owner = np.where(taken >= 0, inv[np.clip(taken, 0, len(inv) - 1)], -1).astype(np.int32)
    return n, owner


@needs_fref
@pytest.mark.parametrize('th', [1.0, 3.0])
def test_frame_grid_and_search_local_points_equal_real_frontend(oracle, synth, th):
    """FrameKTL::PosInGrid + grid fill (src/FrameKTL.cc:250-264,426-436), GetFeaturesInArea (:359-424), isInFrustum (:299-357) and
    ORBmatcher::SearchByProjection (:49-125) of the reference's REAL classes compiled together, against the oracle's grid and
    window search on the same frame"""
    S = local_points_scene(oracle, synth, th)
    ref = R.search_local_points(S['kxyoa'], S['kdesc'], [0, S['W'], 0, S['H']], S['intr'], 8, 1.2, S['T'], S['pos'], S['obs_level'], S['ref_Ow'],
                                S['pdesc'], th, 0.8)
    assert ref['inview'].sum() > 1500 and (ref['inview'] == 0).sum() > 100
    kx, ky, octave = S['kxyoa'][:, 0], S['kxyoa'][:, 1], S['kxyoa'][:, 2].astype(np.int32)
    inv_w = np.float32(64.0) / np.float32(S['W']); inv_h = np.float32(48.0) / np.float32(S['H'])
    start, items = oracle.grid_build(kx, ky, 0.0, 0.0, float(inv_w), float(inv_h))
    assert np.array_equal(start, ref['cell_start']) and np.array_equal(items, ref['cell_items'])      # the grid itself, cell by cell
    search = lambda u, v, r, lo, hi, d: oracle.search_window(0, 100, np.float32(0.8), u, v, r, lo, hi, d, kx, ky, octave, S['kdesc'], start, items,
                                                             0.0, 0.0, float(inv_w), float(inv_h))
    n, owner = replay_local_points(oracle, S, ref, search)
    assert n == ref['n'] and n > 500
    assert np.array_equal(owner, ref['owner'])


@needs_fref
@pytest.mark.gpu
@pytest.mark.parametrize('th', [1.0, 3.0])
def test_cuda_grid_and_search_equal_real_frontend(gpu, oracle, synth, th):
    S = local_points_scene(oracle, synth, th)
    ref = R.search_local_points(S['kxyoa'], S['kdesc'], [0, S['W'], 0, S['H']], S['intr'], 8, 1.2, S['T'], S['pos'], S['obs_level'], S['ref_Ow'],
                                S['pdesc'], th, 0.8)
    kx, ky, octave = S['kxyoa'][:, 0], S['kxyoa'][:, 1], S['kxyoa'][:, 2].astype(np.int32)
    m = gpu.ORBmatcher(0.8, True)
    grid = m.grid_build(kx, ky, (0, S['W'], 0, S['H']))
    assert np.array_equal(grid['start'], ref['cell_start']) and np.array_equal(grid['items'], ref['cell_items'])
    search = lambda u, v, r, lo, hi, d: m.search_frame(0, 100, u, v, r, lo, hi, d, kx, ky, octave, S['kdesc'], (0, S['W'], 0, S['H']))
    n, owner = replay_local_points(oracle, S, ref, search)
    assert n == ref['n'] and np.array_equal(owner, ref['owner'])
