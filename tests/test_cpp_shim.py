"""The C++ drop-in shim (u-vip-slam_b200/host/ORBextractor.h, ORBmatcher.h): same class names and call signatures as
the reference's include/ORBextractor.h / include/ORBmatcher.h, forwarding to the C-ABI.  CPU: it compiles and links
against libuvip_orb.so and fails loudly without a device.  GPU: driven the way Tracking.cc drives it and compared with
the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def shim_exe(pkg, tmp_path_factory):
    pkg.capi.lib()
    out = str(tmp_path_factory.mktemp('shim') / 'test_shim')
    libdir = os.path.join(ROOT, 'u-vip-slam_b200')
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.check_call([cxx, '-std=c++14', '-O2', '-Wall', '-Werror', '-o', out, os.path.join(ROOT, 'tests', 'cpp', 'test_shim.cpp'),
                           '-L', libdir, '-luvip_orb', '-Wl,-rpath,' + libdir])
    return out


def _write_input(path, img, nfeatures, fast_th, mpd, need):
    h, w = img.shape
    with open(path, 'wb') as f:
        f.write(struct.pack('6i', w, h, nfeatures, fast_th, mpd, need))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def test_shim_compiles_and_refuses_without_device(pkg, synth, shim_exe, tmp_path):
    if pkg.capi.lib().uvip_device_count() > 0:
        pytest.skip('a CUDA device is present')
    _write_input(tmp_path / 'in.bin', synth.synth_frame(3, 320, 240), 300, 20, 20, 100)
    r = subprocess.run([shim_exe, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], capture_output=True, text=True)
    assert r.returncode == 3 and 'no CPU fallback' in r.stderr


@pytest.mark.gpu
def test_shim_matches_oracle(pkg, oracle, synth, shim_exe, tmp_path):
    W, H, NF, TH, MPD, NEED = 752, 480, 1000, 20, 20, 300
    img = synth.synth_frame(1, W, H)
    _write_input(tmp_path / 'in.bin', img, NF, TH, MPD, NEED)
    r = subprocess.run([shim_exe, str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(tmp_path / 'out.bin', 'rb').read()
    KP = pkg.capi.KP_DTYPE
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off); off += a.nbytes
        return a

    # 1. full detection
    n = int(take(np.int32, 1)[0]); kps = take(KP, n); desc = take(np.uint8, n * 32).reshape(n, 32)
    oex = oracle.Extractor(NF, 1.2, 8, 1, TH)
    okps, odesc = oex(img)
    assert n == len(okps)
    for fld in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
        assert np.array_equal(kps[fld], okps[fld]), fld
    assert np.abs(kps['angle'] - okps['angle']).max() <= 1e-3 * 180 / np.pi
    assert 1.0 - np.unpackbits(desc ^ odesc).mean() >= 0.999
    # 2. replenishing detection on the caller's occupancy grid
    n2 = int(take(np.int32, 1)[0]); kps2 = take(KP, n2); desc2 = take(np.uint8, n2 * 32).reshape(n2, 32)
    gr, gc = [int(v) for v in take(np.int32, 2)]
    grid = take(np.int32, gr * gc).reshape(gc, gr).T            # column-major
    inc = okps[:20].copy(); inc['x'] = np.floor(inc['x']); inc['y'] = np.floor(inc['y']); inc['octave'] = 0
    og = np.zeros((H // MPD + 2, W // MPD + 2), np.int32, order='F')
    for k in inc:
        og[int(k['y'] / MPD), int(k['x'] / MPD)] += 1
    okps2, odesc2 = oex(img, keypoints=inc, grid=og, min_px_dist=MPD, full_detect=False, num_needed=NEED)
    assert n2 == len(okps2) and (gr, gc) == og.shape and np.array_equal(grid, og)
    for fld in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
        assert np.array_equal(kps2[fld], okps2[fld]), fld
    assert 1.0 - np.unpackbits(desc2 ^ odesc2).mean() >= 0.999
    assert int(take(np.int32, 1)[0]) == 1                        # empty image left the outputs untouched
    # 3. SearchByProjection through the matcher shim, claims included
    nm = int(take(np.int32, 1)[0]); owner = take(np.int32, n)
    i = np.arange(n)
    u = (okps['x'] + ((i % 5) - 2).astype(np.float32) * np.float32(0.7)).astype(np.float32)
    v = (okps['y'] + ((i % 3) - 1).astype(np.float32) * np.float32(0.9)).astype(np.float32)
    lvl = okps['octave'].astype(np.int32)
    cosv = np.where(i & 1, np.float32(0.9990), np.float32(0.9)).astype(np.float32)
    use = ((i % 11) != 0) & ((i % 13) != 0)
    sf = [np.float32(1)]
    for _ in range(7):
        sf.append(np.float32(sf[-1] * np.float32(1.2)))
    sf = np.array(sf, np.float32)
    r_ = np.array([oracle.lib().uo_radius_by_viewing_cos(float(c)) for c in cosv], np.float32) * sf[lvl]
    inv_w = np.float32(64.0) / np.float32(W); inv_h = np.float32(48.0) / np.float32(H)
    start, items = oracle.grid_build(okps['x'], okps['y'], 0.0, 0.0, float(inv_w), float(inv_h))
    q = np.nonzero(use)[0]
    on, omatch, otaken = oracle.search_window(0, 100, np.float32(0.8), u[q], v[q], r_[q].astype(np.float32), lvl[q] - 1, lvl[q], odesc[q],
                                              okps['x'], okps['y'], lvl, odesc, start, items, 0.0, 0.0, float(inv_w), float(inv_h))
    assert nm == on and nm > 500
    expect_owner = np.full(n, -1, np.int32)
    for qi, k in enumerate(omatch):
        if k >= 0:
            expect_owner[k] = q[qi]
    assert np.array_equal(owner, expect_owner)
    dd = int(take(np.int32, 1)[0])
    assert dd == oracle.descriptor_distance(odesc[0], odesc[1])
    # 4. SearchByBoW(KeyFrame*, Frame&) through the shim: node-major query order, claims, TH_LOW, ratio 0.9, histogram
    nb = int(take(np.int32, 1)[0]); bowner = take(np.int32, n)
    node = (okps['x'] / np.float32(64)).astype(np.int64) + 16 * (okps['y'] / np.float32(64)).astype(np.int64)
    fnode = {}
    for k in range(n):
        if k % 17 != 3:
            fnode.setdefault(int(node[k]) + (1000 if k % 29 == 0 else 0), []).append(k)
    queries, cs, ci = [], [0], []
    for nd in sorted(set(node.tolist())):
        if nd not in fnode:
            continue
        for k in np.nonzero(node == nd)[0]:
            if k % 7 == 0 or k % 19 == 0:
                continue                                          # no map point / bad map point
            queries.append(int(k)); ci.extend(fnode[nd]); cs.append(len(ci))
    queries = np.array(queries)
    on, omatch, _ = oracle.search_lists(2, 50, np.float32(0.9), odesc[queries], np.array(cs, np.int32), np.array(ci, np.int32), odesc)
    omatch = oracle.rot_hist_filter(omatch, okps['angle'][queries], okps['angle'])
    expect = np.full(n, -1, np.int32)
    for qi, k in enumerate(omatch):
        if k >= 0:
            expect[k] = queries[qi]
    assert nb == int((omatch >= 0).sum()) and nb > 300
    assert np.array_equal(bowner, expect)
    # 5. ComputeDistinctiveDescriptors batched through the shim (uses the shim's own descriptors, 99.9 % equal to the oracle's)
    npnt = int(take(np.int32, 1)[0]); best = take(np.int32, npnt)
    sizes = [1 + p % 9 for p in range(npnt)]
    start = np.zeros(npnt + 1, np.int32); start[1:] = np.cumsum(sizes)
    rows = np.concatenate([desc[[(p + 7 * j) % n for j in range(sizes[p])]] for p in range(npnt)])
    obest, _ = oracle.distinctive_descriptors(rows, start)
    assert npnt == 64 and np.array_equal(best, obest)
    # 6. SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th=10, ORBdist=100) through the shim: projection with the
    #    current pose (cv::gemm small-matrix float path for R*x+t, double accumulation for -R.t()*t and cv::norm), predicted level by lower_bound on the
    #    scale factors, window levels [l-1, l+1], best-only + claims, rotation histogram rollback (src/ORBmatcher.cc:1622-1746)
    nr = int(take(np.int32, 1)[0]); rowner = take(np.int32, n)
    f32 = np.float32
    fx, fy, cx, cy = f32(458.0), f32(457.0), f32(367.0), f32(248.0)
    c_, s_ = f32(0.99995), f32(0.0099998)
    P = np.array([[c_, -s_, 0, f32(0.01)], [s_, c_, 0, f32(-0.02)], [0, 0, 1, f32(0.03)]], np.float32)
    Rm, tv = P[:, :3], P[:, 3]
    Ow = np.zeros(3, np.float32)
    for c in range(3):
        acc = 0.0
        for r in range(3):
            acc += float(Rm[r, c]) * float(tv[r])
        Ow[c] = f32(-1.0 * acc)
    # the shim used ITS OWN keypoints / descriptors (kps, desc): identical to the oracle's except for the angle / descriptor tolerance
    qu, qv, qr, qlo, qhi, qa, who = [], [], [], [], [], [], []
    for i in range(n):
        if i % 9 == 0 or i % 23 == 0 or i % 10 == 1:
            continue
        z = f32(2.0) + f32(i % 7) * f32(0.5)
        X = np.array([(kps['x'][i] - cx) / fx * z, (kps['y'][i] - cy) / fy * z, z], np.float32)
        xc3 = np.zeros(3, np.float32)
        for r in range(3):                  # cv::gemm 3x3 path: float products and sums, (float)(t + c) in double
            t_ = f32(Rm[r, 0] * X[0]); t_ = f32(t_ + f32(Rm[r, 1] * X[1])); t_ = f32(t_ + f32(Rm[r, 2] * X[2]))
            xc3[r] = f32(float(t_) + float(tv[r]))
        invz = f32(1.0 / float(xc3[2]))
        u_ = fx * xc3[0] * invz + cx; v_ = fy * xc3[1] * invz + cy
        if u_ < 0 or u_ > W or v_ < 0 or v_ > H:
            continue
        po = X - Ow
        d3 = f32(np.sqrt(float(po[0]) * float(po[0]) + float(po[1]) * float(po[1]) + float(po[2]) * float(po[2])))
        mind = z / sf[kps['octave'][i]] * f32(1.05)
        ratio = d3 / mind
        lv = min(int(np.searchsorted(sf, ratio, side='left')), 7)
        qu.append(u_); qv.append(v_); qr.append(f32(10.0) * sf[lv]); qlo.append(lv - 1); qhi.append(lv + 1); qa.append(kps['angle'][i]); who.append(i)
    taken0 = np.where(np.arange(n) % 31 == 5, -2, -1).astype(np.int32)
    start3, items3 = oracle.grid_build(kps['x'], kps['y'], 0.0, 0.0, float(inv_w), float(inv_h))
    on3, om3, _ = oracle.search_window(1, 100, np.float32(0.9), qu, qv, qr, qlo, qhi, desc[who], kps['x'], kps['y'], kps['octave'].astype(np.int32),
                                       desc, start3, items3, 0.0, 0.0, float(inv_w), float(inv_h), taken=taken0)
    kept = oracle.rot_hist_filter(om3, np.array(qa, np.float32), kps['angle'])
    expect3 = np.where(np.arange(n) % 31 == 5, 0, -1).astype(np.int32)
    for qi, k in enumerate(kept):
        if k >= 0:
            expect3[k] = who[qi]
    assert nr == int((kept >= 0).sum()) and nr > 300, (nr, int((kept >= 0).sum()))
    assert np.array_equal(rowner, expect3)
    # 7. SearchByBoW(KeyFrame*, KeyFrame*) through the shim (src/ORBmatcher.cc:715-850): strict best < TH_LOW, ratio 0.9, claims
    #    on the second keyframe, candidates without a good map point dropped, histogram rollback
    nkk = int(take(np.int32, 1)[0]); o12 = take(np.int32, n)
    node_s = (kps['x'] / np.float32(64)).astype(np.int64) + 16 * (kps['y'] / np.float32(64)).astype(np.int64)
    anode, bnode = {}, {}
    for k in range(n):
        anode.setdefault(int(node_s[k]), []).append(k)
        if k % 15 != 4:
            bnode.setdefault(int(node_s[k]) + (2000 if k % 37 == 0 else 0), []).append(k)
    q7, cs7, ci7 = [], [0], []
    for nd in sorted(anode):
        if nd not in bnode:
            continue
        for k in anode[nd]:
            if k % 7 == 0 or k % 19 == 0:
                continue
            q7.append(k); ci7.extend(j for j in bnode[nd] if j % 6 != 1 and j % 21 != 2); cs7.append(len(ci7))
    q7 = np.array(q7)
    on7, om7, _ = oracle.search_lists(3, 50, np.float32(0.9), desc[q7], np.array(cs7, np.int32), np.array(ci7 or [0], np.int32), desc)
    om7 = oracle.rot_hist_filter(om7, kps['angle'][q7], kps['angle'])
    e12 = np.full(n, -1, np.int32); e12[q7[om7 >= 0]] = om7[om7 >= 0]
    assert nkk == int((om7 >= 0).sum()) and nkk > 200, (nkk, int((om7 >= 0).sum()))
    assert np.array_equal(o12, e12)


@pytest.fixture(scope='module')
def threads_exe(pkg, tmp_path_factory):
    pkg.capi.lib()
    out = str(tmp_path_factory.mktemp('shim') / 'test_threads')
    libdir = os.path.join(ROOT, 'u-vip-slam_b200')
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.check_call([cxx, '-std=c++14', '-O2', '-Wall', '-Werror', '-pthread', '-o', out, os.path.join(ROOT, 'tests', 'cpp', 'test_threads.cpp'),
                           '-L', libdir, '-luvip_orb', '-Wl,-rpath,' + libdir])
    return out


def test_threads_driver_compiles_and_refuses_without_device(pkg, threads_exe):
    if pkg.capi.lib().uvip_device_count() > 0:
        pytest.skip('a CUDA device is present')
    r = subprocess.run([threads_exe], capture_output=True, text=True)
    assert r.returncode == 3 and 'no CPU fallback' in r.stderr


@pytest.mark.gpu
def test_three_matcher_threads_and_one_extractor_thread_equal_serial(threads_exe):
    """stack-local ORBmatcher instances used concurrently from three host threads (src/Tracking.cc:2222,
    src/LocalMapping.cc:1230, src/LoopClosing.cc:373) while the Tracking thread extracts; every concurrent result equals the
    serial one.  The shim's device handle is thread-local, so the per-call construct / destruct allocates nothing."""
    import json
    r = subprocess.run([threads_exe, '40'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout, r.stderr)
    rec = json.loads(r.stdout.strip().splitlines()[-1])
    assert rec['mismatches'] == 0 and rec['threads'] == 4 and min(rec['matches']) >= 50
    print('shim call shape (construct + SearchByProjection + destruct):', rec['construct_search_destruct_us'])
