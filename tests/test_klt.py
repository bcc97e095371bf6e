"""next row N1: KLT front end — cv::buildOpticalFlowPyramid (src/FrameKTL.cc:76) + cv::calcOpticalFlowPyrLK
(src/Tracking.cc:1044-1047).  Integer stages bit-exact; positions within 1e-2 px of OpenCV (float reduction order)."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
POS_TOL = 1e-2          # pixels


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _pts(synth):
    import gen_golden
    return gen_golden.klt_points(synth, 500, 752, 480)


def _check(p1, st, err, gp1, gst, gerr, min_status_agree=0.995):
    """Lucas-Kanade iterates in float until |delta| <= 0.01 px, so two correct implementations that sum the window in a
    different order can stop one iteration apart, and ill-conditioned windows (tiny min eigenvalue) amplify rounding.
    Bar: status flags and min-eigenvalues agree; >= 98 % of the tracked points within 1e-2 px, median <= 1e-3 px."""
    agree = (st == gst)
    assert agree.mean() >= min_status_agree, agree.mean()
    both = (st == 1) & (gst == 1)
    assert both.sum() > 300
    d = np.abs(p1 - gp1).max(1)[both]
    assert (d <= POS_TOL).mean() >= 0.98, (d <= POS_TOL).mean()
    assert np.median(d) <= 1e-3
    assert np.abs(err - gerr)[both].max() <= 1e-5 * max(1.0, float(np.abs(gerr[both]).max()))


def test_oracle_klt_against_cv2(oracle, synth, golden):
    a = synth.synth_frame(1, 752, 480); b = synth.synth_frame(1, 752, 480, dx=5, dy=3, noise_seed=2)
    assert sha(oracle.pyr_down(a)) == str(golden['klt_pyrdown_sha'])
    assert sha(oracle.scharr(a)) == str(golden['klt_scharr_sha'])
    p0 = _pts(synth)
    for win, lev in ((21, 5), (9, 3)):
        P0 = oracle.LKPyramid(a, win, lev); P1 = oracle.LKPyramid(b, win, lev)
        p1, st, err = oracle.lk_track(P0, P1, p0, p0 + np.float32([2.0, 1.5]), win, lev, 30, 0.01, 12)
        _check(p1, st, err, golden['klt_p1_w%d' % win], golden['klt_st_w%d' % win], golden['klt_err_w%d' % win])
        flow = (p1 - p0)[(st == 1)]
        assert np.abs(np.median(flow, 0) - np.array([-5.0, -3.0])).max() < 0.05     # the twin frame is shifted by (5, 3)


@pytest.mark.gpu
def test_gpu_klt_against_oracle_and_cv2(pkg, oracle, synth, golden):
    a = synth.synth_frame(1, 752, 480); b = synth.synth_frame(1, 752, 480, dx=5, dy=3, noise_seed=2)
    p0 = _pts(synth)
    for win, lev in ((21, 5), (9, 3)):
        klt = pkg.KLTTracker(752, 480, win, lev, nslots=2)
        n0 = klt.build_pyramid(0, a); n1 = klt.build_pyramid(1, b)
        P0 = oracle.LKPyramid(a, win, lev); P1 = oracle.LKPyramid(b, win, lev)
        assert n0 == n1 == P0.levels()
        for l in range(n0):
            gi, gd = klt.level(0, l); oi, od = P0.level(l)
            assert np.array_equal(gi, oi) and np.array_equal(gd, od), ('pyramid level', l)      # pyrDown + Scharr bit-exact
        p1, st, err = klt.track(0, 1, p0, p0 + np.float32([2.0, 1.5]))
        o1, ost, oerr = oracle.lk_track(P0, P1, p0, p0 + np.float32([2.0, 1.5]), win, lev, 30, 0.01, 12)
        _check(p1, st, err, o1, ost, oerr)
        _check(p1, st, err, golden['klt_p1_w%d' % win], golden['klt_st_w%d' % win], golden['klt_err_w%d' % win])
        # without an initial guess, and reversed roles (a frame's pyramid serves as prev and as next)
        q1, qst, qerr = klt.track(1, 0, p0, p0, flags=8)
        r1, rst, rerr = oracle.lk_track(P1, P0, p0, p0, win, lev, 30, 0.01, 8)
        _check(q1, qst, qerr, r1, rst, rerr)
        klt.close()
    # a smaller frame through a handle sized for a larger one, odd sizes
    c = synth.synth_frame(4, 401, 307); d = synth.synth_frame(4, 401, 307, dx=-3, dy=2, noise_seed=5)
    klt = pkg.KLTTracker(752, 480, 21, 5)
    klt.build_pyramid(0, c); klt.build_pyramid(1, d)
    P0 = oracle.LKPyramid(c, 21, 5); P1 = oracle.LKPyramid(d, 21, 5)
    pts = p0[(p0[:, 0] < 390) & (p0[:, 1] < 300)]
    p1, st, err = klt.track(0, 1, pts, pts, flags=8)
    o1, ost, oerr = oracle.lk_track(P0, P1, pts, pts, 21, 5, 30, 0.01, 8)
    agree = st == ost
    assert agree.mean() > 0.99 and (np.abs(p1 - o1).max(1)[(st == 1) & (ost == 1)] <= POS_TOL).mean() >= 0.98


@pytest.mark.gpu
def test_gpu_klt_refuses_slots_of_another_geometry(pkg, synth):
    """a pyramid built for another frame size invalidates the other slots (they are laid out for the old plan): tracking between
    a stale slot and a fresh one is an argument error, not garbage with status 1"""
    a = synth.synth_frame(1, 752, 480); c = synth.synth_frame(4, 401, 307)
    klt = pkg.KLTTracker(752, 480, 21, 5, nslots=3)
    klt.build_pyramid(0, a); klt.build_pyramid(1, a)
    p0 = _pts(synth)[:50]
    klt.track(0, 1, p0, p0, flags=8)
    klt.build_pyramid(2, c)                                  # new geometry: slots 0 and 1 are stale now
    with pytest.raises(pkg.capi.UvipError) as e:
        klt.track(0, 2, p0, p0, flags=8)
    assert e.value.code == pkg.capi.ERR_ARG
    with pytest.raises(pkg.capi.UvipError):
        klt.track(0, 1, p0, p0, flags=8)
    klt.build_pyramid(0, c)
    pts = p0[(p0[:, 0] < 390) & (p0[:, 1] < 300)]
    p1, st, err = klt.track(0, 2, pts, pts, flags=8)         # same image in both slots: zero flow
    assert st.sum() > 0 and np.abs(p1 - pts)[st == 1].max() < 1e-3
    klt.close()


@pytest.mark.gpu
def test_gpu_klt_21x21_kernel_equals_generic_kernel(pkg, synth):
    """the tracker specialised for the reference's 21x21 window (offsets in registers, second image staged in shared memory, float
    derivative window) returns the bits of the generic kernel: same integer interpolation, same float products, same summation order.
    Large initial errors force the staged region to be re-centred inside a level."""
    a = synth.synth_frame(1, 752, 480); b = synth.synth_frame(1, 752, 480, dx=5, dy=3, noise_seed=2)
    p0 = _pts(synth)
    klt = pkg.KLTTracker(752, 480, 21, 5, nslots=2)
    klt.build_pyramid(0, a); klt.build_pyramid(1, b)
    for guess, flags in ((np.float32([2.0, 1.5]), 12), (np.float32([-14.0, 11.0]), 12), (np.float32([0, 0]), 8)):
        os.environ.pop('UVIP_KLT_GENERIC', None)
        fast = klt.track(0, 1, p0, p0 + guess, flags=flags)
        os.environ['UVIP_KLT_GENERIC'] = '1'
        try:
            gen = klt.track(0, 1, p0, p0 + guess, flags=flags)
        finally:
            os.environ.pop('UVIP_KLT_GENERIC', None)
        for x, y in zip(fast, gen):
            assert np.array_equal(x, y)
        assert fast[1].sum() > 300
    klt.close()


# ---- N1, last third: cv::findFundamentalMat(FM_RANSAC) inlier mask (src/Tracking.cc:1062) ------------------------------------------
RANSAC_CASES = (101, 102, 103, 104)


def _ransac_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'cv2_ransac.npz'))


def _inliers_of(F, p0, p1, thr=1.0):
    """OpenCV's FMEstimatorCallback::computeError + the (float) threshold test, in numpy"""
    p0 = p0.astype(np.float64); p1 = p1.astype(np.float64)
    x0, y0, x1, y1 = p0[:, 0], p0[:, 1], p1[:, 0], p1[:, 1]
    a = F[0, 0] * x0 + F[0, 1] * y0 + F[0, 2]; b = F[1, 0] * x0 + F[1, 1] * y0 + F[1, 2]; c = F[2, 0] * x0 + F[2, 1] * y0 + F[2, 2]
    s2 = 1.0 / (a * a + b * b); d2 = x1 * a + y1 * b + c
    a = F[0, 0] * x1 + F[1, 0] * y1 + F[2, 0]; b = F[0, 1] * x1 + F[1, 1] * y1 + F[2, 1]; c = F[0, 2] * x1 + F[1, 2] * y1 + F[2, 2]
    s1 = 1.0 / (a * a + b * b); d1 = x0 * a + y0 * b + c
    return (np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(np.float32) <= np.float32(thr * thr)).astype(np.uint8)


def test_oracle_ransac_against_cv2(oracle):
    """cv::RNG cannot be reproduced, so the pin is (i) the residual + threshold test: with cv2's own F it reproduces cv2's mask exactly;
    (ii) the consensus: the restatement (2048 hypotheses, no early stop) finds at least cv2's inlier count, its inliers are true
    inliers of the synthetic geometry, and the two masks agree on >= 95 % of the points"""
    g = _ransac_golden()
    for seed in RANSAC_CASES:
        p0, p1, cm, out = g['p0_%d' % seed], g['p1_%d' % seed], g['mask_%d' % seed], g['outlier_%d' % seed]
        assert np.array_equal(_inliers_of(g['F_%d' % seed], p0, p1), cm), seed            # (i)
        cnt, mask, F = oracle.ransac_fundamental(p0, p1, 1.0, nhyp=2048)
        assert cnt == mask.sum() and np.array_equal(_inliers_of(F, p0, p1), mask)
        assert cnt >= cm.sum(), (seed, cnt, int(cm.sum()))                                # (ii)
        assert (mask == cm).mean() >= 0.95, (seed, (mask == cm).mean())
        assert mask[out == 1].sum() <= 0.02 * len(mask) + 2                               # gross outliers are rejected
        assert abs(np.linalg.det(F)) < 1e-9                                               # rank 2 (the cubic's root)
    cnt, mask, F = oracle.ransac_fundamental(g['p0_101'][:14], g['p1_101'][:14])
    assert cnt == -1                                                                      # below 15 points OpenCV runs LMedS


@pytest.mark.gpu
def test_gpu_ransac_equals_oracle(pkg, oracle):
    """one warp per hypothesis on the GPU == the sequential oracle, bit for bit (count, mask and F): the solver uses only + - * / sqrt"""
    g = _ransac_golden()
    klt = pkg.KLTTracker(752, 480, 21, 5)
    for seed in RANSAC_CASES:
        p0, p1 = g['p0_%d' % seed], g['p1_%d' % seed]
        for nh in (64, 2048):
            cnt, mask, F = klt.ransac_fundamental(p0, p1, 1.0, nhyp=nh)
            ocnt, omask, oF = oracle.ransac_fundamental(p0, p1, 1.0, nhyp=nh)
            assert cnt == ocnt and np.array_equal(mask, omask) and np.array_equal(F, oF), (seed, nh, cnt, ocnt)
        assert cnt >= g['mask_%d' % seed].sum()
    with pytest.raises(pkg.capi.UvipError) as e:
        klt.ransac_fundamental(g['p0_101'][:14], g['p1_101'][:14])
    assert e.value.code == pkg.capi.ERR_UNSUPPORTED
    # all points on one degenerate configuration (identical points): no hypothesis, empty mask, no fault
    z = np.zeros((20, 2), np.float32)
    cnt, mask, F = klt.ransac_fundamental(z, z)
    ocnt, omask, oF = oracle.ransac_fundamental(z, z, 1.0, nhyp=2048)
    assert cnt == ocnt and np.array_equal(mask, omask)
    klt.close()


@pytest.mark.gpu
def test_gpu_klt_sequence_device_equals_pairwise_calls(pkg, synth):
    """uvip_klt_track_sequence_device (pyramids of a whole device-resident sequence, all consecutive pairs in one launch) gives exactly
    what build_pyramid + track give pair by pair"""
    import ctypes as C
    import torch
    W, H, NFR, NP = 752, 480, 5, 400
    frames = np.stack([synth.synth_frame(1, W, H, dx=2 * f, dy=f, noise_seed=10 + f) for f in range(NFR)])
    p0 = _pts(synth)[:NP]
    klt = pkg.KLTTracker(W, H, 21, 5, nslots=8)
    want = []
    for f in range(NFR - 1):
        klt.build_pyramid(0, frames[f]); klt.build_pyramid(1, frames[f + 1])
        want.append(klt.track(0, 1, p0, p0 + np.float32([1.0, 0.5])))
    d_fr = torch.from_numpy(frames).cuda()
    prev = torch.from_numpy(np.tile(p0[None], (NFR - 1, 1, 1))).cuda().contiguous()
    nxt = (prev + torch.tensor([1.0, 0.5], device='cuda')).contiguous()
    st = torch.zeros((NFR - 1, NP), dtype=torch.uint8, device='cuda'); er = torch.zeros((NFR - 1, NP), dtype=torch.float32, device='cuda')
    P = lambda t: C.c_void_p(t.data_ptr())
    pkg.capi.check(pkg.capi.lib().uvip_klt_track_sequence_device(klt.h, P(d_fr), NFR, W, H, W, W * H, P(prev), P(nxt), NP, 5, 30, 0.01, 12, 1e-4, P(st), P(er),
                                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)))
    torch.cuda.synchronize()
    for f in range(NFR - 1):
        p1, s1, e1 = want[f]
        assert np.array_equal(nxt[f].cpu().numpy(), p1) and np.array_equal(st[f].cpu().numpy(), s1) and np.array_equal(er[f].cpu().numpy(), e1), f
    assert st.sum().item() > 0.9 * (NFR - 1) * NP
    klt.close()
