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
