"""Generates tests/golden/cv2_golden.npz — known answers for the OpenCV-owned arithmetic of the hot
path, produced by the cv2 wheel in the BUILD container (cv2 4.13.0).  The reference itself has no
tests, fixtures or golden vectors, and cannot be built here, so these are the
only external pins the oracle has.  Inputs come from the deterministic generator in
u-vip-slam_b200/synth.py, so the fixture mostly stores OUTPUTS (or their SHA-256 at full size).

    python tools/gen_golden.py          # rewrites tests/golden/cv2_golden.npz
"""
import hashlib
import importlib.util
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('synth', os.path.join(ROOT, 'u-vip-slam_b200', 'synth.py'))
S = importlib.util.module_from_spec(spec)
spec.loader.exec_module(S)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def level_sizes(W, H, n=8):
    inv = [np.float32(1)]
    isf = np.float32(1.0 / np.float64(np.float32(1.2)))
    for _ in range(1, n):
        inv.append(np.float32(inv[-1] * isf))
    return [(int(np.rint(np.float32(W) * s)), int(np.rint(np.float32(H) * s))) for s in inv]


def klt_points(S, n, W, H):
    """deterministic sub-pixel points, a few of them close to / outside the image border"""
    k = np.arange(1, n + 1, dtype=np.uint64)
    x = (S.draw(71, k) % np.uint64((W + 20) * 100)).astype(np.float32) / np.float32(100) - np.float32(10)
    y = (S.draw(72, k) % np.uint64((H + 20) * 100)).astype(np.float32) / np.float32(100) - np.float32(10)
    return np.stack([x, y], 1).astype(np.float32)


def main():
    cv2.setNumThreads(1)
    G = {'cv2_version': np.array(cv2.__version__)}
    # ---- generator pin
    for (seed, W, H) in [(1, 752, 480), (1000, 640, 512), (5, 320, 240)]:
        G['synth_sha_%d_%dx%d' % (seed, W, H)] = np.array(sha(S.synth_frame(seed, W, H)))
    G['synth_sha_twin'] = np.array(sha(S.synth_frame(1, 752, 480, dx=5, dy=3, noise_seed=2)))
    # ---- resize: explicit small cases
    src = S.synth_frame(11, 160, 120)
    sizes = [(133, 100), (111, 83), (150, 113), (80, 60), (159, 119), (40, 31), (160, 120), (97, 120)]
    G['resize_sizes'] = np.array(sizes, np.int32)
    for i, (w, h) in enumerate(sizes):
        G['resize_out_%d' % i] = cv2.resize(src, (w, h), interpolation=cv2.INTER_LINEAR)
    # ---- full-size pyramid / border / blur / FAST hashes for the three benchmark shapes
    for (seed, W, H) in [(1, 752, 480), (1000, 640, 512), (100000, 1280, 1024)]:
        img = S.synth_frame(seed, W, H)
        cur = img
        tag = '%dx%d' % (W, H)
        lv_sha, bd_sha, bl_sha, f20, f7 = [], [], [], [], []
        for l, (w, h) in enumerate(level_sizes(W, H)):
            if l > 0:
                cur = cv2.resize(cur, (w, h), interpolation=cv2.INTER_LINEAR)
            lv_sha.append(sha(cur))
            bd_sha.append(sha(cv2.copyMakeBorder(cur, 16, 16, 16, 16, cv2.BORDER_REFLECT_101)))
            bl_sha.append(sha(cv2.GaussianBlur(cur, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)))
            for th, acc in ((20, f20), (7, f7)):
                det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                     type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                kp = det.detect(cur)
                arr = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kp], np.int32).reshape(-1, 3)
                acc.append('%d:%s' % (len(arr), sha(arr)))
        G['pyr_sha_' + tag] = np.array(lv_sha)
        G['border_sha_' + tag] = np.array(bd_sha)
        G['blur_sha_' + tag] = np.array(bl_sha)
        G['fast20_sha_' + tag] = np.array(f20)
        G['fast7_sha_' + tag] = np.array(f7)
    # ---- FAST explicit small ROIs (incl. cell-sized 37x37 windows)
    img = S.synth_frame(7, 320, 240)
    rois = [(10, 20, 37, 37), (100, 50, 37, 38), (200, 100, 7, 7), (5, 5, 8, 30), (150, 150, 60, 45), (0, 0, 320, 240)]
    G['fast_rois'] = np.array(rois, np.int32)
    for i, (x0, y0, w, h) in enumerate(rois):
        for th in (20, 7, 1):
            det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                                 type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            kp = det.detect(np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w]))
            G['fast_out_%d_th%d' % (i, th)] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kp],
                                                       np.int32).reshape(-1, 3)
    noise = (S.draw(99, np.arange(1, 64 * 48 + 1, dtype=np.uint64)) % np.uint64(256)).astype(np.uint8).reshape(48, 64)
    det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
    G['fast_noise_th20'] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(noise)],
                                    np.int32).reshape(-1, 3)
    # ---- blur explicit
    small = S.synth_frame(13, 64, 48)
    G['blur_small'] = cv2.GaussianBlur(small, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    G['blur_noise'] = cv2.GaussianBlur(noise, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    # ---- fastAtan2
    k = np.arange(1, 4097, dtype=np.uint64)
    ys = ((S.draw(21, k) % np.uint64(1 << 24)).astype(np.int64) - (1 << 23)).astype(np.float32)
    xs = ((S.draw(22, k) % np.uint64(1 << 24)).astype(np.int64) - (1 << 23)).astype(np.float32)
    ys[:8] = [0, 0, 1, -1, 5, -5, 0, 7]
    xs[:8] = [0, 1, 0, 0, 5, -5, -3, -7]
    G['atan2_y'] = ys
    G['atan2_x'] = xs
    G['atan2_out'] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    # ---- brute-force kNN k=2 (ties included)
    q = S.random_descriptors(31, 200)
    t = S.random_descriptors(32, 2000)
    t[100] = t[7]; t[300] = q[5]; t[301] = q[5]; t[1999] = q[9]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    G['knn_idx'] = np.array([[mm[0].trainIdx, mm[1].trainIdx] for mm in m], np.int32)
    G['knn_dist'] = np.array([[int(mm[0].distance), int(mm[1].distance)] for mm in m], np.int32)
    # ---- IC angle + rotated-BRIEF end to end through cv2.ORB (level-0 keypoints)
    img = S.synth_frame(5, 320, 240)
    orb = cv2.ORB_create(nfeatures=400, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_FAST_SCORE, patchSize=31, fastThreshold=20)
    kps = [k for k in orb.detect(img) if k.octave == 0]
    kps2, desc = orb.compute(img, kps)
    G['orb_xy'] = np.array([(int(k.pt[0]), int(k.pt[1])) for k in kps2], np.int32)
    G['orb_angle'] = np.array([k.angle for k in kps2], np.float32)
    G['orb_desc'] = desc
    # cv2.ORB blurs its pyramid sub-matrix through sepFilter2D with the float Gaussian kernel ("variant C",
    # not the fixed-point kernel GaussianBlur applies to a standalone Mat); store that blurred plane so the
    # descriptor arithmetic can be pinned independently of the blur variant.
    g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    G['orb_blurred_C'] = cv2.sepFilter2D(img, -1, g, g, borderType=cv2.BORDER_REFLECT_101)
    # ---- CLAHE (next row N3): cv::createCLAHE(4, (12,12)) of src/Tracking.cc:425-431, plus a small explicit case
    for (seed, W, H) in [(1, 752, 480), (1000, 640, 512), (100000, 1280, 1024)]:
        G['clahe_sha_%dx%d' % (W, H)] = np.array(sha(cv2.createCLAHE(4.0, (12, 12)).apply(S.synth_frame(seed, W, H))))
    G['clahe_small'] = cv2.createCLAHE(2.0, (4, 3)).apply(S.synth_frame(17, 97, 61))
    # ---- KLT (next row N1): cv::calcOpticalFlowPyrLK on the config-1 frame pair with the reference's parameters
    a = S.synth_frame(1, 752, 480); b = S.synth_frame(1, 752, 480, dx=5, dy=3, noise_seed=2)
    G['klt_pyrdown_sha'] = np.array(sha(cv2.pyrDown(a)))
    G['klt_scharr_sha'] = np.array(sha(np.stack([cv2.Scharr(a, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101),
                                                 cv2.Scharr(a, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101)], 2)))
    p0 = klt_points(S, 500, 752, 480)
    for win, lev in ((21, 5), (9, 3)):
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
        p1, st, err = cv2.calcOpticalFlowPyrLK(a, b, p0.copy(), (p0 + np.float32([2.0, 1.5])).copy(), winSize=(win, win), maxLevel=lev, criteria=crit,
                                               flags=cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS)
        G['klt_p1_w%d' % win] = p1.astype(np.float32); G['klt_st_w%d' % win] = st.ravel().astype(np.uint8); G['klt_err_w%d' % win] = err.ravel().astype(np.float32)
    out = os.path.join(ROOT, 'tests', 'golden', 'cv2_golden.npz')
    np.savez_compressed(out, **G)
    print('wrote', out, os.path.getsize(out), 'bytes;', len(G), 'entries')


if __name__ == '__main__':
    sys.exit(main())
