"""cv2.findFundamentalMat(FM_RANSAC, 1, 0.999) golden vectors for the N1 RANSAC row (src/Tracking.cc:1062) -> tests/golden/cv2_ransac.npz.
Synthetic two-view correspondences (deterministic SplitMix64 draws): 3-D points seen by two cameras, pixel noise, gross outliers.
   python tools/gen_golden_ransac.py          (needs the cv2 wheel of the build container)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def two_view_case(S, seed, n, outlier_pct, noise_px):
    """n correspondences in pixels: (1 - outlier_pct/100) of them consistent with one epipolar geometry up to `noise_px` (uniform), the
    rest uniformly random positions in the second image"""
    d = lambda k, m: (S.draw(seed, np.arange(n, dtype=np.uint64) * np.uint64(16) + np.uint64(k + 1)) % np.uint64(m)).astype(np.float64)
    W, H, fx, fy, cx, cy = 752.0, 480.0, 458.0, 457.0, 367.0, 248.0
    u0 = 20 + d(0, 712) + d(1, 1000) / 1000.0; v0 = 20 + d(2, 440) + d(3, 1000) / 1000.0
    z = 2.0 + d(4, 6000) / 1000.0
    X = np.stack([(u0 - cx) / fx * z, (v0 - cy) / fy * z, z], 1)
    a, b, c = 0.03, -0.05, 0.02
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    Rz = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx; t = np.array([0.25, -0.04, 0.06])
    Y = X @ R.T + t
    u1 = fx * Y[:, 0] / Y[:, 2] + cx + (d(5, 2001) / 1000.0 - 1.0) * noise_px
    v1 = fy * Y[:, 1] / Y[:, 2] + cy + (d(6, 2001) / 1000.0 - 1.0) * noise_px
    out = d(7, 100) < outlier_pct
    u1 = np.where(out, d(8, 752), u1); v1 = np.where(out, d(9, 480), v1)
    return np.stack([u0, v0], 1).astype(np.float32), np.stack([u1, v1], 1).astype(np.float32), out


CASES = [(101, 600, 25, 0.3), (102, 300, 40, 0.5), (103, 1000, 10, 0.2), (104, 40, 20, 0.3)]


def main():
    import cv2
    import __graft_entry__ as ge
    S = ge.load_package().synth
    G = {'cv2_version': np.array(cv2.__version__)}
    for seed, n, op, noise in CASES:
        p0, p1, out = two_view_case(S, seed, n, op, noise)
        F, mask = cv2.findFundamentalMat(p0, p1, cv2.FM_RANSAC, 1.0, 0.999)
        G['p0_%d' % seed] = p0; G['p1_%d' % seed] = p1; G['mask_%d' % seed] = mask.ravel().astype(np.uint8); G['F_%d' % seed] = F
        G['outlier_%d' % seed] = out.astype(np.uint8)
        print(seed, n, 'cv2 inliers', int(mask.sum()), 'true inliers', int((~out).sum()))
    path = os.path.join(ROOT, 'tests', 'golden', 'cv2_ransac.npz')
    np.savez_compressed(path, **G)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    sys.exit(main())
