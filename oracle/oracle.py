"""ctypes binding of the CPU oracle (oracle/libuvip_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([('x', 'f4'), ('y', 'f4'), ('size', 'f4'), ('angle', 'f4'), ('response', 'f4'),
                     ('octave', 'i4'), ('class_id', 'i4')])
assert KP_DTYPE.itemsize == 28


class Params(C.Structure):
    _fields_ = [('nfeatures', C.c_int), ('scale_factor', C.c_float), ('nlevels', C.c_int),
                ('score_type', C.c_int), ('fast_th', C.c_int), ('retry_th', C.c_int), ('cell', C.c_int)]


class SearchParams(C.Structure):
    _fields_ = [('mode', C.c_int), ('th_dist', C.c_int), ('ratio', C.c_float)]


def build(force=False):
    so = os.path.join(_HERE, 'libuvip_oracle.so')
    src = [os.path.join(_HERE, f) for f in ('uvip_oracle.c', 'uvip_oracle.h', 'orb_pattern.inc')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.uo_create.restype = C.c_void_p
        L.uo_create.argtypes = [C.POINTER(Params)]
        L.uo_destroy.argtypes = [C.c_void_p]
        L.uo_fast_atan2.restype = C.c_float
        L.uo_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.uo_ic_angle.restype = C.c_float
        L.uo_radius_by_viewing_cos.restype = C.c_float
        L.uo_radius_by_viewing_cos.argtypes = [C.c_float]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_params(nfeatures=1000, scale_factor=1.2, nlevels=8, score_type=1, fast_th=20, retry_th=7, cell=30):
    return Params(nfeatures, scale_factor, nlevels, score_type, fast_th, retry_th, cell)


class Extractor:
    """Mirror of USLAM::ORBextractor (include/ORBextractor.h:45-94) over the C oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, score_type=1, fast_th=20):
        self.params = make_params(nfeatures, scale_factor, nlevels, score_type, fast_th)
        self.h = C.c_void_p(lib().uo_create(C.byref(self.params)))
        self.nlevels = nlevels
        self.W = self.H = 0

    def __del__(self):
        try:
            lib().uo_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        n = self.nlevels
        sc = np.zeros(n, np.float32); inv = np.zeros(n, np.float32)
        quota = np.zeros(n, np.int32); umax = np.zeros(16, np.int32)
        lib().uo_tables(self.h, _p(sc), _p(inv), _p(quota), _p(umax))
        return sc, inv, quota, umax

    def level_size(self, W, H, level):
        w = C.c_int(); h = C.c_int()
        lib().uo_level_size(self.h, W, H, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def __call__(self, image, keypoints=None, grid=None, min_px_dist=1, full_detect=True, num_needed=0, cap=None):
        """returns (keypoints[KP_DTYPE], descriptors[n,32]); grid (column-major int32 2-D, Fortran order) is updated in place"""
        img = np.ascontiguousarray(image, np.uint8)
        H, W = img.shape
        self.W, self.H = W, H
        n_in = 0 if keypoints is None else len(keypoints)
        if cap is None:
            cap = 4 * self.params.nfeatures + n_in + 4096
        kps = np.zeros(cap, KP_DTYPE)
        if n_in:
            kps[:n_in] = keypoints
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(n_in)
        gr = gc = 0
        gp = None
        if grid is not None:
            assert grid.dtype == np.int32 and grid.flags.f_contiguous
            gr, gc = grid.shape
            gp = _p(grid)
        rc = lib().uo_extract(self.h, _p(img), W, H, W, _p(kps), C.byref(n), cap, _p(desc), gp, gr, gc,
                              int(min_px_dist), int(bool(full_detect)), int(num_needed))
        if rc != 0:
            raise RuntimeError('uo_extract failed: %d' % rc)
        return kps[:n.value].copy(), desc[:n.value].copy()

    def level(self, l, blurred=False):
        w, h = self.level_size(self.W, self.H, l)
        out = np.zeros((h, w), np.uint8)
        rc = lib().uo_get_level(self.h, l, int(blurred), _p(out), w)
        assert rc == 0
        return out

    def padded_level(self, l):
        w, h = self.level_size(self.W, self.H, l)
        out = np.zeros((h + 32, w + 32), np.uint8)
        assert lib().uo_get_padded_level(self.h, l, _p(out)) == 0
        return out

    def raw_corners(self, l):
        cap = 1 << 20
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = lib().uo_get_raw_corners(self.h, l, _p(xs), _p(ys), _p(sc), cap)
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def level_keypoints(self, l):
        cap = 1 << 16
        out = np.zeros(cap, KP_DTYPE)
        n = lib().uo_get_level_keypoints(self.h, l, _p(out), cap)
        return out[:n].copy()


def extract_batch(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, fast_th=20, threads=0, cap=None):
    frames = np.ascontiguousarray(frames, np.uint8)
    nf, H, W = frames.shape
    if cap is None:
        cap = 2 * nfeatures + 512
    p = make_params(nfeatures, scale_factor, nlevels, 1, fast_th)
    kps = np.zeros((nf, cap), KP_DTYPE); desc = np.zeros((nf, cap, 32), np.uint8); n = np.zeros(nf, np.int32)
    rc = lib().uo_extract_batch(C.byref(p), _p(frames), nf, W, H, _p(kps), _p(n), cap, _p(desc), int(threads))
    if rc != 0:
        raise RuntimeError('uo_extract_batch failed: %d' % rc)
    return kps, n, desc


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    sh, sw = src.shape
    dst = np.zeros((dh, dw), np.uint8)
    lib().uo_resize_linear_u8(_p(src), sw, sh, sw, _p(dst), dw, dh, dw)
    return dst


def border_reflect101(img, pad=16):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    buf = np.zeros((h + 2 * pad, w + 2 * pad), np.uint8)
    buf[pad:pad + h, pad:pad + w] = img
    lib().uo_border_reflect101(_p(buf), w, h, w + 2 * pad, pad)
    return buf


def fast9(img, th, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = w * h
    xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
    n = lib().uo_fast9(_p(img), w, w, h, int(th), int(nms), _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def blur7(img):
    """7x7 sigma-2 blur with reflect-101 border of a standalone image."""
    pad = border_reflect101(img, 3)
    h, w = img.shape
    out = np.zeros((h, w), np.uint8)
    inner = pad[3:, 3:]
    lib().uo_blur7(C.c_void_p(pad.ctypes.data + 3 * (w + 6) + 3), w, h, w + 6, _p(out), w)
    return out


def fast_atan2(y, x):
    return float(lib().uo_fast_atan2(float(y), float(x)))


def ic_angle(padded, x, y, umax):
    """padded: 2-D uint8 with enough margin; (x, y) in padded coordinates."""
    padded = np.ascontiguousarray(padded, np.uint8)
    umax = np.ascontiguousarray(umax, np.int32)
    stride = padded.shape[1]
    return float(lib().uo_ic_angle(C.c_void_p(padded.ctypes.data + y * stride + x), stride, _p(umax)))


def descriptor(padded, x, y, angle_deg):
    padded = np.ascontiguousarray(padded, np.uint8)
    stride = padded.shape[1]
    d = np.zeros(32, np.uint8)
    lib().uo_descriptor(C.c_void_p(padded.ctypes.data + y * stride + x), stride, C.c_float(angle_deg), _p(d))
    return d


def harris_responses(level_img, xs, ys, block_size=7, k=0.04):
    """HarrisResponses (src/ORBextractor.cc:80-121) on a level image (points at least block_size//2 + 2 px from its edge)"""
    img = np.ascontiguousarray(level_img, np.uint8)
    xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
    out = np.zeros(len(xs), np.float32)
    lib().uo_harris_responses(_p(img), img.shape[1], _p(xs), _p(ys), len(xs), int(block_size), C.c_float(k), _p(out))
    return out


def distribute_octtree(x, y, resp, minX, maxX, minY, maxY, N):
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    n = len(x)
    out = np.zeros(n + 8, np.int32)
    m = lib().uo_distribute_octtree(_p(x), _p(y), _p(resp), n, minX, maxX, minY, maxY, N, _p(out), n + 8)
    return out[:m].copy()


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(lib().uo_descriptor_distance(_p(a), _p(b)))


def distinctive_descriptors(desc, start):
    """MapPoint::ComputeDistinctiveDescriptors over ragged lists (rows start[p]..start[p+1]-1) -> (best_idx, best_median)"""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
    n = len(start) - 1
    bi = np.zeros(n, np.int32); bm = np.zeros(n, np.int32)
    lib().uo_distinctive_descriptors(_p(desc), _p(start), n, _p(bi), _p(bm))
    return bi, bm


def knn2(q, t, threads=0):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nt = len(q), len(t)
    idx = np.zeros((nq, 2), np.int32); dist = np.zeros((nq, 2), np.int32)
    lib().uo_knn2(_p(q), nq, _p(t), nt, _p(idx), _p(dist), int(threads))
    return idx, dist


def ratio_filter(idx, dist, ratio):
    nq = len(idx)
    m = np.zeros(nq, np.int32)
    idx = np.ascontiguousarray(idx, np.int32); dist = np.ascontiguousarray(dist, np.int32)
    lib().uo_ratio_filter(_p(idx), _p(dist), nq, C.c_double(ratio), _p(m))
    return m


def rot_hist_filter(match, angle_a, angle_b):
    m = np.ascontiguousarray(match, np.int32).copy()
    a = np.ascontiguousarray(angle_a, np.float32); b = np.ascontiguousarray(angle_b, np.float32)
    lib().uo_rot_hist_filter(_p(m), len(m), _p(a), _p(b))
    return m


def grid_build(kx, ky, minX, minY, inv_w, inv_h, cols=64, rows=48):
    kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32)
    n = len(kx)
    start = np.zeros(cols * rows + 1, np.int32); items = np.zeros(max(n, 1), np.int32)
    lib().uo_grid_build(_p(kx), _p(ky), n, C.c_float(minX), C.c_float(minY), C.c_float(inv_w), C.c_float(inv_h),
                        cols, rows, _p(start), _p(items))
    return start, items[:start[-1]].copy()


def features_in_area(kx, ky, octave, start, items, minX, minY, inv_w, inv_h, x, y, r, minL, maxL, cols=64, rows=48):
    kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32)
    octave = np.ascontiguousarray(octave, np.int32)
    out = np.zeros(len(kx) + 1, np.int32)
    n = lib().uo_features_in_area(_p(kx), _p(ky), _p(octave), _p(start), _p(items), C.c_float(minX), C.c_float(minY),
                                  C.c_float(inv_w), C.c_float(inv_h), cols, rows, C.c_float(x), C.c_float(y),
                                  C.c_float(r), int(minL), int(maxL), _p(out), len(out))
    return out[:n].copy()


def search_window(mode, th_dist, ratio, qu, qv, qr, qminL, qmaxL, qdesc, kx, ky, octave, kdesc, start, items,
                  minX, minY, inv_w, inv_h, taken=None, cols=64, rows=48):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    qu, qv, qr, kx, ky = f32(qu), f32(qv), f32(qr), f32(kx), f32(ky)
    qminL, qmaxL, octave = i32(qminL), i32(qmaxL), i32(octave)
    qdesc = np.ascontiguousarray(qdesc, np.uint8); kdesc = np.ascontiguousarray(kdesc, np.uint8)
    nq, nk = len(qu), len(kx)
    tk = np.full(nk, -1, np.int32) if taken is None else i32(taken).copy()
    match = np.zeros(nq, np.int32)
    sp = SearchParams(mode, th_dist, ratio)
    n = lib().uo_search_window(C.byref(sp), _p(qu), _p(qv), _p(qr), _p(qminL), _p(qmaxL), _p(qdesc), nq,
                               _p(kx), _p(ky), _p(octave), _p(kdesc), nk, _p(i32(start)), _p(i32(items)),
                               C.c_float(minX), C.c_float(minY), C.c_float(inv_w), C.c_float(inv_h), cols, rows,
                               _p(tk), _p(match))
    return n, match, tk


def search_lists(mode, th_dist, ratio, qdesc, cand_start, cand_idx, kdesc, taken=None):
    qdesc = np.ascontiguousarray(qdesc, np.uint8); kdesc = np.ascontiguousarray(kdesc, np.uint8)
    cs = np.ascontiguousarray(cand_start, np.int32); ci = np.ascontiguousarray(cand_idx, np.int32)
    nq, nk = len(qdesc), len(kdesc)
    tk = np.full(nk, -1, np.int32) if taken is None else np.ascontiguousarray(taken, np.int32).copy()
    match = np.zeros(nq, np.int32)
    n = lib().uo_search_lists(int(mode), int(th_dist), C.c_float(ratio), _p(qdesc), nq, _p(cs), _p(ci), _p(kdesc), nk, _p(tk), _p(match))
    return n, match, tk


def search_lists_epipolar(th_dist, qdesc, qline, cand_start, cand_idx, kdesc, kx, ky, kthr, taken=None):
    qdesc = np.ascontiguousarray(qdesc, np.uint8); kdesc = np.ascontiguousarray(kdesc, np.uint8)
    ql = np.ascontiguousarray(qline, np.float32); cs = np.ascontiguousarray(cand_start, np.int32); ci = np.ascontiguousarray(cand_idx, np.int32)
    kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32); kt = np.ascontiguousarray(kthr, np.float64)
    nq, nk = len(qdesc), len(kdesc)
    tk = np.full(nk, -1, np.int32) if taken is None else np.ascontiguousarray(taken, np.int32).copy()
    match = np.zeros(nq, np.int32)
    n = lib().uo_search_lists_epipolar(int(th_dist), _p(qdesc), _p(ql), nq, _p(cs), _p(ci), _p(kdesc), _p(kx), _p(ky), _p(kt), nk, _p(tk), _p(match))
    return n, match, tk


def haloc_hash(desc, proj):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); proj = np.ascontiguousarray(proj, np.float32)
    out = np.zeros(proj.shape[0] * 32, np.float32)
    lib().uo_haloc_hash(_p(desc), len(desc), _p(proj), proj.shape[0], proj.shape[1], _p(out))
    return out


def haloc_match(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    L = lib(); L.uo_haloc_match.restype = C.c_float
    return float(L.uo_haloc_match(_p(a), _p(b), len(a)))


def clahe(img, clip_limit=4.0, tiles=(12, 12)):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros_like(img)
    lib().uo_clahe(_p(img), w, h, w, C.c_double(clip_limit), int(tiles[0]), int(tiles[1]), _p(out), w)
    return out


def bow_transform(voc, desc, levelsup=4):
    """voc: dict(child_start, child_ids, desc, weight, word, L) flat vocabulary tree"""
    desc = np.ascontiguousarray(desc, np.uint8)
    n = len(desc)
    wid = np.zeros(n, np.int32); nid = np.zeros(n, np.int32); w = np.zeros(n, np.float64)
    cs = np.ascontiguousarray(voc['child_start'], np.int32); ci = np.ascontiguousarray(voc['child_ids'], np.int32)
    nd = np.ascontiguousarray(voc['desc'], np.uint8); nw = np.ascontiguousarray(voc['weight'], np.float64)
    word = np.ascontiguousarray(voc['word'], np.int32)
    lib().uo_bow_transform(_p(cs), _p(ci), _p(nd), _p(nw), _p(word), int(voc['L']), _p(desc), n, int(levelsup), _p(wid), _p(nid), _p(w))
    return wid, nid, w


class LKPyramid:
    """cv::buildOpticalFlowPyramid(img, pyr, Size(win,win), max_level) with derivatives (FrameKTL.cc:76)"""

    def __init__(self, img, win=21, max_level=5):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        L = lib()
        L.uo_lk_pyramid_build.restype = C.c_void_p
        self.p = C.c_void_p(L.uo_lk_pyramid_build(_p(img), w, h, w, int(win), int(max_level)))
        self.win = win

    def __del__(self):
        try:
            lib().uo_lk_pyramid_free(self.p)
        except Exception:
            pass

    def levels(self):
        return lib().uo_lk_pyramid_levels(self.p)

    def level(self, l):
        w = C.c_int(); h = C.c_int()
        lib().uo_lk_pyramid_get(self.p, l, C.byref(w), C.byref(h), None, None)
        img = np.zeros((h.value, w.value), np.uint8); der = np.zeros((h.value, w.value, 2), np.int16)
        lib().uo_lk_pyramid_get(self.p, l, C.byref(w), C.byref(h), _p(img), _p(der))
        return img, der


def lk_track(pyr0, pyr1, prev_pts, next_pts, win=21, max_level=5, max_iter=30, eps=0.01, flags=12, min_eig_thr=1e-4):
    prev_pts = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    nxt = np.ascontiguousarray(next_pts, np.float32).reshape(-1, 2).copy()
    n = len(prev_pts)
    status = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    lib().uo_lk_track(pyr0.p, pyr1.p, _p(prev_pts), _p(nxt), n, int(win), int(max_level), int(max_iter), C.c_double(eps), int(flags),
                      C.c_double(min_eig_thr), _p(status), _p(err))
    return nxt, status, err


def ransac_fundamental(pts0, pts1, threshold=1.0, nhyp=1024, seed=0x5EED):
    """cv::findFundamentalMat(pts0, pts1, FM_RANSAC, threshold, ...) inlier mask (src/Tracking.cc:1062): (count, mask, F)"""
    p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2); p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    n = len(p0)
    mask = np.zeros(n, np.uint8); F = np.zeros(9, np.float64)
    L = lib()
    L.uo_ransac_fundamental.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_ulonglong, C.c_void_p, C.c_void_p]
    cnt = L.uo_ransac_fundamental(_p(p0), _p(p1), n, float(threshold), int(nhyp), int(seed), _p(mask), _p(F))
    return cnt, mask, F.reshape(3, 3)


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().uo_pyr_down(_p(img), w, h, w, _p(out), out.shape[1])
    return out


def scharr(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h, w, 2), np.int16)
    lib().uo_scharr(_p(img), w, h, w, _p(out))
    return out
