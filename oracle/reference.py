"""ctypes binding of oracle/_ref/libref_orbextractor.so: the reference's OWN USLAM::ORBextractor (compiled from
/root/reference/src/ORBextractor.cc where it lies, against the stand-in OpenCV/Eigen/ROS headers of oracle/ref_shim/).
TEST INFRASTRUCTURE ONLY — used by tests/ to pin the C oracle (and through it the CUDA path) against the reference's
real code.  The library is built by `make -C oracle ref` in the build container (where /root/reference exists) and
travels to the GPU box as a prebuilt file; nothing here reads /root/reference at run time."""
import ctypes as C
import os
import subprocess
import numpy as np

from .oracle import KP_DTYPE, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, '_ref', 'libref_orbextractor.so')
REF_SRC = '/root/reference/src/ORBextractor.cc'
_LIB = None


def build(force=False):
    """(re)build oracle/_ref when the reference checkout is present; returns the path or None when unavailable"""
    if os.path.exists(REF_SRC):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'ref'] + (['-B'] if force else []))
    return SO if os.path.exists(SO) else None


def available():
    return os.path.exists(SO) or os.path.exists(REF_SRC)


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError('oracle/_ref is not built and /root/reference is absent')
        _LIB = C.CDLL(so)
        _LIB.ref_create.restype = C.c_void_p
        _LIB.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.ref_destroy.argtypes = [C.c_void_p]
        _LIB.ref_scale_factor.restype = C.c_float
        _LIB.ref_scale_factor.argtypes = [C.c_void_p]
        _LIB.ref_levels.argtypes = [C.c_void_p]
    return _LIB


class Extractor:
    """USLAM::ORBextractor of the reference (include/ORBextractor.h:45-94), same call shape as oracle.Extractor"""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, score_type=1, fast_th=20):
        self.nfeatures = nfeatures
        self.h = C.c_void_p(lib().ref_create(nfeatures, scale_factor, nlevels, score_type, fast_th))

    def __del__(self):
        try:
            lib().ref_destroy(self.h)
        except Exception:
            pass

    def GetLevels(self):
        return lib().ref_levels(self.h)

    def GetScaleFactor(self):
        return lib().ref_scale_factor(self.h)

    def __call__(self, image, keypoints=None, grid=None, min_px_dist=1, full_detect=True, num_needed=0, cap=None, arena_mb=1024):
        img = np.ascontiguousarray(image, np.uint8)
        H, W = img.shape if img.ndim == 2 else (0, 0)
        n_in = 0 if keypoints is None else len(keypoints)
        if cap is None:
            cap = 4 * self.nfeatures + n_in + 4096
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
        rc = lib().ref_extract(self.h, _p(img), W, H, W, _p(kps), C.byref(n), cap, _p(desc), gp, gr, gc, int(min_px_dist),
                               int(bool(full_detect)), int(num_needed), int(arena_mb))
        if rc != 0:
            raise RuntimeError('ref_extract failed: %d' % rc)
        return kps[:n.value].copy(), desc[:n.value].copy()


def dead_path_keypoints(extractor, image, level, cap=20000):
    """keypoints the reference's dead ComputeKeyPoints path (src/ORBextractor.cc:536-746) leaves on `level` (HarrisResponses as
    response when the extractor was created with score_type 0)"""
    img = np.ascontiguousarray(image, np.uint8)
    H, W = img.shape
    out = np.zeros(cap, KP_DTYPE)
    n = lib().ref_dead_path_keypoints(extractor.h, _p(img), W, H, W, int(level), _p(out), cap)
    if n < 0 or n > cap:
        raise RuntimeError('ref_dead_path_keypoints: %d' % n)
    return out[:n].copy()


def extract_batch(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, fast_th=20, threads=0, cap=None, arena_mb=256):
    """frame-parallel batch over the reference extractor (one instance per OpenMP thread), same shape as oracle.extract_batch.
    arena_mb >= 0 (default): per-thread bump arena = pinned quadtree tie-break, and also the fastest way to run it (glibc
    malloc serialises the reference's many large short-lived vectors across threads: measured 35 vs 233 frames/s on 8 cores)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    nf, H, W = frames.shape
    if cap is None:
        cap = 2 * nfeatures + 512
    kps = np.zeros((nf, cap), KP_DTYPE); desc = np.zeros((nf, cap, 32), np.uint8); n = np.zeros(nf, np.int32)
    L = lib()
    L.ref_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    rc = L.ref_extract_batch(nfeatures, scale_factor, nlevels, 1, fast_th, _p(frames), nf, W, H, _p(kps), _p(n), cap, _p(desc),
                             int(threads) if threads and threads > 0 else (os.cpu_count() or 1), int(arena_mb))
    if rc != 0:
        raise RuntimeError('ref_extract_batch failed: %d' % rc)
    return kps, n, desc


# ------------------------------------------------------------------------------------------------ matcher
MATCHER_SO = os.path.join(_HERE, '_ref', 'libref_orbmatcher.so')
_MLIB = None


def matcher_available():
    return os.path.exists(MATCHER_SO) or os.path.exists('/root/reference/src/ORBmatcher.cc')


def mlib():
    """oracle/_ref/libref_orbmatcher.so: the reference's own src/ORBmatcher.cc over stand-in FrameKTL/KeyFrame/MapPoint types"""
    global _MLIB
    if _MLIB is None:
        build()
        if not os.path.exists(MATCHER_SO):
            raise RuntimeError('oracle/_ref/libref_orbmatcher.so is not built and /root/reference is absent')
        _MLIB = C.CDLL(MATCHER_SO)
        _MLIB.refm_last_call_seconds.restype = C.c_double
    return _MLIB


def last_call_seconds():
    """wall time of the reference member function inside the last matcher entry point (scene construction excluded)"""
    return float(mlib().refm_last_call_seconds())


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def descriptor_distance(a, b):
    return int(mlib().refm_descriptor_distance(_p(_u8(a)), _p(_u8(b))))


def search_by_projection_mps(kx, ky, octave, kdesc, bounds, scale_factors, u, v, level, view_cos, qdesc, th, nnratio, taken=None,
                             in_view=None, bad=None):
    """ORBmatcher::SearchByProjection(FrameKTL&, vector<MapPoint*>&, th) -> (nmatches, owner[nk])"""
    kx, ky, octave, kdesc = _f32(kx), _f32(ky), _i32(octave), _u8(kdesc)
    nk, nq = len(kx), len(u)
    owner = np.zeros(nk, np.int32)
    tk = None if taken is None else _i32(taken)
    iv = None if in_view is None else _u8(in_view)
    bd = None if bad is None else _u8(bad)
    sf = _f32(scale_factors)
    n = mlib().refm_search_by_projection_mps(nk, _p(kx), _p(ky), _p(octave), _p(kdesc), _p(_f32(bounds)), len(sf), _p(sf), _p(tk),
                                             nq, _p(_f32(u)), _p(_f32(v)), _p(_i32(level)), _p(_f32(view_cos)), _p(iv), _p(bd), _p(_u8(qdesc)),
                                             C.c_float(th), C.c_float(nnratio), _p(owner))
    return n, owner


def search_by_projection_kf(kx, ky, octave, kangle, kdesc, bounds, scale_factors, Tcw, intr, has_mp, bad, found, pos, min_dist, pdesc, pangle,
                            th, orb_dist, nnratio, check_ori=True, taken=None):
    """ORBmatcher::SearchByProjection(FrameKTL&, KeyFrame*, sAlreadyFound, th, ORBdist) -> (nmatches, owner[nk])"""
    kx, ky = _f32(kx), _f32(ky)
    nk, npnt = len(kx), len(has_mp)
    owner = np.zeros(nk, np.int32)
    tk = None if taken is None else _i32(taken)
    sf = _f32(scale_factors)
    n = mlib().refm_search_by_projection_kf(nk, _p(kx), _p(ky), _p(_i32(octave)), _p(_f32(kangle)), _p(_u8(kdesc)), _p(_f32(bounds)), len(sf), _p(sf),
                                            _p(tk), _p(_f32(Tcw)), _p(_f32(intr)), npnt, _p(_u8(has_mp)), _p(_u8(bad)), _p(_u8(found)), _p(_f32(pos)),
                                            _p(_f32(min_dist)), _p(_u8(pdesc)), _p(_f32(pangle)), C.c_float(th), int(orb_dist), C.c_float(nnratio),
                                            int(bool(check_ori)), _p(owner))
    return n, owner


def _featvec(fv):
    """dict node id -> list of feature indices  ->  (ids, start, idx) flat arrays in ascending node order"""
    ids = sorted(fv)
    start = np.zeros(len(ids) + 1, np.int32)
    idx = []
    for i, k in enumerate(ids):
        idx.extend(fv[k]); start[i + 1] = len(idx)
    return _i32(ids), start, _i32(idx if idx else [0])


def search_by_bow_kf_frame(kf_desc, kf_angle, has_mp, bad, kf_featvec, f_desc, f_angle, f_featvec, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, FrameKTL&, matches) -> (nmatches, keyframe slot per frame keypoint or -1)"""
    ki, ks, kx = _featvec(kf_featvec); fi, fs, fx = _featvec(f_featvec)
    nk = len(f_desc)
    out = np.zeros(nk, np.int32)
    n = mlib().refm_search_by_bow_kf_frame(len(kf_desc), _p(_u8(kf_desc)), _p(_f32(kf_angle)), _p(_u8(has_mp)), _p(_u8(bad)), len(ki), _p(ki), _p(ks), _p(kx),
                                           nk, _p(_u8(f_desc)), _p(_f32(f_angle)), len(fi), _p(fi), _p(fs), _p(fx), C.c_float(nnratio),
                                           int(bool(check_ori)), _p(out))
    return n, out


def search_by_bow_kf_kf(desc1, angle1, has1, bad1, fv1, desc2, angle2, has2, bad2, fv2, nnratio, check_ori=True):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, matches12) -> (nmatches, slot of keyframe 2 per slot of keyframe 1 or -1)"""
    i1, s1, x1 = _featvec(fv1); i2, s2, x2 = _featvec(fv2)
    out = np.zeros(len(desc1), np.int32)
    n = mlib().refm_search_by_bow_kf_kf(len(desc1), _p(_u8(desc1)), _p(_f32(angle1)), _p(_u8(has1)), _p(_u8(bad1)), len(i1), _p(i1), _p(s1), _p(x1),
                                        len(desc2), _p(_u8(desc2)), _p(_f32(angle2)), _p(_u8(has2)), _p(_u8(bad2)), len(i2), _p(i2), _p(s2), _p(x2),
                                        C.c_float(nnratio), int(bool(check_ori)), _p(out))
    return n, out


def fuse(kx, ky, octave, kdesc, bounds, scale_factors, Rcw, tcw, Ow, intr, kf_has_mp, kf_bad, is_null, bad, in_kf, pos, normal, min_dist,
         max_dist, pdesc, th):
    """ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>&, th) -> (nFused, action[np], target[np]); see ref_matcher_driver.cpp"""
    nk, npnt = len(kx), len(is_null)
    sf = _f32(scale_factors)
    action = np.zeros(npnt, np.int32); target = np.zeros(npnt, np.int32)
    n = mlib().refm_fuse(nk, _p(_f32(kx)), _p(_f32(ky)), _p(_i32(octave)), _p(_u8(kdesc)), _p(_f32(bounds)), len(sf), _p(sf), _p(_f32(Rcw)),
                         _p(_f32(tcw)), _p(_f32(Ow)), _p(_f32(intr)), _p(_u8(kf_has_mp)), _p(_u8(kf_bad)), npnt, _p(_u8(is_null)), _p(_u8(bad)),
                         _p(_u8(in_kf)), _p(_f32(pos)), _p(_f32(normal)), _p(_f32(min_dist)), _p(_f32(max_dist)), _p(_u8(pdesc)), C.c_float(th),
                         _p(action), _p(target))
    return n, action, target


M8_EXE = os.path.join(_HERE, '_ref', 'test_shim_m8')


def write_bundle(path, arrays):
    """array bundle of oracle/ref_shim/m8_scene.h: int32 count, then int32 nbytes + payload per array"""
    import struct
    with open(path, 'wb') as f:
        f.write(struct.pack('i', len(arrays)))
        for a in arrays:
            b = np.ascontiguousarray(a).tobytes()
            f.write(struct.pack('i', len(b))); f.write(b)


def read_bundle(path):
    import struct
    buf = open(path, 'rb').read()
    n = struct.unpack_from('i', buf, 0)[0]; off = 4; out = []
    for _ in range(n):
        nb = struct.unpack_from('i', buf, off)[0]; off += 4
        out.append(np.frombuffer(buf, np.int32, nb // 4, off).copy()); off += nb
    return out


def m8_run(scene_path, out_path, which):
    """the reference's own Fuse / Fuse(Scw) / SearchByProjection(KF,Scw) / SearchBySim3 (which = 0..3) on a scene bundle"""
    r = mlib().refm_m8_run(scene_path.encode(), out_path.encode(), int(which))
    if r == -1000:
        raise RuntimeError('refm_m8_run: I/O failure')
    return r, read_bundle(out_path)


# ------------------------------------------------------------------------------------------------ haloc hash (next row N4)
HASH_SO = os.path.join(_HERE, '_ref', 'libref_hash.so')
_HLIB = None


def hash_available():
    return os.path.exists(HASH_SO) or os.path.exists('/root/reference/src/hash.cpp')


def hlib():
    global _HLIB
    if _HLIB is None:
        build()
        _HLIB = C.CDLL(HASH_SO)
        _HLIB.refh_create.restype = C.c_void_p
        _HLIB.refh_destroy.argtypes = [C.c_void_p]
        _HLIB.refh_get_hash.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _HLIB.refh_projection_length.argtypes = [C.c_void_p]
        _HLIB.refh_get_projections.argtypes = [C.c_void_p, C.c_void_p]
        _HLIB.refh_match.restype = C.c_float
        _HLIB.refh_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return _HLIB


class HalocHash:
    """the reference's own haloc::Hash (src/hash.cpp); its time(NULL)-seeded projection vectors are made reproducible by
    interposing time() inside the library and can be read out with projections()"""

    def __init__(self, num_proj=3):
        self.num_proj = num_proj
        self.h = C.c_void_p(hlib().refh_create(num_proj))

    def __del__(self):
        try:
            hlib().refh_destroy(self.h)
        except Exception:
            pass

    def get_hash(self, desc):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        out = np.zeros(self.num_proj * 32, np.float32)
        n = hlib().refh_get_hash(self.h, _p(desc), len(desc), _p(out))
        assert n == len(out), n
        return out

    def projections(self):
        n = hlib().refh_projection_length(self.h)
        out = np.zeros((self.num_proj, n), np.float32)
        hlib().refh_get_projections(self.h, _p(out))
        return out

    def match(self, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        return float(hlib().refh_match(self.h, _p(a), _p(b), len(a)))


# ------------------------------------------------------------------------------------------------ DBoW2 (next row N2)
DBOW_SO = os.path.join(_HERE, '_ref', 'libref_dbow.so')
_VLIB = None


def dbow_available():
    return os.path.exists(DBOW_SO) or os.path.exists('/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h')


def vlib():
    global _VLIB
    if _VLIB is None:
        build()
        _VLIB = C.CDLL(DBOW_SO)
        _VLIB.refv_load.restype = C.c_void_p
        _VLIB.refv_load.argtypes = [C.c_char_p]
        _VLIB.refv_destroy.argtypes = [C.c_void_p]
        _VLIB.refv_size.argtypes = [C.c_void_p]
    return _VLIB


class Vocabulary:
    """the reference's own ORBVocabulary (DBoW2 TemplatedVocabulary<FORB>) loaded from a text file"""

    def __init__(self, path):
        h = vlib().refv_load(str(path).encode())
        if not h:
            raise RuntimeError('loadFromTextFile failed: %s' % path)
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            vlib().refv_destroy(self.h)
        except Exception:
            pass

    def size(self):
        return vlib().refv_size(self.h)

    def transform(self, desc, levelsup=4):
        """-> (BowVector as {word: value}, FeatureVector as {node: [feature indices]})"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        bw = np.zeros(n + 1, np.int32); bv = np.zeros(n + 1, np.float64); fn = np.zeros(n + 1, np.int32); ff = np.zeros(n + 1, np.int32)
        nb = C.c_int(); nf = C.c_int()
        rc = vlib().refv_transform(self.h, _p(desc), n, int(levelsup), _p(bw), _p(bv), C.byref(nb), n + 1, _p(fn), _p(ff), C.byref(nf), n + 1)
        if rc != 0:
            raise RuntimeError('refv_transform: %d' % rc)
        bow = {int(w): float(v) for w, v in zip(bw[:nb.value], bv[:nb.value])}
        fv = {}
        for node, f in zip(fn[:nf.value], ff[:nf.value]):
            fv.setdefault(int(node), []).append(int(f))
        return bow, fv


# ------------------------------------------------------------------------------------------------ MapPoint (next row N4, first half)
MAPPOINT_SO = os.path.join(_HERE, '_ref', 'libref_mappoint.so')


def mappoint_available():
    return os.path.exists(MAPPOINT_SO) or os.path.exists('/root/reference/src/MapPoint.cc')


def distinctive_descriptors(desc, start):
    """the reference's real MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:197-270) on ragged observation lists:
    -> (chosen descriptor per list [n, 32], chosen flag per list)"""
    build()
    L = C.CDLL(MAPPOINT_SO)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
    n = len(start) - 1
    out = np.zeros((n, 32), np.uint8); chosen = np.zeros(n, np.int32)
    L.refp_distinctive_descriptors(_p(desc), _p(start), n, _p(out), _p(chosen))
    return out, chosen


# ------------------------------------------------------------------------------------------------ mini front-end (real FrameKTL)
FRONTEND_SO = os.path.join(_HERE, '_ref', 'libref_frontend.so')


def frontend_available():
    return os.path.exists(FRONTEND_SO) or os.path.exists('/root/reference/src/FrameKTL.cc')


def search_local_points(kxyoa, kdesc, bounds, intr, nlevels, scale_factor, Tcw, pos, obs_level, ref_Ow, pdesc, th, nnratio, cos_limit=0.5):
    """the reference's REAL FrameKTL + MapPoint + ORBmatcher replaying Tracking::SearchLocalPoints for one frame
    -> dict(n, inview, u, v, level, viewcos, owner, cell_start, cell_items)"""
    build()
    L = C.CDLL(FRONTEND_SO)
    kxyoa = np.ascontiguousarray(kxyoa, np.float32); kdesc = np.ascontiguousarray(kdesc, np.uint8)
    pos = np.ascontiguousarray(pos, np.float32); pdesc = np.ascontiguousarray(pdesc, np.uint8)
    nk, npnt = len(kxyoa), len(pos)
    out = dict(inview=np.zeros(npnt, np.int32), u=np.zeros(npnt, np.float32), v=np.zeros(npnt, np.float32), level=np.zeros(npnt, np.int32),
               viewcos=np.zeros(npnt, np.float32), owner=np.zeros(nk, np.int32), cell_start=np.zeros(64 * 48 + 1, np.int32),
               cell_items=np.zeros(max(nk, 1), np.int32))
    L.reff_search_local_points.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float] + [C.c_void_p] * 8
    out['n'] = L.reff_search_local_points(nk, _p(kxyoa), _p(kdesc), _p(_i32(bounds)), _p(_f32(intr)), int(nlevels), float(scale_factor), _p(_f32(Tcw)),
                                          npnt, _p(pos), _p(_i32(obs_level)), _p(_f32(ref_Ow)), _p(pdesc), float(th), float(nnratio), float(cos_limit),
                                          _p(out['inview']), _p(out['u']), _p(out['v']), _p(out['level']), _p(out['viewcos']), _p(out['owner']),
                                          _p(out['cell_start']), _p(out['cell_items']))
    out['cell_items'] = out['cell_items'][:out['cell_start'][-1]].copy()
    return out
