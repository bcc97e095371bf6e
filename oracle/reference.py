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
