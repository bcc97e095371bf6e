"""Host-side mirror of the two reference classes over the C-ABI, used by tests/, bench.py and smoke().

    ORBextractor  <->  USLAM::ORBextractor  (include/ORBextractor.h:45-94 of the reference)
    ORBmatcher    <->  USLAM::ORBmatcher    (include/ORBmatcher.h:41-94), descriptor path only

Same constructor arguments, same operator() argument meaning, same error behaviour (empty image -> outputs
untouched).  numpy arrays stand in for cv::Mat / std::vector<cv::KeyPoint> / Eigen::MatrixXi (Fortran-ordered
int32).  The C++ shim with the literal reference signatures is host/ORBextractor.h, host/ORBmatcher.h."""
import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE, ExtractorParams, SearchParams, check, lib, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, scoreType=0, fastTh=7,
                 device=0, max_width=1280, max_height=1024, max_batch=1, retry_th=0, cell=0):
        self.params = ExtractorParams(nfeatures, scaleFactor, nlevels, scoreType, fastTh, retry_th, cell, device,
                                      max_width, max_height, max_batch)
        self.h = C.c_void_p()
        check(lib().uvip_extractor_create(C.byref(self.params), C.byref(self.h)))
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            lib().uvip_extractor_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def GetLevels(self):
        return lib().uvip_extractor_levels(self.h)

    def GetScaleFactor(self):
        return lib().uvip_extractor_scale_factor(self.h)

    def tables(self):
        n = self.nlevels
        sc = np.zeros(n, np.float32); inv = np.zeros(n, np.float32)
        quota = np.zeros(n, np.int32); umax = np.zeros(16, np.int32)
        check(lib().uvip_extractor_tables(self.h, ptr(sc), ptr(inv), ptr(quota), ptr(umax)))
        return sc, inv, quota, umax

    def __call__(self, image, mask=None, keypoints=None, grid_2d=None, min_px_dist=1, FullDetect=True,
                 num_featsneeded=0, cap=None):
        """operator()(image, mask, keypoints, descriptors, grid_2d, min_px_dist, FullDetect, num_featsneeded).
        Returns (keypoints, descriptors); an empty image returns the inputs untouched: (keypoints, None)."""
        if image is None or image.size == 0:
            return keypoints, None
        if image.dtype != np.uint8 or image.ndim != 2:
            raise AssertionError('image.type() == CV_8UC1')          # src/ORBextractor.cc:856
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        H, W = image.shape
        n_in = 0 if keypoints is None else len(keypoints)
        if cap is None:      # incoming keypoints in steps of 256: cap is part of the single-frame call shape (CUDA-graph key)
            cap = self.nfeatures + 8 * self.nlevels + 64 + (n_in + 255) // 256 * 256
        kps = np.zeros(cap, KP_DTYPE)
        if n_in:
            kps[:n_in] = keypoints
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(n_in)
        gr = gc = 0
        if grid_2d is not None:
            assert grid_2d.dtype == np.int32 and grid_2d.flags.f_contiguous, 'Eigen::MatrixXi is column-major int32'
            gr, gc = grid_2d.shape
        check(lib().uvip_extract(self.h, ptr(image), W, H, image.strides[0], ptr(kps), C.byref(n), cap, ptr(desc),
                                 ptr(grid_2d), gr, gc, int(min_px_dist), int(bool(FullDetect)), int(num_featsneeded)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, frames, cap=None):
        frames = np.ascontiguousarray(frames, np.uint8)
        nf, H, W = frames.shape
        if cap is None:
            cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros((nf, cap), KP_DTYPE); desc = np.zeros((nf, cap, 32), np.uint8); n = np.zeros(nf, np.int32)
        check(lib().uvip_extract_batch(self.h, ptr(frames), nf, W, H, W, W * H, ptr(kps), ptr(n), cap, ptr(desc)))
        return kps, n, desc

    def extract_batch_submit(self, frames, kps, n, desc):
        """enqueue one batch (caller-owned, ideally pinned, arrays: frames (nf,H,W) u8, kps (nf,cap) KP_DTYPE, n (nf,) i32,
        desc (nf,cap,32) u8) and return its ticket; the arrays are filled when extract_batch_wait(ticket) returns"""
        nf, H, W = frames.shape
        cap = kps.shape[1]
        t = C.c_int(-1)
        check(lib().uvip_extract_batch_submit(self.h, ptr(frames), nf, W, H, W, W * H, ptr(kps), ptr(n), cap, ptr(desc), C.byref(t)))
        return t.value

    def extract_match_batch_submit(self, matcher, frames, kps, n, desc, knn_idx, knn_dist):
        """extract_batch_submit + consecutive-frame kNN2 on the device (descriptors are matched where the extractor wrote them);
        knn_idx / knn_dist: (nf - 1, cap, 2) int32, filled when extract_batch_wait(ticket) returns"""
        nf, H, W = frames.shape
        cap = kps.shape[1]
        t = C.c_int(-1)
        check(lib().uvip_extract_match_batch_submit(self.h, matcher.h, ptr(frames), nf, W, H, W, W * H, ptr(kps), ptr(n), cap, ptr(desc),
                                                    ptr(knn_idx), ptr(knn_dist), C.byref(t)))
        return t.value

    def extract_batch_wait(self, ticket):
        check(lib().uvip_extract_batch_wait(self.h, int(ticket)))

    def extract_batch_device(self, d_frames, nframes, W, H, d_kps, d_n, cap, d_desc, stream=0):
        """all pointers are integers (device addresses, e.g. torch tensor.data_ptr()); asynchronous"""
        check(lib().uvip_extract_batch_device(self.h, ptr(d_frames), nframes, W, H, W, W * H, ptr(d_kps), ptr(d_n), cap,
                                              ptr(d_desc), C.c_void_p(stream) if stream else None))

    def clahe(self, image, clip_limit=4.0, tiles=(12, 12)):
        """cv::createCLAHE(clip_limit, tiles)->apply(im, im), the `Enhance` pre-processing of src/Tracking.cc:425-431"""
        image = np.ascontiguousarray(image, np.uint8)
        H, W = image.shape
        out = np.zeros_like(image)
        check(lib().uvip_clahe(self.h, ptr(image), W, H, W, float(clip_limit), int(tiles[0]), int(tiles[1]), ptr(out), W))
        return out

    def status(self):
        check(lib().uvip_extractor_status(self.h))

    def launch_count(self):
        return int(lib().uvip_extractor_launch_count(self.h))

    def graph_captures(self):
        return int(lib().uvip_extractor_graph_captures(self.h))

    STAGES = ('pyramid', 'fast', 'quadtree', 'blur', 'select', 'describe')

    def profile(self, enable=True):
        check(lib().uvip_extractor_profile(self.h, int(enable)))

    def stage_ms(self):
        ms = np.zeros(len(self.STAGES), np.float32); n = C.c_int()
        check(lib().uvip_extractor_stage_ms(self.h, ptr(ms), C.byref(n)))
        return dict(zip(self.STAGES, ms.tolist())), n.value

    # ---- debug taps
    def level(self, l, blurred=False, frame=0):
        w = C.c_int(); h = C.c_int()
        check(lib().uvip_get_pyramid_level(self.h, frame, l, int(blurred), None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(lib().uvip_get_pyramid_level(self.h, frame, l, int(blurred), ptr(out), w.value, C.byref(w), C.byref(h)))
        return out

    def _list(self, fn, l, frame):
        cap = 1 << 20
        xs = np.zeros(cap, np.int32); ys = np.zeros(cap, np.int32); sc = np.zeros(cap, np.int32)
        n = C.c_int()
        check(fn(self.h, frame, l, ptr(xs), ptr(ys), ptr(sc), cap, C.byref(n)))
        return xs[:n.value].copy(), ys[:n.value].copy(), sc[:n.value].copy()

    def raw_corners(self, l, frame=0):
        return self._list(lib().uvip_get_raw_corners, l, frame)

    def level_keypoints(self, l, frame=0):
        return self._list(lib().uvip_get_level_keypoints, l, frame)

    def compute_keypoints_quota(self, frame=0, cap_per_level=4096):
        """optional mode: the reference's dead ComputeKeyPoints path (src/ORBextractor.cc:536-746) on the last extracted frame
        -> list of per-level keypoint arrays (level coordinates)"""
        nl = self.GetLevels()
        kps = np.zeros((nl, cap_per_level), KP_DTYPE); n = np.zeros(nl, np.int32)
        check(lib().uvip_compute_keypoints_quota(self.h, int(frame), ptr(kps), ptr(n), int(cap_per_level)))
        return [kps[l, :n[l]].copy() for l in range(nl)]

    def harris_responses(self, l, xs, ys, block_size=7, k=0.04, frame=0):
        """HarrisResponses (src/ORBextractor.cc:80-121) at points of pyramid level l of the last extracted frame"""
        xs = np.ascontiguousarray(xs, np.float32); ys = np.ascontiguousarray(ys, np.float32)
        out = np.zeros(len(xs), np.float32)
        check(lib().uvip_harris_responses(self.h, int(frame), int(l), ptr(xs), ptr(ys), len(xs), int(block_size), float(k), ptr(out)))
        return out


class ORBmatcher:
    TH_HIGH = 100
    TH_LOW = 50
    HISTO_LENGTH = 30
    FRAME_GRID_COLS = 64
    FRAME_GRID_ROWS = 48

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.h = C.c_void_p()
        check(lib().uvip_matcher_create(device, C.byref(self.h)))

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            lib().uvip_matcher_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return int(lib().uvip_matcher_launch_count(self.h))

    def DescriptorDistance(self, a, b):
        """static int DescriptorDistance(const cv::Mat&, const cv::Mat&): one pair, or row-wise for 2-D inputs"""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        check(lib().uvip_descriptor_distance(self.h, ptr(a), ptr(b), len(a), ptr(out)))
        return int(out[0]) if len(out) == 1 else out

    def knn2(self, q, t):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.zeros((len(q), 2), np.int32); dist = np.zeros((len(q), 2), np.int32)
        check(lib().uvip_knn2(self.h, ptr(q), len(q), ptr(t), len(t), ptr(idx), ptr(dist)))
        return idx, dist

    def distinctive_descriptors(self, desc, start):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:197-270) for a batch of map points: the observed
        descriptors of point p are rows start[p]..start[p+1]-1 of desc.  Returns (best index inside each list, its median)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
        n = len(start) - 1
        bi = np.zeros(n, np.int32); bm = np.zeros(n, np.int32)
        check(lib().uvip_distinctive_descriptors(self.h, ptr(desc), ptr(start), n, ptr(bi), ptr(bm)))
        return bi, bm

    def ratio_filter(self, idx, dist, ratio=None):
        idx = np.ascontiguousarray(idx, np.int32); dist = np.ascontiguousarray(dist, np.int32)
        m = np.zeros(len(idx), np.int32); n = C.c_int()
        check(lib().uvip_ratio_filter(self.h, ptr(idx), ptr(dist), len(idx), float(self.mfNNratio if ratio is None else ratio),
                                      ptr(m), C.byref(n)))
        return m

    def rot_hist_filter(self, match, angle_a, angle_b):
        m = np.ascontiguousarray(match, np.int32).copy()
        a = np.ascontiguousarray(angle_a, np.float32); b = np.ascontiguousarray(angle_b, np.float32)
        n = C.c_int()
        check(lib().uvip_rot_hist_filter(self.h, ptr(m), len(m), ptr(a), ptr(b), C.byref(n)))
        return m

    def ratioMatching(self, desc1, desc2, ratio, angles1=None, angles2=None):
        """haloc::Utils::ratioMatching (include/utils.h:81-111) + optional rotation-consistency histogram"""
        idx, dist = self.knn2(desc1, desc2)
        m = self.ratio_filter(idx, dist, ratio)
        if self.mbCheckOrientation and angles1 is not None:
            m = self.rot_hist_filter(m, angles1, angles2)
        return m

    def grid_build(self, kx, ky, bounds, cols=FRAME_GRID_COLS, rows=FRAME_GRID_ROWS):
        kx = np.ascontiguousarray(kx, np.float32); ky = np.ascontiguousarray(ky, np.float32)
        minX, maxX, minY, maxY = bounds
        inv_w = np.float32(cols) / np.float32(maxX - minX); inv_h = np.float32(rows) / np.float32(maxY - minY)
        start = np.zeros(cols * rows + 1, np.int32); items = np.zeros(max(len(kx), 1), np.int32)
        check(lib().uvip_grid_build(self.h, ptr(kx), ptr(ky), len(kx), float(minX), float(minY), float(inv_w), float(inv_h),
                                    cols, rows, ptr(start), ptr(items)))
        return dict(start=start, items=items[:start[-1]].copy(), minX=float(minX), minY=float(minY),
                    inv_w=float(inv_w), inv_h=float(inv_h), cols=cols, rows=rows)

    def search_window(self, mode, th_dist, qu, qv, qr, qminL, qmaxL, qdesc, kx, ky, octave, kdesc, grid, taken=None):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        qu, qv, qr, kx, ky = f32(qu), f32(qv), f32(qr), f32(kx), f32(ky)
        qminL, qmaxL, octave = i32(qminL), i32(qmaxL), i32(octave)
        qdesc = np.ascontiguousarray(qdesc, np.uint8); kdesc = np.ascontiguousarray(kdesc, np.uint8)
        nq, nk = len(qu), len(kx)
        tk = np.full(nk, -1, np.int32) if taken is None else i32(taken).copy()
        match = np.full(nq, -1, np.int32)
        sp = SearchParams(mode, th_dist, self.mfNNratio, grid['minX'], grid['minY'], grid['inv_w'], grid['inv_h'],
                          grid['cols'], grid['rows'])
        n = C.c_int()
        check(lib().uvip_search_window(self.h, C.byref(sp), ptr(qu), ptr(qv), ptr(qr), ptr(qminL), ptr(qmaxL), ptr(qdesc), nq,
                                       ptr(kx), ptr(ky), ptr(octave), ptr(kdesc), nk, ptr(i32(grid['start'])),
                                       ptr(i32(grid['items'])), ptr(tk), ptr(match), C.byref(n)))
        return n.value, match, tk

    def search_frame(self, mode, th_dist, qu, qv, qr, qminL, qmaxL, qdesc, kx, ky, octave, kdesc, bounds, taken=None, cols=64, rows=48):
        """uvip_search_frame: uvip_search_window with the frame grid built on the device by the same call"""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        qu, qv, qr, kx, ky = f32(qu), f32(qv), f32(qr), f32(kx), f32(ky)
        qminL, qmaxL, octave = i32(qminL), i32(qmaxL), i32(octave)
        qdesc = np.ascontiguousarray(qdesc, np.uint8); kdesc = np.ascontiguousarray(kdesc, np.uint8)
        nq, nk = len(qu), len(kx)
        tk = np.full(nk, -1, np.int32) if taken is None else i32(taken).copy()
        match = np.full(nq, -1, np.int32)
        minX, maxX, minY, maxY = bounds
        inv_w = np.float32(cols) / np.float32(maxX - minX); inv_h = np.float32(rows) / np.float32(maxY - minY)
        sp = SearchParams(mode, th_dist, self.mfNNratio, float(minX), float(minY), float(inv_w), float(inv_h), cols, rows)
        n = C.c_int()
        check(lib().uvip_search_frame(self.h, C.byref(sp), ptr(qu), ptr(qv), ptr(qr), ptr(qminL), ptr(qmaxL), ptr(qdesc), nq,
                                      ptr(kx), ptr(ky), ptr(octave), ptr(kdesc), nk, ptr(tk), ptr(match), C.byref(n)))
        return n.value, match, tk

    def haloc_hash(self, desc, start, proj):
        """haloc::Hash::getHash (src/hash.cpp:57-85) for CSR descriptor sets; proj = (num_proj, proj_len) float32 -> (nsets, num_proj*32)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
        proj = np.ascontiguousarray(proj, np.float32)
        nsets = len(start) - 1
        out = np.zeros((nsets, proj.shape[0] * 32), np.float32)
        check(lib().uvip_haloc_hash(self.h, ptr(desc), ptr(start), nsets, ptr(proj), proj.shape[0], proj.shape[1], ptr(out)))
        return out

    def haloc_match(self, query, table):
        """haloc::Hash::match (src/hash.cpp:190-206) of one hash against a table of hashes"""
        query = np.ascontiguousarray(query, np.float32); table = np.ascontiguousarray(table, np.float32).reshape(-1, len(query))
        out = np.zeros(len(table), np.float32)
        check(lib().uvip_haloc_match(self.h, ptr(query), ptr(table), len(table), len(query), ptr(out)))
        return out

    def projection_radius(self, view_cos, level, scale_factors, th=1.0):
        """r = RadiusByViewingCos(viewCos) [* th] * mvScaleFactors[level]  (src/ORBmatcher.cc:68-76,127-133), float by float"""
        r = np.array([lib().uvip_radius_by_viewing_cos(float(c)) for c in view_cos], np.float32)
        if th != 1.0:
            r = (r * np.float32(th)).astype(np.float32)
        return (r * np.asarray(scale_factors, np.float32)[np.asarray(level, np.int32)]).astype(np.float32)

    def search_window_batch_device(self, mode, th_dist, bounds, nframes, q_ptrs, d_nq, q_stride, k_ptrs, d_nk, k_stride,
                                   d_taken, d_match, d_counts, stream=None, cols=64, rows=48):
        """uvip_search_window_batch_device over raw device pointers (ints): q_ptrs = (qu, qv, qr, qminL, qmaxL, qdesc),
        k_ptrs = (kx, ky, octave, kdesc); frame f owns [f*stride, f*stride + n[f]).  Asynchronous."""
        minX, maxX, minY, maxY = bounds
        inv_w = np.float32(cols) / np.float32(maxX - minX); inv_h = np.float32(rows) / np.float32(maxY - minY)
        sp = SearchParams(mode, th_dist, self.mfNNratio, float(minX), float(minY), float(inv_w), float(inv_h), cols, rows)
        vp = C.c_void_p
        check(lib().uvip_search_window_batch_device(self.h, C.byref(sp), int(nframes), *[vp(p) for p in q_ptrs], vp(d_nq), int(q_stride),
                                                    *[vp(p) for p in k_ptrs], vp(d_nk), int(k_stride), vp(d_taken), vp(d_match), vp(d_counts),
                                                    vp(stream) if stream else None))

    def search_lists(self, mode, th_dist, qdesc, cand_start, cand_idx, kdesc, taken=None, ratio=None):
        qdesc = np.ascontiguousarray(qdesc, np.uint8); kdesc = np.ascontiguousarray(kdesc, np.uint8)
        cs = np.ascontiguousarray(cand_start, np.int32); ci = np.ascontiguousarray(cand_idx, np.int32)
        nq, nk = len(qdesc), len(kdesc)
        tk = np.full(nk, -1, np.int32) if taken is None else np.ascontiguousarray(taken, np.int32).copy()
        match = np.full(nq, -1, np.int32); n = C.c_int()
        check(lib().uvip_search_lists(self.h, int(mode), int(th_dist), float(self.mfNNratio if ratio is None else ratio), ptr(qdesc), nq,
                                      ptr(cs), ptr(ci), ptr(kdesc), nk, ptr(tk), ptr(match), C.byref(n)))
        return n.value, match, tk

    def SearchByBoW(self, desc1, nodes1, angles1, desc2, nodes2, angles2, keyframe_pair=False):
        """SearchByBoW (src/ORBmatcher.cc:155-284 / :715-850) over flat arrays: nodesX[i] = vocabulary node id of feature i
        (what DBoW2's FeatureVector groups by).  Features of set 1 are the queries, visited node-major (ascending node id,
        then feature index), matched inside the same node of set 2; TH_LOW, ratio, claims, rotation histogram."""
        nodes1 = np.asarray(nodes1); nodes2 = np.asarray(nodes2)
        order = np.lexsort((np.arange(len(nodes1)), nodes1))
        common = np.intersect1d(nodes1, nodes2)
        order = order[np.isin(nodes1[order], common)]
        by_node = {n: np.nonzero(nodes2 == n)[0] for n in common}
        lists = [by_node[nodes1[i]] for i in order]
        cs = np.zeros(len(order) + 1, np.int32); cs[1:] = np.cumsum([len(l) for l in lists])
        ci = np.concatenate(lists).astype(np.int32) if len(lists) else np.zeros(0, np.int32)
        n, match, taken = self.search_lists(3 if keyframe_pair else 2, self.TH_LOW, np.asarray(desc1)[order], cs, ci, desc2)
        full = np.full(len(nodes1), -1, np.int32); full[order] = match
        if self.mbCheckOrientation:
            full = self.rot_hist_filter(full, angles1, angles2)
        return full

    def SearchByProjection(self, frame, map_points, th=1.0, taken=None):
        """SearchByProjection(FrameKTL&, const vector<MapPoint*>&, float th) (src/ORBmatcher.cc:49-125) over flat arrays.
        frame: dict(kx, ky, octave, kdesc, grid, scale_factors); map_points: dict(u, v, level, view_cos, desc)."""
        lvl = np.asarray(map_points['level'], np.int32)
        r = self.projection_radius(map_points['view_cos'], lvl, frame['scale_factors'], th)
        return self.search_window(0, self.TH_HIGH, map_points['u'], map_points['v'], r, lvl - 1, lvl, map_points['desc'],
                                  frame['kx'], frame['ky'], frame['octave'], frame['kdesc'], frame['grid'], taken)


class ORBVocabulary:
    """DBoW2 ORB vocabulary (include/ORBVocabulary.h = TemplatedVocabulary<FORB::TDescriptor, FORB>), next row N2.

    The tree is kept flat (see uvip_vocabulary_create); `transform` sends the descriptors through the CUDA tree descent and
    builds BowVector / FeatureVector on the host exactly like TemplatedVocabulary::transform (:1119-1195): feature order,
    double accumulation, weighting (TF_IDF/TF/IDF/BINARY) and scoring-dependent normalisation."""
    TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
    L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)

    def __init__(self, tree=None, device=0):
        self.h = C.c_void_p()
        self.device = device
        if tree is not None:
            self._set(tree)

    def _set(self, tree):
        self.tree = tree
        self.k, self.L = int(tree['k']), int(tree['L'])
        self.scoring, self.weighting = int(tree.get('scoring', 0)), int(tree.get('weighting', 0))
        cs = np.ascontiguousarray(tree['child_start'], np.int32); ci = np.ascontiguousarray(tree['child_ids'], np.int32)
        nd = np.ascontiguousarray(tree['desc'], np.uint8); nw = np.ascontiguousarray(tree['weight'], np.float64)
        word = np.ascontiguousarray(tree['word'], np.int32)
        self.close()
        check(lib().uvip_vocabulary_create(self.device, len(word), ptr(cs), ptr(ci), ptr(nd), ptr(nw), ptr(word), self.L, C.byref(self.h)))

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            lib().uvip_vocabulary_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def parse_text(path):
        """loadFromTextFile (TemplatedVocabulary.h:1338-1420): 'k L scoring weighting' then one line per node:
        'parent isLeaf b0 .. b31 weight' (node ids are 1.. in file order, 0 is the root)"""
        with open(path) as f:
            k, L, n1, n2 = [int(v) for v in f.readline().split()[:4]]
            if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
                raise ValueError('Vocabulary loading failure: This is not a correct text file!')
            parent, leaf, desc, weight = [0], [0], [np.zeros(32, np.uint8)], [0.0]
            for line in f:
                t = line.split()
                if len(t) < 35:
                    continue
                parent.append(int(t[0])); leaf.append(int(t[1]))
                desc.append(np.array([int(v) for v in t[2:34]], np.uint8)); weight.append(float(t[34]))
        n = len(parent)
        children = [[] for _ in range(n)]
        for i in range(1, n):
            children[parent[i]].append(i)
        cs = np.zeros(n + 1, np.int32); cs[1:] = np.cumsum([len(c) for c in children])
        ci = np.array([c for ch in children for c in ch], np.int32)
        word = np.full(n, -1, np.int32); w = 0
        for i in range(1, n):
            if leaf[i] > 0:
                word[i] = w; w += 1
        return dict(k=k, L=L, scoring=n1, weighting=n2, child_start=cs, child_ids=ci, desc=np.stack(desc), weight=np.array(weight, np.float64),
                    word=word)

    def loadFromTextFile(self, path):
        self._set(self.parse_text(path))
        return True

    def descend(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid = np.zeros(n, np.int32); nid = np.zeros(n, np.int32); w = np.zeros(n, np.float64)
        check(lib().uvip_bow_transform(self.h, ptr(desc), n, int(levelsup), ptr(wid), ptr(nid), ptr(w)))
        return wid, nid, w

    @staticmethod
    def accumulate(wid, nid, w, weighting, scoring):
        """BowVector / FeatureVector bookkeeping of TemplatedVocabulary::transform(features, v, fv, levelsup)"""
        bow, fv = {}, {}
        tf = weighting in (ORBVocabulary.TF_IDF, ORBVocabulary.TF)
        for i in range(len(wid)):
            if w[i] > 0:                                   # not stopped
                k = int(wid[i])
                if tf:
                    bow[k] = bow.get(k, 0.0) + float(w[i])
                elif k not in bow:
                    bow[k] = float(w[i])
                fv.setdefault(int(nid[i]), []).append(i)
        must = scoring != ORBVocabulary.DOT_PRODUCT
        if tf and bow and not must:
            nd = float(len(bow))
            for k in bow:
                bow[k] /= nd
        if must:
            keys = sorted(bow)
            if scoring == ORBVocabulary.L2_NORM:
                norm = 0.0
                for k in keys:
                    norm += bow[k] * bow[k]
                norm = float(np.sqrt(norm))
            else:
                norm = 0.0
                for k in keys:
                    norm += abs(bow[k])
            if norm > 0.0:
                for k in keys:
                    bow[k] /= norm
        return bow, fv

    def transform(self, desc, levelsup=4):
        wid, nid, w = self.descend(desc, levelsup)
        return self.accumulate(wid, nid, w, self.weighting, self.scoring)


class KLTTracker:
    """cv::buildOpticalFlowPyramid + cv::calcOpticalFlowPyrLK as used by FrameKTL / Tracking::perform_matching
    (src/FrameKTL.cc:76, src/Tracking.cc:1044-1047), next row N1."""
    USE_INITIAL_FLOW, GET_MIN_EIGENVALS = 4, 8

    def __init__(self, max_width, max_height, win=21, max_level=5, nslots=2, device=0):
        self.h = C.c_void_p()
        self.win, self.max_level = win, max_level
        check(lib().uvip_klt_create(device, max_width, max_height, win, max_level, nslots, C.byref(self.h)))

    def close(self):
        if getattr(self, 'h', None) is not None and self.h:
            lib().uvip_klt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_pyramid(self, slot, image):
        image = np.ascontiguousarray(image, np.uint8)
        H, W = image.shape
        n = C.c_int()
        check(lib().uvip_klt_build_pyramid(self.h, slot, ptr(image), W, H, W, C.byref(n)))
        return n.value

    def level(self, slot, l):
        w = C.c_int(); h = C.c_int()
        check(lib().uvip_klt_get_level(self.h, slot, l, None, None, C.byref(w), C.byref(h)))
        img = np.zeros((h.value, w.value), np.uint8); der = np.zeros((h.value, w.value, 2), np.int16)
        check(lib().uvip_klt_get_level(self.h, slot, l, ptr(img), ptr(der), C.byref(w), C.byref(h)))
        return img, der

    def ransac_fundamental(self, pts0, pts1, threshold=1.0, nhyp=2048):
        """cv::findFundamentalMat(pts0, pts1, FM_RANSAC, threshold, 0.999, mask) of src/Tracking.cc:1062: (inlier count, mask, F)"""
        p0 = np.ascontiguousarray(pts0, np.float32).reshape(-1, 2); p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        n = len(p0)
        mask = np.zeros(n, np.uint8); F = np.zeros(9, np.float64); cnt = C.c_int()
        check(lib().uvip_klt_ransac_fundamental(self.h, ptr(p0), ptr(p1), n, float(threshold), int(nhyp), ptr(mask), ptr(F), C.byref(cnt)))
        return cnt.value, mask, F.reshape(3, 3)

    def track(self, slot_prev, slot_next, prev_pts, next_pts, max_iter=30, eps=0.01, flags=12, min_eig_thr=1e-4):
        prev_pts = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
        nxt = np.ascontiguousarray(next_pts, np.float32).reshape(-1, 2).copy()
        n = len(prev_pts)
        status = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
        check(lib().uvip_klt_track(self.h, slot_prev, slot_next, ptr(prev_pts), ptr(nxt), n, self.max_level, int(max_iter), float(eps), int(flags),
                                   float(min_eig_thr), ptr(status), ptr(err)))
        return nxt, status, err
