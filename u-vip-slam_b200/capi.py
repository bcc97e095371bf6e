"""ctypes binding of the C-ABI in include/uvip_orb.h (libuvip_orb.so).  No fallback: if the library is missing
or a call fails, a UvipError is raised."""
import ctypes as C
import os
import re

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, '..', 'include', 'uvip_orb.h')
_LIB = None

KP_DTYPE = np.dtype([('x', 'f4'), ('y', 'f4'), ('size', 'f4'), ('angle', 'f4'), ('response', 'f4'),
                     ('octave', 'i4'), ('class_id', 'i4')])
assert KP_DTYPE.itemsize == 28

OK, ERR_ARG, ERR_CAPACITY, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5


class UvipError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, 'libuvip_orb status %d: %s' % (code, text))
        self.code = code


class ExtractorParams(C.Structure):
    _fields_ = [('nfeatures', C.c_int32), ('scale_factor', C.c_float), ('nlevels', C.c_int32), ('score_type', C.c_int32),
                ('fast_th', C.c_int32), ('retry_th', C.c_int32), ('cell', C.c_int32), ('device', C.c_int32),
                ('max_width', C.c_int32), ('max_height', C.c_int32), ('max_batch', C.c_int32)]


class SearchParams(C.Structure):
    _fields_ = [('mode', C.c_int32), ('th_dist', C.c_int32), ('ratio', C.c_float), ('min_x', C.c_float), ('min_y', C.c_float),
                ('inv_w', C.c_float), ('inv_h', C.c_float), ('cols', C.c_int32), ('rows', C.c_int32)]


def declared_symbols():
    """every function the header declares (used by the CPU test that checks the library exports them all)"""
    txt = open(HEADER).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(uvip_[a-z0-9_]+)\s*\(', txt)))


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get('UVIP_LIB') or _build.SO          # UVIP_LIB: A/B builds of the same library while tuning kernels
        if not os.path.exists(so):
            so = _build.build_cuda()
        L = C.CDLL(so)
        L.uvip_last_error.restype = C.c_char_p
        L.uvip_extractor_scale_factor.restype = C.c_float
        L.uvip_radius_by_viewing_cos.restype = C.c_float
        L.uvip_radius_by_viewing_cos.argtypes = [C.c_float]
        L.uvip_extractor_launch_count.restype = C.c_longlong
        L.uvip_matcher_launch_count.restype = C.c_longlong
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        L.uvip_extractor_create.argtypes = [C.POINTER(ExtractorParams), C.POINTER(vp)]
        L.uvip_extractor_destroy.argtypes = [vp]
        L.uvip_extractor_levels.argtypes = [vp]
        L.uvip_extractor_scale_factor.argtypes = [vp]
        L.uvip_extractor_tables.argtypes = [vp, vp, vp, vp, vp]
        L.uvip_extract.argtypes = [vp, vp, i, i, i, vp, C.POINTER(i), i, vp, vp, i, i, i, i, i]
        L.uvip_extract_batch.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp]
        L.uvip_distinctive_descriptors.argtypes = [vp, vp, vp, i, vp, vp]
        L.uvip_extract_batch_submit.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp, C.POINTER(i)]
        L.uvip_extract_batch_wait.argtypes = [vp, i]
        L.uvip_extract_match_batch_submit.argtypes = [vp, vp, vp, i, i, i, i, sz, vp, vp, i, vp, vp, vp, C.POINTER(i)]
        L.uvip_extract_batch_device.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp, vp]
        L.uvip_extractor_status.argtypes = [vp]
        L.uvip_extractor_set_subbatch.argtypes = [vp, i]
        L.uvip_get_pyramid_level.argtypes = [vp, i, i, i, vp, i, C.POINTER(i), C.POINTER(i)]
        L.uvip_get_raw_corners.argtypes = [vp, i, i, vp, vp, vp, i, C.POINTER(i)]
        L.uvip_get_level_keypoints.argtypes = [vp, i, i, vp, vp, vp, i, C.POINTER(i)]
        L.uvip_extractor_launch_count.argtypes = [vp]
        L.uvip_extractor_graph_captures.argtypes = [vp]
        L.uvip_extractor_graph_captures.restype = C.c_longlong
        L.uvip_extractor_stream.argtypes = [vp]
        L.uvip_extractor_stream.restype = vp
        L.uvip_matcher_stream.argtypes = [vp]
        L.uvip_matcher_stream.restype = vp
        L.uvip_matcher_sync.argtypes = [vp]
        L.uvip_extractor_profile.argtypes = [vp, i]
        L.uvip_extractor_stage_ms.argtypes = [vp, vp, C.POINTER(i)]
        L.uvip_clahe.argtypes = [vp, vp, i, i, i, C.c_double, i, i, vp, i]
        L.uvip_clahe_batch_device.argtypes = [vp, vp, i, i, i, i, sz, C.c_double, i, i, vp, i, sz, vp]
        L.uvip_compute_keypoints_quota.argtypes = [vp, i, vp, vp, i]
        L.uvip_harris_responses.argtypes = [vp, i, i, vp, vp, i, i, C.c_float, vp]
        L.uvip_matcher_create.argtypes = [i, C.POINTER(vp)]
        L.uvip_matcher_destroy.argtypes = [vp]
        L.uvip_matcher_launch_count.argtypes = [vp]
        L.uvip_descriptor_distance.argtypes = [vp, vp, vp, i, vp]
        L.uvip_knn2.argtypes = [vp, vp, i, vp, i, vp, vp]
        L.uvip_knn2_device.argtypes = [vp, vp, i, vp, i, i, vp, vp, vp]
        L.uvip_knn2_batch_device.argtypes = [vp, vp, vp, sz, vp, vp, sz, i, i, vp, vp, sz, vp]
        L.uvip_knn2_batch.argtypes = [vp, vp, vp, sz, vp, vp, sz, i, i, vp, vp, sz]
        L.uvip_knn2_merge_device.argtypes = [vp, vp, vp, i, sz, i, vp, vp, vp]
        L.uvip_ratio_filter.argtypes = [vp, vp, vp, i, C.c_double, vp, C.POINTER(i)]
        L.uvip_rot_hist_filter.argtypes = [vp, vp, i, vp, vp, C.POINTER(i)]
        L.uvip_grid_build.argtypes = [vp, vp, vp, i, C.c_float, C.c_float, C.c_float, C.c_float, i, i, vp, vp]
        L.uvip_search_window_batch_device.argtypes = [vp, C.POINTER(SearchParams), i] + [vp] * 7 + [i] + [vp] * 5 + [i] + [vp] * 4
        L.uvip_search_window.argtypes = [vp, C.POINTER(SearchParams), vp, vp, vp, vp, vp, vp, i,
                                         vp, vp, vp, vp, i, vp, vp, vp, vp, C.POINTER(i)]
        L.uvip_haloc_hash.argtypes = [vp, vp, vp, i, vp, i, i, vp]
        L.uvip_haloc_match.argtypes = [vp, vp, vp, i, i, vp]
        L.uvip_search_frame.argtypes = [vp, C.POINTER(SearchParams), vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, C.POINTER(i)]
        L.uvip_search_lists_epipolar.argtypes = [vp, i, vp, vp, i, vp, vp, vp, vp, vp, vp, i, vp, vp, C.POINTER(i)]
        L.uvip_search_lists.argtypes = [vp, i, i, C.c_float, vp, i, vp, vp, vp, i, vp, vp, C.POINTER(i)]
        L.uvip_vocabulary_create.argtypes = [i, i, vp, vp, vp, vp, vp, i, C.POINTER(vp)]
        L.uvip_vocabulary_destroy.argtypes = [vp]
        L.uvip_bow_transform.argtypes = [vp, vp, i, i, vp, vp, vp]
        L.uvip_bow_transform_device.argtypes = [vp, vp, i, i, vp, vp, vp, vp]
        L.uvip_klt_create.argtypes = [i, i, i, i, i, i, C.POINTER(vp)]
        L.uvip_klt_destroy.argtypes = [vp]
        L.uvip_klt_build_pyramid.argtypes = [vp, i, vp, i, i, i, C.POINTER(i)]
        L.uvip_klt_get_level.argtypes = [vp, i, i, vp, vp, C.POINTER(i), C.POINTER(i)]
        L.uvip_klt_track.argtypes = [vp, i, i, vp, vp, i, i, i, C.c_double, i, C.c_double, vp, vp]
        L.uvip_klt_launch_count.argtypes = [vp]
        L.uvip_klt_track_sequence_device.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, i, i, C.c_double, i, C.c_double, vp, vp, vp]
        L.uvip_klt_ransac_fundamental.argtypes = [vp, vp, vp, i, C.c_double, i, vp, vp, C.POINTER(i)]
        L.uvip_klt_launch_count.restype = C.c_longlong
        L.uvip_synth_frames_device.argtypes = [vp, vp, vp, i, i, i, vp, sz, vp]
        L.uvip_popc_peak.argtypes = [i, i, C.POINTER(C.c_double)]
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise UvipError(rc, (lib().uvip_last_error() or b'').decode('utf-8', 'replace'))


def ptr(a):
    """host pointer of a numpy array, or the integer itself (device pointer), or None"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)
