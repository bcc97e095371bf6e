"""Parity pin against the reference's OWN code.

oracle/_ref/libref_orbextractor.so is USLAM::ORBextractor compiled from /root/reference/src/ORBextractor.cc where it
lies (`make -C oracle ref`), against stand-in OpenCV/Eigen/ROS headers (oracle/ref_shim/): cell geometry, empty-cell
retry, DistributeOctTree/DivideNode, IC_Angle, computeOrbDescriptor, operator() orchestration and the occupancy-grid
filter are the reference's compiled code; the OpenCV primitives behind the stand-ins are the C oracle's restatements,
themselves pinned against cv2 (tests/test_oracle_golden.py).

CPU tests: the C oracle must reproduce the reference bit for bit (keypoints incl. order, angles, descriptors, grid).
GPU tests (-m gpu): the CUDA path through the C-ABI against the reference directly.  The library is prebuilt in the
build container and travels to the GPU box; nothing here reads /root/reference at run time."""
import numpy as np
import pytest

from oracle import reference as R

needs_ref = pytest.mark.skipif(not R.available(), reason='oracle/_ref not built and /root/reference absent')

ANGLE_TOL_DEG = 1e-3 * 180.0 / np.pi
FIELDS = ('x', 'y', 'size', 'angle', 'response', 'octave', 'class_id')


def assert_same(a, b, what=''):
    (ak, ad), (bk, bd) = a, b
    assert len(ak) == len(bk), (what, len(ak), len(bk))
    for f in FIELDS:
        assert np.array_equal(ak[f], bk[f]), (what, f)
    assert np.array_equal(ad, bd), (what, 'descriptors')


@needs_ref
def test_reference_builds_and_reports_its_tables(oracle):
    ex = R.Extractor(1000, 1.2, 8, 1, 20)
    assert ex.GetLevels() == 8 and ex.GetScaleFactor() == np.float32(1.2)


@needs_ref
@pytest.mark.parametrize('seed,W,H,nf,th', [(1, 752, 480, 1000, 20), (1000, 640, 512, 1500, 20), (100000, 1280, 1024, 2000, 20),
                                            (7, 320, 240, 400, 7), (8, 968, 608, 800, 10), (9, 401, 307, 600, 1),
                                            (10, 640, 480, 1000, 40)])
def test_oracle_equals_reference_on_benchmark_shapes(oracle, synth, seed, W, H, nf, th):
    img = synth.synth_frame(seed, W, H)
    assert_same(oracle.Extractor(nf, 1.2, 8, 1, th)(img), R.Extractor(nf, 1.2, 8, 1, th)(img), (seed, W, H))


@needs_ref
def test_oracle_equals_reference_parameter_sweep(oracle, synth):
    rng = np.random.default_rng(2026)
    for i in range(16):
        W, H = int(rng.integers(200, 800)), int(rng.integers(160, 600))
        nf = int(rng.integers(100, 2500)); sf = float(rng.choice([1.1, 1.2, 1.25, 1.5, 2.0])); nl = int(rng.integers(1, 9))
        th = int(rng.choice([5, 7, 12, 20, 30, 60]))
        while min(W, H) / sf ** (nl - 1) < 70:          # keep every level larger than the 2 x 16 px border + one cell
            nl -= 1
        img = synth.synth_frame(500 + i, W, H)
        assert_same(oracle.Extractor(nf, sf, nl, i & 1, th)(img), R.Extractor(nf, sf, nl, i & 1, th)(img), (i, W, H, nf, sf, nl, th))


@needs_ref
def test_oracle_equals_reference_on_hard_images(oracle, synth):
    oex, rex = oracle.Extractor(500, 1.2, 8, 1, 20), R.Extractor(500, 1.2, 8, 1, 20)
    flat = np.full((300, 400), 90, np.uint8)
    noise = (synth.draw(5, np.arange(1, 400 * 300 + 1, dtype=np.uint64)) % np.uint64(256)).astype(np.uint8).reshape(300, 400)
    ramp = np.tile(np.arange(400, dtype=np.uint8), (300, 1))
    sparse = np.full((300, 400), 50, np.uint8); sparse[100:140, 120:180] = 58; sparse[200:203, 300:303] = 66
    for name, img in (('flat', flat), ('noise', noise), ('ramp', ramp), ('sparse', sparse)):
        assert_same(oex(img), rex(img), name)
    assert len(rex(flat)[0]) == 0


@needs_ref
def test_oracle_equals_reference_occupancy_grid_sequence(oracle, synth):
    """FullDetect=false over a short sequence that carries the caller's grid from call to call, with incoming keypoints
    (src/ORBextractor.cc:863,872-910; the call site is src/Tracking.cc:901-946)."""
    W, H, mpd = 752, 480, 20
    oex, rex = oracle.Extractor(1000, 1.2, 8, 1, 20), R.Extractor(1000, 1.2, 8, 1, 20)
    g_o = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F'); g_r = g_o.copy(order='F')
    rng = np.random.default_rng(5)
    for step, (n_in, need) in enumerate(((0, 400), (25, 150), (5, 1000), (60, 37))):
        img = synth.synth_frame(21 + step, W, H)
        inc = np.zeros(n_in, oracle.KP_DTYPE)
        inc['x'] = rng.integers(40, W - 40, n_in); inc['y'] = rng.integers(40, H - 40, n_in)
        inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 7
        a = oex(img, keypoints=inc, grid=g_o, min_px_dist=mpd, full_detect=False, num_needed=need)
        b = rex(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=False, num_needed=need)
        assert_same(a, b, step)
        assert np.array_equal(g_o, g_r), step
        assert len(a[0]) >= n_in
    # FullDetect=true drops the incoming keypoints and leaves the grid alone (:911-913)
    before = g_r.copy(order='F')
    a = oex(img, keypoints=inc, grid=g_o, min_px_dist=mpd, full_detect=True, num_needed=10)
    b = rex(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=True, num_needed=10)
    assert_same(a, b, 'full')
    assert np.array_equal(g_r, before)


@needs_ref
def test_reference_pointer_tiebreak_is_the_only_freedom(oracle, synth):
    """With plain malloc addresses the reference's (size, node pointer) sort (src/ORBextractor.cc:1151) may order
    equal-size nodes differently; the result may then differ from the pinned one only in a few keypoints per level."""
    img = synth.synth_frame(1, 752, 480)
    pk, _ = R.Extractor(1000, 1.2, 8, 1, 20)(img)
    mk, _ = R.Extractor(1000, 1.2, 8, 1, 20)(img, arena_mb=-1)
    a = set(zip(pk['x'], pk['y'], pk['octave'])); b = set(zip(mk['x'], mk['y'], mk['octave']))
    assert len(a - b) <= 0.05 * len(a)
    assert abs(len(pk) - len(mk)) <= 8


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope='module')
def gpu(pkg):
    if pkg.capi.lib().uvip_device_count() < 1:
        pytest.fail('no CUDA device: the gpu-marked tests must run on the B200 box')
    return pkg


def assert_cuda_matches(gk, gd, rk, rd, what=''):
    assert len(gk) == len(rk), (what, len(gk), len(rk))
    for f in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
        assert np.array_equal(gk[f], rk[f]), (what, f)
    if len(gk):
        assert np.abs(gk['angle'] - rk['angle']).max() <= ANGLE_TOL_DEG, what
        assert 1.0 - np.unpackbits(gd ^ rd).mean() >= 0.999, what


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize('seed,W,H,nf,th', [(1, 752, 480, 1000, 20), (1000, 640, 512, 1500, 20), (100000, 1280, 1024, 2000, 20),
                                            (9, 401, 307, 600, 1)])
def test_cuda_equals_reference(gpu, synth, seed, W, H, nf, th):
    img = synth.synth_frame(seed, W, H)
    gk, gd = gpu.ORBextractor(nf, 1.2, 8, 1, th, max_width=W, max_height=H)(img)
    rk, rd = R.Extractor(nf, 1.2, 8, 1, th)(img)
    assert_cuda_matches(gk, gd, rk, rd, (seed, W, H))
    assert np.array_equal(gk['angle'], rk['angle']) and np.array_equal(gd, rd)      # measured: bit-identical, not merely within tolerance


@needs_ref
@pytest.mark.gpu
def test_cuda_equals_reference_occupancy_grid(gpu, synth):
    W, H, mpd = 752, 480, 20
    ex = gpu.ORBextractor(1000, 1.2, 8, 1, 20, max_width=W, max_height=H)
    rex = R.Extractor(1000, 1.2, 8, 1, 20)
    g_g = np.zeros((H // mpd + 2, W // mpd + 2), np.int32, order='F'); g_r = g_g.copy(order='F')
    rng = np.random.default_rng(5)
    for step, (n_in, need) in enumerate(((0, 400), (25, 150), (5, 1000))):
        img = synth.synth_frame(21 + step, W, H)
        inc = np.zeros(n_in, gpu.capi.KP_DTYPE)
        inc['x'] = rng.integers(40, W - 40, n_in); inc['y'] = rng.integers(40, H - 40, n_in)
        inc['size'] = 31; inc['angle'] = -1; inc['octave'] = 0; inc['class_id'] = 7
        gk, gd = ex(img, keypoints=inc, grid_2d=g_g, min_px_dist=mpd, FullDetect=False, num_featsneeded=need)
        rk, rd = rex(img, keypoints=inc, grid=g_r, min_px_dist=mpd, full_detect=False, num_needed=need)
        assert_cuda_matches(gk, gd, rk, rd, step)
        assert np.array_equal(g_g, g_r), step
