"""The device front end (csrc/ekf_frontend.cuh).  CPU: the numpy restatement of OpenCV's FAST-9/16 + 3x3 NMS
against cv2.FastFeatureDetector (identical keypoints, order, scores) on synthetic images and on a crop of the reference's
own s3 frame 00090 (tests/golden/s3_frame_crop.npz).  GPU: the CUDA detector / descriptor against that restatement, exact,
and one frame-to-frame run of the whole pipeline from device-detected keypoints."""
import os

import numpy as np
import pytest

from oracle import fast_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "s3_frame_crop.npz")


def images():
    rng = np.random.default_rng(0)
    out = [("s3 crop", np.ascontiguousarray(np.load(GOLD)["crop"]))]
    out.append(("white noise", rng.integers(0, 256, (120, 160), dtype=np.uint8)))
    n = rng.normal(size=(240 + 16, 320 + 16))
    k = np.ones(5) / 5
    for _ in range(2):      # separable box blur: smooth structure with real corners
        n = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 0, n)
        n = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, n)
    n = n[8:-8, 8:-8]
    out.append(("blurred noise", np.clip(128 + 400 * n, 0, 255).astype(np.uint8)))
    out.append(("odd size", np.ascontiguousarray(out[-1][1][3:3 + 187, 5:5 + 251])))   # width not a multiple of 16: row pitch != width
    return out


@pytest.mark.parametrize("threshold", [5, 20, 40])
def test_fast_restatement_equals_opencv(threshold):
    cv2 = pytest.importorskip("cv2")
    det = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    total = 0
    for name, img in images():
        kps = det.detect(img, None)
        ref = np.array([k.pt for k in kps], np.float32).reshape(-1, 2)
        xy, sc = fast_oracle.detect(img, threshold)
        assert np.array_equal(xy, ref), name
        assert np.array_equal(sc, np.array([k.response for k in kps])), name
        total += len(ref)
    assert total > 300


def test_descriptor_restatement_basics():
    img = images()[2][1]
    xy, _ = fast_oracle.detect(img, 10, border=fast_oracle.BORDER)
    d = fast_oracle.describe(img, xy)
    assert d.shape == (len(xy), 32) and len(xy) > 20
    assert np.array_equal(d, fast_oracle.describe(img.copy(), xy))
    shifted = np.roll(img, (3, 5), axis=(0, 1))       # the same structure 5 px right, 3 px down: identical descriptors
    inner = (xy[:, 0] > 30) & (xy[:, 0] < img.shape[1] - 30) & (xy[:, 1] > 30) & (xy[:, 1] < img.shape[0] - 30)
    assert np.array_equal(fast_oracle.describe(shifted, xy[inner] + np.array([5, 3], np.float32)), d[inner])
    bits = np.unpackbits(d, axis=1)
    assert 0.3 < bits.mean() < 0.7                     # comparisons are balanced


@pytest.mark.gpu
@pytest.mark.parametrize("threshold", [8, 25])
def test_gpu_detector_and_descriptor_equal_restatement(threshold):
    from openekfmonoslam_b200.capi import EkfBatch
    from openekfmonoslam_b200.params import synthetic_params
    for name, img in images():
        H, W = img.shape
        gpu = EkfBatch(synthetic_params(W, H), 1, 8, 20000)
        gpu.set_image(0, img)
        n = gpu.detect_keypoints(0, threshold)
        xy, ds = gpu.get_keypoints(0, n)
        ref, _ = fast_oracle.detect(img, threshold, border=fast_oracle.BORDER)
        assert n == len(ref) and np.array_equal(xy, ref), name
        assert np.array_equal(ds, fast_oracle.describe(img, ref)), name
        assert n > 10
    # capacity: the first max_keypoints in raster order
    name, img = images()[1]
    gpu = EkfBatch(synthetic_params(img.shape[1], img.shape[0]), 1, 8, 50)
    gpu.set_image(0, img)
    assert gpu.detect_keypoints(0, 8) == 50
    ref, _ = fast_oracle.detect(img, 8, border=fast_oracle.BORDER)
    assert np.array_equal(gpu.get_keypoints(0, 50)[0], ref[:50])


@pytest.mark.gpu
def test_gpu_frame_to_frame_from_device_keypoints():
    """image -> device keypoints -> map (ekfb_add_features) -> next image (shifted) -> device keypoints -> ekfb_step"""
    from openekfmonoslam_b200.capi import EkfBatch
    from openekfmonoslam_b200.params import synthetic_params
    from oracle.oracle_lib import OracleFilter
    world = images()[2][1]
    world = np.tile(world, (3, 3))[:600, :800]
    W, H = 640, 480
    p = synthetic_params(W, H)
    gpu = EkfBatch(p, 1, 80, 20000)
    orc = OracleFilter(p)
    orc.init()
    x0, P0 = orc.get_state()
    gpu.set_state(0, x0, P0, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 32), np.uint8))
    f0 = np.ascontiguousarray(world[40:40 + H, 60:60 + W])
    gpu.set_image(0, f0)
    n = gpu.detect_keypoints(0, 12)
    xy, ds = gpu.get_keypoints(0, n)
    pick = np.linspace(0, n - 1, 60).astype(int)
    gpu.add_features(0, xy[pick].astype(np.float64), ds[pick])
    for t in range(1, 4):
        ft = np.ascontiguousarray(world[40:40 + H, 60 + t:60 + t + W])      # the scene slides 1 px per frame
        gpu.set_image(0, ft)
        assert gpu.detect_keypoints(0, 12) > 100
        gpu.step()
        info = gpu.frame_info(0)
        assert info["status"] == 0 and info["n_matches"] >= 40 and info["n_inliers"] >= 25, info


@pytest.mark.gpu
def test_gpu_drop_in_class_with_device_front_end(tmp_path):
    """The C++ EKF class through its flat C binding (what the Android JNI layer binds): BGR frames in, detector + descriptor,
    matching, RANSAC, updates and map management on the device, state out."""
    import ctypes
    from openekfmonoslam_b200 import build
    from openekfmonoslam_b200.params import synthetic_params, write_config
    build.build_host()
    host = ctypes.CDLL(build.HOST_OUT)
    host.ekfb_host_ekf_create.restype = ctypes.c_void_p
    W, H = 640, 480
    cfg = str(tmp_path / "config.yml")
    write_config(cfg, synthetic_params(W, H), 60, max_map_size=600,
                 extra={"MapManagementFrequency": 1, "GoodFeatureMatchingPercent": 0.5, "InverseDepthLinearityIndexThreshold": 0.1,
                        "DetectNewFeaturesImageAreasDivideTimes": 2, "DetectNewFeaturesImageMaskEllipseSize": 10})
    world = np.tile(images()[2][1], (3, 3))[:600, :800]
    e = ctypes.c_void_p(host.ekfb_host_ekf_create(cfg.encode(), b"", 0, 12))

    def frame(t, channels):
        g = np.ascontiguousarray(world[40:40 + H, 60 + t:60 + t + W])
        return g if channels == 1 else np.ascontiguousarray(np.repeat(g[:, :, None], channels, axis=2))

    f0 = frame(0, 3)
    assert host.ekfb_host_ekf_init(e, f0.ctypes.data_as(ctypes.c_void_p), W, H, 3) == 0
    x = np.zeros(13); c = np.zeros(8, np.int32)
    host.ekfb_host_ekf_get(e, x.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
    assert c[0] == 60 and c[1] == 13 + 6 * 60 and x[3] == 1.0
    for t in range(1, 6):
        ch = (3, 4, 1)[t % 3]
        ft = frame(t, ch)
        assert host.ekfb_host_ekf_step(e, ft.ctypes.data_as(ctypes.c_void_p), W, H, ch) == 0
        host.ekfb_host_ekf_get(e, x.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
        assert c[2] >= 40 and c[3] >= 25, (t, c)
        assert abs(np.linalg.norm(x[3:7]) - 1.0) < 1e-12
    assert c[0] >= 55
    host.ekfb_host_ekf_destroy(e)


@pytest.mark.gpu
def test_gpu_colour_frame_to_grey_like_cvtcolor():
    """ekfb_set_image_color: BGR / BGRA frames become grey with OpenCV 2.4's fixed-point rule (the reference's OpenCV:
    (B 1868 + G 9617 + R 4899 + 2^13) >> 14; cv2 4.x uses 15-bit coefficients and may differ by one grey level)"""
    cv2 = pytest.importorskip("cv2")
    from openekfmonoslam_b200.capi import EkfBatch
    from openekfmonoslam_b200.params import synthetic_params
    rng = np.random.default_rng(8)
    H, W = 187, 251
    gpu = EkfBatch(synthetic_params(W, H), 1, 4, 64)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    want = ((bgr[:, :, 0].astype(np.int64) * 1868 + bgr[:, :, 1].astype(np.int64) * 9617 + bgr[:, :, 2].astype(np.int64) * 4899 + 8192) >> 14).astype(np.uint8)
    gpu.set_image_color(0, bgr)
    got = gpu.ncc_level(0, 0)
    assert np.array_equal(got, want)
    assert np.abs(got.astype(int) - cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY).astype(int)).max() <= 1
    bgra = np.concatenate([bgr, rng.integers(0, 256, (H, W, 1), dtype=np.uint8)], axis=2)
    gpu.set_image_color(0, bgra)
    assert np.array_equal(gpu.ncc_level(0, 0), want)
