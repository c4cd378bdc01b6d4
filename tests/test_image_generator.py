"""FileSequenceImageGenerator + the image reader of the host side (include/ImageGenerator.h; reference:
kalmanFilter/modules/ImageGenerator/FileSequenceImageGenerator.cpp:60-98, which decodes with cv::imread).
CPU: every PNG colour type / filter choice cv2 writes, binary PGM / PPM -- decoded to the same BGR bytes cv2.imread returns.
GPU: the sample driver in image-sequence mode (numbered PNG frames -> device front end -> filter) follows the same frames fed
through the flat C binding of the EKF class."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def host():
    from openekfmonoslam_b200 import build
    build.build_host()
    lib = ctypes.CDLL(build.HOST_OUT)
    lib.ekfb_host_read_image.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    return lib


def read(host, path):
    w, h = ctypes.c_int(), ctypes.c_int()
    if host.ekfb_host_read_image(path.encode(), None, ctypes.byref(w), ctypes.byref(h)) != 0:
        return None
    out = np.zeros((h.value, w.value, 3), np.uint8)
    assert host.ekfb_host_read_image(path.encode(), out.ctypes.data_as(ctypes.c_void_p), None, None) == 0
    return out


def test_png_and_pnm_reader_equals_cv2_imread(host, tmp_path):
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:97, 0:131]
    smooth = ((np.sin(xx / 9.0) + np.cos(yy / 7.0)) * 60 + 128).astype(np.uint8)       # gradients: Sub / Up / Paeth rows
    cases = {
        "grey_noise.png": rng.integers(0, 256, (97, 131), dtype=np.uint8),
        "grey_smooth.png": smooth,
        "rgb_noise.png": rng.integers(0, 256, (64, 48, 3), dtype=np.uint8),
        "rgba.png": rng.integers(0, 256, (33, 70, 4), dtype=np.uint8),
        "one_pixel.png": np.array([[[1, 2, 3]]], np.uint8),
        "grey.pgm": rng.integers(0, 256, (40, 57), dtype=np.uint8),
        "colour.ppm": rng.integers(0, 256, (40, 57, 3), dtype=np.uint8),
    }
    cases["rgb.png"] = np.ascontiguousarray(np.stack([smooth, np.roll(smooth, 5, axis=1), 255 - smooth], axis=2))
    for name, img in cases.items():
        path = str(tmp_path / name)
        for level in ((0, 3, 9) if name.endswith(".png") else (None,)):
            assert cv2.imwrite(path, img, [] if level is None else [cv2.IMWRITE_PNG_COMPRESSION, level])
            want = cv2.imread(path)              # default flag: 8-bit BGR, alpha dropped, grey replicated
            got = read(host, path)
            assert got is not None and got.shape == want.shape, name
            assert np.array_equal(got, want), name
    assert read(host, str(tmp_path / "missing.png")) is None
    bad = str(tmp_path / "sixteen.png")
    cv2.imwrite(bad, rng.integers(0, 65536, (8, 8), dtype=np.uint16))
    assert read(host, bad) is None               # 16-bit samples are refused, not misread


@pytest.mark.gpu
def test_sample_driver_reads_an_image_sequence(host, tmp_path):
    """ekf_sample config.yml framesDir/ : numbered PNG files -> FileSequenceImageGenerator -> EKF::init / EKF::step with the device
    front end.  The same frames through the flat C binding (BGR arrays from cv2.imread) give the same trajectory."""
    from openekfmonoslam_b200 import build
    from openekfmonoslam_b200.params import synthetic_params, write_config
    from test_frontend import images
    W, H, T = 640, 480, 6
    cfg = str(tmp_path / "config.yml")
    write_config(cfg, synthetic_params(W, H), 60, max_map_size=600,
                 extra={"MapManagementFrequency": 1, "GoodFeatureMatchingPercent": 0.5, "InverseDepthLinearityIndexThreshold": 0.1,
                        "DetectNewFeaturesImageAreasDivideTimes": 2, "DetectNewFeaturesImageMaskEllipseSize": 10})
    world = np.tile(images()[2][1], (3, 3))[:600, :800]
    seq = tmp_path / "frames"
    seq.mkdir()
    for t in range(T):
        g = np.ascontiguousarray(world[40:40 + H, 60 + t:60 + t + W])
        assert cv2.imwrite(str(seq / f"{7 + t:05d}.png"), g if t % 2 else np.repeat(g[:, :, None], 3, axis=2))
    env = dict(os.environ, EKFB_SEQ_BEGIN="7", EKFB_SEQ_END="6550", EKFB_FAST_THRESHOLD="12")
    ctypes.CDLL(None).srand(1)
    run = subprocess.run([build.SAMPLE_OUT, cfg, str(seq) + "/"], capture_output=True, text=True, timeout=300, env=env)
    assert run.returncode == 0, run.stderr
    rows = [ln.split() for ln in run.stdout.splitlines() if ln.startswith("STEP")]
    assert len(rows) == T - 1            # the sequence ends at the first missing file
    # the same frames through the flat C binding
    host.ekfb_host_ekf_create.restype = ctypes.c_void_p
    e = ctypes.c_void_p(host.ekfb_host_ekf_create(cfg.encode(), b"", 0, 12))
    ctypes.CDLL(None).srand(1)
    x = np.zeros(13); c = np.zeros(8, np.int32)
    for t in range(T):
        bgr = np.ascontiguousarray(cv2.imread(str(seq / f"{7 + t:05d}.png")))
        fn = host.ekfb_host_ekf_init if t == 0 else host.ekfb_host_ekf_step
        assert fn(e, bgr.ctypes.data_as(ctypes.c_void_p), W, H, 3) == 0
        if t:
            host.ekfb_host_ekf_get(e, x.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p))
            got = np.array([float(v) for v in rows[t - 1][9:22]])
            assert np.array_equal(got, x), t
            assert int(rows[t - 1][5]) >= 25          # inliers
    host.ekfb_host_ekf_destroy(e)
