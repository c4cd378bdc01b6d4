"""The C++ host side (include/EKF.h, openekfmonoslam_b200/host/): configuration reader and initial-map construction on
the CPU against the oracle, and the sample driver (samples/ekf_main.cpp) end to end on the GPU against the oracle."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

from openekfmonoslam_b200 import build
from openekfmonoslam_b200.params import EkfParams, load_config, synthetic_params, write_config
from openekfmonoslam_b200.scenario import Scenario
from oracle.oracle_lib import OracleFilter

REF_CFG = "/root/reference/kalmanFilter/samples/EKF/config.yml"


@pytest.fixture(scope="module")
def host():
    build.build_host()
    lib = ctypes.CDLL(build.HOST_OUT)
    lib.ekfb_host_load_config.restype = ctypes.c_int
    return lib


def _load(host, path):
    p = EkfParams()
    mm, ms = ctypes.c_int(-1), ctypes.c_int(-1)
    rc = host.ekfb_host_load_config(path.encode(), ctypes.byref(p), ctypes.byref(mm), ctypes.byref(ms))
    return rc, p, mm.value, ms.value


def test_config_reader_round_trip(host, tmp_path):
    p = synthetic_params(640, 480)
    cfg = str(tmp_path / "config.yml")
    write_config(cfg, p, 37, max_map_size=313)
    rc, q, mm, ms = _load(host, cfg)
    assert rc == 0 and mm == 37 and ms == 313
    assert p.as_dict() == q.as_dict()


def test_config_reader_errors(host, tmp_path):
    assert _load(host, str(tmp_path / "missing.yml"))[0] != 0
    bad = tmp_path / "bad.yml"
    bad.write_text('%YAML:1.0\nRunConfiguration:\n  ExtendedKalmanFilter: "EKF"\n  CameraCalibration: "CAM"\n'
                   'ExtendedKalmanFilter:\n  EKF:\n    InitInvDepthRho: "1.0"\n')
    assert _load(host, str(bad))[0] != 0


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not mounted")
def test_config_reader_on_reference_sample_config(host):
    rc, q, mm, ms = _load(host, REF_CFG)
    p, extras = load_config(REF_CFG)
    assert rc == 0
    assert p.as_dict() == q.as_dict()
    assert mm == extras["min_matches_per_image"] and ms == extras["max_map_size"]


def test_host_add_feature_matches_oracle(host):
    sc = Scenario(320, 240, 12)
    kp, ds = sc.frame(0)
    orc = OracleFilter(sc.params)
    orc.init()
    x, P = orc.get_state()
    n = 13
    for i in range(12):
        uv = np.ascontiguousarray(kp[i], np.float64)
        orc.add_feature(uv, ds[i])
        xo = np.zeros(n + 6); Po = np.zeros((n + 6, n + 6))
        host.ekfb_host_add_feature(ctypes.byref(sc.params), uv.ctypes.data_as(ctypes.c_void_p), x.ctypes.data_as(ctypes.c_void_p),
                                   P.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n), xo.ctypes.data_as(ctypes.c_void_p),
                                   Po.ctypes.data_as(ctypes.c_void_p))
        x, P, n = xo, Po, n + 6
        xr, Pr = orc.get_state()
        assert np.array_equal(x, xr)
        assert np.abs(P - Pr).max() <= 1e-15 * np.abs(Pr).max()


def write_kpseq(path, frames):
    """The sample driver's frame source: per frame int32 count, count x (f32 x, f32 y), count x 32 descriptor bytes."""
    with open(path, "wb") as fh:
        for kp, ds in frames:
            fh.write(struct.pack("<i", len(kp)))
            fh.write(np.ascontiguousarray(kp, np.float32).tobytes())
            fh.write(np.ascontiguousarray(ds, np.uint8).tobytes())


@pytest.mark.gpu
def test_sample_driver_matches_oracle(tmp_path):
    """ekf_sample config.yml frames.kpseq: EKF::init on frame 0 (all its keypoints become features, added on the device),
    EKF::step on the rest, no map management; the printed camera state must follow the oracle run the same way."""
    build.build_host()
    N, T = 40, 8
    sc = Scenario(320, 240, N)
    frames = [sc.frame(t) for t in range(T + 1)]
    cfg, seq = str(tmp_path / "config.yml"), str(tmp_path / "frames.kpseq")
    write_config(cfg, sc.params, len(frames[0][0]))   # no more keypoints than requested: all of frame 0, in order
    write_kpseq(seq, frames)
    out = subprocess.run([build.SAMPLE_OUT, cfg, seq], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("STEP")]
    assert len(rows) == T
    orc = OracleFilter(sc.params)
    orc.init()
    kp0, ds0 = frames[0]
    for i in range(len(kp0)):
        orc.add_feature(np.ascontiguousarray(kp0[i], np.float64), ds0[i])
    for t in range(1, T + 1):
        info = orc.step(*frames[t])
        r = rows[t - 1]
        assert int(r[3]) == info["n_matches"] and int(r[5]) == info["n_inliers"] and int(r[7]) == info["n_rescued"]
        xg = np.array([float(v) for v in r[9:22]])
        xo, Po = orc.get_state()
        assert np.abs(xg - xo[:13]).max() <= 1e-9 * np.abs(xo[:13]).max()
        assert abs(float(r[23]) - Po[0, 0]) <= 1e-9 * np.abs(Po).max()
    assert int(rows[-1][5]) > 10


# ---- host-side pieces of the map management, on the CPU ----
class HostConfig(ctypes.Structure):
    from openekfmonoslam_b200.params import MapPolicy as _MP
    _fields_ = [("params", EkfParams), ("policy", _MP), ("map_management_frequency", ctypes.c_int),
                ("divide_times", ctypes.c_int), ("ellipse_size", ctypes.c_double)]


@pytest.mark.skipif(not os.path.exists("/root/reference/experiments/s3/config.yml"), reason="reference tree not mounted")
def test_full_config_of_the_s3_experiment(host):
    c = HostConfig()
    assert host.ekfb_host_load_config_full(b"/root/reference/experiments/s3/config.yml", ctypes.byref(c)) == 0
    pol = c.policy
    assert (pol.min_matches_per_image, pol.max_map_size, pol.max_map_features_count, pol.always_remove_unseen) == (60, 240, 0, 1)
    assert (pol.good_feature_matching_percent, pol.linearity_index_threshold) == (0.5, 0.1)
    assert (c.map_management_frequency, c.divide_times, c.ellipse_size) == (1, 2, 10.0)
    assert c.params.pixels_x == 640 and c.params.fx == 525.060143149240389


def test_new_feature_selection_matches_reference(host):
    """ekfbSelectNewFeatures (host/new_features.cpp) against the reference's own detectNewImageFeatures
    (DetectNewImageFeatures.cpp:321-419) on the same keypoints, predictions and libc rand() stream."""
    from oracle import oracle_lib, ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    libc = ctypes.CDLL(None)
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, _ = sc.init_map()
    r = ref_lib.ReferenceFilter(sc.params)
    W, H = 320, 240
    stamp_R = int(2.0 * np.sqrt(10.0 * 5.9915)) + 2
    D = 2 * stamp_R + 1
    black = oracle_lib.draw_uncertainty_ellipse(np.zeros((D, D), np.uint8), stamp_R, stamp_R, np.diag([10.0, 10.0]), 2 * (W + H))
    stamp = np.where(black != 0, 0, 255).astype(np.uint8)
    rng = np.random.default_rng(11)
    for trial, (nkp, want) in enumerate([(300, 25), (120, 60), (40, 50), (500, 7)]):
        r.set_state(x, P, ft, fo, desc)
        r.predict(); r.measure()
        mo = r.get_measure()
        vis = mo["vis"].astype(bool)
        kp = np.stack([rng.integers(20, W - 20, nkp), rng.integers(20, H - 20, nkp)], 1).astype(np.float32)
        ds = rng.integers(0, 256, (nkp, 32), dtype=np.uint8)
        libc.srand(1000 + trial)
        uv_ref, dd_ref = r.detect_new(kp, ds, want)
        # the mask the device builds: white, every prediction's gate ellipse in black
        blk = np.zeros((H, W), np.uint8)
        for i in np.flatnonzero(vis):
            oracle_lib.draw_uncertainty_ellipse(blk, mo["h"][i, 0], mo["h"][i, 1], mo["S"][i].reshape(2, 2), 2 * (W + H))
        mask = np.where(blk != 0, 0, 255).astype(np.uint8)
        pred = np.ascontiguousarray(mo["h"][vis], np.float64)
        out = np.zeros(want, np.int32)
        libc.srand(1000 + trial)
        k = host.ekfb_host_select_new_features(W, H, 2, mask.ctypes.data_as(ctypes.c_void_p), stamp.ctypes.data_as(ctypes.c_void_p),
                                               stamp_R, kp.ctypes.data_as(ctypes.c_void_p), nkp, pred.ctypes.data_as(ctypes.c_void_p),
                                               len(pred), want, out.ctypes.data_as(ctypes.c_void_p))
        assert k == len(uv_ref), (trial, k, len(uv_ref))
        assert np.array_equal(kp[out[:k]].astype(np.float64), uv_ref) and np.array_equal(ds[out[:k]], dd_ref)
        assert k > 0
    r.close()


class FrameTrace(ctypes.Structure):
    _fields_ = [(k, ctypes.c_double) for k in ("usPrediction", "usMatching", "usRansac", "usUpdateLI", "usRescue", "usUpdateHI",
                                               "usMapManagement")] + \
               [(k, ctypes.c_int) for k in ("totalMatches", "liInliers", "hiInliers", "invDepthCount", "depthCount")] + \
               [("state", ctypes.c_double * 13), ("cov", ctypes.c_double * 169)]


def test_output_yml_is_readable_by_opencv(host, tmp_path):
    """output.yml in the reference's layout (E/EKF.cpp:257-268 ... 618-628): cv::FileStorage must read it back the way
    kalmanFilter/resultReader/main.cpp:82-150 does."""
    cv2 = pytest.importorskip("cv2")
    host.ekfb_host_trace_open.restype = ctypes.c_void_p
    path = str(tmp_path / "output.yml")
    w = ctypes.c_void_p(host.ekfb_host_trace_open(path.encode()))
    assert w
    rng = np.random.default_rng(5)
    frames = []
    for k in range(1, 4):
        t = FrameTrace()
        for i, name in enumerate(("usPrediction", "usMatching", "usRansac", "usUpdateLI", "usRescue", "usUpdateHI", "usMapManagement")):
            setattr(t, name, float(rng.uniform(1, 5000)) if i else 125.0)
        t.totalMatches, t.liInliers, t.hiInliers, t.invDepthCount, t.depthCount = 50 + k, 40 + k, k, 60, k
        st = rng.normal(size=13); cv = rng.normal(size=169) * 1e-7
        st[0] = 2.0; cv[5] = 0.0
        t.state[:] = list(st); t.cov[:] = list(cv)
        host.ekfb_host_trace_frame(w, k, ctypes.byref(t))
        frames.append((t, st, cv))
    host.ekfb_host_trace_close(w)
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    assert fs.isOpened()
    root = fs.root()
    assert list(root.keys()) == ["Frame 1", "Frame 2", "Frame 3"]
    for k, (t, st, cv) in enumerate(frames, 1):
        node = root.getNode(f"Frame {k}")
        assert np.array_equal(node.getNode("StateEstimation").mat(), st.reshape(1, 13))
        assert np.array_equal(node.getNode("StateCovarianceMatrixEstimation").mat(), cv.reshape(13, 13))
        assert int(node.getNode("totalMatches").real()) == t.totalMatches and int(node.getNode("liInliers").real()) == t.liInliers
        assert int(node.getNode("hiInliers").real()) == t.hiInliers
        assert int(node.getNode("MapFeaturesInvDepthCount").real()) == 60 and int(node.getNode("MapFeaturesDepthCount").real()) == k
        for name, key in (("usPrediction", "Prediction"), ("usMatching", "Matching"), ("usRansac", "Ransac"), ("usUpdateLI", "UpdateLI"),
                          ("usRescue", "RescueOutliers"), ("usUpdateHI", "UpdateHI"), ("usMapManagement", "MapManagement")):
            assert node.getNode(key).real() == getattr(t, name)
    fs.release()


@pytest.mark.gpu
def test_sample_driver_with_map_management_follows_the_reference(tmp_path):
    """The whole drop-in: ekf_sample (C++ EKF class -> C ABI -> CUDA, map management on the device, zone-balanced
    new-feature selection on the host, output.yml) against the REFERENCE's own EKF::init / EKF::step (oracle/_ref, map
    management every frame) on the same keypoint sequence and the same libc rand() stream.  Clutter keypoints become
    features, are never matched again and are removed as bad: the map turns over every frame."""
    from openekfmonoslam_b200.params import MapPolicy
    from oracle import ref_lib
    if not ref_lib.available(build=False):
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    build.build_host()
    N, T = 30, 14
    sc = Scenario(320, 240, N)
    frames = [sc.frame(t) for t in range(T + 1)]
    cfg, seq, out_dir = str(tmp_path / "config.yml"), str(tmp_path / "frames.kpseq"), str(tmp_path) + "/"
    pol = MapPolicy(min_matches_per_image=36, max_map_features_count=0, max_map_size=0, always_remove_unseen=1,
                    good_feature_matching_percent=0.5, linearity_index_threshold=0.1)
    write_config(cfg, sc.params, pol.min_matches_per_image,
                 extra={"MapManagementFrequency": 1, "AlwaysRemoveUnseenMapFeatures": "true", "GoodFeatureMatchingPercent": 0.5,
                        "InverseDepthLinearityIndexThreshold": 0.1, "DetectNewFeaturesImageAreasDivideTimes": 2,
                        "DetectNewFeaturesImageMaskEllipseSize": 10})
    write_kpseq(seq, frames)
    out = subprocess.run([build.SAMPLE_OUT, cfg, seq, out_dir], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("STEP")]
    assert len(rows) == T
    # the reference, in this process: libc rand() starts from the default seed like the freshly started sample
    ctypes.CDLL(None).srand(1)
    r = ref_lib.ReferenceFilter(sc.params)
    r.set_policy(pol, 1)
    r.full_init(*frames[0])
    turnover = 0
    for t in range(1, T + 1):
        N_before = r.dims()[1]
        r.step(*frames[t])
        row = rows[t - 1]
        xr, Pr = r.get_state()
        n, Nf = r.dims()
        xg = np.array([float(v) for v in row[9:22]])
        assert np.abs(xg - xr[:13]).max() <= 1e-9 * np.abs(xr[:13]).max(), f"frame {t}"
        assert abs(float(row[23]) - Pr[0, 0]) <= 1e-9 * np.abs(Pr).max()
        assert (int(row[25]), int(row[27])) == (Nf, n), f"frame {t}: map size"
        turnover += int(row[29]) + int(row[33])
    assert turnover > 10, "the scenario must exercise removal and addition"
    # output.yml of the run, read the way kalmanFilter/resultReader does
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(out_dir + "output.yml", cv2.FILE_STORAGE_READ)
    root = fs.root()
    assert len(root.keys()) == T
    last = root.getNode(f"Frame {T}")
    assert np.abs(last.getNode("StateEstimation").mat().ravel() - xr[:13]).max() <= 1e-9 * np.abs(xr[:13]).max()
    assert int(last.getNode("totalMatches").real()) == int(rows[-1][3])
    assert last.getNode("UpdateLI").real() > 0
    fs.release()
    # log.txt (E/EKF.cpp:172-180,246-250,662-665 -> State::showDetailed): one banner per step, the camera line and one line
    # per map feature of the final state, positions to six significant digits
    log = open(out_dir + "log.txt").read()
    assert log.count("~~~~~~~~~~~~ STEP ") == T + 1 and f"STEP {T} ~" in log
    tail = log[log.rindex("~~~~~~~~~~~~ STEP "):]
    cam = [ln for ln in tail.splitlines() if ln.startswith("Posicion de la camara: ")][0]
    assert np.allclose([float(v) for v in cam.split(": ")[1].split(", ")], xr[:3], rtol=1e-5, atol=1e-12)
    assert f"Map Features ({Nf}):" in tail and len([ln for ln in tail.splitlines() if ln[:1].isdigit() and ": " in ln]) == Nf
    r.close()


def test_config_reader_defaults_of_optional_keys(host, tmp_path):
    """keys the reference treats as optional (ExtendedKalmanFilterConfiguration.cpp: MaxMapFeaturesCount, MaxMapSize,
    AlwaysRemoveUnseenMapFeatures default to 0 / false) may be absent"""
    cfg = str(tmp_path / "config.yml")
    write_config(cfg, synthetic_params(320, 240), 25)
    c = HostConfig()
    assert host.ekfb_host_load_config_full(cfg.encode(), ctypes.byref(c)) == 0
    pol = c.policy
    assert (pol.min_matches_per_image, pol.max_map_size, pol.max_map_features_count, pol.always_remove_unseen) == (25, 0, 0, 0)
    assert c.map_management_frequency == 0


def test_new_feature_selection_takes_everything_when_few_keypoints(host):
    """DetectNewImageFeatures.cpp:358-371: no more unmasked keypoints than requested -> all of them, in order, no rand() drawn"""
    libc = ctypes.CDLL(None)
    W, H = 320, 240
    mask = np.full((H, W), 255, np.uint8)
    mask[:, :100] = 0                                   # left part masked
    kp = np.array([[50, 50], [150, 60], [200, 100], [99.6, 10], [310, 230]], np.float32)   # (99.6 + 0.5) -> column 100: kept
    stamp = np.zeros((1, 1), np.uint8)
    out = np.zeros(8, np.int32)
    libc.srand(7); a = libc.rand(); libc.srand(7)
    k = host.ekfb_host_select_new_features(W, H, 2, mask.ctypes.data_as(ctypes.c_void_p), stamp.ctypes.data_as(ctypes.c_void_p), 0,
                                           kp.ctypes.data_as(ctypes.c_void_p), len(kp), None, 0, 8, out.ctypes.data_as(ctypes.c_void_p))
    assert k == 4 and list(out[:4]) == [1, 2, 3, 4]
    assert libc.rand() == a                              # the random stream was not touched
    assert host.ekfb_host_select_new_features(W, H, 2, mask.ctypes.data_as(ctypes.c_void_p), stamp.ctypes.data_as(ctypes.c_void_p), 0,
                                              kp.ctypes.data_as(ctypes.c_void_p), len(kp), None, 0, 0, out.ctypes.data_as(ctypes.c_void_p)) == 0
