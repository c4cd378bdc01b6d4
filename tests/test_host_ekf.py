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
    """ekf_sample config.yml frames.kpseq: EKF::init on frame 0 (host add-feature of the first MinMatchesPerImage
    keypoints), EKF::step on the rest; the printed camera state must follow the oracle run the same way."""
    build.build_host()
    N, T = 40, 8
    sc = Scenario(320, 240, N)
    frames = [sc.frame(t) for t in range(T + 1)]
    cfg, seq = str(tmp_path / "config.yml"), str(tmp_path / "frames.kpseq")
    write_config(cfg, sc.params, N)
    write_kpseq(seq, frames)
    out = subprocess.run([build.SAMPLE_OUT, cfg, seq], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("STEP")]
    assert len(rows) == T
    orc = OracleFilter(sc.params)
    orc.init()
    kp0, ds0 = frames[0]
    for i in range(N):
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
