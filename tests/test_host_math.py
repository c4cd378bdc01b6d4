"""The product's host+device geometry (openekfmonoslam_b200/csrc/ekf_math.cuh) compiled for the host by
tests/hd/hd_harness.cu, compared with the oracle: motion Jacobians, h(x), H_x, H_f, gate ellipse."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.scenario import Scenario
from oracle.oracle_lib import OracleFilter

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hd():
    so = os.path.join(HERE, "hd", "_build", "libhd.so")
    src = os.path.join(HERE, "hd", "hd_harness.cu")
    hdr = os.path.join(HERE, "..", "openekfmonoslam_b200", "csrc", "ekf_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC",
                               "-shared", "-o", so, src])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _state_after(frames=5):
    sc = Scenario(320, 240, 40)
    x, P, ft, fo, desc, _ = sc.init_map()
    f = OracleFilter(sc.params)
    f.set_state(x, P, ft, fo, desc)
    for t in range(1, frames + 1):
        f.step(*sc.frame(t))
    return sc, f, ft, fo


def test_motion_model(hd):
    sc, f, ft, fo = _state_after()
    x0, P0 = f.get_state()
    F = np.zeros((13, 13)); G = np.zeros((13, 13)); xn = np.zeros(13)
    hd.hd_motion(ctypes.byref(sc.params), _p(np.ascontiguousarray(x0[:13])), _p(F), _p(G), _p(xn))
    f.predict()
    x1, P1 = f.get_state()
    Pp = P0.copy()
    Pp[:13, :13] = F @ P0[:13, :13] @ F.T + G
    Pp[:13, 13:] = F @ P0[:13, 13:]
    Pp[13:, :13] = P0[13:, :13] @ F.T
    assert np.abs(xn - x1[:13]).max() < 1e-15
    assert rel_err(Pp, P1) < 1e-13


def test_motion_model_omega_zero_branch(hd):
    # |omega| < EPSILON: F[w,w] diagonal zeroed and G[q,alpha] = 0 (StateAndCovariancePrediction.cpp:176-184,211-212)
    sc = Scenario(320, 240, 4)
    x, P, ft, fo, desc, _ = sc.init_map()
    x[10:13] = 0.0
    f = OracleFilter(sc.params)
    f.set_state(x, P, ft, fo, desc)
    F = np.zeros((13, 13)); G = np.zeros((13, 13)); xn = np.zeros(13)
    hd.hd_motion(ctypes.byref(sc.params), _p(np.ascontiguousarray(x[:13])), _p(F), _p(G), _p(xn))
    assert F[10, 10] == 0 and F[11, 11] == 0 and F[12, 12] == 0 and np.all(G[3:7, 3:7] == 0)
    f.predict()
    _, P1 = f.get_state()
    assert rel_err(F @ P[:13, :13] @ F.T + G, P1[:13, :13]) < 1e-13


def test_measurement_model_inverse_depth(hd):
    sc, f, ft, fo = _state_after()
    f.predict(); f.measure()
    x1, P1 = f.get_state()
    m = f.get_measure()
    N = len(ft)
    vis = np.zeros(N, np.uint8); h = np.zeros((N, 2)); Hx = np.zeros((N, 14)); Hf = np.zeros((N, 12))
    hd.hd_measure(ctypes.byref(sc.params), _p(x1), N, _p(ft), _p(fo), _p(vis), _p(h), _p(Hx), _p(Hf))
    assert np.array_equal(vis, m["vis"])
    assert np.abs(h - m["h"]).max() < 1e-10
    assert rel_err(Hx, m["Hx"]) < 1e-12 and rel_err(Hf, m["Hf"]) < 1e-12
    for j in range(N):
        o = np.zeros(3)
        hd.hd_gate(_p(np.ascontiguousarray(m["S"][j])), _p(o))
        assert np.array_equal(o[:2], m["ell"][j][:2]) and abs(o[2] - m["ell"][j][2]) < 1e-12


def test_measurement_model_xyz(hd):
    sc = Scenario(320, 240, 10)
    rng = np.random.default_rng(0)
    n = 13 + 30
    x = np.zeros(n); x[3] = 1.0; x[0:3] = [0.01, -0.02, 0.03]; x[3:7] = [0.999, 0.01, -0.02, 0.015]
    x[13:] = sc.points.ravel()
    A = rng.normal(size=(n, n)); P = A @ A.T * 1e-4
    ft = np.ones(10, np.int32); fo = (13 + 3 * np.arange(10)).astype(np.int32)
    f = OracleFilter(sc.params)
    f.set_state(x, P, ft, fo, sc.descriptors)
    f.measure()
    m = f.get_measure()
    vis = np.zeros(10, np.uint8); h = np.zeros((10, 2)); Hx = np.zeros((10, 14)); Hf = np.zeros((10, 12))
    hd.hd_measure(ctypes.byref(sc.params), _p(x), 10, _p(ft), _p(fo), _p(vis), _p(h), _p(Hx), _p(Hf))
    assert vis.sum() > 0 and np.array_equal(vis, m["vis"])
    assert np.abs(h - m["h"]).max() < 1e-10 and rel_err(Hx, m["Hx"]) < 1e-12 and rel_err(Hf, m["Hf"]) < 1e-12
