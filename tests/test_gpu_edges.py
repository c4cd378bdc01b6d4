"""Edge cases of the per-frame path, GPU against the oracle: frames without keypoints, without matches, with a single
match, a one-feature map, an empty map, features that are never predicted, and the capacity errors of the C ABI."""
import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.capi import EkfBatch, EkfError
from openekfmonoslam_b200.scenario import Scenario
from oracle.oracle_lib import OracleFilter
from test_gpu_parity import compare_state, make_pair

pytestmark = pytest.mark.gpu
NO_KP = (np.zeros((0, 2), np.float32), np.zeros((0, 32), np.uint8))


def step_both(orc, gpu, kp, ds):
    io = orc.step(kp, ds)
    gpu.set_keypoints(0, kp, ds)
    gpu.step()
    ig = gpu.frame_info(0)
    for k in ("n_predicted", "n_matches", "n_inliers", "n_rescued"):
        assert io[k] == ig[k], (k, io[k], ig[k])
    assert ig["status"] == 0
    compare_state(orc, gpu, "edge")
    d, tp, tm = gpu.get_descriptors(0)
    fo = orc.get_features()
    assert np.array_equal(tp, fo["times_predicted"]) and np.array_equal(tm, fo["times_matched"]) and np.array_equal(d, fo["desc"])
    return ig


def test_frames_without_keypoints_and_without_matches():
    sc, orc, gpu = make_pair(320, 240, 30)
    step_both(orc, gpu, *sc.frame(1))
    ig = step_both(orc, gpu, *NO_KP)                       # nothing detected: prediction only
    assert ig["n_matches"] == 0 and ig["n_predicted"] > 0
    rng = np.random.default_rng(3)                          # keypoints, but none that belongs to the map
    kp = np.stack([rng.integers(0, 320, 40), rng.integers(0, 240, 40)], 1).astype(np.float32)
    ds = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    step_both(orc, gpu, kp, ds)
    step_both(orc, gpu, *sc.frame(4))                       # and the filter picks the track up again
    assert gpu.frame_info(0)["n_inliers"] > 5


def test_single_keypoint_frame():
    """one keypoint: a handful of gates contain it (the reference lets several features claim the same keypoint), the update
    has only a few rows"""
    sc, orc, gpu = make_pair(320, 240, 30)
    kp, ds, owner, _ = sc.frame(1, with_truth=True)
    one = np.flatnonzero(owner == 7)[:1]
    ig = step_both(orc, gpu, kp[one], ds[one])
    assert 1 <= ig["n_matches"] <= 4


def test_one_feature_map_and_empty_map():
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, _ = sc.init_map()
    n1 = 13 + 6
    orc = OracleFilter(sc.params); orc.set_state(x[:n1], P[:n1, :n1], ft[:1], fo[:1], desc[:1])
    gpu = EkfBatch(sc.params, 1, 4, 256); gpu.set_state(0, x[:n1], P[:n1, :n1], ft[:1], fo[:1], desc[:1])
    for t in (1, 2):
        step_both(orc, gpu, *sc.frame(t))
    orc = OracleFilter(sc.params); orc.init()
    x0, P0 = orc.get_state()
    gpu = EkfBatch(sc.params, 1, 4, 256)
    gpu.set_state(0, x0, P0, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 32), np.uint8))
    for t in (1, 2):                                         # no map: the frame is the motion prediction alone
        kp, ds = sc.frame(t)
        orc.step(kp, ds)
        gpu.set_keypoints(0, kp, ds)
        gpu.step()
        xo, Po = orc.get_state(); xg, Pg = gpu.get_state(0)
        assert rel_err(xg, xo) < 1e-12 and rel_err(Pg, Po) < 1e-12 and gpu.frame_info(0)["n_matches"] == 0


def test_never_predicted_features_stay_untouched():
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, _ = sc.init_map()
    x = x.copy()
    for i in range(0, 30, 3):
        x[fo[i] + 3] += np.pi                                # behind the camera
    orc = OracleFilter(sc.params); orc.set_state(x, P, ft, fo, desc)
    gpu = EkfBatch(sc.params, 1, 30, 256); gpu.set_state(0, x, P, ft, fo, desc)
    for t in (1, 2, 3):
        step_both(orc, gpu, *sc.frame(t))
    r = gpu.feature_results(0)
    assert not r["vis"][::3].any() and not r["matched"][::3].any()
    assert (gpu.get_descriptors(0)[1][::3] == 0).all()


def test_capacity_and_argument_errors():
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, _ = sc.init_map()
    gpu = EkfBatch(sc.params, 1, 20, 50)
    with pytest.raises(EkfError):
        gpu.set_state(0, x, P, ft, fo, desc)                 # 30 features into a handle created for 20
    with pytest.raises(EkfError):
        gpu.set_state(1, x[:13], P[:13, :13], ft[:0], fo[:0], desc[:0])   # no such filter
    kp, ds = sc.frame(1)
    with pytest.raises(EkfError):
        gpu.set_keypoints(0, kp, ds)                         # 60 keypoints, capacity 50
    with pytest.raises(EkfError):
        gpu.update(2)
    n = 13 + 6 * 20
    gpu.set_state(0, x[:n], P[:n, :n], ft[:20], fo[:20], desc[:20])       # and the handle is still usable
    gpu.set_keypoints(0, kp[:50], ds[:50]); gpu.step()
    assert gpu.frame_info(0)["status"] == 0


@pytest.mark.parametrize("packed", [False, True])
def test_keypoints_of_all_filters_in_one_call(packed):
    """ekfb_set_keypoints_batch / ekfb_set_keypoints_packed: three filters of one handle fed different frames (one of them
    empty) in one call"""
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, _ = sc.init_map()
    gpu = EkfBatch(sc.params, 3, 30, 256)
    orcs = []
    for f in range(3):
        gpu.set_state(f, x, P, ft, fo, desc)
        o = OracleFilter(sc.params); o.set_state(x, P, ft, fo, desc)
        orcs.append(o)
    for t in (1, 2):
        fr = [sc.frame(t), sc.frame(t + 5), NO_KP if t == 2 else sc.frame(t + 9)]
        keep = [(np.ascontiguousarray(k), np.ascontiguousarray(d)) for k, d in fr]
        if packed:
            xy = np.ascontiguousarray(np.concatenate([k.reshape(-1, 2) for k, _ in keep]), np.float32)
            dd = np.ascontiguousarray(np.concatenate([d.reshape(-1, 32) for _, d in keep]), np.uint8)
            gpu.set_keypoints_packed_raw(xy.ctypes.data, dd.ctypes.data, np.cumsum([0] + [len(k) for k, _ in keep]))
        else:
            gpu.set_keypoints_batch_raw([k.ctypes.data for k, _ in keep], [d.ctypes.data for _, d in keep], [len(k) for k, _ in keep])
        gpu.step()
        for f in range(3):
            io = orcs[f].step(*fr[f])
            ig = gpu.frame_info(f)
            assert (io["n_matches"], io["n_inliers"], io["n_rescued"]) == (ig["n_matches"], ig["n_inliers"], ig["n_rescued"]), (t, f)
            xo, Po = orcs[f].get_state(); xg, Pg = gpu.get_state(f)
            assert rel_err(xg, xo) < 1e-9 and rel_err(Pg, Po) < 1e-9
