"""Device-side map management (kernels in csrc/ekf_map.cuh) against the CPU oracle, whose restatement is
pinned against the reference's own MapManagement.cpp / AddMapFeature.cpp in tests/test_reference_pin.py.
Feature sets, layout (type, covarianceMatrixPos), descriptors and hit counters exact; state and covariance 1e-9."""
import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.params import MapPolicy
from openekfmonoslam_b200.scenario import Scenario
from oracle import oracle_lib
from oracle.oracle_lib import OracleFilter
from test_gpu_parity import compare_state, phase_by_phase

pytestmark = pytest.mark.gpu


def map_pair(N=40, behind=(5, 17, 30), F=1, cap=None, W=320, H=240):
    sc = Scenario(W, H, N)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    x = x.copy()
    for i in behind:
        x[fo[i] + 3] += np.pi          # behind the camera: never predicted -> "unseen"
    orc = OracleFilter(sc.params)
    orc.set_state(x, P, ft, fo, desc)
    gpu = EkfBatch(sc.params, F, cap or N + 8, 4 * N + 64)
    for f in range(F):
        gpu.set_state(f, x, P, ft, fo, desc)
    return sc, orc, gpu


def same_map(orc, gpu, what, f=0):
    assert orc.dims() == gpu.dims(f), what
    xo, Po = orc.get_state()
    xg, Pg = gpu.get_state(f)
    assert rel_err(xg, xo) < 1e-9 and rel_err(Pg, Po) < 1e-9, what
    assert np.array_equal(Pg, Pg.T), f"{what}: covariance must stay exactly symmetric"
    fo = orc.get_features()
    t, off = gpu.feature_layout(f)
    assert np.array_equal(fo["type"], t) and np.array_equal(fo["off"], off), what
    d, tp, tm = gpu.get_descriptors(f)
    assert np.array_equal(d, fo["desc"]) and np.array_equal(tp, fo["times_predicted"]) and np.array_equal(tm, fo["times_matched"])


def test_map_management_remove_convert_add():
    """Same script as tests/test_reference_pin.py::test_map_management_remove_convert_add, GPU against the oracle:
    frame 2 removes two bad and three unseen features and adds four new ones, frames >= 3 convert one inverse-depth
    feature to XYZ each; all other phases keep running on the re-laid-out map."""
    sc, orc, gpu = map_pair()
    rng = np.random.default_rng(7)
    for t in range(1, 9):
        phase_by_phase(sc, orc, gpu, t)
        if t == 2:
            f_ = orc.get_features()
            tp, tm = np.full_like(f_["times_predicted"], 5), np.full_like(f_["times_matched"], 5)
            tm[[3, 11]] = 1
            tp[[5, 17, 30]] = 0; tm[[5, 17, 30]] = 0
            orc.set_hit_counters(tp, tm); gpu.set_hit_counters(0, tp, tm)
        pol = MapPolicy(min_matches_per_image=60 if t == 2 else 0, max_map_size=240,
                        good_feature_matching_percent=0.5 if t == 2 else 0.0,
                        linearity_index_threshold=1e9 if t >= 3 else 1e-9)
        N0 = orc.dims()[1]
        mo = orc.get_measure()      # indexed by the numbering before the map changes
        needed, removed, conv = orc.map_management(pol)
        res = gpu.map_management(pol)[0]
        assert res["new_features_needed"] == needed and res["converted"] == conv
        assert np.array_equal(gpu.removed_flags(0, N0), removed)
        assert res["n_removed_bad"] == (removed == 1).sum() and res["n_removed_unseen"] == (removed == 2).sum()
        same_map(orc, gpu, f"map management {t}")
        if t == 2:
            assert orc.dims()[1] == N0 - 5
            # buildImageMask: every prediction's ellipse in black on white
            mask = gpu.new_feature_mask(0)
            ref = np.full_like(mask, 255)
            black = np.zeros_like(mask)
            for i in np.flatnonzero(mo["vis"]):
                oracle_lib.draw_uncertainty_ellipse(black, mo["h"][i, 0], mo["h"][i, 1], mo["S"][i].reshape(2, 2),
                                                    2 * (mask.shape[0] + mask.shape[1]))
            ref[black != 0] = 0
            assert np.array_equal(mask, ref)
            uv = np.stack([rng.uniform(20, 300, 4).round(), rng.uniform(20, 220, 4).round()], 1)
            dd = rng.integers(0, 256, (4, 32), dtype=np.uint8)
            for a in range(4):
                orc.add_feature(uv[a], dd[a])
            gpu.add_features(0, uv, dd)
            same_map(orc, gpu, "add after removal")
    assert (orc.get_features()["type"] == 1).sum() == 6


def test_add_features_from_empty_map():
    """EKF::init's path (E/EKF.cpp:170-237): initState / initCovariance, then a batch of new features on the device."""
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    orc = OracleFilter(sc.params)
    orc.init()
    x0, P0 = orc.get_state()
    gpu = EkfBatch(sc.params, 1, 40, 256)
    gpu.set_state(0, x0, P0, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 32), np.uint8))
    for a in range(30):
        orc.add_feature(uv0[a], desc[a])
    gpu.add_features(0, uv0[:17], desc[:17])      # two batches: the second one sees the first one's rows
    gpu.add_features(0, uv0[17:30], desc[17:30])
    same_map(orc, gpu, "30 features added to the initial state")
    xs, Ps = sc.init_map()[:2]
    assert rel_err(gpu.get_state(0)[1], Ps) < 1e-12
    with pytest.raises(Exception):
        gpu.add_features(0, uv0[:11], desc[:11])  # capacity 40


def test_map_management_batched_filters_differ():
    """Two filters of one handle with different maps: only one of them changes, both must stay right."""
    sc, orc, gpu = map_pair(N=30, behind=(4,), F=2)
    sc2 = Scenario(320, 240, 30)
    x, P, ft, fo, desc, _ = sc2.init_map()
    orc2 = OracleFilter(sc2.params)
    orc2.set_state(x, P, ft, fo, desc)
    gpu.set_state(1, x, P, ft, fo, desc)
    for t in range(1, 4):
        kp, ds = sc.frame(t)
        orc.step(kp, ds); orc2.step(kp, ds)
        gpu.set_keypoints(0, kp, ds); gpu.set_keypoints(1, kp, ds)
        gpu.step()
        pol = MapPolicy(min_matches_per_image=100, max_map_size=0, max_map_features_count=20, always_remove_unseen=0,
                        good_feature_matching_percent=0.0, linearity_index_threshold=1e-9)
        a = orc.map_management(pol); b = orc2.map_management(pol)
        res = gpu.map_management(pol)
        assert res[0]["new_features_needed"] == a[0] and res[1]["new_features_needed"] == b[0]
        same_map(orc, gpu, f"filter 0 frame {t}", 0)
        same_map(orc2, gpu, f"filter 1 frame {t}", 1)
    assert gpu.dims(0)[1] == 29 and gpu.dims(1)[1] == 30


def test_raster_ellipse_hook_matches_oracle():
    """ekfb_raster_ellipse (the host side's stamp source) against the oracle's cv2-pinned drawUncertaintyEllipse2D."""
    gpu = EkfBatch(Scenario(320, 240, 4).params, 1, 4, 64)
    rng = np.random.default_rng(3)
    for _ in range(40):
        A = rng.normal(size=(2, 2)) * rng.uniform(1, 12)
        S = A @ A.T + np.eye(2)
        cx, cy = rng.uniform(-10, 110, 2)
        a = gpu.raster_ellipse(np.zeros((96, 128), np.uint8), cx, cy, S, 400, 255)
        b = oracle_lib.draw_uncertainty_ellipse(np.zeros((96, 128), np.uint8), cx, cy, S, 400)
        assert np.array_equal(a, b)
    st = gpu.raster_ellipse(np.full((64, 64), 255, np.uint8), 32, 32, np.diag([10.0, 10.0]), 1000, 0)
    assert st[32, 32] == 0 and st[0, 0] == 255


@pytest.mark.parametrize("seed", range(6))
def test_map_management_policy_branches(seed):
    """Random hit counters and every branch of the removal policy (E/EKF.cpp:580-589), with and without a conversion:
    the same script as tests/test_reference_pin.py::test_map_management_policy_branches, GPU against the oracle."""
    rng = np.random.default_rng(100 + seed)
    behind = tuple(sorted(rng.choice(40, size=4, replace=False)))
    sc, orc, gpu = map_pair(40, behind)
    for t in (1, 2):
        phase_by_phase(sc, orc, gpu, t)
    N = orc.dims()[1]
    tp = rng.integers(1, 9, N).astype(np.int32)
    tm = np.minimum(tp, rng.integers(0, 9, N)).astype(np.int32)
    tp[list(behind)] = 0; tm[list(behind)] = 0
    orc.set_hit_counters(tp, tm); gpu.set_hit_counters(0, tp, tm)
    branch = seed % 4
    pol = MapPolicy(min_matches_per_image=int(rng.integers(20, 80)),
                    max_map_features_count=int(rng.integers(20, 45)) if branch == 1 else 0,
                    max_map_size=int(rng.integers(150, 260)) if branch == 2 else 0,
                    always_remove_unseen=1 if branch == 0 else 0,
                    good_feature_matching_percent=float(rng.choice([0.3, 0.5, 0.7])),
                    linearity_index_threshold=1e9 if seed % 2 else 0.1)
    needed, removed, conv = orc.map_management(pol)
    res = gpu.map_management(pol)[0]
    assert (res["new_features_needed"], res["converted"]) == (needed, conv)
    assert np.array_equal(gpu.removed_flags(0, N), removed)
    same_map(orc, gpu, f"policy branch {branch}")
    phase_by_phase(sc, orc, gpu, 3)


def test_map_management_large_map():
    """N = 300 (n = 1813): the plan kernel's ordered scan runs over more than one block-wide chunk, the gather moves a
    3-MB covariance, 50 features are added in one batch."""
    rng = np.random.default_rng(5)
    behind = tuple(sorted(rng.choice(300, size=25, replace=False)))
    sc, orc, gpu = map_pair(300, behind, cap=360, W=640, H=480)
    kp, ds = sc.frame(1)
    orc.step(kp, ds)
    gpu.set_keypoints(0, kp, ds); gpu.step()
    N = orc.dims()[1]
    tp = rng.integers(1, 9, N).astype(np.int32)
    tm = np.minimum(tp, rng.integers(0, 9, N)).astype(np.int32)
    tp[list(behind)] = 0; tm[list(behind)] = 0
    orc.set_hit_counters(tp, tm); gpu.set_hit_counters(0, tp, tm)
    pol = MapPolicy(min_matches_per_image=400, max_map_features_count=0, max_map_size=0, always_remove_unseen=1,
                    good_feature_matching_percent=0.5, linearity_index_threshold=1e9)
    needed, removed, conv = orc.map_management(pol)
    res = gpu.map_management(pol)[0]
    assert (res["new_features_needed"], res["converted"]) == (needed, conv) and conv >= 0
    assert np.array_equal(gpu.removed_flags(0, N), removed) and (removed == 1).sum() > 50 and (removed == 2).sum() == 25
    same_map(orc, gpu, "large map after removal + conversion")
    uv = np.stack([rng.uniform(30, 600, 50).round(), rng.uniform(30, 450, 50).round()], 1)
    dd = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    for a in range(50):
        orc.add_feature(uv[a], dd[a])
    gpu.add_features(0, uv, dd)
    same_map(orc, gpu, "large map after a batch of 50 new features")
    kp, ds = sc.frame(2)
    io = orc.step(kp, ds)
    gpu.set_keypoints(0, kp, ds); gpu.step()
    ig = gpu.frame_info(0)
    assert (io["n_matches"], io["n_inliers"], io["n_rescued"]) == (ig["n_matches"], ig["n_inliers"], ig["n_rescued"])
    compare_state(orc, gpu, "large map, next frame")
