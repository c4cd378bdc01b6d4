"""Pins the oracle against the REFERENCE's own code: oracle/_ref/libref.so is built from the unmodified sources
under /root/reference (kalmanFilter/modules/{Core,1PointRansacEKF,Gui}) against oracle/cvshim (oracle/ref/Makefile).
Every phase of EKF::step, the add-feature path and whole frames through the reference's own EKF::step must agree
with the oracle: sets exactly, floating point to 1e-12 (both are CPU FP64; only summation order differs).
Skipped when the reference tree (and a prebuilt libref.so) is not available."""
import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.scenario import Scenario
from oracle import ref_lib
from oracle.oracle_lib import OracleFilter

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref/libref.so not built (needs /root/reference)")
TOL = 1e-12


def pair(N, W=320, H=240, seed_offset=0):
    sc = Scenario(W, H, N, seed_offset=seed_offset)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    o = OracleFilter(sc.params)
    r = ref_lib.ReferenceFilter(sc.params)
    o.set_state(x, P, ft, fo, desc)
    r.set_state(x, P, ft, fo, desc)
    return sc, o, r, (x, P, ft, fo, desc, uv0)


def same_state(o, r, what, tol=TOL):
    xo, Po = o.get_state()
    xr, Pr = r.get_state()
    assert rel_err(xo, xr) < tol and rel_err(Po, Pr) < tol, what


def phase_by_phase(sc, o, r, t, tol=TOL):
    kp, ds = sc.frame(t)
    o.predict(); r.predict()
    same_state(o, r, f"predict {t}", tol)
    o.measure(); r.measure()
    mo, mr = o.get_measure(), r.get_measure()
    assert np.array_equal(mo["vis"], mr["vis"])
    v = mo["vis"].astype(bool)
    assert np.abs(mo["h"][v] - mr["h"][v]).max() < 1e-10
    for k in ("S", "Hx", "Hf"):
        assert rel_err(mo[k][v], mr[k][v]) < tol, k
    o.match(kp, ds); r.match(kp, ds)
    ao, ar = o.get_match(), r.get_match()
    assert np.array_equal(ao["matched"], ar["matched"])
    m = ao["matched"].astype(bool)
    assert np.array_equal(ao["z"][m], ar["z"][m]) and np.array_equal(ao["dist"][m], ar["dist"][m])
    o.ransac(); r.ransac()
    so, sr = o.get_ransac(), r.get_sets()
    assert np.array_equal(so["inlier"], sr["inlier"]) and np.array_equal(so["outlier"], sr["outlier"])
    o.update_li(); r.update_li()
    same_state(o, r, f"update LI {t}", tol)
    o.rescue(); r.rescue()
    assert np.array_equal(o.get_rescue(), r.get_sets()["rescued"])
    o.update_hi(); r.update_hi()
    same_state(o, r, f"update HI {t}", tol)
    o.update_map_features(); r.update_map_features()
    fo_, fr_ = o.get_features(), r.get_features()
    assert np.array_equal(fo_["desc"], fr_["desc"]) and np.array_equal(fo_["times_predicted"], fr_["times_predicted"])
    assert np.array_equal(fo_["times_matched"], fr_["times_matched"])


def test_add_feature_path():
    sc, o, r, (x, P, ft, fo, desc, uv0) = pair(10)
    o.init(); r.init()
    for i in range(10):
        o.add_feature(uv0[i], desc[i]); r.add_feature(uv0[i], desc[i])
    same_state(o, r, "add feature")
    r.close()


def test_every_phase_inverse_depth():
    sc, o, r, _ = pair(40)
    for t in range(1, 11):
        phase_by_phase(sc, o, r, t)
    r.close()


def test_whole_frames_through_the_references_own_step():
    """EKF::step (the reference's orchestrator, EKF.cpp:242-666) vs orc_step, 25 free-running frames."""
    sc, o, r, _ = pair(30, seed_offset=2)
    for t in range(1, 26):
        kp, ds = sc.frame(t)
        o.step(kp, ds); r.step(kp, ds)
        same_state(o, r, f"frame {t}")
        fo_, fr_ = o.get_features(), r.get_features()
        assert np.array_equal(fo_["desc"], fr_["desc"]) and np.array_equal(fo_["times_matched"], fr_["times_matched"])
    r.close()


def test_every_phase_mixed_xyz_features():
    sc, o, r, (x, P, ft, fo, desc, _) = pair(18)
    for t in range(1, 4):
        o.step(*sc.frame(t))
    x, P = o.get_state()
    xs, types, offs, blocks, n_new = [x[:13]], [], [], [], 13
    for j in range(sc.N):
        oj = 13 + 6 * j
        y = x[oj:oj + 6]
        if j % 2 == 0:
            th, ph, rho = y[3], y[4], y[5]
            m = np.array([np.cos(ph) * np.sin(th), -np.sin(ph), np.cos(ph) * np.cos(th)])
            dth = np.array([np.cos(ph) * np.cos(th), 0.0, -np.cos(ph) * np.sin(th)])
            dph = np.array([-np.sin(ph) * np.sin(th), -np.cos(ph), -np.sin(ph) * np.cos(th)])
            J = np.hstack([np.eye(3), (dth / rho)[:, None], (dph / rho)[:, None], (-m / rho ** 2)[:, None]])
            xs.append(y[:3] + m / rho); blocks.append((oj, J)); types.append(1); offs.append(n_new); n_new += 3
        else:
            xs.append(y); blocks.append((oj, np.eye(6))); types.append(2); offs.append(n_new); n_new += 6
    T = np.zeros((n_new, x.shape[0]))
    T[:13, :13] = np.eye(13)
    for (oj, J), off in zip(blocks, offs):
        T[off:off + J.shape[0], oj:oj + 6] = J
    P2 = T @ P @ T.T
    P2 = 0.5 * (P2 + P2.T)
    x2 = np.concatenate(xs)
    types = np.array(types, np.int32); offs = np.array(offs, np.int32)
    d = o.get_features()["desc"]
    o.set_state(x2, P2, types, offs, d)
    r.set_state(x2, P2, types, offs, d)
    # the hand-converted XYZ covariance has a wide dynamic range (1/rho^2 factors): S_i = H P H^T cancels, so the two
    # CPU summation orders differ at ~1e-11; everything discrete must still be identical
    for t in range(4, 9):
        phase_by_phase(sc, o, r, t, tol=1e-9)
    r.close()


# ---- map management: the oracle's restatement against the reference's own MapManagement.cpp ----
def map_scenario(N=40, behind=(5, 17, 30)):
    """A map whose features `behind` lie behind the camera (never predicted -> "unseen")."""
    from openekfmonoslam_b200.params import MapPolicy
    sc, o, r, (x, P, ft, fo, desc, uv0) = pair(N)
    x = x.copy()
    for i in behind:
        x[fo[i] + 3] += np.pi
    o.set_state(x, P, ft, fo, desc)
    r.set_state(x, P, ft, fo, desc)
    return sc, o, r, MapPolicy


def same_map(o, r, what, tol=TOL):
    assert o.dims() == r.dims(), what
    same_state(o, r, what, tol)
    fo_, fr_ = o.get_features(), r.get_features()
    t, off = r.get_layout()
    assert np.array_equal(fo_["type"], t) and np.array_equal(fo_["off"], off), what
    for k in ("desc", "times_predicted", "times_matched"):
        assert np.array_equal(fo_[k], fr_[k]), (what, k)


def test_map_management_remove_convert_add():
    """removeBadMapFeatures, removal of the unseen features under the MaxMapSize policy, one inverse-depth -> XYZ conversion
    per frame and the reference's own new-feature selection + add, frame after frame (E/EKF.cpp:572-612)."""
    sc, o, r, MapPolicy = map_scenario()
    for t in range(1, 9):
        phase_by_phase(sc, o, r, t, 1e-9 if t > 2 else TOL)   # XYZ features make S ill-conditioned: looser state tolerance
        if t == 2:   # features 3 and 11 become "bad": matched in 1 of 5 predictions
            f_ = o.get_features()
            tp, tm = np.full_like(f_["times_predicted"], 5), np.full_like(f_["times_matched"], 5)
            tm[[3, 11]] = 1
            tp[[5, 17, 30]] = 0; tm[[5, 17, 30]] = 0      # never predicted: 0/0 is not "bad"
            o.set_hit_counters(tp, tm); r.set_hit_counters(tp, tm)
        # t = 1: nothing to do; t = 2: bad + unseen features go; t >= 3: one conversion per frame (huge threshold)
        pol = MapPolicy(min_matches_per_image=60 if t == 2 else 0, max_map_size=240,
                        good_feature_matching_percent=0.5 if t == 2 else 0.0,
                        linearity_index_threshold=1e9 if t >= 3 else 1e-9)
        r.set_policy(pol)
        N0 = o.dims()[1]
        needed, removed, conv = o.map_management(pol)
        assert needed == r.map_management()
        same_map(o, r, f"map management {t}", 1e-9)
        if t == 1:
            assert not removed.any() and conv == -1
        if t == 2:
            assert sorted(np.flatnonzero(removed == 1)) == [3, 11] and sorted(np.flatnonzero(removed == 2)) == [5, 17, 30]
            assert o.dims()[1] == N0 - 5 and needed > 0
            kp, ds = sc.frame(t)
            uv, dd = r.detect_new(kp, ds, 4)     # the reference's zone-balanced selection (libc rand)
            assert len(uv) == 4
            for a in range(len(uv)):
                o.add_feature(uv[a], dd[a]); r.add_feature(uv[a], dd[a])
            same_map(o, r, "add after removal", 1e-9)
        if t >= 3:
            assert conv == t - 3      # the first remaining inverse-depth feature each frame
    assert (o.get_features()["type"] == 1).sum() == 6
    r.close()


@pytest.mark.parametrize("seed", range(6))
def test_map_management_policy_branches(seed):
    """Random hit counters and every branch of the removal policy of E/EKF.cpp:580-589 (AlwaysRemoveUnseenMapFeatures,
    MaxMapFeaturesCount, MaxMapSize, none), with and without a conversion, oracle against the reference's own functions."""
    rng = np.random.default_rng(100 + seed)
    behind = tuple(sorted(rng.choice(40, size=4, replace=False)))
    sc, o, r, MapPolicy = map_scenario(40, behind)
    for t in (1, 2):
        phase_by_phase(sc, o, r, t)
    N = o.dims()[1]
    tp = rng.integers(1, 9, N).astype(np.int32)
    tm = np.minimum(tp, rng.integers(0, 9, N)).astype(np.int32)
    tp[list(behind)] = 0; tm[list(behind)] = 0          # the unseen features were never predicted (0/0 is not "bad")
    o.set_hit_counters(tp, tm); r.set_hit_counters(tp, tm)
    branch = seed % 4
    pol = MapPolicy(min_matches_per_image=int(rng.integers(20, 80)),
                    max_map_features_count=int(rng.integers(20, 45)) if branch == 1 else 0,
                    max_map_size=int(rng.integers(150, 260)) if branch == 2 else 0,
                    always_remove_unseen=1 if branch == 0 else 0,
                    good_feature_matching_percent=float(rng.choice([0.3, 0.5, 0.7])),
                    linearity_index_threshold=1e9 if seed % 2 else 0.1)
    r.set_policy(pol)
    needed, removed, conv = o.map_management(pol)
    assert needed == r.map_management()
    same_map(o, r, f"policy branch {branch}", 1e-9)
    bad = (tm.astype(np.float32) / np.maximum(tp, 1).astype(np.float32) < pol.good_feature_matching_percent) & (tp > 0)
    assert np.array_equal(removed == 1, bad)
    if branch == 3:
        assert not (removed == 2).any()
    phase_by_phase(sc, o, r, 3, 1e-9)                  # and both keep running on the re-laid-out map
    r.close()
