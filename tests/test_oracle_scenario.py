"""Host-logic tests on the CPU: scenario generator, oracle invariants, the algebra the CUDA path relies on."""
import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.scenario import Scenario
from oracle.oracle_lib import OracleFilter, update_dense


@pytest.fixture(scope="module")
def tracked():
    sc = Scenario(320, 240, 30)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    f = OracleFilter(sc.params)
    f.set_state(x, P, ft, fo, desc)
    infos = []
    for t in range(1, 13):
        kp, ds = sc.frame(t)
        infos.append(f.step(kp, ds))
    return sc, f, infos


def test_numpy_map_init_equals_reference_sequential_add():
    sc = Scenario(320, 240, 12)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    f = OracleFilter(sc.params)
    f.init()
    for i in range(sc.N):
        f.add_feature(uv0[i], desc[i])
    xo, Po = f.get_state()
    assert np.abs(xo - x).max() < 1e-14 and rel_err(P, Po) < 1e-13
    feats = f.get_features()
    assert np.array_equal(feats["off"], fo) and np.array_equal(feats["type"], ft)


def test_filter_tracks_and_invariants(tracked):
    sc, f, infos = tracked
    last = infos[-1]
    assert last["n_predicted"] == sc.N
    assert last["n_inliers"] + last["n_rescued"] >= 0.6 * sc.N
    x, P = f.get_state()
    assert abs(np.linalg.norm(x[3:7]) - 1.0) < 1e-12          # q renormalised by update()
    assert rel_err(P, P.T) < 1e-12                             # symmetrised
    assert np.linalg.eigvalsh(0.5 * (P + P.T)).min() > -1e-12
    # the map scale is unobservable; direction of travel is: x grows with the true motion
    assert x[0] > 0


def test_ransac_uses_few_hypotheses(tracked):
    _, _, infos = tracked
    assert max(i["n_hypotheses"] for i in infos) <= 16


def test_joint_update_equals_schur_complement():
    """(I - K H) P with K = P H^T (H P H^T + s I)^-1 (Update.cpp:92-109,214-218) equals P - W W^T with
    S = U^T U, W = P H^T U^-1: the reformulation the CUDA path uses."""
    rng = np.random.default_rng(3)
    n, k = 61, 10
    A = rng.normal(size=(n, n))
    P = A @ A.T / n
    H = np.zeros((k, n))
    for a in range(k // 2):
        H[2 * a:2 * a + 2, :7] = rng.normal(size=(2, 7))
        H[2 * a:2 * a + 2, 13 + 6 * a:19 + 6 * a] = rng.normal(size=(2, 6))
    Pref, K = update_dense(P, H, 1.0)
    S = H @ P @ H.T + np.eye(k)
    U = np.linalg.cholesky(S).T
    W = P @ H.T @ np.linalg.inv(U)
    assert rel_err(P - W @ W.T, Pref) < 1e-12


def test_scenario_is_deterministic():
    a = Scenario(320, 240, 20).frame(7)
    b = Scenario(320, 240, 20).frame(7)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = Scenario(320, 240, 20, seed_offset=1).frame(7)
    assert not np.array_equal(a[0], c[0])
