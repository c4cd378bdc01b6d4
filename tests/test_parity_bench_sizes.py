"""Parity AT THE BENCHMARKED SIZES (BASELINE.json configs 3 and 4), through the C ABI.

* C3 (640x480, 500 inverse-depth features, n = 3013): the B200 path against the REFERENCE's own code (oracle/_ref/libref.so,
  the reference's sources compiled against oracle/cvshim) -- one cold frame from the initial map and steady-state frames that
  start from the GPU filter's own state after 30 frames (k ~ 640 low-innovation rows, ~60 high-innovation rows: the frames
  bench.py times).  Matched feature set, matched pixels, inlier / outlier / rescued sets EXACT; state and the full
  3013 x 3013 covariance within 1e-9 relative (max|a-b| / max|b|, conftest.rel_err).
* C4 shape: 8 filters of 200 features in ONE handle (the batched kernels, paired slab TRSM, k ~ 250-300) against 8 oracles.

The reference costs ~20 s of CPU per C3 frame (its covariance update is a literal n^3 product)."""
import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
from oracle import ref_lib
from oracle.oracle_lib import OracleFilter

pytestmark = pytest.mark.gpu
TOL = 1e-9


def reference_frame(r, kp, ds):
    """the reference's own phase functions in EKF::step order (oracle/ref/ref_glue.cpp; bit-identical to EKF::step, see
    tests/test_s3_sequence.py::test_glued_reference_equals_its_own_step); returns the frame's sets"""
    r.predict(); r.measure(); r.match(kp, ds); r.ransac(); r.update_li(); r.rescue(); r.update_hi()
    m, sets = r.get_match(), r.get_sets()
    r.update_map_features()
    return m, sets


def compare_frame(gpu, f, m, sets, x_ref, P_ref, what):
    g = gpu.feature_results(f)
    assert np.array_equal(g["matched"], m["matched"]), f"{what}: matched feature sets differ"
    mm = m["matched"].astype(bool)
    assert np.array_equal(g["z"][mm], m["z"][mm]), f"{what}: matched pixel locations differ"
    assert np.array_equal(g["dist"][mm], m["dist"][mm]), f"{what}: descriptor distances differ"
    for key in ("inlier", "outlier", "rescued"):
        assert np.array_equal(g[key], sets[key]), f"{what}: {key} sets differ"
    xg, Pg = gpu.get_state(f)
    ex, eP = rel_err(xg, x_ref), rel_err(Pg, P_ref)
    assert ex < TOL and eP < TOL, f"{what}: state rel err {ex:.3e}, covariance rel err {eP:.3e}"
    assert np.array_equal(Pg, Pg.T), f"{what}: GPU covariance must stay exactly symmetric"
    return ex, eP, int(mm.sum()), int(sets["inlier"].sum()), int(sets["rescued"].sum())


def test_c3_against_the_reference():
    if not ref_lib.available(build=False):
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    sc = Scenario(640, 480, 500)
    x, P, ft, fo, desc, _ = sc.init_map()
    gpu = EkfBatch(sc.params, 1, 500, 1256)
    gpu.set_state(0, x, P, ft, fo, desc)
    r = ref_lib.ReferenceFilter(sc.params)
    r.set_state(x, P, ft, fo, desc)
    report = []
    # (a) the cold frame from the initial map
    kp, ds = sc.frame(1)
    gpu.set_keypoints(0, kp, ds); gpu.step()
    m, sets = reference_frame(r, kp, ds)
    report.append(compare_frame(gpu, 0, m, sets, *r.get_state(), "C3 frame 1"))
    # (b) steady state: 29 more frames on the GPU, then both sides continue from the GPU's state
    for t in range(2, 31):
        gpu.set_keypoints(0, *sc.frame(t)); gpu.step()
    xg, Pg = gpu.get_state(0)
    dg, _, _ = gpu.get_descriptors(0)
    gpu.set_state(0, xg, Pg, ft, fo, dg)          # (resets the hit counters on both sides)
    r.set_state(xg, Pg, ft, fo, dg)
    for t in (31, 32):
        kp, ds = sc.frame(t)
        gpu.set_keypoints(0, kp, ds); gpu.step()
        m, sets = reference_frame(r, kp, ds)
        report.append(compare_frame(gpu, 0, m, sets, *r.get_state(), f"C3 frame {t}"))
        d, tp, tm = gpu.get_descriptors(0)
        fr = r.get_features()
        assert np.array_equal(d, fr["desc"]) and np.array_equal(tp, fr["times_predicted"]) and np.array_equal(tm, fr["times_matched"])
    assert gpu.frame_info(0)["status"] == 0
    assert report[-1][3] > 250, "steady-state frames must have the benchmark's update size (k ~ 640)"
    for (ex, eP, nm, ni, nr), t in zip(report, (1, 31, 32)):
        print(f"C3 frame {t}: {nm} matches / {ni} inliers / {nr} rescued identical to the reference; state {ex:.2e}, cov {eP:.2e}")
    r.close()


def test_c4_shape_batch_against_oracles():
    """8 filters x 200 features in one handle, 4 frames, every filter against its own oracle."""
    F, N, T = 8, 200, 4
    scs = [Scenario(640, 480, N, seed_offset=f) for f in range(F)]
    gpu = EkfBatch(scs[0].params, F, N, 2 * N + 256)
    orcs = []
    for f, sc in enumerate(scs):
        x, P, ft, fo, desc, _ = sc.init_map()
        o = OracleFilter(sc.params); o.set_state(x, P, ft, fo, desc)
        orcs.append(o)
        gpu.set_state(f, x, P, ft, fo, desc)
    worst, kmax = (0.0, 0.0), 0
    for t in range(1, T + 1):
        frames = [sc.frame(t) for sc in scs]
        for f, (kp, ds) in enumerate(frames):
            gpu.set_keypoints(f, kp, ds)
        gpu.step()
        for f, (kp, ds) in enumerate(frames):
            o = orcs[f]
            o.step(kp, ds)
            ma, ro = o.get_match(), o.get_ransac()
            sets = dict(inlier=ro["inlier"], outlier=ro["outlier"], rescued=o.get_rescue())
            g = gpu.feature_results(f)
            mm = ma["matched"].astype(bool)
            assert np.array_equal(g["kp"][mm], ma["kp"][mm]), f"filter {f} frame {t}: matched keypoint indices differ"
            ex, eP, nm, ni, nr = compare_frame(gpu, f, ma, sets, *o.get_state(), f"C4 filter {f} frame {t}")
            info = gpu.frame_info(f)
            assert info["n_hypotheses"] == ro["n_hyp"] and info["best_hypothesis"] == ro["best"] and info["status"] == 0
            worst = (max(worst[0], ex), max(worst[1], eP))
            kmax = max(kmax, 2 * ni)
    assert kmax > 192, "the batch must exercise multi-block updates (k > 3 blocks of 64)"
    print(f"C4 shape: {F} filters x {T} frames identical sets; worst state {worst[0]:.2e}, cov {worst[1]:.2e}; largest k {kmax}")
