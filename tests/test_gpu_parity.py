"""Parity of the CUDA hot path (through the C ABI of include/ekf_b200.h) against the CPU oracle.

Bars (BASELINE.json): matched pixel locations, masks, match / inlier / outlier / rescued sets are
compared EXACTLY; state mean and covariance must agree to 1e-9 relative, measured as
max|a-b| / max|b| over the vector / matrix (conftest.rel_err).
"""
import numpy as np
import pytest

from conftest import rel_err
from openekfmonoslam_b200.capi import EkfBatch
from openekfmonoslam_b200.scenario import Scenario
from oracle.oracle_lib import OracleFilter

pytestmark = pytest.mark.gpu
TOL = 1e-9


def make_pair(W, H, N, seed_offset=0, warm=0, max_features=None):
    sc = Scenario(W, H, N, seed_offset=seed_offset)
    x, P, ft, fo, desc, _ = sc.init_map()
    orc = OracleFilter(sc.params)
    orc.set_state(x, P, ft, fo, desc)
    for t in range(1, warm + 1):
        orc.step(*sc.frame(t))
    x, P = orc.get_state()
    feats = orc.get_features()
    if warm:  # restart the oracle from the same snapshot so the hit counters start at zero on both sides
        orc.set_state(x, P, feats["type"], feats["off"], feats["desc"])
    gpu = EkfBatch(sc.params, 1, max_features or N, 4 * N + 64)
    gpu.set_state(0, x, P, feats["type"], feats["off"], feats["desc"])
    return sc, orc, gpu


def compare_state(orc, gpu, what):
    xo, Po = orc.get_state()
    xg, Pg = gpu.get_state(0)
    ex, eP = rel_err(xg, xo), rel_err(Pg, Po)
    assert ex < TOL and eP < TOL, f"{what}: state rel err {ex:.3e}, covariance rel err {eP:.3e}"
    assert np.array_equal(Pg, Pg.T), f"{what}: GPU covariance must stay exactly symmetric"
    return ex, eP


def phase_by_phase(sc, orc, gpu, t):
    kp, ds = sc.frame(t)
    gpu.set_keypoints(0, kp, ds)
    # prediction
    orc.predict(); gpu.predict()
    compare_state(orc, gpu, f"frame {t} predict")
    # measurement prediction + Jacobians
    orc.measure(); gpu.measure()
    mo, r = orc.get_measure(), gpu.feature_results(0)
    assert np.array_equal(mo["vis"], r["vis"])
    v = mo["vis"].astype(bool)
    assert np.abs(r["h"][v] - mo["h"][v]).max() < 1e-9
    for key in ("S", "Hx", "Hf"):
        assert rel_err(r[key][v], mo[key][v]) < TOL, key
    # matching: mask, surviving keypoints, matches -- all exact
    orc.match(kp, ds); gpu.match()
    mask_o, ok_o = orc.get_mask()
    mask_g, ok_g = gpu.get_mask(0, kp.shape[0])
    assert np.array_equal(mask_o, mask_g), f"mask differs in {(mask_o != mask_g).sum()} pixels"
    assert np.array_equal(ok_o, ok_g)
    ma, r = orc.get_match(), gpu.feature_results(0)
    assert np.array_equal(ma["matched"], r["matched"])
    m = ma["matched"].astype(bool)
    assert np.array_equal(ma["kp"][m], r["kp"][m]) and np.array_equal(ma["z"][m], r["z"][m])
    assert np.array_equal(ma["dist"][m], r["dist"][m])
    # RANSAC
    orc.ransac(); gpu.ransac()
    ro, r, info = orc.get_ransac(), gpu.feature_results(0), gpu.frame_info(0)
    assert np.array_equal(ro["inlier"], r["inlier"]) and np.array_equal(ro["outlier"], r["outlier"])
    assert info["n_hypotheses"] == ro["n_hyp"] and info["best_hypothesis"] == ro["best"]
    # low-innovation update
    orc.update_li(); gpu.update(0)
    compare_state(orc, gpu, f"frame {t} update LI")
    # rescue + high-innovation update
    orc.rescue(); gpu.rescue()
    assert np.array_equal(orc.get_rescue(), gpu.feature_results(0)["rescued"])
    orc.update_hi(); gpu.update(1)
    e = compare_state(orc, gpu, f"frame {t} update HI")
    orc.update_map_features(); gpu.update_map_features()
    d, tp, tm = gpu.get_descriptors(0)
    fo = orc.get_features()
    assert np.array_equal(d, fo["desc"]) and np.array_equal(tp, fo["times_predicted"])
    assert np.array_equal(tm, fo["times_matched"])
    assert gpu.frame_info(0)["status"] == 0
    return e


def test_phase_by_phase_c2_shape():
    """BASELINE config 2 shape (320x240, 50 inverse-depth features), every phase of 6 frames."""
    sc, orc, gpu = make_pair(320, 240, 50)
    for t in range(1, 7):
        phase_by_phase(sc, orc, gpu, t)


def test_phase_by_phase_after_warmup_multi_block_update():
    """k > 64 rows so the blocked Cholesky / panel / trailing kernels all run (N = 100, n = 613)."""
    sc, orc, gpu = make_pair(640, 480, 100, warm=3)
    for t in range(4, 7):
        phase_by_phase(sc, orc, gpu, t)
    assert gpu.frame_info(0)["n_inliers"] > 40


@pytest.mark.parametrize("path", [1, 2])
def test_generic_factorisation_path(path):
    """The two paths for updates too large for the shared-memory slab TRSM must agree with the oracle as well; forced here at a
    size where k spans several blocks: 1 = right-looking factorisation over the whole augmented matrix, 2 = S-chain + blocked
    TRSM on the global-memory resident B (tensor-map fed DMMA kernel, ekf_gemm_tma.cuh)."""
    sc, orc, gpu = make_pair(640, 480, 100, warm=3)
    gpu.set_option(1, path)
    for t in range(4, 7):
        phase_by_phase(sc, orc, gpu, t)


def test_ransac_in_small_rounds_and_legacy_tile_options():
    """RANSAC hypotheses two per round (batches use four; the sequential accept / adaptive-cap replay must not depend on the
    round size), the 128x128 downdate tiles and the shallow TRSM ring, all against the oracle."""
    sc, orc, gpu = make_pair(640, 480, 100)
    gpu.set_option(7, 2)      # EKFB_OPT_RANSAC_CHUNK
    gpu.set_option(2, 0)      # EKFB_OPT_DOWNDATE_VARIANT: the cp.async kernels instead of the TMA-fed default
    gpu.set_option(4, 64)     # EKFB_OPT_DOWNDATE_SMALL_K: updates with more than 64 rows use the 128x128 kernel
    gpu.set_option(5, 2)      # EKFB_OPT_TRSM_STAGES
    hyps = []
    for t in range(1, 5):
        phase_by_phase(sc, orc, gpu, t)
        hyps.append(gpu.frame_info(0)["n_hypotheses"])
    assert max(hyps) > 2, hyps     # at least one frame needed a second round


def test_downdate_variant_128x64():
    """The alternative downdate kernel (128x64 tiles, transposed mirror store) against the oracle."""
    sc, orc, gpu = make_pair(640, 480, 100, warm=3)
    gpu.set_option(2, 1)
    for t in range(4, 6):
        phase_by_phase(sc, orc, gpu, t)


def test_whole_step_sequence_c2():
    """ekfb_step against orc_step over 40 frames, both free-running from the same initial map."""
    sc, orc, gpu = make_pair(320, 240, 50)
    for t in range(1, 41):
        kp, ds = sc.frame(t)
        io = orc.step(kp, ds)
        gpu.set_keypoints(0, kp, ds)
        gpu.step()
        ig = gpu.frame_info(0)
        for a, b in (("n_predicted", "n_predicted"), ("n_matches", "n_matches"), ("n_inliers", "n_inliers"),
                     ("n_rescued", "n_rescued"), ("n_hypotheses", "n_hypotheses")):
            assert io[a] == ig[b], f"frame {t}: {a} oracle {io[a]} gpu {ig[b]}"
        compare_state(orc, gpu, f"frame {t}")
        ro, r = orc.get_ransac(), gpu.feature_results(0)
        assert np.array_equal(ro["inlier"], r["inlier"]) and np.array_equal(orc.get_rescue(), r["rescued"])


def test_mixed_xyz_and_inverse_depth_features():
    """XYZ (dim 3) and inverse-depth (dim 6) blocks interleaved in P in map order."""
    sc = Scenario(320, 240, 24)
    x, P, ft, fo, desc, _ = sc.init_map()
    orc = OracleFilter(sc.params)
    orc.set_state(x, P, ft, fo, desc)
    for t in range(1, 4):
        orc.step(*sc.frame(t))
    x, P = orc.get_state()
    # convert every third feature to XYZ by hand: y = r_i + m(theta, phi) / rho, covariance by the Jacobian
    keep, xs, types, offs, Jrows = list(range(13)), [x[:13]], [], [], []
    n_new = 13
    blocks = []
    for j in range(sc.N):
        o = 13 + 6 * j
        y = x[o:o + 6]
        if j % 3 == 0:
            th, ph, rho = y[3], y[4], y[5]
            m = np.array([np.cos(ph) * np.sin(th), -np.sin(ph), np.cos(ph) * np.cos(th)])
            dm_dth = np.array([np.cos(ph) * np.cos(th), 0.0, -np.cos(ph) * np.sin(th)])
            dm_dph = np.array([-np.sin(ph) * np.sin(th), -np.cos(ph), -np.sin(ph) * np.cos(th)])
            J = np.hstack([np.eye(3), (dm_dth / rho)[:, None], (dm_dph / rho)[:, None], (-m / rho ** 2)[:, None]])
            xs.append(y[:3] + m / rho); blocks.append((o, J)); types.append(1); offs.append(n_new); n_new += 3
        else:
            xs.append(y); blocks.append((o, np.eye(6))); types.append(2); offs.append(n_new); n_new += 6
    T = np.zeros((n_new, x.shape[0]))
    T[:13, :13] = np.eye(13)
    for (o, J), off in zip(blocks, offs):
        T[off:off + J.shape[0], o:o + 6] = J
    P2 = T @ P @ T.T
    P2 = 0.5 * (P2 + P2.T)
    x2 = np.concatenate(xs)
    types = np.array(types, np.int32); offs = np.array(offs, np.int32)
    feats = orc.get_features()
    orc.set_state(x2, P2, types, offs, feats["desc"])
    gpu = EkfBatch(sc.params, 1, sc.N, 4 * sc.N + 64)
    gpu.set_state(0, x2, P2, types, offs, feats["desc"])
    for t in range(4, 8):
        phase_by_phase(sc, orc, gpu, t)


def test_batched_filters_match_single_filters():
    """4 independent filters in one handle (BASELINE config 4 shape, scaled down) == 4 oracle filters."""
    F, N = 4, 40
    scs = [Scenario(640, 480, N, seed_offset=i) for i in range(F)]
    gpu = EkfBatch(scs[0].params, F, N, 4 * N + 64)
    orcs = []
    for i, sc in enumerate(scs):
        x, P, ft, fo, desc, _ = sc.init_map()
        o = OracleFilter(sc.params)
        o.set_state(x, P, ft, fo, desc)
        orcs.append(o)
        gpu.set_state(i, x, P, ft, fo, desc)
    for t in range(1, 9):
        for i, sc in enumerate(scs):
            kp, ds = sc.frame(t)
            orcs[i].step(kp, ds)
            gpu.set_keypoints(i, kp, ds)
        gpu.step()
        for i in range(F):
            xo, Po = orcs[i].get_state()
            xg, Pg = gpu.get_state(i)
            assert rel_err(xg, xo) < TOL and rel_err(Pg, Po) < TOL, (t, i)
            assert np.array_equal(orcs[i].get_ransac()["inlier"], gpu.feature_results(i)["inlier"])
    recs = gpu.records()
    for i in range(F):
        xo, Po = orcs[i].get_state()
        assert rel_err(np.array(recs[i].x_cam), xo[:13]) < TOL
        assert rel_err(np.array(recs[i].P_cam).reshape(13, 13), Po[:13, :13]) < TOL


def test_device_resident_sequence_equals_host_fed():
    sc, orc, gpu = make_pair(320, 240, 30)
    frames = [sc.frame(t) for t in range(1, 9)]
    gpu.load_sequence(0, frames)
    for t, (kp, ds) in enumerate(frames):
        orc.step(kp, ds)
        gpu.select_frame(t)
        gpu.step()
        compare_state(orc, gpu, f"seq frame {t}")


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n,k", [(313, 2), (313, 100), (613, 40), (1213, 120), (3013, 200), (3013, 72), (2005, 640)])
def test_covariance_downdate_kernel(n, k, variant):
    """P - W W^T on the FP64 tensor pipe vs numpy, sizes of SURVEY 7.1 step 4 plus the C3 update sizes; exactly symmetric
    output.  variant 0 = 64x64 tiles fed by cp.async, 1 = 128x64 tiles, 2 = persistent TMA-fed kernel (tensor-map loads and
    stores, mbarrier ring, 128-byte swizzle), 3 = the same without swizzle."""
    rng = np.random.default_rng(n + k)
    A = rng.normal(size=(n, 64))
    P = A @ A.T / 64 + np.eye(n)
    Wt = rng.normal(size=(k, n)) * 0.05
    gpu = EkfBatch(Scenario(320, 240, 4).params, 1, (n - 13 + 5) // 6, 64)
    gpu.set_option(2, variant)
    out = gpu.test_downdate(P, Wt)
    ref = P - Wt.T @ Wt
    assert rel_err(out, ref) < 1e-13
    assert np.array_equal(out, out.T)


@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("k", [2, 8, 62, 64, 66, 126, 128, 130, 200, 384, 600, 640])
def test_innovation_covariance_factorisation(k, variant):
    """The S-chain alone: S = U^T U, y = U^-T nu and the inverses of the diagonal 64x64 blocks, against numpy.  k covers
    a single ragged block, exact multiples of 64 (nu alone in the last column tile) and the C3 size.  variant 0 = one
    fused launch per block step, 1 = the panel + trail launch pairs, 3 = the whole chain as one launch (tile dataflow,
    csrc/ekf_chain.cuh); every variant twice on the same handle (the one-launch chain carries a generation counter)."""
    rng = np.random.default_rng(1000 + k)
    A = rng.normal(size=(k, k + 8))
    S = A @ A.T / k + 0.5 * np.eye(k)
    d = np.exp(rng.uniform(-2, 2, size=k))          # rows of very different scale, as pixel / depth innovations have
    S = d[:, None] * S * d[None, :]
    nu = rng.normal(size=k) * d
    gpu = EkfBatch(Scenario(320, 240, 4).params, 1, max(k // 2, 4), 64)
    gpu.set_option(3, variant)
    gpu.test_factor(np.hstack([S, nu[:, None]]))
    U, Ui = gpu.test_factor(np.hstack([S, nu[:, None]]))
    L = np.linalg.cholesky(S)
    Uu, Lt = np.triu(U[:, :k]), L.T.copy()
    if variant == 1:  # the launch-pair variant keeps U_JJ only as its inverse
        for J in range((k + 63) // 64):
            Uu[64 * J:64 * J + 64, 64 * J:64 * J + 64] = 0.0
            Lt[64 * J:64 * J + 64, 64 * J:64 * J + 64] = 0.0
    if k > 64 or variant != 1:
        assert rel_err(Uu, Lt) < 1e-12
    y = np.linalg.solve(L, nu)
    assert rel_err(U[:, k], y) < 1e-11
    for J in range((k + 63) // 64):
        kb = min(64, k - 64 * J)
        blk = L.T[64 * J:64 * J + kb, 64 * J:64 * J + kb]
        assert rel_err(Ui[J, :kb, :kb], np.linalg.inv(blk)) < 1e-11, J
        assert np.array_equal(np.tril(Ui[J, :kb, :kb], -1), np.zeros((kb, kb)))


@pytest.mark.parametrize("variant", [0, 1, 3, 4])
def test_other_schain_variants_whole_update(variant):
    """One launch per block step (option 3 = 0), the panel + trail launch pairs (1), the one-launch chain (3) and the chain
    with the slab TRSM overlapped inside the same launch (4) through the whole update, against the oracle."""
    sc, orc, gpu = make_pair(640, 480, 100, warm=3)
    gpu.set_option(3, variant)
    for t in range(4, 7):
        phase_by_phase(sc, orc, gpu, t)


def test_one_launch_chain_batched_filters():
    """the one-launch chain with several CTAs per filter (4 filters -> 37 CTAs each) and with a single CTA per filter
    (more filters than SMs / 2 cannot be co-resident with more), against per-filter oracles"""
    for F in (4, 80):
        scs = [Scenario(320, 240, 40, seed_offset=f % 4) for f in range(F)]
        gpu = EkfBatch(scs[0].params, F, 40, 256)
        gpu.set_option(3, 3)
        orcs = []
        for f in range(F):
            x, P, ft, fo, desc, _ = scs[f].init_map()
            gpu.set_state(f, x, P, ft, fo, desc)
            if f < 4:
                o = OracleFilter(scs[f].params); o.set_state(x, P, ft, fo, desc)
                orcs.append(o)
        for t in range(1, 4):
            for f in range(F):
                gpu.set_keypoints(f, *scs[f].frame(t))
            gpu.step()
            for f in range(4):
                orcs[f].step(*scs[f].frame(t))
        for f in range(F):
            xo, Po = orcs[f % 4].get_state()
            xg, Pg = gpu.get_state(f)
            assert rel_err(xg, xo) < TOL and rel_err(Pg, Po) < TOL, (F, f)
            assert gpu.frame_info(f)["status"] == 0


@pytest.mark.parametrize("variant", [-1, 0, 3, 4])
def test_numeric_failure_leaves_the_filter_untouched(variant):
    """EKFB_ERR_NUMERIC: when the factorisation of an innovation covariance reports a non-positive pivot (injected here with
    EKFB_OPT_FAULT_INJECT; the reference inverts S by LU, E/Update.cpp:108, and cannot fail this way) the frame's status is set
    and that update -- and the rest of the frame's updates -- are skipped: state and covariance stay exactly as the
    prediction left them.  The next frame starts clean; `reserved` keeps the sticky flag.  variant -1 = the automatic choice (these
    updates have at most 128 rows: k_update_small)."""
    sc, orc, gpu = make_pair(320, 240, 30)
    gpu.set_option(3, variant)
    gpu.set_keypoints(0, *sc.frame(1))
    gpu.predict(); gpu.measure(); gpu.match(); gpu.ransac()
    assert gpu.frame_info(0)["n_inliers"] > 2
    x0, P0 = gpu.get_state(0)
    gpu.set_option(9, 1)
    gpu.update(0); gpu.rescue(); gpu.update(1); gpu.update_map_features()
    info = gpu.frame_info(0)
    assert info["status"] == 4 and info["reserved"] == 4
    x1, P1 = gpu.get_state(0)
    assert np.array_equal(x0, x1) and np.array_equal(P0, P1)
    gpu.set_option(9, 0)
    gpu.set_keypoints(0, *sc.frame(2)); gpu.step()
    info = gpu.frame_info(0)
    assert info["status"] == 0 and info["reserved"] == 4 and info["n_inliers"] > 2
    x2, P2 = gpu.get_state(0)
    assert not np.array_equal(P1, P2) and np.array_equal(P2, P2.T)


def test_full_size_properties_c3():
    """BASELINE config 3 size (N = 500, n = 3013): the oracle's literal update is too slow to run per
    test, so check size-independent properties of one full GPU frame: P stays exactly symmetric and
    PSD on the camera block, |q| = 1, the posterior trace does not exceed the prior trace, and every
    inlier's post-update residual shrinks."""
    sc = Scenario(640, 480, 500)
    x, P, ft, fo, desc, _ = sc.init_map()
    gpu = EkfBatch(sc.params, 1, 500, 2200)
    gpu.set_state(0, x, P, ft, fo, desc)
    for t in range(1, 4):
        kp, ds = sc.frame(t)
        gpu.set_keypoints(0, kp, ds)
        gpu.predict()
        _, Pprior = gpu.get_state(0)
        gpu.measure(); gpu.match(); gpu.ransac(); gpu.update(0); gpu.rescue(); gpu.update(1); gpu.update_map_features()
        info = gpu.frame_info(0)
        assert info["status"] == 0 and info["n_predicted"] == 500 and info["n_inliers"] > 150, info
        xg, Pg = gpu.get_state(0)
        assert np.array_equal(Pg, Pg.T)
        assert abs(np.linalg.norm(xg[3:7]) - 1.0) < 1e-12
        assert np.trace(Pg) <= np.trace(Pprior) * (1 + 1e-12)
        assert np.linalg.eigvalsh(Pg[:13, :13]).min() > -1e-15


def test_decision_margins_of_a_sequence():
    """every floating-point threshold decision of the oracle (field of view, in-frame, foci gate, 2-best ratio,
    RANSAC support distance, rescue chi-square, dead-bands) logs its margin.  The GPU must reproduce every decision -- the
    sets those decisions produce are compared exactly, frame by frame -- and the smallest margin of the sequence must be far
    above the GPU-vs-CPU arithmetic difference (~1e-12 in a predicted pixel), i.e. no decision was decided by rounding.
    (Frame 1 is left out of the margin statistics: the map was initialised ON the integer keypoint positions of frame 0 and the
    camera has not moved, so its predicted pixels are integers +- 1 ulp and keypoints that moved by exactly one pixel sit at
    |distance - 1.0| ~ 1e-14 from the RANSAC threshold -- a degenerate tie of the synthetic scenario, which the cold-frame
    parity tests show both sides resolve identically.)"""
    from oracle import oracle_lib
    sc, orc, gpu = make_pair(640, 480, 100, warm=1)
    oracle_lib.margins_reset()
    worst = 0.0
    for t in range(2, 32):
        kp, ds = sc.frame(t)
        orc.step(kp, ds)
        gpu.set_keypoints(0, kp, ds); gpu.step()
        r, mo = gpu.feature_results(0), orc.get_measure()
        assert np.array_equal(mo["vis"], r["vis"])                                            # field of view + in-frame
        v = mo["vis"].astype(bool)
        worst = max(worst, float(np.abs(r["h"][v] - mo["h"][v]).max()))
        ma, ro = orc.get_match(), orc.get_ransac()
        assert np.array_equal(ma["matched"], r["matched"])                                    # foci gate + ratio test
        mm = ma["matched"].astype(bool)
        assert np.array_equal(ma["kp"][mm], r["kp"][mm])
        assert np.array_equal(ro["inlier"], r["inlier"]) and np.array_equal(ro["outlier"], r["outlier"])   # RANSAC distance
        assert np.array_equal(orc.get_rescue(), r["rescued"])                                 # chi-square gate
        compare_state(orc, gpu, f"frame {t}")                                                 # dead-bands (state)
    mg = oracle_lib.margins()
    print("decision margins over 30 frames (smallest non-zero |margin|, decisions, exactly on threshold):")
    for k, (m, c, z) in mg.items():
        print(f"  {k:22s} {m if m is None else format(m, '.3e')}  n={c}  on-threshold={z}")
    print(f"  largest GPU-vs-oracle difference of a predicted pixel: {worst:.2e}")
    for k in ("field_of_view_deg", "in_frame_px", "foci_gate_px", "ransac_distance_px", "rescue_chi2"):
        m, c, z = mg[k]
        assert c > 0 and z == 0 and m > 1e4 * max(worst, 1e-13), (k, mg[k], worst)
    # the ratio test compares integer Hamming distances (exact on both sides); dead-bands: a value exactly 0 stays 0
    assert mg["ratio_test_hamming"][1] > 0 and mg["dead_band"][1] > 0


def _run_batch(F, N, T, options):
    scs = [Scenario(640, 480, N, seed_offset=i) for i in range(F)]
    gpu = EkfBatch(scs[0].params, F, N, 4 * N + 64)
    for o, val in options:
        gpu.set_option(o, val)
    for i, sc in enumerate(scs):
        x, P, ft, fo, desc, _ = sc.init_map()
        gpu.set_state(i, x, P, ft, fo, desc)
    for t in range(1, T + 1):
        for i, sc in enumerate(scs):
            gpu.set_keypoints(i, *sc.frame(t))
        gpu.step()
    out = []
    for i in range(F):
        x, P = gpu.get_state(i)
        r = gpu.feature_results(i)
        out.append((x, P, r["matched"], r["inlier"], r["rescued"], gpu.frame_info(i)))
    gpu.close()
    return out


@pytest.mark.parametrize("lanes", [2, 3])
def test_lanes_give_bit_identical_filters(lanes):
    """EKFB_OPT_LANES: the filters of a handle run as lanes on their own streams, interleaved by ekfb_step -- the same kernels
    on the same per-filter data, so every filter must come out bit for bit as without lanes (9 filters: uneven split)."""
    ref = _run_batch(9, 40, 6, [(11, 1)])
    got = _run_batch(9, 40, 6, [(11, lanes)])
    for i, (a, b) in enumerate(zip(ref, got)):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), f"filter {i}: state / covariance differ"
        for j in (2, 3, 4):
            assert np.array_equal(a[j], b[j]), f"filter {i}: sets differ"
        assert a[5] == b[5]


@pytest.mark.parametrize("N", [12, 40, 64])
def test_small_update_kernel_against_the_block_step_path(N):
    """k_update_small (factorisation + slab TRSM + dx in one launch, every slab CTA factoring S itself; updates of at most 128
    rows) against the per-block-step launches it replaces, and against the oracle: N = 12 -> one block, N = 40 -> two blocks
    (k ~ 60..80), N = 64 -> the low-innovation update leaves the small path (k > 128) while the rescue update stays on it."""
    sc = Scenario(640, 480, N)
    x, P, ft, fo, desc, _ = sc.init_map()
    orc = OracleFilter(sc.params)
    orc.set_state(x, P, ft, fo, desc)
    a = EkfBatch(sc.params, 1, N, 4 * N + 64)
    b = EkfBatch(sc.params, 1, N, 4 * N + 64)
    b.set_option(10, 0)
    for g in (a, b):
        g.set_state(0, x, P, ft, fo, desc)
    for t in range(1, 9):
        kp, ds = sc.frame(t)
        orc.step(kp, ds)
        for g in (a, b):
            g.set_keypoints(0, kp, ds); g.step()
        (xa, Pa), (xb, Pb) = a.get_state(0), b.get_state(0)
        xo, Po = orc.get_state()
        assert rel_err(xa, xb) < 1e-12 and rel_err(Pa, Pb) < 1e-12, t
        assert rel_err(xa, xo) < TOL and rel_err(Pa, Po) < TOL, t
        assert np.array_equal(Pa, Pa.T)
        ra, rb = a.feature_results(0), b.feature_results(0)
        for key in ("matched", "inlier", "rescued"):
            assert np.array_equal(ra[key], rb[key])
        assert np.array_equal(ra["inlier"], orc.get_ransac()["inlier"])
        assert a.frame_info(0)["status"] == 0


def test_fused_update_with_a_single_buffered_trsm_ring():
    """704 < k <= 832 update rows: the slab of the fused chain + TRSM launch leaves room for one operand buffer only
    (k_update_fused<24, 1>).  Same frames through the per-block-step launches (option 3 = 0): equal to rounding, sets identical."""
    sc = Scenario(640, 480, 500, clutter_ratio=0.0, outlier_frac=0.0, noise_px=0.1, flip_p=0.0)
    x, P, ft, fo, desc, _ = sc.init_map()
    a = EkfBatch(sc.params, 1, 500, 1256)
    b = EkfBatch(sc.params, 1, 500, 1256)
    b.set_option(3, 0)
    for g in (a, b):
        g.set_state(0, x, P, ft, fo, desc)
    ks = []
    for t in range(1, 5):
        kp, ds = sc.frame(t)
        for g in (a, b):
            g.set_keypoints(0, kp, ds); g.step()
        ia, ib = a.frame_info(0), b.frame_info(0)
        assert ia == ib and ia["status"] == 0
        ks.append(2 * ia["n_inliers"])
        (xa, Pa), (xb, Pb) = a.get_state(0), b.get_state(0)
        assert rel_err(xa, xb) < 1e-12 and rel_err(Pa, Pb) < 1e-12, t
        assert np.array_equal(Pa, Pa.T)
        ra, rb = a.feature_results(0), b.feature_results(0)
        for key in ("matched", "inlier", "rescued"):
            assert np.array_equal(ra[key], rb[key])
    assert any(704 < k <= 832 for k in ks), ks


def test_fused_update_with_a_dense_slab_pitch():
    """832 < k <= 1024 update rows: the slab only fits beside the chain's tiles with a dense pitch (k_update_fused<24, 1, 0>).
    500 features on a grid (gates that do not overlap, so the reference's 2-best rule lets nearly all of them through), no
    clutter: k ~ 950.  Same frames through the per-block-step launches: equal to rounding, sets identical."""
    sc = Scenario(640, 480, 500, clutter_ratio=0.0, outlier_frac=0.0, noise_px=0.1, flip_p=0.0)
    p = sc.params
    gx, gy = np.meshgrid(np.linspace(0.16 * 640, 0.84 * 640, 25), np.linspace(0.10 * 480, 0.90 * 480, 20))
    d = np.random.default_rng(3).uniform(3.0, 6.0, 500)
    sc.points = np.stack([(gx.ravel() - p.cx) / p.fx * d, (gy.ravel() - p.cy) / p.fy * d, d], axis=1)
    x, P, ft, fo, desc, _ = sc.init_map()
    a = EkfBatch(sc.params, 1, 500, 1256)
    b = EkfBatch(sc.params, 1, 500, 1256)
    b.set_option(3, 0)
    for g in (a, b):
        g.set_state(0, x, P, ft, fo, desc)
    ks = []
    for t in range(1, 5):
        kp, ds = sc.frame(t)
        for g in (a, b):
            g.set_keypoints(0, kp, ds); g.step()
        ia, ib = a.frame_info(0), b.frame_info(0)
        assert ia == ib and ia["status"] == 0
        ks.append(2 * ia["n_inliers"])
        (xa, Pa), (xb, Pb) = a.get_state(0), b.get_state(0)
        assert rel_err(xa, xb) < 1e-12 and rel_err(Pa, Pb) < 1e-12, t
        assert np.array_equal(Pa, Pa.T)
    assert any(k > 832 for k in ks), ks
