"""NCC active search (csrc/ekf_ncc.cuh; BASELINE.json's north-star matching path -- the reference has none).
CPU: the numpy oracle finds planted patches.  GPU: pyramid bytes, matched flags, matched pixels, start levels and scores
must equal the oracle's exactly (integer sums, correctly rounded double operations)."""
import numpy as np
import pytest

from openekfmonoslam_b200.scenario import Scenario
from oracle import ncc_oracle
from oracle.oracle_lib import OracleFilter


def render(W, H, pos, tex, seed, jitter=0):
    """mid-grey + N(0, 2^2) noise with every feature's 11 x 11 texture pasted at its (integer) pixel"""
    rng = np.random.default_rng(seed)
    img = np.clip(128 + rng.normal(0, 2, (H, W)), 0, 255).astype(np.uint8)
    for (x, y), t in zip(pos, tex):
        x, y = int(x), int(y)
        if 6 <= x < W - 6 and 6 <= y < H - 6:
            img[y - 5:y + 6, x - 5:x + 6] = t.reshape(11, 11)
    return img


def smooth_textures(rng, N):
    """11 x 11 textures with most of their energy at low spatial frequency (a coarse random 3 x 3 pattern, bilinearly
    blown up, plus a little noise), so that they survive two 2x pyramid reductions the way image structure does"""
    out = np.zeros((N, 121), np.uint8)
    g = np.linspace(0, 2, 11)
    i0 = np.minimum(g.astype(int), 1); fr = g - i0
    for k in range(N):
        c = rng.uniform(20, 235, (3, 3))
        rows = c[i0] * (1 - fr)[:, None] + c[i0 + 1] * fr[:, None]              # 11 x 3
        t = rows[:, i0] * (1 - fr)[None, :] + rows[:, i0 + 1] * fr[None, :]      # 11 x 11
        out[k] = np.clip(t + rng.normal(0, 4, (11, 11)), 0, 255).astype(np.uint8).ravel()
    return out


def feature_pixels(sc, t):
    """observed pixel of every map feature in frame t (feature order), from the scenario's own keypoint list"""
    kp, _, owner, _ = sc.frame(t, with_truth=True)
    pos = np.full((sc.N, 2), -100.0)
    pos[owner[owner >= 0]] = kp[owner >= 0]
    return pos


def ncc_case(N=40, W=320, H=240, pscale=1.0):
    sc = Scenario(W, H, N)
    x, P, ft, fo, desc, uv0 = sc.init_map()
    P = P * pscale           # wider gates: pscale 4 starts the search at pyramid level 1, 9 at level 2
    rng = np.random.default_rng(99)
    tex = smooth_textures(rng, N)
    orc = OracleFilter(sc.params)
    orc.set_state(x, P, ft, fo, desc)
    tmpl = ncc_oracle.cut_templates(ncc_oracle.pyramid(render(W, H, uv0, tex, 1)), uv0)
    return sc, orc, (x, P, ft, fo, desc), tex, tmpl


@pytest.mark.parametrize("pscale,min_level", [(1.0, 0), (4.0, 1), (9.0, 2)])
def test_oracle_finds_planted_patches(pscale, min_level):
    sc, orc, _, tex, tmpl = ncc_case(pscale=pscale)
    pos = feature_pixels(sc, 1)
    orc.predict(); orc.measure()
    mo = orc.get_measure()
    img = render(320, 240, pos, tex, 2)
    matched, z, score, level = ncc_oracle.search(ncc_oracle.pyramid(img), tmpl, mo["vis"], mo["h"], mo["ell"][:, :2], mo["ell"][:, 2])
    good = matched.astype(bool)
    assert good.sum() >= 20 and level[mo["vis"].astype(bool)].max() >= min_level
    # a planted texture is found at its pixel unless the 10 % outlier displacement moved it outside the gate
    hit = (np.abs(z[good] - pos[good]).max(axis=1) <= 1)
    assert hit.mean() > 0.8 and (score[good] >= 0.8).all()
    assert set(level[mo["vis"].astype(bool)]) <= {0, 1, 2}


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,N,pscale", [(320, 240, 40, 1.0), (640, 480, 120, 1.0), (640, 480, 60, 4.0), (640, 480, 60, 9.0)])
def test_gpu_ncc_search_equals_oracle(W, H, N, pscale):
    from openekfmonoslam_b200.capi import EkfBatch
    sc, orc, (x, P, ft, fo, desc), tex, tmpl = ncc_case(N, W, H, pscale)
    gpu = EkfBatch(sc.params, 1, N, 4 * N + 64)
    gpu.set_state(0, x, P, ft, fo, desc)
    gpu.ncc_set_templates(0, 0, tmpl)
    levels_seen = set()
    for t in range(1, 4):
        img = render(W, H, feature_pixels(sc, t), tex, 10 + t)
        pyr = ncc_oracle.pyramid(img)
        gpu.ncc_set_image(0, img)
        for l in range(3):
            assert np.array_equal(gpu.ncc_level(0, l), pyr[l]), f"pyramid level {l}"
        orc.predict(); gpu.predict()
        orc.measure(); gpu.measure()
        mo = orc.get_measure()
        matched, z, score, level = ncc_oracle.search(pyr, tmpl, mo["vis"], mo["h"], mo["ell"][:, :2], mo["ell"][:, 2])
        gpu.match_ncc(0.8)
        r = gpu.feature_results(0)
        gs, gl = gpu.ncc_scores(0)
        assert np.array_equal(gl, level)
        assert np.array_equal(gs, score), f"frame {t}: scores differ, max {np.abs(gs - score).max()}"
        assert np.array_equal(r["matched"], matched)
        m = matched.astype(bool)
        assert np.array_equal(r["z"][m], z[m]) and m.sum() > N // 3
        levels_seen |= set(level[mo["vis"].astype(bool)])
        # carry on with the frame from the NCC matches on the GPU and from the same matches on a descriptor-free path:
        # the oracle side just repeats the prediction next frame (no update), so do the same on the GPU
    assert max(levels_seen) >= (0 if pscale == 1.0 else 1 if pscale == 4.0 else 2)
    # the rest of the frame consumes NCC matches like descriptor matches
    gpu.ransac(); gpu.update(0); gpu.rescue(); gpu.update(1); gpu.update_map_features()
    info = gpu.frame_info(0)
    assert info["status"] == 0 and info["n_inliers"] > N // 3
    Pg = gpu.get_state(0)[1]
    assert np.array_equal(Pg, Pg.T)


@pytest.mark.gpu
def test_gpu_pyramid_of_an_odd_sized_image():
    """251 x 187: the device rows are padded to a 16-byte pitch, the levels are floor(W/2) x floor(H/2)"""
    from openekfmonoslam_b200.capi import EkfBatch
    from openekfmonoslam_b200.params import synthetic_params
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (187, 251), dtype=np.uint8)
    gpu = EkfBatch(synthetic_params(251, 187), 1, 4, 64)
    gpu.ncc_set_image(0, img)
    for l, ref in enumerate(ncc_oracle.pyramid(img)):
        assert np.array_equal(gpu.ncc_level(0, l), ref), l


@pytest.mark.gpu
def test_gpu_ncc_search_two_filters_of_one_handle():
    """per-filter images, templates and results: filter 1 sees a different frame of the same scene than filter 0"""
    from openekfmonoslam_b200.capi import EkfBatch
    W, H, N = 320, 240, 40
    sc, orc0, (x, P, ft, fo, desc), tex, tmpl = ncc_case(N, W, H)
    orc1 = OracleFilter(sc.params); orc1.set_state(x, P, ft, fo, desc)
    gpu = EkfBatch(sc.params, 2, N, 4 * N + 64)
    for f in (0, 1):
        gpu.set_state(f, x, P, ft, fo, desc)
        gpu.ncc_set_templates(f, 0, tmpl)
    imgs = [render(W, H, feature_pixels(sc, 1), tex, 21), render(W, H, feature_pixels(sc, 2), tex, 22)]
    for f in (0, 1):
        gpu.ncc_set_image(f, imgs[f])
    gpu.predict(); gpu.measure(); gpu.match_ncc(0.8)
    for f, orc in ((0, orc0), (1, orc1)):
        orc.predict(); orc.measure()
        mo = orc.get_measure()
        matched, z, score, level = ncc_oracle.search(ncc_oracle.pyramid(imgs[f]), tmpl, mo["vis"], mo["h"], mo["ell"][:, :2], mo["ell"][:, 2])
        r = gpu.feature_results(f)
        gs, gl = gpu.ncc_scores(f)
        assert np.array_equal(gs, score) and np.array_equal(gl, level) and np.array_equal(r["matched"], matched)
        assert np.array_equal(r["z"][matched.astype(bool)], z[matched.astype(bool)])
    assert gpu.frame_info(0)["n_matches"] != gpu.frame_info(1)["n_matches"] or not np.array_equal(gpu.ncc_scores(0)[0], gpu.ncc_scores(1)[0])
