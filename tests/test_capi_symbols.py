"""The C-ABI library loads and exports every symbol include/ekf_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from openekfmonoslam_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ekf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ekfb_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_all_declared_symbols():
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ekf_b200.h but not exported"


def test_params_struct_layout_matches_header():
    from openekfmonoslam_b200.params import EkfParams
    assert ctypes.sizeof(EkfParams) == 8 + 22 * 8
    assert capi.RECORD_BYTES == (13 + 169) * 8 + 12 * 4


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from openekfmonoslam_b200.params import synthetic_params
    with pytest.raises(capi.EkfError):
        capi.EkfBatch(synthetic_params(320, 240), 1, 8, 64)


def test_reference_config_parses():
    ref = "/root/reference/experiments/s3/config.yml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted")
    from openekfmonoslam_b200.params import load_config
    p, extras = load_config(ref)
    assert (p.pixels_x, p.pixels_y) == (640, 480) and abs(p.fx - 525.060143149240389) < 1e-12
    assert p.ransac_chi2 == 5.9915 and extras["max_map_size"] == 240 and extras["feature_detector"] == "STAR"
