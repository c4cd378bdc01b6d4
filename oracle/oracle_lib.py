"""ctypes binding of the CPU oracle (oracle/ekf_oracle.cpp).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libekf_oracle.so")


class FrameInfo(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n", "n_features", "n_predicted", "n_matches", "n_hypotheses",
                                              "n_inliers", "n_outliers", "n_rescued")] + \
               [(k, ctypes.c_double) for k in ("us_prediction", "us_matching", "us_ransac", "us_update_li",
                                               "us_rescue", "us_update_hi", "us_map")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    src = os.path.join(_HERE, "ekf_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_create.restype = ctypes.c_void_p
        _lib = L
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleFilter:
    """One reference-style EKF instance on the CPU."""

    def __init__(self, params):
        self.L = lib()
        self.params = params
        self.h = ctypes.c_void_p(self.L.orc_create(ctypes.byref(params)))
        self.n_kp = 0

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # ---- state ----
    def dims(self):
        n, nf = ctypes.c_int32(), ctypes.c_int32()
        self.L.orc_dims(self.h, ctypes.byref(n), ctypes.byref(nf))
        return n.value, nf.value

    def init(self):
        self.L.orc_init(self.h)

    def add_feature(self, uv, desc):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        desc = np.ascontiguousarray(desc, dtype=np.uint8)
        self.L.orc_add_feature(self.h, _p(uv), _p(desc))

    def set_state(self, x, P, ftype, foff, desc):
        x = np.ascontiguousarray(x, dtype=np.float64)
        P = np.ascontiguousarray(P, dtype=np.float64)
        ftype = np.ascontiguousarray(ftype, dtype=np.int32)
        foff = np.ascontiguousarray(foff, dtype=np.int32)
        desc = np.ascontiguousarray(desc, dtype=np.uint8)
        self.L.orc_set_state(self.h, ctypes.c_int32(x.shape[0]), ctypes.c_int32(ftype.shape[0]), _p(x), _p(ftype),
                             _p(foff), _p(P), _p(desc))

    def get_state(self):
        n, _ = self.dims()
        x = np.zeros(n)
        P = np.zeros((n, n))
        self.L.orc_get_state(self.h, _p(x), _p(P))
        return x, P

    def get_features(self):
        _, N = self.dims()
        t = np.zeros(N, np.int32); o = np.zeros(N, np.int32); d = np.zeros((N, 32), np.uint8)
        tp = np.zeros(N, np.int32); tm = np.zeros(N, np.int32)
        self.L.orc_get_features(self.h, _p(t), _p(o), _p(d), _p(tp), _p(tm))
        return dict(type=t, off=o, desc=d, times_predicted=tp, times_matched=tm)

    # ---- map management (E/EKF.cpp:575-592, E/MapManagement.cpp) ----
    def set_hit_counters(self, tp, tm):
        tp = np.ascontiguousarray(tp, np.int32); tm = np.ascontiguousarray(tm, np.int32)
        self.L.orc_set_hit_counters(self.h, _p(tp), _p(tm))

    def map_management(self, policy):
        """policy: any ctypes struct with the layout of orc_map_policy.  Returns (needed, removed flags, converted)"""
        _, N = self.dims()
        rem = np.zeros(max(N, 1), np.uint8)
        conv = ctypes.c_int32(-1)
        self.L.orc_map_management.restype = ctypes.c_int32
        needed = self.L.orc_map_management(self.h, ctypes.byref(policy), _p(rem), ctypes.byref(conv))
        return needed, rem[:N], conv.value

    def remove_bad_features(self, good_pct):
        return self.L.orc_remove_bad_features(self.h, ctypes.c_double(good_pct), None)

    def convert_features(self, threshold):
        self.L.orc_convert_features.restype = ctypes.c_int32
        return self.L.orc_convert_features(self.h, ctypes.c_double(threshold))

    # ---- phases ----
    def predict(self):
        self.L.orc_predict(self.h)

    def measure(self):
        self.L.orc_measure(self.h)

    def match(self, kp_xy, kp_desc):
        kp_xy = np.ascontiguousarray(kp_xy, dtype=np.float32)
        kp_desc = np.ascontiguousarray(kp_desc, dtype=np.uint8)
        self.n_kp = kp_xy.shape[0]
        self.L.orc_match(self.h, _p(kp_xy), _p(kp_desc), ctypes.c_int32(self.n_kp))

    def ransac(self):
        self.L.orc_ransac(self.h)

    def update_li(self):
        self.L.orc_update_li(self.h)

    def rescue(self):
        self.L.orc_rescue(self.h)

    def update_hi(self):
        self.L.orc_update_hi(self.h)

    def update_map_features(self):
        self.L.orc_update_map_features(self.h)

    def step(self, kp_xy, kp_desc):
        kp_xy = np.ascontiguousarray(kp_xy, dtype=np.float32)
        kp_desc = np.ascontiguousarray(kp_desc, dtype=np.uint8)
        self.n_kp = kp_xy.shape[0]
        info = FrameInfo()
        self.L.orc_step(self.h, _p(kp_xy), _p(kp_desc), ctypes.c_int32(self.n_kp), ctypes.byref(info))
        return info.as_dict()

    # ---- per-frame results, indexed by feature ----
    def get_measure(self):
        _, N = self.dims()
        vis = np.zeros(N, np.uint8); h = np.zeros((N, 2)); S = np.zeros((N, 4)); Hx = np.zeros((N, 14))
        Hf = np.zeros((N, 12)); ell = np.zeros((N, 3))
        self.L.orc_get_measure(self.h, _p(vis), _p(h), _p(S), _p(Hx), _p(Hf), _p(ell))
        return dict(vis=vis, h=h, S=S, Hx=Hx, Hf=Hf, ell=ell)

    def get_match(self):
        _, N = self.dims()
        m = np.zeros(N, np.uint8); z = np.zeros((N, 2)); kp = np.zeros(N, np.int32); d = np.zeros(N, np.float32)
        self.L.orc_get_match(self.h, _p(m), _p(z), _p(kp), _p(d))
        return dict(matched=m, z=z, kp=kp, dist=d)

    def get_ransac(self):
        _, N = self.dims()
        inl = np.zeros(N, np.uint8); out = np.zeros(N, np.uint8)
        nh, best = ctypes.c_int32(), ctypes.c_int32()
        counts = np.zeros(max(N, 1), np.int32)
        self.L.orc_get_ransac(self.h, _p(inl), _p(out), ctypes.byref(nh), ctypes.byref(best), _p(counts))
        return dict(inlier=inl, outlier=out, n_hyp=nh.value, best=best.value, counts=counts[:nh.value])

    def get_rescue(self):
        _, N = self.dims()
        r = np.zeros(N, np.uint8)
        self.L.orc_get_rescue(self.h, _p(r))
        return r

    def get_mask(self):
        p = self.params
        mask = np.zeros((p.pixels_y, p.pixels_x), np.uint8)
        ok = np.zeros(max(self.n_kp, 1), np.uint8)
        self.L.orc_get_mask(self.h, _p(mask), _p(ok))
        return mask, ok[:self.n_kp]


# ---- decision margins ----
MARGIN_CLASSES = ("field_of_view_deg", "in_frame_px", "foci_gate_px", "ratio_test_hamming", "ransac_distance_px",
                  "rescue_chi2", "dead_band")


def margins_reset():
    lib().orc_margins_reset()


def margins():
    """{class: (smallest non-zero |margin| since the reset, number of decisions, decisions exactly on the threshold)}"""
    m = np.zeros(7); c = np.zeros(7, np.int64); z = np.zeros(7, np.int64)
    lib().orc_margins_get(_p(m), _p(c), _p(z))
    return {k: (float(m[i]) if c[i] > z[i] else None, int(c[i]), int(z[i])) for i, k in enumerate(MARGIN_CLASSES)}


# ---- OpenCV-primitive restatements (pinned against cv2 fixtures) ----
def eigen2x2(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    w = np.zeros(2); V = np.zeros((2, 2))
    lib().orc_eigen2x2(_p(A), _p(w), _p(V))
    return w, V


def invert(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.zeros_like(A)
    ok = lib().orc_invert(_p(A), ctypes.c_int32(A.shape[0]), _p(out))
    return out, bool(ok)


def fill_ellipse(img, cx, cy, aw, ah, angle_deg):
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    lib().orc_fill_ellipse(_p(img), ctypes.c_int32(img.shape[1]), ctypes.c_int32(img.shape[0]), ctypes.c_int32(cx),
                           ctypes.c_int32(cy), ctypes.c_int32(aw), ctypes.c_int32(ah), ctypes.c_double(angle_deg))
    return img


def ellipse_params(S):
    S = np.ascontiguousarray(S, dtype=np.float64)
    o = np.zeros(3)
    lib().orc_ellipse_params(_p(S), _p(o))
    return o


def draw_uncertainty_ellipse(img, cx, cy, S, max_axes):
    S = np.ascontiguousarray(S, dtype=np.float64)
    lib().orc_draw_uncertainty_ellipse(_p(img), ctypes.c_int32(img.shape[1]), ctypes.c_int32(img.shape[0]),
                                       ctypes.c_double(cx), ctypes.c_double(cy), _p(S), ctypes.c_int32(max_axes))
    return img


def point_in_ellipse(px, py, cx, cy, aw, ah, angle):
    return bool(lib().orc_point_in_ellipse(ctypes.c_float(px), ctypes.c_float(py), ctypes.c_float(cx),
                                           ctypes.c_float(cy), ctypes.c_int32(aw), ctypes.c_int32(ah),
                                           ctypes.c_double(angle)))


def update_dense(P, H, sigma):
    P = np.ascontiguousarray(P, dtype=np.float64).copy()
    H = np.ascontiguousarray(H, dtype=np.float64)
    n, k = P.shape[0], H.shape[0]
    K = np.zeros((n, k))
    lib().orc_update_dense(_p(P), None, ctypes.c_int32(n), _p(H), ctypes.c_int32(k), ctypes.c_double(sigma), _p(K))
    return P, K
