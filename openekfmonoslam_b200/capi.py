"""ctypes binding of libekf_b200.so (include/ekf_b200.h).  No torch types cross this boundary.

``EkfBatch`` is the host-side object a caller drives: a batch of independent filters on one GPU
(n_filters = 1 is the single-filter drop-in for the reference's EKF object).  It raises
``EkfError`` -- never falls back to a CPU path -- when the CUDA library or a usable device is missing.
"""
import ctypes
import os

import numpy as np

from . import build as _build
from .params import EkfParams, MapPolicy, MapResult


class EkfError(RuntimeError):
    pass


class FrameInfo(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n", "n_features", "n_keypoints", "n_predicted", "n_matches",
                                              "n_hypotheses", "best_hypothesis", "n_inliers", "n_outliers",
                                              "n_rescued", "status", "reserved")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Record(ctypes.Structure):
    _fields_ = [("x_cam", ctypes.c_double * 13), ("P_cam", ctypes.c_double * 169), ("info", FrameInfo)]


RECORD_BYTES = ctypes.sizeof(Record)
PROFILE_GROUPS = ("predict", "measure", "match", "ransac", "gain", "chol", "downdate", "rescue", "misc")

_lib = None


def lib_path():
    return _build.OUT


def load(build_if_missing=True):
    """Load libekf_b200.so; build it in-tree with nvcc if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.OUT
    if build_if_missing:
        try:
            path = _build.build()
        except Exception as exc:  # nvcc missing on a box that already has the .so
            if not os.path.exists(path):
                raise EkfError(f"libekf_b200.so is missing and could not be built: {exc}") from exc
    if not os.path.exists(path):
        raise EkfError("libekf_b200.so is missing; run python -m openekfmonoslam_b200.build")
    L = ctypes.CDLL(path)
    L.ekfb_last_error.restype = ctypes.c_char_p
    L.ekfb_kernel_launches.restype = ctypes.c_int64
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class EkfBatch:
    def __init__(self, params: EkfParams, n_filters=1, max_features=64, max_keypoints=4096, device=0):
        self.L = load()
        self.params = params
        self.n_filters, self.max_features, self.max_keypoints = n_filters, max_features, max_keypoints
        self.h = ctypes.c_void_p()
        self._ck(self.L.ekfb_create(ctypes.byref(params), ctypes.c_int(device), ctypes.c_int(n_filters),
                                    ctypes.c_int(max_features), ctypes.c_int(max_keypoints), ctypes.byref(self.h)))
        # developer switches for every handle of the process: EKFB_OPTS="3=4,8=0" -> ekfb_set_option(3, 4), (8, 0)
        for kv in filter(None, os.environ.get("EKFB_OPTS", "").split(",")):
            self.set_option(*(int(x) for x in kv.split("=")))

    def _ck(self, rc):
        if rc != 0:
            msg = self.L.ekfb_last_error()
            raise EkfError(f"libekf_b200 error {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.ekfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----
    def set_state(self, f, x, P, ftype, foff, desc):
        x = np.ascontiguousarray(x, np.float64); P = np.ascontiguousarray(P, np.float64)
        ftype = np.ascontiguousarray(ftype, np.int32); foff = np.ascontiguousarray(foff, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8)
        self._ck(self.L.ekfb_set_state(self.h, ctypes.c_int(f), ctypes.c_int(x.shape[0]), ctypes.c_int(ftype.shape[0]),
                                       _ptr(x), _ptr(ftype), _ptr(foff), _ptr(P), _ptr(desc)))

    def dims(self, f=0):
        n, N = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self.L.ekfb_get_dims(self.h, ctypes.c_int(f), ctypes.byref(n), ctypes.byref(N)))
        return n.value, N.value

    def get_state(self, f=0, cam_only=False):
        n, _ = self.dims(f)
        if cam_only:
            n = 13
        x = np.zeros(n); P = np.zeros((n, n))
        self._ck(self.L.ekfb_get_state(self.h, ctypes.c_int(f), _ptr(x), _ptr(P), ctypes.c_int(1 if cam_only else 0)))
        return x, P

    def get_descriptors(self, f=0):
        _, N = self.dims(f)
        d = np.zeros((N, 32), np.uint8); tp = np.zeros(N, np.int32); tm = np.zeros(N, np.int32)
        self._ck(self.L.ekfb_get_descriptors(self.h, ctypes.c_int(f), _ptr(d), _ptr(tp), _ptr(tm)))
        return d, tp, tm

    # ---- keypoints ----
    def set_keypoints(self, f, xy, desc):
        """xy (Kp,2) float32, desc (Kp,32) uint8 host arrays (numpy, or pinned torch tensors via .numpy())."""
        xy = np.ascontiguousarray(xy, np.float32); desc = np.ascontiguousarray(desc, np.uint8)
        self._ck(self.L.ekfb_set_keypoints(self.h, ctypes.c_int(f), _ptr(xy), _ptr(desc), ctypes.c_int(xy.shape[0])))

    def set_keypoints_raw(self, f, xy_ptr, desc_ptr, n_kp):
        self._ck(self.L.ekfb_set_keypoints(self.h, ctypes.c_int(f), ctypes.c_void_p(xy_ptr), ctypes.c_void_p(desc_ptr),
                                           ctypes.c_int(n_kp)))

    def set_keypoints_batch_raw(self, xy_ptrs, desc_ptrs, counts):
        """all filters in one call: lists of host pointers (ints) and counts"""
        F = self.n_filters
        a = (ctypes.c_void_p * F)(*xy_ptrs); b = (ctypes.c_void_p * F)(*desc_ptrs); n = (ctypes.c_int32 * F)(*counts)
        self._ck(self.L.ekfb_set_keypoints_batch(self.h, a, b, n))

    def set_keypoints_packed_raw(self, xy_ptr, desc_ptr, offsets):
        """all filters from one packed (pinned) host buffer pair; offsets: int32 numpy array of n_filters + 1 entries"""
        offsets = np.ascontiguousarray(offsets, np.int32)
        self._ck(self.L.ekfb_set_keypoints_packed(self.h, ctypes.c_void_p(xy_ptr), ctypes.c_void_p(desc_ptr), _ptr(offsets)))

    def load_sequence(self, f, frames):
        """frames: list of (xy, desc) per frame; uploads them all to device memory."""
        off = np.zeros(len(frames) + 1, np.int32)
        for t, (xy, _) in enumerate(frames):
            off[t + 1] = off[t] + xy.shape[0]
        xy = np.ascontiguousarray(np.concatenate([fr[0] for fr in frames], axis=0), np.float32)
        ds = np.ascontiguousarray(np.concatenate([fr[1] for fr in frames], axis=0), np.uint8)
        self._ck(self.L.ekfb_load_sequence(self.h, ctypes.c_int(f), ctypes.c_int(len(frames)), _ptr(off), _ptr(xy), _ptr(ds)))

    def select_frame(self, t):
        self._ck(self.L.ekfb_select_frame(self.h, ctypes.c_int(t)))

    # ---- phases ----
    def predict(self): self._ck(self.L.ekfb_predict(self.h))
    def measure(self): self._ck(self.L.ekfb_measure(self.h))
    def match(self): self._ck(self.L.ekfb_match(self.h))
    def ransac(self): self._ck(self.L.ekfb_ransac(self.h))
    def update(self, which): self._ck(self.L.ekfb_update(self.h, ctypes.c_int(which)))
    def rescue(self): self._ck(self.L.ekfb_rescue(self.h))
    def update_map_features(self): self._ck(self.L.ekfb_update_map_features(self.h))
    def step(self): self._ck(self.L.ekfb_step(self.h))
    def sync(self): self._ck(self.L.ekfb_sync(self.h))

    # ---- results ----
    def frame_info(self, f=0):
        info = FrameInfo()
        self._ck(self.L.ekfb_get_frame_info(self.h, ctypes.c_int(f), ctypes.byref(info)))
        return info.as_dict()

    def records(self):
        arr = (Record * self.n_filters)()
        self._ck(self.L.ekfb_get_records(self.h, arr))
        return arr

    def write_records_device(self, device_ptr):
        self._ck(self.L.ekfb_write_records_device(self.h, ctypes.c_void_p(device_ptr)))

    def feature_results(self, f=0):
        _, N = self.dims(f)
        r = dict(vis=np.zeros(N, np.uint8), h=np.zeros((N, 2)), S=np.zeros((N, 4)), Hx=np.zeros((N, 14)),
                 Hf=np.zeros((N, 12)), matched=np.zeros(N, np.uint8), z=np.zeros((N, 2)), kp=np.zeros(N, np.int32),
                 dist=np.zeros(N, np.float32), inlier=np.zeros(N, np.uint8), outlier=np.zeros(N, np.uint8),
                 rescued=np.zeros(N, np.uint8))
        self._ck(self.L.ekfb_get_feature_results(self.h, ctypes.c_int(f), _ptr(r["vis"]), _ptr(r["h"]), _ptr(r["S"]),
                                                 _ptr(r["Hx"]), _ptr(r["Hf"]), _ptr(r["matched"]), _ptr(r["z"]),
                                                 _ptr(r["kp"]), _ptr(r["dist"]), _ptr(r["inlier"]), _ptr(r["outlier"]),
                                                 _ptr(r["rescued"])))
        return r

    def get_mask(self, f=0, n_kp=0):
        p = self.params
        mask = np.zeros((p.pixels_y, p.pixels_x), np.uint8)
        ok = np.zeros(max(n_kp, 1), np.uint8)
        self._ck(self.L.ekfb_get_mask(self.h, ctypes.c_int(f), _ptr(mask), _ptr(ok)))
        return mask, ok[:n_kp]

    # ---- map management on the device ----
    def map_management(self, policy: MapPolicy):
        """E/EKF.cpp:575-592 up to the detection of new features; returns one dict per filter"""
        out = (MapResult * self.n_filters)()
        self._ck(self.L.ekfb_map_management(self.h, ctypes.byref(policy), out))
        return [o.as_dict() for o in out]

    def removed_flags(self, f, n_features_before):
        fl = np.zeros(max(n_features_before, 1), np.uint8)
        self._ck(self.L.ekfb_get_removed_flags(self.h, ctypes.c_int(f), ctypes.c_int(n_features_before), _ptr(fl)))
        return fl[:n_features_before]

    def add_features(self, f, uv, desc):
        uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2); desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self._ck(self.L.ekfb_add_features(self.h, ctypes.c_int(f), ctypes.c_int(uv.shape[0]), _ptr(uv), _ptr(desc)))

    def feature_layout(self, f=0):
        _, N = self.dims(f)
        t = np.zeros(max(N, 1), np.int32); o = np.zeros(max(N, 1), np.int32)
        self._ck(self.L.ekfb_get_feature_layout(self.h, ctypes.c_int(f), _ptr(t), _ptr(o)))
        return t[:N], o[:N]

    def set_hit_counters(self, f, tp, tm):
        tp = np.ascontiguousarray(tp, np.int32); tm = np.ascontiguousarray(tm, np.int32)
        self._ck(self.L.ekfb_set_hit_counters(self.h, ctypes.c_int(f), _ptr(tp), _ptr(tm)))

    def new_feature_mask(self, f=0):
        m = np.zeros((self.params.pixels_y, self.params.pixels_x), np.uint8)
        self._ck(self.L.ekfb_get_new_feature_mask(self.h, ctypes.c_int(f), _ptr(m)))
        return m

    def raster_ellipse(self, img, cx, cy, S, max_axes, value=255):
        assert img.dtype == np.uint8 and img.flags.c_contiguous
        S = np.ascontiguousarray(S, np.float64)
        self._ck(self.L.ekfb_raster_ellipse(self.h, ctypes.c_int(img.shape[1]), ctypes.c_int(img.shape[0]), ctypes.c_double(cx),
                                            ctypes.c_double(cy), _ptr(S), ctypes.c_int(max_axes), ctypes.c_int(value), _ptr(img)))
        return img

    # ---- detector + descriptor on the device ----
    def set_image(self, f, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        self._ck(self.L.ekfb_set_image(self.h, ctypes.c_int(f), _ptr(gray), ctypes.c_int(gray.shape[1])))

    def set_image_color(self, f, img):
        """(H, W, 3) BGR or (H, W, 4) BGRA uint8"""
        img = np.ascontiguousarray(img, np.uint8)
        self._ck(self.L.ekfb_set_image_color(self.h, ctypes.c_int(f), _ptr(img), ctypes.c_int(img.shape[1] * img.shape[2]),
                                             ctypes.c_int(img.shape[2])))

    def detect_keypoints(self, f, threshold):
        n = ctypes.c_int32()
        self._ck(self.L.ekfb_detect_keypoints(self.h, ctypes.c_int(f), ctypes.c_int(threshold), ctypes.byref(n)))
        return n.value

    def get_keypoints(self, f, n):
        xy = np.zeros((max(n, 1), 2), np.float32); ds = np.zeros((max(n, 1), 32), np.uint8)
        self._ck(self.L.ekfb_get_keypoints(self.h, ctypes.c_int(f), _ptr(xy), _ptr(ds)))
        return xy[:n], ds[:n]

    # ---- NCC active search (north-star matching path) ----
    def ncc_set_image(self, f, gray):
        gray = np.ascontiguousarray(gray, np.uint8)
        self._ck(self.L.ekfb_ncc_set_image(self.h, ctypes.c_int(f), _ptr(gray), ctypes.c_int(gray.shape[1])))

    def ncc_set_templates(self, f, first, templates):
        t = np.ascontiguousarray(templates, np.uint8).reshape(-1, 3, 121)
        self._ck(self.L.ekfb_ncc_set_templates(self.h, ctypes.c_int(f), ctypes.c_int(first), ctypes.c_int(t.shape[0]), _ptr(t)))

    def ncc_get_templates(self, f, first, count):
        """(templates [count, 3, 121] uint8, anchors [count, 10] float64) of features first .. first + count - 1"""
        t = np.zeros((count, 3, 121), np.uint8); a = np.zeros((count, 10))
        self._ck(self.L.ekfb_ncc_get_templates(self.h, ctypes.c_int(f), ctypes.c_int(first), ctypes.c_int(count), _ptr(t), _ptr(a)))
        return t, a

    def ncc_set_anchors(self, f, first, anchors):
        a = np.ascontiguousarray(anchors, np.float64).reshape(-1, 10)
        self._ck(self.L.ekfb_ncc_set_anchors(self.h, ctypes.c_int(f), ctypes.c_int(first), ctypes.c_int(a.shape[0]), _ptr(a)))

    def ncc_set_threshold(self, ncc_min): self._ck(self.L.ekfb_ncc_set_threshold(self.h, ctypes.c_double(ncc_min)))

    def match_ncc(self, ncc_min=0.8):
        self._ck(self.L.ekfb_match_ncc(self.h, ctypes.c_double(ncc_min)))

    def ncc_scores(self, f=0):
        _, N = self.dims(f)
        s = np.zeros(max(N, 1)); lv = np.zeros(max(N, 1), np.int32)
        self._ck(self.L.ekfb_ncc_get_scores(self.h, ctypes.c_int(f), _ptr(s), _ptr(lv)))
        return s[:N], lv[:N]

    def ncc_level(self, f, level):
        w, h = ctypes.c_int32(), ctypes.c_int32()
        self._ck(self.L.ekfb_ncc_get_level(self.h, ctypes.c_int(f), ctypes.c_int(level), None, ctypes.byref(w), ctypes.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self._ck(self.L.ekfb_ncc_get_level(self.h, ctypes.c_int(f), ctypes.c_int(level), _ptr(out), None, None))
        return out

    # ---- isolated kernels / timing ----
    def test_downdate(self, P, Wt):
        P = np.ascontiguousarray(P, np.float64); Wt = np.ascontiguousarray(Wt, np.float64)
        out = np.zeros_like(P)
        self._ck(self.L.ekfb_test_downdate(self.h, ctypes.c_int(P.shape[0]), ctypes.c_int(Wt.shape[0]), _ptr(P), _ptr(Wt),
                                           _ptr(out)))
        return out

    def test_factor(self, S_aug):
        """[S | nu] (k x (k+1)) -> (U with y in column k, inverses of the diagonal 64x64 blocks)"""
        S_aug = np.ascontiguousarray(S_aug, np.float64)
        k = S_aug.shape[0]
        U = np.zeros_like(S_aug); Ui = np.zeros(((k + 63) // 64, 64, 64))
        self._ck(self.L.ekfb_test_factor(self.h, ctypes.c_int(k), _ptr(S_aug), _ptr(U), _ptr(Ui)))
        return U, Ui

    def time_update(self, which, reps):
        a, b = ctypes.c_float(), ctypes.c_float()
        self._ck(self.L.ekfb_time_update(self.h, ctypes.c_int(which), ctypes.c_int(reps), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def timer_record(self, slot): self._ck(self.L.ekfb_timer_record(self.h, ctypes.c_int(slot)))

    def timer_elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        self._ck(self.L.ekfb_timer_elapsed_ms(self.h, ctypes.c_int(a), ctypes.c_int(b), ctypes.byref(ms)))
        return ms.value

    def profile_enable(self, on=True): self._ck(self.L.ekfb_profile_enable(self.h, ctypes.c_int(1 if on else 0)))

    def profile_read(self):
        ms = (ctypes.c_float * 9)(); ln = (ctypes.c_int32 * 9)()
        self._ck(self.L.ekfb_profile_read(self.h, ms, ln))
        return dict(zip(PROFILE_GROUPS, list(ms))), dict(zip(PROFILE_GROUPS, list(ln)))

    def kernel_launches(self):
        return int(self.L.ekfb_kernel_launches(self.h))

    def set_option(self, option, value): self._ck(self.L.ekfb_set_option(self.h, ctypes.c_int(option), ctypes.c_int(value)))

    def flush_l2(self): self._ck(self.L.ekfb_flush_l2(self.h))

    def downdate_timing(self, on=True): self._ck(self.L.ekfb_downdate_timing(self.h, ctypes.c_int(1 if on else 0)))

    def downdate_launches(self, cap=16384):
        """(ms, flops) numpy arrays, one entry per timed downdate launch; call before downdate_stats()"""
        ms = np.zeros(cap, np.float32); fl = np.zeros(cap, np.float64); n = ctypes.c_int32()
        self._ck(self.L.ekfb_downdate_launches(self.h, ctypes.c_int(cap), _ptr(ms), _ptr(fl), ctypes.byref(n)))
        k = min(n.value, cap)
        return ms[:k], fl[:k]

    def downdate_stats(self):
        ms, fl, by = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        ln = ctypes.c_int64()
        self._ck(self.L.ekfb_downdate_stats(self.h, ctypes.byref(ms), ctypes.byref(ln), ctypes.byref(fl), ctypes.byref(by)))
        return dict(ms=ms.value, launches=ln.value, flops=fl.value, bytes_min=by.value)
