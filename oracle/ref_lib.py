"""ctypes binding of oracle/_ref/libref.so: the REFERENCE's own per-frame sources compiled against oracle/cvshim
(oracle/ref/Makefile).  TEST INFRASTRUCTURE: used to pin the oracle, and as the "reference" CPU arm when present."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref.so")


def available(build=True):
    if os.path.exists(_SO):
        return True
    if build and os.path.isdir("/root/reference/kalmanFilter/modules"):
        try:
            subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref")])
        except Exception:
            return False
    return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref.so is not built (needs /root/reference)")
        L = ctypes.CDLL(_SO)
        L.ref_create.restype = ctypes.c_void_p
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class ReferenceFilter:
    """The reference's EKF object (one per process: its configuration is a process-global singleton)."""

    def __init__(self, params):
        self.L = lib()
        self.params = params
        self.h = ctypes.c_void_p(self.L.ref_create(ctypes.byref(params)))

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def dims(self):
        n, N = ctypes.c_int32(), ctypes.c_int32()
        self.L.ref_dims(self.h, ctypes.byref(n), ctypes.byref(N))
        return n.value, N.value

    def init(self):
        self.L.ref_init(self.h)

    def add_feature(self, uv, desc):
        uv = np.ascontiguousarray(uv, np.float64); desc = np.ascontiguousarray(desc, np.uint8)
        self.L.ref_add_feature(self.h, _p(uv), _p(desc))

    def set_state(self, x, P, ftype, foff, desc):
        x = np.ascontiguousarray(x, np.float64); P = np.ascontiguousarray(P, np.float64)
        ftype = np.ascontiguousarray(ftype, np.int32); foff = np.ascontiguousarray(foff, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8)
        self.L.ref_set_state(self.h, ctypes.c_int32(x.shape[0]), ctypes.c_int32(ftype.shape[0]), _p(x), _p(ftype), _p(foff),
                             _p(P), _p(desc))

    def get_state(self):
        n, _ = self.dims()
        x = np.zeros(n); P = np.zeros((n, n))
        self.L.ref_get_state(self.h, _p(x), _p(P))
        return x, P

    def get_features(self):
        _, N = self.dims()
        d = np.zeros((N, 32), np.uint8); tp = np.zeros(N, np.int32); tm = np.zeros(N, np.int32)
        self.L.ref_get_features(self.h, _p(d), _p(tp), _p(tm))
        return dict(desc=d, times_predicted=tp, times_matched=tm)

    # ---- map management: the reference's own functions (E/MapManagement.cpp, E/DetectNewImageFeatures.cpp) ----
    def set_policy(self, policy, map_management_frequency=0):
        self.L.ref_set_policy(self.h, ctypes.byref(policy), ctypes.c_int32(map_management_frequency))

    def get_layout(self):
        _, N = self.dims()
        t = np.zeros(max(N, 1), np.int32); o = np.zeros(max(N, 1), np.int32)
        self.L.ref_get_layout(self.h, _p(t), _p(o))
        return t[:N], o[:N]

    def set_hit_counters(self, tp, tm):
        tp = np.ascontiguousarray(tp, np.int32); tm = np.ascontiguousarray(tm, np.int32)
        self.L.ref_set_counters(self.h, _p(tp), _p(tm))

    def map_management(self):
        self.L.ref_map_management.restype = ctypes.c_int32
        return self.L.ref_map_management(self.h)

    def remove_bad(self):
        return self.L.ref_remove_bad(self.h)

    def convert(self):
        self.L.ref_convert(self.h)

    def detect_new(self, kp_xy, kp_desc, max_new):
        uv = np.zeros((max(max_new, 1), 2)); ds = np.zeros((max(max_new, 1), 32), np.uint8)
        self.L.ref_detect_new.restype = ctypes.c_int32
        k = self.L.ref_detect_new(self.h, *self._kp(kp_xy, kp_desc), ctypes.c_int32(max_new), _p(uv), _p(ds))
        return uv[:k], ds[:k]

    def _kp(self, kp_xy, kp_desc):
        self._kpxy = np.ascontiguousarray(kp_xy, np.float32)
        self._kpds = np.ascontiguousarray(kp_desc, np.uint8)   # must outlive the frame (descriptors are read lazily)
        return _p(self._kpxy), _p(self._kpds), ctypes.c_int32(self._kpxy.shape[0])

    def full_init(self, kp_xy, kp_desc):
        """the reference's own EKF::init on the injected keypoints"""
        self.L.ref_full_init(self.h, *self._kp(kp_xy, kp_desc))

    def step(self, kp_xy, kp_desc):
        self.L.ref_step(self.h, *self._kp(kp_xy, kp_desc))

    def predict(self): self.L.ref_predict(self.h)
    def measure(self): self.L.ref_measure(self.h)
    def match(self, kp_xy, kp_desc): self.L.ref_match(self.h, *self._kp(kp_xy, kp_desc))
    def ransac(self): self.L.ref_ransac(self.h)
    def update_li(self): self.L.ref_update_li(self.h)
    def rescue(self): self.L.ref_rescue(self.h)
    def update_hi(self): self.L.ref_update_hi(self.h)
    def update_map_features(self): self.L.ref_update_map_features(self.h)

    def get_measure(self):
        _, N = self.dims()
        vis = np.zeros(N, np.uint8); h = np.zeros((N, 2)); S = np.zeros((N, 4)); Hx = np.zeros((N, 14)); Hf = np.zeros((N, 12))
        self.L.ref_get_measure(self.h, _p(vis), _p(h), _p(S), _p(Hx), _p(Hf))
        return dict(vis=vis, h=h, S=S, Hx=Hx, Hf=Hf)

    def get_match(self):
        _, N = self.dims()
        m = np.zeros(N, np.uint8); z = np.zeros((N, 2)); d = np.zeros(N, np.float32)
        self.L.ref_get_match(self.h, _p(m), _p(z), _p(d))
        return dict(matched=m, z=z, dist=d)

    def get_sets(self):
        _, N = self.dims()
        a = np.zeros(N, np.uint8); b = np.zeros(N, np.uint8); c = np.zeros(N, np.uint8)
        self.L.ref_get_sets(self.h, _p(a), _p(b), _p(c))
        return dict(inlier=a, outlier=b, rescued=c)
