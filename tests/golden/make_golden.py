"""Generates tests/golden/opencv_primitives.npz with cv2 (4.13 in this container).

The reference's hot path calls three OpenCV primitives whose exact semantics decide pixels and
inlier sets: Mat::inv() (Update.cpp:108, EKF.cpp:94, MeasurementPrediction.cpp:216,353,672),
cv::eigen on a 2x2 (Core/EKFMath.cpp:277) and a filled cv::ellipse (Gui/Draw.cpp:58).  OpenCV
2.4.3.2 itself is not available; cv2 4.13 is the closest executable statement of those algorithms.
Run:  python tests/golden/make_golden.py      (needs cv2; the committed .npz does not)
"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20241017)
out = {}

# ---- cv::eigen 2x2 symmetric ----
mats = []
for _ in range(200):
    a = rng.normal(size=(2, 2)) * rng.uniform(0.1, 30)
    mats.append(a @ a.T + np.eye(2) * rng.uniform(0.0, 2.0))
mats += [np.array([[9., 2.], [2., 4.]]), np.array([[4., 2.], [2., 9.]]), np.array([[3., 0.], [0., 5.]]),
         np.array([[5., 0.], [0., 3.]]), np.array([[2., 1e-17], [1e-17, 2.]]), np.array([[1., -0.5], [-0.5, 1.]]),
         np.array([[7., -3.], [-3., 2.]]), np.array([[2., 2.], [2., 2.000001]])]
mats = np.array(mats)
ev, evec = [], []
for m in mats:
    ok, w, v = cv2.eigen(m)
    ev.append(w.reshape(2)); evec.append(v)
out["eig_in"], out["eig_w"], out["eig_v"] = mats, np.array(ev), np.array(evec)

# ---- Mat::inv() DECOMP_LU ----
for n in (2, 3, 6, 20):
    A = []
    for _ in range(20):
        a = rng.normal(size=(n, n))
        A.append(a @ a.T + np.eye(n) if n > 3 else a)
    A = np.array(A)
    out[f"inv_in_{n}"] = A
    out[f"inv_out_{n}"] = np.array([cv2.invert(a, flags=cv2.DECOMP_LU)[1] for a in A])

# ---- filled ellipse masks, 96x80 canvas, packed bits ----
W, H = 96, 80
cases, masks = [], []
for i in range(600):
    if i < 400:
        cx, cy = int(rng.integers(20, W - 20)), int(rng.integers(20, H - 20))
        aw, ah = int(rng.integers(0, 19)), int(rng.integers(0, 19))
    else:  # border-crossing and large ellipses
        cx, cy = int(rng.integers(-10, W + 10)), int(rng.integers(-10, H + 10))
        aw, ah = int(rng.integers(0, 40)), int(rng.integers(0, 40))
    ang = float(rng.uniform(-95, 95))
    img = np.zeros((H, W), np.uint8)
    cv2.ellipse(img, (cx, cy), (aw, ah), ang, 0, 360, 255, -1)
    cases.append((cx, cy, aw, ah, ang))
    masks.append(np.packbits(img > 0))
out["ell_cases"] = np.array(cases, dtype=np.float64)
out["ell_masks"] = np.array(masks)
out["ell_shape"] = np.array([H, W])
out["cv2_version"] = np.array(cv2.__version__)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencv_primitives.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
