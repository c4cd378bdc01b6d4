"""Deterministic synthetic monocular sequences (the workloads of BASELINE.json configs 2-5).

The per-frame hot path consumes what the reference's front end produces inside
matchPredictedFeatures (1PointRansacEKF/Matching.cpp:204-215): keypoint pixel positions (float32,
integer valued like STAR's) and one 32-byte binary descriptor per keypoint.  A ``Scenario``
generates exactly that, plus the initial map built the way the reference builds it
(EKF::init -> addFeaturesToStateAndCovariance, AddMapFeature.cpp:116-350).

Scene (with one documented change: the camera x-motion is a bounded sinusoid whose
peak speed equals the reference robot's 0.002904 m/frame (resultReader/main.cpp:42) instead of
an unbounded ramp, so that every point stays in view for arbitrarily long runs):
  * N points, uniform in the image of the first frame, depth U(2, 8) m;
  * camera r(t) = (A sin(2 pi t/400), 0.01 sin(2 pi t/200), 0), A = 0.002904*400/(2 pi),
    yaw(t) = 0.05 sin(2 pi t/300) rad about the camera y axis;
  * keypoint = distorted true projection + N(0, 0.3^2) px, rounded to integers; 10 % of the
    features per frame are displaced by U(5, 15) px (outliers); N clutter keypoints uniform;
  * descriptors: 256 random bits per feature, each observation flips bits with p = 0.05.
Seeds: scene 20130600 + N (+1000 * seed_offset), frame noise (1234 + frame, seed_offset).
"""
import math

import numpy as np

from .params import synthetic_params

EPSILON = 2.22e-16  # Core/EKFMath.h:37


def _quat_to_rot(q):
    r, x, y, z = q
    return np.array([
        [r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)],
        [2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)],
        [2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z]])


def _jac_quat_to_rot(q, a):
    """d(R(q) a)/dq, 3x4 (CommonFunctions.cpp:87-145)."""
    q0, qx, qy, qz = q
    mats = (
        np.array([[2 * q0, -2 * qz, 2 * qy], [2 * qz, 2 * q0, -2 * qx], [-2 * qy, 2 * qx, 2 * q0]]),
        np.array([[2 * qx, 2 * qy, 2 * qz], [2 * qy, -2 * qx, -2 * q0], [2 * qz, 2 * q0, -2 * qx]]),
        np.array([[-2 * qy, 2 * qx, 2 * q0], [2 * qx, 2 * qy, 2 * qz], [-2 * q0, 2 * qz, -2 * qy]]),
        np.array([[-2 * qz, -2 * q0, 2 * qx], [2 * q0, -2 * qz, 2 * qy], [2 * qx, 2 * qy, 2 * qz]]),
    )
    return np.stack([m @ a for m in mats], axis=1)


def distort(p, uv):
    """Radial distortion of ideal pixel coordinates, the reference's 10-step Newton solve
    (MeasurementPrediction.cpp:47-83).  uv: (..., 2) array."""
    uv = np.asarray(uv, dtype=np.float64)
    px, py = uv[..., 0] - p.cx, uv[..., 1] - p.cy
    ddx, ddy = p.dx * px, p.dy * py
    d2 = ddx * ddx + ddy * ddy
    ru = np.sqrt(d2)
    rd = ru / (1.0 + p.k1 * d2 + p.k2 * d2 * d2)
    for _ in range(10):
        rd2 = rd * rd
        f = rd + p.k1 * rd2 * rd + p.k2 * rd2 * rd2 * rd - ru
        fp = 1 + 3 * p.k1 * rd2 + 5 * p.k2 * rd2 * rd2
        rd = rd - f / fp
    rd2 = rd * rd
    d = 1.0 + p.k1 * rd2 + p.k2 * rd2 * rd2
    return np.stack([p.cx + px / d, p.cy + py / d], axis=-1)


def undistort(p, uv):
    """AddMapFeature.cpp:43-59."""
    uv = np.asarray(uv, dtype=np.float64)
    px, py = uv[..., 0] - p.cx, uv[..., 1] - p.cy
    ddx, ddy = p.dx * px, p.dy * py
    rd = ddx * ddx + ddy * ddy
    dist = 1 + p.k1 * rd + p.k2 * rd * rd
    return np.stack([p.cx + px * dist, p.cy + py * dist], axis=-1)


def init_state_and_covariance(p):
    """initState / initCovariance (CommonFunctions.cpp:39-80)."""
    x = np.zeros(13)
    x[3] = 1.0
    x[10:13] = EPSILON
    P = np.zeros((13, 13))
    P[np.arange(7), np.arange(7)] = EPSILON
    P[np.arange(7, 10), np.arange(7, 10)] = p.init_linear_accel_sd ** 2
    P[np.arange(10, 13), np.arange(10, 13)] = p.init_angular_accel_sd ** 2
    return x, P


def add_features_inverse_depth(p, x, P, uvs):
    """Append inverse-depth features observed at pixels ``uvs`` (M, 2) to (x, P) the way
    addFeaturesToStateAndCovariance does (AddMapFeature.cpp:221-350), all at once."""
    uvs = np.asarray(uvs, dtype=np.float64).reshape(-1, 2)
    M = uvs.shape[0]
    n0 = x.shape[0]
    q = x[3:7]
    R = _quat_to_rot(q)
    und = undistort(p, uvs)
    J_all = np.zeros((6 * M, 7))
    noise = np.zeros((6 * M, 6 * M))
    xf = np.zeros(6 * M)
    N3 = np.diag([p.pixel_error_x ** 2, p.pixel_error_y ** 2, p.inverse_depth_rho_sd ** 2])
    for i in range(M):
        gc = np.array([-(p.cx - und[i, 0]) / p.fx, -(p.cy - und[i, 1]) / p.fy, 1.0])
        gw = R @ gc
        xw, yw, zw = gw
        xf[6 * i:6 * i + 3] = x[0:3]
        xf[6 * i + 3] = math.atan2(xw, zw)
        xf[6 * i + 4] = math.atan2(-yw, math.sqrt(xw * xw + zw * zw))
        xf[6 * i + 5] = p.init_inv_depth_rho
        xxzz = xw * xw + zw * zw
        dth = np.array([zw / xxzz, 0.0, -xw / xxzz])
        sq = math.sqrt(xxzz)
        nsq = xxzz + yw * yw
        dph = np.array([xw * yw / (nsq * sq), -sq / nsq, zw * yw / (nsq * sq)])
        dgw_dq = _jac_quat_to_rot(q, gc)
        J = np.zeros((6, 7))
        J[0, 0] = J[1, 1] = J[2, 2] = 1.0
        J[3, 3:7] = dth @ dgw_dq
        J[4, 3:7] = dph @ dgw_dq
        sub = np.stack([dth @ R, dph @ R])                       # 2x3
        dgc_dhu = np.array([[1.0 / p.fx, 0.0], [0.0, 1.0 / p.fy], [0.0, 0.0]])
        ud, vd = uvs[i]
        xd, yd = (ud - p.cx) * p.dx, (vd - p.cy) * p.dy
        rd2 = xd * xd + yd * yd
        k12 = p.k1 + 2.0 * p.k2 * rd2
        k1p = 1.0 + p.k1 * rd2 + p.k2 * rd2 * rd2
        dx2, dy2 = 2.0 * p.dx * p.dx, 2.0 * p.dy * p.dy
        dhu = np.array([[k1p + (ud - p.cx) * k12 * ((ud - p.cx) * dx2), (ud - p.cx) * k12 * ((vd - p.cy) * dy2)],
                        [(vd - p.cy) * k12 * ((ud - p.cx) * dx2), (vd - p.cy) * k12 * ((vd - p.cy) * dy2) + k1p]])
        ab = sub @ dgc_dhu @ dhu                                  # 2x2
        JH = np.zeros((6, 3))
        JH[3, 0:2] = ab[0]
        JH[4, 0:2] = ab[1]
        JH[5, 2] = 1.0
        J_all[6 * i:6 * i + 6] = J
        noise[6 * i:6 * i + 6, 6 * i:6 * i + 6] = JH @ N3 @ JH.T
    n = n0 + 6 * M
    Pn = np.zeros((n, n))
    Pn[:n0, :n0] = P
    C = J_all @ P[0:7, :]                 # (6M, n0)
    Pn[n0:, :n0] = C
    Pn[:n0, n0:] = P[:, 0:7] @ J_all.T
    Pn[n0:, n0:] = C[:, 0:7] @ J_all.T + noise
    xn = np.concatenate([x, xf])
    return xn, Pn


class Scenario:
    def __init__(self, width, height, n_features, seed_offset=0, outlier_frac=0.10, clutter_ratio=1.0,
                 noise_px=0.3, flip_p=0.05):
        self.W, self.H, self.N = int(width), int(height), int(n_features)
        self.seed_offset = int(seed_offset)
        self.outlier_frac, self.clutter_ratio = outlier_frac, clutter_ratio
        self.noise_px, self.flip_p = noise_px, flip_p
        self.params = synthetic_params(width, height)
        p = self.params
        rng = np.random.default_rng(20130600 + self.N + 1000 * self.seed_offset)
        mx, my = 0.16 * self.W, 0.10 * self.H
        u = rng.uniform(mx, self.W - mx, self.N)
        v = rng.uniform(my, self.H - my, self.N)
        d = rng.uniform(2.0, 8.0, self.N)
        self.points = np.stack([(u - p.cx) / p.fx * d, (v - p.cy) / p.fy * d, d], axis=1)
        self.descriptors = rng.integers(0, 256, size=(self.N, 32), dtype=np.uint8)
        self.A = 0.002904 * 400.0 / (2.0 * math.pi)

    # -- ground truth ------------------------------------------------------------------------
    def camera_pose(self, t):
        r = np.array([self.A * math.sin(2 * math.pi * t / 400.0), 0.01 * math.sin(2 * math.pi * t / 200.0), 0.0])
        psi = 0.05 * math.sin(2 * math.pi * t / 300.0)
        c, s = math.cos(psi), math.sin(psi)
        R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
        return r, R

    def true_pixels(self, t):
        p = self.params
        r, R = self.camera_pose(t)
        c = (self.points - r) @ R              # rows: R^T (p - r)
        uv = np.stack([p.cx + p.fx * c[:, 0] / c[:, 2], p.cy + p.fy * c[:, 1] / c[:, 2]], axis=1)
        return distort(p, uv)

    # -- front-end output for frame t ---------------------------------------------------------
    def frame(self, t, with_truth=False):
        rng = np.random.default_rng([1234 + int(t), self.seed_offset])
        N = self.N
        px = self.true_pixels(t) + rng.normal(0.0, self.noise_px, size=(N, 2))
        n_out = int(round(self.outlier_frac * N))
        out_idx = rng.choice(N, size=n_out, replace=False) if n_out > 0 else np.zeros(0, dtype=np.int64)
        ang = rng.uniform(0.0, 2 * math.pi, n_out)
        mag = rng.uniform(5.0, 15.0, n_out)
        px[out_idx] += np.stack([mag * np.cos(ang), mag * np.sin(ang)], axis=1)
        px = np.rint(px)
        flips = rng.random(size=(N, 256)) < self.flip_p
        desc = self.descriptors ^ np.packbits(flips, axis=1)
        n_cl = int(round(self.clutter_ratio * N))
        cl = np.stack([rng.integers(0, self.W, n_cl), rng.integers(0, self.H, n_cl)], axis=1).astype(np.float64)
        cl_desc = rng.integers(0, 256, size=(n_cl, 32), dtype=np.uint8)
        xy = np.concatenate([px, cl], axis=0)
        ds = np.concatenate([desc, cl_desc], axis=0)
        owner = np.concatenate([np.arange(N), -np.ones(n_cl, dtype=np.int64)])
        ok = (xy[:, 0] >= 0) & (xy[:, 0] <= self.W - 1) & (xy[:, 1] >= 0) & (xy[:, 1] <= self.H - 1)
        xy, ds, owner = xy[ok], ds[ok], owner[ok]
        order = np.lexsort((xy[:, 0], xy[:, 1]))   # raster order (row-major), like a response-map scan
        xy, ds, owner = xy[order], ds[order], owner[order]
        kp = np.ascontiguousarray(xy.astype(np.float32))
        ds = np.ascontiguousarray(ds)
        if with_truth:
            is_out = np.zeros(N, dtype=bool)
            is_out[out_idx] = True
            return kp, ds, owner, is_out
        return kp, ds

    # -- initial map ----------------------------------------------------------------------------
    def init_map(self):
        """State after EKF::init on frame 0: returns (x, P, feat_type, feat_off, desc)."""
        p = self.params
        rng = np.random.default_rng([99, self.seed_offset, self.N])
        uv0 = np.rint(self.true_pixels(0) + rng.normal(0.0, self.noise_px, size=(self.N, 2)))
        x0, P0 = init_state_and_covariance(p)
        x, P = add_features_inverse_depth(p, x0, P0, uv0)
        ftype = np.full(self.N, 2, dtype=np.int32)          # MAPFEATURE_TYPE_INVERSE_DEPTH
        foff = (13 + 6 * np.arange(self.N)).astype(np.int32)
        return x, P, ftype, foff, self.descriptors.copy(), uv0
