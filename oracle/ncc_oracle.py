"""CPU restatement (numpy, exact integer sums) of the NCC active-search rule of openekfmonoslam_b200/csrc/ekf_ncc.cuh.
TEST INFRASTRUCTURE: only tests/ and tools/ import it.

Parity status: the reference (segeschecho/OpenEKFMonoSLAM) has NO patch / NCC search -- its matcher is the descriptor path
(kalmanFilter/modules/1PointRansacEKF/Matching.cpp:181-264, restated in ekf_oracle.cpp and pinned against the reference's
own code).  This path exists because BASELINE.json's north star names it; its specification is this repository's own
(csrc/ekf_ncc.cuh header), so the oracle below pins the CUDA kernel to that specification bit for bit, not to the
reference.  The gate is the reference's foci test (Core/EKFMath.cpp:302-351) through the pinned oracle primitive."""
import numpy as np

from . import oracle_lib

P, R_MAX, LEVELS = 11, 12, 3


def pyramid(gray):
    """L(l+1)(y, x) = (a + b + c + d + 2) >> 2 over the 2x2 block; floor(W/2) x floor(H/2)"""
    levels = [np.ascontiguousarray(gray, np.uint8)]
    for _ in range(LEVELS - 1):
        s = levels[-1].astype(np.int32)
        h, w = s.shape[0] // 2, s.shape[1] // 2
        s = s[:2 * h, :2 * w]
        levels.append(((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8))
    return levels


def cut_templates(levels, xy):
    """11 x 11 patch per level centred on (int(x) >> l, int(y) >> l); zero where it leaves the image"""
    out = np.zeros((len(xy), LEVELS, P * P), np.uint8)
    for i, (x, y) in enumerate(xy):
        for l, img in enumerate(levels):
            cx, cy = int(x) >> l, int(y) >> l
            pad = np.zeros((img.shape[0] + 2 * P, img.shape[1] + 2 * P), np.uint8)
            pad[P:-P, P:-P] = img
            out[i, l] = pad[cy + P - 5:cy + P + 6, cx + P - 5:cx + P + 6].ravel()
    return out


def anchors(x, uv):
    """anchor record of features captured at pixels uv while the camera state is x[:13]: r0[3], q0[4], (u0, v0), valid"""
    a = np.zeros((len(uv), 10))
    a[:, :7] = np.asarray(x, np.float64)[:7]
    a[:, 7:9] = np.asarray(uv, np.float64).reshape(-1, 2)
    a[:, 9] = 1.0
    return a


def _rot(q):
    """R(q) of Core/EKFMath.cpp:121-141 (q need not be normalised)"""
    r, x, y, z = q
    return np.array([[r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)],
                     [2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)],
                     [2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z]])


def point_of(ftype, y):
    """world point of a feature: XYZ (type 1) or inverse depth (type 2: anchor + m(theta, phi) / rho)"""
    if ftype == 2:
        th, ph, rho = y[3], y[4], y[5]
        m = np.array([np.cos(ph) * np.sin(th), -np.sin(ph), np.cos(ph) * np.cos(th)])
        return np.asarray(y[:3]) + m / rho
    return np.asarray(y[:3], np.float64)


def warp_matrix(cam, an, X, r1, q1):
    """A = d(current pixel)/d(anchor pixel) of the plane-induced warp (csrc/ekf_ncc.cuh header), or None when the raw template
    is to be used.  cam = (fx, fy, cx, cy)."""
    fx, fy, cx, cy = cam
    R0, R1 = _rot(an[3:7]), _rot(q1)
    nrm = np.asarray(X, np.float64) - an[:3]
    ln = np.sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2])
    if not ln > 0.0:
        return None
    nrm = nrm / ln
    uvw = []
    for e in range(3):
        u, w = an[7] + (1.0 if e == 1 else 0.0), an[8] + (1.0 if e == 2 else 0.0)
        dc = np.array([(u - cx) / fx, (w - cy) / fy, 1.0])
        dw = np.array([R0[a, 0] * dc[0] + R0[a, 1] * dc[1] + R0[a, 2] * dc[2] for a in range(3)])
        den = nrm[0] * dw[0] + nrm[1] * dw[1] + nrm[2] * dw[2]
        t = ln / den
        pw = an[:3] + t * dw - np.asarray(r1, np.float64)
        pc = np.array([R1[0, a] * pw[0] + R1[1, a] * pw[1] + R1[2, a] * pw[2] for a in range(3)])
        uvw.append((cx + fx * pc[0] / pc[2], cy + fy * pc[1] / pc[2]))
    A = np.array([[uvw[1][0] - uvw[0][0], uvw[2][0] - uvw[0][0]], [uvw[1][1] - uvw[0][1], uvw[2][1] - uvw[0][1]]])
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
    if not (0.25 <= det <= 4.0):
        return None
    dev = max(abs(A[0, 0] - 1.0), abs(A[1, 1] - 1.0), abs(A[0, 1]), abs(A[1, 0]))
    return A if dev >= 0.05 else None


def warp_template(T, A):
    """T'(tx, ty) = bilinear sample of the 11 x 11 template T at (5, 5) + A^-1 (tx - 5, ty - 5), clamped, rounded to nearest"""
    det = A[0, 0] * A[1, 1] - A[0, 1] * A[1, 0]
    Ai = np.array([[A[1, 1] / det, -A[0, 1] / det], [-A[1, 0] / det, A[0, 0] / det]])
    T = T.reshape(P, P).astype(np.float64)
    out = np.zeros((P, P), np.uint8)
    for ty in range(P):
        for tx in range(P):
            ox, oy = tx - 5.0, ty - 5.0
            sx = min(max(5.0 + (Ai[0, 0] * ox + Ai[0, 1] * oy), 0.0), 10.0)
            sy = min(max(5.0 + (Ai[1, 0] * ox + Ai[1, 1] * oy), 0.0), 10.0)
            x0, y0 = min(int(sx), 9), min(int(sy), 9)
            fx_, fy_ = sx - x0, sy - y0
            top = T[y0, x0] + fx_ * (T[y0, x0 + 1] - T[y0, x0])
            bot = T[y0 + 1, x0] + fx_ * (T[y0 + 1, x0 + 1] - T[y0 + 1, x0])
            out[ty, tx] = int(top + fy_ * (bot - top) + 0.5)
    return out.ravel()


def warped_templates(templates, anc, cam, x, ftype, foff):
    """the templates the search compares with: warped for the camera state x[:7] where the feature has an anchor and the warp is
    not negligible; returns (templates', warped flags)"""
    out = np.array(templates, np.uint8, copy=True)
    flags = np.zeros(len(templates), bool)
    for j in range(len(templates)):
        if anc is None or anc[j, 9] == 0.0:
            continue
        X = point_of(int(ftype[j]), np.asarray(x[foff[j]:foff[j] + 6]))
        A = warp_matrix(cam, anc[j], X, x[:3], x[3:7])
        if A is None:
            continue
        flags[j] = True
        for l in range(LEVELS):
            out[j, l] = warp_template(templates[j, l], A)
    return out, flags


def _score(img, tmpl, px, py):
    w = img[py - 5:py + 6, px - 5:px + 6].astype(np.int64).ravel()
    t = tmpl.astype(np.int64)
    n = P * P
    dt = n * int((t * t).sum()) - int(t.sum()) ** 2
    dw = n * int((w * w).sum()) - int(w.sum()) ** 2
    if dt == 0 or dw == 0:
        return None
    num = n * int((t * w).sum()) - int(t.sum()) * int(w.sum())
    return float(np.float64(num) / np.sqrt(np.float64(dt) * np.float64(dw)))


def search(levels, templates, vis, h, ellax, ellang, ncc_min=0.8):
    """vis (N), h (N,2) predicted pixels, ellax (N,2) float32 gate axes, ellang (N) gate angle: the measurement prediction.
    Returns matched (N) uint8, z (N,2), score (N) (-2 = none), level (N) (-1 = not predicted)."""
    N = len(vis)
    matched = np.zeros(N, np.uint8); z = np.zeros((N, 2)); score = np.full(N, -2.0); level = np.full(N, -1, np.int32)
    for j in range(N):
        if not vis[j]:
            continue
        aw, ah = int(np.rint(np.float32(ellax[j, 0]))), int(np.rint(np.float32(ellax[j, 1])))
        cxf, cyf = np.float32(h[j, 0]), np.float32(h[j, 1])
        R0 = max(aw, ah)
        lev = 0
        while lev < LEVELS - 1 and ((R0 + (1 << lev) - 1) >> lev) > R_MAX:
            lev += 1
        R = min(R_MAX, (R0 + (1 << lev) - 1) >> lev)
        level[j] = lev
        cx, cy = int(cxf) >> lev, int(cyf) >> lev
        ok, first = True, True
        for l in range(lev, -1, -1):
            img = levels[l]
            Hl, Wl = img.shape
            best = None
            for dy in range(-R, R + 1):
                for dx in range(-R, R + 1):
                    px, py = cx + dx, cy + dy
                    if not (5 <= px < Wl - 5 and 5 <= py < Hl - 5):
                        continue
                    if first and not oracle_lib.point_in_ellipse(float(px << l), float(py << l), float(cxf), float(cyf), aw, ah,
                                                                 float(ellang[j])):
                        continue
                    s = _score(img, templates[j, l], px, py)
                    if s is not None and (best is None or s > best[0]):     # scan order = ties to smaller dy, then dx
                        best = (s, px, py)
            if best is None:
                ok = False
                break
            if l == 0:
                score[j] = best[0]
                cx, cy = best[1], best[2]
            else:
                cx, cy, R, first = 2 * best[1], 2 * best[2], 1, False
        if ok and score[j] >= ncc_min:
            matched[j] = 1
            z[j] = (cx, cy)
    return matched, z, score, level
