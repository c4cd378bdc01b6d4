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
