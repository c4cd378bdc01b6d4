"""Generates tests/golden/s3_keypoints.npz: the front-end output (keypoints + 32-byte binary descriptors) of the
reference's bundled sequence experiments/s3/costado_recto1 (BASELINE.json config 1), frames 00090.. as the reference's
sample driver reads them (kalmanFilter/samples/EKF/main.cpp:50), plus the parameters of experiments/s3/config.yml.

The reference's own front end (STAR + BRIEF, OpenCV 2.4 nonfree/legacy) does not exist in this image; cv2 4.13's ORB
(FAST corners + rotated BRIEF, 32-byte descriptors, Hamming) stands in for it.  The front end is outside the hot path:
both the reference (oracle/_ref) and the B200 library consume these keypoints through the same seam.
Run in the build container (needs /root/reference and cv2):  python tests/golden/make_s3_keypoints.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openekfmonoslam_b200.params import load_config  # noqa: E402

SEQ = "/root/reference/experiments/s3/costado_recto1"
CFG = "/root/reference/experiments/s3/config.yml"
FIRST, COUNT, NFEAT = 90, 631, 300


def main():
    orb = cv2.ORB_create(nfeatures=NFEAT, scaleFactor=1.2, nlevels=4, edgeThreshold=31, fastThreshold=20)
    off, xy, ds = [0], [], []
    for k in range(FIRST, FIRST + COUNT):
        im = cv2.imread(os.path.join(SEQ, f"{k:05d}.png"))
        kps, d = orb.detectAndCompute(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), None)
        if d is None:
            kps, d = [], np.zeros((0, 32), np.uint8)
        xy.append(np.array([kp.pt for kp in kps], np.float32).reshape(-1, 2))
        ds.append(d.astype(np.uint8))
        off.append(off[-1] + len(kps))
    p, extras = load_config(CFG)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "s3_keypoints.npz"), offsets=np.array(off, np.int32),
                        xy=np.concatenate(xy), desc=np.concatenate(ds), first_frame=FIRST,
                        params=np.array([getattr(p, k) for k, _ in p._fields_], np.float64),
                        param_names=np.array([k for k, _ in p._fields_]))
    print("frames", COUNT, "keypoints", off[-1], "per frame", off[-1] / COUNT, "extras", extras)


if __name__ == "__main__":
    main()
