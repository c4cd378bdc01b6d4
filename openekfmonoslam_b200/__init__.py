"""B200-native per-frame EKF hot path of OpenEKFMonoSLAM (1-point RANSAC monocular EKF-SLAM).

The compute path is hand-written CUDA for sm_100a behind the C ABI declared in
``include/ekf_b200.h`` (built into ``openekfmonoslam_b200/lib/libekf_b200.so``).  This package is
the thin host side: parameter structs, the ctypes binding, the Python mirror of the reference's
``EKF`` class (kalmanFilter/modules/1PointRansacEKF/EKF.h:41-63) and the synthetic-sequence
generator used by the tests and bench.  There is no CPU fallback: every entry point raises if the
CUDA library is missing.
"""
from .params import EkfParams, load_config, synthetic_params  # noqa: F401
