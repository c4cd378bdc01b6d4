"""POD mirror of the reference's two configuration structs.

``EkfParams`` has the exact layout of ``ekfb_params`` in include/ekf_b200.h, i.e. the fields of
CameraCalibration (Configuration/ConfigurationDataReader/CameraCalibrationConfiguration/
CameraCalibration.h:37-63) and ExtendedKalmanFilterParameters (.../ExtendedKalmanFilterParameters.h:37-76)
that the per-frame hot path reads.  ``load_config`` reads the reference's YAML 1.0 files
(experiments/s3/config.yml, kalmanFilter/samples/EKF/config.yml): a RunConfiguration block picks one
named entry from each section and every value is a quoted string (ConfigurationManager.cpp:74-111).
"""
import ctypes
import math


class EkfParams(ctypes.Structure):
    _fields_ = [
        ("pixels_x", ctypes.c_int32), ("pixels_y", ctypes.c_int32),
        ("fx", ctypes.c_double), ("fy", ctypes.c_double),
        ("k1", ctypes.c_double), ("k2", ctypes.c_double),
        ("cx", ctypes.c_double), ("cy", ctypes.c_double),
        ("dx", ctypes.c_double), ("dy", ctypes.c_double),
        ("pixel_error_x", ctypes.c_double), ("pixel_error_y", ctypes.c_double),
        ("angular_vision_x", ctypes.c_double), ("angular_vision_y", ctypes.c_double),
        ("init_inv_depth_rho", ctypes.c_double),
        ("init_linear_accel_sd", ctypes.c_double), ("init_angular_accel_sd", ctypes.c_double),
        ("linear_accel_sd", ctypes.c_double), ("angular_accel_sd", ctypes.c_double),
        ("inverse_depth_rho_sd", ctypes.c_double),
        ("matching_coef", ctypes.c_double),
        ("ransac_threshold", ctypes.c_double),
        ("ransac_all_inliers_prob", ctypes.c_double),
        ("ransac_chi2", ctypes.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}



class MapPolicy(ctypes.Structure):
    """ekfb_map_policy (include/ekf_b200.h): the map-management fields of ExtendedKalmanFilterParameters
    (.../ExtendedKalmanFilterParameters.h:37-76).  Defaults = experiments/s3/config.yml."""
    _fields_ = [("min_matches_per_image", ctypes.c_int32), ("max_map_features_count", ctypes.c_int32),
                ("max_map_size", ctypes.c_int32), ("always_remove_unseen", ctypes.c_int32),
                ("good_feature_matching_percent", ctypes.c_double), ("linearity_index_threshold", ctypes.c_double)]

    def __init__(self, min_matches_per_image=60, max_map_features_count=0, max_map_size=240, always_remove_unseen=0,
                 good_feature_matching_percent=0.5, linearity_index_threshold=0.1):
        super().__init__(min_matches_per_image, max_map_features_count, max_map_size, always_remove_unseen,
                         good_feature_matching_percent, linearity_index_threshold)


class MapResult(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n", "n_features", "n_removed_bad", "n_removed_unseen", "converted",
                                              "new_features_needed")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}

# EKF-section defaults of experiments/s3/config.yml:9-37
_S3_EKF = dict(init_inv_depth_rho=1.0, init_linear_accel_sd=0.001, init_angular_accel_sd=0.004,
               linear_accel_sd=0.0007, angular_accel_sd=0.002, inverse_depth_rho_sd=1.0,
               matching_coef=1.0, ransac_threshold=1.0, ransac_all_inliers_prob=0.99,
               ransac_chi2=5.9915)


def synthetic_params(width, height):
    """Camera of SURVEY.md 8d: the S3 calibration (experiments/s3/config.yml:48-63) rescaled to
    width x height with a centred principal point; EKF parameters as in the s3 config."""
    s = width / 640.0
    fx = fy = 525.060143149240389 * s
    p = EkfParams()
    p.pixels_x, p.pixels_y = int(width), int(height)
    p.fx, p.fy = fx, fy
    p.k1, p.k2 = -7.613e-3, 9.388e-4
    p.cx, p.cy = width / 2.0, height / 2.0
    p.dx = p.dy = 0.0070216 / s
    p.pixel_error_x = p.pixel_error_y = 1.0
    p.angular_vision_x = 2.0 * math.atan(width / (2.0 * fx)) * 180.0 / math.pi
    p.angular_vision_y = 2.0 * math.atan(height / (2.0 * fy)) * 180.0 / math.pi
    for k, v in _S3_EKF.items():
        setattr(p, k, v)
    return p


def load_config(path):
    """Parse a reference config.yml into (EkfParams, extras).  extras holds the keys the hot path
    does not read (map-management policy, detector/extractor names)."""
    import yaml
    with open(path, "r") as fh:
        text = fh.read()
    if text.lstrip().startswith("%YAML"):
        text = text.split("\n", 1)[1]
    doc = yaml.safe_load(text)
    run = doc["RunConfiguration"]
    ekf = doc["ExtendedKalmanFilter"][run["ExtendedKalmanFilter"]]
    cam = doc["CameraCalibration"][run["CameraCalibration"]]
    p = EkfParams()
    p.pixels_x, p.pixels_y = int(cam["PixelsX"]), int(cam["PixelsY"])
    for dst, src in (("fx", "FX"), ("fy", "FY"), ("k1", "K1"), ("k2", "K2"), ("cx", "CX"), ("cy", "CY"),
                     ("dx", "DX"), ("dy", "DY"), ("pixel_error_x", "PixelErrorX"),
                     ("pixel_error_y", "PixelErrorY"), ("angular_vision_x", "AngularVisionX"),
                     ("angular_vision_y", "AngularVisionY")):
        setattr(p, dst, float(cam[src]))
    for dst, src in (("init_inv_depth_rho", "InitInvDepthRho"), ("init_linear_accel_sd", "InitLinearAccelSD"),
                     ("init_angular_accel_sd", "InitAngularAccelSD"), ("linear_accel_sd", "LinearAccelSD"),
                     ("angular_accel_sd", "AngularAccelSD"), ("inverse_depth_rho_sd", "InverseDepthRhoSD"),
                     ("matching_coef", "MatchingCompCoefSecondBestVSFirst"),
                     ("ransac_threshold", "RansacThresholdPredictDistance"),
                     ("ransac_all_inliers_prob", "RansacAllInliersProbability"),
                     ("ransac_chi2", "RansacChi2Threshold")):
        setattr(p, dst, float(ekf[src]))
    extras = {
        "max_map_size": int(ekf.get("MaxMapSize", 0)),
        "max_map_features_count": int(ekf.get("MaxMapFeaturesCount", 0)),
        "always_remove_unseen": str(ekf.get("AlwaysRemoveUnseenMapFeatures", "false")).lower() == "true",
        "map_management_frequency": int(ekf.get("MapManagementFrequency", 0)),
        "min_matches_per_image": int(ekf.get("MinMatchesPerImage", 0)),
        "good_feature_matching_percent": float(ekf.get("GoodFeatureMatchingPercent", 0.0)),
        "inverse_depth_linearity_index_threshold": float(ekf.get("InverseDepthLinearityIndexThreshold", 0.0)),
        "feature_detector": run.get("FeatureDetector"),
        "descriptor_extractor": run.get("DescriptorExtractor"),
    }
    return p, extras


_CAM_KEYS = (("PixelsX", "pixels_x"), ("PixelsY", "pixels_y"), ("FX", "fx"), ("FY", "fy"), ("K1", "k1"), ("K2", "k2"),
             ("CX", "cx"), ("CY", "cy"), ("DX", "dx"), ("DY", "dy"), ("PixelErrorX", "pixel_error_x"),
             ("PixelErrorY", "pixel_error_y"), ("AngularVisionX", "angular_vision_x"), ("AngularVisionY", "angular_vision_y"))
_EKF_KEYS = (("InitInvDepthRho", "init_inv_depth_rho"), ("InitLinearAccelSD", "init_linear_accel_sd"),
             ("InitAngularAccelSD", "init_angular_accel_sd"), ("LinearAccelSD", "linear_accel_sd"),
             ("AngularAccelSD", "angular_accel_sd"), ("InverseDepthRhoSD", "inverse_depth_rho_sd"),
             ("MatchingCompCoefSecondBestVSFirst", "matching_coef"), ("RansacThresholdPredictDistance", "ransac_threshold"),
             ("RansacAllInliersProbability", "ransac_all_inliers_prob"), ("RansacChi2Threshold", "ransac_chi2"))


def write_config(path, p, min_matches_per_image, max_map_size=0, extra=None):
    """Write params in the reference's config.yml layout (samples/EKF/config.yml): a RunConfiguration block naming one
    entry per section, every value a quoted string."""
    lines = ["%YAML:1.0", "", "RunConfiguration:", '  ExtendedKalmanFilter: "EKF"', '  FeatureDetector: "STAR"',
             '  DescriptorExtractor: "BRIEF"', '  CameraCalibration: "CAM"', "", "ExtendedKalmanFilter:", "  EKF:"]
    for key, attr in _EKF_KEYS:
        lines.append(f'    {key}: "{getattr(p, attr)!r}"')
    lines.append(f'    MinMatchesPerImage: "{int(min_matches_per_image)}"')
    if max_map_size:
        lines.append(f'    MaxMapSize: "{int(max_map_size)}"')
    for key, val in (extra or {}).items():      # e.g. MapManagementFrequency, GoodFeatureMatchingPercent, ...
        lines.append(f'    {key}: "{val}"')
    lines += ["FeatureDetector:", "  STAR:", '    Type: "STAR"', "DescriptorExtractor:", "  BRIEF:", '    Type: "BRIEF"',
              "CameraCalibration:", "  CAM:"]
    for key, attr in _CAM_KEYS:
        lines.append(f'    {key}: "{getattr(p, attr)!r}"')
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
