/*
 * ekf_b200.h -- C ABI of the B200-native per-frame EKF hot path (libekf_b200.so).
 *
 * Drop-in boundary: the bodies of EKF::init / EKF::step of the reference
 * (kalmanFilter/modules/1PointRansacEKF/EKF.h:41-63, EKF.cpp:170-666) call these entry points;
 * everything behind them runs as hand-written sm_100a CUDA kernels.  Plain pointers and sizes
 * only, no exceptions across the boundary, int status return (0 = ok, otherwise an ekfb_status).
 * One handle owns `n_filters` independent filters (the reference's one-EKF-per-process becomes a
 * batch; n_filters = 1 is the single-filter drop-in), one CUDA stream and all device memory.
 * A handle is not thread-safe; distinct handles are independent.  There is NO CPU fallback: every
 * call fails with EKFB_ERR_CUDA if no sm_100 device is usable.
 *
 * Each entry point cites the reference code it replaces.  "E/" = kalmanFilter/modules/1PointRansacEKF/.
 */
#ifndef EKF_B200_H
#define EKF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ekfb_status {
    EKFB_OK = 0,
    EKFB_ERR_ARG = 1,      /* bad argument (null pointer, size over capacity, bad filter index) */
    EKFB_ERR_CUDA = 2,     /* CUDA runtime error; ekfb_last_error() has the text */
    EKFB_ERR_CAPACITY = 3, /* state / keypoint count exceeds what ekfb_create reserved */
    EKFB_ERR_NUMERIC = 4   /* innovation covariance not positive definite */
} ekfb_status;

/* POD mirror of CameraCalibration (modules/Configuration/ConfigurationDataReader/
 * CameraCalibrationConfiguration/CameraCalibration.h:37-63) and of the fields of
 * ExtendedKalmanFilterParameters (.../ExtendedKalmanFilterParameters.h:37-76) the per-frame path
 * reads.  Replaces the process-global ConfigurationManager singleton
 * (modules/Configuration/ConfigurationManager.h:45-64): parameters are per handle. */
typedef struct ekfb_params {
    int32_t pixels_x, pixels_y;
    double fx, fy, k1, k2, cx, cy, dx, dy;
    double pixel_error_x, pixel_error_y;
    double angular_vision_x, angular_vision_y;
    double init_inv_depth_rho, init_linear_accel_sd, init_angular_accel_sd;
    double linear_accel_sd, angular_accel_sd, inverse_depth_rho_sd;
    double matching_coef;           /* MatchingCompCoefSecondBestVSFirst */
    double ransac_threshold;        /* RansacThresholdPredictDistance    */
    double ransac_all_inliers_prob; /* RansacAllInliersProbability       */
    double ransac_chi2;             /* RansacChi2Threshold               */
} ekfb_params;

/* per-filter counters of the last frame (what EKF::step writes to output.yml: totalMatches,
 * liInliers, hiInliers -- E/EKF.cpp:412-416,515-516) */
typedef struct ekfb_frame_info {
    int32_t n;            /* state dimension 13 + 3*N_xyz + 6*N_id */
    int32_t n_features;
    int32_t n_keypoints;
    int32_t n_predicted;  /* features predicted inside the frame        */
    int32_t n_matches;    /* totalMatches                               */
    int32_t n_hypotheses; /* RANSAC hypotheses the sequential rule used */
    int32_t best_hypothesis;
    int32_t n_inliers;    /* liInliers                                  */
    int32_t n_outliers;
    int32_t n_rescued;    /* hiInliers                                  */
    int32_t status;       /* THIS frame: 0, or EKFB_ERR_NUMERIC if an innovation covariance was not positive definite (that update
                             and the rest of the frame's updates were skipped: x and P are left as they were) */
    int32_t reserved;     /* OR of `status` over all frames since ekfb_set_state (0 = no frame ever failed) */
} ekfb_frame_info;

/* fixed-size per-filter result record (what callers read back: main.cpp:91,137,141 read
 * ekf.state / P; A/jni/EKFNative.cpp:191-193 reads state.position).  This is also the payload of
 * the multi-GPU gather. */
typedef struct ekfb_record {
    double x_cam[13];       /* r(3) q(4: w,x,y,z) v(3) omega(3) */
    double P_cam[13 * 13];  /* stateCovarianceMatrix(0:13, 0:13) */
    ekfb_frame_info info;
} ekfb_record;

typedef struct ekfb_ctx* ekfb_handle;

/* ---- lifetime -------------------------------------------------------------------------------- */
/* Replaces EKF::EKF (E/EKF.cpp:124-144) minus file outputs.  max_features bounds N per filter
 * (state capacity 13 + 6*max_features rows), max_keypoints bounds the per-frame front-end output. */
int ekfb_create(const ekfb_params* params, int device, int n_filters, int max_features, int max_keypoints,
                ekfb_handle* out);
int ekfb_destroy(ekfb_handle h);                 /* EKF::~EKF, State::~State (E/State.cpp:75-103) */
const char* ekfb_last_error(void);
int ekfb_sync(ekfb_handle h);                    /* wait for the handle's stream */

/* ---- state in / out ------------------------------------------------------------------------------ */
/* Upload one filter: x (n doubles: camera 13 + features at their covarianceMatrixPos), feature
 * type (1 = XYZ, 2 = inverse depth; E/MapFeature.h:39-44), feat_off (covarianceMatrixPos,
 * E/MapFeature.h:68), P (n x n row-major, E/EKF.h:50), descriptors (N x 32 bytes, E/MapFeature.h:70).
 * This is the state EKF::init leaves behind (E/EKF.cpp:170-237) or any later state. */
int ekfb_set_state(ekfb_handle h, int filter, int n, int n_features, const double* x, const int32_t* feat_type,
                   const int32_t* feat_off, const double* P, const uint8_t* desc);
/* Download.  x and/or P may be NULL.  cam_block_only != 0 copies x[0:13] and P[0:13,0:13] only. */
int ekfb_get_state(ekfb_handle h, int filter, double* x, double* P, int cam_block_only);
int ekfb_get_descriptors(ekfb_handle h, int filter, uint8_t* desc, int32_t* times_predicted, int32_t* times_matched);
int ekfb_get_dims(ekfb_handle h, int filter, int32_t* n, int32_t* n_features);

/* ---- front-end output (what detector->detect + extractor->compute return, E/Matching.cpp:204-215) */
/* host buffers: xy = n_kp x 2 float32 pixel coordinates, desc = n_kp x 32 bytes; copied to the device */
int ekfb_set_keypoints(ekfb_handle h, int filter, const float* xy, const uint8_t* desc, int n_kp);
/* the same for all filters of the handle in one call (xy[f], desc[f], n_kp[f] per filter): one pointer-table / count update */
int ekfb_set_keypoints_batch(ekfb_handle h, const float* const* xy, const uint8_t* const* desc, const int32_t* n_kp);
/* all filters from ONE packed host buffer pair: filter f's keypoints are entries kp_offset[f] .. kp_offset[f+1]-1 of xy / desc
 * (n_filters + 1 offsets).  Two host->device copies per call whatever the number of filters (pinned buffers make them
 * asynchronous); the batch form above issues two per filter. */
int ekfb_set_keypoints_packed(ekfb_handle h, const float* xy, const uint8_t* desc, const int32_t* kp_offset);
/* device-resident sequences: upload all frames once, then select a frame with no host traffic.
 * kp_offset has n_frames+1 entries (prefix sums of per-frame keypoint counts). */
int ekfb_load_sequence(ekfb_handle h, int filter, int n_frames, const int32_t* kp_offset, const float* xy,
                       const uint8_t* desc);
int ekfb_select_frame(ekfb_handle h, int frame); /* all filters take frame `frame` of their sequence */

/* ---- the hot path, phase by phase, all filters of the handle (asynchronous on the handle's stream) */
int ekfb_predict(ekfb_handle h);   /* stateAndCovariancePrediction, E/StateAndCovariancePrediction.cpp:244-253 */
int ekfb_measure(ekfb_handle h);   /* predictCameraMeasurements on all features, E/MeasurementPrediction.cpp:705-719 */
int ekfb_match(ekfb_handle h);     /* matchPredictedFeatures from the ellipse mask on, E/Matching.cpp:181-264 */
int ekfb_ransac(ekfb_handle h);    /* ransac, E/1PointRansac.cpp:101-234 */
int ekfb_update(ekfb_handle h, int which); /* update(), E/Update.cpp:282-319; which 0 = low-innovation
                                              inliers (E/EKF.cpp:430), 1 = rescued (E/EKF.cpp:527-532) */
int ekfb_rescue(ekfb_handle h);    /* re-prediction of outliers + rescueOutliers, E/EKF.cpp:448-506, :68-119 */
int ekfb_update_map_features(ekfb_handle h); /* updateMapFeatures, E/MapManagement.cpp:77-113 */
/* the whole frame in EKF::step order (E/EKF.cpp:242-572): predict, measure, match, ransac, update LI,
 * rescue, update HI, updateMapFeatures */
int ekfb_step(ekfb_handle h);

/* ---- map management on the device (E/EKF.cpp:572-612; the covariance never leaves HBM) -------------- */
/* POD mirror of the map-management fields of ExtendedKalmanFilterParameters (.../ExtendedKalmanFilterParameters.h:37-76) */
typedef struct ekfb_map_policy {
    int32_t min_matches_per_image;      /* MinMatchesPerImage            */
    int32_t max_map_features_count;     /* MaxMapFeaturesCount (0 = off) */
    int32_t max_map_size;               /* MaxMapSize, rows of the state (0 = off) */
    int32_t always_remove_unseen;       /* AlwaysRemoveUnseenMapFeatures */
    double good_feature_matching_percent;   /* GoodFeatureMatchingPercent            */
    double linearity_index_threshold;       /* InverseDepthLinearityIndexThreshold   */
} ekfb_map_policy;
typedef struct ekfb_map_result {
    int32_t n, n_features;              /* sizes after the call */
    int32_t n_removed_bad, n_removed_unseen;
    int32_t converted;                  /* index (new numbering) of the feature converted to XYZ, or -1 */
    int32_t new_features_needed;        /* newFeaturesNeededCount, E/EKF.cpp:577 (may be <= 0) */
} ekfb_map_result;
/* After ekfb_step (or ekfb_update_map_features) of the same frame, for all filters of the handle:
 * removeBadMapFeatures (E/MapManagement.cpp:279-307), removal of the features not predicted in this frame under the policy of
 * E/EKF.cpp:580-589 (removeFeaturesFromStateAndCovariance, E/MapManagement.cpp:212-259) and convertMapFeaturesInverseDepthToDepth
 * (at most one feature per call, E/MapManagement.cpp:311-524).  Synchronises; out (n_filters entries) may be NULL.
 * Per-feature results of the frame (ekfb_get_feature_results) are indexed by the numbering BEFORE this call. */
int ekfb_map_management(ekfb_handle h, const ekfb_map_policy* policy, ekfb_map_result* out);
/* flags of the last ekfb_map_management for the n_features_before features present before it: 0 kept, 1 removed as bad,
 * 2 removed as unseen (lets the host mirror State::removeFeatures, E/State.cpp:199-206) */
int ekfb_get_removed_flags(ekfb_handle h, int filter, int n_features_before, uint8_t* flags);
/* addFeaturesToStateAndCovariance (E/AddMapFeature.cpp:293-366): `count` new inverse-depth features observed at pixels
 * uv (count x 2 doubles, ImageFeatureMeasurement::imagePos) with descriptors desc (count x 32).  EKFB_ERR_CAPACITY if they do not fit. */
int ekfb_add_features(ekfb_handle h, int filter, int count, const double* uv, const uint8_t* desc);
/* feature type (1 XYZ / 2 inverse depth) and covarianceMatrixPos of the current map (E/MapFeature.h:39-44,68) */
int ekfb_get_feature_layout(ekfb_handle h, int filter, int32_t* type, int32_t* off);
/* overwrite timesPredicted / timesMatched (E/MapFeature.h:72-73); for tests and for restoring a saved map */
int ekfb_set_hit_counters(ekfb_handle h, int filter, const int32_t* times_predicted, const int32_t* times_matched);
/* buildImageMask (E/DetectNewImageFeatures.cpp:101-122): 255 everywhere, 0 inside the gate ellipse of every feature predicted
 * in this frame; built by the last ekfb_map_management when a filter asked for new features (pixels_y x pixels_x bytes) */
int ekfb_get_new_feature_mask(ekfb_handle h, int filter, uint8_t* mask);
/* the inputs of the host's new-feature selection in one round trip: predicted flags and pixels of this frame's measurement
 * (n_features_before = the feature count before ekfb_map_management; either may be NULL) and the new-feature mask (or NULL) */
int ekfb_get_new_feature_inputs(ekfb_handle h, int filter, int n_features_before, uint8_t* predicted, double* hpred, uint8_t* mask);
/* the host mirror of one filter in one round trip (any pointer may be NULL): state vector (n), feature type / covarianceMatrixPos /
 * descriptor / timesPredicted / timesMatched (N each), and the result record */
int ekfb_get_map_snapshot(ekfb_handle h, int filter, double* x, int32_t* type, int32_t* off, uint8_t* desc, int32_t* times_predicted,
                          int32_t* times_matched, ekfb_record* record);
/* the last frame's counters from the host mirror, no device round trip (valid right after ekfb_step / ekfb_map_management) */
int ekfb_peek_frame_info(ekfb_handle h, int filter, ekfb_frame_info* info);
/* one drawUncertaintyEllipse2D(img, (cx,cy), S, max_axes, value, filled) (Gui/Draw.cpp:42-64) into a host W x H byte image;
 * the host side uses it once for the stamp of E/DetectNewImageFeatures.cpp:283-288, tests pin the rasteriser with it */
int ekfb_raster_ellipse(ekfb_handle h, int W, int H, double cx, double cy, const double* S, int max_axes, int value,
                        uint8_t* img_inout);

/* ---- detector + descriptor on the device (what detector->detect + extractor->compute do, E/Matching.cpp:204-215) */
/* the frame as 8-bit grey (pixels_y rows of pixels_x bytes, `stride` bytes apart); also builds the NCC pyramid */
int ekfb_set_image(ekfb_handle h, int filter, const uint8_t* gray, int stride);
/* the same from an interleaved 8-bit colour frame (channels = 3: BGR, 4: BGRA; 1 = grey): converted on the device like
 * cv::cvtColor(COLOR_BGR2GRAY) */
int ekfb_set_image_color(ekfb_handle h, int filter, const uint8_t* pixels, int stride, int channels);
/* FAST-9/16 corners with 3x3 non-maximum suppression (OpenCV's FastFeatureDetector(threshold, true), the "FAST" entry of the
 * reference's FeatureDetectorFactory) in raster order, keypoints closer than 18 pixels to the border dropped, each with a
 * 256-bit BRIEF-style descriptor; they replace ekfb_set_keypoints for this frame.  At most max_keypoints are kept. */
int ekfb_detect_keypoints(ekfb_handle h, int filter, int threshold, int32_t* n_keypoints);
/* the current frame's keypoints back to the host (n x 2 float32, n x 32 bytes; n from ekfb_detect_keypoints / ekfb_set_keypoints) */
int ekfb_get_keypoints(ekfb_handle h, int filter, float* xy, uint8_t* desc);

/* ---- NCC active search (the north star's matching path; the reference has no counterpart) ----------- */
/* Matching by normalised cross-correlation of 11 x 11 templates inside each feature's gate ellipse over a 3-level image
 * pyramid, in place of ekfb_match (csrc/ekf_ncc.cuh states the exact rule; oracle/ncc_oracle.py restates it on the CPU).
 * gray: pixels_y rows of pixels_x bytes, `stride` bytes apart; the pyramid is built on the device. */
int ekfb_ncc_set_image(ekfb_handle h, int filter, const uint8_t* gray, int stride);
/* templates of features [first_feature, first_feature + count): count x 3 levels x 121 bytes (row-major 11 x 11) */
int ekfb_ncc_set_templates(ekfb_handle h, int filter, int first_feature, int count, const uint8_t* templates);
/* after ekfb_measure, instead of ekfb_match: fills the same match arrays (z = matched pixel, distance = 1 - score) */
int ekfb_match_ncc(ekfb_handle h, double ncc_min);
/* templates (count x 3 x 121 bytes) and anchors (count x 10 doubles: camera position, quaternion, pixel at capture, valid flag) of
 * features first_feature .. first_feature + count - 1, as captured by ekfb_add_features / set by ekfb_ncc_set_templates and
 * carried through map management; either output may be NULL */
int ekfb_ncc_get_templates(ekfb_handle h, int filter, int first_feature, int count, uint8_t* templates, double* anchors);
/* anchors for caller-supplied templates (same layout; call after ekfb_ncc_set_templates, which clears them) */
int ekfb_ncc_set_anchors(ekfb_handle h, int filter, int first_feature, int count, const double* anchors);
/* acceptance threshold of the NCC matcher when it runs inside ekfb_match / ekfb_step (EKFB_OPT_MATCHER = 1); default 0.8 */
int ekfb_ncc_set_threshold(ekfb_handle h, double ncc_min);
/* level-0 score (-2: no candidate) and start level (-1: not predicted) per feature of the last ekfb_match_ncc */
int ekfb_ncc_get_scores(ekfb_handle h, int filter, double* score, int32_t* level);
/* one pyramid level back to the host (w x h bytes, tightly packed); out may be NULL to query the size */
int ekfb_ncc_get_level(ekfb_handle h, int filter, int level, uint8_t* out, int32_t* w, int32_t* h_out);

/* ---- results ----------------------------------------------------------------------------------- */
int ekfb_get_frame_info(ekfb_handle h, int filter, ekfb_frame_info* info);
/* per-filter records of all filters, to a host buffer (n_filters records) */
int ekfb_get_records(ekfb_handle h, ekfb_record* out);
/* same, written to a caller-provided DEVICE buffer on the handle's stream (NCCL gather payload) */
int ekfb_write_records_device(ekfb_handle h, void* device_out);

/* per-feature results of the last frame, arrays of length N (any pointer may be NULL):
 * predicted flag, h[2N], S_i[4N], Hx[14N] (2x7), Hf[12N] (2x6), matched flag, z[2N], matched keypoint
 * index, descriptor distance, inlier / outlier / rescued flags.  For parity tests and for the host
 * side map management (E/MapManagement.cpp). */
int ekfb_get_feature_results(ekfb_handle h, int filter, uint8_t* predicted, double* hpred, double* S, double* Hx,
                             double* Hf, uint8_t* matched, double* z, int32_t* kp_index, float* dist,
                             uint8_t* inlier, uint8_t* outlier, uint8_t* rescued);
/* the detector mask of the last ekfb_match (pixels_y x pixels_x bytes) and the per-keypoint pass flag */
int ekfb_get_mask(ekfb_handle h, int filter, uint8_t* mask, uint8_t* kp_ok);

/* ---- isolated kernels for the stress sweep (BASELINE.json config 5) and for tests ---------------- */
/* Covariance downdate alone on filter 0: P <- P - W W^T with W^T given as k x n row-major (host). */
int ekfb_test_downdate(ekfb_handle h, int n, int k, const double* P_in, const double* Wt, double* P_out);
/* Test hook for the factorisation of the innovation covariance alone (U2, Update.cpp:92-99 forms S and inverts it): S_in is
 * k x (k+1) row-major, [S | nu] with S symmetric positive definite.  U_out (k x (k+1)) receives the upper Cholesky factor
 * (S = U^T U; entries below the diagonal are not written) and, in column k, y = U^-T nu.  Uinv_out receives the inverses of
 * the ceil(k/64) diagonal 64x64 blocks of U (row-major, identity-padded). */
int ekfb_test_factor(ekfb_handle h, int k, const double* S_in, double* U_out, double* Uinv_out);
/* Timing hook.  STATE-DESTRUCTIVE: applies `reps` real updates (`which` = 0 low / 1 high innovation) on the update list of the
 * last RANSAC / rescue to the live x and P, then runs the 128x64-tile downdate kernel alone `reps` times; returns the
 * per-repetition times in ms.  Re-upload the state with ekfb_set_state afterwards. */
int ekfb_time_update(ekfb_handle h, int which, int reps, float* ms_total, float* ms_downdate);

/* ---- device timing on the handle's stream (bench.py cannot see this stream from torch) ------------ */
int ekfb_timer_record(ekfb_handle h, int slot);                 /* slot in [0, 62): slots 62 and 63 are used by ekfb_time_update */
int ekfb_timer_elapsed_ms(ekfb_handle h, int slot_a, int slot_b, float* ms);
/* enable per-kernel-group timing of ekfb_step (adds event records; off by default) */
int ekfb_profile_enable(ekfb_handle h, int on);
/* accumulated ms per group since the last call: predict, measure, match, ransac, gain, chol, downdate,
 * rescue, misc (9 floats), plus launch counts (9 ints) */
int ekfb_profile_read(ekfb_handle h, float* ms9, int32_t* launches9);
int64_t ekfb_kernel_launches(ekfb_handle h);     /* kernels launched by this handle so far */
/* tuning / test switches.  EKFB_OPT_FORCE_GENERIC_FACTOR forces the paths used when k is too large for the shared-memory slab
 * TRSM: 2 = S-chain + blocked TRSM on the global-memory resident B (tensor-map fed DMMA kernel; the default for a single
 * filter), 1 = right-looking factorisation over the whole augmented matrix (batches, or no tensor-map support). */
enum { EKFB_OPT_FORCE_GENERIC_FACTOR = 1, EKFB_OPT_DOWNDATE_VARIANT = 2 /* covariance downdate: 0 = 64x64 tiles fed by cp.async, 1 = 128x64 tiles, 2 = single filter: persistent
                                        TMA-fed kernel (tensor-map loads / stores, mbarrier ring, 128-byte swizzle), 3 = as 2 without swizzle */,
       EKFB_OPT_SCHAIN_VARIANT = 3 /* factorisation of S: 0 = one fused launch per 64-row step, 1 = panel + trail launches,
                                      3 = the whole chain in one launch (tile dataflow, ekf_chain.cuh),
                                      4 = as 3, and for a single filter the slab TRSM runs inside the same launch, overlapped with the chain */,
       EKFB_OPT_DOWNDATE_SMALL_K = 4 /* updates with at most this many rows run the downdate as 64x64 tiles, 4 CTAs/SM (default: all); above it 128x128 tiles */,
       EKFB_OPT_TRSM_STAGES = 5 /* upper limit of the slab TRSM's operand-ring depth: 2, 3 or 4 (default 4, as shared memory allows) */,
       EKFB_OPT_TRSM_PAIR = 6 /* batched filters: 1 (default) = slab footprint that lets two CTAs share an SM when possible */,
       EKFB_OPT_RANSAC_CHUNK = 7 /* RANSAC hypotheses evaluated per round (0 = default: 16 single filter, 4 batched) */,
       EKFB_OPT_PDL = 8 /* 1 (default): the frame's kernels are launched with programmatic stream serialisation */,
       EKFB_OPT_FAULT_INJECT = 9 /* test hook: 1 = the next factorisations report a non-positive pivot (EKFB_ERR_NUMERIC path) */,
       EKFB_OPT_SMALL_UPDATE = 10 /* 1 (default): updates of at most 128 rows run factorisation + slab TRSM in one launch, the
                                     factorisation redone by every slab CTA (ekf_update_small.cuh); 0 = per-block-step launches */,
       EKFB_OPT_LANES = 11 /* batched handles: the filters run as this many lanes on their own streams, interleaved by ekfb_step
                              (0 / -1 = automatic: 2 for 8 or more filters; 1 = off) */,
       EKFB_OPT_DOWNDATE_CTAS = 12 /* persistent CTAs per SM of the TMA-fed downdate: 2 (default) or 1 (leaves half of every SM
                                      to the kernels of another lane / stream) */,
       EKFB_OPT_DOWNDATE_PROBE = 13 /* timing probe of the TMA-fed downdate (results are wrong when set): 1 = no DMMA, 2 = no stores,
                                       4 = no mirror store */,
       EKFB_OPT_MATCHER = 14 /* matcher used by ekfb_match / ekfb_step: 0 (default) = the reference's descriptor matcher, 1 = the NCC
                                active search (needs ekfb_ncc_set_image / ekfb_set_image every frame; templates captured on the device
                                by ekfb_add_features or supplied with ekfb_ncc_set_templates) */,
       EKFB_OPT_NCC_WARP = 15 /* 1 (default): templates with an anchor are warped to the current camera before the comparison */,
       EKFB_OPT_NCC_TMA_WINDOW = 16 /* 1 (default): the NCC search stages its window by one tensor-map TMA load per level, 0 = by
                                       bulk row copies (the results are identical) */,
       EKFB_OPT_SLAB_TRSM_MAX_K = 17 /* single filter: updates with more rows than this use the TRSM on the global-memory resident B
                                        instead of 16-column shared-memory slabs (default 1152 = as long as the slabs fit) */ };
int ekfb_set_option(ekfb_handle h, int option, int value);
/* developer aid: 64 device-side cycle counters written by instrumented kernels */
int ekfb_debug_read(ekfb_handle h, long long* out64);
/* write a scratch buffer larger than L2 on the handle's stream (timing hygiene between iterations) */
int ekfb_flush_l2(ekfb_handle h);
/* Per-launch timing of the covariance downdate kernel (the roofline kernel).  enable != 0 starts
 * recording one CUDA-event pair around every downdate launch (no host sync in the stream);
 * ekfb_downdate_stats synchronises and returns the summed kernel time, the number of launches and
 * the algorithmic flops n(n+1)k summed over those launches, then clears the pool. */
int ekfb_downdate_timing(ekfb_handle h, int enable);
int ekfb_downdate_stats(ekfb_handle h, double* ms_total, int64_t* launches, double* flops_total,
                        double* bytes_min_total);
/* the same measurement launch by launch (call before ekfb_downdate_stats, which resets it): ms[i] / flops[i] = duration and
 * algorithmic flop (n (n + 1) K summed over the handle's filters) of the i-th timed launch; *count = number of launches */
int ekfb_downdate_launches(ekfb_handle h, int cap, float* ms, double* flops, int32_t* count);

#ifdef __cplusplus
}
#endif
#endif /* EKF_B200_H */
