/* svo_b200.h -- C-ABI of the B200-native stereo visual-odometry hot path.
 *
 * Drop-in boundary for the data-parallel path of famoreno/stereo-vo's
 * rso::CStereoOdometryEstimator (libstereo-odometry/include/libstereo-odometry.h:147-1047) in its
 * FAST+ORB / smDescBF / ifmDescBF / robust Gauss-Newton configuration.  The reference has no FFI of
 * its own: its boundary is the C++ class API plus the private stage seams.  Every entry point
 * below names the reference interface it replaces (file:line under /root/reference/libstereo-odometry).
 * The C++ host shim with the reference's class/member names lives in stereo_vo_b200/host/.
 *
 * Conventions
 *  - plain C, plain pointers and sizes; no C++/torch types cross this boundary
 *  - every function returns SVO_OK (0) or a negative svo_status; the message of the last failure
 *    is available from svo_last_error().  Numerical / tracking failures are NOT errors: they travel
 *    in svo_result.error_code with the reference's VOErrorCode values (libstereo-odometry.h:142)
 *  - the library never frees or keeps caller memory; outputs go to caller buffers with capacities
 *  - one svo_ctx per GPU; it owns `max_seqs` independent sequence slots (one slot == one
 *    CStereoOdometryEstimator instance's mutable state: prev/cur frame, last pose, ID counter)
 *  - there is no CPU fallback: every entry point fails with SVO_ERR_CUDA when no device is usable
 */
#ifndef SVO_B200_H
#define SVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_B200_ABI_VERSION 4

typedef enum svo_status {
    SVO_OK = 0,
    SVO_ERR_ARG = -1,        /* bad argument (the reference throws via THROW_EXCEPTION / ASSERT_) */
    SVO_ERR_CUDA = -2,       /* CUDA runtime failure or no usable device */
    SVO_ERR_CAPACITY = -3,   /* a per-image capacity of the context was exceeded (see svo_config) */
    SVO_ERR_STATE = -4       /* call not valid in the current state (e.g. no current frame) */
} svo_status;

/* VOErrorCode, libstereo-odometry.h:142 -- same numeric values */
typedef enum svo_vo_error {
    SVO_VOEC_NONE = 0,
    SVO_VOEC_BAD_COND_NUMBER = 1,
    SVO_VOEC_INCR_FUNC_COST_STG1 = 2,
    SVO_VOEC_INCR_FUNC_COST_STG2 = 3,
    SVO_VOEC_FIRST_ITERATION = 4,
    SVO_VOEC_BAD_TRACKING = 5
} svo_vo_error;

/* cv::KeyPoint (28 B) and cv::DMatch (16 B) field order, so the shim can memcpy
 * (TKeyPointList / TDMatchList, libstereo-odometry.h:108-109) */
typedef struct svo_keypoint { float x, y, size, angle, response; int32_t octave, class_id; } svo_keypoint;
typedef struct svo_dmatch { int32_t queryIdx, trainIdx, imgIdx; float distance; } svo_dmatch;

/* the fields of mrpt::utils::TStereoCamera that the path reads (stage5_optimization.cpp:510-515) */
typedef struct svo_camera {
    double fxl, fyl, cxl, cyl;   /* leftCamera.fx(), fy(), cx(), cy() */
    double fxr, fyr, cxr, cyr;   /* rightCamera */
    double baseline;             /* rightCameraPose[0] */
} svo_camera;

/* Parameters of the path; names and defaults follow the reference's param structs
 * (TDetectParams :383-447, TLeftRightMatchParams :449-499, TLeastSquaresParams :313-350,
 *  TGeneralParams :266-283).  svo_default_config() fills the values of SURVEY.md Appendix C. */
typedef struct svo_config {
    int32_t abi_version;          /* SVO_B200_ABI_VERSION */
    int32_t device;               /* CUDA device ordinal */
    int32_t width, height;        /* image size (all slots) */
    int32_t max_seqs;             /* number of sequence slots (batch size upper bound) */
    int32_t max_corners;          /* capacity: FAST corners per image that survive the 31-px border filter */
    /* TDetectParams */
    int32_t orb_nfeats;           /* 500 -> cap 2*orb_nfeats keypoints (stage2_detect.cpp:405, nOctaves=1) */
    int32_t fast_threshold;       /* m_current_fast_th == initial_FAST_threshold, 20 */
    int32_t non_maximal_suppression; /* true */
    int32_t min_distance;         /* 3 (must be >= 2: the reference divides by uint(min_distance/2)) */
    /* TLeftRightMatchParams */
    double  max_y_diff;           /* 1 */
    int32_t orb_threshold;        /* m_current_orb_th == orb_max_distance, 40 */
    int32_t enable_robust_1to1_match; /* false */
    /* inter-frame: the reference always runs both cv::findFundamentalMat filters
     * (stage4_match_consecutive.cpp:202,237); 0 skips them (documented deviation for A/B tests) */
    int32_t fundmat_filter;       /* 1 */
    /* TLeastSquaresParams */
    int32_t use_robust_kernel;    /* true */
    double  kernel_param;         /* 3 */
    int32_t max_iters;            /* 100 */
    int32_t initial_max_iters;    /* 10 */
    double  min_mod_out_vector;   /* 1e-3 */
    int32_t max_incr_cost;        /* 3 */
    double  residual_threshold;   /* 10 */
    int32_t bad_tracking_th;      /* 5 */
    int32_t use_previous_pose_as_initial; /* true */
    /* TGeneralParams */
    int32_t vo_use_matches_ids;   /* false */
    /* 0 = reproduce libstdc++ std::sort's permutation (stage2_detect.cpp:89,242,324) exactly;
     * 1 = stable order (diagnostic; compare with the oracle's sort_mode=1) */
    int32_t sort_mode;
    /* alternative matchers (ABI version 2) */
    int32_t match_method;         /* TLeftRightMatchParams::TSMMethod: 0 = smDescBF (stage3_match_left_right.cpp:83-178),
                                     1 = smDescRbR, the row-by-row / epipolar-band matcher (:185-419) */
    int32_t if_match_method;      /* TInterFrameMatchingParams::TIFMMethod: 0 = ifmDescBF (stage4_match_consecutive.cpp:88-329),
                                     1 = ifmDescWin, the windowed descriptor tracker (:435-738) */
    int32_t ifm_win_w, ifm_win_h; /* window of ifmDescWin (INI keys window_width / window_height; no default in the
                                     reference, 20 / 20 here) */
    double  orb_max_distance;     /* params_lr_match.orb_max_distance, the threshold smDescRbR reads (:206); 40 */
    /* image pyramid (ABI version 3) */
    int32_t n_octaves;            /* params_rectify.nOctaves (stage1_rectify.cpp:28,80): 1..4.  1 here;
                                     the reference's ctor default is 3 */
    int32_t pyr_mode;             /* how CImagePyramid::buildPyramidFast halves an octave (stage1_rectify.cpp:82-83):
                                     0 = as MRPT 1.x picks (2x2 rounded average when the width is a multiple of 16,
                                     cvPyrDown otherwise), 1 = always the 2x2 average, 2 = always cv::pyrDown */
    /* detector (ABI version 4) */
    int32_t detect_method;        /* TDetectParams::TDMethod (libstereo-odometry.h:388): 1 = dmFAST_ORB (stage2_detect.cpp:502-515,
                                     the named path), 0 = dmORB: the full multi-scale cv::ORB::detectAndCompute (:458-496) with
                                     1.5 x orb_nfeats features (with NMS), Harris score, oriented + steered BRIEF; keypoints
                                     then carry float coordinates, angle and octave.  n_octaves must be 1 with dmORB */
    int32_t orb_nlevels;          /* params_detect.orb_nlevels: pyramid levels of cv::ORB (8) */
} svo_config;

/* POD mirror of TStereoOdometryResult (libstereo-odometry.h:235-264).  detected_* / stereo_matches describe octave 0
 * (svo_get_octave_stats has every octave); the tracked counters run over all octaves as in the reference.
 * The variable-length members (outliers, out_residual) are fetched with svo_get_* below. */
typedef struct svo_result {
    double  out_pose[16];         /* outPose as row-major 4x4 (pose of the current frame wrt the previous) */
    double  out_pose_rotvec[3];   /* same pose: rotation vector */
    double  out_pose_t[3];        /*            translation */
    double  out_pose_ypr[3];      /*            yaw, pitch, roll (CPose3D convention) */
    double  delta_pose[6];        /* the optimiser's (w1,w2,w3,t1,t2,t3) = inverse of outPose */
    int32_t num_it, num_it_final;
    int32_t valid;
    int32_t error_code;           /* svo_vo_error */
    int32_t tracked_feats_from_last_KF, tracked_feats_from_last_frame;
    int32_t detected_left, detected_right;   /* detected_feats[0].first / .second */
    int32_t stereo_matches;                  /* stereo_matches[0] */
    int32_t n_outliers;                      /* outliers.size()  (holds the INLIERS, stage5_optimization.cpp:601-610) */
    int32_t n_residuals;                     /* out_residual.size() */
    int32_t status;                          /* svo_status of this slot's step (capacity overflow etc.) */
} svo_result;

typedef struct svo_ctx svo_ctx;

/* ---- lifetime ----------------------------------------------------------------------------- */
void svo_default_config(svo_config* cfg);                       /* ctor defaults + SURVEY Appendix C */
int  svo_create(const svo_config* cfg, svo_ctx** out);          /* CStereoOdometryEstimator(), common.cpp:28-50 */
void svo_destroy(svo_ctx* ctx);                                 /* ~CStereoOdometryEstimator(), common.cpp:52-61 */
const char* svo_last_error(const svo_ctx* ctx);                 /* ctx may be NULL: last creation error */
int  svo_reset_slot(svo_ctx* ctx, int32_t slot);                /* fresh estimator state for one slot (-1: every slot) */

/* setFASTThreshold / setORBThreshold / resetFASTThreshold / resetORBThreshold (libstereo-odometry.h:529-541):
 * m_current_fast_th and m_current_orb_th of one slot (slot = -1: every slot), read at every following call
 * (stage2_detect.cpp:510, stage3_match_left_right.cpp:164, stage4_match_consecutive.cpp:149).  Nothing else of the
 * slot's state changes: external adaptive loops adjust these between frames while tracking continues.  A negative value
 * keeps that threshold; the clamping to [fast_min_th, fast_max_th] / [orb_min_th, orb_max_th] (:531,:538) is the
 * caller's (the host shim does it).  Stream-ordered: applies to steps submitted after this call.
 * svo_reset_slot returns a slot to the configured values (svo_config.fast_threshold / orb_threshold). */
int  svo_set_thresholds(svo_ctx* ctx, int32_t slot, int32_t fast_th, int32_t orb_th);
int  svo_get_thresholds(svo_ctx* ctx, int32_t slot, int32_t* fast_th, int32_t* orb_th);   /* getFASTThreshold / getORBThreshold */

/* ---- the pipeline: processNewImagePair (process_new_image_pair.cpp:41-385) ----------------- */
/* One call == one processNewImagePair per listed slot, all slots advanced together on the GPU.
 * left/right: n pointers to 8-bit images (host memory), row stride in bytes: gray and rectified unless
 * svo_set_input_format / svo_set_rectify_maps say otherwise (stage 1 then runs on the device).
 * slots: n distinct slot indices (NULL = 0..n-1).  cams: n cameras (one per slot).
 * repeat: request_data.repeat per slot (NULL = all false).  results: n entries (host). */
int svo_process_batch(svo_ctx* ctx, int32_t n, const int32_t* slots,
                      const uint8_t* const* left, const uint8_t* const* right, size_t stride_bytes,
                      const svo_camera* cams, const uint8_t* repeat, svo_result* results);

/* The same step split in two for pipelining: svo_submit_batch enqueues the host->device copies (on a
 * copy stream), the kernels and the device->host copy of the results and returns at once; svo_wait_batch
 * blocks until the OLDEST submitted batch is complete and writes its results (in the order of that
 * batch's slot list).  At most three batches may be in flight, so the copies of batches t+1, t+2 overlap the
 * kernels of batch t.  The host image buffers must stay valid (and should be pinned) until the wait. */
int svo_submit_batch(svo_ctx* ctx, int32_t n, const int32_t* slots,
                     const uint8_t* const* left, const uint8_t* const* right, size_t stride_bytes,
                     const svo_camera* cams, const uint8_t* repeat);
int svo_wait_batch(svo_ctx* ctx, svo_result* results);
/* The variable-length members of TStereoOdometryResult (libstereo-odometry.h:239-240) with every batch:
 * after svo_set_result_detail(ctx, 1) each submitted batch also copies out_residual and outliers of its slots to the
 * host (same stream, same event), and svo_wait_batch_detail hands them out: item i's vectors start at
 * residuals[i*cap] / outliers[i*cap], results[i].n_residuals / n_outliers entries are valid (at most cap each;
 * the context's capacity is 2*orb_nfeats per octave).  octave_stats: 16 values per item, 4 per octave as
 * svo_get_octave_stats returns them (detected left, detected right, stereo matches, tracked pairs) -- the state of THAT
 * batch, while svo_get_octave_stats reads the newest step.  Any of the three arrays may be NULL. */
int svo_set_result_detail(svo_ctx* ctx, int32_t on);
int svo_wait_batch_detail(svo_ctx* ctx, svo_result* results, double* residuals, int64_t* outliers, int32_t cap,
                          int32_t* octave_stats);
/* page-locked host memory for image buffers given to svo_submit_batch (cudaHostAlloc, portable); NULL on failure */
void* svo_host_alloc(size_t bytes);
void  svo_host_free(void* p);

/* Same, images already resident in device memory (HBM): d_left/d_right point to n images each,
 * image i at d_left + i*image_pitch_bytes, rows `stride_bytes` apart.  Results stay on the device
 * until svo_fetch_results(); the call only enqueues work on the context's stream. */
int svo_process_batch_device(svo_ctx* ctx, int32_t n, const int32_t* slots,
                             const uint8_t* d_left, const uint8_t* d_right, size_t stride_bytes,
                             size_t image_pitch_bytes, const svo_camera* cams, const uint8_t* repeat);
int svo_fetch_results(svo_ctx* ctx, int32_t n, const int32_t* slots, svo_result* results); /* syncs */
int svo_synchronize(svo_ctx* ctx);

/* Single pair on slot 0: the literal processNewImagePair(request, result) seam. */
int svo_process_pair(svo_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t stride_bytes,
                     const svo_camera* cam, int32_t repeat, svo_result* result);

/* ---- state readers (getValues, libstereo-odometry.h:704-724, and result vectors) ------------ */
/* which: 0 = current frame, 1 = previous frame.  side: 0 = left, 1 = right.
 * Each returns the element count (>= 0) or a negative svo_status; at most `cap` elements are written. */
int svo_get_keypoints(svo_ctx* ctx, int32_t slot, int32_t which, int32_t side, svo_keypoint* out, int32_t cap);
int svo_get_descriptors(svo_ctx* ctx, int32_t slot, int32_t which, int32_t side, uint8_t* out, int32_t cap_rows);
int svo_get_matches(svo_ctx* ctx, int32_t slot, int32_t which, svo_dmatch* out, int32_t cap);
int svo_get_match_ids(svo_ctx* ctx, int32_t slot, int32_t which, int64_t* out, int32_t cap);
int svo_get_tracked_pairs(svo_ctx* ctx, int32_t slot, int64_t* out_pairs /* cap x 2 */, int32_t cap);
int svo_get_residuals(svo_ctx* ctx, int32_t slot, double* out, int32_t cap);
int svo_get_outliers(svo_ctx* ctx, int32_t slot, int64_t* out, int32_t cap);
/* ---- octaves (params_rectify.nOctaves > 1; per-octave vectors of TImagePairData, libstereo-odometry.h:756-811) ---- */
/* The readers above address octave 0; these take the octave.  svo_get_tracked_pairs_oct = tracked_pairs[octave]
 * (stage4_match_consecutive.cpp:262-287); residuals / outliers cover the flattened list of all octaves in octave order
 * (stage5_optimization.cpp:419-461, 601-610). */
int svo_get_keypoints_oct(svo_ctx* ctx, int32_t slot, int32_t octave, int32_t which, int32_t side, svo_keypoint* out, int32_t cap);
int svo_get_descriptors_oct(svo_ctx* ctx, int32_t slot, int32_t octave, int32_t which, int32_t side, uint8_t* out, int32_t cap_rows);
int svo_get_matches_oct(svo_ctx* ctx, int32_t slot, int32_t octave, int32_t which, svo_dmatch* out, int32_t cap);
int svo_get_match_ids_oct(svo_ctx* ctx, int32_t slot, int32_t octave, int32_t which, int64_t* out, int32_t cap);
int svo_get_tracked_pairs_oct(svo_ctx* ctx, int32_t slot, int32_t octave, int64_t* out_pairs /* cap x 2 */, int32_t cap);
/* out4 = {width, height, kps_to_detect[octave] (stage2_detect.cpp:404-407), keypoint capacity}; returns n_octaves */
int svo_get_octave_info(svo_ctx* ctx, int32_t octave, int32_t* out4);
/* per octave {detected left, detected right, stereo matches, tracked pairs} of the slot's last step (result.detected_feats,
 * result.stereo_matches, tracked_pairs[octave].size()); out holds cap_octaves x 4 values; returns n_octaves */
int svo_get_octave_stats(svo_ctx* ctx, int32_t slot, int32_t* out, int32_t cap_octaves);
/* parity seam of the pyramid: the image of octave >= 1 built for batch item `item` (position in the slot list) of the
 * last step, (height >> octave) x (width >> octave) bytes, tightly packed */
int svo_get_pyramid_plane(svo_ctx* ctx, int32_t item, int32_t octave, int32_t side, uint8_t* out);

/* debugging / parity seams inside stage 4: the greedy-filtered matL/matR lists (train indices)
 * and the two RANSAC inlier masks (stage4_match_consecutive.cpp:145-160, 202, 237) */
int svo_get_track_candidates(svo_ctx* ctx, int32_t slot, int32_t* out_query, int32_t* out_trainL,
                             int32_t* out_trainR, uint8_t* out_inlL, uint8_t* out_inlR, int32_t cap);
/* statistics of the two cv::findFundamentalMat calls of the slot's last step, octave 0 (stage4_match_consecutive.cpp:202,237):
 * out6 = {points, inliers, iterations executed} for the left-left and then the right-right problem */
int svo_get_ransac_stats(svo_ctx* ctx, int32_t slot, int32_t* out6);
/* survivors of stage 5's auxiliary non-maximal suppression (stage5_optimization.cpp:465-474) over the flattened tracked
 * pairs of all octaves (:419-461), as the first Gauss-Newton loop starts with them: the rows the reference writes to
 * optimization.txt (:477-492).  Returns the number of tracked pairs. */
int svo_get_stage5_survivors(svo_ctx* ctx, int32_t slot, uint8_t* out, int32_t cap);
/* the same for one octave, plus the Hamming distances of the surviving matL / matR entries (what the reference prints as
 * the fifth column of l-l_%04d.txt / r-r_%04d.txt, stage4_match_consecutive.cpp:193-195, 228-230); any pointer may be NULL */
int svo_get_track_candidates_oct(svo_ctx* ctx, int32_t slot, int32_t octave, int32_t* out_query, int32_t* out_trainL,
                                 int32_t* out_trainR, uint8_t* out_inlL, uint8_t* out_inlR, float* out_distL, float* out_distR,
                                 int32_t cap);

/* ---- match-ID / key-frame control (libstereo-odometry.h:675-726); need vo_use_matches_ids ----- */
int svo_set_this_frame_as_kf(svo_ctx* ctx, int32_t slot);         /* setThisFrameAsKF(): m_last_kf_max_id = max current ID, H:675-683 */
int svo_reset_ids(svo_ctx* ctx, int32_t slot);                    /* resetIds(): m_reset, applied at the next step (P:254-267), H:684 */
int svo_set_ids(svo_ctx* ctx, int32_t slot, const int64_t* ids, int32_t n);   /* setIds(), H:685-694 */
int svo_set_max_match_id(svo_ctx* ctx, int32_t slot, int64_t id); /* setMaxMatchID(), H:726 */

/* ---- stage 1: stage1_prepare_rectify (stage1_rectify.cpp:37-93); the pyramid (:80-83) is built inside the step -- */
/* Input format of svo_process_batch / svo_submit_batch / svo_process_pair: channels = 1 (8-bit gray, the default)
 * or 3 (8-bit BGR, as MRPT / OpenCV hold colour images; row stride in bytes).  Colour input is converted on the
 * device like mrpt::utils::CImage(img, FAST_REF_OR_CONVERT_TO_GRAY) = cv::cvtColor(BGR2GRAY)
 * (stage1_rectify.cpp:51-52). */
int svo_set_input_format(svo_ctx* ctx, int32_t channels);
/* Rectification maps of the stereo camera: what mrpt::vision::CStereoRectifyMap::setFromCamParams precomputes on
 * first use (stage1_rectify.cpp:69-70) with cv::initUndistortRectifyMap(..., CV_16SC2): map_xy = H*W int16 pairs
 * (x, y), map_a = H*W uint16 interpolation-table indices, per side.  Once set, every image given to
 * svo_process_batch / svo_submit_batch / svo_process_pair is rectified on the device like
 * CStereoRectifyMap::rectify = cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) (stage1_rectify.cpp:72-74).
 * Four NULLs: images are rectified already (obs->areImagesRectified(), stage1_rectify.cpp:61-65). */
int svo_set_rectify_maps(svo_ctx* ctx, const int16_t* map_xy_left, const uint16_t* map_a_left,
                         const int16_t* map_xy_right, const uint16_t* map_a_right);
/* stage 1 alone on one host stereo pair: out_left / out_right receive height*width gray (rectified) bytes */
int svo_prepare_rectify(svo_ctx* ctx, const uint8_t* left, const uint8_t* right, size_t stride_bytes, int32_t channels,
                        uint8_t* out_left, uint8_t* out_right);
/* device-resident variant for n pairs (source image i at d_left + i*image_pitch); the outputs are planes with a
 * row pitch of width rounded up to 16 and height*pitch bytes per image: the layout svo_process_batch_device takes
 * with stride = pitch, image_pitch = height*pitch.  Only enqueues work on the context's stream. */
int svo_prepare_rectify_device(svo_ctx* ctx, int32_t n, const uint8_t* d_left, const uint8_t* d_right,
                               size_t stride_bytes, size_t image_pitch, int32_t channels,
                               uint8_t* d_out_left, uint8_t* d_out_right);

/* the image pyramid alone (CImagePyramid::buildPyramidFast, stage1_rectify.cpp:80-83) for n device-resident gray pairs
 * (layout as svo_process_batch_device): octaves 1 .. n_octaves-1 into the context's planes (svo_get_pyramid_plane reads
 * them).  Only enqueues work on the context's stream; needs n_octaves > 1. */
int svo_build_pyramid_device(svo_ctx* ctx, int32_t n, const uint8_t* d_left, const uint8_t* d_right, size_t stride_bytes,
                             size_t image_pitch_bytes);

/* ---- per-stage entries (the internal seams, usable on their own) ---------------------------- */
/* FAST + ORB::compute + m_non_max_sup + m_update_indexes for one host image
 * (stage2_detect.cpp:502-515, 590-597, 618).  Uses slot 0's scratch; does not touch slot state. */
int svo_detect(svo_ctx* ctx, const uint8_t* img, size_t stride_bytes,
               svo_keypoint* out_kps, uint8_t* out_desc, int32_t cap);

/* cv::BFMatcher(NORM_HAMMING,false).match(query, train)
 * (stage3_match_left_right.cpp:88-94, stage4_match_consecutive.cpp:137-142): host buffers,
 * 32-byte descriptors, out has nq entries (0 written when nq == 0 or nt == 0).  Returns count. */
int svo_match_hamming(svo_ctx* ctx, const uint8_t* query, int32_t nq, const uint8_t* train, int32_t nt,
                      svo_dmatch* out);
/* device-resident variant for the microbench (BASELINE.json config 5): `reps` independent problems,
 * problem r at d_query + r*nq*32 / d_train + r*nt*32, packed results (dist << 16 | trainIdx) to d_out. */
int svo_match_hamming_device(svo_ctx* ctx, const uint8_t* d_query, int32_t nq, const uint8_t* d_train,
                             int32_t nt, int32_t reps, uint32_t* d_out);

/* stage3 BF branch: match + 1-to-1 + epipolar/threshold/disparity filter
 * (stage3_match_left_right.cpp:83-178) on caller keypoints/descriptors.  Returns kept count. */
int svo_match_lr(svo_ctx* ctx, const svo_keypoint* kpsL, const uint8_t* descL, int32_t nL,
                 const svo_keypoint* kpsR, const uint8_t* descR, int32_t nR, svo_dmatch* out, int32_t cap);

/* stage4_track (stage4_match_consecutive.cpp:70-760; ifmDescBF :88-329 or ifmDescWin :435-738 as configured) alone on
 * caller data, one octave: the previous and the current frame as stages 2 and 3 leave them -- keypoints (integer pixel
 * coordinates, as FAST gives them), 32-byte descriptors, stereo matches.  out_pairs receives tracked_pairs[0]:
 * cap x 2 (pre_match_idx, cur_match_idx).  Returns the number of tracked pairs.  Runs on the context's scratch slot;
 * no sequence slot is touched and no match IDs are assigned. */
int svo_track(svo_ctx* ctx,
              const svo_keypoint* pre_left, const uint8_t* pre_desc_left, int32_t n_pl,
              const svo_keypoint* pre_right, const uint8_t* pre_desc_right, int32_t n_pr,
              const svo_dmatch* pre_matches, int32_t n_pm,
              const svo_keypoint* cur_left, const uint8_t* cur_desc_left, int32_t n_cl,
              const svo_keypoint* cur_right, const uint8_t* cur_desc_right, int32_t n_cr,
              const svo_dmatch* cur_matches, int32_t n_cm,
              int64_t* out_pairs, int32_t cap);

/* getChangeInPose (libstereo-odometry.h:162-172, common.cpp:355-413): stage 5 alone on caller data.
 * tracked_pairs: n_tracked x 2 (pre_match_idx, cur_match_idx).  ini_estimation: 6 doubles or NULL
 * (used only when use_custom_initial_pose != 0, which also keeps m_last_computed_pose untouched).
 * Runs on `slot` (its m_last_computed_pose is the warm start otherwise).  Returns result->valid. */
int svo_get_change_in_pose(svo_ctx* ctx, int32_t slot,
                           const int64_t* tracked_pairs, int32_t n_tracked,
                           const svo_dmatch* pre_matches, int32_t n_pre, const svo_dmatch* cur_matches, int32_t n_cur,
                           const svo_keypoint* pre_left, int32_t n_pl, const svo_keypoint* pre_right, int32_t n_pr,
                           const svo_keypoint* cur_left, int32_t n_cl, const svo_keypoint* cur_right, int32_t n_cr,
                           const svo_camera* cam, int32_t use_custom_initial_pose, const double* ini_estimation,
                           svo_result* result);

/* stage5_optimize (stage5_optimization.cpp:392-736) under the name SURVEY section 8(b) lists: same call as
 * svo_get_change_in_pose, which is the reference's public seam for stage 5 alone. */
int svo_optimize(svo_ctx* ctx, int32_t slot,
                 const int64_t* tracked_pairs, int32_t n_tracked,
                 const svo_dmatch* pre_matches, int32_t n_pre, const svo_dmatch* cur_matches, int32_t n_cur,
                 const svo_keypoint* pre_left, int32_t n_pl, const svo_keypoint* pre_right, int32_t n_pr,
                 const svo_keypoint* cur_left, int32_t n_cl, const svo_keypoint* cur_right, int32_t n_cr,
                 const svo_camera* cam, int32_t use_custom_initial_pose, const double* ini_estimation,
                 svo_result* result);

/* getProjectedCoords (libstereo-odometry.h:175-182, common.cpp:415-470): the stereo matches of the previous key frame
 * whose other_matches_tracked entry is -1 are triangulated and projected into the left and right image with
 * delta_pose = [w1 w2 w3 t1 t2 t3], the rotation-vector form of the INVERSE of change_pose the reference derives at
 * common.cpp:459-464 (MRPT pose algebra stays with the caller).  out_pixels: 4 floats (xl, yl, xr, yr) per projected
 * match, in match order.  Returns their number. */
int svo_get_projected_coords(svo_ctx* ctx, const svo_dmatch* pre_matches, int32_t n_matches,
                             const svo_keypoint* pre_left, int32_t n_pl, const svo_keypoint* pre_right, int32_t n_pr,
                             const int32_t* other_matches_tracked, const svo_camera* cam, const double* delta_pose,
                             float* out_pixels, int32_t cap);

/* cv::findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask) (stage4_match_consecutive.cpp:202,237):
 * n x 2 float32 points (host), mask n bytes.  Returns the inlier count, 0 (and an all-zero mask)
 * when no model was found or n < 7. */
int svo_fundmat_ransac(svo_ctx* ctx, const float* p1, const float* p2, int32_t n, uint8_t* mask);

/* parity seam inside the call above: the 7-point solver findFundamentalMat runs on every RANSAC sample
 * (stage4_match_consecutive.cpp:202 -> cv::findFundamentalMat -> run7Point; what cv::findFundamentalMat(7 points,
 * FM_7POINT) returns).  p1, p2: n_sets x 7 x 2 float32 (host); F: n_sets x 27 doubles (up to three row-major 3x3
 * models per sample); n_models: n_sets counts.  Returns n_sets. */
int svo_fundmat_7point(svo_ctx* ctx, const float* p1, const float* p2, int32_t n_sets, double* F, int32_t* n_models);

/* parity seam inside stage 2: the intermediate image planes of the last step for (slot, side).
 * plane 0 = the 7x7 sigma=2 blur cv::ORB::compute applies (stage2_detect.cpp:512), plane 1 = FAST score
 * after 3x3 NMS and the 31-px border filter (0 = no corner).  out: height*width bytes, tightly packed. */
int svo_get_plane(svo_ctx* ctx, int32_t slot, int32_t side, int32_t plane, uint8_t* out);

/* parity seams of the dmORB detector (svo_config.detect_method = 0, stage2_detect.cpp:458-496):
 * svo_get_orb_raw_keypoints: the keypoint vector cv::ORB::detectAndCompute returned for (slot, side) in the last step, i.e.
 *   BEFORE the reference's m_non_max_sup / m_update_indexes (cv::KeyPoint layout: pt, size, angle, response, octave).
 * svo_get_orb_level: size of pyramid level `level` (w, h; out may be NULL) and one of its planes: plane 0 = the level image
 *   (cv::resize INTER_LINEAR_EXACT from the level below; levels >= 1, kept per batch item: meaningful after svo_detect or a
 *   batch whose item 0 is this slot), plane 1 = its 7x7 sigma-2 blur.  out: h*w bytes, tightly packed. */
int svo_get_orb_raw_keypoints(svo_ctx* ctx, int32_t slot, int32_t side, svo_keypoint* out, int32_t cap);
int svo_get_orb_level(svo_ctx* ctx, int32_t slot, int32_t side, int32_t level, int32_t plane, uint8_t* out, int32_t* w, int32_t* h);

/* ---- introspection -------------------------------------------------------------------------- */
int svo_abi_version(void);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
int64_t svo_kernel_launches(const svo_ctx* ctx);
/* raw CUDA stream of the context (cudaStream_t as void*), for event timing on the launching stream */
void* svo_stream(svo_ctx* ctx);
/* Concurrency inside one batched step (no counterpart in the reference, which is single-threaded): batches of
 * >= 64 slots are split into `chunks` groups (1..8, default 2) whose kernel sequences run on separate streams,
 * forked from and joined into the context's stream.  Results do not depend on the setting. */
int svo_set_concurrency(svo_ctx* ctx, int32_t chunks);
/* per-stage device times of the last svo_process_batch* call, in ms, measured with CUDA events on
 * the context's stream when enabled with svo_enable_stage_timing(ctx, 1).  names: static strings. */
int svo_enable_stage_timing(svo_ctx* ctx, int32_t on);
/* register-only probes of the POPC and LOP3 issue rates (lane operations per second) on this device: the measured
 * denominators of the Hamming stage's integer-pipe roofline (bench.py) */
int svo_measure_pipe_peaks(svo_ctx* ctx, double* popc_per_s, double* lop3_per_s);
int svo_get_stage_times(svo_ctx* ctx, const char** names, float* ms, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SVO_B200_H */
