// libstereo_odometry_b200.cpp -- host shim: the reference's class API on top of the C-ABI.
// No arithmetic of the path lives here; every stage runs in libsvo_b200.so on the GPU.
#include "libstereo_odometry_b200.h"

#include <algorithm>
#include <cstring>

namespace rso {

CStereoOdometryEstimator::CStereoOdometryEstimator() {}

CStereoOdometryEstimator::~CStereoOdometryEstimator() {
    if (m_ctx) svo_destroy(m_ctx);
}

void CStereoOdometryEstimator::throw_last(svo_ctx* c, const char* what) {
    throw std::runtime_error(std::string("[Visual Odometry Error] ") + what + ": " + svo_last_error(c));
}

// (Re)create the device context when the image size or any parameter changed.  Parameters are plain
// members in the reference too, read at every call; here they are baked into the context, so a change
// costs one re-creation (sequence state is lost, as after constructing a new estimator).
void CStereoOdometryEstimator::ensure_context(int width, int height) {
    if (m_ctx && !m_dirty && width == m_w && height == m_h) return;
    // process_new_image_pair.cpp:54-76: only the FAST+ORB / DescBF / DescBF path exists here
    if (params_detect.detect_method != TDetectParams::dmFAST_ORB)
        throw std::runtime_error("[Visual Odometry Error] Detect method not correct (this build has dmFAST_ORB only).");
    if (params_lr_match.match_method != TLeftRightMatchParams::smDescBF)
        throw std::runtime_error("[Visual Odometry Error] Stereo matching method not correct (this build has smDescBF only).");
    if (params_if_match.ifm_method != TInterFrameMatchingParams::ifmDescBF)
        throw std::runtime_error("[Visual Odometry Error] Interframe matching method not correct (this build has ifmDescBF only).");
    if (params_rectify.nOctaves != 1)
        throw std::runtime_error("[Visual Odometry Error] nOctaves must be 1 on this path.");
    if (params_detect.nmsMethod != TDetectParams::nmsmStandard)
        throw std::runtime_error("[Visual Odometry Error] Invalid non-maximal-suppression method (nmsmStandard only).");
    if (m_ctx) { svo_destroy(m_ctx); m_ctx = nullptr; }
    svo_config cfg;
    svo_default_config(&cfg);
    cfg.device = m_device;
    cfg.width = width;
    cfg.height = height;
    cfg.max_seqs = 1;
    cfg.orb_nfeats = (int32_t)params_detect.orb_nfeats;
    cfg.fast_threshold = m_current_fast_th;
    cfg.non_maximal_suppression = params_detect.non_maximal_suppression ? 1 : 0;
    cfg.min_distance = (int32_t)params_detect.min_distance;
    cfg.max_y_diff = params_lr_match.max_y_diff;
    cfg.orb_threshold = m_current_orb_th;
    cfg.enable_robust_1to1_match = params_lr_match.enable_robust_1to1_match ? 1 : 0;
    cfg.fundmat_filter = 1;
    cfg.use_robust_kernel = params_least_squares.use_robust_kernel ? 1 : 0;
    cfg.kernel_param = params_least_squares.kernel_param;
    cfg.max_iters = (int32_t)params_least_squares.max_iters;
    cfg.initial_max_iters = (int32_t)params_least_squares.initial_max_iters;
    cfg.min_mod_out_vector = params_least_squares.min_mod_out_vector;
    cfg.max_incr_cost = (int32_t)params_least_squares.max_incr_cost;
    cfg.residual_threshold = params_least_squares.residual_threshold;
    cfg.bad_tracking_th = (int32_t)params_least_squares.bad_tracking_th;
    cfg.use_previous_pose_as_initial = params_least_squares.use_previous_pose_as_initial ? 1 : 0;
    cfg.vo_use_matches_ids = params_general.vo_use_matches_ids ? 1 : 0;
    if (svo_create(&cfg, &m_ctx) != SVO_OK) throw_last(nullptr, "svo_create");
    m_w = width;
    m_h = height;
    m_dirty = false;
}

static svo_camera to_cam(const TStereoCamera& s) {
    svo_camera c;
    c.fxl = s.leftCamera.fx(); c.fyl = s.leftCamera.fy(); c.cxl = s.leftCamera.cx(); c.cyl = s.leftCamera.cy();
    c.fxr = s.rightCamera.fx(); c.fyr = s.rightCamera.fy(); c.cxr = s.rightCamera.cx(); c.cyr = s.rightCamera.cy();
    c.baseline = s.rightCameraPose[0];
    return c;
}

static void fill_result(svo_ctx* ctx, const svo_result& r, CStereoOdometryEstimator::TStereoOdometryResult& out, bool pose_valid) {
    out.num_it = r.num_it;
    out.num_it_final = r.num_it_final;
    out.valid = r.valid != 0;
    out.error_code = (VOErrorCode)r.error_code;
    out.tracked_feats_from_last_KF = (size_t)r.tracked_feats_from_last_KF;
    out.tracked_feats_from_last_frame = (size_t)r.tracked_feats_from_last_frame;
    if (pose_valid) {
        std::memcpy(out.outPose.m, r.out_pose, sizeof(r.out_pose));
        std::memcpy(out.outPose.ypr, r.out_pose_ypr, sizeof(r.out_pose_ypr));
    }
    std::vector<int64_t> tmp((size_t)std::max(r.n_outliers, 0));
    if (r.n_outliers > 0) svo_get_outliers(ctx, 0, tmp.data(), r.n_outliers);
    out.outliers.assign(tmp.begin(), tmp.end());
    out.out_residual.resize((size_t)std::max(r.n_residuals, 0));
    if (r.n_residuals > 0) svo_get_residuals(ctx, 0, out.out_residual.data(), r.n_residuals);
}

void CStereoOdometryEstimator::processNewImagePair(TStereoOdometryRequest& request_data, TStereoOdometryResult& result) {
    if (!request_data.stereo_imgs)   // process_new_image_pair.cpp:81
        throw std::runtime_error("Pointer 'request_data.stereo_imgs' must be set to stereo observation data!");
    if (request_data.use_precomputed_data)
        throw std::runtime_error("[Visual Odometry Error] use_precomputed_data is not supported (unusable in the reference too).");
    const CImage& L = request_data.stereo_imgs->imageLeft;
    const CImage& R = request_data.stereo_imgs->imageRight;
    if (!L.data || !R.data || L.width != R.width || L.height != R.height || L.stride != R.stride)
        throw std::runtime_error("[Visual Odometry Error] left/right images must be 8-bit gray with equal geometry.");
    ensure_context(L.width, L.height);
    result.error_code = voecNone;                         // P:50
    const svo_camera cam = to_cam(request_data.stereo_cam);
    svo_result r;
    const int rc = svo_process_pair(m_ctx, L.data, R.data, L.stride, &cam, request_data.repeat ? 1 : 0, &r);
    if (rc != SVO_OK) throw_last(m_ctx, "svo_process_pair");
    result.detected_feats.assign(1, std::make_pair((size_t)r.detected_left, (size_t)r.detected_right));   // P:171-176
    result.stereo_matches.assign(1, (size_t)r.stereo_matches);                                            // P:274-276
    fill_result(m_ctx, r, result, r.n_residuals > 0 && (r.valid || r.error_code == voecIncrFuncCostStg2 || r.error_code == voecIncrFuncCostStg1));
    if (!request_data.repeat) ++m_it_counter;              // P:380-381
}

bool CStereoOdometryEstimator::getChangeInPose(const vector_index_pairs_t& tracked_pairs, const std::vector<DMatch>& pre_matches,
                                               const std::vector<DMatch>& cur_matches, const std::vector<KeyPoint>& pre_left_feats,
                                               const std::vector<KeyPoint>& pre_right_feats, const std::vector<KeyPoint>& cur_left_feats,
                                               const std::vector<KeyPoint>& cur_right_feats, const TStereoCamera& stereo_camera,
                                               TStereoOdometryResult& result, const std::vector<double>& ini_estimation) {
    // common.cpp:402-403: the image size comes from the camera model
    const int w = stereo_camera.leftCamera.ncols ? (int)stereo_camera.leftCamera.ncols : m_w;
    const int h = stereo_camera.leftCamera.nrows ? (int)stereo_camera.leftCamera.nrows : m_h;
    if (w <= 0 || h <= 0) throw std::runtime_error("[Visual Odometry Error] getChangeInPose needs leftCamera.ncols/nrows.");
    ensure_context(w, h);
    std::vector<int64_t> tp(2 * tracked_pairs.size());
    for (size_t i = 0; i < tracked_pairs.size(); ++i) { tp[2 * i] = (int64_t)tracked_pairs[i].first; tp[2 * i + 1] = (int64_t)tracked_pairs[i].second; }
    const svo_camera cam = to_cam(stereo_camera);
    svo_result r;
    const bool custom = params_least_squares.use_custom_initial_pose;
    const int rc = svo_get_change_in_pose(
        m_ctx, 0, tp.data(), (int32_t)tracked_pairs.size(),
        reinterpret_cast<const svo_dmatch*>(pre_matches.data()), (int32_t)pre_matches.size(),
        reinterpret_cast<const svo_dmatch*>(cur_matches.data()), (int32_t)cur_matches.size(),
        reinterpret_cast<const svo_keypoint*>(pre_left_feats.data()), (int32_t)pre_left_feats.size(),
        reinterpret_cast<const svo_keypoint*>(pre_right_feats.data()), (int32_t)pre_right_feats.size(),
        reinterpret_cast<const svo_keypoint*>(cur_left_feats.data()), (int32_t)cur_left_feats.size(),
        reinterpret_cast<const svo_keypoint*>(cur_right_feats.data()), (int32_t)cur_right_feats.size(),
        &cam, custom ? 1 : 0, custom && ini_estimation.size() == 6 ? ini_estimation.data() : nullptr, &r);
    if (rc < 0) throw_last(m_ctx, "svo_get_change_in_pose");
    fill_result(m_ctx, r, result, r.n_residuals > 0);
    return result.valid;
}

void CStereoOdometryEstimator::getValues(std::vector<KeyPoint>& leftKP, std::vector<KeyPoint>& rightKP, Mat& leftDesc, Mat& rightDesc,
                                         std::vector<DMatch>& matches, std::vector<size_t>& matches_id) {
    if (!m_ctx) throw std::runtime_error("[Visual Odometry Error] no current frame.");
    const int cap = 2 * (int)params_detect.orb_nfeats + 16;
    auto kps = [&](int side, std::vector<KeyPoint>& out, Mat& desc) {
        out.resize(cap);
        const int n = svo_get_keypoints(m_ctx, 0, 0, side, reinterpret_cast<svo_keypoint*>(out.data()), cap);
        if (n < 0) throw_last(m_ctx, "svo_get_keypoints");
        out.resize(n);
        desc.rows = n; desc.cols = 32; desc.data.resize((size_t)n * 32);
        if (n && svo_get_descriptors(m_ctx, 0, 0, side, desc.data.data(), n) < 0) throw_last(m_ctx, "svo_get_descriptors");
    };
    kps(0, leftKP, leftDesc);
    kps(1, rightKP, rightDesc);
    matches.resize(cap);
    const int m = svo_get_matches(m_ctx, 0, 0, reinterpret_cast<svo_dmatch*>(matches.data()), cap);
    if (m < 0) throw_last(m_ctx, "svo_get_matches");
    matches.resize(m);
    std::vector<int64_t> ids(cap);
    const int ni = svo_get_match_ids(m_ctx, 0, 0, ids.data(), cap);
    matches_id.assign(ids.begin(), ids.begin() + std::max(ni, 0));
}

}  // namespace rso
