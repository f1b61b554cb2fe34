// ref_glue.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own class, compiled together with the
// reference's UNMODIFIED sources (oracle/Makefile.ref) into oracle/_ref/libsvo_ref.so; driven by oracle/ref.py.
// The class' private members are reached with the usual test-glue trick (this translation unit only); the members
// that gui_thread.cpp would define (GUI parameter / info constructors, the GUI thread) and the two SAD helpers of
// compute_SAD8.cpp / tracking_SAD.cpp are defined here, because those files need the MRPT GUI / SSE headers and are
// off the hot path.
#include <unistd.h>
#include "svo_ref_stubs.h"      // first: the standard headers must be seen before the access trick below
#define private public
#define protected public
#include <libstereo-odometry.h>
#undef private
#undef protected
#include "internal_libstereo-odometry.h"

using namespace rso;

svo_ref_backend g_svo_ref_backend = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

// ---- members defined in gui_thread.cpp in the reference (gui_thread.cpp:34-50) ---------------------------------
CStereoOdometryEstimator::TGUIParams::TGUIParams() : show_gui(true), draw_all_raw_feats(false), draw_lr_pairings(false), draw_tracking(true) {}
CStereoOdometryEstimator::TInfoForTheGUI::TInfoForTheGUI() : timestamp(INVALID_TIMESTAMP), img_left(UNINITIALIZED_IMAGE), img_right(UNINITIALIZED_IMAGE) {}
void CStereoOdometryEstimator::thread_gui() { svo_ref_off_path("GUI thread"); }
bool CStereoOdometryEstimator::keyPressedOnGUI() { return false; }
int CStereoOdometryEstimator::getKeyPressedOnGUI() { return 0; }
namespace rso {
uint32_t compute_SAD8(const uint8_t*, const uint8_t*, const size_t, const mrpt::utils::TPixelCoord&, const mrpt::utils::TPixelCoord&) { svo_ref_off_path("compute_SAD8 (smSAD / ifmSAD)"); }
}

namespace {
std::string g_err;
struct Handle {
    CStereoOdometryEstimator est;
    CStereoOdometryEstimator::TStereoOdometryResult res;
};
mrpt::utils::TStereoCamera make_cam(const double* c, int w, int h) {
    mrpt::utils::TStereoCamera cam;
    cam.leftCamera.setIntrinsicParamsFromValues(c[0], c[1], c[2], c[3]);
    cam.rightCamera.setIntrinsicParamsFromValues(c[4], c[5], c[6], c[7]);
    cam.leftCamera.ncols = cam.rightCamera.ncols = w;
    cam.leftCamera.nrows = cam.rightCamera.nrows = h;
    cam.rightCameraPose[0] = c[8];
    return cam;
}
void fill_result(Handle* H, double* out /* 40 */) {
    const auto& r = H->res;
    out[0] = r.error_code; out[1] = r.valid; out[2] = r.num_it; out[3] = r.num_it_final;
    out[4] = (double)r.tracked_feats_from_last_frame; out[5] = (double)r.tracked_feats_from_last_KF;
    out[6] = (double)r.outliers.size(); out[7] = (double)r.out_residual.size();
    for (int i = 0; i < 9; ++i) out[8 + i] = r.outPose.R[i];
    for (int i = 0; i < 3; ++i) out[17 + i] = r.outPose.t[i];
    out[20] = r.outPose.yaw(); out[21] = r.outPose.pitch(); out[22] = r.outPose.roll();
    for (int i = 0; i < 6; ++i) out[23 + i] = H->est.m_last_computed_pose[i];
    out[29] = H->est.m_error; out[30] = H->est.m_current_fast_th; out[31] = H->est.m_current_orb_th;
    out[32] = (double)H->est.m_last_match_ID; out[33] = (double)H->est.m_last_kf_max_id; out[34] = H->est.m_it_counter;
    out[35] = (double)r.detected_feats.size();
}
const CStereoOdometryEstimator::TImagePairData* frame(Handle* H, int which) {
    const auto& p = which == 0 ? H->est.m_current_imgpair : H->est.m_prev_imgpair;
    return p.present() ? p.pointer() : nullptr;
}
}  // namespace

extern "C" {

void svo_ref_set_backend(const svo_ref_backend* b) { g_svo_ref_backend = *b; }
const char* svo_ref_last_error() { return g_err.c_str(); }
void* svo_ref_create() {
    Handle* H = new Handle();
    H->est.params_gui.show_gui = false;           // no GUI thread, no window (process_new_image_pair.cpp:357)
    H->est.params_gui.draw_tracking = false;
    H->est.setVerbosityLevel(0);
    return H;
}
void svo_ref_destroy(void* h) { delete (Handle*)h; }

// parameters by name: the reference's own field names
int svo_ref_set_param(void* h, const char* name, double v) {
    auto& e = ((Handle*)h)->est;
    const std::string n(name);
#define P(field, expr) if (n == field) { expr; return 0; }
    P("nOctaves", e.params_rectify.nOctaves = (int)v)
    P("detect_method", e.params_detect.detect_method = (CStereoOdometryEstimator::TDetectParams::TDMethod)(int)v)
    P("min_distance", e.params_detect.min_distance = (size_t)v)
    P("orb_nfeats", e.params_detect.orb_nfeats = (size_t)v)
    P("orb_nlevels", e.params_detect.orb_nlevels = (size_t)v)
    P("non_maximal_suppression", e.params_detect.non_maximal_suppression = v != 0)
    P("nmsMethod", e.params_detect.nmsMethod = (CStereoOdometryEstimator::TDetectParams::NMSMethod)(int)v)
    P("initial_FAST_threshold", e.params_detect.initial_FAST_threshold = (int)v)
    P("fast_min_th", e.params_detect.fast_min_th = (int)v)
    P("fast_max_th", e.params_detect.fast_max_th = (int)v)
    P("minimum_ORB_response", e.params_detect.minimum_ORB_response = v)
    P("match_method", e.params_lr_match.match_method = (CStereoOdometryEstimator::TLeftRightMatchParams::TSMMethod)(int)v)
    P("max_y_diff", e.params_lr_match.max_y_diff = v)
    P("enable_robust_1to1_match", e.params_lr_match.enable_robust_1to1_match = v != 0)
    P("orb_max_distance", e.params_lr_match.orb_max_distance = v)
    P("orb_min_th", e.params_lr_match.orb_min_th = (int)v)
    P("orb_max_th", e.params_lr_match.orb_max_th = (int)v)
    P("ifm_method", e.params_if_match.ifm_method = (CStereoOdometryEstimator::TInterFrameMatchingParams::TIFMMethod)(int)v)
    P("ifm_win_w", e.params_if_match.ifm_win_w = (int)v)
    P("ifm_win_h", e.params_if_match.ifm_win_h = (int)v)
    P("use_robust_kernel", e.params_least_squares.use_robust_kernel = v != 0)
    P("kernel_param", e.params_least_squares.kernel_param = v)
    P("max_iters", e.params_least_squares.max_iters = (size_t)v)
    P("initial_max_iters", e.params_least_squares.initial_max_iters = (size_t)v)
    P("min_mod_out_vector", e.params_least_squares.min_mod_out_vector = v)
    P("max_incr_cost", e.params_least_squares.max_incr_cost = (size_t)v)
    P("residual_threshold", e.params_least_squares.residual_threshold = v)
    P("bad_tracking_th", e.params_least_squares.bad_tracking_th = (size_t)v)
    P("use_previous_pose_as_initial", e.params_least_squares.use_previous_pose_as_initial = v != 0)
    P("use_custom_initial_pose", e.params_least_squares.use_custom_initial_pose = v != 0)
    P("vo_use_matches_ids", e.params_general.vo_use_matches_ids = v != 0)
    P("vo_save_files", e.params_general.vo_save_files = v != 0)
    P("vo_debug", e.params_general.vo_debug = v != 0)
    P("verbosity", e.setVerbosityLevel((int)v))
#undef P
    g_err = "unknown parameter " + n;
    return -1;
}
int svo_ref_set_out_dir(void* h, const char* dir) { ((Handle*)h)->est.params_general.vo_out_dir = dir; return 0; }
// libstereo-odometry.h:554-672 with the sections of SURVEY Appendix C
int svo_ref_load_ini(void* h, const char* file, const char* const* sections7) {
    try {
        std::vector<std::string> s(sections7, sections7 + 7);
        ((Handle*)h)->est.loadParamsFromConfigFileName(file, s);
        ((Handle*)h)->est.params_gui.show_gui = false;
        ((Handle*)h)->est.params_gui.draw_tracking = false;
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// libstereo-odometry.h:529-541, 675-726
void svo_ref_reset_thresholds(void* h) { ((Handle*)h)->est.resetFASTThreshold(); ((Handle*)h)->est.resetORBThreshold(); }
void svo_ref_set_fast_threshold(void* h, int v) { ((Handle*)h)->est.setFASTThreshold(v); }
void svo_ref_set_orb_threshold(void* h, int v) { ((Handle*)h)->est.setORBThreshold(v); }
int svo_ref_get_fast_threshold(void* h) { return ((Handle*)h)->est.getFASTThreshold(); }
int svo_ref_get_orb_threshold(void* h) { return ((Handle*)h)->est.getORBThreshold(); }
int svo_ref_set_this_frame_as_kf(void* h) { try { ((Handle*)h)->est.setThisFrameAsKF(); return 0; } catch (std::exception& e) { g_err = e.what(); return -1; } }
void svo_ref_reset_ids(void* h) { ((Handle*)h)->est.resetIds(); }
void svo_ref_set_max_match_id(void* h, int64_t id) { ((Handle*)h)->est.setMaxMatchID((size_t)id); }
void svo_ref_set_ids(void* h, const int64_t* ids, int n) { std::vector<size_t> v(ids, ids + n); ((Handle*)h)->est.setIds(v); }

// processNewImagePair (libstereo-odometry.h:157-159) on two rectified 8-bit gray images.  cam: fxL fyL cxL cyL fxR fyR
// cxR cyR baseline.  out: 40 doubles (fill_result).  Returns 0, or -1 when the reference threw (svo_ref_last_error()).
int svo_ref_process(void* h, const uint8_t* L, const uint8_t* R, int w, int hgt, int stride, const double* camv, int repeat, double* out) {
    Handle* H = (Handle*)h;
    char cwd[4096];
    bool files = H->est.params_general.vo_save_files || H->est.params_general.vo_debug;
    // optimization.txt goes to the working directory (S5:479); the row-by-row matcher ALWAYS writes its debug file
    // dist.txt there (S3:243): keep both out of the caller's directory
    const bool rbr = H->est.params_lr_match.match_method == CStereoOdometryEstimator::TLeftRightMatchParams::smDescRbR;
    if (files || rbr) {
        if (!getcwd(cwd, sizeof cwd)) return -1;
        const char* dir = files ? H->est.params_general.vo_out_dir.c_str() : "/tmp";
        if (chdir(dir) != 0) { g_err = "cannot enter the output directory"; return -1; }
        files = true;
    }
    int rc = 0;
    try {
        CStereoOdometryEstimator::TStereoOdometryRequest req;
        mrpt::obs::CObservationStereoImages* o = new mrpt::obs::CObservationStereoImages();
        o->imageLeft.loadFromMemoryBuffer(w, hgt, L, stride);
        o->imageRight.loadFromMemoryBuffer(w, hgt, R, stride);
        o->timestamp = 1000 + H->est.m_it_counter;
        req.stereo_imgs = CObservationStereoImagesPtr(o);
        req.stereo_cam = make_cam(camv, w, hgt);
        req.repeat = repeat != 0;
        H->res = CStereoOdometryEstimator::TStereoOdometryResult();
        const std::string keep = H->est.params_general.vo_out_dir;
        if (files) H->est.params_general.vo_out_dir = ".";
        try { H->est.processNewImagePair(req, H->res); } catch (...) { H->est.params_general.vo_out_dir = keep; throw; }
        H->est.params_general.vo_out_dir = keep;
        fill_result(H, out);
    } catch (std::exception& e) { g_err = e.what(); rc = -1; }
    if (files && chdir(cwd) != 0) rc = -1;
    return rc;
}

// getChangeInPose (libstereo-odometry.h:162-172, common.cpp:355-413)
int svo_ref_get_change_in_pose(void* h, const int64_t* tracked, int n_tracked, const svo_ref_dm* pre_m, int n_pm, const svo_ref_dm* cur_m, int n_cm,
                               const svo_ref_kp* pl, int n_pl, const svo_ref_kp* pr, int n_pr, const svo_ref_kp* cl, int n_cl,
                               const svo_ref_kp* cr, int n_cr, const double* camv, int w, int hgt, const double* ini6, double* out) {
    Handle* H = (Handle*)h;
    try {
        vector_index_pairs_t tp(n_tracked);
        for (int i = 0; i < n_tracked; ++i) tp[i] = std::make_pair((size_t)tracked[2 * i], (size_t)tracked[2 * i + 1]);
        auto dm = [](const svo_ref_dm* p, int n) { std::vector<cv::DMatch> v(n); if (n) std::memcpy(v.data(), p, (size_t)n * sizeof(cv::DMatch)); return v; };
        auto kp = [](const svo_ref_kp* p, int n) { std::vector<cv::KeyPoint> v(n); if (n) std::memcpy(v.data(), p, (size_t)n * sizeof(cv::KeyPoint)); return v; };
        H->res = CStereoOdometryEstimator::TStereoOdometryResult();
        std::vector<double> ini(ini6, ini6 + 6);
        H->est.getChangeInPose(tp, dm(pre_m, n_pm), dm(cur_m, n_cm), kp(pl, n_pl), kp(pr, n_pr), kp(cl, n_cl), kp(cr, n_cr), make_cam(camv, w, hgt), H->res, ini);
        fill_result(H, out);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// getProjectedCoords (common.cpp:415-470); change_pose as rotation matrix (row-major 9) + translation
int svo_ref_get_projected_coords(void* h, const svo_ref_dm* pre_m, int n_pm, const svo_ref_kp* pl, int n_pl, const svo_ref_kp* pr, int n_pr,
                                 const int32_t* other_tracked, const double* camv, int w, int hgt, const double* R9, const double* t3,
                                 float* out4, int cap) {
    Handle* H = (Handle*)h;
    try {
        std::vector<cv::DMatch> m(n_pm); if (n_pm) std::memcpy(m.data(), pre_m, (size_t)n_pm * sizeof(cv::DMatch));
        std::vector<cv::KeyPoint> a(n_pl), b(n_pr);
        if (n_pl) std::memcpy(a.data(), pl, (size_t)n_pl * sizeof(cv::KeyPoint));
        if (n_pr) std::memcpy(b.data(), pr, (size_t)n_pr * sizeof(cv::KeyPoint));
        std::vector<std::pair<int, float>> ot(n_pm);
        for (int i = 0; i < n_pm; ++i) ot[i] = std::make_pair((int)other_tracked[i], 0.f);
        CPose3D pose;
        std::memcpy(pose.R, R9, sizeof pose.R);
        std::memcpy(pose.t, t3, sizeof pose.t);
        std::vector<std::pair<TPixelCoordf, TPixelCoordf>> px;
        H->est.getProjectedCoords(m, a, b, ot, make_cam(camv, w, hgt), pose, px);
        const int n = (int)px.size();
        for (int i = 0; i < n && i < cap; ++i) { out4[4 * i] = px[i].first.x; out4[4 * i + 1] = px[i].first.y; out4[4 * i + 2] = px[i].second.x; out4[4 * i + 3] = px[i].second.y; }
        return n;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// ---- readers of the frame data the class keeps (which: 0 = m_current_imgpair, 1 = m_prev_imgpair) -------------
int svo_ref_n_octaves(void* h, int which) { const auto* f = frame((Handle*)h, which); return f ? (int)f->left.pyr_feats_kps.size() : 0; }
int svo_ref_get_keypoints(void* h, int which, int side, int oct, svo_ref_kp* out, int cap) {
    const auto* f = frame((Handle*)h, which);
    if (!f) return -1;
    const auto& v = (side ? f->right : f->left).pyr_feats_kps[oct];
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) std::memcpy(out + i, &v[i], sizeof(svo_ref_kp));
    return (int)v.size();
}
int svo_ref_get_descriptors(void* h, int which, int side, int oct, uint8_t* out, int cap_rows) {
    const auto* f = frame((Handle*)h, which);
    if (!f) return -1;
    const cv::Mat& d = (side ? f->right : f->left).pyr_feats_desc[oct];
    for (int r = 0; r < d.rows && r < cap_rows; ++r) std::memcpy(out + (size_t)r * 32, d.ptr(r), 32);
    return d.rows;
}
int svo_ref_get_feats_index(void* h, int which, int side, int oct, int64_t* out, int cap) {
    const auto* f = frame((Handle*)h, which);
    if (!f) return -1;
    const auto& v = (side ? f->right : f->left).pyr_feats_index[oct];
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int64_t)v[i];
    return (int)v.size();
}
int svo_ref_get_matches(void* h, int which, int oct, svo_ref_dm* out, int cap) {
    const auto* f = frame((Handle*)h, which);
    if (!f) return -1;
    const auto& v = f->lr_pairing_data[oct].matches_lr_dm;
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) std::memcpy(out + i, &v[i], sizeof(svo_ref_dm));
    return (int)v.size();
}
int svo_ref_get_ids(void* h, int which, int oct, int64_t* out, int cap) {
    const auto* f = frame((Handle*)h, which);
    if (!f) return -1;
    const auto& v = f->lr_pairing_data[oct].matches_IDs;
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int64_t)v[i];
    return (int)v.size();
}
int svo_ref_get_row_index(void* h, int which, int oct, int64_t* out, int cap) {
    const auto* f = frame((Handle*)h, which);
    if (!f) return -1;
    const auto& v = f->lr_pairing_data[oct].matches_lr_row_index;
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int64_t)v[i];
    return (int)v.size();
}
int svo_ref_get_pyramid(void* h, int which, int side, int oct, uint8_t* out, int* w, int* hgt) {
    const auto* f = frame((Handle*)h, which);
    if (!f || oct >= (int)(side ? f->right : f->left).pyr.images.size()) return -1;
    const auto& im = (side ? f->right : f->left).pyr.images[oct];
    *w = (int)im.getWidth(); *hgt = (int)im.getHeight();
    if (out) std::memcpy(out, im.get_unsafe(0, 0), (size_t)*w * *hgt);
    return 0;
}
// tracked_pairs of the last call (a local of processNewImagePair, process_new_image_pair.cpp:310): stage4_track is run
// again on the retained (previous, current) frames -- it is deterministic -- and its side effects (match IDs, ID
// generator, counters; stage4_match_consecutive.cpp:268-305,743-752) are put back.  out: (first, second) per pair.
int svo_ref_get_tracked(void* h, int oct, int64_t* out, int cap) {
    Handle* H = (Handle*)h;
    auto& e = H->est;
    if (!e.m_prev_imgpair.present() || !e.m_current_imgpair.present()) return 0;
    auto& cur = *e.m_current_imgpair;
    std::vector<std::vector<size_t>> ids;
    for (auto& d : cur.lr_pairing_data) ids.push_back(d.matches_IDs);
    const size_t last_id = e.m_last_match_ID, nf = e.m_num_tracked_pairs_from_last_frame, nk = e.m_num_tracked_pairs_from_last_kf;
    const bool sf = e.params_general.vo_save_files;
    e.params_general.vo_save_files = false;
    int n = -1;
    try {
        CStereoOdometryEstimator::TTrackingData td;
        e.stage4_track(td, *e.m_prev_imgpair, cur);
        if (oct < (int)td.tracked_pairs.size()) {
            const auto& v = td.tracked_pairs[oct];
            n = (int)v.size();
            for (int i = 0; i < n && i < cap; ++i) { out[2 * i] = (int64_t)v[i].first; out[2 * i + 1] = (int64_t)v[i].second; }
        }
    } catch (std::exception& ex) { g_err = ex.what(); }
    e.params_general.vo_save_files = sf;
    for (size_t i = 0; i < ids.size(); ++i) cur.lr_pairing_data[i].matches_IDs = ids[i];
    e.m_last_match_ID = last_id; e.m_num_tracked_pairs_from_last_frame = nf; e.m_num_tracked_pairs_from_last_kf = nk;
    return n;
}

// result vectors of the last call
int svo_ref_get_outliers(void* h, int64_t* out, int cap) { const auto& v = ((Handle*)h)->res.outliers; for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int64_t)v[i]; return (int)v.size(); }
int svo_ref_get_residuals(void* h, double* out, int cap) { const auto& v = ((Handle*)h)->res.out_residual; for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i]; return (int)v.size(); }
int svo_ref_get_counts(void* h, int64_t* out /* 3 per octave: detected L, detected R, stereo matches */, int cap_oct) {
    const auto& r = ((Handle*)h)->res;
    const int n = (int)r.detected_feats.size();
    for (int i = 0; i < n && i < cap_oct; ++i) { out[3 * i] = (int64_t)r.detected_feats[i].first; out[3 * i + 1] = (int64_t)r.detected_feats[i].second; out[3 * i + 2] = i < (int)r.stereo_matches.size() ? (int64_t)r.stereo_matches[i] : -1; }
    return n;
}

// ---- private members called directly: the seams the restated oracle functions are pinned on --------------------
// m_non_max_sup, 8-argument form (stage2_detect.cpp:296-370)
int svo_ref_non_max_sup(void* h, int num_out, const svo_ref_kp* kps, int n, const uint8_t* desc, int H, int W, svo_ref_kp* out_kp, uint8_t* out_desc) {
    auto& e = ((Handle*)h)->est;
    std::vector<cv::KeyPoint> in(n), out;
    if (n) std::memcpy(in.data(), kps, (size_t)n * sizeof(cv::KeyPoint));
    cv::Mat d, od;
    if (desc && n) { d.create(n, 32, CV_8UC1); std::memcpy(d.ptr(), desc, (size_t)n * 32); }
    std::vector<bool> dummy;
    e.m_non_max_sup((size_t)num_out, in, d, out, od, (size_t)H, (size_t)W, dummy);
    for (size_t i = 0; i < out.size(); ++i) std::memcpy(out_kp + i, &out[i], sizeof(svo_ref_kp));
    for (int r = 0; r < od.rows; ++r) std::memcpy(out_desc + (size_t)r * 32, od.ptr(r), 32);
    return (int)out.size();
}
// m_non_max_sup, auxiliary 5-argument form (stage2_detect.cpp:225-283)
int svo_ref_non_max_sup_aux(void* h, const svo_ref_kp* kps, int n, int H, int W, int num_out, uint8_t* survivors) {
    auto& e = ((Handle*)h)->est;
    std::vector<cv::KeyPoint> in(n);
    if (n) std::memcpy(in.data(), kps, (size_t)n * sizeof(cv::KeyPoint));
    std::vector<bool> s(n);
    e.m_non_max_sup(in, s, (size_t)H, (size_t)W, (size_t)num_out);
    for (int i = 0; i < n; ++i) survivors[i] = s[i];
    return n;
}
// m_update_indexes (stage2_detect.cpp:65-130): sorts in place, returns pyr_feats_index
int svo_ref_update_indexes(void* h, svo_ref_kp* kps, int n, uint8_t* desc, int H, int W, int64_t* row_index) {
    auto& e = ((Handle*)h)->est;
    CStereoOdometryEstimator::TImagePairData::img_data_t d;
    d.pyr.images.resize(1);
    std::vector<uint8_t> blank((size_t)W * H, 0);
    d.pyr.images[0].loadFromMemoryBuffer(W, H, blank.data(), W);
    d.pyr_feats_kps.resize(1); d.pyr_feats_desc.resize(1); d.pyr_feats_index.resize(1);
    d.pyr_feats_kps[0].resize(n);
    if (n) std::memcpy(d.pyr_feats_kps[0].data(), kps, (size_t)n * sizeof(cv::KeyPoint));
    d.pyr_feats_desc[0].create(n, 32, CV_8UC1);
    if (n) std::memcpy(d.pyr_feats_desc[0].ptr(), desc, (size_t)n * 32);
    e.m_update_indexes(d, 0, true);
    for (int i = 0; i < n; ++i) std::memcpy(kps + i, &d.pyr_feats_kps[0][i], sizeof(svo_ref_kp));
    for (int r = 0; r < n; ++r) std::memcpy(desc + (size_t)r * 32, d.pyr_feats_desc[0].ptr(r), 32);
    for (int y = 0; y < H; ++y) row_index[y] = (int64_t)d.pyr_feats_index[0][y];
    return n;
}
// m_pinhole_stereo_projection (stage5_optimization.cpp:35-257): pixels (4 floats) and the 4x6 Jacobians per landmark
int svo_ref_pinhole_projection(void* h, const double* lmks, int n, const double* camv, const double* pose6, float* px4, double* jac24) {
    auto& e = ((Handle*)h)->est;
    std::vector<mrpt::math::TPoint3D> l(n);
    for (int i = 0; i < n; ++i) l[i] = mrpt::math::TPoint3D(lmks[3 * i], lmks[3 * i + 1], lmks[3 * i + 2]);
    std::vector<double> dp(pose6, pose6 + 6);
    std::vector<std::pair<TPixelCoordf, TPixelCoordf>> px;
    std::vector<Eigen::MatrixXd> J;
    e.m_pinhole_stereo_projection(l, make_cam(camv, 640, 480), dp, px, J);
    for (int i = 0; i < n; ++i) {
        px4[4 * i] = px[i].first.x; px4[4 * i + 1] = px[i].first.y; px4[4 * i + 2] = px[i].second.x; px4[4 * i + 3] = px[i].second.y;
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 6; ++c) jac24[24 * i + 6 * r + c] = J[i](r, c);
    }
    return n;
}

}  // extern "C"
