// libstereo_odometry_b200.cpp -- host shim: the reference's class API on top of the C-ABI.
// No arithmetic of the path lives here; every stage runs in libsvo_b200.so on the GPU.
#include "libstereo_odometry_b200.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <sys/stat.h>
#include <sys/types.h>

namespace rso {

// ---- CConfigFile ---------------------------------------------------------------------------------------
static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
static std::string lower(std::string s) {
    for (char& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}

void CConfigFile::parse(const std::string& text) {
    std::istringstream in(text);
    std::string line, section;
    while (std::getline(in, line)) {
        size_t cut = line.size();
        for (size_t i = 0; i < line.size(); ++i)
            if (line[i] == ';' || line[i] == '#' || (line[i] == '/' && i + 1 < line.size() && line[i + 1] == '/')) { cut = i; break; }
        line = trim(line.substr(0, cut));
        if (line.empty()) continue;
        if (line.front() == '[' && line.back() == ']') { section = lower(trim(line.substr(1, line.size() - 2))); m_data[section]; continue; }
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        m_data[section][lower(trim(line.substr(0, eq)))] = trim(line.substr(eq + 1));
    }
}

CConfigFile::CConfigFile(const std::string& fileName) {
    std::ifstream f(fileName.c_str());
    if (!f) throw std::runtime_error("CConfigFile: cannot open '" + fileName + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    parse(ss.str());
}

CConfigFile CConfigFile::fromString(const std::string& text) {
    CConfigFile c;
    c.parse(text);
    return c;
}

bool CConfigFile::sectionExists(const std::string& section) const { return m_data.count(lower(section)) != 0; }

bool CConfigFile::find(const std::string& section, const std::string& name, std::string& out, bool failIfNotFound) const {
    auto s = m_data.find(lower(section));
    if (s != m_data.end()) {
        auto k = s->second.find(lower(name));
        if (k != s->second.end()) { out = k->second; return true; }
    }
    if (failIfNotFound) throw std::runtime_error("Value '" + name + "' not found in section '" + section + "' of the configuration file");
    return false;
}

int CConfigFile::read_int(const std::string& section, const std::string& name, int def, bool fail) const {
    std::string v;
    return find(section, name, v, fail) ? (int)std::strtol(v.c_str(), nullptr, 0) : def;
}
double CConfigFile::read_double(const std::string& section, const std::string& name, double def, bool fail) const {
    std::string v;
    return find(section, name, v, fail) ? std::strtod(v.c_str(), nullptr) : def;
}
bool CConfigFile::read_bool(const std::string& section, const std::string& name, bool def, bool fail) const {
    std::string v;
    if (!find(section, name, v, fail)) return def;
    v = lower(v);
    if (v == "true" || v == "yes" || v == "on") return true;
    if (v == "false" || v == "no" || v == "off") return false;
    return std::strtol(v.c_str(), nullptr, 0) != 0;
}
std::string CConfigFile::read_string(const std::string& section, const std::string& name, const std::string& def, bool fail) const {
    std::string v;
    return find(section, name, v, fail) ? v : def;
}

// ---- loadParamsFromConfigFile (H:554-672) ------------------------------------------------------------------
void CStereoOdometryEstimator::loadParamsFromConfigFile(const CConfigFile& ini, const std::vector<std::string>& sections) {
    if (sections.size() != 7) throw std::runtime_error("Assert condition failed: sections.size() == 7");   // ASSERT_, H:556
    if (!sections[0].empty())   // rectify
        params_rectify.nOctaves = ini.read_int(sections[0], "nOctaves", params_rectify.nOctaves);
    if (!sections[1].empty()) { // detection
        TDetectParams& d = params_detect;
        d.detect_method = static_cast<TDetectParams::TDMethod>(ini.read_int(sections[1], "detect_method", d.detect_method));
        d.min_distance = (size_t)ini.read_int(sections[1], "min_distance", (int)d.min_distance);
        d.initial_FAST_threshold = ini.read_int(sections[1], "initial_FAST_threshold", d.initial_FAST_threshold);
        d.fast_min_th = ini.read_int(sections[1], "fast_min_th", d.fast_min_th);
        d.fast_max_th = ini.read_int(sections[1], "fast_max_th", d.fast_max_th);
        d.orb_nfeats = (size_t)ini.read_int(sections[1], "orb_nfeats", (int)d.orb_nfeats);
        d.orb_nlevels = (size_t)ini.read_int(sections[1], "orb_nlevels", (int)d.orb_nlevels);
        d.non_maximal_suppression = ini.read_bool(sections[1], "non_maximal_suppression", d.non_maximal_suppression);
        d.nmsMethod = static_cast<TDetectParams::NMSMethod>(ini.read_int(sections[1], "non_max_supp_method", d.nmsMethod));
    }
    if (!sections[2].empty()) { // left right matching
        TLeftRightMatchParams& m = params_lr_match;
        m.match_method = static_cast<TLeftRightMatchParams::TSMMethod>(ini.read_int(sections[2], "match_method", m.match_method));
        m.max_y_diff = ini.read_double(sections[2], "max_y_diff", m.max_y_diff);
        m.enable_robust_1to1_match = ini.read_bool(sections[2], "enable_robust_1to1_match", m.enable_robust_1to1_match);
        m.rectified_images = ini.read_bool(sections[2], "rectified_images", m.rectified_images);
        m.orb_min_th = ini.read_int(sections[2], "orb_min_th", m.orb_min_th);
        m.orb_max_th = ini.read_int(sections[2], "orb_max_th", m.orb_max_th);
        m.orb_max_distance = ini.read_double(sections[2], "orb_max_distance", m.orb_max_distance);
    }
    if (!sections[3].empty()) { // inter-frame matching (default 0 when the key is missing, H:610)
        params_if_match.ifm_method = static_cast<TInterFrameMatchingParams::TIFMMethod>(ini.read_int(sections[3], "if_match_method", 0));
        params_if_match.filter_fund_matrix = ini.read_bool(sections[3], "filter_fund_matrix", params_if_match.filter_fund_matrix);
        params_if_match.ifm_win_h = ini.read_int(sections[3], "window_height", params_if_match.ifm_win_h);   // H:614-615
        params_if_match.ifm_win_w = ini.read_int(sections[3], "window_width", params_if_match.ifm_win_w);
    }
    if (!sections[4].empty()) { // least squares
        TLeastSquaresParams& l = params_least_squares;
        l.std_noise_pixels = ini.read_double(sections[4], "std_noise_pixels", l.std_noise_pixels);
        l.use_previous_pose_as_initial = ini.read_bool(sections[4], "use_previous_pose_as_initial", l.use_previous_pose_as_initial);
        l.initial_max_iters = (size_t)ini.read_int(sections[4], "initial_max_iters", (int)l.initial_max_iters);
        l.max_iters = (size_t)ini.read_int(sections[4], "max_iters", (int)l.max_iters);
        l.min_mod_out_vector = ini.read_double(sections[4], "min_mod_out_vector", l.min_mod_out_vector);
        l.max_incr_cost = (size_t)ini.read_int(sections[4], "max_incr_cost", (int)l.max_incr_cost);
        l.residual_threshold = ini.read_double(sections[4], "residual_threshold", l.residual_threshold);
        l.bad_tracking_th = (size_t)ini.read_int(sections[4], "bad_tracking_th", (int)l.bad_tracking_th);
        l.use_robust_kernel = ini.read_bool(sections[4], "use_robust_kernel", l.use_robust_kernel);
        l.kernel_param = ini.read_double(sections[4], "kernel_param", l.kernel_param);
    }
    // sections[5] (GUI): there is no GUI in this build, its keys are ignored
    if (!sections[6].empty()) { // GENERAL
        TGeneralParams& g = params_general;
        g.vo_use_matches_ids = ini.read_bool(sections[6], "vo_use_matches_ids", g.vo_use_matches_ids);
        g.vo_save_files = ini.read_bool(sections[6], "vo_save_files", g.vo_save_files);
        g.vo_debug = ini.read_bool(sections[6], "vo_debug", g.vo_debug);
        g.vo_pause_it = ini.read_bool(sections[6], "vo_pause_it", g.vo_pause_it);
        g.vo_out_dir = ini.read_string(sections[6], "vo_out_dir", g.vo_out_dir);
    }
    resetFASTThreshold();       // H:661-662
    resetORBThreshold();
}

void CStereoOdometryEstimator::loadParamsFromConfigFileName(const std::string& fileName, const std::vector<std::string>& sections) {
    CConfigFile ini(fileName);  // ASSERT_(fileExists(fileName)), H:669
    loadParamsFromConfigFile(ini, sections);
}

// ---- match IDs / key frames (H:675-726) ---------------------------------------------------------------------
void CStereoOdometryEstimator::setThisFrameAsKF() {
    if (!m_ctx) throw std::runtime_error("[VO Error -- setThisFrameAsKF] Current frame does not exist");
    if (svo_set_this_frame_as_kf(m_ctx, m_slot) != SVO_OK) throw_last(m_ctx, "setThisFrameAsKF");
}
void CStereoOdometryEstimator::resetIds() {
    if (!m_ctx) throw std::runtime_error("[Visual Odometry Error] no frame has been processed yet.");
    if (svo_reset_ids(m_ctx, m_slot) != SVO_OK) throw_last(m_ctx, "resetIds");
}
void CStereoOdometryEstimator::setIds(const std::vector<size_t>& ids) {
    if (!m_ctx) throw std::runtime_error("[Visual Odometry Error] no frame has been processed yet.");
    std::vector<int64_t> v(ids.begin(), ids.end());
    if (svo_set_ids(m_ctx, m_slot, v.data(), (int32_t)v.size()) != SVO_OK) throw_last(m_ctx, "setIds");
}
void CStereoOdometryEstimator::setMaxMatchID(size_t id) {
    if (!m_ctx) throw std::runtime_error("[Visual Odometry Error] no frame has been processed yet.");
    if (svo_set_max_match_id(m_ctx, m_slot, (int64_t)id) != SVO_OK) throw_last(m_ctx, "setMaxMatchID");
}
std::vector<size_t> CStereoOdometryEstimator::getCopyCurrentIDs(size_t octave) {
    if (!m_ctx || octave >= (size_t)n_octaves()) return std::vector<size_t>();
    const int cap = keypoint_capacity((int)octave);
    std::vector<int64_t> ids((size_t)cap);
    const int n = svo_get_match_ids_oct(m_ctx, m_slot, (int32_t)octave, 0, ids.data(), cap);
    if (n < 0) throw_last(m_ctx, "svo_get_match_ids_oct");
    return std::vector<size_t>(ids.begin(), ids.begin() + std::min(n, cap));
}

// capacity of the per-octave keypoint / match arrays of the context: 2*orb_nfeats-derived quota with the NMS, the corner
// capacity (max_corners) without it (svo_get_octave_info)
int CStereoOdometryEstimator::keypoint_capacity(int octave) const {
    int32_t info[4] = {0, 0, 0, 0};
    if (!m_ctx || svo_get_octave_info(m_ctx, octave, info) < 1) return 0;
    return info[3];
}

// S3:425-445
std::vector<size_t> CStereoOdometryEstimator::getMatchesRowIndex(size_t octave) {
    if (!m_ctx || octave >= (size_t)n_octaves()) return std::vector<size_t>();
    const int cap = keypoint_capacity((int)octave);
    int32_t info[4];
    svo_get_octave_info(m_ctx, (int)octave, info);
    std::vector<svo_keypoint> kl((size_t)cap);
    std::vector<svo_dmatch> m((size_t)cap);
    const int nl = svo_get_keypoints_oct(m_ctx, m_slot, (int)octave, 0, 0, kl.data(), cap);
    const int nm = svo_get_matches_oct(m_ctx, m_slot, (int)octave, 0, m.data(), cap);
    if (nl < 0 || nm < 0) throw_last(m_ctx, "getMatchesRowIndex");
    const size_t imgH = (size_t)info[1], nFeats = (size_t)std::min(nm, cap);
    std::vector<size_t> idx(imgH + 1);
    size_t k = 0;
    for (size_t y = 0; y < imgH; ++y) {
        idx[y] = k;
        while (k < nFeats && kl[(size_t)m[k].queryIdx].y <= (float)(int)y) ++k;
    }
    idx[imgH] = (size_t)std::min(nl, cap);          // "[nRows] = total # of feats": the left FEATURE count (S3:443)
    return idx;
}

void CStereoOdometryEstimator::setRectifyMaps(const std::vector<int16_t>& xyL, const std::vector<uint16_t>& aL,
                                              const std::vector<int16_t>& xyR, const std::vector<uint16_t>& aR) {
    if (xyL.size() != 2 * aL.size() || xyR.size() != 2 * aR.size() || aL.size() != aR.size() || aL.empty())
        throw std::runtime_error("[Visual Odometry Error] setRectifyMaps: map_xy needs 2 and map_a 1 entry per pixel, both sides.");
    m_map_xy[0] = xyL; m_map_a[0] = aL; m_map_xy[1] = xyR; m_map_a[1] = aR;
    m_maps_on_device = false;
}

CStereoOdometryEstimator::CStereoOdometryEstimator() {}

CStereoOdometryEstimator::~CStereoOdometryEstimator() {
    if (m_ctx && !m_pool) svo_destroy(m_ctx);
    if (m_aux_ctx) svo_destroy(m_aux_ctx);
}

void CStereoOdometryEstimator::throw_last(svo_ctx* c, const char* what) {
    throw std::runtime_error(std::string("[Visual Odometry Error] ") + what + ": " + svo_last_error(c));
}

// process_new_image_pair.cpp:54-76: the methods this build has
void CStereoOdometryEstimator::check_methods() const {
    if (params_detect.detect_method != TDetectParams::dmFAST_ORB && params_detect.detect_method != TDetectParams::dmORB)
        throw std::runtime_error("[Visual Odometry Error] Detect method not correct (this build has dmFAST_ORB and dmORB).");
    if (params_lr_match.match_method != TLeftRightMatchParams::smDescBF && params_lr_match.match_method != TLeftRightMatchParams::smDescRbR)
        throw std::runtime_error("[Visual Odometry Error] Stereo matching method not correct (this build has smDescBF and smDescRbR).");
    if (params_if_match.ifm_method != TInterFrameMatchingParams::ifmDescBF && params_if_match.ifm_method != TInterFrameMatchingParams::ifmDescWin)
        throw std::runtime_error("[Visual Odometry Error] Interframe matching method not correct (this build has ifmDescBF and ifmDescWin).");
    if (params_rectify.nOctaves < 1 || params_rectify.nOctaves > 4)
        throw std::runtime_error("[Visual Odometry Error] nOctaves must be within [1, 4].");
    if (params_detect.nmsMethod != TDetectParams::nmsmStandard)
        throw std::runtime_error("[Visual Odometry Error] Invalid non-maximal-suppression method (nmsmStandard only).");
}

svo_config CStereoOdometryEstimator::make_config(int width, int height, int max_seqs) const {
    svo_config cfg;
    svo_default_config(&cfg);
    cfg.device = m_device;
    cfg.width = width;
    cfg.height = height;
    cfg.max_seqs = max_seqs;
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
    cfg.match_method = (int32_t)params_lr_match.match_method;
    cfg.if_match_method = (int32_t)params_if_match.ifm_method;
    cfg.ifm_win_w = params_if_match.ifm_win_w;
    cfg.ifm_win_h = params_if_match.ifm_win_h;
    cfg.orb_max_distance = params_lr_match.orb_max_distance;
    cfg.n_octaves = n_octaves();                  // stage1_rectify.cpp:80: dmORB builds its own pyramid and forces one octave
    cfg.detect_method = (int32_t)params_detect.detect_method;
    cfg.orb_nlevels = (int32_t)params_detect.orb_nlevels;
    cfg.pyr_mode = params_rectify.pyr_mode;
    return cfg;
}

// m_current_fast_th / m_current_orb_th to the device (H:529-541): two ints of this estimator's slot, nothing else
void CStereoOdometryEstimator::push_thresholds() {
    if (!m_ctx || (m_dirty && !m_pool)) return;     // no context yet (or about to be rebuilt): svo_create takes the values
    if (svo_set_thresholds(m_ctx, m_slot, m_current_fast_th, m_current_orb_th) != SVO_OK) throw_last(m_ctx, "svo_set_thresholds");
}

// (Re)create the device context when the image size or a params_* member changed (applyParams / loadParamsFromConfigFile).
// The params_* members are baked into the context (they size its arrays); the two dynamic thresholds are not, see
// push_thresholds().  A pooled estimator never rebuilds: its context belongs to the pool.
void CStereoOdometryEstimator::ensure_context(int width, int height) {
    if (m_pool) {
        if (width != m_w || height != m_h) throw std::runtime_error("[Visual Odometry Error] image size differs from the pool's.");
        return;
    }
    if (m_ctx && !m_dirty && width == m_w && height == m_h) return;
    check_methods();
    if (m_ctx) { svo_destroy(m_ctx); m_ctx = nullptr; }
    const svo_config cfg = make_config(width, height, 1);
    if (svo_create(&cfg, &m_ctx) != SVO_OK) throw_last(nullptr, "svo_create");
    m_w = width;
    m_h = height;
    m_dirty = false;
    m_in_channels = 1;          // a fresh context takes gray, rectified images
    m_maps_on_device = false;
    m_rectify_active = false;
}

// getChangeInPose / getProjectedCoords take the image size from the camera model (common.cpp:402-403).  They run on
// this estimator's own slot when that size is the running one (or nothing runs yet), and on a private scratch context
// otherwise -- a live sequence is never torn down by them.
svo_ctx* CStereoOdometryEstimator::context_for_size(int width, int height, int* slot) {
    if (m_ctx && !(m_dirty && !m_pool) && width == m_w && height == m_h) { *slot = m_slot; return m_ctx; }
    if (!m_ctx && !m_pool) { ensure_context(width, height); *slot = m_slot; return m_ctx; }
    if (!m_aux_ctx || m_aux_w != width || m_aux_h != height) {
        check_methods();
        if (m_aux_ctx) { svo_destroy(m_aux_ctx); m_aux_ctx = nullptr; }
        const svo_config cfg = make_config(width, height, 1);
        if (svo_create(&cfg, &m_aux_ctx) != SVO_OK) throw_last(nullptr, "svo_create");
        m_aux_w = width; m_aux_h = height;
    }
    *slot = 0;
    return m_aux_ctx;
}

static svo_camera to_cam(const TStereoCamera& s) {
    svo_camera c;
    c.fxl = s.leftCamera.fx(); c.fyl = s.leftCamera.fy(); c.cxl = s.leftCamera.cx(); c.cyl = s.leftCamera.cy();
    c.fxr = s.rightCamera.fx(); c.fyr = s.rightCamera.fy(); c.cxr = s.rightCamera.cx(); c.cyr = s.rightCamera.cy();
    c.baseline = s.rightCameraPose[0];
    return c;
}

static void fill_result(svo_ctx* ctx, int slot, const svo_result& r, const double* residuals, const int64_t* outliers,
                        CStereoOdometryEstimator::TStereoOdometryResult& out, bool pose_valid) {
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
    const int no = std::max(r.n_outliers, 0), nr = std::max(r.n_residuals, 0);
    if (outliers) out.outliers.assign(outliers, outliers + no);
    else {
        std::vector<int64_t> tmp((size_t)no);
        if (no > 0) svo_get_outliers(ctx, slot, tmp.data(), no);
        out.outliers.assign(tmp.begin(), tmp.end());
    }
    out.out_residual.resize((size_t)nr);
    if (residuals) std::copy(residuals, residuals + nr, out.out_residual.begin());
    else if (nr > 0) svo_get_residuals(ctx, slot, out.out_residual.data(), nr);
}

// ---- params_general.vo_save_files: the per-frame text dumps of the reference ---------------------------------
// Written from the device state after the step, file names and line formats as in process_new_image_pair.cpp:179-203
// (left/right_feats_%04d.txt, left/right_desc_o%02d_%04d.yml in cv::FileStorage's YAML layout), :278-287
// (matches_%04d.txt), stage4_match_consecutive.cpp:81-82,308-328 (inter_frame_matching_%04d.txt) and
// stage5_optimization.cpp:703-713 (out_residual_%04d.txt and out_residual_final_%04d.txt: the reference writes the
// final residual vector to both), stage4_match_consecutive.cpp:173-234 (l-l_%04d.txt / r-r_%04d.txt: the greedy-filtered
// left-left / right-right lists with their Hamming distances; the reference re-opens the same name for every octave, so
// the LAST octave's list is what stays on disk) and stage5_optimization.cpp:477-492 (optimization.txt in the working
// directory: the eight coordinates of every survivor of stage 5's auxiliary non-maximal suppression).
static FILE* open_dump(const std::string& dir, const char* fmt, int a, int b = -1) {
    char name[128];
    if (b < 0) std::snprintf(name, sizeof(name), fmt, a);
    else std::snprintf(name, sizeof(name), fmt, a, b);
    FILE* f = std::fopen((dir + "/" + name).c_str(), "wt");
    if (!f) throw std::runtime_error("[Visual Odometry Error] cannot write '" + dir + "/" + name + "'");
    return f;
}

void CStereoOdometryEstimator::save_files(const TStereoOdometryResult& result, bool have_prev) {
    const std::string& dir = params_general.vo_out_dir;
    const int it = (int)m_it_counter;
    const int n_oct = (int)result.detected_feats.size();
    struct Oct { std::vector<KeyPoint> kp[2][2]; std::vector<DMatch> m[2]; std::vector<int64_t> tp; };   // [which][side], [which]
    std::vector<Oct> oc((size_t)n_oct);
    auto kps = [&](int octave, int which, int side, std::vector<KeyPoint>& out) {
        const int cap = keypoint_capacity(octave);
        out.resize((size_t)cap);
        const int n = svo_get_keypoints_oct(m_ctx, m_slot, octave, which, side, reinterpret_cast<svo_keypoint*>(out.data()), cap);
        if (n < 0) throw_last(m_ctx, "svo_get_keypoints_oct");
        out.resize((size_t)std::min(n, cap));
    };
    auto matches = [&](int octave, int which, std::vector<DMatch>& out) {
        const int cap = keypoint_capacity(octave);
        out.resize((size_t)cap);
        const int n = svo_get_matches_oct(m_ctx, m_slot, octave, which, reinterpret_cast<svo_dmatch*>(out.data()), cap);
        if (n < 0) throw_last(m_ctx, "svo_get_matches_oct");
        out.resize((size_t)std::min(n, cap));
    };
    for (int k = 0; k < n_oct; ++k) {
        for (int which = 0; which < (have_prev ? 2 : 1); ++which) {
            kps(k, which, 0, oc[k].kp[which][0]);
            kps(k, which, 1, oc[k].kp[which][1]);
            matches(k, which, oc[k].m[which]);
        }
    }
    for (int side = 0; side < 2; ++side) {
        FILE* f = open_dump(dir, side ? "right_feats_%04d.txt" : "left_feats_%04d.txt", it);
        for (int k = 0; k < n_oct; ++k) {
            const std::vector<KeyPoint>& v = oc[k].kp[0][side];
            for (const KeyPoint& p : v) std::fprintf(f, "%d %.2f %.2f %.8f\n", k, p.pt.x, p.pt.y, p.response);
            std::vector<uint8_t> desc(v.size() * 32);
            if (!v.empty() && svo_get_descriptors_oct(m_ctx, m_slot, k, 0, side, desc.data(), (int32_t)v.size()) < 0) throw_last(m_ctx, "svo_get_descriptors_oct");
            FILE* y = open_dump(dir, side ? "right_desc_o%02d_%04d.yml" : "left_desc_o%02d_%04d.yml", k, it);
            std::fprintf(y, "%%YAML:1.0\n---\n%s: !!opencv-matrix\n   rows: %d\n   cols: 32\n   dt: u\n   data: [ ", side ? "rightDescMat" : "leftDescMat", (int)v.size());
            for (size_t i = 0; i < desc.size(); ++i) std::fprintf(y, "%s%d", i ? (i % 16 ? ", " : ",\n       ") : "", (int)desc[i]);
            std::fprintf(y, " ]\n");
            std::fclose(y);
        }
        std::fclose(f);
    }
    FILE* f = open_dump(dir, "matches_%04d.txt", it);
    for (int k = 0; k < n_oct; ++k)
        for (const DMatch& m : oc[k].m[0]) std::fprintf(f, "%d %d %d %.2f\n", k, m.queryIdx, m.trainIdx, m.distance);
    std::fclose(f);
    if (!have_prev) return;
    f = open_dump(dir, "inter_frame_matching_%04d.txt", it);
    for (int k = 0; k < n_oct; ++k) {
        const int cap = keypoint_capacity(k);
        std::vector<int64_t>& tp = oc[k].tp;
        tp.resize((size_t)2 * cap);
        const int n = std::min(svo_get_tracked_pairs_oct(m_ctx, m_slot, k, tp.data(), cap), cap);
        if (n < 0) throw_last(m_ctx, "svo_get_tracked_pairs_oct");
        tp.resize((size_t)2 * n);
        for (int i = 0; i < n; ++i) {
            const DMatch& pm = oc[k].m[1][(size_t)tp[2 * i]];
            const DMatch& cm = oc[k].m[0][(size_t)tp[2 * i + 1]];
            const KeyPoint &pl = oc[k].kp[1][0][pm.queryIdx], &pr = oc[k].kp[1][1][pm.trainIdx];
            const KeyPoint &cl = oc[k].kp[0][0][cm.queryIdx], &cr = oc[k].kp[0][1][cm.trainIdx];
            std::fprintf(f, "%d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f %d %d\n", k, pl.pt.x, pl.pt.y, pr.pt.x, pr.pt.y, cl.pt.x, cl.pt.y,
                         cr.pt.x, cr.pt.y, (int)tp[2 * i], (int)tp[2 * i + 1]);
        }
    }
    std::fclose(f);
    // l-l / r-r (S4:173-234, brute-force tracker only): "prev.x prev.y cur.x cur.y distance" per surviving matL / matR entry
    if (params_if_match.ifm_method == TInterFrameMatchingParams::ifmDescBF) {
        const int k = n_oct - 1;                             // the reference overwrites the file octave after octave
        const int cap = keypoint_capacity(k);
        std::vector<int32_t> q((size_t)cap), tl((size_t)cap), tr((size_t)cap);
        std::vector<float> dl((size_t)cap), dr((size_t)cap);
        const int n = std::min(svo_get_track_candidates_oct(m_ctx, m_slot, k, q.data(), tl.data(), tr.data(), nullptr, nullptr, dl.data(), dr.data(), cap), cap);
        if (n < 0) throw_last(m_ctx, "svo_get_track_candidates_oct");
        for (int side = 0; side < 2; ++side) {
            FILE* fl = open_dump(dir, side ? "r-r_%04d.txt" : "l-l_%04d.txt", it);
            for (int i = 0; i < n; ++i) {
                const DMatch& pm = oc[k].m[1][(size_t)q[i]];
                const DMatch& cm = oc[k].m[0][(size_t)(side ? tr[i] : tl[i])];
                const KeyPoint& a = oc[k].kp[1][side][(size_t)(side ? pm.trainIdx : pm.queryIdx)];
                const KeyPoint& b = oc[k].kp[0][side][(size_t)(side ? cm.trainIdx : cm.queryIdx)];
                std::fprintf(fl, "%.2f %.2f %.2f %.2f %.2f\n", a.pt.x, a.pt.y, b.pt.x, b.pt.y, side ? dr[i] : dl[i]);
            }
            std::fclose(fl);
        }
    }
    if (result.error_code == voecBadTracking) return;       // stage 5 did not run (P:332)
    {   // optimization.txt (S5:477-492): in the working directory, overwritten at every call
        size_t total = 0;
        for (int k = 0; k < n_oct; ++k) total += oc[k].tp.size() / 2;
        std::vector<uint8_t> surv(total + 1);
        const int T = svo_get_stage5_survivors(m_ctx, m_slot, surv.data(), (int32_t)total);
        if (T < 0) throw_last(m_ctx, "svo_get_stage5_survivors");
        FILE* fo = std::fopen("optimization.txt", "wt");
        if (fo) {
            size_t i = 0;
            for (int k = 0; k < n_oct; ++k) {
                const float sc = n_oct > 1 ? (float)(1 << k) : 1.f;          // S5:422-456
                for (size_t j = 0; j + 1 < oc[k].tp.size(); j += 2, ++i) {
                    if (i >= (size_t)T || !surv[i]) continue;
                    const DMatch& pm = oc[k].m[1][(size_t)oc[k].tp[j]];
                    const DMatch& cm = oc[k].m[0][(size_t)oc[k].tp[j + 1]];
                    const KeyPoint &pl = oc[k].kp[1][0][pm.queryIdx], &pr = oc[k].kp[1][1][pm.trainIdx];
                    const KeyPoint &cl = oc[k].kp[0][0][cm.queryIdx], &cr = oc[k].kp[0][1][cm.trainIdx];
                    std::fprintf(fo, "%.2f,%.2f,%.2f,%.2f,%.2f,%.2f,%.2f,%.2f\n", pl.pt.x * sc, pl.pt.y * sc, pr.pt.x * sc, pr.pt.y * sc,
                                 cl.pt.x * sc, cl.pt.y * sc, cr.pt.x * sc, cr.pt.y * sc);
                }
            }
            std::fclose(fo);
        }
    }
    for (const char* fmt : {"out_residual_%04d.txt", "out_residual_final_%04d.txt"}) {
        FILE* fr = open_dump(dir, fmt, it);
        for (double v : result.out_residual) std::fprintf(fr, "%.3f\n", v);
        std::fclose(fr);
    }
}

static void check_request(const CStereoOdometryEstimator::TStereoOdometryRequest& request_data) {
    if (!request_data.stereo_imgs)   // process_new_image_pair.cpp:81
        throw std::runtime_error("Pointer 'request_data.stereo_imgs' must be set to stereo observation data!");
    if (request_data.use_precomputed_data)
        throw std::runtime_error("[Visual Odometry Error] use_precomputed_data is not supported (unusable in the reference too).");
    const CImage& L = request_data.stereo_imgs->imageLeft;
    const CImage& R = request_data.stereo_imgs->imageRight;
    if (!L.data || !R.data || L.width != R.width || L.height != R.height || L.stride != R.stride || L.channels != R.channels ||
        (L.channels != 1 && L.channels != 3))
        throw std::runtime_error("[Visual Odometry Error] left/right images must be 8-bit gray or BGR with equal geometry.");
}

// stage 1 (stage1_rectify.cpp:37-93) runs on the device: tell the context what the images are
void CStereoOdometryEstimator::apply_stage1_format(const CObservationStereoImages& obs) {
    const CImage& L = obs.imageLeft;
    if (L.channels != m_in_channels) {
        if (svo_set_input_format(m_ctx, L.channels) != SVO_OK) throw_last(m_ctx, "svo_set_input_format");
        m_in_channels = L.channels;
    }
    const bool rectify = !obs.areImagesRectified();
    if (rectify && m_map_a[0].size() != (size_t)L.width * L.height)
        throw std::runtime_error("[Visual Odometry Error] unrectified images need setRectifyMaps() with maps of the image size.");
    if (rectify != m_rectify_active || (rectify && !m_maps_on_device)) {
        const int rc1 = rectify ? svo_set_rectify_maps(m_ctx, m_map_xy[0].data(), m_map_a[0].data(), m_map_xy[1].data(), m_map_a[1].data())
                                : svo_set_rectify_maps(m_ctx, nullptr, nullptr, nullptr, nullptr);
        if (rc1 != SVO_OK) throw_last(m_ctx, "svo_set_rectify_maps");
        m_rectify_active = rectify;
        m_maps_on_device = rectify;
    }
}

// everything processNewImagePair does after the stages: result struct (P:171-176, 274-276 + stage 5's fields), dumps, counter
void CStereoOdometryEstimator::finish_step(const svo_result& r, const double* residuals, const int64_t* outliers, const int32_t* batch_stats,
                                           bool repeat, TStereoOdometryResult& result) {
    int32_t stats[4 * 4];
    int n_oct = n_octaves();
    if (batch_stats) std::memcpy(stats, batch_stats, sizeof stats);      // the counters of THIS batch (later ones may be running)
    else n_oct = svo_get_octave_stats(m_ctx, m_slot, stats, 4);
    if (n_oct < 1) throw_last(m_ctx, "svo_get_octave_stats");
    result.detected_feats.resize((size_t)n_oct);                                                           // P:171-176
    result.stereo_matches.resize((size_t)n_oct);                                                           // P:274-276
    for (int k = 0; k < n_oct; ++k) {
        result.detected_feats[k] = std::make_pair((size_t)stats[4 * k], (size_t)stats[4 * k + 1]);
        result.stereo_matches[k] = (size_t)stats[4 * k + 2];
    }
    fill_result(m_ctx, m_slot, r, residuals, outliers, result,
                r.n_residuals > 0 && (r.valid || r.error_code == voecIncrFuncCostStg2 || r.error_code == voecIncrFuncCostStg1));
    if (params_general.vo_save_files) {
        struct stat sb;                                     // P:46-47: the output directory is created on demand
        if (stat(params_general.vo_out_dir.c_str(), &sb) != 0 && mkdir(params_general.vo_out_dir.c_str(), 0777) != 0)
            throw std::runtime_error("[Visual Odometry Error] cannot create '" + params_general.vo_out_dir + "'");
        save_files(result, r.error_code != voecFirstIteration);
    }
    if (!repeat) ++m_it_counter;                            // P:380-381
}

void CStereoOdometryEstimator::processNewImagePair(TStereoOdometryRequest& request_data, TStereoOdometryResult& result) {
    check_request(request_data);
    const CImage& L = request_data.stereo_imgs->imageLeft;
    const CImage& R = request_data.stereo_imgs->imageRight;
    ensure_context(L.width, L.height);
    apply_stage1_format(*request_data.stereo_imgs);
    result.error_code = voecNone;                         // P:50
    const svo_camera cam = to_cam(request_data.stereo_cam);
    svo_result r;
    const int32_t slot = m_slot;
    const uint8_t rep = request_data.repeat ? 1 : 0;
    const uint8_t* lp = L.data;
    const uint8_t* rp = R.data;
    const int rc = svo_process_batch(m_ctx, 1, &slot, &lp, &rp, L.stride, &cam, request_data.repeat ? &rep : nullptr, &r);
    if (rc != SVO_OK) throw_last(m_ctx, "svo_process_batch");
    finish_step(r, nullptr, nullptr, nullptr, request_data.repeat, result);
}

// ---- CStereoOdometryPool -----------------------------------------------------------------------------------------
CStereoOdometryPool::CStereoOdometryPool(const CStereoOdometryEstimator& params, int width, int height, size_t n_estimators, int device) {
    if (n_estimators < 1) throw std::runtime_error("[Visual Odometry Error] a pool needs at least one estimator.");
    params.check_methods();
    svo_config cfg = params.make_config(width, height, (int)n_estimators);
    cfg.device = device;
    if (svo_create(&cfg, &m_ctx) != SVO_OK) CStereoOdometryEstimator::throw_last(nullptr, "svo_create");
    if (svo_set_result_detail(m_ctx, 1) != SVO_OK) { svo_destroy(m_ctx); CStereoOdometryEstimator::throw_last(nullptr, "svo_set_result_detail"); }
    m_w = width; m_h = height;
    int32_t info[4];
    const int n_oct = svo_get_octave_info(m_ctx, 0, info);
    m_cap_total = 0;
    for (int k = 0; k < n_oct; ++k) { svo_get_octave_info(m_ctx, k, info); m_cap_total += info[3]; }
    m_est.reserve(n_estimators);
    for (size_t k = 0; k < n_estimators; ++k) {
        CStereoOdometryEstimator* e = new CStereoOdometryEstimator();
        e->params_rectify = params.params_rectify; e->params_detect = params.params_detect; e->params_lr_match = params.params_lr_match;
        e->params_if_match = params.params_if_match; e->params_least_squares = params.params_least_squares;
        e->params_general = params.params_general;
        e->m_current_fast_th = params.m_current_fast_th; e->m_current_orb_th = params.m_current_orb_th;
        e->m_ctx = m_ctx; e->m_pool = this; e->m_slot = (int)k; e->m_w = width; e->m_h = height; e->m_device = device; e->m_dirty = false;
        m_est.push_back(e);
    }
}

CStereoOdometryPool::~CStereoOdometryPool() {
    for (CStereoOdometryEstimator* e : m_est) delete e;
    if (m_ctx) svo_destroy(m_ctx);
}

void CStereoOdometryPool::submit(const std::vector<size_t>& which, std::vector<Request>& requests) {
    const size_t n = which.size();
    if (n == 0 || n != requests.size()) throw std::runtime_error("[Visual Odometry Error] submit: one request per listed estimator.");
    std::vector<int32_t> slots(n);
    std::vector<const uint8_t*> L(n), R(n);
    std::vector<svo_camera> cams(n);
    Pending pd;
    pd.which = which;
    pd.repeat.resize(n);
    bool any_repeat = false;
    for (size_t i = 0; i < n; ++i) {
        if (which[i] >= m_est.size()) throw std::runtime_error("[Visual Odometry Error] submit: estimator index out of range.");
        check_request(requests[i]);
        const CObservationStereoImages& o = *requests[i].stereo_imgs;
        if (o.imageLeft.width != m_w || o.imageLeft.height != m_h || o.imageLeft.stride != requests[0].stereo_imgs->imageLeft.stride ||
            o.imageLeft.channels != requests[0].stereo_imgs->imageLeft.channels || o.areImagesRectified() != requests[0].stereo_imgs->areImagesRectified())
            throw std::runtime_error("[Visual Odometry Error] submit: all images of a batch share size, stride, format and rectification state.");
        slots[i] = (int32_t)which[i];
        L[i] = o.imageLeft.data; R[i] = o.imageRight.data;
        cams[i] = to_cam(requests[i].stereo_cam);
        pd.repeat[i] = requests[i].repeat ? 1 : 0;
        any_repeat = any_repeat || requests[i].repeat;
    }
    m_est[which[0]]->apply_stage1_format(*requests[0].stereo_imgs);          // the stage-1 format is a property of the shared context
    for (size_t i = 1; i < n; ++i) {
        CStereoOdometryEstimator* e = m_est[which[i]];
        e->m_in_channels = m_est[which[0]]->m_in_channels; e->m_rectify_active = m_est[which[0]]->m_rectify_active;
        e->m_maps_on_device = m_est[which[0]]->m_maps_on_device;
    }
    const int rc = svo_submit_batch(m_ctx, (int32_t)n, slots.data(), L.data(), R.data(), requests[0].stereo_imgs->imageLeft.stride,
                                    cams.data(), any_repeat ? pd.repeat.data() : nullptr);
    if (rc != SVO_OK) CStereoOdometryEstimator::throw_last(m_ctx, "svo_submit_batch");
    m_pending.push_back(pd);
}

void CStereoOdometryPool::wait(std::vector<Result>& results) {
    if (m_pending.empty()) throw std::runtime_error("[Visual Odometry Error] wait: no batch in flight.");
    const Pending pd = m_pending.front();
    m_pending.erase(m_pending.begin());
    const size_t n = pd.which.size();
    m_res.resize(n);
    m_resid.resize(n * (size_t)m_cap_total);
    m_outl.resize(n * (size_t)m_cap_total);
    m_stats.resize(n * 16);
    const int rc = svo_wait_batch_detail(m_ctx, m_res.data(), m_resid.data(), m_outl.data(), m_cap_total, m_stats.data());
    if (rc != SVO_OK) CStereoOdometryEstimator::throw_last(m_ctx, "svo_wait_batch_detail");
    results.resize(n);
    for (size_t i = 0; i < n; ++i) {
        results[i].error_code = voecNone;                  // P:50
        m_est[pd.which[i]]->finish_step(m_res[i], m_resid.data() + i * (size_t)m_cap_total, m_outl.data() + i * (size_t)m_cap_total,
                                        m_stats.data() + i * 16, pd.repeat[i] != 0, results[i]);
    }
}

void CStereoOdometryPool::processNewImagePairs(const std::vector<size_t>& which, std::vector<Request>& requests, std::vector<Result>& results) {
    if (!m_pending.empty()) throw std::runtime_error("[Visual Odometry Error] processNewImagePairs: batches are still in flight.");
    submit(which, requests);
    wait(results);
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
    int slot = 0;
    svo_ctx* ctx = context_for_size(w, h, &slot);
    std::vector<int64_t> tp(2 * tracked_pairs.size());
    for (size_t i = 0; i < tracked_pairs.size(); ++i) { tp[2 * i] = (int64_t)tracked_pairs[i].first; tp[2 * i + 1] = (int64_t)tracked_pairs[i].second; }
    const svo_camera cam = to_cam(stereo_camera);
    svo_result r;
    const bool custom = params_least_squares.use_custom_initial_pose;
    const int rc = svo_get_change_in_pose(
        ctx, slot, tp.data(), (int32_t)tracked_pairs.size(),
        reinterpret_cast<const svo_dmatch*>(pre_matches.data()), (int32_t)pre_matches.size(),
        reinterpret_cast<const svo_dmatch*>(cur_matches.data()), (int32_t)cur_matches.size(),
        reinterpret_cast<const svo_keypoint*>(pre_left_feats.data()), (int32_t)pre_left_feats.size(),
        reinterpret_cast<const svo_keypoint*>(pre_right_feats.data()), (int32_t)pre_right_feats.size(),
        reinterpret_cast<const svo_keypoint*>(cur_left_feats.data()), (int32_t)cur_left_feats.size(),
        reinterpret_cast<const svo_keypoint*>(cur_right_feats.data()), (int32_t)cur_right_feats.size(),
        &cam, custom ? 1 : 0, custom && ini_estimation.size() == 6 ? ini_estimation.data() : nullptr, &r);
    if (rc < 0) throw_last(ctx, "svo_get_change_in_pose");
    fill_result(ctx, slot, r, nullptr, nullptr, result, r.n_residuals > 0);
    return result.valid;
}

void CStereoOdometryEstimator::getProjectedCoords(const std::vector<DMatch>& pre_matches, const std::vector<KeyPoint>& pre_left_feats,
                                                  const std::vector<KeyPoint>& pre_right_feats,
                                                  const std::vector<std::pair<int, float> >& other_matches_tracked,
                                                  const TStereoCamera& stereo_camera, const CPose3D& change_pose,
                                                  std::vector<std::pair<Point2f, Point2f> >& pro_pre_feats) {
    if (other_matches_tracked.size() > pre_matches.size())
        throw std::runtime_error("[Visual Odometry Error] getProjectedCoords: more tracking flags than matches.");
    const int w = stereo_camera.leftCamera.ncols ? (int)stereo_camera.leftCamera.ncols : (m_w ? m_w : 640);
    const int h = stereo_camera.leftCamera.nrows ? (int)stereo_camera.leftCamera.nrows : (m_h ? m_h : 480);
    int slot_unused = 0;
    svo_ctx* ctx = context_for_size(w, h, &slot_unused);
    // C:459-464: auxPose = change_pose^-1, then its rotation-vector form [w1 w2 w3 t1 t2 t3]
    const double* M = change_pose.m;
    double Ri[9], ti[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Ri[3 * i + j] = M[4 * j + i];
        ti[i] = -(M[i] * M[3] + M[4 + i] * M[7] + M[8 + i] * M[11]);
    }
    double dp[6] = {0, 0, 0, ti[0], ti[1], ti[2]};
    const double tr = std::max(-1.0, std::min(3.0, Ri[0] + Ri[4] + Ri[8]));
    const double angle = std::acos(std::max(-1.0, std::min(1.0, (tr - 1.0) / 2.0)));
    const double vx = Ri[7] - Ri[5], vy = Ri[2] - Ri[6], vz = Ri[3] - Ri[1];      // 2 sin(angle) * axis
    if (angle > 1e-12) {
        const double s = std::sin(angle);
        if (std::fabs(s) > 1e-9) { const double k = angle / (2.0 * s); dp[0] = k * vx; dp[1] = k * vy; dp[2] = k * vz; }
        else {   // angle ~ pi: axis from the diagonal
            const double ax = std::sqrt(std::max(0.0, (Ri[0] + 1) / 2)), ay = std::sqrt(std::max(0.0, (Ri[4] + 1) / 2)),
                         az = std::sqrt(std::max(0.0, (Ri[8] + 1) / 2));
            dp[0] = angle * ax; dp[1] = angle * (Ri[1] + Ri[3] < 0 ? -ay : ay); dp[2] = angle * (Ri[2] + Ri[6] < 0 ? -az : az);
        }
    }
    std::vector<int32_t> flags(pre_matches.size(), 0);          // entries beyond other_matches_tracked are not visited (C:428)
    for (size_t m = 0; m < other_matches_tracked.size(); ++m) flags[m] = other_matches_tracked[m].first;
    const svo_camera cam = to_cam(stereo_camera);
    std::vector<float> out(4 * std::max<size_t>(1, pre_matches.size()));
    const int n = svo_get_projected_coords(ctx, reinterpret_cast<const svo_dmatch*>(pre_matches.data()), (int32_t)pre_matches.size(),
                                           reinterpret_cast<const svo_keypoint*>(pre_left_feats.data()), (int32_t)pre_left_feats.size(),
                                           reinterpret_cast<const svo_keypoint*>(pre_right_feats.data()), (int32_t)pre_right_feats.size(),
                                           flags.data(), &cam, dp, out.data(), (int32_t)pre_matches.size());
    if (n < 0) throw_last(ctx, "svo_get_projected_coords");
    pro_pre_feats.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        pro_pre_feats[i].first = Point2f{out[4 * i], out[4 * i + 1]};
        pro_pre_feats[i].second = Point2f{out[4 * i + 2], out[4 * i + 3]};
    }
}

void CStereoOdometryEstimator::getValues(std::vector<KeyPoint>& leftKP, std::vector<KeyPoint>& rightKP, Mat& leftDesc, Mat& rightDesc,
                                         std::vector<DMatch>& matches, std::vector<size_t>& matches_id) {
    if (!m_ctx) throw std::runtime_error("[Visual Odometry Error] no current frame.");
    const int cap = keypoint_capacity(0);               // octave 0 only, as in the reference (H:709)
    auto kps = [&](int side, std::vector<KeyPoint>& out, Mat& desc) {
        out.resize((size_t)cap);
        int n = svo_get_keypoints(m_ctx, m_slot, 0, side, reinterpret_cast<svo_keypoint*>(out.data()), cap);
        if (n < 0) throw_last(m_ctx, "svo_get_keypoints");
        n = std::min(n, cap);
        out.resize((size_t)n);
        desc.rows = n; desc.cols = 32; desc.data.resize((size_t)n * 32);
        if (n && svo_get_descriptors(m_ctx, m_slot, 0, side, desc.data.data(), n) < 0) throw_last(m_ctx, "svo_get_descriptors");
    };
    kps(0, leftKP, leftDesc);
    kps(1, rightKP, rightDesc);
    matches.resize((size_t)cap);
    const int m = svo_get_matches(m_ctx, m_slot, 0, reinterpret_cast<svo_dmatch*>(matches.data()), cap);
    if (m < 0) throw_last(m_ctx, "svo_get_matches");
    matches.resize((size_t)std::min(m, cap));
    std::vector<int64_t> ids((size_t)cap);
    const int ni = svo_get_match_ids(m_ctx, m_slot, 0, ids.data(), cap);
    matches_id.assign(ids.begin(), ids.begin() + std::min(std::max(ni, 0), cap));
}

}  // namespace rso
