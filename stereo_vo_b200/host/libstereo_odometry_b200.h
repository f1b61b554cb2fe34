// libstereo_odometry_b200.h -- C++ host shim with the reference's API shape
// (libstereo-odometry/include/libstereo-odometry.h:147-1047) on top of the C-ABI include/svo_b200.h.
//
// The reference's public boundary is the class rso::CStereoOdometryEstimator with
//   processNewImagePair(TStereoOdometryRequest&, TStereoOdometryResult&)      H:157-159
//   getChangeInPose(...)                                                       H:162-172
//   getValues(...)                                                             H:704-724
// plus its parameter structs.  This header mirrors those names, members and defaults for the
// FAST+ORB / smDescBF / ifmDescBF / Gauss-Newton path, but is MRPT- and OpenCV-free: this image has
// neither.  The stand-in types below keep the field names and layouts the path
// touches (cv::KeyPoint 28 B, cv::DMatch 16 B, TStereoCamera's fx()/cx()/rightCameraPose[0], CPose3D's
// x()/yaw()/...), so that code written against the reference compiles against this header after
// replacing the include; INTEGRATION.md shows the binding inside the real MRPT/OpenCV class.
//
// All arithmetic of the path runs in the sm_100a kernels of libsvo_b200.so; there is no CPU path here.
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/svo_b200.h"

namespace rso {

// ---- stand-ins for the OpenCV / MRPT types the path touches ------------------------------------------
struct Point2f { float x, y; };
struct KeyPoint {                 // == cv::KeyPoint (28 bytes)
    Point2f pt; float size, angle, response; int octave, class_id;
};
struct DMatch {                   // == cv::DMatch (16 bytes)
    int queryIdx, trainIdx, imgIdx; float distance;
};
static_assert(sizeof(KeyPoint) == sizeof(svo_keypoint) && sizeof(DMatch) == sizeof(svo_dmatch), "layout");

struct Mat {                      // descriptor matrix: rows x 32 bytes, CV_8U, continuous
    int rows = 0, cols = 32;
    std::vector<uint8_t> data;
    const uint8_t* ptr(int r) const { return data.data() + (size_t)r * cols; }
};

struct TCamera {                  // mrpt::utils::TCamera: the accessors the path uses
    double fx_ = 0, fy_ = 0, cx_ = 0, cy_ = 0;
    uint32_t ncols = 0, nrows = 0;
    double fx() const { return fx_; }
    double fy() const { return fy_; }
    double cx() const { return cx_; }
    double cy() const { return cy_; }
};
struct TStereoCamera {            // mrpt::utils::TStereoCamera
    TCamera leftCamera, rightCamera;
    double rightCameraPose[7] = {0, 0, 0, 1, 0, 0, 0};   // [0] = baseline (S5:185,515)
};

struct CPose3D {                  // mrpt::poses::CPose3D: what the demo reads (demo-main.cpp:235-253)
    double m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major homogeneous matrix
    double ypr[3] = {0, 0, 0};
    double x() const { return m[3]; }
    double y() const { return m[7]; }
    double z() const { return m[11]; }
    double yaw() const { return ypr[0]; }
    double pitch() const { return ypr[1]; }
    double roll() const { return ypr[2]; }
};

struct CImage {                   // 8-bit view (gray or BGR); the shim never owns or frees pixels
    const uint8_t* data = nullptr;
    int width = 0, height = 0;
    size_t stride = 0;
    int channels = 1;             // 1 = gray, 3 = BGR (converted on the device like FAST_REF_OR_CONVERT_TO_GRAY, S1:51-52)
    size_t getRowStride() const { return stride; }
    int getWidth() const { return width; }
    int getHeight() const { return height; }
    bool isColor() const { return channels == 3; }
};
struct CObservationStereoImages {
    CImage imageLeft, imageRight;
    double timestamp = 0;
    bool rectified = true;        // areImagesRectified() (S1:61): false = rectify with the maps given to setRectifyMaps()
    bool areImagesRectified() const { return rectified; }
};

/** Minimal stand-in for mrpt::utils::CConfigFile: "[SECTION]" headers, "key = value" lines, comments starting with
  * ';', '#' or "//" (anywhere on the line), case-insensitive section and key names, values trimmed. */
class CConfigFile {
public:
    CConfigFile() {}
    explicit CConfigFile(const std::string& fileName);      // throws std::runtime_error when the file cannot be read
    static CConfigFile fromString(const std::string& text);
    bool sectionExists(const std::string& section) const;
    int read_int(const std::string& section, const std::string& name, int defaultValue, bool failIfNotFound = false) const;
    double read_double(const std::string& section, const std::string& name, double defaultValue, bool failIfNotFound = false) const;
    bool read_bool(const std::string& section, const std::string& name, bool defaultValue, bool failIfNotFound = false) const;
    std::string read_string(const std::string& section, const std::string& name, const std::string& defaultValue,
                            bool failIfNotFound = false) const;
private:
    void parse(const std::string& text);
    bool find(const std::string& section, const std::string& name, std::string& out, bool failIfNotFound) const;
    std::map<std::string, std::map<std::string, std::string> > m_data;
};

typedef std::vector<KeyPoint> TKeyPointList;
typedef std::vector<DMatch> TDMatchList;
typedef std::vector<std::pair<size_t, size_t> > vector_index_pairs_t;

enum VOErrorCode { voecNone, voecBadCondNumber, voecIncrFuncCostStg1, voecIncrFuncCostStg2, voecFirstIteration, voecBadTracking };  // H:142

class CStereoOdometryPool;

class CStereoOdometryEstimator {
public:
    struct TStereoOdometryRequest {                 // H:205-233
        const CObservationStereoImages* stereo_imgs = nullptr;
        TStereoCamera stereo_cam;
        bool use_precomputed_data = false;          // the reference's precomputed-data seam is broken at this commit (SURVEY B-11)
        bool repeat = false;
    };
    struct TStereoOdometryResult {                  // H:235-264
        CPose3D outPose;
        std::vector<size_t> outliers;               // holds the INLIERS' current-match indices (S5:601-610)
        std::vector<double> out_residual;
        int num_it = 0, num_it_final = 0;
        bool valid = false;
        VOErrorCode error_code = voecNone;
        size_t tracked_feats_from_last_KF = 0, tracked_feats_from_last_frame = 0;
        std::vector<std::pair<size_t, size_t> > detected_feats;   // one entry per octave (P:171-176)
        std::vector<size_t> stereo_matches;                       // one entry per octave (P:274-276)
    };
    struct TGeneralParams {                         // H:266-283, process_new_image_pair.cpp:34-35
        bool vo_use_matches_ids = false, vo_save_files = false, vo_debug = false, vo_pause_it = false;
        std::string vo_out_dir = "out";
    };
    struct TInterFrameMatchingParams {              // H:285-310
        enum TIFMMethod { ifmDescBF = 0, ifmDescWin, ifmSAD, ifmOpticalFlow };
        TIFMMethod ifm_method = ifmDescBF;          // ifmDescBF and ifmDescWin are built
        int ifm_win_w = 20, ifm_win_h = 20;         // window of ifmDescWin (uninitialised in the reference, C:84)
        bool filter_fund_matrix = true;             // both descriptor branches always filter (S4:171-259, :673-713)
    };
    struct TLeastSquaresParams {                    // H:313-350, common.cpp:69-82
        bool use_robust_kernel = true;
        double kernel_param = 3.;
        size_t max_iters = 100, initial_max_iters = 10;
        double min_mod_out_vector = 1e-3, std_noise_pixels = 1.;
        size_t max_incr_cost = 3;
        double residual_threshold = 10.;
        size_t bad_tracking_th = 5;
        bool use_previous_pose_as_initial = true, use_custom_initial_pose = false;
    };
    struct TRectifyParams {                         // H:352-381, stage1_rectify.cpp:26-30
        int nOctaves = 3;                           // ctor default of the reference; 1..4 here
        // how CImagePyramid::buildPyramidFast halves an octave (MRPT is absent here): 0 = as MRPT 1.x picks (2x2 rounded
        // average when the width is a multiple of 16, cvPyrDown otherwise), 1 = always the average, 2 = always pyrDown
        int pyr_mode = 0;
    };
    struct TDetectParams {                          // H:383-447, stage2_detect.cpp:44-58
        enum NMSMethod { nmsmStandard, nmsmAdaptive };
        enum TDMethod { dmORB, dmFAST_ORB, dmFASTER, dmKLT };
        TDMethod detect_method = dmFAST_ORB;        // the reference's ctor default is dmFASTER (unimplemented there); built here: dmFAST_ORB, dmORB
        bool non_maximal_suppression = true;
        NMSMethod nmsMethod = nmsmStandard;
        size_t min_distance = 3;
        size_t orb_nfeats = 500, orb_nlevels = 8;
        int fast_min_th = 5, fast_max_th = 30, initial_FAST_threshold = 20;
    };
    struct TLeftRightMatchParams {                  // H:449-499, stage3_match_left_right.cpp:46-57
        enum TSMMethod { smDescBF = 0, smDescRbR, smSAD };
        TSMMethod match_method = smDescBF;          // ctor default smSAD; smDescBF and smDescRbR are built
        double orb_max_distance = 40;
        int orb_min_th = 30, orb_max_th = 100;
        bool enable_robust_1to1_match = false, rectified_images = true;
        double max_y_diff = 1;                      // SURVEY Appendix C (ctor default 0)
    };

    TRectifyParams params_rectify;
    TDetectParams params_detect;
    TLeftRightMatchParams params_lr_match;
    TInterFrameMatchingParams params_if_match;
    TLeastSquaresParams params_least_squares;
    TGeneralParams params_general;

    CStereoOdometryEstimator();
    ~CStereoOdometryEstimator();
    CStereoOdometryEstimator(const CStereoOdometryEstimator&) = delete;
    CStereoOdometryEstimator& operator=(const CStereoOdometryEstimator&) = delete;

    /** The main entry point (H:157-159).  Throws std::runtime_error on argument / device errors
      * (the reference throws through MRPT's THROW_EXCEPTION); VO failures travel in result.error_code. */
    void processNewImagePair(TStereoOdometryRequest& request_data, TStereoOdometryResult& result);

    /** Stage 5 alone on caller data (H:162-172, common.cpp:355-413). */
    bool getChangeInPose(const vector_index_pairs_t& tracked_pairs, const std::vector<DMatch>& pre_matches,
                         const std::vector<DMatch>& cur_matches, const std::vector<KeyPoint>& pre_left_feats,
                         const std::vector<KeyPoint>& pre_right_feats, const std::vector<KeyPoint>& cur_left_feats,
                         const std::vector<KeyPoint>& cur_right_feats, const TStereoCamera& stereo_camera,
                         TStereoOdometryResult& result, const std::vector<double>& ini_estimation = std::vector<double>(6, 0));

    /** Projects the stereo matches of the previous key frame that were NOT tracked (other_matches_tracked[m].first == -1)
      * into the current left / right images for the given change in pose (H:175-182, common.cpp:415-470).  The inverse
      * of change_pose is turned into [w t] here (MRPT's CPose3D::inverse + CPose3DRotVec(CPose3D), restated with the
      * standard SO(3) logarithm); triangulation and projection run on the device. */
    void getProjectedCoords(const std::vector<DMatch>& pre_matches, const std::vector<KeyPoint>& pre_left_feats,
                            const std::vector<KeyPoint>& pre_right_feats, const std::vector<std::pair<int, float> >& other_matches_tracked,
                            const TStereoCamera& stereo_camera, const CPose3D& change_pose,
                            std::vector<std::pair<Point2f, Point2f> >& pro_pre_feats);

    /** Copies of the current frame's data (H:704-724). */
    void getValues(std::vector<KeyPoint>& leftKP, std::vector<KeyPoint>& rightKP, Mat& leftDesc, Mat& rightDesc,
                   std::vector<DMatch>& matches, std::vector<size_t>& matches_id);

    // H:529-541.  As in the reference these only change two ints that the next call reads (S2:510, S3:164, S4:149):
    // the previous frame, m_last_computed_pose, the error state and the match IDs stay (svo_set_thresholds).
    int getFASTThreshold() const { return m_current_fast_th; }
    void setFASTThreshold(int v) { m_current_fast_th = std::min(params_detect.fast_max_th, std::max(params_detect.fast_min_th, v)); push_thresholds(); }
    void resetFASTThreshold() { m_current_fast_th = params_detect.initial_FAST_threshold; push_thresholds(); }
    bool isFASTThMin() const { return m_current_fast_th == params_detect.fast_min_th; }
    bool isFASTThMax() const { return m_current_fast_th == params_detect.fast_max_th; }
    int getORBThreshold() const { return m_current_orb_th; }
    void setORBThreshold(int v) { m_current_orb_th = std::min(params_lr_match.orb_max_th, std::max(params_lr_match.orb_min_th, v)); push_thresholds(); }
    void resetORBThreshold() { m_current_orb_th = (int)params_lr_match.orb_max_distance; push_thresholds(); }
    bool isORBThMin() const { return m_current_orb_th == params_lr_match.orb_min_th; }
    bool isORBThMax() const { return m_current_orb_th == params_lr_match.orb_max_th; }
    void setVerbosityLevel(int level) { m_verbose_level = level; }

    /** Loads configuration from an INI file (H:554-663).  Sections must be (in this order) related to:
      * RECTIFY, DETECT, MATCH, IF-MATCH, LEAST_SQUARES, GUI, GENERAL; empty names are skipped.  Keys of methods this
      * build does not have (KLT, SAD, window trackers, GUI) are accepted and ignored.  Ends with
      * resetFASTThreshold() + resetORBThreshold() like the reference (H:661-662). */
    void loadParamsFromConfigFile(const CConfigFile& iniFile, const std::vector<std::string>& sections);
    /** Loads configuration from an INI file from its name (H:666-672). */
    void loadParamsFromConfigFileName(const std::string& fileName, const std::vector<std::string>& sections);

    // Match-ID / key-frame control (H:675-726); IDs exist only with params_general.vo_use_matches_ids
    void setThisFrameAsKF();
    void resetIds();
    void setIds(const std::vector<size_t>& ids);
    void setMaxMatchID(size_t id);
    std::vector<size_t> getCopyCurrentIDs(size_t octave = 0);   // getRefCurrentIDs(octave) by value (state lives on the device)
    /** lr_pairing_data[octave].matches_lr_row_index of the current frame (H:786-789, built at S3:425-445): for every image
      * row the index of the first stereo match at or below it, plus the final entry = number of LEFT features (sic, S3:443).
      * Rebuilt here from the frame's keypoints and matches ("kept in shim"); the device rebuilds it for its own
      * windowed tracker. */
    std::vector<size_t> getMatchesRowIndex(size_t octave = 0);

    /** Rectification maps for observations with rectified == false: what CStereoRectifyMap::setFromCamParams computes
      * on first use (S1:69-70) with cv::initUndistortRectifyMap(..., CV_16SC2): per side H*W (x, y) int16 pairs and
      * H*W uint16 interpolation indices.  The rectification itself (cv::remap) runs on the device. */
    void setRectifyMaps(const std::vector<int16_t>& map_xy_left, const std::vector<uint16_t>& map_a_left,
                        const std::vector<int16_t>& map_xy_right, const std::vector<uint16_t>& map_a_right);

    /** Apply the param structs (what loadParamsFromConfigFile does at H:554-663, minus the INI parsing):
      * re-reads every params_* member and resets both thresholds (H:661-662). */
    void applyParams() { m_dirty = true; resetFASTThreshold(); resetORBThreshold(); }

    /** CUDA device ordinal used by this estimator (default 0); must be set before the first call. */
    void setDevice(int device) { m_device = device; m_dirty = true; }

    /** true when this estimator is one slot of a CStereoOdometryPool (shared device context) */
    bool pooled() const { return m_pool != nullptr; }

private:
    friend class CStereoOdometryPool;
    void ensure_context(int width, int height);
    svo_ctx* context_for_size(int width, int height, int* slot);   // getChangeInPose / getProjectedCoords never tear a live sequence down
    void push_thresholds();
    svo_config make_config(int width, int height, int max_seqs) const;
    void check_methods() const;
    // stage1_rectify.cpp:80: n_octaves = detect_method == dmORB ? 1 : params_rectify.nOctaves
    int n_octaves() const { return params_detect.detect_method == TDetectParams::dmORB ? 1 : params_rectify.nOctaves; }
    void apply_stage1_format(const CObservationStereoImages& obs);
    void finish_step(const svo_result& r, const double* residuals, const int64_t* outliers, const int32_t* batch_stats, bool repeat,
                     TStereoOdometryResult& result);
    int keypoint_capacity(int octave) const;
    void save_files(const TStereoOdometryResult& result, bool have_prev);   // params_general.vo_save_files
    static void throw_last(svo_ctx* c, const char* what);
    svo_ctx* m_ctx = nullptr;           // owned unless m_pool is set
    svo_ctx* m_aux_ctx = nullptr;       // scratch context for getChangeInPose / getProjectedCoords on another image size
    int m_aux_w = 0, m_aux_h = 0;
    CStereoOdometryPool* m_pool = nullptr;
    int m_slot = 0;
    int m_w = 0, m_h = 0, m_device = 0;
    bool m_dirty = true;
    int m_current_fast_th = 20, m_current_orb_th = 60;   // ctor values, common.cpp:35-36
    int m_verbose_level = 1;
    unsigned int m_it_counter = 0;
    int m_in_channels = 1;
    bool m_maps_on_device = false, m_rectify_active = false;
    std::vector<int16_t> m_map_xy[2];
    std::vector<uint16_t> m_map_a[2];
};

/** A pool of estimators on one GPU that share ONE device context.
  *
  * The reference's boundary is one processNewImagePair per estimator (libstereo-odometry.h:157-159).  A B200 advances
  * hundreds of independent stereo sequences in the time of one, so the pool binds estimator k to slot k of a single
  * svo_ctx: every estimator keeps its own previous frame, m_last_computed_pose, error state, thresholds and match IDs
  * (all the members of H:733-754), and all of them are advanced together by one call:
  *   - processNewImagePairs(): one processNewImagePair per listed estimator, synchronous;
  *   - submit() / wait(): the same, asynchronous, up to three batches in flight (uploads of the next batches overlap the
  *     kernels of the current one);
  *   - pool[k].processNewImagePair(req, res) still works for one estimator at a time.
  * Results are complete TStereoOdometryResult objects (outPose, outliers, out_residual, counters, per-octave statistics).
  * All estimators of a pool share the parameter set the pool was created from (thresholds are per estimator). */
class CStereoOdometryPool {
public:
    typedef CStereoOdometryEstimator::TStereoOdometryRequest Request;
    typedef CStereoOdometryEstimator::TStereoOdometryResult Result;
    /** params: an estimator whose params_* members (and current thresholds) configure the pool */
    CStereoOdometryPool(const CStereoOdometryEstimator& params, int width, int height, size_t n_estimators, int device = 0);
    ~CStereoOdometryPool();
    CStereoOdometryPool(const CStereoOdometryPool&) = delete;
    CStereoOdometryPool& operator=(const CStereoOdometryPool&) = delete;
    size_t size() const { return m_est.size(); }
    CStereoOdometryEstimator& operator[](size_t k) { return *m_est[k]; }

    /** requests[i] goes to estimator which[i] (distinct); results[i] receives its TStereoOdometryResult */
    void processNewImagePairs(const std::vector<size_t>& which, std::vector<Request>& requests, std::vector<Result>& results);
    void submit(const std::vector<size_t>& which, std::vector<Request>& requests);
    /** blocks until the OLDEST submitted batch is complete */
    void wait(std::vector<Result>& results);
    size_t in_flight() const { return m_pending.size(); }
    svo_ctx* context() { return m_ctx; }

private:
    struct Pending { std::vector<size_t> which; std::vector<uint8_t> repeat; };
    svo_ctx* m_ctx = nullptr;
    int m_w = 0, m_h = 0, m_cap_total = 0;
    std::vector<CStereoOdometryEstimator*> m_est;
    std::vector<Pending> m_pending;
    std::vector<svo_result> m_res;
    std::vector<double> m_resid;
    std::vector<int64_t> m_outl;
    std::vector<int32_t> m_stats;
};

}  // namespace rso
