// Host unit test of the shim's INI loader (libstereo-odometry.h:554-672 semantics): section order, defaults kept for
// missing keys, if_match_method defaulting to 0, thresholds reset at the end, comments, bool spellings.
#include <cstdio>
#include <string>
#include <vector>
#include "../../stereo_vo_b200/host/libstereo_odometry_b200.h"

using namespace rso;
static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); ++g_fail; } } while (0)

int main(int argc, char** argv) {
    const std::vector<std::string> sec = {"RECTIFY", "DETECT", "MATCH", "IF-MATCH", "LEAST_SQUARES", "GUI", "GENERAL"};
    const char* text =
        "# comment\n[RECTIFY]\nnOctaves = 1\n[detect]\ndetect_method=1 ; dmFAST_ORB\norb_nfeats = 250 // cap 500\n"
        "initial_FAST_threshold = 17\nfast_min_th=6\nnon_maximal_suppression = no\n[MATCH]\nmatch_method = 0\nmax_y_diff = 2.5\n"
        "orb_max_distance = 55\nenable_robust_1to1_match = TRUE\n[IF-MATCH]\n[LEAST_SQUARES]\nmax_iters = 50\nkernel_param=2.5\n"
        "use_robust_kernel = 0\nbad_tracking_th = 7\n[GUI]\nshow_gui = true\n[GENERAL]\nvo_use_matches_ids = yes\nvo_out_dir = results\n";
    CStereoOdometryEstimator e;
    e.params_if_match.ifm_method = CStereoOdometryEstimator::TInterFrameMatchingParams::ifmSAD;   // must be overwritten by the default 0
    e.loadParamsFromConfigFile(CConfigFile::fromString(text), sec);
    CHECK(e.params_rectify.nOctaves == 1);
    CHECK(e.params_detect.detect_method == CStereoOdometryEstimator::TDetectParams::dmFAST_ORB);
    CHECK(e.params_detect.orb_nfeats == 250 && e.params_detect.initial_FAST_threshold == 17 && e.params_detect.fast_min_th == 6);
    CHECK(e.params_detect.fast_max_th == 30 && e.params_detect.min_distance == 3);           // defaults kept
    CHECK(!e.params_detect.non_maximal_suppression);
    CHECK(e.params_lr_match.max_y_diff == 2.5 && e.params_lr_match.orb_max_distance == 55 && e.params_lr_match.enable_robust_1to1_match);
    CHECK(e.params_if_match.ifm_method == CStereoOdometryEstimator::TInterFrameMatchingParams::ifmDescBF);
    CHECK(e.params_least_squares.max_iters == 50 && e.params_least_squares.kernel_param == 2.5 && !e.params_least_squares.use_robust_kernel);
    CHECK(e.params_least_squares.bad_tracking_th == 7 && e.params_least_squares.initial_max_iters == 10);
    CHECK(e.params_general.vo_use_matches_ids && e.params_general.vo_out_dir == "results");
    CHECK(e.getFASTThreshold() == 17 && e.getORBThreshold() == 55);                              // H:661-662
    // empty section names are skipped
    CStereoOdometryEstimator e2;
    e2.loadParamsFromConfigFile(CConfigFile::fromString(text), {"", "", "", "", "", "", ""});
    CHECK(e2.params_detect.orb_nfeats == 500 && e2.getORBThreshold() == 40);
    CHECK(e2.params_rectify.nOctaves == 3);                                                      // ctor default, stage1_rectify.cpp:28
    bool threw = false;
    try { e2.loadParamsFromConfigFile(CConfigFile::fromString(text), {"RECTIFY"}); } catch (const std::exception&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { e2.loadParamsFromConfigFileName("/nonexistent/file.ini", sec); } catch (const std::exception&) { threw = true; }
    CHECK(threw);
    if (argc > 1) {   // the shipped hot-path configuration
        CStereoOdometryEstimator e3;
        e3.loadParamsFromConfigFileName(argv[1], sec);
        CHECK(e3.params_detect.orb_nfeats == 500 && e3.params_lr_match.max_y_diff == 1 && e3.getORBThreshold() == 40);
        CConfigFile ini(argv[1]);
        CHECK(ini.read_string("GENERAL", "camera_pose_on_robot", "") == "[0 0 0 0 0 0]");
    }
    std::printf("failures=%d\n", g_fail);
    return g_fail ? 1 : 0;
}
