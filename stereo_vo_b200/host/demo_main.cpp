// demo_main.cpp -- the per-frame loop of demo-stereo-odometry/demo-main.cpp:210-253 on top of the host
// shim: process every stereo pair of a raw sequence file, compose the camera pose
// (pose <- pose * K * delta * K^-1, demo-main.cpp:235-242, K = camera pose on the robot, identity here)
// and write one "x y z yaw pitch roll" line per frame (demo-main.cpp:251-253).
//
// Sequence file (written by tests / tools, little endian):
//   int32 width, height, n_frames; double fx, fy, cx, cy, baseline; then n_frames x (left, right) images.
//
// usage: svo_demo <sequence.bin> <camera_pose.txt> [orb_nfeats [nOctaves] | config.ini]
//   config.ini: the 7 sections RECTIFY, DETECT, MATCH, IF-MATCH, LEAST_SQUARES, GUI, GENERAL (demo-main.cpp:155-167),
//   GENERAL/vo_verbosity and GENERAL/camera_pose_on_robot = [x y z yaw pitch roll] in metres / degrees (:168-172, :182)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

#include "libstereo_odometry_b200.h"

using namespace rso;

static void matmul4(const double* a, const double* b, double* c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[4 * i + k] * b[4 * k + j];
            c[4 * i + j] = s;
        }
}

// mrpt::poses::CPose3D(x, y, z, yaw, pitch, roll): R = Rz(yaw) * Ry(pitch) * Rx(roll)
static void pose_matrix(const double* v, double* m) {
    const double cy = std::cos(v[3]), sy = std::sin(v[3]), cp = std::cos(v[4]), sp = std::sin(v[4]), cr = std::cos(v[5]), sr = std::sin(v[5]);
    const double r[16] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, v[0],
                          sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, v[1],
                          -sp,     cp * sr,                cp * cr,                v[2],
                          0, 0, 0, 1};
    std::memcpy(m, r, sizeof(r));
}
static void rigid_inverse(const double* a, double* inv) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) inv[4 * i + j] = a[4 * j + i];
        inv[4 * i + 3] = -(a[i] * a[3] + a[4 + i] * a[7] + a[8 + i] * a[11]);
    }
    inv[12] = inv[13] = inv[14] = 0; inv[15] = 1;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <sequence.bin> <camera_pose.txt> [orb_nfeats]\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("sequence"); return 2; }
    int32_t hdr[3];
    double cam[5];
    if (std::fread(hdr, 4, 3, f) != 3 || std::fread(cam, 8, 5, f) != 5) { std::fprintf(stderr, "bad header\n"); return 2; }
    const int W = hdr[0], H = hdr[1], N = hdr[2];
    try {
        CStereoOdometryEstimator engine;
        double K[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, Kinv[16];   // camera pose on the robot
        const bool have_ini = argc > 3 && std::strlen(argv[3]) > 4 && !std::strcmp(argv[3] + std::strlen(argv[3]) - 4, ".ini");
        if (have_ini) {
            const std::vector<std::string> sections = {"RECTIFY", "DETECT", "MATCH", "IF-MATCH", "LEAST_SQUARES", "GUI", "GENERAL"};
            engine.loadParamsFromConfigFileName(argv[3], sections);                     // demo-main.cpp:167
            CConfigFile ini(argv[3]);
            engine.setVerbosityLevel(ini.read_int("GENERAL", "vo_verbosity", 0));       // :168
            std::string vp = ini.read_string("GENERAL", "camera_pose_on_robot", "");    // :171-172, :182
            for (char& ch : vp) if (ch == '[' || ch == ']' || ch == ',') ch = ' ';
            std::istringstream is(vp);
            double v[6] = {0, 0, 0, 0, 0, 0};
            for (int k = 0; k < 6 && (is >> v[k]); ++k) {}
            for (int k = 3; k < 6; ++k) v[k] *= M_PI / 180.0;
            pose_matrix(v, K);
        } else {
            if (argc > 3) engine.params_detect.orb_nfeats = (size_t)std::atoi(argv[3]);
            engine.params_rectify.nOctaves = argc > 4 ? std::atoi(argv[4]) : 1;        // the hot-path INI runs one octave
            engine.applyParams();                     // what loadParamsFromConfigFile ends with (H:661-662)
            engine.setVerbosityLevel(0);
        }
        rigid_inverse(K, Kinv);
        CStereoOdometryEstimator::TStereoOdometryRequest req;
        req.stereo_cam.leftCamera.fx_ = req.stereo_cam.rightCamera.fx_ = cam[0];
        req.stereo_cam.leftCamera.fy_ = req.stereo_cam.rightCamera.fy_ = cam[1];
        req.stereo_cam.leftCamera.cx_ = req.stereo_cam.rightCamera.cx_ = cam[2];
        req.stereo_cam.leftCamera.cy_ = req.stereo_cam.rightCamera.cy_ = cam[3];
        req.stereo_cam.leftCamera.ncols = req.stereo_cam.rightCamera.ncols = (uint32_t)W;
        req.stereo_cam.leftCamera.nrows = req.stereo_cam.rightCamera.nrows = (uint32_t)H;
        req.stereo_cam.rightCameraPose[0] = cam[4];
        std::vector<uint8_t> L((size_t)W * H), R((size_t)W * H);
        double pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        double ypr[3] = {0, 0, 0};
        FILE* out = std::fopen(argv[2], "wt");
        if (!out) { std::perror("output"); return 2; }
        for (int i = 0; i < N; ++i) {
            if (std::fread(L.data(), 1, L.size(), f) != L.size() || std::fread(R.data(), 1, R.size(), f) != R.size()) {
                std::fprintf(stderr, "truncated sequence\n");
                return 2;
            }
            CObservationStereoImages obs;
            obs.imageLeft.data = L.data(); obs.imageLeft.width = W; obs.imageLeft.height = H; obs.imageLeft.stride = (size_t)W;
            obs.imageRight = obs.imageLeft;
            obs.imageRight.data = R.data();
            req.stereo_imgs = &obs;
            CStereoOdometryEstimator::TStereoOdometryResult res;
            engine.processNewImagePair(req, res);
            if (res.valid) {
                double kd[16], kdk[16], next[16];     // pose <- pose * (K * delta * K^-1), demo-main.cpp:235-242
                matmul4(K, res.outPose.m, kd);
                matmul4(kd, Kinv, kdk);
                matmul4(pose, kdk, next);
                for (int k = 0; k < 16; ++k) pose[k] = next[k];
                ypr[1] = std::atan2(-pose[8], std::hypot(pose[0], pose[4]));
                ypr[0] = std::atan2(pose[4], pose[0]);
                ypr[2] = std::atan2(pose[9], pose[10]);
            }
            std::fprintf(out, "%.3f %.3f %.3f %.3f %.3f %.3f\n", pose[3], pose[7], pose[11], ypr[0], ypr[1], ypr[2]);
            std::printf("Frame: %d valid=%d error_code=%d kp=%zu/%zu matches=%zu tracked=%zu it=%d+%d\n", i + 1, (int)res.valid,
                        (int)res.error_code, res.detected_feats[0].first, res.detected_feats[0].second, res.stereo_matches[0],
                        res.tracked_feats_from_last_frame, res.num_it, res.num_it_final);
        }
        std::fclose(out);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    std::fclose(f);
    return 0;
}
