// pool_driver.cpp -- many rso::CStereoOdometryEstimator instances behind one device context (CStereoOdometryPool):
// the throughput of the reference's class API (one processNewImagePair per estimator, libstereo-odometry.h:157-159)
// when the estimators of one GPU are advanced together, and a self-check that pooling changes no estimator's results.
//
// Sequence file (little endian): int32 width, height, n_sequences, n_frames; double fx, fy, cx, cy, baseline;
// then n_sequences x n_frames x (left, right) 8-bit gray images.
//
// usage: svo_pool_driver bench <sequences.bin> <n_estimators> <steps> <warmup> [orb_nfeats] [device]
//          estimator e plays sequence e % n_sequences forward and backward, starting e frames in; every step submits one
//          stereo pair per estimator from pinned host memory (three batches in flight) and receives the complete
//          TStereoOdometryResult of every estimator; prints one JSON object.
//        svo_pool_driver check <sequences.bin> [orb_nfeats]
//          every sequence through (a) its own stand-alone estimator, (b) a pool, synchronously, (c) a pool, asynchronously
//          with three batches in flight; dynamic thresholds changed at frame 3; prints "failures=N".
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "libstereo_odometry_b200.h"

using namespace rso;
typedef CStereoOdometryEstimator::TStereoOdometryRequest Request;
typedef CStereoOdometryEstimator::TStereoOdometryResult Result;

struct Sequences {
    int W = 0, H = 0, S = 0, F = 0;
    double cam[5] = {0, 0, 0, 0, 0};
    uint8_t* px = nullptr;                   // pinned: [S][F][2][H*W]
    const uint8_t* img(int s, int f, int side) const { return px + (((size_t)s * F + f) * 2 + side) * (size_t)W * H; }
    ~Sequences() { svo_host_free(px); }
};

static bool load(const char* path, Sequences& q) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    int32_t hdr[4];
    if (std::fread(hdr, 4, 4, f) != 4 || std::fread(q.cam, 8, 5, f) != 5) { std::fclose(f); return false; }
    q.W = hdr[0]; q.H = hdr[1]; q.S = hdr[2]; q.F = hdr[3];
    const size_t bytes = (size_t)q.S * q.F * 2 * q.W * q.H;
    q.px = (uint8_t*)svo_host_alloc(bytes);
    const bool ok = q.px && std::fread(q.px, 1, bytes, f) == bytes;
    std::fclose(f);
    return ok;
}

static int pingpong(int t, int n) {
    if (n < 2) return 0;
    const int period = 2 * (n - 1);
    t %= period;
    return t < n ? t : period - t;
}

static TStereoCamera camera(const Sequences& q) {
    TStereoCamera c;
    c.leftCamera.fx_ = c.rightCamera.fx_ = q.cam[0]; c.leftCamera.fy_ = c.rightCamera.fy_ = q.cam[1];
    c.leftCamera.cx_ = c.rightCamera.cx_ = q.cam[2]; c.leftCamera.cy_ = c.rightCamera.cy_ = q.cam[3];
    c.leftCamera.ncols = c.rightCamera.ncols = (uint32_t)q.W; c.leftCamera.nrows = c.rightCamera.nrows = (uint32_t)q.H;
    c.rightCameraPose[0] = q.cam[4];
    return c;
}

static void observation(const Sequences& q, int s, int f, CObservationStereoImages& o) {
    o.imageLeft.data = q.img(s, f, 0); o.imageRight.data = q.img(s, f, 1);
    o.imageLeft.width = o.imageRight.width = q.W; o.imageLeft.height = o.imageRight.height = q.H;
    o.imageLeft.stride = o.imageRight.stride = (size_t)q.W;
    o.imageLeft.channels = o.imageRight.channels = 1;
    o.rectified = true;
}

static void configure(CStereoOdometryEstimator& e, int nfeats) {
    e.params_rectify.nOctaves = 1;
    e.params_detect.orb_nfeats = (size_t)nfeats;
    e.applyParams();
}

static bool same(const Result& a, const Result& b) {
    return a.valid == b.valid && a.error_code == b.error_code && a.num_it == b.num_it && a.num_it_final == b.num_it_final &&
           a.tracked_feats_from_last_frame == b.tracked_feats_from_last_frame && a.tracked_feats_from_last_KF == b.tracked_feats_from_last_KF &&
           a.detected_feats == b.detected_feats && a.stereo_matches == b.stereo_matches && a.outliers == b.outliers &&
           a.out_residual == b.out_residual && !std::memcmp(a.outPose.m, b.outPose.m, sizeof a.outPose.m);
}

static int run_check(const Sequences& q, int nfeats) {
    int failures = 0;
    const TStereoCamera cam = camera(q);
    const int S = q.S, F = q.F;
    // (a) stand-alone estimators, the literal reference usage
    std::vector<std::vector<Result> > alone((size_t)S, std::vector<Result>((size_t)F));
    for (int s = 0; s < S; ++s) {
        CStereoOdometryEstimator e;
        configure(e, nfeats);
        for (int f = 0; f < F; ++f) {
            if (f == 3) { e.setFASTThreshold(14); e.setORBThreshold(50); }
            CObservationStereoImages o;
            observation(q, s, f, o);
            Request r;
            r.stereo_imgs = &o; r.stereo_cam = cam;
            e.processNewImagePair(r, alone[s][f]);
        }
        // the setters must not have reset the sequence: frame 3 is tracked against frame 2
        if (F > 3 && !(alone[s][3].tracked_feats_from_last_frame > 0 && alone[s][3].error_code != voecFirstIteration)) { ++failures; std::printf("sequence %d: threshold setters reset the sequence\n", s); }
        if (e.getFASTThreshold() != (F > 3 ? 14 : 20) ) { ++failures; std::printf("sequence %d: FAST threshold getter\n", s); }
    }
    for (int mode = 0; mode < 2; ++mode) {
        CStereoOdometryEstimator proto;
        configure(proto, nfeats);
        CStereoOdometryPool pool(proto, q.W, q.H, (size_t)S);
        std::vector<size_t> which((size_t)S);
        for (int s = 0; s < S; ++s) which[s] = (size_t)(S - 1 - s);        // a non-identity slot order
        std::vector<std::vector<CObservationStereoImages> > obs((size_t)F, std::vector<CObservationStereoImages>((size_t)S));
        std::vector<std::vector<Request> > req((size_t)F, std::vector<Request>((size_t)S));
        for (int f = 0; f < F; ++f)
            for (int i = 0; i < S; ++i) {
                observation(q, (int)which[i], f, obs[f][i]);
                req[f][i].stereo_imgs = &obs[f][i]; req[f][i].stereo_cam = cam;
            }
        std::vector<std::vector<Result> > got((size_t)F);
        if (mode == 0) {
            for (int f = 0; f < F; ++f) {
                if (f == 3) for (int s = 0; s < S; ++s) { pool[s].setFASTThreshold(14); pool[s].setORBThreshold(50); }
                pool.processNewImagePairs(which, req[f], got[f]);
            }
        } else {
            int waited = 0;
            for (int f = 0; f < F; ++f) {
                if (f == 3) for (int s = 0; s < S; ++s) { pool[s].setFASTThreshold(14); pool[s].setORBThreshold(50); }   // stream-ordered
                pool.submit(which, req[f]);
                if (pool.in_flight() == 3) pool.wait(got[waited++]);
            }
            while (pool.in_flight()) pool.wait(got[waited++]);
        }
        for (int f = 0; f < F; ++f)
            for (int i = 0; i < S; ++i)
                if (!same(got[f][i], alone[which[i]][f])) { ++failures; std::printf("%s pool: sequence %d frame %d differs from the stand-alone estimator\n", mode ? "async" : "sync", (int)which[i], f); }
        // one estimator of the pool driven alone afterwards keeps its own state
        CObservationStereoImages o;
        observation(q, 0, F - 1, o);
        Request r; r.stereo_imgs = &o; r.stereo_cam = cam; r.repeat = true;
        Result again;
        pool[0].processNewImagePair(r, again);
        if (!(again.valid == alone[0][F - 1].valid && again.num_it == alone[0][F - 1].num_it)) { ++failures; std::printf("pooled estimator driven alone: repeat differs\n"); }
    }
    std::printf("failures=%d\n", failures);
    return failures ? 1 : 0;
}

static int run_bench(const Sequences& q, int n_est, int steps, int warmup, int nfeats, int device) {
    const TStereoCamera cam = camera(q);
    CStereoOdometryEstimator proto;
    configure(proto, nfeats);
    proto.setDevice(device);
    CStereoOdometryPool pool(proto, q.W, q.H, (size_t)n_est, device);
    std::vector<size_t> which((size_t)n_est);
    for (int e = 0; e < n_est; ++e) which[e] = (size_t)e;
    const int depth = 3;
    // request objects of the batches in flight stay alive until their wait()
    std::vector<std::vector<CObservationStereoImages> > obs((size_t)depth, std::vector<CObservationStereoImages>((size_t)n_est));
    std::vector<std::vector<Request> > req((size_t)depth, std::vector<Request>((size_t)n_est));
    std::vector<Result> res;
    size_t valid = 0, resid_values = 0, outl_values = 0;
    auto submit = [&](int t) {
        const int b = t % depth;
        for (int e = 0; e < n_est; ++e) {
            observation(q, e % q.S, pingpong(t + e, q.F), obs[b][e]);
            req[b][e].stereo_imgs = &obs[b][e]; req[b][e].stereo_cam = cam;
        }
        pool.submit(which, req[b]);
    };
    auto wait = [&](bool count) {
        pool.wait(res);
        if (count) for (const Result& r : res) { valid += r.valid; resid_values += r.out_residual.size(); outl_values += r.outliers.size(); }
    };
    int t = 0;
    for (int k = 0; k < warmup + 2; ++k) { submit(t++); wait(false); }
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; ++k) {
        submit(t++);
        if ((int)pool.in_flight() == depth) wait(true);
    }
    while (pool.in_flight()) wait(true);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"pairs_per_s\": %.1f, \"ms_per_step\": %.4f, \"estimators\": %d, \"steps\": %d, \"valid\": %zu, \"of\": %zu, "
                "\"h2d_bytes_per_step\": %zu, \"d2h_result_values_per_step\": %.1f}\n",
                (double)n_est * steps / sec, 1e3 * sec / steps, n_est, steps, valid, (size_t)n_est * steps,
                (size_t)2 * n_est * q.W * q.H, (double)(resid_values + outl_values) / steps);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s bench|check <sequences.bin> ...\n", argv[0]); return 2; }
    try {
        Sequences q;
        if (!load(argv[2], q)) { std::fprintf(stderr, "cannot read '%s' (or no pinned memory: no CUDA device)\n", argv[2]); return 2; }
        if (!std::strcmp(argv[1], "check")) return run_check(q, argc > 3 ? std::atoi(argv[3]) : 100);
        if (argc < 6) { std::fprintf(stderr, "bench needs <n_estimators> <steps> <warmup>\n"); return 2; }
        return run_bench(q, std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), argc > 6 ? std::atoi(argv[6]) : 500, argc > 7 ? std::atoi(argv[7]) : 0);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
