// svo_internal.cuh -- device-side data layout shared by all kernels of the stereo-VO hot path.
//
// HBM layout (all arrays are pooled per context, indexed by slot / buffer / side):
//   slot  s in [0, max_seqs)   one CStereoOdometryEstimator's mutable state
//   buf   b in {0,1}           ping-pong frame storage: m_current_imgpair / m_prev_imgpair
//                              (libstereo-odometry.h:829); which one is "current" is SlotState.cur
//   side  d in {0,1}           left / right image
// Keypoints are kept as (y<<16|x, score) pairs, descriptors as 32-byte rows (two uint4), matches as
// int4 {queryIdx, trainIdx, distance, 0}; expansion to cv::KeyPoint / cv::DMatch layout happens only
// in the C-ABI readers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svo {

constexpr int kOrbEdge = 31;        // cv::ORB edgeThreshold (border filter inside ORB::compute)
constexpr int kMaxImgDim = 4096;    // x and y are packed in 16 bits each, grid occupancy bitmaps sized from it
constexpr int kTileW = 64, kTileH = 32;        // fast_blur_kernel's tile
constexpr int kTileCap = (kTileW / 2) * (kTileH / 2);   // strict 3x3 maxima: at most one per 2x2 block
constexpr int kBandH = 4;                                // rows of a tile that one warp of fast_blur_kernel lists (256 pixels)
constexpr int kBandCap = (kTileW / 2) * (kBandH / 2);   // list entries of one band (same 2x2 argument): 8 bands x 64 = kTileCap
__host__ __device__ inline int tiles_x_of(int W) { return (W + kTileW - 1) / kTileW; }
__host__ __device__ inline int tiles_y_of(int H) { return (H + kTileH - 1) / kTileH; }
constexpr int kMaxOctaves = 4;      // params_rectify.nOctaves (reference default 3, stage1_rectify.cpp:28)

struct SlotState {                  // mutable members of the estimator (libstereo-odometry.h:733-754)
    int32_t cur;                    // buffer index holding m_current_imgpair
    int32_t has_cur, has_prev;      // m_current_imgpair.present(), m_prev_imgpair.present()
    int32_t m_error;                // VOErrorCode kept across calls (process_new_image_pair.cpp:86-95)
    int32_t repeat;                 // request_data.repeat of the step being processed
    int32_t status;                 // svo_status of the step (capacity overflow ...)
    uint32_t it_counter;            // m_it_counter
    int32_t  reset_ids;             // m_reset: renumber the previous frame's match IDs at the next step (P:254-267)
    int32_t fast_th1, orb_th1;      // m_current_fast_th + 1 / m_current_orb_th + 1 once a setter ran (libstereo-odometry.h:529-541);
                                    // 0 = still the configured value (Params::fast_th / orb_th, H:661-662)
    long long last_match_id;        // m_last_match_ID
    long long last_kf_max_id;       // m_last_kf_max_id
    double last_pose[6];            // m_last_computed_pose
};

struct Camera { double fxl, fyl, cxl, cyl, fxr, fyr, cxr, cyr, baseline; };

struct Params {                     // device copy of svo_config (see include/svo_b200.h)
    int32_t W, H, max_seqs, max_corners, kp_cap;
    int32_t quota;                  // kps_to_detect[0] = 2*orb_nfeats for nOctaves == 1 (stage2_detect.cpp:405)
    int32_t fast_th, nms, min_distance;
    double  max_y_diff;
    int32_t orb_th, robust_1to1, fundmat_filter;
    int32_t use_robust_kernel;
    double  kernel_param;
    int32_t max_iters, initial_max_iters;
    double  min_mod_out_vector;
    int32_t max_incr_cost;
    double  residual_threshold;
    int32_t bad_tracking_th, use_prev_pose, use_ids, sort_mode;
    int32_t match_method;           // TSMMethod: 0 = smDescBF, 1 = smDescRbR (libstereo-odometry.h:454)
    int32_t ifm_method;             // TIFMMethod: 0 = ifmDescBF, 1 = ifmDescWin (libstereo-odometry.h:291)
    int32_t ifm_win_w, ifm_win_h;   // inter-frame search window (libstereo-odometry.h:294)
    int32_t rbr_max_distance;       // size_t(params_lr_match.orb_max_distance) (stage3_match_left_right.cpp:206)
    int32_t detect_method;          // TDMethod: 0 = dmORB (stage2_detect.cpp:458-496), 1 = dmFAST_ORB (:502-515)
};

// Device result record: same layout as svo_result in include/svo_b200.h (static_assert in capi.cu).
struct Result {
    double  out_pose[16];
    double  out_pose_rotvec[3];
    double  out_pose_t[3];
    double  out_pose_ypr[3];
    double  delta_pose[6];
    int32_t num_it, num_it_final;
    int32_t valid;
    int32_t error_code;
    int32_t tracked_kf, tracked_frame;
    int32_t detected_left, detected_right;
    int32_t stereo_matches;
    int32_t n_outliers;
    int32_t n_residuals;
    int32_t status;
};

// m_current_fast_th / m_current_orb_th of a slot: read at every call (S2:510, S3:164, S4:149)
__host__ __device__ inline int slot_fast_th(const Params& p, const SlotState& s) { return s.fast_th1 ? s.fast_th1 - 1 : p.fast_th; }
__host__ __device__ inline int slot_orb_th(const Params& p, const SlotState& s) { return s.orb_th1 ? s.orb_th1 - 1 : p.orb_th; }

// All device pointers of a context, passed by value to kernels.
struct Pools {
    Params p;
    SlotState* state;      // [max_seqs]
    Result* result;        // [max_seqs]
    Camera* cam;           // [max_seqs]
    // per (slot, side) scratch of one step
    uint8_t* blur;         // [max_seqs*2][H*W]   7x7 blurred image
    // FAST corners after 3x3 NMS and the 31-px border filter, as compact per-tile lists (tiles of kTileW x kTileH pixels,
    // fast_blur_kernel's CTA tile): no dense score plane is written or read
    uint16_t* tile_seg;    // [max_seqs*2][tiles_y*kTileH][tiles_x]   one word per SEGMENT (image row y inside tile column tx):
                           //                                 (first entry inside the band's list) << 8 | corners of the segment
    uint16_t* tile_list;   // [max_seqs*2][tiles][kTileCap]   (score << 8 | x in tile); band b at [b * kBandCap ..), rows top to
                           //                                 bottom, x ascending inside the band
    uint32_t* cand_pos;    // [max_seqs*2][max_corners]  y<<16|x of raw corners in row-major order
    int32_t* n_cand;       // [max_seqs*2]
    // per (slot, buf, side) frame data
    int2* kp;              // [max_seqs*2*2][kp_cap]  {y<<16|x, score}
    float4* kpf;           // dmORB only (else nullptr): [max_seqs*2*2][kp_cap] cv::KeyPoint {pt.x, pt.y, response, angle} -- the
                           // multi-scale detector yields float coordinates; kp then holds their truncated copy
    int32_t* kpo;          // dmORB only: [max_seqs*2*2][kp_cap] cv::KeyPoint::octave
    uint4* desc;           // [max_seqs*2*2][kp_cap][2]
    int32_t* n_kp;         // [max_seqs*2*2]
    // per (slot, buf) stereo matches
    int4* match;           // [max_seqs*2][kp_cap]  {queryIdx, trainIdx, distance, 0}
    int32_t* n_match;      // [max_seqs*2]
    uint4* mdesc;          // [max_seqs*2*2][kp_cap][2]  descriptor rows gathered per match (S4:105-131)
    long long* match_id;   // [max_seqs*2][kp_cap]
    // per slot scratch of one step
    uint32_t* lr_best;     // [max_seqs][kp_cap]   (distance<<16 | trainIdx) per left keypoint
    uint32_t* tr_best;     // [max_seqs*2][kp_cap] tracking: side 0 = left-left, 1 = right-right
    int4* tcand;           // [max_seqs][kp_cap]   after the greedy filter: {queryIdx, trainL, trainR, 0}
    int32_t* n_tcand;      // [max_seqs]
    uint8_t* inl;          // [max_seqs*2][kp_cap] RANSAC masks (left-left, right-right)
    int32_t* n_inl;        // [max_seqs*2]         inlier counts
    int2* tracked;         // [max_seqs][kp_cap]   {pre_match_idx, cur_match_idx}
    int32_t* n_tracked;    // [max_seqs]
    double* residual;      // [max_seqs][kp_cap]
    int32_t* outliers;     // [max_seqs][kp_cap]
    // batch description of the running step
    int32_t* slots;        // [max_seqs] slot index of batch item i
    // Octaves (params_rectify.nOctaves > 1): stages 2-4 run once per octave, each with its own Pools (own W, H, quota,
    // kp_cap and frame arrays); state / result / cam / slots are shared by the octaves of a slot.
    int32_t octave, n_octaves;
    int32_t* oct_stats;    // [max_seqs][kMaxOctaves][4] detected left, detected right, stereo matches, tracked pairs
};

// The arrays of all octaves of a context, for the kernels that cross octaves: the ID renumbering of
// process_new_image_pair.cpp:254-267 and the flattening of stage5_optimization.cpp:419-461.
struct OctaveList {
    int32_t n;
    int32_t cap[kMaxOctaves];
    const int2* kp[kMaxOctaves];
    const int4* match[kMaxOctaves];
    const int32_t* n_match[kMaxOctaves];
    long long* match_id[kMaxOctaves];
    const int2* tracked[kMaxOctaves];
    const int32_t* n_tracked[kMaxOctaves];
};

// Optional per-kernel timing: one CUDA event is recorded on the launching stream after every kernel
// of a step (bench.py derives the per-kernel roofline numbers from these).
struct StepTimer {
    cudaEvent_t* ev = nullptr;   // events of the running step, or nullptr when timing is off
    int next = 0;
    void mark(cudaStream_t st) { if (ev) cudaEventRecord(ev[next++], st); }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to (kernel, device) and is shared by every context of the process:
// it is only ever RAISED, so that creating a small context never takes shared memory away from a larger one that is
// still alive.  One table per kernel at the call site (SVO_RAISE_DYN_SMEM).
inline cudaError_t raise_dyn_smem(const void* func, int bytes, int* seen /* [64], zero-initialised */) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && bytes <= seen[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) seen[dev] = bytes;
    return e;
}
#define SVO_RAISE_DYN_SMEM(kernel, bytes) ([&]() { static int seen_[64]; return raise_dyn_smem((const void*)(kernel), (bytes), seen_); }())

__host__ __device__ inline size_t img_idx(int slot, int side) { return (size_t)slot * 2 + side; }
__host__ __device__ inline size_t frame_idx(int slot, int buf, int side) { return ((size_t)slot * 2 + buf) * 2 + side; }
__host__ __device__ inline size_t pair_idx(int slot, int buf) { return (size_t)slot * 2 + buf; }
#if defined(__CUDACC__)
// pt of keypoint i of the frame whose keypoints start at fbase = frame_idx(...) * kp_cap
__device__ __forceinline__ float2 kp_xy(const Pools& P, size_t fbase, int i) {
    if (P.kpf) { const float4 k = P.kpf[fbase + i]; return make_float2(k.x, k.y); }
    const int2 k = P.kp[fbase + i];
    return make_float2((float)(k.x & 0xFFFF), (float)((unsigned)k.x >> 16));
}
#endif
__host__ __device__ inline size_t stat_idx(int slot, int octave, int what) { return ((size_t)slot * kMaxOctaves + octave) * 4 + what; }

}  // namespace svo
