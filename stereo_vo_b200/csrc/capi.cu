// capi.cu -- the C-ABI of include/svo_b200.h: context, memory pools in HBM, stream-ordered
// pipeline of the hot path, state readers.  No CPU implementation of any stage exists in this
// library: every entry point needs a CUDA device and fails with SVO_ERR_CUDA otherwise.
#include "../../include/svo_b200.h"
#include "detect.cuh"
#include "gn.cuh"
#include "match.cuh"
#include "microbench.cuh"
#include "prepare.cuh"
#include "ransac.cuh"
#include "svo_internal.cuh"
#include "orb.cuh"

#include <nvtx3/nvToolsExt.h>   // header-only; ranges are no-ops unless a profiler (nsys / ncu --nvtx) is attached

#include <cmath>
#include <thread>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace svo;

static_assert(sizeof(svo_result) == sizeof(Result), "svo_result and svo::Result must have the same layout");
static_assert(sizeof(svo_camera) == sizeof(Camera), "svo_camera and svo::Camera must have the same layout");
static_assert(sizeof(svo_keypoint) == 28 && sizeof(svo_dmatch) == 16, "cv::KeyPoint / cv::DMatch layout");

static thread_local std::string g_create_error;

// one timed interval per kernel of a step, in launch order (see run_step)
enum { KT_COUNT = 12, KT_DEPTH = 64 };
static const char* kKernelNames[KT_COUNT] = {
    "begin_step", "fast_blur_kernel", "select_kernel", "brief_kernel", "hamming_kernel(lr)", "lr_filter_kernel",
    "hamming_kernel(track)", "track_filter_kernel", "ransac_prepare_kernel", "fundmat_kernel", "consistency_kernel",
    "gn_kernel"};

enum { kMaxChunks = 8, kDefaultChunks = 2, kMinChunkItems = 32 };
enum { kMetaRing = 4 };   // pinned staging sets for the batch description of svo_process_batch_device (see upload_batch_meta)
enum { kDepth = 3 };   // batches in flight (svo_submit_batch / svo_wait_batch): with three staging buffers the copy stream
                       // never waits for the host to see a batch finish before the next upload can start

// octave k >= 1 of a multi-octave context (octave 0 lives in svo_ctx::P / RS and reads the caller's images)
struct OctaveCtx {
    Pools P;
    RansacScratch RS;
    int W = 0, H = 0, pitch = 0;
    uint8_t* imgL = nullptr;      // [n_slots_alloc][pitch * H] pyramid planes of the running step
    uint8_t* imgR = nullptr;
};

struct svo_ctx {
    svo_config cfg;
    Pools P;
    GnScratch G;
    RansacScratch RS;
    // params_rectify.nOctaves > 1: per-octave pools, and the pools / scratch of the flattened stage-5 problem
    // (kp_cap = sum of the octaves' capacities).  With one octave PG == P and GG == G.
    int n_oct = 1;
    OctaveCtx oct[kMaxOctaves];
    Pools PG;
    GnScratch GG;
    OctaveList OL;
    OrbDev orb;                 // detect_method == 0 (dmORB): pyramid levels and scratch of cv::ORB::detectAndCompute
    cudaStream_t stream = nullptr;
    std::vector<void*> dev_allocs;
    std::vector<void*> host_allocs;
    int n_slots_alloc = 0;      // max_seqs + 1 (the last slot serves the per-stage entry points)
    int hidden_slot = 0;
    int pitch = 0;              // row pitch of the internal image planes (W rounded up to 16)
    uint8_t* d_stageL = nullptr;      // == d_stL[0] (used by the synchronous per-stage entries)
    uint8_t* d_stageR = nullptr;
    // kDepth-deep submit/wait pipeline: staging images, pinned meta/result rings, events
    cudaStream_t copy_stream = nullptr;
    enum { kCopyLanes = 4 };           // scattered uploads (one pointer per image) are spread over several streams: more copy
    cudaStream_t copy_lane[kCopyLanes] = {};   // descriptors in flight when the host side of the link is busy (8 ranks copying)
    cudaEvent_t ev_lane[kCopyLanes] = {}, ev_lane_go = nullptr;
    uint8_t* d_stL[kDepth] = {};
    uint8_t* d_stR[kDepth] = {};
    int32_t* r_slots[kDepth] = {};
    uint8_t* r_repeat[kDepth] = {};
    Camera* r_cams[kDepth] = {};
    Result* r_res[kDepth] = {};
    SlotState* r_state[kDepth] = {};
    // svo_set_result_detail(1): the variable-length members of TStereoOdometryResult travel with every batch
    int detail = 0;
    bool batch_copy_ok = true;         // cudaMemcpyBatchAsync accepted by this driver (scattered image uploads)
    uint8_t* h_stage[kDepth] = {};     // pinned staging for images that arrive in pageable host memory (allocated on demand)
    size_t h_stage_bytes = 0;
    double* r_resid[kDepth] = {};      // [n_slots][PG.kp_cap]
    int32_t* r_outl[kDepth] = {};      // [n_slots][PG.kp_cap]
    int32_t* r_stats[kDepth] = {};     // [n_slots][kMaxOctaves*4] per-octave counters (result.detected_feats / stereo_matches)
    cudaEvent_t ev_h2d[kDepth] = {}, ev_done[kDepth] = {};
    int pend_n[kDepth] = {}, pend_lo[kDepth] = {}, pend_hi[kDepth] = {};
    int pend_head = 0, pend_count = 0;
    uint8_t* d_repeat = nullptr;
    Camera* d_cams_batch = nullptr;
    uint32_t* d_occ = nullptr;
    int occ_words = 0;
    uint32_t* d_sort_scratch = nullptr;
    uint2* d_list_scratch = nullptr;
    uint32_t* d_key_scratch = nullptr;
    // pinned host staging of svo_process_batch_device's batch description: a ring, so that a call never overwrites
    // memory an earlier call's asynchronous copy has not read yet
    int32_t* m_slots[kMetaRing] = {};
    uint8_t* m_repeat[kMetaRing] = {};
    Camera* m_cams[kMetaRing] = {};
    cudaEvent_t m_ev[kMetaRing] = {};
    int m_next = 0;
    // pinned host staging (synchronous entry points)
    int32_t* h_slots = nullptr;
    uint8_t* h_repeat = nullptr;
    Camera* h_cams = nullptr;
    Result* h_results = nullptr;
    std::string err;
    int launches_i = 0;
    int64_t launches = 0;
    // stage 1 on the device (gray conversion / rectification): input format, maps, raw staging (allocated on demand)
    int in_channels = 1;
    bool rectify = false;
    short2* d_map_xy[2] = {nullptr, nullptr};
    uint16_t* d_map_a[2] = {nullptr, nullptr};
    uint8_t* d_raw[kDepth][2] = {};                                    // [buffer][side]
    uint8_t* d_gray[2] = {nullptr, nullptr};                           // gray planes between conversion and rectification
    size_t raw_pitch = 0;             // bytes per row of the raw staging images
    int raw_channels = 0;
    // concurrency inside one step: groups of batch items on their own streams (see run_step)
    int chunks = kDefaultChunks;
    cudaStream_t aux[kMaxChunks] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxChunks] = {};
    int timing = 0;
    std::vector<cudaEvent_t> ev;      // KT_DEPTH sets of KT_COUNT+1 events
    int timed_steps = 0;              // steps recorded since timing was enabled
    int marks_last = 0;
};

namespace {

// NVTX range with the section names of the reference's CTimeLogger (m_profiler.enter/leave: "_stg2" ... "_stg5",
// "processNewImagePair"; libstereo-odometry.h:521-524): the ranges cover the ENQUEUE of a stage's kernels on the host,
// profilers project them onto the kernels they launch.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

int fail(svo_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(svo_ctx* c, cudaError_t e, const char* what) {
    return fail(c, SVO_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail((c), e_, #call); } while (0)

template <class T>
cudaError_t dalloc(svo_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    e = cudaMemset(q, 0, count * sizeof(T) + 16);
    c->dev_allocs.push_back(q);
    *p = static_cast<T*>(q);
    return e;
}
template <class T>
cudaError_t halloc(svo_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMallocHost(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    memset(q, 0, count * sizeof(T));
    c->host_allocs.push_back(q);
    *p = static_cast<T*>(q);
    return e;
}

__global__ void scatter_cams_kernel(Pools P, const Camera* cams, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) P.cam[P.slots[i]] = cams[i];
}

// Batch description from pinned host memory (read by the kernel over PCIe, a few KB) into the device arrays.
// A kernel instead of cudaMemcpyAsync: small H2D copies issued on the compute stream would queue on the host->device
// copy engine BEHIND the image uploads of the following batches and delay this batch's kernels by a whole upload.
__global__ void upload_meta_kernel(int32_t* __restrict__ d_slots, Camera* __restrict__ d_cams, uint8_t* __restrict__ d_repeat,
                                   const int32_t* __restrict__ h_slots, const Camera* __restrict__ h_cams,
                                   const uint8_t* __restrict__ h_repeat, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d_slots[i] = h_slots[i];
    d_cams[i] = h_cams[i];
    if (h_repeat) d_repeat[i] = h_repeat[i];
}

__global__ void reset_slot_kernel(Pools P, int slot_lo, int slot_hi) {
    const int slot = slot_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (slot > slot_hi) return;
    SlotState s;
    memset(&s, 0, sizeof(s));
    P.state[slot] = s;
    Result r;
    memset(&r, 0, sizeof(r));
    P.result[slot] = r;
}

// setFASTThreshold / setORBThreshold (libstereo-odometry.h:529-541): only the two ints change, the sequence state stays
__global__ void set_thresholds_kernel(Pools P, int slot_lo, int slot_hi, int fast_th, int orb_th) {
    const int s = slot_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (s > slot_hi) return;
    if (fast_th >= 0) P.state[s].fast_th1 = fast_th + 1;
    if (orb_th >= 0) P.state[s].orb_th1 = orb_th + 1;
}

bool valid_slot(const svo_ctx* c, int slot) { return slot >= 0 && slot < c->cfg.max_seqs; }


// nOctaves > 1 (stage1_rectify.cpp:80-83, stage2_detect.cpp:423-624, stage3_match_left_right.cpp:87-177,
// stage4_match_consecutive.cpp:91-328): after the prev/cur shift and the cross-octave ID renumbering, stages 2-4 run
// octave by octave -- octave k > 0 on the image the pyramid kernel derives from octave k-1 -- with the match-ID
// generator and the tracking counters carried through the shared slot state in octave order; the last octave's
// consistency kernel applies the bad-tracking gate to the total; stage 5 then runs once on the flattened lists.
// P0 / RS0: octave 0's pools with the slot list already advanced to item a.
int enqueue_octaves(svo_ctx* c, const Pools& P0, const RansacScratch& RS0, int a, int n, const uint8_t* d_left,
                    const uint8_t* d_right, size_t stride, size_t image_pitch, cudaStream_t st, int& L) {
    CU(c, launch_reset_ids_octaves(P0, c->OL, n, st, &L));
    const size_t img2 = 2 * (size_t)a;
    const uint8_t* srcL = d_left + a * image_pitch;
    const uint8_t* srcR = d_right + a * image_pitch;
    size_t src_stride = stride, src_pitch = image_pitch;
    int sw = c->cfg.width, sh = c->cfg.height;
    static const char* const kOctaveRange[kMaxOctaves] = {"_stg2-4.oct=0", "_stg2-4.oct=1", "_stg2-4.oct=2", "_stg2-4.oct=3"};
    for (int k = 0; k < c->n_oct; ++k) {
        NvtxRange range(kOctaveRange[k]);
        Pools P = k == 0 ? P0 : c->oct[k].P;
        RansacScratch RS = k == 0 ? RS0 : c->oct[k].RS;
        P.slots = P0.slots;
        RS.slots = P0.slots;
        int pitch_k = c->pitch;
        if (k > 0) {
            OctaveCtx& o = c->oct[k];
            const size_t plane = (size_t)o.pitch * o.H;
            PyramidArgs A;
            A.srcL = srcL; A.srcR = srcR; A.src_stride = src_stride; A.src_image_pitch = src_pitch;
            A.dstL = o.imgL + a * plane; A.dstR = o.imgR + a * plane; A.dst_stride = (size_t)o.pitch; A.dst_image_pitch = plane;
            A.sw = sw; A.sh = sh;
            CU(c, launch_pyramid_level(A, n, pyramid_method(c->cfg.pyr_mode, sw), st, &L));
            srcL = A.dstL; srcR = A.dstR; src_stride = (size_t)o.pitch; src_pitch = plane;
            sw = o.W; sh = o.H;
            pitch_k = o.pitch;
        }
        CU(c, launch_detect(P, n, srcL, srcR, src_stride, src_pitch, pitch_k, c->d_occ + img2 * c->occ_words, c->occ_words,
                            c->d_sort_scratch + img2 * c->cfg.max_corners,
                            c->d_list_scratch + img2 * 3 * (c->cfg.max_corners / 16 + 2),
                            c->d_key_scratch + img2 * c->cfg.max_corners, st, &L, nullptr));
        CU(c, launch_match_lr(P, n, st, &L, nullptr));
        CU(c, launch_track_match(P, n, st, &L, nullptr));
        if (c->cfg.fundmat_filter) CU(c, launch_ransac(P, RS, n, st, &L, nullptr));
        CU(c, launch_consistency(P, n, st, &L, nullptr));
    }
    NvtxRange range("_stg5");
    Pools PG = c->PG;
    PG.slots = P0.slots;
    CU(c, launch_flatten_octaves(PG, c->GG, c->OL, n, st, &L));
    CU(c, launch_gn(PG, c->GG, n, GN_FLATTENED, /*custom pose=*/0, /*assume_fast_kps=*/0, st, &L, nullptr));
    return SVO_OK;
}

// enqueue the kernels of one pipeline step for the batch items [a, a + n) on stream st
int enqueue_items(svo_ctx* c, int a, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride, size_t image_pitch,
                  bool have_repeat, cudaStream_t st, int& L, StepTimer* t) {
    Pools P = c->P;
    P.slots += a;                                   // batch item i of this chunk = item a + i of the step
    RansacScratch RS = c->RS;
    RS.slots = P.slots;
    if (t) t->mark(st);
    scatter_cams_kernel<<<(n + 127) / 128, 128, 0, st>>>(P, c->d_cams_batch + a, n);
    CU(c, launch_begin_step(P, n, have_repeat ? c->d_repeat + a : nullptr, st));
    if (t) t->mark(st);
    L += 2;
    const size_t img2 = 2 * (size_t)a;              // scratch indexed by (item, side)
    if (c->n_oct > 1) return enqueue_octaves(c, P, RS, a, n, d_left, d_right, stride, image_pitch, st, L);
    {
        NvtxRange r("_stg2");
        if (c->cfg.detect_method == 0) {
            CU(c, launch_orb_detect(P, c->orb, n, d_left, d_right, stride, image_pitch, a, c->d_occ + img2 * c->occ_words, c->occ_words,
                                    c->d_sort_scratch + img2 * c->cfg.max_corners,
                                    c->d_list_scratch + img2 * 3 * (c->cfg.max_corners / 16 + 2), st, &L));
            if (t) { t->mark(st); t->mark(st); t->mark(st); }
        } else
        CU(c, launch_detect(P, n, d_left + a * image_pitch, d_right + a * image_pitch, stride, image_pitch, c->pitch,
                            c->d_occ + img2 * c->occ_words, c->occ_words, c->d_sort_scratch + img2 * c->cfg.max_corners,
                            c->d_list_scratch + img2 * 3 * (c->cfg.max_corners / 16 + 2),
                            c->d_key_scratch + img2 * c->cfg.max_corners, st, &L, t));
    }
    {
        NvtxRange r("_stg3");
        CU(c, launch_match_lr(P, n, st, &L, t));
    }
    {
        NvtxRange r("_stg4");
        CU(c, launch_track_match(P, n, st, &L, t));
        if (c->cfg.fundmat_filter) CU(c, launch_ransac(P, RS, n, st, &L, t));
        else if (t) { t->mark(st); t->mark(st); }
        CU(c, launch_consistency(P, n, st, &L, t));
    }
    NvtxRange r("_stg5");
    // dmORB keypoints of different levels can share a grid cell: the aux NMS of S5:465-474 really runs
    CU(c, launch_gn(P, c->G, n, GN_FROM_TRACKED, /*custom pose=*/0, /*assume_fast_kps=*/c->cfg.detect_method == 0 ? 0 : 1, st, &L, t));
    return SVO_OK;
}

// enqueue one pipeline step on device images.  Large batches are split into `chunks` groups of slots that run
// the same kernel sequence on their own streams (forked from / joined into the context's stream): the
// latency-bound kernels of one group (sort emulation, RANSAC, Gauss-Newton) then share the SMs with the
// issue-bound kernels of another (FAST, Hamming) instead of leaving them idle.  With per-kernel timing on, the
// step runs on the one stream so that every kernel is timed alone.
int run_step(svo_ctx* c, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride, size_t image_pitch,
             bool have_repeat) {
    NvtxRange range("processNewImagePair");
    int& L = c->launches_i;
    L = 0;
    StepTimer tm;
    if (c->timing && !c->ev.empty()) tm.ev = c->ev.data() + (size_t)(c->timed_steps % KT_DEPTH) * (KT_COUNT + 1);
    StepTimer* t = (tm.ev && c->n_oct == 1) ? &tm : nullptr;      // the per-kernel event set describes the one-octave step
    int chunks = std::min(c->chunks, n / kMinChunkItems);
    if (t || chunks < 2) {
        int rc = enqueue_items(c, 0, n, d_left, d_right, stride, image_pitch, have_repeat, c->stream, L, t);
        if (rc != SVO_OK) return rc;
        if (t) { c->marks_last = tm.next; ++c->timed_steps; }
    } else {
        CU(c, cudaEventRecord(c->ev_fork, c->stream));
        for (int j = 0; j < chunks; ++j) {
            const int a = (int)((long long)n * j / chunks), b = (int)((long long)n * (j + 1) / chunks);
            CU(c, cudaStreamWaitEvent(c->aux[j], c->ev_fork, 0));
            int rc = enqueue_items(c, a, b - a, d_left, d_right, stride, image_pitch, have_repeat, c->aux[j], L, nullptr);
            if (rc != SVO_OK) return rc;
            CU(c, cudaEventRecord(c->ev_join[j], c->aux[j]));
            CU(c, cudaStreamWaitEvent(c->stream, c->ev_join[j], 0));
        }
    }
    c->launches += L;
    return SVO_OK;
}

int upload_batch_meta(svo_ctx* c, int n, const int32_t* slots, const svo_camera* cams, const uint8_t* repeat) {
    if (n <= 0 || n > c->cfg.max_seqs) return fail(c, SVO_ERR_ARG, "batch size out of range");
    if (!cams) return fail(c, SVO_ERR_ARG, "cams is NULL");
    std::vector<char> seen(c->cfg.max_seqs, 0);
    for (int i = 0; i < n; ++i) {
        const int s = slots ? slots[i] : i;
        if (!valid_slot(c, s)) return fail(c, SVO_ERR_ARG, "slot index out of range");
        if (seen[s]) return fail(c, SVO_ERR_ARG, "duplicate slot in one batch");
        seen[s] = 1;
    }
    // The copies below are asynchronous (pinned memory): the staging set must stay untouched until they have run.
    // Calls may follow each other without a synchronisation (bench.py enqueues hundreds of steps), so each call takes
    // the next set of a small ring and first waits for the copies that last used it (an event that was never
    // recorded counts as complete).
    const int r = c->m_next;
    c->m_next = (r + 1) % kMetaRing;
    CU(c, cudaEventSynchronize(c->m_ev[r]));
    for (int i = 0; i < n; ++i) {
        c->m_slots[r][i] = slots ? slots[i] : i;
        memcpy(&c->m_cams[r][i], &cams[i], sizeof(Camera));
        c->m_repeat[r][i] = repeat ? repeat[i] : 0;
    }
    CU(c, cudaMemcpyAsync(c->P.slots, c->m_slots[r], n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->d_cams_batch, c->m_cams[r], n * sizeof(Camera), cudaMemcpyHostToDevice, c->stream));
    if (repeat) CU(c, cudaMemcpyAsync(c->d_repeat, c->m_repeat[r], n, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaEventRecord(c->m_ev[r], c->stream));
    return SVO_OK;
}

// make `slot` the only batch item (used by readers / per-stage entries that launch kernels)
int set_single_slot(svo_ctx* c, int slot) {
    c->h_slots[0] = slot;
    CU(c, cudaMemcpyAsync(c->P.slots, c->h_slots, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    return SVO_OK;
}

int get_state(svo_ctx* c, int slot, SlotState* s) {
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(s, c->P.state + slot, sizeof(SlotState), cudaMemcpyDeviceToHost));
    return SVO_OK;
}

int frame_buf(const SlotState& s, int which) { return which == 0 ? s.cur : (s.cur ^ 1); }

bool needs_stage1(const svo_ctx* c) { return c->in_channels != 1 || c->rectify; }

// raw staging for inputs that go through stage 1 on the device (colour and / or unrectified images)
int ensure_raw_staging(svo_ctx* c) {
    if (c->d_raw[0][0] && c->raw_channels >= c->in_channels) return SVO_OK;
    for (int b = 0; b < kDepth; ++b) for (int sd = 0; sd < 2; ++sd) c->d_raw[b][sd] = nullptr;   // old buffers stay in dev_allocs
    c->raw_channels = c->in_channels;
    c->raw_pitch = ((size_t)c->cfg.width * c->raw_channels + 15) & ~(size_t)15;
    for (int b = 0; b < kDepth; ++b)
        for (int sd = 0; sd < 2; ++sd) {
            cudaError_t e = dalloc(c, &c->d_raw[b][sd], (size_t)c->n_slots_alloc * c->raw_pitch * c->cfg.height);
            if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc raw staging");
        }
    for (int sd = 0; sd < 2 && c->raw_channels == 3 && !c->d_gray[sd]; ++sd) {
        cudaError_t e = dalloc(c, &c->d_gray[sd], (size_t)c->n_slots_alloc * c->pitch * c->cfg.height);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc gray planes");
    }
    return SVO_OK;
}

PrepareArgs prepare_args(const svo_ctx* c, const uint8_t* srcL, const uint8_t* srcR, size_t stride, size_t image_pitch,
                         uint8_t* dstL, uint8_t* dstR) {
    PrepareArgs A;
    A.srcL = srcL; A.srcR = srcR; A.src_stride = stride; A.src_image_pitch = image_pitch;
    A.dstL = dstL; A.dstR = dstR; A.dst_image_pitch = (size_t)c->pitch * c->cfg.height;
    A.W = c->cfg.width; A.H = c->cfg.height; A.pitch_out = c->pitch;
    for (int sd = 0; sd < 2; ++sd) { A.map_xy[sd] = c->d_map_xy[sd]; A.map_a[sd] = c->d_map_a[sd]; }
    A.tmpL = c->d_gray[0]; A.tmpR = c->d_gray[1];     // room for max_seqs + 1 images (callers with more pass NULL)
    return A;
}

// kps_to_detect[] of stage2_detect.cpp:404-407
void octave_quotas(int orb_nfeats, int n_oct, int* q) {
    q[0] = (int)(size_t)((double)((size_t)orb_nfeats * (2 * (size_t)n_oct)) / (std::pow(2.0, n_oct) - 1.0));
    for (int k = 1; k < n_oct; ++k) q[k] = (int)(size_t)std::round((double)q[0] / std::pow(2.0, k));
}

// the per-(slot, buffer, side) arrays of one octave k >= 1 (octave 0's are allocated in svo_create)
cudaError_t alloc_octave_arrays(svo_ctx* c, Pools& P, size_t S, size_t plane, size_t cap) {
    cudaError_t e;
#define OA(ptr, count) do { if ((e = dalloc(c, &(ptr), (count))) != cudaSuccess) return e; } while (0)
    OA(P.blur, S * 2 * plane);
    OA(P.tile_seg, S * 2 * (size_t)tiles_x_of(P.p.W) * tiles_y_of(P.p.H) * kTileH);
    OA(P.tile_list, S * 2 * (size_t)tiles_x_of(P.p.W) * tiles_y_of(P.p.H) * kTileCap);
    OA(P.cand_pos, S * 2 * (size_t)P.p.max_corners); OA(P.n_cand, S * 2);
    OA(P.kp, S * 4 * cap); OA(P.desc, S * 4 * cap * 2); OA(P.n_kp, S * 4);
    OA(P.match, S * 2 * cap); OA(P.n_match, S * 2); OA(P.mdesc, S * 4 * cap * 2); OA(P.match_id, S * 2 * cap);
    OA(P.lr_best, S * cap); OA(P.tr_best, S * 2 * cap); OA(P.tcand, S * cap); OA(P.n_tcand, S);
    OA(P.inl, S * 2 * cap); OA(P.n_inl, S * 2); OA(P.tracked, S * cap); OA(P.n_tracked, S);
#undef OA
    return cudaSuccess;
}

const Pools* octave_pools(const svo_ctx* c, int octave) {
    if (octave < 0 || octave >= c->n_oct) return nullptr;
    return octave == 0 ? &c->P : &c->oct[octave].P;
}

}  // namespace

extern "C" {

int svo_abi_version(void) { return SVO_B200_ABI_VERSION; }

void svo_default_config(svo_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = SVO_B200_ABI_VERSION;
    cfg->device = 0;
    cfg->width = 640;
    cfg->height = 480;
    cfg->max_seqs = 1;
    cfg->max_corners = 16384;
    cfg->orb_nfeats = 500;               // TDetectParams ctor, stage2_detect.cpp:44-58
    cfg->fast_threshold = 20;
    cfg->non_maximal_suppression = 1;
    cfg->min_distance = 3;
    cfg->max_y_diff = 1.0;               // SURVEY Appendix C (ctor default 0, stage3_match_left_right.cpp:54)
    cfg->orb_threshold = 40;             // orb_max_distance, stage3_match_left_right.cpp:50
    cfg->enable_robust_1to1_match = 0;
    cfg->fundmat_filter = 1;
    cfg->use_robust_kernel = 1;          // TLeastSquaresParams ctor, common.cpp:69-82
    cfg->kernel_param = 3.0;
    cfg->max_iters = 100;
    cfg->initial_max_iters = 10;
    cfg->min_mod_out_vector = 1e-3;
    cfg->max_incr_cost = 3;
    cfg->residual_threshold = 10.0;
    cfg->bad_tracking_th = 5;
    cfg->use_previous_pose_as_initial = 1;
    cfg->vo_use_matches_ids = 0;
    cfg->sort_mode = 0;
    cfg->match_method = 0;               // smDescBF
    cfg->if_match_method = 0;            // ifmDescBF
    cfg->ifm_win_w = 20;
    cfg->ifm_win_h = 20;
    cfg->orb_max_distance = 40.0;
    cfg->n_octaves = 1;                  // SURVEY Appendix C (the reference's ctor default is 3, stage1_rectify.cpp:28)
    cfg->pyr_mode = 0;                   // as MRPT picks: 2x2 average when width % 16 == 0, cvPyrDown otherwise
    cfg->detect_method = 1;              // dmFAST_ORB, the named path (the reference's ctor default is dmFASTER, unimplemented there)
    cfg->orb_nlevels = 8;                // TDetectParams ctor (libstereo-odometry.h:397)
}

const char* svo_last_error(const svo_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int svo_create(const svo_config* cfg, svo_ctx** out) {
    if (!cfg || !out) return fail(nullptr, SVO_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (cfg->abi_version != SVO_B200_ABI_VERSION) return fail(nullptr, SVO_ERR_ARG, "ABI version mismatch");
    if (cfg->width < 64 || cfg->height < 64 || cfg->width > kMaxImgDim || cfg->height > kMaxImgDim)
        return fail(nullptr, SVO_ERR_ARG, "image size must be within [64, 4096]");
    if (cfg->max_seqs < 1 || cfg->max_seqs > 65536) return fail(nullptr, SVO_ERR_ARG, "max_seqs out of range");
    if (cfg->orb_nfeats < 1) return fail(nullptr, SVO_ERR_ARG, "orb_nfeats must be positive");
    if (cfg->min_distance < 2) return fail(nullptr, SVO_ERR_ARG, "min_distance must be >= 2 (the reference divides by uint(min_distance/2))");
    if (cfg->fast_threshold < 1 || cfg->fast_threshold > 254) return fail(nullptr, SVO_ERR_ARG, "fast_threshold out of range");
    if (cfg->max_corners < 16) return fail(nullptr, SVO_ERR_ARG, "max_corners too small");
    if ((cfg->match_method | 1) != 1) return fail(nullptr, SVO_ERR_ARG, "Stereo matching method not correct (smDescBF = 0 or smDescRbR = 1)");
    if ((cfg->if_match_method | 1) != 1) return fail(nullptr, SVO_ERR_ARG, "Interframe matching method not correct (ifmDescBF = 0 or ifmDescWin = 1)");
    if (cfg->ifm_win_w < 0 || cfg->ifm_win_h < 0 || cfg->ifm_win_w > kMaxImgDim || cfg->ifm_win_h > kMaxImgDim)
        return fail(nullptr, SVO_ERR_ARG, "ifm_win_w / ifm_win_h out of range");
    if (!(cfg->orb_max_distance >= 0 && cfg->orb_max_distance < 65536)) return fail(nullptr, SVO_ERR_ARG, "orb_max_distance out of range");
    const int n_oct = cfg->n_octaves;
    if (n_oct < 1 || n_oct > kMaxOctaves) return fail(nullptr, SVO_ERR_ARG, "n_octaves must be within [1, 4]");
    if (cfg->pyr_mode < 0 || cfg->pyr_mode > 2) return fail(nullptr, SVO_ERR_ARG, "pyr_mode must be 0 (as MRPT), 1 (2x2 average) or 2 (pyrDown)");
    if ((cfg->width >> (n_oct - 1)) < 64 || (cfg->height >> (n_oct - 1)) < 64)
        return fail(nullptr, SVO_ERR_ARG, "the smallest octave must be at least 64 x 64");
    if ((cfg->detect_method | 1) != 1) return fail(nullptr, SVO_ERR_ARG, "Unknown detection method (dmORB = 0 or dmFAST_ORB = 1)");
    if (cfg->detect_method == 0) {
        if (n_oct != 1) return fail(nullptr, SVO_ERR_ARG, "dmORB builds its own pyramid: n_octaves must be 1 (stage2_detect.cpp:460)");
        if (cfg->orb_nlevels < 1 || cfg->orb_nlevels > kOrbMaxLevels) return fail(nullptr, SVO_ERR_ARG, "orb_nlevels must be within [1, 8]");
        if (cfg->sort_mode != 0) return fail(nullptr, SVO_ERR_ARG, "sort_mode 1 is not available with dmORB");
        {   // the sort scratch of the context (sized by max_corners) must hold detectAndCompute's whole keypoint list
            const int n_extract = cfg->non_maximal_suppression ? (int)(size_t)(1.5 * cfg->orb_nfeats) : cfg->orb_nfeats;
            if (cfg->max_corners < 2 * n_extract + 64) return fail(nullptr, SVO_ERR_ARG, "max_corners too small for dmORB (needs 2 x extracted features + 64)");
        }
        const OrbLevels G = orb_levels(cfg->width, cfg->height, cfg->orb_nlevels, 1);
        if (G.w[cfg->orb_nlevels - 1] < 64 || G.h[cfg->orb_nlevels - 1] < 64)
            return fail(nullptr, SVO_ERR_ARG, "the smallest ORB level must be at least 64 x 64 (lower orb_nlevels)");
    }
    int quotas[kMaxOctaves];
    octave_quotas(cfg->orb_nfeats, n_oct, quotas);    // kps_to_detect[], stage2_detect.cpp:404-407
    const int quota = quotas[0];                      // 2 * orb_nfeats for nOctaves = 1
    if (quota < 1 || quotas[n_oct - 1] < 1) return fail(nullptr, SVO_ERR_ARG, "orb_nfeats too small for n_octaves");
    const int kp_cap = cfg->non_maximal_suppression ? std::min(quota, cfg->max_corners) : cfg->max_corners;
    if (kp_cap > 65535) return fail(nullptr, SVO_ERR_ARG, "more than 65535 keypoints per image are not supported");
    if (cfg->max_corners >= (1 << 24)) return fail(nullptr, SVO_ERR_ARG, "max_corners too large");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SVO_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                               " (this library has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, SVO_ERR_ARG, "device ordinal out of range");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");

    svo_ctx* c = new (std::nothrow) svo_ctx();
    if (!c) return fail(nullptr, SVO_ERR_ARG, "out of host memory");
    c->cfg = *cfg;
    Params& p = c->P.p;
    p.W = cfg->width; p.H = cfg->height; p.max_seqs = cfg->max_seqs; p.max_corners = cfg->max_corners;
    p.kp_cap = kp_cap; p.quota = quota; p.fast_th = cfg->fast_threshold; p.nms = cfg->non_maximal_suppression;
    p.min_distance = cfg->min_distance; p.max_y_diff = cfg->max_y_diff; p.orb_th = cfg->orb_threshold;
    p.robust_1to1 = cfg->enable_robust_1to1_match; p.fundmat_filter = cfg->fundmat_filter;
    p.use_robust_kernel = cfg->use_robust_kernel; p.kernel_param = cfg->kernel_param; p.max_iters = cfg->max_iters;
    p.initial_max_iters = cfg->initial_max_iters; p.min_mod_out_vector = cfg->min_mod_out_vector;
    p.max_incr_cost = cfg->max_incr_cost; p.residual_threshold = cfg->residual_threshold;
    p.bad_tracking_th = cfg->bad_tracking_th; p.use_prev_pose = cfg->use_previous_pose_as_initial;
    p.use_ids = cfg->vo_use_matches_ids; p.sort_mode = cfg->sort_mode;
    p.match_method = cfg->match_method; p.ifm_method = cfg->if_match_method;
    p.ifm_win_w = cfg->ifm_win_w; p.ifm_win_h = cfg->ifm_win_h;
    p.rbr_max_distance = (int32_t)(size_t)cfg->orb_max_distance;
    p.detect_method = cfg->detect_method;

    const int S = cfg->max_seqs + 1;
    c->n_slots_alloc = S;
    c->hidden_slot = cfg->max_seqs;
    c->pitch = (cfg->width + 15) & ~15;
    const size_t plane = (size_t)c->pitch * cfg->height;
    const size_t cap = (size_t)kp_cap;
    Pools& P = c->P;
#define DA(ptr, count) do { cudaError_t e2 = dalloc(c, &(ptr), (count)); if (e2 != cudaSuccess) { int rc = cuda_fail(nullptr, e2, "cudaMalloc " #ptr); svo_destroy(c); return rc; } } while (0)
#define HA(ptr, count) do { cudaError_t e2 = halloc(c, &(ptr), (count)); if (e2 != cudaSuccess) { int rc = cuda_fail(nullptr, e2, "cudaMallocHost " #ptr); svo_destroy(c); return rc; } } while (0)
    DA(P.state, S); DA(P.result, S); DA(P.cam, S);
    DA(P.blur, S * 2 * plane);
    DA(P.tile_seg, S * 2 * (size_t)tiles_x_of(cfg->width) * tiles_y_of(cfg->height) * kTileH);
    DA(P.tile_list, S * 2 * (size_t)tiles_x_of(cfg->width) * tiles_y_of(cfg->height) * kTileCap);
    DA(P.cand_pos, (size_t)S * 2 * cfg->max_corners); DA(P.n_cand, S * 2);
    DA(P.kp, S * 4 * cap); DA(P.desc, S * 4 * cap * 2); DA(P.n_kp, S * 4);
    P.kpf = nullptr; P.kpo = nullptr;
    if (cfg->detect_method == 0) {
        DA(P.kpf, S * 4 * cap); DA(P.kpo, S * 4 * cap);
        // S2:460-464: more features when the NMS thins them out afterwards
        const int n_extract = cfg->non_maximal_suppression ? (int)(size_t)(1.5 * cfg->orb_nfeats) : cfg->orb_nfeats;
        cudaError_t e2 = alloc_orb(c->orb, p, S, cfg->orb_nlevels, n_extract, c->dev_allocs);
        if (e2 != cudaSuccess) { int rc = cuda_fail(nullptr, e2, "alloc_orb"); svo_destroy(c); return rc; }
    }
    DA(P.match, S * 2 * cap); DA(P.n_match, S * 2); DA(P.mdesc, S * 4 * cap * 2); DA(P.match_id, S * 2 * cap);
    DA(P.lr_best, S * cap); DA(P.tr_best, S * 2 * cap); DA(P.tcand, S * cap); DA(P.n_tcand, S);
    DA(P.inl, S * 2 * cap); DA(P.n_inl, S * 2); DA(P.tracked, S * cap); DA(P.n_tracked, S);
    DA(P.residual, S * cap); DA(P.outliers, S * cap); DA(P.slots, S);
    DA(P.oct_stats, (size_t)S * kMaxOctaves * 4);
    P.octave = 0; P.n_octaves = n_oct;
    c->n_oct = n_oct;
    for (int b = 0; b < kDepth; ++b) { DA(c->d_stL[b], (size_t)S * plane); DA(c->d_stR[b], (size_t)S * plane); }
    c->d_stageL = c->d_stL[0]; c->d_stageR = c->d_stR[0];
    DA(c->d_repeat, S); DA(c->d_cams_batch, S);
    // occupancy grids (only touched when uint(min_distance/2) != 1): worst case 1-px... cells >= 2 px
    {
        const unsigned cell = (unsigned)(cfg->min_distance / 2.0);
        const float inv = 1.0f / (float)cell;
        const unsigned glx = (unsigned)(1 + cfg->width * inv), gly = (unsigned)(1 + cfg->height * inv);
        c->occ_words = (int)((glx * gly + 31) / 32) + 1;
    }
    DA(c->d_occ, (size_t)S * 2 * c->occ_words);
    DA(c->d_sort_scratch, (size_t)S * 2 * cfg->max_corners);
    DA(c->d_list_scratch, (size_t)S * 2 * 3 * (cfg->max_corners / 16 + 2));
    DA(c->d_key_scratch, (size_t)S * 2 * cfg->max_corners);
    GnScratch& G = c->G;
    DA(G.prev, S * cap); DA(G.obs, S * cap); DA(G.lmk, S * cap * 3); DA(G.mask, S * cap); DA(G.mask0, S * cap);
    DA(G.cur_idx, S * cap); DA(G.ini_pose, S * 6);
    G.residual = P.residual;
    // own grid, indexed by slot: stage 5 of one group of slots may overlap stage 2 of another group (indexed by batch item)
    DA(G.occ, (size_t)S * 2 * c->occ_words);
    G.occ_words = 2 * c->occ_words;
    {
        cudaError_t e2 = alloc_ransac(c->RS, S, kp_cap, c->dev_allocs);
        if (e2 != cudaSuccess) { int rc = cuda_fail(nullptr, e2, "alloc_ransac"); svo_destroy(c); return rc; }
        c->RS.slots = P.slots; c->RS.inl = P.inl; c->RS.n_inl = P.n_inl;
    }
    // octaves 1..n-1: own frame arrays and pyramid planes; state / result / cam / slots / stats are shared with octave 0
    c->PG = c->P; c->GG = c->G;
    {
        OctaveList& OL = c->OL;
        memset(&OL, 0, sizeof(OL));
        OL.n = n_oct;
        size_t total_cap = cap;
        for (int k = 1; k < n_oct; ++k) {
            OctaveCtx& o = c->oct[k];
            o.W = cfg->width >> k; o.H = cfg->height >> k; o.pitch = (o.W + 15) & ~15;
            o.P = c->P;
            o.P.octave = k;
            o.P.p.W = o.W; o.P.p.H = o.H; o.P.p.quota = quotas[k];
            o.P.p.kp_cap = cfg->non_maximal_suppression ? std::min(quotas[k], cfg->max_corners) : cfg->max_corners;
            o.P.residual = nullptr; o.P.outliers = nullptr;        // stage 5 runs on the flattened problem only
            const size_t plane_k = (size_t)o.pitch * o.H;
            cudaError_t e2 = alloc_octave_arrays(c, o.P, S, plane_k, (size_t)o.P.p.kp_cap);
            if (e2 == cudaSuccess) e2 = dalloc(c, &o.imgL, (size_t)S * plane_k);
            if (e2 == cudaSuccess) e2 = dalloc(c, &o.imgR, (size_t)S * plane_k);
            if (e2 != cudaSuccess) { int rc = cuda_fail(nullptr, e2, "cudaMalloc octave arrays"); svo_destroy(c); return rc; }
            o.RS = c->RS;
            o.RS.kp_cap = o.P.p.kp_cap; o.RS.inl = o.P.inl; o.RS.n_inl = o.P.n_inl;
            total_cap += (size_t)o.P.p.kp_cap;
        }
        for (int k = 0; k < n_oct; ++k) {
            const Pools& Pk = *octave_pools(c, k);
            OL.cap[k] = Pk.p.kp_cap; OL.kp[k] = Pk.kp; OL.match[k] = Pk.match; OL.n_match[k] = Pk.n_match;
            OL.match_id[k] = Pk.match_id; OL.tracked[k] = Pk.tracked; OL.n_tracked[k] = Pk.n_tracked;
        }
        if (n_oct > 1) {
            if (total_cap > 65535) { svo_destroy(c); return fail(nullptr, SVO_ERR_ARG, "more than 65535 keypoints over all octaves are not supported"); }
            Pools& PG = c->PG;
            GnScratch& GG = c->GG;
            PG.p.kp_cap = (int)total_cap;
            PG.tracked = nullptr;                      // GN_FLATTENED reads GnScratch::cur_idx instead
            DA(PG.n_tracked, S); DA(PG.residual, S * total_cap); DA(PG.outliers, S * total_cap);
            DA(GG.prev, S * total_cap); DA(GG.obs, S * total_cap); DA(GG.lmk, S * total_cap * 3); DA(GG.mask, S * total_cap); DA(GG.mask0, S * total_cap);
            DA(GG.cur_idx, S * total_cap); DA(GG.ini_pose, S * 6);
            DA(GG.occ, (size_t)S * 2 * c->occ_words);      // own grid: stage 5 of one group may overlap stage 2 of another
            GG.residual = PG.residual;
        }
    }
    HA(c->h_slots, S); HA(c->h_repeat, S); HA(c->h_cams, S); HA(c->h_results, S);
    for (int b = 0; b < kDepth; ++b) { HA(c->r_slots[b], S); HA(c->r_repeat[b], S); HA(c->r_cams[b], S); HA(c->r_res[b], S); HA(c->r_state[b], S); }
    for (int r = 0; r < kMetaRing; ++r) { HA(c->m_slots[r], S); HA(c->m_repeat[r], S); HA(c->m_cams[r], S); }
#undef DA
#undef HA
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < svo_ctx::kCopyLanes && e == cudaSuccess; ++k) {
        e = cudaStreamCreateWithFlags(&c->copy_lane[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_lane[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_lane_go, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (int j = 0; j < kMaxChunks && e == cudaSuccess; ++j) {
        e = cudaStreamCreateWithFlags(&c->aux[j], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join[j], cudaEventDisableTiming);
    }
    for (int r = 0; r < kMetaRing && e == cudaSuccess; ++r) e = cudaEventCreateWithFlags(&c->m_ev[r], cudaEventDisableTiming);
    for (int b = 0; b < kDepth && e == cudaSuccess; ++b) {
        e = cudaEventCreateWithFlags(&c->ev_h2d[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { int rc = cuda_fail(nullptr, e, "cudaStreamCreate"); svo_destroy(c); return rc; }
    if ((e = configure_detect(p)) != cudaSuccess || (e = configure_match(p)) != cudaSuccess ||
        (e = configure_gn(c->PG.p)) != cudaSuccess || (e = configure_ransac(kp_cap)) != cudaSuccess) {
        int rc = cuda_fail(nullptr, e, "kernel configuration (shared memory)");
        svo_destroy(c);
        return rc;
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { int rc = cuda_fail(nullptr, e, "cudaDeviceSynchronize"); svo_destroy(c); return rc; }
    *out = c;
    return SVO_OK;
}

void svo_destroy(svo_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (int k = 0; k < svo_ctx::kCopyLanes; ++k) {
        if (c->copy_lane[k]) { cudaStreamSynchronize(c->copy_lane[k]); cudaStreamDestroy(c->copy_lane[k]); }
        if (c->ev_lane[k]) cudaEventDestroy(c->ev_lane[k]);
    }
    if (c->ev_lane_go) cudaEventDestroy(c->ev_lane_go);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (int b = 0; b < kDepth; ++b) { if (c->ev_h2d[b]) cudaEventDestroy(c->ev_h2d[b]); if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]); }
    for (int j = 0; j < kMaxChunks; ++j) {
        if (c->aux[j]) { cudaStreamSynchronize(c->aux[j]); cudaStreamDestroy(c->aux[j]); }
        if (c->ev_join[j]) cudaEventDestroy(c->ev_join[j]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int r = 0; r < kMetaRing; ++r) if (c->m_ev[r]) cudaEventDestroy(c->m_ev[r]);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    for (void* p : c->dev_allocs) cudaFree(p);
    for (void* p : c->host_allocs) cudaFreeHost(p);
    for (int k = 0; k < kDepth; ++k) if (c->h_stage[k]) cudaFreeHost(c->h_stage[k]);
    delete c;
}

int svo_reset_slot(svo_ctx* c, int32_t slot) {
    if (!c) return SVO_ERR_ARG;
    if (slot != -1 && !valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "slot index out of range (-1 = every slot)");
    const int lo = slot < 0 ? 0 : slot, hi = slot < 0 ? c->cfg.max_seqs - 1 : slot;
    reset_slot_kernel<<<(hi - lo + 128) / 128, 128, 0, c->stream>>>(c->P, lo, hi);
    CU(c, cudaGetLastError());
    return SVO_OK;
}

int svo_set_thresholds(svo_ctx* c, int32_t slot, int32_t fast_th, int32_t orb_th) {
    if (!c) return SVO_ERR_ARG;
    if (slot != -1 && !valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "slot index out of range (-1 = every slot)");
    if (fast_th == 0 || fast_th > 254) return fail(c, SVO_ERR_ARG, "fast_th must be within [1, 254] (negative: keep)");
    if (orb_th > 0x7FFE) return fail(c, SVO_ERR_ARG, "orb_th out of range (negative: keep)");
    CU(c, cudaSetDevice(c->cfg.device));
    const int lo = slot < 0 ? 0 : slot, hi = slot < 0 ? c->cfg.max_seqs - 1 : slot;
    set_thresholds_kernel<<<(hi - lo + 128) / 128, 128, 0, c->stream>>>(c->P, lo, hi, fast_th, orb_th);   // stream-ordered
    CU(c, cudaGetLastError());
    return SVO_OK;
}

int svo_get_thresholds(svo_ctx* c, int32_t slot, int32_t* fast_th, int32_t* orb_th) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "slot index out of range");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (fast_th) *fast_th = slot_fast_th(c->P.p, s);
    if (orb_th) *orb_th = slot_orb_th(c->P.p, s);
    return SVO_OK;
}

// page-locked host memory for the image buffers handed to svo_submit_batch (pageable memory works too, but the copies
// then go through the driver's staging buffers and do not overlap)
void* svo_host_alloc(size_t bytes) {
    void* p = nullptr;
    return cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
}
void svo_host_free(void* p) { if (p) cudaFreeHost(p); }

int svo_synchronize(svo_ctx* c) {
    if (!c) return SVO_ERR_ARG;
    CU(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

void* svo_stream(svo_ctx* c) { return c ? (void*)c->stream : nullptr; }

int svo_set_concurrency(svo_ctx* c, int32_t chunks) {
    if (!c) return SVO_ERR_ARG;
    if (chunks < 1 || chunks > kMaxChunks) return fail(c, SVO_ERR_ARG, "chunks must be within [1, 8]");
    c->chunks = chunks;
    return SVO_OK;
}
int64_t svo_kernel_launches(const svo_ctx* c) { return c ? c->launches : 0; }

int svo_measure_pipe_peaks(svo_ctx* c, double* popc_per_s, double* lop3_per_s) {
    if (!c || !popc_per_s || !lop3_per_s) return SVO_ERR_ARG;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, measure_pipe_peaks(popc_per_s, lop3_per_s, c->stream));
    return SVO_OK;
}

int svo_enable_stage_timing(svo_ctx* c, int32_t on) {
    if (!c) return SVO_ERR_ARG;
    if (on && c->ev.empty()) {
        c->ev.resize((size_t)KT_DEPTH * (KT_COUNT + 1));
        for (cudaEvent_t& e : c->ev) CU(c, cudaEventCreate(&e));
    }
    c->timing = on;
    c->timed_steps = 0;
    return SVO_OK;
}

// mean device time per kernel over the steps recorded since timing was enabled (at most the last KT_DEPTH)
int svo_get_stage_times(svo_ctx* c, const char** names, float* ms, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (c->ev.empty() || c->timed_steps == 0) return 0;
    CU(c, cudaStreamSynchronize(c->stream));
    const int steps = std::min(c->timed_steps, (int)KT_DEPTH);
    int n = 0;
    for (int i = 0; i < KT_COUNT && n < cap; ++i) {
        double acc = 0;
        for (int s = 0; s < steps; ++s) {
            const cudaEvent_t* e = c->ev.data() + (size_t)s * (KT_COUNT + 1);
            float t = 0;
            if (cudaEventElapsedTime(&t, e[i], e[i + 1]) == cudaSuccess) acc += t;
        }
        if (names) names[n] = kKernelNames[i];
        if (ms) ms[n] = (float)(acc / steps);
        ++n;
    }
    return n;
}

int svo_set_input_format(svo_ctx* c, int32_t channels) {
    if (!c) return SVO_ERR_ARG;
    if (channels != 1 && channels != 3) return fail(c, SVO_ERR_ARG, "channels must be 1 (gray) or 3 (BGR)");
    if (c->pend_count != 0) return fail(c, SVO_ERR_STATE, "batches are in flight");
    c->in_channels = channels;
    return SVO_OK;
}

int svo_set_rectify_maps(svo_ctx* c, const int16_t* map_xy_left, const uint16_t* map_a_left, const int16_t* map_xy_right,
                         const uint16_t* map_a_right) {
    if (!c) return SVO_ERR_ARG;
    if (c->pend_count != 0) return fail(c, SVO_ERR_STATE, "batches are in flight");
    const int16_t* mxy[2] = {map_xy_left, map_xy_right};
    const uint16_t* ma[2] = {map_a_left, map_a_right};
    const int nnull = (!mxy[0]) + (!mxy[1]) + (!ma[0]) + (!ma[1]);
    if (nnull == 4) { c->rectify = false; return SVO_OK; }          // obs->areImagesRectified()
    if (nnull != 0) return fail(c, SVO_ERR_ARG, "all four maps or none");
    CU(c, cudaSetDevice(c->cfg.device));
    const size_t npx = (size_t)c->cfg.width * c->cfg.height;
    for (int sd = 0; sd < 2; ++sd) {
        if (!c->d_map_xy[sd]) {
            cudaError_t e = dalloc(c, &c->d_map_xy[sd], npx);
            if (e == cudaSuccess) e = dalloc(c, &c->d_map_a[sd], npx);
            if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc rectification maps");
        }
        CU(c, cudaMemcpyAsync(c->d_map_xy[sd], mxy[sd], npx * sizeof(short2), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->d_map_a[sd], ma[sd], npx * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    c->rectify = true;
    return SVO_OK;
}

int svo_prepare_rectify_device(svo_ctx* c, int32_t n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                               size_t image_pitch, int32_t channels, uint8_t* d_out_left, uint8_t* d_out_right) {
    if (!c) return SVO_ERR_ARG;
    if (n <= 0 || !d_left || !d_right || !d_out_left || !d_out_right) return fail(c, SVO_ERR_ARG, "bad argument");
    if (channels != 1 && channels != 3) return fail(c, SVO_ERR_ARG, "channels must be 1 (gray) or 3 (BGR)");
    if (stride < (size_t)c->cfg.width * channels) return fail(c, SVO_ERR_ARG, "stride smaller than one image row");
    CU(c, cudaSetDevice(c->cfg.device));
    PrepareArgs A = prepare_args(c, d_left, d_right, stride, image_pitch, d_out_left, d_out_right);
    if (channels == 3 && c->rectify && !c->d_gray[0]) {      // scratch for the two-pass colour path
        for (int sd = 0; sd < 2; ++sd) {
            cudaError_t e = dalloc(c, &c->d_gray[sd], (size_t)c->n_slots_alloc * c->pitch * c->cfg.height);
            if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc gray planes");
        }
        A.tmpL = c->d_gray[0]; A.tmpR = c->d_gray[1];
    }
    if (n > c->n_slots_alloc) A.tmpL = A.tmpR = nullptr;      // larger than the scratch: fused single pass
    int L1 = 0;
    CU(c, launch_prepare(A, n, channels, c->rectify, c->stream, &L1));
    c->launches += L1;
    return SVO_OK;
}

int svo_prepare_rectify(svo_ctx* c, const uint8_t* left, const uint8_t* right, size_t stride, int32_t channels,
                        uint8_t* out_left, uint8_t* out_right) {
    if (!c) return SVO_ERR_ARG;
    if (!left || !right || !out_left || !out_right) return fail(c, SVO_ERR_ARG, "NULL argument");
    if (channels != 1 && channels != 3) return fail(c, SVO_ERR_ARG, "channels must be 1 (gray) or 3 (BGR)");
    const int W = c->cfg.width, H = c->cfg.height;
    if (stride < (size_t)W * channels) return fail(c, SVO_ERR_ARG, "stride smaller than one image row");
    if (c->pend_count != 0) return fail(c, SVO_ERR_STATE, "batches are in flight");
    CU(c, cudaSetDevice(c->cfg.device));
    const int keep = c->in_channels;
    c->in_channels = std::max(c->in_channels, channels);
    int rc = ensure_raw_staging(c);
    c->in_channels = keep;
    if (rc != SVO_OK) return rc;
    const size_t rp = c->raw_pitch;
    CU(c, cudaMemcpy2DAsync(c->d_raw[0][0], rp, left, stride, (size_t)W * channels, H, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpy2DAsync(c->d_raw[0][1], rp, right, stride, (size_t)W * channels, H, cudaMemcpyHostToDevice, c->stream));
    int L1 = 0;
    CU(c, launch_prepare(prepare_args(c, c->d_raw[0][0], c->d_raw[0][1], rp, rp * H, c->d_stL[0], c->d_stR[0]), 1, channels,
                         c->rectify, c->stream, &L1));
    c->launches += L1;
    CU(c, cudaMemcpy2DAsync(out_left, W, c->d_stL[0], c->pitch, W, H, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpy2DAsync(out_right, W, c->d_stR[0], c->pitch, W, H, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

int svo_build_pyramid_device(svo_ctx* c, int32_t n, const uint8_t* d_left, const uint8_t* d_right, size_t stride, size_t image_pitch) {
    if (!c) return SVO_ERR_ARG;
    if (n <= 0 || n > c->cfg.max_seqs || !d_left || !d_right) return fail(c, SVO_ERR_ARG, "bad argument");
    if (stride < (size_t)c->cfg.width) return fail(c, SVO_ERR_ARG, "stride smaller than the image width");
    if (c->n_oct < 2) return fail(c, SVO_ERR_STATE, "the context has one octave: there is no pyramid to build");
    CU(c, cudaSetDevice(c->cfg.device));
    const uint8_t* srcL = d_left;
    const uint8_t* srcR = d_right;
    size_t src_stride = stride, src_pitch = image_pitch;
    int sw = c->cfg.width, sh = c->cfg.height, L = 0;
    for (int k = 1; k < c->n_oct; ++k) {
        OctaveCtx& o = c->oct[k];
        PyramidArgs A;
        A.srcL = srcL; A.srcR = srcR; A.src_stride = src_stride; A.src_image_pitch = src_pitch;
        A.dstL = o.imgL; A.dstR = o.imgR; A.dst_stride = (size_t)o.pitch; A.dst_image_pitch = (size_t)o.pitch * o.H;
        A.sw = sw; A.sh = sh;
        CU(c, launch_pyramid_level(A, n, pyramid_method(c->cfg.pyr_mode, sw), c->stream, &L));
        srcL = o.imgL; srcR = o.imgR; src_stride = (size_t)o.pitch; src_pitch = A.dst_image_pitch;
        sw = o.W; sh = o.H;
    }
    c->launches += L;
    return SVO_OK;
}

int svo_process_batch_device(svo_ctx* c, int32_t n, const int32_t* slots, const uint8_t* d_left, const uint8_t* d_right,
                             size_t stride, size_t image_pitch, const svo_camera* cams, const uint8_t* repeat) {
    if (!c) return SVO_ERR_ARG;
    if (!d_left || !d_right) return fail(c, SVO_ERR_ARG, "image pointer is NULL");
    if (stride < (size_t)c->cfg.width) return fail(c, SVO_ERR_ARG, "stride smaller than the image width");
    CU(c, cudaSetDevice(c->cfg.device));
    int rc = upload_batch_meta(c, n, slots, cams, repeat);
    if (rc != SVO_OK) return rc;
    return run_step(c, n, d_left, d_right, stride, image_pitch, repeat != nullptr);
}

int svo_fetch_results(svo_ctx* c, int32_t n, const int32_t* slots, svo_result* results) {
    if (!c || !results) return SVO_ERR_ARG;
    if (n <= 0 || n > c->cfg.max_seqs) return fail(c, SVO_ERR_ARG, "batch size out of range");
    int lo = c->cfg.max_seqs, hi = -1;
    for (int i = 0; i < n; ++i) {
        const int s = slots ? slots[i] : i;
        if (!valid_slot(c, s)) return fail(c, SVO_ERR_ARG, "slot index out of range");
        lo = std::min(lo, s); hi = std::max(hi, s);
    }
    CU(c, cudaMemcpyAsync(c->h_results + lo, c->P.result + lo, (size_t)(hi - lo + 1) * sizeof(Result),
                          cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    int rc = SVO_OK;
    for (int i = 0; i < n; ++i) {
        const int s = slots ? slots[i] : i;
        memcpy(&results[i], &c->h_results[s], sizeof(svo_result));
    }
    // per-slot step status (capacity overflow) lives in the slot state
    std::vector<SlotState> st(hi - lo + 1);
    CU(c, cudaMemcpy(st.data(), c->P.state + lo, st.size() * sizeof(SlotState), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const int s = slots ? slots[i] : i;
        results[i].status = st[s - lo].status;
        if (st[s - lo].status != 0) { results[i].valid = 0; rc = st[s - lo].status; }
    }
    if (rc == SVO_ERR_CAPACITY) return fail(c, rc, "max_corners exceeded for at least one image (raise svo_config.max_corners)");
    return rc;
}

int svo_submit_batch(svo_ctx* c, int32_t n, const int32_t* slots, const uint8_t* const* left,
                     const uint8_t* const* right, size_t stride, const svo_camera* cams, const uint8_t* repeat) {
    if (!c) return SVO_ERR_ARG;
    if (!left || !right || !cams) return fail(c, SVO_ERR_ARG, "NULL argument");
    if (n <= 0 || n > c->cfg.max_seqs) return fail(c, SVO_ERR_ARG, "batch size out of range");
    if (c->pend_count == kDepth) return fail(c, SVO_ERR_STATE, "three batches are already in flight: call svo_wait_batch first");
    CU(c, cudaSetDevice(c->cfg.device));
    const int b = (c->pend_head + c->pend_count) % kDepth;
    std::vector<char> seen(c->cfg.max_seqs, 0);
    int lo = c->cfg.max_seqs, hi = -1;
    for (int i = 0; i < n; ++i) {
        const int s = slots ? slots[i] : i;
        if (!valid_slot(c, s)) return fail(c, SVO_ERR_ARG, "slot index out of range");
        if (seen[s]) return fail(c, SVO_ERR_ARG, "duplicate slot in one batch");
        if (!left[i] || !right[i]) return fail(c, SVO_ERR_ARG, "image pointer is NULL");
        seen[s] = 1;
        c->r_slots[b][i] = s;
        memcpy(&c->r_cams[b][i], &cams[i], sizeof(Camera));
        c->r_repeat[b][i] = repeat ? repeat[i] : 0;
        lo = std::min(lo, s); hi = std::max(hi, s);
    }
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t plane = (size_t)c->pitch * H;
    const bool stage1 = needs_stage1(c);
    if (stage1) {
        int rc1 = ensure_raw_staging(c);
        if (rc1 != SVO_OK) return rc1;
    }
    // images: host -> staging buffer b on the copy stream (overlaps the kernels of the previous batch); inputs that
    // need stage 1 (colour / unrectified) land in the raw staging buffer and are converted on the device
    uint8_t* upL = stage1 ? c->d_raw[b][0] : c->d_stL[b];
    uint8_t* upR = stage1 ? c->d_raw[b][1] : c->d_stR[b];
    const size_t up_pitch = stage1 ? c->raw_pitch : (size_t)c->pitch;
    const size_t up_plane = up_pitch * H;
    const size_t row_bytes = (size_t)W * (stage1 ? c->in_channels : 1);
    if (stride < row_bytes) return fail(c, SVO_ERR_ARG, "stride smaller than one image row");
    bool contig = true;
    for (int i = 1; i < n && contig; ++i)
        contig = left[i] == left[0] + (size_t)i * stride * H && right[i] == right[0] + (size_t)i * stride * H;
    // images in ordinary (pageable) host memory -- what an MRPT CImage or a cv::Mat holds: the driver would stage every
    // one of the 2n copies through its own bounce buffer, synchronously.  They are gathered into a pinned block of the
    // library by a few host threads instead (memcpy at memory speed) and go up as two large asynchronous copies.
    bool pageable = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, left[0]) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (pageable) {
        const size_t need = 2 * (size_t)c->cfg.max_seqs * up_plane;
        if (c->h_stage_bytes < need) {
            for (int k = 0; k < kDepth; ++k) {
                if (c->h_stage[k]) { cudaFreeHost(c->h_stage[k]); c->h_stage[k] = nullptr; }
                CU(c, cudaHostAlloc((void**)&c->h_stage[k], need, cudaHostAllocDefault));
            }
            c->h_stage_bytes = need;
        }
        uint8_t* hs = c->h_stage[b];
        const int nthr = std::max(1, std::min(getenv("SVO_GATHER_THREADS") ? atoi(getenv("SVO_GATHER_THREADS")) : 16, std::min(n, (int)std::thread::hardware_concurrency())));
        auto work = [&](int t) {
            for (int i = t; i < n; i += nthr)
                for (int side = 0; side < 2; ++side) {
                    const uint8_t* src = side ? right[i] : left[i];
                    uint8_t* dst = hs + ((size_t)side * n + i) * up_plane;
                    if (stride == up_pitch) memcpy(dst, src, up_plane);
                    else for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * up_pitch, src + (size_t)y * stride, row_bytes);
                }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        CU(c, cudaMemcpyAsync(upL, hs, up_plane * n, cudaMemcpyHostToDevice, c->copy_stream));
        CU(c, cudaMemcpyAsync(upR, hs + (size_t)n * up_plane, up_plane * n, cudaMemcpyHostToDevice, c->copy_stream));
    } else if (contig && stride == up_pitch) {
        CU(c, cudaMemcpyAsync(upL, left[0], up_plane * n, cudaMemcpyHostToDevice, c->copy_stream));
        CU(c, cudaMemcpyAsync(upR, right[0], up_plane * n, cudaMemcpyHostToDevice, c->copy_stream));
    } else if (stride == up_pitch) {
        // separately allocated images with the plane's own row pitch (one buffer per camera image, the usual caller): one
        // batched copy call for the 2n planes; per-image copies if the driver refuses the batch
        std::vector<void*> dsts(2 * (size_t)n), srcs(2 * (size_t)n);
        std::vector<size_t> sizes(2 * (size_t)n, up_plane);
        for (int i = 0; i < n; ++i) {
            dsts[2 * i] = upL + i * up_plane; srcs[2 * i] = const_cast<uint8_t*>(left[i]);
            dsts[2 * i + 1] = upR + i * up_plane; srcs[2 * i + 1] = const_cast<uint8_t*>(right[i]);
        }
        cudaError_t eb = cudaErrorNotSupported;
        if (c->batch_copy_ok) {
            cudaMemcpyAttributes attr;
            memset(&attr, 0, sizeof attr);
            attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;     // the sources stay valid until the matching wait
            size_t attr_idx = 0, fail_idx = 0;
            const int lanes = (n >= 64 && !getenv("SVO_ONE_COPY_LANE")) ? svo_ctx::kCopyLanes : 1;
            if (lanes == 1) {
                eb = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &attr_idx, 1, &fail_idx, c->copy_stream);
            } else {
                // the lanes start behind what the copy stream has queued (the staging buffer's previous user) and meet in it again
                eb = cudaEventRecord(c->ev_lane_go, c->copy_stream);
                for (int k = 0; k < lanes && eb == cudaSuccess; ++k) {
                    const size_t lo = dsts.size() * k / lanes, hi = dsts.size() * (k + 1) / lanes;
                    eb = cudaStreamWaitEvent(c->copy_lane[k], c->ev_lane_go, 0);
                    if (eb == cudaSuccess)
                        eb = cudaMemcpyBatchAsync(dsts.data() + lo, srcs.data() + lo, sizes.data() + lo, hi - lo, &attr, &attr_idx, 1, &fail_idx, c->copy_lane[k]);
                    if (eb == cudaSuccess) eb = cudaEventRecord(c->ev_lane[k], c->copy_lane[k]);
                    if (eb == cudaSuccess) eb = cudaStreamWaitEvent(c->copy_stream, c->ev_lane[k], 0);
                }
            }
            if (eb != cudaSuccess) { cudaGetLastError(); c->batch_copy_ok = false; }
        }
        if (eb != cudaSuccess)
            for (size_t k = 0; k < dsts.size(); ++k) CU(c, cudaMemcpyAsync(dsts[k], srcs[k], up_plane, cudaMemcpyHostToDevice, c->copy_stream));
    } else {
        for (int i = 0; i < n; ++i) {
            CU(c, cudaMemcpy2DAsync(upL + i * up_plane, up_pitch, left[i], stride, row_bytes, H, cudaMemcpyHostToDevice, c->copy_stream));
            CU(c, cudaMemcpy2DAsync(upR + i * up_plane, up_pitch, right[i], stride, row_bytes, H, cudaMemcpyHostToDevice, c->copy_stream));
        }
    }
    CU(c, cudaEventRecord(c->ev_h2d[b], c->copy_stream));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_h2d[b], 0));
    {
        void *ds = nullptr, *dc = nullptr, *dr = nullptr;      // device views of the pinned rings (identical under UVA)
        CU(c, cudaHostGetDevicePointer(&ds, c->r_slots[b], 0));
        CU(c, cudaHostGetDevicePointer(&dc, c->r_cams[b], 0));
        CU(c, cudaHostGetDevicePointer(&dr, c->r_repeat[b], 0));
        upload_meta_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->P.slots, c->d_cams_batch, c->d_repeat, (const int32_t*)ds,
                                                                   (const Camera*)dc, repeat ? (const uint8_t*)dr : nullptr, n);
        CU(c, cudaGetLastError());
        c->launches += 1;
    }
    if (stage1) {   // stage1_prepare_rectify on the device: raw staging -> gray, rectified planes
        int L1 = 0;
        CU(c, launch_prepare(prepare_args(c, upL, upR, up_pitch, up_plane, c->d_stL[b], c->d_stR[b]), n, c->in_channels,
                             c->rectify, c->stream, &L1));
        c->launches += L1;
    }
    int rc = run_step(c, n, c->d_stL[b], c->d_stR[b], c->pitch, plane, repeat != nullptr);
    if (rc != SVO_OK) return rc;
    CU(c, cudaMemcpyAsync(c->r_res[b] + lo, c->P.result + lo, (size_t)(hi - lo + 1) * sizeof(Result), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(c->r_state[b] + lo, c->P.state + lo, (size_t)(hi - lo + 1) * sizeof(SlotState), cudaMemcpyDeviceToHost, c->stream));
    if (c->detail) {   // result.out_residual / result.outliers of every slot of the batch (libstereo-odometry.h:239-240)
        const size_t gc = (size_t)c->PG.p.kp_cap;
        CU(c, cudaMemcpyAsync(c->r_resid[b] + lo * gc, c->PG.residual + lo * gc, (size_t)(hi - lo + 1) * gc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(c->r_outl[b] + lo * gc, c->PG.outliers + lo * gc, (size_t)(hi - lo + 1) * gc * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaMemcpyAsync(c->r_stats[b] + stat_idx(lo, 0, 0), c->P.oct_stats + stat_idx(lo, 0, 0), (size_t)(hi - lo + 1) * kMaxOctaves * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaEventRecord(c->ev_done[b], c->stream));
    c->pend_n[b] = n; c->pend_lo[b] = lo; c->pend_hi[b] = hi;
    ++c->pend_count;
    return SVO_OK;
}

int svo_set_result_detail(svo_ctx* c, int32_t on) {
    if (!c) return SVO_ERR_ARG;
    if (c->pend_count != 0) return fail(c, SVO_ERR_STATE, "batches are in flight");
    if (on && !c->r_resid[0]) {
        const size_t cnt = (size_t)c->n_slots_alloc * c->PG.p.kp_cap;
        for (int b = 0; b < kDepth; ++b) {
            cudaError_t e = halloc(c, &c->r_resid[b], cnt);
            if (e == cudaSuccess) e = halloc(c, &c->r_outl[b], cnt);
            if (e == cudaSuccess) e = halloc(c, &c->r_stats[b], (size_t)c->n_slots_alloc * kMaxOctaves * 4);
            if (e != cudaSuccess) return cuda_fail(c, e, "cudaMallocHost (result detail rings)");
        }
    }
    c->detail = on ? 1 : 0;
    return SVO_OK;
}

int svo_wait_batch(svo_ctx* c, svo_result* results) { return svo_wait_batch_detail(c, results, nullptr, nullptr, 0, nullptr); }

int svo_wait_batch_detail(svo_ctx* c, svo_result* results, double* residuals, int64_t* outliers, int32_t cap, int32_t* octave_stats) {
    if (!c || !results) return SVO_ERR_ARG;
    if (c->pend_count == 0) return fail(c, SVO_ERR_STATE, "no batch in flight");
    if ((residuals || outliers || octave_stats) && !c->detail) return fail(c, SVO_ERR_STATE, "svo_set_result_detail(ctx, 1) was not called before the submit");
    const int b = c->pend_head;
    CU(c, cudaEventSynchronize(c->ev_done[b]));
    int rc = SVO_OK;
    const size_t gc = (size_t)c->PG.p.kp_cap;
    for (int i = 0; i < c->pend_n[b]; ++i) {
        const int s = c->r_slots[b][i];
        memcpy(&results[i], &c->r_res[b][s], sizeof(svo_result));
        const int st = c->r_state[b][s].status;
        results[i].status = st;
        if (st != 0) { results[i].valid = 0; rc = st; }
        if (residuals) memcpy(residuals + (size_t)i * cap, c->r_resid[b] + s * gc, sizeof(double) * (size_t)std::min(std::max(results[i].n_residuals, 0), cap));
        if (outliers) {
            const int m = std::min(std::max(results[i].n_outliers, 0), cap);
            for (int k = 0; k < m; ++k) outliers[(size_t)i * cap + k] = c->r_outl[b][s * gc + k];
        }
        if (octave_stats) memcpy(octave_stats + (size_t)i * kMaxOctaves * 4, c->r_stats[b] + stat_idx(s, 0, 0), kMaxOctaves * 4 * sizeof(int32_t));
    }
    c->pend_head = (c->pend_head + 1) % kDepth;
    --c->pend_count;
    if (rc == SVO_ERR_CAPACITY) return fail(c, rc, "max_corners exceeded for at least one image (raise svo_config.max_corners)");
    return rc;
}

int svo_process_batch(svo_ctx* c, int32_t n, const int32_t* slots, const uint8_t* const* left,
                      const uint8_t* const* right, size_t stride, const svo_camera* cams, const uint8_t* repeat,
                      svo_result* results) {
    if (!c) return SVO_ERR_ARG;
    if (!results) return fail(c, SVO_ERR_ARG, "NULL argument");
    if (c->pend_count != 0) return fail(c, SVO_ERR_STATE, "batches submitted with svo_submit_batch are still in flight");
    const int rc = svo_submit_batch(c, n, slots, left, right, stride, cams, repeat);
    if (rc != SVO_OK) return rc;
    return svo_wait_batch(c, results);
}

int svo_process_pair(svo_ctx* c, const uint8_t* left, const uint8_t* right, size_t stride, const svo_camera* cam,
                     int32_t repeat, svo_result* result) {
    const int32_t slot = 0;
    const uint8_t rep = repeat ? 1 : 0;
    return svo_process_batch(c, 1, &slot, &left, &right, stride, cam, repeat ? &rep : nullptr, result);
}

// ---- readers -------------------------------------------------------------------------------------
int svo_get_keypoints(svo_ctx* c, int32_t slot, int32_t which, int32_t side, svo_keypoint* out, int32_t cap) {
    return svo_get_keypoints_oct(c, slot, 0, which, side, out, cap);
}

int svo_get_keypoints_oct(svo_ctx* c, int32_t slot, int32_t octave, int32_t which, int32_t side, svo_keypoint* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if ((!valid_slot(c, slot) && slot != c->hidden_slot) || (which | 1) != 1 || (side | 1) != 1) return fail(c, SVO_ERR_ARG, "bad slot/which/side");
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (which == 0 ? !s.has_cur : !s.has_prev) return fail(c, SVO_ERR_STATE, "frame not present");
    const size_t f = frame_idx(slot, frame_buf(s, which), side);
    int n = 0;
    CU(c, cudaMemcpy(&n, Pk->n_kp + f, sizeof(int), cudaMemcpyDeviceToHost));
    if (Pk->kpf) {                                     // dmORB: cv::ORB's keypoints
        std::vector<float4> kf(n);
        std::vector<int32_t> ko(n);
        if (n) {
            CU(c, cudaMemcpy(kf.data(), Pk->kpf + f * Pk->p.kp_cap, n * sizeof(float4), cudaMemcpyDeviceToHost));
            CU(c, cudaMemcpy(ko.data(), Pk->kpo + f * Pk->p.kp_cap, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        }
        for (int i = 0; i < n && i < cap; ++i) {
            svo_keypoint k;
            k.x = kf[i].x; k.y = kf[i].y; k.response = kf[i].z; k.angle = kf[i].w; k.octave = ko[i];
            k.size = (float)kOrbPatch * c->orb.L[ko[i]].scale; k.class_id = -1;
            out[i] = k;
        }
        return n;
    }
    std::vector<int2> tmp(n);
    if (n) CU(c, cudaMemcpy(tmp.data(), Pk->kp + f * Pk->p.kp_cap, n * sizeof(int2), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; ++i) {
        svo_keypoint k;
        k.x = (float)(tmp[i].x & 0xFFFF); k.y = (float)((unsigned)tmp[i].x >> 16);
        k.size = 7.f; k.angle = -1.f; k.response = (float)tmp[i].y; k.octave = 0; k.class_id = -1;   // cv::FAST keypoints
        out[i] = k;
    }
    return n;
}

int svo_get_descriptors(svo_ctx* c, int32_t slot, int32_t which, int32_t side, uint8_t* out, int32_t cap_rows) {
    return svo_get_descriptors_oct(c, slot, 0, which, side, out, cap_rows);
}

int svo_get_descriptors_oct(svo_ctx* c, int32_t slot, int32_t octave, int32_t which, int32_t side, uint8_t* out, int32_t cap_rows) {
    if (!c) return SVO_ERR_ARG;
    if ((!valid_slot(c, slot) && slot != c->hidden_slot) || (which | 1) != 1 || (side | 1) != 1) return fail(c, SVO_ERR_ARG, "bad slot/which/side");
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (which == 0 ? !s.has_cur : !s.has_prev) return fail(c, SVO_ERR_STATE, "frame not present");
    const size_t f = frame_idx(slot, frame_buf(s, which), side);
    int n = 0;
    CU(c, cudaMemcpy(&n, Pk->n_kp + f, sizeof(int), cudaMemcpyDeviceToHost));
    const int m = std::min(n, cap_rows);
    if (m > 0) CU(c, cudaMemcpy(out, Pk->desc + f * Pk->p.kp_cap * 2, (size_t)m * 32, cudaMemcpyDeviceToHost));
    return n;
}

int svo_get_matches(svo_ctx* c, int32_t slot, int32_t which, svo_dmatch* out, int32_t cap) {
    return svo_get_matches_oct(c, slot, 0, which, out, cap);
}

int svo_get_matches_oct(svo_ctx* c, int32_t slot, int32_t octave, int32_t which, svo_dmatch* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if ((!valid_slot(c, slot) && slot != c->hidden_slot) || (which | 1) != 1) return fail(c, SVO_ERR_ARG, "bad slot/which");
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (which == 0 ? !s.has_cur : !s.has_prev) return fail(c, SVO_ERR_STATE, "frame not present");
    const size_t pi = pair_idx(slot, frame_buf(s, which));
    int n = 0;
    CU(c, cudaMemcpy(&n, Pk->n_match + pi, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int4> tmp(n);
    if (n) CU(c, cudaMemcpy(tmp.data(), Pk->match + pi * Pk->p.kp_cap, n * sizeof(int4), cudaMemcpyDeviceToHost));
    // imgIdx: 0 from BFMatcher::match (S3:91-94), -1 from the cv::DMatch(q, t, d) constructor of the row-by-row branch (S3:404)
    const int img_idx_field = c->cfg.match_method == 1 ? -1 : 0;
    for (int i = 0; i < n && i < cap; ++i) {
        svo_dmatch m = {tmp[i].x, tmp[i].y, img_idx_field, (float)tmp[i].z};
        out[i] = m;
    }
    return n;
}

int svo_get_match_ids(svo_ctx* c, int32_t slot, int32_t which, int64_t* out, int32_t cap) {
    return svo_get_match_ids_oct(c, slot, 0, which, out, cap);
}

int svo_get_match_ids_oct(svo_ctx* c, int32_t slot, int32_t octave, int32_t which, int64_t* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot) || (which | 1) != 1) return fail(c, SVO_ERR_ARG, "bad slot/which");
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    if (!c->cfg.vo_use_matches_ids) return 0;
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (which == 0 ? !s.has_cur : !s.has_prev) return fail(c, SVO_ERR_STATE, "frame not present");
    const size_t pi = pair_idx(slot, frame_buf(s, which));
    int n = 0;
    CU(c, cudaMemcpy(&n, Pk->n_match + pi, sizeof(int), cudaMemcpyDeviceToHost));
    const int m = std::min(n, cap);
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    if (m > 0) CU(c, cudaMemcpy(out, Pk->match_id + pi * Pk->p.kp_cap, (size_t)m * 8, cudaMemcpyDeviceToHost));
    return n;
}

// ---- match-ID / key-frame control (libstereo-odometry.h:675-726) ------------------------------------
static int put_state(svo_ctx* c, int slot, const SlotState& s) {
    CU(c, cudaMemcpy(c->P.state + slot, &s, sizeof(SlotState), cudaMemcpyHostToDevice));
    return SVO_OK;
}

int svo_set_this_frame_as_kf(svo_ctx* c, int32_t slot) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    if (!c->cfg.vo_use_matches_ids) return fail(c, SVO_ERR_STATE, "vo_use_matches_ids is off: there are no match IDs");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (!s.has_cur) return fail(c, SVO_ERR_STATE, "[VO Error -- setThisFrameAsKF] Current frame does not exist");
    const size_t pi = pair_idx(slot, s.cur);
    int n = 0;
    CU(c, cudaMemcpy(&n, c->P.n_match + pi, sizeof(int), cudaMemcpyDeviceToHost));
    if (n <= 0) return fail(c, SVO_ERR_STATE, "[VO Error -- setThisFrameAsKF] No existing lr_pairing_data");
    std::vector<long long> ids(n);
    CU(c, cudaMemcpy(ids.data(), c->P.match_id + pi * c->P.p.kp_cap, (size_t)n * 8, cudaMemcpyDeviceToHost));
    long long mx = ids[0];
    for (long long v : ids) mx = std::max(mx, v);
    s.last_kf_max_id = mx;                           // *std::max_element(v.begin(), v.end())
    return put_state(c, slot, s);
}

int svo_reset_ids(svo_ctx* c, int32_t slot) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    s.reset_ids = 1;                                 // m_reset = true
    return put_state(c, slot, s);
}

int svo_set_ids(svo_ctx* c, int32_t slot, const int64_t* ids, int32_t n) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot) || n < 0 || (n > 0 && !ids)) return fail(c, SVO_ERR_ARG, "bad argument");
    if (!c->cfg.vo_use_matches_ids) return fail(c, SVO_ERR_STATE, "vo_use_matches_ids is off: there are no match IDs");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    if (!s.has_cur && !s.has_prev) return fail(c, SVO_ERR_STATE, "no frame present");
    // H:685-694: the current frame's IDs if there is one, else the previous frame's
    const size_t pi = pair_idx(slot, s.has_cur ? s.cur : (s.cur ^ 1));
    int nm = 0;
    CU(c, cudaMemcpy(&nm, c->P.n_match + pi, sizeof(int), cudaMemcpyDeviceToHost));
    if (n != nm) return fail(c, SVO_ERR_ARG, "setIds: one ID per stereo match of the frame is required");
    if (n) CU(c, cudaMemcpy(c->P.match_id + pi * c->P.p.kp_cap, ids, (size_t)n * 8, cudaMemcpyHostToDevice));
    return SVO_OK;
}

int svo_set_max_match_id(svo_ctx* c, int32_t slot, int64_t id) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    SlotState s;
    int rc = get_state(c, slot, &s);
    if (rc) return rc;
    s.last_match_id = id;                            // m_last_match_ID = id
    return put_state(c, slot, s);
}

int svo_get_tracked_pairs(svo_ctx* c, int32_t slot, int64_t* out_pairs, int32_t cap) {
    return svo_get_tracked_pairs_oct(c, slot, 0, out_pairs, cap);
}

int svo_get_tracked_pairs_oct(svo_ctx* c, int32_t slot, int32_t octave, int64_t* out_pairs, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    CU(c, cudaStreamSynchronize(c->stream));
    int n = 0;
    CU(c, cudaMemcpy(&n, Pk->n_tracked + slot, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int2> tmp(n);
    if (n) CU(c, cudaMemcpy(tmp.data(), Pk->tracked + (size_t)slot * Pk->p.kp_cap, n * sizeof(int2), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; ++i) { out_pairs[2 * i] = tmp[i].x; out_pairs[2 * i + 1] = tmp[i].y; }
    return n;
}

int svo_get_residuals(svo_ctx* c, int32_t slot, double* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    CU(c, cudaStreamSynchronize(c->stream));
    Result r;
    CU(c, cudaMemcpy(&r, c->P.result + slot, sizeof(Result), cudaMemcpyDeviceToHost));
    const int m = std::min(r.n_residuals, cap);
    if (m > 0) CU(c, cudaMemcpy(out, c->PG.residual + (size_t)slot * c->PG.p.kp_cap, (size_t)m * 8, cudaMemcpyDeviceToHost));
    return r.n_residuals;
}

int svo_get_outliers(svo_ctx* c, int32_t slot, int64_t* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    CU(c, cudaStreamSynchronize(c->stream));
    Result r;
    CU(c, cudaMemcpy(&r, c->P.result + slot, sizeof(Result), cudaMemcpyDeviceToHost));
    std::vector<int32_t> tmp(r.n_outliers);
    if (r.n_outliers) CU(c, cudaMemcpy(tmp.data(), c->PG.outliers + (size_t)slot * c->PG.p.kp_cap, r.n_outliers * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < r.n_outliers && i < cap; ++i) out[i] = tmp[i];
    return r.n_outliers;
}

// statistics of the two findFundamentalMat calls of the slot's last step (octave 0): points, inliers, iterations executed
int svo_get_ransac_stats(svo_ctx* c, int32_t slot, int32_t* out6) {
    if (!c || !out6) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    CU(c, cudaStreamSynchronize(c->stream));
    int32_t a[2], b[2], d[2];
    CU(c, cudaMemcpy(a, c->RS.n_pts + (size_t)slot * 2, 8, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(b, c->RS.n_inl + (size_t)slot * 2, 8, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(d, c->RS.n_iter + (size_t)slot * 2, 8, cudaMemcpyDeviceToHost));
    out6[0] = a[0]; out6[1] = b[0]; out6[2] = d[0]; out6[3] = a[1]; out6[4] = b[1]; out6[5] = d[1];
    return SVO_OK;
}

// survivors of stage 5's auxiliary non-maximal suppression (S5:465-474) over the flattened tracked pairs (S5:419-461),
// as the first Gauss-Newton loop starts with them: what optimization.txt lists (S5:477-492)
int svo_get_stage5_survivors(svo_ctx* c, int32_t slot, uint8_t* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    CU(c, cudaStreamSynchronize(c->stream));
    const GnScratch& G = c->n_oct > 1 ? c->GG : c->G;
    const size_t gc = (size_t)c->PG.p.kp_cap;
    int T = 0;
    CU(c, cudaMemcpy(&T, c->PG.n_tracked + slot, sizeof(int), cudaMemcpyDeviceToHost));
    T = std::max(0, std::min(T, (int)gc));
    const int m = std::min(T, (int)cap);
    if (m > 0 && out) CU(c, cudaMemcpy(out, G.mask0 + slot * gc, (size_t)m, cudaMemcpyDeviceToHost));
    return T;
}

int svo_get_track_candidates(svo_ctx* c, int32_t slot, int32_t* out_query, int32_t* out_trainL, int32_t* out_trainR,
                             uint8_t* out_inlL, uint8_t* out_inlR, int32_t cap) {
    return svo_get_track_candidates_oct(c, slot, 0, out_query, out_trainL, out_trainR, out_inlL, out_inlR, nullptr, nullptr, cap);
}

int svo_get_track_candidates_oct(svo_ctx* c, int32_t slot, int32_t octave, int32_t* out_query, int32_t* out_trainL, int32_t* out_trainR,
                                 uint8_t* out_inlL, uint8_t* out_inlR, float* out_distL, float* out_distR, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    CU(c, cudaStreamSynchronize(c->stream));
    int n = 0;
    CU(c, cudaMemcpy(&n, Pk->n_tcand + slot, sizeof(int), cudaMemcpyDeviceToHost));
    const size_t kc = Pk->p.kp_cap;
    std::vector<int4> tmp(n);
    std::vector<uint8_t> il(n), ir(n);
    std::vector<uint32_t> bl, br;
    if (n) {
        CU(c, cudaMemcpy(tmp.data(), Pk->tcand + (size_t)slot * kc, n * sizeof(int4), cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(il.data(), Pk->inl + ((size_t)slot * 2 + 0) * kc, n, cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(ir.data(), Pk->inl + ((size_t)slot * 2 + 1) * kc, n, cudaMemcpyDeviceToHost));
        if ((out_distL || out_distR) && c->cfg.if_match_method == 0) {   // the BF distances of matL / matR (S4:141-142), by query index
            bl.resize(kc); br.resize(kc);
            CU(c, cudaMemcpy(bl.data(), Pk->tr_best + ((size_t)slot * 2 + 0) * kc, kc * 4, cudaMemcpyDeviceToHost));
            CU(c, cudaMemcpy(br.data(), Pk->tr_best + ((size_t)slot * 2 + 1) * kc, kc * 4, cudaMemcpyDeviceToHost));
        }
    }
    for (int i = 0; i < n && i < cap; ++i) {
        if (out_query) out_query[i] = tmp[i].x;
        if (out_trainL) out_trainL[i] = tmp[i].y;
        if (out_trainR) out_trainR[i] = tmp[i].z;
        if (out_inlL) out_inlL[i] = il[i];
        if (out_inlR) out_inlR[i] = ir[i];
        if (out_distL) out_distL[i] = bl.empty() ? 0.f : (float)(bl[tmp[i].x] >> 16);
        if (out_distR) out_distR[i] = br.empty() ? 0.f : (float)(br[tmp[i].x] >> 16);
    }
    return n;
}

int svo_get_octave_info(svo_ctx* c, int32_t octave, int32_t* out4) {
    if (!c || !out4) return SVO_ERR_ARG;
    const Pools* Pk = octave_pools(c, octave);
    if (!Pk) return fail(c, SVO_ERR_ARG, "octave out of range");
    out4[0] = Pk->p.W; out4[1] = Pk->p.H; out4[2] = Pk->p.quota; out4[3] = Pk->p.kp_cap;
    return c->n_oct;
}

int svo_get_octave_stats(svo_ctx* c, int32_t slot, int32_t* out, int32_t cap_octaves) {
    if (!c || !out) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    CU(c, cudaStreamSynchronize(c->stream));
    int32_t tmp[kMaxOctaves * 4];
    CU(c, cudaMemcpy(tmp, c->P.oct_stats + stat_idx(slot, 0, 0), sizeof(tmp), cudaMemcpyDeviceToHost));
    const int n = std::min(c->n_oct, (int)cap_octaves);
    for (int i = 0; i < n * 4; ++i) out[i] = tmp[i];
    return c->n_oct;
}

int svo_get_pyramid_plane(svo_ctx* c, int32_t item, int32_t octave, int32_t side, uint8_t* out) {
    if (!c || !out) return SVO_ERR_ARG;
    if (octave < 1 || octave >= c->n_oct || item < 0 || item >= c->cfg.max_seqs || (side | 1) != 1) return fail(c, SVO_ERR_ARG, "bad item/octave/side");
    CU(c, cudaStreamSynchronize(c->stream));
    const OctaveCtx& o = c->oct[octave];
    const uint8_t* src = (side ? o.imgR : o.imgL) + (size_t)item * o.pitch * o.H;
    CU(c, cudaMemcpy2D(out, o.W, src, o.pitch, o.W, o.H, cudaMemcpyDeviceToHost));
    return SVO_OK;
}

int svo_get_plane(svo_ctx* c, int32_t slot, int32_t side, int32_t plane, uint8_t* out) {
    if (!c || !out) return SVO_ERR_ARG;
    if (slot == -1) slot = c->hidden_slot;   // the scratch slot used by svo_detect
    if ((!valid_slot(c, slot) && slot != c->hidden_slot) || (side | 1) != 1 || (plane | 1) != 1) return fail(c, SVO_ERR_ARG, "bad slot/side/plane");
    CU(c, cudaStreamSynchronize(c->stream));
    const int W = c->cfg.width, H = c->cfg.height;
    if (plane == 0) {
        const uint8_t* src = c->P.blur + img_idx(slot, side) * (size_t)c->pitch * H;
        CU(c, cudaMemcpy2D(out, W, src, c->pitch, W, H, cudaMemcpyDeviceToHost));
        return SVO_OK;
    }
    // the score plane is not kept on the device: rebuilt here from the per-tile corner lists
    const int tx = tiles_x_of(W), ty = tiles_y_of(H);
    const size_t tiles = (size_t)tx * ty;
    std::vector<uint16_t> seg(tiles * kTileH), lst(tiles * kTileCap);
    CU(c, cudaMemcpy(seg.data(), c->P.tile_seg + img_idx(slot, side) * tiles * kTileH, seg.size() * 2, cudaMemcpyDeviceToHost));
    CU(c, cudaMemcpy(lst.data(), c->P.tile_list + img_idx(slot, side) * tiles * kTileCap, lst.size() * 2, cudaMemcpyDeviceToHost));
    memset(out, 0, (size_t)W * H);
    for (int y = 0; y < H; ++y)
        for (int t = 0; t < tx; ++t) {
            const uint16_t sg = seg[(size_t)y * tx + t];
            const uint16_t* band = &lst[((size_t)(y / kTileH) * tx + t) * kTileCap + (size_t)((y % kTileH) / kBandH) * kBandCap];
            for (int k = sg >> 8; k < (sg >> 8) + (sg & 0xFF) && k < kBandCap; ++k) {
                const int x = t * kTileW + (band[k] & 0xFF);
                if (x < W) out[(size_t)y * W + x] = (uint8_t)(band[k] >> 8);
            }
        }
    return SVO_OK;
}

int svo_get_orb_raw_keypoints(svo_ctx* c, int32_t slot, int32_t side, svo_keypoint* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (c->cfg.detect_method != 0) return fail(c, SVO_ERR_STATE, "the context does not run dmORB");
    if ((!valid_slot(c, slot) && slot != c->hidden_slot) || (side | 1) != 1) return fail(c, SVO_ERR_ARG, "bad slot/side");
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->stream));
    const size_t im = img_idx(slot, side);
    int n = 0;
    CU(c, cudaMemcpy(&n, c->orb.n_raw + im, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float4> kf(n);
    std::vector<int32_t> ko(n);
    if (n) {
        CU(c, cudaMemcpy(kf.data(), c->orb.raw + im * c->orb.raw_cap, n * sizeof(float4), cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(ko.data(), c->orb.raw_oct + im * c->orb.raw_cap, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i < n && i < cap; ++i) {
        svo_keypoint k;
        k.x = kf[i].x; k.y = kf[i].y; k.response = kf[i].z; k.angle = kf[i].w; k.octave = ko[i];
        k.size = (float)kOrbPatch * c->orb.L[ko[i]].scale; k.class_id = -1;
        out[i] = k;
    }
    return n;
}

int svo_get_orb_level(svo_ctx* c, int32_t slot, int32_t side, int32_t level, int32_t plane, uint8_t* out, int32_t* w, int32_t* h) {
    if (!c) return SVO_ERR_ARG;
    if (c->cfg.detect_method != 0) return fail(c, SVO_ERR_STATE, "the context does not run dmORB");
    if ((!valid_slot(c, slot) && slot != c->hidden_slot) || (side | 1) != 1 || level < 0 || level >= c->orb.n_levels || (plane | 1) != 1)
        return fail(c, SVO_ERR_ARG, "bad slot/side/level/plane");
    const OrbLevelDev& L = c->orb.L[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (!out) return SVO_OK;
    CU(c, cudaSetDevice(c->cfg.device));
    CU(c, cudaStreamSynchronize(c->stream));
    const uint8_t* src;
    if (plane == 1) src = L.blur + img_idx(slot, side) * (size_t)L.pitch * L.h;
    else {
        if (level == 0) return fail(c, SVO_ERR_ARG, "level 0 is the caller's image");
        // level images are kept per batch item of the last step: the per-stage entry (svo_detect) uses item 0
        src = (side ? L.imgR : L.imgL);
    }
    CU(c, cudaMemcpy2D(out, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return SVO_OK;
}

// ---- per-stage entries ---------------------------------------------------------------------------
int svo_detect(svo_ctx* c, const uint8_t* img, size_t stride, svo_keypoint* out_kps, uint8_t* out_desc, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (!img) return fail(c, SVO_ERR_ARG, "image pointer is NULL");
    if (stride < (size_t)c->cfg.width) return fail(c, SVO_ERR_ARG, "stride smaller than the image width");
    CU(c, cudaSetDevice(c->cfg.device));
    const int hs = c->hidden_slot;
    const int W = c->cfg.width, H = c->cfg.height;
    const size_t plane = (size_t)c->pitch * H;
    int rc = set_single_slot(c, hs);
    if (rc) return rc;
    CU(c, cudaMemcpy2DAsync(c->d_stageL, c->pitch, img, stride, W, H, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpy2DAsync(c->d_stageR, c->pitch, img, stride, W, H, cudaMemcpyHostToDevice, c->stream));
    CU(c, launch_begin_step(c->P, 1, nullptr, c->stream));
    int L = 1;
    if (c->cfg.detect_method == 0)
        CU(c, launch_orb_detect(c->P, c->orb, 1, c->d_stageL, c->d_stageR, c->pitch, plane, 0, c->d_occ, c->occ_words,
                                c->d_sort_scratch, c->d_list_scratch, c->stream, &L));
    else
    CU(c, launch_detect(c->P, 1, c->d_stageL, c->d_stageR, c->pitch, plane, c->pitch, c->d_occ, c->occ_words,
                        c->d_sort_scratch, c->d_list_scratch, c->d_key_scratch, c->stream, &L));
    c->launches += L;
    const int n = svo_get_keypoints(c, hs, 0, 0, out_kps, cap);
    if (n < 0) return n;
    SlotState s;
    rc = get_state(c, hs, &s);
    if (rc) return rc;
    if (s.status != 0) return fail(c, s.status, "max_corners exceeded (raise svo_config.max_corners)");
    if (out_desc) { const int m = svo_get_descriptors(c, hs, 0, 0, out_desc, cap); if (m < 0) return m; }
    return n;
}

int svo_match_hamming_device(svo_ctx* c, const uint8_t* d_q, int32_t nq, const uint8_t* d_t, int32_t nt, int32_t reps,
                             uint32_t* d_out) {
    if (!c) return SVO_ERR_ARG;
    if (nq < 0 || nt < 0 || nt > 65535 || reps < 0) return fail(c, SVO_ERR_ARG, "bad sizes (train rows must be <= 65535)");
    CU(c, launch_hamming_raw(d_q, nq, d_t, nt, reps, d_out, c->stream));
    if (nq > 0 && reps > 0) c->launches += 1;
    return SVO_OK;
}

int svo_match_hamming(svo_ctx* c, const uint8_t* query, int32_t nq, const uint8_t* train, int32_t nt, svo_dmatch* out) {
    if (!c) return SVO_ERR_ARG;
    if (nq < 0 || nt < 0 || nt > 65535) return fail(c, SVO_ERR_ARG, "bad sizes (train rows must be <= 65535)");
    if (nq == 0 || nt == 0) return 0;   // BFMatcher returns no matches
    if (!query || !train || !out) return fail(c, SVO_ERR_ARG, "NULL argument");
    CU(c, cudaSetDevice(c->cfg.device));
    uint8_t *dq = nullptr, *dt = nullptr;
    uint32_t* dout = nullptr;
    CU(c, cudaMalloc(&dq, (size_t)nq * 32));
    cudaError_t e = cudaMalloc(&dt, (size_t)nt * 32);
    if (e == cudaSuccess) e = cudaMalloc(&dout, (size_t)nq * 4);
    std::vector<uint32_t> h(nq);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dq, query, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dt, train, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = launch_hamming_raw(dq, nq, dt, nt, 1, dout, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.data(), dout, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dq); cudaFree(dt); cudaFree(dout);
    if (e != cudaSuccess) return cuda_fail(c, e, "svo_match_hamming");
    c->launches += 1;
    for (int i = 0; i < nq; ++i) {
        svo_dmatch m = {i, (int32_t)(h[i] & 0xFFFFu), 0, (float)(h[i] >> 16)};
        out[i] = m;
    }
    return nq;
}

// caller keypoints into the frame whose keypoints start at element fbase (= frame_idx * kp_cap) of the octave-0 arrays
static int upload_kps(svo_ctx* c, const svo_keypoint* k, int n, size_t fbase) {
    std::vector<int2> tmp(n);
    const bool float_kp = c->P.kpf != nullptr;          // dmORB context: cv::ORB keypoints, any float coordinates
    for (int i = 0; i < n; ++i) {
        const float x = k[i].x, y = k[i].y;
        if (!(x >= 0 && y >= 0 && x < kMaxImgDim && y < kMaxImgDim) || (!float_kp && (x != std::floor(x) || y != std::floor(y))))
            return fail(c, SVO_ERR_ARG, "keypoint coordinates must lie inside the image (and be integers unless detect_method = dmORB)");
        tmp[i] = make_int2(((int)y << 16) | (int)x, (int)k[i].response);
    }
    if (n) CU(c, cudaMemcpyAsync(c->P.kp + fbase, tmp.data(), n * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    std::vector<float4> kf;
    std::vector<int32_t> ko;
    if (float_kp && n) {
        kf.resize(n); ko.resize(n);
        for (int i = 0; i < n; ++i) { kf[i] = make_float4(k[i].x, k[i].y, k[i].response, k[i].angle); ko[i] = k[i].octave; }
        CU(c, cudaMemcpyAsync(c->P.kpf + fbase, kf.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->P.kpo + fbase, ko.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return SVO_OK;
}

int svo_match_lr(svo_ctx* c, const svo_keypoint* kpsL, const uint8_t* descL, int32_t nL, const svo_keypoint* kpsR,
                 const uint8_t* descR, int32_t nR, svo_dmatch* out, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    const int kc = c->P.p.kp_cap;
    if (nL < 0 || nR < 0 || nL > kc || nR > kc) return fail(c, SVO_ERR_ARG, "more keypoints than the context's capacity (2*orb_nfeats)");
    CU(c, cudaSetDevice(c->cfg.device));
    const int hs = c->hidden_slot;
    int rc = set_single_slot(c, hs);
    if (rc) return rc;
    CU(c, launch_begin_step(c->P, 1, nullptr, c->stream));
    SlotState s;
    rc = get_state(c, hs, &s);
    if (rc) return rc;
    const size_t fL = frame_idx(hs, s.cur, 0), fR = frame_idx(hs, s.cur, 1);
    rc = upload_kps(c, kpsL, nL, fL * kc);
    if (rc) return rc;
    rc = upload_kps(c, kpsR, nR, fR * kc);
    if (rc) return rc;
    if (nL) CU(c, cudaMemcpyAsync(c->P.desc + fL * kc * 2, descL, (size_t)nL * 32, cudaMemcpyHostToDevice, c->stream));
    if (nR) CU(c, cudaMemcpyAsync(c->P.desc + fR * kc * 2, descR, (size_t)nR * 32, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->P.n_kp + fL, &nL, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->P.n_kp + fR, &nR, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    int L = 1;
    CU(c, launch_match_lr(c->P, 1, c->stream, &L));
    c->launches += L;
    return svo_get_matches(c, hs, 0, out, cap);
}

// one frame of caller data into buffer `buf` of the scratch slot: keypoints, descriptors, stereo matches and the
// descriptor rows gathered per match (what lr_filter_kernel leaves for stage 4, S4:105-131)
static int upload_frame(svo_ctx* c, int hs, int buf, const svo_keypoint* kl, const uint8_t* dl, int nl, const svo_keypoint* kr,
                        const uint8_t* dr, int nr, const svo_dmatch* m, int nm) {
    const Pools& P = c->P;
    const int kc = P.p.kp_cap;
    const size_t fL = frame_idx(hs, buf, 0), fR = frame_idx(hs, buf, 1), pi = pair_idx(hs, buf);
    int rc = upload_kps(c, kl, nl, fL * kc);
    if (rc) return rc;
    rc = upload_kps(c, kr, nr, fR * kc);
    if (rc) return rc;
    std::vector<int4> hm((size_t)nm);
    std::vector<uint8_t> gl((size_t)nm * 32), gr((size_t)nm * 32);
    for (int i = 0; i < nm; ++i) {
        if (m[i].queryIdx < 0 || m[i].queryIdx >= nl || m[i].trainIdx < 0 || m[i].trainIdx >= nr)
            return fail(c, SVO_ERR_ARG, "match index outside the keypoint lists");
        hm[i] = make_int4(m[i].queryIdx, m[i].trainIdx, (int)m[i].distance, 0);
        memcpy(&gl[(size_t)i * 32], dl + (size_t)m[i].queryIdx * 32, 32);
        memcpy(&gr[(size_t)i * 32], dr + (size_t)m[i].trainIdx * 32, 32);
    }
    if (nl) CU(c, cudaMemcpyAsync(P.desc + fL * kc * 2, dl, (size_t)nl * 32, cudaMemcpyHostToDevice, c->stream));
    if (nr) CU(c, cudaMemcpyAsync(P.desc + fR * kc * 2, dr, (size_t)nr * 32, cudaMemcpyHostToDevice, c->stream));
    if (nm) {
        CU(c, cudaMemcpyAsync(P.match + pi * kc, hm.data(), (size_t)nm * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(P.mdesc + fL * kc * 2, gl.data(), (size_t)nm * 32, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(P.mdesc + fR * kc * 2, gr.data(), (size_t)nm * 32, cudaMemcpyHostToDevice, c->stream));
    }
    CU(c, cudaMemcpyAsync(P.n_kp + fL, &nl, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(P.n_kp + fR, &nr, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(P.n_match + pi, &nm, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));      // the host vectors above go out of scope
    return SVO_OK;
}

int svo_track(svo_ctx* c, const svo_keypoint* pre_left, const uint8_t* pre_desc_left, int32_t n_pl,
              const svo_keypoint* pre_right, const uint8_t* pre_desc_right, int32_t n_pr, const svo_dmatch* pre_matches,
              int32_t n_pm, const svo_keypoint* cur_left, const uint8_t* cur_desc_left, int32_t n_cl,
              const svo_keypoint* cur_right, const uint8_t* cur_desc_right, int32_t n_cr, const svo_dmatch* cur_matches,
              int32_t n_cm, int64_t* out_pairs, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    const int kc = c->P.p.kp_cap;
    const int32_t counts[6] = {n_pl, n_pr, n_pm, n_cl, n_cr, n_cm};
    for (int32_t v : counts)
        if (v < 0 || v > kc) return fail(c, SVO_ERR_ARG, "more keypoints / matches than the context's capacity (kps_to_detect[0])");
    if ((n_pl && (!pre_left || !pre_desc_left)) || (n_pr && (!pre_right || !pre_desc_right)) || (n_pm && !pre_matches) ||
        (n_cl && (!cur_left || !cur_desc_left)) || (n_cr && (!cur_right || !cur_desc_right)) || (n_cm && !cur_matches) ||
        (cap > 0 && !out_pairs) || cap < 0)
        return fail(c, SVO_ERR_ARG, "NULL argument");
    if (c->pend_count != 0) return fail(c, SVO_ERR_STATE, "batches are in flight");
    CU(c, cudaSetDevice(c->cfg.device));
    const int hs = c->hidden_slot;
    int rc = set_single_slot(c, hs);
    if (rc) return rc;
    SlotState s;
    memset(&s, 0, sizeof(s));
    s.cur = 0; s.has_cur = 1; s.has_prev = 1;      // buffer 0 = current frame, buffer 1 = previous frame
    CU(c, cudaStreamSynchronize(c->stream));
    rc = put_state(c, hs, s);
    if (rc) return rc;
    Result zero;
    memset(&zero, 0, sizeof(zero));
    CU(c, cudaMemcpy(c->P.result + hs, &zero, sizeof(zero), cudaMemcpyHostToDevice));
    rc = upload_frame(c, hs, 1, pre_left, pre_desc_left, n_pl, pre_right, pre_desc_right, n_pr, pre_matches, n_pm);
    if (rc) return rc;
    rc = upload_frame(c, hs, 0, cur_left, cur_desc_left, n_cl, cur_right, cur_desc_right, n_cr, cur_matches, n_cm);
    if (rc) return rc;
    int L = 0;
    RansacScratch RS = c->RS;
    CU(c, launch_track_match(c->P, 1, c->stream, &L));
    if (c->cfg.fundmat_filter) CU(c, launch_ransac(c->P, RS, 1, c->stream, &L));
    CU(c, launch_consistency(c->P, 1, c->stream, &L));
    c->launches += L;
    CU(c, cudaStreamSynchronize(c->stream));
    int n = 0;
    CU(c, cudaMemcpy(&n, c->P.n_tracked + hs, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int2> tmp((size_t)n);
    if (n) CU(c, cudaMemcpy(tmp.data(), c->P.tracked + (size_t)hs * kc, (size_t)n * sizeof(int2), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; ++i) { out_pairs[2 * i] = tmp[i].x; out_pairs[2 * i + 1] = tmp[i].y; }
    return n;
}

int svo_optimize(svo_ctx* c, int32_t slot, const int64_t* tracked_pairs, int32_t n_tracked, const svo_dmatch* pre_matches,
                 int32_t n_pre, const svo_dmatch* cur_matches, int32_t n_cur, const svo_keypoint* pre_left, int32_t n_pl,
                 const svo_keypoint* pre_right, int32_t n_pr, const svo_keypoint* cur_left, int32_t n_cl,
                 const svo_keypoint* cur_right, int32_t n_cr, const svo_camera* cam, int32_t use_custom_initial_pose,
                 const double* ini_estimation, svo_result* result) {
    return svo_get_change_in_pose(c, slot, tracked_pairs, n_tracked, pre_matches, n_pre, cur_matches, n_cur, pre_left, n_pl,
                                  pre_right, n_pr, cur_left, n_cl, cur_right, n_cr, cam, use_custom_initial_pose, ini_estimation,
                                  result);
}

int svo_get_change_in_pose(svo_ctx* c, int32_t slot, const int64_t* tracked_pairs, int32_t n_tracked,
                           const svo_dmatch* pre_matches, int32_t n_pre, const svo_dmatch* cur_matches, int32_t n_cur,
                           const svo_keypoint* pre_left, int32_t n_pl, const svo_keypoint* pre_right, int32_t n_pr,
                           const svo_keypoint* cur_left, int32_t n_cl, const svo_keypoint* cur_right, int32_t n_cr,
                           const svo_camera* cam, int32_t use_custom_initial_pose, const double* ini_estimation,
                           svo_result* result) {
    if (!c) return SVO_ERR_ARG;
    if (!valid_slot(c, slot)) return fail(c, SVO_ERR_ARG, "bad slot");
    if (!cam || !result || n_tracked < 0) return fail(c, SVO_ERR_ARG, "NULL argument");
    const int kc = c->PG.p.kp_cap;      // with n_octaves > 1: the capacity of the flattened problem
    if (n_tracked > kc) return fail(c, SVO_ERR_ARG, "more tracked pairs than the context's capacity");
    CU(c, cudaSetDevice(c->cfg.device));
    // S5:419-461 on the host side of the seam: flatten the index chains into four point lists
    std::vector<float4> prev(n_tracked), obs(n_tracked);
    std::vector<int32_t> cur_idx(n_tracked);
    const float W = (float)c->cfg.width, H = (float)c->cfg.height;
    for (int i = 0; i < n_tracked; ++i) {
        const int64_t a = tracked_pairs[2 * i], b = tracked_pairs[2 * i + 1];
        if (a < 0 || a >= n_pre || b < 0 || b >= n_cur) return fail(c, SVO_ERR_ARG, "tracked pair index out of range");
        const svo_dmatch& pm = pre_matches[a];
        const svo_dmatch& cm = cur_matches[b];
        if (pm.queryIdx < 0 || pm.queryIdx >= n_pl || pm.trainIdx < 0 || pm.trainIdx >= n_pr || cm.queryIdx < 0 ||
            cm.queryIdx >= n_cl || cm.trainIdx < 0 || cm.trainIdx >= n_cr)
            return fail(c, SVO_ERR_ARG, "match index out of range");
        const svo_keypoint& pl = pre_left[pm.queryIdx];
        if (!(pl.x >= 0 && pl.x < W && pl.y >= 0 && pl.y < H)) return fail(c, SVO_ERR_ARG, "previous-left keypoint outside the image");
        prev[i] = make_float4(pl.x, pl.y, pre_right[pm.trainIdx].x, pl.response);
        obs[i] = make_float4(cur_left[cm.queryIdx].x, cur_left[cm.queryIdx].y, cur_right[cm.trainIdx].x, cur_right[cm.trainIdx].y);
        cur_idx[i] = (int32_t)b;
    }
    int rc = set_single_slot(c, slot);
    if (rc) return rc;
    const size_t base = (size_t)slot * kc;
    if (n_tracked) {
        CU(c, cudaMemcpyAsync(c->GG.prev + base, prev.data(), n_tracked * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->GG.obs + base, obs.data(), n_tracked * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(c->GG.cur_idx + base, cur_idx.data(), n_tracked * 4, cudaMemcpyHostToDevice, c->stream));
    }
    CU(c, cudaMemcpyAsync(c->PG.n_tracked + slot, &n_tracked, 4, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->P.cam + slot, cam, sizeof(Camera), cudaMemcpyHostToDevice, c->stream));
    double ini[6] = {0, 0, 0, 0, 0, 0};
    if (use_custom_initial_pose && ini_estimation) memcpy(ini, ini_estimation, sizeof(ini));
    CU(c, cudaMemcpyAsync(c->GG.ini_pose + (size_t)slot * 6, ini, sizeof(ini), cudaMemcpyHostToDevice, c->stream));
    // a fresh TStereoOdometryResult (libstereo-odometry.h:251-263)
    Result zero;
    memset(&zero, 0, sizeof(zero));
    CU(c, cudaMemcpyAsync(c->P.result + slot, &zero, sizeof(zero), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    int L = 0;
    CU(c, launch_gn(c->PG, c->GG, 1, GN_EXTERNAL, use_custom_initial_pose ? 1 : 0, /*assume_fast_kps=*/0, c->stream, &L));
    c->launches += L;
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(result, c->P.result + slot, sizeof(Result), cudaMemcpyDeviceToHost));
    return result->valid;
}

int svo_get_projected_coords(svo_ctx* c, const svo_dmatch* pre_matches, int32_t n_matches, const svo_keypoint* pre_left,
                             int32_t n_pl, const svo_keypoint* pre_right, int32_t n_pr, const int32_t* other_matches_tracked,
                             const svo_camera* cam, const double* delta_pose, float* out_pixels, int32_t cap) {
    if (!c) return SVO_ERR_ARG;
    if (n_matches < 0 || n_pl < 0 || n_pr < 0 || !cam || !delta_pose || (n_matches > 0 && (!pre_matches || !other_matches_tracked)))
        return fail(c, SVO_ERR_ARG, "bad argument");
    std::vector<int32_t> order;
    std::vector<int4> hm((size_t)n_matches);
    for (int m = 0; m < n_matches; ++m) {
        hm[m] = make_int4(pre_matches[m].queryIdx, pre_matches[m].trainIdx, 0, 0);
        if (other_matches_tracked[m] != -1) continue;                 // C:430-431
        if (pre_matches[m].queryIdx < 0 || pre_matches[m].queryIdx >= n_pl || pre_matches[m].trainIdx < 0 || pre_matches[m].trainIdx >= n_pr)
            return fail(c, SVO_ERR_ARG, "match index outside the keypoint lists");
        order.push_back(m);
    }
    const int n = (int)order.size();
    if (n > cap) return fail(c, SVO_ERR_ARG, "output capacity too small");
    if (n == 0) return 0;
    if (!out_pixels || !pre_left || !pre_right) return fail(c, SVO_ERR_ARG, "NULL argument");
    CU(c, cudaSetDevice(c->cfg.device));
    std::vector<float2> hl((size_t)n_pl), hr((size_t)n_pr);
    for (int i = 0; i < n_pl; ++i) hl[i] = make_float2(pre_left[i].x, pre_left[i].y);
    for (int i = 0; i < n_pr; ++i) hr[i] = make_float2(pre_right[i].x, pre_right[i].y);
    int4* dm = nullptr; float2* dl = nullptr; float2* dr = nullptr; int32_t* dord = nullptr; double* dpose = nullptr; float4* dout = nullptr;
    cudaError_t e = cudaMalloc(&dm, hm.size() * sizeof(int4));
    if (e == cudaSuccess) e = cudaMalloc(&dl, std::max<size_t>(1, hl.size()) * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&dr, std::max<size_t>(1, hr.size()) * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&dord, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&dpose, 6 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&dout, (size_t)n * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dm, hm.data(), hm.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_pl) e = cudaMemcpyAsync(dl, hl.data(), hl.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n_pr) e = cudaMemcpyAsync(dr, hr.data(), hr.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dord, order.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpose, delta_pose, 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    Camera cm;
    memcpy(&cm, cam, sizeof(Camera));
    if (e == cudaSuccess) e = launch_project(dm, dl, dr, dord, n, cm, dpose, dout, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_pixels, dout, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dm); cudaFree(dl); cudaFree(dr); cudaFree(dord); cudaFree(dpose); cudaFree(dout);
    if (e != cudaSuccess) return cuda_fail(c, e, "svo_get_projected_coords");
    c->launches += 1;
    return n;
}

int svo_fundmat_ransac(svo_ctx* c, const float* p1, const float* p2, int32_t n, uint8_t* mask) {
    if (!c) return SVO_ERR_ARG;
    if (n < 0 || (n > 0 && (!p1 || !p2 || !mask))) return fail(c, SVO_ERR_ARG, "bad arguments");
    if (n > c->P.p.kp_cap) return fail(c, SVO_ERR_ARG, "more points than the context's capacity (2*orb_nfeats)");
    CU(c, cudaSetDevice(c->cfg.device));
    int inl = 0;
    cudaError_t e = ransac_host_points(c->RS, c->hidden_slot, p1, p2, n, mask, &inl, c->stream);
    if (e != cudaSuccess) return cuda_fail(c, e, "svo_fundmat_ransac");
    c->launches += 1;
    return inl;
}

int svo_fundmat_7point(svo_ctx* c, const float* p1, const float* p2, int32_t n_sets, double* F, int32_t* n_models) {
    if (!c) return SVO_ERR_ARG;
    if (n_sets < 0 || (n_sets > 0 && (!p1 || !p2 || !F || !n_models))) return fail(c, SVO_ERR_ARG, "bad arguments");
    CU(c, cudaSetDevice(c->cfg.device));
    cudaError_t e = fm7_host_points(p1, p2, n_sets, F, n_models, c->stream);
    if (e != cudaSuccess) return cuda_fail(c, e, "svo_fundmat_7point");
    if (n_sets > 0) c->launches += 1;
    return n_sets;
}

}  // extern "C"
