// skps_pipeline: the device-side chain of FaceAna.run (Skps/core/api/facer.py:52-85):
//   frame -> letterbox -> detector -> NMS/un-letterbox -> judge_boxs(track) -> sort_and_filter
//         -> per-face crop+resize -> landmark net -> de-normalise
// One H2D copy of the frame in, one D2H copy of the packed results out, nothing in between
// touches the host.  The temporal smoothing that follows (GroupTrack, facer.py:71-82) is O(K*98)
// host math on the returned landmarks.
#include <string.h>

#include <vector>

#include "../../include/skps_b200.h"
#include "common.h"

using namespace skps;

struct skps_pipeline {
    skps_engine* det = nullptr;
    skps_engine* kps = nullptr;
    skps_pipeline_cfg cfg;
    int device = 0;
    int det_h = 0, det_w = 0, kps_hw = 0, n_points = 0, det_rows = 0;
    int max_det = 256;
    uint8_t* d_frame[2] = {nullptr, nullptr};     // current / previous frame
    int cur = 0;
    int prev_h = 0, prev_w = 0;                   // size of the frame in d_frame[cur^1] (0 = none)
    uint8_t* h_frame = nullptr;                   // pinned staging
    float* d_det_rows = nullptr; int32_t* d_det_idx = nullptr; int32_t* d_det_count = nullptr;
    float* d_track = nullptr;
    float* d_boxes = nullptr; int32_t* d_count = nullptr; int32_t* d_detail = nullptr;
    float* d_kps = nullptr;
    unsigned long long* d_diff = nullptr;
    // packed host result block (pinned)
    struct Host {
        int32_t n_faces, n_det;
        unsigned long long diff;
    };
    Host* h_res = nullptr;
    float* h_boxes = nullptr; float* h_kps = nullptr; float* h_scores = nullptr;
    int32_t* h_det_idx = nullptr; float* h_det_rows = nullptr; float* h_track = nullptr;
};

extern "C" SKPS_API void skps_pipeline_destroy(skps_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int i = 0; i < 2; ++i) if (p->d_frame[i]) cudaFree(p->d_frame[i]);
    if (p->h_frame) cudaFreeHost(p->h_frame);
    void* dev[] = {p->d_det_rows, p->d_det_idx, p->d_det_count, p->d_track, p->d_boxes, p->d_count, p->d_detail,
                   p->d_kps, p->d_diff};
    for (void* q : dev) if (q) cudaFree(q);
    void* host[] = {p->h_res, p->h_boxes, p->h_kps, p->h_scores, p->h_det_idx, p->h_det_rows, p->h_track};
    for (void* q : host) if (q) cudaFreeHost(q);
    delete p;
}

extern "C" SKPS_API int skps_pipeline_create(skps_engine* det, skps_engine* kps, const skps_pipeline_cfg* cfg,
                                    skps_pipeline** out) {
    SKPS_CHECK(det && kps && cfg && out, "pipeline_create: null argument");
    SKPS_CHECK(cfg->top_k > 0 && cfg->top_k <= 64, "pipeline_create: top_k %d outside 1..64", cfg->top_k);
    skps_pipeline* p = new skps_pipeline();
    p->det = det; p->kps = kps; p->cfg = *cfg;
    int c = 0;
    skps_engine_input_dims(det, &p->det_h, &p->det_w, &c);
    int kh = 0, kw = 0;
    skps_engine_input_dims(kps, &kh, &kw, &c);
    SKPS_CHECK(kh == kw, "pipeline_create: landmark input must be square");
    p->kps_hw = kh;
    SKPS_CHECK(skps_engine_num_outputs(det) == 1 && skps_engine_num_outputs(kps) == 2, "pipeline_create: engine outputs");
    p->det_rows = skps_engine_output_elems(det, 0) / 16;
    p->n_points = skps_engine_output_elems(kps, 1);
    SKPS_CHECK(skps_engine_output_elems(kps, 0) == 2 * p->n_points, "pipeline_create: landmark outputs");
    cudaGetDevice(&p->device);
    const size_t fbytes = (size_t)cfg->max_h * cfg->max_w * 3;
    const int K = cfg->top_k, P = p->n_points;
#define PALLOC(ptr, bytes) SKPS_CUDA(cudaMalloc((void**)&(ptr), (bytes)))
#define HALLOC(ptr, bytes) SKPS_CUDA(cudaMallocHost((void**)&(ptr), (bytes)))
    PALLOC(p->d_frame[0], fbytes); PALLOC(p->d_frame[1], fbytes);
    HALLOC(p->h_frame, fbytes);
    PALLOC(p->d_det_rows, sizeof(float) * 16 * p->max_det);
    PALLOC(p->d_det_idx, sizeof(int32_t) * p->max_det);
    PALLOC(p->d_det_count, sizeof(int32_t));
    PALLOC(p->d_track, sizeof(float) * 4 * 256);
    PALLOC(p->d_boxes, sizeof(float) * 4 * K);
    PALLOC(p->d_count, sizeof(int32_t));
    PALLOC(p->d_detail, sizeof(int32_t) * 5 * K);
    PALLOC(p->d_kps, sizeof(float) * 2 * P * K);
    PALLOC(p->d_diff, sizeof(unsigned long long));
    HALLOC(p->h_res, sizeof(skps_pipeline::Host));
    HALLOC(p->h_boxes, sizeof(float) * 4 * K);
    HALLOC(p->h_kps, sizeof(float) * 2 * P * K);
    HALLOC(p->h_scores, sizeof(float) * P * K);
    HALLOC(p->h_det_idx, sizeof(int32_t) * p->max_det);
    HALLOC(p->h_det_rows, sizeof(float) * 16 * p->max_det);
    HALLOC(p->h_track, sizeof(float) * 4 * 256);
    SKPS_CUDA(cudaMemset(p->d_det_count, 0, sizeof(int32_t)));
    *out = p;
    return 0;
}

extern "C" SKPS_API int skps_pipeline_reset(skps_pipeline* p) {
    SKPS_CHECK(p, "pipeline_reset: null");
    p->prev_h = p->prev_w = 0;            // FaceAna.reset (facer.py:200-208): previous_image = None
    return 0;
}

// Upload (or adopt) the frame into d_frame[cur]; returns the device pointer.
static int stage_frame(skps_pipeline* p, const uint8_t* frame, int H, int W, int on_device, cudaStream_t s,
                       const uint8_t** dptr) {
    SKPS_CHECK(H > 0 && W > 0 && (size_t)H * W <= (size_t)p->cfg.max_h * p->cfg.max_w,
               "frame %dx%d has more pixels than the pipeline maximum %dx%d", H, W, p->cfg.max_h, p->cfg.max_w);
    size_t bytes = (size_t)H * W * 3;
    if (on_device) {
        SKPS_CUDA(cudaMemcpyAsync(p->d_frame[p->cur], frame, bytes, cudaMemcpyDeviceToDevice, s));
    } else {
        // frames already in pinned (page-locked) memory go straight to the GPU; pageable frames are first copied
        // into the pipeline's pinned staging buffer so the H2D copy stays asynchronous and at full PCIe rate
        cudaPointerAttributes attr;
        bool pinned = cudaPointerGetAttributes(&attr, frame) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        if (!pinned) {
            cudaGetLastError();          // clear the "invalid value" a pageable pointer may leave behind
            memcpy(p->h_frame, frame, bytes);
        }
        SKPS_CUDA(cudaMemcpyAsync(p->d_frame[p->cur], pinned ? frame : p->h_frame, bytes, cudaMemcpyHostToDevice, s));
    }
    *dptr = p->d_frame[p->cur];
    return 0;
}

extern "C" SKPS_API int skps_pipeline_frame_diff(skps_pipeline* p, const uint8_t* frame, int H, int W, int on_device,
                                        double* mean_diff, void* stream) {
    SKPS_CHECK(p && frame && mean_diff, "frame_diff: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    SKPS_CUDA(cudaSetDevice(p->device));
    const uint8_t* d = nullptr;
    if (stage_frame(p, frame, H, W, on_device, s, &d)) return 1;
    if (p->prev_h != H || p->prev_w != W) {
        *mean_diff = -1.0;
        return 0;
    }
    size_t n = (size_t)H * W * 3;
    if (skps_frame_absdiff_sum(p->d_frame[p->cur ^ 1], d, n, p->d_diff, s)) return 1;
    SKPS_CUDA(cudaMemcpyAsync(&p->h_res->diff, p->d_diff, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    SKPS_CUDA(cudaStreamSynchronize(s));
    // facer.py:113: np.sum(diff)/H/W/3.
    *mean_diff = (double)p->h_res->diff / (double)H / (double)W / 3.0;
    return 0;
}

// The frame staged by skps_pipeline_frame_diff becomes the "previous" frame without running the chain (FaceAna.run on
// a static frame with nothing to track, facer.py:57-62: previous_image is replaced every frame).
extern "C" SKPS_API int skps_pipeline_commit_frame(skps_pipeline* p, int H, int W) {
    SKPS_CHECK(p && H > 0 && W > 0, "commit_frame: bad arguments");
    p->prev_h = H; p->prev_w = W;
    p->cur ^= 1;
    return 0;
}

extern "C" SKPS_API int skps_pipeline_run(skps_pipeline* p, const uint8_t* frame, int H, int W, int frame_on_device,
                                 int run_detector, int rw, int rh, int top, int left, float scale,
                                 const float* track, int n_track, int32_t* n_faces, float* boxes4, float* kps,
                                 float* scores, int32_t* n_det, int32_t* det_idx, float* det_rows, void* stream) {
    SKPS_CHECK(p && n_faces && boxes4 && kps && scores, "pipeline_run: null argument");
    SKPS_CHECK(n_track >= 0 && n_track <= 256, "pipeline_run: n_track %d", n_track);
    cudaStream_t s = (cudaStream_t)stream;
    SKPS_CUDA(cudaSetDevice(p->device));
    const skps_pipeline_cfg& c = p->cfg;
    const int K = c.top_k, P = p->n_points;
    const uint8_t* d_frame = nullptr;
    if (frame) {
        if (stage_frame(p, frame, H, W, frame_on_device, s, &d_frame)) return 1;
    } else {
        d_frame = p->d_frame[p->cur];          // already staged by skps_pipeline_frame_diff
    }
    if (n_track > 0) {
        SKPS_CHECK(track, "pipeline_run: track is null");
        memcpy(p->h_track, track, sizeof(float) * 4 * n_track);
        SKPS_CUDA(cudaMemcpyAsync(p->d_track, p->h_track, sizeof(float) * 4 * n_track, cudaMemcpyHostToDevice, s));
    }
    if (run_detector) {
        uint8_t* det_in = (uint8_t*)skps_engine_input_ptr(p->det);
        if (skps_letterbox(d_frame, H, W, W * 3, det_in, p->det_h, p->det_w, rw, rh, top, left, s)) return 1;
        if (skps_engine_forward(p->det, det_in, 1, nullptr, s)) return 1;
        if (skps_detect_post(skps_engine_output_ptr(p->det, 0), p->det_rows, c.score_thres, c.iou_thres, scale,
                             (float)left, (float)top, p->d_det_rows, p->d_det_idx, p->d_det_count, p->max_det, s))
            return 1;
        // facer.py:58 judge_boxs(track_box, boxes) then :64 sort_and_filter
        if (skps_select_faces(p->d_det_rows, p->d_det_count, 16, n_track > 0 ? p->d_track : nullptr, n_track,
                              c.track_iou, c.alpha, (float)(1.0 - (double)c.alpha), c.min_face, K, p->d_boxes,
                              p->d_count, s))
            return 1;
    } else {
        // facer.py:61: boxes = track_box, then sort_and_filter
        int32_t nt = n_track;
        p->h_res->n_det = nt;
        SKPS_CUDA(cudaMemcpyAsync(p->d_det_count, &p->h_res->n_det, sizeof(int32_t), cudaMemcpyHostToDevice, s));
        if (skps_select_faces(p->d_track, p->d_det_count, 4, nullptr, 0, c.track_iou, c.alpha,
                              (float)(1.0 - (double)c.alpha), c.min_face, K, p->d_boxes, p->d_count, s))
            return 1;
    }
    uint8_t* kps_in = (uint8_t*)skps_engine_input_ptr(p->kps);
    if (skps_crop_resize(d_frame, H, W, W * 3, p->d_boxes, p->d_count, K, c.face_scale, c.kps_min_face, kps_in,
                         p->kps_hw, p->d_detail, s))
        return 1;
    if (skps_engine_forward(p->kps, kps_in, K, nullptr, s)) return 1;
    if (skps_landmark_post(skps_engine_output_ptr(p->kps, 0), p->d_detail, p->d_count, K, P, p->d_kps, s)) return 1;
    SKPS_CUDA(cudaMemcpyAsync(&p->h_res->n_faces, p->d_count, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    SKPS_CUDA(cudaMemcpyAsync(p->h_boxes, p->d_boxes, sizeof(float) * 4 * K, cudaMemcpyDeviceToHost, s));
    SKPS_CUDA(cudaMemcpyAsync(p->h_kps, p->d_kps, sizeof(float) * 2 * P * K, cudaMemcpyDeviceToHost, s));
    SKPS_CUDA(cudaMemcpyAsync(p->h_scores, skps_engine_output_ptr(p->kps, 1), sizeof(float) * P * K,
                              cudaMemcpyDeviceToHost, s));
    if (run_detector) SKPS_CUDA(cudaMemcpyAsync(&p->h_res->n_det, p->d_det_count, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (run_detector && n_det) {
        SKPS_CUDA(cudaMemcpyAsync(p->h_det_idx, p->d_det_idx, sizeof(int32_t) * p->max_det, cudaMemcpyDeviceToHost, s));
        SKPS_CUDA(cudaMemcpyAsync(p->h_det_rows, p->d_det_rows, sizeof(float) * 16 * p->max_det, cudaMemcpyDeviceToHost, s));
    }
    SKPS_CUDA(cudaStreamSynchronize(s));
    SKPS_CHECK(!run_detector || p->h_res->n_det >= 0, "detector produced %d candidates over the score threshold (limit 1024)",
               -p->h_res->n_det);
    const int nf = p->h_res->n_faces;
    *n_faces = nf;
    memcpy(boxes4, p->h_boxes, sizeof(float) * 4 * nf);
    memcpy(kps, p->h_kps, sizeof(float) * 2 * P * nf);
    memcpy(scores, p->h_scores, sizeof(float) * P * nf);
    if (n_det) {
        *n_det = run_detector ? p->h_res->n_det : 0;
        if (run_detector && det_idx) memcpy(det_idx, p->h_det_idx, sizeof(int32_t) * (*n_det));
        if (run_detector && det_rows) memcpy(det_rows, p->h_det_rows, sizeof(float) * 16 * (*n_det));
    }
    // the frame just processed becomes "previous" for the next frame_diff (facer.py:57,62)
    p->prev_h = H; p->prev_w = W;
    p->cur ^= 1;
    return 0;
}
