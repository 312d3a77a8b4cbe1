// skps_mpipe: FaceAna.run (Skps/core/api/facer.py:52-85) for MANY concurrent video streams on one GPU.
//
// One call takes one frame from each of up to S streams and runs, batched across the streams and with every piece of
// per-stream state resident in HBM:
//   H2D (copy stream, overlapped with the previous batch's compute) -> |prev - cur| gate -> letterbox x S ->
//   ONE detector forward (batch S) -> NMS x S -> judge_boxs/sort_and_filter x S (detector rows or track boxes, chosen on
//   the device by the gate) -> crops x S -> ONE landmark forward (batch S * top_k) -> de-normalise -> GroupTrack / One-Euro /
//   track-box EMA x S (temporal.cu, float64 with numpy's promotion rules) -> D2H of the packed results.
// The host never sees a decision: which streams re-detect, how many faces each has and all smoothing state are device
// data.  Two result slots let the caller keep two batches in flight (submit(0) submit(1) wait(0) submit(0) ...): the frames
// of batch i+1 cross PCIe while batch i computes.  SURVEY.md 8f-1 (temporal layer for many streams) and 8f-2 (frame ring).
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/skps_b200.h"
#include "common.h"
#include "mpipe_kernels.h"

using namespace skps;

struct skps_mpipe {
    skps_engine* det = nullptr;
    skps_engine* kps = nullptr;
    skps_pipeline_cfg cfg;
    int device = 0, S = 0, K = 0, P = 0;
    int det_h = 0, det_w = 0, kps_hw = 0, det_rows = 0, max_det = 256;
    size_t frame_bytes = 0;
    cudaStream_t s_copy = nullptr, s_compute = nullptr;
    // per stream: ring of 3 frames (previous / current / next batch's upload)
    std::vector<uint8_t*> d_frame;            // [S*3]
    std::vector<int> ring_pos;                // [S] index of the most recent frame in the ring
    std::vector<int> prev_h, prev_w;          // [S] size of the most recent frame (0 = none: FaceAna.reset)
    // per slot
    struct Slot {
        uint8_t* h_stage = nullptr;           // pinned staging for pageable frames [S][frame_bytes]
        int32_t* h_hw = nullptr; int32_t* h_have_prev = nullptr; int32_t* h_geom = nullptr;     // pinned, uploaded per batch
        MpStreamDesc* h_desc = nullptr;       // pinned: per-stream frame pointers + letterbox geometry of this batch
        int32_t* h_count = nullptr; int32_t* h_flag = nullptr; int32_t* h_det_count = nullptr;
        double* h_box = nullptr; double* h_kps = nullptr; float* h_scores = nullptr;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr;
        int n = 0;
        bool busy = false;
    } slot[2];
    // device scratch (one set: batches are serialised on s_compute)
    int32_t *d_hw = nullptr, *d_have_prev = nullptr, *d_flag = nullptr, *d_det_count = nullptr, *d_det_idx = nullptr;
    int32_t *d_count = nullptr, *d_detail = nullptr;
    unsigned long long* d_diff = nullptr;
    MpStreamDesc* d_desc = nullptr;           // this batch's descriptors (uploaded on the compute stream, in order)
    float *d_det_rows = nullptr, *d_boxes = nullptr, *d_kps_now = nullptr;
    // temporal state
    double *d_prev_lm = nullptr, *d_prev_dx = nullptr, *d_track = nullptr, *d_out_kps = nullptr;
    float* d_track_f32 = nullptr;
    int32_t *d_n_prev = nullptr, *d_prev_f32 = nullptr, *d_state_idx = nullptr, *d_n_track = nullptr;
};

static void letterbox_geometry(int H, int W, int in_h, int in_w, float* scale, int* rw, int* rh, int* top, int* left) {
    // face_detector.py:49-62 (python floats = double; int() truncates; round() half-to-even on x.4 / x.6 never ties)
    const double s = fmin((double)in_h / H, (double)in_w / W);
    *rw = (int)(W * s); *rh = (int)(H * s);
    const double dw = (in_w - *rw) / 2.0, dh = (in_h - *rh) / 2.0;
    *top = (int)nearbyint(dh - 0.1); *left = (int)nearbyint(dw - 0.1);
    *scale = (float)s;
}

extern "C" SKPS_API void skps_mpipe_destroy(skps_mpipe* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->s_compute) cudaStreamSynchronize(p->s_compute);
    if (p->s_copy) cudaStreamSynchronize(p->s_copy);
    for (uint8_t* f : p->d_frame) if (f) cudaFree(f);
    for (auto& sl : p->slot) {
        void* host[] = {sl.h_desc, sl.h_stage, sl.h_hw, sl.h_have_prev, sl.h_geom, sl.h_count, sl.h_flag, sl.h_det_count, sl.h_box,
                        sl.h_kps, sl.h_scores};
        for (void* q : host) if (q) cudaFreeHost(q);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    void* dev[] = {p->d_desc, p->d_hw, p->d_have_prev, p->d_flag, p->d_det_count, p->d_det_idx, p->d_count, p->d_detail, p->d_diff,
                   p->d_det_rows, p->d_boxes, p->d_kps_now, p->d_prev_lm, p->d_prev_dx, p->d_track, p->d_out_kps,
                   p->d_track_f32, p->d_n_prev, p->d_prev_f32, p->d_state_idx, p->d_n_track};
    for (void* q : dev) if (q) cudaFree(q);
    if (p->s_copy) cudaStreamDestroy(p->s_copy);
    if (p->s_compute) cudaStreamDestroy(p->s_compute);
    delete p;
}

extern "C" SKPS_API int skps_mpipe_reset(skps_mpipe* p, int stream) {
    SKPS_CHECK(p && stream >= -1 && stream < p->S, "mpipe_reset: bad stream %d", stream);
    SKPS_CUDA(cudaSetDevice(p->device));
    SKPS_CUDA(cudaStreamSynchronize(p->s_compute));
    const int a = stream < 0 ? 0 : stream, b = stream < 0 ? p->S : stream + 1;
    for (int s = a; s < b; ++s) {
        // FaceAna.reset (facer.py:200-208) + a fresh GroupTrack: no previous frame, no track boxes, no landmark history
        p->prev_h[s] = p->prev_w[s] = 0;
        const int32_t zero = 0, none = -1, one = 1;
        SKPS_CUDA(cudaMemcpy(p->d_n_track + s, &zero, 4, cudaMemcpyHostToDevice));
        SKPS_CUDA(cudaMemcpy(p->d_n_prev + s, &none, 4, cudaMemcpyHostToDevice));
        SKPS_CUDA(cudaMemcpy(p->d_prev_f32 + s, &one, 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" SKPS_API int skps_mpipe_create(skps_engine* det, skps_engine* kps, const skps_pipeline_cfg* cfg, int n_streams,
                                          skps_mpipe** out) {
    SKPS_CHECK(det && kps && cfg && out && n_streams > 0 && n_streams <= 256, "mpipe_create: bad arguments");
    SKPS_CHECK(cfg->top_k > 0 && cfg->top_k <= 64, "mpipe_create: top_k %d outside 1..64", cfg->top_k);
    skps_mpipe* p = new skps_mpipe();
    p->det = det; p->kps = kps; p->cfg = *cfg; p->S = n_streams; p->K = cfg->top_k;
    int c = 0, kh = 0, kw = 0;
    skps_engine_input_dims(det, &p->det_h, &p->det_w, &c);
    skps_engine_input_dims(kps, &kh, &kw, &c);
    p->kps_hw = kh;
    p->det_rows = skps_engine_output_elems(det, 0) / 16;
    p->P = skps_engine_output_elems(kps, 1);
    cudaGetDevice(&p->device);
    auto fail = [&](const char* what) {
        char tmp[900];
        snprintf(tmp, sizeof(tmp), "%s", get_error());
        set_error("mpipe_create: %s: %s", what, tmp);
        skps_mpipe_destroy(p);
        return 1;
    };
    if (kh != kw || skps_engine_num_outputs(det) != 1 || skps_engine_num_outputs(kps) != 2) {
        set_error("engines are not a (detector, landmark) pair");
        return fail("engines");
    }
    const int S = p->S, K = p->K, P = p->P;
    p->frame_bytes = ((size_t)cfg->max_h * cfg->max_w * 3 + 255) & ~(size_t)255;
    p->ring_pos.assign(S, 0); p->prev_h.assign(S, 0); p->prev_w.assign(S, 0);
    p->d_frame.assign((size_t)S * 3, nullptr);
#define MP_DEV(ptr, bytes) if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) { set_error("cudaMalloc %zu bytes", (size_t)(bytes)); return fail("alloc"); }
#define MP_HOST(ptr, bytes) if (cudaMallocHost((void**)&(ptr), (bytes)) != cudaSuccess) { set_error("cudaMallocHost %zu bytes", (size_t)(bytes)); return fail("alloc"); }
    for (auto& f : p->d_frame) MP_DEV(f, p->frame_bytes);
    for (auto& sl : p->slot) {
        MP_HOST(sl.h_stage, p->frame_bytes * S);
        MP_HOST(sl.h_hw, sizeof(int32_t) * 2 * S); MP_HOST(sl.h_have_prev, sizeof(int32_t) * S);
        MP_HOST(sl.h_geom, sizeof(int32_t) * 8 * S);
        MP_HOST(sl.h_desc, sizeof(MpStreamDesc) * S);
        MP_HOST(sl.h_count, sizeof(int32_t) * S); MP_HOST(sl.h_flag, sizeof(int32_t) * S); MP_HOST(sl.h_det_count, sizeof(int32_t) * S);
        MP_HOST(sl.h_box, sizeof(double) * 4 * K * S); MP_HOST(sl.h_kps, sizeof(double) * 2 * P * K * S);
        MP_HOST(sl.h_scores, sizeof(float) * P * K * S);
        if (cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming) != cudaSuccess) { set_error("cudaEventCreate"); return fail("event"); }
    }
    MP_DEV(p->d_desc, sizeof(MpStreamDesc) * S); MP_DEV(p->d_hw, 8 * S); MP_DEV(p->d_have_prev, 4 * S); MP_DEV(p->d_flag, 4 * S); MP_DEV(p->d_det_count, 4 * S);
    MP_DEV(p->d_det_idx, 4 * (size_t)p->max_det * S); MP_DEV(p->d_count, 4 * S); MP_DEV(p->d_detail, 4 * 5 * (size_t)K * S);
    MP_DEV(p->d_diff, 8 * S); MP_DEV(p->d_det_rows, 4 * 16 * (size_t)p->max_det * S); MP_DEV(p->d_boxes, 4 * 4 * (size_t)K * S);
    MP_DEV(p->d_kps_now, 4 * 2 * (size_t)P * K * S);
    MP_DEV(p->d_prev_lm, 8 * 2 * 2 * (size_t)P * K * S); MP_DEV(p->d_prev_dx, 8 * 2 * 2 * (size_t)P * K * S);
    MP_DEV(p->d_track, 8 * 4 * (size_t)K * S); MP_DEV(p->d_out_kps, 8 * 2 * (size_t)P * K * S);
    MP_DEV(p->d_track_f32, 4 * 4 * (size_t)K * S);
    MP_DEV(p->d_n_prev, 4 * S); MP_DEV(p->d_prev_f32, 4 * S); MP_DEV(p->d_state_idx, 4 * S); MP_DEV(p->d_n_track, 4 * S);
#undef MP_DEV
#undef MP_HOST
    cudaMemset(p->d_prev_lm, 0, 8 * 2 * 2 * (size_t)P * K * S); cudaMemset(p->d_prev_dx, 0, 8 * 2 * 2 * (size_t)P * K * S);
    cudaMemset(p->d_state_idx, 0, 4 * S); cudaMemset(p->d_track_f32, 0, 4 * 4 * (size_t)K * S);
    cudaMemset(p->d_track, 0, 8 * 4 * (size_t)K * S);
    if (cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate"); return fail("stream"); }
    if (skps_mpipe_reset(p, -1)) return fail("reset");
    *out = p;
    return 0;
}

extern "C" SKPS_API int skps_mpipe_submit(skps_mpipe* p, int slot_i, const uint8_t* const* frames, const int32_t* hw, int n,
                                          int frames_on_device) {
    SKPS_CHECK(p && frames && hw && (slot_i == 0 || slot_i == 1) && n > 0 && n <= p->S, "mpipe_submit: bad arguments");
    skps_mpipe::Slot& sl = p->slot[slot_i];
    SKPS_CHECK(!sl.busy, "mpipe_submit: slot %d still holds results (call skps_mpipe_wait first)", slot_i);
    SKPS_CUDA(cudaSetDevice(p->device));
    const skps_pipeline_cfg& c = p->cfg;
    const int S = p->S, K = p->K, P = p->P;
    cudaStream_t sc = p->s_copy, sx = p->s_compute;
    // ---- uploads on the copy stream: frame i of stream i into the next ring position
    for (int s = 0; s < n; ++s) {
        const int H = hw[2 * s], W = hw[2 * s + 1];
        SKPS_CHECK(frames[s] && H > 0 && W > 0 && (size_t)H * W * 3 <= p->frame_bytes,
                   "mpipe_submit: frame %d is %dx%d, larger than the pipeline maximum %dx%d", s, H, W, c.max_h, c.max_w);
        const size_t bytes = (size_t)H * W * 3;
        const int pos = (p->ring_pos[s] + 1) % 3;
        uint8_t* dst = p->d_frame[(size_t)s * 3 + pos];
        if (frames_on_device) {
            SKPS_CUDA(cudaMemcpyAsync(dst, frames[s], bytes, cudaMemcpyDeviceToDevice, sc));
        } else {
            cudaPointerAttributes attr;
            const bool pinned = cudaPointerGetAttributes(&attr, frames[s]) == cudaSuccess && attr.type == cudaMemoryTypeHost;
            const uint8_t* src = frames[s];
            if (!pinned) {
                cudaGetLastError();
                memcpy(sl.h_stage + p->frame_bytes * s, frames[s], bytes);
                src = sl.h_stage + p->frame_bytes * s;
            }
            SKPS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, sc));
        }
        sl.h_hw[2 * s] = H; sl.h_hw[2 * s + 1] = W;
        sl.h_have_prev[s] = (p->prev_h[s] == H && p->prev_w[s] == W) ? 1 : 0;
    }
    SKPS_CUDA(cudaEventRecord(sl.ev_in, sc));
    // ---- compute
    SKPS_CUDA(cudaStreamWaitEvent(sx, sl.ev_in, 0));
    SKPS_CUDA(cudaMemcpyAsync(p->d_hw, sl.h_hw, 8 * n, cudaMemcpyHostToDevice, sx));
    SKPS_CUDA(cudaMemcpyAsync(p->d_have_prev, sl.h_have_prev, 4 * n, cudaMemcpyHostToDevice, sx));
    SKPS_CUDA(cudaMemsetAsync(p->d_diff, 0, 8 * n, sx));
    uint8_t* det_in = (uint8_t*)skps_engine_input_ptr(p->det);
    const size_t det_in_bytes = (size_t)p->det_h * p->det_w * 3;
    // per-stream frame pointers + letterbox geometry: one small upload, then every pre/post-processing step is ONE launch for
    // all streams (it was 5 launches per stream and call - ~80 launch gaps of a 5 ms call at 16 streams)
    size_t max_bytes = 0;
    for (int s = 0; s < n; ++s) {
        const int H = hw[2 * s], W = hw[2 * s + 1];
        MpStreamDesc& D = sl.h_desc[s];
        D.cur = p->d_frame[(size_t)s * 3 + (p->ring_pos[s] + 1) % 3];
        D.prev = sl.h_have_prev[s] ? p->d_frame[(size_t)s * 3 + p->ring_pos[s]] : nullptr;
        D.H = H; D.W = W; D.have_prev = sl.h_have_prev[s];
        letterbox_geometry(H, W, p->det_h, p->det_w, &D.scale, &D.rw, &D.rh, &D.top, &D.left);
        memcpy(&sl.h_geom[8 * s], &D.scale, 4); sl.h_geom[8 * s + 1] = D.top; sl.h_geom[8 * s + 2] = D.left;
        if ((size_t)H * W * 3 > max_bytes) max_bytes = (size_t)H * W * 3;
    }
    SKPS_CUDA(cudaMemcpyAsync(p->d_desc, sl.h_desc, sizeof(MpStreamDesc) * n, cudaMemcpyHostToDevice, sx));
    if (launch_mp_absdiff(p->d_desc, p->d_diff, n, max_bytes, sx)) return 1;
    if (launch_mp_letterbox(p->d_desc, det_in, det_in_bytes, p->det_h, p->det_w, n, sx)) return 1;
    if (launch_mp_decide(p->d_diff, p->d_hw, p->d_have_prev, p->d_flag, n, sx)) return 1;
    // the detector runs for every stream of the batch (one launch sequence); the gate only chooses whose rows are used
    if (skps_engine_forward(p->det, det_in, n, nullptr, sx)) return 1;
    const float* det_out = skps_engine_output_ptr(p->det, 0);
    if (launch_mp_detect_post(p->d_desc, det_out, p->det_rows, c.score_thres, c.iou_thres, p->d_det_rows, p->d_det_idx,
                              p->d_det_count, p->max_det, n, sx))
        return 1;
    if (launch_mp_select(p->d_det_rows, p->d_det_count, p->max_det, p->d_flag, p->d_track_f32, p->d_n_track, c.track_iou,
                         c.alpha, (float)(1.0 - (double)c.alpha), c.min_face, K, p->d_boxes, p->d_count, n, sx))
        return 1;
    uint8_t* kps_in = (uint8_t*)skps_engine_input_ptr(p->kps);
    if (launch_mp_crop(p->d_desc, p->d_boxes, p->d_count, K, c.face_scale, c.kps_min_face, kps_in, p->kps_hw, p->d_detail, n, sx))
        return 1;
    if (skps_engine_forward(p->kps, kps_in, n * K, nullptr, sx)) return 1;
    if (launch_mp_landmark_post(skps_engine_output_ptr(p->kps, 0), p->d_detail, p->d_count, K, P, p->d_kps_now, n, sx)) return 1;
    MpTemporalArgs a;
    a.top_k = K; a.n_points = P;
    a.kps_now = p->d_kps_now; a.count = p->d_count; a.flag = p->d_flag; a.hw = p->d_hw; a.boxes4 = p->d_boxes;
    a.prev_lm = p->d_prev_lm; a.prev_dx = p->d_prev_dx; a.n_prev = p->d_n_prev; a.prev_f32 = p->d_prev_f32;
    a.state_idx = p->d_state_idx; a.track_box = p->d_track; a.track_f32 = p->d_track_f32; a.n_track = p->d_n_track;
    a.out_kps = p->d_out_kps;
    // the python floats of lk.py / facer.py, evaluated in the same order
    // (the cfg carries them as float32; Skps.yml's 0.5 / 0.3 come back exactly by rounding to 6 decimals in double)
    a.iou_thres = nearbyint((double)c.track_iou * 1e6) / 1e6;
    a.alpha = nearbyint((double)c.alpha * 1e6) / 1e6;
    a.one_minus_alpha = 1.0 - a.alpha;
    a.two_pi = 2 * 3.141592653589793;
    { const double r = a.two_pi * 1.0 * 1.0; a.a_d = r / (r + 1); a.one_minus_a_d = 1 - a.a_d; }
    a.min_cutoff = 0.15; a.beta = 0.8;
    if (launch_mp_temporal(a, n, sx)) return 1;
    SKPS_CUDA(cudaMemcpyAsync(sl.h_count, p->d_count, 4 * n, cudaMemcpyDeviceToHost, sx));
    SKPS_CUDA(cudaMemcpyAsync(sl.h_flag, p->d_flag, 4 * n, cudaMemcpyDeviceToHost, sx));
    SKPS_CUDA(cudaMemcpyAsync(sl.h_det_count, p->d_det_count, 4 * n, cudaMemcpyDeviceToHost, sx));
    SKPS_CUDA(cudaMemcpyAsync(sl.h_box, p->d_track, 8 * 4 * (size_t)K * n, cudaMemcpyDeviceToHost, sx));
    SKPS_CUDA(cudaMemcpyAsync(sl.h_kps, p->d_out_kps, 8 * 2 * (size_t)P * K * n, cudaMemcpyDeviceToHost, sx));
    SKPS_CUDA(cudaMemcpyAsync(sl.h_scores, skps_engine_output_ptr(p->kps, 1), 4 * (size_t)P * K * n, cudaMemcpyDeviceToHost, sx));
    SKPS_CUDA(cudaEventRecord(sl.ev_done, sx));
    // ring discipline: this batch read positions r (previous) and r+1 (current); the next batch uploads into r+2 (free), the
    // one after into r again - and that one reuses this slot, so the caller has passed skps_mpipe_wait(slot) by then
    for (int s = 0; s < n; ++s) {
        p->ring_pos[s] = (p->ring_pos[s] + 1) % 3;
        p->prev_h[s] = hw[2 * s]; p->prev_w[s] = hw[2 * s + 1];
    }
    sl.n = n;
    sl.busy = true;
    return 0;
}

extern "C" SKPS_API int skps_mpipe_wait(skps_mpipe* p, int slot_i, int32_t* n_faces, double* boxes, double* kps, float* scores,
                                        int32_t* ran_detector) {
    SKPS_CHECK(p && (slot_i == 0 || slot_i == 1) && n_faces && boxes && kps && scores, "mpipe_wait: bad arguments");
    skps_mpipe::Slot& sl = p->slot[slot_i];
    SKPS_CHECK(sl.busy, "mpipe_wait: nothing submitted on slot %d", slot_i);
    SKPS_CUDA(cudaSetDevice(p->device));
    SKPS_CUDA(cudaEventSynchronize(sl.ev_done));
    sl.busy = false;
    const int K = p->K, P = p->P, n = sl.n;
    for (int s = 0; s < n; ++s) {
        SKPS_CHECK(!(sl.h_flag[s] && sl.h_det_count[s] < 0), "stream %d: detector produced %d candidates over the score threshold (limit 1024)",
                   s, -sl.h_det_count[s]);
        n_faces[s] = sl.h_count[s];
        if (ran_detector) ran_detector[s] = sl.h_flag[s];
    }
    memcpy(boxes, sl.h_box, sizeof(double) * 4 * K * n);
    memcpy(kps, sl.h_kps, sizeof(double) * 2 * P * K * n);
    memcpy(scores, sl.h_scores, sizeof(float) * P * K * n);
    return 0;
}

extern "C" SKPS_API int skps_mpipe_dims(const skps_mpipe* p, int* n_streams, int* top_k, int* n_points) {
    SKPS_CHECK(p, "mpipe_dims: null");
    if (n_streams) *n_streams = p->S;
    if (top_k) *top_k = p->K;
    if (n_points) *n_points = p->P;
    return 0;
}
