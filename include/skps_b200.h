/*
 * skps_b200 — C-ABI of the B200-native FaceAna hot path.
 *
 * Drop-in boundary for 610265158/Peppa_Pig_Face_Landmark: the reference reaches
 * all of its neural arithmetic through one Python class,
 *     ONNXEngine(onnx_path)(ndarray) -> [ndarray, ...]
 *         (Skps/core/api/onnx_model_base.py:6-27, called from
 *          Skps/core/api/face_detector.py:29 and Skps/core/api/face_landmark.py:48)
 * and all of its image arithmetic through OpenCV/numpy calls in
 *     FaceDetector.preprocess / py_nms / scale_coords   (face_detector.py:45-136)
 *     FaceLandmark.preprocess / postprocess             (face_landmark.py:66-115)
 *     FaceAna.diff_frames / judge_boxs / sort_and_filter (facer.py:98-189)
 * Each entry point below names the reference function it replaces.  Plain
 * pointers and sizes only; `stream` is a cudaStream_t passed as void*.
 * All functions return 0 on success, non-zero on error; the message is
 * available from skps_last_error() (thread-local).
 *
 * Pointers marked [dev] are device pointers on the engine's device, [host] are
 * host pointers (pinned memory makes the copies asynchronous).
 */
#ifndef SKPS_B200_H
#define SKPS_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SKPS_API __attribute__((visibility("default")))
#else
#define SKPS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skps_engine skps_engine;     /* one compiled network  (ONNXEngine,  onnx_model_base.py:6)  */
typedef struct skps_pipeline skps_pipeline; /* detector+landmark     (FaceAna,     facer.py:25)          */

SKPS_API const char* skps_last_error(void);
SKPS_API int skps_version(void);

/* ------------------------------------------------------------------ ONNXEngine (onnx_model_base.py:7-27) */

/* Build an engine from a lowered plan (peppa_pig_face_landmark_b200.lowering.lower(onnx_path)):
 * `plan_words` describe buffers/ops, `weights` is the float32 blob taken unchanged from the
 * .onnx initializers.  Replaces onnxruntime.InferenceSession(onnx_f) (onnx_model_base.py:14). */
SKPS_API int skps_engine_create(const int32_t* plan_words, size_t n_words,
                       const float* weights, size_t n_floats,
                       int max_batch, int device, skps_engine** out);
SKPS_API void skps_engine_destroy(skps_engine* e);

/* Introspection: input is uint8 NHWC (N,H,W,3); output i has `skps_engine_output_elems`
 * float32 per sample. */
SKPS_API int skps_engine_input_dims(const skps_engine* e, int* h, int* w, int* c);
SKPS_API int skps_engine_num_outputs(const skps_engine* e);
SKPS_API int skps_engine_output_elems(const skps_engine* e, int idx);
/* [dev] pointer of the engine-owned input buffer (N,H,W,3) uint8 — producers (letterbox,
 * crop) write here directly so there is no copy between stages. */
SKPS_API void* skps_engine_input_ptr(skps_engine* e);
/* [dev] pointer of the engine-owned output buffer idx, (max_batch, elems) float32. */
SKPS_API float* skps_engine_output_ptr(skps_engine* e, int idx);

/* session.run on device memory (onnx_model_base.py:23): `input` [dev] uint8 NHWC (batch,H,W,3)
 * (may be skps_engine_input_ptr itself), results left in the engine's output buffers and, when
 * outputs[i] != NULL, copied to outputs[i] [dev].  Asynchronous on `stream`. */
SKPS_API int skps_engine_forward(skps_engine* e, const uint8_t* input, int batch,
                        float* const* outputs, void* stream);

/* session.run with HOST buffers, exactly ONNXEngine.__call__'s contract
 * (onnx_model_base.py:17-27): `input` [host] float32 NCHW (batch,3,H,W) in [0,1] as the
 * reference feeds it; outputs[i] [host] float32.  H2D + compute + D2H, synchronous. */
SKPS_API int skps_engine_forward_host_f32(skps_engine* e, const float* input_nchw, int batch,
                                 float* const* outputs, void* stream);
/* Same with uint8 NHWC host input (crops / letterboxed frame before the /255). */
SKPS_API int skps_engine_forward_host_u8(skps_engine* e, const uint8_t* input_nhwc, int batch,
                                float* const* outputs, void* stream);

/* Streaming host round trip (additive): two slots; while one batch computes the next one's pixels cross PCIe on a
 * second stream.  `input` / `outputs[i]` [host, pinned] must stay valid until skps_engine_wait(slot) returns.
 * Order of use: submit(0) submit(1) wait(0) submit(0) wait(1) ... */
SKPS_API int skps_engine_submit_host_u8(skps_engine* e, int slot, const uint8_t* input_nhwc, int batch,
                                        float* const* outputs);
SKPS_API int skps_engine_wait(skps_engine* e, int slot);

/* Debug/parity: copy internal buffer `buf` (N,H,W,C float32 NHWC) of the last forward to host. */
SKPS_API int skps_engine_num_buffers(const skps_engine* e);
SKPS_API int skps_engine_buffer_dims(const skps_engine* e, int buf, int* h, int* w, int* c, int* dtype);
SKPS_API int skps_engine_read_buffer(skps_engine* e, int buf, int batch, void* dst_host);
/* Number of kernels one forward launches (for bench.py's gpu_launches). */
SKPS_API int skps_engine_launches_per_forward(const skps_engine* e);
/* Kernel launches of one forward at `batch` (segments sweep the batch in L2-sized chunks). */
SKPS_API int skps_engine_launches_for_batch(const skps_engine* e, int batch);
/* Profiling: enqueue only op `op_index` of the plan on the buffers left by the last forward
 * (bench.py times the dominant kernel with CUDA events around this call). */
SKPS_API int skps_engine_run_op(skps_engine* e, int op_index, int batch, void* stream);

/* Unit-test entry for the tcgen05 convolution kernel (csrc/conv_tc.cu): one 'same' conv on host
 * data.  x float32 NHWC, w_hi/w_lo float16 (n_tiles*n_tile, K_pad) as packed by
 * plan.pack_tc_weights, out float32 NHWC. */
SKPS_API int skps_debug_conv_tc(const float* x, int N, int H, int W, int Cin, const void* w_hi, const void* w_lo,
                                const float* bias, int Cout, int ksize, int dil, int act, int n_tile, int n_tiles,
                                float out_scale, const float* residual, int out_split, float* out);
/* Same plus the conv stride (1 or 2: TMA element strides; out is N x H/stride x W/stride x Cout), the residual
 * order (res_first != 0: act(conv + bias + residual), the ResNet/HRNet block of the Teacher, model.py:302-345)
 * and the capacity max_batch >= N of the device buffers (maps smaller than 128 pixels share a tile between
 * images; N not a multiple of that exercises the partial tile). */
SKPS_API int skps_debug_conv_tc2(const float* x, int N, int H, int W, int Cin, const void* w_hi, const void* w_lo,
                                 const float* bias, int Cout, int ksize, int dil, int act, int n_tile, int n_tiles,
                                 float out_scale, const float* residual, int out_split, float* out, int stride,
                                 int res_first, int max_batch);

/* Unit-test entry of the transposed heat-map head kernel (csrc/conv_hm.cu; kps graph /student/hm/Conv + the arg-max half of
 * postp, TRAIN/face_landmark/lib/core/base_trainer/model.py:511-554): x float32 NHWC, w_hi/w_lo (n_tile, K_pad) float16 as
 * packed by plan.pack_tc_weights -> per 256-pixel tile and channel the maximum score and its first pixel index (y*W+x):
 * val / idx are [N][H*W/256][128] float32 / int32. */
SKPS_API int skps_debug_conv_hm(const float* x, int N, int H, int W, int Cin, const void* w_hi, const void* w_lo,
                                const float* bias, int Cout, int n_tile, float out_scale, float* val, int* idx);

/* Debug/unit-test entry of the fused producer -> 1x1 conv kernels (csrc/conv_xf.cu): mode 0 = squeeze-excite scale
 * (x * gate[n,c]) ahead of the conv, mode 1 = depthwise 3x3 [over concat(bilinear_x2(low), x)] ahead of the conv.
 * Replaces, for one layer, what onnxruntime runs for the reference's ONNX nodes Mul->Conv / Resize->Concat->Conv(dw)->Conv
 * (Skps/core/api/onnx_model_base.py:23).  Host float32 NHWC in/out; w_hi/w_lo as packed by plan.pack_tc_weights. */
SKPS_API int skps_debug_conv_xf(int mode, const float* x, int N, int H, int W, int Cx, int x_split, const float* low, int Cl,
                                const float* gate, const float* dww, int dw_act, const void* w_hi, const void* w_lo,
                                const float* bias, int Cout, int act, int n_tile, float out_scale, const float* residual,
                                int res_first, int out_split, float* out, const float* weff);

/* Unit-test entries for two fused kernels that otherwise only run inside a whole network (csrc/debug_ops.cu):
 * the squeeze-excite gate (mean of per-tile channel sums -> FC -> act -> FC -> act; kps_student.onnx .../se/ nodes) and the
 * heat-map decode (arg-max + offsets, TRAIN/face_landmark/lib/core/base_trainer/model.py:511-554).  Host float32. */
SKPS_API int skps_debug_se_fc(const float* part, int N, int tiles, int C, const float* w1t, const float* b1, const float* w2t,
                              const float* b2, int Cr, int act1, int act2, int hw, float* gate);
SKPS_API int skps_debug_hm_decode(const float* hm, int N, int H, int W, int ld, int npts, const float* feat, int K,
                                  const float* w_off, const float* b_off, float* xy, float* score);

/* Unit-test entry for the few-channel 3x3 convolution kernel (csrc/conv_mma.cu; Cin == Cout == C in {24, 40}, the
 * Teacher's HRNet branch convs): x, residual, out float32 NHWC; w_packed float16 [tap][hi/lo][C][C] as packed by
 * plan.pack_mma_weights. */
SKPS_API int skps_debug_conv_mma(const float* x, int N, int H, int W, int C, const void* w_packed, const float* bias,
                                 int act, float out_scale, const float* residual, int res_first, int out_split,
                                 float* out);

/* ------------------------------------------------------------------ image kernels */

/* FaceDetector.preprocess (face_detector.py:45-71): BGR->RGB, cv2.resize INTER_LINEAR
 * (bit-exact 11-bit fixed point) to (rw,rh), pad with 114 to (in_h,in_w).  `frame` [dev]
 * HxWx3 uint8 with row pitch `pitch` bytes; `out` [dev] in_h*in_w*3 uint8 RGB.  The /255 is
 * done inside the first conv. */
SKPS_API int skps_letterbox(const uint8_t* frame, int H, int W, int pitch,
                   uint8_t* out, int in_h, int in_w,
                   int rw, int rh, int top, int left, void* stream);

/* xywh2xyxy + py_nms + scale_coords (face_detector.py:73-136) on the raw (rows,16) output.
 * Writes up to max_det kept rows (16 floats each, cols 0-3 mapped back to frame pixels),
 * their row indices into the raw output, and the count.  All [dev]. */
SKPS_API int skps_detect_post(const float* raw, int rows, float score_thres, float iou_thres,
                     float scale, float pad_x, float pad_y,
                     float* kept_rows, int32_t* kept_idx, int32_t* count, int max_det,
                     void* stream);

/* FaceAna.judge_boxs + sort_and_filter (facer.py:120-189): IoU-match detections against the
 * previous track boxes (EMA alpha), drop area<=min_face, keep top_k by area.  `track` [dev]
 * (n_track,4) may be NULL.  Writes (count,4) boxes. */
SKPS_API int skps_select_faces(const float* det_rows, const int32_t* det_count, int det_stride,
                      const float* track, int n_track,
                      float iou_thres, float alpha, float one_minus_alpha,
                      float min_face, int top_k,
                      float* boxes4, int32_t* count, void* stream);

/* FaceLandmark.preprocess (face_landmark.py:66-104) for all faces at once: zero-pad, square
 * crop, cv2.resize to (out_hw,out_hw), uint8 BGR NHWC.  `boxes4` [dev] (max_faces,4), `count`
 * [dev].  `detail` [dev] (max_faces,5) int32 = [h,w,y1,x1,add].  Faces >= *count are zeroed. */
SKPS_API int skps_crop_resize(const uint8_t* frame, int H, int W, int pitch,
                     const float* boxes4, const int32_t* count, int max_faces,
                     float face_scale /* float32(1+2*extend) */, float min_face,
                     uint8_t* crops, int out_hw, int32_t* detail, void* stream);

/* FaceLandmark.postprocess (face_landmark.py:106-115): x*w + x1 - add, y*h + y1 - add. */
SKPS_API int skps_landmark_post(const float* xy_norm, const int32_t* detail, const int32_t* count,
                       int max_faces, int n_points, float* kps, void* stream);

/* FaceAna.diff_frames (facer.py:98-118): sum |a-b| over n bytes into *sum [dev] (uint64). */
SKPS_API int skps_frame_absdiff_sum(const uint8_t* a, const uint8_t* b, size_t n,
                           unsigned long long* sum, void* stream);

/* ------------------------------------------------------------------ FaceAna.run (facer.py:52-85) */

typedef struct skps_pipeline_cfg {
    float score_thres, iou_thres;       /* Skps.yml Detect.score_thrs / iou_thrs         */
    float min_face;                     /* Detect.min_face (area)                        */
    int   top_k;                        /* Detect.topk                                   */
    float track_iou, alpha;             /* Trace.iou_thres / smooth_box                  */
    float face_scale;                   /* float32(1 + 2*Keypoints.base_extend_range[0]) */
    float kps_min_face;                 /* FaceLandmark.min_face (20)                    */
    int   max_h, max_w;                 /* largest frame accepted                        */
} skps_pipeline_cfg;

SKPS_API int skps_pipeline_create(skps_engine* det, skps_engine* kps, const skps_pipeline_cfg* cfg,
                         skps_pipeline** out);
SKPS_API void skps_pipeline_destroy(skps_pipeline* p);
/* FaceAna.reset (facer.py:200-208): forget the previous frame. */
SKPS_API int skps_pipeline_reset(skps_pipeline* p);

/* One frame, detector path of FaceAna.run (facer.py:56-68): letterbox -> detector -> NMS ->
 * judge_boxs(track) -> sort_and_filter -> crops -> landmark net -> de-normalise, with no host
 * round trip in between.  `frame` [host] (or [dev] when frame_on_device) HxWx3 uint8 BGR;
 * letterbox geometry (rw,rh,top,left,scale) is computed by the caller exactly as
 * face_detector.py:51-62 does.  `track` [host] (n_track,4) float32 previous track boxes or
 * NULL.  If run_detector==0 the `track` boxes are used as the face boxes (facer.py:61).
 * Results [host]: n_faces, boxes (top_k,4) — the boxes handed to the landmark stage
 * (facer.py:66 boxes_return), kps (top_k,n_points,2), scores (top_k,n_points),
 * det_idx (top_k... max_det) kept detector row indices (parity checks).  Synchronous. */
SKPS_API int skps_pipeline_run(skps_pipeline* p, const uint8_t* frame, int H, int W, int frame_on_device,
                      int run_detector, int rw, int rh, int top, int left, float scale,
                      const float* track, int n_track,
                      int32_t* n_faces, float* boxes4, float* kps, float* scores,
                      int32_t* n_det, int32_t* det_idx, float* det_rows, void* stream);

/* Mean absolute frame difference vs the previously submitted frame (facer.py:111-113);
 * returns -1.0 in *mean_diff when there is no previous frame of the same size. */
SKPS_API int skps_pipeline_frame_diff(skps_pipeline* p, const uint8_t* frame, int H, int W,
                             int frame_on_device, double* mean_diff, void* stream);
/* Adopt the frame staged by skps_pipeline_frame_diff as the previous frame without running the chain: the skip path of
 * FaceAna.run (facer.py:57-62 replaces previous_image on every call, also when nothing is detected or tracked). */
SKPS_API int skps_pipeline_commit_frame(skps_pipeline* p, int H, int W);

/* WFLW evaluation helpers (TRAIN/face_landmark/tools/eval_WFLW.py).  skps_crop_rect: zero-bordered rectangular crop of a
 * [dev] BGR frame resized to out_hw x out_hw, bit-exact with copyMakeBorder + slicing + cv2.resize (:38-80, :113-124).
 * skps_nme: per-face normalised mean error, mean_p |pred - gt| / |gt[norm_a] - gt[norm_b]| (:84-95; WFLW: 60, 72); all [dev]. */
SKPS_API int skps_crop_rect(const uint8_t* frame, int H, int W, int pitch, int rx, int ry, int rw, int rh, uint8_t* out,
                            int out_hw, void* stream);
SKPS_API int skps_nme(const float* target, const float* preds, int n, int n_points, int norm_a, int norm_b, float* out,
                      void* stream);

/* Batched head pose (csrc/headpose.cu), the GPU counterpart of Skps/core/headpose/pose.py:48-77 get_head_pose():
 * solvePnP (iterative) on 10 landmark/model point pairs with the camera matrix [[w,0,w//2],[0,w,h//2],[0,0,1]], projection
 * of the 8 cube corners, Euler angles as cv2.decomposeProjectionMatrix reports them.  pts [host] (N,10,2) float32 in the
 * order of pose.py:60-61; object_pts (10,3), cube_pts (8,3) float32.  Outputs [host] float64: rvec, tvec, euler (N,3),
 * reproject (N,8,2). */
SKPS_API int skps_head_pose(const float* pts, int N, int img_w, int img_h, const float* object_pts, const float* cube_pts,
                            double* rvec, double* tvec, double* euler, double* reproject);

/* ---- FaceAna.run for many concurrent video streams (csrc/mpipe.cu) ---------------------------------
 * What one FaceAna instance per stream does on the host in the reference (facer.py:52-85 with GroupTrack / OneEuroFilter /
 * EmaFilter of Skps/core/smoother/lk.py:6-162) happens here for up to n_streams streams per call with all per-stream state
 * on the device: one detector forward and one landmark forward per call, batched over the streams.
 * det needs max_batch >= n_streams, kps needs max_batch >= n_streams * cfg->top_k.  Not thread-safe per handle. */
typedef struct skps_mpipe skps_mpipe;
SKPS_API int skps_mpipe_create(skps_engine* det, skps_engine* kps, const skps_pipeline_cfg* cfg, int n_streams,
                               skps_mpipe** out);
SKPS_API void skps_mpipe_destroy(skps_mpipe* p);
/* FaceAna.reset() for one stream (or all: stream = -1): forget the previous frame, the track boxes, the landmark history. */
SKPS_API int skps_mpipe_reset(skps_mpipe* p, int stream);
SKPS_API int skps_mpipe_dims(const skps_mpipe* p, int* n_streams, int* top_k, int* n_points);
/* Enqueue frame i (HxWx3 uint8 BGR, [host] pinned or pageable, or [dev]) of stream i for i < n; hw = {H0,W0,H1,W1,...}.
 * slot in {0,1}: submit(0) submit(1) wait(0) submit(0) ... keeps two batches in flight (uploads overlap compute).
 * Pinned frames must stay valid until skps_mpipe_wait(slot).  Asynchronous. */
SKPS_API int skps_mpipe_submit(skps_mpipe* p, int slot, const uint8_t* const* frames, const int32_t* hw, int n,
                               int frames_on_device);
/* Block until the slot's results are in host memory.  Per stream s < n: n_faces[s]; boxes (n, top_k, 4) float64 = the
 * refreshed track boxes (the 'box' entries of FaceAna.run); kps (n, top_k, n_points, 2) float64 smoothed landmarks;
 * scores (n, top_k, n_points) float32; ran_detector[s] (may be NULL) = the frame-difference gate's decision. */
SKPS_API int skps_mpipe_wait(skps_mpipe* p, int slot, int32_t* n_faces, double* boxes, double* kps, float* scores,
                             int32_t* ran_detector);

#ifdef __cplusplus
}
#endif
#endif /* SKPS_B200_H */
