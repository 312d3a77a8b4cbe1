// Kernels of the multi-stream pipeline (mpipe.cu): launchers defined in image_ops.cu / temporal.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace skps {

struct MpTemporalArgs {
    int top_k, n_points;
    const float* kps_now;        // [S][K][P][2] float32 landmarks in frame pixels (landmark_post)
    const int* count;            // [S] faces this frame
    const int* flag;             // [S] detector ran this frame
    const int* hw;               // [S][2] frame height, width
    const float* boxes4;         // [S][K][4] boxes the landmark stage used (boxes_return, facer.py:66)
    // state, updated in place
    double* prev_lm;             // [S][2][K][P][2] previous landmark sets (ping-pong)
    double* prev_dx;             // [S][2][K][P][2] previous - filtered
    int* n_prev;                 // [S] sets in prev_lm (-1 = None)
    int* prev_f32;               // [S] previous_landmarks_set is a float32 array (numpy dtype bookkeeping)
    int* state_idx;              // [S] which half of prev_lm/prev_dx is current
    double* track_box;           // [S][K][4] float64 track boxes (returned as 'box')
    float* track_f32;            // [S][K][4] the same, as float32 (next frame's judge_boxs / crop input)
    int* n_track;                // [S]
    // outputs
    double* out_kps;             // [S][K][P][2]
    // constants (python floats computed on the host exactly as lk.py does)
    double iou_thres, alpha, one_minus_alpha, a_d, one_minus_a_d, min_cutoff, beta, two_pi;
};

// One frame of one stream as the batched pre/post-processing kernels see it (filled on the host per call, one upload).
struct MpStreamDesc {
    const uint8_t* cur;          // this call's frame (HxWx3 uint8 BGR, device)
    const uint8_t* prev;         // the previous frame of the stream, or null
    int H, W, have_prev;
    float scale;                 // letterbox geometry (face_detector.py:49-62)
    int rw, rh, top, left;
};
// Batched over the streams of a call (grid z / x = stream): what S x {skps_frame_absdiff_sum, skps_letterbox, skps_detect_post,
// skps_crop_resize, skps_landmark_post} launches did, in five launches.  Same device code per element, bit for bit.
int launch_mp_absdiff(const MpStreamDesc* d, unsigned long long* diff, int n, size_t max_bytes, cudaStream_t s);
int launch_mp_letterbox(const MpStreamDesc* d, uint8_t* out, size_t out_stride, int in_h, int in_w, int n, cudaStream_t s);
int launch_mp_detect_post(const MpStreamDesc* d, const float* raw, int rows, float score_thres, float iou_thres, float* kept_rows,
                          int* kept_idx, int* count, int max_det, int n, cudaStream_t s);
int launch_mp_crop(const MpStreamDesc* d, const float* boxes, const int* count, int K, float face_scale, float min_face,
                   uint8_t* crops, int S, int* detail, int n, cudaStream_t s);
int launch_mp_landmark_post(const float* xy, const int* detail, const int* count, int K, int P, float* kps, int n, cudaStream_t s);

int launch_mp_select(const float* det_rows, const int* det_count, int max_det, const int* flag, const float* track,
                     const int* n_track, float iou_thres, float alpha, float oma, float min_face, int top_k, float* boxes4,
                     int* count, int n_streams, cudaStream_t s);
int launch_mp_decide(const unsigned long long* diff, const int* hw, const int* have_prev, int* flag, int n, cudaStream_t s);
int launch_mp_temporal(const MpTemporalArgs& a, int n_streams, cudaStream_t s);

}  // namespace skps
