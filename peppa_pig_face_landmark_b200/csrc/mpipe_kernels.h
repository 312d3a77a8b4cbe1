// Kernels of the multi-stream pipeline (mpipe.cu): launchers defined in image_ops.cu / temporal.cu.
#pragma once
#include <cuda_runtime.h>

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

int launch_mp_select(const float* det_rows, const int* det_count, int max_det, const int* flag, const float* track,
                     const int* n_track, float iou_thres, float alpha, float oma, float min_face, int top_k, float* boxes4,
                     int* count, int n_streams, cudaStream_t s);
int launch_mp_decide(const unsigned long long* diff, const int* hw, const int* have_prev, int* flag, int n, cudaStream_t s);
int launch_mp_temporal(const MpTemporalArgs& a, int n_streams, cudaStream_t s);

}  // namespace skps
