// Image-side kernels of the FaceAna path (sm_100a): letterbox, per-face crop+resize (both bit-exact
// with cv2.resize INTER_LINEAR on uint8), detector post-processing (score filter, greedy NMS,
// un-letterbox), face selection (IoU track match + EMA, area filter, top-k), landmark
// de-normalisation and the frame-difference gate.  All HBM-bound byte/index work; compiled with
// -fmad=false so float32 expressions round exactly like the numpy expressions they restate.
#include "../../include/skps_b200.h"
#include "common.h"
#include "mpipe_kernels.h"

namespace skps {

// ------------------------------------------------------------------------------------------
// cv2.resize(INTER_LINEAR, uint8) tap: OpenCV resize.cpp (called from face_detector.py:53 and
// face_landmark.py:97).  scale = 1/(dst/src) in double, offset rounded to float32, weights
// rint(w*2048) (INTER_RESIZE_COEF_BITS = 11).
// ------------------------------------------------------------------------------------------
struct Tap { int i0, i1, w0, w1; };

__device__ __forceinline__ Tap linear_tap(int d, int dst, int src, bool is_x) {
    double scale = 1.0 / ((double)dst / (double)src);
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    Tap t;
    if (is_x) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= src - 1) { s = src - 1; f = 0.f; }
        t.i0 = s;
        t.i1 = min(s + 1, src - 1);
    } else {
        t.i0 = min(max(s, 0), src - 1);
        t.i1 = min(max(s + 1, 0), src - 1);
    }
    t.w0 = __float2int_rn((1.f - f) * 2048.f);
    t.w1 = __float2int_rn(f * 2048.f);
    return t;
}

__device__ __forceinline__ int vblend(int h0, int h1, int b0, int b1) {
    return (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
}

// ------------------------------------------------------------------------------------------
// Letterbox (face_detector.py:45-71): BGR->RGB, resize to (rw,rh), pad 114.  One thread per
// output pixel (3 channels); output is uint8 RGB NHWC, /255 happens in the first conv.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void letterbox_px(const uint8_t* __restrict__ frame, int H, int W, int pitch,
                                             uint8_t* __restrict__ out, int in_h, int in_w, int rw, int rh, int top, int left,
                                             int x, int y) {
    if (x >= in_w) return;
    uint8_t* o = out + ((long long)y * in_w + x) * 3;
    int dx = x - left, dy = y - top;
    if (dx < 0 || dx >= rw || dy < 0 || dy >= rh) {
        o[0] = 114; o[1] = 114; o[2] = 114;
        return;
    }
    Tap tx = linear_tap(dx, rw, W, true);
    Tap ty = linear_tap(dy, rh, H, false);
    const uint8_t* r0 = frame + (long long)ty.i0 * pitch;
    const uint8_t* r1 = frame + (long long)ty.i1 * pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int h0 = r0[tx.i0 * 3 + c] * tx.w0 + r0[tx.i1 * 3 + c] * tx.w1;
        int h1 = r1[tx.i0 * 3 + c] * tx.w0 + r1[tx.i1 * 3 + c] * tx.w1;
        o[2 - c] = (uint8_t)vblend(h0, h1, ty.w0, ty.w1);      // BGR -> RGB
    }
}
__global__ void __launch_bounds__(256) letterbox_kernel(const uint8_t* __restrict__ frame, int H, int W, int pitch,
                                                        uint8_t* __restrict__ out, int in_h, int in_w,
                                                        int rw, int rh, int top, int left) {
    letterbox_px(frame, H, W, pitch, out, in_h, in_w, rw, rh, top, left, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y);
}
__global__ void __launch_bounds__(256) mp_letterbox_kernel(const MpStreamDesc* __restrict__ d, uint8_t* __restrict__ out,
                                                           size_t out_stride, int in_h, int in_w) {
    const MpStreamDesc D = d[blockIdx.z];
    letterbox_px(D.cur, D.H, D.W, D.W * 3, out + out_stride * blockIdx.z, in_h, in_w, D.rw, D.rh, D.top, D.left,
                 blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y);
}

// ------------------------------------------------------------------------------------------
// Per-face crop + resize (face_landmark.py:74-98).  Geometry in float32 exactly as numpy>=2
// evaluates it; the zero border of copyMakeBorder is virtual.
// ------------------------------------------------------------------------------------------
struct CropGeo { int add, x1, y1, w, h, ok; };

__device__ __forceinline__ CropGeo crop_geometry(const float* b, int H, int W, float face_scale, float min_face) {
    CropGeo g;
    float bw = b[2] - b[0], bh = b[3] - b[1];
    g.ok = !(bw <= min_face || bh <= min_face);
    int add = (int)fmaxf(bw, bh);
    float fa = (float)add;
    float x0 = b[0] + fa, y0 = b[1] + fa, x1 = b[2] + fa, y1 = b[3] + fa;
    float fw = face_scale * bw;
    float cx = floorf((x0 + x1) / 2.f), cy = floorf((y0 + y1) / 2.f);
    float half = floorf(fw / 2.f);
    int ix1 = (int)(cx - half), iy1 = (int)(cy - half), ix2 = (int)(cx + half), iy2 = (int)(cy + half);
    // numpy slicing of the padded frame clamps the ends (negative starts are not meaningful in the
    // reference either; clamp them to 0)
    int PW = W + 2 * add, PH = H + 2 * add;
    ix1 = max(ix1, 0); iy1 = max(iy1, 0);
    ix2 = min(ix2, PW); iy2 = min(iy2, PH);
    g.add = add; g.x1 = ix1; g.y1 = iy1;
    g.w = max(ix2 - ix1, 0); g.h = max(iy2 - iy1, 0);
    if (g.w <= 0 || g.h <= 0) g.ok = 0;
    return g;
}

__device__ __forceinline__ void crop_px(const uint8_t* __restrict__ frame, int H, int W, int pitch,
                                        const float* __restrict__ boxes, const int* __restrict__ count, float face_scale,
                                        float min_face, uint8_t* __restrict__ crops, int S, int* __restrict__ detail,
                                        int face, int x, int y) {
    if (x >= S) return;
    uint8_t* o = crops + (((long long)face * S + y) * S + x) * 3;
    const int n = *count;
    if (face >= n) {
        o[0] = 0; o[1] = 0; o[2] = 0;
        if (x == 0 && y == 0) { for (int k = 0; k < 5; ++k) detail[face * 5 + k] = 0; }
        return;
    }
    CropGeo g = crop_geometry(boxes + face * 4, H, W, face_scale, min_face);
    if (x == 0 && y == 0) {
        detail[face * 5 + 0] = g.h; detail[face * 5 + 1] = g.w;
        detail[face * 5 + 2] = g.y1; detail[face * 5 + 3] = g.x1; detail[face * 5 + 4] = g.add;
    }
    if (!g.ok) { o[0] = 0; o[1] = 0; o[2] = 0; return; }
    Tap tx = linear_tap(x, S, g.w, true);
    Tap ty = linear_tap(y, S, g.h, false);
    // crop coordinates -> frame coordinates (outside the frame = the zero border)
    int fx0 = g.x1 + tx.i0 - g.add, fx1 = g.x1 + tx.i1 - g.add;
    int fy0 = g.y1 + ty.i0 - g.add, fy1 = g.y1 + ty.i1 - g.add;
    bool vx0 = fx0 >= 0 && fx0 < W, vx1 = fx1 >= 0 && fx1 < W;
    bool vy0 = fy0 >= 0 && fy0 < H, vy1 = fy1 >= 0 && fy1 < H;
    const uint8_t* r0 = frame + (long long)(vy0 ? fy0 : 0) * pitch;
    const uint8_t* r1 = frame + (long long)(vy1 ? fy1 : 0) * pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int p00 = (vy0 && vx0) ? r0[fx0 * 3 + c] : 0, p01 = (vy0 && vx1) ? r0[fx1 * 3 + c] : 0;
        int p10 = (vy1 && vx0) ? r1[fx0 * 3 + c] : 0, p11 = (vy1 && vx1) ? r1[fx1 * 3 + c] : 0;
        int h0 = p00 * tx.w0 + p01 * tx.w1;
        int h1 = p10 * tx.w0 + p11 * tx.w1;
        o[c] = (uint8_t)vblend(h0, h1, ty.w0, ty.w1);          // stays BGR (face_landmark.py:44)
    }
}
__global__ void __launch_bounds__(256) crop_resize_kernel(const uint8_t* __restrict__ frame, int H, int W, int pitch,
                                                          const float* __restrict__ boxes, const int* __restrict__ count,
                                                          float face_scale, float min_face,
                                                          uint8_t* __restrict__ crops, int S, int* __restrict__ detail) {
    crop_px(frame, H, W, pitch, boxes, count, face_scale, min_face, crops, S, detail, blockIdx.z, blockIdx.x * blockDim.x + threadIdx.x,
            blockIdx.y);
}
__global__ void __launch_bounds__(256) mp_crop_kernel(const MpStreamDesc* __restrict__ d, const float* __restrict__ boxes,
                                                      const int* __restrict__ count, int K, float face_scale, float min_face,
                                                      uint8_t* __restrict__ crops, int S, int* __restrict__ detail) {
    const int st = blockIdx.z / K, face = blockIdx.z - st * K;
    const MpStreamDesc D = d[st];
    crop_px(D.cur, D.H, D.W, D.W * 3, boxes + (size_t)4 * K * st, count + st, face_scale, min_face,
            crops + (size_t)S * S * 3 * K * st, S, detail + (size_t)5 * K * st, face, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y);
}

// ------------------------------------------------------------------------------------------
// Rectangular crop + resize of one image (WFLW evaluation, TRAIN/face_landmark/tools/eval_WFLW.py:38-80,113-124:
// copyMakeBorder(zero) -> img[min_y:max_y, min_x:max_x] -> cv2.resize((S, S))).  rect = [x0, y0, w, h] in frame
// coordinates; pixels outside the frame are the zero border.  Same fixed-point bilinear as above, bit-exact with OpenCV.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rect_resize_kernel(const uint8_t* __restrict__ frame, int H, int W, int pitch,
                                                          int rx, int ry, int rw, int rh, uint8_t* __restrict__ out, int S) {
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= S) return;
    uint8_t* o = out + ((long long)y * S + x) * 3;
    Tap tx = linear_tap(x, S, rw, true);
    Tap ty = linear_tap(y, S, rh, false);
    int fx0 = rx + tx.i0, fx1 = rx + tx.i1, fy0 = ry + ty.i0, fy1 = ry + ty.i1;
    bool vx0 = fx0 >= 0 && fx0 < W, vx1 = fx1 >= 0 && fx1 < W;
    bool vy0 = fy0 >= 0 && fy0 < H, vy1 = fy1 >= 0 && fy1 < H;
    const uint8_t* r0 = frame + (long long)(vy0 ? fy0 : 0) * pitch;
    const uint8_t* r1 = frame + (long long)(vy1 ? fy1 : 0) * pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int p00 = (vy0 && vx0) ? r0[fx0 * 3 + c] : 0, p01 = (vy0 && vx1) ? r0[fx1 * 3 + c] : 0;
        int p10 = (vy1 && vx0) ? r1[fx0 * 3 + c] : 0, p11 = (vy1 && vx1) ? r1[fx1 * 3 + c] : 0;
        o[c] = (uint8_t)vblend(p00 * tx.w0 + p01 * tx.w1, p10 * tx.w0 + p11 * tx.w1, ty.w0, ty.w1);
    }
}

// Normalised mean error per face (eval_WFLW.py:84-95): mean_p |pred_p - gt_p| / |gt_60 - gt_72| (inter-ocular), float32
// like numpy's; one warp per face.
__global__ void nme_kernel(const float* __restrict__ target, const float* __restrict__ preds, int n, int P, int ia, int ib,
                           float* __restrict__ out) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n) return;
    const float* t = target + (long long)f * P * 2;
    const float* q = preds + (long long)f * P * 2;
    float s = 0.f;
    for (int p = lane; p < P; p += 32) {
        const float dx = q[2 * p] - t[2 * p], dy = q[2 * p + 1] - t[2 * p + 1];
        s += sqrtf(dx * dx + dy * dy);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const float nx = t[2 * ia] - t[2 * ib], ny = t[2 * ia + 1] - t[2 * ib + 1];
        out[f] = (s / (float)P) / sqrtf(nx * nx + ny * ny);
    }
}

// ------------------------------------------------------------------------------------------
// Detector post-processing (face_detector.py:31-37, 73-136).  One block.
//   1. rows with obj > score_thres -> xyxy candidates (more than MAXC: count = -candidates, nothing else written)
//   2. order = (score desc, row index desc)  [np.argsort(score)[::-1]; ties measure-zero]
//   3. greedy NMS: survivors are those with iou < iou_thres against every kept box
//   4. kept rows copied out with cols 0-3 mapped back: (v - pad) / scale
// ------------------------------------------------------------------------------------------
constexpr int MAXC = 1024;

__device__ __forceinline__ float iou_nms(const float4 a, const float4 b) {
    // face_detector.py:117-130, float32 throughout
    float area = (a.z - a.x) * (a.w - a.y);
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float inter = fmaxf(0.f, yy2 - yy1) * fmaxf(0.f, xx2 - xx1);
    float other = (b.w - b.y) * (b.z - b.x);
    return inter / (area + other - inter);
}

__device__ __forceinline__ void detect_post_body(const float* __restrict__ raw, int rows,
                                                           float score_thres, float iou_thres,
                                                           float scale, float pad_x, float pad_y,
                                                           float* __restrict__ kept_rows, int* __restrict__ kept_idx,
                                                           int* __restrict__ count, int max_det) {
    __shared__ int s_n;
    __shared__ int s_cand[MAXC];         // row index of candidate
    __shared__ float s_score[MAXC];
    __shared__ int s_order[MAXC];        // candidate slot by rank
    __shared__ float4 s_box[MAXC];       // xyxy by rank
    __shared__ unsigned char s_dead[MAXC];
    __shared__ int s_keep[256];
    __shared__ int s_nkeep;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { s_n = 0; s_nkeep = 0; }
    __syncthreads();
    for (int r = tid; r < rows; r += nt) {
        float sc = raw[(long long)r * 16 + 4];
        if (sc > score_thres) {
            int slot = atomicAdd(&s_n, 1);
            if (slot < MAXC) { s_cand[slot] = r; s_score[slot] = sc; }
        }
    }
    __syncthreads();
    if (s_n > MAXC) {
        // more candidates than the kernel can rank: refuse (count = -candidates) rather than drop an order-dependent subset;
        // the host raises.  The reference has no cap (face_detector.py:95-136); 1024 rows over obj 0.5 is a noise frame.
        if (tid == 0) *count = -s_n;
        return;
    }
    const int n = s_n;
    // rank sort: key (score desc, row desc) is a total order, so the result is deterministic
    for (int i = tid; i < n; i += nt) {
        float si = s_score[i];
        int ri = s_cand[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            float sj = s_score[j];
            rank += (sj > si) || (sj == si && s_cand[j] > ri);
        }
        s_order[rank] = i;
    }
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
        const float* r = raw + (long long)s_cand[s_order[k]] * 16;
        float hw = r[2] / 2.f, hh = r[3] / 2.f;                      // xywh2xyxy, face_detector.py:76-79
        s_box[k] = make_float4(r[0] - hw, r[1] - hh, r[0] + hw, r[1] + hh);
        s_dead[k] = 0;
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        if (s_dead[i]) continue;                 // uniform: shared value, read after barrier
        if (tid == 0 && s_nkeep < max_det) s_keep[s_nkeep++] = i;
        const float4 cur = s_box[i];
        for (int j = i + 1 + tid; j < n; j += nt) {
            if (!s_dead[j]) {
                float iou = iou_nms(cur, s_box[j]);
                if (!(iou < iou_thres)) s_dead[j] = 1;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    const int nk = s_nkeep;
    if (tid == 0) *count = nk;
    for (int e = tid; e < nk * 16; e += nt) {
        int k = e / 16, c = e % 16;
        int rank = s_keep[k];
        int row = s_cand[s_order[rank]];
        float v;
        if (c < 4) {
            float4 b = s_box[rank];
            float bv = c == 0 ? b.x : (c == 1 ? b.y : (c == 2 ? b.z : b.w));
            v = (bv - ((c & 1) ? pad_y : pad_x)) / scale;           // scale_coords, face_detector.py:86-91
        } else {
            v = raw[(long long)row * 16 + c];
        }
        kept_rows[k * 16 + c] = v;
        if (c == 0) kept_idx[k] = row;
    }
}
__global__ void __launch_bounds__(1024) detect_post_kernel(const float* __restrict__ raw, int rows, float score_thres,
                                                           float iou_thres, float scale, float pad_x, float pad_y,
                                                           float* __restrict__ kept_rows, int* __restrict__ kept_idx,
                                                           int* __restrict__ count, int max_det) {
    detect_post_body(raw, rows, score_thres, iou_thres, scale, pad_x, pad_y, kept_rows, kept_idx, count, max_det);
}
__global__ void __launch_bounds__(1024) mp_detect_post_kernel(const MpStreamDesc* __restrict__ d, const float* __restrict__ raw,
                                                              int rows, float score_thres, float iou_thres,
                                                              float* __restrict__ kept_rows, int* __restrict__ kept_idx,
                                                              int* __restrict__ count, int max_det) {
    const int st = blockIdx.x;
    const MpStreamDesc D = d[st];
    detect_post_body(raw + (size_t)rows * 16 * st, rows, score_thres, iou_thres, D.scale, (float)D.left, (float)D.top,
                     kept_rows + (size_t)16 * max_det * st, kept_idx + (size_t)max_det * st, count + st, max_det);
}

// ------------------------------------------------------------------------------------------
// judge_boxs + sort_and_filter (facer.py:120-189).  One warp; K <= 256 detections.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float iou_track(const float* r1, const float* r2) {
    float s1 = (r1[2] - r1[0]) * (r1[3] - r1[1]);
    float s2 = (r2[2] - r2[0]) * (r2[3] - r2[1]);
    float sum = s1 + s2;
    float x1 = fmaxf(r1[0], r2[0]), y1 = fmaxf(r1[1], r2[1]);
    float x2 = fminf(r1[2], r2[2]), y2 = fminf(r1[3], r2[3]);
    float inter = fmaxf(0.f, x2 - x1) * fmaxf(0.f, y2 - y1);
    return inter / (sum - inter);
}

__device__ __forceinline__ void select_faces_body(const float* __restrict__ det, int n_det, int det_stride,
                                                  const float* __restrict__ track, int n_track, float iou_thres, float alpha,
                                                  float oma, float min_face, int top_k, float* __restrict__ boxes4,
                                                  int* __restrict__ count) {
    __shared__ float s_box[256][4];
    __shared__ float s_area[256];
    __shared__ int s_sel[256];
    __shared__ int s_m;
    const int tid = threadIdx.x;
    const int n = min(n_det, 256);
    if (tid < n) {
        const float* now = det + (long long)tid * det_stride;
        float b[4] = {now[0], now[1], now[2], now[3]};
        for (int j = 0; j < n_track; ++j) {                           // facer.py:176-181: first match wins
            const float* prev = track + j * 4;
            if (iou_track(now, prev) > iou_thres) {
                for (int c = 0; c < 4; ++c) b[c] = alpha * now[c] + oma * prev[c];   // lk.py:95-96
                break;
            }
        }
        for (int c = 0; c < 4; ++c) s_box[tid][c] = b[c];
        s_area[tid] = (b[2] - b[0]) * (b[3] - b[1]);
    }
    __syncthreads();
    if (tid == 0) {
        // area filter keeps detector order (facer.py:132-136)
        int m = 0;
        for (int i = 0; i < n; ++i)
            if (s_area[i] > min_face) s_sel[m++] = i;
        if (m > top_k) {
            // top_k largest areas, descending (facer.py:138: area.argsort()[-k:][::-1]); ties by later index
            // first, matching a stable ascending sort read backwards.
            int picked[64];
            for (int k = 0; k < top_k; ++k) {
                int best = -1;
                for (int q = 0; q < m; ++q) {
                    int i = s_sel[q];
                    if (i < 0) continue;
                    if (best < 0 || s_area[i] >= s_area[s_sel[best]]) best = q;
                }
                picked[k] = s_sel[best];
                s_sel[best] = -1;
            }
            for (int k = 0; k < top_k; ++k) s_sel[k] = picked[k];
            m = top_k;
        }
        s_m = m;
        *count = m;
    }
    __syncthreads();
    const int m = s_m;
    if (tid < m * 4) boxes4[tid] = s_box[s_sel[tid / 4]][tid % 4];
}

__global__ void __launch_bounds__(256) select_faces_kernel(const float* __restrict__ det, const int* __restrict__ det_count,
                                                           int det_stride, const float* __restrict__ track, int n_track,
                                                           float iou_thres, float alpha, float oma, float min_face,
                                                           int top_k, float* __restrict__ boxes4, int* __restrict__ count) {
    select_faces_body(det, *det_count, det_stride, track, n_track, iou_thres, alpha, oma, min_face, top_k, boxes4, count);
}

// Multi-stream variant (mpipe.cu): block = stream.  flag[s] != 0: this frame ran the detector -> judge_boxs(track, det rows)
// (facer.py:58); else boxes = the stream's track boxes (facer.py:61).  Track boxes and their count live on the device.
__global__ void __launch_bounds__(256) mp_select_kernel(const float* __restrict__ det_rows, const int* __restrict__ det_count,
                                                        int max_det, const int* __restrict__ flag,
                                                        const float* __restrict__ track, const int* __restrict__ n_track,
                                                        float iou_thres, float alpha, float oma, float min_face, int top_k,
                                                        float* __restrict__ boxes4, int* __restrict__ count) {
    const int s = blockIdx.x;
    const float* trk = track + (long long)s * top_k * 4;
    if (flag[s])
        select_faces_body(det_rows + (long long)s * max_det * 16, det_count[s], 16, trk, n_track[s], iou_thres, alpha, oma,
                          min_face, top_k, boxes4 + (long long)s * top_k * 4, count + s);
    else
        select_faces_body(trk, n_track[s], 4, nullptr, 0, iou_thres, alpha, oma, min_face, top_k,
                          boxes4 + (long long)s * top_k * 4, count + s);
}

// ------------------------------------------------------------------------------------------
// FaceLandmark.postprocess (face_landmark.py:106-115): float32 product, then + x1 - add in
// float64, stored as float32 (numpy>=2 promotion of `float32 * int + np.int32 - int`).
// ------------------------------------------------------------------------------------------
__global__ void landmark_post_kernel(const float* __restrict__ xy, const int* __restrict__ detail,
                                     const int* __restrict__ count, int max_faces, int P, float* __restrict__ kps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_faces * P) return;
    int f = i / P;
    float ox = 0.f, oy = 0.f;
    if (f < *count) {
        const int* d = detail + f * 5;      // [h, w, y1, x1, add]
        float px = xy[i * 2] * (float)d[1];
        float py = xy[i * 2 + 1] * (float)d[0];
        ox = (float)((double)px + (double)d[3] - (double)d[4]);
        oy = (float)((double)py + (double)d[2] - (double)d[4]);
    }
    kps[i * 2] = ox;
    kps[i * 2 + 1] = oy;
}

// ------------------------------------------------------------------------------------------
// Frame difference gate (facer.py:111-113): sum |a-b| over all bytes.  uint4 loads, __vsadu4.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absdiff_sum_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                          size_t n, unsigned long long* __restrict__ sum) {
    size_t nv = n / 16;
    unsigned long long local = 0;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    const uint4* b4 = reinterpret_cast<const uint4*>(b);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        uint4 x = a4[i], y = b4[i];
        local += __vsadu4(x.x, y.x) + __vsadu4(x.y, y.y) + __vsadu4(x.z, y.z) + __vsadu4(x.w, y.w);
    }
    if (blockIdx.x == 0) {
        for (size_t i = nv * 16 + threadIdx.x; i < n; i += blockDim.x) {
            int d = (int)a[i] - (int)b[i];
            local += (unsigned)(d < 0 ? -d : d);
        }
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ unsigned long long warp_sum[8];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += warp_sum[w];
        atomicAdd(sum, t);
    }
}

__global__ void mp_landmark_post_kernel(const float* __restrict__ xy, const int* __restrict__ detail, const int* __restrict__ count,
                                        int K, int P, float* __restrict__ kps, int n) {
    // per stream exactly landmark_post_kernel: element i of stream st
    const int per = K * P;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= per * n) return;
    const int st = g / per, i = g - st * per, f = i / P;
    const float* x = xy + (size_t)2 * per * st;
    const int* dt = detail + (size_t)5 * K * st;
    float ox = 0.f, oy = 0.f;
    if (f < count[st]) {
        const int* d = dt + f * 5;          // [h, w, y1, x1, add]
        float px = x[i * 2] * (float)d[1];
        float py = x[i * 2 + 1] * (float)d[0];
        ox = (float)((double)px + (double)d[3] - (double)d[4]);
        oy = (float)((double)py + (double)d[2] - (double)d[4]);
    }
    kps[(size_t)2 * per * st + i * 2] = ox;
    kps[(size_t)2 * per * st + i * 2 + 1] = oy;
}

__global__ void __launch_bounds__(256) mp_absdiff_kernel(const MpStreamDesc* __restrict__ d, unsigned long long* __restrict__ sum) {
    // per stream exactly absdiff_sum_kernel (integer sums: the order of the atomic adds does not matter)
    const MpStreamDesc D = d[blockIdx.y];
    if (!D.have_prev) return;
    const uint8_t* a = D.prev;
    const uint8_t* b = D.cur;
    const size_t n = (size_t)D.H * D.W * 3, nv = n / 16;
    unsigned long long local = 0;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    const uint4* b4 = reinterpret_cast<const uint4*>(b);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x) {
        uint4 x = a4[i], y = b4[i];
        local += __vsadu4(x.x, y.x) + __vsadu4(x.y, y.y) + __vsadu4(x.z, y.z) + __vsadu4(x.w, y.w);
    }
    if (blockIdx.x == 0) {
        for (size_t i = nv * 16 + threadIdx.x; i < n; i += blockDim.x) {
            int dd = (int)a[i] - (int)b[i];
            local += (unsigned)(dd < 0 ? -dd : dd);
        }
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ unsigned long long warp_sum[8];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += warp_sum[w];
        atomicAdd(sum + blockIdx.y, t);
    }
}

int launch_mp_absdiff(const MpStreamDesc* d, unsigned long long* diff, int n, size_t max_bytes, cudaStream_t s) {
    size_t blocks = (max_bytes / 16 + 255) / 256;
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    mp_absdiff_kernel<<<dim3((unsigned)blocks, n), 256, 0, s>>>(d, diff);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}
int launch_mp_letterbox(const MpStreamDesc* d, uint8_t* out, size_t out_stride, int in_h, int in_w, int n, cudaStream_t s) {
    mp_letterbox_kernel<<<dim3((in_w + 255) / 256, in_h, n), 256, 0, s>>>(d, out, out_stride, in_h, in_w);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}
int launch_mp_detect_post(const MpStreamDesc* d, const float* raw, int rows, float score_thres, float iou_thres, float* kept_rows,
                          int* kept_idx, int* count, int max_det, int n, cudaStream_t s) {
    mp_detect_post_kernel<<<n, 1024, 0, s>>>(d, raw, rows, score_thres, iou_thres, kept_rows, kept_idx, count, max_det);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}
int launch_mp_crop(const MpStreamDesc* d, const float* boxes, const int* count, int K, float face_scale, float min_face,
                   uint8_t* crops, int S, int* detail, int n, cudaStream_t s) {
    mp_crop_kernel<<<dim3((S + 255) / 256, S, K * n), 256, 0, s>>>(d, boxes, count, K, face_scale, min_face, crops, S, detail);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}
int launch_mp_landmark_post(const float* xy, const int* detail, const int* count, int K, int P, float* kps, int n, cudaStream_t s) {
    const int total = K * P * n;
    mp_landmark_post_kernel<<<(total + 255) / 256, 256, 0, s>>>(xy, detail, count, K, P, kps, n);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

int launch_mp_select(const float* det_rows, const int* det_count, int max_det, const int* flag, const float* track,
                     const int* n_track, float iou_thres, float alpha, float oma, float min_face, int top_k, float* boxes4,
                     int* count, int n_streams, cudaStream_t s) {
    mp_select_kernel<<<n_streams, 256, 0, s>>>(det_rows, det_count, max_det, flag, track, n_track, iou_thres, alpha, oma,
                                                min_face, top_k, boxes4, count);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace skps

// ============================================================================================
// C-ABI wrappers
// ============================================================================================
using namespace skps;

extern "C" SKPS_API int skps_letterbox(const uint8_t* frame, int H, int W, int pitch, uint8_t* out, int in_h, int in_w,
                              int rw, int rh, int top, int left, void* stream) {
    SKPS_CHECK(frame && out && H > 0 && W > 0 && rw > 0 && rh > 0, "letterbox: bad arguments");
    dim3 grid((in_w + 255) / 256, in_h);
    letterbox_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frame, H, W, pitch, out, in_h, in_w, rw, rh, top, left);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_detect_post(const float* raw, int rows, float score_thres, float iou_thres, float scale,
                                float pad_x, float pad_y, float* kept_rows, int32_t* kept_idx, int32_t* count,
                                int max_det, void* stream) {
    SKPS_CHECK(raw && kept_rows && kept_idx && count && max_det > 0 && max_det <= 256, "detect_post: bad arguments");
    detect_post_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(raw, rows, score_thres, iou_thres, scale, pad_x, pad_y,
                                                            kept_rows, kept_idx, count, max_det);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_select_faces(const float* det_rows, const int32_t* det_count, int det_stride, const float* track,
                                 int n_track, float iou_thres, float alpha, float one_minus_alpha, float min_face,
                                 int top_k, float* boxes4, int32_t* count, void* stream) {
    SKPS_CHECK(det_rows && det_count && boxes4 && count && top_k > 0 && top_k <= 64, "select_faces: bad arguments");
    select_faces_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(det_rows, det_count, det_stride, track,
                                                            track ? n_track : 0, iou_thres, alpha, one_minus_alpha,
                                                            min_face, top_k, boxes4, count);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_crop_resize(const uint8_t* frame, int H, int W, int pitch, const float* boxes4,
                                const int32_t* count, int max_faces, float face_scale, float min_face,
                                uint8_t* crops, int out_hw, int32_t* detail, void* stream) {
    SKPS_CHECK(frame && boxes4 && count && crops && detail && max_faces > 0, "crop_resize: bad arguments");
    dim3 grid((out_hw + 255) / 256, out_hw, max_faces);
    crop_resize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frame, H, W, pitch, boxes4, count, face_scale,
                                                               min_face, crops, out_hw, detail);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_crop_rect(const uint8_t* frame, int H, int W, int pitch, int rx, int ry, int rw, int rh,
                                       uint8_t* out, int out_hw, void* stream) {
    SKPS_CHECK(frame && out && H > 0 && W > 0 && rw > 0 && rh > 0 && out_hw > 0, "crop_rect: bad arguments");
    dim3 grid((out_hw + 255) / 256, out_hw);
    rect_resize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frame, H, W, pitch, rx, ry, rw, rh, out, out_hw);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_nme(const float* target, const float* preds, int n, int n_points, int norm_a, int norm_b,
                                 float* out, void* stream) {
    SKPS_CHECK(target && preds && out && n > 0 && n_points > 0 && norm_a >= 0 && norm_b >= 0 && norm_a < n_points &&
               norm_b < n_points, "nme: bad arguments");
    nme_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(target, preds, n, n_points, norm_a, norm_b, out);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_landmark_post(const float* xy_norm, const int32_t* detail, const int32_t* count, int max_faces,
                                  int n_points, float* kps, void* stream) {
    SKPS_CHECK(xy_norm && detail && count && kps, "landmark_post: bad arguments");
    int total = max_faces * n_points;
    landmark_post_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(xy_norm, detail, count, max_faces,
                                                                               n_points, kps);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

extern "C" SKPS_API int skps_frame_absdiff_sum(const uint8_t* a, const uint8_t* b, size_t n, unsigned long long* sum,
                                      void* stream) {
    SKPS_CHECK(a && b && sum, "absdiff: bad arguments");
    SKPS_CHECK(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0), "absdiff: pointers must be 16-byte aligned");
    SKPS_CUDA(cudaMemsetAsync(sum, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    int blocks = 148 * 8;
    absdiff_sum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n, sum);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}
