// Non-GEMM layers of the two networks (sm_100a): depthwise conv, pooling, resizes, channel-view
// copy, global average pool, BatchNorm affine, scSE apply, and the two decode tails.  All are
// HBM/L2-bound gathers; they read/write NHWC float32 through channel views so concat, slice and
// channel shuffle never cost a pass of their own.
#include "common.h"

namespace skps {

static inline int blocks_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    if (b > (1LL << 30)) b = 1LL << 30;
    return (int)b;
}

// ------------------------------------------------------------------------------------------
// Depthwise kxk conv + bias + activation.  One thread = one output pixel x 4 channels.
// (kps_student.onnx conv_dw nodes, e.g. node 4; yolov5n-0.5.onnx branch*.0/.3 nodes.)
// ------------------------------------------------------------------------------------------
struct DwK {
    const void* in; int in_ld, in_coff, H, W; int in_fmt; long long in_plane;
    void* out; int out_ld, out_coff, out_cstride, Ho, Wo, C; int out_fmt; long long out_plane;
    const float* w; const float* bias;
    int kh, kw, sh, sw, ph, pw, dh, dw, act;
    long long total;   // batch*Ho*Wo*(C/4)
};

__global__ void __launch_bounds__(256) dwconv_kernel(const DwK p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.total) return;
    const int C4 = p.C >> 2;
    int c = (int)(i % C4) * 4;
    long long pix = i / C4;
    int ox = (int)(pix % p.Wo);
    long long t = pix / p.Wo;
    int oy = (int)(t % p.Ho);
    int n = (int)(t / p.Ho);
    float4 acc = *reinterpret_cast<const float4*>(p.bias + c);
    const int iy0 = oy * p.sh - p.ph, ix0 = ox * p.sw - p.pw;
    for (int ky = 0; ky < p.kh; ++ky) {
        int iy = iy0 + ky * p.dh;
        if (iy < 0 || iy >= p.H) continue;
        for (int kx = 0; kx < p.kw; ++kx) {
            int ix = ix0 + kx * p.dw;
            if (ix < 0 || ix >= p.W) continue;
            const float4 v = ld4(p.in, p.in_fmt, p.in_plane,
                                  (((long long)n * p.H + iy) * p.W + ix) * p.in_ld + p.in_coff + c);
            const float4 w = *reinterpret_cast<const float4*>(p.w + (ky * p.kw + kx) * p.C + c);
            acc.x = fmaf(v.x, w.x, acc.x);
            acc.y = fmaf(v.y, w.y, acc.y);
            acc.z = fmaf(v.z, w.z, acc.z);
            acc.w = fmaf(v.w, w.w, acc.w);
        }
    }
    acc.x = apply_act(acc.x, p.act);
    acc.y = apply_act(acc.y, p.act);
    acc.z = apply_act(acc.z, p.act);
    acc.w = apply_act(acc.w, p.act);
    const long long o = pix * p.out_ld + p.out_coff;
    if (p.out_cstride == 1) {
        st4(p.out, p.out_fmt, p.out_plane, o + c, acc);
    } else {
        st1(p.out, p.out_fmt, p.out_plane, o + (long long)(c + 0) * p.out_cstride, acc.x);
        st1(p.out, p.out_fmt, p.out_plane, o + (long long)(c + 1) * p.out_cstride, acc.y);
        st1(p.out, p.out_fmt, p.out_plane, o + (long long)(c + 2) * p.out_cstride, acc.z);
        st1(p.out, p.out_fmt, p.out_plane, o + (long long)(c + 3) * p.out_cstride, acc.w);
    }
}

int launch_dwconv(const DwArgs& a, cudaStream_t s) {
    DwK k;
    k.in = a.in.base; k.in_ld = a.in.ld; k.in_coff = a.in.c_off; k.H = a.in.H; k.W = a.in.W;
    k.in_fmt = a.in.fmt; k.in_plane = a.in.plane;
    k.out = a.out.base; k.out_ld = a.out.ld; k.out_coff = a.out.c_off; k.out_cstride = a.out.c_stride;
    k.Ho = a.out.H; k.Wo = a.out.W; k.C = a.out.C; k.out_fmt = a.out.fmt; k.out_plane = a.out.plane;
    k.w = a.w; k.bias = a.bias;
    k.kh = a.kh; k.kw = a.kw; k.sh = a.sh; k.sw = a.sw; k.ph = a.ph; k.pw = a.pw; k.dh = a.dh; k.dw = a.dw;
    k.act = a.act;
    SKPS_CHECK(a.in.c_stride == 1 && a.in.C == a.out.C, "dwconv: bad input view");
    SKPS_CHECK(k.C % 4 == 0 && k.in_ld % 4 == 0 && k.in_coff % 4 == 0, "dwconv: C/ld/offset must be multiples of 4");
    SKPS_CHECK(k.out_cstride != 1 || (k.out_ld % 4 == 0 && k.out_coff % 4 == 0), "dwconv: unaligned output view");
    k.total = (long long)a.batch * k.Ho * k.Wo * (k.C / 4);
    dwconv_kernel<<<blocks_for(k.total, 256), 256, 0, s>>>(k);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Generic per-element kernels over (n, y, x, c) of the OUTPUT view.
// ------------------------------------------------------------------------------------------
struct EwK {
    const void* in; int in_ld, in_coff, in_cs, H, W; int in_fmt; long long in_plane;
    void* out; int out_ld, out_coff, out_cs, Ho, Wo, C; int out_fmt; long long out_plane;
    const void* a; const void* b;     // op specific
    int a_ld, a_coff, b_ld, b_coff; int a_fmt, b_fmt; long long a_plane, b_plane;
    int act;
    long long total;
};

#define SRC(off) ld1(p.in, p.in_fmt, p.in_plane, sbase + (off))
template <int MODE>   // 0 maxpool2 ceil, 1 nearest, 2 bilinear2x, 3 copy, 4 affine+act, 5 scse, 6 scale by gate[n,c]
__global__ void __launch_bounds__(256) elementwise_kernel(const EwK p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.total) return;
    int c = (int)(i % p.C);
    long long pix = i / p.C;
    int ox = (int)(pix % p.Wo);
    long long t = pix / p.Wo;
    int oy = (int)(t % p.Ho);
    int n = (int)(t / p.Ho);
    const long long sbase = (long long)n * p.H * p.W * p.in_ld + p.in_coff + (long long)c * p.in_cs;
    float v;
    if (MODE == 0) {
        // MaxPool 2x2 stride 2, ceil_mode=1, no padding (yolov5n-0.5.onnx node 9)
        int y0 = oy * 2, x0 = ox * 2;
        v = SRC(((long long)y0 * p.W + x0) * p.in_ld);
        if (x0 + 1 < p.W) v = fmaxf(v, SRC(((long long)y0 * p.W + x0 + 1) * p.in_ld));
        if (y0 + 1 < p.H) {
            v = fmaxf(v, SRC(((long long)(y0 + 1) * p.W + x0) * p.in_ld));
            if (x0 + 1 < p.W) v = fmaxf(v, SRC(((long long)(y0 + 1) * p.W + x0 + 1) * p.in_ld));
        }
    } else if (MODE == 1) {
        // Resize nearest, asymmetric, floor: src = floor(dst * in / out)
        int iy = (int)(((long long)oy * p.H) / p.Ho), ix = (int)(((long long)ox * p.W) / p.Wo);
        v = SRC(((long long)iy * p.W + ix) * p.in_ld);
    } else if (MODE == 2) {
        // Resize linear, half_pixel, scale 2 (kps_student.onnx nodes 176, 193)
        float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.f);
        int y0 = (int)sy, x0 = (int)sx;
        int y1 = min(y0 + 1, p.H - 1), x1 = min(x0 + 1, p.W - 1);
        float ly = sy - y0, lx = sx - x0;
        float hy = 1.f - ly, hx = 1.f - lx;
        float p00 = SRC(((long long)y0 * p.W + x0) * p.in_ld), p01 = SRC(((long long)y0 * p.W + x1) * p.in_ld);
        float p10 = SRC(((long long)y1 * p.W + x0) * p.in_ld), p11 = SRC(((long long)y1 * p.W + x1) * p.in_ld);
        v = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
    } else {
        v = SRC(((long long)oy * p.W + ox) * p.in_ld);
        if (MODE == 4) {
            v = apply_act(fmaf(v, ((const float*)p.a)[c], ((const float*)p.b)[c]), p.act);
        } else if (MODE == 5) {
            // scSE (model.py:117-130): x*cSE[n,c] + x*sSE[n,h,w]
            float cs = ld1(p.a, p.a_fmt, p.a_plane, (long long)n * p.a_ld + p.a_coff + c);
            float ss = ld1(p.b, p.b_fmt, p.b_plane,
                           ((long long)n * p.H * p.W + (long long)oy * p.W + ox) * p.b_ld + p.b_coff);
            v = __fadd_rn(__fmul_rn(v, cs), __fmul_rn(v, ss));
        } else if (MODE == 6) {
            // squeeze-excite gate applied ahead of a tensor-core conv: x * gate[n,c]
            v = v * ld1(p.a, p.a_fmt, p.a_plane, (long long)n * p.a_ld + p.a_coff + c);
        }
    }
    st1(p.out, p.out_fmt, p.out_plane, pix * p.out_ld + p.out_coff + (long long)c * p.out_cs, v);
}
#undef SRC

static EwK make_ew(const TView& in, const TView& out, int batch) {
    EwK k = {};
    k.in = in.base; k.in_ld = in.ld; k.in_coff = in.c_off; k.in_cs = in.c_stride; k.H = in.H; k.W = in.W;
    k.in_fmt = in.fmt; k.in_plane = in.plane;
    k.out = out.base; k.out_ld = out.ld; k.out_coff = out.c_off; k.out_cs = out.c_stride;
    k.Ho = out.H; k.Wo = out.W; k.C = out.C; k.out_fmt = out.fmt; k.out_plane = out.plane;
    k.total = (long long)batch * out.H * out.W * out.C;
    return k;
}

#define LAUNCH_EW(MODE, k, s)                                                     \
    elementwise_kernel<MODE><<<blocks_for((k).total, 256), 256, 0, s>>>(k);       \
    SKPS_CUDA(cudaGetLastError());                                                \
    return 0;

int launch_maxpool2(const TView& in, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(in.C == out.C && out.H == (in.H + 1) / 2 && out.W == (in.W + 1) / 2, "maxpool: shape");
    EwK k = make_ew(in, out, batch);
    LAUNCH_EW(0, k, s)
}
int launch_resize_nearest(const TView& in, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(in.C == out.C, "resize: channels");
    EwK k = make_ew(in, out, batch);
    LAUNCH_EW(1, k, s)
}
int launch_bilinear2x(const TView& in, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(in.C == out.C && out.H == 2 * in.H && out.W == 2 * in.W, "bilinear2x: shape");
    EwK k = make_ew(in, out, batch);
    LAUNCH_EW(2, k, s)
}
int launch_copy(const TView& in, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(in.C == out.C && in.H == out.H && in.W == out.W, "copy: shape");
    EwK k = make_ew(in, out, batch);
    LAUNCH_EW(3, k, s)
}
int launch_affine_act(const TView& in, const TView& out, const float* sc, const float* sh, int act, int batch,
                      cudaStream_t s) {
    SKPS_CHECK(in.C == out.C && in.H == out.H && in.W == out.W, "affine: shape");
    EwK k = make_ew(in, out, batch);
    k.a = sc; k.b = sh; k.act = act;
    LAUNCH_EW(4, k, s)
}
int launch_scse(const TView& x, const TView& cse, const TView& sse, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(x.C == out.C && cse.C == x.C && sse.C == 1 && sse.H == x.H && sse.W == x.W, "scse: shape");
    EwK k = make_ew(x, out, batch);
    k.a = cse.base; k.a_ld = cse.ld; k.a_coff = cse.c_off; k.a_fmt = cse.fmt; k.a_plane = cse.plane;
    k.b = sse.base; k.b_ld = sse.ld; k.b_coff = sse.c_off; k.b_fmt = sse.fmt; k.b_plane = sse.plane;
    LAUNCH_EW(5, k, s)
}
int launch_scale_ch(const TView& x, const TView& gate, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(x.C == out.C && gate.C == x.C && x.H == out.H && x.W == out.W, "scale_ch: shape");
    EwK k = make_ew(x, out, batch);
    k.a = gate.base; k.a_ld = gate.ld; k.a_coff = gate.c_off; k.a_fmt = gate.fmt; k.a_plane = gate.plane;
    LAUNCH_EW(6, k, s)
}

// ------------------------------------------------------------------------------------------
// Global average pool: (N,H,W,C) -> (N,1,1,C).  Block = 32 channels x 8 pixel lanes, fixed
// summation order (deterministic).  ReduceMean(2,3) / GlobalAveragePool nodes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gap_kernel(const void* in, int in_fmt, long long in_plane, int in_ld, int in_coff,
                                                  int HW, int C, float* out, int out_ld, int out_coff) {
    __shared__ float part[8][33];
    int n = blockIdx.y;
    int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < C) {
        const long long sb = (long long)n * HW * in_ld + in_coff + c;
        for (int p = threadIdx.y; p < HW; p += 8) s += ld1(in, in_fmt, in_plane, sb + (long long)p * in_ld);
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += part[j][threadIdx.x];
        out[(long long)n * out_ld + out_coff + c] = t / (float)HW;
    }
}

int launch_gap(const TView& in, const TView& out, int batch, cudaStream_t s) {
    SKPS_CHECK(in.C == out.C && out.H == 1 && out.W == 1 && in.c_stride == 1 && out.c_stride == 1, "gap: shape");
    dim3 grid((in.C + 31) / 32, batch), block(32, 8);
    SKPS_CHECK(out.fmt == DT_F32, "gap: output must be float32");
    gap_kernel<<<grid, block, 0, s>>>(in.base, in.fmt, in.plane, in.ld, in.c_off, in.H * in.W, in.C,
                                      (float*)out.base, out.ld, out.c_off);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// yolov5-face Detect decode (yolov5n-0.5.onnx nodes 502-820): three head tensors (N,H,W,48) ->
// (N, rows, 16) = [cx,cy,w,h,obj, 5x(lx,ly), cls], rows ordered scale, anchor, y, x.
// consts: per scale [stride, aw0,ah0, aw1,ah1, aw2,ah2].
// ------------------------------------------------------------------------------------------
struct DetK {
    const float* head[3]; int ld[3], coff[3], H[3], W[3];
    float consts[21];
    float* out; int rows; long long total;
};

__global__ void __launch_bounds__(256) det_decode_kernel(const DetK p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.total) return;
    int row = (int)(i % p.rows);
    int n = (int)(i / p.rows);
    int si = 0, r = row;
    {
        int c0 = 3 * p.H[0] * p.W[0];
        if (r >= c0) {
            r -= c0; si = 1;
            int c1 = 3 * p.H[1] * p.W[1];
            if (r >= c1) { r -= c1; si = 2; }
        }
    }
    const int H = p.H[si], W = p.W[si];
    int x = r % W;
    int t = r / W;
    int y = t % H;
    int a = t / H;
    const float stride = p.consts[si * 7];
    const float aw = p.consts[si * 7 + 1 + 2 * a], ah = p.consts[si * 7 + 2 + 2 * a];
    const float* src = p.head[si] + (((long long)n * H + y) * W + x) * p.ld[si] + p.coff[si] + a * 16;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = src[k];
    float o[16];
    const float gx = (float)x, gy = (float)y;
    float s0 = sigmoid_f(v[0]), s1 = sigmoid_f(v[1]), s2 = sigmoid_f(v[2]), s3 = sigmoid_f(v[3]);
    o[0] = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(s0, 2.f), -0.5f), gx), stride);
    o[1] = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(s1, 2.f), -0.5f), gy), stride);
    float w2 = __fmul_rn(s2, 2.f), h2 = __fmul_rn(s3, 2.f);
    o[2] = __fmul_rn(__fmul_rn(w2, w2), aw);
    o[3] = __fmul_rn(__fmul_rn(h2, h2), ah);
    o[4] = sigmoid_f(v[4]);
    const float gxs = __fmul_rn(gx, stride), gys = __fmul_rn(gy, stride);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        o[5 + 2 * k] = __fadd_rn(__fmul_rn(v[5 + 2 * k], aw), gxs);
        o[6 + 2 * k] = __fadd_rn(__fmul_rn(v[6 + 2 * k], ah), gys);
    }
    o[15] = sigmoid_f(v[15]);
    float4* dst = reinterpret_cast<float4*>(p.out + i * 16);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    dst[2] = make_float4(o[8], o[9], o[10], o[11]);
    dst[3] = make_float4(o[12], o[13], o[14], o[15]);
}

int launch_det_decode(const TView* heads, const float* consts, const TView& out, int rows, int batch,
                      cudaStream_t s) {
    DetK k;
    int total_rows = 0;
    for (int i = 0; i < 3; ++i) {
        SKPS_CHECK(heads[i].C == 48 && heads[i].c_stride == 1, "det_decode: head view");
        k.head[i] = (const float*)heads[i].base; k.ld[i] = heads[i].ld; k.coff[i] = heads[i].c_off;
        k.H[i] = heads[i].H; k.W[i] = heads[i].W;
        total_rows += 3 * heads[i].H * heads[i].W;
    }
    SKPS_CHECK(total_rows == rows && out.ld == 16 && out.c_off == 0, "det_decode: rows %d vs %d", total_rows, rows);
    for (int i = 0; i < 21; ++i) k.consts[i] = consts[i];
    k.out = (float*)out.base; k.rows = rows; k.total = (long long)batch * rows;
    det_decode_kernel<<<blocks_for(k.total, 256), 256, 0, s>>>(k);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------
// Heat-map decode (kps_student.onnx nodes 201-410; TRAIN/.../model.py:511-554): per landmark c
// arg-max over H*W of hm[:, c] (first maximum), score = max, x = (i%W + hm[:,P+c][i]) / W,
// y = (i/W + hm[:,2P+c][i]) / W.  Block = one sample; 128 channel lanes x 8 position groups.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) hm_decode_kernel(const float* hm, int ld, int coff, int H, int W, int P,
                                                         float* xy, int xy_ld, float* score, int sc_ld) {
    __shared__ float sv[8][128];
    __shared__ int si[8][128];
    const int n = blockIdx.x;
    const int c = threadIdx.x, g = threadIdx.y;
    const int HW = H * W;
    const float* base = hm + (long long)n * HW * ld + coff;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (c < P) {
        int per = (HW + 7) / 8;
        int p0 = g * per, p1 = min(HW, p0 + per);
        for (int p = p0; p < p1; ++p) {
            float v = base[(long long)p * ld + c];
            if (bi == 0x7fffffff || v > best) { best = v; bi = p; }   // ascending p: first maximum wins
        }
    }
    sv[g][c] = best;
    si[g][c] = bi;
    __syncthreads();
    if (g == 0 && c < P) {
        for (int j = 1; j < 8; ++j) {
            float v = sv[j][c];
            int idx = si[j][c];
            if (idx != 0x7fffffff && (v > best || (v == best && idx < bi))) { best = v; bi = idx; }
        }
        float ox = base[(long long)bi * ld + P + c];
        float oy = base[(long long)bi * ld + 2 * P + c];
        float x = __fdiv_rn(__fadd_rn((float)(bi % W), ox), (float)W);
        float y = __fdiv_rn(__fadd_rn((float)(bi / W), oy), (float)W);
        xy[(long long)n * xy_ld + 2 * c] = x;
        xy[(long long)n * xy_ld + 2 * c + 1] = y;
        score[(long long)n * sc_ld + c] = best;
    }
}

int launch_hm_decode(const TView& hm, const TView& xy, const TView& score, int npts, int batch, cudaStream_t s) {
    SKPS_CHECK(hm.C == 3 * npts && npts <= 128 && hm.c_stride == 1 && hm.H == hm.W && hm.fmt == DT_F32,
               "hm_decode: shape/format");
    dim3 block(128, 8);
    hm_decode_kernel<<<batch, block, 0, s>>>((const float*)hm.base, hm.ld, hm.c_off, hm.H, hm.W, npts,
                                             (float*)xy.base, xy.ld, (float*)score.base, score.ld);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace skps
