// Generic fp32 implicit-GEMM convolution on CUDA cores (sm_100a).
//
// Used for every dense conv that is not routed to the tcgen05 path: the detector (batch 1,
// latency bound), squeeze-excite FCs, tiny-K layers.  C[M=N*Ho*Wo][Cout] = A[M][kh*kw*Cin] * W^T
// with A gathered on the fly from the NHWC activation (zero outside the image), fused
// bias + activation + residual + channel-view store (concat / channel-shuffle destinations).
//
// Replaces the Conv (+Sigmoid/Mul, +HardSigmoid/Mul, +Relu, +Add) node groups that the reference
// executes through onnxruntime (Skps/core/api/onnx_model_base.py:23).
#include "common.h"

namespace skps {

constexpr int BK = 16;

struct ConvK {
    const void* in; int in_ld, in_coff; int H, W, Cin; int in_fmt; long long in_plane;
    void* out; int out_ld, out_coff, out_cstride; int Ho, Wo, Cout; int out_fmt; long long out_plane;
    const void* res; int res_ld, res_coff; int res_fmt; long long res_plane;
    const float* gate; int gate_ld, gate_coff;     // per-sample (n, ci) input scale
    const float* w; const float* bias;
    int kh, kw, sh, sw, ph, pw, dh, dw, act;
    int res_first;                                 // act(conv + bias + res) instead of act(conv + bias) + res
    int M;                                         // batch*Ho*Wo
};

template <int BM, int BN, int TM, int TN, bool VEC, bool IN_U8>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
conv_igemm_kernel(const ConvK p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int A_SLOTS = BM * BK / 4;
    constexpr int B_SLOTS = BN * BK / 4;
    constexpr int A_IT = (A_SLOTS + NT - 1) / NT;
    constexpr int B_IT = (B_SLOTS + NT - 1) / NT;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);

    // ---- per-slot row bookkeeping for the A gather
    int a_n[A_IT], a_iy0[A_IT], a_ix0[A_IT];
    bool a_ok[A_IT];
#pragma unroll
    for (int it = 0; it < A_IT; ++it) {
        int slot = tid + it * NT;
        int r = slot / (BK / 4);
        int m = m0 + r;
        bool ok = (slot < A_SLOTS) && (m < p.M);
        int mm = ok ? m : 0;
        int ox = mm % p.Wo;
        int t = mm / p.Wo;
        int oy = t % p.Ho;
        a_n[it] = t / p.Ho;
        a_iy0[it] = oy * p.sh - p.ph;
        a_ix0[it] = ox * p.sw - p.pw;
        a_ok[it] = ok;
    }

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int cchunks = (p.Cin + BK - 1) / BK;
    const int total = p.kh * p.kw * cchunks;
    const int Kw = p.kh * p.kw * p.Cin;            // weight row length

    float4 a_reg[A_IT], b_reg[B_IT];

    auto prefetch = [&](int chunk) {
        int tap = chunk / cchunks;
        int ci0 = (chunk - tap * cchunks) * BK;
        int ky = tap / p.kw, kx = tap - ky * p.kw;
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            int slot = tid + it * NT;
            int kq = slot % (BK / 4);
            int ci = ci0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int iy = a_iy0[it] + ky * p.dh, ix = a_ix0[it] + kx * p.dw;
            if (a_ok[it] && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && ci < p.Cin) {
                long long pix = ((long long)a_n[it] * p.H + iy) * p.W + ix;
                if (IN_U8) {
                    const uint8_t* src = (const uint8_t*)p.in + pix * p.in_ld + p.in_coff + ci;
                    // uint8 -> float32 then true division by 255 (face_detector.py:67, face_landmark.py:46)
                    v.x = __fdiv_rn((float)src[0], 255.f);
                    if (ci + 1 < p.Cin) v.y = __fdiv_rn((float)src[1], 255.f);
                    if (ci + 2 < p.Cin) v.z = __fdiv_rn((float)src[2], 255.f);
                    if (ci + 3 < p.Cin) v.w = __fdiv_rn((float)src[3], 255.f);
                } else {
                    const long long e = pix * p.in_ld + p.in_coff + ci;
                    if (VEC) {
                        v = ld4(p.in, p.in_fmt, p.in_plane, e);
                    } else {
                        v.x = ld1(p.in, p.in_fmt, p.in_plane, e);
                        if (ci + 1 < p.Cin) v.y = ld1(p.in, p.in_fmt, p.in_plane, e + 1);
                        if (ci + 2 < p.Cin) v.z = ld1(p.in, p.in_fmt, p.in_plane, e + 2);
                        if (ci + 3 < p.Cin) v.w = ld1(p.in, p.in_fmt, p.in_plane, e + 3);
                    }
                    if (p.gate) {
                        const float* g = p.gate + (long long)a_n[it] * p.gate_ld + p.gate_coff + ci;
                        v.x *= g[0];
                        if (ci + 1 < p.Cin) v.y *= g[1];
                        if (ci + 2 < p.Cin) v.z *= g[2];
                        if (ci + 3 < p.Cin) v.w *= g[3];
                    }
                }
            }
            a_reg[it] = v;
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            int slot = tid + it * NT;
            int co = n0 + slot / (BK / 4);
            int kq = slot % (BK / 4);
            int ci = ci0 + kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (slot < B_SLOTS && co < p.Cout && ci < p.Cin) {
                const float* src = p.w + (long long)co * Kw + tap * p.Cin + ci;
                if (VEC) {
                    v = *reinterpret_cast<const float4*>(src);
                } else {
                    v.x = src[0];
                    if (ci + 1 < p.Cin) v.y = src[1];
                    if (ci + 2 < p.Cin) v.z = src[2];
                    if (ci + 3 < p.Cin) v.w = src[3];
                }
            }
            b_reg[it] = v;
        }
    };

    prefetch(0);
    for (int chunk = 0; chunk < total; ++chunk) {
#pragma unroll
        for (int it = 0; it < A_IT; ++it) {
            int slot = tid + it * NT;
            if (slot < A_SLOTS) {
                int r = slot / (BK / 4), kq = slot % (BK / 4);
                As[kq * 4 + 0][r] = a_reg[it].x;
                As[kq * 4 + 1][r] = a_reg[it].y;
                As[kq * 4 + 2][r] = a_reg[it].z;
                As[kq * 4 + 3][r] = a_reg[it].w;
            }
        }
#pragma unroll
        for (int it = 0; it < B_IT; ++it) {
            int slot = tid + it * NT;
            if (slot < B_SLOTS) {
                int c = slot / (BK / 4), kq = slot % (BK / 4);
                Bs[kq * 4 + 0][c] = b_reg[it].x;
                Bs[kq * 4 + 1][c] = b_reg[it].y;
                Bs[kq * 4 + 2][c] = b_reg[it].z;
                Bs[kq * 4 + 3][c] = b_reg[it].w;
            }
        }
        __syncthreads();
        if (chunk + 1 < total) prefetch(chunk + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, activation, residual, channel-view store
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        const long long obase = (long long)m * p.out_ld + p.out_coff;
        const long long rbase = (long long)m * p.res_ld + p.res_coff;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int co = n0 + tx * TN + j;
            if (co >= p.Cout) continue;
            float v = acc[i][j];
            if (p.bias) v += p.bias[co];
            const float r = p.res ? ld1(p.res, p.res_fmt, p.res_plane, rbase + co) : 0.f;
            v = p.res_first ? apply_act(v + r, p.act) : apply_act(v, p.act) + r;
            st1(p.out, p.out_fmt, p.out_plane, obase + (long long)co * p.out_cstride, v);
        }
    }
}

template <int BM, int BN, int TM, int TN>
static int launch_cfg(const ConvK& k, bool vec, bool in_u8, cudaStream_t s) {
    dim3 grid((k.M + BM - 1) / BM, (k.Cout + BN - 1) / BN);
    constexpr int NT = (BM / TM) * (BN / TN);
    if (in_u8) conv_igemm_kernel<BM, BN, TM, TN, false, true><<<grid, NT, 0, s>>>(k);
    else if (vec) conv_igemm_kernel<BM, BN, TM, TN, true, false><<<grid, NT, 0, s>>>(k);
    else conv_igemm_kernel<BM, BN, TM, TN, false, false><<<grid, NT, 0, s>>>(k);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

int launch_conv(const ConvArgs& a, cudaStream_t s) {
    if (stem_conv_supported(a)) return launch_stem_conv(a, s);
    if (pw_small_supported(a)) return launch_pw_small(a, s);
    ConvK k;
    k.in = a.in.base; k.in_ld = a.in.ld; k.in_coff = a.in.c_off; k.H = a.in.H; k.W = a.in.W; k.Cin = a.in.C;
    k.in_fmt = a.in.fmt; k.in_plane = a.in.plane;
    k.out = a.out.base; k.out_ld = a.out.ld; k.out_coff = a.out.c_off; k.out_cstride = a.out.c_stride;
    k.Ho = a.out.H; k.Wo = a.out.W; k.Cout = a.out.C; k.out_fmt = a.out.fmt; k.out_plane = a.out.plane;
    k.res = a.res.base; k.res_ld = a.res.ld; k.res_coff = a.res.c_off; k.res_fmt = a.res.fmt; k.res_plane = a.res.plane;
    k.gate = (const float*)a.gate.base; k.gate_ld = a.gate.ld; k.gate_coff = a.gate.c_off;
    k.w = a.w; k.bias = a.bias;
    k.kh = a.kh; k.kw = a.kw; k.sh = a.sh; k.sw = a.sw; k.ph = a.ph; k.pw = a.pw; k.dh = a.dh; k.dw = a.dw;
    k.act = a.act;
    k.res_first = a.res.base ? a.res_first : 0;
    k.M = a.batch * k.Ho * k.Wo;
    SKPS_CHECK(a.in.c_stride == 1, "conv: strided input view");
    SKPS_CHECK(!a.res.base || a.res.c_stride == 1, "conv: strided residual view");
    SKPS_CHECK(!a.gate.base || (a.gate.c_stride == 1 && a.gate.C == k.Cin), "conv: bad gate view");
    bool vec = !a.in_u8 && (k.Cin % 4 == 0) && (k.in_ld % 4 == 0) && (k.in_coff % 4 == 0);
    long long tilesL = (long long)((k.M + 127) / 128) * ((k.Cout + 63) / 64);
    long long tilesM = (long long)((k.M + 63) / 64) * ((k.Cout + 63) / 64);
    if (k.Cout <= 32 && (k.M + 127) / 128 >= 148) return launch_cfg<128, 32, 8, 2>(k, vec, a.in_u8, s);
    if (tilesL >= 148) return launch_cfg<128, 64, 8, 4>(k, vec, a.in_u8, s);
    if (tilesM >= 96 && k.Cout > 32) return launch_cfg<64, 64, 4, 4>(k, vec, a.in_u8, s);
    return launch_cfg<32, 32, 2, 2>(k, vec, a.in_u8, s);
}

}  // namespace skps
