// The full-resolution head of the landmark encoder as ONE kernel (sm_100a; FP32 pipes + one tcgen05 K-step for the expansion):
//
//   uint8 crop (H x W x 3)  -> conv_stem 3x3 s2 (+/255, h-swish)                        16 ch @ H/2      [kps_student.onnx
//                           -> blocks.0.0: depthwise 3x3 + ReLU -> 1x1 16->16 + shortcut 16 ch @ H/2       /student/encoder/
//                           -> blocks.1.0: 1x1 16->E + ReLU -> depthwise 3x3 s2 + ReLU   E ch @ H/4        conv_stem .. blocks.1.0/conv_dw]
//
// (timm mobilenetv3 stem + DepthwiseSeparable block + the expand/depthwise half of the first InvertedResidual;
// TRAIN/face_landmark/lib/core/base_trainer/model.py:247-262.)  Run as separate launches these four layers move 3.5 GB
// of 16- and 64-channel full-resolution tensors through HBM per 256-face batch (1.17 ms of a 7 ms step, r2 launch list);
// fused, a CTA reads a 39 x 71 pixel window of the crop and writes an 8 x 16 tile of the E-channel quarter-resolution
// tensor, everything in between lives in shared memory.  The work is ~1.2 M FMAs per tile on tensors with 3 / 16 input
// channels.  The stem, the block-0 depthwise and its 16->16 pointwise run on the FP32 pipes with every dense weight taken from
// the kernel-parameter (constant) bank (FFMA with a constant operand, each activation read from shared memory once per 16
// outputs); the 16->E expansion - half of the FMAs - is a single tcgen05 K-step per 128 window pixels (TC variant below,
// the default; SKPS_STEM_TC=0 keeps it on the FP32 pipes).  Measured: 1.11 ms all-FP32 -> 1.01 ms with the expansion on
// tensor cores -> 0.97 ms with the next tile's input window prefetched into registers (per 256-face batch).
#include <cuda_fp16.h>
#include <string.h>

#include <stdlib.h>

#include "common.h"
#include "stem_block.h"
#include "tc_ptx.h"

namespace skps {

constexpr int SB_THREADS = 576;
constexpr int SB_TH = 8, SB_TW = 16;                       // output tile (quarter resolution)
constexpr int SB_EH = 2 * SB_TH + 1, SB_EW = 2 * SB_TW + 1;   // 17 x 33: half-resolution window the stride-2 depthwise reads
constexpr int SB_SH = SB_EH + 2, SB_SW = SB_EW + 2;        // 19 x 35: stem outputs the 3x3 depthwise of block 0 reads
constexpr int SB_IH = 2 * SB_SH + 1, SB_IW = 2 * SB_SW + 1;   // 39 x 71: input pixels the stem reads
constexpr int SB_PS = 20;                                  // floats per pixel in shared memory (16 + pad: conflict-free float4 rows)
constexpr int SB_NE = SB_EH * SB_EW, SB_NS = SB_SH * SB_SW, SB_NI = SB_IH * SB_IW;
constexpr int SB_A_FLOATS = (SB_NI * 4 > SB_NE * SB_PS) ? SB_NI * 4 : SB_NE * SB_PS;    // input window, later block-0 depthwise output
constexpr int SB_B_FLOATS = SB_NS * SB_PS;                 // stem output, later one 16-channel chunk of the expanded tensor
constexpr int SB_C_FLOATS = SB_NE * SB_PS;                 // block-0 output
constexpr int SB_SMEM = (SB_A_FLOATS + SB_B_FLOATS + SB_C_FLOATS + 10 * SB_MAX_E + 256) * 4;
// Tensor-core variant (TC = true): the 16 -> E expansion (half of the block's FMAs) is ONE tcgen05 K-step.  S3 writes the
// block-0 output as float16 hi/lo rows (64-byte swizzled rows of which only the first 32 bytes = 16 channels are used) instead
// of float32, five 128-row MMA tiles cover the 561 window pixels, the E-column accumulators live in TMEM and S4 shrinks to
// tcgen05.ld -> bias/ReLU/mask -> shared memory.  Same fp16 hi/lo three-product scheme as conv_tc.cu.
constexpr int SB_T_TILES = (SB_NE + 127) / 128;            // 5
constexpr int SB_T_PLANE = SB_T_TILES * 128 * 64;          // one plane of the A operand: 640 rows x 64 B
constexpr int SB_WB_PLANE = SB_MAX_E * 64;                 // one plane of the B operand: E rows x 64 B
constexpr int SB_SMEM_TC = (SB_A_FLOATS + SB_B_FLOATS + 10 * SB_MAX_E + 256) * 4 + 1024 + 2 * SB_T_PLANE + 2 * SB_WB_PLANE;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 floats -> 16 bytes of float16 hi and 16 bytes of float16 lo (v = hi + lo)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t* hp = reinterpret_cast<uint32_t*>(&hi);
    uint32_t* lp = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 hf = __half22float2(h2);
        const __half2 l2 = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
        hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
        lp[j] = *reinterpret_cast<const uint32_t*>(&l2);
    }
}

__device__ __forceinline__ float hswish_f(float v) { return v * hsigmoid_f(v); }

// block-0 depthwise: 4 consecutive pixels of one window row x channel group G (weights as constant operands)
template <int G>
__device__ __forceinline__ void dw16_strip(const float* __restrict__ s1, float* __restrict__ s2, int row, int x0, const StemBlockW& Wt) {
    float4 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float4(Wt.dw0_b[4 * G], Wt.dw0_b[4 * G + 1], Wt.dw0_b[4 * G + 2], Wt.dw0_b[4 * G + 3]);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        float4 in[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int xx = min(x0 + i, SB_SW - 1);          // the last strip of a row hangs over by up to 3 columns
            in[i] = *reinterpret_cast<const float4*>(s1 + ((row + ky) * SB_SW + xx) * SB_PS + 4 * G);
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int t = ky * 3 + kx;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[q].x = fmaf(in[q + kx].x, Wt.dw0_w[t * 16 + 4 * G], acc[q].x);
                acc[q].y = fmaf(in[q + kx].y, Wt.dw0_w[t * 16 + 4 * G + 1], acc[q].y);
                acc[q].z = fmaf(in[q + kx].z, Wt.dw0_w[t * 16 + 4 * G + 2], acc[q].z);
                acc[q].w = fmaf(in[q + kx].w, Wt.dw0_w[t * 16 + 4 * G + 3], acc[q].w);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (x0 + q >= SB_EW) break;
        float4 v = acc[q];
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        *reinterpret_cast<float4*>(s2 + (row * SB_EW + x0 + q) * SB_PS + 4 * G) = v;
    }
}

// one 16-output slice of a 1x1 conv on a 16-channel pixel: acc[j] = b[j] + sum_ci x[ci] * w[ci][j] (constant operands)
template <int CO, int OFF>
__device__ __forceinline__ void pw16in(const float* __restrict__ x, const float* w, const float* b, float* acc) {
    float in[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(x + 4 * g);
        in[4 * g] = v.x; in[4 * g + 1] = v.y; in[4 * g + 2] = v.z; in[4 * g + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = b[OFF + j];
#pragma unroll
    for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(in[ci], w[ci * CO + OFF + j], acc[j]);
}

template <int E, bool TC>
__global__ void __launch_bounds__(SB_THREADS, 1)
stem_block_kernel(const StemBlockK p, const __grid_constant__ StemBlockW Wt) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_wmax[SB_THREADS / 32];
    float* sA = sm;                                    // input window (float4 per pixel: b, g, r, 0), later s2
    float* sB = sA + SB_A_FLOATS;                      // stem output s1, later the expanded chunk
    float* sC = sB + SB_B_FLOATS;                      // block-0 output s3 (float32 variant only)
    float* sW = TC ? sC : sC + SB_C_FLOATS;            // stride-2 depthwise weights [9][E] + bias [E]
    float* lut = sW + 10 * SB_MAX_E;                   // i / 255
    // TC: A operand (block-0 output as fp16 hi/lo rows), then the B operand (expand weights), 1024-byte aligned
    const uint32_t sT = (smem_u32(lut + 256) + 1023u) & ~1023u;
    const uint32_t sWB = sT + 2u * SB_T_PLANE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += SB_THREADS) lut[i] = __fdiv_rn((float)i, 255.f);
    for (int i = tid; i < 10 * E; i += SB_THREADS) sW[i] = p.dw1[i];
    const int tiles_x = p.Wq / SB_TW, tiles_per_img = tiles_x * (p.Hq / SB_TH);
    const int Hh = p.H / 2, Wh = p.W / 2;              // half-resolution map
    uint32_t tmem_base = 0, mma_phase = 0;
    float w_inv = 1.f;
    if (TC) {
        // expand weights -> fp16 hi/lo B operand [E rows][16 K], pre-multiplied by an exact power of two (undone after the
        // MMA) so that the lo parts stay in float16's normal range, as plan.pack_tc_weights does for the other layers
        float wm = 0.f;
        for (int i = tid; i < 16 * E; i += SB_THREADS) wm = fmaxf(wm, fabsf(Wt.pw1_w[i]));
#pragma unroll
        for (int o = 16; o; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        if (lane == 0) s_wmax[warp] = wm;
        if (tid == 0) {
            mbar_init(smem_u32(&mma_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = tmem_slot;
        wm = 0.f;
#pragma unroll
        for (int i = 0; i < SB_THREADS / 32; ++i) wm = fmaxf(wm, s_wmax[i]);
        const int s_exp = wm > 0.f ? ilogbf(8192.f / wm) : 0;
        const float w_scale = ldexpf(1.f, s_exp);
        w_inv = ldexpf(1.f, -s_exp);
        if (tid < 2 * E) {
            const int co = tid >> 1, c = tid & 1;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = Wt.pw1_w[(8 * c + j) * E + co] * w_scale;
            uint4 hi, lo;
            split8(v, hi, lo);
            const uint32_t a = sWB + (uint32_t)co * 64u + (uint32_t)((c ^ ((co >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + (uint32_t)SB_WB_PLANE), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // uint8 input: the 39 x 71 window of the NEXT tile is fetched into registers (packed b | g << 8 | r << 16 | valid << 24,
    // 5 pixels per thread) while the current tile is in its expansion / depthwise phases, so S0 never waits on HBM
    // (the exposed load latency was 8.6 % of the stall samples, profiles/r2_ncu_top5_ops_stalls.txt)
    constexpr int SB_PRE = (SB_NI + SB_THREADS - 1) / SB_THREADS;
    uint32_t pre[SB_PRE];
    auto prefetch = [&](int tl) {
        const int il = tl / tiles_per_img, tt = tl - il * tiles_per_img;
        const int y0 = 4 * ((tt / tiles_x) * SB_TH) - 5, x0 = 4 * ((tt % tiles_x) * SB_TW) - 5;      // = iy0, ix0 of that tile
        const uint8_t* src = p.in + (long long)(il + p.img0) * p.H * p.W * 3;
#pragma unroll
        for (int k = 0; k < SB_PRE; ++k) {
            const int i = tid + k * SB_THREADS;
            const int r = i / SB_IW, c = i - r * SB_IW, y = y0 + r, x = x0 + c;
            uint32_t v = 0;
            if (i < SB_NI && y >= 0 && y < p.H && x >= 0 && x < p.W) {
                const uint8_t* q = src + ((long long)y * p.W + x) * 3;
                v = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | (1u << 24);
            }
            pre[k] = v;
        }
    };
    if (!p.in_f32 && (int)blockIdx.x < p.n_tiles) prefetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int img_l = tile / tiles_per_img, t = tile - img_l * tiles_per_img, img = img_l + p.img0;
        const int oy0 = (t / tiles_x) * SB_TH, ox0 = (t % tiles_x) * SB_TW;     // quarter-res origin of the tile
        const int ey0 = 2 * oy0 - 1, ex0 = 2 * ox0 - 1;       // half-res origin of the 17 x 33 window
        const int sy0 = ey0 - 1, sx0 = ex0 - 1;               // half-res origin of the 19 x 35 stem window
        const int iy0 = 2 * sy0 - 1, ix0 = 2 * sx0 - 1;       // input origin of the 39 x 71 window
        // ---- S0: input window, /255 (true division, as numpy does), zero outside the crop (the stem's padding)
        if (p.in_f32) {
            for (int i = tid; i < SB_NI; i += SB_THREADS) {
                const int r = i / SB_IW, c = i - r * SB_IW, y = iy0 + r, x = ix0 + c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                    const long long o = (((long long)img * p.H + y) * p.W + x) * 3;
                    v.x = p.in_f32[o]; v.y = p.in_f32[o + 1]; v.z = p.in_f32[o + 2];
                }
                *reinterpret_cast<float4*>(sA + 4 * i) = v;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SB_PRE; ++k) {
                const int i = tid + k * SB_THREADS;
                if (i < SB_NI) {
                    const uint32_t u = pre[k];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (u >> 24) { v.x = lut[u & 255u]; v.y = lut[(u >> 8) & 255u]; v.z = lut[(u >> 16) & 255u]; }
                    *reinterpret_cast<float4*>(sA + 4 * i) = v;
                }
            }
        }
        __syncthreads();
        // ---- S1: stem 3x3 stride 2 + h-swish over the 19 x 35 window; item = (pixel, 8 output channels)
        for (int it = tid; it < 2 * ((SB_NS + 31) / 32) * 32; it += SB_THREADS) {
            const int hf = (it >> 5) & 1, px = ((it >> 6) << 5) | (it & 31);       // the channel half is warp-uniform
            if (px >= SB_NS) continue;
            const int r = px / SB_SW, c = px - r * SB_SW;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            const int y = sy0 + r, x = sx0 + c;
            const bool inside = y >= 0 && y < Hh && x >= 0 && x < Wh;
            if (inside) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 v = *reinterpret_cast<const float4*>(sA + 4 * ((2 * r + ky) * SB_IW + 2 * c + kx));
                        const int tp = (ky * 3 + kx) * 3;
                        if (hf == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                acc[j] = fmaf(v.x, Wt.stem_w[(tp + 0) * 16 + j], acc[j]);
                                acc[j] = fmaf(v.y, Wt.stem_w[(tp + 1) * 16 + j], acc[j]);
                                acc[j] = fmaf(v.z, Wt.stem_w[(tp + 2) * 16 + j], acc[j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                acc[j] = fmaf(v.x, Wt.stem_w[(tp + 0) * 16 + 8 + j], acc[j]);
                                acc[j] = fmaf(v.y, Wt.stem_w[(tp + 1) * 16 + 8 + j], acc[j]);
                                acc[j] = fmaf(v.z, Wt.stem_w[(tp + 2) * 16 + 8 + j], acc[j]);
                            }
                        }
                    }
                if (hf == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = hswish_f(acc[j] + Wt.stem_b[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = hswish_f(acc[j] + Wt.stem_b[8 + j]);
                }
            }
            // outside the half-resolution map the tensor is the NEXT conv's zero padding, not stem(padding)
            float* o = sB + px * SB_PS + 8 * hf;
            *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        __syncthreads();
        // ---- S2: blocks.0.0 depthwise 3x3 + ReLU over the 17 x 33 window; item = (4-pixel strip, 4-channel group)
        {
            constexpr int STRIPS = (SB_EW + 3) / 4;            // 9 strips per row, the last one 1 pixel wide
            for (int it = tid; it < SB_EH * STRIPS * 4; it += SB_THREADS) {
                const int g = it & 3, sidx = it >> 2, row = sidx / STRIPS, x0 = (sidx - row * STRIPS) * 4;
                switch (g) {
                    case 0: dw16_strip<0>(sB, sA, row, x0, Wt); break;
                    case 1: dw16_strip<1>(sB, sA, row, x0, Wt); break;
                    case 2: dw16_strip<2>(sB, sA, row, x0, Wt); break;
                    default: dw16_strip<3>(sB, sA, row, x0, Wt); break;
                }
            }
        }
        __syncthreads();
        // ---- S3: blocks.0.0 pointwise 16->16 (linear) + shortcut (the stem output at the same pixel) -> s3
        for (int px = tid; px < SB_NE; px += SB_THREADS) {
            const int r = px / SB_EW, c = px - r * SB_EW;
            float acc[16];
            pw16in<16, 0>(sA + px * SB_PS, Wt.pw0_w, Wt.pw0_b, acc);
            const float* res = sB + ((r + 1) * SB_SW + c + 1) * SB_PS;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4 rv = *reinterpret_cast<const float4*>(res + 4 * g);
                acc[4 * g] += rv.x; acc[4 * g + 1] += rv.y; acc[4 * g + 2] += rv.z; acc[4 * g + 3] += rv.w;
                if (!TC) *reinterpret_cast<float4*>(sC + px * SB_PS + 4 * g) = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
            }
            if (TC) {
                // row px of the A operand: logical 16-byte chunk c (8 channels) sits at chunk c ^ ((px / 2) % 4) of the 64-byte row
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    uint4 hi, lo;
                    split8(acc + 8 * c2, hi, lo);
                    const uint32_t a = sT + (uint32_t)px * 64u + (uint32_t)((c2 ^ ((px >> 1) & 3)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + (uint32_t)SB_T_PLANE), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
                }
            }
        }
        if (TC) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (TC) {
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = (1u << 4) | ((uint32_t)(E >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint64_t b_hi = make_smem_desc_sw64(sWB), b_lo = make_smem_desc_sw64(sWB + (uint32_t)SB_WB_PLANE);
#pragma unroll 1
                for (int m = 0; m < SB_T_TILES; ++m) {
                    const uint64_t a_hi = make_smem_desc_sw64(sT + (uint32_t)m * 8192u);
                    const uint64_t a_lo = make_smem_desc_sw64(sT + (uint32_t)SB_T_PLANE + (uint32_t)m * 8192u);
                    const uint32_t d = tmem_base + (uint32_t)(m * E);
                    umma_f16(d, a_lo, b_hi, idesc, 0u);
                    umma_f16(d, a_hi, b_lo, idesc, 1u);
                    umma_f16(d, a_hi, b_hi, idesc, 1u);
                }
                umma_commit(smem_u32(&mma_bar));
            }
            mbar_wait(smem_u32(&mma_bar), mma_phase);
            mma_phase ^= 1u;
            tc_fence_after();
        }
        // ---- S4/S5 per 16-channel chunk of the expanded tensor: 1x1 16->E + ReLU into sB, then depthwise 3x3 s2 + ReLU
        if (!p.in_f32 && tile + (int)gridDim.x < p.n_tiles) prefetch(tile + gridDim.x);     // in flight during S4 / S5
        const int y_ok0 = max(0, -ey0), y_ok1 = min(SB_EH, Hh - ey0), x_ok0 = max(0, -ex0), x_ok1 = min(SB_EW, Wh - ex0);
#pragma unroll
        for (int ch = 0; ch < E / 16; ++ch) {
            if (TC) {
                // warp = (TMEM lane quarter q, MMA tile m): the accumulator rows of 32 window pixels, 16 expanded channels
                const int q = warp & 3, m = warp >> 2, px = 128 * m + 32 * q + lane;
                float acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(m * E + 16 * ch), acc);
                if (px < SB_NE) {
                    const int r = px / SB_EW, c = px - r * SB_EW;
                    const bool inside = r >= y_ok0 && r < y_ok1 && c >= x_ok0 && c < x_ok1;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float4 o;
                        o.x = fmaxf(fmaf(acc[4 * g], w_inv, Wt.pw1_b[ch * 16 + 4 * g]), 0.f);
                        o.y = fmaxf(fmaf(acc[4 * g + 1], w_inv, Wt.pw1_b[ch * 16 + 4 * g + 1]), 0.f);
                        o.z = fmaxf(fmaf(acc[4 * g + 2], w_inv, Wt.pw1_b[ch * 16 + 4 * g + 2]), 0.f);
                        o.w = fmaxf(fmaf(acc[4 * g + 3], w_inv, Wt.pw1_b[ch * 16 + 4 * g + 3]), 0.f);
                        *reinterpret_cast<float4*>(sB + px * SB_PS + 4 * g) = inside ? o : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                tc_fence_before();                      // the next tile's MMAs overwrite these TMEM columns after a barrier
            } else
            for (int px = tid; px < SB_NE; px += SB_THREADS) {
                const int r = px / SB_EW, c = px - r * SB_EW;
                float acc[16];
                if (ch == 0) pw16in<E, 0>(sC + px * SB_PS, Wt.pw1_w, Wt.pw1_b, acc);
                else if (ch == 1) pw16in<E, 16>(sC + px * SB_PS, Wt.pw1_w, Wt.pw1_b, acc);
                else if (ch == 2) pw16in<E, 32>(sC + px * SB_PS, Wt.pw1_w, Wt.pw1_b, acc);
                else pw16in<E, 48>(sC + px * SB_PS, Wt.pw1_w, Wt.pw1_b, acc);
                // zero outside the half-resolution map: the stride-2 depthwise's padding
                const bool inside = r >= y_ok0 && r < y_ok1 && c >= x_ok0 && c < x_ok1;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<float4*>(sB + px * SB_PS + 4 * g) = inside
                        ? make_float4(fmaxf(acc[4 * g], 0.f), fmaxf(acc[4 * g + 1], 0.f), fmaxf(acc[4 * g + 2], 0.f), fmaxf(acc[4 * g + 3], 0.f))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            // depthwise 3x3 stride 2: item = (output pixel, 4-channel group) = 128 x 4 = 512 items
            if (tid < SB_TH * SB_TW * 4) {
                const int g = tid & 3, opx = tid >> 2, orow = opx / SB_TW, ocol = opx - orow * SB_TW;
                const int c0 = ch * 16 + 4 * g;
                float4 acc = *reinterpret_cast<const float4*>(sW + 9 * E + c0);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 v = *reinterpret_cast<const float4*>(sB + ((2 * orow + ky) * SB_EW + 2 * ocol + kx) * SB_PS + 4 * g);
                        const float4 w = *reinterpret_cast<const float4*>(sW + (ky * 3 + kx) * E + c0);
                        acc.x = fmaf(v.x, w.x, acc.x); acc.y = fmaf(v.y, w.y, acc.y);
                        acc.z = fmaf(v.z, w.z, acc.z); acc.w = fmaf(v.w, w.w, acc.w);
                    }
                acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
                const long long o = (((long long)img * p.Hq + oy0 + orow) * p.Wq + ox0 + ocol) * p.out_ld + p.out_coff + c0;
                st4(p.out, p.out_fmt, p.out_plane, o, acc);
            }
            __syncthreads();
        }
    }
    if (TC) {
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        }
    }
}

bool stem_block_supported(int H, int W, int E, const TView& out) {
    if (H % 32 || W % 64 || (E != 64)) return false;          // whole 8 x 16 quarter-resolution tiles; E fixed by the template
    return out.base && out.c_stride == 1 && out.C == E && out.H == H / 4 && out.W == W / 4 && !((out.ld | out.c_off) & 3) &&
           (out.fmt == DT_F32 || out.fmt == DT_SPLIT16);
}

int stem_block_launch(const StemBlockK& k, const StemBlockW& w, int num_sms, cudaStream_t s) {
    static int use_tc = -1;
    if (use_tc < 0) {
        const char* e = getenv("SKPS_STEM_TC");
        use_tc = (e && e[0] == '0') ? 0 : 1;
        SKPS_CUDA(cudaFuncSetAttribute(stem_block_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM));
        SKPS_CUDA(cudaFuncSetAttribute(stem_block_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SB_SMEM_TC));
    }
    const int grid = k.n_tiles < num_sms ? k.n_tiles : num_sms;
    if (use_tc) stem_block_kernel<64, true><<<grid, SB_THREADS, SB_SMEM_TC, s>>>(k, w);
    else stem_block_kernel<64, false><<<grid, SB_THREADS, SB_SMEM, s>>>(k, w);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace skps
