// Dense k x k convolution as a TRANSPOSED tcgen05 implicit GEMM (sm_100a): output channels on the TMEM lanes, pixels as N.
//
//   D[M = 128 channel rows][N = 256 pixels] = sum over taps, ci of  W[co][tap][ci] * X[pixel + tap][ci]
//
// Why: conv_tc.cu computes C[128 pixels][Cout]; with Cout = 128 every SS-mode MMA (M = N = 128, K = 16) reads 4 KB of A and
// 4 KB of B from shared memory in 64 cycles - 128 B/clk, the SM's whole shared-memory bandwidth - so the TMA fill and the
// epilogue staging compete with the tensor pipe and the decoder's conv2 (40 % of the student's MACs) sits at 63 % tensor-pipe
// utilisation however its tiles arrive (profiles/r2_ncu_conv2_*.txt).  With the pixels as N = 256 one instruction does twice
// the work on 12 KB of operands (A: the 128 x 16 weight slice, B: 256 pixels x 16 channels) = 96 B/clk.  Same fp16 hi/lo
// three-product scheme, same 4-D TMA activation boxes (padding / dilation = OOB fill) and K-padded weight matrix as conv_tc.
//
// Tile = 256 pixels = bh whole rows of one image (W | 256).  Stage = one (tap, 64-channel chunk): weights 2 x 16 KB +
// activations 2 x 32 KB; two stages, two TMEM accumulator stages (2 x 256 columns).  Epilogue: a thread owns one output channel
// and 32 pixels per tcgen05.ld; bias / activation / fp16 hi-lo split, then the 32 x 32 block is written TRANSPOSED into the
// warp's 4 KB staging slice (one 64-byte pixel row per store instruction, lane = channel) and leaves as one TMA store per
// plane, so the NHWC layout of the output is unchanged.
//
// Halo-row stages (k3, 3x3 / dilation 1 / W in {32, 64}): as in conv_tc's opt-in mode a stage is (kx, 32-channel half): ONE
// (bh+2)-row activation box (64-byte swizzled rows) serves the three ky taps - the B operand of tap ky is the 256 box rows that
// start ky image rows in, a descriptor offset of ky * W * 64 bytes - next to the three ky weight tiles.  L2->SM bytes per tile
// drop by a third (1728 -> 1152 KB at Cin = 128).  Opt-in (SKPS_TCT_K3=1): measured slower than the per-tap stages, see
// tct_prepare.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/skps_b200.h"
#include "common.h"
#include "conv_tct.h"
#include "tc_ptx.h"

namespace skps {

constexpr int TCT_THREADS = 384;         // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr int TCT_M = 128, TCT_N = 256;
constexpr int TCT_W_TILE = TCT_M * 128;  // weights of one k-block, one plane: 128 rows x 128 B
constexpr int TCT_X_TILE = TCT_N * 128;  // activations of one k-block, one plane: 256 pixel rows x 128 B
constexpr int TCT_STAGE = 2 * TCT_W_TILE + 2 * TCT_X_TILE;      // 96 KB
constexpr int TCT_STAGES = 2;

__global__ void __launch_bounds__(TCT_THREADS, 1)
conv_tct_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo, const TctK p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TCT_STAGES], empty_bar[TCT_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_off = base + (uint32_t)TCT_STAGES * TCT_STAGE;       // 8 epilogue warps x 4 KB
    const int tiles = p.m_tiles;
    const int kblocks = p.taps * p.cchunks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TCT_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(smem_u32(&tfull_bar[a]), 1);
            mbar_init(smem_u32(&tempty_bar[a]), 8);          // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int img_l = tile / p.tiles_per_img, t = tile - img_l * p.tiles_per_img;
                const int y0 = t * p.bh;
                if (p.k3) {
                    const int halves = p.Cin >> 5, K_row = p.cchunks * 64;
                    const uint32_t stage_bytes = 6u * 8192u + 2u * (uint32_t)p.xb;
                    for (int kb = 0; kb < 3 * halves; ++kb) {
                        const int kx = kb / halves, h = kb - kx * halves;
                        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                        const uint32_t fb = smem_u32(&full_bar[stage]);
                        mbar_expect_tx(fb, stage_bytes);
                        const uint32_t ss = base + (uint32_t)stage * TCT_STAGE;
                        for (int ky = 0; ky < 3; ++ky) {
                            const int kcol = (ky * 3 + kx) * K_row + h * 32;
                            tma_load_2d(ss + (uint32_t)(2 * ky) * 8192u, &tmW_hi, fb, kcol, 0);
                            tma_load_2d(ss + (uint32_t)(2 * ky + 1) * 8192u, &tmW_lo, fb, kcol, 0);
                        }
                        tma_load_4d(ss + 49152u, &tmX_hi, fb, h * 32, kx - 1, y0 - 1, img_l + p.img0);
                        tma_load_4d(ss + 49152u + (uint32_t)p.xb, &tmX_lo, fb, h * 32, kx - 1, y0 - 1, img_l + p.img0);
                        if (++stage == TCT_STAGES) { stage = 0; phase ^= 1u; }
                    }
                    continue;
                }
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
                    const int ky = tap / p.kw, kx = tap - ky * p.kw;
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(fb, (uint32_t)TCT_STAGE);
                    const uint32_t ss = base + (uint32_t)stage * TCT_STAGE;
                    tma_load_2d(ss, &tmW_hi, fb, kb * 64, 0);
                    tma_load_2d(ss + TCT_W_TILE, &tmW_lo, fb, kb * 64, 0);
                    const int cx = kx * p.dil - p.pad, cy = y0 + ky * p.dil - p.pad;
                    tma_load_4d(ss + 2 * TCT_W_TILE, &tmX_hi, fb, cc * 64, cx, cy, img_l + p.img0);
                    tma_load_4d(ss + 2 * TCT_W_TILE + TCT_X_TILE, &tmX_lo, fb, cc * 64, cx, cy, img_l + p.img0);
                    if (++stage == TCT_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one thread)
        if (lane == 0) {
            // instruction descriptor: D = f32, A = B = f16, both K-major, N = 256 pixels, M = 128 channel rows
            const uint32_t idesc = (1u << 4) | ((uint32_t)(TCT_N >> 3) << 17) | ((uint32_t)(TCT_M >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * (uint32_t)TCT_N;
                if (p.k3) {
                    const int nkb = 3 * (p.Cin >> 5);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(smem_u32(&full_bar[stage]), phase);
                        tc_fence_after();
                        const uint32_t ss = base + (uint32_t)stage * TCT_STAGE;
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint64_t w_hi = make_smem_desc_sw64(ss + (uint32_t)(2 * ky) * 8192u);
                            const uint64_t w_lo = make_smem_desc_sw64(ss + (uint32_t)(2 * ky + 1) * 8192u);
                            // tap row ky: the 256 box rows that start ky image rows in
                            const uint32_t xo = ss + 49152u + (uint32_t)(ky * p.W * 64);
                            const uint64_t x_hi = make_smem_desc_sw64(xo), x_lo = make_smem_desc_sw64(xo + (uint32_t)p.xb);
                            for (int k = 0; k < 2; ++k) {
                                const uint64_t koff = (uint64_t)(k * 2);
                                umma_f16(d_tmem, w_lo + koff, x_hi + koff, idesc, (kb | ky | k) != 0);
                                umma_f16(d_tmem, w_hi + koff, x_lo + koff, idesc, 1u);
                                umma_f16(d_tmem, w_hi + koff, x_hi + koff, idesc, 1u);
                            }
                        }
                        umma_commit(smem_u32(&empty_bar[stage]));
                        if (++stage == TCT_STAGES) { stage = 0; phase ^= 1u; }
                    }
                } else
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    tc_fence_after();
                    const uint32_t ss = base + (uint32_t)stage * TCT_STAGE;
                    const uint64_t w_hi = make_smem_desc(ss), w_lo = make_smem_desc(ss + TCT_W_TILE);
                    const uint64_t x_hi = make_smem_desc(ss + 2 * TCT_W_TILE), x_lo = make_smem_desc(ss + 2 * TCT_W_TILE + TCT_X_TILE);
                    const int cc = kb % p.cchunks;
                    const int ksteps = min(4, (p.Cin - cc * 64 + 15) / 16);
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t koff = (uint64_t)(k * 32 >> 4);
                        // small terms first, then the dominant hi*hi product
                        umma_f16(d_tmem, w_lo + koff, x_hi + koff, idesc, (kb | k) != 0);
                        umma_f16(d_tmem, w_hi + koff, x_lo + koff, idesc, 1u);
                        umma_f16(d_tmem, w_hi + koff, x_hi + koff, idesc, 1u);
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));
                    if (++stage == TCT_STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(smem_u32(&tfull_bar[acc]));
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ================================================================== epilogue: lane = output channel, columns = pixels
        const int q = warp & 3;                        // TMEM lane quarter = channels 32q .. 32q+31
        const int half_id = (warp - 4) >> 2;           // pixel columns [0,128) or [128,256)
        const int c = q * 32 + lane;
        const float bias = c < p.Cout ? __ldg(p.bias + c) : 0.f;
        const uint32_t sbuf = out_off + (uint32_t)(warp - 4) * 4096u;       // [hi: 32 pixel rows x 64 B][lo: same]
        const bool q_ok = q * 32 < p.Cout;             // quarters past Cout hold zero rows: nothing to store
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const int img_l = tile / p.tiles_per_img, t = tile - img_l * p.tiles_per_img;
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (uint32_t)acc * (uint32_t)TCT_N + (uint32_t)(half_id * 128) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {
                float v[32];
                tmem_ld32(t_addr + (uint32_t)(ci * 32), v);
                if (!q_ok) continue;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the slice's previous store has drained
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float f = apply_act(fmaf(v[j], p.out_scale, bias), p.act);
                    const __half h = __float2half_rn(f);
                    const __half l = __float2half_rn(f - __half2float(h));
                    // pixel row j of the block: 64 bytes = 32 channels; the 32 lanes fill one row per store instruction
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(sbuf + (uint32_t)(j * 64 + lane * 2)), "h"(__half_as_ushort(h)) : "memory");
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(sbuf + 2048u + (uint32_t)(j * 64 + lane * 2)), "h"(__half_as_ushort(l)) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const int col0 = half_id * 128 + ci * 32;              // first pixel of the block inside the tile
                    const int x0 = col0 % p.W, y0 = t * p.bh + col0 / p.W;
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(&tmO_hi), "r"(sbuf), "r"(q * 32), "r"(x0), "r"(y0), "r"(img_l + p.img0) : "memory");
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(&tmO_lo), "r"(sbuf + 2048u), "r"(q * 32), "r"(x0), "r"(y0), "r"(img_l + p.img0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tct_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// The layers this kernel is for: tensor-bound k x k convs whose Cout fills the 128 TMEM lanes (else the lanes idle and the
// pixels-on-lanes kernel wins), whole 256-pixel row blocks, split-fp16 contiguous output, no residual.
bool tct_applicable(const TcSetup& s) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SKPS_TCT"); on = (e && e[0] == '0') ? 0 : 1; }
    if (!on) return false;
    const int stride = s.stride > 0 ? s.stride : 1;
    if (stride != 1 || s.kh != s.kw || s.kh < 3 || s.pad != s.dil * (s.kh - 1) / 2) return false;
    // Cout threshold: below 96 the idle TMEM lanes cost more than the operand-bandwidth relief buys (measured with
    // SKPS_TCT_MINC=64 on the ASPP 3x3 convs, Cout 64 on 16x16 maps: 73 vs 70 us on the pixels-on-lanes kernel)
    static int min_c = -1;
    if (min_c < 0) { const char* e = getenv("SKPS_TCT_MINC"); min_c = e ? atoi(e) : 96; }
    if (s.W < 16 || s.W > TCT_N || TCT_N % s.W || s.H % (TCT_N / s.W)) return false;
    if (s.Cin < 64 || (s.Cin % 8) || (s.in_ld % 8) || (s.in_coff % 8)) return false;
    if (s.Cout < min_c || s.Cout > TCT_M || (s.Cout % 8) || s.n_tiles != 1) return false;
    if (s.res || s.hm_val || s.out_fmt != DT_SPLIT16 || s.out_cstride != 1 || (s.out_ld % 8) || (s.out_coff % 8)) return false;
    return true;
}

int tct_prepare(TctLayer& L, const TcSetup& s) {
    EncodeTiledFn enc = tct_encode();
    SKPS_CHECK(enc, "cuTensorMapEncodeTiled entry point not available");
    SKPS_CHECK(tct_applicable(s), "conv_tct: layer not applicable");
    TctK& k = L.k;
    memset(&k, 0, sizeof(k));
    k.W = s.W; k.bh = TCT_N / s.W; k.tiles_per_img = s.H / k.bh;
    k.taps = s.kh * s.kw; k.kw = s.kw; k.dil = s.dil; k.pad = s.pad;
    k.cchunks = (s.Cin + 63) / 64; k.Cin = s.Cin; k.Cout = s.Cout; k.act = s.act; k.out_scale = s.out_scale;
    SKPS_CHECK(s.bias, "conv_tct: bias required");
    k.bias = s.bias;
    L.smem_bytes = TCT_STAGES * TCT_STAGE + 8 * 4096 + 1024;
    {
        // opt-in: measured on B200 the halo-row stages are SLOWER here (conv2 0.706 vs 0.649 ms) as they were in conv_tc
        // (no gain) - both times the 64-byte-row operand layout is the common factor, so the third fewer L2->SM bytes do
        // not pay for it.  Kept (and unit-tested with SKPS_TCT_K3=1) as the starting point for a 128-byte-row variant.
        const char* e = getenv("SKPS_TCT_K3");
        const int k3_on = (e && e[0] == '1') ? 1 : 0;
        k.xb = (k.bh + 2) * s.W * 64;
        k.k3 = (k3_on && s.kh == 3 && s.dil == 1 && (s.W == 32 || s.W == 64) && s.Cin % 32 == 0 &&
                6 * 8192 + 2 * k.xb <= TCT_STAGE) ? 1 : 0;
    }
    for (int plane = 0; plane < 2; ++plane) {
        cuuint64_t dims[4] = {(cuuint64_t)s.Cin, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.max_batch};
        cuuint64_t strides[3] = {(cuuint64_t)s.in_ld * 2, (cuuint64_t)s.W * s.in_ld * 2, (cuuint64_t)s.H * s.W * s.in_ld * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)s.W, (cuuint32_t)k.bh, 1};
        if (k.k3) { box[0] = 32; box[2] = (cuuint32_t)(k.bh + 2); }
        cuuint32_t estr[4] = {1, 1, 1, 1};
        void* base = (void*)((__half*)s.in_base + (plane ? s.in_plane : 0) + s.in_coff);
        CUresult r = enc(plane ? &L.x_lo : &L.x_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, k.k3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SKPS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tct X) failed: %d", (int)r);
    }
    const int K_pad = k.taps * k.cchunks * 64;
    for (int plane = 0; plane < 2; ++plane) {
        cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)s.n_tile};        // rows beyond n_tile: OOB zero fill
        cuuint64_t strides[1] = {(cuuint64_t)K_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)(k.k3 ? 32 : 64), (cuuint32_t)TCT_M};
        cuuint32_t estr[2] = {1, 1};
        void* base = (void*)(plane ? s.w_lo : s.w_hi);
        CUresult r = enc(plane ? &L.w_lo : &L.w_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, k.k3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SKPS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tct W) failed: %d", (int)r);
    }
    // output: one box per epilogue warp and 32-pixel block: 32 channels x 32 consecutive pixels of a row, plain 64-byte rows
    for (int plane = 0; plane < 2; ++plane) {
        cuuint64_t dims[4] = {(cuuint64_t)s.Cout, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.max_batch};
        cuuint64_t strides[3] = {(cuuint64_t)s.out_ld * 2, (cuuint64_t)s.W * s.out_ld * 2, (cuuint64_t)s.H * s.W * s.out_ld * 2};
        const cuuint32_t obw = (cuuint32_t)(s.W < 32 ? s.W : 32);                // 32 consecutive pixels = 32/obw whole rows
        cuuint32_t box[4] = {32, obw, 32 / obw, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        char* base = (char*)s.out + (size_t)s.out_coff * 2 + (plane ? (size_t)s.out_plane * 2 : 0);
        CUresult r = enc(plane ? &L.o_lo : &L.o_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SKPS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tct out) failed: %d", (int)r);
    }
    L.valid = true;
    return 0;
}

int tct_launch(const TctLayer& L, int batch, int img0, int num_sms, cudaStream_t stream) {
    static int attr_bytes = 0;
    if (L.smem_bytes > attr_bytes) {
        SKPS_CUDA(cudaFuncSetAttribute(conv_tct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_bytes));
        attr_bytes = L.smem_bytes;
    }
    TctK k = L.k;
    k.m_tiles = batch * k.tiles_per_img;
    k.img0 = img0;
    const int grid = k.m_tiles < num_sms ? k.m_tiles : num_sms;
    conv_tct_kernel<<<grid, TCT_THREADS, L.smem_bytes, stream>>>(L.x_hi, L.x_lo, L.w_hi, L.w_lo, L.o_hi, L.o_lo, k);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace skps
