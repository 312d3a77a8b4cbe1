// Dense convolution as a tcgen05 implicit GEMM (sm_100a): the landmark network's 1x1 and 3x3
// (incl. dilated) convolutions, 96.7 % of its MACs.
//
//   C[M = N*H*W pixels][Cout] = sum over taps (ky,kx), ci of  A[pixel + tap][ci] * W[co][tap][ci]
//
// * Operands are float16 hi/lo pairs: v = hi + lo with hi = fp16(v), lo = fp16(v - hi) (22 mantissa
//   bits).  Each K-step issues three kind::f16 MMAs into one fp32 TMEM accumulator:
//   hi*hi + hi*lo + lo*hi.  Measured against the fp32 oracle this keeps landmarks within 6e-5 px
//   (budget 1e-3 px); plain fp16/bf16/tf32 single-pass does not (0.02-0.5 px).  Tensor work issued is
//   therefore 3x the algorithmic FLOPs.
// * A tiles (128 pixels x 64 channels) are fetched by 4-D TMA straight from the NHWC activation:
//   box (64 ch, bw, bh) with signed start coordinates, so the conv zero padding and dilation come
//   from TMA out-of-bounds fill; no im2col buffer exists.  B tiles (n_tile x 64) come from the
//   pre-split, K-padded weight matrix.  Both land in 128B-swizzled shared memory.
// * Warp roles: warp 0 TMA producer, warp 1 MMA issuer (single thread), warp 2 TMEM allocator,
//   warps 4-7 epilogue (TMEM -> registers -> bias/act/residual -> global).  Persistent CTAs, one per
//   SM, two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Halo-row mode (k3) for 3x3 / stride 1 / dilation 1 convs on 64-wide maps (the decoder's conv2, 40 % of the student's MACs):
// the per-tap pipeline above fetches every activation row once per tap - 9 x 64 KB of A tiles per 256 pixels, and the kernel
// was L2->SM bound (lts 71-79 % of its cap, tensor pipe 53 %).  In k3 mode a pipeline stage is (kx, 32-channel half-chunk):
// ONE 6-row x 64-pixel box (rows y0-1 .. y0+4, shifted by kx-1 in x; padding = TMA OOB fill) serves the three ky taps of a
// 4-row x 64-pixel output group: tap ky of pixel tile u is the 128 rows that start (2u + ky) image rows into the box, i.e. a
// shared-memory descriptor offset of (2u + ky) x 4 KB - the taps are addressed in place, nothing is copied.  64-byte rows
// (SWIZZLE_64B) keep two 96 KB stages (A 48 KB + the three ky weight tiles 48 KB) inside shared memory.  L2->SM bytes per
// 256 pixels: 1728 KB -> 1152 KB.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "../../include/skps_b200.h"
#include "common.h"
#include "conv_tc.h"
#include "conv_tct.h"
#include "tc_ptx.h"

namespace skps {

// ------------------------------------------------------------------------------------------ kernel
constexpr int TC_BM = 128;            // pixels per tile (UMMA M)
constexpr int TC_BK = 64;             // channels per k-block (one 128-byte swizzle atom of fp16)
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_THREADS = 384;        // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 4;
constexpr int K3_ROWS = 6;                          // input rows per halo box: 4 output rows + 2
constexpr int K3_A_PLANE = K3_ROWS * 64 * 64;       // 6 rows x 64 pixels x 64 B (32 fp16 channels) = 24 KB per plane

// One level of a warp reduce-scatter of per-column (max, first arg-max): 2N column candidates per lane in, N out; after the
// levels 16, 8, 4, 2, 1 lane L holds column L reduced over the warp's 32 rows.  Ties keep the smaller pixel index.
template <int N>
__device__ __forceinline__ void argmax_scatter(float* v, int* id, int off, int lane) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float keep = hi ? v[j + N] : v[j], snd = hi ? v[j] : v[j + N];
        const int kid = hi ? id[j + N] : id[j], sid = hi ? id[j] : id[j + N];
        const float r = __shfl_xor_sync(0xffffffffu, snd, off);
        const int ri = __shfl_xor_sync(0xffffffffu, sid, off);
        const bool take = r > keep || (r == keep && ri < kid);
        v[j] = take ? r : keep;
        id[j] = take ? ri : kid;
    }
}

template <int ACT, bool OUT_SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ CUtensorMap tmO_hi, const __grid_constant__ CUtensorMap tmO_lo, const TcK p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_tile_bytes = (uint32_t)p.n_tile * TC_BK * 2;
    // stage = [A0_hi][A0_lo]([A1_hi][A1_lo])[B_hi][B_lo]: with mt = 2 two pixel tiles share one weight tile
    // k3: stage = [A hi 24 KB][A lo 24 KB][ky = 0,1,2: B hi, B lo of n_tile x 64 B each]
    const uint32_t b3_bytes = (uint32_t)p.n_tile * 64u;
    const uint32_t a_bytes = p.k3 ? 2u * K3_A_PLANE : (uint32_t)p.mt * 2u * A_TILE_BYTES;
    const uint32_t stage_bytes = p.k3 ? a_bytes + 6u * b3_bytes : a_bytes + 2u * b_tile_bytes;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
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
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    // epilogue staging for TMA stores: one 16 KB buffer (128 rows x 32 channels, hi+lo or float32) per warp group
    const uint32_t stage_out = tile_base + (uint32_t)p.stages * stage_bytes;
    const int m_groups = (p.m_tiles + p.mt - 1) / p.mt;
    const int total_tiles = m_groups * p.n_tiles;          // work items: (group of mt pixel tiles) x N tile
    const int kblocks = p.taps * p.cchunks;

    if (warp == 0) {
        // ================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int tiles_x = (p.W + p.bw - 1) / p.bw;       // ragged maps: edge tiles hang over, TMA zero-fills / clips
            if (p.k3) {
                const int halves = p.Cin >> 5, gpi = p.tiles_per_img >> 1, K_row = p.cchunks * TC_BK;
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int img_l = tile / gpi, y0 = (tile - img_l * gpi) * 4;
                    for (int kb = 0; kb < 3 * halves; ++kb) {
                        const int kx = kb / halves, h = kb - kx * halves;
                        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                        const uint32_t fb = smem_u32(&full_bar[stage]);
                        mbar_expect_tx(fb, stage_bytes);
                        const uint32_t sa = tile_base + (uint32_t)stage * stage_bytes;
                        tma_load_4d(sa, &tmA_hi, fb, h * 32, kx - 1, y0 - 1, img_l + p.img0);
                        tma_load_4d(sa + K3_A_PLANE, &tmA_lo, fb, h * 32, kx - 1, y0 - 1, img_l + p.img0);
                        for (int ky = 0; ky < 3; ++ky) {
                            const int kcol = (ky * 3 + kx) * K_row + h * 32;
                            tma_load_2d(sa + a_bytes + (uint32_t)(2 * ky) * b3_bytes, &tmB_hi, fb, kcol, 0);
                            tma_load_2d(sa + a_bytes + (uint32_t)(2 * ky + 1) * b3_bytes, &tmB_lo, fb, kcol, 0);
                        }
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            } else
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int g_idx = tile / p.n_tiles, n_idx = tile - g_idx * p.n_tiles;
                int img[2], y0[2], x0[2];
                for (int u = 0; u < p.mt; ++u) {
                    const int m_idx = min(g_idx * p.mt + u, p.m_tiles - 1);     // odd tail: reload the last tile, result unused
                    const int img_l = p.ipt > 1 ? m_idx * p.ipt : m_idx / p.tiles_per_img;
                    const int t = p.ipt > 1 ? 0 : m_idx - img_l * p.tiles_per_img;
                    img[u] = img_l + p.img0;
                    y0[u] = (t / tiles_x) * p.bh * p.stride;                    // input-space origin of the tile
                    x0[u] = (t % tiles_x) * p.bw * p.stride;
                }
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(fb, stage_bytes);
                    const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
                    const int ky = tap / p.kw, kx = tap - ky * p.kw;
                    const uint32_t sa = tile_base + (uint32_t)stage * stage_bytes;
                    for (int u = 0; u < p.mt; ++u) {
                        const int cx = x0[u] + kx * p.dil - p.pad, cy = y0[u] + ky * p.dil - p.pad;
                        tma_load_4d(sa + (uint32_t)u * 2u * A_TILE_BYTES, &tmA_hi, fb, cc * TC_BK, cx, cy, img[u]);
                        tma_load_4d(sa + (uint32_t)u * 2u * A_TILE_BYTES + A_TILE_BYTES, &tmA_lo, fb, cc * TC_BK, cx, cy, img[u]);
                    }
                    tma_load_2d(sa + a_bytes, &tmB_hi, fb, kb * TC_BK, n_idx * p.n_tile);
                    tma_load_2d(sa + a_bytes + b_tile_bytes, &tmB_lo, fb, kb * TC_BK, n_idx * p.n_tile);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (one thread)
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=f16, K-major both, N = n_tile, M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                if (p.k3) {
                    const int nkb = 3 * (p.Cin >> 5);
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(smem_u32(&full_bar[stage]), phase);
                        tc_fence_after();
                        const uint32_t sa = tile_base + (uint32_t)stage * stage_bytes;
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint64_t b_hi = make_smem_desc_sw64(sa + a_bytes + (uint32_t)(2 * ky) * b3_bytes);
                            const uint64_t b_lo = make_smem_desc_sw64(sa + a_bytes + (uint32_t)(2 * ky + 1) * b3_bytes);
                            for (int u = 0; u < 2; ++u) {
                                // pixel tile u, tap row ky: 128 box rows starting (2u + ky) image rows in = (2u + ky) * 4 KB
                                const uint32_t ao = sa + (uint32_t)(2 * u + ky) * 4096u;
                                const uint64_t a_hi = make_smem_desc_sw64(ao), a_lo = make_smem_desc_sw64(ao + K3_A_PLANE);
                                const uint32_t d_u = d_tmem + (uint32_t)u * 128u;
                                for (int k = 0; k < 2; ++k) {
                                    const uint64_t koff = (uint64_t)(k * 2);             // 16 fp16 = 32 bytes along K
                                    umma_f16(d_u, a_lo + koff, b_hi + koff, idesc, (kb | ky | k) != 0);
                                    umma_f16(d_u, a_hi + koff, b_lo + koff, idesc, 1u);
                                    umma_f16(d_u, a_hi + koff, b_hi + koff, idesc, 1u);
                                }
                            }
                        }
                        umma_commit(smem_u32(&empty_bar[stage]));
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                } else
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase);
                    tc_fence_after();
                    const uint32_t sa = tile_base + (uint32_t)stage * stage_bytes;
                    const uint64_t b_hi = make_smem_desc(sa + a_bytes);
                    const uint64_t b_lo = make_smem_desc(sa + a_bytes + b_tile_bytes);
                    // only the 16-channel steps that hold real channels (TMA zero-fills the rest)
                    const int cc = kb % p.cchunks;
                    const int ksteps = min(TC_BK / 16, (p.Cin - cc * TC_BK + 15) / 16);
                    for (int u = 0; u < p.mt; ++u) {
                        const uint64_t a_hi = make_smem_desc(sa + (uint32_t)u * 2u * A_TILE_BYTES);
                        const uint64_t a_lo = make_smem_desc(sa + (uint32_t)u * 2u * A_TILE_BYTES + A_TILE_BYTES);
                        const uint32_t d_u = d_tmem + (uint32_t)u * 128u;         // second pixel tile: columns +128 (mt=2 needs N <= 128)
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t koff = (uint64_t)(k * 32 >> 4);       // 16 fp16 = 32 bytes along K
                            // small terms first, then the dominant hi*hi product
                            umma_f16(d_u, a_lo + koff, b_hi + koff, idesc, (kb | k) != 0);
                            umma_f16(d_u, a_hi + koff, b_lo + koff, idesc, 1u);
                            umma_f16(d_u, a_hi + koff, b_hi + koff, idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));                // frees the smem stage when the MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(smem_u32(&tfull_bar[acc]));                      // accumulator ready for the epilogue
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ================================================================== epilogue
        // 8 warps: warp%4 selects the TMEM lane quarter (hardware rule), (warp-4)/4 the odd/even
        // 32-column chunks.  One thread = one output pixel x 32 consecutive channels per chunk.
        const int q = warp & 3;
        const int half_id = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        int store_i = 0;                 // TMA stores issued by this warp group so far
        int chunk_ctr = 0;               // 32-column chunks handed out so far (same sequence in both groups)
        const int tiles_x = (p.W + p.bw - 1) / p.bw;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int g_idx = tile / p.n_tiles, n_idx = tile - g_idx * p.n_tiles;
            mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
            tc_fence_after();
          for (int u = 0; u < p.mt; ++u) {
            const int m_idx = g_idx * p.mt + u;
            if (m_idx >= p.m_tiles) break;
            const int img_l = p.ipt > 1 ? m_idx * p.ipt : m_idx / p.tiles_per_img, img = img_l + p.img0;
            const int t = p.ipt > 1 ? 0 : m_idx - img_l * p.tiles_per_img;
            // multi-image tiles (ipt > 1): bw = W, bh = ipt*H, so y runs past H into the next image and the linear
            // pixel index below is still right; rows of images past the batch are computed but never stored
            const int y = (t / tiles_x) * p.bh + row / p.bw, x = (t % tiles_x) * p.bw + row % p.bw;
            // rows of an edge tile that hang over the map (ragged tiling) are computed but never stored
            const bool row_ok = p.ipt == 1 ? (y < p.H && x < p.W) : img + y / p.H < p.img_end;
            const long long pix = row_ok ? ((long long)img * p.H + y) * p.W + x : 0;
            const uint32_t t_addr = tmem_base + (uint32_t)acc * 256u + (uint32_t)u * 128u + ((uint32_t)(q * 32) << 16);
            const int co_tile = n_idx * p.n_tile;
            // 32-column chunks alternate between the two warp groups with a counter that runs on across tiles: layers with an
            // odd number of chunks per tile (Cout 24, 72, 40 ...) would otherwise leave all / two thirds of the epilogue to
            // group 0 (the epilogue, not HBM, bounds those layers)
            const int n_chunks = (min(p.n_tile, p.Cout - co_tile) + 31) >> 5;
            for (int ci = 0; ci < n_chunks; ++ci) {
                if (((chunk_ctr + ci) & 1) != half_id) continue;
                const int c0 = ci * 32;
                float v[32];
                tmem_ld32(t_addr + (uint32_t)c0, v);
                const int co0 = co_tile + c0;
                const int nvalid = min(min(32, p.n_tile - c0), p.Cout - co0);   // n_tile need not be a multiple of 32
                if (p.hm_val) {
                    // heat-map head (model.py:511-554 postp): only max and first arg-max of every score map are needed.
                    // Scores = acc * scale + bias (no activation); reduce the tile's 128 pixels per channel here and write
                    // (max, pixel index) per tile instead of the map (436 MB per 256-face batch that hm_decode re-read).
                    int id[32];
                    const int my = row_ok ? y * p.W + x : 0x7fffffff;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = (row_ok && j < nvalid) ? fmaf(v[j], p.out_scale, __ldg(p.bias + co0 + (j < nvalid ? j : 0))) : -INFINITY;
                        id[j] = my;
                    }
                    argmax_scatter<16>(v, id, 16, lane); argmax_scatter<8>(v, id, 8, lane); argmax_scatter<4>(v, id, 4, lane);
                    argmax_scatter<2>(v, id, 2, lane); argmax_scatter<1>(v, id, 1, lane);
                    // per-warp column maxima are combined across the 4 lane quarters through the (otherwise unused) store
                    // staging area of the dynamic shared memory: [8 warps][32] float then [8][32] int
                    float(*hm_sv)[32] = reinterpret_cast<float(*)[32]>(smem_raw + (stage_out - smem_u32(smem_raw)));
                    int(*hm_si)[32] = reinterpret_cast<int(*)[32]>(smem_raw + (stage_out - smem_u32(smem_raw)) + 1024);
                    hm_sv[warp - 4][lane] = v[0];
                    hm_si[warp - 4][lane] = id[0];
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half_id) : "memory");
                    if (q == 0) {
                        float bv = hm_sv[half_id * 4][lane];
                        int bi = hm_si[half_id * 4][lane];
#pragma unroll
                        for (int w2 = 1; w2 < 4; ++w2) {
                            const float ov = hm_sv[half_id * 4 + w2][lane];
                            const int oi = hm_si[half_id * 4 + w2][lane];
                            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                        }
                        if (lane < nvalid) {
                            const long long o = ((long long)img * p.tiles_per_img + t) * p.hm_ld + co0 + lane;
                            p.hm_val[o] = bv;
                            p.hm_idx[o] = bi;
                        }
                    }
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half_id) : "memory");
                    continue;
                }
                const long long o_el = pix * p.out_ld + p.out_coff + co0;
                // vector path: whole groups of 8 channels (every Cout in the network but the 294-wide heat map and
                // the 1-channel sSE map is a multiple of 8), unit channel stride, 16-byte aligned destination
                // a 32-wide box may only be stored when its columns are all ours: a full chunk, or the chunk that ends
                // the tensor (TMA clips at Cout); a ragged chunk in the middle (n_tile % 32 != 0) takes the direct path
                if (p.tma_store && (nvalid == 32 || co0 + nvalid == p.Cout)) {
                    // ---- bias/act/residual on all 32 columns (columns past Cout are clipped by the TMA store)
                    const long long r_el = pix * p.res_ld + p.res_coff + co0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float* w = v + 8 * g;
                        if (8 * g < nvalid) {                         // nvalid is a multiple of 8 on this path
                            const float4* b4 = reinterpret_cast<const float4*>(p.bias + co0 + 8 * g);
                            epilogue8<ACT>(w, __ldg(b4), __ldg(b4 + 1), p, r_el + 8 * g);
                        }
                    }
                    // ---- the staging buffer of this warp group must have been drained by its previous store
                    // staging buffers: one or two per warp group; with two, the store issued two chunks ago must have
                    // drained (wait_group.read 1), so the store of the previous chunk overlaps this chunk's work
                    const uint32_t sbuf = stage_out + (uint32_t)(half_id * p.out_bufs + (store_i % p.out_bufs)) * 16384u;
                    if (q == 0 && lane == 0) {
                        if (p.out_bufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    ++store_i;
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half_id) : "memory");
                    if (OUT_SPLIT) {
                        // rows of 64 B per plane: [hi plane 8 KB][lo plane 8 KB]
                        const uint32_t rh = sbuf + (uint32_t)row * 64u, rl = rh + 8192u;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint32_t hp[4], lp[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float a0 = v[8 * g + 2 * j], a1 = v[8 * g + 2 * j + 1];
                                const __half2 h2 = __floats2half2_rn(a0, a1);
                                const float2 hf = __half22float2(h2);
                                const __half2 l2 = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
                                hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
                                lp[j] = *reinterpret_cast<const uint32_t*>(&l2);
                            }
                            // 64-byte swizzle (matches the tensor map): 16-byte chunk index ^= (row/2) % 4, so the
                            // 32 rows a warp writes spread over all banks and TMA un-swizzles on the way out
                            const uint32_t slot = (uint32_t)((g ^ (row >> 1)) & 3) * 16u;
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rh + slot), "r"(hp[0]), "r"(hp[1]), "r"(hp[2]), "r"(hp[3]) : "memory");
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rl + slot), "r"(lp[0]), "r"(lp[1]), "r"(lp[2]), "r"(lp[3]) : "memory");
                        }
                    } else {
                        const uint32_t rf = sbuf + (uint32_t)row * 128u;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const uint32_t slot = (uint32_t)((g ^ row) & 7) * 16u;      // 128-byte swizzle: chunk ^= row % 8
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rf + slot), "f"(v[4 * g]), "f"(v[4 * g + 1]), "f"(v[4 * g + 2]), "f"(v[4 * g + 3]) : "memory");
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half_id) : "memory");
                    if (q == 0 && lane == 0) {
                        const int ty0 = (t / tiles_x) * p.bh, tx0 = (t % tiles_x) * p.bw;
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(&tmO_hi), "r"(sbuf), "r"(co0), "r"(tx0), "r"(ty0), "r"(img) : "memory");
                        if (OUT_SPLIT)
                            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                         ::"l"(&tmO_lo), "r"(sbuf + 8192u), "r"(co0), "r"(tx0), "r"(ty0), "r"(img) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    continue;
                }
                if (p.out_cstride != 1 && (nvalid & 7) == 0) {
                    // Strided destination (the detector's ShuffleNetV2 units write straight into channel-shuffled views,
                    // c_stride 2).  A lane owns a pixel, so a direct store instruction would touch 32 different sectors with
                    // 2-4 bytes each (measured: 65-85 us per launch at batch 16 for 10 us of work).  The 32 x 32 block is
                    // transposed through this warp's 4 KB staging slice (XOR-swizzled, conflict-free both ways) and stored
                    // pixel by pixel with lane = channel: 4 sectors per instruction.
                    const long long r_el = pix * p.res_ld + p.res_coff + co0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float* w = v + 8 * g;
                        if (8 * g < nvalid) {
                            const float4* b4 = reinterpret_cast<const float4*>(p.bias + co0 + 8 * g);
                            epilogue8<ACT>(w, __ldg(b4), __ldg(b4 + 1), p, r_el + 8 * g);
                        }
                    }
                    float* xs = reinterpret_cast<float*>(smem_raw + (stage_out - smem_u32(smem_raw))) + (warp - 4) * 1024;
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) xs[lane * 32 + (j ^ lane)] = v[j];
                    __syncwarp();
                    const int pix_lo = (int)(pix & 0xffffffffll), pix_hi = (int)(pix >> 32);
#pragma unroll 4
                    for (int i = 0; i < 32; ++i) {
                        const bool ok_i = __shfl_sync(0xffffffffu, (int)row_ok, i) != 0;
                        const long long pix_i = ((long long)__shfl_sync(0xffffffffu, pix_hi, i) << 32) |
                                                (unsigned int)__shfl_sync(0xffffffffu, pix_lo, i);
                        if (ok_i && lane < nvalid)
                            st1(p.out, OUT_SPLIT ? DT_SPLIT16 : DT_F32, p.out_plane,
                                pix_i * p.out_ld + p.out_coff + (long long)(co0 + lane) * p.out_cstride, xs[i * 32 + (lane ^ i)]);
                    }
                    __syncwarp();
                    continue;
                }
                const bool fast = (nvalid & 7) == 0 && p.out_cstride == 1 && ((p.out_ld | (p.out_coff + co0)) & 7) == 0;
                if (!row_ok) continue;
                if (fast) {
                    const int ng = nvalid >> 3;                       // warp-uniform
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + co0);
                    const long long r_el = pix * p.res_ld + p.res_coff + co0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g < ng) {
                            float* w = v + 8 * g;
                            epilogue8<ACT>(w, __ldg(b4 + 2 * g), __ldg(b4 + 2 * g + 1), p, r_el + 8 * g);
                            if (OUT_SPLIT) {
                                __half* oh = (__half*)p.out + o_el + 8 * g;
                                uint4 hv, lv;
                                uint32_t* hp = reinterpret_cast<uint32_t*>(&hv);
                                uint32_t* lp = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float a0 = w[2 * j], a1 = w[2 * j + 1];
                                    const __half2 h2 = __floats2half2_rn(a0, a1);
                                    const float2 hf = __half22float2(h2);
                                    const __half2 l2 = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
                                    hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
                                    lp[j] = *reinterpret_cast<const uint32_t*>(&l2);
                                }
                                *reinterpret_cast<uint4*>(oh) = hv;
                                *reinterpret_cast<uint4*>(oh + p.out_plane) = lv;
                            } else {
                                float* o = (float*)p.out + o_el + 8 * g;
                                *reinterpret_cast<float4*>(o) = make_float4(w[0], w[1], w[2], w[3]);
                                *reinterpret_cast<float4*>(o + 4) = make_float4(w[4], w[5], w[6], w[7]);
                            }
                        }
                    }
                } else {
                    // ragged tail / unaligned destination: scalar, not unrolled (rare)
                    const long long r_el = pix * p.res_ld + p.res_coff + co0;
#pragma unroll 1
                    for (int j = 0; j < nvalid; ++j) {
                        float f = fmaf(v_at(v, j), p.out_scale, __ldg(p.bias + co0 + j));
                        const float r = p.res ? ld1(p.res, p.res_fmt, p.res_plane, r_el + j) : 0.f;
                        f = p.res_first ? act_t<ACT>(f + r) : act_t<ACT>(f) + r;
                        st1(p.out, OUT_SPLIT ? DT_SPLIT16 : DT_F32, p.out_plane,
                            pix * p.out_ld + p.out_coff + (long long)(co0 + j) * p.out_cstride, f);
                    }
                }
            }
            chunk_ctr += n_chunks;
          }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    }

    if (p.tma_store && warp >= 4 && (warp & 3) == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

static const float* zero_bias() {
    static float* z = nullptr;       // 1024 zeros for bias-free layers (ASPP), per process/device
    if (!z) {
        if (cudaMalloc(&z, 1024 * sizeof(float)) != cudaSuccess) return nullptr;
        cudaMemset(z, 0, 1024 * sizeof(float));
    }
    return z;
}

// Tile width for an Ho x Wo OUTPUT map (0 = no tiling): 128-pixel tiles are bw x (128/bw) blocks of one image.
// Maps whose width divides (or is a multiple of) 128 use row-block tiles (bw = min(W, 128)) that cover the map exactly.
// Any other map (the detector's 48x80 / 24x40 / 12x20, re-targeted @192/@320 exports) gets the bw in {64,32,16,8} with the
// fewest tiles, edge tiles hanging over the right/bottom border: the TMA loads zero-fill and the TMA stores clip what
// lies outside, the direct-store path masks it.  SKPS_TC_ANY_W=0 restores the exact-cover-only rule.
static int tc_pick_bw(int H, int W) {
    static int any_w = -1;
    if (any_w < 0) {
        const char* e = getenv("SKPS_TC_ANY_W");
        any_w = (e && e[0] == '0') ? 0 : 1;
    }
    if (W >= TC_BM && W % TC_BM == 0) return TC_BM;
    if (W < TC_BM && TC_BM % W == 0 && H % (TC_BM / W) == 0) return W;
    if (!any_w) return 0;
    int best = 0, best_tiles = 1 << 30;
    for (int bw = 64; bw >= 8; bw >>= 1) {
        const int bh = TC_BM / bw;
        if (bw > W + 7 || bh > H + 7) continue;                  // keep boxes within ~the map
        const int tiles = ((W + bw - 1) / bw) * ((H + bh - 1) / bh);
        if (tiles < best_tiles) { best_tiles = tiles; best = bw; }
    }
    return best;
}

// Ho x Wo = OUTPUT map: 128-pixel tiles are row blocks of one image (ragged at the border if need be), or whole images
// (Ho*Wo | 128)
bool tc_shape_ok(int H, int W, int Cin, int in_ld, int in_coff) {
    if (W < 8 || (Cin % 8) || (in_ld % 8) || (in_coff % 8)) return false;
    if (H * W < TC_BM) return TC_BM % W == 0 && TC_BM % (H * W) == 0;
    return tc_pick_bw(H, W) != 0;
}

int tc_prepare(TcLayer& L, const TcSetup& s) {
    EncodeTiledFn enc = get_encode();
    SKPS_CHECK(enc, "cuTensorMapEncodeTiled entry point not available");
    const int stride = s.stride > 0 ? s.stride : 1;
    SKPS_CHECK((stride == 1 || stride == 2) && s.H % stride == 0 && s.W % stride == 0, "conv_tc: stride %d on %dx%d", stride,
               s.H, s.W);
    const int Ho = s.H / stride, Wo = s.W / stride;
    SKPS_CHECK(tc_shape_ok(Ho, Wo, s.Cin, s.in_ld, s.in_coff), "conv_tc: unsupported shape %dx%d Cin=%d ld=%d off=%d",
               Ho, Wo, s.Cin, s.in_ld, s.in_coff);
    SKPS_CHECK(s.kh == s.kw && s.pad == s.dil * (s.kh - 1) / 2, "conv_tc: only 'same' square kernels");
    L.k = TcK();
    TcK& k = L.k;
    k.H = Ho; k.W = Wo; k.stride = stride;             // the kernel's H, W are the OUTPUT map
    k.ipt = Ho * Wo < TC_BM ? TC_BM / (Ho * Wo) : 1;
    k.bw = k.ipt > 1 ? Wo : tc_pick_bw(Ho, Wo);
    k.bh = TC_BM / k.bw;
    k.tiles_per_img = k.ipt > 1 ? 1 : ((Ho + k.bh - 1) / k.bh) * ((Wo + k.bw - 1) / k.bw);
    k.taps = s.kh * s.kw; k.kw = s.kw; k.dil = s.dil; k.pad = s.pad;
    k.cchunks = (s.Cin + TC_BK - 1) / TC_BK;
    k.Cout = s.Cout; k.act = s.act; k.out_scale = s.out_scale;
    k.n_tile = s.n_tile; k.n_tiles = s.n_tiles;
    // two pixel tiles per weight-tile load when both accumulators fit one TMEM stage (N <= 128) and two
    // pipeline stages still fit in shared memory: halves the weight traffic from L2
    k.mt = (s.n_tile <= 128 && s.mt_hint != 1) ? 2 : 1;
    {
        // halo-row mode (see the file header): 3x3 / stride 1 / dilation 1 on 64-wide maps, whole 4-row groups, one N tile
        // Off by default: measured on B200 it moves 33 % fewer bytes from L2 (1 152 vs 1 728 KB per 256 pixels) but runs
        // conv2 in the same 0.74 ms - with N = 128 both operands of every MMA come from shared memory at 128 B/clk, which is
        // the SM's whole shared-memory bandwidth, so the TMA fill and the epilogue staging compete with the tensor pipe
        // whichever way the tiles arrive.  SKPS_TC_K3=1 enables it (tests/test_conv_tc_gpu.py does).
        const char* e = getenv("SKPS_TC_K3");
        const int k3_on = (e && e[0] == '1') ? 1 : 0;
        k.k3 = (k3_on && s.kh == 3 && stride == 1 && s.dil == 1 && Wo == 64 && k.bw == 64 && Ho % 4 == 0 && s.Cin % 32 == 0 &&
                s.n_tiles == 1 && k.mt == 2 && k.ipt == 1 && !s.hm_val) ? 1 : 0;
    }
    const size_t stage_bytes = k.k3 ? (size_t)2 * K3_A_PLANE + 6 * (size_t)k.n_tile * 64
                                    : (size_t)k.mt * 2 * (size_t)A_TILE_BYTES + 2 * (size_t)k.n_tile * TC_BK * 2;
    // TMA-store epilogue: full 128-byte lines instead of 16-byte pieces per thread.  Needs a unit-stride,
    // 16-byte aligned destination whose channel count is a multiple of 8 (the swizzle-free box clips at Cout).
    const int oes = s.out_fmt == DT_SPLIT16 ? 2 : 4;
    k.tma_store = (s.out_cstride == 1 && (s.Cout % 8) == 0 && ((size_t)s.out_ld * oes) % 16 == 0 &&
                   ((size_t)s.out_coff * oes) % 16 == 0 && s.tma_store_hint != 1 && !s.hm_val) ? 1 : 0;
    // two staging buffers per epilogue warp group when the pipeline still gets its stages, else one
    const size_t budget = 227 * 1024 - 1024 - 1024;                // minus static smem slack and alignment pad
    const int want_stages = (int)((budget - 2 * 16384) / stage_bytes) > MAX_STAGES ? MAX_STAGES
                                                                                  : (int)((budget - 2 * 16384) / stage_bytes);
    k.out_bufs = (k.tma_store && (budget - 4 * 16384) / stage_bytes >= (size_t)want_stages) ? 2 : 1;
    const size_t out_stage = (k.tma_store || s.out_cstride != 1) ? (size_t)2 * k.out_bufs * 16384 : 0;   // strided outputs transpose through it
    int stages = (int)((budget - out_stage) / stage_bytes);
    k.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    SKPS_CHECK(k.stages >= 2, "conv_tc: tile too large for shared memory");
    L.smem_bytes = (int)(k.stages * stage_bytes + out_stage + 1024 + (s.hm_val ? 2048 : 0));    // + the arg-max exchange area

    // activations: (C, W, H, N) fp16, channel window [in_coff, in_coff+Cin) of rows of in_ld channels
    for (int plane = 0; plane < 2; ++plane) {
        cuuint64_t dims[4] = {(cuuint64_t)s.Cin, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.max_batch};
        cuuint64_t strides[3] = {(cuuint64_t)s.in_ld * 2, (cuuint64_t)s.W * s.in_ld * 2, (cuuint64_t)s.H * s.W * s.in_ld * 2};
        // stride 2: TMA walks W and H with element stride 2; a box of 2*bw x 2*bh source elements lands as bw x bh in smem.
        // small maps: the box spans `ipt` whole images (rows of the A tile = (n, y, x))
        const cuuint32_t box_h = (cuuint32_t)(k.ipt > 1 ? Ho : k.bh);
        cuuint32_t box[4] = {TC_BK, (cuuint32_t)(k.bw * stride), box_h * (cuuint32_t)stride, (cuuint32_t)k.ipt};
        if (k.k3) { box[0] = 32; box[1] = 64; box[2] = K3_ROWS; box[3] = 1; }     // 6 halo rows x 64 pixels x 32 channels
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        void* base = (void*)((__half*)s.in_base + (plane ? s.in_plane : 0) + s.in_coff);
        CUresult r = enc(plane ? &L.a_lo : &L.a_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, k.k3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SKPS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
    }
    // weights: (K_pad, rows) fp16, rows = n_tiles*n_tile (zero rows beyond Cout)
    const int K_pad = k.taps * k.cchunks * TC_BK;
    for (int plane = 0; plane < 2; ++plane) {
        cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)(k.n_tiles * k.n_tile)};
        cuuint64_t strides[1] = {(cuuint64_t)K_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)(k.k3 ? 32 : TC_BK), (cuuint32_t)k.n_tile};
        cuuint32_t estr[2] = {1, 1};
        void* base = (void*)(plane ? s.w_lo : s.w_hi);
        CUresult r = enc(plane ? &L.b_lo : &L.b_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, k.k3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SKPS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
    }
    if (k.tma_store) {
        // output view (C, W, H, N); the staged tile is [row = pixel][32 channels] per plane, 64-byte (float16) or
        // 128-byte (float32) rows written with the matching hardware swizzle (bank-conflict-free)
        for (int plane = 0; plane < (s.out_fmt == DT_SPLIT16 ? 2 : 1); ++plane) {
            cuuint64_t dims[4] = {(cuuint64_t)s.Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)s.max_batch};
            cuuint64_t strides[3] = {(cuuint64_t)s.out_ld * oes, (cuuint64_t)Wo * s.out_ld * oes,
                                     (cuuint64_t)Ho * Wo * s.out_ld * oes};
            cuuint32_t box[4] = {32, (cuuint32_t)k.bw, (cuuint32_t)(k.ipt > 1 ? Ho : k.bh), (cuuint32_t)k.ipt};
            cuuint32_t estr[4] = {1, 1, 1, 1};
            char* base = (char*)s.out + (size_t)s.out_coff * oes + (plane ? (size_t)s.out_plane * 2 : 0);
            CUresult r = enc(plane ? &L.o_lo : &L.o_hi,
                             s.out_fmt == DT_SPLIT16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                             base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             s.out_fmt == DT_SPLIT16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            SKPS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(out) failed: %d", (int)r);
        }
        if (s.out_fmt != DT_SPLIT16) L.o_lo = L.o_hi;
    } else {
        L.o_hi = L.a_hi; L.o_lo = L.a_hi;      // unused placeholders
    }
    k.bias = s.bias ? s.bias : zero_bias();
    SKPS_CHECK(k.bias, "conv_tc: zero-bias allocation failed");
    k.Cin = s.Cin;
    k.out = s.out; k.out_fmt = s.out_fmt; k.out_plane = s.out_plane; k.out_ld = s.out_ld; k.out_coff = s.out_coff;
    k.out_cstride = s.out_cstride;
    k.res = s.res; k.res_fmt = s.res_fmt; k.res_plane = s.res_plane; k.res_ld = s.res_ld; k.res_coff = s.res_coff;
    k.res_first = s.res ? s.res_first : 0;
    k.hm_val = s.hm_val; k.hm_idx = s.hm_idx; k.hm_ld = s.hm_ld;
    if (k.hm_val) {
        SKPS_CHECK(k.ipt == 1 && s.act == ACT_NONE && !s.res && (Ho * Wo) % TC_BM == 0 && k.tiles_per_img * TC_BM == Ho * Wo,
                   "conv_tc: heat-map partials need whole 128-pixel tiles, no activation, no residual");
    }
    return 0;
}

template <int ACT, bool SPLIT>
static int tc_launch_t(const TcLayer& L, const TcK& k, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        SKPS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<ACT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024 - 1024));
        attr_set = true;
    }
    conv_tc_kernel<ACT, SPLIT><<<grid, TC_THREADS, L.smem_bytes, stream>>>(L.a_hi, L.a_lo, L.b_hi, L.b_lo, L.o_hi, L.o_lo, k);
    SKPS_CUDA(cudaGetLastError());
    return 0;
}

int tc_launch(const TcLayer& L, int batch, int img0, int num_sms, cudaStream_t stream) {
    TcK k = L.k;
    k.m_tiles = k.ipt > 1 ? (batch + k.ipt - 1) / k.ipt : batch * k.tiles_per_img;
    k.img0 = img0;
    k.img_end = img0 + batch;
    int total = ((k.m_tiles + k.mt - 1) / k.mt) * k.n_tiles;
    int grid = total < num_sms ? total : num_sms;
    const bool sp = k.out_fmt == DT_SPLIT16;
    switch (k.act) {
        case ACT_NONE: return sp ? tc_launch_t<ACT_NONE, true>(L, k, grid, stream) : tc_launch_t<ACT_NONE, false>(L, k, grid, stream);
        case ACT_RELU: return sp ? tc_launch_t<ACT_RELU, true>(L, k, grid, stream) : tc_launch_t<ACT_RELU, false>(L, k, grid, stream);
        case ACT_HSWISH: return sp ? tc_launch_t<ACT_HSWISH, true>(L, k, grid, stream) : tc_launch_t<ACT_HSWISH, false>(L, k, grid, stream);
        case ACT_SIGMOID: return sp ? tc_launch_t<ACT_SIGMOID, true>(L, k, grid, stream) : tc_launch_t<ACT_SIGMOID, false>(L, k, grid, stream);
        case ACT_SILU: return sp ? tc_launch_t<ACT_SILU, true>(L, k, grid, stream) : tc_launch_t<ACT_SILU, false>(L, k, grid, stream);
        default: break;
    }
    set_error("conv_tc: activation %d not instantiated", k.act);
    return 1;
}

// float32 NHWC -> hi/lo float16 planes (used by the debug entry point and by f32->split conversions)
__global__ void f32_to_split_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo,
                                    long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = src[i];
    __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
}

}  // namespace skps

using namespace skps;

// Debug/unit-test entry: one conv through the tensor-core kernel on host data.
//   x      [host] float32 NHWC (N,H,W,Cin)
//   w_hi/w_lo [host] float16 (rows, K_pad) as packed by plan.py:pack_tc_weights
//   out    [host] float32 NHWC (N,H,W,Cout)
extern "C" SKPS_API int skps_debug_conv_tc2(const float* x, int N, int H, int W, int Cin, const void* w_hi,
                                            const void* w_lo, const float* bias, int Cout, int ksize, int dil, int act,
                                            int n_tile, int n_tiles, float out_scale, const float* residual,
                                            int out_split, float* out, int stride, int res_first, int max_batch);

extern "C" SKPS_API int skps_debug_conv_tc(const float* x, int N, int H, int W, int Cin, const void* w_hi,
                                           const void* w_lo, const float* bias, int Cout, int ksize, int dil, int act,
                                           int n_tile, int n_tiles, float out_scale, const float* residual,
                                           int out_split, float* out) {
    return skps_debug_conv_tc2(x, N, H, W, Cin, w_hi, w_lo, bias, Cout, ksize, dil, act, n_tile, n_tiles, out_scale,
                               residual, out_split, out, 1, 0, N);
}

// Same with a conv stride (1 or 2; out is N x H/stride x W/stride x Cout), the residual order
// (res_first: act(conv + res) instead of act(conv) + res) and a buffer capacity max_batch >= N
// (multi-image tiles write whole tiles: exercises the partial last tile).
extern "C" SKPS_API int skps_debug_conv_tc2(const float* x, int N, int H, int W, int Cin, const void* w_hi,
                                            const void* w_lo, const float* bias, int Cout, int ksize, int dil, int act,
                                            int n_tile, int n_tiles, float out_scale, const float* residual,
                                            int out_split, float* out, int stride, int res_first, int max_batch) {
    SKPS_CHECK(x && w_hi && w_lo && out, "debug_conv_tc: null argument");
    SKPS_CHECK((stride == 1 || stride == 2) && max_batch >= N, "debug_conv_tc: bad stride/max_batch");
    const int Ho = H / stride, Wo = W / stride;
    const long long nin = (long long)N * H * W * Cin, nout = (long long)N * Ho * Wo * Cout;
    const long long cin_cap = (long long)max_batch * H * W * Cin, cout_cap = (long long)max_batch * Ho * Wo * Cout;
    const int cchunks = (Cin + TC_BK - 1) / TC_BK;
    const size_t wbytes = (size_t)n_tiles * n_tile * ksize * ksize * cchunks * TC_BK * 2;
    float *d_x = nullptr, *d_bias = nullptr, *d_out = nullptr, *d_res = nullptr;
    __half *d_split = nullptr, *d_wh = nullptr, *d_wl = nullptr, *d_osplit = nullptr;
    SKPS_CUDA(cudaMalloc(&d_x, nin * 4));
    SKPS_CUDA(cudaMalloc(&d_split, cin_cap * 4));
    SKPS_CUDA(cudaMemset(d_split, 0, cin_cap * 4));
    SKPS_CUDA(cudaMalloc(&d_wh, wbytes));
    SKPS_CUDA(cudaMalloc(&d_wl, wbytes));
    SKPS_CUDA(cudaMalloc(&d_out, cout_cap * 4));
    SKPS_CUDA(cudaMalloc(&d_osplit, cout_cap * 4));
    SKPS_CUDA(cudaMemcpy(d_x, x, nin * 4, cudaMemcpyHostToDevice));
    SKPS_CUDA(cudaMemcpy(d_wh, w_hi, wbytes, cudaMemcpyHostToDevice));
    SKPS_CUDA(cudaMemcpy(d_wl, w_lo, wbytes, cudaMemcpyHostToDevice));
    if (bias) {
        SKPS_CUDA(cudaMalloc(&d_bias, Cout * 4));
        SKPS_CUDA(cudaMemcpy(d_bias, bias, Cout * 4, cudaMemcpyHostToDevice));
    }
    if (residual) {
        SKPS_CUDA(cudaMalloc(&d_res, nout * 4));
        SKPS_CUDA(cudaMemcpy(d_res, residual, nout * 4, cudaMemcpyHostToDevice));
    }
    f32_to_split_kernel<<<(unsigned)((nin + 255) / 256), 256>>>(d_x, d_split, d_split + cin_cap, nin);
    SKPS_CUDA(cudaGetLastError());
    TcSetup s = {};
    s.H = H; s.W = W; s.Cin = Cin; s.in_ld = Cin; s.in_coff = 0; s.max_batch = max_batch;
    s.in_base = d_split; s.in_plane = cin_cap;
    s.kh = s.kw = ksize; s.dil = dil; s.pad = dil * (ksize - 1) / 2; s.stride = stride; s.res_first = res_first;
    s.Cout = Cout; s.act = act; s.n_tile = n_tile; s.n_tiles = n_tiles; s.out_scale = out_scale;
    s.w_hi = d_wh; s.w_lo = d_wl; s.bias = d_bias;
    s.out = out_split ? (void*)d_osplit : (void*)d_out; s.out_fmt = out_split ? DT_SPLIT16 : DT_F32;
    s.out_plane = cout_cap; s.out_ld = Cout; s.out_coff = 0; s.out_cstride = 1;
    s.res = d_res; s.res_fmt = DT_F32; s.res_plane = 0; s.res_ld = Cout; s.res_coff = 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (tct_applicable(s)) {                        // the engine routes such layers to the transposed kernel (conv_tct.cu)
        float* d_zero = nullptr;
        if (!s.bias) {
            SKPS_CUDA(cudaMalloc(&d_zero, 1024 * 4));
            SKPS_CUDA(cudaMemset(d_zero, 0, 1024 * 4));
            s.bias = d_zero;
        }
        TctLayer T;
        if (tct_prepare(T, s)) return 1;
        if (tct_launch(T, N, 0, sms, 0)) return 1;
        SKPS_CUDA(cudaDeviceSynchronize());
        if (d_zero) cudaFree(d_zero);
    } else {
    TcLayer L;
    if (tc_prepare(L, s)) return 1;
    if (tc_launch(L, N, 0, sms, 0)) return 1;
    }
    SKPS_CUDA(cudaDeviceSynchronize());
    if (out_split) {
        // recombine hi+lo on the host side of the test
        __half* tmp = (__half*)malloc(nout * 4);
        SKPS_CUDA(cudaMemcpy(tmp, d_osplit, nout * 2, cudaMemcpyDeviceToHost));
        SKPS_CUDA(cudaMemcpy(tmp + nout, d_osplit + cout_cap, nout * 2, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < nout; ++i) out[i] = __half2float(tmp[i]) + __half2float(tmp[nout + i]);
        free(tmp);
    } else {
        SKPS_CUDA(cudaMemcpy(out, d_out, nout * 4, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_x); cudaFree(d_split); cudaFree(d_wh); cudaFree(d_wl); cudaFree(d_out); cudaFree(d_osplit);
    if (d_bias) cudaFree(d_bias);
    if (d_res) cudaFree(d_res);
    return 0;
}
