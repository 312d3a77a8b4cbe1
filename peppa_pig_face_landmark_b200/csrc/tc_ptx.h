// PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, conv_xf.cu): mbarrier, TMA, tcgen05.mma / commit / ld,
// shared-memory matrix descriptors.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "common.h"

namespace skps {

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread sleeps in hardware until the phase completes (or the hint expires)
// instead of re-issuing try_wait back to back.  Without the hint the spin loops of the waiting roles took 37 % of all issued
// instructions of the fused kernel (profiles/r2_ncu_xf_up2_v1.txt) and starved the transform warps.
constexpr uint32_t MBAR_SUSPEND_NS = 0x989680u;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
}
// Bounded wait for kernels under development: traps (the launch fails with an error) instead of hanging the GPU when a
// pipeline bug leaves a barrier incomplete for ~2 s.
__device__ __forceinline__ void mbar_wait_g(uint32_t bar, uint32_t parity) {
    uint64_t t0 = 0;
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS) : "memory");
        if (ok) return;
        if ((it & 15u) == 15u) {
            uint64_t t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate, single CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
    d |= (uint64_t)1 << 46;                         // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

// Same for 64-byte swizzle: rows of 64 B (32 fp16 along K), 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                         // SWIZZLE_64B
    return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
    if (ACT == ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == ACT_HSWISH) return v * hsigmoid_f(v);
    if (ACT == ACT_SIGMOID) return sigmoid_f(v);
    if (ACT == ACT_SILU) return v * sigmoid_f(v);
    if (ACT == ACT_HSIGMOID) return hsigmoid_f(v);
    return v;
}
// dynamic index into a register array without forcing it to local memory
__device__ __forceinline__ float v_at(const float* v, int j) {
    float r = v[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) r = (j == i) ? v[i] : r;
    return r;
}


// bias + (residual) + activation on 8 consecutive output channels of one pixel.
// res_first = 0: act(acc*scale + bias) + res   (MobileNetV3 linear bottlenecks)
// res_first = 1: act(acc*scale + bias + res)   (ResNet / HRNet blocks: conv-bn, add, relu)
template <int ACT, class PK>
__device__ __forceinline__ void epilogue8(float* w, const float4 b0, const float4 b1, const PK& p, long long r_el) {
    w[0] = fmaf(w[0], p.out_scale, b0.x); w[1] = fmaf(w[1], p.out_scale, b0.y);
    w[2] = fmaf(w[2], p.out_scale, b0.z); w[3] = fmaf(w[3], p.out_scale, b0.w);
    w[4] = fmaf(w[4], p.out_scale, b1.x); w[5] = fmaf(w[5], p.out_scale, b1.y);
    w[6] = fmaf(w[6], p.out_scale, b1.z); w[7] = fmaf(w[7], p.out_scale, b1.w);
    if (p.res) {
        const float4 r0 = ld4(p.res, p.res_fmt, p.res_plane, r_el);
        const float4 r1 = ld4(p.res, p.res_fmt, p.res_plane, r_el + 4);
        const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
        if (p.res_first) {
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = act_t<ACT>(w[j] + r[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = act_t<ACT>(w[j]) + r[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = act_t<ACT>(w[j]);
    }
}


}  // namespace skps
