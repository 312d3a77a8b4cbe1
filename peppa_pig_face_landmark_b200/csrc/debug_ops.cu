// Unit-test entry points for the small fused kernels that otherwise only run inside a whole network:
// the squeeze-excite gate (se_fc_kernel) and the heat-map decode (hm_decode_kernel).  Host float32 in/out.
#include <string.h>

#include "../../include/skps_b200.h"
#include "common.h"

using namespace skps;

namespace {
struct DBuf {
    void* p = nullptr;
    ~DBuf() { if (p) cudaFree(p); }
    int put(const void* src, size_t bytes) {
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return 1;
        if (src && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
        return 0;
    }
};
TView mk(void* base, int C, int H, int W, int ld = 0) {
    TView t;
    memset(&t, 0, sizeof(t));
    t.base = base; t.ld = ld ? ld : C; t.c_off = 0; t.c_stride = 1; t.C = C; t.H = H; t.W = W;
    t.sample = (long long)t.ld * H * W; t.fmt = DT_F32; t.plane = 0;
    return t;
}
}  // namespace

// gate[n][c] = act2(W2 * act1(W1 * (sum_tiles part[n][t][c] / hw) + b1) + b2): the squeeze-excite branch of a
// MobileNetV3 block (kps_student.onnx .../se/conv_reduce, conv_expand; timm SqueezeExcite).  w1t [C][Cr], w2t [Cr][C].
extern "C" SKPS_API int skps_debug_se_fc(const float* part, int N, int tiles, int C, const float* w1t, const float* b1,
                                         const float* w2t, const float* b2, int Cr, int act1, int act2, int hw,
                                         float* gate) {
    SKPS_CHECK(part && w1t && b1 && w2t && b2 && gate, "debug_se_fc: null argument");
    DBuf dp, dw1, db, dw2, dg;
    std::string bb((size_t)(Cr + C) * 4, '\0');
    memcpy(&bb[0], b1, (size_t)Cr * 4);
    memcpy(&bb[(size_t)Cr * 4], b2, (size_t)C * 4);
    SKPS_CHECK(!dp.put(part, (size_t)N * tiles * C * 4) && !dw1.put(w1t, (size_t)C * Cr * 4) && !db.put(bb.data(), bb.size()) &&
               !dw2.put(w2t, (size_t)C * Cr * 4) && !dg.put(nullptr, (size_t)N * C * 4), "debug_se_fc: cudaMalloc/copy failed");
    TView pv = mk(dp.p, C, tiles, 1), gv = mk(dg.p, C, 1, 1);
    if (launch_se_fc(pv, gv, (const float*)dw1.p, (const float*)db.p, (const float*)dw2.p, (const float*)db.p + Cr, Cr, act1,
                     act2, hw, N, 0))
        return 1;
    SKPS_CUDA(cudaDeviceSynchronize());
    SKPS_CUDA(cudaMemcpy(gate, dg.p, (size_t)N * C * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// Heat-map decode (model.py:511-554 postp): per landmark arg-max over H*W (first index on ties), score = the maximum,
// (x, y) = (argmax position + offset) / W.  hm (N,H,W,ld) holds npts score maps [and 2*npts offset maps when feat is
// null]; with feat (N,H,W,K) the offsets are w_off (2*npts,K) . feat[argmax] + b_off.
extern "C" SKPS_API int skps_debug_hm_decode(const float* hm, int N, int H, int W, int ld, int npts, const float* feat, int K,
                                             const float* w_off, const float* b_off, float* xy, float* score) {
    SKPS_CHECK(hm && xy && score, "debug_hm_decode: null argument");
    DBuf dh, df, dw, db, dx, ds;
    SKPS_CHECK(!dh.put(hm, (size_t)N * H * W * ld * 4) && !dx.put(nullptr, (size_t)N * 2 * npts * 4) &&
               !ds.put(nullptr, (size_t)N * npts * 4), "debug_hm_decode: cudaMalloc/copy failed");
    TView hv = mk(dh.p, feat ? npts : 3 * npts, H, W, ld), fv;
    memset(&fv, 0, sizeof(fv));
    if (feat) {
        SKPS_CHECK(!df.put(feat, (size_t)N * H * W * K * 4) && !dw.put(w_off, (size_t)2 * npts * K * 4) &&
                   !db.put(b_off, (size_t)2 * npts * 4), "debug_hm_decode: cudaMalloc/copy failed");
        fv = mk(df.p, K, H, W);
    }
    TView xv = mk(dx.p, 2 * npts, 1, 1), sv = mk(ds.p, npts, 1, 1);
    if (launch_hm_decode(hv, fv, (const float*)dw.p, (const float*)db.p, xv, sv, npts, N, 0)) return 1;
    SKPS_CUDA(cudaDeviceSynchronize());
    SKPS_CUDA(cudaMemcpy(xy, dx.p, (size_t)N * 2 * npts * 4, cudaMemcpyDeviceToHost));
    SKPS_CUDA(cudaMemcpy(score, ds.p, (size_t)N * npts * 4, cudaMemcpyDeviceToHost));
    return 0;
}
