// Host/device structs of the transposed tcgen05 convolution (conv_tct.cu): channels on the TMEM lanes, 256 pixels as N.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc.h"

namespace skps {

struct TctK {                    // kernel parameters
    int W, bh, tiles_per_img, m_tiles, img0;    // tile = bh whole rows = 256 pixels
    int taps, kw, dil, pad, cchunks, Cin, Cout, act;
    int k3;                      // halo-row stages: (kx, 32-channel half) = one (bh+2)-row box + the three ky weight tiles
    int xb;                      // k3: bytes of one plane of the activation box
    float out_scale;             // exact power of two undoing the weight pre-scale
    const float* bias;
};

struct TctLayer {
    CUtensorMap x_hi, x_lo, w_hi, w_lo, o_hi, o_lo;
    TctK k;
    int smem_bytes = 0;
    bool valid = false;
};

bool tct_applicable(const TcSetup& s);          // s.H, s.W: the (stride-1) map; split-fp16 contiguous output, no residual
int tct_prepare(TctLayer& L, const TcSetup& s);
int tct_launch(const TctLayer& L, int batch, int img0, int num_sms, cudaStream_t stream);

}  // namespace skps
