"""Unit tests of the tcgen05 convolution kernel (csrc/conv_tc.cu) through skps_debug_conv_tc,
against torch.nn.functional.conv2d in float32 on the CPU.  Shapes are the landmark network's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(N, H, W, Cin, Cout, k, dil, act, with_bias=True, with_res=False, out_split=False, seed=0, stride=1,
         res_first=False, max_batch=None):
    import torch
    import torch.nn.functional as F
    from peppa_pig_face_landmark_b200 import plan as P, runtime as rt
    lib = rt.load_library()
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32) * 2
    w = (rng.standard_normal((Cout, k, k, Cin)) / np.sqrt(k * k * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) if with_bias else None
    Ho, Wo = H // stride, W // stride
    res = rng.standard_normal((N, Ho, Wo, Cout)).astype(np.float32) if with_res else None
    n_tile, n_tiles = P.tc_tiling(Cout)
    hi, lo, out_scale = P.pack_tc_weights(w, n_tile, n_tiles)
    hi, lo = np.ascontiguousarray(hi), np.ascontiguousarray(lo)
    out = np.empty((N, Ho, Wo, Cout), np.float32)
    rt.check(lib.skps_debug_conv_tc2(x.ctypes.data, N, H, W, Cin, hi.ctypes.data, lo.ctypes.data,
                                     b.ctypes.data if b is not None else None, Cout, k, dil, act, n_tile, n_tiles,
                                     out_scale, res.ctypes.data if res is not None else None, 1 if out_split else 0,
                                     out.ctypes.data, stride, 1 if res_first else 0, max_batch or N))
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w).permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(xt, wt, torch.from_numpy(b) if b is not None else None, stride=stride, padding=dil * (k - 1) // 2,
                 dilation=dil)
    if res is not None and res_first:
        y = y + torch.from_numpy(res).permute(0, 3, 1, 2)
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = y * torch.clamp(y * np.float32(1 / 6) + 0.5, 0, 1)
    y = y.permute(0, 2, 3, 1).numpy()
    if res is not None and not res_first:
        y = y + res
    err = np.abs(out - y).max() / (np.abs(y).max() + 1e-9)
    print('conv_tc', (N, H, W, Cin, Cout, k, dil, act, stride), 'rel err %.3e' % err)
    return err


@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout, k, dil, act
    (2, 64, 64, 128, 128, 3, 1, 1),      # decoder/upsampler2/conv2: 40.7 % of the MACs
    (2, 64, 64, 128, 294, 1, 1, 0),      # hm head (two N tiles, ragged Cout)
    (2, 64, 64, 280, 128, 1, 1, 1),      # K not a multiple of 64
    (3, 32, 32, 296, 256, 1, 1, 1),
    (4, 16, 16, 160, 64, 3, 2, 0),       # ASPP dilated 3x3
    (4, 16, 16, 160, 64, 3, 4, 0),
    (2, 16, 16, 160, 960, 1, 1, 2),      # four N tiles, h-swish
    (2, 16, 16, 960, 160, 1, 1, 0),
    (1, 128, 128, 16, 64, 1, 1, 1),      # W = 128: one image row per tile, tiny K
    (2, 64, 64, 24, 72, 1, 1, 1),
    (2, 32, 32, 40, 120, 1, 1, 1),
    (5, 32, 32, 120, 40, 1, 1, 0),
])
def test_conv_tc_matches_fp32(cfg):
    err = _run(*cfg)
    assert err < 1e-5, (cfg, err)       # hi/lo split ~2^-22 per product + the tensor core's fp32 accumulate over K/16*3 steps


@pytest.mark.parametrize("cfg", [
    # halo-row mode of conv_tc (3x3, 64-wide maps, Cin % 32 == 0): the three ky taps address one staged 6-row box in place
    (3, 8, 64, 32, 24, 3, 1, 0),         # two 4-row groups per image, one 32-channel half-chunk, ragged Cout
    (1, 64, 64, 96, 128, 3, 1, 1),       # three half-chunks
    (2, 12, 64, 64, 64, 3, 1, 1),        # H = 12: three groups, image borders inside the batch
    (2, 64, 64, 128, 128, 3, 1, 2),
])
def test_conv_tc_halo_row_mode(cfg, monkeypatch):
    monkeypatch.setenv("SKPS_TC_K3", "1")              # opt-in mode (read by tc_prepare at every layer setup)
    err = _run(*cfg, with_res=True, out_split=True)
    assert err < 1e-5, (cfg, err)


@pytest.mark.parametrize("cfg", [
    # transposed kernel (csrc/conv_tct.cu): channels on the TMEM lanes, 256 pixels as N; split-fp16 output, no residual
    (2, 64, 64, 128, 128, 3, 1, 1),      # decoder conv2 (halo-row stages: 6-row boxes)
    (2, 32, 32, 96, 128, 3, 1, 1),       # halo-row stages with 10-row boxes, three 32-channel halves
    (1, 32, 32, 64, 96, 3, 2, 0),        # dilated, 8-row tiles, Cout < 128 (the last lane quarter is clipped by the store)
    (3, 8, 128, 72, 104, 3, 1, 2),       # two-row tiles, K tail (72 channels), ragged Cout
    (2, 16, 256, 64, 128, 5, 1, 1),      # 5x5, one row per tile
    (3, 16, 16, 160, 128, 3, 2, 1),      # 16-wide map: a tile is a whole image, store boxes of 2 rows x 16 pixels
])
def test_conv_tct_transposed_kernel(cfg):
    err = _run(*cfg, out_split=True)
    assert err < 1e-5, (cfg, err)


@pytest.mark.parametrize("cfg", [(2, 64, 64, 128, 128, 3, 1, 1), (2, 32, 32, 96, 128, 3, 1, 1)])
def test_conv_tct_halo_row_stages(cfg, monkeypatch):
    monkeypatch.setenv("SKPS_TCT_K3", "1")             # opt-in stage layout (read by tct_prepare at every layer setup)
    err = _run(*cfg, out_split=True)
    assert err < 1e-5, (cfg, err)


def test_conv_tc_residual_and_split_output():
    assert _run(2, 32, 32, 120, 40, 1, 1, 0, with_res=True) < 1e-5
    assert _run(2, 64, 64, 128, 128, 3, 1, 1, out_split=True) < 1e-5
    assert _run(2, 16, 16, 672, 112, 1, 1, 0, with_bias=False, with_res=True, out_split=True) < 1e-5


def test_conv_tc_residual_before_activation():
    """conv-bn, += shortcut, relu of the Teacher's HRNet blocks (timm BasicBlock/Bottleneck; model.py:302-345)."""
    assert _run(2, 64, 64, 24, 24, 3, 1, 1, with_res=True, res_first=True, out_split=True) < 1e-5
    assert _run(2, 16, 16, 72, 72, 3, 1, 1, with_res=True, res_first=True) < 1e-5
    assert _run(2, 64, 64, 64, 256, 1, 1, 1, with_res=True, res_first=True, out_split=True) < 1e-5


@pytest.mark.parametrize("cfg", [
    # N, H, W, Cin, Cout, k, dil, act  -- stride 2 (TMA element strides), HRNet stem / transition / fuse convs
    (2, 128, 128, 64, 64, 3, 1, 1),      # stem conv2: 64 -> 64, 128^2 -> 64^2
    (2, 64, 64, 256, 40, 3, 1, 1),       # transition1 new branch (36 padded to 40)
    (3, 64, 64, 24, 24, 3, 1, 1),        # fuse 18 -> 18 (padded to 24)
    (2, 32, 32, 40, 72, 3, 1, 0),
    (4, 16, 16, 72, 144, 3, 1, 0),       # output 8x8: two images per tile
    (3, 16, 16, 72, 144, 3, 1, 1),       # ... with a partial last tile
])
def test_conv_tc_stride2(cfg):
    err = _run(*cfg, stride=2)
    assert err < 1e-5, (cfg, err)


def test_conv_tc_small_maps_share_a_tile():
    """8x8 maps (HRNet branch 4): one 128-row tile holds two images."""
    assert _run(4, 8, 8, 144, 144, 3, 1, 1, with_res=True, res_first=True, out_split=True) < 1e-5
    assert _run(5, 8, 8, 144, 144, 3, 1, 1) < 1e-5                      # odd batch: partial tile, TMA store clips
    assert _run(3, 8, 8, 144, 72, 1, 1, 0, max_batch=8) < 1e-5         # spare capacity: the partial tile lands in unused slots
    assert _run(3, 8, 8, 144, 24, 1, 1, 0, out_split=True, max_batch=4) < 1e-5


def _run_mma(N, H, W, C, act, with_res=False, res_first=False, out_split=True, seed=0):
    """csrc/conv_mma.cu (few-channel 3x3, halo tile + mma.sync) against torch conv2d in float32."""
    import torch
    import torch.nn.functional as F
    from peppa_pig_face_landmark_b200 import plan as P, runtime as rt
    lib = rt.load_library()
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, H, W, C)).astype(np.float32) * 2
    w = (rng.standard_normal((C, 3, 3, C)) / np.sqrt(9 * C)).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32)
    res = rng.standard_normal((N, H, W, C)).astype(np.float32) if with_res else None
    packed, out_scale = P.pack_mma_weights(w)
    out = np.empty((N, H, W, C), np.float32)
    rt.check(lib.skps_debug_conv_mma(x.ctypes.data, N, H, W, C, packed.ctypes.data, b.ctypes.data, act, out_scale,
                                     res.ctypes.data if res is not None else None, 1 if res_first else 0,
                                     1 if out_split else 0, out.ctypes.data))
    y = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(w).permute(0, 3, 1, 2).contiguous(),
                 torch.from_numpy(b), padding=1)
    if res is not None and res_first:
        y = y + torch.from_numpy(res).permute(0, 3, 1, 2)
    if act == 1:
        y = torch.relu(y)
    y = y.permute(0, 2, 3, 1).numpy()
    if res is not None and not res_first:
        y = y + res
    err = np.abs(out - y).max() / (np.abs(y).max() + 1e-9)
    print('conv_mma', (N, H, W, C, act, with_res, res_first, out_split), 'rel err %.3e' % err)
    return err


def test_conv_mma_small_channel_3x3():
    """HRNet branch convs of the Teacher: 18->24 and 36->40 padded channels, 64x64 / 32x32 maps, odd sizes too."""
    assert _run_mma(3, 64, 64, 24, 1) < 1e-5
    assert _run_mma(2, 64, 64, 24, 1, with_res=True, res_first=True) < 1e-5            # BasicBlock conv2: += shortcut, relu
    assert _run_mma(3, 32, 32, 40, 1, with_res=True, res_first=True, out_split=False) < 1e-5
    assert _run_mma(2, 32, 32, 40, 0, with_res=True, res_first=False) < 1e-5
    assert _run_mma(1, 24, 40, 24, 0) < 1e-5                                             # partial tiles in both directions


@pytest.mark.parametrize("x_scale", [1e-3, 1.0, 1e3])
def test_conv_tc_elementwise_error_bound_small_and_large_inputs(x_scale):
    """Range check of the fp16 hi/lo operand format: inputs of magnitude 1e-5 .. 1e+3 and an
    ELEMENT-wise bound instead of a max-norm one:
        |out - ref| <= 2^-19 * sum_k |x_k| |w_k|  +  2^-23 * sum_k |w_k|
    First term: the backward-error form of a dot product with ~2^-22 relative operand error and fp32 accumulation.
    Second term: the format's ABSOLUTE floor - below |x| ~ 2^-3 the lo plane (x - fp16(x)) is a float16 subnormal, so an
    operand is only good to 2^-25 absolute; it is what bounds tiny activations (measured 188x over the relative term alone
    at |x| ~ 1e-3).  The landmark networks' activations are O(1), where the first term dominates."""
    import torch
    import torch.nn.functional as F
    from peppa_pig_face_landmark_b200 import plan as P, runtime as rt
    lib = rt.load_library()
    rng = np.random.default_rng(17)
    N, H, W, Cin, Cout, k = 2, 32, 32, 120, 40, 3
    x = (rng.standard_normal((N, H, W, Cin)) * x_scale).astype(np.float32)
    x[0, :4] *= 1e-2                                      # a region two more decades down
    w = (rng.standard_normal((Cout, k, k, Cin)) / np.sqrt(k * k * Cin)).astype(np.float32)
    n_tile, n_tiles = P.tc_tiling(Cout)
    hi, lo, out_scale = P.pack_tc_weights(w, n_tile, n_tiles)
    hi, lo = np.ascontiguousarray(hi), np.ascontiguousarray(lo)
    out = np.empty((N, H, W, Cout), np.float32)
    rt.check(lib.skps_debug_conv_tc2(x.ctypes.data, N, H, W, Cin, hi.ctypes.data, lo.ctypes.data, None, Cout, k, 1, 0,
                                     n_tile, n_tiles, out_scale, None, 0, out.ctypes.data, 1, 0, N))
    xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
    wt = torch.from_numpy(w).permute(0, 3, 1, 2).contiguous().double()
    ref = F.conv2d(xt, wt, padding=1).permute(0, 2, 3, 1).numpy()
    mag = F.conv2d(xt.abs(), wt.abs(), padding=1).permute(0, 2, 3, 1).numpy()
    wsum = F.conv2d(torch.ones_like(xt), wt.abs(), padding=1).permute(0, 2, 3, 1).numpy()
    err = np.abs(out - ref)
    bound = 2.0 ** -19 * mag + 2.0 ** -23 * wsum
    worst = (err / bound).max()
    print("conv_tc element-wise error / bound at |x|~%g: %.3f" % (x_scale, worst))
    assert worst <= 1.0


@pytest.mark.parametrize("shape", [(3, 64, 64, 128, 104), (2, 32, 32, 128, 98), (5, 16, 16, 64, 24), (1, 8, 32, 72, 128)])
def test_conv_hm_transposed_head_matches_fp64_argmax(shape):
    """csrc/conv_hm.cu: score maps on the TMEM lanes, per-tile (max, first arg-max) from a per-thread scan.  Against an fp64
    conv: the per-tile maximum within fp32 noise, and the reported pixel must BE a maximum of its tile (its fp64 score within
    noise of the tile's fp64 maximum); exact ties (planted duplicates) must resolve to the first pixel."""
    from peppa_pig_face_landmark_b200 import plan as P, runtime as rt
    lib = rt.load_library()
    N, H, W, Cin, Cout = shape
    rng = np.random.default_rng(5)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    x[0, 1, 3] = x[0, 0, 5]                                  # duplicate pixels: equal scores in every map, first must win
    x[0, 3 % H, 7] = x[0, 0, 5]
    w = (rng.standard_normal((Cout, 1, 1, Cin)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    n_tile, n_tiles = P.tc_tiling(Cout)
    assert n_tiles == 1
    hi, lo, out_scale = P.pack_tc_weights(w, n_tile, n_tiles)
    hi, lo = np.ascontiguousarray(hi), np.ascontiguousarray(lo)
    tiles = H * W // 256
    val = np.zeros((N, tiles, 128), np.float32)
    idx = np.zeros((N, tiles, 128), np.int32)
    rt.check(lib.skps_debug_conv_hm(x.ctypes.data, N, H, W, Cin, hi.ctypes.data, lo.ctypes.data, b.ctypes.data, Cout, n_tile,
                                    out_scale, val.ctypes.data, idx.ctypes.data))
    ref = (x.reshape(N, H * W, Cin).astype(np.float64) @ w.reshape(Cout, Cin).astype(np.float64).T + b).reshape(N, tiles, 256, Cout)
    rmax = ref.max(axis=2)                                   # [N][tiles][Cout]
    assert np.abs(val[..., :Cout] - rmax).max() < 2e-5 * (np.abs(ref).max() + 1)
    loc = idx[..., :Cout] - (np.arange(tiles) * 256)[None, :, None]
    assert loc.min() >= 0 and loc.max() < 256
    picked = np.take_along_axis(ref, loc[:, :, None, :], axis=2)[:, :, 0, :]
    assert np.abs(picked - rmax).max() < 2e-5 * (np.abs(ref).max() + 1)
    # the duplicated pixel: wherever it is the tile maximum, the first copy (pixel 5 of tile 0) must be reported
    dup = np.isclose(ref[0, 0, 5], rmax[0, 0], rtol=0, atol=1e-12)
    assert (loc[0, 0][dup] == 5).all()
