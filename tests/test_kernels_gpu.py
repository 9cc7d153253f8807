"""GPU unit parity of the network-half kernels (csrc/conv_simt.cu, conv_tc.cu, glue.cu) against a
plain torch fp32 reference of the same op (tests/emul_ops.py restates each op with torch)."""
import numpy as np
import pytest
import torch

from oracle.net_ops import TorchEmulOps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from danet_b200.plan import CudaOps
    return CudaOps(DEV)


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, wsets, relu, residual
    (2, 56, 56, 48, 48, 3, 1, 1, 1, 1),
    (2, 28, 28, 96, 96, 3, 1, 1, 1, 1),
    (3, 14, 14, 192, 192, 3, 1, 1, 0, 0),
    (2, 7, 7, 384, 384, 3, 1, 1, 1, 1),
    (2, 56, 56, 64, 256, 1, 1, 1, 1, 1),
    (2, 56, 56, 256, 64, 1, 1, 1, 1, 0),
    (2, 56, 56, 48, 96, 3, 2, 1, 0, 0),
    (1, 224, 224, 4, 64, 3, 2, 1, 1, 0),
    (2, 56, 56, 64, 64, 7, 2, 1, 1, 0),
    (48, 56, 56, 48, 24, 3, 1, 24, 0, 0),
    (48, 4, 4, 256, 128, 3, 2, 24, 1, 0),
    (48, 4, 4, 256, 128, 1, 2, 24, 0, 0),
    (2, 56, 56, 48, 92, 3, 1, 1, 0, 0),
    (2, 56, 56, 76, 64, 1, 1, 1, 1, 0),
    (5, 13, 9, 20, 36, 3, 1, 1, 1, 1),          # ragged: nothing divides the tile sizes
    (2, 14, 14, 64, 64, 3, 1, 1, 1, 1),
    (2, 28, 28, 64, 128, 3, 2, 1, 1, 0),
    (2, 28, 28, 64, 128, 1, 2, 1, 0, 0),
]


def _conv_case(case, algo, tol):
    N, H, W, Cin, Cout, k, s, G, relu, has_res = case
    ops, ref = _ops(), TorchEmulOps()
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    d = dict(N=N, H=H, W=W, Cin=Cin, Cout=Cout, ksize=k, stride=s, pad=k // 2, wsets=G, relu=relu)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    x = torch.randn(N, H, W, Cin, generator=g)
    w = torch.randn(G, k * k * Cin, Cout, generator=g) * (1.0 / (k * k * Cin)) ** 0.5
    b = torch.randn(G, Cout, generator=g) * 0.1
    res = torch.randn(N, Ho, Wo, Cout, generator=g) if has_res else None
    y_ref = torch.empty(N, Ho, Wo, Cout)
    ref.conv2d(d, 0, x, w, b, res, y_ref)
    xc, wc, bc = x.to(DEV), w.to(DEV), b.to(DEV)
    rc = res.to(DEV) if has_res else None
    y = torch.full((N, Ho, Wo, Cout), float("nan"), device=DEV)
    if algo == 1:
        if not ops.conv_tc_supported(d):
            pytest.skip("shape not on the tcgen05 path")
        wc = ops.conv_tc_pack(d, wc)
    ops.conv2d(d, algo, xc, wc, bc, rc, y)
    torch.cuda.synchronize()
    err = (y.cpu() - y_ref).abs().max().item()
    assert err < tol, (case, err)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt(case):
    _conv_case(case, 0, 2e-5)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tc(case):
    # fp16 operands (10-bit mantissa, RN from fp32), fp32 accumulate
    _conv_case(case, 1, 8e-3)


def test_conv_rejects_bad_arguments():
    ops = _ops()
    d = dict(N=1, H=8, W=8, Cin=3, Cout=8, ksize=3, stride=1, pad=1, wsets=1, relu=0)
    x = torch.zeros(1, 8, 8, 3, device=DEV)
    with pytest.raises(RuntimeError):
        ops.conv2d(d, 0, x, x, None, None, x)        # Cin % 4 != 0


def test_glue_kernels_vs_torch():
    ops, ref = _ops(), TorchEmulOps()
    g = torch.Generator().manual_seed(7)
    B, S, C = 3, 56, 48
    # fuse_sum with upsampling
    t0, t1, t2 = torch.randn(B, 28, 28, C, generator=g), torch.randn(B, 14, 14, C, generator=g), torch.randn(B, 7, 7, C, generator=g)
    y_ref = torch.empty(B, 28, 28, C); ref.fuse_sum([t0, t1, t2], [1, 2, 4], True, y_ref)
    y = torch.empty(B, 28, 28, C, device=DEV); ops.fuse_sum([t0.to(DEV), t1.to(DEV), t2.to(DEV)], [1, 2, 4], True, y)
    assert torch.equal(y.cpu(), y_ref)
    # maxpool / avgpool / linear / nchw->nhwc
    x = torch.randn(B, 28, 28, 64, generator=g)
    y_ref = torch.empty(B, 14, 14, 64); ref.maxpool(x, y_ref)
    y = torch.empty(B, 14, 14, 64, device=DEV); ops.maxpool(x.to(DEV), y)
    assert torch.equal(y.cpu(), y_ref)
    x = torch.randn(B, 2, 2, 512, generator=g)
    y_ref = torch.empty(B, 512); ref.avgpool(x, y_ref)
    y = torch.empty(B, 512, device=DEV); ops.avgpool(x.to(DEV), y)
    assert (y.cpu() - y_ref).abs().max() < 1e-6
    w, b, add = torch.randn(13, 512, generator=g), torch.randn(13, generator=g), torch.randn(13, generator=g)
    o_ref = torch.empty(B, 13); ref.linear(y_ref, w, b, add, o_ref)
    o = torch.empty(B, 13, device=DEV); ops.linear(y, w.to(DEV), b.to(DEV), add.to(DEV), o)
    assert (o.cpu() - o_ref).abs().max() < 1e-4
    img = torch.randn(B, 3, 20, 20, generator=g)
    n_ref = torch.empty(B, 20, 20, 4); ref.nchw_to_nhwc(img, n_ref)
    n = torch.empty(B, 20, 20, 4, device=DEV); ops.nchw_to_nhwc(img.to(DEV), n)
    assert torch.equal(n.cpu(), n_ref)


def test_clean_and_stn_kernels_vs_torch():
    ops, ref = _ops(), TorchEmulOps()
    g = torch.Generator().manual_seed(8)
    B, S, C = 3, 56, 48
    heads = torch.randn(B, S, S, 92, generator=g)
    heads[0, 0, 0, 50:75] = 1.0                      # exact tie -> first index
    body_r, amax_r = torch.empty(B, S, S, 76), torch.empty(B, S, S, dtype=torch.uint8)
    vis_r = [torch.empty(B, c, S, S) for c in (25, 25, 25, 15)]
    ref.clean_global(heads, body_r, amax_r, vis_r)
    body, amax = torch.empty(B, S, S, 76, device=DEV), torch.empty(B, S, S, dtype=torch.uint8, device=DEV)
    vis = [torch.empty(B, c, S, S, device=DEV) for c in (25, 25, 25, 15)]
    ops.clean_global(heads.to(DEV), body, amax, vis)
    assert torch.equal(amax.cpu(), amax_r) and torch.equal(body.cpu(), body_r)
    for a, b in zip(vis, vis_r):
        assert torch.equal(a.cpu(), b)
    x = torch.randn(B * 24, S, S, 24, generator=g)
    y_r, raw_r = torch.empty(B * 24, S, S, 24), torch.empty(B * 24, 21, S, S)
    ref.clean_parts(x, y_r, raw_r)
    y, raw = torch.empty(B * 24, S, S, 24, device=DEV), torch.empty(B * 24, 21, S, S, device=DEV)
    ops.clean_parts(x.to(DEV), y, raw)
    assert torch.equal(y.cpu(), y_r) and torch.equal(raw.cpu(), raw_r)
    y16 = torch.full((B * 24, S, S, 24), float("nan"), dtype=torch.float16, device=DEV)      # fp16 output option
    ops.clean_parts(x.to(DEV), y16, None)
    assert torch.equal(y16.cpu(), y_r.to(torch.float16))
    # stn params + sampling, both align_corners conventions
    hm = torch.randn(B, S, S, 24, generator=g) * 0.3
    ratio, offset = torch.rand(24, generator=g) + 0.5, torch.rand(24, generator=g) * 0.2
    xd = torch.randn(B, S, S, C, generator=g)
    for ac in (0, 1):
        c_r, th_r = torch.empty(B, 1, 24, 2), torch.empty(B, 1, 24, 3)
        ref.stn_params(hm, amax_r, ratio, offset, 0.5, ac, c_r, th_r)
        c, th = torch.empty(B, 1, 24, 2, device=DEV), torch.empty(B, 1, 24, 3, device=DEV)
        ops.stn_params(hm.to(DEV), amax, ratio.to(DEV), offset.to(DEV), 0.5, ac, c, th)
        assert (c.cpu() - c_r).abs().max() < 2e-5 and (th.cpu() - th_r).abs().max() < 2e-5
        crops_r = torch.empty(B * 24, S, S, C); ref.stn_sample(xd, th_r, ac, crops_r)
        crops = torch.empty(B * 24, S, S, C, device=DEV); ops.stn_sample(xd.to(DEV), th_r.to(DEV), ac, crops)
        assert (crops.cpu() - crops_r).abs().max() < 2e-4
        crops16 = torch.full((B * 24, S, S, C), float("nan"), dtype=torch.float16, device=DEV)
        ops.stn_sample(xd.to(DEV), th_r.to(DEV), ac, crops16)
        assert torch.equal(crops16, crops.to(torch.float16))           # same values, RN-rounded


def test_gcn_pose_head_vs_torch():
    from danet_b200 import netgraph as ng
    ops, ref = _ops(), TorchEmulOps()
    g = torch.Generator().manual_seed(9)
    B = 5
    gb = ng.graph_buffers()
    adj = torch.stack([torch.from_numpy(gb["r2p_A"][0]), torch.from_numpy(ng.undigraph_normalize(gb["A_mask"][0] + np.eye(24)).astype(np.float32)),
                       torch.from_numpy(gb["p2r_A"][0])]).float()
    dims = [(128, 128), (128, 256), (256, 256), (256, 128), (128, 128)]
    gp = {"adj": adj, "W": [torch.randn(i, o, generator=g) * (1.0 / i) ** 0.5 for i, o in dims],
          "b": [torch.randn(o, generator=g) * 0.1 for _, o in dims],
          "bn_scale": [torch.rand(24, generator=g) + 0.5 for _ in dims], "bn_shift": [torch.randn(24, generator=g) * 0.1 for _ in dims],
          "head_w": torch.randn(144, 128, generator=g) * 0.1, "head_b": torch.randn(144, generator=g) * 0.1,
          "mean_pose": torch.tensor([1., 0, 0, 1, 0, 0] * 24)}
    rot = torch.randn(B * 24, 1, 1, 128, generator=g)
    gpara = torch.randn(B, 1, 1, 13, generator=g)
    p_ref = torch.empty(B, 1, 1, 229); ref.gcn_head(gp, rot, gpara, p_ref)
    gpc = {k: ([t.to(DEV) for t in v] if isinstance(v, list) else v.to(DEV)) for k, v in gp.items()}
    p = torch.empty(B, 1, 1, 229, device=DEV); ops.gcn_head(gpc, rot.to(DEV), gpara.to(DEV), p)
    assert (p.cpu() - p_ref).abs().max() < 5e-5


F16_CASES = [
    # N, H, W, C1, C2, C3, k, stride
    (2, 56, 56, 48, 48, 48, 3, 1),
    (3, 13, 9, 24, 40, 16, 3, 1),          # ragged tiles, narrow channel counts
    (2, 28, 28, 64, 64, 128, 3, 2),        # fp16 input through the stride-2 parity planes
    (2, 20, 20, 24, 64, 64, 7, 2),         # limb_net.0 -> conv1 pattern (1x1 then 7x7 stride 2)
]


@pytest.mark.parametrize("case", F16_CASES)
def test_conv_tc_f16_intermediate_is_bit_identical(case):
    """y = conv2(conv1(x)): storing the intermediate as fp16 (DANET_CONV_Y_F16 / _X_F16) must give the
    same bits as the fp32 intermediate, because the kernel rounds its activations to fp16 (RN) either way."""
    N, H, W, C1, C2, C3, k, s = case
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, H, W, C1, generator=g).to(DEV)
    k1 = 1 if k == 7 else 3
    d1 = dict(N=N, H=H, W=W, Cin=C1, Cout=C2, ksize=k1, stride=1, pad=k1 // 2, wsets=1, relu=1)
    Ho = (H + 2 * (k // 2) - k) // s + 1
    Wo = (W + 2 * (k // 2) - k) // s + 1
    d2 = dict(N=N, H=H, W=W, Cin=C2, Cout=C3, ksize=k, stride=s, pad=k // 2, wsets=1, relu=0)
    w1 = (torch.randn(1, k1 * k1 * C1, C2, generator=g) * 0.1).to(DEV)
    w2 = (torch.randn(1, k * k * C2, C3, generator=g) * 0.05).to(DEV)
    b1 = (torch.randn(1, C2, generator=g) * 0.1).to(DEV)
    b2 = (torch.randn(1, C3, generator=g) * 0.1).to(DEV)
    res = torch.randn(N, Ho, Wo, C3, generator=g).to(DEV)
    p1, p2 = ops.conv_tc_pack(d1, w1), ops.conv_tc_pack(d2, w2)
    # fp32 intermediate
    t32 = torch.empty(N, H, W, C2, device=DEV)
    y32 = torch.empty(N, Ho, Wo, C3, device=DEV)
    ops.conv2d(d1, 1, x, p1, b1, None, t32)
    ops.conv2d(d2, 1, t32, p2, b2, res, y32)
    # fp16 intermediate
    t16 = torch.full((N, H, W, C2), float("nan"), dtype=torch.float16, device=DEV)
    y16 = torch.empty(N, Ho, Wo, C3, device=DEV)
    ops.conv2d(dict(d1, flags=2), 1, x, p1, b1, None, t16)
    ops.conv2d(dict(d2, flags=1), 1, t16, p2, b2, res, y16)
    torch.cuda.synchronize()
    assert torch.equal(t16, t32.to(torch.float16))       # same RN rounding as torch
    assert torch.equal(y16, y32)
    # the fp32 FMA path refuses fp16 tensors
    with pytest.raises(RuntimeError):
        ops.conv2d(dict(d1, flags=2), 0, x, w1, b1, None, t16)


def test_conv2x2_as_gemm_on_the_tensor_core_path():
    """plan.conv2x2_as_gemm: the ResNet tail's 3x3 convolutions on 2x2-pixel maps as ONE dense product through the
    tcgen05 kernel (images = pixels of a 1x1 convolution, nothing moves in memory) vs the exact fp32 kernel."""
    from danet_b200.plan import conv2x2_as_gemm
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    N, C = 64, 512
    x = torch.randn(N, 2, 2, C, generator=g).to(DEV)
    w = (torch.randn(1, 9 * C, C, generator=g) * 0.02)
    b = torch.randn(1, C, generator=g) * 0.1
    res = torch.randn(N, 2, 2, C, generator=g).to(DEV)
    d = dict(N=N, H=2, W=2, Cin=C, Cout=C, ksize=3, stride=1, pad=1, wsets=1, relu=1)
    y_ref = torch.empty(N, 2, 2, C, device=DEV)
    ops.conv2d(d, 0, x, w.to(DEV), b.to(DEV), res, y_ref)
    d2 = dict(N=1, H=N // 8, W=8, Cin=4 * C, Cout=4 * C, ksize=1, stride=1, pad=0, wsets=1, relu=1)
    assert ops.conv_tc_supported(d2)
    w2, b2 = conv2x2_as_gemm(w, b, C, C)
    y = torch.full((N, 2, 2, C), float("nan"), device=DEV)
    ops.conv2d(d2, 1, x, ops.conv_tc_pack(d2, w2.to(DEV)), b2.to(DEV), res, y)
    torch.cuda.synchronize()
    err = (y - y_ref).abs().max().item() / y_ref.abs().max().item()
    assert err < 8e-3, err
