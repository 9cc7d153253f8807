"""GPU parity of the whole network half (DaNet.infer_net) against golden vectors produced by the
reference's own modules (tests/golden/net_w*.npz), plus the end-to-end demo.py-style chain
infer_net -> SMPL -> IUV_Renderer checked against the CPU oracle."""
import numpy as np
import pytest
import torch

from net_common import build, check_against_golden, make_image

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("width", [32, 48])
def test_infer_net_fp32_matches_reference_golden(width):
    """fp32 FMA convolutions: para within 1e-4 of the reference (north_star tolerance for the
    floating-point outputs); argmax maps equal wherever the reference's top-2 margin > 1e-3."""
    net = build(width, DEV, conv_algo="simt")
    out = net.infer_net(make_image(2, 100).to(DEV))
    check_against_golden(out, width, para_tol=1e-4, kps_tol=1e-4, margin_eps=1e-3)


@pytest.mark.parametrize("width", [32, 48])
def test_infer_net_tensor_core_path(width):
    """fp16-operand tensor-core convolutions (where supported): tolerance scaled to the 2^-11 operand rounding
    through ~100 layers; integer maps must agree where the reference margin exceeds 0.05."""
    net = build(width, DEV, conv_algo="tc")
    out = net.infer_net(make_image(2, 100).to(DEV))
    plan = net.plan_for(2, torch.device(DEV))
    if plan.n_tc == 0:
        pytest.skip("tcgen05 conv path not built")
    check_against_golden(out, width, para_tol=3e-2, kps_tol=3e-2, margin_eps=0.05, min_agree=0.999)


def test_batch_invariance_and_cuda_graph():
    net = build(32, DEV, conv_algo="simt")
    img = make_image(4, 5).to(DEV)
    a = net.infer_net(img)["para"]
    b = net.infer_net(img[1:2])["para"]
    assert (a[1:2] - b).abs().max() < 1e-5           # images are independent (BN in eval mode)
    netg = build(32, DEV, conv_algo="simt", use_cuda_graph=True)
    c = netg.infer_net(img)["para"]
    c2 = netg.infer_net(img)["para"]                 # replay
    assert torch.equal(c, c2) and (a - c).abs().max() < 1e-6


def test_demo_chain_matches_oracle():
    """demo.py:109,148-151: infer_net -> SMPL(betas, rotmats, pose2rot=False) -> verts2uvimg."""
    from oracle import lbs, raster, synth
    net = build(32, DEV, conv_algo="simt")
    para = net.infer_net(make_image(2, 100).to(DEV))["para"]
    cam, betas, R = para[:, :3].contiguous(), para[:, 3:13].contiguous(), para[:, 13:].reshape(-1, 24, 3, 3)
    smpl = net.iuv2smpl.smpl
    out = smpl(betas=betas, body_pose=R[:, 1:], global_orient=R[:, :1], pose2rot=False)
    ref = lbs.smpl_forward(synth.make_smpl_model(0), betas.cpu().numpy(), R[:, 1:].cpu().numpy(), R[:, :1].cpu().numpy(),
                           pose2rot=False, dtype=np.float64)
    assert np.abs(out.vertices.cpu().numpy() - ref["vertices"]).max() < 1e-4
    img = net.iuv_renderer.verts2uvimg(out.vertices, cam)
    mesh = synth.make_dp_mesh(0)
    rimg, _, _ = raster.verts2uvimg(out.vertices.cpu().numpy(), cam.cpu().numpy(), mesh, synth.dp_textures(mesh))
    np.testing.assert_array_equal(img.cpu().numpy(), rimg)


def test_evaluate_batch_matches_oracle_mpjpe():
    """eval.py:166-212 chain on the GPU vs the numpy oracle (same para -> same joints -> same MPJPE)."""
    from danet_b200.evaluate import evaluate_batch
    from oracle import lbs, synth
    net = build(32, DEV, conv_algo="simt")
    img = make_image(3, 100).to(DEV)
    g = torch.Generator().manual_seed(3)
    gt = (torch.randn(3, 14, 3, generator=g) * 0.2).to(DEV)
    out = evaluate_batch(net, net.iuv2smpl.smpl, img, gt)
    para = out["para"].cpu().numpy()
    R = para[:, 13:].reshape(-1, 24, 3, 3)
    ref = lbs.smpl_forward(synth.make_smpl_model(0), para[:, 3:13], R[:, 1:], R[:, :1], pose2rot=False, dtype=np.float64)
    want = lbs.mpjpe_h36m(ref["joints_h36m"], gt.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(out["mpjpe"].cpu().numpy(), want, atol=1e-5)
    assert out["pred_j14"].shape == (3, 14, 3)


def test_f16_intermediates_are_bit_identical():
    """conv->conv tensors stored in fp16 (danet_conv_desc.flags): the tensor-core kernel rounds its
    activations to fp16 either way, so the whole forward must not change by a single bit."""
    img = make_image(2, 100).to(DEV)
    net_a = build(32, DEV, conv_algo="tc", f16_intermediates=True)
    net_b = build(32, DEV, conv_algo="tc", f16_intermediates=False)
    pa = net_a.plan_for(2, torch.device(DEV))
    assert len(pa.f16) > 20, "expected the first conv of every residual block to be stored in fp16"
    oa, ob = net_a.infer_net(img), net_b.infer_net(img)
    assert torch.equal(oa["para"], ob["para"])
    for x, y in zip(oa["visualization"]["iuv_pred"], ob["visualization"]["iuv_pred"]):
        assert torch.equal(x, y)


def test_gemm_2x2_option_matches_fma_path():
    """DaNet(gemm_2x2=True, the default): body_net's 2x2-pixel layers go through the tensor-core kernel as dense products
    (needs a batch that is a multiple of 8, >= 32); same parameters within the tensor-core tolerance."""
    img = make_image(32, 100).to(DEV)
    net_a = build(32, DEV, conv_algo="tc", gemm_2x2=True)
    net_b = build(32, DEV, conv_algo="tc", gemm_2x2=False)
    pa, pb = net_a.plan_for(32, torch.device(DEV)), net_b.plan_for(32, torch.device(DEV))
    assert pa.n_tc == pb.n_tc + 3                      # the three layer4 convolutions of body_net moved over
    oa, ob = net_a.infer_net(img), net_b.infer_net(img)
    assert (oa["para"] - ob["para"]).abs().max().item() < 3e-2
