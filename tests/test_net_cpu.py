"""CPU tests of the host logic of the network half (graph wiring, BatchNorm folding, weight packing,
buffer planning, state_dict surface) using the torch test double of the kernel-level ops
(tests/emul_ops.py) against golden vectors produced by the reference's own modules."""
import os

import numpy as np
import pytest
import torch

from oracle.net_ops import TorchEmulOps
from net_common import GOLD, build, check_against_golden, make_image


def test_state_dict_surface_matches_reference():
    """Every key / shape of the reference's img2iuv + smpl_para_Outs state_dict exists here (W48)."""
    import danet_b200
    from danet_b200 import synthetic
    net = danet_b200.DaNet(None, synthetic.make_mean_params(0), pretrained=False, width=48,
                           smpl_model=synthetic.make_smpl_model(0), dp_mesh=synthetic.make_dp_mesh(0))
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items() if not k.startswith("iuv2smpl.smpl.")}
    ref = {}
    for line in open(os.path.join(GOLD, "state_dict_keys_w48.txt")):
        k, shp = line.split()
        ref[k] = () if shp == "scalar" else tuple(int(d) for d in shp.split("x"))
    assert set(ref) == set(mine), (sorted(set(ref) - set(mine))[:5], sorted(set(mine) - set(ref))[:5])
    for k in ref:
        assert ref[k] == mine[k], (k, ref[k], mine[k])
    assert any(k.startswith("iuv2smpl.smpl.") for k in net.state_dict())
    assert net.img2iuv.dp2smpl_mapping[7] == [8, 10, 12, 14, 5, 5]
    assert hasattr(net, "iuv_renderer") and hasattr(net.iuv2smpl, "smpl")


@pytest.mark.parametrize("width", [32, 48])
def test_plan_with_emulated_kernels_matches_reference_golden(width):
    net = build(width, ops=TorchEmulOps())
    out = net.infer_net(make_image(2, 100))
    assert out["para"].shape == (2, 229)
    assert out["visualization"]["part_iuv_pred"].shape == (2, 24, 3, 7, 56, 56)
    check_against_golden(out, width, para_tol=5e-5, kps_tol=5e-5, margin_eps=1e-3)
    # rotation block of para is orthonormal (rot6d_to_rotmat)
    R = out["para"][:, 13:].reshape(-1, 3, 3)
    assert (R @ R.transpose(1, 2) - torch.eye(3)).abs().max() < 1e-5


def test_infer_net_requires_eval_mode_and_cuda():
    net = build(32, ops=TorchEmulOps())
    net.train()
    with pytest.raises(ValueError):
        net.infer_net(make_image(1, 1))
    net.eval()
    net._test_ops = None
    with pytest.raises(RuntimeError):                 # no CPU fallback in the product path
        net.infer_net(make_image(1, 1))
    with pytest.raises(NotImplementedError):
        net({"img": None})


def test_buffer_plan_reuses_memory_without_aliasing_live_tensors():
    net = build(32, ops=TorchEmulOps())
    plan = net.plan_for(1, "cpu", ops=TorchEmulOps())
    g = plan.g
    total = sum(t.nmult * t.H * t.W * t.Cp * 4 for t in g.tensors.values() if t.dtype == "f32")
    assert plan.bytes_alloc < 0.5 * total               # liveness-based reuse is effective
    # an op never reads and writes the same storage
    for op in g.ops:
        y = op.get("y")
        if y is None:
            continue
        ins = [op.get(k) for k in ("x", "res", "hm", "gpara")] + [t for t, _ in op.get("terms", [])]
        for t in ins:
            if t is not None and t.dtype == "f32" and y.dtype == "f32":
                assert plan.buf[t.name].data_ptr() != plan.buf[y.name].data_ptr(), (op["op"], t, y)


def test_pretrained_flag_needs_files():
    import danet_b200
    from danet_b200 import synthetic
    with pytest.raises(ValueError):
        danet_b200.DaNet(None, synthetic.make_mean_params(0), pretrained=True, width=48,
                         smpl_model=synthetic.make_smpl_model(0), dp_mesh=synthetic.make_dp_mesh(0))
    with pytest.raises(ValueError):
        danet_b200.DaNet(None, "/nonexistent/smpl_mean_params.npz", pretrained=False)


def test_f16_tensor_selection_rules():
    """plan._f16_tensors (host logic): only tensors written by a tensor-core conv (or the two glue kernels that
    can write fp16) and read exclusively as the INPUT of tensor-core convs may be stored in fp16."""
    from danet_b200 import netgraph as ng
    from danet_b200.plan import Plan

    class FakeOps(object):
        supports_f16 = True

        def conv_tc_supported(self, d):
            return d["H"] >= 4 and d["W"] >= 4          # what the tcgen05 kernel declines: tiny maps

    g = ng.danet_graph(48)
    plan = Plan.__new__(Plan)
    plan.g, plan.B, plan.conv_algo, plan.ops = g, 2, "tc", FakeOps()
    f16 = plan._f16_tensors()
    assert len(f16) > 100
    readers, writers = {}, {}
    for op in g.ops:
        for key in ("x", "res", "hm", "amax", "theta", "gpara"):
            t = op.get(key)
            if t is not None:
                readers.setdefault(t.name, []).append((op, key))
        for (t, _f) in op.get("terms", []):
            readers.setdefault(t.name, []).append((op, "term"))
        if op.get("y") is not None:
            writers[op["y"].name] = op
    keep = plan._keep()
    for name in f16:
        t = g.tensors[name]
        assert t.Cp % 8 == 0 and name not in keep
        w = writers[name]
        assert w["op"] in ("conv", "stn_sample", "clean_parts")
        for (op, key) in readers[name]:
            assert op["op"] == "conv" and key == "x", (name, op["op"], key)     # never a residual / fuse / glue input
            assert FakeOps().conv_tc_supported(plan._conv_desc(op))
    # the big limb-branch tensors are in: crops (stn_sample -> grouped conv) and limb_net.0 -> conv1
    assert g.outputs["part_iuv"].name in f16
    crops = [op["y"].name for op in g.ops if op["op"] == "stn_sample"][0]
    assert crops in f16
    # nothing is selected on the exact fp32 path
    plan.conv_algo = "simt"
    assert plan._f16_tensors() == set()


def test_plan_wires_2x2_convs_as_dense_products():
    """Host logic of plan.conv2x2_as_gemm: with a test double standing in for the tensor-core kernel, a plan
    that re-expresses 3x3 convolutions on 2x2-pixel maps as 1x1 convolutions over an (N/8) x 8 pixel map
    (same buffers, re-laid weights) must reproduce the direct convolution (+ residual + ReLU)."""
    from danet_b200 import netgraph as ng
    from danet_b200.plan import Plan

    class FakeTcOps(TorchEmulOps):
        def __init__(self):
            self.tc_calls = 0

        def conv_tc_supported(self, d):
            return d["ksize"] == 1 and d["H"] >= 4 and d["W"] >= 4          # only the transformed layers

        def conv_tc_pack(self, d, w):
            return w

        def conv2d(self, d, algo, x, w, bias, res, y):
            if algo == 1:
                self.tc_calls += 1
                shp_i, shp_o = (d["N"], d["H"], d["W"], d["Cin"]), (d["N"], d["H"], d["W"], d["Cout"])
                x, y = x.reshape(shp_i), y.view(shp_o)                         # same memory, other shape
                res = res.reshape(shp_o) if res is not None else None
            super().conv2d(d, algo, x, w, bias, res, y)

    g = ng.Graph()
    C = 16
    img = g.tensor(1, 2, 2, C, name="image")
    g.ops.append(dict(op="input", y=img))
    t1 = g.conv(img, "c1", C, 3, 1, bn="bn1", relu=True)
    t2 = g.conv(t1, "c2", C, 3, 1, bn="bn2", relu=True, res=img)
    g.outputs = dict(heads=t2, para=t2)
    gen = torch.Generator().manual_seed(5)
    sd = {}
    for key, spec in g.params.items():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = torch.rand(spec.shape, generator=gen) + 0.5
        else:
            sd[key] = torch.randn(spec.shape, generator=gen) * 0.2
    B = 32
    x = torch.randn(B, C, 2, 2, generator=gen)
    outs = []
    for algo, flag in (("simt", False), ("tc", True)):
        ops = FakeTcOps()
        plan = Plan(g, sd, B, "cpu", conv_algo=algo, want_vis=False, ops=ops, gemm_2x2=flag)
        plan.run(x)
        outs.append(plan.out("para").clone())
        assert ops.tc_calls == (2 if flag else 0)
    assert (outs[0] - outs[1]).abs().max().item() < 1e-5
    assert outs[0].abs().max().item() > 0.1
