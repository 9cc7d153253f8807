"""Per-tensor comparison of two plans of the same network on the same input (debugging aid): where does the
tensor-core path start to deviate from the fp32 FMA path?  Usage: python tools/layer_diff.py [width] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from net_common import build, make_image
from danet_b200.plan import Plan

width = int(sys.argv[1]) if len(sys.argv) > 1 else 48
prec = sys.argv[2] if len(sys.argv) > 2 else "exact"
dev = torch.device("cuda:0")
net = build(width, dev, conv_algo="simt")
sd = {k: v for k, v in net.state_dict().items() if not k.startswith("iuv2smpl.smpl.")}
img = make_image(2, 100).to(dev)
pa = Plan(net.graph, sd, 2, dev, conv_algo="simt", keep_all=True)
pb = Plan(net.graph, sd, 2, dev, conv_algo="tc", precision=prec, keep_all=True)
pa.run(img); pb.run(img)
torch.cuda.synchronize()
worst = 0.0
for idx, op in enumerate(net.graph.ops):
    y = op.get("y")
    if y is None or y.dtype != "f32":
        continue
    a, b = pa.buf[y.name], pb.buf[y.name]
    va = (a.value() if not torch.is_tensor(a) else a).float()
    vb = (b.value() if not torch.is_tensor(b) else b).float()
    d = (va - vb).abs()
    scale = va.abs().max().item()
    rel = d.max().item() / max(scale, 1e-30)
    mean_rel = d.mean().item() / max(va.abs().mean().item(), 1e-30)
    flag = " <<<" if rel > 3 * max(worst, 1e-6) else ""
    worst = max(worst, rel)
    desc = op["op"]
    if desc == "conv":
        desc = "conv k%d s%d %d->%d @%d g%d%s" % (op["k"], op["stride"], op["x"].C, y.C, op["x"].H, op["groups"], " +res" if op["res"] is not None else "")
    print("%4d %-34s %-10s max|d|=%.3e max|v|=%.3e rel=%.2e mean_rel=%.2e%s" % (idx, desc, y.name, d.max().item(), scale, rel, mean_rel, flag))
pa_, pb_ = pa.buf[net.graph.outputs["para"].name], pb.buf[net.graph.outputs["para"].name]
print("para diff:", (pa_.value().float() - pb_.value().float()).abs().max().item())
